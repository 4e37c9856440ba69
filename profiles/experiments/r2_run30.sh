# round 2, GPU call 30: work distribution of the walk: one contiguous block of chunks per SM (G2GPU_WALK_SM_LOCAL=1) against the global chunk counter (=0)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tree_walk.py tests/test_gpu_group.py -m gpu -q -x > gpurun_out/r2_gpu_tests_30.log 2>&1; tail -3 gpurun_out/r2_gpu_tests_30.log
for sl in 0 1; do
for wl in periodic256 periodic128 hernquist1m; do
  G2GPU_WALK_SM_LOCAL=$sl timeout 600 python bench.py --workload $wl --steps 3 --no-cpu-baseline --no-shim > gpurun_out/r2_bench30_${wl}_sl${sl}.json 2> gpurun_out/r2_bench30_${wl}_sl${sl}.err
done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench30_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "ia", d["ia_per_particle"], "rewalked", d.get("rewalked_targets"))
    except Exception as e: print(f, "ERR", e)
PY
for sl in 0 1; do
G2GPU_WALK_SM_LOCAL=$sl timeout 600 ncu --metrics gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio --clock-control none -k regex:walk_kernel -s 2 -c 1 --csv --log-file gpurun_out/r2_l1_sl${sl}.csv python bench.py --profile --steps 1 --no-cpu-baseline --no-shim > /dev/null 2>&1
tail -6 gpurun_out/r2_l1_sl${sl}.csv
done
