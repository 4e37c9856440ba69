# round 2, GPU call 23: FP32 partial sums flushed into the FP64 accumulators at every descent (mask 0) / every 2nd, 4th, 8th cell index (masks 1, 3, 7)
mkdir -p gpurun_out
for m in 0 1 3 7; do
  for wl in periodic256 hernquist1m; do
    G2GPU_WALK_FLUSH_MASK=$m timeout 600 python bench.py --workload $wl --steps 3 --no-cpu-baseline --no-shim > gpurun_out/r2_bench23_${wl}_mask${m}.json 2> gpurun_out/r2_bench23_${wl}_mask${m}.err
  done
done
G2GPU_WALK_FLUSH_MASK=3 timeout 900 python bench.py --steps 3 --no-shim > gpurun_out/r2_bench23_periodic256_mask3_parity.json 2> gpurun_out/r2_bench23_periodic256_mask3_parity.err
G2GPU_WALK_FLUSH_MASK=7 timeout 900 python bench.py --steps 3 --no-shim > gpurun_out/r2_bench23_periodic256_mask7_parity.json 2> gpurun_out/r2_bench23_periodic256_mask7_parity.err
G2GPU_WALK_FLUSH_MASK=7 timeout 900 python bench.py --workload hernquist1m --steps 3 --no-shim > gpurun_out/r2_bench23_hernquist1m_mask7_parity.json 2> gpurun_out/r2_bench23_hernquist1m_mask7_parity.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench23_*.json")):
    try:
        d=json.load(open(f)); p=d.get("parity") or {}
        print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "parity", p.get("median"), p.get("p999"), p.get("max"), p.get("cost_mismatch"))
    except Exception as e: print(f, "ERR", e)
PY
