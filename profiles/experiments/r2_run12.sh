# round 2, GPU call 12: what the exact-GravCost machinery costs: accumulators (FP64 in shared memory / FP32 in registers) x guard-band flags (on / off), 256^3
# (library built with EXTRA=-DG2_WALK_ACC_MATRIX so that all four instantiations exist); periodic non-PM potential test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_potential.py -m gpu -q -k "periodic" > gpurun_out/r2_gpu_tests_12.log 2>&1; tail -6 gpurun_out/r2_gpu_tests_12.log
for accf in 0 1; do for ex in 1 0; do
  fl=""; [ $accf = 1 ] && fl="--acc-float"
  timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-shim --walk-exact $ex $fl > gpurun_out/r2_bench12_p256_accfloat${accf}_ex${ex}.json 2> gpurun_out/r2_bench12_p256_accfloat${accf}_ex${ex}.err
done; done
timeout 900 python bench.py --steps 3 --no-shim --acc-float --walk-exact 1 > gpurun_out/r2_bench12_p256_accfloat1_ex1_parity.json 2> gpurun_out/r2_bench12_p256_accfloat1_ex1_parity.err
timeout 600 python bench.py --workload hernquist1m --steps 3 --no-shim --acc-float --walk-exact 1 > gpurun_out/r2_bench12_h1m_accfloat1_ex1_parity.json 2> gpurun_out/r2_bench12_h1m_accfloat1_ex1_parity.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench12_*.json")):
    try:
        d=json.load(open(f)); p=d.get("parity") or {}
        print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "rewalked", d.get("rewalked_targets"), "parity", p.get("median"), p.get("p999"), p.get("max"), p.get("cost_mismatch"))
    except Exception as e: print(f, "ERR", e)
PY
