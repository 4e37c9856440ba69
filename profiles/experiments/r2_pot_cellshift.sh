#!/bin/bash
# NOT YET RUN (written at the end of round 1, no GPU minutes left).  Per-cell periodic image shift in the potential walk
# (-DG2_POT_CELLSHIFT, csrc/g2_pot.cu): rebuilds the one object with the macro, relinks, checks parity (tests/test_gpu_potential.py)
# and times the kernel with bench.py on periodic 128^3 and 256^3 against the default build.  Run on the GPU box from the repo root.
# Expectation from the force walk, where the same change was worth ~10 %: 11.3 -> ~10 ms at 128^3; results must stay bit-identical.
set -e
cd gadget-2.0.7-ngravs_b200
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
link() { /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xlinker -soname=libg2gpu.so -o libg2gpu.so csrc/*.o -lcudart -lcufft; }
run() { (cd .. && for w in periodic128 periodic256; do python bench.py --workload $w --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', '$w', l['potential_walk']['ms_per_call'])"; done); }
cp libg2gpu.so /tmp/libg2gpu.so.keep; cp csrc/g2_pot.o /tmp/g2_pot.o.keep
run default
$NV -DG2_POT_CELLSHIFT -c csrc/g2_pot.cu -o csrc/g2_pot.o && link
(cd .. && python -m pytest tests/test_gpu_potential.py -q -m gpu 2>&1 | tail -2)
run cellshift
cp /tmp/g2_pot.o.keep csrc/g2_pot.o; cp /tmp/libg2gpu.so.keep libg2gpu.so
