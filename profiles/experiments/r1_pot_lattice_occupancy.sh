#!/bin/bash
# Resident CTAs per SM of the potential walk and of the lattice-correction walk: rebuilds the one object with another launch bound
# (register cap) and grid, relinks, and times the kernel with bench.py on periodic 128^3.  Run on the GPU box from the repo root.
set -e
cd gadget-2.0.7-ngravs_b200
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
link() { /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xlinker -soname=libg2gpu.so -o libg2gpu.so csrc/*.o -lcudart -lcufft; }
cp libg2gpu.so /tmp/libg2gpu.so.keep; cp csrc/g2_lattice.o /tmp/g2_lattice.o.keep; cp csrc/g2_pot.o /tmp/g2_pot.o.keep
for MB in 8 7; do
  $NV -DLATTICE_MINBLOCKS=$MB -c csrc/g2_lattice.cu -o csrc/g2_lattice.o && link
  (cd .. && python bench.py --workload periodic128nopm --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('LATTICE_MINBLOCKS=$MB', l['lattice_correction']['ms_per_step'], 'walk', l['stages_ms']['walk_kernel_ms'])")
done
cp /tmp/g2_lattice.o.keep csrc/g2_lattice.o
for MB in 8 9; do
  $NV -DPOT_MINBLOCKS=$MB -c csrc/g2_pot.cu -o csrc/g2_pot.o && link
  (cd .. && python bench.py --workload periodic128 --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('POT_MINBLOCKS=$MB', l['potential_walk']['ms_per_call'])")
done
cp /tmp/g2_pot.o.keep csrc/g2_pot.o; cp /tmp/libg2gpu.so.keep libg2gpu.so
