"""One g2gpu_pm_periodic call on the default bench workload (periodic 256^3, D = 2, particles in Peano-Hilbert order), for ncu:
   ncu --set full --clock-control none -k regex:pm_ -o gpurun_out/prof_pm_p256 python profiles/experiments/pm_profile.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "gadget-2.0.7-ngravs_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import g2gpu  # noqa: E402
import g2test  # noqa: E402

n, box, N = 256 ** 3, 100000.0, 256
pos, mass, ptype = g2test.periodic_poisson(n, box)
tg = g2gpu.TreeGravity(max_part=n + 64, n_gravs=2, periodic=True, shortrange=True, unequal_softenings=False)
tg.set_species(g2test.GRAV_D2, g2test.force_softening((1.0,) * 6))
tg.upload(pos, mass, ptype)
tg.domain()
perm = tg.order()
tg.upload(pos[perm], mass[perm], ptype[perm])
for _ in range(2):
    tg.pm_device(N, box)
print("pm_ms", tg.timings()["pm_ms"])
tg.close()
