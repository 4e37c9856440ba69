"""Generates tests/golden/pm_*.npz from the UNMODIFIED reference's periodic PM force (pmforce_periodic, pm_periodic.c:204,
compiled from /root/reference by oracle/ref/Makefile and run single-rank through the rfftwnd_mpi stand-in of
oracle/ref/stubs.c).  Run in the build container only:  python tests/golden/make_golden_pm.py

Each fixture holds the float32 particle set in the order the reference's domain decomposition left it, and P[].GravPM."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import g2test  # noqa: E402
from refrun import RefOracle  # noqa: E402

BOX = 100000.0
CASES = {
    "pm_pm64_d2_poisson4096": dict(variant="pm64_d2_f32", n=4096, ntypes=2, grav=g2test.GRAV_D2),
    "pm_pm64_d4_poisson4096": dict(variant="pm64_d4_f32", n=4096, ntypes=6, grav=(0, 1, 2, 3, 1, 2)),
}


def make(case):
    c = CASES[case]
    pos, mass, ptype = g2test.periodic_poisson(c["n"], BOX, ntypes=c["ntypes"])
    mass = (mass * np.random.default_rng(5).uniform(0.5, 2.0, c["n"])).astype(np.float32)
    ref = RefOracle(c["variant"], int(1.1 * c["n"]) + 64, boxsize=BOX, softening=(100.0,) * 6, gravity=c["grav"])
    ref.load(pos, mass, ptype)
    ref.domain()
    pm = ref.pmforce()
    rp = ref.particles()
    return dict(pos=rp["pos"].astype(np.float32), mass=rp["mass"].astype(np.float32), type=rp["type"].astype(np.int32),
                grav=np.asarray(c["grav"], dtype=np.int32), box=BOX, pmgrid=64, G=1.0, gravpm=pm.astype(np.float32))


if __name__ == "__main__":
    for case in CASES:
        np.savez_compressed(os.path.join(HERE, case + ".npz"), **make(case))
        print("wrote", case)
