"""Generates tests/golden/pmpot_pm64_yuk_poisson4096.npz from the UNMODIFIED reference's periodic PM potential
(pmpotential_periodic, pm_periodic.c:798, compiled from /root/reference by oracle/ref/Makefile, variant pm64_yuk_f32 = the
NGRAVS_YUKAWA_FORCETEST wiring: GreensFxns[i][j] = pgyukawa for i != j, none for i == j -- finite at k = 0, unlike the stock 1/k^2
for which the reference returns infinite potentials).  Run in the build container only:  python tests/golden/make_golden_pmpot.py

The fixture holds the float32 particle set in the order the reference's domain decomposition left it and what the routine added
to P[].Potential."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import g2test  # noqa: E402
from refrun import RefOracle  # noqa: E402

BOX, N, G, YUKAWA_IMASS = 100000.0, 4096, 43007.1, 60.0      # ngravs.c:42


def make():
    pos, mass, ptype = g2test.periodic_poisson(N, BOX, ntypes=2)
    mass = (mass * np.random.default_rng(5).uniform(0.5, 2.0, N)).astype(np.float32)
    ref = RefOracle("pm64_yuk_f32", int(1.1 * N) + 64, boxsize=BOX, softening=(100.0,) * 6, gravity=g2test.GRAV_D2, G=G)
    ref.load(pos, mass, ptype)
    ref.domain()
    pot = ref.pmpotential()
    rp = ref.particles()
    return dict(pos=rp["pos"].astype(np.float32), mass=rp["mass"].astype(np.float32), type=rp["type"].astype(np.int32),
                grav=np.asarray(g2test.GRAV_D2, dtype=np.int32), box=BOX, pmgrid=64, G=G, yukawa_imass=YUKAWA_IMASS, pmpot=pot.astype(np.float32))


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "pmpot_pm64_yuk_poisson4096.npz"), **make())
    print("wrote pmpot_pm64_yuk_poisson4096")
