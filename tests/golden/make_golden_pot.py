"""Generates tests/golden/pot_*.npz from the UNMODIFIED reference's TreePM tree potential walk
(force_treeevaluate_potential_shortrange, forcetree.c:2789-3163, called for every particle like the loop of compute_potential,
potential.c:86-97; compiled from /root/reference by oracle/ref/Makefile).  Run in the build container only:
    python tests/golden/make_golden_pot.py

Each fixture holds the float32 particle set in the order the reference's domain decomposition left it (masses are NOT all equal: the
table term of the reference's particle potential carries no mass factor, forcetree.c:3115, and the fixture must see that), both short-range
tables, and P[].Potential straight after the walk (pre-G, self term included) for a Barnes-Hut pass and for a relative-criterion pass
with the OldAcc of a preceding gravity_tree()."""
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
    "pot_pm64_d2_poisson4096": dict(variant="pm64_d2_f32", n=4096, ntypes=2, grav=g2test.GRAV_D2, seed=42),
    "pot_pm64_d4_poisson4096": dict(variant="pm64_d4_f32", n=4096, ntypes=6, grav=(0, 1, 2, 3, 1, 2), seed=7),
}


def make(case):
    c = CASES[case]
    n = c["n"]
    pos, mass, ptype = g2test.periodic_poisson(n, BOX, seed=c["seed"], ntypes=c["ntypes"])
    mass = (mass * np.random.default_rng(5).uniform(0.5, 2.0, n)).astype(np.float32)
    eps = BOX / 16 / 30.0
    soft = (eps,) * 6                  # the PM variants are built without UNEQUALSOFTENINGS: all softenings must be equal (domain.c)
    ref = RefOracle(c["variant"], int(1.1 * n) + 64, boxsize=BOX, softening=soft, gravity=c["grav"])
    ref.load(pos, mass, ptype)
    ref.domain()
    ref.gravity()                      # Barnes-Hut first pass: builds the tree, leaves OldAcc, switches to the relative criterion
    rp = ref.particles()
    ref.set_opening(0.5, 0.005, 1)
    pot_bh = ref.potential()
    ref.set_opening(0.0, 0.005, 1)
    pot_rel = ref.potential()
    asmth, rcut = ref.pm_split()
    return dict(pos=rp["pos"].astype(np.float32), mass=rp["mass"].astype(np.float32), type=rp["type"].astype(np.int32),
                oldacc=rp["oldacc"].astype(np.float32), grav=np.asarray(c["grav"], dtype=np.int32), soft=np.asarray(soft), box=BOX,
                pmgrid=64, D=ref.D, maxpart=ref.maxpart, asmth=asmth, rcut=rcut, srtable=ref.srtable(), srpot=ref.srpot_table(),
                pot_bh=pot_bh.astype(np.float32), pot_rel=pot_rel.astype(np.float32))


if __name__ == "__main__":
    for case in CASES:
        np.savez_compressed(os.path.join(HERE, case + ".npz"), **make(case))
        print("wrote", case, os.path.getsize(os.path.join(HERE, case + ".npz")))
