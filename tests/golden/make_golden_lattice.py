"""Generates tests/golden/lattice_per_d2_poisson3000.npz from the UNMODIFIED reference built with -DPERIODIC and without PMGRID
(oracle/_ref variant per_d2_f32): gravity_tree() = tree walk + lattice-sum correction walk (forcetree.c:1606-1608, 2077-2455) with the
tables of lattice_init (forcetree.c:3611).  Run in the build container only:  python tests/golden/make_golden_lattice.py

All particles belong to species 0: in this FLOAT = float build only the [0][0] pair table is complete (oracle/refrun.py).  The fixture
holds the float32 particle set in the reference's order, accelerations / GravCost / OldAcc of a Barnes-Hut pass and of a relative-criterion
pass, and 4096 sampled entries of the reference's fcorrx/y/z[0][0] tables (the whole tables would be 6.6 MB)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import g2test  # noqa: E402
from refrun import RefOracle  # noqa: E402


def main():
    n, box = 3000, 1000.0
    pos, mass, ptype = g2test.periodic_poisson(n, box, seed=53)
    mass = (mass * np.random.default_rng(6).uniform(0.5, 2.0, n)).astype(np.float32)
    ptype[:] = 1
    soft, grav = (box / 14 / 30.0,) * 6, g2test.GRAV_D2
    ref = RefOracle("per_d2_f32", int(1.1 * n) + 64, boxsize=box, softening=soft, gravity=grav)
    tabs = ref.lattice_tables()[:, 0, 0]
    ref.load(pos, mass, ptype)
    ref.domain()
    rp = ref.particles()
    ref.gravity()
    r1 = ref.particles()
    ref.set_opening(0.0, 0.005, 1)
    ref.force_rebuild()
    ref.gravity()
    r2 = ref.particles()
    idx = np.random.default_rng(7).integers(0, 65, size=(4096, 3))
    out = dict(pos=rp["pos"].astype(np.float32), mass=rp["mass"].astype(np.float32), type=rp["type"].astype(np.int32), box=box,
               soft=np.asarray(soft), grav=np.asarray(grav, dtype=np.int32), maxpart=ref.maxpart,
               bh_acc=r1["acc"].astype(np.float32), bh_cost=r1["cost"], bh_oldacc=r1["oldacc"].astype(np.float32),
               rel_acc=r2["acc"].astype(np.float32), rel_cost=r2["cost"],
               table_index=idx.astype(np.int32), table_sample=tabs[:, idx[:, 0], idx[:, 1], idx[:, 2]])
    path = os.path.join(HERE, "lattice_per_d2_poisson3000.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
