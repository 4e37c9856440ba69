"""CPU test (no GPU) of the lattice-sum correction oracle: the restatement oracle/g2_oracle.c::lattice_correction against
the UNMODIFIED reference built with -DPERIODIC and without PMGRID (variant per_d2_f32), whose force_treeevaluate runs
force_treeevaluate_lattice_correction (forcetree.c:2077-2455) after its own walk.  The port is fed the reference's own tables as
lattice_init left them -- in a FLOAT = float build only the [0][0] pair table is complete, the pairs that found the cache file of the
first pair have a zero second half (oracle/refrun.py) -- so equal results also pin the [target][source] indexing of the tables.
The reference needs ~30 s to tabulate the Ewald sums (lattice_init, forcetree.c:3611); the tables (3 x 4 x 65^3 doubles) are too large for a
committed fixture, so this test runs where oracle/_ref exists (here, and on the GPU box, where oracle/_ref travels)."""
import os

import numpy as np
import pytest

import g2test
from portrun import PortOracle
from refrun import RefOracle, available

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.skipif(not available("per_d2_f32"), reason="oracle/_ref not built (needs /root/reference)")
def test_port_lattice_correction_matches_reference_build():
    n, box = 6000, 1000.0
    pos, mass, ptype = g2test.periodic_poisson(n, box, seed=31)
    mass = (mass * np.random.default_rng(4).uniform(0.5, 2.0, n)).astype(np.float32)
    soft, grav = (box / 18 / 30.0,) * 6, g2test.GRAV_D2
    ref = RefOracle("per_d2_f32", int(1.1 * n) + 64, boxsize=box, softening=soft, gravity=grav)
    tabs = ref.lattice_tables()
    assert tabs is not None and np.abs(tabs[:, 0, 0]).max() > 0
    ref.load(pos, mass, ptype)
    ref.domain()
    rp = ref.particles()
    o = PortOracle(ref.maxpart, D=2, periodic=True, shortrange=False, unequal=False, boxsize=box, softening=soft, gravity=grav)
    o.set_lattice_tables(tabs)
    o.load(rp["pos"], rp["mass"], rp["type"])
    o.domain()
    assert np.array_equal(o.particles()["id"], np.arange(n))
    ref.gravity()                          # Barnes-Hut first pass
    o.gravity()
    r1, p1 = ref.particles(), o.particles()
    assert np.array_equal(p1["cost"], r1["cost"])
    assert np.array_equal(p1["acc"], r1["acc"].astype(np.float32))
    # the correction is really in there: without the tables the port gives the nearest-image force and a smaller GravCost
    o2 = PortOracle(ref.maxpart, D=2, periodic=True, shortrange=False, unequal=False, boxsize=box, softening=soft, gravity=grav)
    o2.load(rp["pos"], rp["mass"], rp["type"])
    o2.domain()
    o2.gravity()
    assert (o2.particles()["cost"] < p1["cost"]).all() and not np.array_equal(o2.particles()["acc"], p1["acc"])
    # relative criterion with the OldAcc of the first pass
    ref.set_opening(0.0, 0.005, 1)
    ref.force_rebuild()
    ref.gravity()
    o.set_opening(0.0, 0.005)
    o.load(rp["pos"], rp["mass"], rp["type"], oldacc=r1["oldacc"])
    o.domain()
    o.gravity()
    r2, p2 = ref.particles(), o.particles()
    assert np.array_equal(p2["cost"], r2["cost"])
    assert np.array_equal(p2["acc"], r2["acc"].astype(np.float32))


def test_port_lattice_correction_matches_golden_fixture():
    """Needs no reference build: the port tabulates ewald_force itself (g2o_make_ewald_table, the sums of ngravs.c:1170-1236 in the same
    order) and must reproduce the reference's sampled table entries and its results for the committed single-species fixture
    (tests/golden/make_golden_lattice.py).  Tolerances instead of bit patterns only because erfc/exp/sin may differ by an ulp between
    C libraries; on the machine that made the fixture the table is bit-identical."""
    from portrun import make_ewald_table
    g = np.load(os.path.join(GOLD, "lattice_per_d2_poisson3000.npz"))
    box, n = float(g["box"]), len(g["mass"])
    ew = make_ewald_table(64) / (box * box)
    idx = g["table_index"]
    sample = ew[:, idx[:, 0], idx[:, 1], idx[:, 2]]
    assert np.abs(sample - g["table_sample"]).max() <= 1e-12 * np.abs(g["table_sample"]).max()
    o = PortOracle(int(g["maxpart"]), D=2, periodic=True, shortrange=False, unequal=False, boxsize=box, softening=g["soft"], gravity=g["grav"])
    o.set_lattice_tables(np.broadcast_to(ew[:, None, None], (3, 2, 2, 65, 65, 65)).copy())
    o.load(g["pos"], g["mass"], g["type"])
    o.domain()
    assert np.array_equal(o.particles()["id"], np.arange(n))
    o.gravity()
    p1 = o.particles()
    assert np.array_equal(p1["cost"], g["bh_cost"])
    assert g2test.rel_err(p1["acc"], g["bh_acc"]).max() <= 1e-6
    o.set_opening(0.0, 0.005)
    o.load(g["pos"], g["mass"], g["type"], oldacc=g["bh_oldacc"])
    o.domain()
    o.gravity()
    p2 = o.particles()
    assert np.array_equal(p2["cost"], g["rel_cost"])
    assert g2test.rel_err(p2["acc"], g["rel_acc"]).max() <= 1e-6
