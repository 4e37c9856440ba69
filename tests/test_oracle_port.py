"""CPU tests (no GPU): the oracle restatement (oracle/g2_oracle.c) against (a) the golden fixtures generated from the
unmodified reference (tests/golden/make_golden.py) and (b) the reference build itself when oracle/_ref is present.
Everything integer/topological is bit-exact; accelerations are compared as float32 bit patterns too, because the port
performs the reference's double arithmetic in the reference's order."""
import glob
import os

import numpy as np
import pytest

import g2test
from portrun import PortOracle, make_srtable
from refrun import RefOracle, available

GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")) if "config1" not in p and not os.path.basename(p).startswith(("pm_", "pmpot_", "pot_", "lattice_")))   # pm_*: PM fixtures, tests/test_pm_oracle.py


def port_from_fixture(g):
    periodic = float(g["box"]) > 0
    o = PortOracle(int(g["maxpart"]), D=int(g["D"]), periodic=periodic, shortrange=periodic, unequal=not periodic,
                   boxsize=float(g["box"]), pmgrid=int(g["pmgrid"]) if periodic else 0, softening=g["soft"], gravity=g["grav"])
    if periodic:
        o.set_srtable(g["srtable"])
    return o


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_port_reproduces_golden_fixture(path):
    g = np.load(path)
    o = port_from_fixture(g)
    o.load(g["in_pos"], g["in_mass"], g["in_type"])
    o.domain()
    p = o.particles()
    d = o.domain_info()
    assert np.array_equal(np.concatenate([d["corner"], d["center"], [d["len"], d["fac"]]]), g["domain"])
    assert np.array_equal(p["key"], g["keys"])
    # order: identical wherever (species, key) is unique; the reference's qsort leaves ties unordered
    same = p["id"] == g["order_id"]
    if not same.all():
        keys = p["key"]
        tie = np.zeros(len(keys), dtype=bool)
        eq = keys[1:] == keys[:-1]
        tie[1:] |= eq
        tie[:-1] |= eq
        assert tie[~same].all()
    # continue from the reference's own order so that particle indices agree exactly
    o.load(g["in_pos"][g["order_id"]], g["in_mass"][g["order_id"]], g["in_type"][g["order_id"]])
    o.domain()
    assert np.array_equal(o.particles()["id"], np.arange(len(g["keys"])))
    o.gravity()
    t = o.topnodes()
    for k in ("daughter", "size", "startkey", "count", "domain_node_index"):
        assert np.array_equal(t[k], g["top_" + k]), k
    leaf_ref = np.where(g["top_daughter"] == -1, g["top_leaf"], -1)
    assert np.array_equal(t["leaf"], leaf_ref)
    tr = o.tree()
    for k in ("len", "center", "s", "mass", "bitflags", "sibling", "nextnode", "father", "p_nextnode", "p_father"):
        assert np.array_equal(tr[k], g["tree_" + k]), k
    p1 = o.particles()
    assert np.array_equal(p1["cost"], g["bh_cost"])
    assert np.array_equal(p1["acc"], g["bh_acc"])
    assert np.array_equal(p1["oldacc"], g["bh_oldacc"])
    o.set_opening(0.0, 0.005)
    o.load(g["in_pos"][g["order_id"]], g["in_mass"][g["order_id"]], g["in_type"][g["order_id"]], oldacc=g["bh_oldacc"])
    o.domain()
    o.gravity()
    p2 = o.particles()
    assert np.array_equal(p2["cost"], g["rel_cost"])
    assert np.array_equal(p2["acc"], g["rel_acc"])


def test_port_peano_key_known_values():
    g = np.load(GOLD[0])
    o = PortOracle(16)
    # the first level of the curve visits the octants in this order (digit of octant bx*4+by*2+bz)
    first = [o.peano_key(x, y, z, 1) for x in (0, 1) for y in (0, 1) for z in (0, 1)]
    assert sorted(first) == list(range(8))
    # keys are a bijection of the 3-bit grid at 3 bits per dimension and neighbours along the curve are grid neighbours
    pts = {}
    for x in range(8):
        for y in range(8):
            for z in range(8):
                pts[o.peano_key(x, y, z, 3)] = (x, y, z)
    assert sorted(pts) == list(range(512))
    for k in range(511):
        a, b = pts[k], pts[k + 1]
        assert sum(abs(a[i] - b[i]) for i in range(3)) == 1
    assert g["keys"].max() < (1 << 54)


def test_port_srtable_matches_reference_table_and_closed_form():
    """performConvolution restated with our own DFT reproduces the reference's Newtonian table (fixture), and the table
    is the tabulated erfc split of the TreePM force: m/r^2 [erfc(r/2rs) + r/(rs sqrt(pi)) exp(-r^2/4rs^2)]."""
    from math import erfc, exp, pi, sqrt
    g = np.load([p for p in GOLD if "pm64_d2" in p][0])
    tab = make_srtable(2048, 0, 0.0)
    ref = g["srtable"][0, 0]
    assert np.max(np.abs(tab - ref) / np.abs(ref)) < 1e-8
    # at bin centres u = 3(i+1/2)/NTAB = r/(2 rs):  1/r^2 - tab/(4 pi rs^2) == shortrange Newtonian factor / r^2
    for i in (10, 200, 700, 1500):
        u = 3.0 / 2048 * (i + 0.5)
        rs = 1.0
        r = 2 * u * rs
        lhs = 1 / r ** 2 - ref[i] / (4 * pi * rs * rs)
        rhs = (erfc(u) + 2 * u / sqrt(pi) * exp(-u * u)) / r ** 2
        assert abs(lhs - rhs) <= 2e-6 * (1 / r ** 2)


@pytest.mark.skipif(not available("np_d2_f32"), reason="oracle/_ref not built (needs /root/reference)")
def test_port_pair_laws_match_reference_function_pointers():
    cases = [("np_d2_f32", "newtonian", "plummer", None), ("np_yuk_f32", [["none", "yukawa"], ["yukawa", "none"]], [["none", "plummer"], ["plummer", "none"]], "yuk"),
             ("np_bam_f32", [["newtonian", "sourcebambaryon"], ["sourcebaryonbam", "bambam"]],
              [["plummer", "sourcebambaryon_spline"], ["sourcebaryonbam_spline", "bambam_spline"]], "bam")]
    rng = np.random.default_rng(0)
    for variant, acc, spl, kind in cases:
        box = 1000.0
        ref = RefOracle(variant, 64, boxsize=box)
        o = PortOracle(64, boxsize=box)
        par = np.zeros((2, 2, 4))
        par[:, :, 0] = 60.0 / box          # YUKAWA_IMASS / All.BoxSize, ngravs.c:41-43,858
        par[:, :, 1] = 1.31e-6             # BAM_EPSILON, ngravs.c:45-47
        o.set_laws(acc, spl, par)
        for _ in range(200):
            pm, m = rng.uniform(1e-6, 1e-3, 2) if kind == "bam" else rng.uniform(0.1, 10, 2)
            r = 10 ** rng.uniform(-3, 2)
            h = 10 ** rng.uniform(-2, 1)
            for i in range(2):
                for j in range(2):
                    assert o.accel(i, j, pm, m, r * r, r, 3) == ref.accel(i, j, pm, m, r * r, r, 3)
                    assert o.spline(i, j, pm, m, h, min(r, h), 3) == ref.spline(i, j, pm, m, h, min(r, h), 3)


@pytest.mark.skipif(not available("np_d2_f32"), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("case", ["hernquist", "pm64", "np_d1"])
def test_port_matches_reference_build(case):
    n = 12000
    if case == "pm64":
        box = 100000.0
        pos, mass, ptype = g2test.periodic_poisson(n, box, seed=11)
        eps = box / 23 / 30.0
        soft, grav = (eps,) * 6, g2test.GRAV_D2
        ref = RefOracle("pm64_d2_f32", int(1.1 * n) + 64, boxsize=box, softening=soft, gravity=grav)
        o = PortOracle(ref.maxpart, periodic=True, shortrange=True, unequal=False, boxsize=box, pmgrid=64, softening=soft, gravity=grav)
        o.set_srtable(ref.srtable())
    elif case == "np_d1":
        pos, mass, ptype = g2test.gaussian_blobs(n, seed=9)
        soft, grav = g2test.SOFT_NP, (0,) * 6
        ref = RefOracle("np_d1_f32", int(1.1 * n) + 64, softening=soft, gravity=grav)
        o = PortOracle(ref.maxpart, D=1, softening=soft, gravity=grav)
    else:
        pos, mass, ptype = g2test.hernquist(n, seed=5)
        soft, grav = g2test.SOFT_NP, g2test.GRAV_D2
        ref = RefOracle("np_d2_f32", int(1.1 * n) + 64, softening=soft, gravity=grav)
        o = PortOracle(ref.maxpart, softening=soft, gravity=grav)
    ref.load(pos, mass, ptype)
    ref.domain()
    rp = ref.particles()
    o.load(rp["pos"], rp["mass"], rp["type"])
    o.domain()
    assert np.array_equal(o.particles()["id"], np.arange(n))
    assert np.array_equal(o.keys(), ref.keys())
    ref.gravity()
    o.gravity()
    mism = g2test.compare_tree(o.tree(), ref.tree(), ref.D)
    assert all(v == 0 for v in mism.values()), mism
    r1, p1 = ref.particles(), o.particles()
    assert np.array_equal(p1["cost"], r1["cost"])
    assert np.array_equal(p1["acc"], r1["acc"].astype(np.float32))
    # direct summation of the port against the tree (accuracy sanity, reference: median 2.8e-3 on GalaxyCollision)
    if case == "hernquist":
        tg = np.arange(0, n, 97, dtype=np.int32)
        d = o.direct(tg)
        err = g2test.rel_err(p1["accd"][tg], d)
        assert np.median(err) < 2e-2


def test_port_reproduces_config1_galaxy_collision():
    """BASELINE config 1 (the reference's shipped example): port against the fixture made from the unmodified reference."""
    import hashlib
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config1_galaxycollision.npz"))
    o = PortOracle(int(g["maxpart"]), D=2, G=float(g["G"]), softening=g["soft"], gravity=g["grav"])
    o.load(g["in_pos"], g["in_mass"], g["in_type"])
    o.domain()
    p = o.particles()
    assert hashlib.sha256(p["key"].tobytes()).hexdigest() == str(g["keys_sha"])
    assert (p["id"] == g["order_id"]).mean() > 0.999
    oid = g["order_id"]
    o.load(g["in_pos"][oid], g["in_mass"][oid], g["in_type"][oid])
    o.domain()
    o.gravity()
    t = o.tree()
    assert t["numnodes"] == int(g["numnodes"])
    for k in ("len", "center", "s", "mass", "bitflags", "sibling", "nextnode", "father", "p_nextnode", "p_father"):
        assert hashlib.sha256(np.ascontiguousarray(t[k]).tobytes()).hexdigest() == str(g["sha_" + k]), k
    p1 = o.particles()
    assert np.array_equal(p1["cost"], g["bh_cost"].astype(np.float32))
    assert np.array_equal(p1["acc"], g["bh_acc"])
    assert np.array_equal(p1["oldacc"], g["bh_oldacc"])
