"""GPU parity of the tree potential walks: g2gpu_potential against
  * the committed fixtures tests/golden/pot_*.npz = P[].Potential of the UNMODIFIED reference's force_treeevaluate_potential_shortrange
    (forcetree.c:2789-3163) for every particle (potential.c:86-97), Barnes-Hut and relative criterion;
  * a fresh run of the reference where oracle/_ref travelled (D = 3, clustered, unequal masses);
  * the pinned oracle port at a larger size, with and without the table term on node terms (the NGRAVS_ACCUMULATOR variant);
  * FP64 direct sums for the non-PM potential walk (force_treeevaluate_potential, forcetree.c:2467), which the reference itself cannot
    compile, and a size-independent property at full size (the self term, and agreement between 1 rank and 2 rank slices).
Tolerance: FP32 terms on float inputs against the reference's double arithmetic: median relative error <= 1e-5, 99.9th percentile <= 1e-3
(the tolerances of the force walk, BASELINE.json north_star)."""
import os

import numpy as np
import pytest

import g2test
from portrun import PortOracle
from refrun import RefOracle, available

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MEDIAN_TOL = 1.0e-5
P999_TOL = 1.0e-3


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.abs(b)


def check(pot, ref, what):
    e = relerr(pot, ref)
    assert np.median(e) <= MEDIAN_TOL, (what, np.median(e))
    assert np.percentile(e, 99.9) <= P999_TOL, (what, np.percentile(e, 99.9), e.max())
    return e


def gpu_treepm(maxpart, D, grav, soft, srtable, srpot, box, accumulator=False):
    from g2gpu import TreeGravity
    t = TreeGravity(max_part=maxpart, n_gravs=D, periodic=True, shortrange=True, unequal_softenings=False)
    t.set_species(grav, g2test.force_softening(soft))
    t.set_laws()
    t.set_srtable(srtable)
    if accumulator:
        t.set_option("accumulator", 1)
    return t


@pytest.mark.parametrize("case", ["pot_pm64_d2_poisson4096", "pot_pm64_d4_poisson4096"])
def test_shortrange_potential_matches_reference_fixture(case):
    g = np.load(os.path.join(GOLD, case + ".npz"))
    box, n = float(g["box"]), len(g["mass"])
    t = gpu_treepm(int(g["maxpart"]), int(g["D"]), g["grav"], g["soft"], g["srtable"], g["srpot"], box)
    wp = t.walk_params(theta=0.5, errtol=0.005, boxsize=box, asmth=float(g["asmth"]), rcut=float(g["rcut"]))
    t.upload(g["pos"], g["mass"], g["type"], oldacc=g["oldacc"])
    t.domain()
    assert np.array_equal(t.order(), np.arange(n, dtype=np.int32))
    t.treebuild()
    # call order is checked: laws and table must be set first
    from g2gpu import G2Error
    with pytest.raises(G2Error):
        t.potential(wp)
    t.set_potential_laws("newtonian", "plummer")
    with pytest.raises(G2Error):
        t.potential(wp)
    t.set_srpot_table(g["srpot"])
    check(t.potential(wp), g["pot_bh"], "barnes-hut")
    wp.theta = 0.0
    check(t.potential(wp), g["pot_rel"], "relative")


@pytest.mark.skipif(not available("pm64_d3_f32"), reason="oracle/_ref not built")
def test_shortrange_potential_matches_fresh_reference_run():
    n, box = 20000, 100000.0
    rng = np.random.default_rng(23)
    pos = (rng.normal(size=(n, 3)) * box / 12 + box / 2).astype(np.float32) % np.float32(box)
    pos = np.minimum(pos, np.float32(np.nextafter(np.float32(box), np.float32(0))))
    mass = rng.uniform(0.2, 3.0, n).astype(np.float32)
    ptype = rng.integers(1, 6, n).astype(np.int32)
    eps = box / 27 / 30.0
    soft, grav = (eps,) * 6, (0, 0, 1, 2, 1, 0)
    ref = RefOracle("pm64_d3_f32", int(1.1 * n) + 64, boxsize=box, softening=soft, gravity=grav)
    ref.load(pos, mass, ptype)
    ref.domain()
    ref.gravity()
    rp = ref.particles()
    asmth, rcut = ref.pm_split()
    t = gpu_treepm(ref.maxpart, 3, grav, soft, ref.srtable(), ref.srpot_table(), box)
    t.set_potential_laws()
    t.set_srpot_table(ref.srpot_table())
    t.upload(rp["pos"], rp["mass"], rp["type"], oldacc=rp["oldacc"])
    t.domain()
    assert np.array_equal(t.order(), np.arange(n, dtype=np.int32))
    t.treebuild()
    for theta in (0.5, 0.0):
        ref.set_opening(theta, 0.005, 1)
        wp = t.walk_params(theta=theta, errtol=0.005, boxsize=box, asmth=asmth, rcut=rcut)
        check(t.potential(wp), ref.potential(), f"theta={theta}")


@pytest.mark.parametrize("accumulator", [False, True])
def test_shortrange_potential_against_pinned_port(accumulator):
    """200 k particles, D = 2, PMGRID 64 geometry; node terms with the table term (an NGRAVS_ACCUMULATOR build, forcetree.c:3134) and without."""
    n, box = 200000, 100000.0
    pos, mass, ptype = g2test.periodic_poisson(n, box, seed=3)
    mass = (mass * np.random.default_rng(9).uniform(0.5, 2.0, n)).astype(np.float32)
    g = np.load(os.path.join(GOLD, "pot_pm64_d2_poisson4096.npz"))
    eps = box / 58 / 30.0
    soft, grav = (eps,) * 6, g2test.GRAV_D2
    o = PortOracle(int(1.1 * n) + 64, D=2, periodic=True, shortrange=True, unequal=False, boxsize=box, pmgrid=64, softening=soft, gravity=grav)
    o.set_srtable(g["srtable"])
    o.set_srpot_table(g["srpot"])
    o.set_potential_laws("newtonian", "plummer", accumulator)
    o.load(pos, mass, ptype)
    o.domain()
    o.gravity(nthreads=8)                    # Barnes-Hut pass: OldAcc for the relative criterion
    p = o.particles()
    o.load(p["pos"], p["mass"], p["type"], oldacc=p["oldacc"])
    o.domain()
    o.treebuild()
    o.set_opening(0.0, 0.005)
    ref = o.potential(nthreads=8)
    t = gpu_treepm(o.maxpart, 2, grav, soft, g["srtable"], g["srpot"], box, accumulator=accumulator)
    t.set_potential_laws()
    t.set_srpot_table(g["srpot"])
    t.upload(p["pos"], p["mass"], p["type"], oldacc=p["oldacc"])
    t.domain()
    assert np.array_equal(t.order(), np.arange(n, dtype=np.int32))
    t.treebuild()
    wp = t.walk_params(theta=0.0, errtol=0.005, boxsize=box, asmth=o.asmth, rcut=o.rcut)
    check(t.potential(wp), ref, f"accumulator={accumulator}")


def test_nonperiodic_potential_against_direct_sum_and_port():
    """force_treeevaluate_potential (no PM): against the port's restatement of the reference text (not compilable there, so unpinned)
    and against what it approximates, the FP64 softened direct-sum potential."""
    from g2gpu import TreeGravity
    n = 100000
    pos, mass, ptype = g2test.hernquist(n)
    soft, grav = g2test.SOFT_NP, g2test.GRAV_D2
    o = PortOracle(int(1.1 * n) + 64, softening=soft, gravity=grav)
    o.set_potential_laws("newtonian", "plummer")
    o.load(pos, mass, ptype)
    o.domain()
    o.treebuild()
    o.set_opening(0.5, 0.005)
    p = o.particles()
    ref = o.potential(nthreads=8)
    t = TreeGravity(max_part=o.maxpart, n_gravs=2)
    t.set_species(grav, g2test.force_softening(soft))
    t.set_laws()
    t.set_potential_laws()
    t.upload(p["pos"], p["mass"], p["type"])
    t.domain()
    assert np.array_equal(t.order(), np.arange(n, dtype=np.int32))
    t.treebuild()
    pot = t.potential(t.walk_params(theta=0.5, errtol=0.005))
    check(pot, ref, "port")
    tg = np.arange(0, n, 997)
    d = g2test.direct_potential(p["pos"], p["mass"], g2test.force_softening(soft)[p["type"]], tg)
    e = relerr(pot[tg], d)
    assert np.median(e) < 2e-3 and e.max() < 3e-2, (np.median(e), e.max())


def test_periodic_box_without_pm_needs_its_tables():
    """PERIODIC without PMGRID needs the potcorr tables of lattice_pot_corr (forcetree.c:3895): without them the call must fail, not approximate."""
    from g2gpu import G2Error, TreeGravity
    n, box = 5000, 1000.0
    pos, mass, ptype = g2test.periodic_poisson(n, box, seed=5)
    t = TreeGravity(max_part=n + 64, n_gravs=2, periodic=True, shortrange=False, unequal_softenings=False)
    t.set_species(g2test.GRAV_D2, g2test.force_softening((1.0,) * 6))
    t.set_laws()
    t.set_potential_laws()
    t.upload(pos, mass, ptype)
    t.domain()
    t.treebuild()
    with pytest.raises(G2Error):
        t.potential(t.walk_params(theta=0.5, boxsize=box))


def test_periodic_potential_without_pm_against_port_and_ewald_sum():
    """force_treeevaluate_potential of a PERIODIC build without PMGRID: nearest-image terms + mass * lattice_pot_corr per term
    (forcetree.c:2736-2738, 2765-2767, 3895-3941).  The reference cannot compile that walk (oracle/ref/Makefile), so the checker is the
    port, whose table and look-up are pinned bit for bit against the reference build (tests/test_latticepot_oracle.py), and the exact
    periodic potential (FP64 Ewald sum).  The device tabulates ewald_psi itself (g2gpu_make_ewald_pot_table)."""
    from g2gpu import TreeGravity
    from portrun import make_ewald_pot_table
    from test_latticepot_oracle import ewald_potential
    n, box = 20000, 1000.0
    pos, mass, ptype = g2test.periodic_poisson(n, box, seed=12)
    mass = (mass * np.random.default_rng(4).uniform(0.5, 2.0, n)).astype(np.float32)
    soft, grav = (box / 27 / 3000.0,) * 6, g2test.GRAV_D2        # far below the closest pair, so that the Ewald sum of point masses applies
    o = PortOracle(int(1.1 * n) + 64, D=2, periodic=True, shortrange=False, unequal=False, boxsize=box, softening=soft, gravity=grav)
    o.set_potential_laws("newtonian", "plummer")
    ptab = make_ewald_pot_table(64) / box
    o.set_lattice_pot_tables(np.broadcast_to(ptab, (2, 2, 65, 65, 65)).copy())
    o.load(pos, mass, ptype)
    o.domain()
    o.treebuild()
    p = o.particles()
    t = TreeGravity(max_part=o.maxpart, n_gravs=2, periodic=True, shortrange=False, unequal_softenings=False)
    t.set_species(grav, g2test.force_softening(soft))
    t.set_laws()
    t.set_potential_laws()
    dtab = t.set_ewald_pot_lattice(box)
    assert np.abs(dtab - ptab).max() <= 1e-9 * np.abs(ptab).max()          # device erfc/exp/cos against libm
    t.upload(p["pos"], p["mass"], p["type"])
    t.domain()
    assert np.array_equal(t.order(), np.arange(n, dtype=np.int32))
    t.treebuild()
    for theta in (0.5, 0.3):
        o.set_opening(theta, 0.005)
        ref = o.potential(nthreads=8)
        pot = t.potential(t.walk_params(theta=theta, errtol=0.005, boxsize=box))
        check(pot, ref, f"port theta={theta}")
    # relative criterion (OldAcc of a force walk)
    acc, cost, old, perm = t.gravity_tree(p["pos"], p["mass"], p["type"], t.walk_params(theta=0.5, boxsize=box))
    assert np.array_equal(perm, np.arange(n))
    o.set_opening(0.0, 0.005)
    o.load(p["pos"], p["mass"], p["type"], oldacc=old)
    o.domain()
    o.treebuild()
    ref = o.potential(nthreads=8)
    t.upload(p["pos"], p["mass"], p["type"], oldacc=old)
    t.domain()
    t.treebuild()
    check(t.potential(t.walk_params(theta=0.0, errtol=0.005, boxsize=box)), ref, "port relative criterion")
    # against the exact periodic potential (self term removed like potential.c:250-254 does, plus the m LatticeZero / L of the table origin).
    # With an opening angle that opens every cell the walk is a direct sum over nearest images + table terms, and only the trilinear
    # interpolation of the 65^3 table separates it from the Ewald sum.  (At theta = 0.3 the monopole potential of the reference's walk is
    # biased by several per cent of the largest potential: every accepted cell is less bound than its particles, and nothing cancels that
    # in a box whose mean potential is zero.)
    tg = np.arange(0, n, 500)
    exact = ewald_potential(p["pos"], p["mass"], tg, box)
    h = 2.8 * soft[1]
    pot = t.potential(t.walk_params(theta=1.0e-3, errtol=0.005, boxsize=box)).astype(np.float64)
    noself = pot + p["mass"] * 2.8 / h - p["mass"] * float(np.float32(2.8372975)) / box
    e = np.abs(noself[tg] - exact) / np.abs(exact).max()
    assert e.max() < 5e-3, (np.median(e), e.max())
    pot = t.potential(t.walk_params(theta=0.3, errtol=0.005, boxsize=box)).astype(np.float64)
    # the reference's own table (pair [0][0] is the complete one in a FLOAT = float build) gives the same potentials
    if available("per_d2_f32"):
        r = RefOracle("per_d2_f32", int(1.1 * n) + 64, boxsize=box, softening=soft, gravity=grav)
        rt = r.potcorr_tables()[0, 0]
        t.set_lattice_pot_tables(np.broadcast_to(rt, (2, 2, 65, 65, 65)).copy())
        pot2 = t.potential(t.walk_params(theta=0.3, errtol=0.005, boxsize=box)).astype(np.float64)
        assert np.abs(pot2 - pot).max() <= 1e-6 * np.abs(pot).max()
    t.close()


def test_full_size_potential_properties(outdir):
    """Config 3 size (128^3 = 2.1 M particles, PMGRID 256): (1) two rank slices together give exactly the single-rank result; (2) the
    walk result is below the self term -m/eps for every particle (all pair terms of the Newtonian potential are attractive given the
    reference's tables) and finite; (3) 4096 targets agree with the pinned oracle port run on the same 2.1 M particles."""
    import json
    from g2gpu import TreeGravity
    n, box, pmgrid = 128 ** 3, 100000.0, 256
    pos, mass, ptype = g2test.periodic_poisson(n, box)
    g = np.load(os.path.join(GOLD, "pot_pm64_d2_poisson4096.npz"))
    eps = box / 128 / 30.0
    soft, grav = (eps,) * 6, g2test.GRAV_D2
    asmth = 1.25 * box / pmgrid
    rcut = 4.5 * asmth
    res = []
    for nranks in (1, 2):
        parts = []
        for rank in range(nranks):
            t = TreeGravity(max_part=n + 64, n_gravs=2, periodic=True, shortrange=True, unequal_softenings=False, rank=rank, nranks=nranks)
            t.set_species(grav, g2test.force_softening(soft))
            t.set_laws()
            t.set_srtable(g["srtable"])
            t.set_potential_laws()
            t.set_srpot_table(g["srpot"])
            t.upload(pos, mass, ptype)
            t.domain()
            t.treebuild()
            wp = t.walk_params(theta=0.5, errtol=0.005, boxsize=box, asmth=asmth, rcut=rcut)
            pot, ms = t.potential(wp, with_time=True)
            order = t.order()
            parts.append(pot)
            t.close()
        res.append((parts, ms))
    single = res[0][0][0]
    a, b = res[1][0]
    assert np.all((a == 0) | (b == 0))                        # disjoint slices
    assert np.array_equal(a + b, single)
    assert np.isfinite(single).all()
    self_term = -mass[order].astype(np.float64) / eps
    assert np.all(single < self_term * (1 - 1e-6))
    # (3) 4096 targets with the pinned oracle port on the same 2.1 M particles (the port's tree is the reference's tree, bit for bit)
    o = PortOracle(n + 64, D=2, periodic=True, shortrange=True, unequal=False, boxsize=box, pmgrid=pmgrid, softening=soft, gravity=grav)
    o.set_srtable(g["srtable"])
    o.set_srpot_table(g["srpot"])
    o.set_potential_laws("newtonian", "plummer")
    o.load(pos[order], mass[order], ptype[order])
    o.domain()
    assert np.array_equal(o.particles()["id"], np.arange(n))
    o.treebuild()
    o.set_opening(0.5, 0.005)
    tg = np.arange(0, n, n // 4096)[:4096].astype(np.int32)
    e = check(single[tg], o.potential_targets(tg), "full size")
    with open(os.path.join(outdir, "potential_p128.json"), "w") as f:
        json.dump(dict(n=n, kernel_ms=res[0][1], kernel_ms_2ranks=res[1][1], err_vs_port=dict(median=float(np.median(e)), p999=float(np.percentile(e, 99.9)), max=float(e.max()))), f)


@pytest.mark.parametrize("n", [2, 33, 1000])
def test_potential_known_answers_small_and_ragged_sets(n):
    """Sets smaller than a warp and not a multiple of 32: with theta -> 0 every cell is opened, so the tree potential
    IS the softened direct sum (self term -2.8 m/h = -m/eps included), to FP32 term accuracy."""
    from g2gpu import TreeGravity
    rng = np.random.default_rng(100 + n)
    pos = rng.uniform(-10, 10, size=(n, 3)).astype(np.float32)
    mass = rng.uniform(0.5, 2.0, n).astype(np.float32)
    ptype = np.where(np.arange(n) % 3 == 0, 2, 1).astype(np.int32)
    soft, grav = g2test.SOFT_NP, g2test.GRAV_D2
    t = TreeGravity(max_part=max(n + 64, 1024), n_gravs=2)      # MaxNodes = 1.5 MaxPart must hold the top-level tree as well
    t.set_species(grav, g2test.force_softening(soft))
    t.set_laws()
    t.set_potential_laws()
    t.upload(pos, mass, ptype)
    t.domain()
    order = t.order()
    t.treebuild()
    pot = t.potential(t.walk_params(theta=1.0e-3, errtol=0.005)).astype(np.float64)
    p, m, ty = pos[order], mass[order], ptype[order]
    ref = g2test.direct_potential(p, m, g2test.force_softening(soft)[ty], np.arange(n))
    self_term = -m / np.asarray(soft)[ty]                       # -2.8 m / h, what potential.c:250-254 removes afterwards
    assert np.all(np.abs(pot - ref) <= 2e-6 * np.abs(ref)), (pot, ref, self_term)
    assert np.all(pot < self_term), (pot, ref, self_term)


# ---------------------------------------------------------------- BAM potentials (ngravs.c:672-760) -----------------------
BAM_POT = [["newtonian", "sourcebambaryon"], ["sourcebaryonbam", "bambam"]]           # ngravs.c:196-200
BAM_POTSPLINE = [["plummer", "sourcebambaryon"], ["sourcebaryonbam", "bambam"]]
BAM_LAWS = dict(accel=[["newtonian", "sourcebambaryon"], ["sourcebaryonbam", "bambam"]],
                spline=[["plummer", "sourcebambaryon_spline"], ["sourcebaryonbam_spline", "bambam_spline"]])


def _bam_params(D=2):
    par = np.zeros((D, D, 4))
    par[:, :, 1] = 1.31e-6                                 # BAM_EPSILON (ngravs.c:45-47)
    return par


@pytest.mark.skipif(not available("np_bam_f32"), reason="oracle/_ref not built")
def test_bam_potentials_pointwise_against_the_reference_functions():
    """PotentialFxns / PotentialSplines of the BAM wiring (ngravs.c:196-200), pair by pair, against the reference's own functions
    (called by name: a PMGRID build of the reference with that wiring cannot be run, see oracle/ref/harness.c::g2ref_named_pot, so the walk
    itself has no reference to compare with)."""
    from g2gpu import TreeGravity
    ref = RefOracle("np_bam_f32", 1000, softening=(0.05,) * 6, gravity=g2test.GRAV_D2)
    t = TreeGravity(max_part=1000, n_gravs=2, periodic=True, shortrange=True, unequal_softenings=False)
    t.set_laws(BAM_LAWS["accel"], BAM_LAWS["spline"], params=_bam_params())
    t.set_potential_laws(BAM_POT, BAM_POTSPLINE)
    rng = np.random.default_rng(4)
    n = 4000
    pm = rng.uniform(1e-6, 1e-3, n).astype(np.float32)
    m = rng.uniform(1e-6, 1e-3, n).astype(np.float32)
    r = (10.0 ** rng.uniform(-1, 4.5, n)).astype(np.float32)        # r*eta from << 0.1 (Taylor branch) to >> 1
    nn = rng.integers(1, 50, n).astype(np.int32)
    which = {(0, 0): (3, 4), (0, 1): (2, 2), (1, 0): (1, 1), (1, 1): (0, 0)}      # (function, "spline") of the pair, ngravs.c:196-200
    for (tg, sg), (wf, ws) in which.items():
        for h in (0.0, 1.0e9):                                      # r >= h: -PotentialFxns ; r < h: +PotentialSplines
            hh = np.full(n, h if h == 0.0 else 1.0e9, dtype=np.float32)
            got = t.eval_potentials(tg, sg, pm, m, r, hh, nn)
            if h == 0.0:
                want = np.array([-ref.named_pot(wf, float(pm[i]), float(m[i]), h, float(r[i]), int(nn[i])) for i in range(n)])
            else:
                want = np.array([ref.named_pot(ws, float(pm[i]), float(m[i]), float(hh[i]), float(r[i]), int(nn[i])) for i in range(n)])
            e = relerr(got, want)
            assert e.max() <= 2.0e-6, (tg, sg, h, e.max())
    t.close()
