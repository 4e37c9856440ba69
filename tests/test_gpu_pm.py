"""GPU: periodic PM long-range force (g2gpu_pm_periodic, csrc/g2_pm.cu) against the unmodified reference's pmforce_periodic
(pm_periodic.c:204-790; committed fixtures tests/golden/pm_*.npz and, where oracle/_ref travelled, a fresh run of it) and
against the pinned numpy restatement oracle/pm_port.py at the full 256^3 mesh.

Tolerance (floating point; the mesh arithmetic is FP64 on both sides, the result a FLOAT): every component within 1e-5 of the
largest |GravPM| component, median relative error per particle <= 1e-5.  The GPU sums the species-pair contributions in k-space
and rounds to FLOAT once, the reference rounds P[].GravPM after every pair (pm_periodic.c:780)."""
import os

import numpy as np
import pytest

import g2test
import pm_port
from refrun import RefOracle, available

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ABS_TOL = 1.0e-5       # of max |GravPM| component
MEDIAN_TOL = 1.0e-5


def gpu_pm(pos, mass, ptype, grav, D, pmgrid, box, G=1.0, **kw):
    from g2gpu import TreeGravity
    n = len(mass)
    tg = TreeGravity(max_part=n + 64, n_gravs=D, periodic=True, shortrange=True, unequal_softenings=False)
    tg.set_species(grav, g2test.force_softening((1.0,) * 6))
    tg.upload(pos, mass, ptype)
    pm = tg.pm_periodic(pmgrid, box, G=G, **kw)
    t = tg.timings()
    tg.close()
    return pm, t


def check(pm, ref):
    scale = np.abs(ref).max()
    assert np.isfinite(pm).all()
    assert np.abs(pm - ref).max() <= ABS_TOL * scale, np.abs(pm - ref).max() / scale
    assert np.median(g2test.rel_err(pm, ref)) <= MEDIAN_TOL


@pytest.mark.parametrize("case,D", [("pm_pm64_d2_poisson4096", 2), ("pm_pm64_d4_poisson4096", 4)])
def test_pm_matches_reference_fixture(case, D):
    g = np.load(os.path.join(GOLD, case + ".npz"))
    pm, _ = gpu_pm(g["pos"], g["mass"], g["type"], g["grav"], D, int(g["pmgrid"]), float(g["box"]), float(g["G"]))
    check(pm, g["gravpm"])
    # upload order must not matter (the reference needs species blocks, the device version does not)
    perm = np.random.default_rng(3).permutation(len(g["mass"]))
    pm2, _ = gpu_pm(g["pos"][perm], g["mass"][perm], g["type"][perm], g["grav"], D, int(g["pmgrid"]), float(g["box"]), float(g["G"]))
    check(pm2, g["gravpm"][perm])


def test_pm_matches_reference_run_d3():
    if not available("pm64_d3_f32"):
        pytest.skip("oracle/_ref not built")
    n, box = 20000, 80000.0
    pos, mass, ptype = g2test.periodic_poisson(n, box, seed=11, ntypes=6)
    grav = (0, 1, 2, 1, 2, 1)
    ref = RefOracle("pm64_d3_f32", int(1.1 * n) + 64, boxsize=box, softening=(50.0,) * 6, gravity=grav, G=43007.1)
    ref.load(pos, mass, ptype)
    ref.domain()
    rpm = ref.pmforce().astype(np.float32)
    rp = ref.particles()
    pm, _ = gpu_pm(rp["pos"].astype(np.float32), rp["mass"].astype(np.float32), rp["type"], grav, 3, 64, box, G=43007.1)
    check(pm, rpm)


def test_pm_full_mesh_256_against_port(outdir):
    """BASELINE config 3 mesh (PMGRID = 256) with 2^18 particles on two species; the numpy port is the checker."""
    n, box, N = 262144, 100000.0, 256
    pos, mass, ptype = g2test.periodic_poisson(n, box, seed=21)
    species = np.asarray(g2test.GRAV_D2)[ptype]
    port = pm_port.pm_force(pos, mass, species, 2, N, box, 1.0, 1.25 * box / N)
    pm, t = gpu_pm(pos, mass, ptype, g2test.GRAV_D2, 2, N, box)
    check(pm, port)
    tot = (mass[:, None].astype(np.float64) * pm).sum(axis=0)
    assert np.abs(tot).max() <= 1e-4 * np.abs(mass[:, None] * pm).sum()       # momentum conservation
    print(f"PM 256^3, {n} particles: {t['pm_ms']:.3f} ms")


def test_pm_yukawa_greens_and_none():
    """Per-pair Green's functions: Yukawa between the species, nothing within species 1 (GreensFxns wiring of ngravs.c:260-275)."""
    n, box, N = 5000, 1000.0, 64
    pos, mass, ptype = g2test.periodic_poisson(n, box, seed=4)
    species = np.asarray(g2test.GRAV_D2)[ptype]
    ym = 60.0 / (2 * np.pi)                     # YUKAWA_IMASS / (2 pi), ngravs.c:41-43, 871
    asmth = 1.25 * box / N
    asmth2 = ((2 * np.pi) * asmth / box) ** 2
    names = [["pgdelta", "pgyukawa"], ["pgyukawa", "none"]]

    def greens(a, b, k2):
        with np.errstate(divide="ignore"):
            return {"pgdelta": 1.0 / k2, "pgyukawa": 1.0 / (k2 + ym * ym) * np.exp(-ym * ym * asmth2), "none": np.zeros_like(k2)}[names[a][b]]
    port = pm_port.pm_force(pos, mass, species, 2, N, box, 1.0, asmth, greens=greens)
    pm, _ = gpu_pm(pos, mass, ptype, g2test.GRAV_D2, 2, N, box, greens=names, greens_par=[[0, ym], [ym, 0]])
    check(pm, port)


def test_pm_requires_upload():
    from g2gpu import TreeGravity, G2Error
    tg = TreeGravity(max_part=100, n_gravs=2, periodic=True, shortrange=True)
    with pytest.raises(G2Error):
        tg.pm_periodic(64, 100.0)
    tg.close()


def test_treepm_total_force_against_ewald(outdir):
    """Tree short-range walk + PM long-range force = the complete TreePM force (what gravity_forcetest writes as GravAccel + GravPM,
    gravtree_forcetest.c:304-310), against exact Ewald sums for 32 targets.  32^3 particles on a 64^3 mesh: the particles-per-mesh-
    cell density of BASELINE config 3.  The TreePM scheme itself (ASMTH 1.25, RCUT 4.5, CIC mesh) is accurate to about a per cent."""
    from g2gpu import TreeGravity
    PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gadget-2.0.7-ngravs_b200")
    n, box, N = 32768, 100000.0, 64
    pos, mass, ptype = g2test.periodic_poisson(n, box, seed=31)
    eps = box / 32 / 30.0
    tg = TreeGravity(max_part=n + 64, n_gravs=2, periodic=True, shortrange=True, unequal_softenings=False)
    tg.set_species(g2test.GRAV_D2, g2test.force_softening((eps,) * 6))
    tg.set_laws()
    tab = np.load(os.path.join(PKG, "data", "srtable_newton_ntab2048.npy"))
    tg.set_srtable(np.broadcast_to(tab, (2, 2, len(tab))).copy())
    asmth = 1.25 * box / N
    tg.upload(pos, mass, ptype)
    pm = tg.pm_periodic(N, box, G=1.0)
    acc, cost, old, perm = tg.gravity_tree(pos, mass, ptype, tg.walk_params(theta=0.3, errtol=0.005, boxsize=box, G=1.0, asmth=asmth, rcut=4.5 * asmth))
    tree = np.zeros_like(acc)
    tree[perm] = acc
    tg.close()
    targets = np.random.default_rng(2).choice(n, 32, replace=False)
    exact = g2test.ewald_direct(pos, mass, targets, box)
    total = tree[targets].astype(np.float64) + pm[targets]
    err = np.linalg.norm(total - exact, axis=1) / np.linalg.norm(exact, axis=1)
    err_tree_only = np.linalg.norm(tree[targets] - exact, axis=1) / np.linalg.norm(exact, axis=1)
    with open(os.path.join(outdir, "treepm_vs_ewald.txt"), "w") as f:
        f.write(f"TreePM (tree + PM) vs Ewald: median {np.median(err):.3e} max {err.max():.3e}; tree alone: median {np.median(err_tree_only):.3e}\n")
    print(f"TreePM vs Ewald: median {np.median(err):.3e} max {err.max():.3e}; tree alone median {np.median(err_tree_only):.3e}")
    assert np.median(err) <= 1e-2 and err.max() <= 8e-2
    assert np.median(err_tree_only) > 5 * np.median(err)          # the long-range part matters


def test_device_ewald_direct_sum_against_numpy():
    """g2gpu_direct with the option "direct_ewald": the complete periodic Newtonian force (nearest image + exact lattice correction of all
    images, the quantity gravity_forcetest compares the TreePM force with) against the independent numpy Ewald sum of g2test."""
    from g2gpu import TreeGravity
    n, box = 4096, 1000.0
    pos, mass, ptype = g2test.periodic_poisson(n, box, seed=8)
    mass = (mass * np.random.default_rng(1).uniform(0.5, 2.0, n)).astype(np.float32)
    tg = TreeGravity(max_part=n + 2000, n_gravs=2, periodic=True, shortrange=False, unequal_softenings=False)
    tg.set_species(g2test.GRAV_D2, g2test.force_softening((1.0e-6,) * 6))
    tg.set_laws()
    tg.set_option("direct_ewald", 1)
    tg.upload(pos, mass, ptype)
    tg.domain()
    order = tg.order()
    targets = np.arange(0, n, n // 16, dtype=np.int32)[:16]
    d = tg.direct(tg.walk_params(theta=0.5, errtol=0.005, boxsize=box, G=1.0), targets)
    tg.close()
    exact = g2test.ewald_direct(pos, mass, order[targets], box)
    err = np.linalg.norm(d - exact, axis=1) / np.linalg.norm(exact, axis=1)
    assert err.max() < 1e-8, err


@pytest.mark.parametrize("side,ntargets", [(128, 128), (256, 32)])
def test_full_size_treepm_total_force_against_ewald(side, ntargets, outdir):
    """BASELINE configs 3 and 5 at full size (2.1 M / 16.8 M particles, PMGRID 256): tree short-range walk (relative criterion) + device PM
    = the complete TreePM force, against the exact periodic force of random targets (device FP64 direct sum over ALL particles with the
    Ewald lattice correction).  What is left is the accuracy of the TreePM scheme itself (about a per cent, Springel 2005 fig. 3)."""
    import json
    from g2gpu import TreeGravity
    PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gadget-2.0.7-ngravs_b200")
    n, box, N = side ** 3, 100000.0, 256
    pos, mass, ptype = g2test.periodic_poisson(n, box)
    eps = box / side / 30.0
    tg = TreeGravity(max_part=int(1.1 * n) + 64, n_gravs=2, periodic=True, shortrange=True, unequal_softenings=False)
    tg.set_species(g2test.GRAV_D2, g2test.force_softening((eps,) * 6))
    tg.set_laws()
    tg.set_option("direct_ewald", 1)
    tab = np.load(os.path.join(PKG, "data", "srtable_newton_ntab2048.npy"))
    tg.set_srtable(np.broadcast_to(tab, (2, 2, len(tab))).copy())
    asmth = 1.25 * box / N
    wp = dict(boxsize=box, G=1.0, asmth=asmth, rcut=4.5 * asmth)
    tg.upload(pos, mass, ptype)
    pm = tg.pm_periodic(N, box, G=1.0)                                  # upload order
    acc0, cost0, old0, perm = tg.gravity_tree(pos, mass, ptype, tg.walk_params(theta=0.5, errtol=0.005, **wp))
    oldacc_by_id = np.zeros(n, dtype=np.float32)
    oldacc_by_id[perm] = old0
    acc, cost, old, perm = tg.gravity_tree(pos, mass, ptype, tg.walk_params(theta=0.0, errtol=0.005, **wp), oldacc=oldacc_by_id)
    targets = np.sort(np.random.default_rng(11).choice(n, ntargets, replace=False)).astype(np.int32)      # current-order indices
    exact = tg.direct(tg.walk_params(theta=0.0, errtol=0.005, boxsize=box, G=1.0, asmth=0.0, rcut=0.0), targets)
    tg.close()
    total = acc[targets].astype(np.float64) + pm[perm[targets]]
    err = np.linalg.norm(total - exact, axis=1) / np.linalg.norm(exact, axis=1)
    json.dump(dict(n=n, ntargets=ntargets, median=float(np.median(err)), p90=float(np.percentile(err, 90)), max=float(err.max())),
              open(os.path.join(outdir, f"fullsize_treepm_total_vs_ewald_{side}.json"), "w"))
    print(f"TreePM total vs Ewald at {side}^3: median {np.median(err):.3e} p90 {np.percentile(err, 90):.3e} max {err.max():.3e}")
    assert np.median(err) < 1.5e-2 and err.max() < 0.15


# ---- pmpotential_periodic (pm_periodic.c:798), g2gpu_pm_potential_periodic ------------------------------------------------------------
def gpu_pmpot(pos, mass, ptype, grav, D, pmgrid, box, G=1.0, **kw):
    from g2gpu import TreeGravity
    n = len(mass)
    tg = TreeGravity(max_part=n + 64, n_gravs=D, periodic=True, shortrange=True, unequal_softenings=False)
    tg.set_species(grav, g2test.force_softening((1.0,) * 6))
    tg.upload(pos, mass, ptype)
    pot = tg.pm_potential_periodic(pmgrid, box, G=G, **kw)
    tg.close()
    return pot


YUK_GREENS = [["none", "pgyukawa"], ["pgyukawa", "none"]]          # NGRAVS_YUKAWA_FORCETEST (ngravs.c:264-272)


def yuk_par(ymass=60.0):
    return np.full((2, 2), ymass / (2 * np.pi))                   # ngravs.c:871


def check_pot(pot, ref):
    """tolerance: 1e-5 of the largest |potential| and 1e-5 median relative (FP64 mesh on both sides, FLOAT result; the device sums the
    source species in k-space and rounds once, the reference rounds after every corner of every pair, pm_periodic.c:1254-1267)."""
    scale = np.abs(ref).max()
    assert np.isfinite(pot).all() and scale > 0
    assert np.abs(pot - ref).max() <= ABS_TOL * scale, np.abs(pot - ref).max() / scale
    assert np.median(np.abs(pot - ref) / np.maximum(np.abs(ref), 1e-30)) <= MEDIAN_TOL


def test_pm_potential_matches_reference_fixture():
    g = np.load(os.path.join(GOLD, "pmpot_pm64_yuk_poisson4096.npz"))
    args = (g["grav"], 2, int(g["pmgrid"]), float(g["box"]), float(g["G"]))
    pot = gpu_pmpot(g["pos"], g["mass"], g["type"], *args, greens=YUK_GREENS, greens_par=yuk_par(float(g["yukawa_imass"])))
    check_pot(pot, g["pmpot"])
    perm = np.random.default_rng(3).permutation(len(g["mass"]))     # upload order must not matter
    pot2 = gpu_pmpot(g["pos"][perm], g["mass"][perm], g["type"][perm], *args, greens=YUK_GREENS, greens_par=yuk_par(float(g["yukawa_imass"])))
    check_pot(pot2, g["pmpot"][perm])


def test_pm_potential_matches_reference_run():
    if not available("pm64_yuk_f32"):
        pytest.skip("oracle/_ref not built")
    n, box, G = 20000, 80000.0, 43007.1
    pos, mass, ptype = g2test.periodic_poisson(n, box, seed=11)
    mass = (1.0 + 0.5 * np.random.default_rng(1).random(n)).astype(np.float32)
    ref = RefOracle("pm64_yuk_f32", int(1.1 * n) + 64, boxsize=box, softening=(50.0,) * 6, gravity=g2test.GRAV_D2, G=G)
    ref.load(pos, mass, ptype)
    ref.domain()
    rpot = ref.pmpotential().astype(np.float32)
    rp = ref.particles()
    pot = gpu_pmpot(rp["pos"].astype(np.float32), rp["mass"].astype(np.float32), rp["type"], g2test.GRAV_D2, 2, 64, box, G=G, greens=YUK_GREENS, greens_par=yuk_par())
    check_pot(pot, rpot)


def test_pm_potential_256_mesh_against_port_and_newtonian_zero_mode():
    """PMGRID = 256 (BASELINE config 3 mesh), 2^17 particles: Yukawa pairs against the pinned numpy restatement; and the stock 1/k^2, for
    which the reference returns infinite potentials (tests/test_pm_oracle.py): the device result is finite and equals the port's with the
    k = 0 mode removed -- a potential is defined up to a constant, and differences of it are what the reference's force mesh uses."""
    n, box, N = 1 << 17, 100000.0, 256
    pos, mass, ptype = g2test.periodic_poisson(n, box, seed=5)
    species = np.asarray(g2test.GRAV_D2)[ptype]
    asmth = 1.25 * box / N
    ym = 60.0 / (2 * np.pi)
    asmth2 = ((2 * np.pi) * asmth / box) ** 2
    yuk = lambda a, b, k2: (1.0 / (k2 + ym * ym) * np.exp(-ym * ym * asmth2)) if a != b else np.zeros_like(k2)     # noqa: E731
    port = pm_port.pm_potential(pos, mass, species, 2, N, box, 1.0, asmth, yuk)
    pot = gpu_pmpot(pos, mass, ptype, g2test.GRAV_D2, 2, N, box, greens=YUK_GREENS, greens_par=yuk_par())
    check_pot(pot, port)

    def newton_without_zero_mode(a, b, k2):
        with np.errstate(divide="ignore"):
            return np.where(k2 > 0, 1.0 / k2, 0.0)
    port = pm_port.pm_potential(pos, mass, species, 2, N, box, 1.0, asmth, newton_without_zero_mode)
    pot = gpu_pmpot(pos, mass, ptype, g2test.GRAV_D2, 2, N, box)
    check_pot(pot, port)
