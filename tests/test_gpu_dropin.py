"""GPU parity of the DROP-IN: the unmodified reference's own call chain -- domain_Decomposition() (domain.c, CPU) ->
peano_hilbert_order() -> gravity_tree() -> force_treebuild() -- linked with the product's host shim
(gadget-2.0.7-ngravs_b200/host/g2_shim.c, built by integration/Makefile) instead of gravtree.c / peano.c / the replaced
entry points of forcetree.c, against the same chain of the pure-CPU reference build."""
import ctypes as C
import os

import numpy as np
import pytest

import g2test
from refrun import RefOracle, available

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _preload():
    C.CDLL(os.path.join(ROOT, "gadget-2.0.7-ngravs_b200", "libg2gpu.so"), mode=C.RTLD_GLOBAL)


@pytest.mark.parametrize("variant", ["np_d2_f32", "pm64_d2_f32", "per_d2_f32"])
def test_reference_call_chain_runs_on_the_gpu(variant, outdir, monkeypatch):
    if not (available(variant) and available(variant, "g2shim")):
        pytest.skip("oracle/_ref (reference and shim builds) not present")
    monkeypatch.setenv("G2GPU_HOST_MIRROR", "1")
    _preload()
    n = 40000 if variant == "np_d2_f32" else 32768
    if variant == "np_d2_f32":
        pos, mass, ptype = g2test.hernquist(n, seed=77)
        soft, box = g2test.SOFT_NP, 0.0
    elif variant == "per_d2_f32":
        # PERIODIC without PMGRID: gravity_tree() = tree walk + lattice-sum correction walk (forcetree.c:1606-1608).  Both sides
        # tabulate ewald_force themselves (the reference in lattice_init, ~30 s; the shim on the device).  In this float build the
        # reference's tables of the pairs other than [0][0] are half empty (oracle/refrun.py), the shim's are complete, so only
        # species-0 targets and sources may be compared: all particles are given type 1 (species 0).
        box = 1000.0
        pos, mass, ptype = g2test.periodic_poisson(n, box, seed=13)
        ptype[:] = 1
        soft = (box / 32 / 30.0,) * 6
    else:
        box = 100000.0
        pos, mass, ptype = g2test.periodic_poisson(n, box, seed=13)
        soft = (box / 32 / 30.0,) * 6
    kw = dict(boxsize=box, softening=soft, gravity=g2test.GRAV_D2)
    ref = RefOracle(variant, int(1.1 * n) + 64, **kw)
    shim = RefOracle(variant, int(1.1 * n) + 64, prefix="g2shim", **kw)
    vel = np.random.default_rng(3).normal(size=(n, 3)).astype(np.float32)
    for o in (ref, shim):
        o.load(pos, mass, ptype, vel=vel)
        o.domain()                 # reference domain.c on the CPU; peano_hilbert_order() on the GPU for the shim build
    rp, sp = ref.particles(), shim.particles()
    # same particle order (ties between equal keys excepted: the reference's qsort leaves them unordered)
    same = rp["id"] == sp["id"]
    keys = ref.keys()
    if not same.all():
        tie = np.zeros(n, dtype=bool)
        eq = keys[1:] == keys[:-1]
        tie[1:] |= eq
        tie[:-1] |= eq
        assert tie[~same].all()
    # first force computation of a run: Barnes-Hut, then the relative criterion (accel.c:46-49)
    for o in (ref, shim):
        o.gravity()
    r1, s1 = ref.particles(), shim.particles()
    assert shim.lib.g2ref_numnodes() == ref.lib.g2ref_numnodes()
    if same.all():
        st, rt = shim.tree(), ref.tree()
        mism = g2test.compare_tree(st, rt, ref.D)     # host mirror Nodes[]/Nextnode[]/Father[] filled by the shim
        mism["vs"] = int(np.sum(st["vs"] != rt["vs"]))   # Extnodes[].vs
        assert all(v == 0 for v in mism.values()), mism
    ra = np.zeros((n, 3)); ra[r1["id"]] = r1["acc"]
    sa = np.zeros((n, 3)); sa[s1["id"]] = s1["acc"]
    e1 = g2test.rel_err(sa, ra)
    rc = np.zeros(n); rc[r1["id"]] = r1["cost"]
    sc = np.zeros(n); sc[s1["id"]] = s1["cost"]
    for o in (ref, shim):
        o.set_opening(0.0, 0.005, 1)
        o.force_rebuild()
        o.gravity()
    r2, s2 = ref.particles(), shim.particles()
    ra2 = np.zeros((n, 3)); ra2[r2["id"]] = r2["acc"]
    sa2 = np.zeros((n, 3)); sa2[s2["id"]] = s2["acc"]
    e2 = g2test.rel_err(sa2, ra2)
    with open(os.path.join(outdir, f"dropin_{variant}.txt"), "w") as f:
        f.write(f"bh median {np.median(e1):.3e} p99.9 {np.percentile(e1, 99.9):.3e} cost_mismatch {int((rc != sc).sum())}\n")
        f.write(f"rel median {np.median(e2):.3e} p99.9 {np.percentile(e2, 99.9):.3e}\n")
    for e in (e1, e2):
        assert np.median(e) <= 1e-6 and np.percentile(e, 99.9) <= 3e-4
    assert (rc != sc).sum() == 0


def test_reference_pm_call_runs_on_the_gpu(outdir):
    """pm_init_periodic() + pmforce_periodic() of the shim (device PM) against the unmodified pm_periodic.c, then the reference's own
    gravity_tree() on top of it: OldAcc contains GravPM/G (gravtree.c:318-331) on both sides."""
    variant = "pm64_d2_f32"
    if not (available(variant) and available(variant, "g2shim")):
        pytest.skip("oracle/_ref (reference and shim builds) not present")
    _preload()
    n, box = 32768, 100000.0
    pos, mass, ptype = g2test.periodic_poisson(n, box, seed=17)
    kw = dict(boxsize=box, softening=(box / 32 / 30.0,) * 6, gravity=g2test.GRAV_D2, G=43007.1)
    ref = RefOracle(variant, int(1.1 * n) + 64, **kw)
    shim = RefOracle(variant, int(1.1 * n) + 64, prefix="g2shim", **kw)
    out = {}
    for name, o in (("ref", ref), ("shim", shim)):
        o.load(pos, mass, ptype)
        o.domain()
        pm = o.pmforce()
        o.gravity()
        p = o.particles()
        gp = np.zeros((n, 3)); gp[p["id"]] = pm
        old = np.zeros(n); old[p["id"]] = p["oldacc"]
        out[name] = (gp, old)
    scale = np.abs(out["ref"][0]).max()
    assert np.abs(out["shim"][0] - out["ref"][0]).max() <= 1e-5 * scale
    assert np.median(g2test.rel_err(out["shim"][0], out["ref"][0])) <= 1e-5
    rel = np.abs(out["shim"][1] - out["ref"][1]) / np.maximum(out["ref"][1], 1e-30)
    assert np.median(rel) <= 1e-5 and np.percentile(rel, 99.9) <= 1e-3


def test_reference_pm_potential_call_runs_on_the_gpu(outdir):
    """pmpotential_periodic() of the shim (one device pass, g2gpu_pm_potential_periodic) against the unmodified pm_periodic.c:798, in the
    Yukawa wiring, whose k = 0 Green's function is finite (with 1/k^2 the reference's own result is infinite)."""
    variant = "pm64_yuk_f32"
    if not (available(variant) and available(variant, "g2shim")):
        pytest.skip("oracle/_ref (reference and shim builds) not present")
    _preload()
    n, box = 32768, 100000.0
    pos, mass, ptype = g2test.periodic_poisson(n, box, seed=17)
    mass = (1.0 + 0.5 * np.random.default_rng(1).random(n)).astype(np.float32)
    kw = dict(boxsize=box, softening=(box / 32 / 30.0,) * 6, gravity=g2test.GRAV_D2, G=43007.1)
    out = {}
    for name, prefix in (("ref", "g2ref"), ("shim", "g2shim")):
        o = RefOracle(variant, int(1.1 * n) + 64, prefix=prefix, **kw)
        o.load(pos, mass, ptype)
        o.domain()
        pot = o.pmpotential()
        p = o.particles()
        full = np.zeros(n); full[p["id"]] = pot
        out[name] = full
    scale = np.abs(out["ref"]).max()
    assert np.isfinite(out["ref"]).all() and scale > 0
    assert np.abs(out["shim"] - out["ref"]).max() <= 1e-5 * scale
    assert np.median(np.abs(out["shim"] - out["ref"]) / np.maximum(np.abs(out["ref"]), 1e-30)) <= 1e-5


def test_reference_potential_loop_runs_on_the_gpu(outdir):
    """The loop of compute_potential() (potential.c:86-97: force_treeevaluate_potential_shortrange(i, 0) for every particle) through the
    shim's entry point -- one device walk of all particles behind the reference's per-target function -- against the unmodified
    forcetree.c:2789, Barnes-Hut and relative criterion (OldAcc from the reference's own gravity_tree() on both sides)."""
    variant = "pm64_d2_f32"
    if not (available(variant) and available(variant, "g2shim")):
        pytest.skip("oracle/_ref (reference and shim builds) not present")
    _preload()
    n, box = 32768, 100000.0
    pos, mass, ptype = g2test.periodic_poisson(n, box, seed=19)
    mass = (mass * np.random.default_rng(2).uniform(0.5, 2.0, n)).astype(np.float32)
    kw = dict(boxsize=box, softening=(box / 32 / 30.0,) * 6, gravity=g2test.GRAV_D2)
    ref = RefOracle(variant, int(1.1 * n) + 64, **kw)
    shim = RefOracle(variant, int(1.1 * n) + 64, prefix="g2shim", **kw)
    res = {}
    for name, o in (("ref", ref), ("shim", shim)):
        o.load(pos, mass, ptype)
        o.domain()
        o.gravity()
        ids = o.particles()["id"]
        for theta in (0.5, 0.0):
            o.set_opening(theta, 0.005, 1)
            pot = np.zeros(n)
            pot[ids] = o.potential()
            res[name, theta] = pot
    with open(os.path.join(outdir, "dropin_potential.txt"), "w") as f:
        for theta in (0.5, 0.0):
            e = np.abs(res["shim", theta] - res["ref", theta]) / np.abs(res["ref", theta])
            f.write(f"theta {theta}: median {np.median(e):.3e} p99.9 {np.percentile(e, 99.9):.3e} max {e.max():.3e}\n")
            assert np.median(e) <= 1e-5 and np.percentile(e, 99.9) <= 1e-3


def test_reference_forcetest_call_runs_on_the_gpu(outdir, tmp_path, monkeypatch):
    """gravity_forcetest() (gravtree_forcetest.c:28, the -DFORCETEST accuracy check accel.c:52 runs after gravity_tree()): the shim's
    version (one device call, FP64 direct sums) against the unmodified file, and the forcetest.txt lines both write."""
    variant = "np_d2_f32_ft"
    if not (available(variant) and available(variant, "g2shim")):
        pytest.skip("oracle/_ref (reference and shim builds) not present")
    _preload()
    monkeypatch.chdir(tmp_path)
    n = 3000
    pos, mass, ptype = g2test.gaussian_blobs(n, seed=5)
    kw = dict(softening=g2test.SOFT_NP, gravity=g2test.GRAV_D2, G=43007.1)
    res = {}
    for name, prefix in (("ref", "g2ref"), ("shim", "g2shim")):
        d = tmp_path / name
        d.mkdir()
        monkeypatch.chdir(d)
        o = RefOracle(variant, int(1.1 * n) + 64, prefix=prefix, **kw)
        o.load(pos, mass, ptype)
        o.domain()
        o.gravity()
        direct = o.run_forcetest()
        p = o.particles()
        full = np.zeros((n, 3)); full[p["id"]] = direct
        tree = np.zeros((n, 3)); tree[p["id"]] = p["acc"]
        lines = open(d / "forcetest.txt").read().splitlines()
        res[name] = (full, tree, lines)
    err = g2test.rel_err(res["shim"][0], res["ref"][0])
    assert err.max() < 1e-5, err.max()                       # FP32 pair arithmetic inputs, FP64 sums on both sides
    assert len(res["shim"][2]) == len(res["ref"][2]) == n
    assert len(res["shim"][2][0].split()) == len(res["ref"][2][0].split()) == 13
    # and the point of the exercise: the tree force of either build agrees with the direct sums
    assert np.median(g2test.rel_err(res["shim"][1], res["shim"][0])) < 5e-3
