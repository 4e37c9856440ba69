"""GPU parity of the lattice-sum correction walk (periodic box WITHOUT PM): g2gpu_walk with lattice tables against
  * the UNMODIFIED reference built with -DPERIODIC and no PMGRID (oracle/_ref variant per_d2_f32, where it travelled): accelerations to
    the force walk's tolerance (median 1e-5, 99.9th percentile 1e-3), GravCost (tree + lattice interactions) equal for all but
    borderline particles -- with the reference's own tables, whose pairs other than [0][0] are half empty in a float build
    (oracle/refrun.py), so the [target][source] table indexing is pinned too;
  * the device-made Ewald table against the reference's [0][0] table (lattice_init + ewald_force, ngravs.c:1170);
  * the pinned oracle port with the device-made table at a larger size;
  * exact Ewald sums (numpy, FP64): the tree force WITH the correction is the periodic force to tree accuracy, without it it is not."""
import numpy as np
import pytest

import g2test
from portrun import PortOracle
from refrun import RefOracle, available

pytestmark = pytest.mark.gpu


def gpu_periodic(maxpart, grav, soft):
    from g2gpu import TreeGravity
    t = TreeGravity(max_part=maxpart, n_gravs=2, periodic=True, shortrange=False, unequal_softenings=False)
    t.set_species(grav, g2test.force_softening(soft))
    t.set_laws()
    return t


@pytest.mark.skipif(not available("per_d2_f32"), reason="oracle/_ref not built")
def test_lattice_correction_matches_reference_build(outdir):
    n, box = 20000, 1000.0
    pos, mass, ptype = g2test.periodic_poisson(n, box, seed=31)
    mass = (mass * np.random.default_rng(4).uniform(0.5, 2.0, n)).astype(np.float32)
    soft, grav = (box / 27 / 30.0,) * 6, g2test.GRAV_D2
    ref = RefOracle("per_d2_f32", int(1.1 * n) + 64, boxsize=box, softening=soft, gravity=grav)
    tabs = ref.lattice_tables()
    ref.load(pos, mass, ptype)
    ref.domain()
    rp = ref.particles()
    t = gpu_periodic(ref.maxpart, grav, soft)
    # the device-made Ewald table is the reference's complete ([0][0]) table
    ew = t.make_ewald_table(64) / (box * box)
    scale = np.abs(tabs[:, 0, 0]).max()
    assert np.abs(ew - tabs[:, 0, 0]).max() <= 1e-9 * scale
    t.set_lattice_tables(tabs)
    t.upload(rp["pos"], rp["mass"], rp["type"])
    t.domain()
    assert np.array_equal(t.order(), np.arange(n, dtype=np.int32))
    t.treebuild()
    ref.gravity()
    r1 = ref.particles()
    t.walk(t.walk_params(theta=0.5, errtol=0.005, boxsize=box))
    acc, cost, old = t.download_acc()
    e1 = g2test.rel_err(acc, r1["acc"])
    assert np.median(e1) <= 1e-5 and np.percentile(e1, 99.9) <= 1e-3, (np.median(e1), np.percentile(e1, 99.9))
    assert (cost != r1["cost"]).sum() <= 0.002 * n
    assert np.median(np.abs(old - r1["oldacc"]) / r1["oldacc"]) <= 1e-5
    # relative criterion, OldAcc of the first pass
    ref.set_opening(0.0, 0.005, 1)
    ref.force_rebuild()
    ref.gravity()
    r2 = ref.particles()
    t.upload(rp["pos"], rp["mass"], rp["type"], oldacc=r1["oldacc"])
    t.domain()
    t.treebuild()
    t.walk(t.walk_params(theta=0.0, errtol=0.005, boxsize=box))
    acc2, cost2, _ = t.download_acc()
    e2 = g2test.rel_err(acc2, r2["acc"])
    assert np.median(e2) <= 1e-5 and np.percentile(e2, 99.9) <= 1e-3, (np.median(e2), np.percentile(e2, 99.9))
    assert (cost2 != r2["cost"]).sum() <= 0.002 * n
    # removing the tables gives the nearest-image force again
    t.set_lattice_tables(None)
    t.walk(t.walk_params(theta=0.0, errtol=0.005, boxsize=box))
    acc3, cost3, _ = t.download_acc()
    assert (cost3 < cost2).all()
    with open(f"{outdir}/lattice_reference.txt", "w") as f:
        f.write(f"bh median {np.median(e1):.3e} p99.9 {np.percentile(e1, 99.9):.3e} cost_mismatch {int((cost != r1['cost']).sum())}\n")
        f.write(f"rel median {np.median(e2):.3e} p99.9 {np.percentile(e2, 99.9):.3e} cost_mismatch {int((cost2 != r2['cost']).sum())}\n")


def test_lattice_correction_against_port_and_exact_ewald(outdir):
    """100 k particles, device-made tables for all pairs (the stock wiring): (1) parity with the pinned port; (2) against exact Ewald sums the
    corrected tree force is accurate to the opening criterion, the uncorrected one is off by the lattice term."""
    n, box = 100000, 1000.0
    pos, mass, ptype = g2test.periodic_poisson(n, box, seed=37)
    soft, grav = (box / 46 / 30.0,) * 6, g2test.GRAV_D2
    t = gpu_periodic(int(1.1 * n) + 64, grav, soft)
    ew = t.set_ewald_lattice(box)
    fc = np.broadcast_to(ew[:, None, None], (3, 2, 2, 65, 65, 65)).copy()
    o = PortOracle(int(1.1 * n) + 64, D=2, periodic=True, shortrange=False, unequal=False, boxsize=box, softening=soft, gravity=grav)
    o.set_lattice_tables(fc)
    o.load(pos, mass, ptype)
    o.domain()
    o.gravity(nthreads=8)
    p = o.particles()
    t.upload(p["pos"], p["mass"], p["type"])
    t.domain()
    assert np.array_equal(t.order(), np.arange(n, dtype=np.int32))
    t.treebuild()
    wp = t.walk_params(theta=0.5, errtol=0.005, boxsize=box)
    t.walk(wp)
    acc, cost, _ = t.download_acc()
    e = g2test.rel_err(acc, p["acc"])
    assert np.median(e) <= 1e-5 and np.percentile(e, 99.9) <= 1e-3, (np.median(e), np.percentile(e, 99.9))
    assert (cost != p["cost"]).sum() <= 0.002 * n
    tg = np.arange(0, n, n // 16)[:16]
    exact = g2test.ewald_direct(p["pos"], p["mass"], tg, box)
    wp = t.walk_params(theta=0.3, errtol=0.005, boxsize=box)
    t.walk(wp)
    acc_fine, _, _ = t.download_acc()
    err_with = g2test.rel_err(acc_fine[tg], exact)
    t.set_lattice_tables(None)
    t.walk(wp)
    acc0, _, _ = t.download_acc()
    err_without = g2test.rel_err(acc0[tg], exact)
    with open(f"{outdir}/lattice_vs_ewald.txt", "w") as f:
        f.write(f"parity vs port: median {np.median(e):.3e} p99.9 {np.percentile(e, 99.9):.3e}; vs exact Ewald: with correction median "
                f"{np.median(err_with):.3e}, nearest image only {np.median(err_without):.3e}\n")
    assert np.median(err_with) < 5e-2 and np.median(err_without) > 2 * np.median(err_with)


def test_lattice_correction_with_sparse_active_targets():
    """Only particles with Ti_endstep == Ti_Current are walked (gravtree.c:113): a random 10 % active set gets exactly the values of
    the all-active run (tree + lattice correction, GravCost), the others are left untouched; an empty active set is a no-op."""
    n, box = 30000, 1000.0
    pos, mass, ptype = g2test.periodic_poisson(n, box, seed=41)
    soft, grav = (box / 31 / 30.0,) * 6, g2test.GRAV_D2
    t = gpu_periodic(n + 64, grav, soft)
    t.set_ewald_lattice(box)
    wp = t.walk_params(theta=0.5, errtol=0.005, boxsize=box)
    t.upload(pos, mass, ptype)
    t.domain()
    order = t.order()
    t.treebuild()
    t.walk(wp)
    acc_all, cost_all, _ = t.download_acc()
    active = (np.random.default_rng(1).random(n) < 0.1).astype(np.int32)
    t2 = gpu_periodic(n + 64, grav, soft)
    t2.set_ewald_lattice(box)
    t2.upload(pos, mass, ptype, active=active)
    t2.domain()
    assert np.array_equal(t2.order(), order)
    t2.treebuild()
    t2.walk(wp)
    acc, cost, _ = t2.download_acc()
    a = active[order] != 0
    # same per-target interaction lists (GravCost equal); the FP32 partial sums of a target are flushed where ITS warp descends, and the
    # 32 targets that share a warp differ between the two runs, so accelerations agree to rounding, not bit for bit
    assert np.array_equal(cost[a], cost_all[a])
    e = g2test.rel_err(acc[a], acc_all[a])
    assert np.median(e) <= 1e-6 and e.max() <= 1e-4, (np.median(e), e.max())     # measured: median 2e-7, max 1.2e-5
    assert np.all(acc[~a] == 0) and np.all(cost[~a] == 0)
    t2.upload(pos, mass, ptype, active=np.zeros(n, dtype=np.int32))
    t2.domain()
    t2.treebuild()
    t2.walk(wp)
    assert t2.timings()["interactions"] == 0


def test_lattice_correction_matches_golden_fixture():
    """The committed fixture of the unmodified PERIODIC-without-PMGRID reference (tests/golden/make_golden_lattice.py; single species, so
    only its complete [0][0] table is involved) against the device walk with the device-made Ewald table: needs no reference build."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lattice_per_d2_poisson3000.npz"))
    box, n = float(g["box"]), len(g["mass"])
    t = gpu_periodic(int(g["maxpart"]), g["grav"], g["soft"])
    ew = t.set_ewald_lattice(box)
    idx = g["table_index"]
    assert np.abs(ew[:, idx[:, 0], idx[:, 1], idx[:, 2]] - g["table_sample"]).max() <= 1e-9 * np.abs(g["table_sample"]).max()
    t.upload(g["pos"], g["mass"], g["type"])
    t.domain()
    assert np.array_equal(t.order(), np.arange(n, dtype=np.int32))
    t.treebuild()
    t.walk(t.walk_params(theta=0.5, errtol=0.005, boxsize=box))
    acc, cost, old = t.download_acc()
    e = g2test.rel_err(acc, g["bh_acc"])
    assert np.median(e) <= 1e-5 and np.percentile(e, 99.9) <= 1e-3, (np.median(e), e.max())
    assert (cost != g["bh_cost"]).sum() <= max(1, 0.002 * n)
    t.upload(g["pos"], g["mass"], g["type"], oldacc=g["bh_oldacc"])
    t.domain()
    t.treebuild()
    t.walk(t.walk_params(theta=0.0, errtol=0.005, boxsize=box))
    acc2, cost2, _ = t.download_acc()
    e2 = g2test.rel_err(acc2, g["rel_acc"])
    assert np.median(e2) <= 1e-5 and np.percentile(e2, 99.9) <= 1e-3, (np.median(e2), e2.max())
    assert (cost2 != g["rel_cost"]).sum() <= max(1, 0.002 * n)
