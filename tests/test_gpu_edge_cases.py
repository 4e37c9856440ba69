"""GPU: edge cases of the hot path (tiny and ragged inputs, sparse active sets, error paths) and size-independent properties
at a BASELINE size (1 M particles, config 2)."""
import os

import numpy as np
import pytest

import g2test
from refrun import RefOracle, available

pytestmark = pytest.mark.gpu


def make_tg(maxpart, **kw):
    from g2gpu import TreeGravity
    tg = TreeGravity(max_part=maxpart, n_gravs=2, **kw)
    tg.set_species(g2test.GRAV_D2, g2test.force_softening(g2test.SOFT_NP))
    tg.set_laws()
    return tg


@pytest.mark.parametrize("n", [2, 3, 31, 33, 257])
def test_tiny_and_ragged_particle_counts(n):
    """fewer targets than one warp / not a multiple of 32 / a handful of tree nodes: bit-exact tree, forces within tolerance"""
    if not available("np_d2_f32"):
        pytest.skip("oracle/_ref not built")
    pos, mass, ptype = g2test.gaussian_blobs(n, seed=100 + n)
    ref = RefOracle("np_d2_f32", 1024, softening=g2test.SOFT_NP, gravity=g2test.GRAV_D2)
    ref.load(pos, mass, ptype)
    ref.domain()
    rp = ref.particles()
    ref.gravity()
    r1 = ref.particles()
    tg = make_tg(1024)
    tg.upload(rp["pos"], rp["mass"], rp["type"])
    tg.domain()
    assert tg.treebuild() == ref.tree()["numnodes"]
    mism = g2test.compare_tree(tg.tree(), ref.tree(), 2)
    assert all(v == 0 for v in mism.values()), mism
    tg.walk(tg.walk_params(theta=0.5, G=1.0))
    acc, cost, old = tg.download_acc()
    tg.close()
    assert np.array_equal(cost, r1["cost"])
    err = g2test.rel_err(acc, r1["acc"])
    assert np.median(err) <= 1e-5 and err.max() <= 1e-3


def test_sparse_active_set_only_updates_active_particles():
    """individual timesteps: only particles with Ti_endstep == Ti_Current are walked (gravtree.c:113)"""
    if not available("np_d2_f32"):
        pytest.skip("oracle/_ref not built")
    n = 20000
    pos, mass, ptype = g2test.hernquist(n, seed=8)
    ref = RefOracle("np_d2_f32", int(1.1 * n) + 64, softening=g2test.SOFT_NP, gravity=g2test.GRAV_D2)
    ref.load(pos, mass, ptype)
    ref.domain()
    rp = ref.particles()
    rng = np.random.default_rng(4)
    active = (rng.uniform(size=n) < 0.25).astype(np.int32)
    ref.set_active(active)
    ref.gravity()
    r1 = ref.particles()
    tg = make_tg(ref.maxpart)
    tg.upload(rp["pos"], rp["mass"], rp["type"], active=active)
    tg.domain()
    tg.treebuild()
    tg.walk(tg.walk_params(theta=0.5, G=1.0))
    acc, cost, old = tg.download_acc()
    t = tg.timings()
    tg.close()
    a = active.astype(bool)
    assert t["interactions"] == int(r1["cost"][a].sum())
    assert np.array_equal(cost[a], r1["cost"][a])
    assert np.all(acc[~a] == 0) and np.all(cost[~a] == 0)          # untouched
    err = g2test.rel_err(acc[a], r1["acc"][a])
    assert np.median(err) <= 1e-5 and np.percentile(err, 99.9) <= 1e-3


def test_error_paths_are_reported_not_papered_over():
    from g2gpu import G2Error, TreeGravity
    n = 5000
    pos, mass, ptype = g2test.gaussian_blobs(n, seed=2)
    # nine coincident particles do not fit the 8 child slots of the deepest node (the reference with NOTREERND runs out of nodes)
    tg = make_tg(8192)
    p2 = pos.copy()
    p2[10:19] = p2[10]
    tg.upload(p2, mass, ptype)
    tg.domain()
    with pytest.raises(G2Error) as e:
        tg.treebuild()
    assert e.value.code == -5
    tg.close()
    # MaxNodes too small: forcetree.c:249-255 -> endrun(1)
    tg = TreeGravity(max_part=8192, n_gravs=2, tree_alloc_factor=0.02)
    tg.set_species(g2test.GRAV_D2, g2test.force_softening(g2test.SOFT_NP))
    tg.set_laws()
    tg.upload(pos, mass, ptype)
    tg.domain()
    with pytest.raises(G2Error) as e:
        tg.treebuild()
    assert e.value.code == -4
    tg.close()
    # particle type outside 0..5
    tg = make_tg(8192)
    bad = ptype.copy()
    bad[7] = 9
    tg.upload(pos, mass, bad)
    tg.domain()
    with pytest.raises(G2Error) as e:
        tg.treebuild()
    assert e.value.code == -2
    # calls out of order / unwired laws
    with pytest.raises(G2Error):
        tg.walk(tg.walk_params())
    tg.close()
    tg = TreeGravity(max_part=8192, n_gravs=2)
    tg.set_species(g2test.GRAV_D2, g2test.force_softening(g2test.SOFT_NP))
    tg.upload(pos, mass, ptype)
    tg.domain()
    tg.treebuild()
    with pytest.raises(G2Error) as e:
        tg.walk(tg.walk_params())
    assert e.value.code == -7                                        # ngravs_core.c:326-365: every slot must be wired
    with pytest.raises(G2Error):
        tg.set_species((0, 0, 5, 0, 0, 0), g2test.force_softening(g2test.SOFT_NP))   # ngravs_core.c:270-279
    tg.close()


def test_full_size_properties_one_million_particles():
    """BASELINE config 2 size (1 M two-species Hernquist): properties that need no oracle."""
    n = 1_000_000
    pos, mass, ptype = g2test.hernquist(n)
    eps = 0.05
    soft = (0.0, eps, eps, eps, eps, eps)
    from g2gpu import TreeGravity
    tg = TreeGravity(max_part=int(1.1 * n) + 64, n_gravs=2)
    tg.set_species(g2test.GRAV_D2, g2test.force_softening(soft))
    tg.set_laws()
    tg.upload(pos, mass, ptype)
    tg.domain()
    keys, perm = tg.keys(), tg.order()
    # a permutation, species-major, Peano-Hilbert sorted inside each species block
    assert np.array_equal(np.sort(perm), np.arange(n))
    species = np.asarray(g2test.GRAV_D2)[ptype[perm]]
    assert np.all(np.diff(species) >= 0)
    for g in (0, 1):
        assert np.all(np.diff(keys[species == g]) >= 0)
    nn = tg.treebuild()
    t = tg.tree()
    # tree invariants: fathers precede children (insertion order), root holds all mass, per-species masses add up
    assert np.all(t["father"][1:] < tg.max_part + np.arange(1, nn)) and t["father"][0] == -1
    msum = np.array([mass[(np.asarray(g2test.GRAV_D2)[ptype] == g)].astype(np.float64).sum() for g in (0, 1)])
    assert np.allclose(t["mass"][0], msum, rtol=1e-5)
    assert np.all(t["len"] > 0) and np.isfinite(t["s"]).all()
    # every particle is the leaf of exactly one node and is reachable: Nextnode is a permutation-like successor map
    assert t["p_father"].min() >= tg.max_part and t["p_father"].max() < tg.max_part + nn
    wp = tg.walk_params(theta=0.5, G=1.0)
    tg.walk(wp)
    acc1, cost1, old1 = tg.download_acc()
    tm = tg.timings()
    assert tm["interactions"] == int(cost1.astype(np.int64).sum())
    assert np.isfinite(acc1).all() and cost1.min() >= 1
    # momentum: sum m a vanishes up to the tree's force error
    m = mass[perm].astype(np.float64)
    resid = np.linalg.norm((m[:, None] * acc1).sum(0)) / (m * np.linalg.norm(acc1, axis=1)).sum()
    assert resid < 5e-3
    # accuracy at full size against FP64 direct summation (gravtree_forcetest-style, BASELINE config 2): second pass with the
    # relative criterion (alpha = 0.005, OldAcc from the Barnes-Hut pass), 256 random targets
    oldacc_by_id = np.zeros(n, dtype=np.float32)
    oldacc_by_id[perm] = old1
    tg.upload(pos, mass, ptype, oldacc=oldacc_by_id)
    tg.domain()
    tg.treebuild()
    tg.walk(tg.walk_params(theta=0.0, errtol=0.005, G=1.0))
    acc_rel, cost_rel, _ = tg.download_acc()
    sample = np.random.default_rng(5).choice(n, 256, replace=False)
    hsoft = g2test.force_softening(soft)[ptype]
    direct = g2test.direct_sum(pos, mass, hsoft, sample)
    inv = np.empty(n, dtype=np.int64)
    inv[perm] = np.arange(n)
    err = g2test.rel_err(acc_rel[inv[sample]], direct)
    assert np.median(err) < 2e-3 and np.percentile(err, 99) < 3e-2, (np.median(err), np.percentile(err, 99))
    # idempotence / determinism: the same step again gives the same bits
    tg.upload(pos, mass, ptype)
    tg.domain()
    tg.treebuild()
    tg.walk(wp)
    acc1, cost1, old1 = tg.download_acc()
    tg.upload(pos, mass, ptype)
    tg.domain()
    assert tg.treebuild() == nn
    tg.walk(wp)
    acc2, cost2, old2 = tg.download_acc()
    assert np.array_equal(cost1, cost2) and np.array_equal(acc1, acc2) and np.array_equal(old1, old2)
    # feeding the sorted output back is a fixed point of peano_hilbert_order
    tg.upload(pos[perm], mass[perm], ptype[perm])
    tg.domain()
    assert np.array_equal(tg.order(), np.arange(n, dtype=np.int32))
    tg.close()


def test_particles_closer_than_the_deepest_level_share_a_bucket_node():
    """Two/three particles that cannot be separated within 21 octree levels (and exact duplicates) must not stop a run: they become
    direct children of one node of the deepest level.  The reference's own tree is not geometric at that scale either
    (forcetree.c:208-232 picks random subnodes below 1e-3 ForceSoftening), so the check is physical: forces against FP64 direct sums."""
    n = 4000
    pos, mass, ptype = g2test.gaussian_blobs(n, seed=4)
    pos = pos.copy()
    pos[101] = pos[100] + np.float32(1e-6)             # closer than L / 2^21
    pos[201] = pos[200]                                  # exact duplicates (a triple)
    pos[202] = pos[200]
    pos[301] = pos[300] * np.float32(1.0000002)
    tg = make_tg(8192)
    tg.upload(pos, mass, ptype)
    tg.domain()
    perm = tg.order()
    nn = tg.treebuild()
    t = tg.tree()
    assert t["father"][0] == -1 and np.all(t["father"][1:] < tg.max_part + np.arange(1, nn))
    inv = np.empty(n, dtype=np.int64)
    inv[perm] = np.arange(n)
    # the duplicates hang under the same node
    pf = t["p_father"]
    assert pf[inv[200]] == pf[inv[201]] == pf[inv[202]] and pf[inv[100]] == pf[inv[101]]
    tg.walk(tg.walk_params(theta=0.3, G=1.0))
    acc, cost, _ = tg.download_acc()
    tg.close()
    targets = np.array([100, 101, 200, 201, 202, 300, 301, 5, 1777, 3999])
    hsoft = g2test.force_softening(g2test.SOFT_NP)[ptype]
    direct = g2test.direct_sum(pos, mass, hsoft, targets)
    err = g2test.rel_err(acc[inv[targets]], direct)
    assert err.max() < 2e-2, err
    assert np.isfinite(acc).all()
    assert int(cost.sum()) == int(cost.astype(np.int64).sum()) and cost.min() >= 1


@pytest.mark.parametrize("side", [128, 256])
def test_full_size_treepm_against_direct_shortrange_sum(side, outdir):
    """BASELINE configs 3 and 5 at full size (2.1 M and 16.8 M particles, periodic TreePM): the short-range tree force of random
    targets against the FP64 direct short-range sum over ALL particles (g2gpu_direct, same table and cut-off) -- what is left is
    the tree approximation, bounded by the relative opening criterion (alpha = 0.005)."""
    import json
    from g2gpu import TreeGravity
    n = side ** 3
    box = 100000.0
    pos, mass, ptype = g2test.periodic_poisson(n, box)
    eps = box / side / 30.0
    tg = TreeGravity(max_part=int(1.1 * n) + 64, n_gravs=2, periodic=True, shortrange=True, unequal_softenings=False)
    tg.set_species(g2test.GRAV_D2, g2test.force_softening((eps,) * 6))
    tg.set_laws()
    tab = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gadget-2.0.7-ngravs_b200", "data", "srtable_newton_ntab2048.npy"))
    tg.set_srtable(np.broadcast_to(tab, (2, 2, len(tab))).copy())
    asmth = 1.25 * box / 256
    rcut = 4.5 * asmth
    wp_bh = tg.walk_params(theta=0.5, errtol=0.005, boxsize=box, G=1.0, asmth=asmth, rcut=rcut)
    wp_rel = tg.walk_params(theta=0.0, errtol=0.005, boxsize=box, G=1.0, asmth=asmth, rcut=rcut)
    acc0, cost0, old0, perm = tg.gravity_tree(pos, mass, ptype, wp_bh)
    oldacc_by_id = np.zeros(n, dtype=np.float32)
    oldacc_by_id[perm] = old0
    acc, cost, old, perm2 = tg.gravity_tree(pos, mass, ptype, wp_rel, oldacc=oldacc_by_id)
    assert np.array_equal(perm, perm2)
    targets = np.sort(np.random.default_rng(11).choice(n, 512, replace=False)).astype(np.int32)
    direct = tg.direct(wp_rel, targets)
    tm = tg.timings()
    tg.close()
    err = g2test.rel_err(acc[targets], direct)
    json.dump(dict(n=n, median=float(np.median(err)), p99=float(np.percentile(err, 99)), max=float(err.max()),
                   ia_per_particle=tm["interactions"] / n), open(os.path.join(outdir, f"fullsize_treepm_{side}.json"), "w"))
    assert np.isfinite(acc).all() and cost.min() >= 1
    assert np.median(err) < 5e-3 and np.percentile(err, 99) < 5e-2, (float(np.median(err)), float(np.percentile(err, 99)))


def test_update_tree_reattaches_the_last_tree():
    """g2gpu_update_tree (dynamic tree update): with unchanged inputs the walk repeats bit for bit; with every particle and
    every node centre of mass shifted by the same vector (what a uniform drift does to the reference's tree) the Barnes-Hut walk, which
    only looks at distances and node sizes, gives the same interaction lists and forces."""
    from g2gpu import TreeGravity, G2Error
    n = 30000
    pos, mass, ptype = g2test.gaussian_blobs(n, seed=9)
    tg = TreeGravity(max_part=int(1.1 * n) + 64, n_gravs=2)
    tg.set_species(g2test.GRAV_D2, g2test.force_softening(g2test.SOFT_NP))
    tg.set_laws()
    wp = tg.walk_params(theta=0.5, errtol=0.005, G=1.0)
    with pytest.raises(G2Error):
        tg.update_tree(np.zeros(0), np.zeros(0))            # no tree yet
    tg.upload(pos, mass, ptype)
    tg.domain()
    tg.treebuild()
    tg.walk(wp)
    acc0, cost0, _ = tg.download_acc()
    t = tg.tree()
    tg.upload(pos, mass, ptype)
    tg.update_tree(t["len"], t["s"])
    tg.walk(wp)
    acc1, cost1, _ = tg.download_acc()
    assert np.array_equal(acc0.view(np.uint32), acc1.view(np.uint32)) and np.array_equal(cost0, cost1)
    shift = np.array([4.0, -2.0, 1.0], dtype=np.float32)
    tg.upload(pos + shift, mass, ptype)
    tg.update_tree(t["len"], t["s"] + shift[None, :, None])
    tg.walk(wp)
    acc2, cost2, _ = tg.download_acc()
    tg.close()
    assert np.sum(cost2 != cost0) <= 0.002 * n
    err = g2test.rel_err(acc2, acc0)
    assert np.median(err) < 1e-5 and np.percentile(err, 99.9) < 1e-3
