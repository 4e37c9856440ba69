"""GPU: the multi-GPU group inside libg2gpu.so (g2gpu_group_*, csrc/g2_group.cu) and the rank slices of the single-context walk.

* slices: the union of the nranks = 2 / 8 slices of g2gpu_walk equals the nranks = 1 walk BIT FOR BIT (slice boundaries are multiples
  of 32 targets, so the 32-target groups do not depend on the number of ranks);
* a one-device group (sharded upload path, compact slice download, host scatter) equals the single-context calls bit for bit;
* with >= 2 GPUs: a two-device group (NCCL all-gather, two slices, cost-weighted second step) equals one device bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

import g2test

pytestmark = pytest.mark.gpu
PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gadget-2.0.7-ngravs_b200")


def _ngpu():
    import g2gpu
    return int(g2gpu.load_library().g2gpu_device_count())


def _treepm_case(n_side=40):
    n, box = n_side ** 3, 100000.0
    pos, mass, ptype = g2test.periodic_poisson(n, box)
    eps = box / n_side / 30.0
    tab = np.load(os.path.join(PKG, "data", "srtable_newton_ntab2048.npy"))
    asmth = 1.25 * box / 64
    return dict(n=n, box=box, pos=pos, mass=mass, ptype=ptype, soft=(eps,) * 6, tab=np.broadcast_to(tab, (2, 2, len(tab))).copy(), asmth=asmth, rcut=4.5 * asmth)


def _setup(t, c):
    t.set_species(g2test.GRAV_D2, g2test.force_softening(c["soft"]))
    t.set_laws()
    t.set_srtable(c["tab"])


def _single(c, oldacc=None, active=None, theta=0.5):
    from g2gpu import TreeGravity
    t = TreeGravity(max_part=int(1.1 * c["n"]) + 64, n_gravs=2, periodic=True, shortrange=True, unequal_softenings=False)
    _setup(t, c)
    wp = t.walk_params(theta=theta, errtol=0.005, boxsize=c["box"], G=1.0, asmth=c["asmth"], rcut=c["rcut"])
    out = t.gravity_tree(c["pos"], c["mass"], c["ptype"], wp, oldacc=oldacc, active=active)
    return t, wp, out


@pytest.mark.parametrize("nranks", [2, 8])
def test_walk_slices_equal_one_rank(nranks):
    c = _treepm_case()
    t, wp, (acc1, cost1, old1, perm1) = _single(c)
    # overwrite every result with those of a different opening angle, so that an entry can only match acc1 if some rank recomputed it
    t.walk(t.walk_params(theta=0.9, errtol=0.005, boxsize=c["box"], G=1.0, asmth=c["asmth"], rcut=c["rcut"]))
    a0, c0, _ = t.download_acc()
    assert np.mean(c0 != cost1) > 0.5
    seen = 0
    for r in range(nranks):
        t.set_option("nranks", nranks)
        t.set_option("rank", r)
        t.walk(wp)
        lo, hi = t.slice()
        assert lo % 32 == 0 and (hi % 32 == 0 or r == nranks - 1)
        seen += hi - lo
    acc, cost, _ = t.download_acc()
    t.close()
    assert seen == c["n"]
    assert np.array_equal(acc.view(np.uint32), acc1.view(np.uint32))
    assert np.array_equal(cost, cost1)


def test_group_of_one_equals_single_context():
    from g2gpu import TreeGravityGroup
    c = _treepm_case()
    t, wp, (acc1, cost1, old1, perm1) = _single(c)
    t.close()
    g = TreeGravityGroup(max_part=int(1.1 * c["n"]) + 64, n_gravs=2, periodic=True, shortrange=True, unequal_softenings=False, ndev=1)
    _setup(g, c)
    acc, cost, old, perm = g.gravity_tree(c["pos"], c["mass"], c["ptype"], wp)
    assert np.array_equal(perm, perm1)
    assert np.array_equal(acc.view(np.uint32), acc1.view(np.uint32)) and np.array_equal(cost, cost1) and np.array_equal(old, old1)
    # staged calls + a sparse active set: only active targets are written, and they equal a single context's walk of the same set
    active = (np.random.default_rng(5).random(c["n"]) < 0.2).astype(np.int32)
    t, wp, (acc1s, cost1s, old1s, perm1s) = _single(c, active=active)
    t.close()
    g.upload(c["pos"], c["mass"], c["ptype"], oldacc=None, active=active)
    g.domain()
    g.treebuild()
    g.walk(wp)
    a2 = np.full_like(acc1, -7.0)
    c2 = np.full_like(cost1, -7.0)
    g.download_acc(out=(a2, c2, np.zeros_like(cost1)))
    act_dev = active[perm1] != 0
    assert np.array_equal(a2[act_dev].view(np.uint32), acc1s[act_dev].view(np.uint32)) and np.array_equal(c2[act_dev], cost1s[act_dev])
    assert np.array_equal(c2[act_dev], cost1[act_dev])          # GravCost does not depend on how targets are grouped
    assert np.all(c2[~act_dev] == -7.0)
    lo, hi, fr = g.slices()
    assert lo[0] == 0 and hi[0] == int(active.sum())
    g.close()


def _pinned(n):
    import torch
    return (torch.full((n, 3), -7.0, dtype=torch.float32).pin_memory().numpy(), torch.full((n,), -7.0, dtype=torch.float32).pin_memory().numpy(),
            torch.full((n,), -7.0, dtype=torch.float32).pin_memory().numpy(), torch.zeros(n, dtype=torch.int32).pin_memory().numpy())


def test_zero_copy_results_equal_staged_results():
    """Pinned result arrays are written by the walk kernel itself (g2gpu_group_zero_copy): same bits as the staged download, all targets and a
    sparse active set (only active targets are written); pageable arrays take the staged path."""
    from g2gpu import TreeGravityGroup
    c = _treepm_case()
    t, wp, (acc1, cost1, old1, perm1) = _single(c)
    t.close()
    active = (np.random.default_rng(5).random(c["n"]) < 0.2).astype(np.int32)
    t, wp, (acc1s, cost1s, old1s, perm1s) = _single(c, active=active)
    t.close()
    g = TreeGravityGroup(max_part=int(1.1 * c["n"]) + 64, n_gravs=2, periodic=True, shortrange=True, unequal_softenings=False, ndev=1)
    _setup(g, c)
    g.gravity_tree(c["pos"], c["mass"], c["ptype"], wp)
    assert not g.zero_copy()
    out = _pinned(c["n"])
    acc, cost, old, perm = g.gravity_tree(c["pos"], c["mass"], c["ptype"], wp, out=out)
    assert g.zero_copy()
    assert np.array_equal(perm, perm1)
    assert np.array_equal(acc.view(np.uint32), acc1.view(np.uint32)) and np.array_equal(cost, cost1) and np.array_equal(old, old1)
    h2d, d2h, _ = g.io_bytes()
    assert d2h == 24 * c["n"]                                   # 20 B per target stored by the kernel + 4 B of the particle order
    out = _pinned(c["n"])
    acc, cost, old, perm = g.gravity_tree(c["pos"], c["mass"], c["ptype"], wp, active=active, out=out)
    assert g.zero_copy()
    act_dev = active[perm1] != 0
    assert np.array_equal(acc[act_dev].view(np.uint32), acc1s[act_dev].view(np.uint32)) and np.array_equal(cost[act_dev], cost1s[act_dev])
    assert np.array_equal(old[act_dev], old1s[act_dev])
    assert np.all(cost[~act_dev] == -7.0) and np.all(acc[~act_dev] == -7.0)
    g.close()


def test_group_of_two_equals_one_device():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    from g2gpu import TreeGravityGroup
    c = _treepm_case(48)
    t, wp, (acc1, cost1, old1, perm1) = _single(c)
    # relative criterion with the first pass's OldAcc (ids -> upload order)
    oldacc = np.zeros(c["n"], dtype=np.float32)
    oldacc[perm1] = old1
    t.close()
    t, wp2, (acc2, cost2, old2, perm2) = _single(c, oldacc=oldacc, theta=0.0)
    t.close()
    g = TreeGravityGroup(max_part=int(1.1 * c["n"]) + 64, n_gravs=2, periodic=True, shortrange=True, unequal_softenings=False, ndev=2)
    assert g.ndev == 2
    _setup(g, c)
    acc, cost, old, perm = g.gravity_tree(c["pos"], c["mass"], c["ptype"], wp)
    assert np.array_equal(perm, perm1)
    assert np.array_equal(acc.view(np.uint32), acc1.view(np.uint32)) and np.array_equal(cost, cost1) and np.array_equal(old, old1)
    lo, hi, fr = g.slices()
    assert lo[0] == 0 and hi[0] == lo[1] and hi[1] == c["n"] and lo[1] % 32 == 0
    assert abs(fr[1] - 0.5) < 0.1                      # cost-weighted boundary of the NEXT walk
    h2d, d2h, gathered = g.io_bytes()
    # per target: 20 B of results + 4 B particle index in the compact slices, + 4 B of the particle order (each device returns 1/N of it)
    assert h2d == 20 * c["n"] and d2h == 28 * c["n"] and gathered >= 32 * c["n"]
    # second step (slices now cut at equal GravCost): still bit-identical to one device
    acc, cost, old, perm = g.gravity_tree(c["pos"], c["mass"], c["ptype"], wp2, oldacc=oldacc)
    assert np.array_equal(acc.view(np.uint32), acc2.view(np.uint32)) and np.array_equal(cost, cost2)
    assert not g.zero_copy()
    # pinned result arrays: both devices store their slices straight into them; the GravCost profile of the balancer comes from the devices
    fr_staged = g.slices()[2]
    out = _pinned(c["n"])
    acc, cost, old, perm = g.gravity_tree(c["pos"], c["mass"], c["ptype"], wp2, oldacc=oldacc, out=out)
    assert g.zero_copy()
    assert np.array_equal(acc.view(np.uint32), acc2.view(np.uint32)) and np.array_equal(cost, cost2) and np.array_equal(old, old2)
    assert np.array_equal(perm, perm1)
    fr = g.slices()[2]
    assert abs(fr[1] - fr_staged[1]) < 1e-3                   # same profile, summed in another order
    g.close()
