"""GPU: the two walk kernels (walk_mode 0 = one cursor per 32 targets, g2_walk.cu; walk_mode 1 = one warp per target over
the level-order records, g2_walkb.cu) take the same per-target decisions: GravCost identical for every particle, the
accelerations equal up to the order of the FP32 summation.  The parity of either mode against the reference is covered by
running the whole -m gpu suite with G2GPU_WALK_MODE=0/1 (the library default is what test_gpu_tree_walk.py checks)."""
import os

import numpy as np
import pytest

import g2test

pytestmark = pytest.mark.gpu

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gadget-2.0.7-ngravs_b200")


def _run(mode, pos, mass, ptype, D, grav, soft, periodic, sr, unequal, box, pmgrid, theta, oldacc=None, acc_double=1):
    from g2gpu import TreeGravity
    n = len(mass)
    tg = TreeGravity(max_part=max(int(1.1 * n) + 64, 2000), n_gravs=D, periodic=periodic, shortrange=sr, unequal_softenings=unequal)
    tg.set_option("walk_mode", mode)
    tg.set_option("acc_double", acc_double)
    tg.set_species(grav, g2test.force_softening(soft))
    tg.set_laws()
    asmth = 1.25 * box / pmgrid if sr else 0.0
    if sr:
        tab = np.load(os.path.join(PKG, "data", "srtable_newton_ntab2048.npy"))
        tg.set_srtable(np.broadcast_to(tab, (D, D, len(tab))).copy())
    wp = tg.walk_params(theta=theta, errtol=0.005, boxsize=box, G=1.0, asmth=asmth, rcut=4.5 * asmth)
    acc, cost, old, perm = tg.gravity_tree(pos, mass, ptype, wp, oldacc=oldacc)
    t = tg.timings()
    tg.close()
    out = np.zeros_like(acc), np.zeros_like(cost), np.zeros_like(old)
    out[0][perm] = acc
    out[1][perm] = cost
    out[2][perm] = old
    return out + (t,)


CASES = {
    "blobs_d2_unequal": dict(gen=lambda: g2test.gaussian_blobs(30000), D=2, grav=g2test.GRAV_D2, soft=g2test.SOFT_NP, periodic=False, sr=False,
                             unequal=True, box=0.0, pmgrid=0),
    "hernquist_d2": dict(gen=lambda: g2test.hernquist(200000), D=2, grav=g2test.GRAV_D2, soft=(0.0, 0.05, 0.05, 0.05, 0.05, 0.05), periodic=False,
                         sr=False, unequal=True, box=0.0, pmgrid=0),
    "treepm64_d2": dict(gen=lambda: g2test.periodic_poisson(32768), D=2, grav=g2test.GRAV_D2, soft=(100.0,) * 6, periodic=True, sr=True,
                        unequal=False, box=100000.0, pmgrid=64),
    "treepm256_d4": dict(gen=lambda: g2test.periodic_poisson(262144, ntypes=6), D=4, grav=(0, 1, 2, 3, 1, 2), soft=(50.0, 40.0, 30.0, 50.0, 20.0, 10.0),
                         periodic=True, sr=True, unequal=True, box=100000.0, pmgrid=128),
    "single_particle": dict(gen=lambda: (np.zeros((1, 3), np.float32) + 5, np.ones(1, np.float32), np.ones(1, np.int32)), D=2, grav=g2test.GRAV_D2,
                            soft=g2test.SOFT_NP, periodic=False, sr=False, unequal=True, box=0.0, pmgrid=0),
    "thirty_three": dict(gen=lambda: g2test.gaussian_blobs(33), D=2, grav=g2test.GRAV_D2, soft=g2test.SOFT_NP, periodic=False, sr=False,
                         unequal=True, box=0.0, pmgrid=0),
}


@pytest.mark.parametrize("case", list(CASES))
def test_walk_modes_agree(case, outdir):
    c = dict(CASES[case])
    pos, mass, ptype = c.pop("gen")()
    res = {}
    for mode in (0, 1):
        a1, c1, o1, _ = _run(mode, pos, mass, ptype, theta=0.5, **c)             # Barnes-Hut first pass
        a2, c2, o2, t2 = _run(mode, pos, mass, ptype, theta=0.0, oldacc=o1, **c)  # relative criterion
        res[mode] = (a1, c1, a2, c2, t2)
    n = len(mass)
    for k, name in ((1, "Barnes-Hut"), (3, "relative")):
        mism = int(np.sum(res[0][k] != res[1][k]))
        # the two kernels are compiled separately (different FMA contraction of the same expressions) and OldAcc of pass 1 differs in the
        # last bits (summation order): a borderline decision may flip for a few particles in a million, as it does against the reference
        assert mism <= 2 + n // 50000, (name, mism)
    for k in (0, 2):
        if n > 1:
            err = g2test.rel_err(res[1][k], res[0][k])
            assert np.median(err) <= 2e-6 and np.percentile(err, 99.9) <= 1e-4, (np.median(err), err.max())
    print(f"{case}: walk kernel ms mode0={res[0][4]['walk_kernel_ms']:.3f} mode1={res[1][4]['walk_kernel_ms']:.3f} "
          f"ia/part={res[1][3].mean():.1f} steps={res[1][4]['cell_visits']}")


def test_mode1_reproducible_bits():
    pos, mass, ptype = g2test.periodic_poisson(32768)
    c = CASES["treepm64_d2"]
    kw = {k: v for k, v in c.items() if k != "gen"}
    a = _run(1, pos, mass, ptype, theta=0.5, **kw)
    b = _run(1, pos, mass, ptype, theta=0.5, **kw)
    assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.array_equal(a[1], b[1])
