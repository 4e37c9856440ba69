"""GPU parity of the ngravs pair force laws (ngravs.c:344-861): point-wise against the reference's own function pointers, and
through the walk with the reference's Yukawa (NGRAVS_YUKAWA_FORCETEST) and BAM (NGRAVS_ACCUMULATOR_TESTING) wirings."""
import numpy as np
import pytest

import g2test
from refrun import RefOracle, available

pytestmark = pytest.mark.gpu

YUK = dict(accel=[["none", "yukawa"], ["yukawa", "none"]], spline=[["none", "plummer"], ["plummer", "none"]])
BAM = dict(accel=[["newtonian", "sourcebambaryon"], ["sourcebaryonbam", "bambam"]],
           spline=[["plummer", "sourcebambaryon_spline"], ["sourcebaryonbam_spline", "bambam_spline"]])


def law_params(box):
    par = np.zeros((2, 2, 4))
    par[:, :, 0] = 60.0 / box if box > 0 else 0.0      # YUKAWA_IMASS / All.BoxSize (ngravs.c:41-43, 858)
    par[:, :, 1] = 1.31e-6                             # BAM_EPSILON (ngravs.c:45-47)
    return par


@pytest.mark.parametrize("variant,wiring,box", [("np_d2_f32", dict(accel="newtonian", spline="plummer"), 0.0),
                                                ("np_yuk_f32", YUK, 1000.0), ("np_bam_f32", BAM, 0.0)])
def test_pair_laws_pointwise(variant, wiring, box):
    if not available(variant):
        pytest.skip("oracle/_ref not built")
    from g2gpu import TreeGravity
    ref = RefOracle(variant, 64, boxsize=box)
    tg = TreeGravity(max_part=1024, n_gravs=2)
    tg.set_species(g2test.GRAV_D2, g2test.force_softening(g2test.SOFT_NP))
    tg.set_laws(wiring["accel"], wiring["spline"], law_params(box))
    rng = np.random.default_rng(1)
    n = 400
    bam = variant == "np_bam_f32"
    pm = (rng.uniform(1e-6, 1e-3, n) if bam else rng.uniform(0.1, 10, n)).astype(np.float32)
    m = (rng.uniform(1e-6, 1e-3, n) if bam else rng.uniform(0.1, 10, n)).astype(np.float32)
    r = (10 ** rng.uniform(-2, 2, n)).astype(np.float32)
    h = (10 ** rng.uniform(-1.5, 1, n)).astype(np.float32)
    for i in range(2):
        for j in range(2):
            got = tg.eval_pairs(i, j, pm, m, r, h)
            want = np.array([ref.accel(i, j, float(pm[k]), float(m[k]), float(r[k]) ** 2, float(r[k]), 1) / float(r[k]) if r[k] >= h[k]
                             else ref.spline(i, j, float(pm[k]), float(m[k]), float(h[k]), float(r[k]), 1) for k in range(n)])
            scale = np.maximum(np.abs(want), 1e-30)
            err = np.abs(got - want) / scale
            ok = (np.abs(want) < 1e-30) | (err < 2e-5)
            assert ok.all(), (variant, i, j, float(err.max()))
    tg.close()


@pytest.mark.parametrize("variant,wiring", [("np_yuk_f32", YUK), ("np_bam_f32", BAM), ("np_bamacc_f32", BAM)])
def test_walk_with_non_newtonian_wiring(variant, wiring, outdir):
    if not available(variant):
        pytest.skip("oracle/_ref not built")
    from g2gpu import TreeGravity
    n = 20000
    box = 1000.0 if variant == "np_yuk_f32" else 0.0
    pos, mass, ptype = g2test.gaussian_blobs(n, seed=21)
    if variant.startswith("np_bam"):
        mass = (mass * 1e-3).astype(np.float32)        # BAM scale radii 4 pi eps / m comparable to the separations
    soft, grav = g2test.SOFT_NP, g2test.GRAV_D2
    ref = RefOracle(variant, int(1.1 * n) + 64, boxsize=box, softening=soft, gravity=grav)
    ref.load(pos, mass, ptype)
    ref.domain()
    rp = ref.particles()
    ref.gravity()
    r1 = ref.particles()
    tg = TreeGravity(max_part=ref.maxpart, n_gravs=2)
    tg.set_species(grav, g2test.force_softening(soft))
    tg.set_laws(wiring["accel"], wiring["spline"], law_params(box))
    if variant == "np_bamacc_f32":
        tg.set_option("accumulator", 1)                # the reference built with -DNGRAVS_ACCUMULATOR: laws get N of the node
    tg.upload(rp["pos"], rp["mass"], rp["type"])
    tg.domain()
    assert tg.treebuild() == ref.tree()["numnodes"]
    if variant == "np_bamacc_f32":
        assert np.array_equal(tg.nparticles(), ref.nparticles())       # Nodes[].u.d.Nparticles, bit-exact
    tg.walk(tg.walk_params(theta=0.5, boxsize=box, G=1.0))
    acc, cost, old = tg.download_acc()
    tg.close()
    err = g2test.rel_err(acc, r1["acc"])
    assert int(np.sum(cost != r1["cost"])) == 0
    assert np.median(err) <= 1e-5 and np.percentile(err, 99.9) <= 1e-3, (float(np.median(err)), float(np.percentile(err, 99.9)))
