"""GPU parity AT THE BASELINE SIZES against the unmodified reference (oracle/_ref): configs 2-5 of BASELINE.json.

For each: the reference's domain_Decomposition + force_treebuild at full size, the device's keys / order / tree compared bit for bit
(Nodes[], Nextnode[], Father[]), and for a strided sample of targets walked by the reference's own force_treeevaluate[_shortrange]
(relative criterion with the OldAcc of a Barnes-Hut first pass, as in a run): accelerations to median <= 1e-6 / p99.9 <= 3e-4 and GravCost
EXACTLY.  The reference is serial for domain + build: ~2 s at 1 M, ~6 s at 2.1 M, ~50 s at 16.8 M particles."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import g2parity
import g2test

pytestmark = pytest.mark.gpu
PKG = os.path.join(ROOT, "gadget-2.0.7-ngravs_b200")

# workload key of bench.make_workload, sample stride
CASES = [("hernquist1m", 40), ("periodic128", 40), ("periodic256x4", 400), ("periodic256", 400)]


@pytest.mark.parametrize("key,stride", CASES)
def test_fullsize_parity(key, stride, outdir):
    import bench
    from g2gpu import TreeGravity
    from refrun import available
    w = bench.make_workload(key)
    if not available(w["ref_variant"]):
        pytest.skip("oracle/_ref not built")
    n = w["n"]
    tg = TreeGravity(max_part=int(1.1 * n) + 64, n_gravs=w["D"], periodic=w["periodic"], shortrange=w["shortrange"], unequal_softenings=w["unequal"])
    tg.set_species(w["grav"], g2test.force_softening(w["soft"]))
    tg.set_laws()
    asmth, rcut = bench.pm_split(w)
    if w["shortrange"]:
        tab = np.load(os.path.join(PKG, "data", "srtable_newton_ntab2048.npy"))
        tg.set_srtable(np.broadcast_to(tab, (w["D"], w["D"], len(tab))).copy())
    # first force computation of a run: Barnes-Hut, OldAcc = 0 (accel.c:46-49) -> OldAcc of the relative criterion
    acc0, cost0, old0, perm0 = tg.gravity_tree(w["pos"], w["mass"], w["ptype"], tg.walk_params(theta=0.5, errtol=0.005, boxsize=w["box"], G=1.0, asmth=asmth, rcut=rcut))
    oldacc_by_id = np.zeros(n, dtype=np.float32)
    oldacc_by_id[perm0] = old0
    del acc0, cost0, old0
    acc, cost, old, perm = tg.gravity_tree(w["pos"], w["mass"], w["ptype"], tg.walk_params(theta=0.0, errtol=0.005, boxsize=w["box"], G=1.0, asmth=asmth, rcut=rcut),
                                           oldacc=oldacc_by_id)
    rewalked = tg.timings()["rewalked"]
    st = g2parity.reference_state(w, oldacc_by_id, stride, len(os.sched_getaffinity(0)))
    par = g2parity.compare_with_device(tg, st, acc, cost, perm)
    par["rewalked_targets"] = int(rewalked)
    par["reference_s"] = dict(domain=st["t_domain"], build=st["t_build"], walk_sample=st["t_walk_sample"])
    tg.close()
    with open(os.path.join(outdir, f"fullsize_parity_{key}.json"), "w") as f:
        json.dump(par, f, indent=1)
    assert par["order_equal"], par
    assert par["tree_equal"], par
    assert par["median"] <= 1.0e-6 and par["p999"] <= 3.0e-4, par
    assert par["cost_mismatch"] == 0, par
