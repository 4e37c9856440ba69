"""Full-size parity against the UNMODIFIED reference (oracle/_ref), shared by tests/test_gpu_fullsize.py and the cpu_baseline leg of
bench.py (TEST INFRASTRUCTURE: the only places that touch oracle/).

The reference's domain_Decomposition + force_treebuild run once at full size (serial, as in the reference); its own per-target walk
function (force_treeevaluate / force_treeevaluate_shortrange) then walks a strided SAMPLE of the targets on all host cores.  Compared
with the device: particle order (Peano-Hilbert, species-major), every node record and link of the tree (Nodes[], Nextnode[], Father[]:
bit-exact), and for the sampled targets the accelerations (relative error) and GravCost (exactly)."""
import time

import numpy as np

import g2test


def reference_state(w, oldacc_by_id, stride, cores, theta=0.0, errtol=0.005):
    """w: a workload dict of bench.make_workload.  Returns the reference after domain + build + sampled walk."""
    from refrun import RefOracle, available
    if not available(w["ref_variant"]):
        raise RuntimeError("oracle/_ref is not built")
    n = w["n"]
    ref = RefOracle(w["ref_variant"], int(1.1 * n) + 64, boxsize=w["box"], softening=w["soft"], gravity=w["grav"], theta=theta, errtol=errtol,
                    criterion=1, buffer_mb=64)
    ref.load(w["pos"], w["mass"], w["ptype"])
    t0 = time.time()
    ref.domain()                                   # extent + keys + qsort + top tree + peano_hilbert_order
    t_domain = time.time() - t0
    rp = ref.particles()
    t0 = time.time()
    ref.treebuild()
    t_build = time.time() - t0
    active = np.zeros(n, dtype=np.int32)
    active[::stride] = 1
    ref.set_active(active)
    if oldacc_by_id is not None:
        ref.set_oldacc(oldacc_by_id[rp["id"]])
    ref.set_opening(theta, errtol, 1)
    t_walk, cost = ref.walk_threads(cores)
    return dict(ref=ref, ids=rp["id"].astype(np.int64), active=active, t_domain=t_domain, t_build=t_build, t_walk_sample=t_walk,
                cost_sum=cost, nsample=int(active.sum()))


def compare_with_device(tg, st, acc_dev, cost_dev, perm_dev, check_tree=True):
    """tg: the device context holding the tree of the same particle set; acc_dev/cost_dev: its walk results in device order; perm_dev:
    device index -> input index.  st: reference_state().  Returns the parity record."""
    ref = st["ref"]
    n = len(perm_dev)
    out = dict(particles=n, sample=st["nsample"])
    out["order_equal"] = bool(np.array_equal(np.asarray(perm_dev, dtype=np.int64), st["ids"]))
    if check_tree:
        rt = ref.tree()
        gt = tg.tree()
        mism = g2test.compare_tree(gt, rt, ref.D)
        out["numnodes"] = int(rt["numnodes"])
        out["tree_mismatches"] = {k: int(v) for k, v in mism.items() if v}
        out["tree_equal"] = bool(all(v == 0 for v in mism.values()))
        del rt, gt
    rp = ref.particles()
    sel = np.nonzero(st["active"])[0]                 # reference order
    by_id_acc = np.zeros((n, 3), dtype=np.float32)
    by_id_cost = np.zeros(n, dtype=np.float32)
    by_id_acc[perm_dev] = acc_dev
    by_id_cost[perm_dev] = cost_dev
    ids = st["ids"][sel]
    err = g2test.rel_err(by_id_acc[ids], rp["acc"][sel])
    out["median"] = float(np.median(err))
    out["p999"] = float(np.percentile(err, 99.9))
    out["max"] = float(err.max())
    out["cost_mismatch"] = int(np.sum(by_id_cost[ids] != rp["cost"][sel]))
    out["ia_per_part_ref"] = float(rp["cost"][sel].mean())
    out["ia_per_part_gpu"] = float(by_id_cost[ids].mean())
    return out
