#!/usr/bin/env python
"""bench.py — tree-force throughput of the B200 path (and of the reference's CPU path beside it).

    python bench.py --gpus N --steps K --warmup W [--workload hernquist1m|periodic128|periodic256|periodic256x4] [--active-frac F]
                    [--impl reference]

A "step" is one pass of the hot path over the particle set: domain_findExtent + Peano-Hilbert keys + radix sort +
top-level tree + species-major order, force_treebuild (tree + per-species moments) and the tree walk of every active
particle with the relative opening criterion (OldAcc from an untimed Barnes-Hut first pass, as in a run: accel.c:46-49,
gravtree.c:334-335).  `value` = particles updated per second with the inputs resident in HBM; `e2e` = the same through
g2gpu_gravity_tree() with pinned HOST buffers (H2D of positions/masses/types/OldAcc and D2H of accelerations/GravCost/
OldAcc inside the timed region).  For N > 1 (torchrun, one process per GPU) every rank holds 1/N of the particle
records, one NCCL all-gather per step replicates them, every rank rebuilds the same tree and walks its 1/N slice of the
tree-ordered targets (strong scaling of a fixed particle set).

The reference arm (--impl reference) and the cpu_baseline object time the UNMODIFIED reference (oracle/_ref) on the host
cores; those are the only places this file touches oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "gadget-2.0.7-ngravs_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "tree-force particle updates/s"
UNIT = "updates/s"

# nominal FP32 peak of the CUDA cores: SMs x 128 lanes x 2 flop (FMA) x max SM clock
FLOP_PER_TERM = {"np": 20.0, "per": 26.0, "sr": 31.0}


# ------------------------------------------------------------------------------------------------ workloads -----
def make_workload(name):
    import g2test
    if name == "hernquist1m" or name.startswith("hernquist"):
        n = 1_000_000 if name == "hernquist1m" else int(name[len("hernquist"):])
        pos, mass, ptype = g2test.hernquist(n)
        eps = 0.05
        return dict(name=f"synthetic {n}-particle two-species Hernquist halo (a=30, r<1000, eps=0.05, seed 20261018), tree-only, "
                         "non-periodic, UNEQUALSOFTENINGS, relative criterion alpha=0.005",
                    key=name, n=n, pos=pos, mass=mass, ptype=ptype, soft=(0.0, eps, eps, eps, eps, eps), grav=(0, 0, 1, 0, 0, 0),
                    D=2, periodic=False, shortrange=False, unequal=True, box=0.0, pmgrid=0, ref_variant="np_d2_f32", flop="np")
    if name.startswith("periodic") and name.endswith("nopm"):
        # PERIODIC without PMGRID: tree walk with nearest images + lattice-sum correction walk (forcetree.c:1606-1608)
        side = int(name[len("periodic"):-4])
        n, box = side ** 3, 100000.0
        pos, mass, ptype = g2test.periodic_poisson(n, box, ntypes=2)
        eps = box / side / 30.0
        return dict(name=f"periodic {side}^3 Poisson box, two species, no PM: tree walk (nearest image) + lattice-sum correction walk "
                         "(Ewald table 65^3, made on the device), relative criterion alpha=0.005",
                    key=name, n=n, pos=pos, mass=mass, ptype=ptype, soft=(eps,) * 6, grav=(0, 0, 1, 0, 0, 0), D=2, periodic=True,
                    shortrange=False, unequal=False, box=box, pmgrid=0, ref_variant="per_d2_f32", flop="per")
    if name.startswith("periodic"):
        four = name.endswith("x4")                      # BASELINE config 4: all 6 particle types mapped onto 4 species
        side = int(name[len("periodic"):-2] if four else name[len("periodic"):])
        n = side ** 3
        box = 100000.0
        pos, mass, ptype = g2test.periodic_poisson(n, box, ntypes=6 if four else 2)
        pmgrid = 256 if side >= 128 else 64
        eps = box / side / 30.0
        D = 4 if four else 2
        return dict(name=f"periodic {side}^3 Poisson box, " + ("6 particle types (i mod 6, gas first) on 4 gravitational species (Gravity* = 0,1,2,3,1,2)"
                                                               if four else "two species") +
                         f", TreePM short-range walk, PMGRID={pmgrid}, ASMTH=1.25, RCUT=4.5, relative criterion alpha=0.005",
                    key=name, n=n, pos=pos, mass=mass, ptype=ptype, soft=(eps,) * 6, grav=(0, 1, 2, 3, 1, 2) if four else (0, 0, 1, 0, 0, 0), D=D,
                    periodic=True, shortrange=True, unequal=False, box=box, pmgrid=pmgrid,
                    ref_variant=("pm_d%d_f32" if pmgrid == 256 else "pm64_d%d_f32") % D, flop="sr")
    raise SystemExit(f"unknown workload {name}")


def pm_split(w):
    asmth = 1.25 * w["box"] / w["pmgrid"] if w["pmgrid"] else 0.0      # pm_periodic.c:59-60
    return asmth, 4.5 * asmth


# ------------------------------------------------------------------------------------------------ clocks --------
class ClockSampler(threading.Thread):
    """Samples nvidia-smi continuously; result() keeps the samples whose query overlapped the timed region (a query takes
    ~0.1 s, longer than a short timed region, so overlap rather than containment is the criterion)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.maxmhz, self.stop_flag = gpu, [], None, False
        self.t0 = self.t1 = None

    def run(self):
        while not self.stop_flag:
            ts = time.time()
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=10).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                reasons = [name for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6])
                           if v.lower().startswith("active")]
                self.samples.append((ts, time.time(), float(f[0]), reasons))
                self.maxmhz = float(f[1])
            except Exception:
                pass
            time.sleep(0.02)

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()
        time.sleep(0.25)                                 # let the query that overlaps the end of the region finish
        self.stop_flag = True

    def result(self):
        sel = [s for s in self.samples if self.t0 is not None and s[0] <= self.t1 and s[1] >= self.t0] or self.samples[-1:]
        if not sel:
            return {"sm_mhz": None, "sm_max_mhz": self.maxmhz, "reasons": []}
        reasons = sorted({r for s in sel for r in s[3]})
        return {"sm_mhz": float(np.median([s[2] for s in sel])), "sm_max_mhz": self.maxmhz, "reasons": reasons, "samples": len(sel)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------ reference arm --
def reference_run(w, oldacc_by_id, steps, warmup, sample_frac=None, nthreads=None):
    """Times the UNMODIFIED reference (oracle/_ref): domain_Decomposition + force_treebuild on one core (they are serial in the
    reference) and the reference's own per-target walk function on all host cores over a strided sample of the targets."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from refrun import RefOracle, available
    kind = "reference"
    n = w["n"]
    cores = nthreads or len(os.sched_getaffinity(0))
    if not available(w["ref_variant"]):
        raise RuntimeError("oracle/_ref is not built")
    ref = RefOracle(w["ref_variant"], int(1.1 * n) + 64, boxsize=w["box"], softening=w["soft"], gravity=w["grav"], theta=0.0, errtol=0.005,
                    criterion=1, buffer_mb=64)
    # interactions per particle ~600 (tree) / ~180 (TreePM) at ~3e7 / 1e7 per core-second: aim at ~10 s of walking
    ia = 600.0 if not w["shortrange"] else (800.0 if w["n"] > 4_000_000 else 180.0)
    rate = (3.0e7 if not w["shortrange"] else 1.0e7) * cores
    if sample_frac is None:
        sample_frac = min(1.0, 6.0 * rate / (ia * n))
    stride = max(1, int(round(1.0 / sample_frac)))
    results = []
    t_domain = t_build = None
    for it in range(warmup + steps):
        if t_domain is None:
            # domain_Decomposition + force_treebuild are timed ONCE at full size (serial in the reference, ~10-60 s at 16.8M);
            # later steps repeat only the sampled walk on the same tree so that a --steps/--warmup run stays bounded
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
            else:
                # untimed Barnes-Hut pass over the sample gives the sample's OldAcc (only a target's own OldAcc enters its walk)
                ref.set_opening(0.5, 0.005, 1)
                ref.walk_threads(cores)
                ref.set_oldacc(np.linalg.norm(ref.particles()["acc"], axis=1))
            ref.set_opening(0.0, 0.005, 1)
        t_walk_s, cost = ref.walk_threads(cores)
        ns = int(active.sum())
        t_walk = t_walk_s * n / ns
        results.append(dict(t_domain=t_domain, t_build=t_build, t_walk_sample=t_walk_s, t_walk_full=t_walk, ia_per_part=cost / ns, nsample=ns))
    res = results[warmup:]
    t_step = float(np.mean([r["t_domain"] + r["t_build"] + r["t_walk_full"] for r in res]))
    last = res[-1]
    sample = (f"{last['nsample']} of {n} targets (every {stride}th in reference order) walked by the reference's force_treeevaluate"
              f"{'_shortrange' if w['shortrange'] else ''} on {cores} threads ({last['t_walk_sample']:.2f} s, {last['ia_per_part']:.1f} ia/part), "
              f"scaled to all targets; domain_Decomposition {last['t_domain']:.2f} s + force_treebuild {last['t_build']:.2f} s on 1 core, full size")
    return dict(value=n / t_step, unit=UNIT, cores=cores, kind=kind, sample=sample, ms_per_step=1e3 * t_step,
                interactions_per_s=last["ia_per_part"] * n / t_step, detail=last)


# ------------------------------------------------------------------------------------------------ GPU arm --------
def main():
    # stdout carries exactly ONE JSON line: everything else that native libraries print there (e.g. the NCCL version banner)
    # is sent to stderr by pointing fd 1 at fd 2 and keeping a private copy of the real stdout for the result line
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="periodic256")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--active-frac", type=float, default=1.0, help="fraction of particles with Ti_endstep == Ti_Current (random, seed 7); "
                    "the metric then counts the ACTIVE particles only (SURVEY.md §8d 'sparse active' case)")
    ap.add_argument("--acc-float", action="store_true", help="FP32 accumulators in the walk (default FP64)")
    ap.add_argument("--walk-exact", type=int, default=1, help="1 (default): borderline decisions re-walked in FP64 (exact GravCost); 0: FP32 decisions only")
    ap.add_argument("--profile", action="store_true", help="short run for ncu: 1 warm-up step, no e2e / cpu_baseline legs")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    W = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.profile:
        W = 1
    K = args.steps

    w = make_workload(args.workload)
    n = w["n"]

    if args.impl == "reference":
        if rank != 0:
            return
        r = reference_run(w, None, K, W)
        line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
                "config": {"workload": w["name"], "particles": n},
                "interactions_per_s": r["interactions_per_s"],
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), file=real_stdout, flush=True)
        return

    import torch
    import torch.distributed as dist
    import g2gpu

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    tg = g2gpu.TreeGravity(max_part=int(1.1 * n) + 64, n_gravs=w["D"], periodic=w["periodic"], shortrange=w["shortrange"],
                           unequal_softenings=w["unequal"], device=local_rank, rank=rank, nranks=world)
    import g2test
    tg.set_species(w["grav"], g2test.force_softening(w["soft"]))
    tg.set_laws()
    if w["periodic"] and not w["shortrange"]:
        tg.set_ewald_lattice(w["box"])
    if args.acc_float:
        tg.set_option("acc_double", 0)
    tg.set_option("walk_exact", args.walk_exact)
    asmth, rcut = pm_split(w)
    if w["shortrange"]:
        tab = np.load(os.path.join(PKG, "data", "srtable_newton_ntab2048.npy"))
        tg.set_srtable(np.broadcast_to(tab, (w["D"], w["D"], len(tab))).copy())
    wp_bh = tg.walk_params(theta=0.5, errtol=0.005, boxsize=w["box"], G=1.0, asmth=asmth, rcut=rcut)
    wp_rel = tg.walk_params(theta=0.0, errtol=0.005, boxsize=w["box"], G=1.0, asmth=asmth, rcut=rcut)
    lib_stream = torch.cuda.ExternalStream(tg.stream, device=dev)

    # ---- untimed first force computation (Barnes-Hut, whole set on every rank) -> OldAcc for the relative criterion
    tg.set_option("nranks", 1)
    tg.set_option("rank", 0)
    acc0, cost0, old0, perm0 = tg.gravity_tree(w["pos"], w["mass"], w["ptype"], wp_bh)
    oldacc_by_id = np.zeros(n, dtype=np.float32)
    oldacc_by_id[perm0] = old0
    bh_stats = tg.timings()
    tg.set_option("nranks", world)
    tg.set_option("rank", rank)

    # ---- device-resident inputs.  N > 1: every rank owns the slice [lo,hi) and all-gathers the rest each step.
    import multigpu
    lo, hi, per = multigpu.owner_slice(n, rank, world)
    ex = multigpu.ParticleExchange(n, dev, world)
    active = None
    n_active = n
    if args.active_frac < 1.0:
        active = (np.random.default_rng(7).random(n) < args.active_frac).astype(np.int32)
        n_active = int(active.sum())
    ex.set_local(multigpu.pack_records(torch.from_numpy(w["pos"][lo:hi]).to(dev), torch.from_numpy(w["mass"][lo:hi]).to(dev),
                                       torch.from_numpy(w["ptype"][lo:hi].astype(np.int32)).to(dev), torch.from_numpy(oldacc_by_id[lo:hi]).to(dev),
                                       None if active is None else torch.from_numpy(active[lo:hi]).to(dev)))
    if world == 1:
        ex.gather()                                     # single GPU: the records simply stay resident in HBM
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)      # > 126 MB L2
    torch.cuda.synchronize()

    def step_resident():
        if world > 1:
            ex.gather()                                 # ONE NCCL all-gather of the 32-byte particle records over NVLink
            torch.cuda.current_stream().synchronize()
        tg.bind_inputs(n, ex.g_rec.data_ptr())
        tg.domain()
        tg.treebuild()
        tg.walk(wp_rel)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(W):
        flush.zero_()
        step_resident()
    tg.sync()
    barrier()
    sampler.begin()
    tg.reset_counters()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    stage = dict(domain_ms=0.0, build_ms=0.0, walk_ms=0.0, walk_kernel_ms=0.0, sort_ms=0.0)
    wall0 = time.time()
    for k in range(K):
        flush.zero_()                                   # L2 flush between timed iterations (untimed)
        torch.cuda.synchronize()
        ev[k][0].record(lib_stream)
        t_host0 = time.time()
        step_resident()
        ev[k][1].record(lib_stream)
        tg.sync()
        t = tg.timings()
        for key in stage:
            stage[key] += t[key]
        if k == K - 1:
            last = t
    barrier()
    wall = time.time() - wall0
    sampler.end()
    launches = tg.timings()["launches"]
    # one more, untimed step with the instrumented walk kernel: species terms, cell visits and decisions for the roofline figures
    tg.set_option("walk_stats", 1)
    step_resident()
    tg.sync()
    last = tg.timings()
    tg.set_option("walk_stats", 0)
    ms_dev = sum(a.elapsed_time(b) for a, b in ev)
    t_ms = torch.tensor([ms_dev], dtype=torch.float64, device=dev)
    inter = torch.tensor([float(last["interactions"]), float(last["species_terms"]), float(last["cell_visits"]), float(last["decisions"])],
                         dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(inter, op=dist.ReduceOp.SUM)
    ms_step = float(t_ms.item()) / K
    interactions, terms, visits, decisions = (float(x) for x in inter.tolist())
    value = n_active / (ms_step * 1e-3)

    # ---- e2e: host buffers -> g2gpu_gravity_tree -> host buffers (N == 1: whole set; N > 1: reported for rank 0's view)
    e2e = None
    if world == 1 and not args.profile and active is None:
        h_pos = torch.from_numpy(w["pos"]).pin_memory().numpy()
        h_mass = torch.from_numpy(w["mass"]).pin_memory().numpy()
        h_type = torch.from_numpy(w["ptype"].astype(np.int32)).pin_memory().numpy()
        h_old = torch.from_numpy(oldacc_by_id).pin_memory().numpy()
        outbuf = (torch.zeros((n, 3), dtype=torch.float32).pin_memory().numpy(), torch.zeros(n, dtype=torch.float32).pin_memory().numpy(),
                  torch.zeros(n, dtype=torch.float32).pin_memory().numpy(), torch.zeros(n, dtype=torch.int32).pin_memory().numpy())
        for _ in range(2):
            tg.gravity_tree(h_pos, h_mass, h_type, wp_rel, oldacc=h_old, out=outbuf)
        e_ms = 0.0
        for k in range(K):
            flush.zero_()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(lib_stream)
            acc_e, cost_e, old_e, perm_e = tg.gravity_tree(h_pos, h_mass, h_type, wp_rel, oldacc=h_old, out=outbuf)
            b.record(lib_stream)
            tg.sync()
            e_ms += a.elapsed_time(b)
        h2d, d2h = tg.io_bytes()
        d2h += 4 * n                                   # the permutation (perm) read back with the results
        e2e = {"value": n / (e_ms / K * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e_ms / K,
               "checksum": float(np.abs(acc_e).sum())}

    if world > 1 and not args.profile and active is None:
        # N > 1: every rank copies ITS OWN 1/N of the particle records from pinned host memory, the all-gather replicates them, and every
        # rank reads the result arrays back (its own targets are the entries that changed); max over ranks
        cnt = hi - lo
        h_rec = multigpu.pack_records(torch.from_numpy(w["pos"][lo:hi]), torch.from_numpy(w["mass"][lo:hi]), torch.from_numpy(w["ptype"][lo:hi].astype(np.int32)),
                                      torch.from_numpy(oldacc_by_id[lo:hi])).pin_memory()
        outbuf = (torch.zeros((n, 3), dtype=torch.float32).pin_memory().numpy(), torch.zeros(n, dtype=torch.float32).pin_memory().numpy(),
                  torch.zeros(n, dtype=torch.float32).pin_memory().numpy())

        def step_e2e():
            ex.s_rec[:cnt].copy_(h_rec, non_blocking=True)
            step_resident()
            return tg.download_acc(out=outbuf)
        for _ in range(2):
            step_e2e()
        barrier()
        e_ms = 0.0
        for k in range(K):
            flush.zero_()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(torch.cuda.current_stream())
            acc_e, cost_e, old_e = step_e2e()
            b.record(lib_stream)
            tg.sync()
            e_ms += a.elapsed_time(b)
        t_e = torch.tensor([e_ms / K], dtype=torch.float64, device=dev)
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        e_ms = float(t_e.item())
        e2e = {"value": n / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": cnt * 32, "d2h_bytes_per_step": 20 * n, "ms_per_step": e_ms,
               "note": "per rank: H2D of its 1/N of the 32-byte records, D2H of the whole result arrays; max over ranks",
               "checksum": float(np.abs(acc_e).sum())}

    # ---- the tree potential of every particle (compute_potential -> force_treeevaluate_potential[_shortrange]) on the tree
    # of the last step, timed by the library's own CUDA events around the kernel; not part of `value`
    potw = None
    if world == 1 and not args.profile and active is None and (w["shortrange"] or not w["periodic"]):
        tg.set_potential_laws()
        if w["shortrange"]:
            ptab = np.load(os.path.join(PKG, "data", "srpot_newton_ntab2048.npy"))
            tg.set_srpot_table(np.broadcast_to(ptab, (w["D"], w["D"], len(ptab))).copy())
        pot_ms = []
        for _ in range(3):
            flush.zero_()
            torch.cuda.synchronize()
            _, ms1 = tg.potential(wp_rel, with_time=True)
            pot_ms.append(ms1)
        potw = {"kernel": "pot_kernel", "ms_per_call": float(np.median(pot_ms)), "particles_per_s": n / (float(np.median(pot_ms)) * 1e-3),
                "criterion": "relative (OldAcc of the first pass)", "note": "all particles are targets (potential.c:86)"}

    # ---- the long-range complement of the TreePM split (pmforce_periodic on the device); reported beside the tree numbers, not part
    # of the step.  Device time from the library's own events around the PM stage, inputs resident (uploaded once).
    pm = None
    if world == 1 and w["shortrange"] and not args.profile and active is None:
        # P[] is in Peano-Hilbert order when the reference runs its PM step (domain_Decomposition precedes it, domain.c:65-72)
        tg.upload(w["pos"][perm0], w["mass"][perm0], w["ptype"][perm0])
        for _ in range(2):
            tg.pm_device(w["pmgrid"], w["box"])
        pm_ms = 0.0
        for _ in range(K):
            flush.zero_()
            torch.cuda.synchronize()
            tg.pm_device(w["pmgrid"], w["box"])
            pm_ms += tg.timings()["pm_ms"]
        pm_ms /= K
        N3, D = float(w["pmgrid"]) ** 3, w["D"]
        # algorithmic DRAM bytes of one call, every mesh touched the minimum number of times: clear (8 N^3 per species),
        # deposit (32 B per particle + one read-modify-write of the mesh), forward transform (8 N^3 in, 8 N^3 out per species), filter
        # (D + 1 spectra of 8 N^3 per target species), inverse transform (8 N^3 + 8 N^3), gather (32 B read + 12 B written per particle +
        # the potential mesh once).  The 56 potential values a particle gathers are served by L1/L2 and not counted.
        pm_bytes = D * 8 * N3 + n * 32 + D * 16 * N3 + D * 16 * N3 + D * (D + 1) * 8 * N3 + D * 16 * N3 + n * 44 + D * 8 * N3
        hbm = (measured_peaks() or {}).get("hbm_gbs", 6650.0)
        pm = {"kernel": "pm_deposit/cuFFT D2Z/pm_filter/cuFFT Z2D/pm_gather", "pmgrid": w["pmgrid"], "ms_per_call": pm_ms, "particle_order": "species-major Peano-Hilbert (as after domain_Decomposition)",
              "particles_per_s": n / (pm_ms * 1e-3), "bound": "hbm", "algorithmic_bytes": pm_bytes,
              "achieved": pm_bytes / (pm_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": pm_bytes / (pm_ms * 1e-3) / 1e9 / hbm}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    prop = torch.cuda.get_device_properties(local_rank)
    peaks = measured_peaks()
    sm_max = (peaks or {}).get("sm_max_mhz", 1965.0)
    fp32_peak = prop.multi_processor_count * 128 * 2 * sm_max * 1e6 / 1e12
    walk_ms = stage["walk_kernel_ms"] / K
    flops_alg = terms * FLOP_PER_TERM[w["flop"]]
    achieved = flops_alg / (walk_ms * 1e-3) / 1e12 * (1.0 if world == 1 else 1.0 / world)
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json"))).get(w["key"])
    except Exception:
        traffic = None
    # visit-inclusive figure of SURVEY.md §8d: every opening decision of a target (accepted or not) costs 16 D + 6 flop (+ 12 TreePM cull)
    flops_dec = decisions * (16.0 * w["D"] + 6.0 + (12.0 if w["shortrange"] else 0.0))
    achieved_incl = (flops_alg + flops_dec) / (walk_ms * 1e-3) / 1e12 * (1.0 if world == 1 else 1.0 / world)
    roofline = {"kernel": "walk_kernel", "bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                "traffic": traffic, "share_of_step": walk_ms / ms_step,
                "visit_inclusive": {"achieved": achieved_incl, "frac": achieved_incl / fp32_peak, "decisions_per_particle": decisions / max(n_active, 1),
                                    "flop_per_decision": 16.0 * w["D"] + 6.0 + (12.0 if w["shortrange"] else 0.0)},
                "note": f"algorithmic flops = species terms x {FLOP_PER_TERM[w['flop']]:.0f} (SURVEY.md §8d), cell-opening arithmetic excluded; peak = "
                        f"{prop.multi_processor_count} SMs x 128 lanes x 2 x {sm_max:.0f} MHz (nominal CUDA-core FP32, of MEASURED_PEAKS sm_max_mhz)"}
    # ncu's own utilisation figures of the same kernel on the same workload, from the committed capture (never measured under this run)
    for suffix in ("r1e", "r1b", "final", "v1"):
        prof = os.path.join(ROOT, "profiles", f"r1_walk_{w['key']}_{suffix}.txt")
        if os.path.exists(prof):
            ncu = {"source": os.path.relpath(prof, ROOT)}
            for ln in open(prof):
                f = ln.split()
                if len(f) >= 2 and f[0] in ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                                            "smsp__thread_inst_executed_per_inst_executed.ratio"):
                    ncu[f[0]] = float(f[1])
            roofline["ncu"] = ncu
            break
    hbm_peak = (peaks or {}).get("hbm_gbs", 6650.0)
    try:
        sort_traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json"))).get("sort_stage1_" + w["key"])
    except Exception:
        sort_traffic = None
    # radix sorts of stage 1: 7 passes x (8 B hist read + 12 B scatter read + 12 B write) per pair
    sort_bytes = n * 7 * 32.0
    roofline_sort = {"kernel": "sort_hist_kernel+sort_scatter_kernel (stage 1, 7 passes)", "bound": "hbm",
                     "achieved": sort_bytes / (stage["sort_ms"] / K * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": sort_bytes / (stage["sort_ms"] / K * 1e-3) / 1e9 / hbm_peak, "traffic": sort_traffic,
                     "note": "of measured" if peaks else "of fallback"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"] + ("" if active is None else f", {n_active} of {n} particles active (random {args.active_frac:g})"),
                       "particles": n, "active": n_active, "l2": "flushed between timed iterations (256 MiB write)",
                       "parallelism": f"replicated tree, {world} equal tree-order target slices" + (", NCCL all-gather of particle records per step" if world > 1 else "")},
            "interactions_per_s": interactions / (ms_step * 1e-3), "ia_per_particle": interactions / n_active,
            "stages_ms": {k: v / K for k, v in stage.items()}, "cell_visits_per_warp_step": visits,
            "first_pass_barnes_hut": {"ia_per_particle": bh_stats["interactions"] / n, "walk_kernel_ms": bh_stats["walk_kernel_ms"]},
            "wall_ms_per_step_incl_flush": 1e3 * wall / K,
            "allgather_bytes_per_step": ex.bytes_per_step() if world > 1 else 0,
            "gpu_launches": int(launches), "clocks": sampler.result(), "roofline": roofline, "roofline_sort": roofline_sort}
    if e2e is not None:
        line["e2e"] = e2e
    if pm is not None:
        line["pm_long_range"] = pm
    if potw is not None:
        line["potential_walk"] = potw
    if w["periodic"] and not w["shortrange"]:
        line["lattice_correction"] = {"kernel": "lattice_kernel", "ms_per_step": (stage["walk_ms"] - stage["walk_kernel_ms"]) / K,
                                      "note": "walk stage minus walk kernel: target compaction + lattice-sum correction walk of all targets"}
    if world == 1 and not args.no_cpu_baseline and not args.profile and active is None:
        try:
            r = reference_run(w, oldacc_by_id, 1, 0)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                                    "interactions_per_s": r["interactions_per_s"]}
        except Exception as e:                           # the baseline is a reported number, never a dependency of the GPU arm
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
    print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
