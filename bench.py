#!/usr/bin/env python
"""bench.py — tree-force throughput of the B200 path (and of the reference's CPU path beside it).

    python bench.py --gpus N --steps K --warmup W [--workload hernquist1m|periodic128|periodic256|periodic256x4] [--active-frac F]
                    [--impl reference] [--ref-threads T]

A "step" is one pass of the hot path over the particle set: domain_findExtent + Peano-Hilbert keys + radix sort +
top-level tree + species-major order, force_treebuild (tree + per-species moments) and the tree walk of every active
particle with the relative opening criterion (OldAcc from an untimed Barnes-Hut first pass, as in a run: accel.c:46-49,
gravtree.c:334-335).  `value` = particles updated per second with the inputs resident in HBM (CUDA events on every device's
stream, maximum over the devices); `e2e` = the same through g2gpu_group_gravity_tree() with pinned HOST buffers, wall clock
around the call (sharded H2D of positions/masses/types/OldAcc, slice D2H of accelerations/GravCost/OldAcc and the host scatter
inside the timed region); `e2e_shim` = the reference's own gravity_tree() entry point served by the host shim on the
reference's AoS P[] (what accel.c calls).

N > 1 runs through the C library's multi-GPU group (csrc/g2_group.cu): ONE process, one host thread + context + stream per
device, sharded upload, one ncclAllGather of the 32-byte records per step, replicated tree, one target slice per device
(strong scaling of a fixed particle set).  Under torchrun (the driver's launch for N > 1) rank 0 drives all N devices through
that C path; the other ranks only take part in the barriers.

The reference arm (--impl reference), the cpu_baseline / parity objects and the e2e_shim leg are the only places this file
touches oracle/ (the unmodified reference compiled under oracle/_ref, and its harness as the caller of the shim).
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
# radix sort of stage 1: 7 passes x (8 B histogram read + 12 B scatter read + 12 B write) per pair
SORT_BYTES_PER_PAIR = 7 * 24.0 + 8.0     # one-sweep: 12 B read + 12 B written per pair and pass, keys read once more for the histograms of all passes
SORT_KERNELS = "os_hist_kernel+os_pass_kernel (domain sort, 7 one-sweep passes of 9 bits)"


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


def config_of(w, n_active, active_frac, ngpu):
    """The `config` object of the JSON line: IDENTICAL for the GPU arm and the reference arm of one workload."""
    return {"workload": w["name"] + ("" if n_active == w["n"] else f", {n_active} of {w['n']} particles active (random {active_frac:g})"),
            "particles": w["n"], "active": n_active, "gpus": ngpu}


# ------------------------------------------------------------------------------------------------ reference arm --
def reference_run(w, oldacc_by_id, steps, warmup, nthreads=None, stride=None, keep_state=False):
    """Times the UNMODIFIED reference (oracle/_ref): domain_Decomposition + force_treebuild at full size on one core (they are serial
    in the reference; timed ONCE per run), and in every step the reference's own per-target walk function on `nthreads` host threads over
    a strided SAMPLE of the targets.  `value` extrapolates the sample to all targets (formula in the record); `ms_timed_per_step` is
    what the host really spent per step."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import g2parity
    n = w["n"]
    cores = nthreads or len(os.sched_getaffinity(0))
    # interactions per particle ~600 (tree) / ~180..800 (TreePM) at ~3e7 / 1e7 per core-second: aim at ~2.5 s of walking per step
    ia = 600.0 if not w["shortrange"] else (800.0 if n > 4_000_000 else 180.0)
    rate = (3.0e7 if not w["shortrange"] else 1.0e7) * cores
    if stride is None:
        stride = max(1, int(round(ia * n / (2.5 * rate))))
    t_run0 = time.time()
    st = g2parity.reference_state(w, oldacc_by_id, stride, cores, theta=0.0 if oldacc_by_id is not None else 0.5)
    ref = st["ref"]
    if oldacc_by_id is None:
        # the sample's OldAcc from its own (untimed) Barnes-Hut walk: only a target's own OldAcc enters its walk
        ref.set_oldacc(np.linalg.norm(ref.particles()["acc"], axis=1))
        ref.set_opening(0.0, 0.005, 1)
        st["t_walk_sample"], st["cost_sum"] = ref.walk_threads(cores)
    walks = [st["t_walk_sample"]]
    for _ in range(warmup + steps - 1):
        t, c = ref.walk_threads(cores)
        walks.append(t)
    t_total = time.time() - t_run0
    ns = st["nsample"]
    t_walk_s = float(np.mean(walks[warmup:])) if len(walks) > warmup else float(walks[-1])
    t_walk_full = t_walk_s * n / ns
    t_step = st["t_domain"] + st["t_build"] + t_walk_full
    sample = (f"{ns} of {n} targets (every {stride}th in reference order) walked by the reference's force_treeevaluate"
              f"{'_shortrange' if w['shortrange'] else ''} on {cores} threads ({t_walk_s:.2f} s per step, {st['cost_sum'] / ns:.1f} ia/part); "
              f"domain_Decomposition {st['t_domain']:.2f} s + force_treebuild {st['t_build']:.2f} s on 1 core at full size, once per run")
    out = dict(value=n / t_step, unit=UNIT, cores=cores, kind="reference", sample=sample, sampled=True, sample_stride=stride, sample_targets=ns,
               formula="particles / (t_domain + t_build + t_walk_sample * particles / sample_targets)",
               ms_per_step_extrapolated=1e3 * t_step, ms_timed_per_step=1e3 * t_total / max(1, warmup + steps),
               seconds=dict(domain=st["t_domain"], build=st["t_build"], walk_sample=t_walk_s, walk_extrapolated=t_walk_full, run=t_total),
               walk_only=dict(value=n / t_walk_full, unit=UNIT, note="walk alone (domain_Decomposition and force_treebuild excluded), extrapolated from the sample"),
               interactions_per_s=st["cost_sum"] / ns * n / t_step, ia_per_particle=st["cost_sum"] / ns)
    if keep_state:
        out["state"] = st
    return out


# ------------------------------------------------------------------------------------------------ e2e through the shim --
SHIM_VARIANT = {"periodic256": "pm_d2_f32", "periodic128": "pm_d2_f32", "hernquist1m": "np_d2_f32"}


def shim_e2e(w, oldacc_by_id, steps, ngpu):
    """The reference's own entry point: the unmodified reference (domain.c, allvars.c, ...) linked with the product's host shim
    (oracle/_ref/libg2shim_<variant>.so, integration/Makefile), its AoS P[] filled with the workload, gravity_tree() called as accel.c
    calls it.  Timed per call, wall clock: pack of P[] into pinned records by host threads, sharded H2D, all-gather, keys + sort + tree +
    walk on the device(s), slice D2H, host threads writing GravAccel/GravCost/OldAcc back into P[]."""
    variant = SHIM_VARIANT.get(w["key"])
    if variant is None:
        return None
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from refrun import RefOracle, available
    if not available(variant, "g2shim"):
        return {"unavailable": f"oracle/_ref/libg2shim_{variant}.so not built"}
    os.environ["G2GPU_NGPU"] = str(ngpu)
    n = w["n"]
    shim = RefOracle(variant, int(1.1 * n) + 64, prefix="g2shim", boxsize=w["box"], softening=w["soft"], gravity=w["grav"], theta=0.0, errtol=0.005,
                     criterion=1, buffer_mb=64)
    shim.load(w["pos"], w["mass"], w["ptype"])
    t0 = time.time()
    shim.domain()                      # the reference's domain_Decomposition on the host; its peano_hilbert_order() is the shim's (device)
    t_domain = time.time() - t0
    sp = shim.particles()
    shim.set_oldacc(oldacc_by_id[sp["id"]])
    shim.set_opening(0.0, 0.005, 1)
    times = []
    for k in range(2 + steps):
        shim.force_rebuild()           # TreeReconstructFlag: gravity_tree() builds the tree from the current P[] (as after every domain decomposition)
        t0 = time.perf_counter()
        shim.gravity()
        times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times[2:]))
    cost = shim.particles()["cost"]
    return {"value": n / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "entry_point": "gravity_tree() of host/g2_shim.c on the reference's P[] (AoS)",
            "library": f"libg2shim_{variant}.so", "ia_per_particle": float(cost.mean()), "gpus": ngpu,
            "domain_Decomposition_s": t_domain, "note": "wall clock per call; force_treebuild + walk + write-back into P[]"}


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
    ap.add_argument("--no-shim", action="store_true", help="skip the e2e_shim leg")
    ap.add_argument("--ref-threads", type=int, default=0, help="host threads of the reference arm / cpu_baseline (0 = all cores this process may use)")
    ap.add_argument("--active-frac", type=float, default=1.0, help="fraction of particles with Ti_endstep == Ti_Current (random, seed 7); "
                    "the metric then counts the ACTIVE particles only (SURVEY.md §8d 'sparse active' case)")
    ap.add_argument("--acc-float", action="store_true", help="FP32 accumulators and FP32-only decisions in the walk (default: FP64 accumulators, exact GravCost)")
    ap.add_argument("--walk-exact", type=int, default=1, help="1 (default): borderline decisions checked in FP64 (exact GravCost); 0: FP32 decisions only")
    ap.add_argument("--equal-slices", action="store_true", help="N > 1: equal-count target slices instead of equal GravCost")
    ap.add_argument("--profile", action="store_true", help="short run for ncu: 1 warm-up step, no e2e / cpu_baseline legs")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ngpu = max(args.gpus, 1)
    W = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.profile:
        W = 1
    K = args.steps
    ref_threads = args.ref_threads or None

    # torchrun starts one process per GPU; the data plane lives in ONE process (rank 0 drives every device through the C library's
    # group), so the other ranks only meet rank 0 at the barriers
    dist = None
    if world > 1:
        import datetime
        import torch.distributed as dist
        from torch.distributed.distributed_c10d import _get_default_store
        dist.init_process_group("gloo")
        # no collective at the end of the run: a rank that leaves a gloo barrier first closes its sockets under the others.  Rank 0
        # posts a key in the launcher's store when it is done and the other ranks wait for it.
        if rank != 0:
            _get_default_store().wait(["g2gpu_bench_done"], datetime.timedelta(hours=8))
            dist.destroy_process_group()
            return

    w = make_workload(args.workload)
    n = w["n"]
    active = None
    n_active = n
    if args.active_frac < 1.0:
        active = (np.random.default_rng(7).random(n) < args.active_frac).astype(np.int32)
        n_active = int(active.sum())
    config = config_of(w, n_active, args.active_frac, ngpu)

    if args.impl == "reference":
        r = reference_run(w, None, K, W, nthreads=ref_threads)
        line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": ngpu, "steps": K, "warmup": W, "ms_per_step": r["ms_timed_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
                "config": config, "sampled": True, "sample_stride": r["sample_stride"], "sample_targets": r["sample_targets"],
                "value_formula": r["formula"], "ms_per_step_extrapolated": r["ms_per_step_extrapolated"], "seconds": r["seconds"],
                "walk_only": r["walk_only"], "interactions_per_s": r["interactions_per_s"],
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), file=real_stdout, flush=True)
        if dist is not None:
            _get_default_store().set("g2gpu_bench_done", "1")
            dist.destroy_process_group()
        return

    import torch
    import g2gpu
    import g2test

    grp = g2gpu.TreeGravityGroup(max_part=int(1.1 * n) + 64, n_gravs=w["D"], periodic=w["periodic"], shortrange=w["shortrange"],
                                 unequal_softenings=w["unequal"], ndev=ngpu)
    grp.set_species(w["grav"], g2test.force_softening(w["soft"]))
    grp.set_laws()
    tg0 = grp.ctx(0)
    if w["periodic"] and not w["shortrange"]:
        for i in range(ngpu):
            grp.ctx(i).set_ewald_lattice(w["box"])
    if args.acc_float:
        grp.set_option("acc_double", 0)
    grp.set_option("walk_exact", args.walk_exact)
    if args.equal_slices:
        grp.set_option("cost_weighted", 0)
    asmth, rcut = pm_split(w)
    if w["shortrange"]:
        tab = np.load(os.path.join(PKG, "data", "srtable_newton_ntab2048.npy"))
        grp.set_srtable(np.broadcast_to(tab, (w["D"], w["D"], len(tab))).copy())
    wp_bh = grp.walk_params(theta=0.5, errtol=0.005, boxsize=w["box"], G=1.0, asmth=asmth, rcut=rcut)
    wp_rel = grp.walk_params(theta=0.0, errtol=0.005, boxsize=w["box"], G=1.0, asmth=asmth, rcut=rcut)
    devs = [torch.device("cuda", i) for i in range(ngpu)]
    streams = [torch.cuda.ExternalStream(grp.stream(i), device=devs[i]) for i in range(ngpu)]

    # ---- untimed first force computation (Barnes-Hut, all particles) -> OldAcc for the relative criterion
    acc0, cost0, old0, perm0 = grp.gravity_tree(w["pos"], w["mass"], w["ptype"], wp_bh)
    oldacc_by_id = np.zeros(n, dtype=np.float32)
    oldacc_by_id[perm0] = old0
    bh_stats = grp.timings()
    del acc0, cost0

    # ---- device-resident inputs: every device keeps ITS shard of the records; a step re-runs the all-gather from the shards
    grp.upload(w["pos"], w["mass"], w["ptype"], oldacc=oldacc_by_id, active=active)
    flush = []
    for d in devs:
        with torch.cuda.device(d):
            flush.append(torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=d))      # > 126 MB L2
    for d in devs:
        torch.cuda.synchronize(d)

    def flush_l2():
        for f in flush:
            f.zero_()
        for d in devs:
            torch.cuda.synchronize(d)

    def barrier():
        grp.sync()
        for d in devs:
            torch.cuda.synchronize(d)

    sampler = ClockSampler(0)
    sampler.start()
    for _ in range(W):
        flush_l2()
        grp.step_resident(n, wp_rel)
    barrier()
    sampler.begin()
    stage = dict(domain_ms=0.0, build_ms=0.0, walk_ms=0.0, walk_kernel_ms=0.0, sort_ms=0.0)
    launches0 = grp.timings()["launches"]
    ms_dev = 0.0
    wall0 = time.time()
    for k in range(K):
        flush_l2()                                      # L2 flush between timed iterations (untimed)
        ev = []
        for i, d in enumerate(devs):
            with torch.cuda.device(d):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(streams[i])
                ev.append((a, b))
        grp.step_resident(n, wp_rel)
        for i, d in enumerate(devs):
            with torch.cuda.device(d):
                ev[i][1].record(streams[i])
        barrier()
        ms_dev += max(a.elapsed_time(b) for a, b in ev)  # the step ends when the slowest device is done
        t = grp.timings()
        for key in stage:
            stage[key] += t[key]
    wall = time.time() - wall0
    sampler.end()
    launches = grp.timings()["launches"] - launches0
    ms_step = ms_dev / K
    value = n_active / (ms_step * 1e-3)
    slices_lo, slices_hi, next_frac = grp.slices()
    # one more, untimed step with the instrumented walk kernel: species terms, cell visits and decisions for the roofline figures
    grp.set_option("walk_stats", 1)
    grp.step_resident(n, wp_rel)
    barrier()
    last = grp.timings()
    grp.set_option("walk_stats", 0)
    interactions, terms, visits, decisions = float(last["interactions"]), float(last["species_terms"]), float(last["cell_visits"]), float(last["decisions"])

    # ---- e2e: pinned host buffers -> g2gpu_group_gravity_tree -> host buffers, wall clock around the call
    e2e = None
    acc_e = cost_e = perm_e = None
    if not args.profile and active is None:
        h_pos = torch.from_numpy(w["pos"]).pin_memory().numpy()
        h_mass = torch.from_numpy(w["mass"]).pin_memory().numpy()
        h_type = torch.from_numpy(w["ptype"].astype(np.int32)).pin_memory().numpy()
        h_old = torch.from_numpy(oldacc_by_id).pin_memory().numpy()
        outbuf = (torch.zeros((n, 3), dtype=torch.float32).pin_memory().numpy(), torch.zeros(n, dtype=torch.float32).pin_memory().numpy(),
                  torch.zeros(n, dtype=torch.float32).pin_memory().numpy(), torch.zeros(n, dtype=torch.int32).pin_memory().numpy())
        for _ in range(2):
            grp.gravity_tree(h_pos, h_mass, h_type, wp_rel, oldacc=h_old, out=outbuf)
        e_s = 0.0
        for k in range(K):
            flush_l2()
            t0 = time.perf_counter()
            acc_e, cost_e, old_e, perm_e = grp.gravity_tree(h_pos, h_mass, h_type, wp_rel, oldacc=h_old, out=outbuf)
            e_s += time.perf_counter() - t0
        h2d, d2h, gathered = grp.io_bytes()
        e2e = {"value": n / (e_s / K), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e_s / K,
               "allgather_bytes_per_device": gathered, "clock": "host wall clock around g2gpu_group_gravity_tree (results in host memory on return)",
               "checksum": float(np.abs(acc_e).sum()),
               "results": ("stored by the walk kernel straight into the pinned result arrays (zero-copy; d2h bytes = those stores + the particle order)"
                           if grp.zero_copy() else "staged on the devices, downloaded and scattered by host threads")}

    # ---- the same through the entry point the reference calls (gravity_tree() of the host shim on the reference's P[])
    e2e_shim = None
    if not args.profile and not args.no_shim and active is None:
        try:
            e2e_shim = shim_e2e(w, oldacc_by_id, min(K, 5), ngpu)
        except Exception as e:
            e2e_shim = {"unavailable": str(e)[:300]}

    # ---- the tree potential of every particle (compute_potential -> force_treeevaluate_potential[_shortrange]) on the tree
    # of the last step, timed by the library's own CUDA events around the kernel; not part of `value`
    potw = None
    if ngpu == 1 and not args.profile and active is None and (w["shortrange"] or not w["periodic"]):
        tg0.n = n
        tg0.set_option("nranks", 1)
        tg0.set_potential_laws()
        if w["shortrange"]:
            ptab = np.load(os.path.join(PKG, "data", "srpot_newton_ntab2048.npy"))
            tg0.set_srpot_table(np.broadcast_to(ptab, (w["D"], w["D"], len(ptab))).copy())
        pot_ms = []
        for _ in range(3):
            flush_l2()
            _, ms1 = tg0.potential(wp_rel, with_time=True)
            pot_ms.append(ms1)
        potw = {"kernel": "pot_kernel", "ms_per_call": float(np.median(pot_ms)), "particles_per_s": n / (float(np.median(pot_ms)) * 1e-3),
                "criterion": "relative (OldAcc of the first pass)", "note": "all particles are targets (potential.c:86)"}

    # ---- cpu_baseline + parity at FULL size: the unmodified reference's domain + build + sampled walk on the same particles, compared with
    # the device's order, tree (bit for bit) and the sampled targets' accelerations / GravCost
    cpu = parity = None
    if ngpu == 1 and not args.no_cpu_baseline and not args.profile and active is None:
        try:
            import g2parity
            r = reference_run(w, oldacc_by_id, 1, 0, nthreads=ref_threads, keep_state=True)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"], "sampled": True,
                   "sample_stride": r["sample_stride"], "value_formula": r["formula"], "seconds": r["seconds"], "walk_only": r["walk_only"],
                   "interactions_per_s": r["interactions_per_s"]}
            tg0.n = n
            parity = g2parity.compare_with_device(tg0, r["state"], acc_e, cost_e, perm_e)
            parity["tolerance"] = "median <= 1e-6, p99.9 <= 3e-4 (FP32 terms vs the reference's double), tree and GravCost exact"
            del r
        except Exception as e:                           # the baseline is a reported number, never a dependency of the GPU arm
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}

    # ---- the long-range complement of the TreePM split (pmforce_periodic on the device); reported beside the tree numbers, not part
    # of the step.  Device time from the library's own events around the PM stage, inputs resident (uploaded once).
    pm = None
    if ngpu == 1 and w["shortrange"] and not args.profile and active is None:
        # P[] is in Peano-Hilbert order when the reference runs its PM step (domain_Decomposition precedes it, domain.c:65-72)
        tg0.upload(w["pos"][perm0], w["mass"][perm0], w["ptype"][perm0])
        for _ in range(2):
            tg0.pm_device(w["pmgrid"], w["box"])
        pm_ms = 0.0
        for _ in range(K):
            flush_l2()
            tg0.pm_device(w["pmgrid"], w["box"])
            pm_ms += tg0.timings()["pm_ms"]
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

    prop = torch.cuda.get_device_properties(0)
    peaks = measured_peaks()
    sm_max = (peaks or {}).get("sm_max_mhz", 1965.0)
    fp32_peak = prop.multi_processor_count * 128 * 2 * sm_max * 1e6 / 1e12
    walk_ms = stage["walk_kernel_ms"] / K
    flops_alg = terms * FLOP_PER_TERM[w["flop"]]
    achieved = flops_alg / (walk_ms * 1e-3) / 1e12 / ngpu
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json"))).get(w["key"])
    except Exception:
        traffic = None
    # visit-inclusive figure of SURVEY.md §8d: every opening decision of a target (accepted or not) costs 16 D + 6 flop (+ 12 TreePM cull)
    flops_dec = decisions * (16.0 * w["D"] + 6.0 + (12.0 if w["shortrange"] else 0.0))
    achieved_incl = (flops_alg + flops_dec) / (walk_ms * 1e-3) / 1e12 / ngpu
    roofline = {"kernel": "walk_kernel", "bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                "traffic": traffic, "share_of_step": walk_ms / ms_step,
                "visit_inclusive": {"achieved": achieved_incl, "frac": achieved_incl / fp32_peak, "decisions_per_particle": decisions / max(n_active, 1),
                                    "flop_per_decision": 16.0 * w["D"] + 6.0 + (12.0 if w["shortrange"] else 0.0)},
                "species_terms_per_particle": terms / max(n_active, 1), "cell_visits_per_32_targets": visits / max(n_active / 32.0, 1.0),
                "note": f"per GPU; algorithmic flops = species terms x {FLOP_PER_TERM[w['flop']]:.0f} (SURVEY.md §8d), cell-opening arithmetic excluded; peak = "
                        f"{prop.multi_processor_count} SMs x 128 lanes x 2 x {sm_max:.0f} MHz (nominal CUDA-core FP32, of MEASURED_PEAKS sm_max_mhz); "
                        "the ceiling under per-particle reference decisions is derived in DESIGN.md §5"}
    # ncu's own utilisation figures of the same kernel on the same workload, from the committed capture (never measured under this run)
    for suffix in ("r2",):
        prof = os.path.join(ROOT, "profiles", f"r2_walk_{w['key']}.txt")
        if os.path.exists(prof):
            ncu = {"source": os.path.relpath(prof, ROOT),
                   "note": "committed capture (taken on the 172.9 ms state of the kernel at 256^3; the final kernel issues ~7 % fewer instructions, DESIGN.md section 5)"}
            for ln in open(prof):
                f = ln.split()
                if len(f) >= 2 and f[0] in ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                                            "smsp__thread_inst_executed_per_inst_executed.ratio"):
                    ncu[f[0]] = float(f[1])
            roofline["ncu"] = ncu
    hbm_peak = (peaks or {}).get("hbm_gbs", 6650.0)
    try:
        sort_traffic = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json"))).get("sort_stage1_" + w["key"])
    except Exception:
        sort_traffic = None
    sort_bytes = n * SORT_BYTES_PER_PAIR
    roofline_sort = {"kernel": SORT_KERNELS, "bound": "hbm",
                     "achieved": sort_bytes / (stage["sort_ms"] / K * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": sort_bytes / (stage["sort_ms"] / K * 1e-3) / 1e9 / hbm_peak, "traffic": sort_traffic,
                     "note": "of measured" if peaks else "of fallback"}

    config = dict(config)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ngpu, "steps": K, "warmup": W, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "l2": "flushed between timed iterations (256 MiB write per device)",
            "parallelism": (f"one process, {ngpu} devices (g2gpu_group): replicated tree, {ngpu} Peano-Hilbert target slices of equal "
                            f"{'count' if args.equal_slices else 'GravCost'}, one ncclAllGather of the particle records per step") if ngpu > 1 else "single device",
            "interactions_per_s": interactions / (ms_step * 1e-3), "ia_per_particle": interactions / n_active,
            "stages_ms": {k: v / K for k, v in stage.items()},
            "first_pass_barnes_hut": {"ia_per_particle": bh_stats["interactions"] / n, "walk_kernel_ms": bh_stats["walk_kernel_ms"]},
            "wall_ms_per_step_incl_flush": 1e3 * wall / K,
            "allgather_bytes_per_step": grp.io_bytes()[2] if ngpu > 1 else 0,
            "target_slices": {"lo": [int(x) for x in slices_lo], "hi": [int(x) for x in slices_hi]} if ngpu > 1 else None,
            "fp64_checked_comparisons": int(last.get("border_checked", 0)), "rewalked_targets": int(last.get("rewalked", 0)),
            "gpu_launches": int(launches), "clocks": sampler.result(), "roofline": roofline, "roofline_sort": roofline_sort}
    if e2e is not None:
        line["e2e"] = e2e
    if e2e_shim is not None:
        line["e2e_shim"] = e2e_shim
    if pm is not None:
        line["pm_long_range"] = pm
    if potw is not None:
        line["potential_walk"] = potw
    if w["periodic"] and not w["shortrange"]:
        line["lattice_correction"] = {"kernel": "lattice_kernel", "ms_per_step": (stage["walk_ms"] - stage["walk_kernel_ms"]) / K,
                                      "note": "walk stage minus walk kernel: target compaction + lattice-sum correction walk of all targets"}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if parity is not None:
        line["parity"] = parity
    print(json.dumps(line), file=real_stdout, flush=True)
    grp.close()
    if dist is not None:
        _get_default_store().set("g2gpu_bench_done", "1")
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
