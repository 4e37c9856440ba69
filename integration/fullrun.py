"""integration/fullrun.py — TEST INFRASTRUCTURE: runs the COMPLETE reference program (main.c -> run.c -> accel.c -> gravity_tree,
every source file unmodified) on the shipped example of BASELINE config 1, once as pure CPU code (oracle/_ref/Gadget2_ref_*) and
once with the host shim + libg2gpu in place of gravtree.c / peano.c / the tree entry points (oracle/_ref/Gadget2_g2gpu_*), both
built by integration/Makefile.  The initial conditions are rebuilt from tests/golden (positions, velocities, ids of the example;
/root/reference does not exist on the GPU box) in the reference's IC format 1 (read_ic.c, allvars.h:685-708); the parameter file
carries the values of the reference's shipped configuration for that example.

    python integration/fullrun.py [--time-max 0.05] [--prec f32|f64] [--only ref|g2gpu]      prints one JSON line
"""
import json
import os
import struct
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
GOLD = os.path.join(ROOT, "tests", "golden")

# values of the example's configuration (tree gravity, two species: disk particles are species 1)
PARAMS = dict(
    InitCondFile="ic.dat", OutputDir="./", EnergyFile="energy.txt", InfoFile="info.txt", TimingsFile="timings.txt", CpuFile="cpu.txt",
    RestartFile="restart", SnapshotFileBase="snapshot", OutputListFilename="none.txt", TimeLimitCPU=36000, ResubmitOn=0,
    ResubmitCommand="none", ICFormat=1, SnapFormat=1, ComovingIntegrationOn=0, TypeOfTimestepCriterion=0, OutputListOn=0,
    PeriodicBoundariesOn=0, TimeBegin=0.0, TimeMax=0.05, Omega0=0, OmegaLambda=0, OmegaBaryon=0, HubbleParam=1.0, BoxSize=0,
    TimeBetSnapshot=0.05, TimeOfFirstSnapshot=0, CpuTimeBetRestartFile=36000.0, TimeBetStatistics=0.05, NumFilesPerSnapshot=1,
    NumFilesWrittenInParallel=1, ErrTolIntAccuracy=0.025, CourantFac=0.15, MaxSizeTimestep=0.01, MinSizeTimestep=0.0, ErrTolTheta=0.5,
    TypeOfOpeningCriterion=1, ErrTolForceAcc=0.005, TreeDomainUpdateFrequency=0.1, DesNumNgb=50, MaxNumNgbDeviation=2,
    ArtBulkViscConst=0.8, InitGasTemp=0, MinGasTemp=0, PartAllocFactor=1.5, TreeAllocFactor=8, BufferSize=25,
    UnitLength_in_cm=3.085678e21, UnitMass_in_g=1.989e43, UnitVelocity_in_cm_per_s=1e5, GravityConstantInternal=0,
    GravityGas=0, GravityHalo=0, GravityDisk=1, GravityBulge=0, GravityStars=0, GravityBndry=0, MinGasHsmlFractional=0.25,
    SofteningGas=0, SofteningHalo=1.0, SofteningDisk=0.4, SofteningBulge=1.0, SofteningStars=1.0, SofteningBndry=1.0,
    SofteningGasMaxPhys=0, SofteningHaloMaxPhys=1.0, SofteningDiskMaxPhys=0.4, SofteningBulgeMaxPhys=1.0, SofteningStarsMaxPhys=1.0,
    SofteningBndryMaxPhys=1.0, MaxRMSDisplacementFac=0.2)


def write_param(path, **over):
    p = dict(PARAMS)
    p.update(over)
    with open(path, "w") as f:
        for k, v in p.items():
            f.write(f"{k:32s} {v}\n")


def _block(f, payload):
    f.write(struct.pack("i", len(payload)))
    f.write(payload)
    f.write(struct.pack("i", len(payload)))


def write_ic(path, pos, vel, ids, npart, masstab):
    """IC format 1: Fortran records header(256 B) / POS / VEL / ID; masses from the header's table (all non-zero here)."""
    h = struct.pack("6i", *npart) + struct.pack("6d", *masstab) + struct.pack("2d", 0.0, 0.0) + struct.pack("2i", 0, 0)
    h += struct.pack("6I", *npart) + struct.pack("2i", 0, 1) + struct.pack("4d", 0.0, 0.0, 0.0, 1.0)
    h += b"\0" * (256 - len(h))
    with open(path, "wb") as f:
        _block(f, h)
        _block(f, np.ascontiguousarray(pos, dtype=np.float32).tobytes())
        _block(f, np.ascontiguousarray(vel, dtype=np.float32).tobytes())
        _block(f, np.ascontiguousarray(ids, dtype=np.uint32).tobytes())


def read_snapshot(path):
    with open(path, "rb") as f:
        def block():
            n = struct.unpack("i", f.read(4))[0]
            d = f.read(n)
            assert struct.unpack("i", f.read(4))[0] == n
            return d
        h = block()
        npart = struct.unpack("6i", h[:24])
        t = struct.unpack("d", h[72:80])[0]
        n = sum(npart)
        pos = np.frombuffer(block(), dtype=np.float32).reshape(n, 3)
        vel = np.frombuffer(block(), dtype=np.float32).reshape(n, 3)
        ids = np.frombuffer(block(), dtype=np.uint32)
    o = np.argsort(ids, kind="stable")
    return dict(time=t, pos=pos[o].copy(), vel=vel[o].copy(), ids=ids[o].copy())


def example_ic():
    g = np.load(os.path.join(GOLD, "config1_galaxycollision.npz"))
    v = np.load(os.path.join(GOLD, "config1_fullrun.npz"))
    ptype = g["in_type"]
    npart = [int((ptype == t).sum()) for t in range(6)]
    # the example's own ID block is not unique across its two galaxies; unique ids 1..N (original order) let two runs be matched
    return g["in_pos"], v["in_vel"], np.arange(1, len(ptype) + 1, dtype=np.uint32), npart, [float(x) for x in v["masstab"]]


def run(kind, prec, workdir, time_max=0.05, timeout=3000, env=None, **over):
    """kind = 'ref' (pure CPU reference) or 'g2gpu' (reference + shim on the GPU).  Returns wall time, step count, the reference's own
    cpu.txt totals (CPU_Total, CPU_Gravity, ...), the timings.txt particle rates, energies and the last snapshot."""
    exe = os.path.join(REFDIR, f"Gadget2_{kind}_{prec}")
    if not os.path.exists(exe):
        raise FileNotFoundError(exe)
    os.makedirs(workdir, exist_ok=True)
    pos, vel, ids, npart, masstab = example_ic()
    write_ic(os.path.join(workdir, "ic.dat"), pos, vel, ids, npart, masstab)
    write_param(os.path.join(workdir, "param.txt"), **dict(dict(TimeMax=time_max, TimeBetSnapshot=time_max, TimeBetStatistics=time_max), **over))
    t0 = time.time()
    with open(os.path.join(workdir, "log.txt"), "w") as log:
        rc = subprocess.run([exe, "param.txt"], cwd=workdir, stdout=log, stderr=subprocess.STDOUT, timeout=timeout,
                            env=dict(os.environ, **(env or {}))).returncode
    wall = time.time() - t0
    if rc != 0:
        raise RuntimeError(f"{exe} exited with {rc}: " + open(os.path.join(workdir, "log.txt")).read()[-2000:])
    cpu = [ln.split() for ln in open(os.path.join(workdir, "cpu.txt")) if not ln.startswith("Step")]
    last = [float(x) for x in cpu[-1]]
    steps = sum(1 for ln in open(os.path.join(workdir, "cpu.txt")) if ln.startswith("Step"))
    rates = [float(ln.split("=")[1].split("|")[0]) for ln in open(os.path.join(workdir, "timings.txt")) if ln.startswith("part/sec")]
    ia = [float(ln.split("ia/part=")[1].split()[0]) for ln in open(os.path.join(workdir, "timings.txt")) if ln.startswith("part/sec")]
    energy = np.loadtxt(os.path.join(workdir, "energy.txt"), ndmin=2)
    snaps = sorted(f for f in os.listdir(workdir) if f.startswith("snapshot_"))
    snap = read_snapshot(os.path.join(workdir, snaps[-1]))
    return dict(wall_s=wall, steps=steps, cpu_total_s=last[0], cpu_gravity_s=last[1], cpu_domain_s=last[3], cpu_predict_s=last[5],
                cpu_timeline_s=last[6], cpu_treewalk_s=last[8], cpu_treebuild_s=last[9], part_per_s_median=float(np.median(rates)) if rates else 0.0,
                ia_per_part_mean=float(np.mean(ia)) if ia else 0.0, force_computations=len(rates), energy=energy, snap=snap)


def main():
    import argparse
    import tempfile
    ap = argparse.ArgumentParser()
    ap.add_argument("--time-max", type=float, default=0.05)
    ap.add_argument("--prec", default="f32")
    ap.add_argument("--only", default="")
    ap.add_argument("--tree-update-frequency", type=float, default=0.1, help="TreeDomainUpdateFrequency (0 = new tree at every step)")
    args = ap.parse_args()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        res = {}
        for kind in ("ref", "g2gpu"):
            if args.only and kind != args.only:
                continue
            r = run(kind, args.prec, os.path.join(tmp, kind), args.time_max, TreeDomainUpdateFrequency=args.tree_update_frequency)
            res[kind] = r
            out[kind] = {k: v for k, v in r.items() if k not in ("energy", "snap")}
            out[kind]["kinetic_energy_end"] = float(r["energy"][-1][3])
        if len(res) == 2:
            a, b = res["ref"]["snap"], res["g2gpu"]["snap"]
            d = np.linalg.norm(a["pos"] - b["pos"], axis=1)
            move = np.linalg.norm(a["pos"] - example_ic()[0], axis=1)
            out["compare"] = dict(pos_diff_max=float(d.max()), pos_diff_median=float(np.median(d)), displacement_median=float(np.median(move)),
                                  speedup_wall=res["ref"]["wall_s"] / res["g2gpu"]["wall_s"],
                                  speedup_gravity=res["ref"]["cpu_gravity_s"] / max(res["g2gpu"]["cpu_gravity_s"], 1e-9))
    print(json.dumps(out))


if __name__ == "__main__":
    sys.exit(main())
