"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built from /root/reference by
oracle/ref/Makefile).  Run in the build container only:  python tests/golden/make_golden.py

Each fixture holds float32 inputs and what the reference computed from them: Peano-Hilbert keys, the particle order
after peano_hilbert_order, TopNodes, the complete tree (node records, Nextnode, Father), accelerations/GravCost/OldAcc of
a Barnes-Hut pass (first force computation of a run) and of a relative-criterion pass, and the short-range table."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import g2test  # noqa: E402
from refrun import RefOracle  # noqa: E402


def run(variant, pos, mass, ptype, soft, grav, box=0.0):
    n = len(mass)
    ref = RefOracle(variant, int(1.1 * n) + 64, boxsize=box, softening=soft, gravity=grav)
    ref.load(pos, mass, ptype)
    ref.domain()
    out = dict(in_pos=pos, in_mass=mass, in_type=ptype, soft=np.asarray(soft), grav=np.asarray(grav), box=box,
               maxpart=ref.maxpart, D=ref.D)
    p0 = ref.particles()
    out["order_id"] = p0["id"].astype(np.int32)
    out["keys"] = ref.keys()
    d = ref.domain_info()
    out["domain"] = np.concatenate([d["corner"], d["center"], [d["len"], d["fac"]]])
    ref.gravity()
    t = ref.topnodes()
    for k in ("daughter", "leaf", "size", "startkey", "count", "domain_node_index"):
        out["top_" + k] = t[k]
    tr = ref.tree()
    for k in ("len", "center", "s", "mass"):
        out["tree_" + k] = tr[k].astype(np.float32)
    for k in ("bitflags", "sibling", "nextnode", "father", "p_nextnode", "p_father"):
        out["tree_" + k] = tr[k]
    p1 = ref.particles()
    out["bh_acc"], out["bh_cost"], out["bh_oldacc"] = p1["acc"].astype(np.float32), p1["cost"], p1["oldacc"].astype(np.float32)
    ref.set_opening(0.0, 0.005, 1)
    ref.gravity()
    p2 = ref.particles()
    out["rel_acc"], out["rel_cost"], out["rel_oldacc"] = p2["acc"].astype(np.float32), p2["cost"], p2["oldacc"].astype(np.float32)
    if ref.pmgrid:
        out["srtable"] = ref.srtable()
        out["asmth"], out["rcut"] = ref.pm_split()
        out["pmgrid"] = ref.pmgrid
    return out


def galaxy_collision():
    """BASELINE config 1: the reference's shipped example GalaxyCollision.IC (IC format 1, read_ic.c) with
    Configuration.reference (theta 0.5, relative criterion alpha 0.005, eps_halo 1.0, eps_disk 0.4, type 2 -> species 1,
    internal G from the unit system).  Tree arrays are stored as SHA-256 digests (they must match bit-for-bit anyway)."""
    import hashlib
    import struct
    f = open("/root/reference/GalaxyCollision.IC", "rb")

    def block():
        n = struct.unpack("i", f.read(4))[0]
        d = f.read(n)
        assert struct.unpack("i", f.read(4))[0] == n
        return d
    h = block()
    npart = struct.unpack("6i", h[:24])
    mtab = struct.unpack("6d", h[24:72])
    n = sum(npart)
    pos = np.frombuffer(block(), dtype=np.float32).reshape(n, 3).copy()
    ptype = np.repeat(np.arange(6), npart).astype(np.int32)
    mass = np.asarray(mtab, dtype=np.float32)[ptype]
    soft = (0.0, 1.0, 0.4, 1.0, 1.0, 1.0)
    grav = (0, 0, 1, 0, 0, 0)
    G = 6.672e-8 / (3.085678e21) ** 3 * 1.989e43 * (3.085678e21 / 1e5) ** 2       # begrun.c:97-103
    ref = RefOracle("np_d2_f32", int(1.1 * n) + 64, G=G, softening=soft, gravity=grav)
    ref.load(pos, mass, ptype)
    ref.domain()
    out = dict(in_pos=pos, in_mass=mass, in_type=ptype, soft=np.asarray(soft), grav=np.asarray(grav), box=0.0, maxpart=ref.maxpart, D=2, G=G)
    out["order_id"] = ref.particles()["id"].astype(np.int32)
    out["keys_sha"] = hashlib.sha256(ref.keys().tobytes()).hexdigest()
    ref.gravity()
    tr = ref.tree()
    out["numnodes"] = tr["numnodes"]
    for k in ("len", "center", "s", "mass"):
        out["sha_" + k] = hashlib.sha256(tr[k].astype(np.float32).tobytes()).hexdigest()
    for k in ("bitflags", "sibling", "nextnode", "father", "p_nextnode", "p_father"):
        out["sha_" + k] = hashlib.sha256(tr[k].astype(np.int32).tobytes()).hexdigest()
    p1 = ref.particles()
    out["bh_acc"], out["bh_cost"], out["bh_oldacc"] = p1["acc"].astype(np.float32), p1["cost"].astype(np.uint16), p1["oldacc"].astype(np.float32)
    ref.set_opening(0.0, 0.005, 1)
    ref.gravity()
    p2 = ref.particles()
    out["rel_acc"], out["rel_cost"] = p2["acc"].astype(np.float32), p2["cost"].astype(np.uint16)
    return out


def main():
    np.savez_compressed(os.path.join(HERE, "config1_galaxycollision.npz"), **galaxy_collision())
    pos, mass, ptype = g2test.hernquist(3000)
    np.savez_compressed(os.path.join(HERE, "np_d2_hernquist3000.npz"), **run("np_d2_f32", pos, mass, ptype, g2test.SOFT_NP, g2test.GRAV_D2))
    pos, mass, ptype = g2test.gaussian_blobs(2500, types=(1, 2, 3, 4, 5))
    np.savez_compressed(os.path.join(HERE, "np_d2_blobs2500.npz"), **run("np_d2_f32", pos, mass, ptype, (0.0, 0.05, 0.02, 0.03, 0.05, 0.01), (0, 0, 1, 0, 1, 1)))
    box = 100000.0
    pos, mass, ptype = g2test.periodic_poisson(4096, box)
    eps = box / 16 / 30.0
    np.savez_compressed(os.path.join(HERE, "pm64_d2_poisson4096.npz"), **run("pm64_d2_f32", pos, mass, ptype, (eps,) * 6, g2test.GRAV_D2, box=box))
    pos, mass, ptype = g2test.periodic_poisson(4096, box, seed=7, ntypes=6)
    np.savez_compressed(os.path.join(HERE, "pm64_d4_poisson4096.npz"), **run("pm64_d4_f32", pos, mass, ptype, (eps,) * 6, (0, 1, 2, 3, 1, 2), box=box))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
