"""Shared helpers of the parity tests: synthetic particle sets and oracle/GPU drivers."""
import numpy as np

SOFT_NP = (0.0, 0.05, 0.02, 0.05, 0.05, 0.05)     # Plummer-equivalent softenings per type (types 1 and 2 differ)
GRAV_D2 = (0, 0, 1, 0, 0, 0)                      # Gravity{Gas..Bndry}: type 2 -> species 1 (Configuration.reference)


def hernquist(n, a=30.0, rmax=1000.0, seed=20261018, two_types=True):
    """Two-species Hernquist halo: r = a sqrt(u)/(1-sqrt(u)), isotropic; float32 positions."""
    rng = np.random.default_rng(seed)
    umax = (rmax / (rmax + a)) ** 2
    u = rng.uniform(0, umax, n)
    r = a * np.sqrt(u) / (1 - np.sqrt(u))
    mu = rng.uniform(-1, 1, n)
    phi = rng.uniform(0, 2 * np.pi, n)
    st = np.sqrt(1 - mu * mu)
    pos = np.stack([r * st * np.cos(phi), r * st * np.sin(phi), r * mu], axis=1).astype(np.float32)
    mass = np.full(n, 1.0e-4, dtype=np.float32)
    ptype = np.ones(n, dtype=np.int32)
    if two_types:
        ptype[n // 2:] = 2
    return pos, mass, ptype


def gaussian_blobs(n, seed=3, types=(1, 2)):
    rng = np.random.default_rng(seed)
    centers = rng.uniform(-50, 50, size=(4, 3))
    which = rng.integers(0, 4, n)
    pos = (centers[which] + rng.normal(size=(n, 3)) * rng.uniform(1, 8, size=(4, 1))[which]).astype(np.float32)
    mass = rng.uniform(0.5e-3, 2e-3, n).astype(np.float32)
    ptype = np.asarray(types, dtype=np.int32)[rng.integers(0, len(types), n)]
    return pos, mass, ptype


def periodic_poisson(n, box=100000.0, seed=42, ntypes=2):
    """Periodic Poisson box: species by index parity (types 1/2), unit masses."""
    rng = np.random.default_rng(seed)
    pos = rng.uniform(0, box, size=(n, 3)).astype(np.float32)
    pos = np.minimum(pos, np.float32(np.nextafter(np.float32(box), np.float32(0))))
    mass = np.ones(n, dtype=np.float32)
    if ntypes == 2:
        ptype = np.where(np.arange(n) % 2 == 0, 1, 2).astype(np.int32)
    else:
        ptype = (np.arange(n) % 6).astype(np.int32)
        ptype = np.sort(ptype, kind="stable")          # gas must sit at the head of P[] (peano.c:47-67)
    return pos, mass, ptype


def force_softening(soft):
    return 2.8 * np.asarray(soft, dtype=np.float64)   # gravtree.c:514-515


def rel_err(a, b):
    """per-particle |a-b| / |b|"""
    num = np.linalg.norm(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64), axis=1)
    den = np.linalg.norm(np.asarray(b, dtype=np.float64), axis=1)
    return num / np.maximum(den, 1e-300)


def compare_tree(gt, rt, D):
    """Returns a dict of mismatch counts between a GPU tree mirror and the oracle's tree (same numbering)."""
    out = {}
    out["numnodes"] = int(gt["numnodes"] != rt["numnodes"])
    n = min(gt["numnodes"], rt["numnodes"])
    out["len"] = int(np.sum(gt["len"][:n] != rt["len"][:n].astype(np.float32)))
    out["center"] = int(np.sum(gt["center"][:n] != rt["center"][:n].astype(np.float32)))
    out["mass"] = int(np.sum(gt["mass"][:n] != rt["mass"][:n].astype(np.float32)))
    out["s"] = int(np.sum(gt["s"][:n] != rt["s"][:n].astype(np.float32)))
    for k in ("bitflags", "sibling", "nextnode", "father"):
        out[k] = int(np.sum(gt[k][:n] != rt[k][:n]))
    out["p_nextnode"] = int(np.sum(gt["p_nextnode"] != rt["p_nextnode"]))
    out["p_father"] = int(np.sum(gt["p_father"] != rt["p_father"]))
    return out


def direct_sum(pos, mass, h, targets, chunk=None):
    """FP64 direct summation with the GADGET-2 spline softening (the stock pair law of every BASELINE config: Newtonian for
    r >= h, ngravs.c:420-434 below; h = max of the two particles' 2.8 eps as in forcetree.c:3428-3548); G = 1, pre-G."""
    pos = np.asarray(pos, dtype=np.float64)
    mass = np.asarray(mass, dtype=np.float64)
    h = np.asarray(h, dtype=np.float64)
    out = np.zeros((len(targets), 3))
    chunk = chunk or max(1, int(1.6e7 // len(mass)))       # ~0.4 GB of temporaries per chunk
    for c0 in range(0, len(targets), chunk):
        t = np.asarray(targets[c0:c0 + chunk])
        d = pos[None, :, :] - pos[t][:, None, :]
        r2 = np.einsum("tnk,tnk->tn", d, d)
        r = np.sqrt(r2)
        hh = np.maximum(h[None, :], h[t][:, None])
        u = r / hh
        with np.errstate(divide="ignore", invalid="ignore"):
            newton = 1.0 / (r2 * r)
            inner = (10.666666666667 + u * u * (32.0 * u - 38.4)) / hh ** 3
            outer = (21.333333333333 - 48.0 * u + 38.4 * u * u - 10.666666666667 * u ** 3 - 0.066666666667 / u ** 3) / hh ** 3
        fac = np.where(r >= hh, newton, np.where(u < 0.5, inner, outer))
        fac[r2 == 0.0] = 0.0
        out[c0:c0 + len(t)] = np.einsum("tn,tnk->tk", fac * mass[None, :], d)
    return out


def direct_potential(pos, mass, h, targets):
    """FP64 direct-sum potential with the GADGET-2 spline (newtonian_pot / plummer_pot, ngravs.c:368, 459-471; h = max of the two
    particles' 2.8 eps), self term INCLUDED, pre-G: what the tree potential walks approximate (forcetree.c:2724-2768)."""
    pos = np.asarray(pos, dtype=np.float64)
    mass = np.asarray(mass, dtype=np.float64)
    h = np.asarray(h, dtype=np.float64)
    out = np.zeros(len(targets))
    for k, t in enumerate(targets):
        d = pos - pos[t]
        r = np.sqrt(np.einsum("nk,nk->n", d, d))
        hh = np.maximum(h, h[t])
        u = r / hh
        with np.errstate(divide="ignore", invalid="ignore"):
            newton = -1.0 / r
            inner = (-2.8 + u * u * (5.333333333333 + u * u * (6.4 * u - 9.6))) / hh
            outer = (-3.2 + 0.066666666667 / u + u * u * (10.666666666667 + u * (-16.0 + u * (9.6 - 2.133333333333 * u)))) / hh
        out[k] = np.sum(mass * np.where(r >= hh, newton, np.where(u < 0.5, inner, outer)))
    return out


def ewald_direct(pos, mass, targets, box, alpha_l=2.0, nmax=3, hmax2=10):
    """Exact periodic Newtonian acceleration (G = 1, point masses) of the particles `targets` by Ewald summation, FP64: the ground
    truth the reference's FORCETEST uses for periodic boxes (direct sum + lattice correction, forcetree.c:3428-3548 with
    ewald_force, ngravs.c:1170).  alpha = alpha_l / box; real-space images |n| <= nmax, reciprocal vectors |h|^2 < hmax2."""
    from math import erfc  # noqa: F401  (scipy.special.erfc is used vectorised below)
    from scipy.special import erfc as verfc
    pos = np.asarray(pos, dtype=np.float64)
    mass = np.asarray(mass, dtype=np.float64)
    alpha = alpha_l / box
    rng = np.arange(-nmax, nmax + 1)
    images = np.array([(a, b, c) for a in rng for b in rng for c in rng], dtype=np.float64) * box
    hs = np.array([(a, b, c) for a in rng for b in rng for c in rng if 0 < a * a + b * b + c * c < hmax2], dtype=np.float64)
    ks = 2 * np.pi * hs / box
    k2 = (ks * ks).sum(axis=1)
    kfac = 4 * np.pi / box ** 3 * np.exp(-k2 / (4 * alpha * alpha)) / k2
    out = np.zeros((len(targets), 3))
    for ti, t in enumerate(targets):
        x = pos[t] - pos                       # target minus source
        x -= box * np.rint(x / box)
        keep = np.arange(len(mass)) != t
        x, m = x[keep], mass[keep]
        acc = np.zeros(3)
        for c0 in range(0, len(m), 4096):
            xc, mc = x[c0:c0 + 4096], m[c0:c0 + 4096]
            d = xc[:, None, :] - images[None, :, :]
            r = np.sqrt((d * d).sum(axis=2))
            f = (verfc(alpha * r) + 2 * alpha * r / np.sqrt(np.pi) * np.exp(-(alpha * r) ** 2)) / r ** 3
            acc -= np.einsum("s,sn,snk->k", mc, f, d)
            ph = xc @ ks.T
            acc -= np.einsum("s,sh,hk->k", mc, np.sin(ph) * kfac[None, :], ks)
        out[ti] = acc
    return out
