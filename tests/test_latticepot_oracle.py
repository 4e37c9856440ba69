"""CPU test (no GPU) of the periodic non-PM potential oracle: lattice_pot_corr (forcetree.c:3895-3941) and the table it
interpolates, potcorr of lattice_init (forcetree.c:3697-3702, 3759) = ewald_psi (ngravs.c:761-816) for the stock wiring.

What CAN be pinned against the unmodified reference (variant per_d2_f32, -DPERIODIC without PMGRID): the table of pair [0][0] and the
look-up, point by point.  What cannot: force_treeevaluate_potential (forcetree.c:2467-2776), the only caller, does not compile in the
reference (oracle/ref/Makefile), so the walk with the lattice term is PARITY UNPINNED and checked against an independent FP64 Ewald sum of
the periodic potential instead."""
import numpy as np
import pytest

import g2test
from portrun import PortOracle, make_ewald_pot_table
from refrun import RefOracle, available


def ewald_potential(pos, mass, targets, box, alpha_l=2.0, nmax=3, hmax2=12):
    """Exact periodic Newtonian potential (G = 1, point masses, neutralising background, self term excluded; sign: phi = -sum m/r ...)
    at the particles `targets`, FP64 Ewald sum."""
    from scipy.special import erfc
    pos, mass = np.asarray(pos, dtype=np.float64), np.asarray(mass, dtype=np.float64)
    alpha = alpha_l / box
    rng = np.arange(-nmax, nmax + 1)
    images = np.array([(a, b, c) for a in rng for b in rng for c in rng], dtype=np.float64) * box
    hs = np.array([(a, b, c) for a in rng for b in rng for c in rng if 0 < a * a + b * b + c * c < hmax2], dtype=np.float64)
    ks = 2 * np.pi * hs / box
    k2 = (ks * ks).sum(axis=1)
    kfac = 4 * np.pi / box ** 3 * np.exp(-k2 / (4 * alpha * alpha)) / k2
    out = np.zeros(len(targets))
    for ti, t in enumerate(targets):
        x = pos[t] - pos
        x -= box * np.rint(x / box)
        keep = np.arange(len(mass)) != t
        xs, ms = x[keep], mass[keep]
        d = xs[:, None, :] - images[None, :, :]
        r = np.sqrt((d * d).sum(axis=2))
        phi = -(ms[:, None] * erfc(alpha * r) / r).sum()
        phi -= (ms[:, None] * np.cos(xs @ ks.T) * kfac[None, :]).sum()
        phi += np.pi / (alpha * alpha * box ** 3) * ms.sum()
        out[ti] = phi
    return out


@pytest.mark.skipif(not available("per_d2_f32"), reason="oracle/_ref not built (needs /root/reference)")
def test_port_potential_table_and_lookup_match_reference_build():
    n, box = 500, 1000.0
    ref = RefOracle("per_d2_f32", int(1.1 * n) + 64, boxsize=box, softening=(1.0,) * 6, gravity=g2test.GRAV_D2)
    tabs = ref.potcorr_tables()
    assert tabs is not None and np.abs(tabs[0, 0]).max() > 0
    port = make_ewald_pot_table(64) / box                       # forcetree.c:3759
    assert np.array_equal(port, tabs[0, 0])                      # same sums in the same order: bit-identical
    o = PortOracle(int(1.1 * n) + 64, D=2, periodic=True, shortrange=False, unequal=False, boxsize=box, softening=(1.0,) * 6, gravity=g2test.GRAV_D2)
    o.set_lattice_pot_tables(tabs)
    rng = np.random.default_rng(8)
    for _ in range(2000):
        d = rng.uniform(-0.5 * box, 0.5 * box, 3)
        for t, s in ((0, 0), (1, 0), (0, 1)):
            assert o.lattice_pot_corr(d[0], d[1], d[2], t, s) == ref.lattice_pot_corr(d[0], d[1], d[2], t, s)
    for d in ([0.0, 0.0, 0.0], [0.5 * box, -0.5 * box, 0.5 * box], [0.5 * box, 0.0, 1e-9]):          # grid corners and the clamp at EN
        assert o.lattice_pot_corr(d[0], d[1], d[2], 0, 0) == ref.lattice_pot_corr(d[0], d[1], d[2], 0, 0)


def test_port_periodic_potential_agrees_with_ewald_sum():
    """The port's walk (nearest-image tree potential + mass * lattice_pot_corr per term) against the exact periodic potential.  With the
    opening angle of a direct sum (theta tiny) only the table interpolation separates the two."""
    n, box = 600, 1000.0
    pos, mass, ptype = g2test.periodic_poisson(n, box, seed=12)
    mass = (mass * np.random.default_rng(4).uniform(0.5, 2.0, n)).astype(np.float32)
    soft = (1.0e-3,) * 6
    tab = make_ewald_pot_table(64) / box
    o = PortOracle(int(1.1 * n) + 64, D=2, periodic=True, shortrange=False, unequal=False, boxsize=box, theta=1.0e-3, softening=soft, gravity=g2test.GRAV_D2)
    o.set_potential_laws("newtonian", "plummer")
    o.set_lattice_pot_tables(np.broadcast_to(tab, (2, 2, 65, 65, 65)).copy())
    o.load(pos, mass, ptype)
    o.domain()
    o.treebuild()
    pot = o.potential(nthreads=4).astype(np.float64)
    p = o.particles()
    # remove the self term like potential.c:250-254 (the walk meets the target itself at r = 0: spline value -2.8 m/h, plus
    # mass * potcorr(0) = m * LatticeZero / L)
    h = 2.8 * soft[1]
    pot_noself = pot + p["mass"] * 2.8 / h - p["mass"] * float(np.float32(2.8372975)) / box
    targets = np.arange(0, n, 25)
    exact = ewald_potential(p["pos"], p["mass"], targets, box)
    scale = np.abs(exact).max()
    assert np.abs(pot_noself[targets] - exact).max() <= 2e-3 * scale, np.abs(pot_noself[targets] - exact).max() / scale
    # without the tables the same walk is the nearest-image potential, which is off by the missing images (tens of per cent)
    o2 = PortOracle(int(1.1 * n) + 64, D=2, periodic=True, shortrange=False, unequal=False, boxsize=box, theta=1.0e-3, softening=soft, gravity=g2test.GRAV_D2)
    o2.set_potential_laws("newtonian", "plummer")
    o2.load(pos, mass, ptype)
    assert o2.L.g2o_potential(o2.h, 1, np.zeros(n, dtype=np.float32).ctypes.data_as(__import__("ctypes").c_void_p)) != 0     # refused: needs the tables
