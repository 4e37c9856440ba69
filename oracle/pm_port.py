"""oracle/pm_port.py — TEST INFRASTRUCTURE ONLY: numpy restatement of the reference's periodic PM long-range force
(pmforce_periodic, pm_periodic.c:204-790) for a single rank.  PINNED: tests/test_pm_oracle.py checks it against the
unmodified reference (oracle/_ref, driven through the rfftwnd_mpi stand-in of oracle/ref/stubs.c) and against
tests/golden/pm_*.npz generated from that reference build (tests/golden/make_golden_pm.py).

Per ordered species pair (nA sources, nB targets), as the reference does (pm_periodic.c:268-786):
  CIC mass assignment of species nA (:285-316, float position scaling `to_slab_fac * Pos`, weights in double)
  -> forward FFT -> multiply by GreensFxns[nA][nB](k) * (-exp(-k^2 asmth2)) / (sinc_x sinc_y sinc_z)^4 (:468-523), k = 0 mode
  zeroed (:525-526) -> inverse FFT -> 4-point finite difference (:726-737) scaled by fac = G/(pi L) / (2 L/PMGRID) (:236-237)
  -> CIC interpolation to the particles of species nB (:739-781), GravPM (FLOAT) += acc (double).
"""
import numpy as np


def pgdelta(k2):            # ngravs.c:390-397  (Newton: 4 pi G / k^2 -> 1/k2 in mesh units)
    with np.errstate(divide="ignore"):
        return 1.0 / k2


def pgyukawa(k2, ym_grid):  # ngravs.c:869-878 style: 1/(k2 + m^2) in mesh units
    return 1.0 / (k2 + ym_grid * ym_grid)


def cic_index(pos, pmgrid, boxsize, float_bytes=4):
    """slab index and in-cell offset exactly as pm_periodic.c:286-310 (to_slab_fac is a FLOAT, the product is a FLOAT)."""
    ft = np.float32 if float_bytes == 4 else np.float64
    fac = ft(pmgrid / boxsize)
    u = (fac * pos.astype(ft)).astype(ft)
    slab = u.astype(np.int64)
    slab = np.minimum(slab, pmgrid - 1)
    d = (u - slab.astype(ft)).astype(np.float64)
    return slab, d


def pm_force(pos, mass, species, D, pmgrid, boxsize, G, asmth, greens=None, float_bytes=4):
    """pos[n,3], mass[n] (FLOAT values), species[n] in 0..D-1.  Returns GravPM[n,3] as the reference stores it (FLOAT)."""
    ft = np.float32 if float_bytes == 4 else np.float64
    n = len(mass)
    N = pmgrid
    slab, d = cic_index(pos, N, boxsize, float_bytes)
    m = mass.astype(np.float64)
    asmth2 = ((2 * np.pi) * asmth / boxsize) ** 2
    fac = G / (np.pi * boxsize) * (1 / (2 * boxsize / N))
    kk = np.fft.fftfreq(N, 1.0 / N)
    kx, ky, kz = np.meshgrid(kk, kk, kk[:N // 2 + 1].copy(), indexing="ij")
    kz = np.abs(kz)
    kz[:, :, N // 2] = N // 2
    kx = np.where(kx == -N // 2, N // 2, kx)      # the reference maps index N/2 to +N/2 (x > PMGRID/2 ? x - PMGRID : x)
    ky = np.where(ky == -N // 2, N // 2, ky)
    k2 = kx * kx + ky * ky + kz * kz

    def sinc(k):
        a = np.pi * k / N
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.where(k != 0, np.sin(a) / a, 1.0)
    ff = 1.0 / (sinc(kx) * sinc(ky) * sinc(kz))
    base = -np.exp(-k2 * asmth2) * ff ** 4
    w = [(1.0 - d, d)]
    out = np.zeros((n, 3), dtype=ft)
    corners = [(a, b, c) for a in (0, 1) for b in (0, 1) for c in (0, 1)]
    for nA in range(D):
        selA = species == nA
        rho = np.zeros((N, N, N))
        for (a, b, c) in corners:
            wx = d[selA, 0] if a else 1.0 - d[selA, 0]
            wy = d[selA, 1] if b else 1.0 - d[selA, 1]
            wz = d[selA, 2] if c else 1.0 - d[selA, 2]
            np.add.at(rho, ((slab[selA, 0] + a) % N, (slab[selA, 1] + b) % N, (slab[selA, 2] + c) % N), m[selA] * wx * wy * wz)
        rk = np.fft.rfftn(rho)
        for nB in range(D):
            selB = np.nonzero(species == nB)[0]
            if len(selB) == 0:
                continue
            g = pgdelta(k2) if greens is None else greens(nA, nB, k2)
            smth = np.where(k2 > 0, g * base, 0.0)
            phi = np.fft.irfftn(rk * smth, s=(N, N, N), axes=(0, 1, 2)) * N ** 3        # FFTW-2 is unnormalised in both directions
            for dim in range(3):
                f = fac * ((4.0 / 3) * (np.roll(phi, 1, dim) - np.roll(phi, -1, dim)) - (1.0 / 6) * (np.roll(phi, 2, dim) - np.roll(phi, -2, dim)))
                acc = np.zeros(len(selB))
                for (a, b, c) in [(0, 0, 0), (0, 1, 0), (0, 0, 1), (0, 1, 1), (1, 0, 0), (1, 1, 0), (1, 0, 1), (1, 1, 1)]:   # order of :770-778
                    wx = d[selB, 0] if a else 1.0 - d[selB, 0]
                    wy = d[selB, 1] if b else 1.0 - d[selB, 1]
                    wz = d[selB, 2] if c else 1.0 - d[selB, 2]
                    acc += f[(slab[selB, 0] + a) % N, (slab[selB, 1] + b) % N, (slab[selB, 2] + c) % N] * wx * wy * wz
                out[selB, dim] = (out[selB, dim].astype(np.float64) + acc).astype(ft)
    return out


def pm_potential(pos, mass, species, D, pmgrid, boxsize, G, asmth, greens, float_bytes=4):
    """pmpotential_periodic (pm_periodic.c:798-1290) restated for a single rank: what the routine ADDS to P[].Potential (FLOAT).

    Per ordered pair (nA sources, nB targets): CIC assignment of species nA (:886-916) -> forward FFT (:1010) -> multiply by
    -exp(-k^2 asmth2) * GreensFxns[nA][nB](k) * fac / (sinc sinc sinc)^4 with fac = G / (pi L) (:832, 1014-1056); the k = 0 mode is KEPT
    (:1033-1035) and only reset when its real part is NaN (:1060-1063) -> inverse FFT (:1068) -> CIC interpolation of the potential
    itself to the particles of species nB, eight FLOAT `+=` per source species (:1254-1267).
    greens(nA, nB, k2) must be finite at k2 = 0 for a finite result (the stock 1/k^2 is not: the reference then returns NaN/inf)."""
    ft = np.float32 if float_bytes == 4 else np.float64
    n = len(mass)
    N = pmgrid
    slab, d = cic_index(pos, N, boxsize, float_bytes)
    m = mass.astype(np.float64)
    asmth2 = ((2 * np.pi) * asmth / boxsize) ** 2
    fac = G / (np.pi * boxsize)
    kk = np.fft.fftfreq(N, 1.0 / N)
    kx, ky, kz = np.meshgrid(kk, kk, kk[:N // 2 + 1].copy(), indexing="ij")
    kz = np.abs(kz)
    kz[:, :, N // 2] = N // 2
    kx = np.where(kx == -N // 2, N // 2, kx)
    ky = np.where(ky == -N // 2, N // 2, ky)
    k2 = kx * kx + ky * ky + kz * kz

    def sinc(k):
        a = np.pi * k / N
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.where(k != 0, np.sin(a) / a, 1.0)
    ff = 1.0 / (sinc(kx) * sinc(ky) * sinc(kz))
    out = np.zeros(n, dtype=ft)
    for nA in range(D):
        selA = species == nA
        rho = np.zeros((N, N, N))
        for (a, b, c) in [(a, b, c) for a in (0, 1) for b in (0, 1) for c in (0, 1)]:
            wx = d[selA, 0] if a else 1.0 - d[selA, 0]
            wy = d[selA, 1] if b else 1.0 - d[selA, 1]
            wz = d[selA, 2] if c else 1.0 - d[selA, 2]
            np.add.at(rho, ((slab[selA, 0] + a) % N, (slab[selA, 1] + b) % N, (slab[selA, 2] + c) % N), m[selA] * wx * wy * wz)
        rk = np.fft.rfftn(rho)
        for nB in range(D):
            selB = np.nonzero(species == nB)[0]
            if len(selB) == 0:
                continue
            with np.errstate(invalid="ignore", divide="ignore"):
                smth = -np.exp(-k2 * asmth2) * greens(nA, nB, k2) * fac * ff ** 4
                pk = rk * smth
            if np.isnan(pk[0, 0, 0].real):
                pk[0, 0, 0] = 0.0
            phi = np.fft.irfftn(pk, s=(N, N, N), axes=(0, 1, 2)) * N ** 3        # FFTW-2 is unnormalised in both directions
            acc = out[selB]
            for (a, b, c) in [(0, 0, 0), (0, 1, 0), (0, 0, 1), (0, 1, 1), (1, 0, 0), (1, 1, 0), (1, 0, 1), (1, 1, 1)]:   # order of :1254-1267
                wx = d[selB, 0] if a else 1.0 - d[selB, 0]
                wy = d[selB, 1] if b else 1.0 - d[selB, 1]
                wz = d[selB, 2] if c else 1.0 - d[selB, 2]
                acc = (acc.astype(np.float64) + phi[(slab[selB, 0] + a) % N, (slab[selB, 1] + b) % N, (slab[selB, 2] + c) % N] * wx * wy * wz).astype(ft)
            out[selB] = acc
    return out
