"""ctypes driver for oracle/libg2oracle.so (the CPU restatement, oracle/g2_oracle.c) — TEST INFRASTRUCTURE ONLY.
Mirrors oracle/refrun.RefOracle so that tests can run either oracle through the same calls."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libg2oracle.so")

LAW = dict(none=0, newtonian=1, neg_newtonian=2, yukawa=3, coloyuk=4, bambam=5, sourcebambaryon=6, sourcebaryonbam=7)
SPLINE = dict(none=16, plummer=17, neg_plummer=18, bambam_spline=19, sourcebambaryon_spline=20, sourcebaryonbam_spline=21)

POT = dict(none=32, newtonian=33, neg_newtonian=34)
POTSPLINE = dict(none=48, plummer=49, neg_plummer=50)

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise FileNotFoundError(f"{LIB} not built (make -C oracle)")
        L = C.CDLL(LIB)
        L.g2o_create.restype = C.c_void_p
        L.g2o_peano_key.restype = C.c_longlong
        L.g2o_walk_range.restype = C.c_double
        L.g2o_gravity_tree.restype = C.c_double
        L.g2o_accel.restype = C.c_double
        L.g2o_spline.restype = C.c_double
        L.g2o_accel.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_long]
        L.g2o_spline.argtypes = L.g2o_accel.argtypes
        L.g2o_set_params.argtypes = [C.c_void_p] + [C.c_double] * 6
        L.g2o_make_srtable.argtypes = [C.c_int, C.c_int, C.c_double, C.c_void_p]
        L.g2o_make_srtables.argtypes = [C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        L.g2o_potfxn.restype = C.c_double
        L.g2o_potspline.restype = C.c_double
        L.g2o_potfxn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        L.g2o_potspline.argtypes = L.g2o_potfxn.argtypes
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_srtable(ntab=2048, kind=0, ym=0.0):
    """shortrange_fourier_force for one pair law (kind 0 Newtonian, 1 Yukawa); ~1 s."""
    t = np.zeros(ntab)
    rc = lib().g2o_make_srtable(ntab, kind, ym, _p(t))
    assert rc == 0
    return t


def make_srtables(ntab=2048, kind=0, ym=0.0):
    """(shortrange_fourier_force, shortrange_fourier_pot) of one pair law (forcetree.c:3335-3353)."""
    f, p = np.zeros(ntab), np.zeros(ntab)
    rc = lib().g2o_make_srtables(ntab, kind, ym, _p(f), _p(p))
    assert rc == 0
    return f, p


def make_ewald_table(en=64, nthreads=8):
    """ewald_force (ngravs.c:1170) on the (en+1)^3 grid of lattice_init, dimensionless: (3, en+1, en+1, en+1); ~5 s on 8 threads."""
    out = np.zeros((3, en + 1, en + 1, en + 1))
    rc = lib().g2o_make_ewald_table(int(en), int(nthreads), _p(out))
    assert rc == 0
    return out


def make_ewald_pot_table(en=64, nthreads=8, latticezero=float(np.float32(2.8372975))):
    """potcorr of the stock wiring before the division by BoxSize (forcetree.c:3697-3702): ewald_psi (ngravs.c:761) at x = 0.5 (i,j,k)/en,
    LatticeZero (a FLOAT; ngravs.c:133) at the origin; shape (en+1, en+1, en+1)."""
    out = np.zeros((en + 1, en + 1, en + 1))
    rc = lib().g2o_make_ewald_pot_table(int(en), int(nthreads), C.c_double(latticezero), _p(out))
    assert rc == 0
    return out


class PortOracle:
    def __init__(self, maxpart, D=2, periodic=False, shortrange=False, unequal=True, ntab=2048, boxsize=0.0, G=1.0, theta=0.5,
                 errtol=0.005, softening=(0.0, 1.0, 1.0, 1.0, 1.0, 1.0), gravity=(0, 0, 1, 0, 0, 0), tree_alloc=1.5, pmgrid=0,
                 asmth_cells=1.25, rcut_cells=4.5):
        self.L = lib()
        self.D, self.periodic, self.shortrange, self.unequal, self.ntab = D, periodic, shortrange, unequal, ntab
        self.maxpart = int(maxpart)
        self.h = C.c_void_p(self.L.g2o_create(D, int(periodic), int(shortrange), int(unequal), ntab, self.maxpart, int(tree_alloc * maxpart)))
        t2g = np.ascontiguousarray(gravity, dtype=np.int32)
        fs = 2.8 * np.ascontiguousarray(softening, dtype=np.float64)         # set_softenings, gravtree.c:514
        self.fsoft = fs
        self.L.g2o_set_species(self.h, _p(t2g), _p(fs))
        self.boxsize, self.G, self.theta, self.errtol = boxsize, G, theta, errtol
        self.asmth = asmth_cells * boxsize / pmgrid if pmgrid else 0.0    # pm_periodic.c:59-60
        self.rcut = rcut_cells * self.asmth
        self._push()
        self.n = 0

    def _push(self):
        self.L.g2o_set_params(self.h, self.boxsize, self.G, self.theta, self.errtol, self.asmth, self.rcut)

    def close(self):
        if self.h:
            self.L.g2o_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_laws(self, accel="newtonian", spline="plummer", params=None):
        D = self.D

        def grid(x, table):
            if isinstance(x, str):
                return np.full((D, D), table[x], dtype=np.int32)
            return np.array([[table[v] for v in row] for row in x], dtype=np.int32)
        a, s = grid(accel, LAW), grid(spline, SPLINE)
        par = np.zeros((D, D, 4)) if params is None else np.ascontiguousarray(params, dtype=np.float64)
        self.L.g2o_set_laws(self.h, _p(a), _p(s), _p(par))

    def set_srtable(self, table):
        t = np.ascontiguousarray(table, dtype=np.float64)
        assert t.shape == (self.D, self.D, self.ntab)
        self.L.g2o_set_srtable(self.h, _p(t))

    def set_lattice_tables(self, fcorr, en=64):
        """fcorrx/y/z after lattice_init, shape (3, D, D, en+1, en+1, en+1) (PERIODIC without PMGRID): gravity() then runs the
        lattice-correction walk after the tree walk of every target (forcetree.c:1606-1608)."""
        t = np.ascontiguousarray(fcorr, dtype=np.float64)
        assert t.shape == (3, self.D, self.D, en + 1, en + 1, en + 1)
        self.L.g2o_set_lattice_tables(self.h, int(en), _p(t))

    def set_lattice_pot_tables(self, potcorr, en=64):
        """potcorr[tgt][src] after lattice_init (divided by BoxSize), shape (D, D, en+1, en+1, en+1): potential() of a periodic box without
        PM then adds mass * lattice_pot_corr per term (forcetree.c:2736-2738, 2765-2767)."""
        t = np.ascontiguousarray(potcorr, dtype=np.float64)
        assert t.shape == (self.D, self.D, en + 1, en + 1, en + 1)
        self.L.g2o_set_lattice_pot_tables(self.h, int(en), _p(t))

    def lattice_pot_corr(self, dx, dy, dz, tgt, src):
        self.L.g2o_lattice_pot_corr.restype = C.c_double
        self.L.g2o_lattice_pot_corr.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
        return self.L.g2o_lattice_pot_corr(self.h, dx, dy, dz, tgt, src)

    def set_potential_laws(self, pot="newtonian", spline="plummer", node_table_term=False):
        """PotentialFxns / PotentialSplines [target][source]; node_table_term = a -DNGRAVS_ACCUMULATOR build (forcetree.c:3134-3140)."""
        D = self.D

        def grid(x, table):
            if isinstance(x, str):
                return np.full((D, D), table[x], dtype=np.int32)
            return np.array([[table[v] for v in row] for row in x], dtype=np.int32)
        a, s = grid(pot, POT), grid(spline, POTSPLINE)
        self.L.g2o_set_potlaws(self.h, _p(a), _p(s), int(node_table_term))

    def set_srpot_table(self, table):
        t = np.ascontiguousarray(table, dtype=np.float64)
        assert t.shape == (self.D, self.D, self.ntab)
        self.L.g2o_set_srpot(self.h, _p(t))

    def potential(self, nthreads=8):
        """P[].Potential of every particle (current order) as the potential walk leaves it (pre-G, self term included)."""
        out = np.zeros(self.n, dtype=np.float32)
        rc = self.L.g2o_potential(self.h, int(nthreads), _p(out))
        if rc != 0:
            raise RuntimeError(f"g2o_potential failed ({rc})")
        return out

    def potential_targets(self, targets):
        t = np.ascontiguousarray(targets, dtype=np.int32)
        out = np.zeros(len(t), dtype=np.float32)
        rc = self.L.g2o_potential_targets(self.h, len(t), _p(t), _p(out))
        if rc != 0:
            raise RuntimeError(f"g2o_potential_targets failed ({rc})")
        return out

    def potfxn(self, tgt, src, m, h, r):
        return self.L.g2o_potfxn(self.h, tgt, src, m, h, r)

    def potspline(self, tgt, src, m, h, r):
        return self.L.g2o_potspline(self.h, tgt, src, m, h, r)

    def set_opening(self, theta, errtol, criterion=1):
        self.theta, self.errtol = theta, errtol
        self._push()

    def load(self, pos, mass, ptype, oldacc=None, active=None):
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        mass = np.ascontiguousarray(mass, dtype=np.float32)
        ptype = np.ascontiguousarray(ptype, dtype=np.int32)
        old = None if oldacc is None else np.ascontiguousarray(oldacc, dtype=np.float32)
        act = None if active is None else np.ascontiguousarray(active, dtype=np.int32)
        n = len(mass)
        rc = self.L.g2o_load(self.h, n, _p(pos), _p(mass), _p(ptype), _p(old), _p(act))
        assert rc == 0
        self.n = n

    def domain(self):
        self.L.g2o_domain(self.h)

    def treebuild(self):
        return self.L.g2o_treebuild(self.h)

    def gravity(self, nthreads=1):
        """gravity_tree(): build + walk + epilogue; returns the interaction count."""
        return self.L.g2o_gravity_tree(self.h, int(nthreads))

    def walk_threads(self, nthreads, lo=0, hi=None):
        cost = C.c_double(0)
        hi = self.n if hi is None else hi
        dt = self.L.g2o_walk_range(self.h, int(lo), int(hi), int(nthreads), C.byref(cost))
        return dt, cost.value

    def peano_key(self, x, y, z, bits):
        return int(self.L.g2o_peano_key(int(x), int(y), int(z), int(bits)))

    def counts(self):
        c = np.zeros(4, dtype=np.int32)
        self.L.g2o_counts(self.h, _p(c))
        return dict(n=int(c[0]), ntop=int(c[1]), ntopleaves=int(c[2]), numnodes=int(c[3]))

    def domain_info(self):
        d = np.zeros(8)
        self.L.g2o_get_domain(self.h, _p(d))
        return dict(corner=d[0:3].copy(), center=d[3:6].copy(), len=d[6], fac=d[7])

    def particles(self):
        n = self.n
        pos = np.zeros((n, 3), dtype=np.float32)
        mass = np.zeros(n, dtype=np.float32)
        ptype = np.zeros(n, dtype=np.int32)
        pid = np.zeros(n, dtype=np.int32)
        key = np.zeros(n, dtype=np.int64)
        acc = np.zeros((n, 3), dtype=np.float32)
        cost = np.zeros(n, dtype=np.float32)
        old = np.zeros(n, dtype=np.float32)
        accd = np.zeros((n, 3))
        self.L.g2o_get_particles(self.h, *(_p(a) for a in (pos, mass, ptype, pid, key, acc, cost, old, accd)))
        return dict(pos=pos, mass=mass, type=ptype, id=pid, key=key, acc=acc, cost=cost, oldacc=old, accd=accd)

    def keys(self):
        return self.particles()["key"]

    def topnodes(self):
        c = self.counts()
        tn = np.zeros((c["ntop"], 5), dtype=np.int64)
        dni = np.zeros(c["ntopleaves"], dtype=np.int32)
        self.L.g2o_get_topnodes(self.h, _p(tn), _p(dni))
        return dict(daughter=tn[:, 0].copy(), leaf=tn[:, 1].copy(), size=tn[:, 2].copy(), startkey=tn[:, 3].copy(), count=tn[:, 4].copy(),
                    domain_node_index=dni, ntopleaves=c["ntopleaves"])

    def tree(self):
        nn, D, n = self.counts()["numnodes"], self.D, self.n
        ln = np.zeros(nn, dtype=np.float32)
        ce = np.zeros((nn, 3), dtype=np.float32)
        s = np.zeros((nn, 3, D), dtype=np.float32)
        m = np.zeros((nn, D), dtype=np.float32)
        link = np.zeros((nn, 4), dtype=np.int32)
        pn = np.zeros(n, dtype=np.int32)
        pf = np.zeros(n, dtype=np.int32)
        self.L.g2o_get_tree(self.h, _p(ln), _p(ce), _p(s), _p(m), _p(link), _p(pn), _p(pf))
        return dict(numnodes=nn, maxpart=self.maxpart, len=ln, center=ce, s=s, mass=m, bitflags=link[:, 0].copy(), sibling=link[:, 1].copy(),
                    nextnode=link[:, 2].copy(), father=link[:, 3].copy(), p_nextnode=pn, p_father=pf)

    def timings(self):
        d = np.zeros(3)
        self.L.g2o_timings(self.h, _p(d))
        return dict(domain=d[0], build=d[1], walk=d[2])

    def accel(self, tgt, src, pm, m, r2, r, n=1):
        return self.L.g2o_accel(self.h, tgt, src, pm, m, r2, r, n)

    def spline(self, tgt, src, pm, m, h, r, n=1):
        return self.L.g2o_spline(self.h, tgt, src, pm, m, h, r, n)

    def direct(self, targets):
        t = np.ascontiguousarray(targets, dtype=np.int32)
        acc = np.zeros((len(t), 3))
        self.L.g2o_direct(self.h, len(t), _p(t), _p(acc))
        return acc

    def pm_split(self):
        return self.asmth, self.rcut
