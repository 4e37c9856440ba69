"""ctypes driver for oracle/_ref/libg2ref_<variant>.so — TEST INFRASTRUCTURE ONLY.

The libraries are the UNMODIFIED reference sources compiled by oracle/ref/Makefile.  The reference keeps
all state in globals and can be initialised once per image, so every RefOracle instance loads a private
copy of the shared object.
"""
import ctypes as C
import os
import shutil
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def available(variant, prefix="g2ref"):
    return os.path.exists(os.path.join(REF_DIR, f"lib{prefix}_{variant}.so"))


class RefOracle:
    """One single-rank instance of the reference code (variant = compile-time flag set)."""

    def __init__(self, variant, maxpart, boxsize=0.0, G=1.0, theta=0.5, errtol=0.005, criterion=1,
                 softening=(0.0, 1.0, 1.0, 1.0, 1.0, 1.0), gravity=(0, 0, 1, 0, 0, 0), tree_alloc=1.5,
                 buffer_mb=32, prefix="g2ref"):
        """prefix="g2ref": the unmodified reference; prefix="g2shim": the same reference linked with the product's host shim
        (integration/Makefile), i.e. its gravity_tree()/peano_hilbert_order()/force_treebuild() run on the GPU."""
        src = os.path.join(REF_DIR, f"lib{prefix}_{variant}.so")
        if not os.path.exists(src):
            raise FileNotFoundError(src)
        fd, self._tmp = tempfile.mkstemp(suffix=".so", prefix=f"g2ref_{variant}_")
        os.close(fd)
        shutil.copyfile(src, self._tmp)
        self.lib = lib = C.CDLL(self._tmp)
        os.unlink(self._tmp)
        cfg = np.zeros(16, dtype=np.int32)
        lib.g2ref_config(cfg.ctypes.data_as(C.c_void_p))
        self.D, self.float_bytes, self.periodic, self.pmgrid, self.ntab, self.unequal, self.forcetest = (
            int(x) for x in cfg[:7])
        self.sizeof_particle, self.sizeof_node, self.sizeof_extnode, self.bits = (int(x) for x in cfg[7:11])
        self.fdtype = np.float32 if self.float_bytes == 4 else np.float64
        par = np.zeros(20)
        par[0] = maxpart
        par[1] = boxsize
        par[2] = G
        par[3] = theta
        par[4] = errtol
        par[5] = criterion
        par[6:12] = softening
        par[12:18] = gravity
        par[18] = tree_alloc
        par[19] = buffer_mb
        lib.g2ref_walk_threads.restype = C.c_double
        lib.g2ref_peano_key.restype = C.c_longlong
        lib.g2ref_accel.restype = C.c_double
        lib.g2ref_spline.restype = C.c_double
        lib.g2ref_accel.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_long]
        lib.g2ref_spline.argtypes = lib.g2ref_accel.argtypes
        # lattice_init() (PERIODIC without PMGRID) caches its tables as files in the working directory (forcetree.c:3637-3745).  In a
        # FLOAT = float build that cache is unusable: the tables are `static double` (forcetree.c:49-52) but are written and read with
        # sizeof(FLOAT) per entry, so a run that FINDS the file gets only the first half of every table (the rest stays zero).  The
        # oracle therefore always runs in a fresh scratch directory: tables are computed (~1 min), the file is discarded.
        self._scratch = tempfile.mkdtemp(prefix="g2ref_lattice_")
        os.environ["G2REF_CACHE_DIR"] = self._scratch
        # the reference prints progress with printf; silence fd 1 during calls
        rc = self._quiet(lib.g2ref_setup, par.ctypes.data_as(C.c_void_p))
        shutil.rmtree(self._scratch, ignore_errors=True)
        if rc != 0:
            raise RuntimeError("g2ref_setup failed")
        self.maxpart = maxpart
        self.n = 0

    @staticmethod
    def _quiet(fn, *args):
        import sys
        if os.environ.get("G2REF_VERBOSE"):
            return fn(*args)
        sys.stdout.flush()
        saved = os.dup(1)
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 1)
        try:
            return fn(*args)
        finally:
            C.CDLL(None).fflush(None)
            os.dup2(saved, 1)
            os.close(saved)
            os.close(devnull)

    # ---- inputs -------------------------------------------------------------------------------------
    def load(self, pos, mass, ptype, vel=None):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        mass = np.ascontiguousarray(mass, dtype=np.float64)
        ptype = np.ascontiguousarray(ptype, dtype=np.int32)
        n = len(mass)
        velp = None
        if vel is not None:
            vel = np.ascontiguousarray(vel, dtype=np.float64)
            velp = vel.ctypes.data_as(C.c_void_p)
        rc = self.lib.g2ref_load(n, pos.ctypes.data_as(C.c_void_p), velp, mass.ctypes.data_as(C.c_void_p),
                                 ptype.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise RuntimeError(f"g2ref_load failed ({rc})")
        self.n = n

    def set_active(self, active):
        a = np.ascontiguousarray(active, dtype=np.int32)
        self.lib.g2ref_set_active(a.ctypes.data_as(C.c_void_p))

    def set_opening(self, theta, errtol, criterion):
        self.lib.g2ref_set_opening(C.c_double(theta), C.c_double(errtol), int(criterion))

    def set_oldacc(self, oldacc):
        a = np.ascontiguousarray(oldacc, dtype=np.float64)
        self.lib.g2ref_set_oldacc(a.ctypes.data_as(C.c_void_p))

    # ---- reference entry points -----------------------------------------------------------------------
    def domain(self):
        """domain_Decomposition(): extent, keys, top tree, species-major PH reorder of P[]."""
        self._quiet(self.lib.g2ref_domain)

    def gravity(self):
        """gravity_tree(): build if flagged + walk of active particles + OldAcc + G scaling."""
        self._quiet(self.lib.g2ref_gravity)

    def treebuild(self):
        return self._quiet(self.lib.g2ref_treebuild)

    def force_rebuild(self):
        self.lib.g2ref_force_rebuild()

    def walk_threads(self, nthreads, lo=0, hi=None):
        cost = C.c_double(0)
        hi = self.n if hi is None else hi
        dt = self.lib.g2ref_walk_threads(int(lo), int(hi), int(nthreads), C.byref(cost))
        return dt, cost.value

    # ---- outputs --------------------------------------------------------------------------------------
    def keys(self):
        k = np.zeros(self.n, dtype=np.int64)
        self.lib.g2ref_keys(k.ctypes.data_as(C.c_void_p))
        return k

    def peano_key(self, x, y, z, bits):
        return int(self.lib.g2ref_peano_key(int(x), int(y), int(z), int(bits)))

    def domain_info(self):
        d = np.zeros(8)
        self.lib.g2ref_get_domain(d.ctypes.data_as(C.c_void_p))
        return dict(corner=d[0:3].copy(), center=d[3:6].copy(), len=d[6], fac=d[7])

    def particles(self):
        n = self.n
        pos = np.zeros((n, 3))
        mass = np.zeros(n)
        ptype = np.zeros(n, dtype=np.int32)
        pid = np.zeros(n, dtype=np.uint32)
        acc = np.zeros((n, 3))
        cost = np.zeros(n, dtype=np.float32)
        oldacc = np.zeros(n)
        self.lib.g2ref_get_particles(*(a.ctypes.data_as(C.c_void_p) for a in (pos, mass, ptype, pid, acc, cost, oldacc)))
        return dict(pos=pos, mass=mass, type=ptype, id=pid, acc=acc, cost=cost, oldacc=oldacc)

    def run_forcetest(self):
        """gravity_forcetest() (gravtree_forcetest.c:28): P[].GravAccelDirect of every particle, current order; FORCETEST variants only."""
        out = np.zeros((self.n, 3))
        if self.lib.g2ref_forcetest(out.ctypes.data_as(C.c_void_p)) != 0:
            raise RuntimeError("this oracle variant was compiled without FORCETEST (or with PERIODIC)")
        return out

    def pmforce(self):
        """GravPM of every particle in the current order of P[] (pmforce_periodic, pm_periodic.c:204); PM variants only."""
        out = np.zeros((self.n, 3))
        rc = self.lib.g2ref_pmforce(out.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise RuntimeError("this oracle variant was compiled without PMGRID/PERIODIC")
        return out

    def potcorr_tables(self):
        """potcorr[tgt][src][EN+1]^3 after lattice_init (PERIODIC without PMGRID), shape (D, D, EN+1, EN+1, EN+1); else None."""
        en1 = 65
        out = np.zeros((self.D, self.D, en1, en1, en1))
        got = self.lib.g2ref_get_potcorr(out.ctypes.data_as(C.c_void_p))
        return out if got == en1 else None

    def lattice_pot_corr(self, dx, dy, dz, tgt, src):
        self.lib.g2ref_lattice_pot_corr.restype = C.c_double
        self.lib.g2ref_lattice_pot_corr.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
        return self.lib.g2ref_lattice_pot_corr(dx, dy, dz, tgt, src)

    def pmpotential(self):
        """What pmpotential_periodic (pm_periodic.c:798) adds to P[].Potential, current order of P[]; PM variants only."""
        out = np.zeros(self.n)
        rc = self.lib.g2ref_pmpotential(out.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise RuntimeError("this oracle variant was compiled without PMGRID/PERIODIC")
        return out

    def topnodes(self):
        nt = self.lib.g2ref_ntopnodes()
        nl = self.lib.g2ref_ntopleaves()
        tn = np.zeros((nt, 7), dtype=np.int64)
        dni = np.zeros(nl, dtype=np.int32)
        self.lib.g2ref_get_topnodes(tn.ctypes.data_as(C.c_void_p), dni.ctypes.data_as(C.c_void_p))
        return dict(daughter=tn[:, 0].copy(), leaf=tn[:, 3].copy(), size=tn[:, 4].copy(), startkey=tn[:, 5].copy(),
                    count=tn[:, 6].copy(), domain_node_index=dni, ntopleaves=nl)

    def tree(self):
        nn = self.lib.g2ref_numnodes()
        D = self.D
        geom = np.zeros((nn, 4))
        s = np.zeros((nn, 3, D))
        mass = np.zeros((nn, D))
        vs = np.zeros((nn, 3, D))
        link = np.zeros((nn, 4), dtype=np.int32)
        nextnode = np.zeros(self.n, dtype=np.int32)
        father = np.zeros(self.n, dtype=np.int32)
        self.lib.g2ref_get_tree(*(a.ctypes.data_as(C.c_void_p) for a in (geom, s, mass, vs, link, nextnode, father)))
        return dict(numnodes=nn, maxpart=self.maxpart, len=geom[:, 0].copy(), center=geom[:, 1:4].copy(), s=s, mass=mass,
                    vs=vs, bitflags=link[:, 0].copy(), sibling=link[:, 1].copy(), nextnode=link[:, 2].copy(),
                    father=link[:, 3].copy(), p_nextnode=nextnode, p_father=father)

    def nparticles(self):
        """Nodes[].u.d.Nparticles (variants built with -DNGRAVS_ACCUMULATOR), else None."""
        cnt = np.zeros((self.lib.g2ref_numnodes(), self.D), dtype=np.int64)
        return cnt if self.lib.g2ref_get_nparticles(cnt.ctypes.data_as(C.c_void_p)) else None

    def srtable(self):
        if not self.pmgrid:
            return None
        t = np.zeros((self.D, self.D, self.ntab))
        self.lib.g2ref_get_srtable(t.ctypes.data_as(C.c_void_p))
        return t

    def lattice_tables(self):
        """fcorrx/y/z[tgt][src][EN+1]^3 after lattice_init (PERIODIC without PMGRID), shape (3, D, D, EN+1, EN+1, EN+1); else None."""
        en1 = 65
        out = np.zeros((3, self.D, self.D, en1, en1, en1))
        got = self.lib.g2ref_get_lattice_tables(out.ctypes.data_as(C.c_void_p))
        if got == 0:
            return None
        assert got == en1
        return out

    def srpot_table(self):
        """shortrange_fourier_pot[tgt][src][NTAB] (forcetree.c:34, filled at 3346); PM variants only."""
        if not self.pmgrid:
            return None
        t = np.zeros((self.D, self.D, self.ntab))
        self.lib.g2ref_get_srpot_table(t.ctypes.data_as(C.c_void_p))
        return t

    def potential(self, nthreads=8):
        """P[].Potential of every particle (current order) straight after the reference's own per-target potential walk
        (force_treeevaluate_potential_shortrange, forcetree.c:2789): pre-G, self term included.  PM variants only."""
        out = np.zeros(self.n)
        if self.lib.g2ref_potential(out.ctypes.data_as(C.c_void_p), int(nthreads)) != 0:
            raise RuntimeError("this oracle variant has no potential walk (needs PMGRID)")
        return out

    def potfxn(self, tgt, src, pm, m, h, r, n=1):
        self.lib.g2ref_potfxn.restype = C.c_double
        self.lib.g2ref_potfxn.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_long]
        return self.lib.g2ref_potfxn(tgt, src, pm, m, h, r, n)

    def named_pot(self, which, pm, m, h, r, n=1):
        """bambam_pot (0), sourcebaryonbam_pot (1), sourcebambaryon_pot (2), newtonian_pot (3), plummer_pot (4) called directly."""
        self.lib.g2ref_named_pot.restype = C.c_double
        self.lib.g2ref_named_pot.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_long]
        return self.lib.g2ref_named_pot(int(which), pm, m, h, r, n)

    def potspline(self, tgt, src, pm, m, h, r, n=1):
        self.lib.g2ref_potspline.restype = C.c_double
        self.lib.g2ref_potspline.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_long]
        return self.lib.g2ref_potspline(tgt, src, pm, m, h, r, n)

    def pm_split(self):
        d = np.zeros(2)
        self.lib.g2ref_get_pm_split(d.ctypes.data_as(C.c_void_p))
        return d[0], d[1]

    def softening(self):
        d = np.zeros(6)
        self.lib.g2ref_get_softening(d.ctypes.data_as(C.c_void_p))
        return d

    def timings(self):
        d = np.zeros(6)
        self.lib.g2ref_timings(d.ctypes.data_as(C.c_void_p))
        return dict(domain=d[0], gravity=d[1], build=d[2], walk=d[3], peano=d[4], cpu_domain=d[5])

    def accel(self, tgt, src, pm, m, r2, r, n=1):
        return self.lib.g2ref_accel(tgt, src, pm, m, r2, r, n)

    def spline(self, tgt, src, pm, m, h, r, n=1):
        return self.lib.g2ref_spline(tgt, src, pm, m, h, r, n)

    def direct(self, targets):
        t = np.ascontiguousarray(targets, dtype=np.int32)
        acc = np.zeros((len(t), 3))
        rc = self.lib.g2ref_direct(len(t), t.ctypes.data_as(C.c_void_p), acc.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise RuntimeError("this oracle variant has no direct summation (needs FORCETEST, non-periodic)")
        return acc
