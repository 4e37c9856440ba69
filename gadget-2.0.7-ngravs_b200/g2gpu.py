"""ctypes mirror of include/g2gpu.h — the host-side (Python) view of the reference-facing C ABI.

Names follow the reference: `gravity_tree`, `force_treebuild`, `peano_hilbert_order`, ... are thin methods of
`TreeGravity` that call the corresponding C-ABI stage.  There is no CPU fallback: constructing `TreeGravity`
raises if libg2gpu.so is missing or no sm_100 device is present.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("G2GPU_LIB") or os.path.join(HERE, "libg2gpu.so")   # G2GPU_LIB: a variant build (profiles/experiments)

LAW = dict(none=0, newtonian=1, neg_newtonian=2, yukawa=3, coloyuk=4, bambam=5, sourcebambaryon=6, sourcebaryonbam=7)
SPLINE = dict(none=16, plummer=17, neg_plummer=18, bambam_spline=19, sourcebambaryon_spline=20,
              sourcebaryonbam_spline=21)

POT = dict(none=32, newtonian=33, neg_newtonian=34, bambam=35, sourcebaryonbam=36, sourcebambaryon=37)
POTSPLINE = dict(none=48, plummer=49, neg_plummer=50, bambam=51, sourcebaryonbam=52, sourcebambaryon=53)

EXPORTED = [
    "g2gpu_create", "g2gpu_destroy", "g2gpu_last_error", "g2gpu_device_count", "g2gpu_set_species", "g2gpu_set_laws",
    "g2gpu_set_srtable", "g2gpu_upload", "g2gpu_upload_aos", "g2gpu_input_buffers", "g2gpu_inputs_ready", "g2gpu_bind_inputs", "g2gpu_io_bytes", "g2gpu_domain",
    "g2gpu_get_domain", "g2gpu_get_keys", "g2gpu_get_order", "g2gpu_get_topnodes", "g2gpu_treebuild", "g2gpu_download_tree", "g2gpu_download_extnodes", "g2gpu_download_nparticles",
    "g2gpu_walk", "g2gpu_direct", "g2gpu_download_acc", "g2gpu_slice", "g2gpu_gravity_tree", "g2gpu_set_option", "g2gpu_timings", "g2gpu_get_counts",
    "g2gpu_update_tree", "g2gpu_pm_periodic", "g2gpu_pm_potential_periodic", "g2gpu_download_gravpm", "g2gpu_reset_counters", "g2gpu_stream", "g2gpu_sync", "g2gpu_peano_keys", "g2gpu_sort_pairs", "g2gpu_eval_pairs",
    "g2gpu_set_lattice_tables", "g2gpu_make_ewald_table", "g2gpu_set_lattice_pot_tables", "g2gpu_make_ewald_pot_table",
    "g2gpu_set_potential_laws", "g2gpu_set_srpot_table", "g2gpu_potential", "g2gpu_download_potential", "g2gpu_eval_potentials",
    "g2gpu_group_create", "g2gpu_group_destroy", "g2gpu_group_size", "g2gpu_group_ctx", "g2gpu_group_set_species", "g2gpu_group_set_laws",
    "g2gpu_group_set_srtable", "g2gpu_group_set_lattice_tables", "g2gpu_group_set_option", "g2gpu_group_upload", "g2gpu_group_upload_aos",
    "g2gpu_group_gather_resident", "g2gpu_group_shard", "g2gpu_group_domain", "g2gpu_group_treebuild", "g2gpu_group_update_tree",
    "g2gpu_group_walk", "g2gpu_group_download_acc", "g2gpu_group_download_aos", "g2gpu_group_get_order", "g2gpu_group_gravity_tree",
    "g2gpu_group_step_resident", "g2gpu_group_sync", "g2gpu_group_timings", "g2gpu_group_io_bytes", "g2gpu_group_slices", "g2gpu_group_zero_copy", "g2gpu_group_bind_results_aos",
]


class Config(C.Structure):
    _fields_ = [("device", C.c_int), ("n_gravs", C.c_int), ("periodic", C.c_int), ("shortrange", C.c_int),
                ("ntab", C.c_int), ("unequal_softenings", C.c_int), ("max_part", C.c_int), ("max_nodes", C.c_int),
                ("rank", C.c_int), ("nranks", C.c_int)]


class WalkParams(C.Structure):
    _fields_ = [("theta", C.c_double), ("errtol_force_acc", C.c_double), ("boxsize", C.c_double), ("G", C.c_double),
                ("asmth", C.c_double), ("rcut", C.c_double), ("pos_fac_pre_g", C.c_double),
                ("pos_fac_post_g", C.c_double), ("use_gravpm", C.c_int)]


GREENS = dict(none=0, pgdelta=1, neg_pgdelta=2, pgyukawa=3, pgcoloyuk=4)
MAX_GRAVS = 6


class PMParams(C.Structure):
    _fields_ = [("pmgrid", C.c_int), ("boxsize", C.c_double), ("asmth", C.c_double), ("G", C.c_double),
                ("greens_id", C.c_int * (MAX_GRAVS * MAX_GRAVS)), ("greens_par", C.c_double * (MAX_GRAVS * MAX_GRAVS))]


class G2Error(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"g2gpu error {code}: {text}")
        self.code = code


_lib = None


def load_library():
    """dlopen libg2gpu.so (built in-tree by __graft_entry__.build()); fails loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(LIB_PATH)
        lib.g2gpu_last_error.restype = C.c_char_p
        lib.g2gpu_stream.restype = C.c_void_p
        lib.g2gpu_group_ctx.restype = C.c_void_p
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


class TreeGravity:
    """One GPU's tree-gravity context (the device twin of the reference's global tree state)."""

    def __init__(self, max_part, n_gravs=2, periodic=False, shortrange=False, ntab=2048, unequal_softenings=True,
                 tree_alloc_factor=1.5, device=0, rank=0, nranks=1):
        self.lib = load_library()
        self.cfg = Config(device, n_gravs, int(periodic), int(shortrange), ntab, int(unequal_softenings), int(max_part),
                          int(tree_alloc_factor * max_part), rank, nranks)
        self.ctx = C.c_void_p()
        self._chk(self.lib.g2gpu_create(C.byref(self.ctx), C.byref(self.cfg)))
        self.D = n_gravs
        self.max_part = int(max_part)
        self.n = 0
        self.numnodes = 0

    def _chk(self, rc):
        if rc != 0:
            raise G2Error(rc, self.lib.g2gpu_last_error().decode())

    def close(self):
        if self.ctx:
            self.lib.g2gpu_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- init_grav_maps / wire_grav_maps (ngravs_core.c:201, ngravs.c:64) -----------------------------------
    def set_species(self, type_to_grav, force_softening):
        t2g = _i32(type_to_grav)
        fs = np.ascontiguousarray(force_softening, dtype=np.float64)
        self._chk(self.lib.g2gpu_set_species(self.ctx, _p(t2g), _p(fs)))

    def set_laws(self, accel="newtonian", spline="plummer", params=None):
        """accel/spline: a name for all pairs, or a D x D nested list of names indexed [target][source]."""
        D = self.D

        def grid(x, table):
            if isinstance(x, str):
                return np.full((D, D), table[x], dtype=np.int32)
            return np.array([[table[v] for v in row] for row in x], dtype=np.int32)
        a = grid(accel, LAW)
        s = grid(spline, SPLINE)
        par = np.zeros((D, D, 4)) if params is None else np.ascontiguousarray(params, dtype=np.float64)
        self._chk(self.lib.g2gpu_set_laws(self.ctx, _p(a), _p(s), _p(par)))

    def set_srtable(self, table):
        t = np.ascontiguousarray(table, dtype=np.float64)
        assert t.shape == (self.D, self.D, self.cfg.ntab)
        self._chk(self.lib.g2gpu_set_srtable(self.ctx, _p(t)))

    def set_option(self, name, value):
        self._chk(self.lib.g2gpu_set_option(self.ctx, name.encode(), int(value)))

    # ---- particle_data upload ----------------------------------------------------------------------------------
    def upload(self, pos, mass, ptype, oldacc=None, vel=None, gravpm=None, active=None):
        pos, mass, ptype = _f32(pos), _f32(mass), _i32(ptype)
        n = len(mass)
        self._chk(self.lib.g2gpu_upload(self.ctx, n, _p(pos), _p(mass), _p(ptype), _p(_f32(oldacc)), _p(_f32(vel)),
                                        _p(_f32(gravpm)), _p(_i32(active))))
        self.n = n

    # ---- domain_Decomposition: extent + keys + top tree + peano_hilbert_order -------------------------------------
    def domain(self):
        self._chk(self.lib.g2gpu_domain(self.ctx))

    peano_hilbert_order = domain

    def domain_info(self):
        d = np.zeros(8)
        self._chk(self.lib.g2gpu_get_domain(self.ctx, _p(d)))
        return dict(corner=d[0:3].copy(), center=d[3:6].copy(), len=d[6], fac=d[7])

    def keys(self):
        k = np.zeros(self.n, dtype=np.int64)
        self._chk(self.lib.g2gpu_get_keys(self.ctx, _p(k)))
        return k

    def order(self):
        p = np.zeros(self.n, dtype=np.int32)
        self._chk(self.lib.g2gpu_get_order(self.ctx, _p(p)))
        return p

    def topnodes(self):
        cap = 32768
        nt, nl = C.c_int(), C.c_int()
        d = np.zeros(cap, dtype=np.int32)
        lf = np.zeros(cap, dtype=np.int32)
        sk = np.zeros(cap, dtype=np.int64)
        sz = np.zeros(cap, dtype=np.int64)
        ct = np.zeros(cap, dtype=np.int64)
        dni = np.zeros(cap, dtype=np.int32)
        self._chk(self.lib.g2gpu_get_topnodes(self.ctx, C.byref(nt), C.byref(nl), _p(d), _p(lf), _p(sk), _p(sz), _p(ct), _p(dni)))
        n, l = nt.value, nl.value
        return dict(daughter=d[:n].astype(np.int64), leaf=lf[:n].astype(np.int64), startkey=sk[:n], size=sz[:n], count=ct[:n],
                    domain_node_index=dni[:l].copy(), ntopleaves=l)

    # ---- force_treebuild ----------------------------------------------------------------------------------------
    def treebuild(self):
        nn = C.c_int()
        self._chk(self.lib.g2gpu_treebuild(self.ctx, C.byref(nn)))
        self.numnodes = nn.value
        return nn.value

    force_treebuild = treebuild

    def counts(self):
        c = np.zeros(4, dtype=np.int32)
        self._chk(self.lib.g2gpu_get_counts(self.ctx, _p(c)))
        return dict(n=int(c[0]), numnodes=int(c[1]), ntargets=int(c[2]), stage=int(c[3]))

    def tree(self):
        cn = self.counts()
        self.numnodes, self.n = cn["numnodes"], cn["n"]
        nn, D, n = self.numnodes, self.D, self.n
        ln = np.zeros(nn, dtype=np.float32)
        ce = np.zeros((nn, 3), dtype=np.float32)
        s = np.zeros((nn, 3, D), dtype=np.float32)
        m = np.zeros((nn, D), dtype=np.float32)
        bf, sib, nxt, fat = (np.zeros(nn, dtype=np.int32) for _ in range(4))
        pn, pf = (np.zeros(n, dtype=np.int32) for _ in range(2))
        self._chk(self.lib.g2gpu_download_tree(self.ctx, _p(ln), _p(ce), _p(s), _p(m), _p(bf), _p(sib), _p(nxt), _p(fat), _p(pn), _p(pf)))
        return dict(numnodes=nn, maxpart=self.max_part, len=ln, center=ce, s=s, mass=m, bitflags=bf, sibling=sib, nextnode=nxt,
                    father=fat, p_nextnode=pn, p_father=pf)

    def extnodes_vs(self):
        vs = np.zeros((self.numnodes, 3, self.D), dtype=np.float32)
        self._chk(self.lib.g2gpu_download_extnodes(self.ctx, _p(vs)))
        return vs

    def nparticles(self):
        """Nodes[].u.d.Nparticles[g] of a -DNGRAVS_ACCUMULATOR build (allvars.h:645-648)."""
        cnt = np.zeros((self.numnodes, self.D), dtype=np.int64)
        self._chk(self.lib.g2gpu_download_nparticles(self.ctx, _p(cnt)))
        return cnt

    # ---- gravity_tree -------------------------------------------------------------------------------------------
    @staticmethod
    def walk_params(theta=0.5, errtol=0.005, boxsize=0.0, G=1.0, asmth=0.0, rcut=0.0, use_gravpm=0):
        return WalkParams(theta, errtol, boxsize, G, asmth, rcut, 0.0, 0.0, use_gravpm)

    def walk(self, wp):
        self._chk(self.lib.g2gpu_walk(self.ctx, C.byref(wp)))

    def direct(self, wp, targets):
        """force_treeevaluate_direct for the given current-order indices: (len(targets), 3) float64, pre-G."""
        t = _i32(targets)
        out = np.zeros((len(t), 3))
        self._chk(self.lib.g2gpu_direct(self.ctx, C.byref(wp), len(t), _p(t), _p(out)))
        return out

    # ---- lattice_init / force_treeevaluate_lattice_correction (forcetree.c:3611, 2077): periodic box without PM -------------
    def make_ewald_table(self, en=64):
        """ewald_force (ngravs.c:1170) at x = 0.5 (i,j,k)/en, computed on the device in FP64: (3, en+1, en+1, en+1), dimensionless."""
        out = np.zeros((3, en + 1, en + 1, en + 1))
        self._chk(self.lib.g2gpu_make_ewald_table(self.ctx, int(en), _p(out)))
        return out

    def set_lattice_tables(self, fcorr, en=64):
        """fcorr: (3, D, D, en+1, en+1, en+1) = fcorrx/y/z[target][source] after lattice_init (divided by BoxSize^2), or None to go
        back to the nearest-image force.  walk()/gravity_tree() then include the lattice-sum correction walk."""
        if fcorr is None:
            self._chk(self.lib.g2gpu_set_lattice_tables(self.ctx, int(en), None))
            return
        t = np.ascontiguousarray(fcorr, dtype=np.float64)
        assert t.shape == (3, self.D, self.D, en + 1, en + 1, en + 1)
        self._chk(self.lib.g2gpu_set_lattice_tables(self.ctx, int(en), _p(t)))

    def set_ewald_lattice(self, boxsize, en=64):
        """The stock wiring (LatticeForce = ewald_force for every pair, ngravs.c:131): device-made table / BoxSize^2 for all pairs."""
        t = self.make_ewald_table(en) / (boxsize * boxsize)
        self.set_lattice_tables(np.broadcast_to(t[:, None, None], (3, self.D, self.D) + t.shape[1:]).copy(), en)
        return t

    def make_ewald_pot_table(self, en=64, latticezero=float(np.float32(2.8372975))):
        """ewald_psi (ngravs.c:761) at x = 0.5 (i,j,k)/en on the device in FP64, LatticeZero (a FLOAT, ngravs.c:133) at the origin:
        (en+1, en+1, en+1), not yet divided by BoxSize."""
        out = np.zeros((en + 1, en + 1, en + 1))
        self.lib.g2gpu_make_ewald_pot_table.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        self._chk(self.lib.g2gpu_make_ewald_pot_table(self.ctx, int(en), float(latticezero), _p(out)))
        return out

    def set_lattice_pot_tables(self, potcorr, en=64):
        """potcorr: (D, D, en+1, en+1, en+1) = potcorr[target][source] after lattice_init (divided by BoxSize), or None.  potential() of a
        periodic box without PM needs it (lattice_pot_corr, forcetree.c:3895)."""
        if potcorr is None:
            self._chk(self.lib.g2gpu_set_lattice_pot_tables(self.ctx, int(en), None))
            return
        t = np.ascontiguousarray(potcorr, dtype=np.float64)
        assert t.shape == (self.D, self.D, en + 1, en + 1, en + 1)
        self._chk(self.lib.g2gpu_set_lattice_pot_tables(self.ctx, int(en), _p(t)))

    def set_ewald_pot_lattice(self, boxsize, en=64):
        """The stock wiring (LatticePotential = ewald_psi for every pair, ngravs.c:132): device-made table / BoxSize for all pairs."""
        t = self.make_ewald_pot_table(en) / boxsize
        self.set_lattice_pot_tables(np.broadcast_to(t, (self.D, self.D) + t.shape).copy(), en)
        return t

    # ---- compute_potential (potential.c:22): the tree potential of every particle -----------------------------------------
    def set_potential_laws(self, pot="newtonian", spline="plummer"):
        """PotentialFxns / PotentialSplines: a name for all pairs, or a D x D nested list of names indexed [target][source]."""
        D = self.D

        def grid(x, table):
            if isinstance(x, str):
                return np.full((D, D), table[x], dtype=np.int32)
            return np.array([[table[v] for v in row] for row in x], dtype=np.int32)
        a, s = grid(pot, POT), grid(spline, POTSPLINE)
        self._chk(self.lib.g2gpu_set_potential_laws(self.ctx, _p(a), _p(s)))

    def set_srpot_table(self, table):
        t = np.ascontiguousarray(table, dtype=np.float64)
        assert t.shape == (self.D, self.D, self.cfg.ntab)
        self._chk(self.lib.g2gpu_set_srpot_table(self.ctx, _p(t)))

    def potential(self, wp, with_time=False):
        """force_treeevaluate_potential[_shortrange] for every particle: P[].Potential as the walk leaves it (pre-G, self term included),
        float32, current particle order."""
        self._chk(self.lib.g2gpu_potential(self.ctx, C.byref(wp)))
        out = np.zeros(self.n, dtype=np.float32)
        ms = C.c_double(0)
        self._chk(self.lib.g2gpu_download_potential(self.ctx, _p(out), C.byref(ms)))
        return (out, ms.value) if with_time else out

    def download_acc(self, out=None):
        """out = (acc[n,3], cost[n], oldacc[n]) float32 arrays (e.g. pinned) to receive the results; allocated when None."""
        if out is not None:
            acc, cost, old = out
        else:
            acc = np.zeros((self.n, 3), dtype=np.float32)
            cost = np.zeros(self.n, dtype=np.float32)
            old = np.zeros(self.n, dtype=np.float32)
        self._chk(self.lib.g2gpu_download_acc(self.ctx, _p(acc), _p(cost), _p(old)))
        return acc, cost, old

    def gravity_tree(self, pos, mass, ptype, wp, oldacc=None, active=None, out=None):
        """Whole step from host buffers (upload, domain, build, walk, download): the e2e path.
        out = (acc[n,3] f32, cost[n] f32, oldacc[n] f32, perm[n] i32) lets the caller supply (pinned) result buffers."""
        pos, mass, ptype = _f32(pos), _f32(mass), _i32(ptype)
        n = len(mass)
        if out is not None:
            acc, cost, old, perm = out
        else:
            acc = np.zeros((n, 3), dtype=np.float32)
            cost = np.zeros(n, dtype=np.float32)
            old = np.zeros(n, dtype=np.float32)
            perm = np.zeros(n, dtype=np.int32)
        self._chk(self.lib.g2gpu_gravity_tree(self.ctx, n, _p(pos), _p(mass), _p(ptype), _p(_f32(oldacc)), _p(_i32(active)),
                                              C.byref(wp), _p(acc), _p(cost), _p(old), _p(perm)))
        self.n = n
        return acc, cost, old, perm

    def slice(self):
        lo, hi = C.c_int(), C.c_int()
        self._chk(self.lib.g2gpu_slice(self.ctx, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def input_buffers(self, n):
        ptr = C.c_void_p()
        self._chk(self.lib.g2gpu_input_buffers(self.ctx, int(n), C.byref(ptr)))
        return int(ptr.value)

    def bind_inputs(self, n, records_ptr):
        """Use a caller-owned device array of n 32-byte g2gpu_particle records (raw pointer, e.g. a torch tensor's data_ptr())."""
        self._chk(self.lib.g2gpu_bind_inputs(self.ctx, int(n), C.c_void_p(records_ptr)))
        self.n = int(n)

    def io_bytes(self):
        b = np.zeros(2, dtype=np.int64)
        self._chk(self.lib.g2gpu_io_bytes(self.ctx, _p(b)))
        return int(b[0]), int(b[1])

    def inputs_ready(self, n):
        self._chk(self.lib.g2gpu_inputs_ready(self.ctx, int(n)))
        self.n = int(n)

    def update_tree(self, length, s):
        """Dynamic tree update: re-attach the last tree to freshly uploaded particles with the host's node len / s (reference numbering)."""
        l = _f32(length)
        ss = _f32(s)
        assert len(l) == self.numnodes and ss.size == 3 * self.numnodes * self.D
        self._chk(self.lib.g2gpu_update_tree(self.ctx, _p(l), _p(ss)))

    # ---- long_range_force -> pmforce_periodic (longrange.c:56, pm_periodic.c:204) --------------------------------------
    def _pm_params(self, pmgrid, boxsize, G, asmth, greens, greens_par):
        D = self.D
        pp = PMParams()
        pp.pmgrid, pp.boxsize, pp.G = int(pmgrid), float(boxsize), float(G)
        pp.asmth = float(1.25 * boxsize / pmgrid if asmth is None else asmth)      # ASMTH = 1.25 (allvars.h:82)
        for a in range(D):
            for b in range(D):
                name = greens if isinstance(greens, str) else greens[a][b]
                pp.greens_id[a * D + b] = GREENS[name]
                pp.greens_par[a * D + b] = 0.0 if greens_par is None else float(np.asarray(greens_par).reshape(D, D)[a, b])
        return pp

    def pm_potential_periodic(self, pmgrid, boxsize, G=1.0, asmth=None, greens="pgdelta", greens_par=None):
        """pmpotential_periodic (pm_periodic.c:798): what it adds to P[].Potential, float32[n] in upload order."""
        pp = self._pm_params(pmgrid, boxsize, G, asmth, greens, greens_par)
        out = np.zeros(self.n, dtype=np.float32)
        self._chk(self.lib.g2gpu_pm_potential_periodic(self.ctx, C.byref(pp), _p(out)))
        return out

    def pm_periodic(self, pmgrid, boxsize, G=1.0, asmth=None, greens="pgdelta", greens_par=None, download=True):
        """Periodic PM long-range force of the particles uploaded last; returns GravPM[n,3] in upload order.
        greens: a name for all pairs or a D x D nested list indexed [source][target] like GreensFxns[nA][nB]."""
        pp = self._pm_params(pmgrid, boxsize, G, asmth, greens, greens_par)
        self._chk(self.lib.g2gpu_pm_periodic(self.ctx, C.byref(pp)))
        if not download:
            return None
        out = np.zeros((self.n, 3), dtype=np.float32)
        self._chk(self.lib.g2gpu_download_gravpm(self.ctx, _p(out)))
        return out

    def pm_device(self, pmgrid, boxsize, **kw):
        """pm_periodic without the device->host copy: GravPM stays on the device as the input of the next domain()/walk()."""
        return self.pm_periodic(pmgrid, boxsize, download=False, **kw)

    def timings(self):
        ms = np.zeros(8)
        cnt = np.zeros(8, dtype=np.int64)
        self._chk(self.lib.g2gpu_timings(self.ctx, _p(ms), _p(cnt)))
        return dict(domain_ms=ms[0], build_ms=ms[1], walk_ms=ms[2], walk_kernel_ms=ms[3], sort_ms=ms[4], h2d_ms=ms[5], d2h_ms=ms[6], pm_ms=ms[7],
                    launches=int(cnt[0]), interactions=int(cnt[1]), cell_visits=int(cnt[2]), species_terms=int(cnt[3]), decisions=int(cnt[4]), rewalked=int(cnt[5]), border_checked=int(cnt[6]))

    def reset_counters(self):
        self.lib.g2gpu_reset_counters(self.ctx)

    def sync(self):
        self._chk(self.lib.g2gpu_sync(self.ctx))

    @property
    def stream(self):
        return self.lib.g2gpu_stream(self.ctx)

    # ---- stand-alone kernels -------------------------------------------------------------------------------------
    def peano_keys(self, xyz, bits):
        xyz = _i32(xyz)
        k = np.zeros(len(xyz), dtype=np.int64)
        self._chk(self.lib.g2gpu_peano_keys(self.ctx, len(xyz), _p(xyz), int(bits), _p(k)))
        return k

    def sort_pairs(self, keys, vals, begin_bit=0, end_bit=64):
        k = np.ascontiguousarray(keys, dtype=np.uint64).copy()
        v = np.ascontiguousarray(vals, dtype=np.uint32).copy()
        self._chk(self.lib.g2gpu_sort_pairs(self.ctx, len(k), _p(k), _p(v), begin_bit, end_bit))
        return k, v

    def eval_pairs(self, tgt, src, pm, m, r, h, nn=None):
        pm, m, r, h = _f32(pm), _f32(m), _f32(r), _f32(h)
        fac = np.zeros(len(r), dtype=np.float32)
        self._chk(self.lib.g2gpu_eval_pairs(self.ctx, len(r), tgt, src, _p(pm), _p(m), _p(r), _p(h), _p(_i32(nn)), _p(fac)))
        return fac

    def eval_potentials(self, tgt, src, pm, m, r, h, nn=None):
        """What the potential walk adds to pot for each pair (no TreePM table term): -PotentialFxns for r >= h, +PotentialSplines for r < h."""
        pm, m, r, h = _f32(pm), _f32(m), _f32(r), _f32(h)
        out = np.zeros(len(r), dtype=np.float32)
        self._chk(self.lib.g2gpu_eval_potentials(self.ctx, len(r), tgt, src, _p(pm), _p(m), _p(r), _p(h), _p(_i32(nn)), _p(out)))
        return out


class TreeGravityGroup:
    """Several GPUs of one node behind the same stage calls (g2gpu_group_*, csrc/g2_group.cu): ONE process, one context + stream +
    host thread per device, sharded upload + NCCL all-gather of the particle records, replicated tree, one target slice per device.
    It is the device twin of the parallel driver inside the reference's gravity_tree() (gravtree.c:102-285)."""

    def __init__(self, max_part, n_gravs=2, periodic=False, shortrange=False, ntab=2048, unequal_softenings=True, tree_alloc_factor=1.5,
                 ndev=0, devices=None):
        self.lib = load_library()
        self.cfg = Config(0, n_gravs, int(periodic), int(shortrange), ntab, int(unequal_softenings), int(max_part),
                          int(tree_alloc_factor * max_part), 0, 1)
        self.grp = C.c_void_p()
        dv = None if devices is None else _i32(devices)
        self._chk(self.lib.g2gpu_group_create(C.byref(self.grp), C.byref(self.cfg), int(ndev if devices is None else len(devices)), _p(dv)))
        self.ndev = int(self.lib.g2gpu_group_size(self.grp))
        self.D = n_gravs
        self.max_part = int(max_part)
        self.n = 0
        self.numnodes = 0

    def _chk(self, rc):
        if rc != 0:
            raise G2Error(rc, self.lib.g2gpu_last_error().decode())

    def close(self):
        if self.grp:
            self.lib.g2gpu_group_destroy(self.grp)
            self.grp = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def ctx(self, i=0):
        """Device i's context as a (non-owning) TreeGravity: tree mirror, potential, PM, direct sums are served by device 0."""
        t = TreeGravity.__new__(TreeGravity)
        t.lib, t.cfg, t.D, t.max_part, t.n, t.numnodes = self.lib, self.cfg, self.D, self.max_part, self.n, self.numnodes
        t.ctx = C.c_void_p(self.lib.g2gpu_group_ctx(self.grp, int(i)))
        t.close = lambda: None
        return t

    def set_species(self, type_to_grav, force_softening):
        t2g = _i32(type_to_grav)
        fs = np.ascontiguousarray(force_softening, dtype=np.float64)
        self._chk(self.lib.g2gpu_group_set_species(self.grp, _p(t2g), _p(fs)))

    def set_laws(self, accel="newtonian", spline="plummer", params=None):
        D = self.D

        def grid(x, table):
            if isinstance(x, str):
                return np.full((D, D), table[x], dtype=np.int32)
            return np.array([[table[v] for v in row] for row in x], dtype=np.int32)
        a, s = grid(accel, LAW), grid(spline, SPLINE)
        par = np.zeros((D, D, 4)) if params is None else np.ascontiguousarray(params, dtype=np.float64)
        self._chk(self.lib.g2gpu_group_set_laws(self.grp, _p(a), _p(s), _p(par)))

    def set_srtable(self, table):
        t = np.ascontiguousarray(table, dtype=np.float64)
        assert t.shape == (self.D, self.D, self.cfg.ntab)
        self._chk(self.lib.g2gpu_group_set_srtable(self.grp, _p(t)))

    def set_option(self, name, value):
        self._chk(self.lib.g2gpu_group_set_option(self.grp, name.encode(), int(value)))

    walk_params = staticmethod(lambda **kw: TreeGravity.walk_params(**kw))

    def upload(self, pos, mass, ptype, oldacc=None, active=None):
        pos, mass, ptype = _f32(pos), _f32(mass), _i32(ptype)
        self.n = len(mass)
        self._chk(self.lib.g2gpu_group_upload(self.grp, self.n, _p(pos), _p(mass), _p(ptype), _p(_f32(oldacc)), _p(_i32(active))))

    def shard(self, n, i):
        lo, cnt = C.c_int(), C.c_int()
        self._chk(self.lib.g2gpu_group_shard(self.grp, int(n), int(i), C.byref(lo), C.byref(cnt)))
        return lo.value, cnt.value

    def gather_resident(self, n):
        self._chk(self.lib.g2gpu_group_gather_resident(self.grp, int(n)))
        self.n = int(n)

    def domain(self):
        self._chk(self.lib.g2gpu_group_domain(self.grp))

    def treebuild(self):
        nn = C.c_int()
        self._chk(self.lib.g2gpu_group_treebuild(self.grp, C.byref(nn)))
        self.numnodes = nn.value
        return nn.value

    def walk(self, wp):
        self._chk(self.lib.g2gpu_group_walk(self.grp, C.byref(wp)))

    def download_acc(self, out=None):
        if out is not None:
            acc, cost, old = out
        else:
            acc = np.zeros((self.n, 3), dtype=np.float32)
            cost = np.zeros(self.n, dtype=np.float32)
            old = np.zeros(self.n, dtype=np.float32)
        self._chk(self.lib.g2gpu_group_download_acc(self.grp, _p(acc), _p(cost), _p(old)))
        return acc, cost, old

    def get_order(self):
        perm = np.zeros(self.n, dtype=np.int32)
        self._chk(self.lib.g2gpu_group_get_order(self.grp, _p(perm)))
        return perm

    def gravity_tree(self, pos, mass, ptype, wp, oldacc=None, active=None, out=None):
        """Whole step from host buffers on all devices (the e2e path); results in CURRENT (device) order + perm, like TreeGravity.gravity_tree."""
        pos, mass, ptype = _f32(pos), _f32(mass), _i32(ptype)
        n = len(mass)
        if out is not None:
            acc, cost, old, perm = out
        else:
            acc = np.zeros((n, 3), dtype=np.float32)
            cost = np.zeros(n, dtype=np.float32)
            old = np.zeros(n, dtype=np.float32)
            perm = np.zeros(n, dtype=np.int32)
        self._chk(self.lib.g2gpu_group_gravity_tree(self.grp, n, _p(pos), _p(mass), _p(ptype), _p(_f32(oldacc)), _p(_i32(active)),
                                                    C.byref(wp), _p(acc), _p(cost), _p(old), _p(perm)))
        self.n = n
        return acc, cost, old, perm

    def step_resident(self, n, wp):
        """all-gather of the resident shards -> domain -> treebuild -> walk on every device (results stay there)."""
        self._chk(self.lib.g2gpu_group_step_resident(self.grp, int(n), C.byref(wp)))
        self.n = int(n)

    def stream(self, i=0):
        return self.lib.g2gpu_stream(C.c_void_p(self.lib.g2gpu_group_ctx(self.grp, int(i))))

    def sync(self):
        self._chk(self.lib.g2gpu_group_sync(self.grp))

    def timings(self):
        ms = np.zeros(8)
        cnt = np.zeros(8, dtype=np.int64)
        self._chk(self.lib.g2gpu_group_timings(self.grp, _p(ms), _p(cnt)))
        return dict(domain_ms=ms[0], build_ms=ms[1], walk_ms=ms[2], walk_kernel_ms=ms[3], sort_ms=ms[4], h2d_ms=ms[5], d2h_ms=ms[6], pm_ms=ms[7],
                    launches=int(cnt[0]), interactions=int(cnt[1]), cell_visits=int(cnt[2]), species_terms=int(cnt[3]), decisions=int(cnt[4]),
                    rewalked=int(cnt[5]), border_checked=int(cnt[6]))

    def io_bytes(self):
        b = np.zeros(3, dtype=np.int64)
        self._chk(self.lib.g2gpu_group_io_bytes(self.grp, _p(b)))
        return int(b[0]), int(b[1]), int(b[2])

    def zero_copy(self):
        """True when the last gravity_tree() let the walk kernel store the results straight into the (pinned) `out` arrays."""
        return bool(self.lib.g2gpu_group_zero_copy(self.grp))

    def slices(self):
        lo, hi = np.zeros(self.ndev, dtype=np.int32), np.zeros(self.ndev, dtype=np.int32)
        fr = np.zeros(self.ndev + 1)
        self._chk(self.lib.g2gpu_group_slices(self.grp, _p(lo), _p(hi), _p(fr)))
        return lo, hi, fr
