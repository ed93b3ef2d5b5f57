"""Host-side mirror of the reference interface for the element-local hot path, on top of the C ABI.

The reference computes one element at a time:
    call elem(Mdle, Itest,Itrial)         problems/<PROB>/elem.F90:20        -> ALOC/BLOC  (src/modules/assembly.F90:36-37)
    call stc_fwd_wrapper(Iel,Mdle)        src/modules/stc.F90:182            -> condensed ALOC/BLOC + CLOC(Iel)%ASchur/BSchur
inside the OpenMP element loop of par_mumps_sc (src/solver/par_mumps/par_mumps_sc.F90:347-357).  Here the same two
steps run for all elements of a subdomain in one call (`ElemEngine.elem_stc_batch`), and `stc_bwd_batch` mirrors
stc_bwd (stc.F90:661-677).  Argument names follow the reference (norder, norient_edge, norient_face, xnod).

Everything numeric happens in hp3d_b200/libhp3d_gpu.so (CUDA, sm_100a).  There is no CPU fallback: without the
library or without a GPU these calls raise.
"""
import ctypes as C

import numpy as np

from . import _lib

POIS_GAL, POIS_PDPG, MAXW_GAL, MAXW_UW = 1, 2, 3, 4
GRAPH_NORM, MATH_NORM, GRAPH_DIAG = 1, 2, 3
SRC_ZERO, SRC_SIN, SRC_TABLE = 0, 1, 9
MDLB, MDLP = 1, 3   # element types (src/modules/node_types.F90:8-10)


def _ptr(a, t=C.c_void_p):
    return None if a is None else a.ctypes.data_as(t)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def pinned_empty(shape, dtype):
    """Page-locked host array (hp3d_gpu_host_alloc) so that result copies are asynchronous DMA transfers."""
    L = _lib.lib()
    L.hp3d_gpu_host_alloc.restype = C.c_void_p
    L.hp3d_gpu_host_alloc.argtypes = [C.c_longlong]
    L.hp3d_gpu_host_free.argtypes = [C.c_void_p]
    dt = np.dtype(dtype)
    n = int(np.prod(shape))
    p = L.hp3d_gpu_host_alloc(max(n * dt.itemsize, 8))
    if not p:
        raise MemoryError("hp3d_gpu_host_alloc failed")
    buf = (C.c_char * (n * dt.itemsize)).from_address(p)
    arr = np.frombuffer(buf, dtype=dt, count=n).reshape(shape)
    arr.flags.writeable = True
    return _Pinned(arr, p)


class _Pinned:
    """Owner of a pinned allocation; `.a` is the numpy view."""

    def __init__(self, arr, ptr):
        self.a, self._ptr = arr, ptr

    def free(self):
        if self._ptr:
            _lib.lib().hp3d_gpu_host_free(C.c_void_p(self._ptr))
            self._ptr = None
            self.a = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Physics(C.Structure):
    """Mirror of `hp3d_physics` (include/hp3d_gpu.h): the problem's physics table (src/modules/physics.F90)."""
    _fields_ = [("nphys", C.c_int), ("dtype", C.c_int * 8), ("ncomp", C.c_int * 8), ("adres", C.c_int * 8), ("nrvar", C.c_int * 3)]


def physics_default(kind):
    ph = Physics()
    _lib.check(_lib.lib().hp3d_gpu_physics_default(int(kind), C.byref(ph)))
    return ph


def celem_pack(ph, nrdofl, nrcon, nac, constr, nrdofm_f):
    """Flat per-modified-dof lists (cptr, cidx, cval) of ONE element from the output of `logic` (hp3d_gpu_celem_pack).
    nrcon/nac/constr: per family (H1, H(curl), H(div)) arrays (nk,), (nk,nacdim), (nk,nacdim); nac holds 1-based indices."""
    L = _lib.lib()
    nrdofl, nrdofm_f = _i32(nrdofl), _i32(nrdofm_f)
    nacdim = max([np.asarray(a).shape[1] for a in nac if np.asarray(a).ndim == 2 and np.asarray(a).size] + [1])
    arrs = []
    for f in range(3):
        k = int(nrdofl[f])
        rc = np.zeros(k, np.int32); na = np.zeros((k, nacdim), np.int32); co = np.zeros((k, nacdim))
        if k:
            rc[:] = np.asarray(nrcon[f])[:k]
            a = np.asarray(nac[f]); c = np.asarray(constr[f])
            na[:, :a.shape[1]] = a[:k]; co[:, :c.shape[1]] = c[:k]
        arrs += [rc, na, co]
    nm = int(nrdofm_f.sum())
    cptr = np.zeros(nm + 1, np.int64)
    f_ = L.hp3d_gpu_celem_pack
    f_.restype = C.c_longlong
    f_.argtypes = [C.c_void_p] * 11 + [C.c_int] + [C.c_void_p] * 4 + [C.c_longlong]
    args = [C.byref(ph), _ptr(nrdofl)] + [_ptr(a) for a in arrs] + [int(nacdim), _ptr(nrdofm_f), _ptr(cptr)]
    n = f_(*args, None, None, 0)
    if n < 0:
        _lib.check(int(n))
    cidx = np.zeros(max(n, 1), np.int32); cval = np.zeros(max(n, 1))
    n2 = f_(*args, _ptr(cidx), _ptr(cval), int(n))
    if n2 < 0:
        _lib.check(int(n2))
    return cptr, cidx[:n], cval[:n]


class ElemEngine:
    """One plan = one problem (`elem` plugin of the reference) with fixed parameters."""

    _initialised = None

    def __init__(self, kind, device=0, **params):
        self.L = _lib.lib()
        L = self.L
        if ElemEngine._initialised != device:
            _lib.check(L.hp3d_gpu_init(int(device)))
            ElemEngine._initialised = device
        prm = _lib.Params()
        L.hp3d_gpu_params_default(C.byref(prm))
        for k, v in params.items():
            if not hasattr(prm, k):
                raise TypeError(f"unknown parameter {k}")
            if k == "eps_tensor":   # (3,3) complex, eps_t[i, j]; the struct holds it column-major, (re, im) interleaved
                t = np.asarray(v, dtype=np.complex128).reshape(3, 3).T.copy().view(np.float64).ravel()
                for i in range(18):
                    prm.eps_tensor[i] = t[i]
            else:
                setattr(prm, k, v)
        self.kind, self.prm = kind, prm
        self.complex = kind >= MAXW_GAL
        self.dtype = np.complex128 if self.complex else np.float64
        self.plan = L.hp3d_gpu_plan(int(kind), C.byref(prm))
        if self.plan < 0:
            _lib.check(self.plan)

    def close(self):
        if self.plan is not None and self.plan >= 0:
            self.L.hp3d_gpu_plan_destroy(self.plan)
            self.plan = None

    def sizes(self, norder, etype=MDLB):
        """(ni, nb, nint, nrdofH) for one element order vector, as stc_get_nrdof (stc.F90:94) / set_3D_int / celndof."""
        norder = _i32(norder)
        v = [C.c_int() for _ in range(4)]
        _lib.check(self.L.hp3d_gpu_sizes_t(self.plan, int(etype), _ptr(norder), *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def sig_dims(self, norder, norient_edge, norient_face, etype=MDLB):
        """dict(ntest, ni, nb, nint, nrdofH, np, nbp, nip) of one element signature (host only)."""
        d = np.zeros(8, np.int32)
        _lib.check(self.L.hp3d_gpu_sig_dims(self.plan, int(etype), _ptr(_i32(norder)), _ptr(_i32(norient_edge)), _ptr(_i32(norient_face)), _ptr(d)))
        return dict(zip(("ntest", "ni", "nb", "nint", "nrdofH", "np", "nbp", "nip"), (int(v) for v in d)))

    def quad_points(self, norder, norient_edge, norient_face, xnod, etype=None):
        norder, noe, nof = _i32(norder).reshape(-1, 19), _i32(norient_edge).reshape(-1, 12), _i32(norient_face).reshape(-1, 6)
        xnod = np.ascontiguousarray(xnod, dtype=np.float64)
        nel = norder.shape[0]
        et = None if etype is None else _i32(np.broadcast_to(etype, (nel,)))
        nint = max(self.sizes(norder[e], MDLB if et is None else et[e])[2] for e in range(nel))
        xq = np.zeros((nel, nint, 3))
        _lib.check(self.L.hp3d_gpu_quad_points(self.plan, nel, _ptr(et), _ptr(norder), _ptr(noe), _ptr(nof), _ptr(xnod),
                                               int(xnod[0].size), _ptr(xq), C.c_longlong(nint * 3)))
        return xq

    def elem_stc_batch(self, norder, norient_edge, norient_face, xnod, source_qp=None, out=None, etype=None):
        """elem + stc_fwd_wrapper for nel elements.

        norder (nel,19), norient_edge (nel,12), norient_face (nel,6): find_order / find_orient output;
        xnod (nel, nrdofH_max, 3): nodcor output (geometry dofs; row k = coordinates of dof k).
        etype: None (all bricks), a scalar, or (nel,) element types MDLB / MDLP; a prism uses the first 15 / 9 / 5 entries
        of its norder / norient_edge / norient_face rows.
        Returns dict(Aii (nel,ni,ni) [Fortran order per element: Aii[e].T is the column-major block], Bi, ASchur, BSchur,
        ni, nb, info).  All elements are padded to the largest ni/nb of the batch.
        """
        L = self.L
        norder, noe, nof = _i32(norder).reshape(-1, 19), _i32(norient_edge).reshape(-1, 12), _i32(norient_face).reshape(-1, 6)
        xnod = np.ascontiguousarray(xnod, dtype=np.float64)
        nel = norder.shape[0]
        assert xnod.shape[0] == nel
        et = None if etype is None else _i32(np.broadcast_to(etype, (nel,)))
        sz = {}
        for e in range(nel):
            t = MDLB if et is None else int(et[e])
            k = bytes([t]) + norder[e].tobytes()
            if k not in sz:
                sz[k] = self.sizes(norder[e], t)
        ni = max((s[0] for s in sz.values()), default=0)
        nb = max((s[1] for s in sz.values()), default=0)
        if out is None:
            nr = max(1, int(self.prm.nr_rhs))   # Bi (ni, NR_RHS), BSchur (nb, NR_RHS) per element, column-major
            out = dict(Aii=np.zeros((nel, self.aii_len(ni)), self.dtype), Bi=np.zeros((nel, ni * nr), self.dtype),
                       ASchur=np.zeros((nel, max(nb * ni, 1)), self.dtype), BSchur=np.zeros((nel, max(nb * nr, 1)), self.dtype))
        Aii, Bi, AS, BS = out["Aii"], out["Bi"], out["ASchur"], out["BSchur"]
        nio = np.zeros(nel, np.int32); nbo = np.zeros(nel, np.int32); info = np.zeros(nel, np.int32)
        src_ld = 0
        if source_qp is not None:
            source_qp = np.ascontiguousarray(source_qp)
            src_ld = source_qp[0].size * (2 if np.iscomplexobj(source_qp) else 1)
        L.hp3d_gpu_elem_batch.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong,
                                          C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]
        rc = L.hp3d_gpu_elem_batch(self.plan, nel, _ptr(et), _ptr(norder), _ptr(noe), _ptr(nof), _ptr(xnod), int(np.prod(xnod.shape[1:])),
                                   _ptr(source_qp), src_ld, _ptr(Aii), int(np.prod(Aii.shape[1:])), _ptr(Bi), int(np.prod(Bi.shape[1:])),
                                   _ptr(AS), int(np.prod(AS.shape[1:])), _ptr(BS), int(np.prod(BS.shape[1:])), _ptr(nio), _ptr(nbo), _ptr(info))
        _lib.check(rc)
        return dict(Aii=Aii, Bi=Bi, ASchur=AS, BSchur=BS, ni=nio, nb=nbo, info=info)

    # ---- device-resident CLOC (STORE_STC = .true. with the factors kept in HBM; include/hp3d_gpu.h "Device-resident CLOC")
    def cloc_create(self, limit_bytes=0):
        h = self.L.hp3d_gpu_cloc_create(self.plan, C.c_longlong(int(limit_bytes)))
        if h < 0:
            _lib.check(h)
        return h

    def cloc_destroy(self, cloc):
        _lib.check(self.L.hp3d_gpu_cloc_destroy(int(cloc)))

    def cloc_clear(self, cloc):
        _lib.check(self.L.hp3d_gpu_cloc_clear(int(cloc)))

    def cloc_stats(self, cloc):
        st = np.zeros(4, np.int64)
        _lib.check(self.L.hp3d_gpu_cloc_stats(int(cloc), _ptr(st)))
        return dict(resident=int(st[0]), spilled=int(st[1]), bytes=int(st[2]), limit=int(st[3]))

    def aii_len(self, ni):
        """scalars of one element's Aii block: ni^2, or ni (ni+1)/2 with aii_packed"""
        return ni * (ni + 1) // 2 if self.prm.aii_packed == 1 else ni * ni

    def elem_stc_batch_cloc(self, cloc, norder, norient_edge, norient_face, xnod, iel=None, source_qp=None, out=None, etype=None):
        """elem + stc_fwd_wrapper with the Schur factors filed in the device-resident store `cloc` under iel[e] (default e):
        only Aii (the packed lower triangle if the plan has aii_packed = 1) and Bi return.  dict(Aii, Bi, ni, nb, info)."""
        norder, noe, nof, xnod, nel, et = self._descr(norder, norient_edge, norient_face, xnod, etype)
        ni = max((self.sizes(norder[e], MDLB if et is None else int(et[e]))[0] for e in range(nel)), default=0) if out is None else 0
        if out is None:
            out = dict(Aii=np.zeros((nel, self.aii_len(ni)), self.dtype), Bi=np.zeros((nel, ni * max(1, int(self.prm.nr_rhs))), self.dtype))
        Aii, Bi = out["Aii"], out["Bi"]
        iel = None if iel is None else np.ascontiguousarray(iel, dtype=np.int64)
        nio = np.zeros(nel, np.int32); nbo = np.zeros(nel, np.int32); info = np.zeros(nel, np.int32)
        src_ld = 0
        if source_qp is not None:
            source_qp = np.ascontiguousarray(source_qp)
            src_ld = source_qp[0].size * (2 if np.iscomplexobj(source_qp) else 1)
        f = self.L.hp3d_gpu_elem_batch_cloc
        ll = C.c_longlong
        f.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_void_p, ll, C.c_void_p, ll, C.c_void_p, ll] + [C.c_void_p] * 3
        _lib.check(f(self.plan, int(cloc), nel, _ptr(iel), _ptr(et), _ptr(norder), _ptr(noe), _ptr(nof), _ptr(xnod), int(np.prod(xnod.shape[1:])),
                     _ptr(source_qp), src_ld, _ptr(Aii), int(np.prod(Aii.shape[1:])), _ptr(Bi), int(np.prod(Bi.shape[1:])), _ptr(nio), _ptr(nbo), _ptr(info)))
        return dict(Aii=Aii, Bi=Bi, ni=nio, nb=nbo, info=info)

    def cloc_bwd_batch(self, cloc, xi, iel=None, nb_max=None):
        """stc_bwd (stc.F90:661-677) on the device-resident factors: xb = BSchur - ASchur xi for the elements iel (default 0..nel-1).
        xi (nel, ni_max).  Returns dict(xb (nel, nb_max), nb, info)."""
        xi = np.ascontiguousarray(xi, dtype=self.dtype)
        nel = xi.shape[0]
        iel = None if iel is None else np.ascontiguousarray(iel, dtype=np.int64)
        if nb_max is None:
            nb_max = 0
            for e in range(nel):
                nb_max = max(nb_max, self.cloc_fetch(cloc, e if iel is None else int(iel[e]), sizes_only=True)[1])
        xb = np.zeros((nel, max(nb_max * max(1, int(self.prm.nr_rhs)), 1)), self.dtype)
        nbo = np.zeros(nel, np.int32); info = np.zeros(nel, np.int32)
        f = self.L.hp3d_gpu_cloc_bwd_batch
        ll = C.c_longlong
        f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, ll, C.c_void_p, ll, C.c_void_p, C.c_void_p]
        _lib.check(f(int(cloc), nel, _ptr(iel), _ptr(xi), xi[0].size if nel else 0, _ptr(xb), xb[0].size if nel else 0, _ptr(nbo), _ptr(info)))
        return dict(xb=xb, nb=nbo, info=info)

    def cloc_fetch(self, cloc, iel, sizes_only=False):
        """(ASchur (nb, ni) as a Fortran-ordered array, BSchur (nb,)) of one stored element, or None if it is spilled;
        sizes_only: (ni, nb) of a resident element, (0, 0) of a spilled one."""
        f = self.L.hp3d_gpu_cloc_fetch
        f.argtypes = [C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        ni, nb = C.c_int(0), C.c_int(0)
        rc = f(int(cloc), int(iel), None, None, C.byref(ni), C.byref(nb))
        if rc < 0:
            _lib.check(rc)
        if sizes_only:
            return (ni.value, nb.value)
        if rc == 1:
            return None
        nr = max(1, int(self.prm.nr_rhs))
        AS = np.zeros(max(ni.value * nb.value, 1), self.dtype); BS = np.zeros(max(nb.value * nr, 1), self.dtype)
        _lib.check(f(int(cloc), int(iel), _ptr(AS), _ptr(BS), None, None))
        return AS[:ni.value * nb.value].reshape(ni.value, nb.value).T, BS[:nb.value * nr]

    def hermitian_unpack(self, AP, ni, out=None, threads=0):
        """Full (nel, ni*ni) column-major Hermitian blocks from packed lower triangles (hp3d_gpu_hermitian_unpack_batch; host only).
        ni: an int or an (nel,) array of per-element sizes."""
        AP = np.ascontiguousarray(AP, dtype=self.dtype)
        nel = AP.shape[0]
        nie = None if np.isscalar(ni) else _i32(ni)
        nmax = int(ni) if nie is None else int(nie.max(initial=0))
        if out is None:
            out = np.zeros((nel, nmax * nmax), self.dtype)
        f = self.L.hp3d_gpu_hermitian_unpack_batch
        ll = C.c_longlong
        f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, ll, C.c_void_p, ll, C.c_int]
        _lib.check(f(int(self.complex), nel, nmax, _ptr(nie), _ptr(AP), int(np.prod(AP.shape[1:])), _ptr(out), int(np.prod(out.shape[1:])), int(threads)))
        return out

    def _descr(self, norder, norient_edge, norient_face, xnod, etype):
        norder, noe, nof = _i32(norder).reshape(-1, 19), _i32(norient_edge).reshape(-1, 12), _i32(norient_face).reshape(-1, 6)
        xnod = np.ascontiguousarray(xnod, dtype=np.float64)
        nel = norder.shape[0]
        et = None if etype is None else _i32(np.broadcast_to(etype, (nel,)))
        return norder, noe, nof, xnod, nel, et

    def elem_bwd_batch(self, norder, norient_edge, norient_face, xnod, xi, etype=None):
        """stc_bwd without stored factors (hp3d_gpu_elem_bwd_batch): recompute the elements on the device and return
        xb = BSchur - ASchur xi, (nel, nb_max) [+ nb per element].  xi: (nel, ni_max) interface dofs (rows of Aii)."""
        norder, noe, nof, xnod, nel, et = self._descr(norder, norient_edge, norient_face, xnod, etype)
        xi = np.ascontiguousarray(xi, dtype=self.dtype)
        nb = max(self.sizes(norder[e], MDLB if et is None else int(et[e]))[1] for e in range(nel))
        xb = np.zeros((nel, max(nb * max(1, int(self.prm.nr_rhs)), 1)), self.dtype)
        nbo = np.zeros(nel, np.int32); info = np.zeros(nel, np.int32)
        f = self.L.hp3d_gpu_elem_bwd_batch
        ll = C.c_longlong
        f.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p, ll, C.c_void_p, ll, C.c_void_p, ll, C.c_void_p, C.c_void_p]
        _lib.check(f(self.plan, nel, _ptr(et), _ptr(norder), _ptr(noe), _ptr(nof), _ptr(xnod), int(xnod[0].size), None, 0,
                     _ptr(xi), xi[0].size, _ptr(xb), xb[0].size, _ptr(nbo), _ptr(info)))
        return dict(xb=xb, nb=nbo, info=info)

    def elem_residual_batch(self, norder, norient_edge, norient_face, xnod, xi, xb, etype=None):
        """DPG element residuals ||l - B u||^2_{V'} (hp3d_gpu_elem_residual_batch) for u = (xi | xb); returns (nel,) float64."""
        norder, noe, nof, xnod, nel, et = self._descr(norder, norient_edge, norient_face, xnod, etype)
        xi = np.ascontiguousarray(xi, dtype=self.dtype); xb = np.ascontiguousarray(xb, dtype=self.dtype)
        res = np.zeros(nel); info = np.zeros(nel, np.int32)
        f = self.L.hp3d_gpu_elem_residual_batch
        ll = C.c_longlong
        f.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p, ll, C.c_void_p, ll, C.c_void_p, ll, C.c_void_p, C.c_void_p]
        _lib.check(f(self.plan, nel, _ptr(et), _ptr(norder), _ptr(noe), _ptr(nof), _ptr(xnod), int(xnod[0].size), None, 0,
                     _ptr(xi), xi[0].size, _ptr(xb), xb[0].size, _ptr(res), _ptr(info)))
        return dict(resid=res, info=info)

    def celem_batch(self, norder, norient_edge, norient_face, xnod, cons, isym_flag=2, want_coo=False, want_schur=False,
                    source_qp=None, etype=None, out=None, packed=None, cloc=None, iel=None):
        """elem + stc_fwd_wrapper + the rest of celem_systemI (constraints, Dirichlet lift, compression; :543-785) for nel
        elements (hp3d_gpu_celem_batch).  cons: one dict per element with the flat lists of `celem_pack` (cptr, cidx, cval) and
        idbc, zdofd (Nrdofm,), nextract (Nrdofc,) [1-based], optionally lcon (Nrdofc,) global dof numbers.
        out: optional dict of preallocated (e.g. pinned) result arrays zbload/zastif/irn/jcn/ASchur/BSchur;
        packed: the dict `pack_constraints(cons, isym_flag, want_coo)` returns, to keep the Python-side packing out of a timed call.
        Returns dict(zbload, zastif [flat, element e at aptr[e]], xptr, aptr, irn, jcn, ASchur, BSchur, ni, nb, info)."""
        norder, noe, nof, xnod, nel, et = self._descr(norder, norient_edge, norient_face, xnod, etype)
        pk = packed if packed is not None else self.pack_constraints(cons, isym_flag, want_coo)
        mptr, xptr, aptr, cptr = pk["mptr"], pk["xptr"], pk["aptr"], pk["cptr"]
        assert len(mptr) == nel + 1
        cidx, cval, idbc, zdofd, nextract, lcon = pk["cidx"], pk["cval"], pk["idbc"], pk["zdofd"], pk["nextract"], pk["lcon"]
        out = out or {}
        zb = out.get("zbload"); za = out.get("zastif"); irn = out.get("irn"); jcn = out.get("jcn")
        if zb is None:
            zb = np.zeros(max(int(xptr[-1]), 1), self.dtype)
        if za is None:
            za = np.zeros(max(int(aptr[-1]), 1), self.dtype)
        if want_coo and irn is None:
            irn = np.zeros(max(int(aptr[-1]), 1), np.int32); jcn = np.zeros(max(int(aptr[-1]), 1), np.int32)
        AS, BS = out.get("ASchur"), out.get("BSchur")
        sAS = sBS = 0
        if want_schur and AS is None:
            szs = [self.sizes(norder[e], MDLB if et is None else int(et[e])) for e in range(nel)]
            ni = max(s_[0] for s_ in szs); nb = max(s_[1] for s_ in szs)
            AS = np.zeros((nel, max(nb * ni, 1)), self.dtype); BS = np.zeros((nel, max(nb, 1)), self.dtype)
        if want_schur:
            sAS, sBS = AS[0].size, BS[0].size
        else:
            AS = BS = None
        nio = np.zeros(nel, np.int32); nbo = np.zeros(nel, np.int32); info = np.zeros(nel, np.int32)
        src_ld = 0
        if source_qp is not None:
            source_qp = np.ascontiguousarray(source_qp)
            src_ld = source_qp[0].size * (2 if np.iscomplexobj(source_qp) else 1)
        ll = C.c_longlong
        if cloc is not None:   # Schur factors into the device-resident store (hp3d_gpu_celem_batch_cloc): nothing but the compressed systems returns
            iel = None if iel is None else np.ascontiguousarray(iel, dtype=np.int64)
            g = self.L.hp3d_gpu_celem_batch_cloc
            g.argtypes = ([C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_void_p, ll] + [C.c_void_p] * 9 + [C.c_int] + [C.c_void_p] * 5 + [C.c_void_p] * 3)
            _lib.check(g(self.plan, int(cloc), nel, _ptr(iel), _ptr(et), _ptr(norder), _ptr(noe), _ptr(nof), _ptr(xnod), int(np.prod(xnod.shape[1:])), _ptr(source_qp), src_ld,
                         _ptr(mptr), _ptr(cptr), _ptr(cidx), _ptr(cval), _ptr(idbc), _ptr(zdofd), _ptr(xptr), _ptr(nextract), _ptr(lcon), int(isym_flag),
                         _ptr(aptr), _ptr(zb), _ptr(za), _ptr(irn), _ptr(jcn), _ptr(nio), _ptr(nbo), _ptr(info)))
            return dict(zbload=zb, zastif=za, xptr=xptr, aptr=aptr, irn=irn, jcn=jcn, ASchur=None, BSchur=None, ni=nio, nb=nbo, info=info)
        f = self.L.hp3d_gpu_celem_batch
        f.argtypes = ([C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p, ll] + [C.c_void_p] * 9 + [C.c_int] + [C.c_void_p] * 5 +
                      [C.c_void_p, ll, C.c_void_p, ll] + [C.c_void_p] * 3)
        _lib.check(f(self.plan, nel, _ptr(et), _ptr(norder), _ptr(noe), _ptr(nof), _ptr(xnod), int(np.prod(xnod.shape[1:])), _ptr(source_qp), src_ld,
                     _ptr(mptr), _ptr(cptr), _ptr(cidx), _ptr(cval), _ptr(idbc), _ptr(zdofd), _ptr(xptr), _ptr(nextract), _ptr(lcon), int(isym_flag),
                     _ptr(aptr), _ptr(zb), _ptr(za), _ptr(irn), _ptr(jcn), _ptr(AS), sAS, _ptr(BS), sBS, _ptr(nio), _ptr(nbo), _ptr(info)))
        return dict(zbload=zb, zastif=za, xptr=xptr, aptr=aptr, irn=irn, jcn=jcn, ASchur=AS, BSchur=BS, ni=nio, nb=nbo, info=info)

    def pack_constraints(self, cons, isym_flag=2, want_coo=False):
        """Concatenate per-element constraint dicts into the flat arrays of hp3d_gpu_celem_batch (what the Fortran shim builds)."""
        nel = len(cons)
        mptr = np.zeros(nel + 1, np.int64); xptr = np.zeros(nel + 1, np.int64); aptr = np.zeros(nel + 1, np.int64)
        for e, c in enumerate(cons):
            nm, nc = len(c["idbc"]), len(c["nextract"])
            mptr[e + 1] = mptr[e] + nm; xptr[e + 1] = xptr[e] + nc
            aptr[e + 1] = aptr[e] + (nc * (nc + 1) // 2 if isym_flag == 1 else nc * nc)
        if nel and all("cptr" not in c for c in cons):   # regular mesh: the library builds the identity lists itself
            cat0 = lambda k, dt: np.ascontiguousarray(np.concatenate([np.asarray(c[k]).ravel() for c in cons]), dtype=dt)  # noqa: E731
            return dict(mptr=mptr, xptr=xptr, aptr=aptr, cptr=None, cidx=None, cval=None, idbc=cat0("idbc", np.int32), zdofd=cat0("zdofd", self.dtype),
                        nextract=cat0("nextract", np.int32), lcon=cat0("lcon", np.int32) if want_coo else None)
        cptr = np.zeros(mptr[-1] + 1, np.int64)
        base = 0
        for e, c in enumerate(cons):
            cp = np.asarray(c["cptr"], np.int64)
            cptr[mptr[e]:mptr[e + 1] + 1] = base + cp
            base += int(cp[-1])
        cat = lambda k, dt: (np.ascontiguousarray(np.concatenate([np.asarray(c[k]).ravel() for c in cons]), dtype=dt) if nel else np.zeros(0, dt))  # noqa: E731
        cidx, cval = cat("cidx", np.int32), cat("cval", np.float64)
        idbc, zdofd, nextract = cat("idbc", np.int32), cat("zdofd", self.dtype), cat("nextract", np.int32)
        lcon = cat("lcon", np.int32) if want_coo else None
        return dict(mptr=mptr, xptr=xptr, aptr=aptr, cptr=cptr, cidx=cidx, cval=cval, idbc=idbc, zdofd=zdofd, nextract=nextract, lcon=lcon)

    def elem_error_batch(self, norder, norient_edge, norient_face, xnod, zdof, exact_qp=None, l2proj=False, etype=None):
        """element_error (compute_error.F90:226) for the field variable of the problem (hp3d_gpu_elem_error_batch).
        zdof (nel, nrdof_max, ncomp): the variable's dofs, interface dofs first then the middle node's; exact_qp: None (built-in
        manufactured solution) or (nel, nint_max, nvals) exact values at `error_points`.  Returns dict(err, rnorm, info)."""
        norder, noe, nof, xnod, nel, et = self._descr(norder, norient_edge, norient_face, xnod, etype)
        zdof = np.ascontiguousarray(zdof, dtype=self.dtype)
        ex = None if exact_qp is None else np.ascontiguousarray(exact_qp, dtype=self.dtype)
        err = np.zeros(nel); rn = np.zeros(nel); info = np.zeros(nel, np.int32)
        f = self.L.hp3d_gpu_elem_error_batch
        ll = C.c_longlong
        f.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p, ll, C.c_void_p, ll, C.c_int] + [C.c_void_p] * 3
        _lib.check(f(self.plan, nel, _ptr(et), _ptr(norder), _ptr(noe), _ptr(nof), _ptr(xnod), int(np.prod(xnod.shape[1:])),
                     _ptr(zdof), int(np.prod(zdof.shape[1:])) if nel else 0, _ptr(ex), 0 if ex is None else int(np.prod(ex.shape[1:])), int(bool(l2proj)),
                     _ptr(err), _ptr(rn), _ptr(info)))
        return dict(err=err, rnorm=rn, info=info)

    def error_points(self, norder, norient_edge, norient_face, xnod, etype=None):
        """Physical coordinates of element_error's quadrature points: (xq (nel, nint_max, 3), nint (nel,))."""
        norder, noe, nof, xnod, nel, et = self._descr(norder, norient_edge, norient_face, xnod, etype)
        nint = np.zeros(nel, np.int32)
        f = self.L.hp3d_gpu_error_points
        f.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_longlong, C.c_void_p]
        _lib.check(f(self.plan, nel, _ptr(et), _ptr(norder), _ptr(noe), _ptr(nof), _ptr(xnod), int(np.prod(xnod.shape[1:])), None, 0, _ptr(nint)))
        xq = np.zeros((nel, int(nint.max()) if nel else 0, 3))
        _lib.check(f(self.plan, nel, _ptr(et), _ptr(norder), _ptr(noe), _ptr(nof), _ptr(xnod), int(np.prod(xnod.shape[1:])), _ptr(xq),
                     int(np.prod(xq.shape[1:])), _ptr(nint)))
        return xq, nint

    @staticmethod
    def unpack(res, e):
        """Element e of an elem_stc_batch result as (Aii (ni,ni), Bi (ni), ASchur (nb,ni), BSchur (nb)) numpy arrays."""
        ni, nb = int(res["ni"][e]), int(res["nb"][e])
        Aii = res["Aii"][e][: ni * ni].reshape(ni, ni).T
        AS = res["ASchur"][e][: nb * ni].reshape(ni, nb).T
        return Aii, res["Bi"][e][:ni], AS, res["BSchur"][e][:nb]

    def stc_bwd_batch(self, ASchur, BSchur, xi):
        """xb = BSchur - ASchur xi (stc_bwd, stc.F90:661-677).  ASchur (nel,nb,ni) logical, passed column-major per element."""
        nel, nb, ni = ASchur.shape
        A = np.ascontiguousarray(np.transpose(ASchur, (0, 2, 1)), dtype=self.dtype)   # per element column-major
        B = np.ascontiguousarray(BSchur, dtype=self.dtype); x = np.ascontiguousarray(xi, dtype=self.dtype)
        xb = np.zeros((nel, nb), self.dtype)
        ll = C.c_longlong
        self.L.hp3d_gpu_stc_bwd_batch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, ll, C.c_void_p, ll, C.c_void_p, ll, C.c_void_p, ll]
        _lib.check(self.L.hp3d_gpu_stc_bwd_batch(int(self.complex), nel, ni, nb, _ptr(A), nb * ni, _ptr(B), nb, _ptr(x), ni, _ptr(xb), nb))
        return xb

    def integrate_debug(self, norder, norient_edge, norient_face, xnod, source_qp=None, etype=MDLB):
        """Raw dense-phase input of one element (test hook): returns (W (planes,R,np), dims dict)."""
        norder, noe, nof = _i32(norder), _i32(norient_edge), _i32(norient_face)
        xnod = np.ascontiguousarray(xnod, dtype=np.float64)
        dims = np.zeros(8, np.int32)
        g = self.L.hp3d_gpu_integrate_debug_t
        g.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]
        f = lambda plan, *a: g(plan, int(etype), *a)   # noqa: E731
        _lib.check(f(self.plan, _ptr(norder), _ptr(noe), _ptr(nof), _ptr(xnod), _ptr(source_qp), None, 0, _ptr(dims)))
        np_, nbp, nip, n, nb, ni, R, P = [int(v) for v in dims]
        ld = np_ if np_ else R
        W = np.zeros((P, R, ld))
        _lib.check(f(self.plan, _ptr(norder), _ptr(noe), _ptr(nof), _ptr(xnod), _ptr(source_qp), _ptr(W), W.size, _ptr(dims)))
        return W, dict(np=np_, nbp=nbp, nip=nip, n=n, nb=nb, ni=ni, R=R, planes=P)

    def bench(self, norder, norient_edge, norient_face, xnod, reps=1, max_chunk=0, lanes=2, etype=None):
        """Device-resident throughput run (hp3d_gpu_bench): returns dict(ms_total, ms_integ, ms_dense, launches)."""
        norder, noe, nof = _i32(norder).reshape(-1, 19), _i32(norient_edge).reshape(-1, 12), _i32(norient_face).reshape(-1, 6)
        xnod = np.ascontiguousarray(xnod, dtype=np.float64)
        nel = norder.shape[0]
        t = [C.c_double() for _ in range(3)]
        ln = C.c_longlong()
        et = None if etype is None else _i32(np.broadcast_to(etype, (nel,)))
        _lib.check(self.L.hp3d_gpu_bench_t(self.plan, nel, _ptr(et), _ptr(norder), _ptr(noe), _ptr(nof), _ptr(xnod), int(xnod[0].size), int(reps),
                                         int(max_chunk), int(lanes), C.byref(t[0]), C.byref(t[1]), C.byref(t[2]), C.byref(ln)))
        return dict(ms_total=t[0].value, ms_integ=t[1].value, ms_dense=t[2].value, launches=ln.value)


# ---- H1 projection-based interpolation: update_gdof / update_Ddof (src/hpinterp), SURVEY 8f row f4 -----------------------------
def _pbi_descr(norder, norient_edge, norient_face, etype):
    norder = _i32(norder).reshape(-1, 19); noe = _i32(norient_edge).reshape(-1, 12); nof = _i32(norient_face).reshape(-1, 6)
    nel = norder.shape[0]
    et = None if etype is None else _i32(np.broadcast_to(np.asarray(etype, np.int32), (nel,)))
    return norder, noe, nof, nel, et


def pbi_points(norder, norient_edge, norient_face, integration=0, maxp=9, etype=None):
    """Where the host evaluates the interpolated function (hp3d_gpu_pbi_points; host only, no GPU needed).
    Returns dict(xi (nel, npts_max, 3) master coordinates, npts (nel,), nrdofH (nel,), nodes (nel, 27, 4) = first dof, # dofs,
    first point, # points of every node (vertices, edges, faces, middle))."""
    norder, noe, nof, nel, et = _pbi_descr(norder, norient_edge, norient_face, etype)
    L = _lib.lib()
    f = L.hp3d_gpu_pbi_points
    f.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_void_p, C.c_longlong] + [C.c_void_p] * 3
    npts = np.zeros(nel, np.int32); nH = np.zeros(nel, np.int32); nodes = np.zeros((nel, 27, 4), np.int32)
    _lib.check(f(nel, _ptr(et), _ptr(norder), _ptr(noe), _ptr(nof), int(integration), int(maxp), None, 0, _ptr(npts), _ptr(nH), _ptr(nodes)))
    xi = np.zeros((nel, int(npts.max()) if nel else 0, 3))
    _lib.check(f(nel, _ptr(et), _ptr(norder), _ptr(noe), _ptr(nof), int(integration), int(maxp), _ptr(xi), int(np.prod(xi.shape[1:])), None, None, None))
    return dict(xi=xi, npts=npts, nrdofH=nH, nodes=nodes)


def pbi_h1_batch(norder, norient_edge, norient_face, etav, fvert, fgrad, integration=0, maxp=9, mask=None, dof=None, etype=None, device=0):
    """hpvert/hpedge/hpface_opt/hpmdle_opt (integration=0, the GMP map) or dhpvert/dhpedgeH/dhpfaceH_opt (integration=1, Dirichlet
    data) for all elements at once (hp3d_gpu_pbi_h1_batch).
    etav (nel, 8, 3); fvert (nel, 8, ncomp); fgrad (nel, npts_max, 3, ncomp) = d g_c / d eta_i at `pbi_points`; mask (nel,) node bits
    or None; dof (nel, nrdofH_max, ncomp) incoming dofs of the unselected nodes or None.  Returns dict(dof, info)."""
    norder, noe, nof, nel, et = _pbi_descr(norder, norient_edge, norient_face, etype)
    L = _lib.lib()
    _lib.check(L.hp3d_gpu_init(int(device)))
    etav = np.ascontiguousarray(etav, dtype=np.float64).reshape(nel, 8, 3)
    fvert = np.ascontiguousarray(fvert, dtype=np.float64)
    ncomp = int(fvert.shape[-1])
    fvert = fvert.reshape(nel, 8, ncomp)
    fgrad = np.ascontiguousarray(fgrad, dtype=np.float64).reshape(nel, -1, 3, ncomp)
    nH = np.zeros(nel, np.int32)
    fp = L.hp3d_gpu_pbi_points
    fp.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_void_p, C.c_longlong] + [C.c_void_p] * 3
    _lib.check(fp(nel, _ptr(et), _ptr(norder), _ptr(noe), _ptr(nof), int(integration), int(maxp), None, 0, None, _ptr(nH), None))
    nHmax = int(nH.max()) if nel else 0
    out = np.zeros((nel, nHmax, ncomp)) if dof is None else np.ascontiguousarray(dof, dtype=np.float64).reshape(nel, -1, ncomp).copy()
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint32)
    info = np.zeros(nel, np.int32)
    f = L.hp3d_gpu_pbi_h1_batch
    ll = C.c_longlong
    f.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, ll, C.c_void_p, C.c_void_p, ll, C.c_void_p]
    _lib.check(f(nel, _ptr(et), _ptr(norder), _ptr(noe), _ptr(nof), int(integration), int(maxp), _ptr(etav), ncomp, _ptr(fvert), _ptr(fgrad),
                 int(np.prod(fgrad.shape[1:])), _ptr(m), _ptr(out), int(np.prod(out.shape[1:])), _ptr(info)))
    return dict(dof=out, info=info)


def pbi_hcurl_points(norder, norient_edge, norient_face, maxp=9, etype=None):
    """Points of the H(curl) Dirichlet interpolation (hp3d_gpu_pbi_hcurl_points; host only): dict(xi, npts, nrdofE, nodes)."""
    norder, noe, nof, nel, et = _pbi_descr(norder, norient_edge, norient_face, etype)
    f = _lib.lib().hp3d_gpu_pbi_hcurl_points
    f.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_longlong] + [C.c_void_p] * 3
    npts = np.zeros(nel, np.int32); nE = np.zeros(nel, np.int32); nodes = np.zeros((nel, 27, 4), np.int32)
    _lib.check(f(nel, _ptr(et), _ptr(norder), _ptr(noe), _ptr(nof), int(maxp), None, 0, _ptr(npts), _ptr(nE), _ptr(nodes)))
    xi = np.zeros((nel, int(npts.max()) if nel else 0, 3))
    _lib.check(f(nel, _ptr(et), _ptr(norder), _ptr(noe), _ptr(nof), int(maxp), _ptr(xi), int(np.prod(xi.shape[1:])), None, None, None))
    return dict(xi=xi, npts=npts, nrdofE=nE, nodes=nodes)


def pbi_hcurl_batch(norder, norient_edge, norient_face, etav, fval, fcurl, maxp=9, mask=None, dof=None, etype=None, device=0):
    """dhpedgeE / dhpfaceE_opt for all elements at once (hp3d_gpu_pbi_hcurl_batch).  fval, fcurl (nel, npts_max, 3, ncomp): the datum
    and its curl pulled back to eta at `pbi_hcurl_points`; dof (nel, nrdofE_max, ncomp) incoming dofs or None.  Returns dict(dof, info)."""
    norder, noe, nof, nel, et = _pbi_descr(norder, norient_edge, norient_face, etype)
    L = _lib.lib()
    _lib.check(L.hp3d_gpu_init(int(device)))
    etav = np.ascontiguousarray(etav, dtype=np.float64).reshape(nel, 8, 3)
    fval = np.ascontiguousarray(fval, dtype=np.float64)
    ncomp = int(fval.shape[-1])
    fval = fval.reshape(nel, -1, 3, ncomp)
    fcurl = np.ascontiguousarray(fcurl, dtype=np.float64).reshape(fval.shape)
    nE = pbi_hcurl_points(norder, noe, nof, maxp=maxp, etype=et)["nrdofE"]
    out = np.zeros((nel, int(nE.max()) if nel else 0, ncomp)) if dof is None else np.ascontiguousarray(dof, dtype=np.float64).reshape(nel, -1, ncomp).copy()
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint32)
    info = np.zeros(nel, np.int32)
    f = L.hp3d_gpu_pbi_hcurl_batch
    ll = C.c_longlong
    f.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, ll, C.c_void_p, C.c_void_p, ll, C.c_void_p]
    _lib.check(f(nel, _ptr(et), _ptr(norder), _ptr(noe), _ptr(nof), int(maxp), _ptr(etav), ncomp, _ptr(fval), _ptr(fcurl), int(np.prod(fval.shape[1:])),
                 _ptr(m), _ptr(out), int(np.prod(out.shape[1:])), _ptr(info)))
    return dict(dof=out, info=info)


def pbi_hdiv_points(norder, norient_edge, norient_face, maxp=9, etype=None):
    """Points of the H(div) Dirichlet interpolation (hp3d_gpu_pbi_hdiv_points; host only): dict(xi, npts, nrdofV, nodes)."""
    norder, noe, nof, nel, et = _pbi_descr(norder, norient_edge, norient_face, etype)
    f = _lib.lib().hp3d_gpu_pbi_hdiv_points
    f.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_longlong] + [C.c_void_p] * 3
    npts = np.zeros(nel, np.int32); nV = np.zeros(nel, np.int32); nodes = np.zeros((nel, 27, 4), np.int32)
    _lib.check(f(nel, _ptr(et), _ptr(norder), _ptr(noe), _ptr(nof), int(maxp), None, 0, _ptr(npts), _ptr(nV), _ptr(nodes)))
    xi = np.zeros((nel, int(npts.max()) if nel else 0, 3))
    _lib.check(f(nel, _ptr(et), _ptr(norder), _ptr(noe), _ptr(nof), int(maxp), _ptr(xi), int(np.prod(xi.shape[1:])), None, None, None))
    return dict(xi=xi, npts=npts, nrdofV=nV, nodes=nodes)


def pbi_hdiv_batch(norder, norient_edge, norient_face, etav, fval, maxp=9, mask=None, dof=None, etype=None, device=0):
    """dhpfaceV_opt for all elements at once (hp3d_gpu_pbi_hdiv_batch).  fval (nel, npts_max, 3, ncomp): the datum pulled back to eta
    (Piola) at `pbi_hdiv_points`.  Returns dict(dof (nel, nrdofV_max, ncomp), info)."""
    norder, noe, nof, nel, et = _pbi_descr(norder, norient_edge, norient_face, etype)
    L = _lib.lib()
    _lib.check(L.hp3d_gpu_init(int(device)))
    etav = np.ascontiguousarray(etav, dtype=np.float64).reshape(nel, 8, 3)
    fval = np.ascontiguousarray(fval, dtype=np.float64)
    ncomp = int(fval.shape[-1])
    fval = fval.reshape(nel, -1, 3, ncomp)
    nV = pbi_hdiv_points(norder, noe, nof, maxp=maxp, etype=et)["nrdofV"]
    out = np.zeros((nel, int(nV.max()) if nel else 0, ncomp)) if dof is None else np.ascontiguousarray(dof, dtype=np.float64).reshape(nel, -1, ncomp).copy()
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint32)
    info = np.zeros(nel, np.int32)
    f = L.hp3d_gpu_pbi_hdiv_batch
    ll = C.c_longlong
    f.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_int, C.c_void_p, ll, C.c_void_p, C.c_void_p, ll, C.c_void_p]
    _lib.check(f(nel, _ptr(et), _ptr(norder), _ptr(noe), _ptr(nof), int(maxp), _ptr(etav), ncomp, _ptr(fval), int(np.prod(fval.shape[1:])),
                 _ptr(m), _ptr(out), int(np.prod(out.shape[1:])), _ptr(info)))
    return dict(dof=out, info=info)
