"""ctypes front-end of the CPU oracle (oracle/libhp3d_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (hp3d_b200) never imports this module.
"""
import ctypes as C
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

POIS_GAL, POIS_PDPG, MAXW_GAL, MAXW_UW = 1, 2, 3, 4
MDLB, MDLP = 1, 3   # element types (src/modules/node_types.F90:8-10)


class Params(C.Structure):
    _fields_ = [("nord_add", C.c_int), ("test_norm", C.c_int), ("alpha_norm", C.c_double),
                ("omega", C.c_double), ("eps", C.c_double), ("mu", C.c_double), ("sigma", C.c_double),
                ("eps_tensor", C.c_double * 18), ("source", C.c_int), ("icomp_exact", C.c_int),
                ("source_table", C.c_void_p)]


def build(force=False):
    so = os.path.join(_HERE, "libhp3d_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("shape.c", "shape_prism.c", "quad_geom.c", "etype.c", "tri_rules.h", "dense.c", "elem.c", "celem.c", "soleval.c", "pbi.c", "hp3d_oracle.h", "dense.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libhp3d_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def find_openblas():
    """The OpenBLAS bundled with scipy in this image (LP64, symbols prefixed scipy_)."""
    try:
        import scipy
        cands = glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs", "libscipy_openblas*.so"))
        cands = [c for c in cands if "64_" not in os.path.basename(c)]
        return cands[0] if cands else None
    except Exception:
        return None


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_dense_use_blas.argtypes = [C.c_char_p]
        L.orc_dense_use_blas.restype = C.c_int
        _LIB = L
    return _LIB


def use_blas(on=True, threads=1):
    L = lib()
    if not on:
        L.orc_dense_use_blas(None)
        return False
    path = find_openblas()
    ok = bool(path) and bool(L.orc_dense_use_blas(path.encode()))
    if ok:
        L.orc_dense_set_threads(int(threads))
    return ok


def _ip(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _pad(a, n):
    """descriptor arrays keep the brick layout (19/12/6); a prism uses the first 15/9/5 entries"""
    a = _ip(a).ravel()
    if a.size >= n:
        return a
    out = np.zeros(n, dtype=np.int32)
    out[:a.size] = a
    return out


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _d(a):
    return a.ctypes.data_as(C.c_void_p)


def default_params(**kw):
    p = Params()
    lib().orc_params_default(C.byref(p))
    for k, v in kw.items():
        if k == "eps_tensor":
            t = np.asarray(v, dtype=np.complex128).reshape(3, 3).T.copy().view(np.float64).ravel()  # column-major
            for i in range(18):
                p.eps_tensor[i] = t[i]
        else:
            setattr(p, k, v)
    return p


def set_maxp(maxp):
    lib().orc_set_maxp(int(maxp))


def uniform_order(p, etype=MDLB, pz=None):
    """norder(19) of an isotropic order-p hexa / an order-(p, pz) prism (find_order.F90:41-58 layout)."""
    if etype == MDLP:
        pz = p if pz is None else pz
        return np.array([p] * 6 + [pz] * 3 + [p] * 2 + [10 * p + pz] * 3 + [10 * p + pz] + [0] * 4, dtype=np.int32)
    return np.array([p] * 12 + [11 * p] * 6 + [111 * p], dtype=np.int32)


def mid_index(etype):
    return 14 if etype == MDLP else 18


def celndof(norder, etype=MDLB):
    h, e, v, q = (C.c_int() for _ in range(4))
    lib().orc_celndof(int(etype), _i(_pad(norder, 19)), C.byref(h), C.byref(e), C.byref(v), C.byref(q))
    return h.value, e.value, v.value, q.value


def ndof_mdl(nord, etype=MDLB):
    h, e, v, q = (C.c_int() for _ in range(4))
    lib().orc_ndof_nod_mid(int(etype), int(nord), C.byref(h), C.byref(e), C.byref(v), C.byref(q))
    return h.value, e.value, v.value, q.value


def gauss1(n):
    x = np.zeros(n)
    w = np.zeros(n)
    lib().orc_gauss1(n, _d(x), _d(w))
    return x, w


def shape3DH(xi, norder, norie, norif, etype=MDLB):
    nH = celndof(norder, etype)[0]
    s = np.zeros(nH)
    g = np.zeros((nH, 3))
    xi = np.ascontiguousarray(xi, dtype=np.float64)
    n = lib().orc_shape3DH(int(etype), _d(xi), _i(_pad(norder, 19)), _i(_pad(norie, 12)), _i(_pad(norif, 6)), _d(s), _d(g))
    assert n == nH
    return s, g


def shape3DE(xi, norder, norie, norif, etype=MDLB):
    nE = celndof(norder, etype)[1]
    s = np.zeros((nE, 3))
    c = np.zeros((nE, 3))
    xi = np.ascontiguousarray(xi, dtype=np.float64)
    n = lib().orc_shape3DE(int(etype), _d(xi), _i(_pad(norder, 19)), _i(_pad(norie, 12)), _i(_pad(norif, 6)), _d(s), _d(c))
    assert n == nE
    return s, c


def shape3DV(xi, norder, norif, etype=MDLB):
    nV = celndof(norder, etype)[2]
    s = np.zeros((nV, 3))
    d = np.zeros(nV)
    xi = np.ascontiguousarray(xi, dtype=np.float64)
    n = lib().orc_shape3DV(int(etype), _d(xi), _i(_pad(norder, 19)), _i(_pad(norif, 6)), _d(s), _d(d))
    assert n == nV
    return s, d


def shape3DQ(xi, norder, etype=MDLB):
    nQ = celndof(norder, etype)[3]
    s = np.zeros(nQ)
    xi = np.ascontiguousarray(xi, dtype=np.float64)
    n = lib().orc_shape3DQ(int(etype), _d(xi), _i(_pad(norder, 19)), _d(s))
    assert n == nQ
    return s


def shape3HH(xi, nordM, etype=MDLB):
    n = celndof(enriched_order(nordM, etype), etype)[0]
    s = np.zeros(n)
    g = np.zeros((n, 3))
    xi = np.ascontiguousarray(xi, dtype=np.float64)
    m = lib().orc_shape3HH(int(etype), _d(xi), int(nordM), _d(s), _d(g))
    assert m == n
    return s, g


def shape3EE(xi, nordM, etype=MDLB):
    n = celndof(enriched_order(nordM, etype), etype)[1]
    s = np.zeros((n, 3))
    c = np.zeros((n, 3))
    xi = np.ascontiguousarray(xi, dtype=np.float64)
    m = lib().orc_shape3EE(int(etype), _d(xi), int(nordM), _d(s), _d(c))
    assert m == n
    return s, c


def quad3(norder, norif, integration, maxp, etype=MDLB):
    xi = np.zeros((1000, 3))
    w = np.zeros(1000)
    n = lib().orc_set_3D_int(int(etype), _i(_pad(norder, 19)), _i(_pad(norif, 6)), int(integration), int(maxp), _d(xi), _d(w))
    return xi[:n].copy(), w[:n].copy()


def stc_partition(kind, norder, etype=MDLB):
    perm = np.zeros(8192, dtype=np.int32)
    ni, nb = C.c_int(), C.c_int()
    r = lib().orc_stc_partition_t(int(etype), int(kind), _i(_pad(norder, 19)), _i(perm), C.byref(ni), C.byref(nb))
    assert r == 0
    return perm[: ni.value + nb.value].copy(), ni.value, nb.value


def elem(kind, norder, norie, norif, xnod, prm, want_dpg=False, etype=MDLB):
    """Full (uncondensed) local matrix/load of one element, reference dof ordering. xnod: (nrdofH,3)."""
    L = lib()
    norder, norie, norif = _pad(norder, 19), _pad(norie, 12), _pad(norif, 6)
    xnod = np.ascontiguousarray(xnod, dtype=np.float64)
    et = int(etype)
    nH, nE, nV, nQ = celndof(norder, et)
    bH, bE, bV, bQ = ndof_mdl(int(norder[mid_index(et)]), et)
    a1, a2 = C.c_int(), C.c_int()
    if kind == POIS_GAL:
        n = nH
        A = np.zeros((n, n), order="F"); b = np.zeros(n)
        r = L.orc_elem_poisson_galerkin_t(et, _i(norder), _i(norie), _i(norif), _d(xnod), C.byref(prm), _d(A), _d(b), C.byref(a1))
    elif kind == POIS_PDPG:
        n = nH + nV - bV
        A = np.zeros((n, n), order="F"); b = np.zeros(n)
        r = L.orc_elem_poisson_primal_dpg_t(et, _i(norder), _i(norie), _i(norif), _d(xnod), C.byref(prm), _d(A), _d(b), C.byref(a1), C.byref(a2))
    elif kind == MAXW_GAL:
        n = nE
        A = np.zeros((n, n), order="F", dtype=np.complex128); b = np.zeros(n, dtype=np.complex128)
        r = L.orc_elem_maxwell_galerkin_t(et, _i(norder), _i(norie), _i(norif), _d(xnod), C.byref(prm), _d(A), _d(b), C.byref(a1))
    elif kind == MAXW_UW:
        n = 2 * (nE - bE) + 6 * nQ
        A = np.zeros((n, n), order="F", dtype=np.complex128); b = np.zeros(n, dtype=np.complex128)
        gram = stiff = None
        gp = sp = None
        if want_dpg:
            dp = prm.nord_add
            nEE = celndof(enriched_order(int(norder[mid_index(et)]) + dp * (11 if et == MDLP else 111), et), et)[1]
            gram = np.zeros((2 * nEE, 2 * nEE), order="F", dtype=np.complex128)
            stiff = np.zeros((2 * nEE, n + 1), order="F", dtype=np.complex128)
            gp, sp = _d(gram), _d(stiff)
        r = L.orc_elem_maxwell_uw_dpg_t(et, _i(norder), _i(norie), _i(norif), _d(xnod), C.byref(prm), _d(A), _d(b), C.byref(a1), C.byref(a2), gp, sp)
        if want_dpg:
            assert r == 0
            return A, b, gram, stiff
    else:
        raise ValueError(kind)
    assert r == 0, r
    return A, b


def elem_uw_scalar(norder, norie, norif, xnod, prm, etype=MDLB):
    """The reference's scalar-loop twin of the ultraweak Maxwell element (elem_maxwell.F90, oracle/elem.c:
    orc_elem_maxwell_uw_scalar_t) -> (A, b, Gram [upper triangle filled], enriched stiffness [B | l])."""
    norder, norie, norif = _pad(norder, 19), _pad(norie, 12), _pad(norif, 6)
    xnod = np.ascontiguousarray(xnod, dtype=np.float64)
    et = int(etype)
    nH, nE, nV, nQ = celndof(norder, et)
    bE = ndof_mdl(int(norder[mid_index(et)]), et)[1]
    n = 2 * (nE - bE) + 6 * nQ
    nEE = celndof(enriched_order(int(norder[mid_index(et)]) + prm.nord_add * (11 if et == MDLP else 111), et), et)[1]
    A = np.zeros((n, n), order="F", dtype=np.complex128); b = np.zeros(n, dtype=np.complex128)
    gram = np.zeros((2 * nEE, 2 * nEE), order="F", dtype=np.complex128)
    stiff = np.zeros((2 * nEE, n + 1), order="F", dtype=np.complex128)
    r = lib().orc_elem_maxwell_uw_scalar_t(et, _i(norder), _i(norie), _i(norif), _d(xnod), C.byref(prm), _d(A), _d(b), _d(gram), _d(stiff))
    assert r == 0, r
    return A, b, gram, stiff


def enriched_order(nordP, etype=MDLB):
    no = np.zeros(19, dtype=np.int32)
    lib().orc_compute_enriched_order(int(etype), int(nordP), _i(no))
    return no


def condensed(kind, norder, norie, norif, xnod, prm, etype=MDLB):
    """elem + stc_fwd_wrapper for one element -> Aii, Bi, ASchur, BSchur."""
    L = lib()
    norder, norie, norif = _pad(norder, 19), _pad(norie, 12), _pad(norif, 6)
    xnod = np.ascontiguousarray(xnod, dtype=np.float64)
    _, ni, nb = stc_partition(kind, norder, etype)
    dt = np.complex128 if kind >= 3 else np.float64
    Aii = np.zeros((ni, ni), order="F", dtype=dt); Bi = np.zeros(ni, dtype=dt)
    AS = np.zeros((nb, ni), order="F", dtype=dt); BS = np.zeros(nb, dtype=dt)
    a, b = C.c_int(), C.c_int()
    r = L.orc_condensed_element_t(int(etype), int(kind), _i(norder), _i(norie), _i(norif), _d(xnod), C.byref(prm), _d(Aii), _d(Bi), _d(AS), _d(BS), C.byref(a), C.byref(b))
    assert r == 0, r
    return Aii, Bi, AS, BS


def condensed_batch(kind, norder, norie, norif, xnod, prm, nthreads=1):
    """OpenMP element loop (par_mumps_sc.F90:318-357 shape).  norder (nel,19), xnod (nel,nrdofH,3)."""
    L = lib()
    norder, norie, norif = _ip(norder), _ip(norie), _ip(norif)
    xnod = np.ascontiguousarray(xnod, dtype=np.float64)
    nel = norder.shape[0]
    _, ni, nb = stc_partition(kind, norder[0])
    dt = np.complex128 if kind >= 3 else np.float64
    Aii = np.zeros((nel, ni, ni), dtype=dt); Bi = np.zeros((nel, ni), dtype=dt)
    AS = np.zeros((nel, ni, nb), dtype=dt); BS = np.zeros((nel, nb), dtype=dt)
    info = np.zeros(nel, dtype=np.int32)
    L.orc_condensed_batch.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_long, C.c_long,
                                      C.c_void_p, C.c_int]
    bad = L.orc_condensed_batch(int(kind), nel, _d(norder), _d(norie), _d(norif), _d(xnod), int(xnod[0].size), C.byref(prm),
                                _d(Aii), _d(Bi), _d(AS), _d(BS), ni * ni, ni, ni * nb, nb, _d(info), int(nthreads))
    # per-element blocks are column-major: expose them as (nel, rows, cols)
    return (np.transpose(Aii, (0, 2, 1)), Bi, np.transpose(AS.reshape(nel, ni, nb), (0, 2, 1)), BS, info, bad)


# ---- celem_systemI after the element (celem.c): constraints, Dirichlet lift, compression, COO fill ----------------
class Physics(C.Structure):
    _fields_ = [("nphys", C.c_int), ("dtype", C.c_int * 8), ("ncomp", C.c_int * 8), ("adres", C.c_int * 8), ("active", C.c_int * 8),
                ("nrvar", C.c_int * 3)]


def physics_of(kind):
    """Physics table of the four problems (problems/<PROB>/input/physics): D_TYPE 0 contin, 1 tangen, 2 normal, 3 discon."""
    table = {POIS_GAL: [(0, 1)], POIS_PDPG: [(0, 1), (2, 1)], MAXW_GAL: [(1, 1)], MAXW_UW: [(1, 2), (3, 6)]}[kind]
    ph = Physics()
    ph.nphys = len(table)
    nvar = [0, 0, 0, 0]
    for i, (dt, nc) in enumerate(table):
        ph.dtype[i], ph.ncomp[i], ph.adres[i], ph.active[i] = dt, nc, nvar[dt], int(dt != 3)
        nvar[dt] += nc
    for f in range(3):
        ph.nrvar[f] = nvar[f]
    return ph


def celem_modify(ph, nrdofl, nrcon, nac, constr, nrdofm_f, A, b, idbc, zdofd, nextract, isym, want_zamod=False):
    """celem_systemI.F90:543-785 on the condensed system (A (ni,ni), b (ni)); nrcon/nac/constr: per family (H,E,V) arrays
    (nk,), (nk,nacdim), (nk,nacdim) as `logic` returns them (1-based nac).  Returns zbload, zastif[, zamod]."""
    L = lib()
    ni = A.shape[0]
    cplx = np.iscomplexobj(A) or np.iscomplexobj(zdofd)
    Ac = np.asfortranarray(A, dtype=np.complex128); bc = np.ascontiguousarray(b, dtype=np.complex128)
    nacdim = max([np.asarray(x).shape[1] for x in nac if np.asarray(x).ndim == 2 and np.asarray(x).size] + [1])
    keep = []

    def fam(arrs, dt, two_d):
        out = (C.c_void_p * 3)()
        for f in range(3):
            a = np.asarray(arrs[f], dtype=dt)
            if two_d:
                full = np.zeros((a.shape[0] if a.ndim == 2 else 0, nacdim), dtype=dt)
                if a.size:
                    full[:, :a.shape[1]] = a
                a = full
            a = np.ascontiguousarray(a)
            keep.append(a)
            out[f] = a.ctypes.data if a.size else None
        return out
    p_nrcon, p_nac, p_con = fam(nrcon, np.int32, False), fam(nac, np.int32, True), fam(constr, np.float64, True)
    nrdofl = _ip(nrdofl); nrdofm_f = _ip(nrdofm_f)
    nrdofm = int(nrdofm_f.sum())
    idbc = _ip(idbc); zd = np.ascontiguousarray(zdofd, dtype=np.complex128); nx = _ip(nextract)
    nc = nx.size
    zb = np.zeros(nc, np.complex128)
    za = np.zeros(nc * (nc + 1) // 2 if isym == 1 else nc * nc, np.complex128)
    zm = np.zeros((nrdofm, nrdofm), np.complex128, order="F") if want_zamod else None
    r = L.orc_celem_modify(C.byref(ph), _i(nrdofl), p_nrcon, p_nac, p_con, int(nacdim), _i(nrdofm_f), int(ni), _d(Ac), _d(bc), _i(idbc), _d(zd),
                           int(nc), _i(nx), int(isym), _d(zb), _d(za), _d(zm) if want_zamod else None)
    assert r == 0, r
    if not cplx:
        zb, za = zb.real.copy(), za.real.copy()
        zm = zm.real.copy() if want_zamod else None
    return (zb, za, zm) if want_zamod else (zb, za)


def coo_fill(lcon, ztemp, zload, ndof_global):
    """par_mumps_sc.F90:419-448 for one element: (A_loc, IRN, JCN, RHS contribution)."""
    lcon = _ip(lcon)
    n = lcon.size
    zt = np.ascontiguousarray(ztemp, dtype=np.complex128); zl = np.ascontiguousarray(zload, dtype=np.complex128)
    a = np.zeros(n * n, np.complex128); irn = np.zeros(n * n, np.int32); jcn = np.zeros(n * n, np.int32)
    rhs = np.zeros(ndof_global, np.complex128)
    lib().orc_coo_fill(int(n), _i(lcon), _d(zt), _d(zl), _d(a), _i(irn), _i(jcn), _d(rhs))
    return a, irn, jcn, rhs


# ---- solution evaluation / element error (soleval.c): soleval.F90, compute_error.F90:226 -------------------------------------
def error_nvals(kind):
    return int(lib().orc_error_nvals(int(kind)))


def element_error(kind, norder, norie, norif, xnod, zdof, prm, exact_tab=None, l2proj=False, etype=MDLB):
    """element_error for the field variable of problem `kind`; zdof (nrdof, ncomp).  Returns (err, rnorm, nint)."""
    norder, norie, norif = _pad(norder, 19), _pad(norie, 12), _pad(norif, 6)
    xnod = np.ascontiguousarray(xnod, dtype=np.float64)
    z = np.ascontiguousarray(zdof, dtype=np.complex128)
    tab = None if exact_tab is None else np.ascontiguousarray(exact_tab, dtype=np.complex128)
    e, r = C.c_double(), C.c_double()
    n = lib().orc_element_error(int(etype), int(kind), _i(norder), _i(norie), _i(norif), _d(xnod), _d(z), C.byref(prm),
                                None if tab is None else _d(tab), int(bool(l2proj)), C.byref(e), C.byref(r))
    assert n > 0, n
    return e.value, r.value, n


def error_points(norder, norie, norif, xnod, etype=MDLB):
    norder, norie, norif = _pad(norder, 19), _pad(norie, 12), _pad(norif, 6)
    xnod = np.ascontiguousarray(xnod, dtype=np.float64)
    xq = np.zeros((2000, 3))
    n = lib().orc_error_points(int(etype), _i(norder), _i(norie), _i(norif), _d(xnod), _d(xq))
    return xq[:n].copy()


def exact_field(kind, prm, x):
    v = np.zeros(error_nvals(kind), np.complex128)
    x = np.ascontiguousarray(x, dtype=np.float64)
    lib().orc_exact_field(int(kind), C.byref(prm), _d(x), _d(v))
    return v


# ---- H1 projection-based interpolation (pbi.c): hpvert/hpedge/hpface_opt/hpmdle_opt, dhpvert/dhpedgeH/dhpfaceH_opt ------------
PBI_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)


def pbi_offsets(norder, etype=MDLB):
    """offsets of the nodes' H1 dofs (vertices, edges, faces, middle) + total"""
    off = np.zeros(32, np.int32)
    lib().orc_pbi_offsets(int(etype), _i(_pad(norder, 19)), _i(off))
    nn = (6 + 9 + 5 + 1) if etype == MDLP else 27
    return off[:nn + 1].copy()


def pbi_element(norder, norie, norif, etav, fun, ncomp, integration=0, maxp=9, mask=None, dof=None, etype=MDLB):
    """PB interpolation of fun(eta) -> (val[ncomp], dval[ncomp, 3]) on one element; returns dof (nrdofH, ncomp)."""
    norder, norie, norif = _pad(norder, 19), _pad(norie, 12), _pad(norif, 6)
    etav = np.ascontiguousarray(etav, dtype=np.float64)
    off = pbi_offsets(norder, etype)
    nn = off.size - 1
    out = np.zeros((int(off[-1]), ncomp)) if dof is None else np.ascontiguousarray(dof, dtype=np.float64).copy()
    mask = (1 << nn) - 1 if mask is None else int(mask)

    def cb(eta, val, dval, ctx):
        v, dv = fun(np.array([eta[0], eta[1], eta[2]]))
        v = np.atleast_1d(v); dv = np.asarray(dv).reshape(ncomp, 3)
        for c in range(ncomp):
            val[c] = v[c]
            for i in range(3):
                dval[c + ncomp * i] = dv[c, i]
    r = lib().orc_pbi_element(int(etype), _i(norder), _i(norie), _i(norif), _d(etav), int(ncomp), int(integration), int(maxp),
                              C.c_uint(mask), PBI_FN(cb), None, _d(out))
    assert r == 0, r
    return out


PBI_FNE = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)


def pbi_offsets_E(norder, etype=MDLB):
    """offsets of the edge and face nodes' H(curl) dofs + their total"""
    off = np.zeros(32, np.int32)
    lib().orc_pbi_offsets_E(int(etype), _i(_pad(norder, 19)), _i(off))
    nn = (9 + 5) if etype == MDLP else 18
    return off[:nn + 1].copy()


def pbi_hcurl_element(norder, norie, norif, etav, fun, ncomp, maxp=9, mask=None, dof=None, etype=MDLB):
    """dhpedgeE / dhpfaceE_opt on one element.  fun(eta) -> (E[ncomp, 3], curlE[ncomp, 3], dxdeta[3, 3]) in physical components.
    Returns dofE (n_edge_and_face_dofs, ncomp)."""
    norder, norie, norif = _pad(norder, 19), _pad(norie, 12), _pad(norif, 6)
    etav = np.ascontiguousarray(etav, dtype=np.float64)
    off = pbi_offsets_E(norder, etype)
    nn = off.size - 1
    out = np.zeros((int(off[-1]), ncomp)) if dof is None else np.ascontiguousarray(dof, dtype=np.float64).copy()
    mask = (1 << nn) - 1 if mask is None else int(mask)

    def cb(eta, E, cE, dxdeta, ctx):
        e, c, J = fun(np.array([eta[0], eta[1], eta[2]]))
        e = np.asarray(e).reshape(ncomp, 3); c = np.asarray(c).reshape(ncomp, 3); J = np.asarray(J).reshape(3, 3)
        for n in range(ncomp):
            for j in range(3):
                E[n + ncomp * j] = e[n, j]; cE[n + ncomp * j] = c[n, j]
        for j in range(3):
            for i in range(3):
                dxdeta[j + 3 * i] = J[j, i]
    r = lib().orc_pbi_hcurl_element(int(etype), _i(norder), _i(norie), _i(norif), _d(etav), int(ncomp), int(maxp), C.c_uint(mask), PBI_FNE(cb),
                                    None, _d(out))
    assert r == 0, r
    return out


def pbi_offsets_V(norder, etype=MDLB):
    off = np.zeros(8, np.int32)
    lib().orc_pbi_offsets_V(int(etype), _i(_pad(norder, 19)), _i(off))
    return off[:(5 if etype == MDLP else 6) + 1].copy()


def pbi_hdiv_element(norder, norie, norif, etav, fun, ncomp, maxp=9, mask=None, dof=None, etype=MDLB):
    """dhpfaceV_opt on one element.  fun(eta) -> (V[ncomp, 3], ignored, dxdeta[3, 3]).  Returns dofV (n_face_dofs, ncomp)."""
    norder, norie, norif = _pad(norder, 19), _pad(norie, 12), _pad(norif, 6)
    etav = np.ascontiguousarray(etav, dtype=np.float64)
    off = pbi_offsets_V(norder, etype)
    out = np.zeros((int(off[-1]), ncomp)) if dof is None else np.ascontiguousarray(dof, dtype=np.float64).copy()
    mask = (1 << (off.size - 1)) - 1 if mask is None else int(mask)

    def cb(eta, E, cE, dxdeta, ctx):
        e, _, J = fun(np.array([eta[0], eta[1], eta[2]]))
        e = np.asarray(e).reshape(ncomp, 3); J = np.asarray(J).reshape(3, 3)
        for n in range(ncomp):
            for j in range(3):
                E[n + ncomp * j] = e[n, j]; cE[n + ncomp * j] = 0.0
        for j in range(3):
            for i in range(3):
                dxdeta[j + 3 * i] = J[j, i]
    r = lib().orc_pbi_hdiv_element(int(etype), _i(norder), _i(norie), _i(norif), _d(etav), int(ncomp), int(maxp), C.c_uint(mask), PBI_FNE(cb),
                                   None, _d(out))
    assert r == 0, r
    return out
