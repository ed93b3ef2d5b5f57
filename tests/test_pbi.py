"""H1 projection-based interpolation on the device (hp3d_gpu_pbi_h1_batch; SURVEY 8f row f4, interpolation half) against the
oracle's restatement of hpvert/hpedge/hpface_opt/hpmdle_opt and dhpvert/dhpedgeH/dhpfaceH_opt (oracle/pbi.c).
The oracle evaluates the interpolated function through a callback at the points it visits, like the reference calls the GMP
routines; the device path gets the same function tabulated at hp3d_gpu_pbi_points -- a different point order or a different
rule breaks the comparison.  Tolerance: 1e-11 relative to the largest dof of the element (two Cholesky orderings of
stiffness matrices with condition numbers up to ~1e5)."""
import numpy as np
import pytest

from hp3d_b200 import api, synth
from tests.test_pbi_oracle import poly, smooth

MDLB, MDLP = 1, 3
BRICK_M = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
PRISM_M = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1]], float)
TOL = 1e-11


def vertex_shape(et, xi):
    x, y, z = xi
    if et == MDLB:
        return np.array([(x if m[0] else 1 - x) * (y if m[1] else 1 - y) * (z if m[2] else 1 - z) for m in BRICK_M])
    lam = [1 - x - y, x, y]
    return np.array([lam[v % 3] * (z if v >= 3 else 1 - z) for v in range(6)])


def tabulate(fun, ncomp, pts, etav, etype):
    """fvert (nel, 8, ncomp), fgrad (nel, npts_max, 3, ncomp) of fun at the product's points"""
    nel = etav.shape[0]
    fv = np.zeros((nel, 8, ncomp)); fg = np.zeros((nel, pts["xi"].shape[1], 3, ncomp))
    for e in range(nel):
        et = int(etype[e]); nv = 8 if et == MDLB else 6
        for v in range(nv):
            fv[e, v] = fun(etav[e, v])[0]
        for l in range(int(pts["npts"][e])):
            eta = vertex_shape(et, pts["xi"][e, l]) @ etav[e, :nv]
            fg[e, l] = np.asarray(fun(eta)[1]).reshape(ncomp, 3).T
    return fv, fg


def random_brick(rng, lo=2, hi=5):
    no = np.array(list(rng.integers(lo, hi + 1, 12)) + [10 * int(rng.integers(lo, hi + 1)) + int(rng.integers(lo, hi + 1)) for _ in range(6)]
                  + [100 * int(rng.integers(lo, hi + 1)) + 10 * int(rng.integers(lo, hi + 1)) + int(rng.integers(lo, hi + 1))], np.int32)
    return no, rng.integers(0, 2, 12).astype(np.int32), rng.integers(0, 8, 6).astype(np.int32)


def random_prism(rng, lo=2, hi=5):
    no = np.zeros(19, np.int32); noe = np.zeros(12, np.int32); nof = np.zeros(6, np.int32)
    no[:9] = rng.integers(lo, hi + 1, 9)
    no[9:11] = rng.integers(lo, hi + 1, 2)
    no[11:14] = [10 * int(rng.integers(lo, hi + 1)) + int(rng.integers(lo, hi + 1)) for _ in range(3)]
    no[14] = 10 * int(rng.integers(lo, hi + 1)) + int(rng.integers(lo, hi + 1))
    noe[:9] = rng.integers(0, 2, 9); nof[:2] = rng.integers(0, 6, 2); nof[2:5] = rng.integers(0, 8, 3)
    return no, noe, nof


def warped_vertices(rng, et):
    M = BRICK_M if et == MDLB else PRISM_M
    out = np.zeros((8, 3))
    out[:len(M)] = np.array([0.2, 0.1, 0.3]) + M * np.array([0.6, 0.5, 0.4]) + rng.uniform(-0.05, 0.05, M.shape)
    return out


# ---------------------------------------------------------------------------------------------------------------- host logic (CPU)
@pytest.mark.parametrize("integration", [0, 1])
def test_points_are_the_ones_the_reference_loops_visit(oracle, gpulib, integration):
    """the callback of the oracle records eta in the order hpvert / hpedge / hpface / hpmdle evaluate the GMP map"""
    oracle.set_maxp(9)
    rng = np.random.default_rng(1)
    for et, gen in ((MDLB, random_brick), (MDLP, random_prism)):
        for _ in range(3):
            no, noe, nof = gen(rng, 1, 5)   # order-1 nodes have no dofs and must contribute no points
            etav = warped_vertices(rng, et)
            nv = 8 if et == MDLB else 6
            seen = []

            def fun(eta):
                seen.append(eta.copy())
                return np.zeros(1), np.zeros((1, 3))
            oracle.pbi_element(no, noe, nof, etav[:nv], fun, 1, integration=integration, etype=et)
            pts = api.pbi_points(no, noe, nof, integration=integration, etype=et)
            n = int(pts["npts"][0])
            assert len(seen) == nv + n
            mine = np.array([vertex_shape(et, x) @ etav[:nv] for x in pts["xi"][0, :n]])
            assert np.abs(np.array(seen[nv:]) - mine).max() < 1e-14
            off = oracle.pbi_offsets(no, et)
            nn = len(off) - 1
            assert np.array_equal(pts["nodes"][0, :nn, 0], off[:-1]) and int(pts["nrdofH"][0]) == int(off[-1])
            assert np.array_equal(pts["nodes"][0, :nn, 1], np.diff(off))


def test_points_reject_bad_descriptors(gpulib):
    no = synth.uniform_order(3).copy(); no[3] = 12
    with pytest.raises(RuntimeError, match="edge order"):
        api.pbi_points(no, np.zeros(12), np.zeros(6))


# ------------------------------------------------------------------------------------------------------------------ device parity
def _compare(oracle, fun, ncomp, norder, noe, nof, etav, etype, integration, mask=None, dof_in=None):
    nel = norder.shape[0]
    pts = api.pbi_points(norder, noe, nof, integration=integration, etype=etype)
    fv, fg = tabulate(fun, ncomp, pts, etav, etype)
    res = api.pbi_h1_batch(norder, noe, nof, etav, fv, fg, integration=integration, mask=mask, dof=dof_in, etype=etype)
    assert not res["info"].any()
    worst = 0.0
    for e in range(nel):
        et = int(etype[e]); nv = 8 if et == MDLB else 6; nH = int(pts["nrdofH"][e])
        ref = oracle.pbi_element(norder[e], noe[e], nof[e], etav[e, :nv], fun, ncomp, integration=integration, etype=et,
                                 mask=None if mask is None else int(mask[e]), dof=None if dof_in is None else dof_in[e, :nH])
        worst = max(worst, float(np.abs(res["dof"][e, :nH] - ref).max() / max(1.0, np.abs(ref).max())))
    return worst, res, pts


@pytest.mark.gpu
@pytest.mark.parametrize("integration", [0, 1])
def test_gpu_matches_oracle_random_elements(oracle, gpu, integration):
    """bricks and prisms, anisotropic orders 2..5, all orientation codes, warped (non-affine) vertex coordinates, smooth g"""
    oracle.set_maxp(9)
    rng = np.random.default_rng(7)
    rows = [(MDLB,) + random_brick(rng) for _ in range(4)] + [(MDLP,) + random_prism(rng) for _ in range(4)]
    rows += [rows[0], rows[5]]   # two elements sharing a signature: grouped into one launch
    etype = np.array([r[0] for r in rows], np.int32)
    norder = np.array([r[1] for r in rows]); noe = np.array([r[2] for r in rows]); nof = np.array([r[3] for r in rows])
    etav = np.array([warped_vertices(rng, int(t)) for t in etype])
    worst, _, _ = _compare(oracle, smooth, 2, norder, noe, nof, etav, etype, integration)
    assert worst < TOL


@pytest.mark.gpu
def test_gpu_geometry_dofs_on_hp_mesh(oracle, gpu):
    """update_gdof on the mixed hexa/prism hp mesh of config 4 (orders 2..6): a curved GMP map x(eta), 3 components, every
    element its own signature; parity with the oracle and identical dofs on shared entities"""
    from tests.mini_fem_hp import build_space
    oracle.set_maxp(9)
    m = synth.hp_mesh(2, prism_frac=0.45, pmin=2, pmax=6, seed_p=21, seed_g=9)
    nel = len(m["etype"])
    etav = np.zeros((nel, 8, 3))
    for e in range(nel):
        nv = 8 if m["etype"][e] == MDLB else 6
        etav[e, :nv] = m["xnod"][e, :nv]

    def gmp(eta):   # a smooth curved block
        x, y, z = eta
        v = np.array([x + 0.1 * np.sin(2.0 * y) * z, y + 0.05 * x * x, z + 0.1 * np.cos(x + y)])
        d = np.array([[1.0, 0.2 * np.cos(2.0 * y) * z, 0.1 * np.sin(2.0 * y)], [0.1 * x, 1.0, 0.0], [-0.1 * np.sin(x + y), -0.1 * np.sin(x + y), 1.0]])
        return v, d
    worst, res, pts = _compare(oracle, gmp, 3, m["norder"], m["norient_edge"], m["norient_face"], etav, m["etype"], 0)
    assert worst < TOL
    keys, l2g, nloc, _, _ = build_space(m)
    U = np.full((len(keys), 3), np.nan); spread = 0.0
    for e in range(nel):
        for k, g in enumerate(l2g[e]):
            if np.isnan(U[g, 0]):
                U[g] = res["dof"][e, k]
            else:
                spread = max(spread, float(np.abs(U[g] - res["dof"][e, k]).max()))
    assert spread < 1e-11


@pytest.mark.gpu
def test_gpu_dirichlet_mask_and_incoming_dofs(oracle, gpu):
    """update_Ddof: only the nodes of one face (its vertices, edges and the face node) are interpolated; the other entries of
    dof are left exactly as they came"""
    oracle.set_maxp(9)
    rng = np.random.default_rng(11)
    no, noe, nof = random_brick(rng, 3, 4)
    etav = warped_vertices(rng, MDLB)[None]
    face = 3   # face 4 (0-based 3): vertices 2,3,7,6 ; edges 2,11,6,10 (1-based)
    mask = 0
    for v in (1, 2, 6, 5):
        mask |= 1 << v
    for e1 in (2, 11, 6, 10):
        mask |= 1 << (8 + e1 - 1)
    mask |= 1 << (8 + 12 + face)
    pts = api.pbi_points(no, noe, nof, integration=1)
    nH = int(pts["nrdofH"][0])
    dof_in = rng.standard_normal((1, nH, 2))
    worst, res, _ = _compare(oracle, smooth, 2, no[None], noe[None], nof[None], etav, np.array([MDLB], np.int32), 1,
                             mask=np.array([mask], np.uint32), dof_in=dof_in)
    assert worst < TOL
    touched = np.zeros(nH, bool)
    for i in range(27):
        if mask >> i & 1:
            t0, n = pts["nodes"][0, i, 0], pts["nodes"][0, i, 1]
            touched[t0:t0 + n] = True
    assert np.array_equal(res["dof"][0, ~touched], dof_in[0, ~touched])
    assert not np.array_equal(res["dof"][0, touched], dof_in[0, touched])


@pytest.mark.gpu
def test_gpu_polynomial_reproduction_full_size(oracle, gpu):
    """size-independent property at the largest orders of config 4 (p = 7 bricks, 216-dof middle nodes; p = 6 prisms), 256
    elements in one call: a polynomial of the space is reproduced -- dofs evaluated back through the oracle's shape3DH"""
    oracle.set_maxp(9)
    rng = np.random.default_rng(5)
    nel = 256
    nb = synth.uniform_order(7)
    npz = oracle.uniform_order(6, MDLP, 6)
    sigs = []
    for et, no in ((MDLB, nb), (MDLP, npz)):
        for _ in range(2):
            noe = np.zeros(12, np.int32); nof = np.zeros(6, np.int32)
            if et == MDLB:
                noe[:] = rng.integers(0, 2, 12); nof[:] = rng.integers(0, 8, 6)
            else:
                noe[:9] = rng.integers(0, 2, 9); nof[:2] = rng.integers(0, 6, 2); nof[2:5] = rng.integers(0, 8, 3)
            sigs.append((et, no, noe, nof))
    pick = rng.integers(0, len(sigs), nel)
    etype = np.array([sigs[i][0] for i in pick], np.int32)
    norder = np.array([sigs[i][1] for i in pick]); noe = np.array([sigs[i][2] for i in pick]); nof = np.array([sigs[i][3] for i in pick])
    etav = np.zeros((nel, 8, 3))
    for e in range(nel):   # affine images of the master element (random boxes / sheared prisms)
        M = BRICK_M if etype[e] == MDLB else PRISM_M
        A = np.diag(rng.uniform(0.2, 0.6, 3))
        if etype[e] == MDLP:
            A[0, 1] = rng.uniform(-0.1, 0.1)
        etav[e, :len(M)] = rng.uniform(0, 0.4, 3) + M @ A.T
    pts = api.pbi_points(norder, noe, nof, integration=0, etype=etype)
    # tabulate per signature and element (vectorised over points)
    fv = np.zeros((nel, 8, 3)); fg = np.zeros((nel, pts["xi"].shape[1], 3, 3))
    for e in range(nel):
        et = int(etype[e]); nv = 8 if et == MDLB else 6
        for v in range(nv):
            fv[e, v] = poly(etav[e, v])[0]
        n = int(pts["npts"][e])
        S = np.array([vertex_shape(et, x) for x in pts["xi"][e, :n]])
        eta = S @ etav[e, :nv]
        for l in range(n):
            fg[e, l] = poly(eta[l])[1].T
    res = api.pbi_h1_batch(norder, noe, nof, etav, fv, fg, integration=0, etype=etype)
    assert not res["info"].any()
    worst = 0.0
    for e in rng.choice(nel, 24, replace=False):
        et = int(etype[e]); nv = 8 if et == MDLB else 6; nH = int(pts["nrdofH"][e])
        for _ in range(4):
            xi = rng.random(3)
            if et == MDLP and xi[0] + xi[1] > 1:
                xi[:2] = 1 - xi[:2]
            s, _ = oracle.shape3DH(xi, norder[e], noe[e], nof[e], et)
            worst = max(worst, float(np.abs(s @ res["dof"][e, :nH] - poly(s[:nv] @ etav[e, :nv])[0]).max()))
    assert worst < 1e-11


@pytest.mark.gpu
def test_gpu_pbi_errors_are_loud(gpu):
    no = synth.uniform_order(2)
    z = np.zeros
    with pytest.raises(RuntimeError, match="ncomp"):
        api.pbi_h1_batch(no, z(12), z(6), z((1, 8, 3)), z((1, 8, 13)), z((1, 50, 3, 13)))
    with pytest.raises(RuntimeError, match="fgrad_ld"):
        api.pbi_h1_batch(no, z(12), z(6), BRICK_M[None], z((1, 8, 1)), z((1, 2, 3, 1)))
    # a degenerate element (all vertices equal): info = -1, no exception
    pts = api.pbi_points(no, z(12), z(6))
    res = api.pbi_h1_batch(no, z(12), z(6), z((1, 8, 3)), z((1, 8, 1)), z((1, int(pts["npts"][0]), 3, 1)))
    assert res["info"][0] != 0


# ------------------------------------------------------------------------------------------- H(curl) Dirichlet dofs (dhpedgeE, dhpfaceE_opt)
from tests.test_pbi_oracle import nedelec_poly, smooth_E  # noqa: E402


def curved_E(eta):
    """smooth datum on a curved GMP block: physical components, curl, dx/deta (not the identity: both pullbacks matter)"""
    E, cE, _ = smooth_E(eta)
    x, y, z = eta
    J = np.array([[1.0, 0.2 * np.cos(2.0 * y) * z, 0.1 * np.sin(2.0 * y)], [0.1 * x, 1.0, 0.0], [-0.1 * np.sin(x + y), -0.1 * np.sin(x + y), 1.0]])
    return E, cE, J


def tabulate_E(fun, ncomp, pts, etav, etype):
    """what the Fortran shim does: E_eta = dxdeta^T E, curl_eta = det(dxdeta) dxdeta^-1 curl E at the product's points"""
    nel = etav.shape[0]
    fv = np.zeros((nel, pts["xi"].shape[1], 3, ncomp)); fc = np.zeros_like(fv)
    for e in range(nel):
        et = int(etype[e]); nv = 8 if et == MDLB else 6
        for l in range(int(pts["npts"][e])):
            eta = vertex_shape(et, pts["xi"][e, l]) @ etav[e, :nv]
            E, cE, J = fun(eta)
            E = np.asarray(E).reshape(ncomp, 3); cE = np.asarray(cE).reshape(ncomp, 3)
            fv[e, l] = (E @ J).T
            fc[e, l] = (np.linalg.det(J) * np.linalg.solve(J, cE.T))
    return fv, fc


@pytest.mark.parametrize("et", [MDLB, MDLP])
def test_hcurl_points_are_the_ones_the_reference_loops_visit(oracle, gpulib, et):
    oracle.set_maxp(9)
    rng = np.random.default_rng(2)
    for _ in range(3):
        no, noe, nof = (random_brick if et == MDLB else random_prism)(rng, 1, 4)
        etav = warped_vertices(rng, et)
        nv = 8 if et == MDLB else 6
        seen = []

        def fun(eta):
            seen.append(eta.copy())
            return np.zeros((1, 3)), np.zeros((1, 3)), np.eye(3)
        oracle.pbi_hcurl_element(no, noe, nof, etav[:nv], fun, 1, etype=et)
        pts = api.pbi_hcurl_points(no, noe, nof, etype=et)
        n = int(pts["npts"][0])
        assert len(seen) == n
        mine = np.array([vertex_shape(et, x) @ etav[:nv] for x in pts["xi"][0, :n]])
        assert np.abs(np.array(seen) - mine).max() < 1e-14
        off = oracle.pbi_offsets_E(no, et)
        nn = len(off) - 1
        assert np.array_equal(pts["nodes"][0, nv:nv + nn, 0], off[:-1]) and int(pts["nrdofE"][0]) == int(off[-1])


def _compare_E(oracle, fun, ncomp, norder, noe, nof, etav, etype, mask=None, dof_in=None):
    nel = norder.shape[0]
    pts = api.pbi_hcurl_points(norder, noe, nof, etype=etype)
    fv, fc = tabulate_E(fun, ncomp, pts, etav, etype)
    res = api.pbi_hcurl_batch(norder, noe, nof, etav, fv, fc, mask=mask, dof=dof_in, etype=etype)
    assert not res["info"].any()
    worst = 0.0
    for e in range(nel):
        et = int(etype[e]); nv = 8 if et == MDLB else 6; nE = int(pts["nrdofE"][e])
        om = None if mask is None else (int(mask[e]) >> nv)   # the oracle numbers edges, then faces
        ref = oracle.pbi_hcurl_element(norder[e], noe[e], nof[e], etav[e, :nv], fun, ncomp, etype=et, mask=om,
                                       dof=None if dof_in is None else dof_in[e, :nE])
        worst = max(worst, float(np.abs(res["dof"][e, :nE] - ref).max() / max(1.0, np.abs(ref).max())))
    return worst, res, pts


@pytest.mark.gpu
def test_gpu_hcurl_matches_oracle_random_elements(oracle, gpu):
    """bricks and prisms, anisotropic orders 1..5 (order-1 edges and (2,1) faces own H(curl) dofs but no H1 bubbles), all
    orientation codes, warped vertex coordinates, curved GMP map, two components"""
    oracle.set_maxp(9)
    rng = np.random.default_rng(17)
    rows = [(MDLB,) + random_brick(rng, 1, 5) for _ in range(4)] + [(MDLP,) + random_prism(rng, 1, 5) for _ in range(4)]
    rows += [rows[1], rows[6]]
    etype = np.array([r[0] for r in rows], np.int32)
    norder = np.array([r[1] for r in rows]); noe = np.array([r[2] for r in rows]); nof = np.array([r[3] for r in rows])
    etav = np.array([warped_vertices(rng, int(t)) for t in etype])
    worst, _, _ = _compare_E(oracle, curved_E, 2, norder, noe, nof, etav, etype)
    assert worst < 1e-10


@pytest.mark.gpu
def test_gpu_hcurl_mask_and_incoming_dofs(oracle, gpu):
    """only one face and two of its edges are interpolated; the face projection uses the incoming dofs of the other two edges"""
    oracle.set_maxp(9)
    rng = np.random.default_rng(19)
    no, noe, nof = random_brick(rng, 2, 4)
    etav = warped_vertices(rng, MDLB)[None]
    mask = (1 << (8 + 2 - 1)) | (1 << (8 + 11 - 1)) | (1 << (8 + 12 + 3))   # edges 2, 11 and face 4
    pts = api.pbi_hcurl_points(no, noe, nof)
    nE = int(pts["nrdofE"][0])
    dof_in = rng.standard_normal((1, nE, 2))
    worst, res, _ = _compare_E(oracle, curved_E, 2, no[None], noe[None], nof[None], etav, np.array([MDLB], np.int32),
                               mask=np.array([mask], np.uint32), dof_in=dof_in)
    assert worst < 1e-10
    touched = np.zeros(nE, bool)
    for i in range(27):
        if mask >> i & 1:
            t0, n = pts["nodes"][0, i, 0], pts["nodes"][0, i, 1]
            touched[t0:t0 + n] = True
    assert np.array_equal(res["dof"][0, ~touched], dof_in[0, ~touched])


@pytest.mark.gpu
def test_gpu_hcurl_polynomial_trace_full_size(oracle, gpu):
    """p = 5 bricks (the headline order), 128 elements: the tangential trace of a field of the Nedelec space is reproduced on
    every face (size-independent property; evaluated back through the oracle's shape3DE)"""
    oracle.set_maxp(9)
    rng = np.random.default_rng(23)
    nel = 128
    no = np.tile(synth.uniform_order(5), (nel, 1))
    noe = rng.integers(0, 2, (nel, 12)).astype(np.int32); nof = rng.integers(0, 8, (nel, 6)).astype(np.int32)
    noe[nel // 2:] = noe[0]; nof[nel // 2:] = nof[0]   # half of the elements share one signature
    etav = np.zeros((nel, 8, 3))
    for e in range(nel):
        etav[e] = rng.uniform(0, 0.4, 3) + BRICK_M * rng.uniform(0.2, 0.6, 3)
    etype = np.full(nel, MDLB, np.int32)
    pts = api.pbi_hcurl_points(no, noe, nof)
    fv, fc = tabulate_E(nedelec_poly, 1, pts, etav, etype)
    res = api.pbi_hcurl_batch(no, noe, nof, etav, fv, fc)
    assert not res["info"].any()
    worst = 0.0
    for e in rng.choice(nel, 16, replace=False):
        nEF = int(pts["nrdofE"][e])
        for _ in range(4):
            xi = rng.random(3); ax = int(rng.integers(0, 3)); xi[ax] = float(rng.integers(0, 2))
            sE, _ = oracle.shape3DE(xi, no[e], noe[e], nof[e])
            s, g = oracle.shape3DH(xi, no[e], noe[e], nof[e])
            J = etav[e].T @ g[:8]
            u_eta = np.linalg.solve(J.T, res["dof"][e, :nEF, 0] @ sE[:nEF])
            E, _, A = nedelec_poly(s[:8] @ etav[e])
            want = A.T @ E[0]
            t = [a for a in range(3) if a != ax]
            worst = max(worst, float(np.abs(u_eta[t] - want[t]).max()))
    assert worst < 1e-10


# ----------------------------------------------------------------------------------------------------- H(div) Dirichlet dofs (dhpfaceV_opt)
def curved_V(eta):
    E, _, _ = smooth_E(eta)
    return E, np.zeros_like(E), curved_E(eta)[2]


def tabulate_V(fun, ncomp, pts, etav, etype):
    """V_eta = det(dxdeta) dxdeta^-1 V at the product's points"""
    nel = etav.shape[0]
    fv = np.zeros((nel, pts["xi"].shape[1], 3, ncomp))
    for e in range(nel):
        et = int(etype[e]); nv = 8 if et == MDLB else 6
        for l in range(int(pts["npts"][e])):
            V, _, J = fun(vertex_shape(et, pts["xi"][e, l]) @ etav[e, :nv])
            fv[e, l] = np.linalg.det(J) * np.linalg.solve(J, np.asarray(V).reshape(ncomp, 3).T)
    return fv


def test_hdiv_points_are_the_ones_the_reference_loops_visit(oracle, gpulib):
    oracle.set_maxp(9)
    rng = np.random.default_rng(3)
    for et, gen in ((MDLB, random_brick), (MDLP, random_prism)):
        no, noe, nof = gen(rng, 1, 4)
        etav = warped_vertices(rng, et)
        nv, ne = (8, 12) if et == MDLB else (6, 9)
        seen = []

        def fun(eta):
            seen.append(eta.copy())
            return np.zeros((1, 3)), np.zeros((1, 3)), np.eye(3)
        oracle.pbi_hdiv_element(no, noe, nof, etav[:nv], fun, 1, etype=et)
        pts = api.pbi_hdiv_points(no, noe, nof, etype=et)
        n = int(pts["npts"][0])
        assert len(seen) == n
        mine = np.array([vertex_shape(et, x) @ etav[:nv] for x in pts["xi"][0, :n]])
        assert np.abs(np.array(seen) - mine).max() < 1e-14
        off = oracle.pbi_offsets_V(no, et)
        assert np.array_equal(pts["nodes"][0, nv + ne:nv + ne + len(off) - 1, 0], off[:-1]) and int(pts["nrdofV"][0]) == int(off[-1])


@pytest.mark.gpu
def test_gpu_hdiv_matches_oracle_random_elements(oracle, gpu):
    """bricks and prisms (triangle and quad faces), anisotropic orders 1..5, all orientations, warped vertices, curved GMP map;
    one element with a face mask: unselected faces keep their incoming dofs"""
    oracle.set_maxp(9)
    rng = np.random.default_rng(29)
    rows = [(MDLB,) + random_brick(rng, 1, 5) for _ in range(3)] + [(MDLP,) + random_prism(rng, 1, 5) for _ in range(3)] 
    rows += [rows[0]]
    etype = np.array([r[0] for r in rows], np.int32)
    norder = np.array([r[1] for r in rows]); noe = np.array([r[2] for r in rows]); nof = np.array([r[3] for r in rows])
    etav = np.array([warped_vertices(rng, int(t)) for t in etype])
    nel = len(rows)
    pts = api.pbi_hdiv_points(norder, noe, nof, etype=etype)
    fv = tabulate_V(curved_V, 2, pts, etav, etype)
    mask = np.full(nel, 0xFFFFFFFF, np.uint32)
    mask[nel - 1] = (1 << (8 + 12 + 1)) | (1 << (8 + 12 + 4))   # faces 2 and 5 of the last brick
    dof_in = rng.standard_normal((nel, int(pts["nrdofV"].max()), 2))
    res = api.pbi_hdiv_batch(norder, noe, nof, etav, fv, mask=mask, dof=dof_in, etype=etype)
    assert not res["info"].any()
    for e in range(nel):
        et = int(etype[e]); nv, ne = (8, 12) if et == MDLB else (6, 9); nV = int(pts["nrdofV"][e])
        ref = oracle.pbi_hdiv_element(norder[e], noe[e], nof[e], etav[e, :nv], curved_V, 2, etype=et, mask=(int(mask[e]) >> (nv + ne)) & 0x3F,
                                      dof=dof_in[e, :nV])
        assert np.abs(res["dof"][e, :nV] - ref).max() / max(1.0, np.abs(ref).max()) < 1e-11
    assert np.array_equal(res["dof"][nel - 1, :int(pts["nodes"][nel - 1, 20 + 1, 0])], dof_in[nel - 1, :int(pts["nodes"][nel - 1, 20 + 1, 0])])


# ------------------------------------------------------------------------------------------ golden fixtures (tools/make_golden_pbi.py)
import glob  # noqa: E402
import os  # noqa: E402

PBI_FIXTURES = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pbi_*.npz")))


def gmp(eta):
    """the curved GMP block of the fixtures: x(eta) and dx/deta"""
    x, y, z = eta
    v = np.array([x + 0.1 * np.sin(2.0 * y) * z, y + 0.05 * x * x, z + 0.1 * np.cos(x + y)])
    d = np.array([[1.0, 0.2 * np.cos(2.0 * y) * z, 0.1 * np.sin(2.0 * y)], [0.1 * x, 1.0, 0.0], [-0.1 * np.sin(x + y), -0.1 * np.sin(x + y), 1.0]])
    return v, d


@pytest.mark.parametrize("path", PBI_FIXTURES, ids=[os.path.basename(f)[:-4] for f in PBI_FIXTURES])
def test_oracle_reproduces_pbi_golden(oracle, path):
    g = np.load(path)
    oracle.set_maxp(9); oracle.use_blas(False)
    et = int(g["etype"]); nv = 8 if et == MDLB else 6
    for e in range(2):
        no, noe, nof, etav = g[f"norder{e}"], g[f"norie{e}"], g[f"norif{e}"], g[f"etav{e}"]
        assert np.abs(oracle.pbi_element(no, noe, nof, etav[:nv], gmp, 3, etype=et) - g[f"h1_dof{e}"]).max() < 1e-14
        assert np.abs(oracle.pbi_hcurl_element(no, noe, nof, etav[:nv], curved_E, 2, etype=et) - g[f"e_dof{e}"]).max() < 1e-14
        assert np.abs(oracle.pbi_hdiv_element(no, noe, nof, etav[:nv], curved_V, 2, etype=et) - g[f"v_dof{e}"]).max() < 1e-14


@pytest.mark.gpu
@pytest.mark.parametrize("path", PBI_FIXTURES, ids=[os.path.basename(f)[:-4] for f in PBI_FIXTURES])
def test_gpu_reproduces_pbi_golden(gpu, path):
    """the device path fed with the fixture's tabulated data (no oracle involved) reproduces the fixture's dofs"""
    g = np.load(path)
    et = int(g["etype"]); nv = 8 if et == MDLB else 6
    for e in range(2):
        no, noe, nof, etav = g[f"norder{e}"], g[f"norie{e}"], g[f"norif{e}"], g[f"etav{e}"]
        etype = np.array([et], np.int32)
        # H1
        pts = api.pbi_points(no, noe, nof, etype=etype); n = int(pts["npts"][0])
        mine = np.array([vertex_shape(et, x) @ etav[:nv] for x in pts["xi"][0, :n]]).reshape(-1, 3)
        assert n == len(g[f"h1_eta{e}"]) and np.abs(mine - g[f"h1_eta{e}"]).max() < 1e-14
        fv = np.zeros((1, 8, 3)); fv[0, :nv] = g[f"h1_fvert{e}"]
        res = api.pbi_h1_batch(no, noe, nof, etav[None], fv, g[f"h1_fgrad{e}"][None], etype=etype)
        assert not res["info"].any() and np.abs(res["dof"][0] - g[f"h1_dof{e}"]).max() < 1e-11
        # H(curl)
        pts = api.pbi_hcurl_points(no, noe, nof, etype=etype); n = int(pts["npts"][0])
        mine = np.array([vertex_shape(et, x) @ etav[:nv] for x in pts["xi"][0, :n]]).reshape(-1, 3)
        assert n == len(g[f"e_eta{e}"]) and np.abs(mine - g[f"e_eta{e}"]).max() < 1e-14
        res = api.pbi_hcurl_batch(no, noe, nof, etav[None], g[f"e_fval{e}"][None], g[f"e_fcurl{e}"][None], etype=etype)
        assert not res["info"].any() and np.abs(res["dof"][0] - g[f"e_dof{e}"]).max() < 1e-10
        # H(div)
        pts = api.pbi_hdiv_points(no, noe, nof, etype=etype); n = int(pts["npts"][0])
        mine = np.array([vertex_shape(et, x) @ etav[:nv] for x in pts["xi"][0, :n]]).reshape(-1, 3)
        assert n == len(g[f"v_eta{e}"]) and np.abs(mine - g[f"v_eta{e}"]).max() < 1e-14
        res = api.pbi_hdiv_batch(no, noe, nof, etav[None], g[f"v_fval{e}"][None], etype=etype)
        assert not res["info"].any() and np.abs(res["dof"][0] - g[f"v_dof{e}"]).max() < 1e-11


@pytest.mark.gpu
def test_gpu_table_cache_trim_keeps_results(oracle, gpu):
    """with a 1-byte limit every call drops the cached device tables first and rebuilds them: same dofs as with the cache"""
    import ctypes as C
    rng = np.random.default_rng(31)
    no, noe, nof = random_brick(rng, 2, 4)
    etav = warped_vertices(rng, MDLB)[None]
    etype = np.array([MDLB], np.int32)
    pts = api.pbi_points(no, noe, nof, etype=etype)
    fv, fg = tabulate(smooth, 2, pts, etav, etype)
    ref = api.pbi_h1_batch(no, noe, nof, etav, fv, fg, etype=etype)["dof"]
    gpu.hp3d_gpu_pbi_cache_limit.argtypes = [C.c_longlong]
    try:
        assert gpu.hp3d_gpu_pbi_cache_limit(1) == 0
        for _ in range(2):
            assert np.array_equal(api.pbi_h1_batch(no, noe, nof, etav, fv, fg, etype=etype)["dof"], ref)
    finally:
        gpu.hp3d_gpu_pbi_cache_limit(8 << 30)


# ------------------------------------------------------------------------------------------------- more host logic (CPU): MAXP cap, tables
@pytest.mark.parametrize("et", [MDLB, MDLP])
def test_points_respect_the_maxp_cap(oracle, gpulib, et):
    """order + INTEGRATION is capped at MAXP (set_1D_int.F90:41-43 and its 2-D / 3-D twins): with MAXP = 4 an order-4 node keeps 5 points
    per direction under INTEGRATION = 1; the oracle visits the same points for all three families"""
    oracle.set_maxp(4)
    try:
        rng = np.random.default_rng(41)
        no, noe, nof = (random_brick if et == MDLB else random_prism)(rng, 3, 4)
        etav = warped_vertices(rng, et); nv = 8 if et == MDLB else 6
        for fam in ("h1", "hcurl", "hdiv"):
            seen = []
            if fam == "h1":
                oracle.pbi_element(no, noe, nof, etav[:nv], lambda eta: (seen.append(eta.copy()), (np.zeros(1), np.zeros((1, 3))))[1], 1,
                                   integration=1, maxp=4, etype=et)
                seen = seen[nv:]
                pts = api.pbi_points(no, noe, nof, integration=1, maxp=4, etype=et)
            else:
                f = lambda eta: (seen.append(eta.copy()), (np.zeros((1, 3)), np.zeros((1, 3)), np.eye(3)))[1]   # noqa: E731
                (oracle.pbi_hcurl_element if fam == "hcurl" else oracle.pbi_hdiv_element)(no, noe, nof, etav[:nv], f, 1, maxp=4, etype=et)
                pts = (api.pbi_hcurl_points if fam == "hcurl" else api.pbi_hdiv_points)(no, noe, nof, maxp=4, etype=et)
            n = int(pts["npts"][0])
            assert n == len(seen), fam
            mine = np.array([vertex_shape(et, x) @ etav[:nv] for x in pts["xi"][0, :n]])
            assert np.abs(np.array(seen) - mine).max() < 1e-14, fam
            uncapped = (api.pbi_points(no, noe, nof, integration=1, maxp=9, etype=et) if fam == "h1" else
                        (api.pbi_hcurl_points if fam == "hcurl" else api.pbi_hdiv_points)(no, noe, nof, maxp=9, etype=et))
            assert int(uncapped["npts"][0]) > n, fam
    finally:
        oracle.set_maxp(9)


def test_node_tables_are_consistent(gpulib):
    """for random descriptors: dof ranges of the nodes tile [0, nrdof), point ranges tile [0, npts), nodes without dofs have no points"""
    rng = np.random.default_rng(43)
    for _ in range(20):
        et = MDLB if rng.random() < 0.5 else MDLP
        no, noe, nof = (random_brick if et == MDLB else random_prism)(rng, 1, 6)
        nn = 27 if et == MDLB else 21
        for pts, key in ((api.pbi_points(no, noe, nof, integration=int(rng.integers(0, 2)), etype=et), "nrdofH"),
                         (api.pbi_hcurl_points(no, noe, nof, etype=et), "nrdofE"), (api.pbi_hdiv_points(no, noe, nof, etype=et), "nrdofV")):
            t = pts["nodes"][0]
            d0 = 0; p0 = 0
            for i in range(nn):
                if key != "nrdofH" and i == nn - 1:
                    assert t[i, 1] == 0 and t[i, 3] == 0      # no Dirichlet data on middle nodes
                    continue
                assert t[i, 0] == d0
                d0 += t[i, 1]
                if t[i, 3]:
                    assert t[i, 1] > 0 and t[i, 2] == p0
                    p0 += t[i, 3]
                elif key == "nrdofH" and i >= (8 if et == MDLB else 6):
                    assert t[i, 1] == 0
            assert d0 == int(pts[key][0]) and p0 == int(pts["npts"][0])
            assert not t[nn:].any()
