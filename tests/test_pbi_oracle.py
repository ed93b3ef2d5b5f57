"""Pins of the projection-based interpolation oracle (oracle/pbi.c = hpvert / hpedge / hpface_opt / hpmdle_opt and
dhpvert / dhpedgeH / dhpfaceH_opt).  The reference's tests hold no numeric vectors for src/hpinterp;
trunk/test/poly_pois.F90 exercises update_gdof + update_Ddof and asserts that a polynomial is reproduced -- pinned here
directly, together with the conformity that the node-by-node construction of update_gdof relies on (the dofs of a shared
vertex / edge / face are the same whichever adjacent element computes them)."""
import numpy as np
import pytest

from hp3d_b200 import synth
from tests.mini_fem_hp import build_space

MDLB, MDLP = 1, 3


def poly(eta):
    x, y, z = eta
    v = np.array([1.0 + 0.5 * x - y + x * x + 2.0 * y * z - x * z + 0.25 * z * z,
                  x * y * z - 0.3 * z * z + y,
                  2.0 - x + 0.7 * x * y])
    d = np.array([[0.5 + 2 * x - z, -1.0 + 2 * z, 2 * y - x + 0.5 * z],
                  [y * z, x * z + 1.0, x * y - 0.6 * z],
                  [-1.0 + 0.7 * y, 0.7 * x, 0.0]])
    return v, d


def smooth(eta):
    x, y, z = eta
    v = np.array([np.sin(1.3 * x + 0.4) * np.cos(0.9 * y) * np.exp(0.5 * z), x + np.sin(2.0 * y * z)])
    d = np.array([[1.3 * np.cos(1.3 * x + 0.4) * np.cos(0.9 * y) * np.exp(0.5 * z),
                   -0.9 * np.sin(1.3 * x + 0.4) * np.sin(0.9 * y) * np.exp(0.5 * z),
                   0.5 * np.sin(1.3 * x + 0.4) * np.cos(0.9 * y) * np.exp(0.5 * z)],
                  [1.0, 2.0 * z * np.cos(2.0 * y * z), 2.0 * y * np.cos(2.0 * y * z)]])
    return v, d


def _eval(oracle, et, no, noe, nof, dof, etav, xi):
    s, _ = oracle.shape3DH(xi, no, noe, nof, et)
    nv = 8 if et == MDLB else 6
    return s @ dof, s[:nv] @ etav


@pytest.mark.parametrize("integration", [0, 1])
def test_polynomial_reproduction_brick(oracle, integration):
    """degree <= 2 per variable (3 for xyz): reproduced by any brick whose nodes all have order >= 3, any orientation,
    on an axis-aligned sub-box of the reference block (a refined element's Etav)"""
    oracle.set_maxp(9)
    rng = np.random.default_rng(3)
    box0, box1 = np.array([0.25, 0.0, 0.5]), np.array([0.75, 0.5, 1.0])
    M = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
    etav = box0 + M * (box1 - box0)
    for trial in range(3):
        no = np.array(list(rng.integers(3, 6, 12)) + [10 * int(rng.integers(3, 6)) + int(rng.integers(3, 6)) for _ in range(6)]
                      + [100 * int(rng.integers(3, 6)) + 10 * int(rng.integers(3, 6)) + int(rng.integers(3, 6))], np.int32)
        noe = rng.integers(0, 2, 12).astype(np.int32); nof = rng.integers(0, 8, 6).astype(np.int32)
        dof = oracle.pbi_element(no, noe, nof, etav, poly, 3, integration=integration)
        for _ in range(20):
            xi = rng.random(3)
            u, eta = _eval(oracle, MDLB, no, noe, nof, dof, etav, xi)
            assert np.abs(u - poly(eta)[0]).max() < 1e-13


def test_polynomial_reproduction_prism(oracle):
    oracle.set_maxp(9)
    rng = np.random.default_rng(4)
    etav = np.array([[0.2, 0.1, 0.0], [0.9, 0.2, 0.0], [0.3, 0.8, 0.0], [0.2, 0.1, 0.6], [0.9, 0.2, 0.6], [0.3, 0.8, 0.6]])
    for trial in range(3):
        p, pz = int(rng.integers(3, 6)), int(rng.integers(3, 6))
        no = np.array([p] * 6 + [pz] * 3 + [p, p] + [10 * p + pz] * 3 + [10 * p + pz], np.int32)
        noe = rng.integers(0, 2, 9).astype(np.int32)
        nof = np.array(list(rng.integers(0, 6, 2)) + list(rng.integers(0, 8, 3)), np.int32)
        dof = oracle.pbi_element(no, noe, nof, etav, poly, 3, etype=MDLP)
        for _ in range(20):
            xi = rng.random(3)
            if xi[0] + xi[1] > 1:
                xi[:2] = 1 - xi[:2]
            u, eta = _eval(oracle, MDLP, no, noe, nof, dof, etav, xi)
            assert np.abs(u - poly(eta)[0]).max() < 1e-13


def test_trilinear_map_has_no_higher_order_dofs(oracle):
    """update_gdof on an element of a trilinear GMP block: x(eta) trilinear => edge / face / middle dofs vanish"""
    oracle.set_maxp(9)
    A = np.array([[1.0, 0.2, 0.0], [0.1, 1.5, 0.3], [0.0, -0.2, 0.8]])

    def tri(eta):
        x, y, z = eta
        v = A @ eta + np.array([0.3 * x * y, 0.2 * y * z, 0.1 * x * y * z])
        d = A + np.array([[0.3 * y, 0.3 * x, 0.0], [0.0, 0.2 * z, 0.2 * y], [0.1 * y * z, 0.1 * x * z, 0.1 * x * y]])
        return v, d
    no = synth.uniform_order(4)
    M = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
    dof = oracle.pbi_element(no, np.zeros(12, np.int32), np.zeros(6, np.int32), M, tri, 3)
    assert np.abs(dof[8:]).max() < 1e-14
    assert np.abs(dof[:8] - np.array([tri(m)[0] for m in M])).max() == 0.0


def test_shared_entities_get_identical_dofs(oracle):
    """a smooth (non-polynomial) function on a mixed hexa/prism hp mesh: dofs with the same global key coincide"""
    oracle.set_maxp(9)
    m = synth.hp_mesh(2, prism_frac=0.45, pmin=2, pmax=4, seed_p=11, seed_g=5)
    keys, l2g, nloc, _, _ = build_space(m)
    U = np.full((len(keys), 2), np.nan)
    worst, shared = 0.0, 0
    for e in range(len(m["etype"])):
        et = int(m["etype"][e]); nv = 8 if et == MDLB else 6
        dof = oracle.pbi_element(m["norder"][e], m["norient_edge"][e], m["norient_face"][e], m["xnod"][e, :nv], smooth, 2,
                                 integration=1, etype=et)
        for k, g in enumerate(l2g[e]):
            if np.isnan(U[g, 0]):
                U[g] = dof[k]
            else:
                worst = max(worst, float(np.abs(U[g] - dof[k]).max())); shared += 1
    assert shared > 50
    assert worst < 1e-12


# ---- H(curl) Dirichlet dofs (dhpedgeE, dhpfaceE_opt) -----------------------------------------------------------------------------
SCALE = np.array([1.5, 0.8, 1.2])   # a linear GMP block x = SCALE * eta: exercises both pullbacks


def nedelec_poly(eta):
    """a field whose pullback lies in the order-3 Nedelec space; physical components, curl, dx/deta"""
    x, y, z = SCALE * eta
    E = np.array([[x * x * y * z - y, x * z * z + x * y, z * z * x * y + 1.0]])
    cE = np.array([[z * z * x - 2.0 * x * z, x * x * y - z * z * y, z * z + y - x * x * z + 1.0]])
    return E, cE, np.diag(SCALE)


def smooth_E(eta):
    x, y, z = eta
    E = np.array([[np.sin(y + 0.3) * z, np.cos(x) * np.exp(0.3 * z), x * y + np.sin(z)],
                  [y * y, 0.5 * x * z, np.cos(x + y)]])
    cE = np.array([[x - 0.3 * np.cos(x) * np.exp(0.3 * z), np.sin(y + 0.3) - y, -np.sin(x) * np.exp(0.3 * z) - np.cos(y + 0.3) * z],
                   [-np.sin(x + y) - 0.5 * x, np.sin(x + y), 0.5 * z - 2.0 * y]])
    return E, cE, np.eye(3)


def test_nedelec_polynomial_tangential_trace_reproduced(oracle):
    oracle.set_maxp(9)
    rng = np.random.default_rng(8)
    M = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
    etav = np.array([0.1, 0.2, 0.0]) + M * np.array([0.5, 0.4, 0.6])
    no = synth.uniform_order(3)
    for trial in range(3):
        noe = rng.integers(0, 2, 12).astype(np.int32); nof = rng.integers(0, 8, 6).astype(np.int32)
        dof = oracle.pbi_hcurl_element(no, noe, nof, etav, nedelec_poly, 1)
        nEF = dof.shape[0]
        for _ in range(12):
            xi = rng.random(3); ax = int(rng.integers(0, 3)); xi[ax] = float(rng.integers(0, 2))   # a point on a face
            sE, _ = oracle.shape3DE(xi, no, noe, nof)
            s, g = oracle.shape3DH(xi, no, noe, nof)
            J = etav.T @ g[:8]                      # d eta / d xi
            u_eta = np.linalg.solve(J.T, (dof[:, 0] @ sE[:nEF]))   # J^-T E^
            E, _, A = nedelec_poly(s[:8] @ etav)
            want = A.T @ E[0]
            t = [a for a in range(3) if a != ax]
            assert np.abs(u_eta[t] - want[t]).max() < 1e-12


def test_hcurl_shared_entities_get_identical_dofs(oracle):
    from tests.test_hp_mesh_conformity import entity_blocks, topo
    oracle.set_maxp(9)
    m = synth.hp_mesh(2, prism_frac=0.45, pmin=1, pmax=4, seed_p=13, seed_g=6)
    seen, worst, shared = {}, 0.0, 0
    for e in range(len(m["etype"])):
        et = int(m["etype"][e]); nv = 8 if et == MDLB else 6
        E, F, _ = topo(et)
        v = [int(x) for x in m["verts"][e] if x >= 0]
        dof = oracle.pbi_hcurl_element(m["norder"][e], m["norient_edge"][e], m["norient_face"][e], m["xnod"][e, :nv], smooth_E, 2, etype=et)
        blocks, ntot = entity_blocks(et, m["norder"][e], "E")
        assert ntot == dof.shape[0]
        for kind, idx, m0, n in blocks:
            ent = (kind, frozenset((v[E[idx][0]], v[E[idx][1]]) if kind == "e" else tuple(v[i] for i in F[idx])))
            for k in range(n):
                key = ent + (k,)
                if key in seen:
                    worst = max(worst, float(np.abs(seen[key] - dof[m0 + k]).max())); shared += 1
                else:
                    seen[key] = dof[m0 + k]
    assert shared > 50
    assert worst < 1e-11


# ---- H(div) Dirichlet dofs (dhpfaceV_opt) ---------------------------------------------------------------------------------------
def rt_poly(eta):
    """a field whose Piola pullback lies in the order-3 Raviart-Thomas space of the brick (linear GMP block x = SCALE * eta)"""
    x, y, z = SCALE * eta
    V = np.array([[x * x * x * y * z + y * y, x * y * y * z * z - 1.0, z * z * z * x + x * y]])
    return V, np.zeros((1, 3)), np.diag(SCALE)


def test_raviart_thomas_normal_trace_reproduced(oracle):
    oracle.set_maxp(9)
    rng = np.random.default_rng(9)
    M = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
    etav = np.array([0.1, 0.2, 0.0]) + M * np.array([0.5, 0.4, 0.6])
    no = synth.uniform_order(3)
    for trial in range(3):
        noe = rng.integers(0, 2, 12).astype(np.int32); nof = rng.integers(0, 8, 6).astype(np.int32)
        dof = oracle.pbi_hdiv_element(no, noe, nof, etav, rt_poly, 1)
        nVF = dof.shape[0]
        for _ in range(12):
            xi = rng.random(3); ax = int(rng.integers(0, 3)); xi[ax] = float(rng.integers(0, 2))
            sV, _ = oracle.shape3DV(xi, no, nof)
            s, g = oracle.shape3DH(xi, no, noe, nof)
            J = etav.T @ g[:8]
            u_eta = J @ (dof[:, 0] @ sV[:nVF]) / np.linalg.det(J)      # Piola: J V^ / det
            V, _, A = rt_poly(s[:8] @ etav)
            want = np.linalg.det(A) * np.linalg.solve(A, V[0])         # det(dxdeta) dxdeta^-1 V
            assert abs(u_eta[ax] - want[ax]) < 1e-12


# ---- variational characterisation with an independent quadrature -----------------------------------------------------------------
def test_h1_interpolant_satisfies_its_variational_definition(oracle):
    """The PB interpolant is DEFINED by Galerkin orthogonality in the H1 seminorm, node by node: on every edge
    int d_t(g - u_h) d_t(phi_j) = 0 for the edge's bubbles, on every face int grad_s(g - u_h).grad_s(phi_j) = 0 for the face's
    bubbles, inside int grad(g - u_h).grad(phi_j) = 0 for the middle node's bubbles.  Checked for a polynomial g that is NOT in the
    element's space (degree p+1 per variable, so the oracle's p+1-point rules are still exact) with numpy's own Gauss-Legendre rule
    (12 points per direction) -- independent of the oracle's quadrature tables, system assembly and solvers."""
    oracle.set_maxp(9)
    rng = np.random.default_rng(12)
    p = 3
    no = synth.uniform_order(p)
    noe = rng.integers(0, 2, 12).astype(np.int32); nof = rng.integers(0, 8, 6).astype(np.int32)
    M = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
    h = np.array([0.5, 0.4, 0.6]); o = np.array([0.1, 0.2, 0.0])
    etav = o + M * h

    def g(eta):
        x, y, z = eta
        return (np.array([x ** 4 * y * y * z + x * y ** 4 + z ** 4 * x * x]),
                np.array([[4 * x ** 3 * y * y * z + y ** 4 + 2 * x * z ** 4, 2 * x ** 4 * y * z + 4 * x * y ** 3, x ** 4 * y * y + 4 * z ** 3 * x * x]]))
    dof = oracle.pbi_element(no, noe, nof, etav, g, 1)[:, 0]
    off = oracle.pbi_offsets(no)
    t, w = np.polynomial.legendre.leggauss(12)
    t = 0.5 * (t + 1.0); w = 0.5 * w

    def resid(xi):   # grad_eta (g - u_h) and grad_eta of all shape functions at a master point
        s, gr = oracle.shape3DH(xi, no, noe, nof)
        gr = gr / h
        return g(o + xi * h)[1][0] - dof @ gr, gr
    worst = 0.0
    EV = [(0, 1), (1, 2), (3, 2), (0, 3), (4, 5), (5, 6), (7, 6), (4, 7), (0, 4), (1, 5), (2, 6), (3, 7)]
    for e, (a, b) in enumerate(EV):                                  # edges: tangential derivative
        d = M[b] - M[a]; tau = d * h; L = np.linalg.norm(tau); tau = tau / L
        for j in range(off[8 + e], off[8 + e + 1]):
            r = sum(wi * L * (resid(M[a] + ti * d)[0] @ tau) * (resid(M[a] + ti * d)[1][j] @ tau) for ti, wi in zip(t, w))
            worst = max(worst, abs(r))
    FV = [(0, 1, 3), (4, 5, 7), (0, 1, 4), (1, 2, 5), (3, 2, 7), (0, 3, 4)]
    for f, (a, b, c) in enumerate(FV):                               # faces: surface gradient
        d1, d2 = M[b] - M[a], M[c] - M[a]
        n = np.cross(d1 * h, d2 * h); area = np.linalg.norm(n); n = n / area
        P = np.eye(3) - np.outer(n, n)
        acc = np.zeros(off[20 + f + 1] - off[20 + f])
        for t1, w1 in zip(t, w):
            for t2, w2 in zip(t, w):
                R, G = resid(M[a] + t1 * d1 + t2 * d2)
                acc += w1 * w2 * area * (G[off[20 + f]:off[20 + f + 1]] @ (P @ R))
        worst = max(worst, float(np.abs(acc).max()))
    acc = np.zeros(off[27] - off[26])                                # middle node: full gradient
    for t1, w1 in zip(t, w):
        for t2, w2 in zip(t, w):
            for t3, w3 in zip(t, w):
                R, G = resid(np.array([t1, t2, t3]))
                acc += w1 * w2 * w3 * np.prod(h) * (G[off[26]:off[27]] @ R)
    worst = max(worst, float(np.abs(acc).max()))
    assert worst < 1e-12
    assert np.abs(dof[8:]).max() > 1e-3      # g is not in the span of the vertex functions: the test is not vacuous


def test_hcurl_interpolant_satisfies_its_variational_definition(oracle):
    """H(curl) Dirichlet interpolant, same idea: on an edge the tangential residual is L2-orthogonal to the edge's functions; on a
    face the normal curl of the residual is orthogonal to the normal curls of the face's functions AND the tangential residual is
    orthogonal to the surface gradients of the face's H1 bubbles (the two block rows of dhpfaceE_opt's saddle-point system; the
    multiplier vanishes).  E is a polynomial of degree p+1 per variable (not in the space), numpy's Gauss-Legendre rule."""
    oracle.set_maxp(9)
    rng = np.random.default_rng(14)
    p = 2
    no = synth.uniform_order(p)
    noe = rng.integers(0, 2, 12).astype(np.int32); nof = rng.integers(0, 8, 6).astype(np.int32)
    M = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
    h = np.array([0.5, 0.4, 0.6]); o = np.array([0.1, 0.2, 0.0])
    etav = o + M * h

    def E(eta):
        x, y, z = eta
        e = np.array([[x ** 3 * y + z ** 3, y ** 3 * z * x - x, z ** 3 + x * y * y * z]])
        c = np.array([[2 * x * y * z - x * y ** 3, 3 * z * z - y * y * z, y ** 3 * z - 1.0 - x ** 3]])
        return e, c, np.eye(3)
    dof = oracle.pbi_hcurl_element(no, noe, nof, etav, E, 1)[:, 0]
    offE = oracle.pbi_offsets_E(no); offH = oracle.pbi_offsets(no)
    nEF = int(offE[-1])
    t, w = np.polynomial.legendre.leggauss(12)
    t = 0.5 * (t + 1.0); w = 0.5 * w
    J = np.diag(h); detJ = float(np.prod(h))

    def at(xi):
        sE, cE = oracle.shape3DE(xi, no, noe, nof)
        _, gH = oracle.shape3DH(xi, no, noe, nof)
        u = sE[:nEF] / h              # J^-T E^ (J diagonal)
        cu = cE[:nEF] * h / detJ      # J curl^ / det
        e, c, _ = E(o + xi * h)
        return e[0] - dof @ u, c[0] - dof @ cu, u, cu, gH / h
    worst = 0.0
    EV = [(0, 1), (1, 2), (3, 2), (0, 3), (4, 5), (5, 6), (7, 6), (4, 7), (0, 4), (1, 5), (2, 6), (3, 7)]
    for e, (a, b) in enumerate(EV):
        d = M[b] - M[a]; tau = d * h; L = np.linalg.norm(tau); tau = tau / L
        acc = np.zeros(offE[e + 1] - offE[e])
        for ti, wi in zip(t, w):
            r, _, u, _, _ = at(M[a] + ti * d)
            acc += wi * L * (r @ tau) * (u[offE[e]:offE[e + 1]] @ tau)
        worst = max(worst, float(np.abs(acc).max()))
    FV = [(0, 1, 3), (4, 5, 7), (0, 1, 4), (1, 2, 5), (3, 2, 7), (0, 3, 4)]
    for f, (a, b, c3) in enumerate(FV):
        d1, d2 = M[b] - M[a], M[c3] - M[a]
        n = np.cross(d1 * h, d2 * h); area = np.linalg.norm(n); n = n / area
        P = np.eye(3) - np.outer(n, n)
        je = slice(offE[12 + f], offE[12 + f + 1]); jh = slice(offH[20 + f], offH[20 + f + 1])
        acc_c = np.zeros(je.stop - je.start); acc_g = np.zeros(jh.stop - jh.start)
        for t1, w1 in zip(t, w):
            for t2, w2 in zip(t, w):
                r, rc, u, cu, gH = at(M[a] + t1 * d1 + t2 * d2)
                acc_c += w1 * w2 * area * (rc @ n) * (cu[je] @ n)
                acc_g += w1 * w2 * area * (gH[jh] @ (P @ r))
        worst = max(worst, float(np.abs(acc_c).max()), float(np.abs(acc_g).max()) if acc_g.size else 0.0)
    assert worst < 1e-12
    assert np.abs(dof).max() > 1e-3


def test_hdiv_interpolant_satisfies_its_variational_definition(oracle):
    """H(div) Dirichlet interpolant: on every face the normal component of V - u_h is L2-orthogonal to the normal components of the
    face's functions (dhpfaceV_opt); V of degree p per variable (one more than the face space holds), numpy's Gauss-Legendre rule."""
    oracle.set_maxp(9)
    rng = np.random.default_rng(15)
    p = 2
    no = synth.uniform_order(p)
    noe = rng.integers(0, 2, 12).astype(np.int32); nof = rng.integers(0, 8, 6).astype(np.int32)
    M = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
    h = np.array([0.5, 0.4, 0.6]); o = np.array([0.1, 0.2, 0.0])
    etav = o + M * h

    def V(eta):
        x, y, z = eta
        return np.array([[x * y * y * z * z + 1.0, x * x * z * z - y, x * x * y * y + z * z]]), np.zeros((1, 3)), np.eye(3)
    dof = oracle.pbi_hdiv_element(no, noe, nof, etav, V, 1)[:, 0]
    offV = oracle.pbi_offsets_V(no)
    nVF = int(offV[-1])
    t, w = np.polynomial.legendre.leggauss(12)
    t = 0.5 * (t + 1.0); w = 0.5 * w
    detJ = float(np.prod(h))
    worst = 0.0
    FV = [(0, 1, 3), (4, 5, 7), (0, 1, 4), (1, 2, 5), (3, 2, 7), (0, 3, 4)]
    for f, (a, b, c3) in enumerate(FV):
        d1, d2 = M[b] - M[a], M[c3] - M[a]
        n = np.cross(d1 * h, d2 * h); area = np.linalg.norm(n); n = n / area
        jv = slice(offV[f], offV[f + 1])
        acc = np.zeros(jv.stop - jv.start)
        for t1, w1 in zip(t, w):
            for t2, w2 in zip(t, w):
                xi = M[a] + t1 * d1 + t2 * d2
                sV, _ = oracle.shape3DV(xi, no, nof)
                u = sV[:nVF] * h / detJ          # Piola: J V^ / det
                acc += w1 * w2 * area * ((V(o + xi * h)[0][0] - dof @ u) @ n) * (u[jv] @ n)
        worst = max(worst, float(np.abs(acc).max()))
    assert worst < 1e-12
    assert np.abs(dof).max() > 1e-3
