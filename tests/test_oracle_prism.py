"""Pins of the oracle's triangle / prism branch (oracle/shape_prism.c, etype.c, tri_rules.h).

The reference holds no numeric vectors for prisms; the pins are analytic:
  * the Dunavant rules in hp3D's point order integrate monomials of degree 2p exactly (the property NSELECT is chosen
    for, gauss_quadrature.F90:47-48; debug self-test set_3D_int.F90 `test_set_3Dint`);
  * dof counts == ndof_nod/celndof (element_data.F90:808-870), partition of unity, gradients / curls / divergences ==
    finite differences of the values, exact-sequence inclusions grad H1 c H(curl), curl H(curl) c H(div), div H(div) = L2
    with the dimension count of each space (Prism.F90:38-1131);
  * element level: constants are in the kernel of the Poisson stiffness, x^T A x = |K| for u = x on an affine prism.
"""
import math

import numpy as np
import pytest

QSWAP = [0, 1, 0, 1, 1, 0, 1, 0]
PV = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1]], float)


def prism_signature(rng, p, pz, uniform=True):
    """(norder, norie, norif) of a prism: quad-face digits in the face's own oriented frame (find_order.F90:41-58)."""
    no = np.zeros(19, np.int32)
    ne = np.zeros(12, np.int32); nf = np.zeros(6, np.int32)
    ne[:9] = rng.integers(0, 2, 9); nf[:2] = rng.integers(0, 6, 2); nf[2:5] = rng.integers(0, 8, 3)
    pick = (lambda hi: hi) if uniform else (lambda hi: int(rng.integers(1, hi + 1)))
    no[:6] = [pick(p) for _ in range(6)]
    no[6:9] = [pick(pz) for _ in range(3)]
    no[9:11] = [pick(p) for _ in range(2)]
    for f in range(3):
        a, b = pick(p), pick(pz)
        if QSWAP[nf[2 + f]]:
            a, b = b, a
        no[11 + f] = 10 * a + b
    no[14] = 10 * p + pz
    return no, ne, nf


def rand_point(rng):
    while True:
        x = rng.random(3)
        if x[0] + x[1] < 0.95:
            return x


def test_triangle_rules_exact(oracle):
    for p in range(1, 10):
        no = oracle.uniform_order(p, oracle.MDLP, 1)
        xi, w = oracle.quad3(no, np.zeros(6, np.int32), 0, 9, oracle.MDLP)
        assert abs(w.sum() - 0.5) < 1e-14
        z0 = xi[0, 2]
        sel = np.abs(xi[:, 2] - z0) < 1e-15
        x, y, wt = xi[sel, 0], xi[sel, 1], w[sel] / w[sel].sum() * 0.5
        for a in range(2 * p + 1):
            for b in range(2 * p + 1 - a):
                ex = math.factorial(a) * math.factorial(b) / math.factorial(a + b + 2)
                assert abs((wt * x ** a * y ** b).sum() - ex) < 1e-13 * max(ex, 1e-3), (p, a, b)


@pytest.mark.parametrize("seed", range(6))
def test_prism_shape_derivatives(oracle, seed):
    oracle.set_maxp(8)
    P = oracle.MDLP
    rng = np.random.default_rng(seed)
    p, pz = int(rng.integers(1, 5)), int(rng.integers(1, 5))
    no, ne, nf = prism_signature(rng, p, pz, uniform=False)
    xi = rand_point(rng)
    s, g = oracle.shape3DH(xi, no, ne, nf, P)
    assert abs(s[:6].sum() - 1) < 1e-14
    h = 1e-6
    E, c = oracle.shape3DE(xi, no, ne, nf, P)
    V, dv = oracle.shape3DV(xi, no, nf, P)
    J = np.zeros((E.shape[0], 3, 3)); dfd = np.zeros(V.shape[0])
    for d in range(3):
        e = np.zeros(3); e[d] = h
        sp, _ = oracle.shape3DH(xi + e, no, ne, nf, P); sm, _ = oracle.shape3DH(xi - e, no, ne, nf, P)
        assert np.abs((sp - sm) / (2 * h) - g[:, d]).max() < 1e-7
        Ep, _ = oracle.shape3DE(xi + e, no, ne, nf, P); Em, _ = oracle.shape3DE(xi - e, no, ne, nf, P)
        J[:, :, d] = (Ep - Em) / (2 * h)
        Vp, _ = oracle.shape3DV(xi + e, no, nf, P); Vm, _ = oracle.shape3DV(xi - e, no, nf, P)
        dfd += (Vp[:, d] - Vm[:, d]) / (2 * h)
    cfd = np.stack([J[:, 2, 1] - J[:, 1, 2], J[:, 0, 2] - J[:, 2, 0], J[:, 1, 0] - J[:, 0, 1]], 1)
    assert np.abs(cfd - c).max() < 1e-6
    assert np.abs(dfd - dv).max() < 1e-6
    # broken test spaces: counts of the enriched order
    nordM = int(no[14]) + 11
    pe, pze = p + 1, pz + 1
    sH, _ = oracle.shape3HH(xi, nordM, P); sE, _ = oracle.shape3EE(xi, nordM, P)
    assert sH.size == (pe + 1) * (pe + 2) // 2 * (pze + 1)
    assert sE.shape[0] == pe * (pe + 2) * (pze + 1) + (pe + 1) * (pe + 2) // 2 * pze


@pytest.mark.parametrize("p,pz", [(2, 2), (3, 2), (2, 3)])
def test_prism_exact_sequence(oracle, p, pz):
    oracle.set_maxp(8)
    P = oracle.MDLP
    rng = np.random.default_rng(10 * p + pz)
    no, ne, nf = prism_signature(rng, p, pz)
    G, E, Cc, V, D, Q = [], [], [], [], [], []
    for _ in range(300):
        xi = rand_point(rng)
        s, g = oracle.shape3DH(xi, no, ne, nf, P); e, c = oracle.shape3DE(xi, no, ne, nf, P)
        v, d = oracle.shape3DV(xi, no, nf, P); q = oracle.shape3DQ(xi, no, P)
        G.append(g.T); E.append(e.T); Cc.append(c.T); V.append(v.T); D.append(d[None]); Q.append(q[None])
    G, E, Cc, V, D, Q = (np.concatenate(a, 0) for a in (G, E, Cc, V, D, Q))
    tri = (p + 1) * (p + 2) // 2
    assert G.shape[1] == tri * (pz + 1)
    assert E.shape[1] == p * (p + 2) * (pz + 1) + tri * pz
    assert V.shape[1] == p * (p + 2) * pz + p * (p + 1) // 2 * (pz + 1)
    assert Q.shape[1] == p * (p + 1) // 2 * pz

    def resid(A, B):
        x = np.linalg.lstsq(A, B, rcond=None)[0]
        return np.abs(A @ x - B).max()
    assert resid(E, G) < 1e-12 and resid(V, Cc) < 1e-12 and resid(Q, D) < 1e-12
    assert np.linalg.matrix_rank(E) == E.shape[1] and np.linalg.matrix_rank(V) == V.shape[1]
    assert np.linalg.matrix_rank(D) == Q.shape[1]            # div is onto L2


@pytest.mark.parametrize("p,pz", [(1, 1), (2, 2), (3, 2)])
def test_prism_elements(oracle, p, pz):
    oracle.set_maxp(8)
    P = oracle.MDLP
    rng = np.random.default_rng(p + 7 * pz)
    no, ne, nf = prism_signature(rng, p, pz)
    nH = oracle.celndof(no, P)[0]
    T = rng.normal(size=(3, 3)) * 0.2 + np.eye(3)
    X = np.zeros((nH, 3)); X[:6] = PV @ T.T + 0.3
    prm = oracle.default_params(omega=1.0)
    A, b = oracle.elem(1, no, ne, nf, X, prm, etype=P)
    one = np.zeros(nH); one[:6] = 1
    ux = np.zeros(nH); ux[:6] = X[:6, 0]
    assert np.abs(A @ one).max() < 1e-14
    assert abs(ux @ A @ ux - abs(np.linalg.det(T)) / 2) < 1e-14
    for kind in (2, 3, 4):
        prm = oracle.default_params(omega=2 * np.pi if kind == 4 else 1.0)
        Aii, Bi, AS, BS = oracle.condensed(kind, no, ne, nf, X, prm, etype=P)
        assert np.isfinite(Aii).all() and np.isfinite(AS).all()
        if kind != 3:
            assert np.abs(Aii - Aii.conj().T).max() < 1e-13 * np.abs(Aii).max()
