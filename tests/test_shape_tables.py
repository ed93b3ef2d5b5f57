"""Host-side parity (no GPU): the product's signed tensor-product description of the hexahedral shape functions
and its 1-D tables against the oracle's restatement of src/element/shape_1/Hexahedron.F90 -- dof ORDER, orientation
SIGNS and values.  Index/sign data must match exactly; values to rounding (1e-14)."""
import ctypes as C

import numpy as np
import pytest

from tests.util import dof_map, random_signature, tables_1d


def test_gauss_tables_bit_exact(oracle, gpulib):
    # gauss_quadrature.F90:518-651,749-750: the product regenerates the reference's literals exactly
    for n in range(1, 11):
        x, w, *_ = tables_1d(gpulib, 3, n)
        xo, wo = oracle.gauss1(n)
        assert np.array_equal(x, xo), (n, x - xo)
        assert np.array_equal(w, wo), (n, w - wo)


def _eval(gpulib, space, norder, norie, norif, nq=4):
    """Evaluate the product's description on the tensor grid of nq Gauss points; returns values in the oracle's layout."""
    fam, idx, sgn = dof_map(gpulib, space, norder, norie, norif)
    p = 9
    x, w, H, dH, Q = tables_1d(gpulib, p, nq)
    pts = np.array([[x[i], x[j], x[k]] for k in range(nq) for j in range(nq) for i in range(nq)])
    qi = np.array([[i, j, k] for k in range(nq) for j in range(nq) for i in range(nq)])
    n = len(sgn)
    kind = {0: lambda f, d: 0, 3: lambda f, d: 1, 1: lambda f, d: 1 if d == f else 0, 2: lambda f, d: 0 if d == f else 1}[space]
    val = np.zeros((len(pts), n, 3)); der = np.zeros((len(pts), n, 3))
    for k in range(n):
        f = fam[k]
        T = [(H if kind(f, d) == 0 else Q)[idx[k, d]][qi[:, d]] for d in range(3)]
        dT = [dH[idx[k, d]][qi[:, d]] if kind(f, d) == 0 else None for d in range(3)]
        psi = sgn[k] * T[0] * T[1] * T[2]
        def dpsi(d):
            t = list(T); t[d] = dT[d]
            return sgn[k] * t[0] * t[1] * t[2]
        if space == 0:      # value + gradient
            val[:, k, 0] = psi
            for d in range(3):
                der[:, k, d] = dpsi(d)
        elif space == 3:
            val[:, k, 0] = psi
        elif space == 1:    # E = psi e_f ; curl = grad psi x e_f
            val[:, k, f] = psi
            b, c = (f + 1) % 3, (f + 2) % 3
            der[:, k, b] = dpsi(c)
            der[:, k, c] = -dpsi(b)
        else:               # V = psi e_f ; div = d_f psi
            val[:, k, f] = psi
            der[:, k, 0] = dpsi(f)
    return pts, val, der


@pytest.mark.parametrize("seed", range(6))
def test_dof_maps_against_oracle(oracle, gpulib, seed):
    rng = np.random.default_rng(seed)
    norder, norie, norif = random_signature(rng, pmax=4 if seed else 2, uniform=(seed == 0))
    oracle.set_maxp(9)
    nH, nE, nV, nQ = oracle.celndof(norder)
    for space, n_ref in ((0, nH), (1, nE), (2, nV), (3, nQ)):
        pts, val, der = _eval(gpulib, space, norder, norie, norif)
        assert val.shape[1] == n_ref
        for ip in range(0, len(pts), 7):
            xi = pts[ip]
            if space == 0:
                s, g = oracle.shape3DH(xi, norder, norie, norif)
                np.testing.assert_allclose(val[ip, :, 0], s, rtol=0, atol=2e-14)
                np.testing.assert_allclose(der[ip], g, rtol=0, atol=2e-13)
            elif space == 1:
                s, c = oracle.shape3DE(xi, norder, norie, norif)
                np.testing.assert_allclose(val[ip], s, rtol=0, atol=2e-14)
                np.testing.assert_allclose(der[ip], c, rtol=0, atol=2e-13)
            elif space == 2:
                s, d = oracle.shape3DV(xi, norder, norif)
                np.testing.assert_allclose(val[ip], s, rtol=0, atol=2e-14)
                np.testing.assert_allclose(der[ip, :, 0], d, rtol=0, atol=2e-13)
            else:
                s = oracle.shape3DQ(xi, norder)
                np.testing.assert_allclose(val[ip, :, 0], s, rtol=0, atol=2e-14)
