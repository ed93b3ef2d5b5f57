"""Host-side parity of the product's prism description (hp3d_b200/csrc/prism_space.hpp + tri_space.hpp: every prism shape
function as sign * T(x,y) * Z(z)) against the oracle's pointwise restatement of Prism.F90 -- same dof ORDER, same SIGNS
(exact: a wrong order or sign is an O(1) error), values to rounding (two independent evaluations of the polynomials)."""
import ctypes as C

import numpy as np
import pytest

from tests.test_oracle_prism import prism_signature, rand_point
from tests.util import _p, i32


def prism_shape(L, space, no, ne, nf, xi):
    no, ne, nf = i32(no), i32(ne), i32(nf)
    xi = np.ascontiguousarray(xi, dtype=np.float64)
    n = L.hp3d_gpu_prism_shape(space, _p(no), _p(ne), _p(nf), _p(xi, C.c_double), 0, None, None)
    assert n >= 0
    val = np.zeros((n, 3)); der = np.zeros((n, 3))
    m = L.hp3d_gpu_prism_shape(space, _p(no), _p(ne), _p(nf), _p(xi, C.c_double), n, _p(val, C.c_double), _p(der, C.c_double))
    assert m == n
    return val, der


@pytest.mark.parametrize("seed", range(12))
def test_prism_decomposition_vs_oracle(oracle, gpulib, seed):
    oracle.set_maxp(8)
    P = oracle.MDLP
    rng = np.random.default_rng(seed)
    p, pz = int(rng.integers(1, 7)), int(rng.integers(1, 7))
    no, ne, nf = prism_signature(rng, p, pz, uniform=(seed % 3 == 0))
    for _ in range(3):
        xi = rand_point(rng)
        tol = 2e-13
        v, d = prism_shape(gpulib, 0, no, ne, nf, xi)
        s, g = oracle.shape3DH(xi, no, ne, nf, P)
        assert v.shape[0] == s.size
        assert np.abs(v[:, 0] - s).max() < tol and np.abs(d - g).max() < tol * 50
        v, d = prism_shape(gpulib, 1, no, ne, nf, xi)
        E, c = oracle.shape3DE(xi, no, ne, nf, P)
        assert v.shape == E.shape
        assert np.abs(v - E).max() < tol and np.abs(d - c).max() < tol * 50
        v, d = prism_shape(gpulib, 3, no, ne, nf, xi)
        q = oracle.shape3DQ(xi, no, P)
        assert v.shape[0] == q.size and np.abs(v[:, 0] - q).max() < tol
        # H(div): the product only describes the face functions (normal traces); they come first in the reference order
        noi = no.copy(); noi[14] = 11
        v, d = prism_shape(gpulib, 2, noi, ne, nf, xi)
        V, dv = oracle.shape3DV(xi, noi, nf, P)
        assert v.shape[0] == V.shape[0]
        assert np.abs(v - V).max() < tol and np.abs(d[:, 0] - dv).max() < tol * 50
