"""poly_pois (trunk/test/poly_pois.F90) generalised to BASELINE.json configs[4]: a conforming hp mesh of hexahedra and prisms
(variable order, min rule, orientations from a random global vertex numbering).  A quadratic polynomial lies in the
conforming FE space, so
  (1) a least-squares fit of its GLOBAL coefficient vector (dofs shared through (entity, k) keys) must be exact -- global
      conformity of the oracle's shape functions under the reference's dof order and orientation rules (CPU test);
  (2) the GPU's condensed element matrices, assembled with those keys and solved with Dirichlet data, must return the same
      interface coefficients, and the stored Schur factors the same bubble coefficients (elem + stc + solve + stc_bwd)."""
import numpy as np
import pytest

from hp3d_b200 import synth
from tests.mini_fem_hp import F_SOURCE, build_space, fit_polynomial


def _mesh():
    return synth.hp_mesh(2, prism_frac=0.5, pmin=2, pmax=3, seed_p=21, seed_g=4, jitter=0.0)


def test_polynomial_lies_in_the_conforming_space(oracle):
    oracle.set_maxp(8)
    m = _mesh()
    assert (m["etype"] == 1).any() and (m["etype"] == 3).any()
    keys, l2g, nloc, is_b, is_bub = build_space(m)
    U, res = fit_polynomial(oracle, m, l2g, len(keys))
    assert res < 1e-12, res


@pytest.mark.gpu
def test_poisson_polynomial_on_hp_mesh(oracle, gpu):
    from hp3d_b200.api import ElemEngine
    oracle.set_maxp(8)
    m = _mesh()
    nel = len(m["etype"])
    keys, l2g, nloc, is_b, is_bub = build_space(m)
    ndof = len(keys)
    U, res = fit_polynomial(oracle, m, l2g, ndof)
    assert res < 1e-12
    eng = ElemEngine(1, source=9, maxp=8)
    nint_max = max(eng.sizes(m["norder"][e], int(m["etype"][e]))[2] for e in range(nel))
    src = np.full((nel, nint_max), F_SOURCE)
    res = eng.elem_stc_batch(m["norder"], m["norient_edge"], m["norient_face"], m["xnod"], source_qp=src, etype=m["etype"])
    assert (res["info"] == 0).all()
    K = np.zeros((ndof, ndof)); Fv = np.zeros(ndof)
    for e in range(nel):
        Aii, Bi, _, _ = eng.unpack(res, e)
        g = l2g[e][:nloc[e]]
        assert Aii.shape[0] == nloc[e]
        K[np.ix_(g, g)] += Aii
        Fv[g] += Bi
    u = np.zeros(ndof)
    u[is_b] = U[is_b]
    free = ~is_b & ~is_bub
    u[free] = np.linalg.solve(K[np.ix_(free, free)], Fv[free] - K[np.ix_(free, is_b)] @ u[is_b])
    assert np.abs(u[free] - U[free]).max() < 1e-11, np.abs(u[free] - U[free]).max()
    for e in range(nel):
        _, _, AS, BS = eng.unpack(res, e)
        if AS.shape[0]:
            xb = BS - AS @ u[l2g[e][:nloc[e]]]
            assert np.abs(xb - U[l2g[e][nloc[e]:]]).max() < 1e-11
    eng.close()
