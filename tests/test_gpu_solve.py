"""Mini end-to-end solves around the drop-in boundary (north_star: final solution within 1e-10 of the reference path):
the same assembly + dense solve + stc back-substitution harness fed once with the oracle's element matrices and once with
the GPU's, on the structured cube mesh of trunk/test/poly_pois.F90 / conv_pois.F90."""
import numpy as np
import pytest

from tests.mini_fem import CubeMeshH1

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,p", [(4, 1), (3, 2), (2, 3), (2, 4)])
def test_poisson_solution_matches_oracle_path(oracle, gpu, N, p):
    from hp3d_b200.api import ElemEngine
    oracle.set_maxp(6)
    mesh = CubeMeshH1(gpu, N, p)
    no, oe, of, X = mesh.descriptors()
    prm = oracle.default_params(source=1)
    ref = [oracle.condensed(oracle.POIS_GAL, no[e], oe[e], of[e], X[e], prm) for e in range(len(no))]
    u_ref = mesh.solve(np.array([r[0] for r in ref]), np.array([r[1] for r in ref]))
    eng = ElemEngine(1)
    res = eng.elem_stc_batch(no, oe, of, X)
    assert (res["info"] == 0).all()
    out = [eng.unpack(res, e) for e in range(len(no))]
    u_gpu = mesh.solve(np.array([o[0] for o in out]), np.array([o[1] for o in out]))
    assert np.abs(u_gpu - u_ref).max() < 1e-10 * max(1.0, np.abs(u_ref).max())
    if mesh.nbub:
        # bubbles: stc_bwd on the GPU with the GPU's own Schur factors vs the oracle path
        xi = mesh.local_interface(u_gpu)
        AS = np.array([o[2] for o in out]); BS = np.array([o[3] for o in out])
        xb = eng.stc_bwd_batch(AS, BS, xi)
        xb_ref = np.array([r[3] - r[2] @ mesh.local_interface(u_ref)[e] for e, r in enumerate(ref)])
        assert np.abs(xb - xb_ref).max() < 1e-10 * max(1.0, np.abs(xb_ref).max())
    eng.close()


def test_poly_pois_through_gpu(gpu):
    """trunk/test/poly_pois.F90 with the GPU as the element routine: u = xyz reproduced to 1e-14."""
    from hp3d_b200.api import ElemEngine
    mesh = CubeMeshH1(gpu, 4, 1)
    no, oe, of, X = mesh.descriptors()
    eng = ElemEngine(1, source=0)
    res = eng.elem_stc_batch(no, oe, of, X)
    out = [eng.unpack(res, e) for e in range(len(no))]
    u = mesh.solve(np.array([o[0] for o in out]), np.array([o[1] for o in out]), dirichlet=lambda x: x[0] * x[1] * x[2])
    err = max(abs(u[g] - x[0] * x[1] * x[2]) for g, x in mesh.vertex_xyz.items())
    assert err < 1e-14
    eng.close()


@pytest.mark.parametrize("kind,N,p", [(3, 2, 2), (3, 2, 3), (4, 2, 2), (4, 2, 3)])
def test_maxwell_solution_matches_oracle_path(oracle, gpu, kind, N, p):
    """Complex builds (configs 2 and 3 of BASELINE.json): the global skeleton system of a PEC cavity assembled from the GPU's
    condensed matrices gives the same solution as the one assembled from the oracle's (north_star: 1e-10), and the bubble
    dofs recovered with the Schur factors agree too.  Galerkin: H(curl) field; ultraweak DPG: E/H traces (2 components)."""
    from hp3d_b200.api import ElemEngine
    from tests.mini_fem import CubeMeshHcurl
    oracle.set_maxp(6)
    om = 2 * np.pi if kind == 4 else np.pi
    mesh = CubeMeshHcurl(gpu, N, p, ncomp=2 if kind == 4 else 1)
    no, oe, of, X = mesh.descriptors()
    prm = oracle.default_params(omega=om)
    ref = [oracle.condensed(kind, no[e], oe[e], of[e], X[e], prm) for e in range(len(no))]
    assert ref[0][0].shape[0] == mesh.l2g.shape[1]
    u_ref, K, F, free = mesh.solve(np.array([r[0] for r in ref]), np.array([r[1] for r in ref]))
    eng = ElemEngine(kind, omega=om)
    res = eng.elem_stc_batch(no, oe, of, X)
    assert (res["info"] == 0).all()
    out = [eng.unpack(res, e) for e in range(len(no))]
    u_gpu, Kg, Fg, _ = mesh.solve(np.array([o[0] for o in out]), np.array([o[1] for o in out]))
    scale = max(1.0, np.abs(u_ref).max())
    assert np.abs(u_gpu - u_ref).max() < 1e-10 * scale
    # the GPU-assembled system is solved by the oracle path's solution to rounding
    r = Kg[np.ix_(free, free)] @ u_ref[free] - Fg[free]
    assert np.linalg.norm(r) < 1e-10 * max(np.linalg.norm(Fg[free]), 1e-300)
    xi = u_gpu[mesh.l2g]
    AS = np.array([o[2] for o in out]); BS = np.array([o[3] for o in out])
    xb = eng.stc_bwd_batch(AS, BS, xi)
    xb_ref = np.array([r_[3] - r_[2] @ u_ref[mesh.l2g[e]] for e, r_ in enumerate(ref)])
    assert np.abs(xb - xb_ref).max() < 1e-9 * max(1.0, np.abs(xb_ref).max())
    eng.close()
