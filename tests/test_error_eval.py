"""SURVEY 8f row f4 (error-evaluation half): soleval + element_error.

CPU: pins of the oracle's restatement (oracle/soleval.c) -- closed-form norms of the manufactured solutions, exact
reproduction of polynomial fields set through their dofs, Piola maps against the definition.  GPU: hp3d_gpu_elem_error_batch
against the oracle (1e-12 relative), bricks and prisms, all four problem kinds, built-in and tabulated exact solutions, and the
error of an actual GPU solve (conv_pois.F90's quantity)."""
import numpy as np
import pytest

from tests.util import VERT, hexa_xnod, uniform_order


def _unit_cube(O, p):
    no = uniform_order(p)
    nH = O.celndof(no)[0]
    X = np.zeros((nH, 3)); X[:8] = VERT
    return no, np.zeros(12, np.int32), np.zeros(6, np.int32), X


def test_norms_of_manufactured_solutions(oracle):
    """rnorm over the unit cube = closed form (quadrature of order p+2: 1e-5): Poisson 1/8 + 3 pi^2/8 ; ultraweak Maxwell
    |E|^2 + |H|^2 = 1/4 + 1/2 ; Galerkin Maxwell |E|^2 + |curl E|^2 = 1/4 + w^2/2."""
    O = oracle
    O.set_maxp(6)
    no, oe, of, X = _unit_cube(O, 3)
    nH, nE, _, nQ = O.celndof(no)
    e, r, n = O.element_error(1, no, oe, of, X, np.zeros((nH, 1)), O.default_params())
    assert n == 216 and abs(r - (1 / 8 + 3 * np.pi ** 2 / 8)) < 1e-5 and e == r      # zero dofs: error = norm
    e, r, _ = O.element_error(1, no, oe, of, X, np.zeros((nH, 1)), O.default_params(), l2proj=True)
    assert abs(r - 1 / 8) < 1e-6
    om = 2 * np.pi
    e, r, _ = O.element_error(4, no, oe, of, X, np.zeros((nQ, 6)), O.default_params(omega=om))
    assert abs(r - 0.75) < 2e-3
    e, r, _ = O.element_error(3, no, oe, of, X, np.zeros((nE, 1)), O.default_params(omega=np.pi))
    assert abs(r - (0.25 + np.pi ** 2 / 2)) < 1e-4


def test_polynomials_are_evaluated_exactly(oracle):
    """A trilinear field set through the vertex dofs of a (distorted, straight-edged) brick is its own interpolant: the error
    against a table of the same polynomial is rounding; an L2 field with only the constant Legendre dof equals c/det."""
    O = oracle
    O.set_maxp(6)
    rng = np.random.default_rng(3)
    no = uniform_order(2)
    nH, _, _, nQ = O.celndof(no)
    X = hexa_xnod(nH, h=0.7, jitter=0.1, rng=rng)
    oe = rng.integers(0, 2, 12).astype(np.int32); of = rng.integers(0, 8, 6).astype(np.int32)
    u = lambda x: 1.0 + 2 * x[0] - x[1] + 0.5 * x[2]      # noqa: E731  (affine: exactly representable on any trilinear brick)
    z = np.zeros((nH, 1)); z[:8, 0] = [u(X[v]) for v in range(8)]
    xq = O.error_points(no, oe, of, X)
    tab = np.array([[u(x), 2.0, -1.0, 0.5] for x in xq])
    e, r, n = O.element_error(1, no, oe, of, X, z, O.default_params(), exact_tab=tab)
    assert n == len(xq) and e < 1e-26 * max(r, 1) and r > 0
    # L2: dof of P0 P0 P0 only -> u_h = c / det J ; on the affine unit cube det = 1
    no1, oe1, of1, X1 = _unit_cube(O, 2)
    zq = np.zeros((nQ, 6), complex); zq[0] = [1, 2j, 3, 4, 5, 6 - 1j]
    xq1 = O.error_points(no1, oe1, of1, X1)
    tabq = np.tile(zq[0], (len(xq1), 1))
    e, r, _ = O.element_error(4, no1, oe1, of1, X1, zq, O.default_params(omega=1.0), exact_tab=tabq)
    assert e < 1e-26 * r


# ----------------------------------------------------------------------------------------------------------------------------
def relerr(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,p", [(1, 2), (1, 4), (2, 3), (3, 2), (3, 3), (4, 2), (4, 3)])
def test_gpu_element_error_vs_oracle_bricks(oracle, gpu, kind, p):
    from hp3d_b200.api import ElemEngine
    O = oracle
    O.set_maxp(6)
    rng = np.random.default_rng(900 + 10 * kind + p)
    nel = 3
    no = np.tile(uniform_order(p), (nel, 1))
    oe = rng.integers(0, 2, (nel, 12)).astype(np.int32); of = rng.integers(0, 8, (nel, 6)).astype(np.int32)
    nH, nE, _, nQ = O.celndof(no[0])
    X = np.stack([hexa_xnod(nH, h=0.5, origin=(0.1 * e, 0.2, 0.1), jitter=0.12, curved=0.01, rng=rng) for e in range(nel)])
    nF, nc = {1: (nH, 1), 2: (nH, 1), 3: (nE, 1), 4: (nQ, 6)}[kind]
    cplx = kind >= 3
    z = rng.standard_normal((nel, nF, nc)) + (1j * rng.standard_normal((nel, nF, nc)) if cplx else 0)
    om = 2 * np.pi if kind == 4 else (np.pi if kind == 3 else 1.0)
    prm = O.default_params(omega=om, icomp_exact=2)
    eng = ElemEngine(kind, omega=om, icomp_exact=2)
    for l2 in (False, True):
        res = eng.elem_error_batch(no, oe, of, X, z, l2proj=l2)
        assert (res["info"] == 0).all()
        for e in range(nel):
            er, rn, _ = O.element_error(kind, no[e], oe[e], of[e], X[e], z[e], prm, l2proj=l2)
            assert relerr(res["err"][e], er) < 1e-12 and relerr(res["rnorm"][e], rn) < 1e-12, (l2, e)
    # tabulated exact solution at the points the library reports
    xq, nint = eng.error_points(no, oe, of, X)
    nv = O.error_nvals(kind)
    tab = rng.standard_normal((nel, xq.shape[1], nv)) + (1j * rng.standard_normal((nel, xq.shape[1], nv)) if cplx else 0)
    res = eng.elem_error_batch(no, oe, of, X, z, exact_qp=tab)
    for e in range(nel):
        assert np.abs(xq[e, :nint[e]] - O.error_points(no[e], oe[e], of[e], X[e])).max() < 1e-13
        er, rn, _ = O.element_error(kind, no[e], oe[e], of[e], X[e], z[e], prm, exact_tab=tab[e, :nint[e]])
        assert relerr(res["err"][e], er) < 1e-12 and relerr(res["rnorm"][e], rn) < 1e-12
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", [1, 3, 4])
def test_gpu_element_error_vs_oracle_prisms_mixed(oracle, gpu, kind):
    """Prisms and bricks of different orders in one call (grouped by signature internally)."""
    from hp3d_b200.api import ElemEngine
    from tests.test_gpu_prism import prism_xnod
    from tests.test_oracle_prism import prism_signature
    O = oracle
    O.set_maxp(8)
    rng = np.random.default_rng(77 + kind)
    cplx = kind >= 3
    specs = []
    for (et, p, pz) in [(O.MDLP, 2, 2), (O.MDLB, 2, 0), (O.MDLP, 3, 2), (O.MDLP, 2, 2)]:
        if et == O.MDLP:
            no, ne, nf = prism_signature(rng, p, pz, uniform=True)
            nHe = O.celndof(no, et)[0]
            X = prism_xnod(nHe, rng, curved=0.01)
        else:
            no = uniform_order(p); ne = rng.integers(0, 2, 12); nf = rng.integers(0, 8, 6)
            nHe = O.celndof(no)[0]
            X = hexa_xnod(nHe, h=0.5, jitter=0.1, rng=rng)
        specs.append((et, no, ne, nf, X))
    nel = len(specs)
    NO = np.zeros((nel, 19), np.int32); OE = np.zeros((nel, 12), np.int32); OF = np.zeros((nel, 6), np.int32); ET = np.zeros(nel, np.int32)
    cnt = [O.celndof(s[1], s[0]) for s in specs]
    nFs = [{1: c[0], 3: c[1], 4: c[3]}[kind] for c in cnt]
    nc = 6 if kind == 4 else 1
    XX = np.zeros((nel, max(c[0] for c in cnt), 3)); Z = np.zeros((nel, max(nFs), nc), complex if cplx else float)
    for e, (et, no, ne, nf, X) in enumerate(specs):
        ET[e] = et; NO[e, :len(no)] = no; OE[e, :len(ne)] = ne; OF[e, :len(nf)] = nf; XX[e, :len(X)] = X
        Z[e, :nFs[e]] = rng.standard_normal((nFs[e], nc)) + (1j * rng.standard_normal((nFs[e], nc)) if cplx else 0)
    om = 2 * np.pi if kind == 4 else (np.pi if kind == 3 else 1.0)
    prm = O.default_params(omega=om)
    eng = ElemEngine(kind, omega=om, maxp=8)
    res = eng.elem_error_batch(NO, OE, OF, XX, Z, etype=ET)
    assert (res["info"] == 0).all()
    for e, (et, no, ne, nf, X) in enumerate(specs):
        er, rn, _ = O.element_error(kind, no, ne, nf, X, Z[e, :nFs[e]], prm, etype=et)
        assert relerr(res["err"][e], er) < 1e-12 and relerr(res["rnorm"][e], rn) < 1e-12, e
    eng.close()


@pytest.mark.gpu
def test_error_of_a_gpu_poisson_solve_converges(oracle, gpu):
    """conv_pois.F90's quantity with every numeric step on the device: element matrices + condensation, (host) assembly and
    solve, bubbles by stc_bwd, H1 error by hp3d_gpu_elem_error_batch: rate ~ p between N = 2 and N = 4 at p = 2."""
    from hp3d_b200.api import ElemEngine
    from tests.mini_fem import CubeMeshH1
    errs = []
    for N in (2, 4):
        mesh = CubeMeshH1(gpu, N, 2)
        no, oe, of, X = mesh.descriptors()
        eng = ElemEngine(1)
        res = eng.elem_stc_batch(no, oe, of, X)
        out = [eng.unpack(res, e) for e in range(len(no))]
        u = mesh.solve(np.array([o[0] for o in out]), np.array([o[1] for o in out]))
        xi = mesh.local_interface(u)
        xb = eng.stc_bwd_batch(np.array([o[2] for o in out]), np.array([o[3] for o in out]), xi)
        z = np.concatenate([xi, xb], axis=1)[:, :, None]
        r = eng.elem_error_batch(no, oe, of, X, z)
        errs.append(np.sqrt(r["err"].sum() / r["rnorm"].sum()))
        eng.close()
    rate = np.log(errs[0] / errs[1]) / np.log(2.0)
    assert errs[1] < 0.03 and 1.7 < rate < 2.4, (errs, rate)
