"""GPU parity of hp3d_gpu_celem_batch: elem + stc_fwd_wrapper + constrained-approximation transform,
Dirichlet lift, compression (celem_systemI.F90:543-785) and the COO indices of par_mumps_sc.F90:433-448, all on the
device, against the oracle's restatement of the host loops applied to the condensed systems of hp3d_gpu_elem_batch.

Bar: BIT-EXACT.  The transform is a gather with sums formed in the reference's loop order and without fused
multiply-adds (index work + exact products for unconstrained dofs), so Zastif / Zbload / IRN / JCN must equal the oracle's
to the last bit -- also for constrained dofs and the Dirichlet lift.
"""
import numpy as np
import pytest

from tests import celem_util as CU
from tests.util import hexa_xnod, uniform_order

pytestmark = pytest.mark.gpu


def _batch(O, rng, kind, specs):
    """specs: list of (etype, norder); returns descriptor arrays with jittered geometry."""
    nel = len(specs)
    no = np.zeros((nel, 19), np.int32); noe = np.zeros((nel, 12), np.int32); nof = np.zeros((nel, 6), np.int32)
    et = np.zeros(nel, np.int32)
    nHs = [O.celndof(s[1], s[0])[0] for s in specs]
    X = np.zeros((nel, max(nHs), 3))
    for e, (t, o) in enumerate(specs):
        et[e] = t; no[e, :len(o)] = o
        if t == O.MDLB:
            noe[e] = rng.integers(0, 2, 12); nof[e] = rng.integers(0, 8, 6)
            X[e, :nHs[e]] = hexa_xnod(nHs[e], h=0.5, jitter=0.1, rng=rng)
        else:
            noe[e, :9] = rng.integers(0, 2, 9); nof[e, :2] = rng.integers(0, 6, 2); nof[e, 2:5] = rng.integers(0, 8, 3)
            X[e, :6] = 0.5 * np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1]], float) + 0.1 + rng.uniform(-0.03, 0.03, (6, 3))
    return et, no, noe, nof, X


@pytest.mark.parametrize("kind,p", [(1, 3), (2, 2), (3, 2), (4, 2), (4, 3)])
def test_celem_bit_exact_vs_oracle(oracle, gpu, kind, p):
    from hp3d_b200 import api
    O = oracle
    O.set_maxp(6)
    rng = np.random.default_rng(300 + 10 * kind + p)
    cplx = kind >= 3
    nel = 7
    et, no, noe, nof, X = _batch(O, rng, kind, [(O.MDLB, uniform_order(p))] * nel)
    eng = api.ElemEngine(kind, omega=2 * np.pi if kind == 4 else 1.0)
    ref = eng.elem_stc_batch(no, noe, nof, X)
    assert (ref["info"] == 0).all()
    cons = [CU.random_constraints(rng, O, api, kind, no[e], O.MDLB, cplx, frac_con=(0.0 if e == 0 else 0.3), frac_dbc=(0.0 if e == 1 else 0.2), dof0=1 + 1000 * e)
            for e in range(nel)]
    for isym in (2, 1, 3):
        coo = isym != 1
        res = eng.celem_batch(no, noe, nof, X, cons, isym_flag=isym, want_coo=coo, want_schur=(isym == 2))
        assert (res["info"] == 0).all()
        for e in range(nel):
            Aii, Bi, AS, BS = eng.unpack(ref, e)
            zb, za = CU.oracle_celem(O, cons[e], Aii, Bi, isym)
            x0, x1, a0, a1 = res["xptr"][e], res["xptr"][e + 1], res["aptr"][e], res["aptr"][e + 1]
            assert np.array_equal(res["zbload"][x0:x1], zb), (isym, e)
            assert np.array_equal(res["zastif"][a0:a1], za), (isym, e)
            if coo:
                n = x1 - x0
                lc = cons[e]["lcon"]
                R, Cc = np.repeat(lc[:, None], n, 1), np.repeat(lc[None, :], n, 0)
                if isym == 3:
                    R, Cc = R.T, Cc.T
                assert np.array_equal(res["irn"][a0:a1].reshape(n, n), R) and np.array_equal(res["jcn"][a0:a1].reshape(n, n), Cc)
            if isym == 2:   # the Schur factors still come back unchanged
                ni, nb = int(res["ni"][e]), int(res["nb"][e])
                assert np.array_equal(res["ASchur"][e][: nb * ni], ref["ASchur"][e][: nb * ni]) and np.array_equal(res["BSchur"][e][:nb], ref["BSchur"][e][:nb])
    eng.close()


def test_celem_mixed_orders_and_prisms(oracle, gpu):
    """Heterogeneous call: bricks of different orders and prisms, several chunks (max chunk 3), one shared Zastif array."""
    from hp3d_b200 import _lib, api
    O = oracle
    O.set_maxp(6)
    rng = np.random.default_rng(77)
    specs = [(O.MDLB, uniform_order(2)), (O.MDLP, O.uniform_order(2, O.MDLP, 2)), (O.MDLB, uniform_order(3)), (O.MDLB, uniform_order(2)),
             (O.MDLP, O.uniform_order(3, O.MDLP, 2)), (O.MDLB, uniform_order(2)), (O.MDLB, uniform_order(2)), (O.MDLB, uniform_order(3)),
             (O.MDLB, uniform_order(2)), (O.MDLB, uniform_order(2)), (O.MDLB, uniform_order(2))]
    et, no, noe, nof, X = _batch(O, rng, 4, specs)
    eng = api.ElemEngine(4, omega=2 * np.pi)
    ref = eng.elem_stc_batch(no, noe, nof, X, etype=et)
    cons = [CU.random_constraints(rng, O, api, 4, no[e], int(et[e]), True, dof0=1 + 5000 * e) for e in range(len(specs))]
    _lib.check(gpu.hp3d_gpu_set_chunk(3))
    try:
        res = eng.celem_batch(no, noe, nof, X, cons, isym_flag=2, want_coo=True, etype=et)
    finally:
        _lib.check(gpu.hp3d_gpu_set_chunk(0))
    assert (res["info"] == 0).all()
    for e in range(len(specs)):
        Aii, Bi, _, _ = eng.unpack(ref, e)
        zb, za = CU.oracle_celem(O, cons[e], Aii, Bi, 2)
        assert np.array_equal(res["zbload"][res["xptr"][e]:res["xptr"][e + 1]], zb), e
        assert np.array_equal(res["zastif"][res["aptr"][e]:res["aptr"][e + 1]], za), e
        a, irn, jcn, _ = O.coo_fill(cons[e]["lcon"], za, zb, int(cons[e]["lcon"].max()))
        assert np.array_equal(res["irn"][res["aptr"][e]:res["aptr"][e + 1]], irn) and np.array_equal(res["jcn"][res["aptr"][e]:res["aptr"][e + 1]], jcn)
    eng.close()


def test_hanging_nodes_reproduce_xyz_gpu(oracle, gpu):
    """poly_pois.F90's criterion on a 1-irregular mesh, with the element matrices, the constraint transform and the
    Dirichlet lift all computed on the device."""
    from hp3d_b200 import api
    from tests.test_celem_oracle import hanging_solve
    O = oracle
    mesh = CU.HangingMesh(N=3, ref=(1, 1, 1))
    uex = lambda x: x[0] * x[1] * x[2]           # noqa: E731
    free = np.flatnonzero(~mesh.bdry)
    numbering = {int(g): i + 1 for i, g in enumerate(free)}
    no, noe, nof, X = mesh.descriptors()
    cons = mesh.constraints(api, O, uex, numbering)
    eng = api.ElemEngine(1, source=api.SRC_ZERO)
    res = eng.celem_batch(no, noe, nof, X, cons, isym_flag=2)
    assert (res["info"] == 0).all()
    systems = [(cons[e], (res["zbload"][res["xptr"][e]:res["xptr"][e + 1]], res["zastif"][res["aptr"][e]:res["aptr"][e + 1]])) for e in range(len(cons))]
    u = hanging_solve(O, mesh, systems)
    exact = np.array([uex(mesh.xyz[g]) for g in free])
    assert np.abs(u - exact).max() < 1e-13
    eng.close()


def test_celem_rejects_bad_indices(oracle, gpu):
    from hp3d_b200 import api
    O = oracle
    rng = np.random.default_rng(5)
    et, no, noe, nof, X = _batch(O, rng, 1, [(O.MDLB, uniform_order(2))] * 2)
    eng = api.ElemEngine(1)
    cons = [CU.random_constraints(rng, O, api, 1, no[e], O.MDLB, False) for e in range(2)]
    bad = [dict(c) for c in cons]
    bad[1]["cidx"] = bad[1]["cidx"].copy(); bad[1]["cidx"][0] = 10 ** 6
    with pytest.raises(RuntimeError, match="cidx"):
        eng.celem_batch(no, noe, nof, X, bad)
    bad = [dict(c) for c in cons]
    bad[0]["nextract"] = bad[0]["nextract"].copy(); bad[0]["nextract"][0] = 0
    with pytest.raises(RuntimeError, match="NEXTRACT"):
        eng.celem_batch(no, noe, nof, X, bad)
    with pytest.raises(RuntimeError, match="ISYM_FLAG"):
        eng.celem_batch(no, noe, nof, X, cons, isym_flag=4)
    with pytest.raises(RuntimeError, match="IRN/JCN"):
        eng.celem_batch(no, noe, nof, X, cons, isym_flag=1, want_coo=True)
    res = eng.celem_batch(no[:0], noe[:0], nof[:0], X[:0], [])      # empty batch
    assert res["info"].size == 0
    eng.close()


def test_celem_regular_mesh_without_lists(oracle, gpu):
    """cptr = NULL: on a regular mesh the caller passes no constraint lists at all; the result equals the explicit identity lists
    (and is a pure permutation / extraction of the condensed matrices)."""
    from hp3d_b200 import api
    O = oracle
    rng = np.random.default_rng(8)
    nel = 5
    et, no, noe, nof, X = _batch(O, rng, 4, [(O.MDLB, uniform_order(2))] * nel)
    eng = api.ElemEngine(4, omega=2 * np.pi)
    full = [CU.random_constraints(rng, O, api, 4, no[e], O.MDLB, True, frac_con=0.0, frac_dbc=0.25, extra=0, dof0=1 + 700 * e) for e in range(nel)]
    # make the explicit lists the identity (random_constraints permutes): modified dof ll <- element dof ll
    for c in full:
        n = len(c["idbc"])
        c["cptr"] = np.arange(n + 1, dtype=np.int64); c["cidx"] = np.arange(1, n + 1, dtype=np.int32); c["cval"] = np.ones(n)
    bare = [{k: v for k, v in c.items() if k not in ("cptr", "cidx", "cval")} for c in full]
    a = eng.celem_batch(no, noe, nof, X, full, isym_flag=2, want_coo=True)
    b = eng.celem_batch(no, noe, nof, X, bare, isym_flag=2, want_coo=True)
    for k in ("zbload", "zastif", "irn", "jcn"):
        assert np.array_equal(a[k], b[k]), k
    ref = eng.elem_stc_batch(no, noe, nof, X)
    for e in range(nel):
        Aii, Bi, _, _ = eng.unpack(ref, e)
        x = full[e]["nextract"] - 1
        if not full[e]["idbc"].any():
            assert np.array_equal(b["zastif"][b["aptr"][e]:b["aptr"][e + 1]].reshape(len(x), len(x)), Aii[np.ix_(x, x)])
    eng.close()
