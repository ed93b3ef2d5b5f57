"""NR_RHS > 1 (src/modules/stc.F90:223-257: Bi(ni,NR_RHS), CLOC%BSchur(nb,NR_RHS) travel through the condensation together): the
q-th column of Bi / BSchur of ONE call with nr_rhs load vectors == what the oracle gives for the q-th source table alone, the matrices
are those of the single-load call, and the back-substitutions (stored factors, device-resident store, recompute) act column by column."""
import ctypes as C

import numpy as np
import pytest

from tests.test_gpu_bwd_residual import _batch, relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,rr", [(2, 1), (4, 1), (4, 0)])
def test_several_load_vectors(oracle, gpu, kind, rr):
    from hp3d_b200.api import ElemEngine
    oracle.set_maxp(8)
    oracle.use_blas(True)
    rng = np.random.default_rng(210 + kind + rr)
    items, et, norder, norie, norif, X = _batch(oracle, rng, nel=4, pmax=3)   # bricks and prisms mixed
    nel, nr = len(items), 3
    om = 2 * np.pi if kind == 4 else 1.0
    cplx = kind >= 3
    vals = 3 if cplx else 1
    one = ElemEngine(kind, omega=om, maxp=8, real_reduction=rr, source=9)
    nint = np.array([one.sig_dims(norder[e], norie[e], norif[e], int(et[e]))["nint"] for e in range(nel)])   # the signature's own rule (orientation dependent)
    nmax = int(nint.max())
    # element e: nr consecutive tables of nint[e] points each (the layout hp3d_params.nr_rhs documents)
    J = np.zeros((nel, nr * nmax * vals), np.complex128 if cplx else np.float64)
    tabs = []
    for e in range(nel):
        t = rng.standard_normal((nr, nint[e], vals)) + (1j * rng.standard_normal((nr, nint[e], vals)) if cplx else 0)
        tabs.append(t)
        J[e, :nr * nint[e] * vals] = t.reshape(-1)
    eng = ElemEngine(kind, omega=om, maxp=8, real_reduction=rr, source=9, nr_rhs=nr)
    res = eng.elem_stc_batch(norder, norie, norif, X, source_qp=J, etype=et)
    assert (res["info"] == 0).all()
    for e, it in enumerate(items):
        ni, nb = int(res["ni"][e]), int(res["nb"][e])
        Aii = res["Aii"][e, :ni * ni].reshape(ni, ni).T
        AS = res["ASchur"][e, :nb * ni].reshape(ni, nb).T
        for q in range(nr):
            tab = np.ascontiguousarray(tabs[e][q].reshape(nint[e], vals) if cplx else tabs[e][q].reshape(nint[e]))
            prm = oracle.default_params(omega=om, source=9, source_table=tab.ctypes.data_as(C.c_void_p))
            rA, rB, rAS, rBS = oracle.condensed(kind, it[1], it[2], it[3], it[4], prm, etype=it[0])
            assert relerr(res["Bi"][e, q * ni:(q + 1) * ni], rB) < 1e-12, (e, q)
            if nb:
                assert relerr(res["BSchur"][e, q * nb:(q + 1) * nb], rBS) < 1e-9, (e, q)
            if q == 0:
                assert relerr(Aii, rA) < 1e-12
                if nb:
                    assert relerr(AS, rAS) < 1e-9
    # back-substitution, three ways, on nr columns at once
    nim, nbm = int(res["ni"].max()), int(res["nb"].max())
    xi = np.zeros((nel, nim * nr), eng.dtype)
    for e in range(nel):
        ni = int(res["ni"][e])
        xi[e, :ni * nr] = rng.standard_normal(ni * nr) + (1j * rng.standard_normal(ni * nr) if cplx else 0)

    def expect(e, q):
        ni, nb = int(res["ni"][e]), int(res["nb"][e])
        AS = res["ASchur"][e, :nb * ni].reshape(ni, nb).T
        return res["BSchur"][e, q * nb:(q + 1) * nb] - AS @ xi[e, q * ni:(q + 1) * ni]

    cl = eng.cloc_create()
    r2 = eng.elem_stc_batch_cloc(cl, norder, norie, norif, X, source_qp=J, etype=et)
    assert np.array_equal(r2["Bi"], res["Bi"])
    out = eng.cloc_bwd_batch(cl, xi, nb_max=nbm)
    for e in range(nel):
        nb = int(res["nb"][e])
        for q in range(nr):
            if nb:
                assert relerr(out["xb"][e, q * nb:(q + 1) * nb], expect(e, q)) < 1e-12, (e, q)
        if nb:
            _, BS = eng.cloc_fetch(cl, e)
            assert np.array_equal(BS, res["BSchur"][e, :nb * nr])
    eng.cloc_destroy(cl)
    # a store too small for anything: every element is spilled and recomputed, all columns
    cl = eng.cloc_create(limit_bytes=600)
    eng.elem_stc_batch_cloc(cl, norder, norie, norif, X, source_qp=J, etype=et)
    assert eng.cloc_stats(cl)["spilled"] >= 1
    out = eng.cloc_bwd_batch(cl, xi, nb_max=nbm)
    for e in range(nel):
        nb = int(res["nb"][e])
        for q in range(nr):
            if nb:
                assert relerr(out["xb"][e, q * nb:(q + 1) * nb], expect(e, q)) < 1e-11, (e, q)
    eng.cloc_destroy(cl)
    eng.close(); one.close()


def test_nr_rhs_is_refused_where_it_is_not_carried(gpu):
    from hp3d_b200.api import ElemEngine
    with pytest.raises(RuntimeError, match="nr_rhs"):
        ElemEngine(1, nr_rhs=2, source=9)            # pivoted-LU condensation
    with pytest.raises(RuntimeError, match="nr_rhs"):
        ElemEngine(4, nr_rhs=2)                      # built-in manufactured source
    eng = ElemEngine(2, nr_rhs=2, source=9)
    from tests.util import hexa_xnod, uniform_order
    no = uniform_order(2)[None]; z12 = np.zeros((1, 12), np.int32); z6 = np.zeros((1, 6), np.int32)
    X = hexa_xnod(27, h=0.5, jitter=0.0, rng=np.random.default_rng(0))[None]
    with pytest.raises(RuntimeError, match="one load vector"):
        eng.elem_residual_batch(no, z12, z6, X, np.zeros((1, 400)), np.zeros((1, 64)))
    eng.close()
