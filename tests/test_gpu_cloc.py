"""Device-resident CLOC (stc.F90:45-58,273-277,529-677 with STORE_STC = .true., the factors kept in HBM) and the packed
Hermitian Aii (hp3d_params.aii_packed): both must reproduce what hp3d_gpu_elem_batch returns to the host, and that is what the
parity tests compare with the oracle."""
import numpy as np
import pytest

from tests.test_gpu_bwd_residual import _batch, relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", [1, 2, 3, 4])
def test_cloc_store_matches_host_factors(oracle, gpu, kind):
    """factors filed in the store == factors returned to the host (bit for bit); stc_bwd on the store == BSchur - ASchur xi of the
    oracle's factors; elements addressed through arbitrary caller indices, in a different order than they were condensed"""
    from hp3d_b200.api import ElemEngine
    oracle.set_maxp(8)
    rng = np.random.default_rng(140 + kind)
    items, et, norder, norie, norif, X = _batch(oracle, rng, nel=6)
    om = 2 * np.pi if kind == 4 else 1.0
    eng = ElemEngine(kind, omega=om, maxp=8)
    ref = eng.elem_stc_batch(norder, norie, norif, X, etype=et)
    cl = eng.cloc_create()
    iel = np.array([1007, 3, 55, 20000000000, 9, 12], np.int64)
    res = eng.elem_stc_batch_cloc(cl, norder, norie, norif, X, iel=iel, etype=et)
    assert (res["info"] == 0).all() and (res["ni"] == ref["ni"]).all() and (res["nb"] == ref["nb"]).all()
    nel = len(items)
    for e in range(nel):
        ni, nb = int(ref["ni"][e]), int(ref["nb"][e])
        A0, b0, AS0, BS0 = eng.unpack(ref, e)
        assert np.array_equal(res["Aii"][e, :ni * ni], ref["Aii"][e, :ni * ni]) and np.array_equal(res["Bi"][e, :ni], ref["Bi"][e, :ni])
        if nb:
            AS, BS = eng.cloc_fetch(cl, int(iel[e]))
            assert np.array_equal(AS, AS0) and np.array_equal(BS, BS0)
    st = eng.cloc_stats(cl)
    assert st["resident"] == nel and st["spilled"] == 0
    ni_max = int(ref["ni"].max())
    xi = rng.normal(size=(nel, ni_max)) + (1j * rng.normal(size=(nel, ni_max)) if kind >= 3 else 0)
    perm = rng.permutation(nel)
    out = eng.cloc_bwd_batch(cl, xi[perm], iel=iel[perm])
    prm = oracle.default_params(omega=om)
    for k, e in enumerate(perm):
        ni, nb = int(ref["ni"][e]), int(ref["nb"][e])
        assert out["nb"][k] == nb
        if nb == 0:
            continue
        _, _, AS0, BS0 = eng.unpack(ref, e)
        assert relerr(out["xb"][k, :nb], BS0 - AS0 @ xi[e, :ni]) < 1e-13
        it = items[e]
        _, _, rAS, rBS = oracle.condensed(kind, it[1], it[2], it[3], it[4], prm, etype=it[0])
        assert relerr(out["xb"][k, :nb], rBS - rAS @ xi[e, :ni]) < 1e-9
    with pytest.raises(RuntimeError, match="has not been condensed"):
        eng.cloc_bwd_batch(cl, xi[:1], iel=np.array([4242], np.int64), nb_max=1)
    eng.cloc_destroy(cl)
    eng.close()


@pytest.mark.parametrize("kind", [2, 4])
def test_cloc_spill_recomputes(oracle, gpu, kind):
    """a store too small for the batch: the first elements are resident, the rest spilled and recomputed by stc_bwd -- same xb"""
    from hp3d_b200.api import ElemEngine
    oracle.set_maxp(8)
    om = 2 * np.pi if kind == 4 else 1.0
    eng = ElemEngine(kind, omega=om, maxp=8)
    es = 16 if kind >= 3 else 8
    for seed in range(150, 170):   # a batch in which at least two elements have bubbles (low-order H1 elements have none)
        rng = np.random.default_rng(seed + kind)
        items, et, norder, norie, norif, X = _batch(oracle, rng, nel=6, pmax=4)
        ref = eng.elem_stc_batch(norder, norie, norif, X, etype=et)
        need = [es * int(ref["nb"][e]) * (int(ref["ni"][e]) + 1) for e in range(6)]
        if sum(1 for v in need if v > 0) >= 2:
            break
    j = max(e for e in range(6) if need[e] > 0)   # the last element that needs room does not get it: it and everything after it spills
    assert j >= 1
    cl = eng.cloc_create(limit_bytes=sum(need[:j]) + 512 + 8)
    eng.elem_stc_batch_cloc(cl, norder, norie, norif, X, etype=et)
    st = eng.cloc_stats(cl)
    spilled = [e for e in range(6) if eng.cloc_fetch(cl, e) is None]
    assert spilled == list(range(j, 6)) and st["spilled"] == len(spilled) and st["resident"] == j   # slots are granted in caller order
    ni_max = int(ref["ni"].max())
    xi = rng.normal(size=(6, ni_max)) + (1j * rng.normal(size=(6, ni_max)) if kind >= 3 else 0)
    out = eng.cloc_bwd_batch(cl, xi, nb_max=int(ref["nb"].max()))
    assert (out["info"] == 0).all()
    for e in range(6):
        ni, nb = int(ref["ni"][e]), int(ref["nb"][e])
        if nb == 0:
            continue
        _, _, AS0, BS0 = eng.unpack(ref, e)
        assert relerr(out["xb"][e, :nb], BS0 - AS0 @ xi[e, :ni]) < 1e-12
    # condensing again after a clear with room for everything makes all of them resident
    eng.cloc_clear(cl)
    assert eng.cloc_stats(cl)["resident"] == 0
    eng.cloc_destroy(cl)
    eng.close()


@pytest.mark.parametrize("kind,rr", [(2, 1), (4, 1), (4, 0)])
def test_aii_packed_is_the_lower_triangle(oracle, gpu, kind, rr):
    """aii_packed = 1: AP(i + (j-1)(2n-j)/2) = A(i,j), i >= j, bit for bit the full matrix's lower triangle; the host unpack
    restores the full Hermitian matrix (the upper triangle of the full output is the exact conjugate mirror already)"""
    from hp3d_b200.api import ElemEngine
    oracle.set_maxp(8)
    rng = np.random.default_rng(160 + kind)
    items, et, norder, norie, norif, X = _batch(oracle, rng, nel=5)
    om = 2 * np.pi if kind == 4 else 1.0
    full = ElemEngine(kind, omega=om, maxp=8, real_reduction=rr)
    ref = full.elem_stc_batch(norder, norie, norif, X, etype=et)
    full.close()
    eng = ElemEngine(kind, omega=om, maxp=8, real_reduction=rr, aii_packed=1)
    res = eng.elem_stc_batch(norder, norie, norif, X, etype=et)
    assert (res["info"] == 0).all()
    nimax = int(ref["ni"].max())
    assert res["Aii"].shape[1] == nimax * (nimax + 1) // 2
    for e in range(5):
        n = int(ref["ni"][e])
        A = ref["Aii"][e, :n * n].reshape(n, n).T    # A[i, j]
        ap = res["Aii"][e, :n * (n + 1) // 2]
        k = 0
        for j in range(n):
            assert np.array_equal(ap[k:k + n - j], A[j:, j])
            k += n - j
        assert np.array_equal(res["Bi"][e, :n], ref["Bi"][e, :n])
    un = eng.hermitian_unpack(res["Aii"], res["ni"])
    for e in range(5):
        n = int(ref["ni"][e])
        assert np.array_equal(un[e, :n * n], ref["Aii"][e, :n * n])
    with pytest.raises(RuntimeError, match="stride"):
        eng.elem_stc_batch(norder, norie, norif, X, etype=et, out=dict(Aii=np.zeros((5, 10), eng.dtype), Bi=res["Bi"], ASchur=res["ASchur"], BSchur=res["BSchur"]))
    eng.close()
    with pytest.raises(RuntimeError, match="Hermitian"):
        ElemEngine(1, aii_packed=1)


def test_short_strides_are_refused(oracle, gpu):
    """every caller stride is validated before anything is queued (source_ld, sBi, sAS, sBS, sxi)"""
    from hp3d_b200.api import ElemEngine
    rng = np.random.default_rng(170)
    items, et, norder, norie, norif, X = _batch(oracle, rng, nel=2)
    eng = ElemEngine(2, maxp=8)
    ref = eng.elem_stc_batch(norder, norie, norif, X, etype=et)
    good = dict(Aii=ref["Aii"], Bi=ref["Bi"], ASchur=ref["ASchur"], BSchur=ref["BSchur"])
    for key in ("Bi", "ASchur", "BSchur"):
        bad = dict(good)
        bad[key] = np.zeros((2, 1), eng.dtype)
        with pytest.raises(RuntimeError, match="stride"):
            eng.elem_stc_batch(norder, norie, norif, X, etype=et, out=bad)
    with pytest.raises(RuntimeError, match="stride"):
        eng.elem_bwd_batch(norder, norie, norif, X, np.zeros((2, 1)), etype=et)
    eng.close()
    eng = ElemEngine(2, maxp=8, source=9)
    with pytest.raises(RuntimeError, match="source_ld"):
        eng.elem_stc_batch(norder, norie, norif, X, etype=et, source_qp=np.zeros((2, 1)))
    eng.close()


@pytest.mark.parametrize("kind,rr,uniform", [(2, 1, False), (4, 1, False), (4, 0, False), (4, 1, True), (2, 1, True)])
def test_aii_lower_trapezoids_mirrored_on_host(oracle, gpu, kind, rr, uniform):
    """aii_packed = 2: the caller's FULL Aii blocks, bit for bit what aii_packed = 0 returns, although only the lower block
    trapezoids crossed PCIe (output pre-filled with NaN: every entry must have been written by the copy or by the mirror threads).
    uniform: equal sizes and whole-column strides (one strided 3-D copy per block column); else per-element 2-D copies."""
    from hp3d_b200.api import ElemEngine
    from tests.util import hexa_xnod, uniform_order
    oracle.set_maxp(8)
    rng = np.random.default_rng(180 + kind)
    if uniform:
        nel, p = 7, (4 if kind == 2 else 3)
        no = uniform_order(p)
        norder = np.tile(no, (nel, 1)); norie = np.zeros((nel, 12), np.int32); norif = np.zeros((nel, 6), np.int32)
        X = np.stack([hexa_xnod(oracle.celndof(no, oracle.MDLB)[0], h=0.4, jitter=0.1, rng=rng) for _ in range(nel)])
        et = None
    else:
        items, et, norder, norie, norif, X = _batch(oracle, rng, nel=7, pmax=4)
        nel = 7
    om = 2 * np.pi if kind == 4 else 1.0
    full = ElemEngine(kind, omega=om, maxp=8, real_reduction=rr)
    ref = full.elem_stc_batch(norder, norie, norif, X, etype=et)
    full.close()
    assert int(ref["ni"].max()) > 64   # otherwise the trapezoid path is not exercised
    eng = ElemEngine(kind, omega=om, maxp=8, real_reduction=rr, aii_packed=2)
    out = {k: np.full_like(v, np.nan) for k, v in ref.items() if k in ("Aii", "Bi", "ASchur", "BSchur")}
    res = eng.elem_stc_batch(norder, norie, norif, X, etype=et, out=out)
    assert (res["info"] == 0).all()
    for e in range(nel):
        n, nb = int(ref["ni"][e]), int(ref["nb"][e])
        assert np.array_equal(res["Aii"][e, :n * n], ref["Aii"][e, :n * n])
        assert np.array_equal(res["Bi"][e, :n], ref["Bi"][e, :n]) and np.array_equal(res["ASchur"][e, :n * nb], ref["ASchur"][e, :n * nb])
    eng.close()


@pytest.mark.parametrize("kind", [2, 4])
def test_celem_batch_with_device_resident_factors(oracle, gpu, kind):
    """hp3d_gpu_celem_batch_cloc: the compressed systems are bit for bit those of hp3d_gpu_celem_batch, and the factors the fused
    call files in the store give the same back-substitution as the factors the plain call returns to the host"""
    from hp3d_b200 import api
    from tests import celem_util as CU
    from tests.test_gpu_celem import _batch as celem_batch
    from tests.util import uniform_order
    O = oracle
    O.set_maxp(6)
    rng = np.random.default_rng(190 + kind)
    nel, p = 5, 2
    et, no, noe, nof, X = celem_batch(O, rng, kind, [(O.MDLB, uniform_order(p))] * nel)
    eng = api.ElemEngine(kind, omega=2 * np.pi if kind == 4 else 1.0)
    cons = [CU.random_constraints(rng, O, api, kind, no[e], O.MDLB, kind >= 3, frac_con=0.3, frac_dbc=0.2, dof0=1 + 1000 * e) for e in range(nel)]
    ref = eng.celem_batch(no, noe, nof, X, cons, isym_flag=2, want_coo=True, want_schur=True)
    cl = eng.cloc_create()
    iel = np.arange(nel, dtype=np.int64) * 3 + 11
    res = eng.celem_batch(no, noe, nof, X, cons, isym_flag=2, want_coo=True, cloc=cl, iel=iel)
    assert (res["info"] == 0).all()
    for key in ("zbload", "zastif", "irn", "jcn"):
        assert np.array_equal(res[key], ref[key]), key
    ni, nb = int(ref["ni"][0]), int(ref["nb"][0])
    xi = rng.normal(size=(nel, ni)) + (1j * rng.normal(size=(nel, ni)) if kind >= 3 else 0)
    out = eng.cloc_bwd_batch(cl, xi, iel=iel)
    for e in range(nel):
        AS = ref["ASchur"][e][:nb * ni].reshape(ni, nb).T; BS = ref["BSchur"][e][:nb]
        assert relerr(out["xb"][e, :nb], BS - AS @ xi[e]) < 1e-13
    with pytest.raises(RuntimeError, match="no such Schur store"):
        eng.cloc_stats(cl + 17)
    eng.cloc_destroy(cl)
    eng.close()
