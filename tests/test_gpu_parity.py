"""GPU parity tests proper: the CUDA path through the C ABI against the CPU oracle on the same seeded inputs.

Tolerances (north_star): element and condensed matrices within 1e-12 relative Frobenius error.  The Schur
back-substitution factors ASchur = A_bb^-1 A_bi are a *solution* of a linear system, so their forward error scales with
cond(A_bb); they are checked by residual (||A_bb ASchur - A_bi|| / ||A_bi|| via the oracle's own uncondensed matrix)
and by a conditioning-scaled forward bound.
"""
import numpy as np
import pytest

from tests.util import complex_W, hexa_xnod, random_signature, uniform_order

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _engine(kind, **kw):
    from hp3d_b200.api import ElemEngine
    return ElemEngine(kind, **kw)


def _oracle_params(oracle, **kw):
    return oracle.default_params(**kw)


@pytest.mark.parametrize("rr", [1, 0])
@pytest.mark.parametrize("p,curved", [(1, 0.0), (2, 0.0), (2, 0.03), (3, 0.02)])
def test_uw_maxwell_integration_vs_oracle(oracle, gpu, p, curved, rr):
    """Gram matrix and enriched stiffness of ultraweak Maxwell straight out of the integration kernels
    (MAXWELL/ULTRAWEAK_DPG/elem/elem_opt.F90:236-768) vs the oracle's BLAS3 restatement."""
    oracle.set_maxp(6)
    rng = np.random.default_rng(100 + p)
    norder, _, _ = random_signature(rng, uniform=True)
    norder = uniform_order(p)
    norie = rng.integers(0, 2, 12).astype(np.int32); norif = rng.integers(0, 8, 6).astype(np.int32)
    nH = oracle.celndof(norder)[0]
    X = hexa_xnod(nH, h=0.5, jitter=0.15, curved=curved, rng=rng)
    om = 2 * np.pi
    prm = _oracle_params(oracle, omega=om)
    A, b, G, S = oracle.elem(oracle.MAXW_UW, norder, norie, norif, X, prm, want_dpg=True)
    eng = _engine(4, omega=om, real_reduction=rr)   # rr = 1: real-structured storage (one plane), 0: general complex layout
    W, d = eng.integrate_debug(norder, norie, norif, X)
    assert W.shape[0] == (1 if rr else 2)
    n, nb, ni, np_, nbp = d["n"], d["nb"], d["ni"], d["np"], d["nbp"]
    nEE = n // 2
    Wc = complex_W(W, d)
    Gi = np.tril(Wc[:n, :n]); Gi = Gi + np.tril(Gi, -1).conj().T
    perm = np.empty(n, int); perm[0::2] = np.arange(nEE); perm[1::2] = nEE + np.arange(nEE)
    Gg = Gi[np.ix_(perm, perm)]
    Gu = np.triu(G); Go = Gu + np.triu(Gu, 1).conj().T
    assert relerr(Gg, Go) < 1e-13
    rows = np.r_[np_ + nbp + np.arange(ni), np_ + np.arange(nb), np_ + nbp + d["nip"] - 1]   # load: last padded interface row
    Bg = Wc[rows][:, :n].conj().T[perm]
    assert relerr(Bg[:, :ni], S[:, :ni]) < 1e-13          # trace pairings
    assert relerr(Bg[:, ni:ni + nb], S[:, ni:ni + nb]) < 1e-13
    assert relerr(Bg[:, -1], S[:, -1]) < 1e-13            # load
    eng.close()


CASES = [
    # kind, p, nel, tolerances on (Aii, Bi)
    (1, 1, 3), (1, 2, 3), (1, 3, 4), (1, 4, 2),
    (2, 1, 3), (2, 2, 3), (2, 3, 2), (2, 4, 2),
    (3, 1, 3), (3, 2, 3), (3, 3, 2), (3, 4, 2), (3, 5, 1),
    (4, 1, 3), (4, 2, 3), (4, 3, 2), (4, 4, 2),
]


@pytest.mark.parametrize("kind,p,nel", CASES)
def test_condensed_vs_oracle(oracle, gpu, kind, p, nel):
    """elem + stc_fwd_wrapper through hp3d_gpu_elem_batch vs the oracle, uniform order p, random orientations,
    jittered + slightly curved geometry, several elements per call."""
    oracle.set_maxp(6)
    oracle.use_blas(True)
    rng = np.random.default_rng(1000 * kind + p)
    norder = np.tile(uniform_order(p), (nel, 1))
    norie = rng.integers(0, 2, (nel, 12)).astype(np.int32); norif = rng.integers(0, 8, (nel, 6)).astype(np.int32)
    norie[0] = 0; norif[0] = 0
    nH = oracle.celndof(norder[0])[0]
    X = np.stack([hexa_xnod(nH, h=0.5, origin=(0.1 * e, 0.2, 0.3), jitter=0.15, curved=0.01 if p > 1 else 0.0, rng=rng) for e in range(nel)])
    om = 2 * np.pi if kind == 4 else (np.pi if kind == 3 else 1.0)
    prm = _oracle_params(oracle, omega=om)
    eng = _engine(kind, omega=om)
    res = eng.elem_stc_batch(norder, norie, norif, X)
    assert (res["info"] == 0).all()
    for e in range(nel):
        Aii, Bi, AS, BS = eng.unpack(res, e)
        rA, rB, rAS, rBS = oracle.condensed(kind, norder[e], norie[e], norif[e], X[e], prm)
        assert Aii.shape == rA.shape and AS.shape == rAS.shape
        assert relerr(Aii, rA) < 1e-12, (e, relerr(Aii, rA))
        assert relerr(Bi, rB) < 1e-12, (e, relerr(Bi, rB))
        if AS.size:
            # residual check of the stored factors against the oracle's uncondensed element matrix
            Afull, bfull = oracle.elem(kind, norder[e], norie[e], norif[e], X[e], prm)
            perm, ni, nb = oracle.stc_partition(kind, norder[e])
            Ap = Afull[np.ix_(perm, perm)]; bp = bfull[perm]
            Abb, Abi = Ap[ni:, ni:], Ap[ni:, :ni]
            assert relerr(Abb @ AS, Abi) < 1e-12
            assert relerr(Abb @ BS, bp[ni:]) < 1e-12
            cond = np.linalg.cond(Abb)
            assert relerr(AS, rAS) < 1e-15 * cond * 50 + 1e-12
            assert relerr(BS, rBS) < 1e-15 * cond * 50 + 1e-12
    eng.close()


@pytest.mark.parametrize("p", [1, 2, 3])
def test_general_complex_path_matches_real_reduction(oracle, gpu, p):
    """Ultraweak Maxwell has two implementations of the dense phase: the real-structured one (default for real eps, mu and
    sigma = 0: A = T A~ T^H with A~ real) and the general complex kernels (real_reduction = 0, the reference's ZPOTRF / ZTRTRS /
    ZHERK sequence).  Both must match the oracle, and each other, on every output incl. the Schur factors and residuals."""
    oracle.set_maxp(6)
    oracle.use_blas(True)
    rng = np.random.default_rng(4242 + p)
    nel = 3
    norder = np.tile(uniform_order(p), (nel, 1))
    norie = rng.integers(0, 2, (nel, 12)).astype(np.int32); norif = rng.integers(0, 8, (nel, 6)).astype(np.int32)
    nH = oracle.celndof(norder[0])[0]
    X = np.stack([hexa_xnod(nH, h=0.5, jitter=0.15, curved=0.01 if p > 1 else 0.0, rng=rng) for e in range(nel)])
    om = 2 * np.pi
    prm = _oracle_params(oracle, omega=om)
    outs = []
    for rr in (1, 0):
        eng = _engine(4, omega=om, real_reduction=rr)
        res = eng.elem_stc_batch(norder, norie, norif, X)
        assert (res["info"] == 0).all()
        u = [eng.unpack(res, e) for e in range(nel)]
        xi = np.array([rng.standard_normal(u[e][1].size) + 1j * rng.standard_normal(u[e][1].size) for e in range(nel)])
        rng2 = np.random.default_rng(7)
        xi = np.array([rng2.standard_normal(u[e][1].size) + 1j * rng2.standard_normal(u[e][1].size) for e in range(nel)])
        xb = eng.elem_bwd_batch(norder, norie, norif, X, xi)["xb"]
        eta = eng.elem_residual_batch(norder, norie, norif, X, xi, xb)["resid"]
        outs.append((u, xb, eta))
        eng.close()
    for e in range(nel):
        rA, rB, rAS, rBS = oracle.condensed(4, norder[e], norie[e], norif[e], X[e], prm)
        for k in range(2):
            Aii, Bi, AS, BS = outs[k][0][e]
            assert relerr(Aii, rA) < 1e-12 and relerr(Bi, rB) < 1e-12
        for a, b in zip(outs[0][0][e], outs[1][0][e]):
            if a.size:
                assert relerr(a, b) < 1e-10
    assert relerr(outs[0][1], outs[1][1]) < 1e-10          # recomputed back-substitution
    assert np.abs(outs[0][2] - outs[1][2]).max() < 1e-9 * np.abs(outs[1][2]).max()   # DPG residuals


def test_mixed_signatures_one_call(oracle, gpu):
    """Elements of different order and orientation in one batch (grouped by signature internally)."""
    oracle.set_maxp(6)
    rng = np.random.default_rng(7)
    sigs = [random_signature(rng, pmax=3) for _ in range(3)] + [(uniform_order(2), np.zeros(12, np.int32), np.zeros(6, np.int32))]
    order = [0, 3, 1, 0, 2, 3, 1]
    nel = len(order)
    norder = np.stack([sigs[i][0] for i in order]); norie = np.stack([sigs[i][1] for i in order]); norif = np.stack([sigs[i][2] for i in order])
    nHmax = max(oracle.celndof(s[0])[0] for s in sigs)
    X = np.zeros((nel, nHmax, 3))
    for e in range(nel):
        nH = oracle.celndof(norder[e])[0]
        X[e, :nH] = hexa_xnod(nH, h=0.4, jitter=0.1, rng=rng)
    for kind in (1, 2, 3, 4):
        om = 2 * np.pi if kind == 4 else 1.0
        prm = _oracle_params(oracle, omega=om)
        eng = _engine(kind, omega=om)
        res = eng.elem_stc_batch(norder, norie, norif, X)
        assert (res["info"] == 0).all()
        for e in range(nel):
            nH = oracle.celndof(norder[e])[0]
            Aii, Bi, AS, BS = eng.unpack(res, e)
            rA, rB, rAS, rBS = oracle.condensed(kind, norder[e], norie[e], norif[e], X[e, :nH], prm)
            assert Aii.shape == rA.shape
            assert relerr(Aii, rA) < 1e-12, (kind, e, relerr(Aii, rA))
            assert relerr(Bi, rB) < 1e-12, (kind, e)
        eng.close()


def test_negative_jacobian_flag(gpu):
    """geom3D.F90:92-109: a negative Jacobian is reported per element (info = -1) instead of stopping."""
    eng = _engine(1)
    norder = uniform_order(2)[None]
    X = hexa_xnod(27)[None].copy()
    X[0, :8, 0] *= -1.0   # mirror -> negative determinant
    res = eng.elem_stc_batch(norder, np.zeros((1, 12), np.int32), np.zeros((1, 6), np.int32), X)
    assert res["info"][0] == -1
    eng.close()


def test_full_size_p5_uw_maxwell(oracle, gpu):
    """BASELINE.json configs[3] at full size (p=5, dp=1: 1764 test / 1350 trial dofs, ni=600, nb=750), two elements of the
    bench workload, against the oracle; plus the size-independent properties used when the oracle is too slow."""
    from hp3d_b200 import synth
    oracle.set_maxp(6)
    oracle.use_blas(True)
    nel = 2
    norder, noe, nof, xnod = synth.cube_mesh(nel, 5, first=3)
    om = 2 * np.pi
    eng = _engine(4, omega=om)
    res = eng.elem_stc_batch(norder, noe, nof, xnod)
    assert (res["info"] == 0).all() and (res["ni"] == 600).all() and (res["nb"] == 750).all()
    prm = _oracle_params(oracle, omega=om)
    for e in range(nel):
        Aii, Bi, AS, BS = eng.unpack(res, e)
        rA, rB, rAS, rBS = oracle.condensed(4, norder[e], noe[e], nof[e], xnod[e], prm)
        assert relerr(Aii, rA) < 1e-12, relerr(Aii, rA)
        assert relerr(Bi, rB) < 1e-12, relerr(Bi, rB)
        assert relerr(Aii, Aii.conj().T) < 1e-14            # Hermitian (ZHERK + mirror, elem_opt.F90:862-869)
        assert relerr(AS, rAS) < 1e-9 and relerr(BS, rBS) < 1e-9   # forward error ~ cond(A_bb) eps; the 1e-12 residual check is test_schur_factor_residual_at_full_size
    eng.close()


def test_properties_at_full_size(gpu):
    """Size-independent properties on a larger p=5 batch (no oracle): results do not depend on the batch composition or
    chunking, translation of an element leaves its matrix unchanged and only rephases nothing for a zero source, and the
    condensed matrix is Hermitian positive semi-definite (it is a Schur complement of B^H G^-1 B)."""
    from hp3d_b200 import synth
    nel = 6
    norder, noe, nof, xnod = synth.cube_mesh(nel, 5)
    eng = _engine(4, omega=2 * np.pi, source=0)
    a = eng.elem_stc_batch(norder, noe, nof, xnod)
    gpu.hp3d_gpu_set_chunk(2)
    b = eng.elem_stc_batch(norder[::-1].copy(), noe, nof, xnod[::-1].copy())
    gpu.hp3d_gpu_set_chunk(0)
    for e in range(nel):
        A1 = eng.unpack(a, e)[0]; A2 = eng.unpack(b, nel - 1 - e)[0]
        assert np.array_equal(A1, A2)                       # bitwise: same kernels, same data, different batch slot
        assert np.abs(eng.unpack(a, e)[1]).max() == 0.0    # zero source -> zero load
    xs = xnod.copy(); xs[:, :8, :] += np.array([0.25, -0.5, 1.0])
    c = eng.elem_stc_batch(norder, noe, nof, xs)
    for e in range(nel):
        assert relerr(eng.unpack(c, e)[0], eng.unpack(a, e)[0]) < 1e-11
    w = np.linalg.eigvalsh(eng.unpack(a, 0)[0])
    assert w.min() > -1e-10 * w.max()
    eng.close()


@pytest.mark.parametrize("kind,rr", [(3, 1), (3, 0), (4, 1), (4, 0)])
def test_maxwell_caller_source_table(oracle, gpu, kind, rr):
    """HP3D_SRC_TABLE for the complex problems: a caller-supplied (complex, random) source J at the quadrature points
    (`getf` evaluated on the host) must give the same load vectors as the oracle fed with the same table -- through the real
    form (two real load rows) and through the general complex kernels."""
    import ctypes as C
    oracle.set_maxp(6)
    rng = np.random.default_rng(31 + kind)
    p, nel = 2, 2
    norder = np.tile(uniform_order(p), (nel, 1))
    norie = rng.integers(0, 2, (nel, 12)).astype(np.int32); norif = rng.integers(0, 8, (nel, 6)).astype(np.int32)
    nH = oracle.celndof(norder[0])[0]
    X = np.stack([hexa_xnod(nH, h=0.5, jitter=0.12, rng=rng) for e in range(nel)])
    om = 2 * np.pi if kind == 4 else np.pi
    eng = _engine(kind, omega=om, source=9, real_reduction=rr)
    nint = eng.sizes(norder[0])[2]
    J = rng.standard_normal((nel, nint, 3)) + 1j * rng.standard_normal((nel, nint, 3))
    res = eng.elem_stc_batch(norder, norie, norif, X, source_qp=J)
    assert (res["info"] == 0).all()
    for e in range(nel):
        tab = np.ascontiguousarray(J[e])
        prm = _oracle_params(oracle, omega=om, source=9, source_table=tab.ctypes.data_as(C.c_void_p))
        rA, rB, rAS, rBS = oracle.condensed(kind, norder[e], norie[e], norif[e], X[e], prm)
        Aii, Bi, AS, BS = eng.unpack(res, e)
        assert relerr(Aii, rA) < 1e-12 and relerr(Bi, rB) < 1e-12, (e, relerr(Bi, rB))
        if BS.size:
            assert relerr(BS, rBS) < 1e-10
    eng.close()


@pytest.mark.parametrize("test_norm", [1, 2, 3])
@pytest.mark.parametrize("rr", [1, 0])
def test_uw_maxwell_material_and_norm_parameters(oracle, gpu, test_norm, rr):
    """Ultraweak Maxwell away from the defaults: eps, mu != 1, ALPHA_NORM != 1 and the three test norms (adjoint graph,
    mathematician's, diagonal graph), second manufactured component -- through the real form and the complex kernels."""
    oracle.set_maxp(6)
    oracle.use_blas(True)
    rng = np.random.default_rng(555 + test_norm)
    p, nel = 2, 2
    norder = np.tile(uniform_order(p), (nel, 1))
    norie = rng.integers(0, 2, (nel, 12)).astype(np.int32); norif = rng.integers(0, 8, (nel, 6)).astype(np.int32)
    nH = oracle.celndof(norder[0])[0]
    X = np.stack([hexa_xnod(nH, h=0.5, jitter=0.12, curved=0.01, rng=rng) for e in range(nel)])
    kw = dict(omega=1.7 * np.pi, eps=2.5, mu=0.7, alpha_norm=0.3, test_norm=test_norm, icomp_exact=3)
    prm = _oracle_params(oracle, **kw)
    eng = _engine(4, real_reduction=rr, **kw)
    res = eng.elem_stc_batch(norder, norie, norif, X)
    assert (res["info"] == 0).all()
    for e in range(nel):
        Aii, Bi, AS, BS = eng.unpack(res, e)
        rA, rB, rAS, rBS = oracle.condensed(4, norder[e], norie[e], norif[e], X[e], prm)
        assert relerr(Aii, rA) < 1e-12 and relerr(Bi, rB) < 1e-12, (e, relerr(Aii, rA), relerr(Bi, rB))
        assert relerr(AS, rAS) < 1e-9 and relerr(BS, rBS) < 1e-9
    eng.close()


@pytest.mark.parametrize("kind", [2, 4])
def test_without_schur_factors(oracle, gpu, kind):
    """STORE_STC off (hp3d_params.store_schur = 0): the condensed system is unchanged, the back-substitution factors are
    neither formed (the Z = Y L^-1 steps are skipped) nor copied."""
    oracle.set_maxp(6)
    rng = np.random.default_rng(99 + kind)
    p, nel = 3, 3
    norder = np.tile(uniform_order(p), (nel, 1))
    norie = rng.integers(0, 2, (nel, 12)).astype(np.int32); norif = rng.integers(0, 8, (nel, 6)).astype(np.int32)
    nH = oracle.celndof(norder[0])[0]
    X = np.stack([hexa_xnod(nH, h=0.5, jitter=0.12, rng=rng) for e in range(nel)])
    om = 2 * np.pi if kind == 4 else 1.0
    full = _engine(kind, omega=om)
    ref = full.elem_stc_batch(norder, norie, norif, X)
    eng = _engine(kind, omega=om, store_schur=0)
    res = eng.elem_stc_batch(norder, norie, norif, X)
    assert (res["info"] == 0).all()
    for e in range(nel):
        a, b = eng.unpack(res, e), full.unpack(ref, e)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])     # same kernels, same order: bit-identical
        assert not a[2].any() and not a[3].any()                             # the Schur arrays are left untouched
    full.close(); eng.close()


@pytest.mark.parametrize("test_norm", [1, 2, 3])
@pytest.mark.parametrize("tensor,rr", [("real", 1), ("real", 0), ("complex", 1), ("complex", 0)])
def test_uw_maxwell_permittivity_tensor(oracle, gpu, test_norm, tensor, rr):
    """get_permittivity other than the identity (elem_opt.F90:260-266: za = i w eps * eps_t, a full 3x3 tensor): the Gram term
    (za^H F, za^H F), the cross term (curl G, za^H F) and the stiffness -(za E, F) against the oracle's BLAS3 elem_opt with the same
    tensor.  A real tensor keeps the real form (rr = 1), a complex one takes the general complex kernels whatever rr says."""
    import ctypes as C
    oracle.set_maxp(6)
    oracle.use_blas(True)
    rng = np.random.default_rng(777 + test_norm)
    p, nel = 2, 2
    norder = np.tile(uniform_order(p), (nel, 1))
    norie = rng.integers(0, 2, (nel, 12)).astype(np.int32); norif = rng.integers(0, 8, (nel, 6)).astype(np.int32)
    nH = oracle.celndof(norder[0])[0]
    X = np.stack([hexa_xnod(nH, h=0.5, jitter=0.12, curved=0.01, rng=rng) for e in range(nel)])
    T = np.eye(3) + 0.3 * rng.standard_normal((3, 3))
    if tensor == "complex":
        T = T + 0.2j * rng.standard_normal((3, 3))
    kw = dict(omega=1.3 * np.pi, eps=1.5, mu=0.8, alpha_norm=0.6, test_norm=test_norm, eps_tensor=T)
    eng = _engine(4, real_reduction=rr, source=9, **kw)
    nint = eng.sizes(norder[0])[2]
    J = rng.standard_normal((nel, nint, 3)) + 1j * rng.standard_normal((nel, nint, 3))
    res = eng.elem_stc_batch(norder, norie, norif, X, source_qp=J)
    assert (res["info"] == 0).all()
    for e in range(nel):
        tab = np.ascontiguousarray(J[e])
        prm = _oracle_params(oracle, source=9, source_table=tab.ctypes.data_as(C.c_void_p), **kw)
        Aii, Bi, AS, BS = eng.unpack(res, e)
        rA, rB, rAS, rBS = oracle.condensed(4, norder[e], norie[e], norif[e], X[e], prm)
        assert relerr(Aii, rA) < 1e-12 and relerr(Bi, rB) < 1e-12, (e, relerr(Aii, rA), relerr(Bi, rB))
        assert relerr(AS, rAS) < 1e-9 and relerr(BS, rBS) < 1e-9
    eng.close()
    with pytest.raises(RuntimeError, match="HP3D_SRC_TABLE"):
        _engine(4, eps_tensor=T)                      # the built-in manufactured source assumes the identity tensor
    with pytest.raises(RuntimeError, match="ultraweak Maxwell only"):
        _engine(3, eps_tensor=T, source=9)
