"""GPU parity for TRIANGULAR PRISMS (BASELINE.json configs[4]: hp meshes of hexahedra and prisms): the CUDA path through the
C ABI against the CPU oracle's prism branch on the same seeded inputs.  Tolerances as in test_gpu_parity.py."""
import numpy as np
import pytest

from tests.test_oracle_prism import PV, prism_signature
from tests.util import complex_W

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def prism_xnod(nH, rng, h=0.5, curved=0.0):
    """Geometry dofs of a prism: an affine image of the master prism, jittered vertices, optional higher-order dofs."""
    T = np.eye(3) + rng.normal(size=(3, 3)) * 0.15
    X = np.zeros((nH, 3))
    X[:6] = h * (PV @ T.T) + rng.uniform(0, 0.3, 3) + rng.uniform(-0.05 * h, 0.05 * h, (6, 3))
    if curved:
        X[6:] = rng.uniform(-curved * h, curved * h, (nH - 6, 3))
    return X


def _engine(kind, **kw):
    from hp3d_b200.api import ElemEngine
    return ElemEngine(kind, **kw)


@pytest.mark.parametrize("rr", [1, 0])
@pytest.mark.parametrize("p,pz,curved", [(1, 1, 0.0), (2, 2, 0.0), (2, 3, 0.02), (3, 2, 0.02)])
def test_prism_uw_integration_vs_oracle(oracle, gpu, p, pz, curved, rr):
    """Gram matrix, enriched stiffness, trace pairings and load of ultraweak Maxwell on a prism, straight out of the
    integration kernels, vs the oracle's BLAS3 restatement of elem_opt.F90:236-768."""
    oracle.set_maxp(8)
    P = oracle.MDLP
    rng = np.random.default_rng(300 + 10 * p + pz)
    no, ne, nf = prism_signature(rng, p, pz)
    nH = oracle.celndof(no, P)[0]
    X = prism_xnod(nH, rng, curved=curved)
    om = 2 * np.pi
    prm = oracle.default_params(omega=om)
    A, b, G, S = oracle.elem(oracle.MAXW_UW, no, ne, nf, X, prm, want_dpg=True, etype=P)
    eng = _engine(4, omega=om, maxp=8, real_reduction=rr)
    W, d = eng.integrate_debug(no, ne, nf, X, etype=P)
    assert W.shape[0] == (1 if rr else 2)
    n, nb, ni, np_, nbp = d["n"], d["nb"], d["ni"], d["np"], d["nbp"]
    nEE = n // 2
    assert G.shape[0] == n and S.shape[1] == ni + nb + 1
    Wc = complex_W(W, d)
    Gi = np.tril(Wc[:n, :n]); Gi = Gi + np.tril(Gi, -1).conj().T
    perm = np.empty(n, int); perm[0::2] = np.arange(nEE); perm[1::2] = nEE + np.arange(nEE)
    Gg = Gi[np.ix_(perm, perm)]
    Gu = np.triu(G); Go = Gu + np.triu(Gu, 1).conj().T
    assert relerr(Gg, Go) < 1e-13, relerr(Gg, Go)
    rows = np.r_[np_ + nbp + np.arange(ni), np_ + np.arange(nb), np_ + nbp + d["nip"] - 1]   # load: last padded interface row
    Bg = Wc[rows][:, :n].conj().T[perm]
    assert relerr(Bg[:, :ni], S[:, :ni]) < 1e-13, relerr(Bg[:, :ni], S[:, :ni])          # trace pairings
    assert relerr(Bg[:, ni:ni + nb], S[:, ni:ni + nb]) < 1e-13
    assert relerr(Bg[:, -1], S[:, -1]) < 1e-13                                             # load
    eng.close()


CASES = [(1, 1, 1), (1, 2, 2), (1, 3, 2), (1, 4, 3),
         (2, 1, 1), (2, 2, 2), (2, 3, 2),
         (3, 1, 1), (3, 2, 2), (3, 3, 3), (3, 4, 2),
         (4, 1, 1), (4, 2, 2), (4, 3, 2), (4, 2, 3)]


@pytest.mark.parametrize("kind,p,pz", CASES)
def test_prism_condensed_vs_oracle(oracle, gpu, kind, p, pz):
    """elem + stc_fwd_wrapper for prisms through hp3d_gpu_elem_batch vs the oracle: random edge / face orientations,
    non-uniform node orders for the second element, jittered and slightly curved geometry."""
    oracle.set_maxp(8)
    oracle.use_blas(True)
    P = oracle.MDLP
    rng = np.random.default_rng(5000 + 100 * kind + 10 * p + pz)
    nel = 3
    sig = [prism_signature(rng, p, pz, uniform=(e != 1)) for e in range(nel)]
    norder = np.stack([s[0] for s in sig]); norie = np.stack([s[1] for s in sig]); norif = np.stack([s[2] for s in sig])
    nHs = [oracle.celndof(norder[e], P)[0] for e in range(nel)]
    X = np.zeros((nel, max(nHs), 3))
    for e in range(nel):
        X[e, :nHs[e]] = prism_xnod(nHs[e], rng, curved=0.01 if p > 1 else 0.0)
    om = 2 * np.pi if kind == 4 else (np.pi if kind == 3 else 1.0)
    prm = oracle.default_params(omega=om)
    eng = _engine(kind, omega=om, maxp=8)
    res = eng.elem_stc_batch(norder, norie, norif, X, etype=P)
    assert (res["info"] == 0).all(), res["info"]
    for e in range(nel):
        Aii, Bi, AS, BS = eng.unpack(res, e)
        rA, rB, rAS, rBS = oracle.condensed(kind, norder[e], norie[e], norif[e], X[e, :nHs[e]], prm, etype=P)
        assert Aii.shape == rA.shape and AS.shape == rAS.shape
        assert relerr(Aii, rA) < 1e-12, (e, relerr(Aii, rA))
        assert relerr(Bi, rB) < 1e-12, (e, relerr(Bi, rB))
        if AS.size:
            Afull, bfull = oracle.elem(kind, norder[e], norie[e], norif[e], X[e, :nHs[e]], prm, etype=P)
            perm, ni, nb = oracle.stc_partition(kind, norder[e], P)
            Ap = Afull[np.ix_(perm, perm)]; bp = bfull[perm]
            Abb, Abi = Ap[ni:, ni:], Ap[ni:, :ni]
            assert relerr(Abb @ AS, Abi) < 1e-12
            assert relerr(Abb @ BS, bp[ni:]) < 1e-12
    eng.close()


def test_mixed_hexa_prism_batch(oracle, gpu):
    """One call holding bricks and prisms of different orders (the hp-mesh situation of configs[4])."""
    from tests.util import hexa_xnod, random_signature
    oracle.set_maxp(8)
    rng = np.random.default_rng(77)
    B, P = oracle.MDLB, oracle.MDLP
    items = []
    for e in range(6):
        if e % 2 == 0:
            no, ne, nf = random_signature(rng, pmax=3)
            nH = oracle.celndof(no, B)[0]
            items.append((B, no, ne, nf, hexa_xnod(nH, h=0.4, jitter=0.1, rng=rng)))
        else:
            no, ne, nf = prism_signature(rng, int(rng.integers(1, 4)), int(rng.integers(1, 4)), uniform=False)
            nH = oracle.celndof(no, P)[0]
            items.append((P, no, ne, nf, prism_xnod(nH, rng)))
    nel = len(items)
    et = np.array([it[0] for it in items], np.int32)
    norder = np.stack([it[1] for it in items]); norie = np.stack([it[2] for it in items]); norif = np.stack([it[3] for it in items])
    X = np.zeros((nel, max(it[4].shape[0] for it in items), 3))
    for e, it in enumerate(items):
        X[e, :it[4].shape[0]] = it[4]
    for kind in (1, 4):
        om = 2 * np.pi if kind == 4 else 1.0
        prm = oracle.default_params(omega=om)
        eng = _engine(kind, omega=om, maxp=8)
        res = eng.elem_stc_batch(norder, norie, norif, X, etype=et)
        assert (res["info"] == 0).all()
        for e, it in enumerate(items):
            Aii, Bi, AS, BS = eng.unpack(res, e)
            rA, rB, _, _ = oracle.condensed(kind, it[1], it[2], it[3], it[4], prm, etype=it[0])
            assert Aii.shape == rA.shape
            assert relerr(Aii, rA) < 1e-12, (kind, e, relerr(Aii, rA))
            assert relerr(Bi, rB) < 1e-12, (kind, e)
        eng.close()


def test_hp_mesh_batch_vs_oracle(oracle, gpu):
    """BASELINE.json configs[4] in miniature: a conforming hp mesh of hexahedra and prisms (min-rule orders 2..4, orientations
    from a random global vertex numbering, ~30 distinct signatures) through ONE hp3d_gpu_elem_batch call, every element
    against the oracle; ultraweak Maxwell."""
    from hp3d_b200 import synth
    oracle.set_maxp(8)
    oracle.use_blas(True)
    m = synth.hp_mesh(2, prism_frac=0.5, pmin=2, pmax=4, seed_p=3, seed_g=9, jitter=0.1)
    nel = len(m["etype"])
    assert (m["etype"] == 3).any() and (m["etype"] == 1).any()
    om = 2 * np.pi
    prm = oracle.default_params(omega=om)
    eng = _engine(4, omega=om, maxp=8)
    res = eng.elem_stc_batch(m["norder"], m["norient_edge"], m["norient_face"], m["xnod"], etype=m["etype"])
    assert (res["info"] == 0).all()
    for e in range(nel):
        nH = int(m["nrdofH"][e])
        Aii, Bi, AS, BS = eng.unpack(res, e)
        rA, rB, _, _ = oracle.condensed(4, m["norder"][e], m["norient_edge"][e], m["norient_face"][e], m["xnod"][e, :nH], prm,
                                        etype=int(m["etype"][e]))
        assert Aii.shape == rA.shape
        assert relerr(Aii, rA) < 1e-12, (e, relerr(Aii, rA))
        assert relerr(Bi, rB) < 1e-12, (e, relerr(Bi, rB))
    eng.close()


def test_high_order_elements(oracle, gpu):
    """The upper end of configs[4] (orders up to 7, enriched to 8): one p=6 brick and one p=(6,6) prism against the oracle,
    and p=7 elements through size-independent properties (Hermitian, positive semi-definite condensed matrix, info = 0)."""
    from tests.util import hexa_xnod, uniform_order
    oracle.set_maxp(8)
    oracle.use_blas(True, threads=8)
    rng = np.random.default_rng(99)
    om = 2 * np.pi
    prm = oracle.default_params(omega=om)
    eng = _engine(4, omega=om, maxp=8)
    B, P = oracle.MDLB, oracle.MDLP
    for p, check in ((6, True), (7, False)):
        nob = uniform_order(p); nop = oracle.uniform_order(p, P, p)
        neb = rng.integers(0, 2, 12).astype(np.int32); nfb = rng.integers(0, 8, 6).astype(np.int32)
        _, nep, nfp = prism_signature(rng, p, p)
        nHb, nHp = oracle.celndof(nob, B)[0], oracle.celndof(nop, P)[0]
        X = np.zeros((2, max(nHb, nHp), 3))
        X[0, :nHb] = hexa_xnod(nHb, h=0.3, jitter=0.1, rng=rng)
        X[1, :nHp] = prism_xnod(nHp, rng, h=0.3)
        res = eng.elem_stc_batch(np.stack([nob, nop]), np.stack([neb, nep]), np.stack([nfb, nfp]), X, etype=np.array([B, P], np.int32))
        assert (res["info"] == 0).all(), res["info"]
        for e, (et, no, ne, nf, nH) in enumerate(((B, nob, neb, nfb, nHb), (P, nop, nep, nfp, nHp))):
            Aii, Bi, AS, BS = eng.unpack(res, e)
            assert relerr(Aii, Aii.conj().T) < 1e-13
            w = np.linalg.eigvalsh(Aii)
            assert w.min() > -1e-9 * w.max()
            if check:
                rA, rB, _, _ = oracle.condensed(4, no, ne, nf, X[e, :nH], prm, etype=et)
                # the p=6 prism Gram matrix has cond(G) = 2.5e10 (hexa: 4e9): two correct FP64 evaluations of B^H G^-1 B agree
                # to ~cond*eps*1e-4; the bar stays 1e-12 for the brick and is 5e-12 for this prism
                tol = 1e-12 if et == B else 5e-12
                assert relerr(Aii, rA) < tol, (p, e, relerr(Aii, rA))
                assert relerr(Bi, rB) < tol, (p, e, relerr(Bi, rB))
    oracle.use_blas(True, threads=1)
    eng.close()


def test_edge_cases(gpu):
    """Empty batch, unknown element type, invalid descriptors (loud errors, no silent fallback), negative Jacobian on a prism."""
    from hp3d_b200.api import ElemEngine
    eng = _engine(4, omega=1.0)
    z = lambda *s: np.zeros(s, np.int32)   # noqa: E731
    res = eng.elem_stc_batch(z(0, 19), z(0, 12), z(0, 6), np.zeros((0, 8, 3)))
    assert res["info"].size == 0
    P = 3
    rng = np.random.default_rng(1)
    no, ne, nf = prism_signature(rng, 2, 2)
    X = prism_xnod(18, rng)[None]
    with pytest.raises(RuntimeError, match="element type"):
        eng.elem_stc_batch(no[None], ne[None], nf[None], X, etype=2)          # MDLN (tetrahedron): not implemented
    bad = no.copy(); bad[0] = 0
    with pytest.raises(RuntimeError, match="order"):
        eng.elem_stc_batch(bad[None], ne[None], nf[None], X, etype=P)
    badf = nf.copy(); badf[0] = 6
    with pytest.raises(RuntimeError, match="orientation"):
        eng.elem_stc_batch(no[None], ne[None], badf[None], X, etype=P)
    Xm = X.copy(); Xm[0, :6, 2] *= -1.0                                        # mirrored prism: negative Jacobian
    res = eng.elem_stc_batch(no[None], ne[None], nf[None], Xm, etype=P)
    assert res["info"][0] == -1
    with pytest.raises(RuntimeError, match="xnod_ld"):
        eng.elem_stc_batch(no[None], ne[None], nf[None], X[:, :10], etype=P)   # too few geometry dofs
    with pytest.raises(RuntimeError, match="DPG"):
        ElemEngine(1).elem_residual_batch(no[None], ne[None], nf[None], X, np.zeros((1, 4)), np.zeros((1, 4)), etype=P)
    eng.close()


@pytest.mark.parametrize("test_norm", [1, 2, 3])
@pytest.mark.parametrize("tensor,rr", [("real", 1), ("complex", 1), ("complex", 0)])
def test_prism_uw_maxwell_permittivity_tensor(oracle, gpu, test_norm, tensor, rr):
    """the permittivity tensor of tests/test_gpu_parity.py::test_uw_maxwell_permittivity_tensor on PRISMS (mixed with a brick in the
    same call): Gram, cross and stiffness terms through the (triangle function) x (z table) families against the oracle"""
    import ctypes as C
    from hp3d_b200.api import ElemEngine
    from tests.test_oracle_prism import prism_signature
    from tests.util import hexa_xnod, random_signature
    oracle.set_maxp(6)
    oracle.use_blas(True)
    rng = np.random.default_rng(888 + test_norm)
    B, P = oracle.MDLB, oracle.MDLP
    items = []
    for e in range(3):
        if e == 1:
            no, ne, nf = random_signature(rng, pmax=2)
            items.append((B, no, ne, nf, hexa_xnod(oracle.celndof(no, B)[0], h=0.4, jitter=0.1, rng=rng)))
        else:
            no, ne, nf = prism_signature(rng, 2, int(rng.integers(1, 3)), uniform=False)
            items.append((P, no, ne, nf, prism_xnod(oracle.celndof(no, P)[0], rng)))
    nel = len(items)
    et = np.array([it[0] for it in items], np.int32)
    norder = np.stack([it[1] for it in items]); norie = np.stack([it[2] for it in items]); norif = np.stack([it[3] for it in items])
    X = np.zeros((nel, max(it[4].shape[0] for it in items), 3))
    for e, it in enumerate(items):
        X[e, :it[4].shape[0]] = it[4]
    T = np.eye(3) + 0.3 * rng.standard_normal((3, 3))
    if tensor == "complex":
        T = T + 0.2j * rng.standard_normal((3, 3))
    kw = dict(omega=1.3 * np.pi, eps=1.5, mu=0.8, alpha_norm=0.6, test_norm=test_norm, eps_tensor=T)
    eng = ElemEngine(4, real_reduction=rr, source=9, maxp=6, **kw)
    nint = max(eng.sizes(norder[e], int(et[e]))[2] for e in range(nel))
    J = rng.standard_normal((nel, nint, 3)) + 1j * rng.standard_normal((nel, nint, 3))
    res = eng.elem_stc_batch(norder, norie, norif, X, source_qp=J, etype=et)
    assert (res["info"] == 0).all()
    for e, it in enumerate(items):
        ni_e = eng.sizes(norder[e], int(et[e]))[2]
        tab = np.ascontiguousarray(J[e, :ni_e])
        prm = oracle.default_params(source=9, source_table=tab.ctypes.data_as(C.c_void_p), **kw)
        Aii, Bi, AS, BS = eng.unpack(res, e)
        rA, rB, rAS, rBS = oracle.condensed(4, it[1], it[2], it[3], it[4], prm, etype=it[0])
        assert relerr(Aii, rA) < 1e-12 and relerr(Bi, rB) < 1e-12, (e, it[0], relerr(Aii, rA), relerr(Bi, rB))
        assert relerr(AS, rAS) < 1e-9 and relerr(BS, rBS) < 1e-9
    eng.close()
