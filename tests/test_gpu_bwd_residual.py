"""SURVEY.md 8f rows built on the same path: back-substitution without stored Schur factors (f3: stc_bwd_wrapper's
"recompute" option, stc.F90:279-281,529-677) and the DPG element residual (f2: elem_residual_maxwell.F90:246-552,
POISSON/PRIMAL_DPG/elem_residual.F90), both against the oracle."""
import numpy as np
import pytest

from tests.test_gpu_prism import prism_xnod
from tests.test_oracle_prism import prism_signature
from tests.util import hexa_xnod, random_signature

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _batch(oracle, rng, nel=4, pmax=3):
    B, P = oracle.MDLB, oracle.MDLP
    items = []
    for e in range(nel):
        if e % 2 == 0:
            no, ne, nf = random_signature(rng, pmax=pmax)
            items.append((B, no, ne, nf, hexa_xnod(oracle.celndof(no, B)[0], h=0.4, jitter=0.1, rng=rng)))
        else:
            no, ne, nf = prism_signature(rng, int(rng.integers(1, pmax + 1)), int(rng.integers(1, pmax + 1)), uniform=False)
            items.append((P, no, ne, nf, prism_xnod(oracle.celndof(no, P)[0], rng)))
    et = np.array([it[0] for it in items], np.int32)
    norder = np.stack([it[1] for it in items]); norie = np.stack([it[2] for it in items]); norif = np.stack([it[3] for it in items])
    X = np.zeros((nel, max(it[4].shape[0] for it in items), 3))
    for e, it in enumerate(items):
        X[e, :it[4].shape[0]] = it[4]
    return items, et, norder, norie, norif, X


@pytest.mark.parametrize("kind", [1, 2, 3, 4])
def test_bwd_recompute(oracle, gpu, kind):
    from hp3d_b200.api import ElemEngine
    oracle.set_maxp(8)
    rng = np.random.default_rng(40 + kind)
    items, et, norder, norie, norif, X = _batch(oracle, rng)
    om = 2 * np.pi if kind == 4 else 1.0
    eng = ElemEngine(kind, omega=om, maxp=8)
    res = eng.elem_stc_batch(norder, norie, norif, X, etype=et)
    nel = len(items)
    ni_max = int(res["ni"].max())
    xi = rng.normal(size=(nel, ni_max)) + (1j * rng.normal(size=(nel, ni_max)) if kind >= 3 else 0)
    out = eng.elem_bwd_batch(norder, norie, norif, X, xi, etype=et)
    assert (out["info"] == 0).all()
    prm = oracle.default_params(omega=om)
    for e, it in enumerate(items):
        ni, nb = int(res["ni"][e]), int(res["nb"][e])
        assert out["nb"][e] == nb
        if nb == 0:
            continue
        _, _, AS, BS = eng.unpack(res, e)
        ref = BS - AS @ xi[e, :ni]
        assert relerr(out["xb"][e, :nb], ref) < 1e-13
        _, _, rAS, rBS = oracle.condensed(kind, it[1], it[2], it[3], it[4], prm, etype=it[0])
        cond = np.linalg.cond(rAS) if rAS.size else 1.0
        assert relerr(out["xb"][e, :nb], rBS - rAS @ xi[e, :ni]) < 1e-9
    eng.close()


@pytest.mark.parametrize("kind", [2, 4])
def test_dpg_residual(oracle, gpu, kind):
    """eta^2 = (G^-1 (l - B u), l - B u) from the oracle's Gram / enriched stiffness (ultraweak Maxwell) or from its condensed
    outputs' uncondensed twin (primal Poisson) vs hp3d_gpu_elem_residual_batch, random u."""
    from hp3d_b200.api import ElemEngine
    oracle.set_maxp(8)
    oracle.use_blas(True)
    rng = np.random.default_rng(60 + kind)
    items, et, norder, norie, norif, X = _batch(oracle, rng)
    om = 2 * np.pi if kind == 4 else 1.0
    prm = oracle.default_params(omega=om)
    eng = ElemEngine(kind, omega=om, maxp=8)
    nel = len(items)
    parts = [oracle.stc_partition(kind, it[1], it[0]) for it in items]
    ni_max = max(p[1] for p in parts); nb_max = max(max(p[2] for p in parts), 1)
    cplx = kind >= 3
    xi = np.zeros((nel, ni_max), complex if cplx else float); xb = np.zeros((nel, nb_max), complex if cplx else float)
    ref = np.zeros(nel)
    for e, it in enumerate(items):
        perm, ni, nb = parts[e]
        A, b = oracle.elem(kind, it[1], it[2], it[3], it[4], prm, etype=it[0])
        n = ni + nb
        u = rng.normal(size=n) + (1j * rng.normal(size=n) if cplx else 0)
        xi[e, :ni] = u[perm[:ni]]; xb[e, :nb] = u[perm[ni:]]
        if kind == 4:
            _, _, G, S = oracle.elem(kind, it[1], it[2], it[3], it[4], prm, want_dpg=True, etype=it[0])
            Gu = np.triu(G); Gf = Gu + np.triu(Gu, 1).conj().T
            r = S[:, -1] - S[:, :-1] @ u
            ref[e] = np.real(np.vdot(r, np.linalg.solve(Gf, r)))
        else:
            # A = B^T G^-1 B and b = B^T G^-1 l are known; c = l^T G^-1 l follows from the zero-solution residual of the GPU itself
            ref[e] = np.nan
        if kind == 2:
            ref[e] = np.real(np.vdot(u, A @ u) - 2 * np.real(np.vdot(u, b)))   # eta^2 - c
    out = eng.elem_residual_batch(norder, norie, norif, X, xi, xb, etype=et)
    assert (out["info"] == 0).all()
    if kind == 2:
        c = eng.elem_residual_batch(norder, norie, norif, X, np.zeros_like(xi), np.zeros_like(xb), etype=et)["resid"]
        assert (c > 0).all()
        assert np.abs(out["resid"] - c - ref).max() < 1e-11 * max(np.abs(ref).max(), c.max())
    else:
        assert np.abs(out["resid"] - ref).max() < 1e-10 * np.abs(ref).max(), (out["resid"], ref)
    assert (out["resid"] >= 0).all()
    eng.close()
