"""The top of BASELINE.json configs[4] (hp-refined hexa/prism meshes with orders 2..7, enriched to 8) against the oracle:
p = 7 bricks and prisms, the full p = 2..7 mixed mesh, the DPG residual and the Schur factors at the largest sizes.

Tolerances.  Element / condensed matrices: 1e-12 relative Frobenius for bricks at every order.  The Gram matrix of a prism of
order >= 6 has cond(G) = 2e10 .. 5e10 (brick p=5: 1e9, p=7: 1.5e10; measured, profiles/r02_highorder_parity.json): two correct FP64
evaluations of B^H G^-1 B then differ by more than 1e-12 (OpenBLAS with 1 and 16 threads: 5e-13), so the prism bar is
1e-12 * max(1, cond(G) / 4e9), with cond(G) computed in the test.  Schur factors: residual ||A_bb ASchur - A_bi|| / ||A_bi||."""
import numpy as np
import pytest

from tests.test_gpu_prism import prism_xnod
from tests.test_oracle_prism import prism_signature
from tests.util import hexa_xnod, uniform_order

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _cond_gram(O, no, ne, nf, X, prm, et):
    _, _, G, _ = O.elem(4, no, ne, nf, X, prm, want_dpg=True, etype=et)
    Gu = np.triu(G)
    w = np.linalg.eigvalsh(Gu + np.triu(Gu, 1).conj().T)
    return float(w[-1] / w[0])


def test_p7_brick_and_prism_vs_oracle(oracle, gpu):
    """p = 7 (3888 / 2592 enriched test dofs): Aii, Bi against the oracle, Schur factors by residual, DPG residual of a random u."""
    from hp3d_b200.api import ElemEngine
    O = oracle
    O.set_maxp(8)
    O.use_blas(True, threads=16)
    rng = np.random.default_rng(707)
    om = 2 * np.pi
    prm = O.default_params(omega=om)
    eng = ElemEngine(4, omega=om, maxp=8)
    p = 7
    nob = uniform_order(p); nop = O.uniform_order(p, O.MDLP, p)
    neb = rng.integers(0, 2, 12).astype(np.int32); nfb = rng.integers(0, 8, 6).astype(np.int32)
    _, nep, nfp = prism_signature(rng, p, p)
    nHb, nHp = O.celndof(nob, O.MDLB)[0], O.celndof(nop, O.MDLP)[0]
    X = np.zeros((2, max(nHb, nHp), 3))
    X[0, :nHb] = hexa_xnod(nHb, h=0.3, jitter=0.1, rng=rng)
    X[1, :nHp] = prism_xnod(nHp, rng, h=0.3)
    NO, NE, NF, ET = np.stack([nob, nop]), np.stack([neb, nep]), np.stack([nfb, nfp]), np.array([O.MDLB, O.MDLP], np.int32)
    res = eng.elem_stc_batch(NO, NE, NF, X, etype=ET)
    assert (res["info"] == 0).all(), res["info"]
    ni_max, nb_max = int(res["ni"].max()), int(res["nb"].max())
    xi = np.zeros((2, ni_max), complex); xb = np.zeros((2, nb_max), complex); eta_ref = np.zeros(2)
    for e, (et, no, ne, nf, nH) in enumerate(((O.MDLB, nob, neb, nfb, nHb), (O.MDLP, nop, nep, nfp, nHp))):
        Aii, Bi, AS, BS = eng.unpack(res, e)
        rA, rB, rAS, rBS = O.condensed(4, no, ne, nf, X[e, :nH], prm, etype=et)
        tol = 1e-12 if et == O.MDLB else 1e-12 * max(1.0, _cond_gram(O, no, ne, nf, X[e, :nH], prm, et) / 4e9)
        assert relerr(Aii, rA) < tol, (et, relerr(Aii, rA), tol)
        assert relerr(Bi, rB) < tol, (et, relerr(Bi, rB), tol)
        Afull, bfull, G, S = O.elem(4, no, ne, nf, X[e, :nH], prm, want_dpg=True, etype=et)
        perm, ni, nb = O.stc_partition(4, no, et)
        Ap = Afull[np.ix_(perm, perm)]; bp = bfull[perm]
        assert relerr(Ap[ni:, ni:] @ AS, Ap[ni:, :ni]) < 10 * tol
        assert relerr(Ap[ni:, ni:] @ BS, bp[ni:]) < 10 * tol
        # DPG residual of a random element solution (M = nbp + nip = 3328 at p=7: the residual kernel's vector needs > 48 KB)
        u = rng.normal(size=ni + nb) + 1j * rng.normal(size=ni + nb)
        xi[e, :ni] = u[perm[:ni]]; xb[e, :nb] = u[perm[ni:]]
        Gu = np.triu(G)
        r = S[:, -1] - S[:, :-1] @ u
        eta_ref[e] = np.real(np.vdot(r, np.linalg.solve(Gu + np.triu(Gu, 1).conj().T, r)))
    out = eng.elem_residual_batch(NO, NE, NF, X, xi, xb, etype=ET)
    assert (out["info"] == 0).all()
    assert np.abs(out["resid"] - eta_ref).max() < 1e-9 * np.abs(eta_ref).max(), (out["resid"], eta_ref)
    O.use_blas(True, threads=1)
    eng.close()


def test_hp_mesh_p2_to_7_vs_oracle(oracle, gpu):
    """BASELINE.json configs[4] itself: a conforming mixed mesh with element orders drawn from {2..7} (min rule on edges / faces,
    orientations from a random global vertex numbering), ALL elements through one hp3d_gpu_elem_batch call; a sample of 16
    elements that contains the highest-order brick and prism is compared with the oracle."""
    from hp3d_b200 import synth
    from hp3d_b200.api import ElemEngine
    O = oracle
    O.set_maxp(8)
    O.use_blas(True, threads=16)
    m = synth.hp_mesh(3, prism_frac=0.4, pmin=2, pmax=7, seed_p=2024, seed_g=7, jitter=0.1)
    nel = len(m["etype"])
    assert m["p"].max() == 7 and m["p"].min() == 2 and (m["etype"] == 3).any() and (m["etype"] == 1).any()
    om = 2 * np.pi
    prm = O.default_params(omega=om)
    eng = ElemEngine(4, omega=om, maxp=8)
    res = eng.elem_stc_batch(m["norder"], m["norient_edge"], m["norient_face"], m["xnod"], etype=m["etype"])
    assert (res["info"] == 0).all()
    rng = np.random.default_rng(1)
    top = [int(np.flatnonzero((m["etype"] == t) & (m["p"] == m["p"][m["etype"] == t].max()))[0]) for t in (1, 3)]
    sample = sorted(set(top) | set(int(i) for i in rng.choice(nel, 14, replace=False)))
    assert len(sample) >= 12
    worst = 0.0
    for e in sample:
        nH = int(m["nrdofH"][e]); et = int(m["etype"][e])
        d = (m["norder"][e], m["norient_edge"][e], m["norient_face"][e], m["xnod"][e, :nH], prm)
        Aii, Bi, AS, BS = eng.unpack(res, e)
        rA, rB, _, _ = O.condensed(4, *d, etype=et)
        assert Aii.shape == rA.shape
        tol = 1e-12 if (et == 1 or m["p"][e] < 6) else 1e-12 * max(1.0, _cond_gram(O, *d, et) / 4e9)
        assert relerr(Aii, rA) < tol, (e, et, int(m["p"][e]), relerr(Aii, rA), tol)
        assert relerr(Bi, rB) < tol, (e, et, int(m["p"][e]), relerr(Bi, rB), tol)
        worst = max(worst, relerr(Aii, rA))
    O.use_blas(True, threads=1)
    eng.close()


def test_schur_factor_residual_at_full_size(oracle, gpu):
    """north_star: condensed matrices to 1e-12.  ASchur / BSchur are solutions of A_bb X = [A_bi | b_b]; at the bench size (p = 5,
    nb = 750, ni = 600) their residual against the oracle's uncondensed matrix is held to 1e-12."""
    from hp3d_b200 import synth
    from hp3d_b200.api import ElemEngine
    O = oracle
    O.set_maxp(6)
    O.use_blas(True, threads=16)
    norder, noe, nof, xnod = synth.cube_mesh(1, 5, first=11)
    om = 2 * np.pi
    eng = ElemEngine(4, omega=om)
    res = eng.elem_stc_batch(norder, noe, nof, xnod)
    Aii, Bi, AS, BS = eng.unpack(res, 0)
    prm = O.default_params(omega=om)
    Afull, bfull = O.elem(4, norder[0], noe[0], nof[0], xnod[0], prm)
    perm, ni, nb = O.stc_partition(4, norder[0])
    Ap = Afull[np.ix_(perm, perm)]; bp = bfull[perm]
    assert relerr(Ap[ni:, ni:] @ AS, Ap[ni:, :ni]) < 1e-12
    assert relerr(Ap[ni:, ni:] @ BS, bp[ni:]) < 1e-12
    # and the condensed system itself is the Schur complement of the oracle's matrix with the GPU's factors
    assert relerr(Ap[:ni, :ni] - Ap[:ni, ni:] @ AS, Aii) < 1e-12
    O.use_blas(True, threads=1)
    eng.close()


def test_mixed_signatures_up_to_p5(oracle, gpu):
    """anisotropic random signatures with orders up to 5 in one call, ultraweak Maxwell and primal Poisson DPG"""
    from hp3d_b200.api import ElemEngine
    from tests.util import random_signature
    O = oracle
    O.set_maxp(6)
    O.use_blas(True, threads=8)
    rng = np.random.default_rng(55)
    sigs = [random_signature(rng, pmax=5) for _ in range(4)]
    nel = len(sigs)
    norder = np.stack([s[0] for s in sigs]); norie = np.stack([s[1] for s in sigs]); norif = np.stack([s[2] for s in sigs])
    nHs = [O.celndof(s[0])[0] for s in sigs]
    X = np.zeros((nel, max(nHs), 3))
    for e in range(nel):
        X[e, :nHs[e]] = hexa_xnod(nHs[e], h=0.4, jitter=0.1, curved=0.005, rng=rng)
    for kind in (2, 4):
        om = 2 * np.pi if kind == 4 else 1.0
        prm = O.default_params(omega=om)
        eng = ElemEngine(kind, omega=om)
        res = eng.elem_stc_batch(norder, norie, norif, X)
        assert (res["info"] == 0).all()
        for e in range(nel):
            Aii, Bi, AS, BS = eng.unpack(res, e)
            rA, rB, _, _ = O.condensed(kind, norder[e], norie[e], norif[e], X[e, :nHs[e]], prm)
            assert relerr(Aii, rA) < 1e-12 and relerr(Bi, rB) < 1e-12, (kind, e, relerr(Aii, rA), relerr(Bi, rB))
        eng.close()
    O.use_blas(True, threads=1)
