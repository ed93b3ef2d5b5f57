"""Pins of the headline path by MATHEMATICS rather than by the port (SURVEY.md 7 step 0 iii, 8c):

  1. the reference's two Maxwell solve tests replayed GLOBALLY -- trunk/test/conv_maxw.F90:96-116 (p = 1, three refinements, H(curl)
     error rate in [0.90, 1.10]) and trunk/test/poly_maxw.F90:101 (p = 2, polynomial field reproduced to 1e-14);
  2. the ultraweak DPG Maxwell problem with the reference's manufactured sin solution (mfd_solutions.F90:80-100, isol = 1) on
     meshes of bricks, of prisms and of both with a random global vertex numbering: the relative L2 error of (E, H) decays at
     rate p, the DPG residual with it;
  3. the reference's scalar-loop twin elem_maxwell.F90 == its BLAS3 elem_opt.F90 (Gram matrix, enriched stiffness, element
     matrices; arbitrary complex permittivity tensor, all test norms, bricks and prisms).

Every test exists twice: with the CPU oracle as the provider of the element matrices (pins the oracle; runs without a GPU) and
with the CUDA library through the C ABI (pins the product; `-m gpu`)."""
import numpy as np
import pytest

from hp3d_b200 import synth
from tests import mms
from tests.util import hexa_xnod, random_signature


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def rate(a, b):
    return float(np.log(a / b) / np.log(2.0))


# ---- 3. scalar-loop twin ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("test_norm", [1, 2, 3])
@pytest.mark.parametrize("etype", [1, 3])
def test_scalar_twin_equals_blas3_formulation(oracle, test_norm, etype):
    from tests.test_gpu_prism import prism_xnod
    from tests.test_oracle_prism import prism_signature
    O = oracle
    O.set_maxp(6); O.use_blas(True)
    rng = np.random.default_rng(10 * test_norm + etype)
    for trial in range(2):
        if etype == 1:
            no, ne, nf = random_signature(rng, pmax=3)
            X = hexa_xnod(O.celndof(no)[0], h=0.5, jitter=0.15, curved=0.02, rng=rng)
        else:
            no, ne, nf = prism_signature(rng, int(rng.integers(1, 4)), int(rng.integers(1, 4)), uniform=False)
            X = prism_xnod(O.celndof(no, 3)[0], rng, curved=0.01)
        kw = dict(omega=1.7 * np.pi, eps=1.3, mu=0.8, alpha_norm=0.6, test_norm=test_norm, icomp_exact=2)
        if trial:   # get_permittivity returning a general complex tensor (elem_opt.F90:260-266, elem_maxwell.F90:246-253)
            kw["eps_tensor"] = np.eye(3) + 0.2 * rng.standard_normal((3, 3)) + 0.1j * rng.standard_normal((3, 3))
        prm = O.default_params(**kw)
        A, b, G, S = O.elem(4, no, ne, nf, X, prm, want_dpg=True, etype=etype)
        A2, b2, G2, S2 = O.elem_uw_scalar(no, ne, nf, X, prm, etype=etype)
        assert rel(S2, S) < 1e-14
        if test_norm == 1:
            assert rel(np.triu(G2), np.triu(G)) < 1e-14
        assert rel(A2, A) < 1e-12 and rel(b2, b) < 1e-12, (rel(A2, A), rel(b2, b))


# ---- 1. conv_maxw / poly_maxw -------------------------------------------------------------------------------------------------------
def _conv_maxw(prov):
    errs = [mms.maxwell_galerkin_poly_error(prov, mms.structured_mesh(N, 1))[0] for N in (2, 4, 8)]   # nref = 3
    for a, b in zip(errs, errs[1:]):
        assert 0.90 < rate(a, b) < 1.10, (errs, rate(a, b))                                            # conv_maxw.F90:114
    return errs


def _poly_maxw(prov, tol):
    for N in (2, 4):                                                                                    # nref = 2
        err, nrm = mms.maxwell_galerkin_poly_error(prov, mms.structured_mesh(N, 2))
        assert abs(nrm - 7.0 / 30.0) < 1e-12           # (|E|^2 + |curl E|^2 of the polynomial field over the unit cube)^(1/2) = 7/30
        assert err < tol, (N, err)                     # poly_maxw.F90:101 (1e-14 on its own error norm)


def test_conv_maxw_oracle(oracle):
    _conv_maxw(mms.OracleProvider(oracle, 3, omega=np.pi))


def test_poly_maxw_oracle(oracle):
    _poly_maxw(mms.OracleProvider(oracle, 3, omega=np.pi), 1e-14)


@pytest.mark.gpu
def test_conv_maxw_gpu(oracle, gpu):
    prov = mms.GpuProvider(3, omega=np.pi)
    errs = _conv_maxw(prov)
    ref = [mms.maxwell_galerkin_poly_error(mms.OracleProvider(oracle, 3, omega=np.pi), mms.structured_mesh(N, 1))[0] for N in (2, 4)]
    assert abs(errs[0] - ref[0]) < 1e-10 * ref[0] and abs(errs[1] - ref[1]) < 1e-10 * ref[1]   # north_star: solution error within 1e-10
    prov.close()


@pytest.mark.gpu
def test_poly_maxw_gpu(gpu):
    prov = mms.GpuProvider(3, omega=np.pi)
    _poly_maxw(prov, 1e-13)
    prov.close()


# ---- 2. ultraweak DPG Maxwell, manufactured solution ----------------------------------------------------------------------------------
MESHES = {"bricks": dict(prism_frac=0.0, jitter=0.0), "prisms": dict(prism_frac=1.0, jitter=0.0), "mixed": dict(prism_frac=0.5, jitter=0.1)}


def _uw_rates(prov, which, p, Ns):
    out = []
    for N in Ns:
        m = synth.hp_mesh(N, pmin=p, pmax=p, seed_p=5, seed_g=11, **MESHES[which])
        if which != "bricks":
            assert (m["etype"] == 3).any()
        out.append(mms.uw_maxwell_sin_solution(prov, m))
    return out


def _check_uw(out, p, asymptotic):
    (e0, r0), (e1, r1) = out[-2], out[-1]
    lo = 0.9 * p if asymptotic else 0.85 * p
    assert lo < rate(e0, e1) < 1.15 * p + 0.1, (out, rate(e0, e1))         # L2 error of (E, H): O(h^p) (L2 fields of degree p-1)
    assert rate(r0, r1) > (0.7 if asymptotic else 0.55) * p, (out, rate(r0, r1))   # the residual (energy error) decays with it
    assert e1 < e0 and r1 < r0


@pytest.mark.parametrize("which", ["bricks", "prisms", "mixed"])
def test_uw_maxwell_manufactured_solution_oracle(oracle, which):
    prov = mms.OracleProvider(oracle, 4, omega=np.pi)
    _check_uw(_uw_rates(prov, which, 1, (4, 8)), 1, True)
    _check_uw(_uw_rates(prov, which, 2, (2, 4)), 2, False)


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["bricks", "prisms", "mixed"])
def test_uw_maxwell_manufactured_solution_gpu(oracle, gpu, which):
    prov = mms.GpuProvider(4, omega=np.pi)
    _check_uw(_uw_rates(prov, which, 1, (4, 8)), 1, True)
    out = _uw_rates(prov, which, 2, (2, 4))
    _check_uw(out, 2, False)
    if which == "bricks":
        _check_uw(_uw_rates(prov, which, 3, (2, 4)), 3, False)
    # the same solve with the oracle's matrices: error and residual agree (north_star: final solution error within 1e-10)
    ref = _uw_rates(mms.OracleProvider(oracle, 4, omega=np.pi), which, 2, (2,))[0]
    assert abs(out[0][0] - ref[0]) < 1e-10 * ref[0], (out[0], ref)
    assert abs(out[0][1] - ref[1]) < 1e-8 * ref[1], (out[0], ref)
    prov.close()


@pytest.mark.gpu
def test_uw_maxwell_manufactured_solution_gpu_complex_kernels(gpu):
    """the general complex dense phase (real_reduction = 0: the reference's ZPOTRF / ZTRTRS / ZHERK sequence) through the same pin"""
    prov = mms.GpuProvider(4, omega=np.pi, real_reduction=0)
    _check_uw(_uw_rates(prov, "mixed", 2, (2, 4)), 2, False)
    prov.close()
