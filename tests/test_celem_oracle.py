"""Pins of the oracle's restatement of celem_systemI.F90:543-785 + par_mumps_sc.F90:419-448 (oracle/celem.c) and of the
product's host-side packing (hp3d_gpu_celem_pack) -- CPU only.

The reference's tests hold no vectors for this routine (it is exercised through the solves of poly_pois/poly_maxw on
REGULAR meshes, where the transform is the identity).  Pins: (a) identity case == permutation of ALOC, exact; (b) the
defining algebra ZAMOD = C^T A C, ZBMOD = C^T b - ZAMOD z_D with C built from the PRODUCT's packed lists; (c) a 1-irregular
mesh with hanging nodes on which u = xyz must be reproduced (test/poly_pois.F90's criterion, 1e-13).
"""
import numpy as np
import pytest

from hp3d_b200 import api
from tests import celem_util as CU


def dense_C(c, ni):
    nm = len(c["idbc"])
    Cm = np.zeros((ni, nm))
    for g in range(nm):
        for q in range(c["cptr"][g], c["cptr"][g + 1]):
            Cm[c["cidx"][q] - 1, g] += c["cval"][q]
    return Cm


@pytest.mark.parametrize("kind,cplx", [(1, False), (2, False), (3, True), (4, True)])
def test_identity_is_a_permutation(oracle, gpulib, kind, cplx):
    O = oracle
    rng = np.random.default_rng(kind)
    no = O.uniform_order(2)
    c = CU.random_constraints(rng, O, api, kind, no, O.MDLB, cplx, frac_con=0.0, frac_dbc=0.0, extra=0)
    ni = sum(n * v for n, v in zip(c["nrdofl"], [max(c["ph"].nrvar[f], 0) for f in range(3)]))
    A = rng.standard_normal((ni, ni)) + (1j * rng.standard_normal((ni, ni)) if cplx else 0)
    b = rng.standard_normal(ni) + (1j * rng.standard_normal(ni) if cplx else 0)
    Cm = dense_C(c, ni)
    assert ((Cm != 0).sum(0) == 1).all() and ((Cm != 0).sum(1) == 1).all()
    rowof = np.argmax(Cm, axis=0)           # modified dof -> element dof
    idx = rowof[c["nextract"] - 1]
    for isym in (1, 2, 3):
        zb, za = CU.oracle_celem(O, c, A, b, isym)
        assert np.array_equal(zb, b[idx])
        S = A[np.ix_(idx, idx)]
        if isym == 2:
            assert np.array_equal(za.reshape(len(idx), len(idx)), S)
        elif isym == 3:
            assert np.array_equal(za.reshape(len(idx), len(idx)).T, S)
        else:
            il = np.tril_indices(len(idx))
            assert np.array_equal(za, ((S + S.T) / 2.0)[il])


@pytest.mark.parametrize("kind,cplx,p", [(1, False, 3), (2, False, 2), (3, True, 2), (4, True, 2)])
def test_transform_algebra(oracle, gpulib, kind, cplx, p):
    O = oracle
    rng = np.random.default_rng(10 + kind)
    no = O.uniform_order(p)
    c = CU.random_constraints(rng, O, api, kind, no, O.MDLB, cplx)
    ni = int(max(c["cidx"]))
    A = rng.standard_normal((ni, ni)) + (1j * rng.standard_normal((ni, ni)) if cplx else 0)
    b = rng.standard_normal(ni) + (1j * rng.standard_normal(ni) if cplx else 0)
    Cm = dense_C(c, ni)
    Zm = Cm.T @ A @ Cm
    zbm = Cm.T @ b - Zm @ c["zdofd"]
    x = c["nextract"] - 1
    zb, za, zam = O.celem_modify(c["ph"], c["nrdofl"], c["nrcon"], c["nac"], c["constr"], c["nrdofm_f"], A, b, c["idbc"], c["zdofd"],
                                 c["nextract"], 2, want_zamod=True)
    sc = np.abs(Zm).max()
    assert np.abs(zam - Zm).max() < 1e-13 * sc
    assert np.abs(za.reshape(len(x), len(x)) - Zm[np.ix_(x, x)]).max() < 1e-13 * sc
    assert np.abs(zb - zbm[x]).max() < 1e-12 * max(np.abs(zbm).max(), 1)
    # COO fill: triplets in the element loop's order, load accumulated into the global vector
    a, irn, jcn, rhs = O.coo_fill(c["lcon"], za, zb, int(c["lcon"].max()))
    n = len(x)
    assert np.array_equal(irn.reshape(n, n), np.repeat(c["lcon"][:, None], n, 1)) and np.array_equal(jcn.reshape(n, n), np.repeat(c["lcon"][None, :], n, 0))
    assert np.array_equal(a, za.astype(np.complex128)) and np.allclose(rhs[c["lcon"] - 1], zb)


def hanging_solve(O, mesh, systems):
    """Assemble the compressed element systems through LCON and solve (the role of MUMPS)."""
    nfree = int((~mesh.bdry).sum())
    K = np.zeros((nfree, nfree)); F = np.zeros(nfree)
    for c, (zb, za) in systems:
        n = len(c["lcon"])
        a, irn, jcn, rhs = O.coo_fill(c["lcon"], za, zb, nfree)
        np.add.at(K, (irn - 1, jcn - 1), a.real)
        F += rhs.real
    return np.linalg.solve(K, F)


def test_hanging_nodes_reproduce_xyz(oracle, gpulib):
    """poly_pois.F90 on a 1-irregular mesh: u = xyz (Laplace u = 0) is in the constrained p=1 space, so the constrained
    assembly + Dirichlet lift must return it at every free regular vertex."""
    O = oracle
    mesh = CU.HangingMesh(N=3, ref=(1, 1, 1))
    assert len(mesh.parents) == 12 + 6           # all edge midpoints and face centres of the interior brick hang
    uex = lambda x: x[0] * x[1] * x[2]           # noqa: E731
    free = np.flatnonzero(~mesh.bdry)
    numbering = {int(g): i + 1 for i, g in enumerate(free)}
    no, noe, nof, X = mesh.descriptors()
    cons = mesh.constraints(api, O, uex, numbering)
    prm = O.default_params(source=0)
    systems = []
    for e in range(len(cons)):
        A, b, _, _ = O.condensed(O.POIS_GAL, no[e], noe[e], nof[e], X[e], prm)
        systems.append((cons[e], CU.oracle_celem(O, cons[e], A, b, 2)))
    u = hanging_solve(O, mesh, systems)
    exact = np.array([uex(mesh.xyz[g]) for g in free])
    assert len(free) == 8 + 1 and np.abs(u - exact).max() < 1e-13
    # boundary-touching refinement: some edge midpoints / face centres are regular boundary nodes instead
    mesh2 = CU.HangingMesh(N=2, ref=(0, 0, 0))
    assert len(mesh2.parents) == 9 + 3


def test_pack_rejects_bad_input(gpulib):
    ph = api.physics_default(1)
    z = np.zeros((0, 2))
    with pytest.raises(RuntimeError, match="outside the modified element"):
        api.celem_pack(ph, [2, 0, 0], [[1, 1], [], []], [np.array([[1, 0], [5, 0]]), z, z], [np.ones((2, 2)), z, z], [3, 0, 0])
