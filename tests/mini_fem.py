"""A tiny single-process H1 finite-element harness on structured cube meshes (test infrastructure).

It plays the role of the reference's celem_systemI + mumps_sc + solout around the hot path for the known-answer tests
trunk/test/poly_pois.F90 and conv_pois.F90: take per-element (condensed) matrices from a provider (the CPU oracle or the
GPU library), assemble the interface system with Dirichlet data, solve it densely, and recover the bubble dofs with the
stored Schur factors (stc_bwd, src/modules/stc.F90:661-677).

On an N^3 structured mesh every local edge/face axis points along +x/+y/+z (element_data.F90:67-93), so all orientations
are 0 and a global dof is identified by, per axis, either a grid vertex coordinate or (cell, 1-D mode index >= 2).
"""
import numpy as np

from tests.util import VERT, dof_map, uniform_order


class CubeMeshH1:
    def __init__(self, gpulib, N, p):
        self.N, self.p, self.h = N, p, 1.0 / N
        self.norder = uniform_order(p)
        _, idx, sgn = dof_map(gpulib, 0, self.norder, np.zeros(12, np.int32), np.zeros(6, np.int32))
        assert (sgn == 1).all()
        self.idx = idx
        self.nloc = len(idx)
        self.nbub = (p - 1) ** 3
        self.nint_loc = self.nloc - self.nbub      # interface dofs come first in the local order
        keys = {}
        self.l2g = np.zeros((N ** 3, self.nint_loc), dtype=np.int64)
        self.cells = [(i, j, k) for k in range(N) for j in range(N) for i in range(N)]
        for e, c in enumerate(self.cells):
            for a in range(self.nint_loc):
                key = tuple(("v", c[d] + idx[a, d]) if idx[a, d] < 2 else ("m", c[d], idx[a, d]) for d in range(3))
                self.l2g[e, a] = keys.setdefault(key, len(keys))
        self.keys = keys
        self.ndof = len(keys)
        # Dirichlet dofs: any axis sits on the cube boundary
        self.bdry = np.zeros(self.ndof, bool)
        self.vertex_xyz = {}
        for key, g in keys.items():
            if any(t[0] == "v" and t[1] in (0, N) for t in key):
                self.bdry[g] = True
            if all(t[0] == "v" for t in key):
                self.vertex_xyz[g] = np.array([t[1] for t in key]) * self.h

    def xnod(self):
        X = np.zeros((self.N ** 3, self.nloc, 3))
        for e, c in enumerate(self.cells):
            X[e, :8] = (np.array(c) + VERT) * self.h
        return X

    def descriptors(self):
        nel = self.N ** 3
        return np.tile(self.norder, (nel, 1)), np.zeros((nel, 12), np.int32), np.zeros((nel, 6), np.int32), self.xnod()

    def solve(self, Aii, Bi, dirichlet=None):
        """Aii (nel,ni,ni), Bi (nel,ni) -> global interface solution (ndof,)."""
        K = np.zeros((self.ndof, self.ndof)); F = np.zeros(self.ndof)
        for e in range(self.N ** 3):
            g = self.l2g[e]
            K[np.ix_(g, g)] += Aii[e]
            F[g] += Bi[e]
        u = np.zeros(self.ndof)
        if dirichlet is not None:
            for g, x in self.vertex_xyz.items():
                if self.bdry[g]:
                    u[g] = dirichlet(x)
        free = ~self.bdry
        rhs = F[free] - K[np.ix_(free, self.bdry)] @ u[self.bdry]
        u[free] = np.linalg.solve(K[np.ix_(free, free)], rhs)
        return u

    def local_interface(self, u):
        return u[self.l2g]


def h1_seminorm_error_sq(oracle, mesh, u_loc_full, grad_exact, nq=4):
    """sum_K int_K |grad(u_h - u)|^2 with an nq^3 Gauss rule, shape functions from the oracle (element_error analogue)."""
    x1, w1 = oracle.gauss1(nq)
    z12, z6 = np.zeros(12, np.int32), np.zeros(6, np.int32)
    shp = []
    for k in range(nq):
        for j in range(nq):
            for i in range(nq):
                xi = np.array([x1[i], x1[j], x1[k]])
                s, g = oracle.shape3DH(xi, mesh.norder, z12, z6)
                shp.append((xi, w1[i] * w1[j] * w1[k], g))
    err = 0.0
    h = mesh.h
    for e, c in enumerate(mesh.cells):
        for xi, w, g in shp:
            x = (np.array(c) + xi) * h
            gh = (u_loc_full[e] @ g) / h            # affine cube cell: J = h I
            d = gh - grad_exact(x)
            err += w * h ** 3 * float(d @ d)
    return err


class CubeMeshHcurl:
    """H(curl) analogue of CubeMeshH1 for the Maxwell problems: global tangential dofs of the N^3 structured mesh (all
    orientations 0), `ncomp` components per dof interleaved as the reference stores multi-component variables
    (kk = (k-1)*NR_COMP + ivar; ultraweak Maxwell: ncomp = 2, the E- and H-traces)."""

    def __init__(self, gpulib, N, p, ncomp=1):
        self.N, self.p, self.h, self.ncomp = N, p, 1.0 / N, ncomp
        self.norder = uniform_order(p)
        fam, idx, sgn = dof_map(gpulib, 1, self.norder, np.zeros(12, np.int32), np.zeros(6, np.int32))
        assert (sgn == 1).all()
        nbub = 3 * p * (p - 1) ** 2
        self.nE = len(idx)
        self.nEi = self.nE - nbub                 # interface dofs come first (edges, faces)
        self.cells = [(i, j, k) for k in range(N) for j in range(N) for i in range(N)]
        keys = {}
        self.l2g = np.zeros((N ** 3, self.nEi * ncomp), dtype=np.int64)
        for e, c in enumerate(self.cells):
            for a in range(self.nEi):
                key = (int(fam[a]),) + tuple(("q", c[d], idx[a, d]) if d == fam[a] else (("v", c[d] + idx[a, d]) if idx[a, d] < 2 else ("m", c[d], idx[a, d]))
                                             for d in range(3))
                g = keys.setdefault(key, len(keys))
                for iv in range(ncomp):
                    self.l2g[e, a * ncomp + iv] = g * ncomp + iv
        self.ndof = len(keys) * ncomp
        on_bdry = np.zeros(len(keys), bool)
        for key, g in keys.items():
            if any(t[0] == "v" and t[1] in (0, N) for d, t in enumerate(key[1:]) if d != key[0]):
                on_bdry[g] = True
        self.bdry_scalar = on_bdry

    def descriptors(self):
        nel = self.N ** 3
        X = np.zeros((nel, (self.p + 1) ** 3, 3))
        for e, c in enumerate(self.cells):
            X[e, :8] = (np.array(c) + VERT) * self.h
        return np.tile(self.norder, (nel, 1)), np.zeros((nel, 12), np.int32), np.zeros((nel, 6), np.int32), X

    def solve(self, Aii, Bi, dirichlet_comps=(0,)):
        """Assemble the condensed systems and solve with homogeneous Dirichlet data on the boundary dofs of the listed
        components (the electric trace / field: a PEC cavity)."""
        K = np.zeros((self.ndof, self.ndof), complex); F = np.zeros(self.ndof, complex)
        for e in range(self.N ** 3):
            g = self.l2g[e]
            K[np.ix_(g, g)] += Aii[e]
            F[g] += Bi[e]
        fixed = np.zeros(self.ndof, bool)
        for iv in dirichlet_comps:
            fixed[iv::self.ncomp] = self.bdry_scalar
        free = ~fixed
        u = np.zeros(self.ndof, complex)
        u[free] = np.linalg.solve(K[np.ix_(free, free)], F[free])
        return u, K, F, free
