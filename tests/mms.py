"""Manufactured-solution harness around the drop-in boundary (test infrastructure).

It plays celem_systemI + mumps_sc + solout + element_error for GLOBAL known-answer solves whose element matrices come from a
*provider* -- the CPU oracle or the GPU library through the C ABI -- so that the same mathematics pins both:

  * trunk/test/conv_maxw.F90:96-116 and poly_maxw.F90:101 (Maxwell Galerkin on the refined unit cube, polynomial field with
    homogeneous tangential trace, files/mesh/hexa_orient_0): H(curl) error rate in [0.90, 1.10] at p = 1, <= 1e-13 at p = 2;
  * the ultraweak DPG Maxwell problem with the manufactured sin solution (problems/MAXWELL/ULTRAWEAK_DPG/common/
    mfd_solutions.F90:80-100, isol = 1) on meshes of hexahedra AND prisms with a random global vertex numbering
    (hp3d_b200.synth.hp_mesh): the L2 error of (E, H) and the DPG residual decay at rate p.

A wrong sign, dof order, orientation table, Piola map, Gram entry or condensation step breaks these at O(1); they do not
depend on the oracle being a faithful port.  Global H(curl) dofs are (entity, k) keys as in tests/mini_fem_hp.py.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from tests.test_hp_mesh_conformity import entity_blocks, topo


# ---- the polynomial field of conv_maxw.F90 / poly_maxw.F90 (subroutine exact) and its getf -----------------------------------------
def poly_E(x):
    X, Y, Z = x[..., 0], x[..., 1], x[..., 2]
    return np.stack([Y * (1 - Y) * Z * (1 - Z), Y * X * (1 - X) * Z * (1 - Z), X * (1 - X) * Y * (1 - Y)], -1)


def poly_curlE(x):
    X, Y, Z = x[..., 0], x[..., 1], x[..., 2]
    return np.stack([X * (1 - X) * (1 - 2 * Y) - Y * X * (1 - X) * (1 - 2 * Z),
                     Y * (1 - Y) * (1 - 2 * Z) - (1 - 2 * X) * Y * (1 - Y),
                     Y * (1 - 2 * X) * Z * (1 - Z) - (1 - 2 * Y) * Z * (1 - Z)], -1)


def poly_curlcurlE(x):
    X, Y, Z = x[..., 0], x[..., 1], x[..., 2]
    return np.stack([(1 - 2 * X) * Z * (1 - Z) + 2 * Z * (1 - Z) + 2 * Y * (1 - Y),
                     2 * Y * Z * (1 - Z) + 2 * Y * X * (1 - X),
                     X * (1 - X) * (1 - 2 * Z) + 2 * Y * (1 - Y) + 2 * X * (1 - X)], -1)


def poly_J(x, omega, eps=1.0, mu=1.0):
    """getf of the two tests: -i w J = curl(1/mu curl E) - w^2 eps E  (sigma = 0)."""
    return (poly_curlcurlE(x) / mu - omega * omega * eps * poly_E(x)) / (-1j * omega)


# ---- providers ---------------------------------------------------------------------------------------------------------------------
class OracleProvider:
    """Element matrices / errors / residuals from the CPU oracle (one element at a time)."""
    name = "oracle"

    def __init__(self, O, kind, maxp=8, **prm):
        self.O, self.kind, self.prm_kw, self.maxp = O, kind, prm, maxp
        O.set_maxp(maxp)
        O.use_blas(True)

    def _prm(self, **extra):
        return self.O.default_params(**{**self.prm_kw, **extra})

    def quad_points(self, m):
        out = []
        for e in range(len(m["etype"])):
            nH = int(m["nrdofH"][e]); et = int(m["etype"][e])
            xi, _ = self.O.quad3(m["norder"][e], m["norient_face"][e], 0, self.maxp, et)
            S = np.array([self.O.shape3DH(x, m["norder"][e], m["norient_edge"][e], m["norient_face"][e], et)[0] for x in xi])
            out.append(S[:, :nH] @ m["xnod"][e, :nH])
        return out

    def condensed(self, m, source=None):
        import ctypes as C
        res = []
        for e in range(len(m["etype"])):
            nH = int(m["nrdofH"][e])
            if source is None:
                prm = self._prm()
            else:
                tab = np.ascontiguousarray(source[e], dtype=np.complex128)
                prm = self._prm(source=9, source_table=tab.ctypes.data_as(C.c_void_p))
            res.append(self.O.condensed(self.kind, m["norder"][e], m["norient_edge"][e], m["norient_face"][e], m["xnod"][e, :nH], prm,
                                        etype=int(m["etype"][e])))
        return res

    def error_points(self, m):
        return [self.O.error_points(m["norder"][e], m["norient_edge"][e], m["norient_face"][e], m["xnod"][e, :int(m["nrdofH"][e])],
                                    int(m["etype"][e])) for e in range(len(m["etype"]))]

    def error(self, m, zdof, exact=None):
        err = rn = 0.0
        prm = self._prm()
        for e in range(len(m["etype"])):
            a, b, _ = self.O.element_error(self.kind, m["norder"][e], m["norient_edge"][e], m["norient_face"][e],
                                           m["xnod"][e, :int(m["nrdofH"][e])], zdof[e], prm,
                                           exact_tab=None if exact is None else exact[e], etype=int(m["etype"][e]))
            err += a; rn += b
        return err, rn

    def residual(self, m, xi, xb):
        """sum_K (G^-1 (l - B u), l - B u) from the oracle's Gram matrix and enriched stiffness (ultraweak Maxwell)."""
        tot = 0.0
        prm = self._prm()
        for e in range(len(m["etype"])):
            et = int(m["etype"][e]); nH = int(m["nrdofH"][e])
            args = (m["norder"][e], m["norient_edge"][e], m["norient_face"][e], m["xnod"][e, :nH], prm)
            _, _, G, S = self.O.elem(self.kind, *args, want_dpg=True, etype=et)
            perm, ni, nb = self.O.stc_partition(self.kind, m["norder"][e], et)
            u = np.zeros(ni + nb, complex)
            u[perm[:ni]] = xi[e][:ni]; u[perm[ni:]] = xb[e][:nb]
            Gu = np.triu(G); Gf = Gu + np.triu(Gu, 1).conj().T
            r = S[:, -1] - S[:, :-1] @ u
            tot += float(np.real(np.vdot(r, np.linalg.solve(Gf, r))))
        return tot


class GpuProvider:
    """The same quantities from the CUDA library through the C ABI (hp3d_gpu_elem_batch, hp3d_gpu_stc_bwd_batch semantics on the
    returned Schur factors, hp3d_gpu_elem_error_batch, hp3d_gpu_elem_residual_batch)."""
    name = "gpu"

    def __init__(self, kind, maxp=8, **prm):
        from hp3d_b200.api import ElemEngine
        self.kind, self.prm_kw, self.maxp = kind, prm, maxp
        self.eng = ElemEngine(kind, maxp=maxp, **prm)
        self.eng_tab = None

    def close(self):
        self.eng.close()
        if self.eng_tab is not None:
            self.eng_tab.close()

    def _d(self, m):
        return m["norder"], m["norient_edge"], m["norient_face"], m["xnod"]

    def quad_points(self, m):
        xq = self.eng.quad_points(*self._d(m), etype=m["etype"])
        return [xq[e, :self.eng.sizes(m["norder"][e], int(m["etype"][e]))[2]] for e in range(len(m["etype"]))]

    def condensed(self, m, source=None):
        from hp3d_b200.api import ElemEngine
        eng, src = self.eng, None
        if source is not None:
            if self.eng_tab is None:
                self.eng_tab = ElemEngine(self.kind, maxp=self.maxp, **{**self.prm_kw, "source": 9})
            eng = self.eng_tab
            nmax = max(len(s) for s in source)
            src = np.zeros((len(source), nmax, 3), complex)
            for e, s in enumerate(source):
                src[e, :len(s)] = s
        res = eng.elem_stc_batch(*self._d(m), source_qp=src, etype=m["etype"])
        assert (res["info"] == 0).all(), res["info"]
        return [eng.unpack(res, e) for e in range(len(m["etype"]))]

    def error_points(self, m):
        xq, nint = self.eng.error_points(*self._d(m), etype=m["etype"])
        return [xq[e, :nint[e]] for e in range(len(nint))]

    def error(self, m, zdof, exact=None):
        nel = len(m["etype"])
        nF = max(z.shape[0] for z in zdof); nc = zdof[0].shape[1]
        Z = np.zeros((nel, nF, nc), complex)
        for e, z in enumerate(zdof):
            Z[e, :z.shape[0]] = z
        tab = None
        if exact is not None:
            nmax = max(len(t) for t in exact)
            tab = np.zeros((nel, nmax, exact[0].shape[1]), complex)
            for e, t in enumerate(exact):
                tab[e, :len(t)] = t
        r = self.eng.elem_error_batch(*self._d(m), Z, exact_qp=tab, etype=m["etype"])
        assert (r["info"] == 0).all()
        return float(r["err"].sum()), float(r["rnorm"].sum())

    def residual(self, m, xi, xb):
        nel = len(m["etype"])
        XI = np.zeros((nel, max(len(x) for x in xi)), complex); XB = np.zeros((nel, max(max(len(x) for x in xb), 1)), complex)
        for e in range(nel):
            XI[e, :len(xi[e])] = xi[e]; XB[e, :len(xb[e])] = xb[e]
        r = self.eng.elem_residual_batch(*self._d(m), XI, XB, etype=m["etype"])
        assert (r["info"] == 0).all()
        return float(r["resid"].sum())


# ---- global H(curl) space on a synth.hp_mesh (entity keys) ---------------------------------------------------------------------------
class HcurlSpace:
    """Global tangential dofs of a conforming hexa/prism mesh: dof = ((entity), k) with `ncomp` interleaved components
    (kk = (k-1)*NR_COMP + ivar, celem_system.F90:595).  l2g[e] lists the global indices of element e's INTERFACE dofs in the
    reference's local order (edges, then faces)."""

    def __init__(self, m, ncomp=1):
        self.m, self.ncomp = m, ncomp
        coords = m["coords"]
        keys, self.l2g, bdry = {}, [], set()

        def on_boundary(vs):
            P = coords[list(vs)]
            return any(np.all(np.abs(P[:, a] - s) < 1e-12) for a in range(3) for s in (0.0, 1.0))

        for e in range(len(m["etype"])):
            et = int(m["etype"][e]); E, F, _ = topo(et)
            v = [int(x) for x in m["verts"][e] if x >= 0]
            blocks, nint = entity_blocks(et, m["norder"][e], "E")
            g = np.zeros(nint * ncomp, np.int64)
            for kind, idx, m0, n in blocks:
                vs = (v[E[idx][0]], v[E[idx][1]]) if kind == "e" else tuple(v[i] for i in F[idx])
                for k in range(n):
                    key = (kind, frozenset(vs), k)
                    gi = keys.setdefault(key, len(keys))
                    if on_boundary(vs):
                        bdry.add(gi)
                    for iv in range(ncomp):
                        g[(m0 + k) * ncomp + iv] = gi * ncomp + iv
            self.l2g.append(g)
        self.nscalar = len(keys)
        self.ndof = self.nscalar * ncomp
        self.bdry_scalar = np.zeros(self.nscalar, bool)
        self.bdry_scalar[list(bdry)] = True

    def solve(self, mats, dirichlet_comps=(0,)):
        """Assemble the condensed systems (Aii, Bi) and solve with homogeneous Dirichlet data on the boundary dofs of the listed
        components; returns the global vector."""
        rows, cols, vals = [], [], []
        F = np.zeros(self.ndof, complex)
        for e, (Aii, Bi, _, _) in enumerate(mats):
            g = self.l2g[e]
            assert Aii.shape[0] == len(g), (Aii.shape, len(g))
            rows.append(np.repeat(g, len(g))); cols.append(np.tile(g, len(g))); vals.append(np.asarray(Aii).ravel())
            np.add.at(F, g, Bi)
        K = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(self.ndof, self.ndof)).tocsc()
        fixed = np.zeros(self.ndof, bool)
        for iv in dirichlet_comps:
            fixed[iv::self.ncomp] = self.bdry_scalar
        free = np.flatnonzero(~fixed)
        u = np.zeros(self.ndof, complex)
        u[free] = spla.splu(K[free][:, free]).solve(F[free])
        return u


def structured_mesh(N, p):
    """The reference's test mesh (files/mesh/hexa_orient_0 after global_href): N^3 congruent bricks, every local axis along +x/+y/+z, so all
    orientations are 0 -- in the dict layout of synth.hp_mesh."""
    from hp3d_b200 import synth
    m = synth.hp_mesh(N, prism_frac=0.0, pmin=p, pmax=p, seed_g=0, rotate_local=False)
    n1 = N + 1
    m["gid"] = np.arange(n1 ** 3)
    m["norient_edge"][:] = 0; m["norient_face"][:] = 0
    return m


# ---- the three global known-answer solves -------------------------------------------------------------------------------------------
def maxwell_galerkin_poly_error(prov, m, omega=np.pi):
    """One solve of conv_maxw / poly_maxw: Maxwell Galerkin with getf from the polynomial field, homogeneous tangential data;
    returns sqrt(sum errorE) as element_error accumulates it (values + curl)."""
    space = HcurlSpace(m, 1)
    src = [poly_J(x, omega) for x in prov.quad_points(m)]
    mats = prov.condensed(m, source=src)
    u = space.solve(mats)
    zdof = []
    for e, (Aii, Bi, AS, BS) in enumerate(mats):
        xi = u[space.l2g[e]]
        xb = BS - AS @ xi if AS.shape[0] else np.zeros(0, complex)
        zdof.append(np.concatenate([xi, xb])[:, None])
    exact = [np.concatenate([poly_E(x), poly_curlE(x)], -1).astype(complex) for x in prov.error_points(m)]
    err, rn = prov.error(m, zdof, exact)
    return np.sqrt(err), np.sqrt(rn)


def uw_maxwell_sin_solution(prov, m):
    """Ultraweak DPG Maxwell with the built-in manufactured sin solution: global trace solve, bubbles (the L2 fields E, H) through the
    Schur factors, then (relative L2 error of (E,H), sqrt of the summed DPG residual)."""
    space = HcurlSpace(m, 2)
    mats = prov.condensed(m)
    u = space.solve(mats, dirichlet_comps=(0,))
    xi = [u[space.l2g[e]] for e in range(len(mats))]
    xb = [mats[e][3] - mats[e][2] @ xi[e] for e in range(len(mats))]
    zdof = [b.reshape(-1, 6) for b in xb]
    err, rn = prov.error(m, zdof)
    return np.sqrt(err / rn), np.sqrt(prov.residual(m, xi, xb))
