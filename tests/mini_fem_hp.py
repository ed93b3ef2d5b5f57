"""Global H1 dof numbering and a least-squares polynomial fit on the mixed hexa/prism hp meshes of hp3d_b200.synth.hp_mesh
(test infrastructure).  A global dof is (entity, k): the k-th function of a vertex / edge / face in the reference's local
order -- orientation-embedded shape functions make that identification valid from every adjacent element."""
import numpy as np

from tests.test_hp_mesh_conformity import entity_blocks, topo


def u_exact(x):
    return 1.0 + 0.5 * x[..., 0] - x[..., 1] + x[..., 0] * x[..., 0] + 2.0 * x[..., 1] * x[..., 2] - x[..., 0] * x[..., 2] + 0.25 * x[..., 2] ** 2


F_SOURCE = -(2.0 + 0.5)   # f = -Laplace(u_exact)


def build_space(m):
    """-> keys {(entity..., k): global index}, l2g [per element global indices, interface dofs first], nint [# interface dofs],
    is_bdry, is_bubble (ndof,)"""
    nel = len(m["etype"])
    keys, l2g, nloc, bdry_keys = {}, [], [], set()
    coords = m["coords"]

    def on_boundary(vs):
        P = coords[list(vs)]
        return any(np.all(np.abs(P[:, a] - s) < 1e-12) for a in range(3) for s in (0.0, 1.0))

    for e in range(nel):
        et = int(m["etype"][e]); E, F, M = topo(et)
        v = [int(x) for x in m["verts"][e] if x >= 0]
        blocks, nint = entity_blocks(et, m["norder"][e], "H")
        g = []
        for kind, idx, m0, n in blocks:
            if kind == "v":
                ent = ("v", v[idx]); vs = (v[idx],)
            elif kind == "e":
                vs = (v[E[idx][0]], v[E[idx][1]]); ent = ("e", frozenset(vs))
            else:
                vs = tuple(v[i] for i in F[idx]); ent = ("f", frozenset(vs))
            for k in range(n):
                key = ent + (k,)
                g.append(keys.setdefault(key, len(keys)))
                if on_boundary(vs):
                    bdry_keys.add(key)
        for k in range(int(m["nrdofH"][e]) - nint):
            g.append(keys.setdefault(("b", e, k), len(keys)))
        l2g.append(np.array(g)); nloc.append(nint)
    ndof = len(keys)
    is_b = np.zeros(ndof, bool); is_bub = np.zeros(ndof, bool)
    for key, gidx in keys.items():
        is_b[gidx] = key in bdry_keys
        is_bub[gidx] = key[0] == "b"
    return keys, l2g, nloc, is_b, is_bub


def fit_polynomial(oracle, m, l2g, ndof, seed=5):
    """Least-squares fit of u_exact in the global conforming space; returns (U, max residual)."""
    rng = np.random.default_rng(seed)
    rows, rhs = [], []
    for e in range(len(m["etype"])):
        et = int(m["etype"][e]); nH = int(m["nrdofH"][e]); nv = 8 if et == 1 else 6
        for _ in range(2 * nH):
            xi = rng.random(3)
            if et == 3 and xi[0] + xi[1] > 1:
                xi[:2] = 1 - xi[:2]
            s, _ = oracle.shape3DH(xi, m["norder"][e], m["norient_edge"][e], m["norient_face"][e], et)
            x = s[:nv] @ m["xnod"][e, :nv]
            r = np.zeros(ndof); r[l2g[e]] = s
            rows.append(r); rhs.append(u_exact(x))
    A = np.array(rows); b = np.array(rhs)
    U = np.linalg.lstsq(A, b, rcond=None)[0]
    return U, float(np.abs(A @ U - b).max())
