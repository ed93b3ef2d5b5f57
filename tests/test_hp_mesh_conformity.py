"""Conformity pin for the orientation handling (oracle): on a conforming hp mesh of hexahedra and prisms with a random
global vertex numbering (hp3d_b200.synth.hp_mesh: min-rule orders, edge / face orientations derived from the numbering as
Orient.F90 defines them), the H1 functions and the tangential H(curl) traces of the two elements sharing a face must
coincide function by function on that face, and every function not attached to the face's closure must vanish there.
This is what `find_orient` + the orientation-embedded shape functions guarantee in the reference (global dof
connectivity relies on it); a wrong sign, dof order or orientation table in the restatement breaks it at O(1)."""
import itertools

import numpy as np

from hp3d_b200 import synth

MDLB, MDLP = 1, 3
BRICK_M = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)


def topo(et):
    return (synth.BRICK_EDGE, synth.BRICK_FACE, BRICK_M) if et == MDLB else (synth.PRISM_EDGE, synth.PRISM_FACE, synth.PRISM_VERT)


def entity_blocks(et, no, space):
    """[(kind, local index, first dof, count)] of the vertex / edge / face blocks in the reference's dof order."""
    E, F, _ = topo(et)
    out, m = [], 0
    if space == "H":
        for v in range(len(_verts(et))):
            out.append(("v", v, m, 1)); m += 1
    for ie in range(len(E)):
        n = no[ie] - 1 if space == "H" else no[ie]
        out.append(("e", ie, m, n)); m += n
    for jf, f in enumerate(F):
        o = no[len(E) + jf]
        if len(f) == 3:
            n = (o - 1) * (o - 2) // 2 if space == "H" else o * (o - 1)
        else:
            a, b = o // 10, o % 10
            n = (a - 1) * (b - 1) if space == "H" else a * (b - 1) + (a - 1) * b
        out.append(("f", jf, m, n)); m += n
    return out, m


def _verts(et):
    return range(8) if et == MDLB else range(6)


def test_shared_face_conformity(oracle):
    oracle.set_maxp(8)
    m = synth.hp_mesh(3, prism_frac=0.45, pmin=1, pmax=4, seed_p=11, seed_g=5)
    nel = len(m["etype"])
    faces = {}
    for e in range(nel):
        et = int(m["etype"][e]); E, F, M = topo(et)
        v = m["verts"][e]
        for jf, f in enumerate(F):
            faces.setdefault(frozenset(int(v[i]) for i in f), []).append((e, jf))
    rng = np.random.default_rng(0)
    checked = {"qq": 0, "tt": 0, "mixed": 0}
    for key, owners in faces.items():
        if len(owners) != 2:
            continue
        gv = sorted(key)
        w = rng.dirichlet(np.ones(len(gv)))
        if len(gv) == 4:   # a point of the (planar parallelogram) face: bilinear weights in a consistent cyclic order
            e0, jf0 = owners[0]
            cyc = [int(m["verts"][e0][i]) for i in topo(int(m["etype"][e0]))[1][jf0]]
            s, t = rng.random(2)
            w = dict(zip(cyc, [(1 - s) * (1 - t), s * (1 - t), s * t, (1 - s) * t]))
        else:
            w = dict(zip(gv, w))
        vals = []
        for (e, jf) in owners:
            et = int(m["etype"][e]); E, F, M = topo(et)
            v = [int(x) for x in m["verts"][e] if x >= 0]
            no, ne, nf = m["norder"][e], m["norient_edge"][e], m["norient_face"][e]
            xi = sum(w[g] * M[v.index(g)] for g in key)
            sH, gH = oracle.shape3DH(xi, no, ne, nf, et)
            sE, cE = oracle.shape3DE(xi, no, ne, nf, et)
            # master tangents of the face: towards the other face vertices from the smallest-id vertex
            tang = [M[v.index(g)] - M[v.index(gv[0])] for g in gv[1:]]
            rec = {}
            for space, arr in (("H", sH), ("E", sE)):
                blocks, ntot = entity_blocks(et, no, space)
                on_face = np.zeros(arr.shape[0], bool)
                for kind, idx, m0, n in blocks:
                    if kind == "v":
                        gk = ("v", v[idx]); inc = v[idx] in key
                    elif kind == "e":
                        a, b = E[idx]; gk = ("e", frozenset((v[a], v[b]))); inc = {v[a], v[b]} <= key
                    else:
                        gk = ("f", frozenset(v[i] for i in F[idx])); inc = gk[1] == key
                    if inc:
                        on_face[m0:m0 + n] = True
                        if space == "H":
                            rec[(space,) + gk] = arr[m0:m0 + n].copy()
                        else:
                            rec[(space,) + gk] = np.stack([arr[m0:m0 + n] @ t for t in tang], 1)
                rest = arr[:ntot][~on_face[:ntot]]
                if space == "H":
                    assert np.abs(rest).max(initial=0.0) < 1e-13
                    assert np.abs(arr[ntot:]).max(initial=0.0) < 1e-13          # bubbles vanish on the boundary
                else:
                    assert np.abs(np.stack([rest @ t for t in tang], 1)).max(initial=0.0) < 1e-13
                    assert np.abs(np.stack([arr[ntot:] @ t for t in tang], 1)).max(initial=0.0) < 1e-13
            vals.append(rec)
        a, b = vals
        assert a.keys() == b.keys()
        for k in a:
            assert a[k].shape == b[k].shape, k
            assert np.abs(a[k] - b[k]).max(initial=0.0) < 1e-13, (k, a[k], b[k])
        ets = sorted(int(m["etype"][e]) for e, _ in owners)
        checked["qq" if len(gv) == 4 and ets == [1, 1] else "tt" if len(gv) == 3 else "mixed"] += 1
    assert checked["qq"] > 0 and checked["tt"] > 0 and checked["mixed"] > 0, checked
