"""Synthetic element descriptors for benchmarks and tests.

The inputs are exactly what the reference's host code hands to `elem` for each element of a subdomain:
`norder(19)` (find_order), `norient_edge(12)` / `norient_face(6)` (find_orient) and the geometry dofs `xnod(3,nrdofH)`
(nodcor).  *Uniform*: the unit cube split into N^3 congruent hexahedra of order p, all orientations 0.  *Perturbed*: the
same mesh with every vertex jittered by U(-0.15h, 0.15h) (seed 12345) so that the Jacobian is neither constant nor
diagonal and no two elements are congruent.
"""
import numpy as np

VERT = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], dtype=np.int64)


def uniform_order(p):
    """norder(19) of an isotropic order-p brick: 12 edges p, 6 faces 10p+p, middle node 100p+10p+p."""
    return np.array([p] * 12 + [11 * p] * 6 + [111 * p], dtype=np.int32)


def nrdof_h1(p):
    return (p + 1) ** 3


def cube_mesh(nel, p, jitter=0.15, seed=12345, first=0, total=None):
    """Descriptors of elements first..first+nel-1 (lexicographic, x fastest) of the N^3 cube mesh, N = ceil(cbrt(total))
    (total defaults to first+nel; ranks of one job pass the job's element count so that they all see the same mesh).

    Returns norder (nel,19), norient_edge (nel,12), norient_face (nel,6), xnod (nel,(p+1)^3,3)."""
    end = first + nel
    total = max(int(total or 0), end)
    N = max(1, int(np.ceil(total ** (1.0 / 3.0) - 1e-9)))
    while N ** 3 < total:
        N += 1
    h = 1.0 / N
    rng = np.random.default_rng(seed)
    # one jitter vector per mesh vertex (shared by the neighbouring elements); boundary vertices stay on the cube
    g = rng.uniform(-jitter * h, jitter * h, (N + 1, N + 1, N + 1, 3))
    for d in range(3):
        idx = [slice(None)] * 3
        for side in (0, N):
            idx[d] = side
            g[tuple(idx) + (d,)] = 0.0
            idx[d] = slice(None)
    ids = np.arange(first, end)
    ix, iy, iz = ids % N, (ids // N) % N, ids // (N * N)
    xnod = np.zeros((nel, nrdof_h1(p), 3))
    for v in range(8):
        vx, vy, vz = ix + VERT[v, 0], iy + VERT[v, 1], iz + VERT[v, 2]
        xnod[:, v, 0] = vx * h
        xnod[:, v, 1] = vy * h
        xnod[:, v, 2] = vz * h
        xnod[:, v, :] += g[vx, vy, vz]
    norder = np.tile(uniform_order(p), (nel, 1))
    return norder, np.zeros((nel, 12), np.int32), np.zeros((nel, 6), np.int32), xnod


def dense_flops(kind, n, m, ni, nb, nrhs=1):
    """ALGORITHMIC real flops of the dense phase of one element:
    c*[n^3/3 + n^2(m+1) + n(m+1)^2] (DPG normal equations) + c*[nb^3/3 (Cholesky; 2nb^3/3 for LU) + 2nb^2(ni+r) + 2 ni nb (ni+r)]."""
    c = 4.0 if kind >= 3 else 1.0
    dpg = kind in (2, 4)
    lu = kind in (1, 3)
    f = 0.0
    if dpg:
        f += n ** 3 / 3.0 + n ** 2 * (m + 1.0) + n * (m + 1.0) ** 2
    f += (2.0 if lu else 1.0) * nb ** 3 / 3.0 + 2.0 * nb ** 2 * (ni + nrhs) + 2.0 * ni * nb * (ni + nrhs)
    return c * f


def dense_flops_real_form(n, m, ni, nb):
    """Real flops of the dense phase of one lossless ultraweak-Maxwell element in its REAL form (A = T A~ T^H, forms.hpp):
    the same factor / solve / rank-k sequence on real matrices (c = 1), with the complex load as TWO real right-hand sides."""
    r = 2.0
    return (n ** 3 / 3.0 + n ** 2 * (m + r) + n * (m + r) ** 2) + (nb ** 3 / 3.0 + 2.0 * nb ** 2 * (ni + r) + 2.0 * ni * nb * (ni + r))


def problem_sizes(kind, p, dp=1):
    """(ntest, ntrial, ni, nb) for an isotropic order-p brick."""
    H, E, V, Q = (p + 1) ** 3, 3 * p * (p + 1) ** 2, 3 * p * p * (p + 1), p ** 3
    bH, bE, bV = (p - 1) ** 3, 3 * p * (p - 1) ** 2, 3 * p * p * (p - 1)
    pe = p + dp
    if kind == 1:
        return 0, H, H - bH, bH
    if kind == 2:
        return (pe + 1) ** 3, H + (V - bV), (H - bH) + (V - bV), bH
    if kind == 3:
        return 0, E, E - bE, bE
    if kind == 4:
        return 2 * 3 * pe * (pe + 1) ** 2, 2 * (E - bE) + 6 * Q, 2 * (E - bE), 6 * Q
    raise ValueError(kind)


# ----------------------------------------------------------------------------------------------------------------------
# hp-refined mixed hexahedron / prism mesh (BASELINE.json configs[4])
# master-element topology, 0-based (src/modules/element_data.F90:25-35, 62-70, 85-93)
MDLB, MDLP = 1, 3
BRICK_EDGE = [(0, 1), (1, 2), (3, 2), (0, 3), (4, 5), (5, 6), (7, 6), (4, 7), (0, 4), (1, 5), (2, 6), (3, 7)]
BRICK_FACE = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (1, 2, 6, 5), (3, 2, 6, 7), (0, 3, 7, 4)]
PRISM_VERT = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1]], dtype=np.float64)
PRISM_EDGE = [(0, 1), (1, 2), (0, 2), (3, 4), (4, 5), (3, 5), (0, 3), (1, 4), (2, 5)]
PRISM_FACE = [(0, 1, 2), (3, 4, 5), (0, 1, 4, 3), (1, 2, 5, 4), (0, 2, 5, 3)]
TRI_ORIENT = {(0, 1, 2): 0, (1, 2, 0): 1, (2, 0, 1): 2, (0, 2, 1): 3, (1, 0, 2): 4, (2, 1, 0): 5}   # Orient.F90:119-160


def quad_orientation(gids):
    """Orientation code 0..7 (Orient.F90:39-99) of a quad face whose element-local cyclic vertices carry the global ids
    `gids`: the face's own frame has its origin at the vertex with the smallest id and its first axis towards the
    neighbour with the smaller id.  Codes 0-3: origin at local vertex 0..3, first axis towards the NEXT vertex of the
    cycle; 4-7: towards the PREVIOUS one."""
    o = int(np.argmin(gids))
    nxt, prv = gids[(o + 1) % 4], gids[(o - 1) % 4]
    return o if nxt < prv else 4 + o


def tri_orientation(gids):
    """Orientation code 0..5 (Orient.F90:119-160): global vertex k of the face = local vertex perm[k], ids ascending."""
    return TRI_ORIENT[tuple(int(i) for i in np.argsort(gids))]


def nrdof_h1_prism(norder):
    n = 6 + sum(int(norder[e]) - 1 for e in range(9))
    for f in (9, 10):
        n += (norder[f] - 1) * (norder[f] - 2) // 2
    for f in (11, 12, 13):
        n += (norder[f] // 10 - 1) * (norder[f] % 10 - 1)
    p, pz = norder[14] // 10, norder[14] % 10
    return int(n + (p - 1) * (p - 2) // 2 * (pz - 1))


def nrdof_h1_brick(norder):
    n = 8 + sum(int(norder[e]) - 1 for e in range(12))
    for f in range(12, 18):
        n += (norder[f] // 10 - 1) * (norder[f] % 10 - 1)
    m = int(norder[18])
    return int(n + (m // 100 - 1) * ((m // 10) % 10 - 1) * (m % 10 - 1))


def hp_mesh(N, prism_frac=0.3, pmin=2, pmax=7, seed_p=2024, seed_g=7, jitter=0.0, seed_x=12345, rotate_local=True):
    """A conforming mixed mesh of the unit cube: N^3 cells; whole (i,j) columns of cells are split into two prisms each
    (along the vertical diagonal plane) so that about `prism_frac` of the ELEMENTS are prisms, the others are hexahedra.  Every element draws an
    isotropic order uniformly from {pmin..pmax} (seed_p); edges and faces get the MINIMUM order of the adjacent elements
    (hp3D's min rule); edge and face orientations follow from a random global vertex numbering (seed_g): an edge runs from
    the smaller to the larger vertex id, a face's frame starts at its smallest id (see quad_orientation / tri_orientation).
    With `rotate_local` the element-local vertex numbering of every element is rotated about its vertical axis by a random
    multiple of 90 (brick) / 120 (prism) degrees, so that neighbours see a shared face with different local numberings.

    Returns dict(etype (nel,), norder (nel,19), norient_edge (nel,12), norient_face (nel,6), xnod (nel,nHmax,3),
    nrdofH (nel,), verts (nel,8) global vertex ids in element-local order (-1 padded), p (nel,))."""
    rng_p, rng_g, rng_x = np.random.default_rng(seed_p), np.random.default_rng(seed_g), np.random.default_rng(seed_x)
    n1 = N + 1
    gid = rng_g.permutation(n1 ** 3)
    h = 1.0 / N
    vid = lambda i, j, k: i + n1 * (j + n1 * k)   # noqa: E731
    coords = np.zeros((n1 ** 3, 3))
    for k in range(n1):
        for j in range(n1):
            for i in range(n1):
                c = np.array([i, j, k]) * h
                if jitter:
                    d = rng_x.uniform(-jitter * h, jitter * h, 3)
                    for a, idx in enumerate((i, j, k)):
                        if idx in (0, N):
                            d[a] = 0.0
                    c = c + d
                coords[vid(i, j, k)] = c
    split = rng_p.random((N, N)) < prism_frac / (2.0 - prism_frac)   # a split cell yields two prisms
    elems = []   # (etype, [vertex ids in element-local order])
    for k in range(N):
        for j in range(N):
            for i in range(N):
                c = [vid(i, j, k), vid(i + 1, j, k), vid(i + 1, j + 1, k), vid(i, j + 1, k),
                     vid(i, j, k + 1), vid(i + 1, j, k + 1), vid(i + 1, j + 1, k + 1), vid(i, j + 1, k + 1)]
                if split[i, j]:
                    elems.append((MDLP, [c[0], c[1], c[2], c[4], c[5], c[6]]))
                    elems.append((MDLP, [c[0], c[2], c[3], c[4], c[6], c[7]]))
                else:
                    elems.append((MDLB, c))
    nel = len(elems)
    if rotate_local:
        for e, (et, v) in enumerate(elems):
            n = 4 if et == MDLB else 3
            r = int(rng_g.integers(0, n))
            elems[e] = (et, [v[(i + r) % n] for i in range(n)] + [v[n + (i + r) % n] for i in range(n)])
    p = rng_p.integers(pmin, pmax + 1, nel)
    # min rule over the elements adjacent to each edge / face
    emin, fmin = {}, {}
    for e, (et, v) in enumerate(elems):
        E, F = (BRICK_EDGE, BRICK_FACE) if et == MDLB else (PRISM_EDGE, PRISM_FACE)
        for a, b in E:
            key = frozenset((v[a], v[b]))
            emin[key] = min(emin.get(key, 99), int(p[e]))
        for f in F:
            key = frozenset(v[i] for i in f)
            fmin[key] = min(fmin.get(key, 99), int(p[e]))
    etype = np.zeros(nel, np.int32); norder = np.zeros((nel, 19), np.int32)
    noe = np.zeros((nel, 12), np.int32); nof = np.zeros((nel, 6), np.int32)
    verts = -np.ones((nel, 8), np.int64); nH = np.zeros(nel, np.int32)
    for e, (et, v) in enumerate(elems):
        etype[e] = et
        verts[e, :len(v)] = v
        E, F = (BRICK_EDGE, BRICK_FACE) if et == MDLB else (PRISM_EDGE, PRISM_FACE)
        g = gid[np.array(v)]
        for ie, (a, b) in enumerate(E):
            norder[e, ie] = emin[frozenset((v[a], v[b]))]
            noe[e, ie] = 0 if g[a] < g[b] else 1
        for jf, f in enumerate(F):
            pf = fmin[frozenset(v[i] for i in f)]
            if len(f) == 3:
                norder[e, len(E) + jf] = pf
                nof[e, jf] = tri_orientation(g[list(f)])
            else:
                norder[e, len(E) + jf] = 11 * pf
                nof[e, jf] = quad_orientation(g[list(f)])
        if et == MDLB:
            norder[e, 18] = 111 * int(p[e]); nH[e] = nrdof_h1_brick(norder[e])
        else:
            norder[e, 14] = 11 * int(p[e]); nH[e] = nrdof_h1_prism(norder[e])
    xnod = np.zeros((nel, int(nH.max()), 3))
    for e, (et, v) in enumerate(elems):
        xnod[e, :len(v)] = coords[np.array(v)]
    return dict(etype=etype, norder=norder, norient_edge=noe, norient_face=nof, xnod=xnod, nrdofH=nH, verts=verts, p=p,
                gid=gid, coords=coords)


def synthetic_constraints(kind, ni, nel, frac_hanging=0.1, frac_boundary=0.2, seed=99, nacdim=4):
    """Per-element input of celem_systemI's transform/compression step for a synthetic mesh: most elements
    are regular (every element dof is its own modified dof), `frac_hanging` of them have 30 % of their dofs constrained to
    2..nacdim parents (1-irregular refinement), `frac_boundary` of them carry Dirichlet data on 15 % of their dofs.
    Single-family problems only (kinds 1, 3, 4).  Returns a list of dicts for ElemEngine.celem_batch."""
    from . import api
    if kind == 2:
        raise ValueError("synthetic_constraints: single-family problems only")
    ph = api.physics_default(kind)
    fam = ph.dtype[0]
    ncomp = ph.ncomp[0]
    nk = ni // ncomp
    cplx = kind >= 3
    rng = np.random.default_rng(seed)
    out = []
    gdof = 1
    for e in range(nel):
        hang = rng.random() < frac_hanging
        nm = nk + (nk // 10 if hang else 0)
        rc = np.ones(nk, np.int32); na = np.zeros((nk, nacdim), np.int32); co = np.zeros((nk, nacdim))
        na[:, 0] = np.arange(1, nk + 1); co[:, 0] = 1.0
        if hang:
            for k in np.flatnonzero(rng.random(nk) < 0.3):
                m = int(rng.integers(2, nacdim + 1))
                rc[k] = m; na[k, :m] = rng.choice(nm, m, replace=False) + 1; co[k, :m] = 1.0 / m
        nrdofl = [0, 0, 0]; nrdofm_f = [0, 0, 0]
        nrdofl[fam] = nk; nrdofm_f[fam] = nm * ncomp
        z = np.zeros((0, nacdim))
        nrcon = [np.zeros(0, np.int32)] * 3; nac = [z.astype(np.int32)] * 3; con = [z] * 3
        nrcon[fam], nac[fam], con[fam] = rc, na, co
        cptr, cidx, cval = api.celem_pack(ph, nrdofl, nrcon, nac, con, nrdofm_f)
        nrdofm = nm * ncomp
        idbc = np.zeros(nrdofm, np.int32)
        if rng.random() < frac_boundary:
            idbc[rng.random(nrdofm) < 0.15] = 1
        zd = np.where(idbc == 1, rng.standard_normal(nrdofm) + (1j * rng.standard_normal(nrdofm) if cplx else 0), 0)
        zd = zd.astype(np.complex128 if cplx else np.float64)
        nextract = (np.flatnonzero(idbc == 0)[::-1] + 1).astype(np.int32)
        lcon = (gdof + np.arange(len(nextract))).astype(np.int32)
        gdof += len(nextract) // 2          # neighbours share dofs
        out.append(dict(nrdofl=nrdofl, nrcon=nrcon, nac=nac, constr=con, nrdofm_f=nrdofm_f, idbc=idbc, zdofd=zd, nextract=nextract, lcon=lcon,
                        cptr=cptr, cidx=cidx, cval=cval))
    return out
