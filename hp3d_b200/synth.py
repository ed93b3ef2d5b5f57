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
