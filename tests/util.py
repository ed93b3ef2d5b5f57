"""Shared helpers of the test-suite (synthetic element descriptors, ctypes plumbing)."""
import ctypes as C

import numpy as np

VERT = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
# face -> its four edges (element_data.F90:106-109, 0-based); [0],[2] along the first face axis
FACE_EDGES = [[0, 1, 2, 3], [4, 5, 6, 7], [0, 9, 4, 8], [1, 10, 5, 9], [2, 10, 6, 11], [3, 11, 7, 8]]
FACE_AXES = [(0, 1), (0, 1), (0, 2), (1, 2), (0, 2), (1, 2)]
EDGE_AXIS = [0, 1, 0, 1, 0, 1, 0, 1, 2, 2, 2, 2]
QSWAP = [0, 1, 0, 1, 1, 0, 1, 0]


def _p(a, t=C.c_int):
    return a.ctypes.data_as(C.POINTER(t))


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def uniform_order(p):
    return np.array([p] * 12 + [11 * p] * 6 + [111 * p], dtype=np.int32)


def random_signature(rng, pmax=4, uniform=False):
    """A valid (norder, norient_edge, norient_face): interior order per axis, then min-rule-like lower orders on
    faces and edges, face digits stored in the face's oriented frame (find_order.F90:41-58)."""
    if uniform:
        p = int(rng.integers(1, pmax + 1))
        return uniform_order(p), np.zeros(12, np.int32), np.zeros(6, np.int32)
    mid = rng.integers(1, pmax + 1, 3)
    norie = rng.integers(0, 2, 12).astype(np.int32)
    norif = rng.integers(0, 8, 6).astype(np.int32)
    norder = np.zeros(19, np.int32)
    for e in range(12):
        norder[e] = rng.integers(1, mid[EDGE_AXIS[e]] + 1)
    for f in range(6):
        a, b = FACE_AXES[f]
        pa, pb = rng.integers(1, mid[a] + 1), rng.integers(1, mid[b] + 1)
        if QSWAP[norif[f]]:
            pa, pb = pb, pa       # digits are kept in the face's own (oriented) frame
        norder[12 + f] = 10 * pa + pb
    norder[18] = 100 * mid[0] + 10 * mid[1] + mid[2]
    return norder, norie, norif


def dof_map(L, space, norder, norie, norif):
    norder, norie, norif = i32(norder), i32(norie), i32(norif)
    n = L.hp3d_gpu_dof_map(space, _p(norder), _p(norie), _p(norif), 0, None, None, None)
    assert n >= 0
    fam = np.zeros(n, np.int32); idx = np.zeros((n, 3), np.int32); sgn = np.zeros(n, np.int32)
    m = L.hp3d_gpu_dof_map(space, _p(norder), _p(norie), _p(norif), n, _p(fam), _p(idx), _p(sgn))
    assert m == n
    return fam, idx, sgn


def tables_1d(L, p, nq):
    x = np.zeros(nq); w = np.zeros(nq); H = np.zeros((p + 1, nq)); dH = np.zeros((p + 1, nq)); Q = np.zeros((p, nq))
    rc = L.hp3d_gpu_tables_1d(p, nq, _p(x, C.c_double), _p(w, C.c_double), _p(H, C.c_double), _p(dH, C.c_double), _p(Q, C.c_double))
    assert rc == 0
    return x, w, H, dH, Q


def hexa_xnod(nH, h=0.25, origin=(0.0, 0.0, 0.0), jitter=0.0, seed=0, curved=0.0, rng=None):
    """Geometry dofs (nrdofH,3): trilinear vertex coordinates (+ vertex jitter), optional random higher-order dofs."""
    rng = rng or np.random.default_rng(seed)
    X = np.zeros((nH, 3))
    X[:8] = np.asarray(origin) + h * VERT + rng.uniform(-jitter * h, jitter * h, (8, 3))
    if curved:
        X[8:] = rng.uniform(-curved * h, curved * h, (nH - 8, 3))
    return X


def complex_W(W, d):
    """The complex dense-phase input [G ; B^H ; load] from hp3d_gpu_integrate_debug.  Two planes: the general complex layout.
    One plane: the REAL-STRUCTURED storage of lossless ultraweak Maxwell (forms.hpp): W~[r,c] = kappa conj(i^pr) i^pc W[r,c],
    kappa = 1 on Gram rows / i on trial rows, test phases 0 (F) / 1 (G), trial phases 0 (E-type) / 1 (H-type), the load as two
    real rows (Re, Im) -- undone here so that callers can compare with the oracle's complex matrices.  The complex load is
    returned in the LAST padded interface row in both cases."""
    if W.shape[0] == 2:
        return W[0] + 1j * W[1]
    n, nb, ni, np_, nbp, nip = d["n"], d["nb"], d["ni"], d["np"], d["nbp"], d["nip"]
    nEE = n // 2
    Wt = W[0]
    R = Wt.shape[0]
    sig = np.ones(np_, complex); sig[nEE:n] = 1j
    rowph = np.ones(R, complex); kap = np.ones(R, complex)
    rowph[:np_] = sig
    kap[np_:] = 1j
    b = np.arange(nb); rowph[np_ + b] = np.where(b % 6 >= 3, 1j, 1.0)
    i = np.arange(ni); rowph[np_ + nbp + i] = np.where(i % 2 == 1, 1j, 1.0)
    # W = conj(kappa) * i^pr * conj(i^pc) * W~
    Wc = (np.conj(kap) * rowph)[:, None] * Wt * np.conj(sig)[None, :]
    l0 = np_ + nbp + nip - 2
    lam = Wt[l0] + 1j * Wt[l0 + 1]
    Wc[l0] = 0.0
    Wc[l0 + 1] = lam * np.conj(sig)
    return Wc
