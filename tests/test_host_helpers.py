"""Host-only entry points of the C ABI that need no GPU: the packed-Hermitian unpack the Fortran shim runs before copying Aii
into ALOC (hp3d_gpu_hermitian_unpack_batch), against LAPACK's packed-storage definition AP(i + (j-1)(2n-j)/2) = A(i,j), i >= j."""
import ctypes as C

import numpy as np
import pytest

from hp3d_b200 import _lib


@pytest.mark.parametrize("cplx", [0, 1])
@pytest.mark.parametrize("threads", [1, 0])
def test_hermitian_unpack(cplx, threads):
    L = _lib.lib()
    rng = np.random.default_rng(7 + cplx)
    dt = np.complex128 if cplx else np.float64
    nis = np.array([1, 5, 33, 64, 0, 70], np.int32)
    nmax = int(nis.max())
    AP = np.zeros((len(nis), nmax * (nmax + 1) // 2 + 3), dt)
    full = []
    for e, n in enumerate(nis):
        A = rng.normal(size=(n, n)) + (1j * rng.normal(size=(n, n)) if cplx else 0)
        A = A + A.conj().T
        k = 0
        for j in range(n):
            AP[e, k:k + n - j] = A[j:, j]
            k += n - j
        full.append(A)
    out = np.full((len(nis), nmax * nmax + 2), 777.0, dt)
    f = L.hp3d_gpu_hermitian_unpack_batch
    f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int]
    assert f(cplx, len(nis), nmax, nis.ctypes.data, AP.ctypes.data, AP.shape[1], out.ctypes.data, out.shape[1], threads) == 0
    for e, n in enumerate(nis):
        got = out[e, :n * n].reshape(n, n).T
        assert np.array_equal(got, full[e])
        assert (out[e, n * n:] == 777.0).all()   # nothing beyond the element's block is touched
    # uniform size without the per-element array
    n = 33
    out2 = np.zeros((2, n * n), dt)
    assert f(cplx, 2, n, None, AP[2:4].ctypes.data, AP.shape[1], out2.ctypes.data, n * n, threads) == 0
    assert np.array_equal(out2[0].reshape(n, n).T, full[2])
    assert f(cplx, 1, n, None, out2.ctypes.data, n * n, out2.ctypes.data, n * n, threads) != 0   # in place is refused
