import sys, time, numpy as np
sys.path.insert(0, '.')
from hp3d_b200 import synth, _lib
from hp3d_b200.api import ElemEngine, pinned_empty
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
no, oe, of, x = synth.cube_mesh(B, 5)
n, m, ni, nb = synth.problem_sizes(4, 5)
L = _lib.lib()
dt = np.complex128
bufs = [pinned_empty((B, ni * ni), dt), pinned_empty((B, ni), dt), pinned_empty((B, nb * ni), dt), pinned_empty((B, nb), dt)]
out = dict(Aii=bufs[0].a, Bi=bufs[1].a, ASchur=bufs[2].a, BSchur=bufs[3].a)
for schur in (1, 0):
    eng = ElemEngine(4, omega=2 * np.pi, store_schur=schur)
    for mc in (128, 64, 32):
        L.hp3d_gpu_set_chunk(mc)
        eng.elem_stc_batch(no, oe, of, x, out=out)
        t0 = time.perf_counter()
        for _ in range(2): eng.elem_stc_batch(no, oe, of, x, out=out)
        t = time.perf_counter() - t0
        print(f"e2e pinned store_schur={schur} max_chunk={mc:4d}: {2*B/t:8.1f} el/s  ({t/2*1e3:.1f} ms/call)")
    eng.close()
