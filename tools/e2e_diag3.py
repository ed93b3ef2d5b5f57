import sys, time, numpy as np
sys.path.insert(0, '.')
from hp3d_b200 import synth, _lib
from hp3d_b200.api import ElemEngine, pinned_empty
B = 256
no, oe, of, x = synth.cube_mesh(B, 5)
n, m, ni, nb = synth.problem_sizes(4, 5)
L = _lib.lib()
dt = np.complex128
bufs = [pinned_empty((B, ni * ni), dt), pinned_empty((B, ni), dt), pinned_empty((B, nb * ni), dt), pinned_empty((B, nb), dt)]
out = dict(Aii=bufs[0].a, Bi=bufs[1].a, ASchur=bufs[2].a, BSchur=bufs[3].a)
eng = ElemEngine(4, omega=2 * np.pi)
L.hp3d_gpu_set_chunk(32)
eng.elem_stc_batch(no, oe, of, x, out=out)
print("---- timed call", file=sys.stderr)
t0 = time.perf_counter(); eng.elem_stc_batch(no, oe, of, x, out=out); print("call ms", (time.perf_counter() - t0) * 1e3)
