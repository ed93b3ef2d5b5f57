"""HP3D_TRACE=1 python tools/e2e_trace.py [B]: host-side phase times of hp3d_gpu_elem_batch on the headline workload."""
import sys, time, numpy as np
sys.path.insert(0, '.')
from hp3d_b200 import synth, _lib
from hp3d_b200.api import ElemEngine, pinned_empty
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
no, oe, of, x = synth.cube_mesh(B, 5)
n, m, ni, nb = synth.problem_sizes(4, 5)
dt = np.complex128
bufs = [pinned_empty((B, ni * ni), dt), pinned_empty((B, ni), dt), pinned_empty((B, nb * ni), dt), pinned_empty((B, nb), dt)]
out = dict(Aii=bufs[0].a, Bi=bufs[1].a, ASchur=bufs[2].a, BSchur=bufs[3].a)
eng = ElemEngine(4, omega=2 * np.pi)
L = _lib.lib()
for mc in (0, 32, 16):
    L.hp3d_gpu_set_chunk(mc)
    eng.elem_stc_batch(no, oe, of, x, out=out)
    for _ in range(2):
        t0 = time.perf_counter(); eng.elem_stc_batch(no, oe, of, x, out=out); t = time.perf_counter() - t0
        print(f"max_chunk {mc}: call {t*1e3:.1f} ms = {B/t:.0f} el/s", file=sys.stderr)
# no Schur factors (STORE_STC off): 5.8 MB instead of 13 MB per element
eng2 = ElemEngine(4, omega=2 * np.pi, store_schur=0)
L.hp3d_gpu_set_chunk(0)
eng2.elem_stc_batch(no, oe, of, x, out=out)
t0 = time.perf_counter(); eng2.elem_stc_batch(no, oe, of, x, out=out); t = time.perf_counter() - t0
print(f"no schur: call {t*1e3:.1f} ms = {B/t:.0f} el/s", file=sys.stderr)
r = eng.bench(no, oe, of, x, reps=2, lanes=4, max_chunk=64)
print("device-resident 4 lanes x 64:", 2 * B / r["ms_total"] * 1e3, file=sys.stderr)
