"""Diagnostics of the end-to-end pipeline: chunk size sweep, device-only vs host I/O."""
import sys, time, numpy as np
sys.path.insert(0, '.')
from hp3d_b200 import synth, _lib
from hp3d_b200.api import ElemEngine, pinned_empty
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = ElemEngine(4, omega=2 * np.pi)
no, oe, of, x = synth.cube_mesh(B, 5)
n, m, ni, nb = synth.problem_sizes(4, 5)
for _ in range(2): eng.bench(no, oe, of, x, reps=1, lanes=2)
for mc in (0, 128, 64, 32):
    for lanes in (1, 2):
        r = eng.bench(no, oe, of, x, reps=2, max_chunk=mc, lanes=lanes)
        print(f"device-only max_chunk={mc:4d} lanes={lanes}: {2*B/r['ms_total']*1e3:8.1f} el/s")
dt = np.complex128
bufs = [pinned_empty((B, ni * ni), dt), pinned_empty((B, ni), dt), pinned_empty((B, nb * ni), dt), pinned_empty((B, nb), dt)]
out = dict(Aii=bufs[0].a, Bi=bufs[1].a, ASchur=bufs[2].a, BSchur=bufs[3].a)
L = _lib.lib()
for mc in (0, 128, 64, 32, 16):
    L.hp3d_gpu_set_chunk(mc)
    eng.elem_stc_batch(no, oe, of, x, out=out)
    t0 = time.perf_counter()
    for _ in range(2): eng.elem_stc_batch(no, oe, of, x, out=out)
    t = time.perf_counter() - t0
    print(f"e2e pinned   max_chunk={mc:4d}: {2*B/t:8.1f} el/s")
# D2H bandwidth of one big pinned copy
import torch
d = torch.empty(1 << 30, dtype=torch.uint8, device='cuda'); h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); print("D2H GB/s", 1.0737 / (time.perf_counter() - t0))
