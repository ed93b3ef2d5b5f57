"""Where does hp3d_gpu_celem_batch spend its time?  Variants of the call on the p=5 ultraweak workload (one B200)."""
import sys, time, numpy as np
sys.path.insert(0, '.')
from hp3d_b200 import synth
from hp3d_b200.api import ElemEngine, pinned_empty
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = ElemEngine(4, omega=2 * np.pi)
no, oe, of, x = synth.cube_mesh(B, 5)
n, m, ni, nb = synth.problem_sizes(4, 5)
dt = np.complex128
AS = pinned_empty((B, nb * ni), dt); BS = pinned_empty((B, nb), dt); A = pinned_empty((B, ni * ni), dt); Bi = pinned_empty((B, ni), dt)
outE = dict(Aii=A.a, Bi=Bi.a, ASchur=AS.a, BSchur=BS.a)
def timeit(f, reps=2):
    f(); t0 = time.perf_counter()
    for _ in range(reps): f()
    return B * reps / (time.perf_counter() - t0)
print(f"elem_stc_batch (Aii+Schur)      : {timeit(lambda: eng.elem_stc_batch(no, oe, of, x, out=outE)):8.1f} el/s")
for name, fh, fb in (("regular", 0.0, 0.0), ("10% hanging, 20% Dirichlet", 0.1, 0.2)):
    cons = synth.synthetic_constraints(4, ni, B, frac_hanging=fh, frac_boundary=fb)
    for coo, schur in ((False, False), (True, False), (False, True), (True, True)):
        pk = eng.pack_constraints(cons, 2, coo)
        nz, nx = int(pk["aptr"][-1]), int(pk["xptr"][-1])
        za = pinned_empty((nz,), dt); zb = pinned_empty((nx,), dt)
        out = dict(zastif=za.a, zbload=zb.a, ASchur=AS.a, BSchur=BS.a)
        if coo:
            irn = pinned_empty((nz,), np.int32); jcn = pinned_empty((nz,), np.int32); out.update(irn=irn.a, jcn=jcn.a)
        r = timeit(lambda: eng.celem_batch(no, oe, of, x, None, isym_flag=2, want_coo=coo, want_schur=schur, out=out, packed=pk))
        print(f"celem_batch {name:28s} coo={coo!s:5} schur={schur!s:5}: {r:8.1f} el/s")
        za.free(); zb.free()
        if coo: irn.free(); jcn.free()
import torch
d = torch.empty(1 << 30, dtype=torch.uint8, device='cuda'); h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); print("D2H GB/s", 1.0737 / (time.perf_counter() - t0))
