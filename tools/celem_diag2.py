import sys, time, os, numpy as np
sys.path.insert(0, '.')
from hp3d_b200 import synth
from hp3d_b200.api import ElemEngine, pinned_empty
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = ElemEngine(4, omega=2 * np.pi)
no, oe, of, x = synth.cube_mesh(B, 5)
n, m, ni, nb = synth.problem_sizes(4, 5)
dt = np.complex128
AS = pinned_empty((B, nb * ni), dt); BS = pinned_empty((B, nb), dt)
cons = synth.synthetic_constraints(4, ni, B)
pk = eng.pack_constraints(cons, 2, True)
nz, nx = int(pk["aptr"][-1]), int(pk["xptr"][-1])
za = pinned_empty((nz,), dt); zb = pinned_empty((nx,), dt); irn = pinned_empty((nz,), np.int32); jcn = pinned_empty((nz,), np.int32)
out = dict(zastif=za.a, zbload=zb.a, ASchur=AS.a, BSchur=BS.a, irn=irn.a, jcn=jcn.a)
for i in range(3):
    t0 = time.perf_counter()
    eng.celem_batch(no, oe, of, x, None, isym_flag=2, want_coo=True, want_schur=True, out=out, packed=pk)
    print("call ms", (time.perf_counter() - t0) * 1e3, file=sys.stderr)
