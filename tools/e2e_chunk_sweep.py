import sys, time, numpy as np
sys.path.insert(0, '.')
from hp3d_b200 import synth, _lib
from hp3d_b200.api import ElemEngine, pinned_empty
Be=1024
norder, noe, nof, xnod = synth.cube_mesh(Be, 5)
eng = ElemEngine(4, omega=2*np.pi, aii_packed=1)
ntest, ntrial, ni, nb = synth.problem_sizes(4, 5)
bufs=[pinned_empty((Be, ni*(ni+1)//2), np.complex128), pinned_empty((Be, ni), np.complex128)]
out=dict(Aii=bufs[0].a, Bi=bufs[1].a)
xs=pinned_empty(xnod.shape, np.float64); xs.a[...]=xnod
cl=eng.cloc_create()
for mc in (0, 96, 128, 192, 256):
    _lib.check(eng.L.hp3d_gpu_set_chunk(mc))
    for _ in range(2): eng.elem_stc_batch_cloc(cl, norder, noe, nof, xs.a, out=out)
    t0=time.perf_counter()
    for _ in range(3): eng.elem_stc_batch_cloc(cl, norder, noe, nof, xs.a, out=out)
    print(mc, 3*Be/(time.perf_counter()-t0))
