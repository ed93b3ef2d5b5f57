"""(Measurement script, not a test; it lives under tests/ because it times the oracle beside the device path.)
Throughput of the batched H1 projection-based interpolation (hp3d_gpu_pbi_h1_batch) through the C ABI with host buffers,
beside the oracle's OpenMP element loop on the host cores (update_gdof.F90:409-435 shape).  Prints one JSON line.
usage: python tests/bench_pbi.py [p] [nel] [path of an alternative libhp3d_gpu.so]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hp3d_b200 import _lib, api, synth  # noqa: E402

if len(sys.argv) > 3:   # an alternative build of the library (kernel experiments)
    _lib.LIB_PATH = os.path.abspath(sys.argv[3])

p = int(sys.argv[1]) if len(sys.argv) > 1 else 5
nel = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
rng = np.random.default_rng(0)
no = np.tile(synth.uniform_order(p), (nel, 1)); noe = np.zeros((nel, 12), np.int32); nof = np.zeros((nel, 6), np.int32)
M = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
etav = M[None] * 0.5 + rng.uniform(0, 0.5, (nel, 1, 3)) + rng.uniform(-0.02, 0.02, (nel, 8, 3))
pts = api.pbi_points(no[:1], noe[:1], nof[:1])
npts, nH = int(pts["npts"][0]), int(pts["nrdofH"][0])
xi = pts["xi"][0, :npts]
S = np.array([[(x[0] if m[0] else 1 - x[0]) * (x[1] if m[1] else 1 - x[1]) * (x[2] if m[2] else 1 - x[2]) for m in M] for x in xi])
eta = np.einsum("lv,evc->elc", S, etav)


def g(eta):
    x, y, z = eta[..., 0], eta[..., 1], eta[..., 2]
    v = np.stack([x + 0.1 * np.sin(2 * y) * z, y + 0.05 * x * x, z + 0.1 * np.cos(x + y)], -1)
    o, zz = np.ones_like(x), np.zeros_like(x)
    d = np.stack([np.stack([o, 0.1 * x, -0.1 * np.sin(x + y)], -1), np.stack([0.2 * np.cos(2 * y) * z, o, -0.1 * np.sin(x + y)], -1),
                  np.stack([0.1 * np.sin(2 * y), zz, o], -1)], -2)   # d[..., i, c] = d g_c / d eta_i
    return v, d


fv = g(etav)[0]; fg = g(eta)[1]
api.pbi_h1_batch(no[:8], noe[:8], nof[:8], etav[:8], fv[:8], fg[:8])   # warm-up: signature tables, context
ts = []
for _ in range(5):
    t = time.perf_counter(); res = api.pbi_h1_batch(no, noe, nof, etav, fv, fg); ts.append(time.perf_counter() - t)
assert not res["info"].any()
# the same call on page-locked caller arrays (hp3d_gpu_host_alloc): the copies become asynchronous DMA transfers
L = _lib.lib()
tp = []
try:
    keep = []

    def pinned(a):
        h = api.pinned_empty(a.shape, a.dtype); h.a[...] = a; keep.append(h); return h.a
    pe, pfv, pfg = pinned(etav), pinned(fv), pinned(fg)
    pd = pinned(np.zeros((nel, nH, 3))); info = np.zeros(nel, np.int32)
    f = L.hp3d_gpu_pbi_h1_batch
    for _ in range(5):
        t = time.perf_counter()
        rc = f(nel, None, api._ptr(no), api._ptr(noe), api._ptr(nof), 0, 9, api._ptr(pe), 3, api._ptr(pfv), api._ptr(pfg), int(np.prod(pfg.shape[1:])),
               None, api._ptr(pd), 3 * nH, api._ptr(info))
        tp.append(time.perf_counter() - t)
        assert rc == 0 and not info.any()
    assert np.array_equal(pd, res["dof"])
except Exception as ex:
    tp = []
    print("pinned leg failed:", repr(ex), file=sys.stderr)
out = {"what": "hp3d_gpu_pbi_h1_batch (update_gdof, 3 components), hexa p=%d, host buffers" % p, "elements": nel, "nrdofH": nH, "points_per_element": npts,
       "e2e_elements_per_s": nel / min(ts), "ms_per_call": 1e3 * min(ts),
       "e2e_pinned_elements_per_s": (nel / min(tp)) if tp else None, "lib": os.path.basename(os.path.dirname(_lib.LIB_PATH)) + "/" + os.path.basename(_lib.LIB_PATH)}
# H(curl) / H(div) Dirichlet interpolation of the same elements (two real components; the systems do not depend on the data)
for name, pts_f, run in (("hcurl", api.pbi_hcurl_points, lambda a, b: api.pbi_hcurl_batch(no, noe, nof, etav, a, b)),
                         ("hdiv", api.pbi_hdiv_points, lambda a, b: api.pbi_hdiv_batch(no, noe, nof, etav, a))):
    n2 = int(pts_f(no[:1], noe[:1], nof[:1])["npts"][0])
    a = rng.standard_normal((nel, n2, 3, 2)); b = rng.standard_normal((nel, n2, 3, 2))
    run(a, b)
    tt = []
    for _ in range(3):
        t = time.perf_counter(); r2 = run(a, b); tt.append(time.perf_counter() - t)
    assert not r2["info"].any()
    out[name + "_e2e_elements_per_s"] = nel / min(tt)
try:
    from oracle import oracle as O
    L = O.lib(); O.set_maxp(9)
    ns = min(nel, 256); nthr = os.cpu_count() or 1
    dof = np.zeros((ns, nH, 3))
    f = L.orc_pbi_batch_sample
    f.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_int, C.c_void_p, C.c_long, C.c_int]
    ev = np.ascontiguousarray(etav[:ns])
    t = time.perf_counter()
    bad = f(ns, None, no[:ns].ctypes.data, noe[:ns].ctypes.data, nof[:ns].ctypes.data, ev.ctypes.data, 0, 9, dof.ctypes.data, 3 * nH, nthr)
    dt = time.perf_counter() - t
    out["cpu_port"] = {"elements_per_s": ns / dt, "cores": nthr, "sample": "%d elements, OpenMP over elements" % ns, "bad": int(bad)}
    out["max_rel_diff_vs_oracle"] = float(np.abs(dof - res["dof"][:ns]).max() / np.abs(dof).max())
except Exception as ex:   # the oracle is optional here
    out["cpu_port"] = {"unavailable": repr(ex)}
print(json.dumps(out))
