"""D2H bandwidth into cudaMallocHost memory (the library's hp3d_gpu_host_alloc) vs copy size and number of streams."""
import ctypes as C, sys, time
sys.path.insert(0, '.')
import numpy as np
from hp3d_b200 import _lib
L = _lib.lib(); _lib.check(L.hp3d_gpu_init(0))
L.hp3d_gpu_host_alloc.restype = C.c_void_p; L.hp3d_gpu_host_alloc.argtypes = [C.c_longlong]
rt = C.CDLL("libcudart.so.12")
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaStreamCreate.argtypes = [C.POINTER(C.c_void_p)]
rt.cudaStreamSynchronize.argtypes = [C.c_void_p]
G = 1 << 30
N = 4 * G
d = C.c_void_p(); assert rt.cudaMalloc(C.byref(d), N) == 0
h = L.hp3d_gpu_host_alloc(N); assert h
st = [C.c_void_p() for _ in range(2)]
for s in st: rt.cudaStreamCreate(C.byref(s))
for rep in range(2):
    for size in (8 << 20, 64 << 20, 368 << 20, 1 << 30):
        n = N // size
        t0 = time.perf_counter()
        for i in range(n): rt.cudaMemcpyAsync(C.c_void_p(h + i * size), C.c_void_p(d.value + i * size), size, 2, st[0])
        rt.cudaStreamSynchronize(st[0]); t = time.perf_counter() - t0
        print(f"D2H {size>>20:5d} MB x {n:4d} one stream: {N/t/1e9:6.1f} GB/s")
    t0 = time.perf_counter()
    size = 64 << 20; n = N // size
    for i in range(n): rt.cudaMemcpyAsync(C.c_void_p(h + i * size), C.c_void_p(d.value + i * size), size, 2, st[i & 1])
    for s in st: rt.cudaStreamSynchronize(s)
    t = time.perf_counter() - t0
    print(f"D2H 64 MB two streams: {N/t/1e9:6.1f} GB/s")
import subprocess
print(subprocess.run("nvidia-smi topo -m | head -8; lscpu | grep -i 'numa\\|socket\\|model name' | head; nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv", shell=True, capture_output=True, text=True).stdout)
