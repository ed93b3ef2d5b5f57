"""Measure the box's dense FP64 rates (cuBLAS DGEMM/ZGEMM via torch) -> the FP64 roofline denominator.
MEASURED_PEAKS.json (driver-written) has no FP64 entry; this script records one under gpurun_out/."""
import json, sys, torch
def best_ms(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize(); best = 1e30
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); e1.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
out = {"gpu": torch.cuda.get_device_name(0)}
N = 8192
a = torch.randn(N, N, dtype=torch.float64, device="cuda"); b = torch.randn(N, N, dtype=torch.float64, device="cuda")
ms = best_ms(lambda: torch.matmul(a, b)); out["dgemm_8192_tflops"] = 2 * N**3 / ms / 1e9
# sustained: back to back for ~3 s
import time
torch.cuda.synchronize(); t0 = time.time(); k = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
while time.time() - t0 < 3.0:
    for _ in range(5): torch.matmul(a, b)
    k += 5; torch.cuda.synchronize()
e1.record(); e1.synchronize(); out["dgemm_8192_tflops_sustained"] = k * 2 * N**3 / e0.elapsed_time(e1) / 1e9
del a, b
N = 4096
a = torch.randn(N, N, dtype=torch.complex128, device="cuda"); b = torch.randn(N, N, dtype=torch.complex128, device="cuda")
ms = best_ms(lambda: torch.matmul(a, b)); out["zgemm_4096_tflops"] = 8 * N**3 / ms / 1e9
print(json.dumps(out))
json.dump(out, open("gpurun_out/fp64_peak.json", "w"))
