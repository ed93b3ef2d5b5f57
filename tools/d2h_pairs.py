#!/usr/bin/env python
"""D2H bandwidth of every GPU of the box alone, of GPU pairs and of all GPUs at once (pinned host memory, one process): which
GPUs share an uplink / what the host takes in total.  Run with `gpurun --gpus 8 -- python tools/d2h_pairs.py`; one JSON line."""
import itertools
import json
import subprocess
import time

import torch

n = torch.cuda.device_count()
SZ = 1 << 30
dev = [torch.empty(SZ, dtype=torch.uint8, device=f"cuda:{i}") for i in range(n)]
host = [torch.empty(SZ, dtype=torch.uint8).pin_memory() for i in range(n)]
streams = [torch.cuda.Stream(device=i) for i in range(n)]


def run(gpus, reps=3):
    best = 0.0
    for _ in range(reps):
        for i in gpus:
            torch.cuda.synchronize(i)
        t0 = time.perf_counter()
        for i in gpus:
            with torch.cuda.stream(streams[i]):
                host[i].copy_(dev[i], non_blocking=True)
        for i in gpus:
            streams[i].synchronize()
        dt = time.perf_counter() - t0
        best = max(best, len(gpus) * SZ / dt / 1e9)
    return best


for i in range(n):
    run([i], 1)
out = {"gpus": n, "single_gbs": [round(run([i]), 1) for i in range(n)]}
out["pairs_gbs"] = {f"{a}+{b}": round(run([a, b]), 1) for a, b in itertools.combinations(range(n), 2)} if n <= 8 else {}
for k in (2, 4, 8):
    if k <= n:
        out[f"first_{k}_gbs"] = round(run(list(range(k))), 1)
try:
    out["topo"] = subprocess.run("nvidia-smi topo -m | head -12", shell=True, capture_output=True, text=True).stdout
    out["cpu"] = subprocess.run("lscpu | grep -i 'model name\\|socket\\|numa node(s)\\|^CPU(s)'", shell=True, capture_output=True, text=True).stdout
except Exception:
    pass
print(json.dumps(out))
