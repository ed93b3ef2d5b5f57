#!/usr/bin/env python
"""Device-resident throughput of the headline workload vs (elements per step, lanes, chunk)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hp3d_b200 import synth
from hp3d_b200.api import ElemEngine
eng = ElemEngine(4, omega=2 * np.pi)
for B, lanes, chunk in [(256, 1, 128), (256, 2, 128), (256, 4, 64), (512, 4, 128), (512, 2, 128), (384, 3, 128), (256, 2, 64), (256,4,32)]:
    a = synth.cube_mesh(B, 5)
    eng.bench(*a, reps=1, lanes=lanes, max_chunk=chunk)
    r = eng.bench(*a, reps=2, lanes=lanes, max_chunk=chunk)
    print(B, lanes, chunk, "el/s %.1f" % (B * 2 / (r["ms_total"] * 1e-3)), flush=True)
