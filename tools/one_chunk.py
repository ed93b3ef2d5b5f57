#!/usr/bin/env python
"""Two single-stream passes over ONE chunk of B elements of the headline workload (the second one is what ncu captures:
`ncu -s <launches of one pass> -c <same>`)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hp3d_b200 import synth  # noqa: E402
from hp3d_b200.api import ElemEngine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
p = int(sys.argv[2]) if len(sys.argv) > 2 else 5
eng = ElemEngine(4, omega=2 * np.pi)
a = synth.cube_mesh(B, p)
for _ in range(2):
    r = eng.bench(*a, reps=1, lanes=1, max_chunk=B)
print(r)
