"""N>1 host logic on CPU: block partition (par_mesh.F90:66-82) and the max-over-ranks timing reduction of bench.py,
exercised with a real 2-process gloo group."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from hp3d_b200 import partition, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_block_partition_matches_reference_rule():
    for n, p in [(10, 3), (8, 8), (7, 2), (1000, 8), (5, 8)]:
        own = partition.block_partition(n, p)
        sizes = np.bincount(own, minlength=p)
        base, rem = divmod(n, p)
        assert list(sizes) == [base + 1] * rem + [base] * (p - rem)
        assert (np.diff(own) >= 0).all()
        for r in range(p):
            a, b = partition.block_range(n, p, r)
            assert list(partition.elem_subd(own, r)) == list(range(a, b))


def test_weighted_partition_balances():
    rng = np.random.default_rng(2024)
    w = np.array([synth.dense_flops(4, *synth.problem_sizes(4, int(p))) for p in rng.integers(2, 8, 400)])
    own = partition.weighted_partition(w, 8)
    load = np.bincount(own, weights=w, minlength=8)
    assert (np.diff(own) >= 0).all() and load.max() / load.mean() < 1.15


def test_mesh_blocks_tile_the_serial_mesh():
    no, oe, of, x = synth.cube_mesh(27, 2)
    parts = [synth.cube_mesh(b - a, 2, first=a, total=27) for a, b in (partition.block_range(27, 2, r) for r in range(2))]
    # the per-rank generator must see the same global mesh: same N for every rank
    xs = np.concatenate([q[3] for q in parts])
    assert np.array_equal(xs, x)


WORKER = textwrap.dedent("""
    import os, sys, numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from hp3d_b200 import partition
    dist.init_process_group("gloo")
    r, w = dist.get_rank(), dist.get_world_size()
    a, b = partition.block_range(37, w, r)
    t = torch.tensor([float(b - a), 10.0 + r], dtype=torch.float64)   # [elements done, my time]
    n = t[:1].clone(); dist.all_reduce(n, op=dist.ReduceOp.SUM)
    tm = t[1:].clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    if r == 0:
        print("RESULT", int(n.item()), tm.item(), flush=True)
    dist.barrier(); dist.destroy_process_group()
""")


def test_two_rank_gloo_reduction(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "w.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), stdout=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    line = [ln for ln in outs[0].splitlines() if ln.startswith("RESULT")][0].split()
    assert int(line[1]) == 37 and float(line[2]) == 11.0     # all elements covered once; time = max over ranks


def test_chunk_plan_covers_every_batch_size(gpulib):
    """The chunk plan of hp3d_gpu_elem_batch (ramped start, geometric taper; host logic, no GPU): for every batch size the
    chunks are non-empty, at most `cap` elements, and add up to the batch; large batches start with a ramp and end with the
    8-element chunks that keep the exposed result copy small."""
    import ctypes as C
    f = gpulib.hp3d_gpu_chunk_plan_debug
    f.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    buf = (C.c_longlong * 8192)()
    for cap in (1, 3, 4, 16, 25, 64):
        for nl in (1, 2, 4):
            for n in list(range(0, 300)) + [511, 512, 513, 1000, 1024, 4097]:
                k = f(n, cap, nl, 0, buf, 8192)
                sizes = list(buf[:k])
                assert sum(sizes) == n and all(1 <= s <= cap for s in sizes), (cap, nl, n, sizes)
    k = f(1024, 64, 4, 0, buf, 4096)
    sizes = list(buf[:k])
    assert sizes[:4] == [16, 32, 48, 64] and sizes[-4:] == [8, 8, 8, 8] and sizes[-8:-4] == [16, 16, 16, 16]
    k = f(100, 64, 4, 7, buf, 4096)      # forced chunk size (hp3d_gpu_set_chunk): plain chunks of `cap`
    assert sum(buf[:k]) == 100
