#!/usr/bin/env python
"""Throughput of the hot path on BASELINE.json configs[4]: an hp-refined mixed hexahedron / prism mesh, orders 2..7,
orientations from a random global vertex numbering (hp3d_b200.synth.hp_mesh).  Not the headline metric (bench.py is); it
exercises heterogeneous batching: one signature per (element type, orders, orientations) combination.

  python tools/bench_hp.py [--N 8] [--reps 2] [--gpus-split 8]

Prints one JSON line: elements/s device-resident and end to end (host buffers), signatures, host compile time of the
signatures (first call), achieved dense TFLOP/s, and the flop-weighted contiguous partition imbalance for --gpus-split ranks
(partition.weighted_partition, the role of Zoltan's OBJ_WEIGHT in zoltan_wrapper.F90:563).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=8)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--kind", type=int, default=4)
    ap.add_argument("--pmin", type=int, default=2)
    ap.add_argument("--pmax", type=int, default=7)
    ap.add_argument("--gpus-split", type=int, default=8)
    ap.add_argument("--sort", action="store_true", help="sort the element list by signature size before batching")
    args = ap.parse_args()
    from hp3d_b200 import partition, synth
    from hp3d_b200.api import ElemEngine
    m = synth.hp_mesh(args.N, pmin=args.pmin, pmax=args.pmax, jitter=0.1)
    nel = len(m["etype"])
    eng = ElemEngine(args.kind, omega=2 * np.pi if args.kind == 4 else 1.0, maxp=8)
    t0 = time.perf_counter()
    dims = [eng.sig_dims(m["norder"][e], m["norient_edge"][e], m["norient_face"][e], int(m["etype"][e])) for e in range(nel)]
    t_compile = time.perf_counter() - t0
    flops = np.array([synth.dense_flops(args.kind, d["ntest"], d["ni"] + d["nb"], d["ni"], d["nb"]) for d in dims])   # reference count
    flops_exec = np.array([synth.dense_flops_real_form(d["ntest"], d["ni"] + d["nb"], d["ni"], d["nb"]) for d in dims]) if args.kind == 4 else flops
    nsig = len({(int(m["etype"][e]),) + tuple(m["norder"][e]) + tuple(m["norient_edge"][e]) + tuple(m["norient_face"][e]) for e in range(nel)})
    owner = partition.weighted_partition(flops, args.gpus_split)
    loads = np.array([flops[owner == r].sum() for r in range(args.gpus_split)])
    a = (m["norder"], m["norient_edge"], m["norient_face"], m["xnod"])
    eng.bench(*a, reps=1, lanes=4, etype=m["etype"])            # warm-up: uploads the signature tables
    r = eng.bench(*a, reps=args.reps, lanes=4, etype=m["etype"])
    ms = r["ms_total"] / args.reps
    from hp3d_b200.api import pinned_empty
    ni_max = max(d["ni"] for d in dims); nb_max = max(d["nb"] for d in dims)
    bufs = [pinned_empty((nel, ni_max * ni_max), eng.dtype), pinned_empty((nel, ni_max), eng.dtype),
            pinned_empty((nel, max(nb_max * ni_max, 1)), eng.dtype), pinned_empty((nel, max(nb_max, 1)), eng.dtype)]
    out = dict(Aii=bufs[0].a, Bi=bufs[1].a, ASchur=bufs[2].a, BSchur=bufs[3].a)
    eng.elem_stc_batch(*a, etype=m["etype"], out=out)
    t0 = time.perf_counter()
    res = eng.elem_stc_batch(*a, etype=m["etype"], out=out)
    te = time.perf_counter() - t0
    assert (res["info"] == 0).all()
    # cold call: a fresh plan has no signature compiled; the call compiles them on all host cores, uploads them and computes
    eng2 = ElemEngine(args.kind, omega=2 * np.pi if args.kind == 4 else 1.0, maxp=8)
    t0 = time.perf_counter()
    res2 = eng2.elem_stc_batch(*a, etype=m["etype"], out=out)
    tcold = time.perf_counter() - t0
    assert (res2["info"] == 0).all()
    print(json.dumps({
        "workload": f"hp mesh N={args.N}: {nel} elements ({int((m['etype'] == 3).sum())} prisms), orders {args.pmin}..{args.pmax}, kind {args.kind}",
        "signatures": nsig, "host_compile_s": t_compile, "elements_per_s": nel / (ms * 1e-3), "ms_per_pass": ms,
        "dense_tflops": flops_exec.sum() / (ms * 1e-3) / 1e12, "dense_tflops_on_reference_count": flops.sum() / (ms * 1e-3) / 1e12, "launches_per_pass": r["launches"] / args.reps,
        "e2e_elements_per_s": nel / te, "cold_e2e_elements_per_s": nel / tcold, "cold_call_s": tcold, "host_cores": os.cpu_count(), "partition_imbalance_max_over_mean": float(loads.max() / loads.mean()),
        "ranks": args.gpus_split}))


if __name__ == "__main__":
    main()
