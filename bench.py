#!/usr/bin/env python
"""Benchmark of the element-local hot path: condensed element matrices per second (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--elements B] [--p 5] [--kind 4]

Workload (config[3] of BASELINE.json, the configuration the metric is quoted on): ultraweak DPG Maxwell, hexahedra of
order p=5, enrichment dp=1, complex FP64, perturbed (jittered) cube mesh so no two elements are congruent; one "step" =
one pass of the hot path (descriptors -> integration -> DPG normal equations -> static condensation) over a batch of B
elements per GPU.  Elements are block-partitioned over the ranks (hp3D's ZOLTAN_LB=0 split); there is no collective on
the path, so scaling is "weak" (fixed B per GPU).

  value     whole-job elements/s with the geometry dofs already resident in HBM (device time, CUDA events on the
            launching stream inside the library, max over ranks)
  e2e       the same metric through the C-ABI call hp3d_gpu_elem_batch with HOST buffers: H2D of the geometry dofs and
            D2H of Aii, Bi, ASchur, BSchur inside the timed region
  roofline  FP64 tensor (DMMA) roofline of the dense phase: algorithmic flops / time of the dense phase
  cpu_baseline  the CPU oracle (restatement of the reference's elem_opt + stc_fwd_herm, OpenMP over elements like
            par_mumps_sc.F90:347, single-threaded OpenBLAS per element) on the host's cores, bounded sample

--impl reference times that CPU path alone (the reference is Fortran + PETSc/MUMPS and cannot be built in this image).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GEMM_DRAM_BYTES_PER_ELEMENT_LAUNCH = 589.1e6 / 103   # ncu, p=5 ultraweak Maxwell, real-form dense phase (profiles/r01_rs_launches_traffic_b32_summary.csv)
DMMA_PEAK_TFLOPS = 37.05   # measured on this pool's B200: raw mma.sync m16n8k16.f64 loop (profiles/r01_dmma_probe.jsonl);
                           # cuBLAS DGEMM 8192^3 reaches 35.5, ZGEMM 4096^3 36.8 (profiles/r01_fp64_peak.json)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for i, nm in enumerate(names):
                if f[5 + i].lower().startswith("active"):
                    reasons.add(nm)
        # "under load" = samples drawing real power
        load = [s for s, p in zip(sm, pw) if p > 300] or sm
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


def cpu_reference_rate(kind, p, nsample, threads, omega):
    """Elements/s of the CPU restatement (oracle) on `threads` host threads over `nsample` elements of the same workload."""
    from oracle import oracle as O
    from hp3d_b200 import synth
    O.set_maxp(6)
    blas = O.use_blas(True, threads=1)
    norder, noe, nof, xnod = synth.cube_mesh(nsample, p)
    prm = O.default_params(omega=omega)
    t0 = time.perf_counter()
    out = O.condensed_batch(kind, norder, noe, nof, xnod, prm, nthreads=threads)
    dt = time.perf_counter() - t0
    assert out[-1] == 0
    return nsample / dt, dt, blas


def run_reference(args, out_stream):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    nsample = args.cpu_sample or max(8 * cores, 32)   # ~8 s of CPU work per step on the 16-core box
    omega = 2 * np.pi
    for _ in range(min(args.warmup, 1)):
        cpu_reference_rate(args.kind, args.p, max(cores, 1), cores, omega)
    t_tot, n_tot = 0.0, 0
    for _ in range(args.steps):
        r, dt, blas = cpu_reference_rate(args.kind, args.p, nsample, cores, omega)
        t_tot += dt; n_tot += nsample
    val = n_tot / t_tot
    sample = f"{nsample} elements/step of the same perturbed-cube workload, {cores} OpenMP threads over elements, single-threaded {'OpenBLAS' if blas else 'built-in loops'} per element"
    print(json.dumps({
        "impl": "reference", "metric": "condensed element matrices/sec", "value": val, "unit": "elements/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c128" if args.kind >= 3 else "f64", "data": "synthetic",
        "config": workload_config(args, nsample),
        "cpu_baseline": {"value": val, "unit": "elements/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), file=out_stream, flush=True)


def celem_leg(args, eng, norder, noe, nof, xs, ni, nb, bufs, rank):
    """End-to-end elements/s of hp3d_gpu_celem_batch (elem + stc + constraints + Dirichlet lift + compression + IRN/JCN) on the
    e2e workload, plus -- on rank 0 -- the host cost it removes: the oracle's restatement of celem_systemI.F90:543-785 and the
    COO fill (par_mumps_sc.F90:433-448) timed on one core over a bounded sample."""
    from hp3d_b200 import synth
    from hp3d_b200.api import pinned_empty
    Be = norder.shape[0]
    cons = synth.synthetic_constraints(args.kind, ni, Be)
    pk = eng.pack_constraints(cons, 2, True)
    nz, nx = int(pk["aptr"][-1]), int(pk["xptr"][-1])
    za = pinned_empty((nz,), eng.dtype); zb = pinned_empty((nx,), eng.dtype); irn = pinned_empty((nz,), np.int32); jcn = pinned_empty((nz,), np.int32)
    out = dict(zastif=za.a, zbload=zb.a, irn=irn.a, jcn=jcn.a, ASchur=bufs[2].a, BSchur=bufs[3].a)
    kw = dict(isym_flag=2, want_coo=True, want_schur=True, out=out, packed=pk)
    for _ in range(2):
        eng.celem_batch(norder, noe, nof, xs, None, **kw)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = eng.celem_batch(norder, noe, nof, xs, None, **kw)
    te = time.perf_counter() - t0
    assert (res["info"] == 0).all()
    es = 16 if args.kind >= 3 else 8
    r = {"e2e_value": Be * args.steps / te, "unit": "elements/s (this rank)", "elements_per_step": Be,
         "d2h_bytes_per_step": int(es * (nz + nx + Be * (nb * ni + nb)) + 8 * nz + 4 * Be),
         "what": "hp3d_gpu_celem_batch: Zastif (row-major) + Zbload + IRN/JCN + Schur factors to pinned host arrays; 10% of the elements with hanging-node constraints, 20% with Dirichlet data"}
    if rank == 0 and not args.no_cpu:
        from oracle import oracle as O
        pho = O.physics_of(args.kind)
        Aii = np.zeros((ni, ni), eng.dtype); Aii[...] = np.arange(ni)[:, None] + 1.0
        Bi = np.ones(ni, eng.dtype)
        ns = min(Be, 16)
        t0 = time.perf_counter()
        for c in cons[:ns]:
            zbl, zas = O.celem_modify(pho, c["nrdofl"], c["nrcon"], c["nac"], c["constr"], c["nrdofm_f"], Aii, Bi, c["idbc"], c["zdofd"], c["nextract"], 2)
            O.coo_fill(c["lcon"], zas, zbl, int(c["lcon"].max()))
        tc = time.perf_counter() - t0
        r["cpu_port"] = {"value": ns / tc, "unit": "elements/s", "cores": 1, "kind": "port",
                         "sample": f"{ns} elements: oracle celem_modify + coo_fill (restatement of the host loops), one thread"}
    for b in (za, zb, irn, jcn):
        b.free()
    return r


def workload_config(args, B):
    names = {1: "Poisson Galerkin", 2: "Poisson primal DPG (dp=1)", 3: "Maxwell Galerkin", 4: "Maxwell ultraweak DPG (dp=1, adjoint-graph norm)"}
    return {"workload": f"{names[args.kind]}, hexa p={args.p}, complex FP64, perturbed cube mesh (jitter 0.15h, seed 12345)" if args.kind >= 3
            else f"{names[args.kind]}, hexa p={args.p}, real FP64, perturbed cube mesh (jitter 0.15h, seed 12345)",
            "elements_per_gpu_per_step": B, "partition": "contiguous blocks of the element list per rank (ZOLTAN_LB=0)",
            "l2": "per-step working set (>100 MB per element at p=5) far exceeds the 126 MB L2; no flush needed"}


def claim_stdout():
    """stdout must carry ONE JSON line: keep a private handle on it and point fd 1 at stderr for everything else (NCCL prints its
    version banner on stdout, other libraries may too)."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def bind_to_gpu_numa(gpu_index):
    """Pin this rank's host threads to the CPUs next to its GPU (NVML's affinity mask) BEFORE any pinned allocation, so that the
    result buffers live on the GPU's NUMA node: with several ranks per host the D2H copies otherwise cross the socket
    interconnect.  (What `mpirun --bind-to` / numactl does for the reference's MPI ranks.)  Returns the CPU list or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return allowed
    except Exception:
        pass
    return None


def main():
    out_stream = claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--elements", type=int, default=256, help="elements per GPU per step")
    ap.add_argument("--e2e-elements", type=int, default=0, help="elements per GPU per end-to-end step (a subdomain slice; default 1024 = 13.3 GB of results at p=5 on one GPU, 512 per GPU on several: the ranks share the host's pinned memory)")
    ap.add_argument("--complex-kernels", action="store_true", help="force the general complex dense phase (hp3d_params.real_reduction = 0), the reference's ZPOTRF/ZTRTRS/ZHERK sequence")
    ap.add_argument("--p", type=int, default=5)
    ap.add_argument("--kind", type=int, default=4)
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--condensed-only", action="store_true", help="also time the end-to-end step with hp3d_params.store_schur = 0: only the condensed system returns (5.8 instead of 13.0 MB per element at p=5); the bubbles are recovered by hp3d_gpu_elem_bwd_batch")
    ap.add_argument("--celem", action="store_true", help="also time SURVEY 8f row f1 (constraints + compression + COO fused into the batched call)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, out_stream)

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    numa_cpus = bind_to_gpu_numa(local) if world > 1 and not os.environ.get("HP3D_NO_NUMA_BIND") else None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from hp3d_b200 import synth
    from hp3d_b200.api import ElemEngine, pinned_empty
    omega = 2 * np.pi if args.kind == 4 else 1.0
    eng = ElemEngine(args.kind, device=local, omega=omega, real_reduction=0 if args.complex_kernels else 1)
    B = args.elements
    norder, noe, nof, xnod = synth.cube_mesh(B, args.p, first=rank * B, total=world * B)
    ntest, ntrial, ni, nb = synth.problem_sizes(args.kind, args.p)
    F_ref = synth.dense_flops(args.kind, ntest, ntrial, ni, nb)          # the reference's algorithm (complex: c = 4)
    real_form = args.kind == 4 and not args.complex_kernels               # lossless ultraweak Maxwell: A = T A~ T^H, A~ real
    F_dense = synth.dense_flops_real_form(ntest, ntrial, ni, nb) if real_form else F_ref

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up (also builds the signature tables and allocates the workspaces)
    for _ in range(max(args.warmup, 3)):
        eng.bench(norder, noe, nof, xnod, reps=1, lanes=2)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    # K steps; each step = the B elements as two half-batches on two streams (hp3d_gpu_elem_batch rotates its chunks over four);
    # CUDA events on the launching stream bracket the K steps (the second stream is fenced inside that interval)
    r = eng.bench(norder, noe, nof, xnod, reps=args.steps, lanes=2)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = r["ms_total"]
    # stage breakdown (integration / dense phase) from a single-stream pass of the same K steps: stage events are only
    # meaningful when nothing overlaps
    r1 = eng.bench(norder, noe, nof, xnod, reps=args.steps, lanes=1)
    ms_dense, ms_integ, ms_single = r1["ms_dense"], r1["ms_integ"], r1["ms_total"]
    if dist is not None:
        import torch
        t = torch.tensor([ms, ms_dense, ms_integ, ms_single], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_dense, ms_integ, ms_single = [float(v) for v in t.tolist()]
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end to end through hp3d_gpu_elem_batch with pinned host buffers
    e2e = None
    if not args.no_e2e:
        Be = args.e2e_elements or (1024 if world == 1 else 512)
        if Be > B:
            norder, noe, nof, xnod = synth.cube_mesh(Be, args.p, first=rank * Be, total=world * Be)
        dt = eng.dtype
        bufs = [pinned_empty((Be, ni * ni), dt), pinned_empty((Be, ni), dt), pinned_empty((Be, max(nb * ni, 1)), dt), pinned_empty((Be, max(nb, 1)), dt)]
        out = dict(Aii=bufs[0].a, Bi=bufs[1].a, ASchur=bufs[2].a, BSchur=bufs[3].a)
        xs = pinned_empty(xnod[:Be].shape, np.float64)
        xs.a[...] = xnod[:Be]
        for _ in range(2):
            eng.elem_stc_batch(norder[:Be], noe[:Be], nof[:Be], xs.a, out=out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = eng.elem_stc_batch(norder[:Be], noe[:Be], nof[:Be], xs.a, out=out)   # returns after the last D2H completed
        te = time.perf_counter() - t0
        barrier()
        assert (res["info"] == 0).all()
        if dist is not None:
            import torch
            t = torch.tensor([te], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t.item())
        es = 16 if args.kind >= 3 else 8
        e2e = {"value": world * Be * args.steps / te, "unit": "elements/s", "elements_per_gpu_per_step": Be,
               "h2d_bytes_per_step": int(xs.a.nbytes), "d2h_bytes_per_step": int(Be * (es * (ni * ni + ni + nb * ni + nb) + 4)),
               "timer": "host wall clock around the synchronous C-ABI calls (they return after the last D2H)",
               "host_binding": f"rank bound to the {len(numa_cpus)} CPUs NVML lists for its GPU" if numa_cpus else "none"}
        # ---- optional: the same step with celem_systemI's transform / compression / COO indices fused in (hp3d_gpu_celem_batch)
        celem = None
        if args.celem:
            celem = celem_leg(args, eng, norder[:Be], noe[:Be], nof[:Be], xs.a, ni, nb, bufs, rank)
        # ---- optional: STORE_STC off -- only Aii / Bi cross PCIe (SURVEY 8f row f3 recomputes the factors on the device when needed)
        cond_only = None
        if args.condensed_only:
            eng2 = ElemEngine(args.kind, device=local, omega=omega, real_reduction=0 if args.complex_kernels else 1, store_schur=0)
            tiny = [pinned_empty((Be, 1), dt), pinned_empty((Be, 1), dt)]
            out2 = dict(Aii=bufs[0].a, Bi=bufs[1].a, ASchur=tiny[0].a, BSchur=tiny[1].a)
            for _ in range(2):
                eng2.elem_stc_batch(norder[:Be], noe[:Be], nof[:Be], xs.a, out=out2)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                res2 = eng2.elem_stc_batch(norder[:Be], noe[:Be], nof[:Be], xs.a, out=out2)
            t2 = time.perf_counter() - t0
            barrier()
            assert (res2["info"] == 0).all()
            if dist is not None:
                import torch
                t = torch.tensor([t2], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                t2 = float(t.item())
            cond_only = {"value": world * Be * args.steps / t2, "unit": "elements/s", "elements_per_gpu_per_step": Be,
                         "d2h_bytes_per_step": int(Be * (es * (ni * ni + ni) + 4)),
                         "what": "hp3d_gpu_elem_batch with store_schur = 0: the condensed system only (stc.F90 STORE_STC = .false.)"}
            for b in tiny:
                b.free()
            eng2.close()
        for b in bufs + [xs]:
            b.free()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    achieved = F_dense * B * args.steps / (ms_dense * 1e-3) / 1e12
    out = {
        "metric": "condensed element matrices/sec", "value": value, "unit": "elements/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "c128" if args.kind >= 3 else "f64", "data": "synthetic", "config": workload_config(args, B),
        "gpu_launches": int(r["launches"]), "clocks": clocks, "e2e": e2e,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": DMMA_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / DMMA_PEAK_TFLOPS,
                     "traffic": GEMM_DRAM_BYTES_PER_ELEMENT_LAUNCH * ((B + 1) // 2) if args.kind == 4 and args.p == 5 else None,
                     "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, average over the 103 gemm_nc_kernel launches of one 32-element chunk, scaled to this run's chunk (profiles/r01_launches_traffic_b32_summary.csv): bytes per launch; the operands stream from HBM once per panel, 12 % of HBM bandwidth",
                     "kernel": ("gemm_nc_kernel<real>" if (real_form or args.kind < 3) else "gemm_nc_kernel<complex>") + " (all launches of the dense phase: Cholesky panels, solves, HERK, Schur)",
                     "flops_per_element": F_dense,
                     "algorithm": ("real form of the lossless ultraweak Maxwell system (A = T A~ T^H, T = diag(i^k), A~ real: DESIGN.md 2.7): the flops counted are "
                                   "those of the real factor/solve/rank-k sequence with two load columns; the reference's complex ZPOTRF/ZTRTRS/ZHERK sequence "
                                   "costs reference_flops_per_element for the same result") if real_form else "the reference's factor/solve/rank-k sequence (SURVEY 8d)",
                     "reference_flops_per_element": F_ref,
                     "tflops_on_reference_count": F_ref * B * args.steps / (ms_dense * 1e-3) / 1e12, "ms_dense_per_step": ms_dense / args.steps, "ms_integration_per_step": ms_integ / args.steps,
                     "ms_per_step_single_stream": ms_single / args.steps,
                     "timing": "achieved = algorithmic dense flops / CUDA-event time of the dense phase in a single-stream pass of the same K steps; whole_step_frac uses the two-stream step time that `value` reports",
                     "peak_source": "own probe: raw FP64 DMMA loop on this pool's B200 (profiles/r01_dmma_probe.jsonl); MEASURED_PEAKS.json has no FP64 entry",
                     "whole_step_frac": F_dense * B * args.steps / (ms * 1e-3) / 1e12 / DMMA_PEAK_TFLOPS},
    }
    if not args.no_e2e and args.celem:
        out["celem"] = celem
    if not args.no_e2e and args.condensed_only:
        out["e2e_condensed_only"] = cond_only
    if not args.no_cpu and world == 1:   # CPU baseline beside the GPU number: rank 0 at N=1 only
        cores = host_cores()
        ns = args.cpu_sample or max(16 * cores, 64)   # ~15 s of CPU work (bounded sample)
        v, dtc, blas = cpu_reference_rate(args.kind, args.p, ns, cores, omega)
        out["cpu_baseline"] = {"value": v, "unit": "elements/s", "cores": cores, "kind": "port",
                               "sample": f"{ns} elements of the same workload in {dtc:.1f} s; OpenMP over elements ({cores} threads), single-threaded {'OpenBLAS' if blas else 'built-in loops'} per element"}
    print(json.dumps(out), file=out_stream, flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
