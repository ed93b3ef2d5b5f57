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
  e2e       the same metric through the C-ABI call hp3d_gpu_elem_batch_cloc with HOST buffers: H2D of the geometry dofs, D2H of
            the condensed system (the Hermitian Aii as its LAPACK-packed lower triangle, Bi) inside the timed region; the Schur
            factors (CLOC) stay in HBM, where the back-substitution reads them (stc_bwd_on_store).  e2e_variants: the same step
            with the full ni x ni Aii on the host; --host-factors: round 1's step (the Schur factors to the host as well)
  configs   sub-records: the other BASELINE.json configs and the general complex dense phase, device-resident, same rules
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

TRAFFIC_FILE = os.path.join(ROOT, "profiles", "gemm_traffic.json")   # ncu dram bytes of the dominant kernel (tools/profile_gpu.sh writes it)
DMMA_PEAK_FALLBACK_TFLOPS = 37.05   # only if the in-run probe fails: raw mma.sync f64 loop on this pool's B200 (profiles/r01_dmma_probe.jsonl)
TRAP_W = 64   # block-column width of the lower trapezoids hp3d_params.aii_packed = 2 moves over PCIe (csrc/hp3d_gpu.cu)


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def trapezoid_scalars(ni):
    """scalars of one element's Aii that cross PCIe with aii_packed = 2"""
    if ni <= TRAP_W:
        return ni * ni
    return sum(min(TRAP_W, ni - c0) * (ni - c0) for c0 in range(0, ni, TRAP_W))


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
    """The reference arm: the CPU restatement of elem_opt + stc_fwd_herm (the reference itself is Fortran + PETSc/MUMPS/Zoltan and
    cannot be built here) on all host cores, on the same config as the GPU arm: `--elements` elements per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    nsample = args.cpu_sample or args.elements
    omega = 2 * np.pi if args.kind == 4 else 1.0
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


def cpu_rate_mixed(kind, mesh, idx, threads, omega):
    """Elements/s of the CPU oracle over the elements `idx` of a mixed hexa/prism mesh (one ctypes call per element from a thread
    pool: the calls release the GIL, so this is the OpenMP-over-elements structure of par_mumps_sc.F90:347)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    O.set_maxp(8)
    blas = O.use_blas(True, threads=1)
    prm = O.default_params(omega=omega)

    def one(e):
        et = int(mesh["etype"][e])
        nH = O.celndof(mesh["norder"][e], et)[0]
        O.condensed(kind, mesh["norder"][e], mesh["norient_edge"][e], mesh["norient_face"][e], mesh["xnod"][e, :nH], prm, etype=et)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(one, idx))
    dt = time.perf_counter() - t0
    return len(idx) / dt, dt, blas


def sub_config(name, kind, p, B, steps, local, rank, world, maxr, peak_tf, hbm_gbs, complex_kernels=False, mesh=None, lanes=2, cpu_sample=0):
    """Device-resident throughput of another BASELINE.json config on this rank's GPU (same timing rules as the headline:
    warm-up, CUDA events on the launching stream inside the library, max over ranks): value, dense TFLOP/s against the measured
    FP64 tensor peak, and the algorithmic output bytes per second against the measured HBM bandwidth."""
    from hp3d_b200 import synth
    from hp3d_b200.api import ElemEngine
    omega = 2 * np.pi if kind == 4 else 1.0
    eng = ElemEngine(kind, device=local, omega=omega, maxp=8 if mesh is not None else 6, real_reduction=0 if complex_kernels else 1)
    es = 16 if kind >= 3 else 8
    if mesh is None:
        norder, noe, nof, xnod = synth.cube_mesh(B, p, first=rank * B, total=world * B)
        et = None
        ntest, ntrial, ni, nb = synth.problem_sizes(kind, p)
        real_form = kind == 4 and not complex_kernels
        F = (synth.dense_flops_real_form(ntest, ntrial, ni, nb) if real_form else synth.dense_flops(kind, ntest, ntrial, ni, nb)) * B
        F_ref = synth.dense_flops(kind, ntest, ntrial, ni, nb) * B
        out_bytes = es * (ni * ni + ni + nb * ni + nb) * B
    else:
        norder, noe, nof, xnod, et = mesh["norder"], mesh["norient_edge"], mesh["norient_face"], mesh["xnod"], mesh["etype"]
        B = len(et)
        dims = [eng.sig_dims(norder[e], noe[e], nof[e], int(et[e])) for e in range(B)]
        F = sum(synth.dense_flops_real_form(d["ntest"], d["ni"] + d["nb"], d["ni"], d["nb"]) if kind == 4 and not complex_kernels
                else synth.dense_flops(kind, d["ntest"], d["ni"] + d["nb"], d["ni"], d["nb"]) for d in dims)
        F_ref = sum(synth.dense_flops(kind, d["ntest"], d["ni"] + d["nb"], d["ni"], d["nb"]) for d in dims)
        out_bytes = sum(es * (d["ni"] ** 2 + d["ni"] + d["nb"] * d["ni"] + d["nb"]) for d in dims)
    for _ in range(3):
        eng.bench(norder, noe, nof, xnod, reps=1, lanes=lanes, etype=et)
    r = eng.bench(norder, noe, nof, xnod, reps=steps, lanes=lanes, etype=et)
    ms = maxr(r["ms_total"]) / steps
    eng.close()
    e2e = None
    if mesh is not None:   # the same mesh end to end: host descriptors in, condensed systems (packed Hermitian Aii, Bi) out, Schur factors resident in HBM
        from hp3d_b200.api import pinned_empty
        eng_e = ElemEngine(kind, device=local, omega=omega, maxp=8, real_reduction=0 if complex_kernels else 1, aii_packed=1)
        ni_max = max(d["ni"] for d in dims)
        bufs = [pinned_empty((B, ni_max * (ni_max + 1) // 2), eng_e.dtype), pinned_empty((B, ni_max), eng_e.dtype)]
        out = dict(Aii=bufs[0].a, Bi=bufs[1].a)
        cl = eng_e.cloc_create()
        for _ in range(2):
            eng_e.elem_stc_batch_cloc(cl, norder, noe, nof, xnod, etype=et, out=out)
        t0 = time.perf_counter()
        for _ in range(steps):
            re_ = eng_e.elem_stc_batch_cloc(cl, norder, noe, nof, xnod, etype=et, out=out)
        te = maxr(time.perf_counter() - t0)
        assert (re_["info"] == 0).all()
        e2e = {"value": world * B * steps / te, "unit": "elements/s",
               "d2h_bytes_per_step": int(sum(es * (d["ni"] * (d["ni"] + 1) // 2 + d["ni"]) + 4 for d in dims)),
               "call": "hp3d_gpu_elem_batch_cloc (aii_packed = 1), pinned host arrays, Schur factors resident in HBM"}
        eng_e.cloc_destroy(cl)
        eng_e.close()
        for b in bufs:
            b.free()
    tf = F / (ms * 1e-3) / 1e12
    gbs = out_bytes / (ms * 1e-3) / 1e9
    cpu = None
    if cpu_sample and world == 1:   # the CPU oracle beside it: all host cores, bounded sample of the same workload
        cores = host_cores()
        if mesh is None:
            v, dtc, blas = cpu_reference_rate(kind, p, cpu_sample, cores, omega)
            what = f"{cpu_sample} elements of the same workload"
        else:
            idx = np.random.default_rng(5).choice(B, size=min(cpu_sample, B), replace=False)
            v, dtc, blas = cpu_rate_mixed(kind, mesh, idx, cores, omega)
            what = f"{len(idx)} randomly chosen elements of the same mesh"
        cpu = {"value": v, "unit": "elements/s", "cores": cores, "kind": "port", "sample": f"{what} in {dtc:.1f} s; {cores} threads over elements, single-threaded {'OpenBLAS' if blas else 'built-in loops'} per element"}
    return {"workload": name, "value": world * B / (ms * 1e-3), "unit": "elements/s", "elements_per_gpu_per_step": B, "steps": steps, "ms_per_step": ms, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches_per_step": r["launches"] / steps, "dense_tflops_per_gpu": tf, "frac_fp64_tensor_peak": tf / peak_tf,
            "dense_tflops_on_reference_count": F_ref / (ms * 1e-3) / 1e12,
            "algorithmic_output_gbs_per_gpu": gbs, "frac_hbm": gbs / hbm_gbs,
            "bound": "tensor" if tf / peak_tf > gbs / hbm_gbs else "hbm"}


def main():
    out_stream = claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--elements", type=int, default=512, help="elements per GPU per step (two lane chunks of half that size: 256-element launches fill the GPU for 10-20 waves)")
    ap.add_argument("--e2e-elements", type=int, default=0, help="elements per GPU per end-to-end step (a subdomain slice; default 1024)")
    ap.add_argument("--complex-kernels", action="store_true", help="force the general complex dense phase (hp3d_params.real_reduction = 0), the reference's ZPOTRF/ZTRTRS/ZHERK sequence")
    ap.add_argument("--p", type=int, default=5)
    ap.add_argument("--kind", type=int, default=4)
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the sub-records of the other BASELINE.json configs")
    ap.add_argument("--host-factors", action="store_true", help="also time the round-1 end-to-end step: full Aii AND the Schur factors to the host (13.0 MB per element at p=5)")
    ap.add_argument("--no-variants", action="store_true", help="skip the end-to-end variants that deliver the full Aii (aii_packed = 0 and 2)")
    ap.add_argument("--celem", action="store_true", help="also time SURVEY 8f row f1 (constraints + compression + COO fused into the batched call)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, out_stream)

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    numa_cpus = bind_to_gpu_numa(local) if world > 1 and not os.environ.get("HP3D_NO_NUMA_BIND") else None
    # the library's host mirror threads (aii_packed = 2): this rank's share of the cores it may run on, one left for the submitting thread
    os.environ.setdefault("HP3D_HOST_THREADS", str(max(1, min(8, host_cores() // max(world, 1) - 1))))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from hp3d_b200 import synth
    from hp3d_b200.api import ElemEngine, pinned_empty
    import ctypes as C
    omega = 2 * np.pi if args.kind == 4 else 1.0
    rr = 0 if args.complex_kernels else 1
    eng = ElemEngine(args.kind, device=local, omega=omega, real_reduction=rr)
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

    def maxr(x):   # max over ranks of a host float (device times are already CUDA-event times)
        if dist is None:
            return float(x)
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def minr(x):
        return -maxr(-x)

    # ---- warm-up (also builds the signature tables and allocates the workspaces)
    for _ in range(max(args.warmup, 3)):
        eng.bench(norder, noe, nof, xnod, reps=1, lanes=2)
    # ---- the roofline denominator, measured on this run's device: raw FP64 DMMA issue rate (0.1 s)
    tf_, ms_ = C.c_double(0), C.c_double(0)
    peak_src = "measured in this run: hp3d_gpu_fp64_peak_probe (mma.sync m8n8k4 f64 loop, 2 CTAs x 8 warps per SM, best of 4); MEASURED_PEAKS.json has no FP64 entry"
    if eng.L.hp3d_gpu_fp64_peak_probe(C.byref(tf_), C.byref(ms_)) != 0 or not (tf_.value > 1.0):
        tf_.value = DMMA_PEAK_FALLBACK_TFLOPS
        peak_src = "fallback constant (the in-run probe failed): profiles/r01_dmma_probe.jsonl"
    peak_tf = minr(tf_.value)
    hbm_gbs = float(measured_peaks().get("hbm_gbs", 6539.5))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    # K steps; each step = the B elements as two half-batches on two streams (hp3d_gpu_elem_batch rotates its chunks over four);
    # CUDA events on the launching stream bracket the K steps (the second stream is fenced inside that interval)
    r = eng.bench(norder, noe, nof, xnod, reps=args.steps, lanes=2)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = maxr(r["ms_total"])
    # stage breakdown (integration / dense phase) from a single-stream pass of the same K steps: stage events are only
    # meaningful when nothing overlaps
    r1 = eng.bench(norder, noe, nof, xnod, reps=args.steps, lanes=1)
    ms_dense, ms_integ, ms_single = maxr(r1["ms_dense"]), maxr(r1["ms_integ"]), maxr(r1["ms_total"])
    value = world * B * args.steps / (ms * 1e-3)
    es = 16 if args.kind >= 3 else 8
    herm = args.kind in (2, 4)

    # ---- end to end through the C ABI with HOST buffers: hp3d_gpu_elem_batch_cloc -- geometry dofs H2D, the condensed system
    # (Aii, Bi) D2H into the caller's full arrays; the Schur factors (CLOC) stay in HBM where stc_bwd reads them (device-resident
    # store); for the Hermitian problems only the lower block-trapezoids of Aii cross PCIe (aii_packed = 2) and the library's host
    # threads mirror the rest inside the call
    e2e = celem = host_factors = bwd = variants = None
    if not args.no_e2e:
        Be = args.e2e_elements or 1024
        if Be > B:
            norder, noe, nof, xnod = synth.cube_mesh(Be, args.p, first=rank * Be, total=world * Be)
        dt = eng.dtype
        eng_e = ElemEngine(args.kind, device=local, omega=omega, real_reduction=rr, aii_packed=1 if herm else 0)
        cl = eng_e.cloc_create()
        bufs = [pinned_empty((Be, ni * ni), dt), pinned_empty((Be, ni), dt)]
        out = dict(Aii=bufs[0].a, Bi=bufs[1].a)
        xs = pinned_empty(xnod[:Be].shape, np.float64)
        xs.a[...] = xnod[:Be]
        iel = np.arange(Be, dtype=np.int64) + rank * Be
        for _ in range(2):
            eng_e.elem_stc_batch_cloc(cl, norder[:Be], noe[:Be], nof[:Be], xs.a, iel=iel, out=out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = eng_e.elem_stc_batch_cloc(cl, norder[:Be], noe[:Be], nof[:Be], xs.a, iel=iel, out=out)   # returns after the last D2H
        te = maxr(time.perf_counter() - t0)
        barrier()
        assert (res["info"] == 0).all()
        st = eng_e.cloc_stats(cl)
        aii_scalars = ni * (ni + 1) // 2 if herm else ni * ni
        e2e = {"value": world * Be * args.steps / te, "unit": "elements/s", "elements_per_gpu_per_step": Be,
               "h2d_bytes_per_step": int(xs.a.nbytes), "d2h_bytes_per_step": int(Be * (es * (aii_scalars + ni) + 4)),
               "call": "hp3d_gpu_elem_batch_cloc (aii_packed = %d): %s and Bi to pinned host arrays; ASchur/BSchur stay in HBM (CLOC store: %d resident, %d spilled elements, %.2f GB) where hp3d_gpu_cloc_bwd_batch reads them"
                       % (1 if herm else 0, "the Hermitian Aii as its LAPACK-packed lower triangle ('L')" if herm else "Aii", st["resident"], st["spilled"], st["bytes"] / 1e9),
               "host_work_in_timed_region": "descriptor staging",
               "timer": "host wall clock around the synchronous C-ABI calls (they return after the last D2H)",
               "host_binding": f"rank bound to the {len(numa_cpus)} CPUs NVML lists for its GPU" if numa_cpus else "none"}
        # ---- stc_bwd on the device-resident store (stc.F90:661-677): xi H2D, xb D2H
        xi = np.ones((Be, ni), dt)
        eng_e.cloc_bwd_batch(cl, xi, iel=iel, nb_max=nb)
        barrier()
        t0 = time.perf_counter()
        ob = eng_e.cloc_bwd_batch(cl, xi, iel=iel, nb_max=nb)
        tb = maxr(time.perf_counter() - t0)
        barrier()
        assert (ob["info"] == 0).all()
        bwd = {"value": world * Be / tb, "unit": "elements/s", "what": "hp3d_gpu_cloc_bwd_batch: xb = BSchur - ASchur xi on the device-resident factors, host xi -> host xb",
               "factor_gbs_per_gpu": es * nb * ni * Be / tb / 1e9}
        eng_e.cloc_destroy(cl)
        eng_e.close()
        # ---- the same step with the FULL ni x ni Aii delivered to the host: whole over PCIe (aii_packed = 0), or as lower block
        # trapezoids + mirror by the library's host threads inside the call (2).  A full matrix costs the host 5.8 MB of DRAM
        # writes per element whoever writes them; with eight ranks on one host that, not the GPUs, is the limit.
        if herm and not args.no_variants:
            variants = {}
            for mode_, nm in ((0, "full_aii_whole_over_pcie"), (2, "full_aii_lower_trapezoids_over_pcie_host_mirror_in_call")):
                ev = ElemEngine(args.kind, device=local, omega=omega, real_reduction=rr, aii_packed=mode_)
                clv = ev.cloc_create()
                for _ in range(2):
                    ev.elem_stc_batch_cloc(clv, norder[:Be], noe[:Be], nof[:Be], xs.a, iel=iel, out=out)
                barrier()
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    rv = ev.elem_stc_batch_cloc(clv, norder[:Be], noe[:Be], nof[:Be], xs.a, iel=iel, out=out)
                tv = maxr(time.perf_counter() - t0)
                barrier()
                assert (rv["info"] == 0).all()
                variants[nm] = {"value": world * Be * args.steps / tv, "unit": "elements/s",
                                "d2h_bytes_per_step": int(Be * (es * ((ni * ni if mode_ == 0 else trapezoid_scalars(ni)) + ni) + 4))}
                if mode_ == 2:
                    variants[nm]["host_mirror_threads"] = int(os.environ["HP3D_HOST_THREADS"])
                ev.cloc_destroy(clv)
                ev.close()
        # ---- optional: the round-1 step, everything to the host (full Aii and the Schur factors)
        if args.host_factors:
            sb = [pinned_empty((Be, max(nb * ni, 1)), dt), pinned_empty((Be, max(nb, 1)), dt)]
            out2 = dict(Aii=bufs[0].a, Bi=bufs[1].a, ASchur=sb[0].a, BSchur=sb[1].a)
            for _ in range(2):
                eng.elem_stc_batch(norder[:Be], noe[:Be], nof[:Be], xs.a, out=out2)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                res2 = eng.elem_stc_batch(norder[:Be], noe[:Be], nof[:Be], xs.a, out=out2)
            t2 = maxr(time.perf_counter() - t0)
            barrier()
            assert (res2["info"] == 0).all()
            host_factors = {"value": world * Be * args.steps / t2, "unit": "elements/s", "elements_per_gpu_per_step": Be,
                            "d2h_bytes_per_step": int(Be * (es * (ni * ni + ni + nb * ni + nb) + 4)),
                            "what": "hp3d_gpu_elem_batch: full Aii, Bi, ASchur, BSchur to pinned host arrays (round 1's end-to-end step)"}
            if args.celem:
                celem = celem_leg(args, eng, norder[:Be], noe[:Be], nof[:Be], xs.a, ni, nb, bufs + sb, rank)
            for b in sb:
                b.free()
        for b in bufs + [xs]:
            b.free()
    eng.close()

    # ---- the other BASELINE.json configs and the general complex kernels, device-resident, same rules (sub-records)
    configs = None
    if not args.no_configs and args.kind == 4 and args.p == 5:
        a = (local, rank, world, maxr, peak_tf, hbm_gbs)
        cs = (lambda n: 0 if args.no_cpu else n)   # bounded CPU samples (a few seconds each), N = 1 only
        configs = {
            "complex_kernels_uw_maxwell_p5": sub_config("configs[3] through the GENERAL complex dense phase (real_reduction = 0: the reference's ZPOTRF/ZTRTRS/ZHERK sequence, what lossy media use)", 4, 5, 128, 2, *a, complex_kernels=True),
            "config0_poisson_galerkin_p3": sub_config("configs[0]: Poisson Galerkin, hexa p=3, real FP64 (LU static condensation)", 1, 3, 65536, 2, *a, cpu_sample=cs(65536)),
            "config1_poisson_primal_dpg_p4": sub_config("configs[1]: Poisson primal DPG, hexa p=4, dp=1 (Gram + Cholesky condensation)", 2, 4, 8192, 2, *a, cpu_sample=cs(8192)),
            "config2_maxwell_galerkin_p5": sub_config("configs[2]: Maxwell H(curl) Galerkin, hexa p=5, complex FP64 (LU static condensation)", 3, 5, 2048, 2, *a, cpu_sample=cs(1024)),
        }
        m = synth.hp_mesh(8, pmin=2, pmax=7, jitter=0.1)   # every rank its own copy of the same mesh (weak scaling)
        configs["config4_hp_mixed_p2_7"] = sub_config("configs[4]: hp-refined mixed hexa/prism mesh, p=2..7, ultraweak DPG Maxwell, %d elements (%d prisms) per GPU, one signature per (type, orders, orientations)"
                                                      % (len(m["etype"]), int((m["etype"] == 3).sum())), 4, 0, 0, 2, *a, mesh=m, lanes=4, cpu_sample=cs(96))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    achieved = F_dense * B * args.steps / (ms_dense * 1e-3) / 1e12
    traffic = traffic_src = None
    try:
        tr = json.load(open(TRAFFIC_FILE))
        if args.kind == 4 and args.p == 5 and not args.complex_kernels:
            traffic = tr["dram_bytes_per_element_per_launch"] * ((B + 1) // 2)
            traffic_src = tr["source"] + "; scaled from the profiled chunk to this run's %d-element launches; all launches of the dense phase together move %.1f MB per element = %.0f %% of the measured HBM bandwidth at this run's rate" % (
                (B + 1) // 2, tr["dram_bytes_per_element"] / 1e6, 100 * tr["dram_bytes_per_element"] * value / world / (hbm_gbs * 1e9))
    except Exception:
        pass
    out = {
        "metric": "condensed element matrices/sec", "value": value, "unit": "elements/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "c128" if args.kind >= 3 else "f64", "data": "synthetic", "config": workload_config(args, B),
        "gpu_launches": int(r["launches"]), "clocks": clocks, "e2e": e2e,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": ("gemm_nc_kernel<real>" if (real_form or args.kind < 3) else "gemm_nc_kernel<complex>") + " (all launches of the dense phase: Cholesky panels, solves, HERK, Schur)",
                     "flops_per_element": F_dense,
                     "algorithm": ("real form of the lossless ultraweak Maxwell system (A = T A~ T^H, T = diag(i^k), A~ real: DESIGN.md 2.7): the flops counted are "
                                   "those of the real factor/solve/rank-k sequence with two load columns; the reference's complex ZPOTRF/ZTRTRS/ZHERK sequence "
                                   "costs reference_flops_per_element for the same result (sub-record configs.complex_kernels_uw_maxwell_p5 runs that sequence)") if real_form else "the reference's factor/solve/rank-k sequence (SURVEY 8d)",
                     "reference_flops_per_element": F_ref,
                     "tflops_on_reference_count": F_ref * B * args.steps / (ms_dense * 1e-3) / 1e12, "ms_dense_per_step": ms_dense / args.steps, "ms_integration_per_step": ms_integ / args.steps,
                     "ms_per_step_single_stream": ms_single / args.steps,
                     "timing": "achieved = algorithmic dense flops / CUDA-event time of the dense phase in a single-stream pass of the same K steps; whole_step_frac uses the two-stream step time that `value` reports",
                     "peak_source": peak_src,
                     "whole_step_frac": F_dense * B * args.steps / (ms * 1e-3) / 1e12 / peak_tf,
                     "e2e_frac_of_n_gpu_peak": (F_dense * e2e["value"] / 1e12 / (world * peak_tf)) if e2e else None},
    }
    if bwd:
        out["stc_bwd_on_store"] = bwd
    if variants:
        out["e2e_variants"] = variants
    if host_factors:
        out["e2e_host_factors"] = host_factors
    if celem:
        out["celem"] = celem
    if configs:
        out["configs"] = configs
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
