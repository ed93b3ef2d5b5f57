"""Loader of the C-ABI shared library (hp3d_b200/libhp3d_gpu.so).

There is deliberately NO fallback: if the CUDA library is missing or cannot be loaded the import of any
compute entry point raises.  Build it with `python -c "import __graft_entry__ as g; g.build()"`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhp3d_gpu.so")
_LIB = None


class Params(C.Structure):
    """Mirror of `hp3d_params` (include/hp3d_gpu.h)."""
    _fields_ = [("nord_add", C.c_int), ("maxp", C.c_int), ("test_norm", C.c_int), ("alpha_norm", C.c_double),
                ("omega", C.c_double), ("eps", C.c_double), ("mu", C.c_double), ("sigma", C.c_double),
                ("eps_tensor", C.c_double * 18), ("source", C.c_int), ("icomp_exact", C.c_int),
                ("store_schur", C.c_int), ("real_reduction", C.c_int), ("aii_packed", C.c_int), ("nr_rhs", C.c_int)]


# every symbol include/hp3d_gpu.h declares (tests check that the library exports all of them)
EXPORTS = [
    "hp3d_gpu_params_default", "hp3d_gpu_init", "hp3d_gpu_finalize", "hp3d_gpu_last_error", "hp3d_gpu_plan",
    "hp3d_gpu_plan_destroy", "hp3d_gpu_sizes", "hp3d_gpu_elem_batch", "hp3d_gpu_quad_points",
    "hp3d_gpu_stc_bwd_batch", "hp3d_gpu_bench", "hp3d_gpu_dense_debug", "hp3d_gpu_host_alloc", "hp3d_gpu_host_free",
    "hp3d_gpu_set_chunk", "hp3d_gpu_dof_map", "hp3d_gpu_tables_1d", "hp3d_gpu_integrate_debug",
    "hp3d_gpu_sizes_t", "hp3d_gpu_bench_t", "hp3d_gpu_integrate_debug_t", "hp3d_gpu_prism_shape", "hp3d_gpu_sig_dims", "hp3d_gpu_elem_bwd_batch", "hp3d_gpu_elem_residual_batch",
    "hp3d_gpu_physics_default", "hp3d_gpu_celem_pack", "hp3d_gpu_celem_batch",
    "hp3d_gpu_elem_error_batch", "hp3d_gpu_error_points", "hp3d_gpu_chunk_plan_debug",
    "hp3d_gpu_pbi_points", "hp3d_gpu_pbi_h1_batch", "hp3d_gpu_pbi_hcurl_points", "hp3d_gpu_pbi_hcurl_batch", "hp3d_gpu_pbi_hdiv_points", "hp3d_gpu_pbi_hdiv_batch", "hp3d_gpu_pbi_cache_limit",
    "hp3d_gpu_cloc_create", "hp3d_gpu_cloc_clear", "hp3d_gpu_cloc_destroy", "hp3d_gpu_cloc_stats", "hp3d_gpu_elem_batch_cloc",
    "hp3d_gpu_celem_batch_cloc", "hp3d_gpu_cloc_bwd_batch", "hp3d_gpu_cloc_fetch", "hp3d_gpu_hermitian_unpack_batch", "hp3d_gpu_fp64_peak_probe",
]


def build(verbose=False):
    """Compile the library for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    import subprocess
    src = os.path.join(_HERE, "csrc", "hp3d_gpu.cu")
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC,-pthread", "-o", LIB_PATH, src]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return LIB_PATH


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    csrc = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc)] + [os.path.join(_HERE, "..", "include", "hp3d_gpu.h")]
    return any(os.path.getmtime(s) > t for s in srcs)


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: the CUDA extension has not been built (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        L.hp3d_gpu_last_error.restype = C.c_char_p
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        raise RuntimeError(f"hp3d_gpu error {rc}: {lib().hp3d_gpu_last_error().decode()}")
