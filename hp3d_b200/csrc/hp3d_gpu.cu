// hp3d_gpu.cu -- C-ABI entry points of the B200 element engine (see include/hp3d_gpu.h).
#include "../../include/hp3d_gpu.h"
#include "engine.cuh"

#include <cstdarg>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

using namespace hp3d;

namespace {
std::mutex g_mu;
std::string g_err;
int g_device = -1;
std::vector<Plan *> g_plans;

int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CUDA_TRY(x)                                                                     \
  do {                                                                                  \
    cudaError_t e_ = (x);                                                               \
    if (e_ != cudaSuccess) return fail(HP3D_ENODEV, "%s: %s", #x, cudaGetErrorString(e_)); \
  } while (0)
}  // namespace

extern "C" {

const char *hp3d_gpu_last_error(void) { return g_err.c_str(); }

void hp3d_gpu_params_default(hp3d_params *p) {
  memset(p, 0, sizeof *p);
  p->nord_add = 1; p->maxp = 6; p->test_norm = HP3D_GRAPH_NORM; p->alpha_norm = 1.0;
  p->omega = 1.0; p->eps = 1.0; p->mu = 1.0; p->sigma = 0.0;
  p->eps_tensor[0] = p->eps_tensor[8] = p->eps_tensor[16] = 1.0;
  p->source = HP3D_SRC_SIN; p->icomp_exact = 1; p->store_schur = 1;
}

int hp3d_gpu_init(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) return fail(HP3D_ENODEV, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(HP3D_EINVAL, "device %d out of range (0..%d)", device, n - 1);
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(HP3D_ENODEV, "device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor);
  CUDA_TRY(dense_configure<true>());
  CUDA_TRY(dense_configure<false>());
  g_device = device;
  return HP3D_OK;
}

int hp3d_gpu_finalize(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (Plan *p : g_plans) delete p;
  g_plans.clear();
  g_device = -1;
  return HP3D_OK;
}

int hp3d_gpu_dense_debug(int cplx, int nel, int n, int nb, int ni, const void *G, const void *Bm, void *Aii, void *Bi,
                         void *ASchur, void *BSchur, int *info) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_device < 0) return fail(HP3D_ENODEV, "hp3d_gpu_init has not been called");
  if (nel <= 0 || n <= 0 || nb < 0 || ni <= 0) return fail(HP3D_EINVAL, "bad sizes");
  std::string err;
  int rc = cplx ? dense_debug_run<true>(nel, n, nb, ni, G, Bm, Aii, Bi, ASchur, BSchur, info, err)
                : dense_debug_run<false>(nel, n, nb, ni, G, Bm, Aii, Bi, ASchur, BSchur, info, err);
  if (rc) return fail(rc, "%s", err.c_str());
  return HP3D_OK;
}

}  // extern "C"
