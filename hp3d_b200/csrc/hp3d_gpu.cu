// hp3d_gpu.cu -- C-ABI entry points of the B200 element engine (see include/hp3d_gpu.h).
#include "../../include/hp3d_gpu.h"
#include "engine.cuh"
#include "error_eval.cuh"
#include "pbi.cuh"

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <sched.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <cstdarg>
#include <deque>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

using namespace hp3d;

namespace {
// Locking: g_mu serialises the entry points that touch the DEVICE engine (lanes, arena, streams, signature uploads) and the
// plan table; host-only entry points (celem_pack, dof_map, tables_1d, prism_shape, physics_default, chunk_plan_debug) take
// no lock at all and may run from the reference's OpenMP threads while a batch is in flight; the size queries
// (hp3d_gpu_sizes*, hp3d_gpu_sig_dims) only take the plan's own mutex.  The last error message is per calling thread.
std::recursive_mutex g_mu;
thread_local std::string g_err;
int g_device = -1;
std::vector<Plan *> g_plans;
cudaStream_t g_lane_stream[LaneSet::NLANE] = {nullptr, nullptr, nullptr, nullptr}, g_copy = nullptr;
#define g_compute g_lane_stream[0]
int g_max_chunk = 0;

struct CelemStore {   // grow-only device buffers for the constraint arrays of hp3d_gpu_celem_batch
  void *p[10] = {nullptr}; size_t cap[10] = {0};
  void *get(int i, size_t bytes) {
    if (bytes > cap[i]) {
      cudaDeviceSynchronize();
      cudaFree(p[i]); p[i] = nullptr; cap[i] = 0;
      const size_t want = bytes + bytes / 4;
      if (cudaMalloc(&p[i], want) != cudaSuccess) return nullptr;
      cap[i] = want;
    }
    return p[i];
  }
  void release() { for (int i = 0; i < 10; i++) { cudaFree(p[i]); p[i] = nullptr; cap[i] = 0; } }
} g_celem_store;

void release_error_signatures();   // error-evaluation tables (defined with hp3d_gpu_elem_error_batch below)
void release_pbi_signatures();     // interpolation tables (defined with hp3d_gpu_pbi_h1_batch below)

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

void pool_shutdown();           // host mirror threads (aii_packed = 2), defined with HostPool below
void release_clocs(int plan);   // device-resident Schur stores of one plan (-1: all), defined with ClocStore below
std::mutex g_plans_mu;   // the plan table and the lifetime of its entries (held briefly; never while waiting for the device)
Plan *plan_of(int id) {
  std::lock_guard<std::mutex> lk(g_plans_mu);
  return (id >= 0 && id < (int)g_plans.size()) ? g_plans[id] : nullptr;
}
// CUDA's current device is per host thread: every device entry point may be called from any of the reference's OpenMP threads
int enter_device() {
  if (g_device < 0) return fail(HP3D_ENODEV, "hp3d_gpu_init has not been called");
  cudaError_t e = cudaSetDevice(g_device);
  if (e != cudaSuccess) return fail(HP3D_ENODEV, "cudaSetDevice(%d): %s", g_device, cudaGetErrorString(e));
  return HP3D_OK;
}

// largest chunk (elements per lane) of this shape that fits in the free device memory
int chunk_capacity(const ChunkShape &sh, int want, int nlanes = 2) {
  size_t fre = 0, tot = 0;
  cudaMemGetInfo(&fre, &tot);
  const size_t per = g_lanes.bytes_per_element(sh);
  double budget = 0.80 * (double)(fre + g_arena.dcap);   // the shared arena is reusable; the lanes share the budget
  long long cap = (long long)(budget / (double)(nlanes * per));
  if (cap > 1024) cap = 1024;
  if (cap > want) cap = want;
  return (int)cap;
}

// Chunk sizes for `ntot` elements of one dense class, at most `cap` per chunk, rotated over NL lanes (max_chunk > 0: plain
// chunks of that size, used by tests and sweeps).
std::vector<size_t> chunk_plan(size_t ntot, int cap, int NL, int max_chunk) {
      std::vector<size_t> sizes;
      if (max_chunk > 0) {
        for (size_t left = ntot; left;) { const size_t n = std::min(left, (size_t)cap); sizes.push_back(n); left -= n; }
      } else {
        // RAMPED start: the lanes share the SMs evenly, so equal first chunks would all finish at the same moment and their
        // result copies would pile up behind the compute; first chunks of cap/NL, 2cap/NL, ... keep the lanes out of phase and
        // D2H streams continuously under the kernels of the other lanes.
        // GEOMETRIC taper: whatever the lanes compute last is copied after the compute has ended, so the last NL chunks are 8
        // elements, the NL before them 16, then 32 (the result copy of an element costs half its compute time).
        std::vector<size_t> head, tail;
        size_t hsum = 0, tsum = 0;
        if (ntot >= (size_t)2 * NL * cap && cap >= 4 * NL)
          for (int k = 0; k < NL; k++) { head.push_back(std::min((size_t)cap, std::max((size_t)8, (size_t)cap * (k + 1) / NL))); hsum += head.back(); }
        const size_t avail = ntot - hsum;
        for (size_t sz : {(size_t)8, (size_t)16, (size_t)32})
          if (sz < (size_t)cap)
            for (int i = 0; i < NL; i++)
              if (tsum + sz <= avail / 2) { tail.push_back(sz); tsum += sz; }
        const size_t body = avail - tsum, nbody = (body + cap - 1) / cap;
        sizes = head;
        for (size_t i = 0, left = body; i < nbody; i++) { const size_t n = (left + (nbody - i) - 1) / (nbody - i); sizes.push_back(n); left -= n; }
        for (size_t i = tail.size(); i-- > 0;) sizes.push_back(tail[i]);
      }
  return sizes;
}

// The elements of one call grouped into dense classes; inside a class sorted by signature so that a chunk is a short
// list of equal-signature segments.
struct ClassGroup {
  ChunkShape shape;
  std::vector<int> el;            // caller element indices, signature-sorted
  std::vector<Signature *> sig;   // signature of el[i]
};
int build_classes(Plan *p, int nel, const int *etype, const int *norder, const int *norie, const int *norif, bool device,
                  std::vector<ClassGroup> &out, std::string &err) {
  std::map<std::string, std::vector<int>> bysig;
  for (int e = 0; e < nel; e++) {
    const int et = etype ? etype[e] : HP3D_MDLB;
    if (et != HP3D_MDLB && et != HP3D_MDLP) { err = "element " + std::to_string(e) + ": element type " + std::to_string(et) + " is not implemented (bricks and prisms are)"; return HP3D_EINVAL; }
    bysig[Plan::key(et, norder + 19 * e, norie + 12 * e, norif + 6 * e)].push_back(e);
  }
  // signatures -> dense classes.  Signatures whose padded extents (np, nbp) agree but whose padded interface extent nip differs
  // (hp meshes: the min rule gives every element its own trace orders) are MERGED into a class of the largest nip as long as
  // the class stays small: a class is a chain of ~100 dependent launches per chunk, so many tiny classes leave the GPU idle,
  // while the zero rows a merged element carries cost a few per cent of flops.
  struct SigGroup { Signature *S; const std::vector<int> *el; };
  std::map<std::string, std::vector<SigGroup>> base;
  {
    std::vector<std::pair<std::string, int>> missing;
    for (auto &g : bysig) if (!p->find(g.first)) missing.emplace_back(g.first, g.second[0]);
    const auto t0 = std::chrono::steady_clock::now();
    if (p->compile_missing(missing, etype, norder, norie, norif, err)) return HP3D_EINVAL;
    if (getenv("HP3D_TRACE") && !missing.empty())
      fprintf(stderr, "[hp3d] compiled %zu signatures in %.1f ms\n", missing.size(),
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  }
  const auto t_up0 = std::chrono::steady_clock::now();
  if (device) {   // tables of all signatures that are not resident yet: one allocation, one copy
    std::vector<Signature *> fresh;
    for (auto &g : bysig) { Signature *S = p->find(g.first); if (S && !S->d_tab) fresh.push_back(S); }
    if (Signature::upload_many(fresh, err)) return HP3D_ENOMEM;
  }
  for (auto &g : bysig) {
    const int e0 = g.second[0];
    Signature *S = p->get(etype ? etype[e0] : HP3D_MDLB, norder + 19 * e0, norie + 12 * e0, norif + 6 * e0, device, err);
    if (!S) { err = "element " + std::to_string(e0) + ": " + err; return HP3D_EINVAL; }
    base[ChunkShape::key(S->h)].push_back(SigGroup{S, &g.second});
  }
  if (getenv("HP3D_TRACE"))
    fprintf(stderr, "[hp3d] signature lookup + upload: %.1f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_up0).count());
  size_t MERGE_TARGET = 16;   // elements (576-element hp mesh, round-2 scheduler with per-lane class binding: 8 -> 1495, 16 -> 1503, 32 -> 1476, 64 -> 1410, 128 -> 1374 elements/s; round 1: 32)
  if (const char *mt = getenv("HP3D_MERGE_TARGET")) MERGE_TARGET = (size_t)atoi(mt);
  for (auto &b : base) {
    std::vector<SigGroup> &v = b.second;
    std::stable_sort(v.begin(), v.end(), [](const SigGroup &a, const SigGroup &c) { return a.S->h.dims.nip < c.S->h.dims.nip; });
    size_t i = 0;
    while (i < v.size()) {
      out.emplace_back();
      ClassGroup &C = out.back();
      size_t cnt = 0;
      int nip_open = -1;
      // take whole nip levels until the class holds MERGE_TARGET elements
      while (i < v.size() && (cnt < MERGE_TARGET || v[i].S->h.dims.nip == nip_open)) {
        nip_open = v[i].S->h.dims.nip;
        C.shape.absorb(v[i].S->h);
        for (int e : *v[i].el) { C.el.push_back(e); C.sig.push_back(v[i].S); }
        cnt += v[i].el->size();
        i++;
      }
    }
  }
  return HP3D_OK;
}
// segments of the chunk [c0, c0+n) of a class
void chunk_segments(const ClassGroup &C, size_t c0, int n, std::vector<Seg> &segs) {
  segs.clear();
  for (int i = 0; i < n;) {
    int j = i + 1;
    while (j < n && C.sig[c0 + j] == C.sig[c0 + i]) j++;
    segs.push_back(Seg{C.sig[c0 + i], i, j - i});
    i = j;
  }
}

}  // namespace

extern "C" {

const char *hp3d_gpu_last_error(void) { return g_err.c_str(); }

void hp3d_gpu_params_default(hp3d_params *p) {
  memset(p, 0, sizeof *p);
  p->nord_add = 1; p->maxp = 6; p->test_norm = HP3D_GRAPH_NORM; p->alpha_norm = 1.0;
  p->omega = 1.0; p->eps = 1.0; p->mu = 1.0; p->sigma = 0.0;
  p->eps_tensor[0] = p->eps_tensor[8] = p->eps_tensor[16] = 1.0;
  p->source = HP3D_SRC_SIN; p->icomp_exact = 1; p->store_schur = 1; p->real_reduction = 1; p->nr_rhs = 1;
}

int hp3d_gpu_init(int device) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) return fail(HP3D_ENODEV, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(HP3D_EINVAL, "device %d out of range (0..%d)", device, n - 1);
  // one device per process (hp3D runs one MPI rank per GPU): streams, workspaces and cached tables belong to the first device
  if (g_device >= 0 && g_device != device)
    return fail(HP3D_EINVAL, "already initialised on device %d: call hp3d_gpu_finalize before selecting device %d (one device per process)", g_device, device);
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(HP3D_ENODEV, "device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor);
  CUDA_TRY(dense_configure<true>());
  CUDA_TRY(dense_configure<false>());
  CUDA_TRY(formats_configure());
  CUDA_TRY(tp3_configure<4>()); CUDA_TRY(tp3_configure<6>()); CUDA_TRY(tp3_configure<8>()); CUDA_TRY(tp3_configure<10>());
  CUDA_TRY(tp2_configure<4>()); CUDA_TRY(tp2_configure<6>()); CUDA_TRY(tp2_configure<8>()); CUDA_TRY(tp2_configure<10>());
  // (descending stream priorities per lane were tried to stagger the chunk completions: 4 % slower device-resident and 8 %
  // slower end to end than equal priorities; the chunk plan's ramped start does the staggering instead)
  for (int i = 0; i < LaneSet::NLANE; i++)
    if (!g_lane_stream[i]) CUDA_TRY(cudaStreamCreateWithFlags(&g_lane_stream[i], cudaStreamNonBlocking));
  if (!g_copy) CUDA_TRY(cudaStreamCreateWithFlags(&g_copy, cudaStreamNonBlocking));
  g_device = device;
  return HP3D_OK;
}

int hp3d_gpu_finalize(void) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (g_device >= 0) cudaSetDevice(g_device);
  {
    std::lock_guard<std::mutex> lp(g_plans_mu);
    for (Plan *p : g_plans) delete p;
    g_plans.clear();
  }
  cudaDeviceSynchronize();
  pool_shutdown();
  release_clocs(-1);
  g_lanes.release();
  g_arena.release();
  g_celem_store.release();
  release_error_signatures();
  release_pbi_signatures();
  for (int i = 0; i < LaneSet::NLANE; i++)
    if (g_lane_stream[i]) { cudaStreamDestroy(g_lane_stream[i]); g_lane_stream[i] = nullptr; }
  if (g_copy) { cudaStreamDestroy(g_copy); g_copy = nullptr; }
  g_device = -1;
  return HP3D_OK;
}

int hp3d_gpu_plan(int problem_kind, const hp3d_params *prm) {
  if (!prm) return fail(HP3D_EINVAL, "null params");
  if (problem_kind < HP3D_POIS_GAL || problem_kind > HP3D_MAXW_UW) return fail(HP3D_EINVAL, "unknown problem kind %d", problem_kind);
  // get_permittivity (problems/MAXWELL/ULTRAWEAK_DPG/common/commonRoutines.F90:126-150; used at elem_opt.F90:260-266): a CONSTANT
  // complex 3x3 tensor for ultraweak Maxwell (real tensors keep the real form, complex ones take the general complex kernels);
  // the other problems of the reference have a scalar permittivity only
  bool tensor = false;
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < 3; i++) {
      double re = prm->eps_tensor[2 * (i + 3 * j)], im = prm->eps_tensor[2 * (i + 3 * j) + 1];
      if (re != (i == j ? 1.0 : 0.0) || im != 0.0) tensor = true;
    }
  if (tensor && problem_kind != HP3D_MAXW_UW) return fail(HP3D_EINVAL, "a permittivity tensor other than the identity is defined for ultraweak Maxwell only (scale with eps)");
  if (tensor && prm->source == HP3D_SRC_SIN)
    return fail(HP3D_EINVAL, "the built-in manufactured source assumes the identity permittivity tensor: pass the source through HP3D_SRC_TABLE");
  if (prm->maxp < 1 || prm->maxp > 9) return fail(HP3D_EINVAL, "maxp out of range");
  if (prm->icomp_exact < 1 || prm->icomp_exact > 3) return fail(HP3D_EINVAL, "icomp_exact out of range");
  Plan *p = new Plan();
  p->fp.kind = problem_kind; p->fp.nord_add = prm->nord_add; p->fp.maxp = prm->maxp; p->fp.test_norm = prm->test_norm;
  p->fp.alpha_norm = prm->alpha_norm; p->fp.omega = prm->omega; p->fp.eps = prm->eps; p->fp.mu = prm->mu; p->fp.sigma = prm->sigma;
  p->fp.source = prm->source; p->fp.icomp = prm->icomp_exact - 1;
  p->store_schur = prm->store_schur;
  p->fp.real_struct = prm->real_reduction != 0;
  p->fp.nrhs = prm->nr_rhs > 0 ? prm->nr_rhs : 1;   // 0 (a zero-initialised struct) means the default
  if (p->fp.nrhs > 1) {
    const char *why = nullptr;
    if (problem_kind != HP3D_POIS_PDPG && problem_kind != HP3D_MAXW_UW) why = "the Cholesky condensation of the DPG problems";
    else if (prm->source == HP3D_SRC_SIN) why = "sources given through HP3D_SRC_TABLE (the built-in manufactured source is one load)";
    else if (p->fp.nrhs > 16) why = "at most 16 load vectors";
    if (why) { const int n = p->fp.nrhs; delete p; return fail(HP3D_EINVAL, "nr_rhs = %d: more than one load vector is implemented for %s", n, why); }
  }
  p->fp.tensor = tensor;
  for (int i = 0; i < 9; i++) p->fp.epst[i] = std::complex<double>(prm->eps_tensor[2 * i], prm->eps_tensor[2 * i + 1]);
  p->aii_packed = prm->aii_packed;
  if (p->aii_packed < 0 || p->aii_packed > 2) { delete p; return fail(HP3D_EINVAL, "aii_packed must be 0, 1 or 2"); }
  if (p->aii_packed && problem_kind != HP3D_POIS_PDPG && problem_kind != HP3D_MAXW_UW) {
    delete p;
    return fail(HP3D_EINVAL, "aii_packed is defined for the Hermitian (DPG) problems only");
  }
  std::lock_guard<std::mutex> lk(g_plans_mu);
  for (size_t i = 0; i < g_plans.size(); i++)
    if (!g_plans[i]) { g_plans[i] = p; return (int)i; }
  g_plans.push_back(p);
  return (int)g_plans.size() - 1;
}

int hp3d_gpu_set_chunk(int max_elements) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (max_elements < 0) return fail(HP3D_EINVAL, "negative chunk size");
  g_max_chunk = max_elements;
  return HP3D_OK;
}

int hp3d_gpu_plan_destroy(int plan) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);   // waits for a batch in flight: it may be using this plan's tables
  Plan *p = nullptr;
  {
    std::lock_guard<std::mutex> lp(g_plans_mu);
    if (plan >= 0 && plan < (int)g_plans.size()) { p = g_plans[plan]; g_plans[plan] = nullptr; }
  }
  if (!p) return fail(HP3D_ENOPLAN, "no such plan %d", plan);
  if (g_device >= 0) cudaSetDevice(g_device);
  release_clocs(plan);
  delete p;
  return HP3D_OK;
}

int hp3d_gpu_sizes(int plan, const int *norder, int *ni, int *nb, int *nint, int *nrdofH) {
  return hp3d_gpu_sizes_t(plan, HP3D_MDLB, norder, ni, nb, nint, nrdofH);
}

int hp3d_gpu_sizes_t(int plan, int etype, const int *norder, int *ni, int *nb, int *nint, int *nrdofH) {
  std::lock_guard<std::mutex> lp(g_plans_mu);   // host only: keeps the plan alive, does not wait for the device engine
  Plan *p = (plan >= 0 && plan < (int)g_plans.size()) ? g_plans[plan] : nullptr;
  if (!p) return fail(HP3D_ENOPLAN, "no such plan %d", plan);
  const int z12[12] = {0}, z6[6] = {0};
  std::string err;
  SigHost h;
  if (!p->sizes(etype, norder, z12, z6, h, err)) return fail(HP3D_EINVAL, "%s", err.c_str());
  if (ni) *ni = h.ni;
  if (nb) *nb = h.nb;
  if (nint) *nint = h.nint;
  if (nrdofH) *nrdofH = h.nH;
  return HP3D_OK;
}

int hp3d_gpu_sig_dims(int plan, int etype, const int *norder, const int *norie, const int *norif, int *dims) {
  std::lock_guard<std::mutex> lp(g_plans_mu);
  Plan *p = (plan >= 0 && plan < (int)g_plans.size()) ? g_plans[plan] : nullptr;
  if (!p) return fail(HP3D_ENOPLAN, "no such plan %d", plan);
  std::string err;
  SigHost h;
  if (!p->sizes(etype, norder, norie, norif, h, err)) return fail(HP3D_EINVAL, "%s", err.c_str());
  dims[0] = h.ntest; dims[1] = h.ni; dims[2] = h.nb; dims[3] = h.nint; dims[4] = h.nH; dims[5] = h.dims.np; dims[6] = h.dims.nbp; dims[7] = h.dims.nip;
  return HP3D_OK;
}

int hp3d_gpu_dof_map(int space, const int *norder, const int *norie, const int *norif, int cap, int *fam, int *idx, int *sgn) {
  std::vector<TensorDof> d;
  switch (space) {
    case 0: d = hexa_dofs_H1(norder, norie, norif); break;
    case 1: d = hexa_dofs_Hcurl(norder, norie, norif); break;
    case 2: d = hexa_dofs_Hdiv(norder, norif); break;
    case 3: d = hexa_dofs_L2(norder); break;
    default: return fail(HP3D_EINVAL, "unknown space %d", space);
  }
  const int n = (int)d.size();
  if (!fam && !idx && !sgn) return n;
  if (cap < n) return fail(HP3D_EINVAL, "dof_map: capacity %d < %d", cap, n);
  for (int k = 0; k < n; k++) {
    if (fam) fam[k] = d[k].fam;
    if (sgn) sgn[k] = d[k].sgn;
    if (idx) for (int a = 0; a < 3; a++) idx[3 * k + a] = d[k].idx[a];
  }
  return n;
}

// pointwise values of the prism shape functions through the T x Z decomposition (host only; parity tests)
int hp3d_gpu_prism_shape(int space, const int *norder, const int *norie, const int *norif, const double *xi, int cap, double *val, double *der) {
  using namespace hp3d::detail;
  TriList T0, T1;
  std::vector<PrismDof> d;
  switch (space) {
    case 0: d = prism_dofs_H1(norder, norie, norif, T0); break;
    case 1: d = prism_dofs_Hcurl(norder, norie, norif, T0, T1); break;
    case 2: d = prism_dofs_Hdiv_faces(norder, norif, T0, T1); break;
    case 3: d = prism_dofs_L2(norder, T0); break;
    default: return fail(HP3D_EINVAL, "unknown space %d", space);
  }
  const int n = (int)d.size();
  if (!val) return n;
  if (cap < n) return fail(HP3D_EINVAL, "prism_shape: capacity %d < %d", cap, n);
  const TriVals v0 = eval_list(T0, xi[0], xi[1]), v1 = eval_list(T1, xi[0], xi[1]);
  const ZVals Z = eval_z(MAXN1D - 1, xi[2]);
  for (int k = 0; k < n; k++) {
    const PrismDof &q = d[k];
    const double *t = (q.list == 0 ? v0 : v1).at(q.t), s = q.sgn;
    double *V = val + 3 * k, *D = der ? der + 3 * k : nullptr;
    if (space == 0) {          // val[0] = value ; der = gradient
      V[0] = s * t[0] * Z.H[q.zi]; V[1] = V[2] = 0.0;
      if (D) { D[0] = s * t[1] * Z.H[q.zi]; D[1] = s * t[2] * Z.H[q.zi]; D[2] = s * t[0] * Z.dH[q.zi]; }
    } else if (space == 1) {   // val = E ; der = curl E
      if (q.list == 0) { V[0] = s * t[0] * Z.H[q.zi]; V[1] = s * t[1] * Z.H[q.zi]; V[2] = 0.0;
        if (D) { D[0] = -s * t[1] * Z.dH[q.zi]; D[1] = s * t[0] * Z.dH[q.zi]; D[2] = s * t[2] * Z.H[q.zi]; } }
      else { V[0] = V[1] = 0.0; V[2] = s * t[0] * Z.Q[q.zi];
        if (D) { D[0] = s * t[2] * Z.Q[q.zi]; D[1] = -s * t[1] * Z.Q[q.zi]; D[2] = 0.0; } }
    } else if (space == 2) {   // val = V (face functions only) ; der[0] = div V
      if (q.list == 0) { V[0] = V[1] = 0.0; V[2] = s * t[0] * Z.H[q.zi]; if (D) { D[0] = s * t[0] * Z.dH[q.zi]; D[1] = D[2] = 0.0; } }
      else { V[0] = s * t[1] * Z.Q[q.zi]; V[1] = -s * t[0] * Z.Q[q.zi]; V[2] = 0.0; if (D) { D[0] = s * t[2] * Z.Q[q.zi]; D[1] = D[2] = 0.0; } }
    } else {
      V[0] = s * t[0] * Z.Q[q.zi]; V[1] = V[2] = 0.0;
      if (D) D[0] = D[1] = D[2] = 0.0;
    }
  }
  return n;
}

int hp3d_gpu_tables_1d(int p, int nq, double *x, double *w, double *H, double *dH, double *Q) {
  if (p < 1 || p > MAXN1D - 1 || nq < 1 || nq > MAXN1D) return fail(HP3D_EINVAL, "tables_1d: p=%d nq=%d out of range", p, nq);
  Tables1D t = make_tables(p, nq);
  if (x) memcpy(x, t.x.data(), sizeof(double) * nq);
  if (w) memcpy(w, t.w.data(), sizeof(double) * nq);
  if (H) memcpy(H, t.H.data(), sizeof(double) * (p + 1) * nq);
  if (dH) memcpy(dH, t.dH.data(), sizeof(double) * (p + 1) * nq);
  if (Q) memcpy(Q, t.Q.data(), sizeof(double) * p * nq);
  return HP3D_OK;
}

void *hp3d_gpu_host_alloc(long long bytes) {
  void *p = nullptr;
  if (bytes <= 0 || cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) return nullptr;
  return p;
}
void hp3d_gpu_host_free(void *p) { if (p) cudaFreeHost(p); }

// ------------------------------------------------------------------------------------------------
}  // extern "C"

namespace {
// ------------------------------------------------------------------------------------------------
// Device-resident CLOC: the back-substitution factors of stc_fwd_wrapper, CLOC(iel)%ASchur (nb x ni) and %BSchur (nb)
// (src/modules/stc.F90:45-58,273-277), kept in HBM under the caller's element index instead of travelling to the host
// (7.2 of the 13.0 MB an ultraweak Maxwell p=5 element produces); hp3d_gpu_cloc_bwd_batch is stc_bwd (stc.F90:661-677) on them.
// Storage: one slab per hp3d_gpu_elem_batch_cloc call ([all ASchur blocks | all BSchur blocks] of the elements that did not
// have a slot of the right size yet), column-major complex(8)/real(8) blocks exactly as the host arrays would hold them.
// Elements that do not fit under the store's byte limit are SPILLED: their descriptors are kept on the host and stc_bwd
// recomputes them through the MODE_BWD pipeline (the "recompute" option the reference leaves unimplemented, stc.F90:279-281).
struct ClocStore {
  int plan = -1;
  bool cplx = false;
  int nrhs = 1;   // NR_RHS of the plan: BSchur blocks hold nb x nrhs values
  size_t limit = 0, bytes = 0;
  struct Slot { double *AS, *BS; int ni, nb; };
  struct Spill { int etype; int norder[19], norie[12], norif[6]; std::vector<double> xnod; std::vector<double> src; };
  std::unordered_map<long long, Slot> slots;
  std::unordered_map<long long, Spill> spilled;
  std::vector<void *> slabs;
  void release() { for (void *q : slabs) cudaFree(q); slabs.clear(); slots.clear(); spilled.clear(); bytes = 0; }
};
std::vector<ClocStore *> g_clocs;
ClocStore *cloc_of(int id) { return (id >= 0 && id < (int)g_clocs.size()) ? g_clocs[id] : nullptr; }
void release_clocs(int plan) {
  for (ClocStore *&c : g_clocs)
    if (c && (plan < 0 || c->plan == plan)) { cudaDeviceSynchronize(); c->release(); delete c; c = nullptr; }
}

struct EventSet {   // the events of one pipeline call; destroyed on every exit path
  std::vector<cudaEvent_t> ev;
  cudaError_t add(cudaEvent_t *e) { cudaError_t rc = cudaEventCreateWithFlags(e, cudaEventDisableTiming); if (rc == cudaSuccess) ev.push_back(*e); return rc; }
  ~EventSet() { for (cudaEvent_t e : ev) cudaEventDestroy(e); }
};

// ------------------------------------------------------------------------------------------------
// hp3d_params.aii_packed = 2: a Hermitian Aii crosses PCIe as the block-trapezoids of its lower triangle (block columns of
// TRAP_W columns, rows from the block's first row down: 55 % of the matrix at ni = 600), placed by the copy engine straight
// into their final position in the caller's full ni x ni block; the strictly-upper blocks are the conjugate transposes and are
// written by these host threads while the device works on the next chunks.  The caller sees the full matrix, as before.
constexpr int TRAP_W = 64;
struct MirrorTask { double *a; int n; bool cplx; };
static void mirror_upper(const MirrorTask &t) {
  const int n = t.n;
  constexpr int T = 16;
  const bool nt = t.cplx && ((uintptr_t)t.a & 15) == 0;
  (void)nt;
  // source tiles: rows [R, R+T) x cols [C, C+T) strictly below the TRAP_W block diagonal; destination (C.., R..) = conj transpose
  for (int cb = 0; cb < n; cb += TRAP_W) {
    const int ce = std::min(n, cb + TRAP_W);
    for (int R = ce; R < n; R += T) {
      const int Re = std::min(n, R + T);
      for (int C = cb; C < ce; C += T) {
        const int Ce = std::min(ce, C + T);
        if (t.cplx) {
          for (int i = R; i < Re; i++) {           // destination column i, rows C..Ce (contiguous)
            double *dst = t.a + 2 * ((size_t)i * n + C);
            const double *src = t.a + 2 * ((size_t)C * n + i);
#if defined(__SSE2__)
            if (nt) {   // streaming stores: the upper triangle is written once and not read here (no read-for-ownership traffic)
              const __m128d conj = _mm_set_pd(-0.0, 0.0);
              for (int j = 0; j < Ce - C; j++) _mm_stream_pd(dst + 2 * j, _mm_xor_pd(_mm_loadu_pd(src + 2 * (size_t)j * n), conj));
              continue;
            }
#endif
            for (int j = 0; j < Ce - C; j++) { dst[2 * j] = src[2 * (size_t)j * n]; dst[2 * j + 1] = -src[2 * (size_t)j * n + 1]; }
          }
        } else {
          for (int i = R; i < Re; i++) {
            double *dst = t.a + (size_t)i * n + C;
            const double *src = t.a + (size_t)C * n + i;
            for (int j = 0; j < Ce - C; j++) dst[j] = src[(size_t)j * n];
          }
        }
      }
    }
  }
#if defined(__SSE2__)
  if (nt) _mm_sfence();
#endif
}
struct HostPool {
  std::vector<std::thread> th;
  std::mutex m;
  std::condition_variable cv, cv_done;
  std::deque<MirrorTask> q;
  size_t pending = 0;
  bool stop = false;
  void start() {
    if (!th.empty()) return;
    // default: the CPUs this process may run on (MPI ranks are usually bound to their share of the node), at most 8, one left
    // for the submitting thread
    int n = 8;
    cpu_set_t cs;
    if (sched_getaffinity(0, sizeof cs, &cs) == 0) n = std::min(8, std::max(1, CPU_COUNT(&cs) - 1));
    if (const char *e = getenv("HP3D_HOST_THREADS")) n = std::max(1, atoi(e));
    for (int i = 0; i < n; i++)
      th.emplace_back([this] {
        for (;;) {
          MirrorTask t;
          {
            std::unique_lock<std::mutex> lk(m);
            cv.wait(lk, [this] { return stop || !q.empty(); });
            if (q.empty()) return;
            t = q.front(); q.pop_front();
          }
          mirror_upper(t);
          {
            std::lock_guard<std::mutex> lk(m);
            if (--pending == 0) cv_done.notify_all();
          }
        }
      });
  }
  void push(const MirrorTask &t) {
    { std::lock_guard<std::mutex> lk(m); q.push_back(t); pending++; }
    cv.notify_one();
  }
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    cv_done.wait(lk, [this] { return pending == 0; });
  }
  void shutdown() {
    { std::lock_guard<std::mutex> lk(m); stop = true; }
    cv.notify_all();
    for (std::thread &t : th) t.join();
    th.clear(); stop = false;
  }
  ~HostPool() { shutdown(); }
};
HostPool g_pool;
void pool_shutdown() { g_pool.shutdown(); }

// The chunked, multi-lane pipeline behind hp3d_gpu_elem_batch (MODE_ELEM), hp3d_gpu_elem_bwd_batch (MODE_BWD: recompute the
// element, return only xb = BSchur - ASchur xi) and hp3d_gpu_elem_residual_batch (MODE_RESID: DPG residual per element).
// cloc != nullptr: the Schur factors of element e go to the device-resident store under the index iel[e] (e if iel == nullptr).
int batch_impl(int mode, int plan, int nel, const int *etype, const int *norder, const int *norie, const int *norif,
               const double *xnod, int xnod_ld, const void *source_qp, long long source_ld, void *Aii, long long sAii,
               void *Bi, long long sBi, void *ASchur, long long sAS, void *BSchur, long long sBS, int *ni_out, int *nb_out,
               int *info, const void *xi, long long sxi, void *xb, long long sxb, double *resid, const CelemCall *cc = nullptr,
               ClocStore *cloc = nullptr, const long long *iel = nullptr) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (int drc = enter_device()) return drc;
  Plan *p = plan_of(plan);
  if (!p) return fail(HP3D_ENOPLAN, "no such plan %d", plan);
  const long long nrhs = p->fp.nrhs;
  if (nrhs > 1 && mode != MODE_ELEM && mode != MODE_BWD) return fail(HP3D_EINVAL, "nr_rhs = %lld: this entry point carries one load vector", nrhs);
  if (nel < 0 || !norder || !norie || !norif || !xnod) return fail(HP3D_EINVAL, "null argument");
  if (mode == MODE_ELEM && (!Aii || !Bi)) return fail(HP3D_EINVAL, "null argument");
  if (mode == MODE_CELEM && !cc) return fail(HP3D_EINVAL, "null argument");
  const bool big = mode == MODE_ELEM || mode == MODE_CELEM;   // results go straight from the device staging into the caller's arrays
  if (mode == MODE_BWD && (!xi || !xb)) return fail(HP3D_EINVAL, "null argument");
  if (mode == MODE_RESID && (!xi || !resid)) return fail(HP3D_EINVAL, "null argument");
  if (mode == MODE_RESID && p->fp.kind != HP3D_POIS_PDPG && p->fp.kind != HP3D_MAXW_UW) return fail(HP3D_EINVAL, "the element residual is defined for the DPG problems only");
  if (p->fp.source == HP3D_SRC_TABLE && !source_qp) return fail(HP3D_EINVAL, "source == HP3D_SRC_TABLE needs source_qp");
  if (cloc && !big) return fail(HP3D_EINVAL, "the device-resident Schur store is filled by the element / celem calls only");

  const bool to_host_schur = big && !cloc && p->store_schur && ASchur && BSchur;
  const bool want_schur = mode == MODE_BWD || to_host_schur || cloc != nullptr;
  const bool packed = mode == MODE_ELEM && p->aii_packed == 1;
  const bool trap = mode == MODE_ELEM && p->aii_packed == 2;   // lower block-trapezoids over PCIe, upper triangle mirrored on the host
  std::vector<ClassGroup> classes;
  std::string err;
  if (int brc = build_classes(p, nel, etype, norder, norie, norif, true, classes, err)) return fail(brc, "%s", err.c_str());
  if (mode == MODE_CELEM)
    for (ClassGroup &C : classes) {
      for (int e : C.el) {
        C.shape.nz_max = std::max(C.shape.nz_max, (size_t)cc->nz(e));
        C.shape.nc_max = std::max(C.shape.nc_max, (size_t)(cc->xptr[e + 1] - cc->xptr[e]));
      }
      C.shape.coo = cc->irn != nullptr;
    }
  // ---- every caller stride is checked BEFORE anything is queued (a short stride would read / write outside the caller's arrays)
  for (const ClassGroup &C : classes) {
    const ChunkShape &sh = C.shape;
    const long long ni = sh.d.ni, nb = sh.d.nb;
    if (xnod_ld < 3 * sh.nH_max) return fail(HP3D_EINVAL, "xnod_ld=%d < 3*nrdofH=%d", xnod_ld, 3 * sh.nH_max);
    if (p->fp.source == HP3D_SRC_TABLE && source_ld < (long long)sh.src_max)
      return fail(HP3D_EINVAL, "source_ld=%lld < %zu doubles (nint x %d values per point)", source_ld, sh.src_max, sh.d.cplx || sh.d.rs ? 6 : 1);
    if (mode == MODE_ELEM) {
      const long long need = packed ? ni * (ni + 1) / 2 : ni * ni;
      if (sAii < need) return fail(HP3D_EINVAL, "Aii stride %lld < %lld scalars (%s ni = %lld)", sAii, need, packed ? "packed triangle of" : "ni^2,", ni);
      if (sBi < ni * nrhs) return fail(HP3D_EINVAL, "Bi stride %lld < ni * nr_rhs = %lld", sBi, ni * nrhs);
    }
    if (to_host_schur && nb > 0) {
      if (sAS < nb * ni) return fail(HP3D_EINVAL, "ASchur stride %lld < nb*ni = %lld", sAS, nb * ni);
      if (sBS < nb * nrhs) return fail(HP3D_EINVAL, "BSchur stride %lld < nb * nr_rhs = %lld", sBS, nb * nrhs);
    }
    if (!big) {
      if (sxi < ni * nrhs) return fail(HP3D_EINVAL, "xi stride %lld < ni * nr_rhs = %lld", sxi, ni * nrhs);
      if (xb && sxb < nb * nrhs) return fail(HP3D_EINVAL, "xb stride %lld < nb * nr_rhs = %lld", sxb, nb * nrhs);
      if (mode == MODE_RESID && nb > 0 && !xb) return fail(HP3D_EINVAL, "residual: xb (bubble dofs) is required for elements with bubbles");
    }
    if (mode == MODE_RESID && sizeof(double) * 2 * (size_t)sh.d.M() > (size_t)RESID_SMEM_MAX)
      return fail(HP3D_EINVAL, "element residual: %d padded trial dofs exceed the kernel's shared-memory vector (%d)", sh.d.M(), RESID_SMEM_MAX / 16);
  }
  // ---- device-resident Schur store: a slot per element (reused when the sizes agree), new ones carved from one slab
  const size_t esz = sizeof(double) * ((p->fp.kind >= HP3D_MAXW_GAL) ? 2 : 1);
  std::vector<ClocStore::Slot> eslot;   // per caller element; AS == nullptr: spilled
  if (cloc) {
    if (cloc->plan != plan) return fail(HP3D_EINVAL, "this Schur store belongs to plan %d", cloc->plan);
    eslot.assign(nel, ClocStore::Slot{nullptr, nullptr, 0, 0});
    std::vector<int> fresh;
    size_t needA = 0, needB = 0;
    for (const ClassGroup &C : classes)
      for (size_t i = 0; i < C.el.size(); i++) {
        const int e = C.el[i];
        const SigHost &h = C.sig[i]->h;
        const long long id = iel ? iel[e] : e;
        cloc->spilled.erase(id);
        auto it = cloc->slots.find(id);
        if (it != cloc->slots.end() && it->second.ni == h.ni && it->second.nb == h.nb) { eslot[e] = it->second; continue; }
        if (it != cloc->slots.end()) cloc->slots.erase(it);   // sizes changed (refinement): the old block stays in its slab until the store is cleared
        eslot[e].ni = h.ni; eslot[e].nb = h.nb;
        fresh.push_back(e);
      }
    // elements are granted in caller order until the limit is reached; the rest is spilled
    std::sort(fresh.begin(), fresh.end());
    size_t granted = 0;
    for (; granted < fresh.size(); granted++) {
      const ClocStore::Slot &sl = eslot[fresh[granted]];
      const size_t a = esz * (size_t)sl.nb * sl.ni, b2 = esz * (size_t)sl.nb * (size_t)nrhs;
      if (cloc->bytes + needA + needB + a + b2 + 512 > cloc->limit) break;
      needA += a; needB += b2;
    }
    if (granted) {
      char *slab = nullptr;
      needA = (needA + 255) & ~(size_t)255;
      cudaError_t ce = cudaMalloc((void **)&slab, needA + needB + 256);
      if (ce != cudaSuccess) { cudaGetLastError(); granted = 0; }   // no room after all: everything new is spilled
      else {
        cloc->slabs.push_back(slab); cloc->bytes += needA + needB + 256;
        size_t oa = 0, ob = needA;
        for (size_t k = 0; k < granted; k++) {
          ClocStore::Slot &sl = eslot[fresh[k]];
          sl.AS = (double *)(slab + oa); sl.BS = (double *)(slab + ob);
          oa += esz * (size_t)sl.nb * sl.ni; ob += esz * (size_t)sl.nb * (size_t)nrhs;
          cloc->slots[iel ? iel[fresh[k]] : fresh[k]] = sl;
        }
      }
    }
    for (size_t k = granted; k < fresh.size(); k++) {   // spilled: keep what stc_bwd needs to recompute the element
      const int e = fresh[k];
      ClocStore::Spill sp;
      sp.etype = etype ? etype[e] : HP3D_MDLB;
      memcpy(sp.norder, norder + 19 * e, sizeof sp.norder); memcpy(sp.norie, norie + 12 * e, sizeof sp.norie); memcpy(sp.norif, norif + 6 * e, sizeof sp.norif);
      sp.xnod.assign(xnod + (size_t)e * xnod_ld, xnod + (size_t)e * xnod_ld + xnod_ld);
      if (p->fp.source == HP3D_SRC_TABLE) sp.src.assign((const double *)source_qp + (size_t)e * source_ld, (const double *)source_qp + (size_t)(e + 1) * source_ld);
      cloc->spilled[iel ? iel[e] : e] = std::move(sp);
    }
  }
  const GeomParams gp = p->geom();
  // slot = (lane, output buffer): chunk k runs on lane k % NL and writes output buffer (k / NL) & 1 of that lane.
  // Four lanes of modest chunks keep the GPU as busy as two lanes of large ones (the latency-bound tile factorizations and
  // the last partial wave of every GEMM launch of one lane are filled by the other lanes) while results stream to the host
  // in smaller pieces.
  constexpr int NL = LaneSet::NLANE, NSLOT = 2 * NL;
  if (trap) g_pool.start();
  cudaEvent_t evCompute[NSLOT], evCopy[NSLOT], evH2D[NL];
  EventSet events;
  for (int i = 0; i < NSLOT; i++) { CUDA_TRY(events.add(&evCompute[i])); CUDA_TRY(events.add(&evCopy[i])); }
  for (int i = 0; i < NL; i++) CUDA_TRY(events.add(&evH2D[i]));
  int rc = HP3D_OK;
  std::vector<Seg> segs;
  const bool trace = getenv("HP3D_TRACE") != nullptr;   // host-side phase times of the call on stderr
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_begin = now();
  double t_wait = 0.0, t_stage = 0.0;
  // ---- chunk plan of the whole call.  Every lane owns a partition of the arena (LaneSet::bind_lane) and is bound to the class
  // of the chunk it runs, so chunks of DIFFERENT dense classes (hp meshes) follow each other on the lanes without draining the
  // device in between; classes go in order of decreasing work.  A chunk is identified by its global number k: lane k % NL,
  // output buffer (k / NL) & 1.
  struct CInfo { ClassGroup *C; int cap; size_t NS, es, nx, nsrc, sA, sB, sS, sT; };
  struct Rec { int ci; size_t c0; int n; const int *h_info; const double *h_xb, *h_res; bool collected; };
  std::vector<CInfo> cinfo;
  std::vector<Rec> recs;
  {
    std::vector<size_t> ord(classes.size());
    for (size_t i = 0; i < ord.size(); i++) ord[i] = i;
    auto work = [&](const ClassGroup &C) { const double n = C.shape.d.np, m = C.shape.d.M(); return (double)C.el.size() * (n * n * n / 3.0 + n * n * m + n * m * m + m * m * m / 3.0 + 1e6); };
    std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return work(classes[a]) > work(classes[b]); });
    size_t part_dev = 0, part_host = 0;
    int need_iota = 1;
    for (size_t oi : ord) {
      ClassGroup &C = classes[oi];
      const ChunkShape &sh = C.shape;
      // chunks of up to 64 elements round-robin over the lanes, with a ramped start and a tapered end (chunk_plan)
      int want = (int)C.el.size();
      if (g_max_chunk > 0) want = std::min(want, g_max_chunk);
      else want = std::min(64, std::max(4, (want + NL - 1) / NL));   // small groups are spread over the lanes
      if (classes.size() > 1 && g_max_chunk <= 0) want = std::min(64, (int)C.el.size());   // several classes: a small class stays whole
      const int cap = chunk_capacity(sh, want, NL);
      if (cap < 1) return fail(HP3D_ENOMEM, "not enough device memory for one element");
      size_t db, hb;
      LaneSet::lane_bytes(sh, cap, db, hb);
      part_dev = std::max(part_dev, db); part_host = std::max(part_host, hb);
      need_iota = std::max(need_iota, std::max(sh.d.ni, sh.d.nb) + 1);
      CInfo ci;
      ci.C = &C; ci.cap = cap; ci.NS = sh.ns(); ci.es = sizeof(double) * ci.NS;
      ci.nx = 3 * (size_t)sh.nH_max; ci.nsrc = sh.src_max;
      ci.sA = (size_t)sh.d.ni * sh.d.ni; ci.sB = (size_t)sh.d.ni * nrhs; ci.sS = (size_t)sh.d.nb * sh.d.ni; ci.sT = (size_t)sh.d.nb * nrhs;   // device staging strides
      const std::vector<size_t> sizes = chunk_plan(C.el.size(), cap, NL, g_max_chunk);
      size_t c0 = 0;
      for (size_t n : sizes) { recs.push_back(Rec{(int)cinfo.size(), c0, (int)n, nullptr, nullptr, nullptr, false}); c0 += n; }
      cinfo.push_back(ci);
    }
    if (LaneSet::ensure_partitions(part_dev, part_host, err) || g_lanes.ensure_iota(need_iota, err)) return fail(HP3D_ENOMEM, "%s", err.c_str());
  }
  const int nrec = (int)recs.size();
  int next_mirror = 0;
  int bound[NL];
  for (int i = 0; i < NL; i++) bound[i] = -1;
  auto mirror_chunk = [&](int k) {   // aii_packed = 2: chunk k's trapezoids are on the host; hand its elements to the mirror threads
    const Rec &R = recs[k];
    const CInfo &I = cinfo[R.ci];
    for (size_t i = R.c0; i < R.c0 + R.n; i++)
      if (I.C->sig[i]->h.ni > TRAP_W) g_pool.push(MirrorTask{(double *)((char *)Aii + I.es * (size_t)sAii * I.C->el[i]), I.C->sig[i]->h.ni, I.NS == 2});
    next_mirror = k + 1;
  };
  auto collect_info = [&](int k) {   // host side of chunk k: wait for its D2H, publish info[] (and the small results)
    Rec &R = recs[k];
    if (R.collected) return;
    R.collected = true;
    const CInfo &I = cinfo[R.ci];
    const int slot = (k % NL) * 2 + ((k / NL) & 1);
    cudaEventSynchronize(evCopy[slot]);
    if (trap) while (next_mirror <= k) mirror_chunk(next_mirror);   // the copy stream is in order: earlier chunks have landed too
    for (int i = 0; i < R.n; i++) {
      const int e = I.C->el[R.c0 + i];
      if (info) info[e] = R.h_info[i];
      if (mode == MODE_BWD) memcpy((char *)xb + I.es * sxb * e, R.h_xb + I.NS * (I.sT + 1) * i, I.es * I.C->sig[R.c0 + i]->h.nb * nrhs);
      if (mode == MODE_RESID) resid[e] = R.h_res[i];
    }
  };
  for (int k = 0; k < nrec; k++) {
    Rec &R = recs[k];
    const CInfo &I = cinfo[R.ci];
    ClassGroup &C = *I.C;
    const std::vector<int> &el = C.el;
    const ChunkShape &sh = C.shape;
    const size_t NS = I.NS, es = I.es, nx = I.nx, nsrc = I.nsrc, sA = I.sA, sB = I.sB, sS = I.sS, sT = I.sT, c0 = R.c0;
    const int n = R.n, ln = k % NL, ob = (k / NL) & 1, slot = ln * 2 + ob, lcap = I.cap;
    cudaStream_t st = g_lane_stream[ln];
    const double tw0 = now();
    if (big) { if (k >= NSLOT) collect_info(k - NSLOT); }   // this slot's previous results are on the host
    else if (k >= NL) collect_info(k - NL);                 // small results are staged per LANE: drain before the lane is reused
    if (bound[ln] != R.ci) {
      // the lane changes class: its buffers move inside its partition, so everything it still has in flight (the other output
      // buffer's copy) must have landed first; the other lanes keep the device busy meanwhile
      if (k >= NL) collect_info(k - NL);
      g_lanes.bind_lane(ln, sh, lcap);
      bound[ln] = R.ci;
    }
    Lane &L = g_lanes.lane[ln];
    if (k >= NL) cudaEventSynchronize(evH2D[ln]);        // the lane's pinned input staging has been consumed
    if (trap)   // chunks whose copies have landed meanwhile (chunks complete in order on the copy stream)
      while (next_mirror < k && cudaEventQuery(evCopy[(next_mirror % NL) * 2 + ((next_mirror / NL) & 1)]) == cudaSuccess) mirror_chunk(next_mirror);
    const double tw1 = now();
    t_wait += tw1 - tw0;
    for (int i = 0; i < n; i++) {
      const int e = el[c0 + i];
      const SigHost &h = C.sig[c0 + i]->h;
      memcpy(L.h_xnod + (size_t)i * nx, xnod + (size_t)e * xnod_ld, sizeof(double) * 3 * h.nH);
      if (gp.source == HP3D_SRC_TABLE)
        memcpy(L.h_src + (size_t)i * nsrc, (const double *)source_qp + (size_t)e * source_ld, sizeof(double) * h.nint * (h.cplx ? 6 : 1) * nrhs);
      L.h_cnt[i] = h.ni; L.h_cnt[lcap + i] = h.nb; L.h_cnt[2 * lcap + i] = h.dims.nil;
      if (mode == MODE_CELEM) L.h_cel[i] = e;
      if (!big) {
        memcpy(L.h_xi + NS * sB * i, (const char *)xi + es * sxi * e, es * h.ni * nrhs);
        if (mode == MODE_RESID && h.nb > 0) memcpy(L.h_xb + NS * (sT + 1) * i, (const char *)xb + es * sxb * e, es * h.nb);
      }
    }
    t_stage += now() - tw1;
    cudaMemcpyAsync(L.d_xnod, L.h_xnod, sizeof(double) * nx * n, cudaMemcpyHostToDevice, st);
    if (gp.source == HP3D_SRC_TABLE) cudaMemcpyAsync(L.d_src, L.h_src, sizeof(double) * nsrc * n, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(L.ws.b.ni_e, L.h_cnt, sizeof(int) * n, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(L.ws.b.nb_e, L.h_cnt + lcap, sizeof(int) * n, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(L.ws.b.nip_e, L.h_cnt + 2 * lcap, sizeof(int) * n, cudaMemcpyHostToDevice, st);
    if (!big) cudaMemcpyAsync(L.d_xi, L.h_xi, es * sB * n, cudaMemcpyHostToDevice, st);
    if (mode == MODE_CELEM) cudaMemcpyAsync(L.d_cel, L.h_cel, sizeof(int) * n, cudaMemcpyHostToDevice, st);
    if (mode == MODE_RESID) cudaMemcpyAsync(L.d_xb, L.h_xb, es * (sT + 1) * n, cudaMemcpyHostToDevice, st);
    cudaEventRecord(evH2D[ln], st);
    chunk_segments(C, c0, n, segs);
    run_chunk(sh, L, ob, gp, segs, n, L.d_xnod, (long long)nx, L.d_src, (long long)nsrc, want_schur, st, nullptr, mode, packed);
    if (mode == MODE_CELEM) run_celem(sh, L, L.out[ob], *cc, n, st);
    const Lane::Out &o = L.out[ob];
    R.h_info = o.h_info; R.h_xb = L.h_xb; R.h_res = L.h_res;
    if (cloc && sh.d.nb > 0) {   // Schur factors: device staging -> their slots (same stream; runs of neighbouring slots are merged)
      for (int i = 0; i < n;) {
        const ClocStore::Slot &s0 = eslot[el[c0 + i]];
        if (!s0.AS) { i++; continue; }
        const size_t ba = (size_t)s0.nb * s0.ni, bb = (size_t)s0.nb * (size_t)nrhs;
        int j = i + 1;
        if (ba == sS && bb == sT)
          while (j < n && eslot[el[c0 + j]].AS == s0.AS + NS * ba * (j - i) && eslot[el[c0 + j]].BS == s0.BS + NS * bb * (j - i) &&
                 eslot[el[c0 + j]].ni == s0.ni && eslot[el[c0 + j]].nb == s0.nb) j++;
        cudaMemcpyAsync(s0.AS, o.AS + NS * sS * i, es * ba * (j - i), cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(s0.BS, o.BS + NS * sT * i, es * bb * (j - i), cudaMemcpyDeviceToDevice, st);
        i = j;
      }
    }
    cudaEventRecord(evCompute[slot], st);
    cudaStreamWaitEvent(g_copy, evCompute[slot], 0);
    if (!big) {   // small results: staged through pinned memory, scattered to the caller in collect_info
      if (mode == MODE_BWD) cudaMemcpyAsync(L.h_xb, L.d_xb, es * (sT + 1) * n, cudaMemcpyDeviceToHost, g_copy);
      else cudaMemcpyAsync(L.h_res, L.d_res, sizeof(double) * n, cudaMemcpyDeviceToHost, g_copy);
      cudaMemcpyAsync(o.h_info, o.info, sizeof(int) * n, cudaMemcpyDeviceToHost, g_copy);
      cudaEventRecord(evCopy[slot], g_copy);
      continue;
    }
    // D2H straight into the caller's arrays; runs of consecutive elements with equal sizes and dense strides are merged
    if (mode == MODE_CELEM) {   // compressed systems: Zastif / IRN / JCN at the caller's offsets, Zbload at xptr
      const size_t zmax = sh.nz_max, cmax = sh.nc_max;
      for (int i = 0; i < n;) {
        const int e = el[c0 + i];
        int j = i + 1;   // merge elements that are adjacent in the caller's arrays AND fill their staging slots completely
        while (j < n && el[c0 + j] == el[c0 + j - 1] + 1 && (size_t)cc->nz(el[c0 + j - 1]) == zmax &&
               (size_t)(cc->xptr[el[c0 + j - 1] + 1] - cc->xptr[el[c0 + j - 1]]) == cmax &&
               cc->aoff[el[c0 + j]] == cc->aoff[el[c0 + j - 1]] + (long long)zmax) j++;
        const int el_last = el[c0 + j - 1];
        const size_t nzr = (size_t)(j - 1 - i) * zmax + (size_t)cc->nz(el_last), ncr = (size_t)(j - 1 - i) * cmax + (size_t)(cc->xptr[el_last + 1] - cc->xptr[el_last]);
        if (nzr) {
          cudaMemcpyAsync((char *)cc->zastif + es * cc->aoff[e], o.Z + NS * zmax * i, es * nzr, cudaMemcpyDeviceToHost, g_copy);
          if (cc->irn) {
            cudaMemcpyAsync(cc->irn + cc->aoff[e], o.irn + zmax * i, sizeof(int) * nzr, cudaMemcpyDeviceToHost, g_copy);
            cudaMemcpyAsync(cc->jcn + cc->aoff[e], o.jcn + zmax * i, sizeof(int) * nzr, cudaMemcpyDeviceToHost, g_copy);
          }
        }
        if (ncr) cudaMemcpyAsync((char *)cc->zbload + es * cc->xptr[e], o.zb + NS * cmax * i, es * ncr, cudaMemcpyDeviceToHost, g_copy);
        i = j;
      }
    }
    for (int i = 0; i < n;) {
      const SigHost &h = C.sig[c0 + i]->h;
      int j = i + 1;
      while (j < n && el[c0 + j] == el[c0 + j - 1] + 1 && C.sig[c0 + j]->h.ni == h.ni && C.sig[c0 + j]->h.nb == h.nb) j++;
      const int e = el[c0 + i], run = j - i;
      auto copy = [&](void *dst, long long stride, const double *src, size_t dstride, size_t blk) {
        if (blk == 0) return;
        if ((size_t)stride == blk && dstride == blk)
          cudaMemcpyAsync((char *)dst + es * stride * e, src + NS * dstride * i, es * blk * run, cudaMemcpyDeviceToHost, g_copy);
        else
          cudaMemcpy2DAsync((char *)dst + es * stride * e, es * stride, src + NS * dstride * i, es * dstride, es * blk, run, cudaMemcpyDeviceToHost, g_copy);
      };
      if (mode == MODE_ELEM) {
        if (trap && h.ni > TRAP_W) {
          // block column b of every element of the run in one strided copy when both element strides are whole columns
          const size_t nn = (size_t)h.ni, pitch = es * nn;
          const bool whole = sA % nn == 0 && (size_t)sAii % nn == 0;
          for (int cb = 0; cb < h.ni; cb += TRAP_W) {
            const size_t w = std::min(TRAP_W, h.ni - cb), rows = nn - cb;
            if (whole) {
              cudaMemcpy3DParms q;
              memset(&q, 0, sizeof q);
              q.srcPtr = make_cudaPitchedPtr((void *)(o.Aii + NS * sA * i), pitch, pitch, sA / nn);
              q.dstPtr = make_cudaPitchedPtr((char *)Aii + es * (size_t)sAii * e, pitch, pitch, (size_t)sAii / nn);
              q.srcPos = make_cudaPos(es * cb, cb, 0); q.dstPos = q.srcPos;
              q.extent = make_cudaExtent(es * rows, w, run);
              q.kind = cudaMemcpyDeviceToHost;
              cudaMemcpy3DAsync(&q, g_copy);
            } else
              for (int kk = 0; kk < run; kk++)
                cudaMemcpy2DAsync((char *)Aii + es * ((size_t)sAii * (e + kk) + (size_t)cb * nn + cb), pitch,
                                  o.Aii + NS * (sA * (i + kk) + (size_t)cb * nn + cb), pitch, es * rows, w, cudaMemcpyDeviceToHost, g_copy);
          }
        } else
          copy(Aii, sAii, o.Aii, sA, packed ? (size_t)h.ni * (h.ni + 1) / 2 : (size_t)h.ni * h.ni);
        copy(Bi, sBi, o.Bi, sB, (size_t)h.ni * nrhs);
      }
      if (to_host_schur) { copy(ASchur, sAS, o.AS, sS, (size_t)h.nb * h.ni); copy(BSchur, sBS, o.BS, sT, (size_t)h.nb * nrhs); }
      i = j;
    }
    cudaMemcpyAsync(o.h_info, o.info, sizeof(int) * n, cudaMemcpyDeviceToHost, g_copy);   // pinned: stays asynchronous
    cudaEventRecord(evCopy[slot], g_copy);
  }
  for (int k = 0; k < nrec; k++) collect_info(k);
  for (ClassGroup &C : classes)
    for (size_t i = 0; i < C.el.size(); i++) { const SigHost &h = C.sig[i]->h; if (ni_out) ni_out[C.el[i]] = h.ni; if (nb_out) nb_out[C.el[i]] = h.nb; }
  {
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) rc = fail(HP3D_ENODEV, "CUDA error in elem_batch: %s", cudaGetErrorString(ce));
  }
  const double t_submitted = now();
  if (trap) g_pool.wait();
  for (int i = 0; i < NL; i++) cudaStreamSynchronize(g_lane_stream[i]);
  const double t_computed = now();
  cudaStreamSynchronize(g_copy);
  if (trace)
    fprintf(stderr, "[hp3d] mode %d nel %d: submit %.1f ms (waits %.1f, staging %.1f), compute drained +%.1f ms, copies drained +%.1f ms\n", mode, nel,
            t_submitted - t_begin, t_wait, t_stage, t_computed - t_submitted, now() - t_computed);
  if (rc == HP3D_OK) {
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) rc = fail(HP3D_ENODEV, "CUDA error in elem_batch: %s", cudaGetErrorString(ce));
  }
  return rc;
}

}  // namespace

extern "C" {

int hp3d_gpu_elem_batch(int plan, int nel, const int *etype, const int *norder, const int *norie, const int *norif,
                        const double *xnod, int xnod_ld, const void *source_qp, long long source_ld, void *Aii, long long sAii,
                        void *Bi, long long sBi, void *ASchur, long long sAS, void *BSchur, long long sBS, int *ni_out, int *nb_out,
                        int *info) {
  return batch_impl(MODE_ELEM, plan, nel, etype, norder, norie, norif, xnod, xnod_ld, source_qp, source_ld, Aii, sAii, Bi, sBi, ASchur, sAS,
                    BSchur, sBS, ni_out, nb_out, info, nullptr, 0, nullptr, 0, nullptr);
}

int hp3d_gpu_elem_bwd_batch(int plan, int nel, const int *etype, const int *norder, const int *norie, const int *norif,
                            const double *xnod, int xnod_ld, const void *source_qp, long long source_ld, const void *xi, long long sxi,
                            void *xb, long long sxb, int *nb_out, int *info) {
  return batch_impl(MODE_BWD, plan, nel, etype, norder, norie, norif, xnod, xnod_ld, source_qp, source_ld, nullptr, 0, nullptr, 0, nullptr, 0,
                    nullptr, 0, nullptr, nb_out, info, xi, sxi, xb, sxb, nullptr);
}

int hp3d_gpu_elem_residual_batch(int plan, int nel, const int *etype, const int *norder, const int *norie, const int *norif,
                                 const double *xnod, int xnod_ld, const void *source_qp, long long source_ld, const void *xi,
                                 long long sxi, const void *xb, long long sxb, double *resid, int *info) {
  return batch_impl(MODE_RESID, plan, nel, etype, norder, norie, norif, xnod, xnod_ld, source_qp, source_ld, nullptr, 0, nullptr, 0, nullptr, 0,
                    nullptr, 0, nullptr, nullptr, info, xi, sxi, const_cast<void *>(xb), sxb, resid);
}

int hp3d_gpu_physics_default(int problem_kind, hp3d_physics *ph) {
  if (!ph) return fail(HP3D_EINVAL, "null argument");
  memset(ph, 0, sizeof *ph);
  // problems/<PROB>/input/physics: (D_TYPE, NR_COMP) per attribute
  static const int tab[4][2][2] = {{{0, 1}, {-1, 0}}, {{0, 1}, {2, 1}}, {{1, 1}, {-1, 0}}, {{1, 2}, {3, 6}}};
  if (problem_kind < HP3D_POIS_GAL || problem_kind > HP3D_MAXW_UW) return fail(HP3D_EINVAL, "unknown problem kind %d", problem_kind);
  int nvar[4] = {0, 0, 0, 0};
  for (int i = 0; i < 2; i++) {
    const int dt = tab[problem_kind - 1][i][0], nc = tab[problem_kind - 1][i][1];
    if (dt < 0) break;
    ph->dtype[i] = dt; ph->ncomp[i] = nc; ph->adres[i] = nvar[dt]; nvar[dt] += nc; ph->nphys++;
  }
  for (int f = 0; f < 3; f++) ph->nrvar[f] = nvar[f];
  return HP3D_OK;
}

long long hp3d_gpu_celem_pack(const hp3d_physics *ph, const int *nrdofl, const int *nrconH, const int *nacH, const double *constrH,
                              const int *nrconE, const int *nacE, const double *constrE, const int *nrconV, const int *nacV,
                              const double *constrV, int nacdim, const int *nrdofm_f, long long *cptr, int *cidx, double *cval, long long cap) {
  if (!ph || !nrdofl || !nrdofm_f || !cptr || nacdim < 1 || ph->nphys < 1 || ph->nphys > HP3D_MAXPHYS) return fail(HP3D_EINVAL, "celem_pack: bad argument");
  const int *nrcon[3] = {nrconH, nrconE, nrconV}, *nac[3] = {nacH, nacE, nacV};
  const double *con[3] = {constrH, constrE, constrV};
  const int nrdofm = nrdofm_f[0] + nrdofm_f[1] + nrdofm_f[2];
  const int base[3] = {0, nrdofm_f[0], nrdofm_f[0] + nrdofm_f[1]};
  // rows of the condensed system are physics-blocked (stc.F90:305-323): off(i) = sum_{j<i} Nrdofs(j), Nrdofs = nrdofl(family) * NR_COMP
  int off[HP3D_MAXPHYS], tot = 0;
  for (int i = 0; i < ph->nphys; i++) {
    off[i] = tot;
    if (ph->dtype[i] >= 0 && ph->dtype[i] <= 2) tot += nrdofl[ph->dtype[i]] * ph->ncomp[i];
  }
  // the reference's loop nest (celem_systemI.F90:553-647): physics, element dof k, connected dof kp, component ivar
  auto visit = [&](auto &&fn) -> int {
    for (int p1 = 0; p1 < ph->nphys; p1++) {
      const int f = ph->dtype[p1];
      if (f < 0 || f > 2 || nrdofl[f] == 0) continue;
      if (!nrcon[f] || !nac[f] || !con[f]) return -1;
      for (int k = 1; k <= nrdofl[f]; k++)
        for (int kp = 1; kp <= nrcon[f][k - 1]; kp++) {
          if (kp > nacdim) return -2;
          const int l = nac[f][(kp - 1) + (long long)nacdim * (k - 1)];
          const double c = con[f][(kp - 1) + (long long)nacdim * (k - 1)];
          for (int ivar = 1; ivar <= ph->ncomp[p1]; ivar++) {
            const int ll = base[f] + (l - 1) * ph->nrvar[f] + ph->adres[p1] + ivar;
            if (l < 1 || ll > base[f] + nrdofm_f[f]) return -3;
            fn(ll, off[p1] + (k - 1) * ph->ncomp[p1] + ivar, c);
          }
        }
    }
    return 0;
  };
  std::vector<long long> cnt(nrdofm + 1, 0);
  int rc = visit([&](int ll, int, double) { cnt[ll]++; });
  if (rc == -1) return fail(HP3D_EINVAL, "celem_pack: missing constraint arrays for a family with dofs");
  if (rc == -2) return fail(HP3D_EINVAL, "celem_pack: nrcon exceeds nacdim");
  if (rc == -3) return fail(HP3D_EINVAL, "celem_pack: nac points outside the modified element");
  cptr[0] = 0;
  for (int ll = 1; ll <= nrdofm; ll++) cptr[ll] = cptr[ll - 1] + cnt[ll];
  const long long nent = cptr[nrdofm];
  if (!cidx || !cval) return nent;
  if (nent > cap) return fail(HP3D_EINVAL, "celem_pack: capacity %lld < %lld entries", cap, nent);
  std::vector<long long> pos(cptr, cptr + nrdofm);
  visit([&](int ll, int row, double c) { const long long q = pos[ll - 1]++; cidx[q] = row; cval[q] = c; });
  return nent;
}

}  // extern "C"
namespace {
int celem_impl(int plan, int nel, const int *etype, const int *norder, const int *norie, const int *norif, const double *xnod,
               int xnod_ld, const void *source_qp, long long source_ld, const long long *mptr, const long long *cptr,
               const int *cidx, const double *cval, const int *idbc, const void *zdofd, const long long *xptr,
               const int *nextract, const int *lcon, int isym_flag, const long long *aptr, void *zbload, void *zastif, int *irn,
               int *jcn, void *ASchur, long long sAS, void *BSchur, long long sBS, int *ni_out, int *nb_out, int *info,
               ClocStore *cloc, const long long *iel) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (int drc = enter_device()) return drc;
  Plan *p = plan_of(plan);
  if (!p) return fail(HP3D_ENOPLAN, "no such plan %d", plan);
  if (isym_flag < 1 || isym_flag > 3) return fail(HP3D_EINVAL, "celem_batch: ISYM_FLAG must be 1, 2 or 3");
  if ((irn == nullptr) != (jcn == nullptr)) return fail(HP3D_EINVAL, "celem_batch: irn and jcn go together");
  if (irn && (isym_flag == 1 || !lcon)) return fail(HP3D_EINVAL, "celem_batch: IRN/JCN need lcon and an unsymmetric ISYM_FLAG (2 or 3)");
  if (nel == 0) return HP3D_OK;
  if (nel < 0 || !norder || !norie || !norif || !xnod || !mptr || !idbc || !zdofd || !xptr || !zbload || !zastif)
    return fail(HP3D_EINVAL, "celem_batch: null argument");
  // cptr == NULL: a REGULAR mesh (no constrained dofs): modified dof ll of every element is its element dof ll with coefficient 1
  std::vector<long long> id_cptr;
  std::vector<int> id_cidx;
  std::vector<double> id_cval;
  if (!cptr) {
    if (cidx || cval) return fail(HP3D_EINVAL, "celem_batch: cidx / cval without cptr");
    const long long nm0 = mptr[nel];
    id_cptr.resize(nm0 + 1); id_cidx.resize(nm0); id_cval.assign(nm0, 1.0);
    for (long long g = 0; g <= nm0; g++) id_cptr[g] = g;
    for (int e = 0; e < nel; e++)
      for (long long g = mptr[e]; g < mptr[e + 1]; g++) id_cidx[g] = (int)(g - mptr[e]) + 1;
    cptr = id_cptr.data(); cidx = id_cidx.data(); cval = id_cval.data();
  }
  const bool cplx = p->fp.kind >= HP3D_MAXW_GAL;
  const bool trace = getenv("HP3D_TRACE") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
  const long long nm = mptr[nel], nx = xptr[nel], nent = cptr[nm];
  if (mptr[0] != 0 || xptr[0] != 0 || cptr[0] != 0) return fail(HP3D_EINVAL, "celem_batch: prefix arrays must start at 0");
  if (nent > 0 && (!cidx || !cval)) return fail(HP3D_EINVAL, "celem_batch: null argument");
  if (nx > 0 && !nextract) return fail(HP3D_EINVAL, "celem_batch: null argument");
  // validate the index data against the element sizes (a bad index would read outside the condensed matrix on the device)
  std::vector<long long> dptr(nel + 1, 0);   // Dirichlet dofs per element (compact, ascending)
  std::vector<int> dlist;
  std::string err;
  CelemCall cc;
  cc.aoff.resize(nel + 1);
  cc.xptr = xptr; cc.isym = isym_flag;
  {   // compile the signatures this call meets for the first time CONCURRENTLY before the (serial) validation loop looks them up
    std::map<std::string, int> first;
    for (int e = 0; e < nel; e++) {
      const int et = etype ? etype[e] : HP3D_MDLB;
      if (et != HP3D_MDLB && et != HP3D_MDLP) return fail(HP3D_EINVAL, "element %d: unknown element type %d", e, et);
      first.emplace(Plan::key(et, norder + 19 * e, norie + 12 * e, norif + 6 * e), e);
    }
    std::vector<std::pair<std::string, int>> missing;
    for (auto &kv : first) if (!p->find(kv.first)) missing.emplace_back(kv.first, kv.second);
    if (p->compile_missing(missing, etype, norder, norie, norif, err)) return fail(HP3D_EINVAL, "%s", err.c_str());
  }
  for (int e = 0; e < nel; e++) {
    if (mptr[e + 1] < mptr[e] || xptr[e + 1] < xptr[e]) return fail(HP3D_EINVAL, "celem_batch: element %d: prefix arrays must be non-decreasing", e);
    Signature *S = p->get(etype ? etype[e] : HP3D_MDLB, norder + 19 * e, norie + 12 * e, norif + 6 * e, false, err);
    if (!S) return fail(HP3D_EINVAL, "element %d: %s", e, err.c_str());
    const int ni = S->h.ni;
    const long long nme = mptr[e + 1] - mptr[e];
    for (long long g = mptr[e]; g < mptr[e + 1]; g++) {
      if (cptr[g + 1] < cptr[g]) return fail(HP3D_EINVAL, "celem_batch: element %d: cptr must be non-decreasing", e);
      for (long long q = cptr[g]; q < cptr[g + 1]; q++)
        if (cidx[q] < 1 || cidx[q] > ni) return fail(HP3D_EINVAL, "celem_batch: element %d: cidx %d outside 1..ni=%d", e, cidx[q], ni);
      if (idbc[g] == 1) dlist.push_back((int)(g - mptr[e]));
    }
    for (long long l = xptr[e]; l < xptr[e + 1]; l++)
      if (nextract[l] < 1 || nextract[l] > nme) return fail(HP3D_EINVAL, "celem_batch: element %d: NEXTRACT %d outside 1..Nrdofm=%lld", e, nextract[l], nme);
    dptr[e + 1] = (long long)dlist.size();
    cc.aoff[e] = aptr ? aptr[e] : (e ? cc.aoff[e - 1] + cc.nz(e - 1) : 0);
  }
  cc.zbload = zbload; cc.zastif = zastif; cc.irn = irn; cc.jcn = jcn;
  const double t1 = now();
  // the constraint arrays live in grow-only device buffers owned by the library (no cudaMalloc / cudaFree per call)
  auto up = [&](int slot, const void *h, size_t bytes) -> void * {
    void *d = g_celem_store.get(slot, bytes ? bytes : 8);
    if (d && bytes && cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, g_compute) != cudaSuccess) return nullptr;
    return d;
  };
  const size_t es = sizeof(double) * (cplx ? 2 : 1);
  cc.d_mptr = (long long *)up(0, mptr, sizeof(long long) * (nel + 1)); cc.d_cptr = (long long *)up(1, cptr, sizeof(long long) * (nm + 1));
  cc.d_xptr = (long long *)up(2, xptr, sizeof(long long) * (nel + 1)); cc.d_cidx = (int *)up(3, cidx, sizeof(int) * nent);
  cc.d_cval = (double *)up(4, cval, sizeof(double) * nent); cc.d_dlist = (int *)up(5, dlist.data(), sizeof(int) * dlist.size());
  cc.d_zdofd = (double *)up(6, zdofd, es * nm); cc.d_nextract = (int *)up(7, nextract, sizeof(int) * nx);
  cc.d_lcon = (int *)up(8, lcon, lcon ? sizeof(int) * nx : 0); cc.d_dptr = (long long *)up(9, dptr.data(), sizeof(long long) * (nel + 1));
  bool ok = cc.d_mptr && cc.d_cptr && cc.d_xptr && cc.d_cidx && cc.d_cval && cc.d_dlist && cc.d_zdofd && cc.d_nextract && cc.d_lcon && cc.d_dptr;
  // the lanes read these arrays: every lane stream waits for the uploads (pageable sources: the copies are staged by the driver
  // before cudaMemcpyAsync returns, so the local vectors may go out of scope)
  if (ok) ok = cudaStreamSynchronize(g_compute) == cudaSuccess;
  int rc;
  const double t2 = now();
  if (!ok) rc = fail(HP3D_ENOMEM, "celem_batch: cannot place the constraint data on the device: %s", cudaGetErrorString(cudaGetLastError()));
  else
    rc = batch_impl(MODE_CELEM, plan, nel, etype, norder, norie, norif, xnod, xnod_ld, source_qp, source_ld, nullptr, 0, nullptr, 0, ASchur, sAS,
                    BSchur, sBS, ni_out, nb_out, info, nullptr, 0, nullptr, 0, nullptr, &cc, cloc, iel);
  const double t3 = now();
  if (trace) fprintf(stderr, "[hp3d] celem_batch: validate %.1f ms, upload %.1f ms, pipeline %.1f ms, release %.1f ms\n", t1 - t0, t2 - t1, t3 - t2, now() - t3);
  return rc;
}
}  // namespace
extern "C" {

int hp3d_gpu_celem_batch(int plan, int nel, const int *etype, const int *norder, const int *norie, const int *norif, const double *xnod,
                         int xnod_ld, const void *source_qp, long long source_ld, const long long *mptr, const long long *cptr,
                         const int *cidx, const double *cval, const int *idbc, const void *zdofd, const long long *xptr,
                         const int *nextract, const int *lcon, int isym_flag, const long long *aptr, void *zbload, void *zastif, int *irn,
                         int *jcn, void *ASchur, long long sAS, void *BSchur, long long sBS, int *ni_out, int *nb_out, int *info) {
  return celem_impl(plan, nel, etype, norder, norie, norif, xnod, xnod_ld, source_qp, source_ld, mptr, cptr, cidx, cval, idbc, zdofd, xptr, nextract,
                    lcon, isym_flag, aptr, zbload, zastif, irn, jcn, ASchur, sAS, BSchur, sBS, ni_out, nb_out, info, nullptr, nullptr);
}

// ---- device-resident CLOC (see the ClocStore comment above batch_impl) ----
int hp3d_gpu_cloc_create(int plan, long long limit_bytes) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (int drc = enter_device()) return drc;
  Plan *p = plan_of(plan);
  if (!p) return fail(HP3D_ENOPLAN, "no such plan %d", plan);
  if (limit_bytes < 0) return fail(HP3D_EINVAL, "cloc_create: negative byte limit");
  if (limit_bytes == 0) {
    size_t fr = 0, tot = 0;
    CUDA_TRY(cudaMemGetInfo(&fr, &tot));
    limit_bytes = (long long)(0.6 * (double)fr);
  }
  ClocStore *c = new ClocStore();
  c->plan = plan; c->cplx = p->fp.kind >= HP3D_MAXW_GAL; c->nrhs = p->fp.nrhs; c->limit = (size_t)limit_bytes;
  for (size_t i = 0; i < g_clocs.size(); i++)
    if (!g_clocs[i]) { g_clocs[i] = c; return (int)i; }
  g_clocs.push_back(c);
  return (int)g_clocs.size() - 1;
}

int hp3d_gpu_cloc_clear(int cloc) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (int drc = enter_device()) return drc;
  ClocStore *c = cloc_of(cloc);
  if (!c) return fail(HP3D_EINVAL, "no such Schur store %d", cloc);
  for (int i = 0; i < LaneSet::NLANE; i++) cudaStreamSynchronize(g_lane_stream[i]);
  c->release();
  return HP3D_OK;
}

int hp3d_gpu_cloc_destroy(int cloc) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (int rc = hp3d_gpu_cloc_clear(cloc)) return rc;
  delete g_clocs[cloc];
  g_clocs[cloc] = nullptr;
  return HP3D_OK;
}

int hp3d_gpu_cloc_stats(int cloc, long long *stats) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  ClocStore *c = cloc_of(cloc);
  if (!c || !stats) return fail(HP3D_EINVAL, "no such Schur store %d", cloc);
  stats[0] = (long long)c->slots.size(); stats[1] = (long long)c->spilled.size(); stats[2] = (long long)c->bytes; stats[3] = (long long)c->limit;
  return HP3D_OK;
}

int hp3d_gpu_elem_batch_cloc(int plan, int cloc, int nel, const long long *iel, const int *etype, const int *norder, const int *norie,
                             const int *norif, const double *xnod, int xnod_ld, const void *source_qp, long long source_ld, void *Aii,
                             long long sAii, void *Bi, long long sBi, int *ni_out, int *nb_out, int *info) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  ClocStore *c = cloc_of(cloc);
  if (!c) return fail(HP3D_EINVAL, "no such Schur store %d", cloc);
  return batch_impl(MODE_ELEM, plan, nel, etype, norder, norie, norif, xnod, xnod_ld, source_qp, source_ld, Aii, sAii, Bi, sBi, nullptr, 0,
                    nullptr, 0, ni_out, nb_out, info, nullptr, 0, nullptr, 0, nullptr, nullptr, c, iel);
}

int hp3d_gpu_celem_batch_cloc(int plan, int cloc, int nel, const long long *iel, const int *etype, const int *norder, const int *norie,
                              const int *norif, const double *xnod, int xnod_ld, const void *source_qp, long long source_ld,
                              const long long *mptr, const long long *cptr, const int *cidx, const double *cval, const int *idbc,
                              const void *zdofd, const long long *xptr, const int *nextract, const int *lcon, int isym_flag,
                              const long long *aptr, void *zbload, void *zastif, int *irn, int *jcn, int *ni_out, int *nb_out, int *info) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  ClocStore *c = cloc_of(cloc);
  if (!c) return fail(HP3D_EINVAL, "no such Schur store %d", cloc);
  return celem_impl(plan, nel, etype, norder, norie, norif, xnod, xnod_ld, source_qp, source_ld, mptr, cptr, cidx, cval, idbc, zdofd, xptr, nextract,
                    lcon, isym_flag, aptr, zbload, zastif, irn, jcn, nullptr, 0, nullptr, 0, ni_out, nb_out, info, c, iel);
}

int hp3d_gpu_cloc_fetch(int cloc, long long iel, void *ASchur, void *BSchur, int *ni, int *nb) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (int drc = enter_device()) return drc;
  ClocStore *c = cloc_of(cloc);
  if (!c) return fail(HP3D_EINVAL, "no such Schur store %d", cloc);
  if (c->spilled.count(iel)) return 1;
  auto it = c->slots.find(iel);
  if (it == c->slots.end()) return fail(HP3D_EINVAL, "cloc_fetch: element %lld has not been condensed into this store", iel);
  const ClocStore::Slot &sl = it->second;
  const size_t es = sizeof(double) * (c->cplx ? 2 : 1);
  if (ni) *ni = sl.ni;
  if (nb) *nb = sl.nb;
  for (int i = 0; i < LaneSet::NLANE; i++) cudaStreamSynchronize(g_lane_stream[i]);
  if (ASchur && sl.nb) CUDA_TRY(cudaMemcpy(ASchur, sl.AS, es * (size_t)sl.nb * sl.ni, cudaMemcpyDeviceToHost));
  if (BSchur && sl.nb) CUDA_TRY(cudaMemcpy(BSchur, sl.BS, es * (size_t)sl.nb * c->nrhs, cudaMemcpyDeviceToHost));
  return HP3D_OK;
}

int hp3d_gpu_cloc_bwd_batch(int cloc, int nel, const long long *iel, const void *xi, long long sxi, void *xb, long long sxb, int *nb_out,
                            int *info) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (int drc = enter_device()) return drc;
  ClocStore *c = cloc_of(cloc);
  if (!c) return fail(HP3D_EINVAL, "no such Schur store %d", cloc);
  if (nel < 0 || (nel > 0 && (!xi || !xb))) return fail(HP3D_EINVAL, "cloc_bwd: null argument");
  const int NS = c->cplx ? 2 : 1;
  const size_t es = sizeof(double) * NS, nr = (size_t)c->nrhs;   // xi / xb hold nr columns per element (column q at q*ni / q*nb)
  std::vector<int> res, spl;   // caller positions of resident / spilled elements
  int nimax = 0, nbmax = 0;
  for (int e = 0; e < nel; e++) {
    const long long id = iel ? iel[e] : e;
    auto it = c->slots.find(id);
    if (it != c->slots.end()) {
      res.push_back(e);
      nimax = std::max(nimax, it->second.ni); nbmax = std::max(nbmax, it->second.nb);
      if (sxi < (long long)(it->second.ni * nr) || sxb < (long long)(it->second.nb * nr))
        return fail(HP3D_EINVAL, "cloc_bwd: element %lld needs strides >= (%d, %d) x nr_rhs", id, it->second.ni, it->second.nb);
      if (nb_out) nb_out[e] = it->second.nb;
      if (info) info[e] = 0;
    } else if (c->spilled.count(id)) spl.push_back(e);
    else return fail(HP3D_EINVAL, "cloc_bwd: element %lld has not been condensed into this store", id);
  }
  for (int i = 0; i < LaneSet::NLANE; i++) cudaStreamSynchronize(g_lane_stream[i]);   // the factors were written on the lane streams
  // resident elements: one warp per bubble row reads its factors in place; chunks bounded by the grid limit and a 256 MB staging budget
  if (!res.empty() && nbmax > 0) {
    const size_t wi = (size_t)nimax * nr, wb = (size_t)nbmax * nr;   // staging widths per element
    const size_t per = es * (wi + wb) + 2 * sizeof(void *) + 2 * sizeof(int);
    const int chunk = (int)std::min<size_t>(std::min<size_t>(res.size(), 32768), std::max<size_t>(1, ((size_t)256 << 20) / per));
    struct Bufs {
      void *d = nullptr, *h = nullptr;
      ~Bufs() { if (d) cudaFree(d); if (h) cudaFreeHost(h); }
    } bufs;
    const size_t oX = 0, oY = oX + es * wi * chunk, oPA = (oY + es * wb * chunk + 15) & ~(size_t)15,
                 oPB = oPA + sizeof(void *) * chunk, oNI = oPB + sizeof(void *) * chunk, oNB = oNI + sizeof(int) * chunk, total = oNB + sizeof(int) * chunk;
    if (cudaMalloc(&bufs.d, total) != cudaSuccess || cudaMallocHost(&bufs.h, total) != cudaSuccess) { cudaGetLastError(); return fail(HP3D_ENOMEM, "cloc_bwd: staging buffers"); }
    char *h = (char *)bufs.h, *d = (char *)bufs.d;
    for (size_t c0 = 0; c0 < res.size(); c0 += chunk) {
      const int n = (int)std::min<size_t>(chunk, res.size() - c0);
      for (int i = 0; i < n; i++) {
        const int e = res[c0 + i];
        const ClocStore::Slot &sl = c->slots.find(iel ? iel[e] : e)->second;
        memcpy(h + oX + es * wi * i, (const char *)xi + es * sxi * e, es * sl.ni * nr);
        ((const double **)(h + oPA))[i] = sl.AS; ((const double **)(h + oPB))[i] = sl.BS;
        ((int *)(h + oNI))[i] = sl.ni; ((int *)(h + oNB))[i] = sl.nb;
      }
      CUDA_TRY(cudaMemcpyAsync(d, h, es * wi * n, cudaMemcpyHostToDevice, g_compute));
      CUDA_TRY(cudaMemcpyAsync(d + oPA, h + oPA, total - oPA, cudaMemcpyHostToDevice, g_compute));
      dim3 grid((nbmax + 7) / 8, n);
      if (c->cplx)
        stc_bwd_ptr_kernel<true><<<grid, 256, 0, g_compute>>>((const double *const *)(d + oPA), (const double *const *)(d + oPB), (const int *)(d + oNI),
                                                              (const int *)(d + oNB), (const double *)(d + oX), (long long)wi, (double *)(d + oY), (long long)wb, c->nrhs);
      else
        stc_bwd_ptr_kernel<false><<<grid, 256, 0, g_compute>>>((const double *const *)(d + oPA), (const double *const *)(d + oPB), (const int *)(d + oNI),
                                                               (const int *)(d + oNB), (const double *)(d + oX), (long long)wi, (double *)(d + oY), (long long)wb, c->nrhs);
      g_launches++;
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaMemcpyAsync(h + oY, d + oY, es * wb * n, cudaMemcpyDeviceToHost, g_compute));
      CUDA_TRY(cudaStreamSynchronize(g_compute));
      for (int i = 0; i < n; i++) {
        const int e = res[c0 + i];
        memcpy((char *)xb + es * sxb * e, h + oY + es * wb * i, es * ((int *)(h + oNB))[i] * nr);
      }
    }
  }
  // spilled elements: recompute (hp3d_gpu_elem_bwd_batch's pipeline) from the descriptors kept at condensation time
  if (!spl.empty()) {
    const int m = (int)spl.size();
    Plan *p = plan_of(c->plan);
    if (!p) return fail(HP3D_ENOPLAN, "cloc_bwd: the store's plan %d is gone", c->plan);
    size_t xl = 0, sl = 0;
    for (int e : spl) { const ClocStore::Spill &sp = c->spilled.find(iel ? iel[e] : e)->second; xl = std::max(xl, sp.xnod.size()); sl = std::max(sl, sp.src.size()); }
    std::vector<int> et(m), no(19 * (size_t)m), oe(12 * (size_t)m), of(6 * (size_t)m), nbo(m), inf(m);
    std::vector<double> xn(xl * m, 0.0), src(sl * m, 0.0);
    std::vector<char> xis(es * (size_t)sxi * m), xbs(es * (size_t)sxb * m);
    for (int k = 0; k < m; k++) {
      const int e = spl[k];
      const ClocStore::Spill &sp = c->spilled.find(iel ? iel[e] : e)->second;
      et[k] = sp.etype;
      memcpy(&no[19 * (size_t)k], sp.norder, sizeof sp.norder); memcpy(&oe[12 * (size_t)k], sp.norie, sizeof sp.norie); memcpy(&of[6 * (size_t)k], sp.norif, sizeof sp.norif);
      std::copy(sp.xnod.begin(), sp.xnod.end(), xn.begin() + xl * k);
      std::copy(sp.src.begin(), sp.src.end(), src.begin() + sl * k);
      memcpy(&xis[es * (size_t)sxi * k], (const char *)xi + es * sxi * e, es * sxi);
    }
    int rc = batch_impl(MODE_BWD, c->plan, m, et.data(), no.data(), oe.data(), of.data(), xn.data(), (int)xl, sl ? src.data() : nullptr, (long long)sl,
                        nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, nbo.data(), inf.data(), xis.data(), sxi, xbs.data(), sxb, nullptr);
    if (rc) return rc;
    for (int k = 0; k < m; k++) {
      const int e = spl[k];
      memcpy((char *)xb + es * sxb * e, &xbs[es * (size_t)sxb * k], es * nbo[k] * nr);
      if (nb_out) nb_out[e] = nbo[k];
      if (info) info[e] = inf[k];
    }
  }
  return HP3D_OK;
}

int hp3d_gpu_hermitian_unpack_batch(int cplx, int nel, int ni, const int *ni_e, const void *AP, long long sAP, void *A, long long sA, int threads) {
  if (nel < 0 || (nel > 0 && (!AP || !A)) || (!ni_e && ni < 0)) return fail(HP3D_EINVAL, "hermitian_unpack: bad argument");
  if (AP == A) return fail(HP3D_EINVAL, "hermitian_unpack: in place is not supported");
  const int NS = cplx ? 2 : 1;
  // work unit = (element, block of 32 columns): column c is one contiguous run of the packed array (rows c..n-1) copied into
  // the lower triangle, then mirrored (conjugated) into row c of the upper triangle while it is still in cache
  std::atomic<long long> next{0};
  int nimax = ni;
  if (ni_e) { nimax = 0; for (int e = 0; e < nel; e++) nimax = std::max(nimax, ni_e[e]); }
  const long long nblk = (nimax + 31) / 32, nwork = (long long)nel * nblk;
  auto work = [&] {
    for (;;) {
      const long long w = next.fetch_add(1);
      if (w >= nwork) return;
      const int e = (int)(w / nblk), n = ni_e ? ni_e[e] : ni;
      const int cb = (int)(w % nblk) * 32, ce = std::min(n, cb + 32);
      const double *ap = (const double *)AP + (size_t)e * sAP * NS;
      double *a = (double *)A + (size_t)e * sA * NS;
      for (int c = cb; c < ce; c++) {
        const double *src = ap + ((size_t)c * (2 * (size_t)n - c - 1) / 2 + c) * NS;   // AP index of (c,c)
        double *col = a + ((size_t)c * n + c) * NS;
        memcpy(col, src, sizeof(double) * NS * (n - c));
        if (cplx) { col[1] = 0.0; for (int r = c + 1; r < n; r++) { double *u = a + ((size_t)r * n + c) * 2; u[0] = src[2 * (r - c)]; u[1] = -src[2 * (r - c) + 1]; } }
        else for (int r = c + 1; r < n; r++) a[(size_t)r * n + c] = src[r - c];
      }
    }
  };
  unsigned nt = threads > 0 ? (unsigned)threads : std::max(1u, std::thread::hardware_concurrency());
  nt = (unsigned)std::min<long long>(nt, std::max<long long>(1, nwork / 4));
  std::vector<std::thread> th;
  for (unsigned t = 1; t < nt; t++) th.emplace_back(work);
  work();
  for (std::thread &t : th) t.join();
  return HP3D_OK;
}

int hp3d_gpu_quad_points(int plan, int nel, const int *etype, const int *norder, const int *norie, const int *norif,
                         const double *xnod, int xnod_ld, double *xq, long long sxq) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (int drc = enter_device()) return drc;
  Plan *p = plan_of(plan);
  if (!p) return fail(HP3D_ENOPLAN, "no such plan %d", plan);
  GeomParams gp = p->geom();
  gp.source = HP3D_SRC_ZERO;
  std::string err;
  std::vector<double> f;
  for (int e = 0; e < nel; e++) {  // not a hot path: one element at a time
    Signature *S = p->get(etype ? etype[e] : HP3D_MDLB, norder + 19 * e, norie + 12 * e, norif + 6 * e, true, err);
    if (!S) return fail(HP3D_EINVAL, "element %d: %s", e, err.c_str());
    ChunkShape sh; sh.absorb(S->h);
    if (g_lanes.reserve(sh, 1, err)) return fail(HP3D_ENOMEM, "%s", err.c_str());
    const SigHost &h = S->h;
    Lane &L = g_lanes.lane[0];
    CUDA_TRY(cudaMemcpyAsync(L.d_xnod, xnod + (size_t)e * xnod_ld, sizeof(double) * 3 * h.nH, cudaMemcpyHostToDevice, g_compute));
    SigTables sg;
    sg.tab = S->d_tab; sg.wq = S->d_wq; sg.hdof = S->d_hdof; sg.nH = h.nH; sg.nint = h.nint;
    sg.ttab = S->d_ttab ? S->d_ttab + h.geo_toff : nullptr; sg.nT = h.geo_nT;
    for (int i = 0; i < 3; i++) sg.nq[i] = h.nq[i];
    if (h.etype == HP3D_MDLP) geom_fields_prism_kernel<<<(h.nint + 127) / 128, 128, 0, g_compute>>>(sg, gp, 1, L.d_xnod, 3LL * h.nH, nullptr, 0, L.d_WF, L.ws.b.info);
    else geom_fields_kernel<<<(h.nint + 127) / 128, 128, 0, g_compute>>>(sg, gp, 1, L.d_xnod, 3LL * h.nH, nullptr, 0, L.d_WF, L.ws.b.info);
    const size_t fs = (size_t)wf_stride(h.nint);
    f.resize(3 * fs);
    CUDA_TRY(cudaMemcpyAsync(f.data(), L.d_WF + (size_t)F_X * fs, sizeof(double) * 3 * fs, cudaMemcpyDeviceToHost, g_compute));
    CUDA_TRY(cudaStreamSynchronize(g_compute));
    for (int q = 0; q < h.nint; q++)
      for (int c = 0; c < 3; c++) xq[(size_t)e * sxq + 3 * q + c] = f[(size_t)c * fs + q];
  }
  return HP3D_OK;
}

int hp3d_gpu_stc_bwd_batch(int cplx, int nel, int ni, int nb, const void *ASchur, long long sAS, const void *BSchur, long long sBS,
                           const void *xi, long long sxi, void *xb, long long sxb) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (int drc = enter_device()) return drc;
  if (nel <= 0 || ni <= 0 || nb <= 0) return fail(HP3D_EINVAL, "bad sizes");
  if (!ASchur || !BSchur || !xi || !xb) return fail(HP3D_EINVAL, "null argument");
  if (sAS < (long long)nb * ni || sBS < nb || sxi < ni || sxb < nb) return fail(HP3D_EINVAL, "stc_bwd_batch: a stride is shorter than its block");
  const size_t es = sizeof(double) * (cplx ? 2 : 1);
  // chunks bounded by the grid limit (65535 elements) and by a quarter of the free device memory; buffers released on every path
  size_t fr = 0, tot = 0;
  CUDA_TRY(cudaMemGetInfo(&fr, &tot));
  const size_t per = es * ((size_t)nb * ni + 2 * (size_t)nb + ni);
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(nel, 65535), (fr / 4) / per));
  struct Bufs {
    double *p[4] = {nullptr, nullptr, nullptr, nullptr};
    ~Bufs() { for (double *q : p) if (q) cudaFree(q); }
  } b;
  CUDA_TRY(cudaMalloc(&b.p[0], es * (size_t)nb * ni * chunk)); CUDA_TRY(cudaMalloc(&b.p[1], es * (size_t)nb * chunk));
  CUDA_TRY(cudaMalloc(&b.p[2], es * (size_t)ni * chunk)); CUDA_TRY(cudaMalloc(&b.p[3], es * (size_t)nb * chunk));
  for (int c0 = 0; c0 < nel; c0 += chunk) {
    const int n = std::min(chunk, nel - c0);
    // element e of each host array sits at e*stride; copy the used extents as 2-D blocks
    CUDA_TRY(cudaMemcpy2DAsync(b.p[0], es * nb * ni, (const char *)ASchur + es * sAS * c0, es * sAS, es * nb * ni, n, cudaMemcpyHostToDevice, g_compute));
    CUDA_TRY(cudaMemcpy2DAsync(b.p[1], es * nb, (const char *)BSchur + es * sBS * c0, es * sBS, es * nb, n, cudaMemcpyHostToDevice, g_compute));
    CUDA_TRY(cudaMemcpy2DAsync(b.p[2], es * ni, (const char *)xi + es * sxi * c0, es * sxi, es * ni, n, cudaMemcpyHostToDevice, g_compute));
    dim3 grid((nb + 7) / 8, n);
    if (cplx) stc_bwd_kernel<true><<<grid, 256, 0, g_compute>>>(nullptr, nullptr, ni, nb, b.p[0], (long long)nb * ni, b.p[1], nb, b.p[2], ni, b.p[3], nb);
    else stc_bwd_kernel<false><<<grid, 256, 0, g_compute>>>(nullptr, nullptr, ni, nb, b.p[0], (long long)nb * ni, b.p[1], nb, b.p[2], ni, b.p[3], nb);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy2DAsync((char *)xb + es * sxb * c0, es * sxb, b.p[3], es * nb, es * nb, n, cudaMemcpyDeviceToHost, g_compute));
    CUDA_TRY(cudaStreamSynchronize(g_compute));
  }
  return HP3D_OK;
}

// ---- FP64 tensor-pipe probe: the roofline denominator bench.py reports against, measured on the device the run uses
namespace {
__global__ void __launch_bounds__(256) dmma_rate_kernel(double *out, int iters, double seed) {
  // 8 independent m8n8k4 accumulator tiles per warp cover the pipe latency; nothing but DMMA in the loop
  double acc[8][2], a[8], b[4];
  for (int i = 0; i < 8; i++) { acc[i][0] = seed * i; acc[i][1] = seed * (i + 1); a[i] = seed + threadIdx.x * 1e-9 * i; }
  for (int i = 0; i < 4; i++) b[i] = seed * 0.5 + i * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) dmma884(acc[i][0], acc[i][1], a[i], b[i & 3]);
  }
  double sum = 0;
  for (int i = 0; i < 8; i++) sum += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}
}  // namespace

int hp3d_gpu_fp64_peak_probe(double *tflops, double *ms_best) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (int drc = enter_device()) return drc;
  if (!tflops) return fail(HP3D_EINVAL, "null argument");
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, g_device));
  const int grid = prop.multiProcessorCount * 2, iters = 40000;
  double *out = nullptr;
  CUDA_TRY(cudaMalloc(&out, sizeof(double) * grid * 256));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  dmma_rate_kernel<<<grid, 256, 0, g_compute>>>(out, 2000, 1.0);
  float best = 1e30f;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(e0, g_compute);
    dmma_rate_kernel<<<grid, 256, 0, g_compute>>>(out, iters, 1.0);
    cudaEventRecord(e1, g_compute);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, ms);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  CUDA_TRY(cudaGetLastError());
  *tflops = (double)grid * 8 /*warps*/ * iters * 8.0 /*mma per iteration*/ * (2.0 * 8 * 8 * 4) / (best * 1e-3) / 1e12;
  if (ms_best) *ms_best = best;
  return HP3D_OK;
}

int hp3d_gpu_bench(int plan, int nel, const int *norder, const int *norie, const int *norif, const double *xnod, int xnod_ld, int reps,
                   int max_chunk, int lanes, double *ms_total, double *ms_integ, double *ms_dense, long long *launches) {
  return hp3d_gpu_bench_t(plan, nel, nullptr, norder, norie, norif, xnod, xnod_ld, reps, max_chunk, lanes, ms_total, ms_integ, ms_dense, launches);
}

int hp3d_gpu_bench_t(int plan, int nel, const int *etype, const int *norder, const int *norie, const int *norif, const double *xnod,
                     int xnod_ld, int reps, int max_chunk, int lanes, double *ms_total, double *ms_integ, double *ms_dense,
                     long long *launches) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (int drc = enter_device()) return drc;
  Plan *p = plan_of(plan);
  if (!p) return fail(HP3D_ENOPLAN, "no such plan %d", plan);
  if (nel <= 0 || reps <= 0 || lanes < 1 || lanes > LaneSet::NLANE) return fail(HP3D_EINVAL, "bad sizes");
  if (p->fp.source == HP3D_SRC_TABLE) return fail(HP3D_EINVAL, "bench: table sources are not supported");
  std::vector<ClassGroup> classes;
  std::string err;
  if (int brc = build_classes(p, nel, etype, norder, norie, norif, true, classes, err)) return fail(brc, "%s", err.c_str());
  const GeomParams gp = p->geom();
  // resident inputs per class: geometry dofs (uniform stride 3*nH_max) and the per-element dof counts
  struct Grp { ClassGroup *C; double *dx; int *dcnt; int n, chunk; double work; };
  std::vector<Grp> gs;
  // Several dense classes (hp meshes): every lane owns a partition of the arena and chunks of DIFFERENT classes run side by side
  // on different lanes, no drain between classes.  Classes go in order of decreasing work; a class that carries a large share of
  // the call is still split over the lanes (it would otherwise be the tail on one lane), a small one stays whole: fewer, larger launches.
  const bool multi = classes.size() > 1 && lanes > 1;
  double work_total = 0.0;
  std::vector<double> work(classes.size(), 0.0);
  for (size_t i = 0; i < classes.size(); i++) {
    const DenseDims &d = classes[i].shape.d;
    const double n = d.np, m = d.M();
    work[i] = (double)classes[i].el.size() * (n * n * n / 3.0 + n * n * m + n * m * m + m * m * m / 3.0 + 1e6);
    work_total += work[i];
  }
  if (multi) {
    std::vector<size_t> ord(classes.size());
    for (size_t i = 0; i < ord.size(); i++) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return work[a] > work[b]; });
    std::vector<ClassGroup> sorted; std::vector<double> w2;
    for (size_t i : ord) { sorted.push_back(std::move(classes[i])); w2.push_back(work[i]); }
    classes.swap(sorted); work.swap(w2);
  }
  size_t part_dev = 0, part_host = 0;
  for (size_t ci = 0; ci < classes.size(); ci++) {
    ClassGroup &C = classes[ci];
    int want = (int)C.el.size();
    static const double split_frac = getenv("HP3D_SPLIT_FRAC") ? atof(getenv("HP3D_SPLIT_FRAC")) : 2.0;   // of one lane's share of the call
    if (!multi || work[ci] > split_frac * work_total / lanes) want = (want + lanes - 1) / lanes;
    if (max_chunk > 0 && want > max_chunk) want = max_chunk;
    int cap = chunk_capacity(C.shape, want, lanes);
    if (cap < 1) return fail(HP3D_ENOMEM, "not enough device memory");
    if (multi) {
      size_t db, hb;
      LaneSet::lane_bytes(C.shape, cap, db, hb);
      part_dev = std::max(part_dev, db); part_host = std::max(part_host, hb);
    } else if (g_lanes.reserve(C.shape, cap, err, lanes)) return fail(HP3D_ENOMEM, "%s", err.c_str());   // grows the arena to the largest class
    const size_t nx = 3 * (size_t)C.shape.nH_max, n = C.el.size();
    std::vector<double> hx(nx * n, 0.0);
    std::vector<int> hc(3 * n);
    for (size_t i = 0; i < n; i++) {
      memcpy(hx.data() + i * nx, xnod + (size_t)C.el[i] * xnod_ld, sizeof(double) * 3 * C.sig[i]->h.nH);
      hc[i] = C.sig[i]->h.ni; hc[n + i] = C.sig[i]->h.nb; hc[2 * n + i] = C.sig[i]->h.dims.nil;
    }
    double *dx; int *dc;
    CUDA_TRY(cudaMalloc(&dx, sizeof(double) * hx.size()));
    CUDA_TRY(cudaMemcpy(dx, hx.data(), sizeof(double) * hx.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&dc, sizeof(int) * hc.size()));
    CUDA_TRY(cudaMemcpy(dc, hc.data(), sizeof(int) * hc.size(), cudaMemcpyHostToDevice));
    gs.push_back(Grp{&C, dx, dc, (int)n, cap, work[ci]});
  }
  if (multi) {
    int need = 1;
    for (ClassGroup &C : classes) need = std::max(need, std::max(C.shape.d.ni, C.shape.d.nb) + 1);
    if (LaneSet::ensure_partitions(part_dev, part_host, err) || g_lanes.ensure_iota(need, err)) return fail(HP3D_ENOMEM, "%s", err.c_str());
  }
  std::vector<StageEvents> evs;
  cudaEvent_t t0, t1, tj, tsw[LaneSet::NLANE];
  CUDA_TRY(cudaEventCreate(&t0)); CUDA_TRY(cudaEventCreate(&t1)); CUDA_TRY(cudaEventCreateWithFlags(&tj, cudaEventDisableTiming));
  for (int i = 0; i < LaneSet::NLANE; i++) CUDA_TRY(cudaEventCreateWithFlags(&tsw[i], cudaEventDisableTiming));
  const long long l0 = g_launches;
  for (int i = 0; i < LaneSet::NLANE; i++) CUDA_TRY(cudaStreamSynchronize(g_lane_stream[i]));
  CUDA_TRY(cudaEventRecord(t0, g_lane_stream[0]));
  for (int i = 1; i < lanes; i++) CUDA_TRY(cudaStreamWaitEvent(g_lane_stream[i], t0, 0));   // the other lanes start inside the timed region
  int k = 0;
  std::vector<Seg> segs;
  for (int r = 0; r < reps; r++)
    for (Grp &g : gs) {
      if (gs.size() > 1 && !multi) {   // another class's buffers occupy the arena: both lanes must drain before they are re-bound
        for (int i = 0; i < lanes; i++) cudaEventRecord(tsw[i], g_lane_stream[i]);
        for (int i = 0; i < lanes; i++)
          for (int j = 0; j < lanes; j++) if (i != j) cudaStreamWaitEvent(g_lane_stream[i], tsw[j], 0);
        g_arena.owner = nullptr;   // force the re-layout for this class (pointer arithmetic only: the arena is large enough)
        if (g_lanes.reserve(g.C->shape, g.chunk, err, lanes)) return fail(HP3D_ENOMEM, "%s", err.c_str());
      }
      const long long nx = 3LL * g.C->shape.nH_max;
      for (int c0 = 0; c0 < g.n; c0 += g.chunk, k++) {
        const int n = std::min(g.n - c0, g.chunk), ln = k % lanes;
        if (multi) g_lanes.bind_lane(ln, g.C->shape, g.chunk);   // this lane's partition, ordered after the lane's earlier work by its stream
        Lane &L = g_lanes.lane[ln];
        StageEvents ev;
        ev.on = (lanes == 1);
        if (ev.on) for (int i = 0; i < 4; i++) cudaEventCreate(&ev.e[i]);
        cudaMemcpyAsync(L.ws.b.ni_e, g.dcnt + c0, sizeof(int) * n, cudaMemcpyDeviceToDevice, g_lane_stream[ln]);
        cudaMemcpyAsync(L.ws.b.nb_e, g.dcnt + g.n + c0, sizeof(int) * n, cudaMemcpyDeviceToDevice, g_lane_stream[ln]);
        cudaMemcpyAsync(L.ws.b.nip_e, g.dcnt + 2 * g.n + c0, sizeof(int) * n, cudaMemcpyDeviceToDevice, g_lane_stream[ln]);
        chunk_segments(*g.C, (size_t)c0, n, segs);
        run_chunk(g.C->shape, L, 0, gp, segs, n, g.dx + (size_t)c0 * nx, nx, nullptr, 0, p->store_schur != 0, g_lane_stream[ln], &ev);
        if (ev.on) evs.push_back(ev);
      }
    }
  for (int i = 1; i < lanes; i++) { CUDA_TRY(cudaEventRecord(tj, g_lane_stream[i])); CUDA_TRY(cudaStreamWaitEvent(g_lane_stream[0], tj, 0)); }
  CUDA_TRY(cudaEventRecord(t1, g_lane_stream[0]));
  for (int i = 0; i < LaneSet::NLANE; i++) CUDA_TRY(cudaStreamSynchronize(g_lane_stream[i]));
  CUDA_TRY(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, t0, t1);
  double mi = 0, md = 0;
  for (StageEvents &ev : evs) {
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, ev.e[0], ev.e[1]);
    cudaEventElapsedTime(&b, ev.e[1], ev.e[2]);
    mi += a; md += b;
    for (int i = 0; i < 4; i++) cudaEventDestroy(ev.e[i]);
  }
  if (ms_total) *ms_total = ms;
  if (ms_integ) *ms_integ = mi;
  if (ms_dense) *ms_dense = md;
  if (launches) *launches = g_launches - l0;
  cudaEventDestroy(t0); cudaEventDestroy(t1); cudaEventDestroy(tj);
  for (int i = 0; i < LaneSet::NLANE; i++) cudaEventDestroy(tsw[i]);
  for (Grp &g : gs) { cudaFree(g.dx); cudaFree(g.dcnt); }
  return HP3D_OK;
}

int hp3d_gpu_integrate_debug(int plan, const int *norder, const int *norie, const int *norif, const double *xnod, const void *source_qp,
                             double *W, long long cap_doubles, int *dims) {
  return hp3d_gpu_integrate_debug_t(plan, HP3D_MDLB, norder, norie, norif, xnod, source_qp, W, cap_doubles, dims);
}

int hp3d_gpu_integrate_debug_t(int plan, int etype, const int *norder, const int *norie, const int *norif, const double *xnod,
                               const void *source_qp, double *W, long long cap_doubles, int *dims) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (int drc = enter_device()) return drc;
  Plan *p = plan_of(plan);
  if (!p) return fail(HP3D_ENOPLAN, "no such plan %d", plan);
  std::string err;
  Signature *S = p->get(etype, norder, norie, norif, true, err);
  if (!S) return fail(HP3D_EINVAL, "%s", err.c_str());
  ChunkShape sh; sh.absorb(S->h);
  if (g_lanes.reserve(sh, 1, err)) return fail(HP3D_ENOMEM, "%s", err.c_str());
  const SigHost &h = S->h;
  const DenseDims &d = h.dims;
  const size_t P = d.planes();
  const size_t need = P * (d.dpg ? d.w_plane() : d.a_plane());
  if (dims) { dims[0] = d.np; dims[1] = d.nbp; dims[2] = d.nil;   /* interface rows that carry data: the load rows end at nbp + dims[2] */ dims[3] = d.n; dims[4] = d.nb; dims[5] = d.ni; dims[6] = d.dpg ? d.R() : d.M(); dims[7] = (int)P; }
  if (!W) return HP3D_OK;
  if ((size_t)cap_doubles < need) return fail(HP3D_EINVAL, "integrate_debug: need %zu doubles", need);
  Lane &L = g_lanes.lane[0];
  CUDA_TRY(cudaMemcpyAsync(L.d_xnod, xnod, sizeof(double) * 3 * h.nH, cudaMemcpyHostToDevice, g_compute));
  if (p->fp.source == HP3D_SRC_TABLE) {
    if (!source_qp) return fail(HP3D_EINVAL, "source table missing");
    CUDA_TRY(cudaMemcpyAsync(L.d_src, source_qp, sizeof(double) * S->src_doubles(), cudaMemcpyHostToDevice, g_compute));
  }
  std::vector<Seg> segs(1, Seg{S, 0, 1});
  run_integration(sh, L, p->geom(), segs, 1, L.d_xnod, 3LL * h.nH, L.d_src, (long long)S->src_doubles(), g_compute);
  CUDA_TRY(cudaMemcpyAsync(W, d.dpg ? L.ws.b.W : L.ws.b.Am, sizeof(double) * need, cudaMemcpyDeviceToHost, g_compute));
  CUDA_TRY(cudaStreamSynchronize(g_compute));
  CUDA_TRY(cudaGetLastError());
  return HP3D_OK;
}

int hp3d_gpu_dense_debug(int cplx, int nel, int n, int nb, int ni, const void *G, const void *Bm, void *Aii, void *Bi,
                         void *ASchur, void *BSchur, int *info) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (int drc = enter_device()) return drc;
  if (nel <= 0 || n <= 0 || nb < 0 || ni <= 0) return fail(HP3D_EINVAL, "bad sizes");
  std::string err;
  int rc = cplx ? dense_debug_run<true>(nel, n, nb, ni, G, Bm, Aii, Bi, ASchur, BSchur, info, err)
                : dense_debug_run<false>(nel, n, nb, ni, G, Bm, Aii, Bi, ASchur, BSchur, info, err);
  if (rc) return fail(rc, "%s", err.c_str());
  return HP3D_OK;
}

}  // extern "C"

namespace {
// device-resident tables of one error signature (cached in g_errsigs for the life of the library)
struct ErrSig {
  ErrSigHost h;
  double *d_w = nullptr, *d_tabH = nullptr, *d_tabF = nullptr;
  ~ErrSig() { cudaFree(d_w); cudaFree(d_tabH); if (d_tabF != d_tabH) cudaFree(d_tabF); }
};
std::map<std::string, std::unique_ptr<ErrSig>> g_errsigs;
}  // namespace
namespace {
void release_error_signatures() { g_errsigs.clear(); }

// shared implementation of hp3d_gpu_elem_error_batch / hp3d_gpu_error_points
int error_impl(int plan, int nel, const int *etype, const int *norder, const int *norie, const int *norif, const double *xnod, int xnod_ld,
               const void *zdof, long long szdof, const void *exact_qp, long long exact_ld, int l2proj, double *err, double *rnorm, int *info,
               double *xq, long long sxq, int *nint_out) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (int drc = enter_device()) return drc;
  Plan *p = plan_of(plan);
  if (!p) return fail(HP3D_ENOPLAN, "no such plan %d", plan);
  if (nel < 0 || (nel > 0 && (!norder || !norie || !norif || !xnod))) return fail(HP3D_EINVAL, "null argument");
  const int kind = p->fp.kind;
  const bool cplx = kind >= HP3D_MAXW_GAL;
  const size_t es = sizeof(double) * (cplx ? 2 : 1);
  // group by signature (element type, orders, orientations)
  std::map<std::string, std::vector<int>> bysig;
  for (int e = 0; e < nel; e++) {
    const int et = etype ? etype[e] : HP3D_MDLB;
    bysig[std::to_string(kind) + "/" + std::to_string(p->fp.maxp) + "/" + Plan::key(et, norder + 19 * e, norie + 12 * e, norif + 6 * e)].push_back(e);
  }
  GeomParams gp = p->geom();
  for (auto &g : bysig) {
    const std::vector<int> &el = g.second;
    const int e0 = el[0];
    auto it = g_errsigs.find(g.first);
    if (it == g_errsigs.end()) {
      std::unique_ptr<ErrSig> S(new ErrSig());
      if (!compile_error_signature(kind, p->fp.maxp, etype ? etype[e0] : HP3D_MDLB, norder + 19 * e0, norie + 12 * e0, norif + 6 * e0, S->h))
        return fail(HP3D_EINVAL, "element %d: %s", e0, S->h.err.c_str());
      std::string uerr;
      if (dev_upload(S->h.w, &S->d_w, uerr) || dev_upload(S->h.tabH, &S->d_tabH, uerr)) return fail(HP3D_ENOMEM, "%s", uerr.c_str());
      if (S->h.space == ES_H1) S->d_tabF = S->d_tabH;
      else if (dev_upload(S->h.tabF, &S->d_tabF, uerr)) return fail(HP3D_ENOMEM, "%s", uerr.c_str());
      it = g_errsigs.emplace(g.first, std::move(S)).first;
    }
    const ErrSig &S = *it->second;
    const ErrSigHost &h = S.h;
    if (xnod_ld < 3 * h.nH) return fail(HP3D_EINVAL, "xnod_ld=%d < 3*nrdofH=%d", xnod_ld, 3 * h.nH);
    const int n = (int)el.size(), nv = h.space == ES_H1 ? 4 : 6;
    const size_t nz = (size_t)h.ncomp * h.nF, nex = (size_t)nv * h.nint;
    if (zdof && szdof < (long long)nz) return fail(HP3D_EINVAL, "zdof stride %lld < ncomp*nrdof = %zu", szdof, nz);
    if (exact_qp && exact_ld < (long long)nex) return fail(HP3D_EINVAL, "exact_ld %lld < nvals*nint = %zu", exact_ld, nex);
    if (xq && sxq < 3LL * h.nint) return fail(HP3D_EINVAL, "xq stride %lld < 3*nint = %d", sxq, 3 * h.nint);
    // gather the group's inputs (host) and run it as one launch
    std::vector<double> hx((size_t)3 * h.nH * n), hz(zdof ? nz * n * (cplx ? 2 : 1) : 0), hex(exact_qp ? nex * n * (cplx ? 2 : 1) : 0);
    for (int i = 0; i < n; i++) {
      memcpy(hx.data() + (size_t)i * 3 * h.nH, xnod + (size_t)el[i] * xnod_ld, sizeof(double) * 3 * h.nH);
      if (zdof) memcpy((char *)hz.data() + es * nz * i, (const char *)zdof + es * szdof * el[i], es * nz);
      if (exact_qp) memcpy((char *)hex.data() + es * nex * i, (const char *)exact_qp + es * exact_ld * el[i], es * nex);
    }
    double *dx = nullptr, *dz = nullptr, *dex = nullptr, *dout = nullptr, *dxq = nullptr;
    int *dinfo = nullptr;
    std::string uerr;
    if (dev_upload(hx, &dx, uerr) || dev_upload(hz, &dz, uerr) || dev_upload(hex, &dex, uerr)) return fail(HP3D_ENOMEM, "%s", uerr.c_str());
    CUDA_TRY(cudaMalloc(&dout, sizeof(double) * 2 * n)); CUDA_TRY(cudaMalloc(&dinfo, sizeof(int) * n));
    CUDA_TRY(cudaMemsetAsync(dinfo, 0, sizeof(int) * n, g_compute));
    CUDA_TRY(cudaMemsetAsync(dout, 0, sizeof(double) * 2 * n, g_compute));
    if (xq) CUDA_TRY(cudaMalloc(&dxq, sizeof(double) * 3 * h.nint * n));
    ErrArgs A;
    A.w = S.d_w; A.tabH = S.d_tabH; A.tabF = S.d_tabF; A.nint = h.nint; A.nH = h.nH; A.nF = h.nF; A.space = h.space; A.ncomp = h.ncomp;
    A.gp = gp; A.nel = n; A.xnod = dx; A.xnod_ld = 3LL * h.nH; A.zdof = dz; A.szd = (long long)nz; A.exact = dex; A.sex = (long long)nex;
    A.l2proj = l2proj; A.err = dout; A.rnorm = dout + n; A.info = dinfo; A.want_points = xq != nullptr; A.xq = dxq; A.sxq = 3LL * h.nint;
    if (cplx) elem_error_kernel<true><<<n, 256, 0, g_compute>>>(A);
    else elem_error_kernel<false><<<n, 256, 0, g_compute>>>(A);
    g_launches++;
    std::vector<double> hout(2 * (size_t)n), hxq(xq ? (size_t)3 * h.nint * n : 0);
    std::vector<int> hinfo(n);
    CUDA_TRY(cudaMemcpyAsync(hout.data(), dout, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, g_compute));
    CUDA_TRY(cudaMemcpyAsync(hinfo.data(), dinfo, sizeof(int) * n, cudaMemcpyDeviceToHost, g_compute));
    if (xq) CUDA_TRY(cudaMemcpyAsync(hxq.data(), dxq, sizeof(double) * hxq.size(), cudaMemcpyDeviceToHost, g_compute));
    CUDA_TRY(cudaStreamSynchronize(g_compute));
    CUDA_TRY(cudaGetLastError());
    for (int i = 0; i < n; i++) {
      if (err) err[el[i]] = hout[i];
      if (rnorm) rnorm[el[i]] = hout[n + i];
      if (info) info[el[i]] = hinfo[i];
      if (nint_out) nint_out[el[i]] = h.nint;
      if (xq) memcpy(xq + (size_t)el[i] * sxq, hxq.data() + (size_t)i * 3 * h.nint, sizeof(double) * 3 * h.nint);
    }
    cudaFree(dx); cudaFree(dz); cudaFree(dex); cudaFree(dout); cudaFree(dinfo); cudaFree(dxq);
  }
  return HP3D_OK;
}
}  // namespace

extern "C" {

int hp3d_gpu_elem_error_batch(int plan, int nel, const int *etype, const int *norder, const int *norie, const int *norif, const double *xnod,
                              int xnod_ld, const void *zdof, long long szdof, const void *exact_qp, long long exact_ld, int l2proj, double *err,
                              double *rnorm, int *info) {
  if (!zdof || !err || !rnorm) return fail(HP3D_EINVAL, "elem_error: null argument");
  return error_impl(plan, nel, etype, norder, norie, norif, xnod, xnod_ld, zdof, szdof, exact_qp, exact_ld, l2proj, err, rnorm, info, nullptr, 0, nullptr);
}

int hp3d_gpu_error_points(int plan, int nel, const int *etype, const int *norder, const int *norie, const int *norif, const double *xnod,
                          int xnod_ld, double *xq, long long sxq, int *nint_out) {
  if (!xq && !nint_out) return fail(HP3D_EINVAL, "error_points: null argument");
  if (!xq) {   // sizes only
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    Plan *p = plan_of(plan);
    if (!p) return fail(HP3D_ENOPLAN, "no such plan %d", plan);
    for (int e = 0; e < nel; e++) {
      ErrSigHost h;
      if (!compile_error_signature(p->fp.kind, p->fp.maxp, etype ? etype[e] : HP3D_MDLB, norder + 19 * e, norie + 12 * e, norif + 6 * e, h))
        return fail(HP3D_EINVAL, "element %d: %s", e, h.err.c_str());
      nint_out[e] = h.nint;
    }
    return HP3D_OK;
  }
  return error_impl(plan, nel, etype, norder, norie, norif, xnod, xnod_ld, nullptr, 0, nullptr, 0, 0, nullptr, nullptr, nullptr, xq, sxq, nint_out);
}

}  // extern "C"

// ---- H1 projection-based interpolation (pbi.cuh) -----------------------------------------------------------------------------
namespace {
struct PbiSig {
  PbiSigHost h;   // node descriptors + points; the gradient table lives on the device only
  bool on_device = false;
  double *d_wa = nullptr, *d_tan = nullptr, *d_grad = nullptr, *d_tabE = nullptr;
  PbiNode *d_nodes = nullptr;
  ~PbiSig() { cudaFree(d_wa); cudaFree(d_tan); cudaFree(d_grad); cudaFree(d_tabE); cudaFree(d_nodes); }
};
std::map<std::string, std::unique_ptr<PbiSig>> g_pbisigs;
CelemStore g_pbi_store;   // grow-only device buffers of hp3d_gpu_pbi_h1_batch
size_t g_pbi_table_bytes = 0;                        // device bytes held by the cached signature tables
size_t PBI_TABLE_CACHE_BYTES = 8ull << 30;           // hp meshes produce thousands of signatures (2.6 MB each at p=5): bound the cache
void release_pbi_signatures() { g_pbisigs.clear(); g_pbi_store.release(); g_pbi_table_bytes = 0; }
// called at the start of a batch call (no kernel of this library is in flight on the tables then): drop the device tables, keep the
// host-side descriptors; the signatures a call needs are rebuilt on demand
void pbi_trim_table_cache() {
  if (g_pbi_table_bytes <= PBI_TABLE_CACHE_BYTES) return;
  cudaDeviceSynchronize();
  for (auto &kv : g_pbisigs) {
    PbiSig &S = *kv.second;
    cudaFree(S.d_wa); cudaFree(S.d_tan); cudaFree(S.d_grad); cudaFree(S.d_tabE); cudaFree(S.d_nodes);
    S.d_wa = S.d_tan = S.d_grad = S.d_tabE = nullptr; S.d_nodes = nullptr; S.on_device = false;
  }
  g_pbi_table_bytes = 0;
}
// static + dynamic shared memory of the large-node variants exceeds the 48 KB default: opt in once per context
int pbi_opt_in_smem() {
  cudaError_t e = cudaFuncSetAttribute(pbi_node_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PBI_GSMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(pbi_node_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PBI_GSMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(pbi_hcurl_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PBI_GSMEM_BYTES);
  if (e != cudaSuccess) return fail(HP3D_ENODEV, "cudaFuncSetAttribute (projection-based interpolation kernels): %s", cudaGetErrorString(e));
  return HP3D_OK;
}

// descriptors + points of a signature (host only); `tables` also builds and uploads the gradient table
PbiSig *pbi_signature(int et, const int *norder, const int *norie, const int *norif, int integration, int maxp, bool tables, std::string &err,
                      int space = PBI_H1) {
  const std::string key = std::to_string(space) + "/" + std::to_string(integration) + "/" + std::to_string(maxp) + "/" + Plan::key(et, norder, norie, norif);
  auto it = g_pbisigs.find(key);
  if (it == g_pbisigs.end()) {
    std::unique_ptr<PbiSig> S(new PbiSig());
    if (!compile_pbi_signature(et, norder, norie, norif, integration, maxp, false, S->h, space)) { err = S->h.err; return nullptr; }
    it = g_pbisigs.emplace(key, std::move(S)).first;
  }
  PbiSig *S = it->second.get();
  if (tables && !S->on_device) {
    PbiSigHost full;
    if (!compile_pbi_signature(et, norder, norie, norif, integration, maxp, true, full, space)) { err = full.err; return nullptr; }
    std::vector<PbiNode> nodes(full.node, full.node + full.nnode);
    if (dev_upload(full.wa, &S->d_wa, err) || dev_upload(full.tan, &S->d_tan, err) || dev_upload(full.grad, &S->d_grad, err) ||
        dev_upload(full.tabE, &S->d_tabE, err) || dev_upload(nodes, &S->d_nodes, err))
      return nullptr;
    S->on_device = true;
    g_pbi_table_bytes += sizeof(double) * (full.wa.size() + full.tan.size() + full.grad.size() + full.tabE.size()) + sizeof(PbiNode) * nodes.size();
  }
  return S;
}

// Compile the signatures a call meets for the first time CONCURRENTLY on the host's cores (as Plan::compile_missing does for the
// element engine): on an hp mesh nearly every element has its own interpolation signature and one compilation with tables costs
// milliseconds.  Afterwards pbi_signature() finds the descriptors in the cache; with `tables` the device tables are uploaded here too.
int pbi_precompile(int nel, const int *etype, const int *norder, const int *norie, const int *norif, int integration, int maxp, bool tables,
                   int space, std::string &err) {
  struct Job { std::string key; int e; bool need_desc; std::unique_ptr<PbiSig> S; PbiSigHost full; bool ok_desc = true, ok_full = true; };
  std::vector<Job> jobs;
  {
    std::map<std::string, int> seen;
    for (int e = 0; e < nel; e++) {
      const int et = etype ? etype[e] : HP3D_MDLB;
      std::string key = std::to_string(space) + "/" + std::to_string(integration) + "/" + std::to_string(maxp) + "/" + Plan::key(et, norder + 19 * e, norie + 12 * e, norif + 6 * e);
      if (!seen.emplace(key, e).second) continue;
      auto it = g_pbisigs.find(key);
      const bool need_desc = it == g_pbisigs.end(), need_tab = tables && (need_desc || !it->second->on_device);
      if (!need_desc && !need_tab) continue;
      Job j; j.key = std::move(key); j.e = e; j.need_desc = need_desc;
      jobs.push_back(std::move(j));
    }
  }
  if (jobs.size() < 2) return 0;   // nothing to gain: the serial path handles it
  std::atomic<size_t> next{0};
  auto work = [&]() {
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= jobs.size()) break;
      Job &j = jobs[i];
      const int e = j.e, et = etype ? etype[e] : HP3D_MDLB;
      if (j.need_desc) {
        j.S.reset(new PbiSig());
        j.ok_desc = compile_pbi_signature(et, norder + 19 * e, norie + 12 * e, norif + 6 * e, integration, maxp, false, j.S->h, space);
      }
      if (tables && j.ok_desc) j.ok_full = compile_pbi_signature(et, norder + 19 * e, norie + 12 * e, norif + 6 * e, integration, maxp, true, j.full, space);
    }
  };
  unsigned nt = std::thread::hardware_concurrency();
  nt = std::max(1u, std::min(nt ? nt : 4u, std::min((unsigned)jobs.size(), 32u)));
  std::vector<std::thread> th;
  for (unsigned t = 1; t < nt; t++) th.emplace_back(work);
  work();
  for (std::thread &t : th) t.join();
  for (Job &j : jobs) {
    if (!j.ok_desc) { err = "element " + std::to_string(j.e) + ": " + j.S->h.err; return -1; }
    if (!j.ok_full) { err = "element " + std::to_string(j.e) + ": " + j.full.err; return -1; }
    PbiSig *S = j.need_desc ? g_pbisigs.emplace(j.key, std::move(j.S)).first->second.get() : g_pbisigs.find(j.key)->second.get();
    if (tables && !S->on_device) {
      std::vector<PbiNode> nodes(j.full.node, j.full.node + j.full.nnode);
      if (dev_upload(j.full.wa, &S->d_wa, err) || dev_upload(j.full.tan, &S->d_tan, err) || dev_upload(j.full.grad, &S->d_grad, err) ||
          dev_upload(j.full.tabE, &S->d_tabE, err) || dev_upload(nodes, &S->d_nodes, err))
        return -1;
      S->on_device = true;
      g_pbi_table_bytes += sizeof(double) * (j.full.wa.size() + j.full.tan.size() + j.full.grad.size() + j.full.tabE.size()) + sizeof(PbiNode) * nodes.size();
    }
  }
  return 0;
}
}  // namespace

extern "C" {

int hp3d_gpu_pbi_points(int nel, const int *etype, const int *norder, const int *norie, const int *norif, int integration, int maxp,
                        double *xi, long long xi_ld, int *npts, int *nrdofH, int *nodes) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (nel < 0 || (nel > 0 && (!norder || !norie || !norif))) return fail(HP3D_EINVAL, "pbi_points: null argument");
  { std::string perr; if (pbi_precompile(nel, etype, norder, norie, norif, integration, maxp, false, PBI_H1, perr)) return fail(HP3D_EINVAL, "%s", perr.c_str()); }
  for (int e = 0; e < nel; e++) {
    std::string err;
    const PbiSig *S = pbi_signature(etype ? etype[e] : HP3D_MDLB, norder + 19 * e, norie + 12 * e, norif + 6 * e, integration, maxp, false, err);
    if (!S) return fail(HP3D_EINVAL, "element %d: %s", e, err.c_str());
    const PbiSigHost &h = S->h;
    if (npts) npts[e] = h.npts;
    if (nrdofH) nrdofH[e] = h.nH;
    if (nodes) {
      int *q = nodes + (size_t)e * 4 * PBI_MAXNODE;
      for (int i = 0; i < 4 * PBI_MAXNODE; i++) q[i] = 0;
      for (int i = 0; i < h.nnode; i++) { q[4 * i] = h.node[i].t0; q[4 * i + 1] = h.node[i].n; q[4 * i + 2] = h.node[i].p0; q[4 * i + 3] = h.node[i].np; }
    }
    if (xi) {
      if (xi_ld < 3LL * h.npts) return fail(HP3D_EINVAL, "pbi_points: xi_ld %lld < 3*npts = %d (element %d)", xi_ld, 3 * h.npts, e);
      memcpy(xi + (size_t)e * xi_ld, h.xi.data(), sizeof(double) * 3 * h.npts);
    }
  }
  return HP3D_OK;
}

int hp3d_gpu_pbi_h1_batch(int nel, const int *etype, const int *norder, const int *norie, const int *norif, int integration, int maxp,
                          const double *etav, int ncomp, const double *fvert, const double *fgrad, long long fgrad_ld,
                          const unsigned *mask, double *dof, long long dof_ld, int *info) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (int drc = enter_device()) return drc;
  if (nel < 0 || (nel > 0 && (!norder || !norie || !norif || !etav || !fvert || !fgrad || !dof))) return fail(HP3D_EINVAL, "pbi_h1: null argument");
  if (ncomp < 1 || ncomp > PBI_MAXCOMP) return fail(HP3D_EINVAL, "pbi_h1: ncomp = %d outside 1..%d (split the components over several calls)", ncomp, PBI_MAXCOMP);
  if (nel == 0) return HP3D_OK;
  if (int rc = pbi_opt_in_smem()) return rc;
  pbi_trim_table_cache();
  { std::string perr; if (pbi_precompile(nel, etype, norder, norie, norif, integration, maxp, true, PBI_H1, perr)) return fail(HP3D_EINVAL, "%s", perr.c_str()); }
  // ---- signature groups (elements are addressed in place through an index list: no host-side gather)
  struct Group { PbiSig *S; std::vector<int> el; };
  std::map<const PbiSig *, size_t> where;
  std::vector<Group> groups;
  for (int e = 0; e < nel; e++) {
    std::string err;
    PbiSig *S = pbi_signature(etype ? etype[e] : HP3D_MDLB, norder + 19 * e, norie + 12 * e, norif + 6 * e, integration, maxp, true, err);
    if (!S) return fail(HP3D_EINVAL, "element %d: %s", e, err.c_str());
    if (dof_ld < (long long)ncomp * S->h.nH) return fail(HP3D_EINVAL, "pbi_h1: dof_ld %lld < ncomp*nrdofH = %d (element %d)", dof_ld, ncomp * S->h.nH, e);
    if (fgrad_ld < 3LL * ncomp * S->h.npts) return fail(HP3D_EINVAL, "pbi_h1: fgrad_ld %lld < 3*ncomp*npts = %d (element %d)", fgrad_ld, 3 * ncomp * S->h.npts, e);
    auto w = where.find(S);
    if (w == where.end()) { w = where.emplace(S, groups.size()).first; groups.push_back(Group{S, {}}); }
    groups[w->second].el.push_back(e);
  }
  // ---- workspace: (n + ncomp)(3 np + n) doubles per CTA, at most 1 GiB per launch
  struct LaunchDims { long long stride[3]; int ny[3]; };
  std::vector<LaunchDims> dims(groups.size());
  long long need_ws = 0;
  std::vector<int> elist; elist.reserve(nel);
  for (size_t g = 0; g < groups.size(); g++) {
    const PbiSigHost &h = groups[g].S->h;
    const int n = (int)groups[g].el.size();
    const int first[4] = {h.nrv, h.nrv + h.nre, h.nrv + h.nre + h.nrf, h.nnode};
    for (int s = 0; s < 3; s++) {
      long long st = 0;
      for (int i = first[s]; i < first[s + 1]; i++)
        if (h.node[i].n > 0) st = std::max(st, (long long)(h.node[i].n + ncomp) * (3LL * h.node[i].np + h.node[i].n));
      dims[g].stride[s] = st; dims[g].ny[s] = 0;
      if (!st) continue;
      const long long per_row = st * (first[s + 1] - first[s]) * (long long)sizeof(double);
      dims[g].ny[s] = (int)std::max(1LL, std::min(std::min((long long)n, 65535LL), (1LL << 30) / per_row));
      if (st * (long long)sizeof(double) > PBI_SMALL_BYTES) need_ws = std::max(need_ws, per_row * dims[g].ny[s]);   // small nodes live in shared memory
    }
    elist.insert(elist.end(), groups[g].el.begin(), groups[g].el.end());
  }
  const size_t b_ev = sizeof(double) * 24 * (size_t)nel, b_fv = sizeof(double) * 8 * ncomp * (size_t)nel,
               b_fg = std::max(sizeof(double) * (size_t)fgrad_ld * nel, sizeof(double)), b_d = sizeof(double) * (size_t)dof_ld * nel;
  double *dev_ev = (double *)g_pbi_store.get(0, b_ev), *dev_fv = (double *)g_pbi_store.get(1, b_fv), *dev_fg = (double *)g_pbi_store.get(2, b_fg),
         *dev_d = (double *)g_pbi_store.get(3, b_d), *dev_ws = (double *)g_pbi_store.get(4, (size_t)std::max(need_ws, 8LL));
  unsigned *dev_m = mask ? (unsigned *)g_pbi_store.get(5, sizeof(unsigned) * nel) : nullptr;
  int *dinfo = (int *)g_pbi_store.get(6, sizeof(int) * nel), *dev_el = (int *)g_pbi_store.get(7, sizeof(int) * nel);
  if (!dev_ev || !dev_fv || !dev_fg || !dev_d || !dev_ws || (mask && !dev_m) || !dinfo || !dev_el)
    return fail(HP3D_ENOMEM, "pbi_h1: out of device memory (%zu bytes of inputs, %lld bytes of workspace)", b_ev + b_fv + b_fg + b_d, need_ws);
  CUDA_TRY(cudaMemcpyAsync(dev_ev, etav, b_ev, cudaMemcpyHostToDevice, g_compute));
  CUDA_TRY(cudaMemcpyAsync(dev_fv, fvert, b_fv, cudaMemcpyHostToDevice, g_compute));
  if (fgrad_ld > 0) CUDA_TRY(cudaMemcpyAsync(dev_fg, fgrad, sizeof(double) * (size_t)fgrad_ld * nel, cudaMemcpyHostToDevice, g_compute));
  CUDA_TRY(cudaMemcpyAsync(dev_d, dof, b_d, cudaMemcpyHostToDevice, g_compute));   // nodes outside the mask keep (and contribute) their entries
  if (mask) CUDA_TRY(cudaMemcpyAsync(dev_m, mask, sizeof(unsigned) * nel, cudaMemcpyHostToDevice, g_compute));
  CUDA_TRY(cudaMemcpyAsync(dev_el, elist.data(), sizeof(int) * nel, cudaMemcpyHostToDevice, g_compute));
  CUDA_TRY(cudaMemsetAsync(dinfo, 0, sizeof(int) * nel, g_compute));
  // ---- per group: vertices, then the three dependent launches (edges, faces, middle node)
  size_t pos = 0;
  for (size_t g = 0; g < groups.size(); g++) {
    const PbiSig &S = *groups[g].S;
    const PbiSigHost &h = S.h;
    const int n = (int)groups[g].el.size();
    const int first[4] = {h.nrv, h.nrv + h.nre, h.nrv + h.nre + h.nrf, h.nnode};
    PbiArgs A;
    A.wa = S.d_wa; A.tan = S.d_tan; A.grad = S.d_grad; A.nodes = S.d_nodes; A.nH = h.nH; A.nrv = h.nrv; A.npts = h.npts; A.ncomp = ncomp;
    A.node0 = 0; A.nel = n; A.elems = dev_el + pos; A.etav = dev_ev; A.fgrad = dev_fg; A.fvert = dev_fv; A.mask = dev_m; A.dof = dev_d;
    A.fgrad_ld = fgrad_ld; A.dof_ld = dof_ld; A.ws = dev_ws; A.ws_stride = 0; A.g_in_smem = 0; A.info = dinfo;
    pbi_vertex_kernel<<<(n * h.nrv + 127) / 128, 128, 0, g_compute>>>(A);
    g_launches++;
    for (int s = 0; s < 3; s++) {
      if (!dims[g].stride[s]) continue;
      A.node0 = first[s]; A.ws_stride = dims[g].stride[s];
      int nmax = 0;
      for (int i = first[s]; i < first[s + 1]; i++) nmax = std::max(nmax, h.node[i].n);
      const size_t small_bytes = sizeof(double) * (size_t)dims[g].stride[s];
      if (small_bytes <= (size_t)PBI_SMALL_BYTES)   // D and G of every node of the launch fit in shared memory: 64-thread CTAs, no workspace traffic
        pbi_node_kernel<4, true><<<dim3(first[s + 1] - first[s], std::min(n, 65535)), 64, small_bytes, g_compute>>>(A);
      // (the system matrix of these larger nodes stays in the workspace: keeping it in shared memory was measured 9 % slower at p=5,
      //  the middle-node launch is bound by the D traffic of phases A and B, not by the factorisation's barriers)
      else if (nmax <= 64) pbi_node_kernel<4, false><<<dim3(first[s + 1] - first[s], dims[g].ny[s]), 256, 0, g_compute>>>(A);
      else pbi_node_kernel<2, false><<<dim3(first[s + 1] - first[s], dims[g].ny[s]), 256, 0, g_compute>>>(A);
      g_launches++;
    }
    pos += n;
  }
  CUDA_TRY(cudaGetLastError());
  std::vector<int> hinfo(info ? 0 : nel);
  CUDA_TRY(cudaMemcpyAsync(dof, dev_d, b_d, cudaMemcpyDeviceToHost, g_compute));
  CUDA_TRY(cudaMemcpyAsync(info ? info : hinfo.data(), dinfo, sizeof(int) * nel, cudaMemcpyDeviceToHost, g_compute));
  CUDA_TRY(cudaStreamSynchronize(g_compute));
  return HP3D_OK;
}

}  // extern "C"

extern "C" {

static int pbi_vec_points(int space, int nel, const int *etype, const int *norder, const int *norie, const int *norif, int maxp, double *xi,
                          long long xi_ld, int *npts, int *nrdofE, int *nodes) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (nel < 0 || (nel > 0 && (!norder || !norie || !norif))) return fail(HP3D_EINVAL, "pbi_hcurl_points: null argument");
  { std::string perr; if (pbi_precompile(nel, etype, norder, norie, norif, 1, maxp, false, space, perr)) return fail(HP3D_EINVAL, "%s", perr.c_str()); }
  for (int e = 0; e < nel; e++) {
    std::string err;
    const PbiSig *S = pbi_signature(etype ? etype[e] : HP3D_MDLB, norder + 19 * e, norie + 12 * e, norif + 6 * e, 1, maxp, false, err, space);
    if (!S) return fail(HP3D_EINVAL, "element %d: %s", e, err.c_str());
    const PbiSigHost &h = S->h;
    if (npts) npts[e] = h.npts;
    if (nrdofE) nrdofE[e] = h.nEF;
    if (nodes) {
      int *q = nodes + (size_t)e * 4 * PBI_MAXNODE;
      for (int i = 0; i < 4 * PBI_MAXNODE; i++) q[i] = 0;
      for (int i = 0; i < h.nnode; i++) { q[4 * i] = h.node[i].t0; q[4 * i + 1] = h.node[i].n; q[4 * i + 2] = h.node[i].p0; q[4 * i + 3] = h.node[i].np; }
    }
    if (xi) {
      if (xi_ld < 3LL * h.npts) return fail(HP3D_EINVAL, "pbi_hcurl_points: xi_ld %lld < 3*npts = %d (element %d)", xi_ld, 3 * h.npts, e);
      memcpy(xi + (size_t)e * xi_ld, h.xi.data(), sizeof(double) * 3 * h.npts);
    }
  }
  return HP3D_OK;
}

static int pbi_vec_batch(int space, int nel, const int *etype, const int *norder, const int *norie, const int *norif, int maxp, const double *etav,
                         int ncomp, const double *fval, const double *fcurl, long long f_ld, const unsigned *mask, double *dof,
                         long long dof_ld, int *info) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (int drc = enter_device()) return drc;
  if (nel < 0 || (nel > 0 && (!norder || !norie || !norif || !etav || !fval || (space == PBI_HCURL && !fcurl) || !dof))) return fail(HP3D_EINVAL, "pbi_hcurl/hdiv: null argument");
  if (ncomp < 1 || ncomp > PBI_MAXCOMP) return fail(HP3D_EINVAL, "pbi_hcurl: ncomp = %d outside 1..%d (split the components over several calls)", ncomp, PBI_MAXCOMP);
  if (nel == 0) return HP3D_OK;
  if (int rc = pbi_opt_in_smem()) return rc;
  pbi_trim_table_cache();
  { std::string perr; if (pbi_precompile(nel, etype, norder, norie, norif, 1, maxp, true, space, perr)) return fail(HP3D_EINVAL, "%s", perr.c_str()); }
  struct Group { PbiSig *S; std::vector<int> el; };
  std::map<const PbiSig *, size_t> where;
  std::vector<Group> groups;
  for (int e = 0; e < nel; e++) {
    std::string err;
    PbiSig *S = pbi_signature(etype ? etype[e] : HP3D_MDLB, norder + 19 * e, norie + 12 * e, norif + 6 * e, 1, maxp, true, err, space);
    if (!S) return fail(HP3D_EINVAL, "element %d: %s", e, err.c_str());
    if (dof_ld < (long long)ncomp * S->h.nEF) return fail(HP3D_EINVAL, "pbi_hcurl: dof_ld %lld < ncomp*(edge+face dofs) = %d (element %d)", dof_ld, ncomp * S->h.nEF, e);
    if (f_ld < 3LL * ncomp * S->h.npts) return fail(HP3D_EINVAL, "pbi_hcurl: f_ld %lld < 3*ncomp*npts = %d (element %d)", f_ld, 3 * ncomp * S->h.npts, e);
    auto w = where.find(S);
    if (w == where.end()) { w = where.emplace(S, groups.size()).first; groups.push_back(Group{S, {}}); }
    groups[w->second].el.push_back(e);
  }
  // workspace per CTA: rows [CE | E | GH | Rc | Rv] x 3 np doubles + the (nt + ncomp) x nt system
  struct LaunchDims { long long stride[2]; int ny[2]; };
  std::vector<LaunchDims> dims(groups.size());
  long long need_ws = 0;
  std::vector<int> elist; elist.reserve(nel);
  for (size_t g = 0; g < groups.size(); g++) {
    const PbiSigHost &h = groups[g].S->h;
    const int n = (int)groups[g].el.size();
    const int first[3] = {h.nrv, h.nrv + h.nre, h.nrv + h.nre + h.nrf};
    for (int s = 0; s < 2; s++) {
      long long st = 0;
      for (int i = first[s]; i < first[s + 1]; i++) {
        const PbiNode &nd = h.node[i];
        if (nd.n <= 0) continue;
        const bool saddle = s == 1 && space == PBI_HCURL;   // H(curl) faces carry curl rows and curl residuals besides the value rows
        const long long nt = nd.n + nd.nh, rows = (saddle ? 2LL * nd.n : nd.n) + nd.nh + (saddle ? 2 : 1) * ncomp;
        st = std::max(st, rows * 3LL * nd.np + (nt + ncomp) * nt);
      }
      dims[g].stride[s] = st; dims[g].ny[s] = 0;
      if (!st) continue;
      const long long per_row = st * (first[s + 1] - first[s]) * (long long)sizeof(double);
      dims[g].ny[s] = (int)std::max(1LL, std::min(std::min((long long)n, 65535LL), (1LL << 30) / per_row));
      if (st * (long long)sizeof(double) > PBI_SMALL_BYTES) need_ws = std::max(need_ws, per_row * dims[g].ny[s]);
    }
    elist.insert(elist.end(), groups[g].el.begin(), groups[g].el.end());
  }
  const size_t b_ev = sizeof(double) * 24 * (size_t)nel, b_f = std::max(sizeof(double) * (size_t)f_ld * nel, sizeof(double)), b_d = sizeof(double) * (size_t)dof_ld * nel;
  double *dev_ev = (double *)g_pbi_store.get(0, b_ev), *dev_fv = (double *)g_pbi_store.get(1, b_f), *dev_fc = (double *)g_pbi_store.get(2, b_f),
         *dev_d = (double *)g_pbi_store.get(3, std::max(b_d, sizeof(double))), *dev_ws = (double *)g_pbi_store.get(4, (size_t)std::max(need_ws, 8LL));
  unsigned *dev_m = mask ? (unsigned *)g_pbi_store.get(5, sizeof(unsigned) * nel) : nullptr;
  int *dinfo = (int *)g_pbi_store.get(6, sizeof(int) * nel), *dev_el = (int *)g_pbi_store.get(7, sizeof(int) * nel);
  if (!dev_ev || !dev_fv || !dev_fc || !dev_d || !dev_ws || (mask && !dev_m) || !dinfo || !dev_el)
    return fail(HP3D_ENOMEM, "pbi_hcurl: out of device memory (%zu bytes of inputs, %lld bytes of workspace)", b_ev + 2 * b_f + b_d, need_ws);
  CUDA_TRY(cudaMemcpyAsync(dev_ev, etav, b_ev, cudaMemcpyHostToDevice, g_compute));
  if (f_ld > 0) {
    CUDA_TRY(cudaMemcpyAsync(dev_fv, fval, sizeof(double) * (size_t)f_ld * nel, cudaMemcpyHostToDevice, g_compute));
    if (fcurl) CUDA_TRY(cudaMemcpyAsync(dev_fc, fcurl, sizeof(double) * (size_t)f_ld * nel, cudaMemcpyHostToDevice, g_compute));
  }
  if (b_d) CUDA_TRY(cudaMemcpyAsync(dev_d, dof, b_d, cudaMemcpyHostToDevice, g_compute));   // edges outside the mask keep (and contribute) their dofs
  if (mask) CUDA_TRY(cudaMemcpyAsync(dev_m, mask, sizeof(unsigned) * nel, cudaMemcpyHostToDevice, g_compute));
  CUDA_TRY(cudaMemcpyAsync(dev_el, elist.data(), sizeof(int) * nel, cudaMemcpyHostToDevice, g_compute));
  CUDA_TRY(cudaMemsetAsync(dinfo, 0, sizeof(int) * nel, g_compute));
  size_t pos = 0;
  for (size_t g = 0; g < groups.size(); g++) {
    const PbiSig &S = *groups[g].S;
    const PbiSigHost &h = S.h;
    const int n = (int)groups[g].el.size();
    const int first[3] = {h.nrv, h.nrv + h.nre, h.nrv + h.nre + h.nrf};
    PbiEArgs A;
    A.wa = S.d_wa; A.tan = S.d_tan; A.grad = S.d_grad; A.tabE = S.d_tabE; A.nodes = S.d_nodes; A.nH = h.nH; A.nEF = h.nEF; A.nrv = h.nrv; A.npts = h.npts;
    A.ncomp = ncomp; A.node0 = 0; A.nel = n; A.space = space; A.elems = dev_el + pos; A.etav = dev_ev; A.fval = dev_fv; A.fcurl = dev_fc; A.mask = dev_m; A.dof = dev_d;
    A.f_ld = f_ld; A.dof_ld = dof_ld; A.ws = dev_ws; A.ws_stride = 0; A.g_in_smem = 0; A.info = dinfo;
    for (int s = 0; s < 2; s++) {   // edges, then faces
      if (!dims[g].stride[s]) continue;
      A.node0 = first[s]; A.ws_stride = dims[g].stride[s];
      const size_t small_bytes = sizeof(double) * (size_t)dims[g].stride[s];
      if (small_bytes <= (size_t)PBI_SMALL_BYTES) pbi_hcurl_kernel<true><<<dim3(first[s + 1] - first[s], std::min(n, 65535)), 64, small_bytes, g_compute>>>(A);
      else {
        size_t gbytes = 0;
        for (int i = first[s]; i < first[s + 1]; i++) { const size_t nt = (size_t)h.node[i].n + h.node[i].nh; gbytes = std::max(gbytes, sizeof(double) * (nt + ncomp) * nt); }
        A.g_in_smem = gbytes <= (size_t)PBI_GSMEM_BYTES;
        pbi_hcurl_kernel<false><<<dim3(first[s + 1] - first[s], dims[g].ny[s]), 256, A.g_in_smem ? gbytes : 0, g_compute>>>(A);
      }
      g_launches++;
    }
    pos += n;
  }
  CUDA_TRY(cudaGetLastError());
  std::vector<int> hinfo(info ? 0 : nel);
  if (b_d) CUDA_TRY(cudaMemcpyAsync(dof, dev_d, b_d, cudaMemcpyDeviceToHost, g_compute));
  CUDA_TRY(cudaMemcpyAsync(info ? info : hinfo.data(), dinfo, sizeof(int) * nel, cudaMemcpyDeviceToHost, g_compute));
  CUDA_TRY(cudaStreamSynchronize(g_compute));
  return HP3D_OK;
}

}  // extern "C"

extern "C" {
int hp3d_gpu_pbi_hcurl_points(int nel, const int *etype, const int *norder, const int *norie, const int *norif, int maxp, double *xi,
                              long long xi_ld, int *npts, int *nrdofE, int *nodes) {
  return pbi_vec_points(PBI_HCURL, nel, etype, norder, norie, norif, maxp, xi, xi_ld, npts, nrdofE, nodes);
}
int hp3d_gpu_pbi_hcurl_batch(int nel, const int *etype, const int *norder, const int *norie, const int *norif, int maxp, const double *etav,
                             int ncomp, const double *fval, const double *fcurl, long long f_ld, const unsigned *mask, double *dof,
                             long long dof_ld, int *info) {
  return pbi_vec_batch(PBI_HCURL, nel, etype, norder, norie, norif, maxp, etav, ncomp, fval, fcurl, f_ld, mask, dof, dof_ld, info);
}
int hp3d_gpu_pbi_hdiv_points(int nel, const int *etype, const int *norder, const int *norie, const int *norif, int maxp, double *xi,
                             long long xi_ld, int *npts, int *nrdofV, int *nodes) {
  return pbi_vec_points(PBI_HDIV, nel, etype, norder, norie, norif, maxp, xi, xi_ld, npts, nrdofV, nodes);
}
int hp3d_gpu_pbi_hdiv_batch(int nel, const int *etype, const int *norder, const int *norie, const int *norif, int maxp, const double *etav,
                            int ncomp, const double *fval, long long f_ld, const unsigned *mask, double *dof, long long dof_ld, int *info) {
  return pbi_vec_batch(PBI_HDIV, nel, etype, norder, norie, norif, maxp, etav, ncomp, fval, nullptr, f_ld, mask, dof, dof_ld, info);
}
}  // extern "C"

extern "C" int hp3d_gpu_pbi_cache_limit(long long bytes) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (bytes < 0) return fail(HP3D_EINVAL, "pbi_cache_limit: negative size");
  PBI_TABLE_CACHE_BYTES = (size_t)bytes;
  return HP3D_OK;
}

extern "C" int hp3d_gpu_chunk_plan_debug(long long ntot, int cap, int nlanes, int max_chunk, long long *sizes, int cap_sizes) {
  if (ntot < 0 || cap < 1 || nlanes < 1) return fail(HP3D_EINVAL, "chunk_plan: bad argument");
  const std::vector<size_t> v = chunk_plan((size_t)ntot, cap, nlanes, max_chunk);
  if (sizes) for (size_t i = 0; i < v.size() && (int)i < cap_sizes; i++) sizes[i] = (long long)v[i];
  return (int)v.size();
}
