// engine.cuh -- batch engine behind the C ABI: per-signature device programs, workspaces, the per-chunk pipeline
//   descriptors -> geometry/weight fields -> sum-factorised integration -> dense phase -> condensed outputs.
#pragma once
#include "celem_kernels.cuh"
#include "dense_pipeline.cuh"
#include "formats.cuh"
#include "forms.hpp"
#include "forms_prism.hpp"
#include "integ_kernels.cuh"

#include <atomic>
#include <complex>
#include <map>
#include <thread>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace hp3d {

#define HP3D_CK(x)                                                                          \
  do {                                                                                      \
    cudaError_t e_ = (x);                                                                   \
    if (e_ != cudaSuccess) { err = std::string(#x) + ": " + cudaGetErrorString(e_); return -2; } \
  } while (0)

static long long g_launches = 0;  // kernels launched by this library (reported by hp3d_gpu_bench)

template <class T> static int dev_upload(const std::vector<T> &h, T **d, std::string &err) {
  *d = nullptr;
  if (h.empty()) return 0;
  HP3D_CK(cudaMalloc((void **)d, sizeof(T) * h.size()));
  HP3D_CK(cudaMemcpy(*d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
  return 0;
}

// One device allocation + one pinned host allocation shared by ALL signatures of the process: element groups are
// processed one after the other, so a signature only *binds* its chunk buffers into the arena (pointer arithmetic, no
// cudaMalloc / cudaFree per signature).  hp meshes produce thousands of signatures (order x orientation combinations).
struct Arena {
  char *d = nullptr, *h = nullptr;
  size_t dcap = 0, hcap = 0;
  const void *owner = nullptr;   // signature whose buffers are currently bound
  int owner_batch = 0;
  int ensure(size_t dbytes, size_t hbytes, std::string &err) {
    if (dbytes > dcap) {
      cudaDeviceSynchronize();
      cudaFree(d); d = nullptr; dcap = 0; owner = nullptr;
      HP3D_CK(cudaMalloc((void **)&d, dbytes));
      dcap = dbytes;
    }
    if (hbytes > hcap) {
      cudaDeviceSynchronize();
      cudaFreeHost(h); h = nullptr; hcap = 0; owner = nullptr;
      HP3D_CK(cudaMallocHost((void **)&h, hbytes));
      hcap = hbytes;
    }
    return 0;
  }
  void release() { cudaFree(d); cudaFreeHost(h); d = h = nullptr; dcap = hcap = 0; owner = nullptr; }
};
static Arena g_arena;

struct Bump {   // 256-byte aligned carving; with base == nullptr it only measures
  char *base; size_t off = 0;
  explicit Bump(char *b) : base(b) {}
  template <class T> T *take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T *p = base ? (T *)(base + off) : nullptr;
    off += sizeof(T) * n;
    return p;
  }
};

struct DenseWorkspace {
  DenseDims d;
  DenseBuffers b;
  char *owned_base = nullptr;   // non-null: stand-alone allocation (test hook); otherwise the buffers are bound into the arena
  void release() { cudaFree(owned_base); owned_base = nullptr; b = DenseBuffers{}; }
  void bind(const DenseDims &dims, int batch, Bump &m) {
    d = dims;
    const size_t P = d.planes(), lp = d.linv_plane(), ns = d.nsteps_stc() ? d.nsteps_stc() : 1;
    b.W = d.dpg ? m.take<double>(P * d.w_plane() * batch) : nullptr;
    b.Am = m.take<double>(P * d.a_plane() * batch);
    b.LH = m.take<double>(P * (d.lh_plane() ? d.lh_plane() : 1) * batch);
    b.Linv = m.take<double>(P * lp * batch);
    b.LinvH = m.take<double>(P * lp * batch);
    b.LinvS = m.take<double>(P * lp * ns * batch);
    b.LinvSH = m.take<double>(P * lp * ns * batch);
    b.info = m.take<int>(batch);
  }
  int reserve(const DenseDims &dims, int batch, std::string &err) {   // stand-alone allocation (dense_debug_run)
    release();
    Bump meas(nullptr);
    bind(dims, batch, meas);
    char *base = nullptr;
    HP3D_CK(cudaMalloc((void **)&base, meas.off + 256));
    Bump m(base);
    bind(dims, batch, m);
    owned_base = base;
    return 0;
  }
};

// ------------------------------------------------------------------------------------------------
// One element signature (element type + node orders + orientations) compiled and resident on the device: only the
// small tables the integration kernels read.  Chunk buffers belong to the "dense class" (below), not to the signature.
struct Signature {
  SigHost h;
  char *d_blob = nullptr;   // ONE device allocation holding all tables of the signature (hp meshes have thousands of signatures)
  double *d_tab = nullptr, *d_wq = nullptr, *d_CW = nullptr, *d_ttab = nullptr;
  int *d_hdof = nullptr, *d_maps = nullptr, *d_crow = nullptr;
  FamilyDesc *d_fam = nullptr; TermDesc *d_term = nullptr; SlotDesc *d_slot = nullptr; BlockDesc *d_block = nullptr; WorkItem *d_work = nullptr;
  std::shared_ptr<char> slab;   // device allocation shared with the other signatures uploaded by the same call (upload_many)
  size_t off_[12] = {0};
  ~Signature() { cudaFree(d_blob); }
  int ns() const { return h.cplx ? 2 : 1; }
  size_t src_doubles() const { return (size_t)h.nint * (h.cplx ? 6 : 1); }
  // the 12 table sections of the signature: (host address, bytes)
  void sections(const void *src[12], size_t bytes[12]) const {
    src[0] = h.tab.data(); bytes[0] = sizeof(double) * h.tab.size(); src[1] = h.wq.data(); bytes[1] = sizeof(double) * h.wq.size();
    src[2] = h.hdof.data(); bytes[2] = sizeof(int) * h.hdof.size(); src[3] = h.maps.data(); bytes[3] = sizeof(int) * h.maps.size();
    src[4] = h.fam.data(); bytes[4] = sizeof(FamilyDesc) * h.fam.size(); src[5] = h.term.data(); bytes[5] = sizeof(TermDesc) * h.term.size();
    src[6] = h.slot.data(); bytes[6] = sizeof(SlotDesc) * h.slot.size(); src[7] = h.block.data(); bytes[7] = sizeof(BlockDesc) * h.block.size();
    src[8] = h.work.data(); bytes[8] = sizeof(WorkItem) * h.work.size(); src[9] = h.crow.data(); bytes[9] = sizeof(int) * h.crow.size();
    src[10] = h.CW.data(); bytes[10] = sizeof(double) * h.CW.size(); src[11] = h.ttab.data(); bytes[11] = sizeof(double) * h.ttab.size();
  }
  // reserve room for the sections behind `off` bytes (256-byte aligned each); returns the new end
  size_t plan_offsets(size_t off) {
    const void *src[12]; size_t bytes[12];
    sections(src, bytes);
    for (int i = 0; i < 12; i++) { off = (off + 255) & ~(size_t)255; off_[i] = off; off += bytes[i]; }
    return off;
  }
  int copy_sections(char *base, std::string &err) const {   // straight from the host tables (no intermediate blob: the trace-pairing rows CW are tens of MB)
    const void *src[12]; size_t bytes[12];
    sections(src, bytes);
    for (int i = 0; i < 12; i++)
      if (bytes[i]) HP3D_CK(cudaMemcpy(base + off_[i], src[i], bytes[i], cudaMemcpyHostToDevice));
    return 0;
  }
  void bind(char *base) {   // base = device address of the blob's first byte
    auto at = [&](size_t off, size_t n) -> char * { return n ? base + off : nullptr; };
    d_tab = (double *)at(off_[0], h.tab.size()); d_wq = (double *)at(off_[1], h.wq.size()); d_hdof = (int *)at(off_[2], h.hdof.size());
    d_maps = (int *)at(off_[3], h.maps.size()); d_fam = (FamilyDesc *)at(off_[4], h.fam.size()); d_term = (TermDesc *)at(off_[5], h.term.size());
    d_slot = (SlotDesc *)at(off_[6], h.slot.size()); d_block = (BlockDesc *)at(off_[7], h.block.size()); d_work = (WorkItem *)at(off_[8], h.work.size());
    d_crow = (int *)at(off_[9], h.crow.size()); d_CW = (double *)at(off_[10], h.CW.size()); d_ttab = (double *)at(off_[11], h.ttab.size());
  }
  int upload(std::string &err) {
    const size_t total = plan_offsets(0);
    HP3D_CK(cudaMalloc((void **)&d_blob, total + 256));
    if (copy_sections(d_blob, err)) return -1;
    bind(d_blob);
    return 0;
  }
  // all tables of the signatures a call meets for the first time in ONE device allocation (an hp mesh brings hundreds of
  // signatures per call)
  static int upload_many(const std::vector<Signature *> &sigs, std::string &err) {
    if (sigs.empty()) return 0;
    size_t total = 0;
    for (Signature *S : sigs) total = S->plan_offsets(total);
    char *base = nullptr;
    HP3D_CK(cudaMalloc((void **)&base, total + 256));
    std::shared_ptr<char> slab(base, [](char *q) { cudaFree(q); });
    for (Signature *S : sigs) { if (S->copy_sections(base, err)) return -1; S->bind(base); S->slab = slab; }
    return 0;
  }
};

// ------------------------------------------------------------------------------------------------
// Dense class: all signatures whose PADDED dense-phase extents (np, nbp, nip) agree run through the dense phase in the
// same batch, whatever their element type, orders below the padding granularity and orientations -- this is the
// heterogeneous batching of hp meshes.  ChunkShape carries the class extents plus the per-element maxima that size the
// chunk buffers.
struct ChunkShape {
  DenseDims d;                 // np/nbp/nip of the class; n/nb/ni = maxima over the members seen in this call
  bool gen_stc = false;
  int nint_max = 0, nH_max = 0;
  size_t src_max = 0;          // doubles per element of a caller-supplied source table
  size_t nz_max = 0, nc_max = 0;   // hp3d_gpu_celem_batch only: scalars of Zastif / Zbload per element (maxima of the group)
  bool coo = false;                //   ... and IRN/JCN staging
  int ns() const { return (d.cplx || d.rs) ? 2 : 1; }   // doubles per scalar of the caller's value type
  bool covers(const ChunkShape &o) const {
    return d.cplx == o.d.cplx && d.rs == o.d.rs && d.nload == o.d.nload && d.dpg == o.d.dpg && gen_stc == o.gen_stc && d.np == o.d.np && d.nbp == o.d.nbp && d.nip == o.d.nip &&
           d.ni >= o.d.ni && d.nb >= o.d.nb && nint_max >= o.nint_max && nH_max >= o.nH_max && src_max >= o.src_max &&
           nz_max >= o.nz_max && nc_max >= o.nc_max && (coo || !o.coo);
  }
  void absorb(const SigHost &h) {
    if (d.np == 0 && d.nip == 0) { d = h.dims; gen_stc = h.gen_stc; }
    d.n = std::max(d.n, h.dims.n); d.nb = std::max(d.nb, h.nb); d.ni = std::max(d.ni, h.ni);
    d.nip = std::max(d.nip, h.dims.nip); d.nil = std::max(d.nil, h.dims.nil);   // classes merged across nip (Cholesky path): every element keeps its own row layout (nip_e)
    nint_max = std::max(nint_max, h.nint); nH_max = std::max(nH_max, h.nH);
    src_max = std::max(src_max, (size_t)h.nint * (h.cplx ? 6 : 1) * (size_t)std::max(1, h.dims.nrhs()));   // NR_RHS tables per element
  }
  // classes may merge across nip when the condensation is the Cholesky pipeline (the load rows are located per element);
  // the pivoted-LU kernel addresses the load column through the class extent, so its classes keep nip in the key
  static bool nip_mergeable(const SigHost &h) { return !h.gen_stc; }
  static std::string key(const SigHost &h) {   // base key: everything but nip
    char b[96];
    snprintf(b, sizeof b, "%d/%d/%d/%d/%d/%d/%d", (int)h.cplx, (int)h.dims.rs, (int)h.dpg, (int)h.gen_stc, h.dims.np, h.dims.nbp, nip_mergeable(h) ? -1 : h.dims.nip);
    return b;
  }
};

// A lane = one complete set of per-chunk buffers.  The lanes (four in hp3d_gpu_elem_batch) run on their own streams so that the latency-bound steps of one
// chunk (64x64 tile factorizations, launch tails) overlap the GEMMs of the other, and D2H of a finished chunk overlaps compute.
struct Lane {
  DenseWorkspace ws;
  double *d_WF = nullptr, *d_xnod = nullptr, *d_src = nullptr;         // chunk inputs / weight fields
  struct Out {
    double *Aii = nullptr, *Bi = nullptr, *AS = nullptr, *BS = nullptr; int *info = nullptr, *h_info = nullptr;   // h_info: pinned
    double *Z = nullptr, *zb = nullptr; int *irn = nullptr, *jcn = nullptr;   // compressed system (hp3d_gpu_celem_batch)
  };
  Out out[2];   // chunk outputs (device staging), double-buffered: the D2H of one chunk overlaps the lane's next chunk
  double *h_xnod = nullptr, *h_src = nullptr;                          // pinned host staging of the chunk inputs
  int *h_cnt = nullptr;                                                // pinned [3][batch]: ni_e | nb_e | nip_e
  int *d_cel = nullptr, *h_cel = nullptr;                              // caller element index of each slot (celem mode)
  // back-substitution / residual modes: solution dofs in (xi | xb), results out (xb or one residual per element)
  double *d_xi = nullptr, *d_xb = nullptr, *d_res = nullptr, *h_xi = nullptr, *h_xb = nullptr, *h_res = nullptr;
};
enum ChunkMode { MODE_ELEM = 0, MODE_BWD = 1, MODE_RESID = 2, MODE_CELEM = 3 };
struct LaneSet {
  static constexpr int NLANE = 4;   // lanes that can be bound; hp3d_gpu_elem_batch uses two
  Lane lane[NLANE];
  ChunkShape shape;
  int cap = 0;   // elements per lane currently bound in the arena
  int nl = 2;    // lanes currently bound
  double *d_ones = nullptr; int *d_iota = nullptr; int n_iota = 0;   // identity output maps (grown on demand, plain cudaMalloc)
  void layout(const ChunkShape &sh, int batch, Bump &dm, Bump &hm, int nlanes = 2) {
    for (int i = 0; i < nlanes; i++) layout_lane(lane[i], sh, batch, dm, hm);
  }
  // Lane i bound on its own inside partition i of the arena (NLANE equal partitions), whatever the other lanes hold: chunks of
  // DIFFERENT dense classes run side by side on different lanes without draining each other (hp meshes).  The caller makes
  // sure the partitions are large enough (ensure_partitions) and orders re-binds by the lane's stream.
  static int ensure_partitions(size_t dev_bytes_per_lane, size_t host_bytes_per_lane, std::string &err) {
    const size_t dneed = (dev_bytes_per_lane + 4096) * NLANE, hneed = (host_bytes_per_lane + 4096) * NLANE;
    return g_arena.ensure(std::max(dneed, g_arena.dcap), std::max(hneed, g_arena.hcap), err);
  }
  static void lane_bytes(const ChunkShape &sh, int batch, size_t &dev, size_t &host) {
    Bump dm(nullptr), hm(nullptr);
    Lane tmp;
    layout_lane(tmp, sh, batch, dm, hm);
    dev = dm.off + 256; host = hm.off + 256;
  }
  void bind_lane(int i, const ChunkShape &sh, int batch) {
    const size_t dpart = (g_arena.dcap / NLANE) & ~(size_t)255, hpart = (g_arena.hcap / NLANE) & ~(size_t)255;
    Bump dm((char *)g_arena.d + dpart * i), hm((char *)g_arena.h + hpart * i);
    layout_lane(lane[i], sh, batch, dm, hm);
    g_arena.owner = nullptr; cap = 0;   // the shared single-shape layout (reserve) is no longer valid
  }
  static void layout_lane(Lane &L, const ChunkShape &sh, int batch, Bump &dm, Bump &hm) {
    const size_t NS = sh.ns();
    const DenseDims &d = sh.d;
    const size_t nr = (size_t)std::max(1, d.nrhs());   // NR_RHS columns of Bi / BSchur / xi / xb
    {
      L.ws.bind(d, batch, dm);
      L.ws.b.ni_e = dm.take<int>(batch); L.ws.b.nb_e = dm.take<int>(batch); L.ws.b.nip_e = dm.take<int>(batch);
      L.d_WF = dm.take<double>((size_t)NFIELD * wf_stride(sh.nint_max) * batch);
      L.d_xnod = dm.take<double>((size_t)3 * sh.nH_max * batch);
      L.d_src = dm.take<double>(sh.src_max * batch);
      for (int o = 0; o < 2; o++) {
        L.out[o].Aii = dm.take<double>(NS * (size_t)d.ni * d.ni * batch);
        L.out[o].Bi = dm.take<double>(NS * (size_t)d.ni * nr * batch);
        L.out[o].AS = dm.take<double>(NS * ((size_t)d.nb * d.ni + 1) * batch);
        L.out[o].BS = dm.take<double>(NS * ((size_t)d.nb * nr + 1) * batch);
        L.out[o].info = dm.take<int>(batch);
        L.out[o].h_info = hm.take<int>(batch);
        L.out[o].Z = dm.take<double>(NS * sh.nz_max * batch);
        L.out[o].zb = dm.take<double>(NS * sh.nc_max * batch);
        L.out[o].irn = dm.take<int>(sh.coo ? sh.nz_max * batch : 0);
        L.out[o].jcn = dm.take<int>(sh.coo ? sh.nz_max * batch : 0);
      }
      L.h_xnod = hm.take<double>((size_t)3 * sh.nH_max * batch);
      L.h_src = hm.take<double>(sh.src_max * batch);
      L.h_cnt = hm.take<int>(3 * (size_t)batch);
      L.d_cel = dm.take<int>(batch); L.h_cel = hm.take<int>(batch);
      L.d_xi = dm.take<double>(NS * (size_t)d.ni * nr * batch); L.d_xb = dm.take<double>(NS * ((size_t)d.nb * nr + 1) * batch);
      L.d_res = dm.take<double>(batch);
      L.h_xi = hm.take<double>(NS * (size_t)d.ni * nr * batch); L.h_xb = hm.take<double>(NS * ((size_t)d.nb * nr + 1) * batch);
      L.h_res = hm.take<double>(batch);
    }
  }
  size_t bytes_per_element(const ChunkShape &sh) {   // device bytes per element of ONE lane (alignment slack excluded)
    Bump d1(nullptr), h1(nullptr), d2(nullptr), h2(nullptr);
    LaneSet tmp;
    tmp.layout(sh, 1, d1, h1, 1); tmp.layout(sh, 65, d2, h2, 1);
    return (d2.off - d1.off) / 64 + 1;
  }
  // bind the chunk buffers for `batch` elements per lane of shape `sh` into the shared arena (no-op if still bound)
  int reserve(const ChunkShape &sh, int batch, std::string &err, int nlanes = 2) {
    if (cap >= batch && nl >= nlanes && shape.covers(sh) && g_arena.owner == this) return 0;
    Bump dmeas(nullptr), hmeas(nullptr);
    LaneSet tmp;
    tmp.layout(sh, batch, dmeas, hmeas, nlanes);
    size_t dneed = dmeas.off + 256, hneed = hmeas.off + 256;
    if (dneed > g_arena.dcap) dneed = std::max(dneed, g_arena.dcap + g_arena.dcap / 4);   // geometric growth
    if (hneed > g_arena.hcap) hneed = std::max(hneed, g_arena.hcap + g_arena.hcap / 4);
    if (int rc = g_arena.ensure(dneed, hneed, err)) { cap = 0; return rc; }
    Bump dm(g_arena.d), hm(g_arena.h);
    layout(sh, batch, dm, hm, nlanes);
    nl = nlanes;
    if (ensure_iota(std::max(sh.d.ni, sh.d.nb) + 1, err)) return -2;
    g_arena.owner = this; shape = sh; cap = batch;
    return 0;
  }
  int ensure_iota(int need, std::string &err) {   // identity output maps for at least `need` dofs
    if (need > n_iota) {
      cudaDeviceSynchronize();
      cudaFree(d_ones); cudaFree(d_iota);
      std::vector<int> iota(need); std::vector<double> ones(need, 1.0);
      for (int i = 0; i < need; i++) iota[i] = i;
      if (dev_upload(iota, &d_iota, err) || dev_upload(ones, &d_ones, err)) return -2;
      n_iota = need;
    }
    return 0;
  }
  void release() { cudaFree(d_ones); cudaFree(d_iota); d_ones = nullptr; d_iota = nullptr; n_iota = 0; cap = 0; shape = ChunkShape(); }
};
static LaneSet g_lanes;

// a run of consecutive chunk slots [start, start+n) holding elements of ONE signature
struct Seg { Signature *S; int start, n; };

template <int NMAX> static cudaError_t tp3_configure() {
  return cudaFuncSetAttribute(tp3_kernel<NMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}
template <int NMAX> static cudaError_t tp2_configure() {
  return cudaFuncSetAttribute(tp2_kernel<NMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}
static void launch_tp3(const Signature &S, const Tp3Args &A, int nel, cudaStream_t st) {
  dim3 grid((unsigned)S.h.work.size(), nel);
  const int off = (int)S.h.smem_u_off, offF = (int)S.h.smem_f_off, offT1 = (int)S.h.smem_t1_off, offQ = (int)S.h.smem_q_off;
  if (S.h.etype == 3) {   // prism: (triangle list) x (z table) families
    switch (S.h.nmax) {
      case 4: tp2_kernel<4><<<grid, S.h.threads, S.h.smem_bytes, st>>>(A, S.d_ttab, off); break;
      case 6: tp2_kernel<6><<<grid, S.h.threads, S.h.smem_bytes, st>>>(A, S.d_ttab, off); break;
      case 8: tp2_kernel<8><<<grid, S.h.threads, S.h.smem_bytes, st>>>(A, S.d_ttab, off); break;
      default: tp2_kernel<10><<<grid, S.h.threads, S.h.smem_bytes, st>>>(A, S.d_ttab, off); break;
    }
    g_launches++;
    return;
  }
  switch (S.h.nmax) {
    case 4: tp3_kernel<4><<<grid, S.h.threads, S.h.smem_bytes, st>>>(A, offF, offT1, off, offQ); break;
    case 6: tp3_kernel<6><<<grid, S.h.threads, S.h.smem_bytes, st>>>(A, offF, offT1, off, offQ); break;
    case 8: tp3_kernel<8><<<grid, S.h.threads, S.h.smem_bytes, st>>>(A, offF, offT1, off, offQ); break;
    default: tp3_kernel<10><<<grid, S.h.threads, S.h.smem_bytes, st>>>(A, offF, offT1, off, offQ); break;
  }
  g_launches++;
}

struct StageEvents { cudaEvent_t e[4]; bool on = false; };  // start, after integration, after dense, after scatter

// Integration of the `nel` elements of a chunk (segments of equal signature) into the dense phase's input buffers.
// Element slot i reads its geometry dofs at d_xnod + i*xnod_ld (and its source table at d_src + i*src_ld).
static void run_integration(const ChunkShape &sh, Lane &L, const GeomParams &gp, const std::vector<Seg> &segs, int nel, const double *d_xnod,
                            long long xnod_ld, const double *d_src, long long src_ld, cudaStream_t st) {
  const DenseDims &d = sh.d;
  const long long P = d.cplx ? 2 : 1;
  cudaMemsetAsync(L.ws.b.info, 0, sizeof(int) * nel, st);
  if (d.dpg) { dim3 g((d.R() + 15) / 16, (unsigned)(nel * P)); zero_w_kernel<<<g, 256, 0, st>>>(L.ws.b.W, (long long)d.w_plane(), d.R(), d.np); g_launches++; }
  else cudaMemsetAsync(L.ws.b.Am, 0, sizeof(double) * P * d.a_plane() * nel, st);
  size_t wf_off = 0;
  for (const Seg &sg_ : segs) {
    Signature &S = *sg_.S;
    const SigHost &h = S.h;
    SigTables sg;
    sg.tab = S.d_tab; sg.wq = S.d_wq; sg.hdof = S.d_hdof; sg.nH = h.nH; sg.nint = h.nint;
    sg.ttab = S.d_ttab ? S.d_ttab + h.geo_toff : nullptr; sg.nT = h.geo_nT;
    for (int i = 0; i < 3; i++) sg.nq[i] = h.nq[i];
    double *WF = L.d_WF + wf_off;
    wf_off += (size_t)NFIELD * wf_stride(h.nint) * sg_.n;
    const double *xn = d_xnod + (long long)sg_.start * xnod_ld, *src = d_src ? d_src + (long long)sg_.start * src_ld : nullptr;
    int *info = L.ws.b.info + sg_.start;
    const long long npts = (long long)sg_.n * h.nint;
    if (h.etype == 3) geom_fields_prism_kernel<<<(unsigned)((npts + 127) / 128), 128, 0, st>>>(sg, gp, sg_.n, xn, xnod_ld, src, src_ld, WF, info);
    else geom_fields_kernel<<<(unsigned)((npts + 127) / 128), 128, 0, st>>>(sg, gp, sg_.n, xn, xnod_ld, src, src_ld, WF, info);
    g_launches++;
    Tp3Args A;
    A.tab = S.d_tab; A.fam = S.d_fam; A.term = S.d_term; A.slot = S.d_slot; A.block = S.d_block; A.work = S.d_work; A.maps = S.d_maps;
    A.WF = WF; A.nint = h.nint;
    for (int i = 0; i < 3; i++) A.nq[i] = h.nq[i];
    const long long wb = P * (long long)d.w_plane(), ab = P * (long long)d.a_plane();
    A.mat[0] = MatTarget{d.dpg ? L.ws.b.W + wb * sg_.start : nullptr, wb, (long long)d.w_plane(), d.np};
    A.mat[1] = MatTarget{L.ws.b.Am + ab * sg_.start, ab, (long long)d.a_plane(), d.M()};
    if (d.dpg && d.np > h.dims.n) { dim3 g((d.np - h.dims.n + 63) / 64, sg_.n); unit_diag_kernel<<<g, 64, 0, st>>>(A.mat[0], h.dims.n, d.np); g_launches++; }
    A.load_only = 0; A.load_shift = 0;
    launch_tp3(S, A, sg_.n, st);
    // NR_RHS > 1: the q-th load vector comes from the q-th source table of every element: the weight fields are rebuilt with
    // that source and only the load blocks are integrated again, their rows moved down to the q-th load's rows
    for (int q = 1; q < d.nrhs(); q++) {
      const double *srcq = src ? src + (size_t)q * h.nint * (h.cplx ? 6 : 1) : nullptr;
      if (h.etype == 3) geom_fields_prism_kernel<<<(unsigned)((npts + 127) / 128), 128, 0, st>>>(sg, gp, sg_.n, xn, xnod_ld, srcq, src_ld, WF, info);
      else geom_fields_kernel<<<(unsigned)((npts + 127) / 128), 128, 0, st>>>(sg, gp, sg_.n, xn, xnod_ld, srcq, src_ld, WF, info);
      g_launches++;
      A.load_only = 1; A.load_shift = q * (d.rs ? 2 : 1);
      launch_tp3(S, A, sg_.n, st);
    }
    if (!h.crow.empty()) {
      dim3 g((d.np + 255) / 256, (unsigned)h.crow.size(), sg_.n);
      const_rows_kernel<<<g, 256, 0, st>>>(S.d_CW, S.d_crow, d.np, A.mat[0]);
      g_launches++;
    }
  }
}

static long long dense_phase_launches(const DenseDims &d, bool want_z = true) {  // mirrors the launch structure of dense_phase()
  long long n = 0;
  auto chol = [&](int nt_r, int nt_c) { for (int j = 0; j < nt_c; j++) { if (j > 0) n++; n++; if (nt_r - j - 1 > 0) n++; } };
  if (d.dpg) { chol(d.R() / TILE, d.np / TILE); n++; }
  if (d.nbp == 0) return n;
  n++;
  chol(d.M() / TILE, d.nbp / TILE);
  n += 1;
  if (!want_z) return n;
  n += 1;
  const int ns = d.nsteps_stc();
  for (int j = ns - 1; j >= 0; j--) { if (j < ns - 1) n++; n++; }
  return n;
}

// CPLX: arithmetic of the dense phase; RS: real-structured complex problem (real dense phase, complex outputs: DenseDims::rs)
template <bool CPLX, bool RS>
static void run_dense_and_scatter(const ChunkShape &sh, Lane &L, const Lane::Out &o, int nel, bool want_schur, cudaStream_t st, StageEvents *ev,
                                  int mode, bool packed) {
  static_assert(!(CPLX && RS), "the real-structured path runs the dense phase in real arithmetic");
  constexpr bool OUTC = CPLX || RS;   // value type of the outputs
  const DenseDims &d = sh.d;
  if (mode == MODE_RESID) {   // uncondensed DPG system, then eta^2 = v^H A v
    dense_phase<CPLX>(d, L.ws.b, nel, st, true);
    if (RS) dpg_residual_rs_kernel<<<nel, 256, sizeof(double) * 2 * d.M(), st>>>(d, L.ws.b.Am, L.ws.b.ni_e, L.ws.b.nb_e, L.ws.b.nip_e, L.d_xi, (long long)d.ni, L.d_xb,
                                                                                (long long)d.nb + 1, L.d_res);
    else dpg_residual_kernel<CPLX><<<nel, 256, sizeof(double) * 2 * d.M(), st>>>(d, L.ws.b.Am, L.ws.b.ni_e, L.ws.b.nb_e, L.ws.b.nip_e, L.d_xi, (long long)d.ni, L.d_xb,
                                                                                (long long)d.nb + 1, L.d_res);
    g_launches += 2;
    cudaMemcpyAsync(o.info, L.ws.b.info, sizeof(int) * nel, cudaMemcpyDeviceToDevice, st);
    return;
  }
  if (sh.gen_stc) {   // pivoted-LU condensation: one kernel, writes the caller-layout outputs itself
    const long long P = CPLX ? 2 : 1;
    stc_gen_kernel<CPLX, RS><<<nel, 512, stc_gen_smem(d.M(), CPLX), st>>>(L.ws.b.nb_e, d.nbp, L.ws.b.ni_e, d.M(), L.ws.b.Am, (long long)d.a_plane(),
                                                                           P * (long long)d.a_plane(), o.Aii, o.Bi, o.AS, o.BS, (long long)d.ni * d.ni,
                                                                           (long long)d.ni, (long long)d.nb * d.ni, (long long)d.nb, want_schur ? 1 : 0,
                                                                           L.ws.b.info, stc_gen_block(d.M(), CPLX));
    g_launches++;
    if (ev && ev->on) cudaEventRecord(ev->e[2], st);
    cudaMemcpyAsync(o.info, L.ws.b.info, sizeof(int) * nel, cudaMemcpyDeviceToDevice, st);
    if (mode == MODE_BWD && d.nb > 0) {
      dim3 gb((d.nb + 7) / 8, nel);
      stc_bwd_kernel<OUTC><<<gb, 256, 0, st>>>(L.ws.b.ni_e, L.ws.b.nb_e, 0, 0, o.AS, (long long)d.nb * d.ni, o.BS, (long long)d.nb, L.d_xi, (long long)d.ni,
                                               L.d_xb, (long long)d.nb + 1);
      g_launches++;
    }
    return;
  }
  dense_phase<CPLX>(d, L.ws.b, nel, st, false, want_schur);
  g_launches += dense_phase_launches(d, want_schur);
  if (ev && ev->on) cudaEventRecord(ev->e[2], st);
  OutMaps mp{g_lanes.d_iota, g_lanes.d_iota, g_lanes.d_ones, g_lanes.d_ones, 0, 0, 0, 0, L.ws.b.ni_e, L.ws.b.nb_e, L.ws.b.nip_e};
  dim3 blk(16, 16), g1((d.ni + 15) / 16, (d.ni + 15) / 16, nel);
  const long long nr = std::max(1, d.nrhs());
  if (RS) scatter_condensed_rs_kernel<<<g1, blk, 0, st>>>(d, L.ws.b.Am, mp, o.Aii, o.Bi, (long long)d.ni * d.ni, (long long)d.ni * nr, packed ? 1 : 0);
  else scatter_condensed_kernel<CPLX><<<g1, blk, 0, st>>>(d, L.ws.b.Am, mp, o.Aii, o.Bi, (long long)d.ni * d.ni, (long long)d.ni * nr, packed ? 1 : 0);
  g_launches++;
  if (d.nb > 0 && want_schur) {
    dim3 g2((d.nb + 15) / 16, (d.ni + 15) / 16, nel);
    if (RS) scatter_schur_rs_kernel<<<g2, blk, 0, st>>>(d, L.ws.b.Am, mp, o.AS, o.BS, (long long)d.nb * d.ni, (long long)d.nb * nr);
    else scatter_schur_kernel<CPLX><<<g2, blk, 0, st>>>(d, L.ws.b.Am, mp, o.AS, o.BS, (long long)d.nb * d.ni, (long long)d.nb * nr);
    g_launches++;
  }
  cudaMemcpyAsync(o.info, L.ws.b.info, sizeof(int) * nel, cudaMemcpyDeviceToDevice, st);
  if (mode == MODE_BWD && d.nb > 0) {
    dim3 gb((d.nb + 7) / 8, nel);
    stc_bwd_kernel<OUTC><<<gb, 256, 0, st>>>(L.ws.b.ni_e, L.ws.b.nb_e, 0, 0, o.AS, (long long)d.nb * d.ni, o.BS, (long long)d.nb * nr, L.d_xi, (long long)d.ni * nr,
                                             L.d_xb, (long long)d.nb * nr + 1, (int)nr);
    g_launches++;
  }
}

// One chunk through the whole pipeline.  The caller has already queued the copy of the per-element dof counts into
// L.ws.b.ni_e / nb_e on `st`.
static void run_chunk(const ChunkShape &sh, Lane &L, int ob, const GeomParams &gp, const std::vector<Seg> &segs, int nel, const double *d_xnod,
                      long long xnod_ld, const double *d_src, long long src_ld, bool want_schur, cudaStream_t st, StageEvents *ev = nullptr,
                      int mode = MODE_ELEM, bool packed = false) {
  if (ev && ev->on) cudaEventRecord(ev->e[0], st);
  run_integration(sh, L, gp, segs, nel, d_xnod, xnod_ld, d_src, src_ld, st);
  if (ev && ev->on) cudaEventRecord(ev->e[1], st);
  if (sh.d.rs) run_dense_and_scatter<false, true>(sh, L, L.out[ob], nel, want_schur, st, ev, mode, packed);
  else if (sh.d.cplx) run_dense_and_scatter<true, false>(sh, L, L.out[ob], nel, want_schur, st, ev, mode, packed);
  else run_dense_and_scatter<false, false>(sh, L, L.out[ob], nel, want_schur, st, ev, mode, packed);
  if (ev && ev->on) cudaEventRecord(ev->e[3], st);
}

// The constraint data of one hp3d_gpu_celem_batch call, resident on the device for the duration of the call (buffers owned by
// the library's grow-only store).
struct CelemCall {
  long long *d_mptr = nullptr, *d_cptr = nullptr, *d_xptr = nullptr;
  int *d_cidx = nullptr, *d_dlist = nullptr, *d_nextract = nullptr, *d_lcon = nullptr;
  long long *d_dptr = nullptr;
  double *d_cval = nullptr, *d_zdofd = nullptr;
  const long long *xptr = nullptr;   // host
  std::vector<long long> aoff;       // host: scalar offset of every element's Zastif in the caller's array
  int isym = 2;
  void *zbload = nullptr, *zastif = nullptr; int *irn = nullptr, *jcn = nullptr;   // caller's (host) outputs
  long long nz(int e) const { const long long n = xptr[e + 1] - xptr[e]; return isym == 1 ? n * (n + 1) / 2 : n * n; }
};

// constraint transform + Dirichlet lift + compression of the `nel` condensed systems a chunk left in o.Aii / o.Bi
static void run_celem(const ChunkShape &sh, Lane &L, const Lane::Out &o, const CelemCall &cc, int nel, cudaStream_t st) {
  CelemArgs a;
  a.mptr = cc.d_mptr; a.cptr = cc.d_cptr; a.xptr = cc.d_xptr; a.cidx = cc.d_cidx; a.cval = cc.d_cval; a.zdofd = cc.d_zdofd;
  a.nextract = cc.d_nextract; a.lcon = cc.d_lcon; a.dptr = cc.d_dptr; a.dlist = cc.d_dlist;
  a.cel = L.d_cel; a.ni_e = L.ws.b.ni_e;
  a.Aii = o.Aii; a.Bi = o.Bi; a.sA = (long long)sh.d.ni * sh.d.ni; a.sB = sh.d.ni;
  a.Z = o.Z; a.zb = o.zb; a.irn = cc.irn ? o.irn : nullptr; a.jcn = cc.irn ? o.jcn : nullptr; a.sZ = (long long)sh.nz_max; a.sZb = (long long)sh.nc_max;
  a.isym = cc.isym;
  if (sh.nc_max == 0) return;
  const unsigned nt = (unsigned)((sh.nc_max + 31) / 32);
  dim3 gl((unsigned)((sh.nc_max + 127) / 128), nel), gc(nt, nt, nel), bc(32, 8);
  if (sh.ns() == 2) { celem_load_kernel<true><<<gl, 128, 0, st>>>(a); celem_compress_kernel<true><<<gc, bc, 0, st>>>(a); }
  else { celem_load_kernel<false><<<gl, 128, 0, st>>>(a); celem_compress_kernel<false><<<gc, bc, 0, st>>>(a); }
  g_launches += 2;
}

// ------------------------------------------------------------------------------------------------
struct Plan {
  FormParams fp;
  int store_schur = 1;
  int aii_packed = 0;   // hp3d_params.aii_packed: Aii returns as the packed lower triangle
  // signature cache; `mu` guards the map itself (nodes are never erased while the plan lives, so Signature pointers stay valid):
  // host-only size queries from other threads may run while a batch that holds the device engine is in flight
  std::map<std::string, std::unique_ptr<Signature>> sigs;
  std::mutex mu;
  Signature *find(const std::string &k) {
    std::lock_guard<std::mutex> lk(mu);
    auto it = sigs.find(k);
    return it == sigs.end() ? nullptr : it->second.get();
  }
  Signature *insert(const std::string &k, std::unique_ptr<Signature> s) {   // keeps the first one if two threads raced
    std::lock_guard<std::mutex> lk(mu);
    return sigs.emplace(k, std::move(s)).first->second.get();
  }
  GeomParams geom() const {
    GeomParams g; g.kind = fp.kind; g.source = fp.source; g.icomp = fp.icomp; g.omega = fp.omega; g.eps = fp.eps; g.mu = fp.mu; g.sigma = fp.sigma;
    g.tensor = fp.tensor ? 1 : 0;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        std::complex<double> m(0.0, 0.0);   // (eps_t eps_t^H)(i,j)
        for (int c = 0; c < 3; c++) m += fp.epst[i + 3 * c] * std::conj(fp.epst[j + 3 * c]);
        g.tR[3 * i + j] = m.real(); g.tS[3 * i + j] = m.imag();
        g.ter[3 * i + j] = fp.epst[i + 3 * j].real(); g.tei[3 * i + j] = fp.epst[i + 3 * j].imag();
      }
    return g;
  }
  // descriptor arrays keep the brick layout (19/12/6 ints per element); a prism uses the first 15/9/5 entries
  static std::string key(int etype, const int *norder, const int *norie, const int *norif) {
    const bool pr = etype == 3;
    std::string k(1, (char)etype);
    k.append((const char *)norder, (pr ? 15 : 19) * sizeof(int));
    k.append((const char *)norie, (pr ? 9 : 12) * sizeof(int));
    k.append((const char *)norif, (pr ? 5 : 6) * sizeof(int));
    return k;
  }
  // compile the not yet cached signatures of a call CONCURRENTLY on the host's cores: on an hp mesh nearly every element has
  // its own signature (orientation combinations) and one compilation costs 1-100 ms (trace pairings by host quadrature)
  int compile_missing(const std::vector<std::pair<std::string, int>> &missing, const int *etype, const int *norder, const int *norie,
                      const int *norif, std::string &err) {
    if (missing.empty()) return 0;
    std::vector<std::unique_ptr<Signature>> built(missing.size());
    std::atomic<size_t> next{0};
    auto work = [&]() {
      for (;;) {
        const size_t i = next.fetch_add(1);
        if (i >= missing.size()) break;
        const int e0 = missing[i].second, et = etype ? etype[e0] : 1;
        std::unique_ptr<Signature> s(new Signature());
        if (et == 3) compile_signature_prism(fp, norder + 19 * e0, norie + 12 * e0, norif + 6 * e0, s->h);
        else compile_signature(fp, norder + 19 * e0, norie + 12 * e0, norif + 6 * e0, s->h);
        built[i] = std::move(s);
      }
    };
    unsigned nt = std::thread::hardware_concurrency();
    nt = std::max(1u, std::min(nt ? nt : 4u, std::min((unsigned)missing.size(), 32u)));
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (std::thread &t : th) t.join();
    for (size_t i = 0; i < missing.size(); i++)
      if (!built[i]->h.err.empty()) { err = "element " + std::to_string(missing[i].second) + ": " + built[i]->h.err; return -1; }
    for (size_t i = 0; i < missing.size(); i++) insert(missing[i].first, std::move(built[i]));
    return 0;
  }
  // dof counts / quadrature size / padded extents of a signature without compiling it (cached signatures are reused)
  bool sizes(int etype, const int *norder, const int *norie, const int *norif, SigHost &out, std::string &err) {
    if (etype != 1 && etype != 3) { err = "unknown element type (HP3D_MDLB = 1 and HP3D_MDLP = 3 are implemented)"; return false; }
    if (const Signature *sg = find(key(etype, norder, norie, norif))) { const SigHost &h = sg->h; out = SigHost(); out.ni = h.ni; out.nb = h.nb; out.nint = h.nint; out.nH = h.nH; out.ntest = h.ntest; out.dims = h.dims; return true; }
    const bool ok = etype == 3 ? compile_signature_prism(fp, norder, norie, norif, out, true) : compile_signature(fp, norder, norie, norif, out, true);
    if (!ok) err = out.err;
    return ok;
  }
  // find or compile; device upload only when `device` is set
  Signature *get(int etype, const int *norder, const int *norie, const int *norif, bool device, std::string &err) {
    if (etype != 1 && etype != 3) { err = "unknown element type (HP3D_MDLB = 1 and HP3D_MDLP = 3 are implemented)"; return nullptr; }
    const std::string k = key(etype, norder, norie, norif);
    Signature *s = find(k);
    if (!s) {
      std::unique_ptr<Signature> ns(new Signature());
      const bool ok = etype == 3 ? compile_signature_prism(fp, norder, norie, norif, ns->h) : compile_signature(fp, norder, norie, norif, ns->h);
      if (!ok) { err = ns->h.err; return nullptr; }
      s = insert(k, std::move(ns));
    }
    if (device && !s->d_tab) { if (s->upload(err)) return nullptr; }   // device uploads only happen under the engine lock
    return s;
  }
};

// ------------------------------------------------------------------------------------------------
// Test hook: dense phase only, host matrices in / host matrices out (identity dof maps).
template <bool CPLX>
int dense_debug_run(int nel, int n, int nb, int ni, const void *Gv, const void *Bv, void *Aii, void *Bi, void *AS, void *BS,
                    int *info, std::string &err) {
  typedef typename std::conditional<CPLX, std::complex<double>, double>::type T;
  const T *G = (const T *)Gv, *Bm = (const T *)Bv;
  DenseDims d;
  d.cplx = CPLX; d.dpg = true; d.n = n; d.nb = nb; d.ni = ni; d.finish();
  DenseWorkspace ws;
  if (int rc = ws.reserve(d, nel, err)) return rc;
  const size_t P = d.planes(), wpl = d.w_plane(), m1 = (size_t)nb + ni + 1;
  std::vector<double> hW(P * wpl * nel, 0.0);
  auto re = [](const T &z) { return std::real(z); };
  auto im = [](const T &z) { return std::imag(z); };
  for (int e = 0; e < nel; e++) {
    double *Wr = hW.data() + (size_t)e * P * wpl, *Wi = Wr + wpl;
    const T *Ge = G + (size_t)e * n * n, *Be = Bm + (size_t)e * n * m1;
    for (int r = 0; r < d.np; r++)
      for (int c = 0; c <= r; c++) {
        if (r < n) { T v = Ge[(size_t)c + (size_t)n * r]; /* upper entry (c,r); lower (r,c) = conj */
          Wr[(size_t)r * d.np + c] = re(v); if (CPLX) Wi[(size_t)r * d.np + c] = -im(v); }
        else if (r == c) Wr[(size_t)r * d.np + c] = 1.0;
      }
    for (size_t c = 0; c < m1; c++) {
      size_t row = d.np + (c < (size_t)nb ? c : (c < (size_t)nb + ni ? d.nbp + (c - nb) : (size_t)d.nbp + d.nil - 1));
      for (int k = 0; k < n; k++) {
        T v = Be[(size_t)k + (size_t)n * c];
        Wr[row * d.np + k] = re(v); if (CPLX) Wi[row * d.np + k] = -im(v);
      }
    }
  }
  HP3D_CK(cudaMemcpy(ws.b.W, hW.data(), sizeof(double) * hW.size(), cudaMemcpyHostToDevice));
  HP3D_CK(cudaMemset(ws.b.info, 0, sizeof(int) * nel));
  HP3D_CK(cudaMemset(ws.b.Am, 0, sizeof(double) * P * d.a_plane() * nel));
  std::vector<int> cnt_i(nel, ni), cnt_b(nel, nb);
  int *d_cnt = nullptr;
  HP3D_CK(cudaMalloc((void **)&d_cnt, sizeof(int) * 2 * nel));
  HP3D_CK(cudaMemcpy(d_cnt, cnt_i.data(), sizeof(int) * nel, cudaMemcpyHostToDevice));
  HP3D_CK(cudaMemcpy(d_cnt + nel, cnt_b.data(), sizeof(int) * nel, cudaMemcpyHostToDevice));
  ws.b.ni_e = d_cnt; ws.b.nb_e = d_cnt + nel;
  cudaStream_t st = 0;
  dense_phase<CPLX>(d, ws.b, nel, st);
  HP3D_CK(cudaGetLastError());
  // identity maps
  std::vector<int> pi(ni), pb(nb ? nb : 1);
  std::vector<double> si(ni, 1.0), sb(nb ? nb : 1, 1.0);
  for (int i = 0; i < ni; i++) pi[i] = i;
  for (int i = 0; i < nb; i++) pb[i] = i;
  int *dpi, *dpb; double *dsi, *dsb;
  HP3D_CK(cudaMalloc(&dpi, sizeof(int) * pi.size())); HP3D_CK(cudaMalloc(&dpb, sizeof(int) * pb.size()));
  HP3D_CK(cudaMalloc(&dsi, sizeof(double) * si.size())); HP3D_CK(cudaMalloc(&dsb, sizeof(double) * sb.size()));
  cudaMemcpy(dpi, pi.data(), sizeof(int) * pi.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dpb, pb.data(), sizeof(int) * pb.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dsi, si.data(), sizeof(double) * si.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dsb, sb.data(), sizeof(double) * sb.size(), cudaMemcpyHostToDevice);
  OutMaps mp{dpi, dpb, dsi, dsb, 0, 0, 0, 0};
  constexpr int NS = CPLX ? 2 : 1;
  double *dA, *dB, *dAS, *dBS;
  const size_t nA = (size_t)ni * ni, nAS = (size_t)(nb ? nb : 1) * ni;
  HP3D_CK(cudaMalloc(&dA, sizeof(double) * NS * nA * nel)); HP3D_CK(cudaMalloc(&dB, sizeof(double) * NS * ni * nel));
  HP3D_CK(cudaMalloc(&dAS, sizeof(double) * NS * nAS * nel)); HP3D_CK(cudaMalloc(&dBS, sizeof(double) * NS * (nb ? nb : 1) * nel));
  dim3 blk(16, 16), g1((ni + 15) / 16, (ni + 15) / 16, nel);
  scatter_condensed_kernel<CPLX><<<g1, blk, 0, st>>>(d, ws.b.Am, mp, dA, dB, (long long)nA, (long long)ni, 0);
  if (nb > 0) {
    dim3 g2((nb + 15) / 16, (ni + 15) / 16, nel);
    scatter_schur_kernel<CPLX><<<g2, blk, 0, st>>>(d, ws.b.Am, mp, dAS, dBS, (long long)nAS, (long long)nb);
  }
  HP3D_CK(cudaDeviceSynchronize());
  HP3D_CK(cudaMemcpy(Aii, dA, sizeof(double) * NS * nA * nel, cudaMemcpyDeviceToHost));
  HP3D_CK(cudaMemcpy(Bi, dB, sizeof(double) * NS * ni * nel, cudaMemcpyDeviceToHost));
  if (nb > 0) {
    HP3D_CK(cudaMemcpy(AS, dAS, sizeof(double) * NS * nAS * nel, cudaMemcpyDeviceToHost));
    HP3D_CK(cudaMemcpy(BS, dBS, sizeof(double) * NS * nb * nel, cudaMemcpyDeviceToHost));
  }
  if (info) HP3D_CK(cudaMemcpy(info, ws.b.info, sizeof(int) * nel, cudaMemcpyDeviceToHost));
  cudaFree(dpi); cudaFree(dpb); cudaFree(dsi); cudaFree(dsb); cudaFree(dA); cudaFree(dB); cudaFree(dAS); cudaFree(dBS); cudaFree(d_cnt);
  ws.release();
  return 0;
}

}  // namespace hp3d
