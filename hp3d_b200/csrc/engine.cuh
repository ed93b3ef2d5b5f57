// engine.cuh -- batch engine behind the C ABI: workspaces, the per-batch pipeline, debug drivers.
#pragma once
#include "dense_pipeline.cuh"
#include "formats.cuh"

#include <complex>
#include <string>
#include <vector>

namespace hp3d {

#define HP3D_CK(x)                                                                          \
  do {                                                                                      \
    cudaError_t e_ = (x);                                                                   \
    if (e_ != cudaSuccess) { err = std::string(#x) + ": " + cudaGetErrorString(e_); return -2; } \
  } while (0)

struct DenseWorkspace {
  DenseDims d;
  DenseBuffers b;
  int cap = 0;  // elements
  void release() {
    cudaFree(b.W); cudaFree(b.Am); cudaFree(b.LH); cudaFree(b.Linv); cudaFree(b.LinvH); cudaFree(b.LinvS); cudaFree(b.LinvSH); cudaFree(b.info);
    b = DenseBuffers{}; cap = 0;
  }
  int reserve(const DenseDims &dims, int batch, std::string &err) {
    release();
    d = dims;
    const size_t P = d.planes(), lp = d.linv_plane();
    if (d.dpg) HP3D_CK(cudaMalloc(&b.W, sizeof(double) * P * d.w_plane() * batch));
    HP3D_CK(cudaMalloc(&b.Am, sizeof(double) * P * d.a_plane() * batch));
    HP3D_CK(cudaMalloc(&b.LH, sizeof(double) * P * (d.lh_plane() ? d.lh_plane() : 1) * batch));
    HP3D_CK(cudaMalloc(&b.Linv, sizeof(double) * P * lp * batch));
    HP3D_CK(cudaMalloc(&b.LinvH, sizeof(double) * P * lp * batch));
    const size_t ns = d.nsteps_stc() ? d.nsteps_stc() : 1;
    HP3D_CK(cudaMalloc(&b.LinvS, sizeof(double) * P * lp * ns * batch));
    HP3D_CK(cudaMalloc(&b.LinvSH, sizeof(double) * P * lp * ns * batch));
    HP3D_CK(cudaMalloc(&b.info, sizeof(int) * batch));
    cap = batch;
    return 0;
  }
};

struct Plan;  // defined in plan.cuh

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
      size_t row = d.np + (c < (size_t)nb ? c : d.nbp + (c - nb));
      for (int k = 0; k < n; k++) {
        T v = Be[(size_t)k + (size_t)n * c];
        Wr[row * d.np + k] = re(v); if (CPLX) Wi[row * d.np + k] = -im(v);
      }
    }
  }
  HP3D_CK(cudaMemcpy(ws.b.W, hW.data(), sizeof(double) * hW.size(), cudaMemcpyHostToDevice));
  HP3D_CK(cudaMemset(ws.b.info, 0, sizeof(int) * nel));
  HP3D_CK(cudaMemset(ws.b.Am, 0, sizeof(double) * P * d.a_plane() * nel));
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
  scatter_condensed_kernel<CPLX><<<g1, blk, 0, st>>>(d, ws.b.Am, mp, dA, dB, (long long)nA, (long long)ni);
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
  cudaFree(dpi); cudaFree(dpb); cudaFree(dsi); cudaFree(dsb); cudaFree(dA); cudaFree(dB); cudaFree(dAS); cudaFree(dBS);
  ws.release();
  return 0;
}

struct Plan {
  int kind = 0;
  virtual ~Plan() {}
};

}  // namespace hp3d
