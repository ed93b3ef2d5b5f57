// dense_pipeline.cuh -- host-side orchestration of the batched dense phase for one batch of elements:
//   (DPG only)  G = L L^H  (left-looking blocked Cholesky),  B~^H = B^H L^-H  (rows appended below G, so the
//               triangular solve rides in the same panel update),  A = B~^H (B~^H)^H   (HERK)
//   (all)       static condensation of the bubble block:  A_bb = L L^H,  Y~ = A_ib L^-H,
//               A_ii -= Y~ Y~^H,  ASchur^H = Y~ L^-1   (reference: src/modules/stc.F90:338-414)
// Internal trial ordering is [bubble dofs | interface dofs | load], each padded to a multiple of 64; the
// signed permutation back to the reference dof ordering happens in the output kernel (formats.cuh).
#pragma once
#include "dense_kernels.cuh"
#include <algorithm>
#include <cstdio>

namespace hp3d {

inline int pad64(int n) { return (n + TILE - 1) / TILE * TILE; }
inline int pad32(int n) { return (n + 31) / 32 * 32; }
inline int pad16(int n) { return (n + 15) / 16 * 16; }

struct DenseDims {
  bool cplx = true;
  bool dpg = true;     // true: Gram + enriched stiffness present; false: A is given directly
  // "real-structured" complex problem: the element system is T A~ T^H with A~ REAL and T a diagonal matrix of powers of i
  // (lossless ultraweak Maxwell, see forms.hpp); the dense phase then runs in real arithmetic (cplx = false, one plane) with the
  // complex load carried as TWO real rows (nload = 2), and the output kernels re-apply the phases.
  bool rs = false;
  int nload = 1;       // load rows: the last `nload` padded interface rows
  int n = 0;           // test dofs (rows of the Gram)
  int nb = 0, ni = 0;  // bubble / interface trial dofs
  int np = 0, nbp = 0, nip = 0;  // padded LAYOUT extents: np=pad64(n), nbp=pad64(nb), nip=pad64(nil)
  int nil = 0;                   // interface rows that carry data: the load sits at interface index nil-nload.  pad32(ni+nload) for the
                                 // Cholesky pipeline (the GEMM skips the 32-row quadrants beyond it), pad64 for the pivoted-LU kernel
  // A batch may mix elements with different n / nb / ni as long as the PADDED extents agree ("dense class"): the kernels
  // below only use np/nbp/nip plus the per-element counts in DenseBuffers::ni_e / nb_e; n, nb, ni here are the class maxima.
  __host__ __device__ int M() const { return nbp + nip; }
  __host__ __device__ int R() const { return np + nbp + nip; }
  __host__ __device__ int nrhs() const { return nload / (rs ? 2 : 1); }   // load vectors (NR_RHS); one (complex / real) or two (real form) rows each
  __host__ __device__ int Mv() const { return nbp + nil; }        // rows / columns of A that carry data
  __host__ __device__ int Rv() const { return np + nbp + nil; }   // rows of W that carry data
  void finish() { np = dpg ? pad64(n) : 0; nbp = pad64(nb); nil = dpg ? pad32(ni + nload) : pad64(ni + nload); nip = pad64(nil); }
  // doubles per element
  __host__ __device__ size_t planes() const { return cplx ? 2 : 1; }
  __host__ __device__ size_t w_plane() const { return (size_t)R() * np; }
  __host__ __device__ size_t a_plane() const { return (size_t)M() * M(); }
  __host__ __device__ size_t lh_plane() const { return (size_t)nbp * nbp; }
  __host__ __device__ size_t linv_plane() const { return (size_t)TILE * TILE; }
  __host__ __device__ int nsteps_stc() const { return nbp / TILE; }
};

struct DenseBuffers {
  double *W = nullptr;      // [batch][planes][R][np]     Gram (lower) on top of B^H
  double *Am = nullptr;     // [batch][planes][M][M]      A = B~^H B~ (lower), then L / Y~ / Schur / Z in place
  double *LH = nullptr;     // [batch][planes][nbp][nbp]  L^H of the bubble block (for the backward solve)
  double *Linv = nullptr;   // [batch][planes][64][64]    scratch for the Gram factorization steps
  double *LinvH = nullptr;
  double *LinvS = nullptr;  // [batch][nsteps][planes][64][64]  kept for the stc backward solve
  double *LinvSH = nullptr;
  int *info = nullptr;      // [batch]
  int *ni_e = nullptr, *nb_e = nullptr;  // [batch] per-element interface / bubble dof counts
  int *nip_e = nullptr;                  // [batch] per-element OWN padded interface extent: the element's load rows sit at interface
                                         // index nip_e - nload (classes merged across nip keep each signature's row layout)
};

template <bool CPLX>
static void launch_gemm(const GemmArgs &g, int mt, int nt, int batch, cudaStream_t st) {
  if (mt <= 0 || nt <= 0 || batch <= 0) return;
  dim3 grid(mt, nt, batch);
  gemm_nc_kernel<CPLX><<<grid, GEMM_THREADS, gemm_smem_bytes<CPLX>(), st>>>(g);
}

template <bool CPLX> static cudaError_t dense_configure() {
  cudaError_t e = cudaFuncSetAttribute(gemm_nc_kernel<CPLX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<CPLX>());
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(potrf_inv_tile_kernel<CPLX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)potrf_smem_bytes());
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(stc_gen_kernel<CPLX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess || CPLX) return e;
  return cudaFuncSetAttribute(stc_gen_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

// Left-looking blocked Cholesky of the leading `ncol` columns of a lower-trapezoidal row-major matrix with
// `nrow` rows (rows >= ncol are "appended right-hand sides": they come out as X L^-H).
// If LinvKeep != nullptr the inverted diagonal blocks are stored per step (stride keep_step doubles).
template <bool CPLX>
static void chol_trapezoid(double *Mbase, long long plane, long long batch_stride, int ld, int nrow, int ncol, int batch,
                           double *Linv, double *LinvH, long long linv_batch, long long linv_step, int *info, cudaStream_t st, int nrow_valid) {
  const int nt_r = nrow / TILE, nt_c = ncol / TILE;
  const long long lp = (long long)TILE * TILE;
  for (int j = 0; j < nt_c; j++) {
    MatRef Dj{Mbase + (long long)j * TILE * ld + (long long)j * TILE, plane, batch_stride, ld};
    if (j > 0) {  // panel update: rows j.. , block column j :  W(i,j) -= sum_{k<j} L(i,k) L(j,k)^H
      GemmArgs g{};
      g.A = MatRef{Mbase + (long long)j * TILE * ld, plane, batch_stride, ld};
      g.B = MatRef{Mbase + (long long)j * TILE * ld, plane, batch_stride, ld};
      g.Cin = Dj; g.Cout = Dj;
      g.K = j * TILE; g.lower_only = 0; g.diag_shift = 0; g.use_cin = 1; g.alpha = -1.0;
      g.rows_valid = nrow_valid - j * TILE; g.diag_first = 1;
      launch_gemm<CPLX>(g, nt_r - j, 1, batch, st);
    }
    MatRef Li{Linv + j * linv_step, lp, linv_batch, TILE}, LiH{LinvH + j * linv_step, lp, linv_batch, TILE};
    if (CPLX) potrf_inv_tile_kernel<CPLX><<<batch, 256, potrf_smem_bytes(), st>>>(Dj, Li, LiH, info, 1, j * TILE);
    else potrf_inv_tile_real_kernel<<<batch, 256, 0, st>>>(Dj, Li, LiH, info, 1, j * TILE);
    if (nt_r - j - 1 > 0) {  // rows below the diagonal block:  L(i,j) = P(i,j) * Linv_jj^H
      GemmArgs g{};
      MatRef Pj{Mbase + (long long)(j + 1) * TILE * ld + (long long)j * TILE, plane, batch_stride, ld};
      g.A = Pj; g.B = Li; g.Cin = Pj; g.Cout = Pj;
      g.K = TILE; g.use_cin = 0; g.alpha = 1.0;
      g.rows_valid = nrow_valid - (j + 1) * TILE; g.tri = 1;   // Linv is lower triangular
      launch_gemm<CPLX>(g, nt_r - j - 1, 1, batch, st);
    }
  }
}

// The whole dense phase for `batch` elements whose W (DPG) or Am (Galerkin) buffers have been filled.
// `normal_eq_only`: stop after A = B~^H B~ (the uncondensed DPG system incl. the load row: what the residual needs).
// `want_z`: also form Z = Y~ L^-1 (ASchur^H and the bubble load), 6 % of the flops at config 3.
template <bool CPLX>
static void dense_phase(const DenseDims &d, const DenseBuffers &b, int batch, cudaStream_t st, bool normal_eq_only = false, bool want_z = true) {
  const long long P = CPLX ? 2 : 1;
  const long long lp = (long long)TILE * TILE;
  const int M = d.M();
  if (d.dpg) {
    const long long wpl = (long long)d.w_plane();
    chol_trapezoid<CPLX>(b.W, wpl, P * wpl, d.np, d.R(), d.np, batch, b.Linv, b.LinvH, P * lp, 0, b.info, st, d.Rv());
    // A = B~^H (B~^H)^H  (lower tiles only)
    GemmArgs g{};
    MatRef Bt{b.W + (long long)d.np * d.np, wpl, P * wpl, d.np};
    g.A = Bt; g.B = Bt;
    g.Cout = MatRef{b.Am, (long long)d.a_plane(), P * (long long)d.a_plane(), M}; g.Cin = g.Cout;
    g.K = std::min(d.np, pad16(d.n)); g.lower_only = 1; g.diag_shift = 0; g.use_cin = 0; g.alpha = 1.0;   // columns >= n of B~^H are zero
    g.rows_valid = d.Mv(); g.cols_valid = d.Mv();
    launch_gemm<CPLX>(g, M / TILE, M / TILE, batch, st);
  }
  if (d.nb == 0 || normal_eq_only) return;
  const long long apl = (long long)d.a_plane(), ab = P * apl;
  {  // padded bubble rows [nb_e, nbp): unit diagonal keeps the factorization regular
    dim3 grid(d.nbp / 64, batch);
    pad_diag_kernel<<<grid, 64, 0, st>>>(b.Am, ab, M, b.nb_e, d.nbp);
  }
  const int ns = d.nsteps_stc();
  // A_bb = L L^H ; rows below become Y~ = A_ib L^-H (and the load row y_b^H)
  chol_trapezoid<CPLX>(b.Am, apl, ab, M, M, d.nbp, batch, b.LinvS, b.LinvSH, (long long)ns * P * lp, P * lp, b.info, st, d.Mv());
  {  // Schur complement: S = A_ii - Y~ Y~^H (lower tiles)
    GemmArgs g{};
    MatRef Y{b.Am + (long long)d.nbp * M, apl, ab, M};
    MatRef S{b.Am + (long long)d.nbp * M + d.nbp, apl, ab, M};
    g.A = Y; g.B = Y; g.Cin = S; g.Cout = S;
    g.K = d.nbp; g.lower_only = 1; g.use_cin = 1; g.alpha = -1.0;
    g.rows_valid = d.nil; g.cols_valid = d.nil;
    launch_gemm<CPLX>(g, d.nip / TILE, d.nip / TILE, batch, st);
  }
  if (!want_z) return;   // the Schur factors are not wanted (STORE_STC off): the condensed system is complete
  {  // LH = L^H (upper, row-major) for the backward solve
    MatRef In{b.Am, apl, ab, M}, Out{b.LH, (long long)d.lh_plane(), P * (long long)d.lh_plane(), d.nbp};
    dim3 grid(d.nbp / 32, d.nbp / 32, batch), blk(32, 8);
    conj_transpose_kernel<CPLX><<<grid, blk, 0, st>>>(In, Out, d.nbp, d.nbp);
  }
  // Z = Y~ L^-1  (ASchur^H), block columns from last to first, in place
  for (int j = ns - 1; j >= 0; j--) {
    MatRef Zj{b.Am + (long long)d.nbp * M + (long long)j * TILE, apl, ab, M};
    if (j < ns - 1) {
      GemmArgs g{};
      g.A = MatRef{b.Am + (long long)d.nbp * M + (long long)(j + 1) * TILE, apl, ab, M};
      g.B = MatRef{b.LH + (long long)j * TILE * d.nbp + (long long)(j + 1) * TILE, (long long)d.lh_plane(), P * (long long)d.lh_plane(), d.nbp};
      g.Cin = Zj; g.Cout = Zj;
      g.K = d.nbp - (j + 1) * TILE; g.use_cin = 1; g.alpha = -1.0;
      g.rows_valid = d.nil;
      launch_gemm<CPLX>(g, d.nip / TILE, 1, batch, st);
    }
    GemmArgs g{};
    g.A = Zj; g.B = MatRef{b.LinvSH + (long long)j * P * lp, lp, (long long)ns * P * lp, TILE};
    g.Cin = Zj; g.Cout = Zj; g.K = TILE; g.use_cin = 0; g.alpha = 1.0;
    g.rows_valid = d.nil; g.tri = 2;   // B = Linv^H: upper triangular
    launch_gemm<CPLX>(g, d.nip / TILE, 1, batch, st);
  }
}

}  // namespace hp3d
