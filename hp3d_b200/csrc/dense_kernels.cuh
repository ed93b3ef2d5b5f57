// dense_kernels.cuh -- batched FP64 dense kernels for the DPG normal equations and the element-local
// static condensation (north-star subsystem 3), hand-written for sm_100a.
//
// Replaces, for a whole batch of elements at once, the LAPACK/BLAS calls of
//   problems/MAXWELL/ULTRAWEAK_DPG/elem/elem_opt.F90:841 (ZPOTRF), :852 (ZTRTRS), :862 (ZHERK)
//   problems/POISSON/PRIMAL_DPG/elem_opt.F90:417 (DPFTRF), :424 (DTFSM), :430 (DSYRK)
//   src/modules/stc.F90:355-414 (?TRTTF/?PFTRF/?PFTRS/?GEMM)
//
// Layout: every matrix is ROW-major with the contraction index contiguous, complex matrices are PLANAR
// (a real plane and an imaginary plane `im_off` doubles apart) so that each complex product is four real
// DMMA (mma.sync.m8n8k4.f64) tile products.  All extents are padded to multiples of TILE=64.
// The only product form needed anywhere on the path is   C(i,j) = Cin(i,j) + alpha * sum_k A(i,k) conj(B(j,k))
// (Cholesky panel update, triangular solve by the inverted diagonal block, HERK, Schur update), so there
// is ONE tensor-core kernel (gemm_nc) plus a small in-shared-memory factor/invert kernel for 64x64 tiles.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hp3d {

constexpr int TILE = 64;      // block size of the factorizations == CTA tile edge
constexpr int KC = 16;        // k-chunk per pipeline stage
constexpr int LDS_K = KC + 4; // padded smem row (20 doubles: rows 160 B apart -> conflict-free 8-byte fragment loads)
constexpr int GEMM_THREADS = 128;

struct MatRef {
  double *re;          // real plane, already offset to the (0,0) entry of the sub-block addressed by this launch
  long long im_off;    // imaginary plane = re + im_off (ignored for real problems)
  long long batch;     // stride between consecutive elements of the batch (doubles)
  int ld;              // row stride (doubles)
};

struct GemmArgs {
  MatRef A, B, Cin, Cout;
  int K;          // contraction length (multiple of KC)
  int lower_only; // skip tiles strictly above the block diagonal (blockIdx.y > blockIdx.x + diag_shift)
  int diag_shift;
  int use_cin;    // 0: C = alpha*S ; 1: C = Cin + alpha*S
  double alpha;
  // work that is known to be zero or unused is dropped per CTA (see gemm_nc_kernel):
  int rows_valid, cols_valid;  // rows / columns of this launch (from its origin) that carry data, multiples of 32; 0: all.
                               // The 32 rows / columns beyond them in the last tile are padding: not computed, written as zeros
                               // when use_cin == 0
  int diag_first;  // the first row tile of the launch is a diagonal tile of a Hermitian matrix whose upper triangle is never
                   // read (Cholesky panel update): only the 8x8 tiles on or below its diagonal are computed and written.  The
                   // same holds for the diagonal tiles of a lower_only launch (blockIdx.y == blockIdx.x + diag_shift)
  int tri;         // K == TILE products with a triangular B: 1 lower (B(j,k) = 0 for k > j: quadrants wn == 0 stop at k = 32),
                   // 2 upper (B(j,k) = 0 for k < j: quadrants wn == 1 start at k = 32)
};

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double dneg(double x) { return __longlong_as_double(__double_as_longlong(x) ^ 0x8000000000000000ULL); }
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// pipeline depth (a third stage for the real kernel was measured: 94.8 vs 93.7 ms per 256 elements, no gain -- the DMMA pipe is
// already 88 % active and the rest is not load latency)
template <bool CPLX> __host__ __device__ constexpr int gemm_stages() { return 2; }
// k-chunk per pipeline stage (32 for the real kernel was measured: 3 CTAs/SM instead of 4, dense 95.3 vs 91.4 ms: worse)
template <bool CPLX> __host__ __device__ constexpr int gemm_kc() { return 16; }

// One 64x64 output tile per CTA, 4 warps.  Work that is known to be zero or never read is dropped per CTA (dropping it per
// warp gains nothing: the warps of a CTA meet at a barrier every k-chunk, so the CTA is as slow as its busiest warp):
//   FULL      2x2 warps, each a 32x32 quadrant = 4x4 DMMA tiles
//   half rows / half columns (last tile of a padded-to-32 extent): the valid 32 rows (columns) are split over the warps,
//             16 per warp: 2x4 (4x2, 2x2) DMMA tiles per warp, half (a quarter of) the work
//   DIAG      diagonal tile of a Hermitian result of which only the lower triangle is read (Cholesky panel update, HERK,
//             Schur update): the 36 of 64 DMMA tiles on or below the diagonal; warp w owns the 8-row strips w and 7-w
//             (w+1 and 8-w tiles: 9 per warp)
// grid = (row tiles, col tiles, batch)
template <bool CPLX>
__global__ void __launch_bounds__(GEMM_THREADS, CPLX ? 2 : 4) gemm_nc_kernel(const GemmArgs g) {
  const int ti = blockIdx.x, tj = blockIdx.y, e = blockIdx.z;
  if (g.lower_only && tj > ti + g.diag_shift) return;
  constexpr int NP = CPLX ? 2 : 1;
  constexpr int KC = gemm_kc<CPLX>(), LDS_K = KC + 4;   // shadow the file-level defaults
  constexpr int NSTAGE = gemm_stages<CPLX>();
  extern __shared__ __align__(16) double smem[];
  // smem: [stage NSTAGE][operand 2][plane NP][TILE][LDS_K]
  auto sm = [&](int stage, int op, int pl) { return smem + (size_t)((stage * 2 + op) * NP + pl) * TILE * LDS_K; };

  const double *Ag = g.A.re + (long long)e * g.A.batch + (long long)ti * TILE * g.A.ld;
  const double *Bg = g.B.re + (long long)e * g.B.batch + (long long)tj * TILE * g.B.ld;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 1, wn = warp & 1, gq = lane >> 2, tq = lane & 3;
  // CTA-uniform shape of the work
  const bool diag = (g.lower_only && tj == ti + g.diag_shift) || (g.diag_first && ti == 0);
  const bool half_r = !diag && g.rows_valid && ti * TILE + 32 >= g.rows_valid;
  const bool half_c = !diag && g.cols_valid && tj * TILE + 32 >= g.cols_valid;
  const int rbase = half_r ? 16 * wm : 32 * wm, cbase = half_c ? 16 * wn : 32 * wn;   // first row / column of the warp (not DIAG)
  const int mcnt = half_r ? 2 : 4, ncnt = half_c ? 2 : 4;
  const int sA_ = warp, sB_ = 7 - warp;   // DIAG: the warp's two 8-row strips

  double cr[4][4][2], ci[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) { cr[a][b][0] = cr[a][b][1] = 0.0; ci[a][b][0] = ci[a][b][1] = 0.0; }

  auto load_stage = [&](int stage, int k0) {
    // each plane of each operand: 64 rows x KC doubles = 64 x KC/2 chunks of 16 B
#pragma unroll
    for (int it = 0; it < (TILE * (KC / 2)) / GEMM_THREADS; it++) {
      int c = tid + it * GEMM_THREADS;
      int row = c / (KC / 2), ch = c % (KC / 2);
#pragma unroll
      for (int pl = 0; pl < NP; pl++) {
        cp_async16(sm(stage, 0, pl) + row * LDS_K + ch * 2, Ag + (pl ? g.A.im_off : 0) + (long long)row * g.A.ld + k0 + ch * 2);
        cp_async16(sm(stage, 1, pl) + row * LDS_K + ch * 2, Bg + (pl ? g.B.im_off : 0) + (long long)row * g.B.ld + k0 + ch * 2);
      }
    }
    cp_async_commit();
  };
  // one complex (or real) 8x8x4 tile product: c += a * conj(b)
  auto cmma = [&](double (&xr)[2], double (&xi)[2], double ar, double ai, double br, double bi) {
    dmma884(xr[0], xr[1], ar, br);
    if (CPLX) {
      dmma884(xr[0], xr[1], ai, bi);          // + Ai*Bi   (conj(B) flips the sign of Bi)
      dmma884(xi[0], xi[1], ai, br);          // + Ai*Br
      dmma884(xi[0], xi[1], dneg(ar), bi);    // - Ar*Bi
    }
  };

  // NSTAGE-deep cp.async pipeline, one barrier per k-chunk: at the top of iteration kc the group of chunk kc has landed
  // (NSTAGE-2 younger groups may still be in flight) and every warp has finished chunk kc-1, whose buffer the next load reuses.
  const int nk = g.K / KC;
#pragma unroll
  for (int s0 = 0; s0 < NSTAGE - 1; s0++) { if (s0 < nk) load_stage(s0, s0 * KC); else cp_async_commit(); }
  for (int kc = 0; kc < nk; kc++) {
    const int st = kc % NSTAGE;
    cp_async_wait<NSTAGE - 2>();
    __syncthreads();
    if (kc + NSTAGE - 1 < nk) load_stage((kc + NSTAGE - 1) % NSTAGE, (kc + NSTAGE - 1) * KC); else cp_async_commit();
    if ((g.tri == 1 && wn == 0 && kc * KC >= 32) || (g.tri == 2 && wn == 1 && kc * KC < 32)) continue;   // triangular B: zero half
    const double *As0 = sm(st, 0, 0) + gq * LDS_K + tq, *Bs0 = sm(st, 1, 0) + gq * LDS_K + tq;
    const double *As1 = CPLX ? sm(st, 0, 1) + gq * LDS_K + tq : nullptr, *Bs1 = CPLX ? sm(st, 1, 1) + gq * LDS_K + tq : nullptr;
    if (diag) {
#pragma unroll
      for (int k4 = 0; k4 < KC / 4; k4++) {
        const double a0r = As0[8 * sA_ * LDS_K + k4 * 4], a1r = As0[8 * sB_ * LDS_K + k4 * 4];
        const double a0i = CPLX ? As1[8 * sA_ * LDS_K + k4 * 4] : 0.0, a1i = CPLX ? As1[8 * sB_ * LDS_K + k4 * 4] : 0.0;
#pragma unroll
        for (int n = 0; n < 8; n++) {
          if (n <= sB_) {
            const double br = Bs0[8 * n * LDS_K + k4 * 4], bi = CPLX ? Bs1[8 * n * LDS_K + k4 * 4] : 0.0;
            if (n < 4) { if (n <= sA_) cmma(cr[0][n], ci[0][n], a0r, a0i, br, bi); }
            if (n < 4) cmma(cr[1][n], ci[1][n], a1r, a1i, br, bi);
            else cmma(cr[2][n - 4], ci[2][n - 4], a1r, a1i, br, bi);
          }
        }
      }
    } else if (mcnt == 4 && ncnt == 4) {
      const double *Ar = As0 + rbase * LDS_K, *Br = Bs0 + cbase * LDS_K;
      const double *Ai = CPLX ? As1 + rbase * LDS_K : nullptr, *Bi = CPLX ? Bs1 + cbase * LDS_K : nullptr;
#pragma unroll
      for (int k4 = 0; k4 < KC / 4; k4++) {
        double ar[4], ai[4], br[4], bi[4];
#pragma unroll
        for (int m = 0; m < 4; m++) {
          ar[m] = Ar[m * 8 * LDS_K + k4 * 4];
          br[m] = Br[m * 8 * LDS_K + k4 * 4];
          if (CPLX) { ai[m] = Ai[m * 8 * LDS_K + k4 * 4]; bi[m] = Bi[m * 8 * LDS_K + k4 * 4]; } else { ai[m] = 0.0; bi[m] = 0.0; }
        }
#pragma unroll
        for (int m = 0; m < 4; m++)
#pragma unroll
          for (int n = 0; n < 4; n++) cmma(cr[m][n], ci[m][n], ar[m], ai[m], br[n], bi[n]);
      }
    } else {   // half tiles: 2 (of 4) DMMA tiles per warp along the halved direction(s)
      const double *Ar = As0 + rbase * LDS_K, *Br = Bs0 + cbase * LDS_K;
      const double *Ai = CPLX ? As1 + rbase * LDS_K : nullptr, *Bi = CPLX ? Bs1 + cbase * LDS_K : nullptr;
#pragma unroll
      for (int k4 = 0; k4 < KC / 4; k4++) {
        double ar[4], ai[4], br[4], bi[4];
#pragma unroll
        for (int m = 0; m < 4; m++) {
          ar[m] = m < mcnt ? Ar[m * 8 * LDS_K + k4 * 4] : 0.0;
          br[m] = m < ncnt ? Br[m * 8 * LDS_K + k4 * 4] : 0.0;
          ai[m] = (CPLX && m < mcnt) ? Ai[m * 8 * LDS_K + k4 * 4] : 0.0;
          bi[m] = (CPLX && m < ncnt) ? Bi[m * 8 * LDS_K + k4 * 4] : 0.0;
        }
#pragma unroll
        for (int m = 0; m < 4; m++)
#pragma unroll
          for (int n = 0; n < 4; n++)
            if (m < mcnt && n < ncnt) cmma(cr[m][n], ci[m][n], ar[m], ai[m], br[n], bi[n]);
      }
    }
  }

  // epilogue: C = [Cin] + alpha*S ; a thread owns entries (row r, cols c, c+1) of each of its DMMA tiles
  const long long crow0 = (long long)ti * TILE, ccol0 = (long long)tj * TILE;
  double *Co = g.Cout.re + (long long)e * g.Cout.batch;
  const double *Cn = g.use_cin ? g.Cin.re + (long long)e * g.Cin.batch : nullptr;
  auto store = [&](int rt, int ct, const double (&xr)[2], const double (&xi)[2]) {   // rt, ct: row / column of the 8x8 tile inside the CTA tile
    const long long r = crow0 + rt + gq, c = ccol0 + ct + 2 * tq;
    double2 vr = make_double2(g.alpha * xr[0], g.alpha * xr[1]);
    double2 vi = make_double2(g.alpha * xi[0], g.alpha * xi[1]);
    if (g.use_cin) {
      double2 o = *reinterpret_cast<const double2 *>(Cn + r * g.Cin.ld + c);
      vr.x += o.x; vr.y += o.y;
      if (CPLX) { double2 oi = *reinterpret_cast<const double2 *>(Cn + g.Cin.im_off + r * g.Cin.ld + c); vi.x += oi.x; vi.y += oi.y; }
    }
    *reinterpret_cast<double2 *>(Co + r * g.Cout.ld + c) = vr;
    if (CPLX) *reinterpret_cast<double2 *>(Co + g.Cout.im_off + r * g.Cout.ld + c) = vi;
  };
  if (diag) {
#pragma unroll
    for (int n = 0; n < 8; n++) {
      if (n <= sB_) {
        if (n < 4) { if (n <= sA_) store(8 * sA_, 8 * n, cr[0][n], ci[0][n]); }
        if (n < 4) store(8 * sB_, 8 * n, cr[1][n], ci[1][n]);
        else store(8 * sB_, 8 * n, cr[2][n - 4], ci[2][n - 4]);
      }
    }
    return;
  }
#pragma unroll
  for (int m = 0; m < 4; m++)
#pragma unroll
    for (int n = 0; n < 4; n++)
      if (m < mcnt && n < ncnt) store(rbase + 8 * m, cbase + 8 * n, cr[m][n], ci[m][n]);
  // the padding half of a half tile is not computed: it stays as it is (zero) when the result is accumulated in place,
  // and is written as zeros when the launch creates the matrix
  if (!g.use_cin && (half_r || half_c)) {
    const double2 z = make_double2(0.0, 0.0);
    for (int idx = tid; idx < TILE * TILE / 2; idx += GEMM_THREADS) {
      const int r = idx / (TILE / 2), c = (idx % (TILE / 2)) * 2;
      if ((half_r && r >= 32) || (half_c && c >= 32)) {
        *reinterpret_cast<double2 *>(Co + (crow0 + r) * g.Cout.ld + ccol0 + c) = z;
        if (CPLX) *reinterpret_cast<double2 *>(Co + g.Cout.im_off + (crow0 + r) * g.Cout.ld + ccol0 + c) = z;
      }
    }
  }
}

// pivots per block and dynamic shared memory of stc_gen_kernel for padded extent M
inline int stc_gen_block(int M, bool cplx) { const size_t row = (size_t)(cplx ? 2 : 1) * M * sizeof(double); int r = (int)((size_t)(160 * 1024) / row); return r > 8 ? 8 : (r < 1 ? 1 : r); }
inline size_t stc_gen_smem(int M, bool cplx) { const size_t a = (size_t)(cplx ? 2 : 1) * stc_gen_block(M, cplx) * M * sizeof(double), b = (size_t)2 * M * sizeof(double); return a > b ? a : b; }

constexpr size_t potrf_smem_bytes() { return (size_t)4 * TILE * (TILE + 1) * sizeof(double); }
template <bool CPLX> constexpr size_t gemm_smem_bytes() { return (size_t)gemm_stages<CPLX>() * 2 * (CPLX ? 2 : 1) * TILE * (gemm_kc<CPLX>() + 4) * sizeof(double); }

// ---------------------------------------------------------------------------------------------------
// potrf_inv_tile_real: the REAL 64x64 diagonal-tile step (factor L L^T, write L back with the upper part zeroed, write
// Linv = L^-1 and LinvH = L^-T), one CTA of 256 threads per element, register-blocked: thread (tx,ty) of a 16x16 grid owns
// the 4x4 entries (ty+16i, tx+16j) (cyclic, so the shrinking trailing matrix stays balanced).  Each of the 64 elimination
// steps is one rank-1 update from a column (factorization) / row (inversion) broadcast through a double-buffered shared
// vector: ONE barrier per step and no integer division -- ~15 us per launch instead of ~100 us for the generic kernel below,
// and 35 KB of shared memory instead of 133 KB, so it co-resides with the GEMM CTAs of the other lanes.
__global__ void __launch_bounds__(256) potrf_inv_tile_real_kernel(MatRef D, MatRef Linv, MatRef LinvH, int *info, int info_stride, int col0) {
  constexpr int N = TILE, LD = TILE + 1;
  __shared__ double Ls[N * LD];
  __shared__ double vec[2][N];
  __shared__ double rs[N];      // 1 / sqrt(pivot)
  int bad = 0;   // first non-positive pivot (1-based); every thread sees the same pivots, so a register per thread will do
  const int e = blockIdx.x, tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  double *Dg = D.re + (long long)e * D.batch;
  double a[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) a[i][j] = Dg[(long long)(ty + 16 * i) * D.ld + tx + 16 * j];
  // ---- right-looking Cholesky, columns scaled at the end: A(r,c) -= A(r,k) A(c,k) / d_k  for r >= c > k
#pragma unroll 1
  for (int k4 = 0; k4 < 4; k4++) {
#pragma unroll 1
    for (int kk = 0; kk < 16; kk++) {
      const int k = 16 * k4 + kk, buf = k & 1;
      if (tx == kk) {
#pragma unroll
        for (int i = 0; i < 4; i++) {   // a[i][k4] with a compile-time second index
          double v = k4 == 0 ? a[i][0] : (k4 == 1 ? a[i][1] : (k4 == 2 ? a[i][2] : a[i][3]));
          vec[buf][ty + 16 * i] = v;
        }
      }
      __syncthreads();
      double d = vec[buf][k];
      if (!(d > 0.0)) { if (bad == 0) bad = k + 1; d = 1.0; }
      const double dinv = 1.0 / d;
      if (tid == 0) rs[k] = 1.0 / sqrt(d);
      double cr[4], cc[4];
#pragma unroll
      for (int i = 0; i < 4; i++) { cr[i] = vec[buf][ty + 16 * i]; cc[i] = vec[buf][tx + 16 * i] * dinv; }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int r = ty + 16 * i, c = tx + 16 * j;
          if (c > k && r >= c) a[i][j] -= cr[i] * cc[j];
        }
    }
  }
  __syncthreads();
  // ---- L = A * diag(rs) (lower), to shared memory and back to the matrix
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = ty + 16 * i, c = tx + 16 * j;
      const double l = r >= c ? a[i][j] * rs[c] : 0.0;
      Ls[r * LD + c] = l;
      Dg[(long long)r * D.ld + c] = l;
    }
  // ---- X = L^-1: residual R = I; for k: X(k,:) = R(k,:) / L(k,k); R(r,:) -= L(r,k) X(k,:) for r > k
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) a[i][j] = (ty + 16 * i == tx + 16 * j) ? 1.0 : 0.0;
  __syncthreads();
#pragma unroll 1
  for (int k4 = 0; k4 < 4; k4++) {
#pragma unroll 1
    for (int kk = 0; kk < 16; kk++) {
      const int k = 16 * k4 + kk, buf = k & 1;
      if (ty == kk) {
        const double s = rs[k];   // 1 / L(k,k)
#pragma unroll
        for (int j = 0; j < 4; j++) {
          double v = (k4 == 0 ? a[0][j] : (k4 == 1 ? a[1][j] : (k4 == 2 ? a[2][j] : a[3][j]))) * s;
          vec[buf][tx + 16 * j] = v;
          if (k4 == 0) a[0][j] = v; else if (k4 == 1) a[1][j] = v; else if (k4 == 2) a[2][j] = v; else a[3][j] = v;
        }
      }
      __syncthreads();
      double xr[4];
#pragma unroll
      for (int j = 0; j < 4; j++) xr[j] = vec[buf][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int r = ty + 16 * i;
        if (r > k) {
          const double l = Ls[r * LD + k];
#pragma unroll
          for (int j = 0; j < 4; j++) a[i][j] -= l * xr[j];
        }
      }
    }
  }
  __syncthreads();   // all reads of Ls are done: reuse it for the transpose
  double *Xg = Linv.re + (long long)e * Linv.batch, *XHg = LinvH.re + (long long)e * LinvH.batch;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = ty + 16 * i, c = tx + 16 * j;
      const double x = r >= c ? a[i][j] : 0.0;
      Xg[(long long)r * Linv.ld + c] = x;
      Ls[c * LD + r] = x;
    }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = ty + 16 * i, c = tx + 16 * j;
      XHg[(long long)r * LinvH.ld + c] = Ls[r * LD + c];
    }
  if (tid == 0 && bad && info && info[(long long)e * info_stride] == 0) info[(long long)e * info_stride] = col0 + bad;
}

// ---------------------------------------------------------------------------------------------------
// potrf_inv_tile: factor the Hermitian positive definite 64x64 diagonal tile (lower, L L^H), write L back
// (upper part zeroed), and write Linv = L^-1 and LinvH = L^-H (both row-major planar 64x64) for the
// "triangular solve = GEMM with the inverted block" steps.  One CTA (256 threads) per element.
// info[e] = first non-positive pivot (1-based, offset by `col0`) like LAPACK ?POTRF, 0 if ok.
template <bool CPLX>
__global__ void __launch_bounds__(256) potrf_inv_tile_kernel(MatRef D, MatRef Linv, MatRef LinvH, int *info, int info_stride, int col0) {
  constexpr int N = TILE, LD = TILE + 1;
  extern __shared__ __align__(16) double smem[];  // 4 planes of N*LD doubles (133 KB: opt-in dynamic smem)
  double *sr = smem, *si = smem + N * LD, *xr = smem + 2 * N * LD, *xi = smem + 3 * N * LD;
  int bad = 0;   // first non-positive pivot (1-based); every thread sees the same pivots, so a register per thread will do
  const int e = blockIdx.x, tid = threadIdx.x;
  double *Dg = D.re + (long long)e * D.batch;
  for (int idx = tid; idx < N * N; idx += blockDim.x) {
    int r = idx / N, c = idx % N;
    sr[r * LD + c] = Dg[(long long)r * D.ld + c];
    if (CPLX) si[r * LD + c] = Dg[D.im_off + (long long)r * D.ld + c];
  }
  __syncthreads();
  // right-looking Cholesky on the lower triangle
  for (int k = 0; k < N; k++) {
    double d = sr[k * LD + k];
    if (!(d > 0.0)) { if (bad == 0) bad = k + 1; d = 1.0; }
    double rs = 1.0 / sqrt(d);
    __syncthreads();
    if (tid >= k && tid < N) { // scale column k (entry (tid,k)); the diagonal becomes sqrt(d)
      if (tid == k) { sr[k * LD + k] = sqrt(d); if (CPLX) si[k * LD + k] = 0.0; }
      else { sr[tid * LD + k] *= rs; if (CPLX) si[tid * LD + k] *= rs; }
    }
    __syncthreads();
    const int rem = N - k - 1;
    for (int idx = tid; idx < rem * rem; idx += blockDim.x) {
      int r = k + 1 + idx / rem, c = k + 1 + idx % rem;
      if (c > r) continue;
      double ar = sr[r * LD + k], br = sr[c * LD + k];
      if (CPLX) {
        double ai = si[r * LD + k], bi = si[c * LD + k];
        sr[r * LD + c] -= ar * br + ai * bi;   // a * conj(b)
        si[r * LD + c] -= ai * br - ar * bi;
      } else sr[r * LD + c] -= ar * br;
    }
    __syncthreads();
  }
  // X = L^-1 by forward substitution, one column per thread
  if (tid < N) {
    const int j = tid;
    for (int i = 0; i < N; i++) {
      if (i < j) { xr[i * LD + j] = 0.0; if (CPLX) xi[i * LD + j] = 0.0; continue; }
      double accr = (i == j) ? 1.0 : 0.0, acci = 0.0;
      for (int k = j; k < i; k++) {
        double lr = sr[i * LD + k], xr_ = xr[k * LD + j];
        if (CPLX) { double li = si[i * LD + k], xi_ = xi[k * LD + j]; accr -= lr * xr_ - li * xi_; acci -= lr * xi_ + li * xr_; }
        else accr -= lr * xr_;
      }
      double dinv = 1.0 / sr[i * LD + i];
      xr[i * LD + j] = accr * dinv;
      if (CPLX) xi[i * LD + j] = acci * dinv;
    }
  }
  __syncthreads();
  double *Xg = Linv.re + (long long)e * Linv.batch, *XHg = LinvH.re + (long long)e * LinvH.batch;
  for (int idx = tid; idx < N * N; idx += blockDim.x) {
    int r = idx / N, c = idx % N;
    bool low = c <= r;
    Dg[(long long)r * D.ld + c] = low ? sr[r * LD + c] : 0.0;
    Xg[(long long)r * Linv.ld + c] = xr[r * LD + c];
    XHg[(long long)r * LinvH.ld + c] = xr[c * LD + r];
    if (CPLX) {
      Dg[D.im_off + (long long)r * D.ld + c] = low ? si[r * LD + c] : 0.0;
      Xg[Linv.im_off + (long long)r * Linv.ld + c] = xi[r * LD + c];
      XHg[LinvH.im_off + (long long)r * LinvH.ld + c] = -xi[c * LD + r];
    }
  }
  if (tid == 0 && bad && info && info[(long long)e * info_stride] == 0) info[(long long)e * info_stride] = col0 + bad;
}

// Out(j,i) = conj(In(i,j)) for an R x C (row-major planar) block; grid = (C/32, R/32, batch), block (32,8)
template <bool CPLX>
__global__ void conj_transpose_kernel(MatRef In, MatRef Out, int R, int Cn) {
  __shared__ double tr[32][33], tim[32][33];
  const int e = blockIdx.z, bx = blockIdx.x * 32, by = blockIdx.y * 32;
  const double *I = In.re + (long long)e * In.batch;
  double *O = Out.re + (long long)e * Out.batch;
  for (int y = threadIdx.y; y < 32; y += 8) {
    int r = by + y, c = bx + threadIdx.x;
    if (r < R && c < Cn) { tr[y][threadIdx.x] = I[(long long)r * In.ld + c]; if (CPLX) tim[y][threadIdx.x] = I[In.im_off + (long long)r * In.ld + c]; }
  }
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += 8) {
    int r = bx + y, c = by + threadIdx.x; // output row = input col
    if (r < Cn && c < R) { O[(long long)r * Out.ld + c] = tr[threadIdx.x][y]; if (CPLX) O[Out.im_off + (long long)r * Out.ld + c] = -tim[threadIdx.x][y]; }
  }
}

// unit diagonal on the padded rows [n0[e], n1) of a square row-major planar matrix (keeps the factorization regular)
__global__ void pad_diag_kernel(double *A, long long batch_stride, int ld, const int *__restrict__ n0, int n1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n0[blockIdx.y] && i < n1) A[(long long)blockIdx.y * batch_stride + (long long)i * ld + i] = 1.0;
}

}  // namespace hp3d

namespace hp3d {

// ---------------------------------------------------------------------------------------------------
// stc_gen_kernel: static condensation with a PIVOTED LU of the bubble block, for element matrices that are not
// Hermitian positive definite (complex-symmetric Maxwell Galerkin: HERM_STC = .false.).  Mirrors
// stc_fwd_gen (src/modules/stc.F90:443-507: ?GETRF, 2 x ?GETRS, 2 x ?GEMM) as one Gaussian elimination of the full
// element matrix [A_bb A_bi b_b ; A_ib A_ii b_i] with partial pivoting restricted to the bubble rows: after nb steps the
// interface rows hold the Schur complement and the condensed load, and a back substitution on the bubble rows gives
// ASchur = A_bb^-1 A_bi, BSchur = A_bb^-1 b_b.  One CTA per element, matrix in global memory (L2-resident), planar;
// the elimination is blocked (rblk pivots per sweep of the trailing matrix, U block rows in shared memory).
// Layout of Am: [M][M] row-major, bubbles at rows/cols [0,nb), interface at [nbp, nbp+ni), load COLUMN M-1; nb, ni per element.
// Outputs are written directly in the caller's layout (column-major, interleaved complex).
// RS2 (with CPLX = false): REAL matrix with a COMPLEX load carried as two real columns (M-2: Re, M-1: Im) -- lossless Maxwell
// Galerkin; the elimination runs in real arithmetic on all columns, the outputs are written as complex numbers.
template <bool CPLX, bool RS2 = false>
__global__ void __launch_bounds__(512) stc_gen_kernel(const int *__restrict__ nb_e, int nbp, const int *__restrict__ ni_e, int M, double *Am,
                                                      long long a_plane, long long a_batch, double *Aii, double *Bi, double *AS, double *BS,
                                                      long long sA, long long sB, long long sAS, long long sBS, int want_schur, int *info, int rblk) {
  static_assert(!(CPLX && RS2), "RS2 is a real elimination");
  constexpr int NS = (CPLX || RS2) ? 2 : 1;   // scalars of the OUTPUT value type
  constexpr int RMAX = 8;                     // pivot columns per block (rblk <= RMAX, chosen by the host from the shared-memory budget)
  extern __shared__ __align__(16) double sh[];   // block rows: [planes][rblk][M]
  __shared__ double red_v[16];
  __shared__ int red_i[16];
  __shared__ int s_piv;
  __shared__ double s_inv[2];

  const int e = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const int nb = nb_e[e], ni = ni_e[e];
  double *Ar = Am + (long long)e * a_batch, *Ai = Ar + a_plane;
  const int ncol = M, lc = RS2 ? M - 2 : M - 1;   // all columns (padding columns are zero); load column(s) = last padded interface column(s)
  double *Ur = sh, *Ui = sh + (size_t)rblk * M;   // U block rows (elimination) / X block rows (back substitution), [t][j]
  int bad = 0;
  // Blocked right-looking elimination: the trailing matrix is swept once per block of rblk pivots instead of once per pivot
  // (the unblocked sweep streams it through L2 for every column and is bandwidth-bound there).  Pivoting: partial, over the
  // bubble rows, on fully updated panel columns -- the pivot sequence of the unblocked algorithm (stc_fwd_gen: ?GETRF on A_bb).
  for (int kb = 0; kb < nb; kb += rblk) {
    const int rb = min(rblk, nb - kb);
    // ---- A. panel: columns kb .. kb+rb-1
    for (int t = 0; t < rb; t++) {
      const int k = kb + t;
      double best = -1.0; int bi = k;
      for (int i = k + tid; i < nb; i += nt) {
        double vr = Ar[(long long)i * M + k], vi = CPLX ? Ai[(long long)i * M + k] : 0.0;
        double v = vr * vr + vi * vi;
        if (v > best) { best = v; bi = i; }
      }
      for (int o = 16; o; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, best, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (lane == 0) { red_v[warp] = best; red_i[warp] = bi; }
      __syncthreads();
      if (tid == 0) {
        double b = red_v[0]; int p = red_i[0];
        for (int w = 1; w < nw; w++) if (red_v[w] > b || (red_v[w] == b && red_i[w] < p)) { b = red_v[w]; p = red_i[w]; }
        s_piv = p;
        if (!(b > 0.0)) { if (!bad) bad = k + 1; }
      }
      __syncthreads();
      const int p = s_piv;
      // swap rows k <-> p: the block's multipliers (columns kb..) and everything to the right
      if (p != k)
        for (int j = kb + tid; j < ncol; j += nt) {
          if (j >= nb && j < nbp) continue;
          double a = Ar[(long long)p * M + j]; Ar[(long long)p * M + j] = Ar[(long long)k * M + j]; Ar[(long long)k * M + j] = a;
          if (CPLX) { double b = Ai[(long long)p * M + j]; Ai[(long long)p * M + j] = Ai[(long long)k * M + j]; Ai[(long long)k * M + j] = b; }
        }
      __syncthreads();
      if (tid == 0) {
        double dr = Ar[(long long)k * M + k], di = CPLX ? Ai[(long long)k * M + k] : 0.0, dn = dr * dr + di * di;
        if (!(dn > 0.0)) { dr = 1.0; di = 0.0; dn = 1.0; }
        s_inv[0] = dr / dn; s_inv[1] = -di / dn;
      }
      __syncthreads();
      const double ir = s_inv[0], ii = s_inv[1];
      // multipliers of column k (stored in place) and the update of the remaining PANEL columns; one thread per row
      const int nrows = (nb - k - 1) + ni;
      for (int r = tid; r < nrows; r += nt) {
        const int i = (r < nb - k - 1) ? k + 1 + r : nbp + (r - (nb - k - 1));
        double *rr = Ar + (long long)i * M, *ri = Ai + (long long)i * M;
        const double ar = rr[k], ai = CPLX ? ri[k] : 0.0;
        const double lr = ar * ir - ai * ii, li = ar * ii + ai * ir;   // l = a(i,k) / pivot
        rr[k] = lr; if (CPLX) ri[k] = li;
        for (int j = k + 1; j < kb + rb; j++) {
          const double ur = Ar[(long long)k * M + j], ui = CPLX ? Ai[(long long)k * M + j] : 0.0;
          if (CPLX) { rr[j] -= lr * ur - li * ui; ri[j] -= lr * ui + li * ur; }
          else rr[j] -= lr * ur;
        }
      }
      __syncthreads();
    }
    // ---- B. U block rows: row kb+t, columns right of the panel, forward-substituted with the unit lower triangle of the panel
    const int j0 = kb + rb;
    for (int t = 0; t < rb; t++) {
      const double *lrow_r = Ar + (long long)(kb + t) * M + kb, *lrow_i = Ai + (long long)(kb + t) * M + kb;
      for (int j = j0 + tid; j < ncol; j += nt) {
        if (j >= nb && j < nbp) { Ur[(size_t)t * M + j] = 0.0; if (CPLX) Ui[(size_t)t * M + j] = 0.0; continue; }
        double vr = Ar[(long long)(kb + t) * M + j], vi = CPLX ? Ai[(long long)(kb + t) * M + j] : 0.0;
        for (int s2 = 0; s2 < t; s2++) {
          const double lr = lrow_r[s2], li = CPLX ? lrow_i[s2] : 0.0, ur = Ur[(size_t)s2 * M + j], ui = CPLX ? Ui[(size_t)s2 * M + j] : 0.0;
          if (CPLX) { vr -= lr * ur - li * ui; vi -= lr * ui + li * ur; }
          else vr -= lr * ur;
        }
        Ur[(size_t)t * M + j] = vr; Ar[(long long)(kb + t) * M + j] = vr;
        if (CPLX) { Ui[(size_t)t * M + j] = vi; Ai[(long long)(kb + t) * M + j] = vi; }
      }
      // column j of row t only depends on column j of the rows above: no barrier needed between the t steps
    }
    __syncthreads();
    // ---- C. trailing update: one sweep, rank-rb
    const int nrows = (nb - j0) + ni;
    for (int r = warp; r < nrows; r += nw) {
      const int i = (r < nb - j0) ? j0 + r : nbp + (r - (nb - j0));
      double *rr = Ar + (long long)i * M, *ri = Ai + (long long)i * M;
      double lr[RMAX], li[RMAX];
#pragma unroll
      for (int t = 0; t < RMAX; t++) { lr[t] = t < rb ? rr[kb + t] : 0.0; li[t] = (CPLX && t < rb) ? ri[kb + t] : 0.0; }
      for (int j = j0 + lane; j < ncol; j += 32) {
        if (j >= nb && j < nbp) continue;
        double vr = rr[j], vi = CPLX ? ri[j] : 0.0;
#pragma unroll
        for (int t = 0; t < RMAX; t++) {
          if (t < rb) {
            const double ur = Ur[(size_t)t * M + j], ui = CPLX ? Ui[(size_t)t * M + j] : 0.0;
            if (CPLX) { vr -= lr[t] * ur - li[t] * ui; vi -= lr[t] * ui + li[t] * ur; }
            else vr -= lr[t] * ur;
          }
        }
        rr[j] = vr; if (CPLX) ri[j] = vi;
      }
    }
    __syncthreads();
  }
  // ---- condensed system straight from the interface rows
  double *oA = Aii + (long long)e * sA * NS, *oB = Bi + (long long)e * sB * NS;
  for (int idx = tid; idx < ni * ni; idx += nt) {
    const int r = idx % ni, c = idx / ni;   // column-major output
    oA[(long long)idx * NS] = Ar[(long long)(nbp + r) * M + nbp + c];
    if (CPLX) oA[(long long)idx * NS + 1] = Ai[(long long)(nbp + r) * M + nbp + c];
    if (RS2) oA[(long long)idx * NS + 1] = 0.0;
  }
  for (int r = tid; r < ni; r += nt) {
    oB[(long long)r * NS] = Ar[(long long)(nbp + r) * M + lc];
    if (CPLX) oB[(long long)r * NS + 1] = Ai[(long long)(nbp + r) * M + lc];
    if (RS2) oB[(long long)r * NS + 1] = Ar[(long long)(nbp + r) * M + lc + 1];
  }
  if (tid == 0 && bad && info[e] == 0) info[e] = bad;
  if (!want_schur || nb == 0) return;
  // ---- back substitution on the bubble rows: X = U_bb^-1 [U_bi | y_b], in place in columns [nbp, ncol); blocked like the
  // elimination: the rblk rows of a block are solved into shared memory, then the rows above are swept once (rank-rb update)
  __syncthreads();
  for (int kt = nb; kt > 0; kt -= rblk) {
    const int kb = max(kt - rblk, 0), rb = kt - kb;
    for (int t = rb - 1; t >= 0; t--) {   // rows of the block, bottom up; thread-owned columns: no barrier between the t steps
      const int k = kb + t;
      double dr = Ar[(long long)k * M + k], di = CPLX ? Ai[(long long)k * M + k] : 0.0, dn = dr * dr + di * di;
      if (!(dn > 0.0)) { dr = 1.0; di = 0.0; dn = 1.0; }
      const double ir = dr / dn, ii = -di / dn;
      for (int j = nbp + tid; j < ncol; j += nt) {
        double vr = Ar[(long long)k * M + j], vi = CPLX ? Ai[(long long)k * M + j] : 0.0;
        for (int s2 = t + 1; s2 < rb; s2++) {
          const double ur = Ar[(long long)k * M + kb + s2], ui = CPLX ? Ai[(long long)k * M + kb + s2] : 0.0;
          const double xr = Ur[(size_t)s2 * M + j], xi = CPLX ? Ui[(size_t)s2 * M + j] : 0.0;
          if (CPLX) { vr -= ur * xr - ui * xi; vi -= ur * xi + ui * xr; }
          else vr -= ur * xr;
        }
        const double xr = vr * ir - vi * ii, xi = vr * ii + vi * ir;
        Ur[(size_t)t * M + j] = xr; Ar[(long long)k * M + j] = xr;
        if (CPLX) { Ui[(size_t)t * M + j] = xi; Ai[(long long)k * M + j] = xi; }
      }
    }
    __syncthreads();
    for (int i = warp; i < kb; i += nw) {   // rows above the block
      double *rr = Ar + (long long)i * M, *ri = Ai + (long long)i * M;
      double ur[RMAX], ui[RMAX];
#pragma unroll
      for (int t = 0; t < RMAX; t++) { ur[t] = t < rb ? rr[kb + t] : 0.0; ui[t] = (CPLX && t < rb) ? ri[kb + t] : 0.0; }
      for (int j = nbp + lane; j < ncol; j += 32) {
        double vr = rr[j], vi = CPLX ? ri[j] : 0.0;
#pragma unroll
        for (int t = 0; t < RMAX; t++) {
          if (t < rb) {
            const double xr = Ur[(size_t)t * M + j], xi = CPLX ? Ui[(size_t)t * M + j] : 0.0;
            if (CPLX) { vr -= ur[t] * xr - ui[t] * xi; vi -= ur[t] * xi + ui[t] * xr; }
            else vr -= ur[t] * xr;
          }
        }
        rr[j] = vr; if (CPLX) ri[j] = vi;
      }
    }
    __syncthreads();
  }
  double *oS = AS + (long long)e * sAS * NS, *oT = BS + (long long)e * sBS * NS;
  for (int idx = tid; idx < nb * ni; idx += nt) {
    const int b = idx % nb, c = idx / nb;
    oS[(long long)idx * NS] = Ar[(long long)b * M + nbp + c];
    if (CPLX) oS[(long long)idx * NS + 1] = Ai[(long long)b * M + nbp + c];
    if (RS2) oS[(long long)idx * NS + 1] = 0.0;
  }
  for (int b = tid; b < nb; b += nt) {
    oT[(long long)b * NS] = Ar[(long long)b * M + lc];
    if (CPLX) oT[(long long)b * NS + 1] = Ai[(long long)b * M + lc];
    if (RS2) oT[(long long)b * NS + 1] = Ar[(long long)b * M + lc + 1];
  }
}

}  // namespace hp3d
