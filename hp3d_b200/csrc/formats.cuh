// formats.cuh -- layout conversions at the boundary of the dense phase.
//   * scatter of condensed results from the internal planar / padded / [bubble|interface|load] layout into
//     the caller's layout: column-major, interleaved complex(8), reference dof ordering with orientation
//     signs (ALOC/BLOC after stc_fwd_wrapper and CLOC(iel)%ASchur/BSchur, src/modules/stc.F90:273-305).
#pragma once
#include "dense_pipeline.cuh"

namespace hp3d {

struct OutMaps {
  const int *perm_i;   // [ni]  reference interface dof -> internal interface index (0..ni-1)
  const int *perm_b;   // [nb]  reference bubble dof    -> internal bubble index
  const double *sgn_i; // [ni]  +-1 (orientation sign);  per element when sgn_stride != 0
  const double *sgn_b; // [nb]
  int sgn_stride_i, sgn_stride_b;  // 0: shared by the batch ; else per-element stride
  int perm_stride_i, perm_stride_b;
  const int *ni_e, *nb_e;   // per-element dof counts (nullptr: the DenseDims values)
  const int *nip_e;         // per-element padded interface extent (position of the load rows; nullptr: DenseDims::nip)
};

// grid = (ceil(ni/16), ceil(ni/16)+extras, batch), block (16,16).  Writes Aii (ni x ni), Bi (ni).
// packed: only the lower triangle, LAPACK 'L' packed column-major storage AP[r + c(2 ni - c - 1)/2] = A(r,c), r >= c (0-based):
// what ZHERK('L') + ZTRTTP would hold; the upper triangle is its conjugate mirror (hp3d_gpu_hermitian_unpack_batch).
template <bool CPLX>
__global__ void scatter_condensed_kernel(DenseDims d, const double *Am, OutMaps mp, double *Aii, double *Bi, long long sA, long long sB, int packed) {
  const int e = blockIdx.z;
  const int r = blockIdx.x * 16 + threadIdx.x, c = blockIdx.y * 16 + threadIdx.y;
  const int M = d.M();
  const long long apl = (long long)d.a_plane();
  const double *S = Am + (long long)e * (CPLX ? 2 : 1) * apl + (long long)d.nbp * M + d.nbp;
  const int *pi = mp.perm_i + (long long)e * mp.perm_stride_i;
  const double *si = mp.sgn_i + (long long)e * mp.sgn_stride_i;
  constexpr int NS = CPLX ? 2 : 1;
  const int ni = mp.ni_e ? mp.ni_e[e] : d.ni, lrow0 = (mp.nip_e ? mp.nip_e[e] : d.nil) - d.nload;   // first load row
  if (r < ni && c < ni && (!packed || r >= c)) {
    int ir = pi[r], ic = pi[c];
    double s = si[r] * si[c];
    int a = ir >= ic ? ir : ic, b = ir >= ic ? ic : ir;
    double vr = S[(long long)a * M + b], vi = 0.0;
    if (CPLX) { vi = S[apl + (long long)a * M + b]; if (ir < ic) vi = -vi; if (ir == ic) vi = 0.0; }
    const long long at = packed ? (long long)r + ((long long)c * (2LL * ni - c - 1)) / 2 : (long long)r + (long long)ni * c;
    double *o = Aii + (long long)e * sA * NS + at * NS;
    o[0] = s * vr;
    if (CPLX) o[1] = s * vi;
  }
  if (blockIdx.y == 0 && threadIdx.y == 0 && r < ni) {
    int ir = pi[r];
    for (int q = 0; q < d.nload; q++) {   // Bi(ni, NR_RHS), column q at offset q*ni
      const int lrow = lrow0 + q;
      double *o = Bi + (long long)e * sB * NS + ((long long)q * ni + r) * NS;
      o[0] = si[r] * S[(long long)lrow * M + ir];
      if (CPLX) o[1] = -si[r] * S[apl + (long long)lrow * M + ir];   // b_i = conj(load row)
    }
  }
}

// grid = (ceil(nb/16), ceil(ni/16), batch), block (16,16).  ASchur (nb x ni) = conj(Z)^T, BSchur (nb).
template <bool CPLX>
__global__ void scatter_schur_kernel(DenseDims d, const double *Am, OutMaps mp, double *AS, double *BS, long long sAS, long long sBS) {
  const int e = blockIdx.z;
  const int bq = blockIdx.x * 16 + threadIdx.x, iq = blockIdx.y * 16 + threadIdx.y;
  const int M = d.M();
  const long long apl = (long long)d.a_plane();
  const double *Z = Am + (long long)e * (CPLX ? 2 : 1) * apl + (long long)d.nbp * M;
  const int *pi = mp.perm_i + (long long)e * mp.perm_stride_i, *pb = mp.perm_b + (long long)e * mp.perm_stride_b;
  const double *si = mp.sgn_i + (long long)e * mp.sgn_stride_i, *sb = mp.sgn_b + (long long)e * mp.sgn_stride_b;
  constexpr int NS = CPLX ? 2 : 1;
  const int ni = mp.ni_e ? mp.ni_e[e] : d.ni, nb = mp.nb_e ? mp.nb_e[e] : d.nb, lrow0 = (mp.nip_e ? mp.nip_e[e] : d.nil) - d.nload;
  if (bq < nb && iq < ni) {
    int ib = pb[bq], ii = pi[iq];
    double s = sb[bq] * si[iq];
    double *o = AS + (long long)e * sAS * NS + ((long long)bq + (long long)nb * iq) * NS;
    o[0] = s * Z[(long long)ii * M + ib];
    if (CPLX) o[1] = -s * Z[apl + (long long)ii * M + ib];
  }
  if (blockIdx.y == 0 && threadIdx.y == 0 && bq < nb) {
    int ib = pb[bq];
    for (int q = 0; q < d.nload; q++) {   // BSchur(nb, NR_RHS)
      const int lrow = lrow0 + q;
      double *o = BS + (long long)e * sBS * NS + ((long long)q * nb + bq) * NS;
      o[0] = sb[bq] * Z[(long long)lrow * M + ib];
      if (CPLX) o[1] = -sb[bq] * Z[apl + (long long)lrow * M + ib];
    }
  }
}

// xb = BSchur - ASchur * xi (stc_bwd, src/modules/stc.F90:661-677), one warp per (bubble row, element); ASchur (nb x ni) and
// xi / xb in the caller's layout (column-major, interleaved complex), per-element sizes.
template <bool CPLX>
__global__ void stc_bwd_kernel(const int *__restrict__ ni_e, const int *__restrict__ nb_e, int ni_u, int nb_u, const double *AS, long long sAS,
                               const double *BS, long long sBS, const double *xi, long long sxi, double *xb, long long sxb, int nrhs = 1) {
  constexpr int NS = CPLX ? 2 : 1;
  const int e = blockIdx.y, r = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
  const int ni = ni_e ? ni_e[e] : ni_u, nb = nb_e ? nb_e[e] : nb_u;
  if (r >= nb) return;
  const double *A = AS + (long long)e * sAS * NS;
  for (int q = 0; q < nrhs; q++) {   // xi(ni, NR_RHS), xb(nb, NR_RHS): column q at offset q*ni / q*nb
  const double *x = xi + ((long long)e * sxi + (long long)q * ni) * NS;
  double sr = 0, si = 0;
  for (int c = lane; c < ni; c += 32) {
    const double *a = A + ((long long)r + (long long)nb * c) * NS;
    if (CPLX) { sr += a[0] * x[2 * c] - a[1] * x[2 * c + 1]; si += a[0] * x[2 * c + 1] + a[1] * x[2 * c]; }
    else sr += a[0] * x[c];
  }
  for (int o = 16; o; o >>= 1) { sr += __shfl_xor_sync(0xffffffffu, sr, o); if (CPLX) si += __shfl_xor_sync(0xffffffffu, si, o); }
  if (lane == 0) {
    const double *b = BS + (long long)e * sBS * NS + ((long long)q * nb + r) * NS;
    double *o = xb + (long long)e * sxb * NS + ((long long)q * nb + r) * NS;
    o[0] = b[0] - sr;
    if (CPLX) o[1] = b[1] - si;
  }
  }
}

// stc_bwd on the device-resident store: element e's factors at AS[e] / BS[e] (one warp per bubble row); grid (ceil(nbmax/8), nel)
template <bool CPLX>
__global__ void stc_bwd_ptr_kernel(const double *const *__restrict__ AS, const double *const *__restrict__ BS, const int *__restrict__ ni_e,
                                   const int *__restrict__ nb_e, const double *xi, long long sxi, double *xb, long long sxb, int nrhs) {
  constexpr int NS = CPLX ? 2 : 1;
  const int e = blockIdx.y, r = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
  const int ni = ni_e[e], nb = nb_e[e];
  if (r >= nb) return;
  const double *A = AS[e];
  for (int q = 0; q < nrhs; q++) {
  const double *x = xi + ((long long)e * sxi + (long long)q * ni) * NS;
  double sr = 0, si = 0;
  for (int c = lane; c < ni; c += 32) {
    const double *a = A + ((long long)r + (long long)nb * c) * NS;
    if (CPLX) { sr += a[0] * x[2 * c] - a[1] * x[2 * c + 1]; si += a[0] * x[2 * c + 1] + a[1] * x[2 * c]; }
    else sr += a[0] * x[c];
  }
  for (int o = 16; o; o >>= 1) { sr += __shfl_xor_sync(0xffffffffu, sr, o); if (CPLX) si += __shfl_xor_sync(0xffffffffu, si, o); }
  if (lane == 0) {
    const double *b = BS[e] + ((long long)q * nb + r) * NS;
    double *o = xb + (long long)e * sxb * NS + ((long long)q * nb + r) * NS;
    o[0] = b[0] - sr;
    if (CPLX) o[1] = b[1] - si;
  }
  }
}

// DPG element residual  eta^2 = || l~ - B~ u ||^2 = v^H A v,  v = [u ; -1],  A = [B~ l~]^H [B~ l~]  (lower triangle of Am after
// the normal equations; load at index M-1): the quantity elem_residual computes with its own Gram factorization
// (problems/MAXWELL/ULTRAWEAK_DPG/elem/elem_residual_maxwell.F90:246-552, POISSON/PRIMAL_DPG/elem_residual.F90).
// u is given in the caller's layout as xi (interface dofs) and xb (bubble dofs).  One CTA (256 threads) per element.
template <bool CPLX>
__global__ void __launch_bounds__(256) dpg_residual_kernel(DenseDims d, const double *Am, const int *__restrict__ ni_e, const int *__restrict__ nb_e,
                                                           const int *__restrict__ nip_e, const double *xi, long long sxi, const double *xb, long long sxb,
                                                           double *res) {
  constexpr int NS = CPLX ? 2 : 1;
  extern __shared__ double sv[];   // v: [2][M]
  __shared__ double red[8];
  const int e = blockIdx.x, M = d.M(), tid = threadIdx.x;
  const long long apl = (long long)d.a_plane();
  const double *Ar = Am + (long long)e * NS * apl, *Ai = Ar + apl;
  const int ni = ni_e[e], nb = nb_e[e];
  double *vr = sv, *vi = sv + M;
  for (int i = tid; i < M; i += blockDim.x) {
    double r = 0.0, im = 0.0;
    if (i < nb) { r = xb[((long long)e * sxb + i) * NS]; if (CPLX) im = xb[((long long)e * sxb + i) * NS + 1]; }
    else if (i >= d.nbp && i < d.nbp + ni) { const int k = i - d.nbp; r = xi[((long long)e * sxi + k) * NS]; if (CPLX) im = xi[((long long)e * sxi + k) * NS + 1]; }
    else if (i == d.nbp + (nip_e ? nip_e[e] : d.nil) - 1) r = -1.0;   // the element's load row
    vr[i] = r; vi[i] = im;
  }
  __syncthreads();
  // eta^2 = sum_i A_ii |v_i|^2 + 2 Re sum_{i>j} conj(v_i) A_ij v_j
  double acc = 0.0;
  const int warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  for (int i = warp; i < M; i += nw) {
    const double *ar = Ar + (long long)i * M, *ai = Ai + (long long)i * M;
    double yr = 0.0, yi = 0.0;   // y = sum_{j<i} A_ij v_j
    for (int j = lane; j < i; j += 32) {
      const double a = ar[j], b = CPLX ? ai[j] : 0.0;
      yr += a * vr[j] - b * vi[j]; yi += a * vi[j] + b * vr[j];
    }
    for (int o = 16; o; o >>= 1) { yr += __shfl_xor_sync(0xffffffffu, yr, o); yi += __shfl_xor_sync(0xffffffffu, yi, o); }
    if (lane == 0) acc += ar[i] * (vr[i] * vr[i] + vi[i] * vi[i]) + 2.0 * (vr[i] * yr + vi[i] * yi);
  }
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (tid == 0) { double s = 0.0; for (int w = 0; w < nw; w++) s += red[w]; res[e] = s; }
}

// ---------------------------------------------------------------------------------------------------------------
// Real-structured complex problems (DenseDims::rs): the dense phase worked on A~ (real) with A = T A~ T^H, T = diag(i^ph);
// the output kernels re-apply the phases.  ph of reference dof k: ultraweak Maxwell interface dofs 2j+ivar -> ivar (E-trace 0,
// H-trace 1), bubble dofs 6j+c -> (c >= 3) (E components 0, H components 1).  The complex load is carried as two real rows
// (lrow, lrow+1) = W~_B G~^-1 (Re, Im of the stored load row); b_t = -i * i^ph_t * (y_re - i y_im)   (see forms.hpp).
__device__ __forceinline__ int rs_phase_i(int k) { return k & 1; }
__device__ __forceinline__ int rs_phase_b(int k) { return (k % 6) >= 3; }
// multiply the real number v by i^(pr - pc) (pr, pc in {0,1})
__device__ __forceinline__ void rs_apply(double v, int pr, int pc, double &re, double &im) {
  const int d = pr - pc;
  re = d == 0 ? v : 0.0;
  im = d == 0 ? 0.0 : (d > 0 ? v : -v);
}
__device__ __forceinline__ void rs_load(double yre, double yim, int ph, double &re, double &im) {
  if (ph == 0) { re = -yim; im = -yre; }   // -i (y_re - i y_im)
  else { re = yre; im = -yim; }            //  i * -i (y_re - i y_im)
}

// grid = (ceil(ni/16), ceil(ni/16), batch), block (16,16): complex Aii (ni x ni), Bi (ni) from the real Schur complement
__global__ void scatter_condensed_rs_kernel(DenseDims d, const double *Am, OutMaps mp, double *Aii, double *Bi, long long sA, long long sB, int packed) {
  const int e = blockIdx.z;
  const int r = blockIdx.x * 16 + threadIdx.x, c = blockIdx.y * 16 + threadIdx.y;
  const int M = d.M();
  const double *S = Am + (long long)e * (long long)d.a_plane() + (long long)d.nbp * M + d.nbp;
  const int ni = mp.ni_e ? mp.ni_e[e] : d.ni, lrow0 = (mp.nip_e ? mp.nip_e[e] : d.nil) - d.nload;   // first pair of load rows
  if (r < ni && c < ni && (!packed || r >= c)) {
    const int a = r >= c ? r : c, b = r >= c ? c : r;
    double re, im;
    rs_apply(S[(long long)a * M + b], rs_phase_i(r), rs_phase_i(c), re, im);
    const long long at = packed ? (long long)r + ((long long)c * (2LL * ni - c - 1)) / 2 : (long long)r + (long long)ni * c;
    double *o = Aii + (long long)e * sA * 2 + at * 2;
    o[0] = re; o[1] = im;
  }
  if (blockIdx.y == 0 && threadIdx.y == 0 && r < ni) {
    for (int q = 0; q < d.nload / 2; q++) {
      const int lrow = lrow0 + 2 * q;
      double re, im;
      rs_load(S[(long long)lrow * M + r], S[(long long)(lrow + 1) * M + r], rs_phase_i(r), re, im);
      double *o = Bi + (long long)e * sB * 2 + ((long long)q * ni + r) * 2;
      o[0] = re; o[1] = im;
    }
  }
}

// grid = (ceil(nb/16), ceil(ni/16), batch), block (16,16): ASchur[b,i] = conj(Z[i,b]) = i^(ph_b - ph_i) Z~[i,b], BSchur like Bi
__global__ void scatter_schur_rs_kernel(DenseDims d, const double *Am, OutMaps mp, double *AS, double *BS, long long sAS, long long sBS) {
  const int e = blockIdx.z;
  const int bq = blockIdx.x * 16 + threadIdx.x, iq = blockIdx.y * 16 + threadIdx.y;
  const int M = d.M();
  const double *Z = Am + (long long)e * (long long)d.a_plane() + (long long)d.nbp * M;
  const int ni = mp.ni_e ? mp.ni_e[e] : d.ni, nb = mp.nb_e ? mp.nb_e[e] : d.nb, lrow0 = (mp.nip_e ? mp.nip_e[e] : d.nil) - d.nload;
  if (bq < nb && iq < ni) {
    double re, im;
    rs_apply(Z[(long long)iq * M + bq], rs_phase_b(bq), rs_phase_i(iq), re, im);
    double *o = AS + (long long)e * sAS * 2 + ((long long)bq + (long long)nb * iq) * 2;
    o[0] = re; o[1] = im;
  }
  if (blockIdx.y == 0 && threadIdx.y == 0 && bq < nb) {
    for (int q = 0; q < d.nload / 2; q++) {
      const int lrow = lrow0 + 2 * q;
      double re, im;
      rs_load(Z[(long long)lrow * M + bq], Z[(long long)(lrow + 1) * M + bq], rs_phase_b(bq), re, im);
      double *o = BS + (long long)e * sBS * 2 + ((long long)q * nb + bq) * 2;
      o[0] = re; o[1] = im;
    }
  }
}

// DPG residual in the real-structured form: with u~ = T^H u,  eta^2 = v1^T A~ v1 + v2^T A~ v2 on the uncondensed real normal
// equations (two load rows at M-2, M-1): with phi = i W~_B^T u~ - conj(load),  Im phi = [W~_B; l_re; l_im]^T v1, -Re phi = [..]^T v2,
// v1 = [Re u~ ; 0 ; 1], v2 = [Im u~ ; 1 ; 0].  One CTA (256 threads) per element.
__global__ void __launch_bounds__(256) dpg_residual_rs_kernel(DenseDims d, const double *Am, const int *__restrict__ ni_e, const int *__restrict__ nb_e,
                                                              const int *__restrict__ nip_e, const double *xi, long long sxi, const double *xb, long long sxb,
                                                              double *res) {
  extern __shared__ double sv[];   // v1, v2: [2][M]
  __shared__ double red[8];
  const int e = blockIdx.x, M = d.M(), tid = threadIdx.x;
  const double *A = Am + (long long)e * (long long)d.a_plane();
  const int ni = ni_e[e], nb = nb_e[e], l0 = d.nbp + (nip_e ? nip_e[e] : d.nil) - 2;   // the element's two load rows
  double *v1 = sv, *v2 = sv + M;
  for (int i = tid; i < M; i += blockDim.x) {
    double ur = 0.0, ui = 0.0;
    int ph = 0;
    if (i < nb) { ur = xb[((long long)e * sxb + i) * 2]; ui = xb[((long long)e * sxb + i) * 2 + 1]; ph = rs_phase_b(i); }
    else if (i >= d.nbp && i < d.nbp + ni) { const int k = i - d.nbp; ur = xi[((long long)e * sxi + k) * 2]; ui = xi[((long long)e * sxi + k) * 2 + 1]; ph = rs_phase_i(k); }
    // u~ = conj(i^ph) u
    double a = ph ? ui : ur, b = ph ? -ur : ui;
    if (i == l0) { a = 0.0; b = 1.0; }
    if (i == l0 + 1) { a = 1.0; b = 0.0; }
    v1[i] = a; v2[i] = b;
  }
  __syncthreads();
  double acc = 0.0;
  const int warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  for (int i = warp; i < M; i += nw) {
    const double *ar = A + (long long)i * M;
    double y1 = 0.0, y2 = 0.0;
    for (int j = lane; j < i; j += 32) { const double a = ar[j]; y1 += a * v1[j]; y2 += a * v2[j]; }
    for (int o = 16; o; o >>= 1) { y1 += __shfl_xor_sync(0xffffffffu, y1, o); y2 += __shfl_xor_sync(0xffffffffu, y2, o); }
    if (lane == 0) acc += ar[i] * (v1[i] * v1[i] + v2[i] * v2[i]) + 2.0 * (v1[i] * y1 + v2[i] * y2);
  }
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (tid == 0) { double s = 0.0; for (int w = 0; w < nw; w++) s += red[w]; res[e] = s; }
}

// opt-in of the residual kernels' dynamic shared memory (v: 2 x M doubles; M = 3328 at p = 7 ultraweak Maxwell exceeds the 48 KB default)
constexpr int RESID_SMEM_MAX = 160 * 1024;
static cudaError_t formats_configure() {
  cudaError_t e = cudaFuncSetAttribute(dpg_residual_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RESID_SMEM_MAX);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(dpg_residual_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RESID_SMEM_MAX);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(dpg_residual_rs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RESID_SMEM_MAX);
}

}  // namespace hp3d
