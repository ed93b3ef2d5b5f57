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
};

// grid = (ceil(ni/16), ceil(ni/16)+extras, batch), block (16,16).  Writes Aii (ni x ni), Bi (ni).
template <bool CPLX>
__global__ void scatter_condensed_kernel(DenseDims d, const double *Am, OutMaps mp, double *Aii, double *Bi, long long sA, long long sB) {
  const int e = blockIdx.z;
  const int r = blockIdx.x * 16 + threadIdx.x, c = blockIdx.y * 16 + threadIdx.y;
  const int M = d.M();
  const long long apl = (long long)d.a_plane();
  const double *S = Am + (long long)e * (CPLX ? 2 : 1) * apl + (long long)d.nbp * M + d.nbp;
  const int *pi = mp.perm_i + (long long)e * mp.perm_stride_i;
  const double *si = mp.sgn_i + (long long)e * mp.sgn_stride_i;
  constexpr int NS = CPLX ? 2 : 1;
  const int ni = mp.ni_e ? mp.ni_e[e] : d.ni, lrow = d.nip - 1;
  if (r < ni && c < ni) {
    int ir = pi[r], ic = pi[c];
    double s = si[r] * si[c];
    int a = ir >= ic ? ir : ic, b = ir >= ic ? ic : ir;
    double vr = S[(long long)a * M + b], vi = 0.0;
    if (CPLX) { vi = S[apl + (long long)a * M + b]; if (ir < ic) vi = -vi; if (ir == ic) vi = 0.0; }
    double *o = Aii + (long long)e * sA * NS + ((long long)r + (long long)ni * c) * NS;
    o[0] = s * vr;
    if (CPLX) o[1] = s * vi;
  }
  if (blockIdx.y == 0 && threadIdx.y == 0 && r < ni) {
    int ir = pi[r];
    double *o = Bi + (long long)e * sB * NS + (long long)r * NS;
    o[0] = si[r] * S[(long long)lrow * M + ir];
    if (CPLX) o[1] = -si[r] * S[apl + (long long)lrow * M + ir];   // b_i = conj(load row)
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
  const int ni = mp.ni_e ? mp.ni_e[e] : d.ni, nb = mp.nb_e ? mp.nb_e[e] : d.nb, lrow = d.nip - 1;
  if (bq < nb && iq < ni) {
    int ib = pb[bq], ii = pi[iq];
    double s = sb[bq] * si[iq];
    double *o = AS + (long long)e * sAS * NS + ((long long)bq + (long long)nb * iq) * NS;
    o[0] = s * Z[(long long)ii * M + ib];
    if (CPLX) o[1] = -s * Z[apl + (long long)ii * M + ib];
  }
  if (blockIdx.y == 0 && threadIdx.y == 0 && bq < nb) {
    int ib = pb[bq];
    double *o = BS + (long long)e * sBS * NS + (long long)bq * NS;
    o[0] = sb[bq] * Z[(long long)lrow * M + ib];
    if (CPLX) o[1] = -sb[bq] * Z[apl + (long long)lrow * M + ib];
  }
}

// xb = BSchur - ASchur * xi (stc_bwd, src/modules/stc.F90:661-677), one warp per (bubble row, element); ASchur (nb x ni) and
// xi / xb in the caller's layout (column-major, interleaved complex), per-element sizes.
template <bool CPLX>
__global__ void stc_bwd_kernel(const int *__restrict__ ni_e, const int *__restrict__ nb_e, int ni_u, int nb_u, const double *AS, long long sAS,
                               const double *BS, long long sBS, const double *xi, long long sxi, double *xb, long long sxb) {
  constexpr int NS = CPLX ? 2 : 1;
  const int e = blockIdx.y, r = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
  const int ni = ni_e ? ni_e[e] : ni_u, nb = nb_e ? nb_e[e] : nb_u;
  if (r >= nb) return;
  const double *A = AS + (long long)e * sAS * NS, *x = xi + (long long)e * sxi * NS;
  double sr = 0, si = 0;
  for (int c = lane; c < ni; c += 32) {
    const double *a = A + ((long long)r + (long long)nb * c) * NS;
    if (CPLX) { sr += a[0] * x[2 * c] - a[1] * x[2 * c + 1]; si += a[0] * x[2 * c + 1] + a[1] * x[2 * c]; }
    else sr += a[0] * x[c];
  }
  for (int o = 16; o; o >>= 1) { sr += __shfl_xor_sync(0xffffffffu, sr, o); if (CPLX) si += __shfl_xor_sync(0xffffffffu, si, o); }
  if (lane == 0) {
    const double *b = BS + (long long)e * sBS * NS + (long long)r * NS;
    double *o = xb + (long long)e * sxb * NS + (long long)r * NS;
    o[0] = b[0] - sr;
    if (CPLX) o[1] = b[1] - si;
  }
}

// DPG element residual  eta^2 = || l~ - B~ u ||^2 = v^H A v,  v = [u ; -1],  A = [B~ l~]^H [B~ l~]  (lower triangle of Am after
// the normal equations; load at index M-1): the quantity elem_residual computes with its own Gram factorization
// (problems/MAXWELL/ULTRAWEAK_DPG/elem/elem_residual_maxwell.F90:246-552, POISSON/PRIMAL_DPG/elem_residual.F90).
// u is given in the caller's layout as xi (interface dofs) and xb (bubble dofs).  One CTA (256 threads) per element.
template <bool CPLX>
__global__ void __launch_bounds__(256) dpg_residual_kernel(DenseDims d, const double *Am, const int *__restrict__ ni_e, const int *__restrict__ nb_e,
                                                           const double *xi, long long sxi, const double *xb, long long sxb, double *res) {
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
    else if (i == M - 1) r = -1.0;
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

}  // namespace hp3d
