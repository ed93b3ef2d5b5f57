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

}  // namespace hp3d
