// celem_kernels.cuh -- what celem_systemI does after elem + stc_fwd_wrapper (src/constrs/celem_systemI.F90:543-785), on the
// condensed element systems still resident in HBM:
//   ZAMOD = C^T A C, ZBMOD = C^T b   constrained-approximation transform (:553-716)
//   ZBMOD -= ZAMOD(:,k2) ZDOFD(k2)   Dirichlet lift (:720-731)
//   Zbload / Zastif                  compression through NEXTRACT, ISYM_FLAG 1/2/3 (:735-781)
//   IRN / JCN                        optional COO indices of par_mumps_sc.F90:433-448
// HBM-bound gather: every entry of the compressed matrix is formed from the |list(k1)| x |list(k2)| connected entries of A
// (1 x 1 for an unconstrained dof), summed in the order the reference's loops meet them and WITHOUT fused multiply-adds, so
// the result is bit-identical to the host code it replaces.
#pragma once
#include <cuda_runtime.h>

namespace hp3d {

struct CelemArgs {
  // constraint data of the whole call (device copies of the caller's arrays; index VALUES are the reference's 1-based ones)
  const long long *mptr, *cptr, *xptr;
  const int *cidx; const double *cval;
  const double *zdofd;
  const int *nextract, *lcon;
  const long long *dptr; const int *dlist;   // Dirichlet dofs of element e: dlist[dptr[e] .. dptr[e+1]) = ll-1, ascending
  // this chunk
  const int *cel;        // [n] caller element index of slot i
  const int *ni_e;       // [n] interface dofs of slot i (leading dimension of its Aii)
  const double *Aii, *Bi; long long sA, sB;   // condensed systems, slot strides in scalars
  double *Z, *zb; int *irn, *jcn; long long sZ, sZb;   // staging of Zastif / Zbload / IRN / JCN, slot strides in scalars
  int isym;
};

// ZAMOD(g1,g2) = sum_b ( sum_a A(r_a, r_b) c_a ) c_b  -- the order of celem_systemI's two passes (AAUX, then ZAMOD)
template <bool CPLX>
__device__ __forceinline__ void zamod_entry(const CelemArgs &a, const double *A, int ni, long long g1, long long g2, double &re, double &im) {
  constexpr int NS = CPLX ? 2 : 1;
  re = 0.0; im = 0.0;
  const long long a0 = a.cptr[g1], a1 = a.cptr[g1 + 1], b0 = a.cptr[g2], b1 = a.cptr[g2 + 1];
  for (long long b = b0; b < b1; b++) {
    const long long col = (long long)(a.cidx[b] - 1) * ni;
    const double vb = a.cval[b];
    double xr = 0.0, xi = 0.0;
    for (long long q = a0; q < a1; q++) {
      const double *p = A + (col + (a.cidx[q] - 1)) * NS;
      const double va = a.cval[q];
      if (CPLX) {   // one 16-byte load per complex entry (the staging buffers are 256-byte aligned)
        const double2 v = *reinterpret_cast<const double2 *>(p);
        xr = __dadd_rn(xr, __dmul_rn(v.x, va)); xi = __dadd_rn(xi, __dmul_rn(v.y, va));
      } else xr = __dadd_rn(xr, __dmul_rn(p[0], va));
    }
    re = __dadd_rn(re, __dmul_rn(xr, vb));
    if (CPLX) im = __dadd_rn(im, __dmul_rn(xi, vb));
  }
}

// Zbload: grid (ceil(nc_max/128), n), block 128; one thread per compressed dof
template <bool CPLX>
__global__ void __launch_bounds__(128) celem_load_kernel(CelemArgs a) {
  constexpr int NS = CPLX ? 2 : 1;
  const int i = blockIdx.y, e = a.cel[i];
  const long long x0 = a.xptr[e], m0 = a.mptr[e];
  const int nc = (int)(a.xptr[e + 1] - x0), ni = a.ni_e[i];
  const int l1 = blockIdx.x * 128 + threadIdx.x;
  if (l1 >= nc) return;
  const double *A = a.Aii + (long long)i * a.sA * NS, *B = a.Bi + (long long)i * a.sB * NS;
  const long long g1 = m0 + a.nextract[x0 + l1] - 1;
  double br = 0.0, bi = 0.0;
  for (long long q = a.cptr[g1]; q < a.cptr[g1 + 1]; q++) {
    const double *p = B + (long long)(a.cidx[q] - 1) * NS;
    const double va = a.cval[q];
    br = __dadd_rn(br, __dmul_rn(p[0], va));
    if (CPLX) bi = __dadd_rn(bi, __dmul_rn(p[1], va));
  }
  for (long long q = a.dptr[e]; q < a.dptr[e + 1]; q++) {   // Dirichlet lift, k2 ascending as celem_systemI.F90:720-731
    const long long g2 = m0 + a.dlist[q];
    double zr, zi;
    zamod_entry<CPLX>(a, A, ni, g1, g2, zr, zi);
    const double *d = a.zdofd + g2 * NS;
    if (CPLX) {
      const double pr = __dsub_rn(__dmul_rn(zr, d[0]), __dmul_rn(zi, d[1])), pi = __dadd_rn(__dmul_rn(zr, d[1]), __dmul_rn(zi, d[0]));
      br = __dsub_rn(br, pr); bi = __dsub_rn(bi, pi);
    } else br = __dsub_rn(br, __dmul_rn(zr, d[0]));
  }
  double *o = a.zb + ((long long)i * a.sZb + l1) * NS;
  o[0] = br;
  if (CPLX) o[1] = bi;
}

// one modified dof's list, staged in shared memory: begin / length in cidx,cval and the first entry (the whole list for an
// unconstrained dof)
struct DofList { long long b; int n, r; double c; };

// ZAMOD(k1,k2) from two staged lists; same summation order as zamod_entry
template <bool CPLX>
__device__ __forceinline__ void zamod_lists(const CelemArgs &a, const double *A, int ni, const DofList &L1, const DofList &L2, double &re, double &im) {
  constexpr int NS = CPLX ? 2 : 1;
  if (L1.n == 1 && L2.n == 1) {   // regular dofs: (0 + A c1) then (0 + x c2)
    const double *p = A + ((long long)L2.r * ni + L1.r) * NS;
    if (CPLX) { const double2 v = *reinterpret_cast<const double2 *>(p); re = __dmul_rn(__dmul_rn(v.x, L1.c), L2.c); im = __dmul_rn(__dmul_rn(v.y, L1.c), L2.c); }
    else { re = __dmul_rn(__dmul_rn(p[0], L1.c), L2.c); im = 0.0; }
    return;
  }
  re = 0.0; im = 0.0;
  for (int jb = 0; jb < L2.n; jb++) {
    const long long col = (long long)(jb ? a.cidx[L2.b + jb] - 1 : L2.r) * ni;
    const double vb = jb ? a.cval[L2.b + jb] : L2.c;
    double xr = 0.0, xi = 0.0;
    for (int ja = 0; ja < L1.n; ja++) {
      const double *p = A + (col + (ja ? a.cidx[L1.b + ja] - 1 : L1.r)) * NS;
      const double va = ja ? a.cval[L1.b + ja] : L1.c;
      if (CPLX) { const double2 v = *reinterpret_cast<const double2 *>(p); xr = __dadd_rn(xr, __dmul_rn(v.x, va)); xi = __dadd_rn(xi, __dmul_rn(v.y, va)); }
      else xr = __dadd_rn(xr, __dmul_rn(p[0], va));
    }
    re = __dadd_rn(re, __dmul_rn(xr, vb));
    if (CPLX) im = __dadd_rn(im, __dmul_rn(xi, vb));
  }
}

// Zastif (+ IRN/JCN): grid (ceil(nc_max/32), ceil(nc_max/32), n), block (32,8); a 32x32 tile of (l1,l2) per CTA.
// The 32 + 32 dof lists of the tile are staged in shared memory once (index metadata would otherwise cost more traffic than
// the matrix itself).  A is read with the threads along l1 (its rows: coalesced in the column-major condensed matrix); the
// row-major / packed outputs are written with the threads along l2 after a transpose through shared memory.
template <bool CPLX>
__global__ void __launch_bounds__(256) celem_compress_kernel(CelemArgs a) {
  constexpr int NS = CPLX ? 2 : 1;
  const int i = blockIdx.z, e = a.cel[i];
  const long long x0 = a.xptr[e], m0 = a.mptr[e];
  const int nc = (int)(a.xptr[e + 1] - x0), ni = a.ni_e[i];
  const int t1 = blockIdx.x, t2 = blockIdx.y;
  if (t1 * 32 >= nc || t2 * 32 >= nc) return;
  if (a.isym == 1 && t2 > t1) return;
  const double *A = a.Aii + (long long)i * a.sA * NS;
  double *Z = a.Z + (long long)i * a.sZ * NS;
  int *irn = a.irn ? a.irn + (long long)i * a.sZ : nullptr, *jcn = a.jcn ? a.jcn + (long long)i * a.sZ : nullptr;
  __shared__ double sre[32][33], sim[CPLX ? 32 : 1][33];
  __shared__ DofList sl[2][32];
  __shared__ int slc[2][32];   // LCON of the tile's rows / columns
  const int tx = threadIdx.x, ty = threadIdx.y;
  if (ty < 2) {
    const int l = (ty == 0 ? t1 : t2) * 32 + tx;
    DofList d{0, 0, 0, 0.0};
    if (l < nc) {
      const long long g = m0 + a.nextract[x0 + l] - 1;
      d.b = a.cptr[g]; d.n = (int)(a.cptr[g + 1] - d.b);
      if (d.n > 0) { d.r = a.cidx[d.b] - 1; d.c = a.cval[d.b]; }
      if (irn) slc[ty][tx] = a.lcon[x0 + l];
    }
    sl[ty][tx] = d;
  }
  __syncthreads();
  {
    const int l1 = t1 * 32 + tx;
    const DofList L1 = sl[0][tx];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int c = ty + 8 * j, l2 = t2 * 32 + c;
      if (l1 >= nc || l2 >= nc || (a.isym == 1 && l2 > l1)) continue;
      const DofList L2 = sl[1][c];
      double zr, zi;
      zamod_lists<CPLX>(a, A, ni, L1, L2, zr, zi);
      if (a.isym == 1) {   // (ZAMOD(k1,k2) + ZAMOD(k2,k1)) / 2
        double wr, wi;
        zamod_lists<CPLX>(a, A, ni, L2, L1, wr, wi);
        zr = __ddiv_rn(__dadd_rn(zr, wr), 2.0);
        if (CPLX) zi = __ddiv_rn(__dadd_rn(zi, wi), 2.0);
      }
      if (a.isym == 3) {   // column-major: k = l2*nc + l1, the threads already run along l1
        const long long k = (long long)l2 * nc + l1;
        if (CPLX) *reinterpret_cast<double2 *>(Z + k * NS) = make_double2(zr, zi);
        else Z[k] = zr;
        if (irn) { irn[k] = slc[0][tx]; jcn[k] = slc[1][c]; }
      } else {
        sre[c][tx] = zr;
        if (CPLX) sim[c][tx] = zi;
      }
    }
  }
  if (a.isym == 3) return;
  __syncthreads();
  const int l2 = t2 * 32 + tx;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int r = ty + 8 * j, l1 = t1 * 32 + r;
    if (l1 >= nc || l2 >= nc || (a.isym == 1 && l2 > l1)) continue;
    const long long k = a.isym == 2 ? (long long)l1 * nc + l2 : (long long)l1 * (l1 + 1) / 2 + l2;
    if (CPLX) *reinterpret_cast<double2 *>(Z + k * NS) = make_double2(sre[tx][r], sim[tx][r]);
    else Z[k] = sre[tx][r];
    if (irn) { irn[k] = slc[0][r]; jcn[k] = slc[1][tx]; }
  }
}

}  // namespace hp3d
