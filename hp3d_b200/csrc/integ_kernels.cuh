// integ_kernels.cuh -- batched, sum-factorised FP64 element integration for hexahedra (north-star subsystem 2).
//
// What it replaces in the reference (paths relative to trunk/):
//   * the quadrature-point loop of every elem_opt: shape3DH + geom3D (src/element/util/geom3D.F90:30-110,
//     geom.F90:30-113), Piola maps (e.g. problems/MAXWELL/ULTRAWEAK_DPG/elem/elem_opt.F90:236-325), user source getf;
//   * the BLAS3 integration calls that follow it (MAXWELL/ULTRAWEAK_DPG/elem/elem_opt.F90:339-470,
//     POISSON/PRIMAL_DPG/elem_opt.F90:219-269, POISSON/GALERKIN/elem_opt.F90:110-137, MAXWELL/GALERKIN/elem_opt.F90:120-149)
//   by the sum-factorised evaluation the reference itself uses in its fast path
//   (problems/LASER/UW_COUPLED/elem/elem_maxwell_fi_hexa.F90:534-1690): every integral is
//        M[(iA,jA,kA),(iB,jB,kB)] = sum_q  XA[iA][qx] XB[iB][qx] * YA[jA][qy] YB[jB][qy] * ZA[kA][qz] ZB[kB][qz] * F(q)
//   with 1-D tables X,Y,Z (values or derivatives of the H / Q bases, tables.hpp) and a per-point weight field F(q)
//   built from the Jacobian (metric tensors w*det*J^-1 J^-T, w/det*J^T J, w*J^-1, w*J/det) and material constants.
//
// Kernels:
//   geom_fields_kernel : one thread per (element, quadrature point): x, J, J^-1, det from the geometry dofs
//                        (tensor-product H1 functions), then all weight fields + source terms -> WF[e][field][q].
//   tp3_kernel         : one CTA per (element, block, iA, jA): x- and y-contractions into shared memory, z-contraction
//                        in registers, result written straight into the dense phase's buffers (W or Am).
//   const_rows_kernel  : copies the element-independent rows (trace pairings, see forms.hpp) into W.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hp3d {

__device__ __forceinline__ void dmma_tile(double &c0, double &c1, double a, double b) {   // 8x8x4 FP64 tensor-core tile product
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int MAXQ = 10;           // points / functions per axis (Gauss table limit of the reference)
constexpr int TABSZ = MAXQ * MAXQ; // one 1-D table
enum TabType { T_H = 0, T_DH = 1, T_Q = 2, T_ONE = 3 };

// weight fields per quadrature point
enum Field {
  F_D = 0,     // 6: w*det*(J^-1 J^-T), symmetric (00,11,22,01,02,12)     (H1 gradients / H(curl) values)
  F_C = 6,     // 6: w/det*(J^T J), symmetric                             (H(curl) curls)
  F_W = 12,    // w
  F_WDET = 13, // w*det
  F_WJI = 14,  // 9: w*dxi_a/dx_c  at [14+3a+c]
  F_WJD = 23,  // 9: w*dx_c/dxi_b/det at [23+3c+b]
  F_SRC = 32,  // 6: source terms (real problems: [32] = w*det*f ; Maxwell: re/im of w*det*(J^-1 zJ)_a at [32+2a], [33+2a])
  F_X = 38,    // 3: physical coordinates
  // constant permittivity TENSOR eps_t of ultraweak Maxwell (get_permittivity, elem_opt.F90:260-266), only filled when it is not
  // the identity: with eps_t eps_t^H = R + i S (R symmetric, S antisymmetric) and eps_t = er + i ei
  F_TD = 41,   // 6: w*det*(J^-1 R J^-T), symmetric (00,11,22,01,02,12)             Gram (za^H F, za^H F), real part
  F_TS = 47,   // 3: w*det*(J^-1 S J^-T) at (1,0),(2,0),(2,1)                        ... imaginary part
  F_T1R = 50,  // 9: w*(J^T er^T J^-T)(b,a) at [50+3b+a]                             Gram cross term (curl G, za^H F)
  F_T1I = 59,  // 9: the same with ei
  NFIELD = 68
};
// weight fields are stored [element][field][fs] with fs = nint rounded up to even: every field starts on a 16-byte boundary,
// so one field is one TMA bulk copy (cp.async.bulk needs 16-byte aligned addresses and sizes)
__host__ __device__ inline int wf_stride(int nint) { return (nint + 1) & ~1; }

__host__ __device__ inline int sym_idx(int a, int b) { return a == b ? a : (a + b + 2); }  // 01->3, 02->4, 12->5

struct GeomParams {
  int kind;            // HP3D_* problem kind
  int source;          // HP3D_SRC_*
  int icomp;           // 0-based component of the manufactured Maxwell solution
  double omega, eps, mu, sigma;
  int tensor;          // permittivity tensor fields wanted
  double tR[9], tS[9], ter[9], tei[9];   // R, S, er, ei (see Field), entry (i,j) at [3*i + j]
};

struct SigTables {               // element-signature tables, device memory
  const double *tab;             // [3 axes][4 types][TABSZ] ; entry [i*nq + q]
  const double *wq;              // [3][MAXQ] 1-D weights
  const int *hdof;               // [nH] geometry dofs: idx0 | idx1<<8 | idx2<<16 | (sign<0)<<24   (prism: t | zi<<8 | (sign<0)<<24)
  const double *ttab;            // prism: triangle tables of the geometry list [3][nT][nqt]
  int nT;
  int nq[3];
  int nH;
  int nint;
};

__device__ __forceinline__ void dev_sincos(double x, double &s, double &c) { sincos(x, &s, &c); }

// manufactured "sin" solutions of the reference's problem directories (isol = 1):
//   POISSON: u = sin(pi x) sin(pi y) sin(pi z), f = -Laplace u          (problems/POISSON/*/common/exact.F90, getf.F90)
//   MAXWELL: E = p e_ic, p = (1+i) sin(w x) sin(w y) sin(w z)            (MAXWELL/ULTRAWEAK_DPG/common/mfd_solutions.F90:80-100)
__device__ inline void sin_potential_hess(double a, const double x[3], double &p, double h[9]) {
  double s[3], c[3];
  for (int i = 0; i < 3; i++) dev_sincos(a * x[i], s[i], c[i]);
  const double a2 = a * a;
  p = s[0] * s[1] * s[2];
  h[0] = h[4] = h[8] = -a2 * p;
  h[1] = h[3] = a2 * c[0] * c[1] * s[2];
  h[2] = h[6] = a2 * c[0] * c[2] * s[1];
  h[5] = h[7] = a2 * c[1] * c[2] * s[0];
}

// All weight fields of one quadrature point from (x, J, w): determinant (Sarrus) and inverse by cofactors as
// src/element/util/geom.F90:57-113, then the Piola maps folded into per-point fields.  J[c + 3*d] = dx_c/dxi_d.
__device__ __forceinline__ void emit_fields(const GeomParams &gp, int e, int q, int nint, const double x[3], const double J[9], double w,
                                            const double *__restrict__ src_tab, long long src_ld, double *__restrict__ WF, int *__restrict__ info) {
  const double det = J[0] * J[4] * J[8] + J[1] * J[5] * J[6] + J[2] * J[3] * J[7] - J[2] * J[4] * J[6] - J[0] * J[5] * J[7] - J[1] * J[3] * J[8];
  if (!(det > 0.0)) info[e] = -1;
  double Ji[9];  // Ji[a + 3*c] = dxi_a/dx_c
  Ji[0] = (J[4] * J[8] - J[5] * J[7]) / det;
  Ji[1] = (-J[1] * J[8] + J[2] * J[7]) / det;
  Ji[2] = (J[1] * J[5] - J[2] * J[4]) / det;
  Ji[3] = (J[5] * J[6] - J[3] * J[8]) / det;
  Ji[4] = (J[0] * J[8] - J[2] * J[6]) / det;
  Ji[5] = (-J[0] * J[5] + J[2] * J[3]) / det;
  Ji[6] = (J[3] * J[7] - J[4] * J[6]) / det;
  Ji[7] = (-J[0] * J[7] + J[1] * J[6]) / det;
  Ji[8] = (J[0] * J[4] - J[1] * J[3]) / det;
  const long long fs = wf_stride(nint);
  double *F = WF + (long long)e * NFIELD * fs + q;
  const double wd = w * det, wod = w / det;
  for (int a = 0; a < 3; a++)
    for (int b = a; b < 3; b++) {
      double dd = 0, cc = 0;
      for (int c = 0; c < 3; c++) { dd += Ji[a + 3 * c] * Ji[b + 3 * c]; cc += J[c + 3 * a] * J[c + 3 * b]; }
      F[(F_D + sym_idx(a, b)) * fs] = wd * dd;
      F[(F_C + sym_idx(a, b)) * fs] = wod * cc;
    }
  F[F_W * fs] = w;
  F[F_WDET * fs] = wd;
  for (int a = 0; a < 3; a++)
    for (int c = 0; c < 3; c++) {
      F[(F_WJI + 3 * a + c) * fs] = w * Ji[a + 3 * c];
      F[(F_WJD + 3 * c + a) * fs] = wod * J[c + 3 * a];
    }
  for (int c = 0; c < 3; c++) F[(F_X + c) * fs] = x[c];
  if (gp.tensor) {
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) {
        double dd = 0, ss = 0, tr = 0, ti = 0;
        for (int d = 0; d < 3; d++)
          for (int d2 = 0; d2 < 3; d2++) {
            dd += Ji[a + 3 * d] * gp.tR[3 * d + d2] * Ji[b + 3 * d2];
            ss += Ji[a + 3 * d] * gp.tS[3 * d + d2] * Ji[b + 3 * d2];
            tr += J[d + 3 * a] * gp.ter[3 * d2 + d] * Ji[b + 3 * d2];   // (J^T er^T J^-T)(a,b)
            ti += J[d + 3 * a] * gp.tei[3 * d2 + d] * Ji[b + 3 * d2];
          }
        if (b >= a) F[(F_TD + sym_idx(a, b)) * fs] = wd * dd;
        if (a > b) F[(F_TS + (a == 1 ? 0 : b + 1)) * fs] = wd * ss;      // (1,0) -> 0, (2,0) -> 1, (2,1) -> 2
        F[(F_T1R + 3 * a + b) * fs] = w * tr;
        F[(F_T1I + 3 * a + b) * fs] = w * ti;
      }
  }
  // ---- source term
  double s[6] = {0, 0, 0, 0, 0, 0};
  if (gp.kind == 1 || gp.kind == 2) {  // Poisson: f(x)
    double f = 0.0;
    if (gp.source == 9) f = src_tab[(long long)e * src_ld + q];
    else if (gp.source == 1) { double p, h[9]; sin_potential_hess(3.14159265358979323846, x, p, h); f = -(h[0] + h[4] + h[8]); }
    s[0] = wd * f;
  } else {  // Maxwell: complex vector zJ(x)
    double jr[3] = {0, 0, 0}, ji[3] = {0, 0, 0};
    if (gp.source == 9) {
      const double *t = src_tab + (long long)e * src_ld + (long long)q * 6;
      for (int c = 0; c < 3; c++) { jr[c] = t[2 * c]; ji[c] = t[2 * c + 1]; }
      if (gp.kind == 3)   // Galerkin Maxwell: the load is -i w (J,F) (MAXWELL/GALERKIN/elem_opt.F90): store g = -i w J like the built-in source below
        for (int c = 0; c < 3; c++) { const double a = jr[c]; jr[c] = gp.omega * ji[c]; ji[c] = -gp.omega * a; }
    } else if (gp.source == 1) {
      double p, h[9], cc[3];
      sin_potential_hess(gp.omega, x, p, h);      // real profile; amplitude (1+i) applied below
      const int ic = gp.icomp;
      for (int c = 0; c < 3; c++) cc[c] = h[c + 3 * ic];
      cc[ic] -= h[0] + h[4] + h[8];               // curl curl (p e_ic) = grad(d_ic p) - Laplace(p) e_ic
      if (gp.kind == 4) {
        // J = curl H - i w eps E, H = curl E / (-i w mu)  (MAXWELL/ULTRAWEAK_DPG/getf.F90)  => J = (i/(w mu)) cc(1+i) - i w eps p(1+i) e_ic
        const double a = 1.0 / (gp.omega * gp.mu), b = gp.omega * gp.eps;
        for (int c = 0; c < 3; c++) { jr[c] = -a * cc[c]; ji[c] = a * cc[c]; }   // i*(1+i) = -1 + i
        jr[ic] += b * p; ji[ic] -= b * p;                                           // -i*(1+i) = 1 - i
      } else {
        // -i w J = curl(1/mu curl E) - (w^2 eps - i w sigma) E   (MAXWELL/GALERKIN/common/getf.F90:50-73)
        // store g = -i w J directly (the load vector is -i w (J,F))
        const double zr = gp.omega * gp.omega * gp.eps, zi = -gp.omega * gp.sigma;  // zb = zr + i zi
        for (int c = 0; c < 3; c++) { jr[c] = cc[c] / gp.mu; ji[c] = cc[c] / gp.mu; }
        // zb*(1+i)*p = (zr - zi) + i (zr + zi)
        jr[ic] -= (zr - zi) * p; ji[ic] -= (zr + zi) * p;
      }
    }
    for (int a = 0; a < 3; a++) {
      double gr = 0, gi = 0;
      for (int c = 0; c < 3; c++) { gr += Ji[a + 3 * c] * jr[c]; gi += Ji[a + 3 * c] * ji[c]; }
      s[2 * a] = wd * gr; s[2 * a + 1] = wd * gi;
    }
  }
  for (int i = 0; i < 6; i++) F[(F_SRC + i) * fs] = s[i];
}

// grid: ceil(nel*nint/128) x 1, block 128
__global__ void __launch_bounds__(128) geom_fields_kernel(SigTables sg, GeomParams gp, int nel, const double *__restrict__ xnod,
                                                          long long xnod_ld, const double *__restrict__ src_tab, long long src_ld,
                                                          double *__restrict__ WF, int *__restrict__ info) {
  __shared__ double sH[3][TABSZ], sdH[3][TABSZ];
  for (int i = threadIdx.x; i < 3 * TABSZ; i += blockDim.x) {
    int ax = i / TABSZ, r = i % TABSZ;
    sH[ax][r] = sg.tab[(ax * 4 + T_H) * TABSZ + r];
    sdH[ax][r] = sg.tab[(ax * 4 + T_DH) * TABSZ + r];
  }
  __syncthreads();
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)nel * sg.nint) return;
  const int e = (int)(gid / sg.nint), q = (int)(gid % sg.nint);
  const int qx = q % sg.nq[0], qy = (q / sg.nq[0]) % sg.nq[1], qz = q / (sg.nq[0] * sg.nq[1]);
  const double *xn = xnod + (long long)e * xnod_ld;
  double x[3] = {0, 0, 0}, J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // J[c + 3*d] = dx_c/dxi_d  (column-major like dxdxi)
  for (int k = 0; k < sg.nH; k++) {
    const int d = sg.hdof[k];
    const int i0 = d & 255, i1 = (d >> 8) & 255, i2 = (d >> 16) & 255;
    const double sgn = (d >> 24) ? -1.0 : 1.0;
    const double h0 = sH[0][i0 * sg.nq[0] + qx], h1 = sH[1][i1 * sg.nq[1] + qy], h2 = sH[2][i2 * sg.nq[2] + qz];
    const double g0 = sdH[0][i0 * sg.nq[0] + qx], g1 = sdH[1][i1 * sg.nq[1] + qy], g2 = sdH[2][i2 * sg.nq[2] + qz];
    const double v = sgn * h0 * h1 * h2, d0 = sgn * g0 * h1 * h2, d1 = sgn * h0 * g1 * h2, d2 = sgn * h0 * h1 * g2;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const double xc = xn[3 * k + c];
      x[c] += xc * v;
      J[c] += xc * d0; J[c + 3] += xc * d1; J[c + 6] += xc * d2;
    }
  }
  const double w = sg.wq[qx] * sg.wq[MAXQ + qy] * sg.wq[2 * MAXQ + qz];
  emit_fields(gp, e, q, sg.nint, x, J, w, src_tab, src_ld, WF, info);
}

// Prism variant: geometry dofs are sign * T[t](x,y) * H[zi](z) with the triangle tables (value, d/dx, d/dy) of the
// signature's geometry list in global memory; quadrature point q = qt + nqt*qz, weights wq[0..nqt) | wq[nqt..nqt+nqz).
// grid: ceil(nel*nint/128), block 128
__global__ void __launch_bounds__(128) geom_fields_prism_kernel(SigTables sg, GeomParams gp, int nel, const double *__restrict__ xnod,
                                                                long long xnod_ld, const double *__restrict__ src_tab, long long src_ld,
                                                                double *__restrict__ WF, int *__restrict__ info) {
  __shared__ double sH[TABSZ], sdH[TABSZ];
  for (int i = threadIdx.x; i < TABSZ; i += blockDim.x) {
    sH[i] = sg.tab[(2 * 4 + T_H) * TABSZ + i];
    sdH[i] = sg.tab[(2 * 4 + T_DH) * TABSZ + i];
  }
  __syncthreads();
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)nel * sg.nint) return;
  const int e = (int)(gid / sg.nint), q = (int)(gid % sg.nint);
  const int nqt = sg.nq[0], nqz = sg.nq[2], qt = q % nqt, qz = q / nqt;
  const double *xn = xnod + (long long)e * xnod_ld;
  const double *Tv = sg.ttab, *Tx = sg.ttab + (long long)sg.nT * nqt, *Ty = sg.ttab + 2LL * sg.nT * nqt;
  double x[3] = {0, 0, 0}, J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = 0; k < sg.nH; k++) {
    const int d = sg.hdof[k];
    const int t = d & 255, zi = (d >> 8) & 255;
    const double sgn = (d >> 24) ? -1.0 : 1.0;
    const double h = sgn * sH[zi * nqz + qz], dh = sgn * sdH[zi * nqz + qz];
    const double tv = __ldg(Tv + t * nqt + qt), tx = __ldg(Tx + t * nqt + qt), ty = __ldg(Ty + t * nqt + qt);
    const double v = tv * h, d0 = tx * h, d1 = ty * h, d2 = tv * dh;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const double xc = xn[3 * k + c];
      x[c] += xc * v;
      J[c] += xc * d0; J[c + 3] += xc * d1; J[c + 6] += xc * d2;
    }
  }
  const double w = sg.wq[qt] * sg.wq[nqt + qz];
  emit_fields(gp, e, q, sg.nint, x, J, w, src_tab, src_ld, WF, info);
}

// ------------------------------------------------------------------------------------------------------------
struct FamilyDesc { int n[3]; int tab[3]; };             // grid extents and base table type per axis
struct TermDesc { int dA, dB, field, slot; double coef; };  // derivative axis of the A / B factor (-1: none)
struct SlotDesc { int zA, zB; double c[2]; };            // z-axis table types; contribution to the two output channels
struct ChannelDesc {
  int mat;        // 0: W, 1: Am ; -1: unused
  int plane;      // 0 real, 1 imaginary
  int row0, col0; // affine target (row = row0 + lA) unless a map is given
  int rowmap, colmap;  // offsets into the map pool or -1.  map entry m: 0 skip, else target = |m|-1, sign = sgn(m)
};
struct BlockDesc { int famA, famB, t0, nt, s0, ns; ChannelDesc ch[2]; int is_load; };   // is_load: the block integrates a load vector (rows of the unit family)
struct WorkItem { short block, iA, jA, pad; };

struct MatTarget { double *base; long long batch, plane; int ld; };

struct Tp3Args {
  const double *tab;            // signature tables [3][4][TABSZ]
  const FamilyDesc *fam;
  const TermDesc *term;
  const SlotDesc *slot;
  const BlockDesc *block;
  const WorkItem *work;
  const int *maps;
  const double *WF;             // [nel][NFIELD][nint]
  int nq[3];
  int nint;
  MatTarget mat[2];
  int load_only;                // NR_RHS > 1, extra passes: only the load blocks run ...
  int load_shift;               // ... and their rows move down by this many rows (the q-th load's rows)
};

// z contraction shared by the hexahedron and prism kernels; thread item = (kA, (iB,jB)), registers over kB.
// sZ: the four z tables re-laid out [type][NMAX][NMAX] and ZERO padded (build_ztab), sU: [ns][NMAX][nij] partial sums with
// the rows >= nqz zeroed: every inner loop has compile-time bounds and immediate shared-memory offsets (the padding adds
// zeros instead of being predicated off).  lA = lA0 + strideA*kA, lB = ij + nij*kB.
template <int NMAX>
__device__ __forceinline__ void build_ztab(const double *sTabZ, double *sZ, int nqz) {
  for (int i = threadIdx.x; i < 4 * NMAX * NMAX; i += blockDim.x) {
    const int type = i / (NMAX * NMAX), r = (i / NMAX) % NMAX, q = i % NMAX;
    sZ[i] = (q < nqz && r * nqz + q < TABSZ) ? sTabZ[type * TABSZ + r * nqz + q] : 0.0;
  }
}
// z contraction on the FP64 tensor pipe.  For one (iA,jA) [hexahedron] / tA [prism] and every slot s
//   D_s[(kA,kB)][ij] = sum_qz (ZA_s[kA][qz] ZB_s[kB][qz]) U_s[qz][ij]
// is a small matrix product (M = nAz*nBz rows, K = nqz, N = nij columns): DMMA m8n8k4 tiles with the A fragment
// a = ZA[kA][qz] * ZB[kB][qz] built in registers once per CTA (it depends on the block only: Stage2Frag), the B fragment read
// from the y-contracted sums U in shared memory, and the two output channels formed from D_s with the slot's coefficients.
// An m-tile is one kA and eight kB; a warp owns ONE m-tile (the warps sharing an m-tile split its n-tiles; the host sizes the
// CTA as a multiple of the m-tile count).  Rows kB >= nBz and columns ij >= nij of a tile are computed and dropped.
constexpr int TP_SMAX = 5;
constexpr int TP_TMAX = 16;  // terms of a block kept in shared memory (more: read from global memory)   // slots whose A fragments a thread keeps in registers; forms with more slots per block are refused
template <int NMAX> struct Stage2Frag {
  static constexpr int KS = (NMAX + 3) / 4, TB = (NMAX + 7) / 8;
  double a[TP_SMAX][KS];
  int mt, part, nsplit;   // m-tile of the warp (-1: none), its share of the n-tiles
};
template <int NMAX>
__device__ __forceinline__ void tp_stage2_prepare(const Tp3Args &A, const BlockDesc &B, const double *sZ, int nAz, Stage2Frag<NMAX> &F) {
  constexpr int KS = Stage2Frag<NMAX>::KS, TB = Stage2Frag<NMAX>::TB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3, nwarp = blockDim.x >> 5;
  const int nM = nAz * TB;
  // warps per m-tile: floor(nwarp / nM), the first (nwarp mod nM) m-tiles get one more; with fewer warps than m-tiles a warp
  // walks several m-tiles and rebuilds its fragments (tp_stage2 handles that case itself)
  F.mt = -1; F.part = 0; F.nsplit = 1;
  if (nM <= nwarp) {
    const int base = nwarp / nM, extra = nwarp % nM;
    int w = warp, m = 0;
    if (w < extra * (base + 1)) { m = w / (base + 1); F.part = w % (base + 1); F.nsplit = base + 1; }
    else { w -= extra * (base + 1); m = extra + w / base; F.part = w % base; F.nsplit = base; }
    F.mt = m;
    const int kA = m / TB, kB = 8 * (m % TB) + gq;
#pragma unroll
    for (int s = 0; s < TP_SMAX; s++)
#pragma unroll
      for (int ks = 0; ks < KS; ks++) {
        const int qz = 4 * ks + tq;
        double v = 0.0;
        if (s < B.ns && kB < NMAX && qz < NMAX) { const SlotDesc S = A.slot[B.s0 + s]; v = sZ[(S.zA * NMAX + kA) * NMAX + qz] * sZ[(S.zB * NMAX + kB) * NMAX + qz]; }
        F.a[s][ks] = v;
      }
  }
}
template <int NMAX>
__device__ __forceinline__ void tp_stage2(const Tp3Args &A, const BlockDesc &B, const double *sZ, const double *sU, const double *sC, int e, int nij,
                                          int nBz, int nAz, int lA0, int strideA, Stage2Frag<NMAX> &F) {   // sC[2*s + ch]: coefficient of slot s in channel ch
  constexpr int KS = Stage2Frag<NMAX>::KS, TB = Stage2Frag<NMAX>::TB;
  const int lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int nM = nAz * TB, nN = (nij + 7) >> 3;
  const bool own = F.mt >= 0;   // fragments prepared once per CTA
  const ChannelDesc c0 = B.ch[0], c1 = B.ch[1];
  const bool has1 = c1.mat >= 0;
  for (int mt = own ? F.mt : warp; mt < nM; mt += (own ? nM : nwarp)) {
    const int kA = mt / TB, kB = 8 * (mt % TB) + gq;
    if (!own) {   // fewer warps than m-tiles: rebuild the fragments for this m-tile
#pragma unroll
      for (int s = 0; s < TP_SMAX; s++)
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
          const int qz = 4 * ks + tq;
          double v = 0.0;
          if (s < B.ns && kB < NMAX && qz < NMAX) { const SlotDesc S = A.slot[B.s0 + s]; v = sZ[(S.zA * NMAX + kA) * NMAX + qz] * sZ[(S.zB * NMAX + kB) * NMAX + qz]; }
          F.a[s][ks] = v;
        }
    }
    // output rows of this m-tile (one kA) in the two channels
    const int lA = lA0 + strideA * kA;
    double *dst[2]; double sg[2];
#pragma unroll
    for (int ch = 0; ch < 2; ch++) {
      const ChannelDesc C = ch ? c1 : c0;
      dst[ch] = nullptr; sg[ch] = 1.0;
      if (C.mat < 0 || kB >= nBz) continue;
      long long row;
      if (C.rowmap >= 0) { const int m = A.maps[C.rowmap + lA]; if (m == 0) continue; row = (m < 0 ? -m : m) - 1; if (m < 0) sg[ch] = -1.0; }
      else row = C.row0 + lA;
      if (B.is_load) row += A.load_shift;
      const MatTarget M = A.mat[C.mat];
      dst[ch] = M.base + (long long)e * M.batch + (long long)C.plane * M.plane + row * M.ld;
    }
    // column of the thread's first entry inside the output row for ij0 = 0 (32-bit arithmetic; the maps, if any, are applied per entry)
    const int cb0 = nij * kB + 2 * tq;
    const int cm0 = c0.colmap, cm1 = c1.colmap;
    int roff[KS];
#pragma unroll
    for (int ks = 0; ks < KS; ks++) roff[ks] = min(4 * ks + tq, NMAX - 1) * nij;   // rows qz = 4 ks + tq (clamped: the A fragment is zero there)
    const int ustride = NMAX * nij;
    for (int nt = own ? F.part : 0; nt < nN; nt += (own ? F.nsplit : 1)) {
      const int ij0 = 8 * nt;
      // B fragment: column ij0+gq (clamped: columns >= nij are dropped at the store)
      const double *Up = sU + min(ij0 + gq, nij - 1);
      double acc0[2] = {0.0, 0.0}, acc1[2] = {0.0, 0.0};
#pragma unroll
      for (int s = 0; s < TP_SMAX; s++) {
        if (s < B.ns) {
          double d0 = 0.0, d1 = 0.0;
#pragma unroll
          for (int ks = 0; ks < KS; ks++) dmma_tile(d0, d1, F.a[s][ks], Up[roff[ks]]);
          const double2 k = *reinterpret_cast<const double2 *>(sC + 2 * s);
          acc0[0] += k.x * d0; acc0[1] += k.x * d1;
          if (has1) { acc1[0] += k.y * d0; acc1[1] += k.y * d1; }
          Up += ustride;
        }
      }
      // the thread holds (kB, ij0 + 2 tq) and (kB, ij0 + 2 tq + 1) of both channels
      const int ij = ij0 + 2 * tq;
      if (ij >= nij) continue;
      const bool two = ij + 1 < nij;
#pragma unroll
      for (int ch = 0; ch < 2; ch++) {
        double *row = dst[ch];
        if (!row) continue;
        const int cm = ch ? cm1 : cm0;
        const double v0 = sg[ch] * (ch ? acc1[0] : acc0[0]), v1 = sg[ch] * (ch ? acc1[1] : acc0[1]);
        if (cm < 0) {
          double *q = row + (ch ? c1.col0 : c0.col0) + cb0 + ij0;
          q[0] = v0;
          if (two) q[1] = v1;
        } else {
          const int lB = cb0 + ij0;
          int m = A.maps[cm + lB];
          if (m != 0) row[(m < 0 ? -m : m) - 1] = m < 0 ? -v0 : v0;
          if (two) { m = A.maps[cm + lB + 1]; if (m != 0) row[(m < 0 ? -m : m) - 1] = m < 0 ? -v1 : v1; }
        }
      }
    }
  }
}

// ---- TMA (bulk async copy) + mbarrier helpers: global -> shared staging of the geometry weight fields
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
      "r"(parity)
      : "memory");
}

// Hexahedron kernel: one CTA per (element, block, iA); the CTA loops over jA.
//   * the weight fields of the block's terms (geometry Jacobians folded with the quadrature weights, geom_fields_kernel) are
//     staged in shared memory by TMA bulk copies, one per term, completing on one mbarrier while the 1-D tables are loaded;
//   * the x contraction depends on iA only: it is done ONCE per term (T1) and reused by all jA;
//   * per jA: y contraction of all terms into U (registers per slot, no accumulation through shared memory), then the z
//     contraction with the output column in registers (tp_stage2).
// dynamic smem: tables 12*TABSZ | F [nt][fs] | T1 [nt][nBx*nqy*nqz] | U [ns][nqz][nBx*nBy]   (offsets from the host, per signature)
template <int NMAX>
__global__ void __launch_bounds__(448, 2) tp3_kernel(Tp3Args A, int off_F, int off_T1, int off_U, int off_Q) {
  constexpr int NQP = NMAX + 2;   // padded qy extent of the T1 / Q rows (even: 16-byte loads; 80-byte row stride at NMAX = 8: conflict-free)
  extern __shared__ __align__(16) double sm[];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ int sbeg[TP_SMAX + 1];                 // terms of slot s: [sbeg[s], sbeg[s+1])  (terms are sorted by slot, forms.hpp BlockBuilder::finish)
  __shared__ __align__(16) double sC[TP_SMAX * 2];  // slot coefficients of the two channels
  double *sTab = sm, *sZ = sm + 12 * TABSZ, *sF = sm + off_F, *sT1 = sm + off_T1, *sU = sm + off_U, *sQ = sm + off_Q;   // sZ: 4*NMAX*NMAX
  const int e = blockIdx.y, tid = threadIdx.x;
  const WorkItem wi = A.work[blockIdx.x];
  // the block and its terms are read all over the kernel: shared-memory copies (global reads would be re-issued after every
  // global store of the output, which the compiler must assume to alias them)
  __shared__ BlockDesc sBlk;
  __shared__ TermDesc sTerm[TP_TMAX];
  if (tid == 0) sBlk = A.block[wi.block];
  __syncthreads();
  const BlockDesc &B = sBlk;
  if (A.load_only && !B.is_load) return;   // extra right-hand sides: only the load blocks are integrated again (CTA-uniform)
  for (int t = tid; t < B.nt && t < TP_TMAX; t += blockDim.x) sTerm[t] = A.term[B.t0 + t];
  const TermDesc *term = B.nt <= TP_TMAX ? sTerm : A.term + B.t0;
  const FamilyDesc fa = A.fam[B.famA], fb = A.fam[B.famB];
  const int iA = wi.iA;
  const int nqx = A.nq[0], nqy = A.nq[1], nqz = A.nq[2];
  const int nBx = fb.n[0], nBy = fb.n[1], nBz = fb.n[2], nAy = fa.n[1], nAz = fa.n[2];
  const int nij = nBx * nBy, t1sz = nBx * nqz * NQP, fs = wf_stride(A.nint), qsz = B.nt * nBy * NQP;
  const double *WFe = A.WF + (long long)e * NFIELD * fs;
  if (tid == 0) mbar_init(&mbar, 1);
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&mbar, (unsigned)(B.nt * fs * sizeof(double)));
    for (int t = 0; t < B.nt; t++) tma_bulk_g2s(sF + (size_t)t * fs, WFe + (long long)A.term[B.t0 + t].field * fs, (unsigned)(fs * sizeof(double)), &mbar);
  }
  if (tid <= B.ns) {   // first term of every slot
    int b = 0;
    while (b < B.nt && A.term[B.t0 + b].slot < tid) b++;   // (global copy: sTerm is not complete before the next barrier)
    sbeg[tid] = b;
    if (tid < B.ns) { sC[2 * tid] = A.slot[B.s0 + tid].c[0]; sC[2 * tid + 1] = A.slot[B.s0 + tid].c[1]; }
  }
  for (int i = tid; i < 12 * TABSZ; i += blockDim.x) sTab[i] = A.tab[i];
  for (int i = tid; i < B.ns * NMAX * nij; i += blockDim.x) sU[i] = 0.0;   // rows >= nqz stay zero
  for (int i = tid; i < B.nt * t1sz; i += blockDim.x) sT1[i] = 0.0;        // qy padding stays zero
  __syncthreads();
  build_ztab<NMAX>(sTab + 8 * TABSZ, sZ, nqz);
  auto tabp = [&](int axis, int type) { return sTab + (axis * 4 + type) * TABSZ; };
  // Q[jA parity][t][jB][qy] = YA_t[jA][qy] * YB_t[jB][qy] (zero for qy >= nqy): the y factors of every term for one jA, double-buffered
  // (the next jA's table is written while the z contraction of the current one runs)
  auto build_q = [&](int jA) {
    double *Q = sQ + (size_t)(jA & 1) * qsz;
    for (int o = tid; o < qsz; o += blockDim.x) {
      const int t = o / (nBy * NQP), r = o - t * nBy * NQP, jB = r / NQP, qy = r - jB * NQP;
      const TermDesc T = term[t];
      Q[o] = qy < nqy ? tabp(1, T.dA == 1 ? T_DH : fa.tab[1])[jA * nqy + qy] * tabp(1, T.dB == 1 ? T_DH : fb.tab[1])[jB * nqy + qy] : 0.0;
    }
  };
  build_q(0);
  mbar_wait(&mbar, 0);
  // ---- x contraction of every term (shared by all jA): T1[t][qz][iB][qy]
  for (int o = tid; o < B.nt * nBx * nqy * nqz; o += blockDim.x) {
    const int t = o / (nBx * nqy * nqz), r = o - t * nBx * nqy * nqz, iB = r % nBx, qyz = r / nBx, qy = qyz % nqy, qz = qyz / nqy;
    const TermDesc T = term[t];
    const double *XA = tabp(0, T.dA == 0 ? T_DH : fa.tab[0]) + iA * nqx;
    const double *XB = tabp(0, T.dB == 0 ? T_DH : fb.tab[0]) + iB * nqx;
    const double *f = sF + (size_t)t * fs + qyz * nqx;
    double sacc = 0.0;
    for (int qx = 0; qx < nqx; qx++) sacc += XA[qx] * XB[qx] * f[qx];
    sT1[(size_t)t * t1sz + ((size_t)qz * nBx + iB) * NQP + qy] = sacc * T.coef;
  }
  __syncthreads();   // sZ, sT1, Q(0) complete
  Stage2Frag<NMAX> frag;
  tp_stage2_prepare<NMAX>(A, B, sZ, nAz, frag);
  // the y-contraction item of this thread is the same for every jA (the CTA covers all items in one pass whenever it can)
  const int nitems = nij * nqz;
  const bool one_pass = nitems <= (int)blockDim.x;
  int my_q0 = 0, my_t0 = 0;
  if (one_pass && tid < nitems) { const int ij = tid % nij, qz = tid / nij; my_q0 = (ij / nBx) * NQP; my_t0 = (qz * nBx + ij % nBx) * NQP; }
  const int qstride = nBy * NQP;
  for (int jA = 0; jA < nAy; jA++) {
    // ---- y contraction: U[slot][qz][ij] = sum over the slot's terms and qy of Q[t][jB][qy] * T1[t][qz][iB][qy]
    const double *Q = sQ + (size_t)(jA & 1) * qsz;
    for (int o = tid; o < nitems; o += blockDim.x) {
      int q0off = my_q0, t0off = my_t0;
      if (!one_pass) { const int ij = o % nij, qz = o / nij; q0off = (ij / nBx) * NQP; t0off = (qz * nBx + ij % nBx) * NQP; }
      const double *q0 = Q + q0off, *t0 = sT1 + t0off;
      for (int sl = 0; sl < B.ns; sl++) {
        double acc = 0.0;
        for (int t = sbeg[sl]; t < sbeg[sl + 1]; t++) {
          const double2 *q = reinterpret_cast<const double2 *>(q0 + t * qstride), *t1 = reinterpret_cast<const double2 *>(t0 + t * t1sz);
#pragma unroll
          for (int h = 0; h < NMAX / 2; h++) { const double2 a = q[h], b = t1[h]; acc += a.x * b.x; acc += a.y * b.y; }
        }
        sU[sl * NMAX * nij + o] = acc;   // o = qz*nij + ij, qz < nqz
      }
    }
    __syncthreads();
    // ---- z contraction and output (tensor pipe), then the next jA's Q
    tp_stage2<NMAX>(A, B, sZ, sU, sC, e, nij, nBz, nAz, iA + fa.n[0] * jA, fa.n[0] * fa.n[1], frag);
    if (jA + 1 < nAy) build_q(jA + 1);
    __syncthreads();
  }
}

// Prism kernel: one CTA per (element, block, tA).  A family is a (triangle list) x (z table) grid:
//   M[(tA,kA),(tB,kB)] = sum_{qt,qz} TA_cA[tA][qt] TB_cB[tB][qt] ZA[kA][qz] ZB[kB][qz] F(qt + nqt*qz)
// stage 1 (per term): G[qt][qz] = TA_cA[tA][qt] * F[qt,qz] in shared memory, then U[slot][qz][tB] += coef * sum_qt TB_cB[tB][qt] G[qt][qz];
// stage 2: the z contraction of the hexahedron kernel.  Triangle tables live in global memory (L1/L2-resident, read-only).
// dynamic smem: z tables 4*TABSZ | G [nqt*nqz] | U [ns][nqz][nTB]
template <int NMAX>
__global__ void __launch_bounds__(384, 2) tp2_kernel(Tp3Args A, const double *__restrict__ ttab, int smem_u_off) {
  extern __shared__ __align__(16) double sm[];
  double *sTabZ = sm, *sZ = sm + 4 * TABSZ, *sG = sZ + 4 * NMAX * NMAX, *sU = sm + smem_u_off;
  __shared__ __align__(16) double sC[TP_SMAX * 2];   // slot coefficients of the two channels
  const int e = blockIdx.y;
  const WorkItem wi = A.work[blockIdx.x];
  __shared__ BlockDesc sBlk;
  if (threadIdx.x == 0) sBlk = A.block[wi.block];
  __syncthreads();
  const BlockDesc &B = sBlk;
  if (A.load_only && !B.is_load) return;   // extra right-hand sides: only the load blocks are integrated again (CTA-uniform)
  const FamilyDesc fa = A.fam[B.famA], fb = A.fam[B.famB];
  const int tA = wi.iA;
  const int nqt = A.nq[0], nqz = A.nq[2];
  const int nTA = fa.n[0], nTB = fb.n[0], nBz = fb.n[2], nAz = fa.n[2];
  for (int i = threadIdx.x; i < 4 * TABSZ; i += blockDim.x) sTabZ[i] = A.tab[8 * TABSZ + i];
  if (threadIdx.x < B.ns) { sC[2 * threadIdx.x] = A.slot[B.s0 + threadIdx.x].c[0]; sC[2 * threadIdx.x + 1] = A.slot[B.s0 + threadIdx.x].c[1]; }
  for (int i = threadIdx.x; i < B.ns * NMAX * nTB; i += blockDim.x) sU[i] = 0.0;
  __syncthreads();
  build_ztab<NMAX>(sTabZ, sZ, nqz);
  __syncthreads();
  const int fs = wf_stride(A.nint);
  const double *WFe = A.WF + (long long)e * NFIELD * fs;
  for (int t = 0; t < B.nt; t++) {
    const TermDesc T = A.term[B.t0 + t];
    const double *TA = ttab + fa.tab[0] + ((long long)T.dA * nTA + tA) * nqt;
    const double *TB = ttab + fb.tab[0] + (long long)T.dB * nTB * nqt;
    const double *Fq = WFe + (long long)T.field * fs;
    for (int o = threadIdx.x; o < nqt * nqz; o += blockDim.x) sG[o] = __ldg(TA + o % nqt) * Fq[o] * T.coef;   // [qz][qt]
    __syncthreads();
    double *U = sU + T.slot * NMAX * nTB;
    for (int o = threadIdx.x; o < nTB * nqz; o += blockDim.x) {
      const int tB = o % nTB, qz = o / nTB;
      const double *tb = TB + (long long)tB * nqt, *g = sG + qz * nqt;
      double s = 0.0;
      for (int qt = 0; qt < nqt; qt++) s += __ldg(tb + qt) * g[qt];
      U[qz * nTB + tB] += s;
    }
    __syncthreads();
  }
  Stage2Frag<NMAX> frag;
  tp_stage2_prepare<NMAX>(A, B, sZ, nAz, frag);
  tp_stage2<NMAX>(A, B, sZ, sU, sC, e, nTB, nBz, nAz, tA, nTA, frag);
}

// Element-independent rows of W (trace pairings): W[e][plane 0][crow[r]][0..ncol) = CW[r][0..ncol)
// grid (ceil(ncol/256), nrows, nel)
__global__ void const_rows_kernel(const double *__restrict__ CW, const int *__restrict__ crow, int ncol, MatTarget M) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncol) return;
  const int r = blockIdx.y, e = blockIdx.z;
  M.base[(long long)e * M.batch + (long long)crow[r] * M.ld + c] = CW[(long long)r * ncol + c];
}

// Zero the part of W the dense phase reads before the integration writes into it (structural zeros of the forms, padding rows
// and columns): Gram rows r < np up to the end of their diagonal 64-tile, the trial rows r >= np completely.  The upper
// block-triangle of the Gram (28 % of W at config 3) is never read and is left alone.  grid (R/16, nel * planes), 256 threads.
__global__ void __launch_bounds__(256) zero_w_kernel(double *W, long long plane, int R, int np) {
  double *base = W + (long long)blockIdx.y * plane;
  const int r0 = blockIdx.x * 16;
  const double2 z = make_double2(0.0, 0.0);
  for (int r = r0; r < min(R, r0 + 16); r++) {
    const int w = r < np ? min(np, (r / 64 + 1) * 64) : np;
    double2 *row = reinterpret_cast<double2 *>(base + (long long)r * np);
    for (int c = threadIdx.x; c < w / 2; c += 256) row[c] = z;
  }
}

// unit diagonal on rows [n0,n1) of the real plane (keeps padded Gram rows regular); grid (ceil((n1-n0)/64), nel)
__global__ void unit_diag_kernel(MatTarget M, int n0, int n1) {
  const int i = n0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n1) M.base[(long long)blockIdx.y * M.batch + (long long)i * M.ld + i] = 1.0;
}

}  // namespace hp3d
