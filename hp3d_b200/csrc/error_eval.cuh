// error_eval.cuh -- batched solution evaluation and element error (error-evaluation half):
//   soleval        src/element/util/soleval.F90:30-283        geometry map + solution of one family at a master point
//   element_error  src/element/util/compute_error.F90:226-579 error / norm of one physical attribute over an element,
//                  quadrature set_3Dint with INTEGRATION = 2 (:305-307), weight wa*rjac
// for the FIELD variable of the plan's problem: H1 (Poisson), H(curl) (Maxwell Galerkin), L2 x 6 (ultraweak Maxwell).
// Host: per element signature the shape functions of the geometry (H1) and of the field family are tabulated ONCE at the
// master quadrature points (hexahedron: signed tensor products of the 1-D tables; prism: triangle function x 1-D table).
// Device: one CTA per element; a thread owns a quadrature point: x, J = sum xnod_k (phi_k, grad phi_k), Sarrus determinant and
// cofactor inverse (geom.F90:57-113), solution in master coordinates = table x dofs, Piola maps (J^-T for gradients and H(curl)
// values, J/det for curls, 1/det for L2), exact solution (built-in isol = 1 or a caller table), block reduction.
#pragma once
#include "forms.hpp"
#include "forms_prism.hpp"
#include "integ_kernels.cuh"

#include <string>
#include <vector>

namespace hp3d {

enum ErrSpace { ES_H1 = 0, ES_HCURL = 1, ES_L2 = 3 };

struct ErrSigHost {
  int etype = 1, space = 0, ncomp = 1, nint = 0, nH = 0, nF = 0;
  std::vector<double> w;       // [nint]
  std::vector<double> tabH;    // [4][nH][nint]: phi, d/dxi1, d/dxi2, d/dxi3
  std::vector<double> tabF;    // H1: aliases tabH ; H(curl): [6][nF][nint] (E^ 3, curl^ 3) ; L2: [1][nF][nint]
  std::string err;
};

namespace detail {

// values / master-coordinate derivatives of the hexahedron's functions at one point, reference dof order
//   space H1:    val[k] = phi, der[3k..] = grad ; H(curl): val[3k..] = E, der[3k..] = curl ; L2: val[k] = q
inline void hexa_shape_at(int space, const std::vector<TensorDof> &dofs, int ptab, const double xi[3], double *val, double *der) {
  ZVals Z[3];
  for (int a = 0; a < 3; a++) Z[a] = eval_z(ptab, xi[a]);
  for (size_t k = 0; k < dofs.size(); k++) {
    const TensorDof &d = dofs[k];
    const double s = d.sgn;
    if (space == ES_H1) {
      const double h0 = Z[0].H[d.idx[0]], h1 = Z[1].H[d.idx[1]], h2 = Z[2].H[d.idx[2]];
      val[k] = s * h0 * h1 * h2;
      der[3 * k] = s * Z[0].dH[d.idx[0]] * h1 * h2; der[3 * k + 1] = s * h0 * Z[1].dH[d.idx[1]] * h2; der[3 * k + 2] = s * h0 * h1 * Z[2].dH[d.idx[2]];
    } else if (space == ES_HCURL) {
      const int a = d.fam, b = (a + 1) % 3, c = (a + 2) % 3;
      const double q = Z[a].Q[d.idx[a]], hb = Z[b].H[d.idx[b]], hc = Z[c].H[d.idx[c]];
      const double f_b = q * Z[b].dH[d.idx[b]] * hc, f_c = q * hb * Z[c].dH[d.idx[c]];   // d f / d xi_b, d f / d xi_c, f = Q H H
      for (int i = 0; i < 3; i++) { val[3 * k + i] = 0.0; der[3 * k + i] = 0.0; }
      val[3 * k + a] = s * q * hb * hc;
      der[3 * k + b] = s * f_c;      // curl (f e_a): component a+1 = + d/d(a+2), component a+2 = - d/d(a+1)
      der[3 * k + c] = -s * f_b;
    } else {
      val[k] = s * Z[0].Q[d.idx[0]] * Z[1].Q[d.idx[1]] * Z[2].Q[d.idx[2]];
    }
  }
}

// prism analogue through the (triangle function) x (1-D table) decomposition (prism_space.hpp)
inline void prism_shape_at(int space, const std::vector<PrismDof> &d, const TriList &T0, const TriList &T1, int ptab, const double xi[3],
                           double *val, double *der) {
  const TriVals v0 = eval_list(T0, xi[0], xi[1]), v1 = eval_list(T1, xi[0], xi[1]);
  const ZVals Z = eval_z(ptab, xi[2]);
  for (size_t k = 0; k < d.size(); k++) {
    const PrismDof &q = d[k];
    const double *t = (q.list == 0 ? v0 : v1).at(q.t), s = q.sgn;
    if (space == ES_H1) {
      val[k] = s * t[0] * Z.H[q.zi];
      der[3 * k] = s * t[1] * Z.H[q.zi]; der[3 * k + 1] = s * t[2] * Z.H[q.zi]; der[3 * k + 2] = s * t[0] * Z.dH[q.zi];
    } else if (space == ES_HCURL) {
      double *V = val + 3 * k, *D = der + 3 * k;
      if (q.list == 0) { V[0] = s * t[0] * Z.H[q.zi]; V[1] = s * t[1] * Z.H[q.zi]; V[2] = 0.0;
        D[0] = -s * t[1] * Z.dH[q.zi]; D[1] = s * t[0] * Z.dH[q.zi]; D[2] = s * t[2] * Z.H[q.zi]; }
      else { V[0] = V[1] = 0.0; V[2] = s * t[0] * Z.Q[q.zi];
        D[0] = s * t[2] * Z.Q[q.zi]; D[1] = -s * t[1] * Z.Q[q.zi]; D[2] = 0.0; }
    } else {
      val[k] = s * t[0] * Z.Q[q.zi];
    }
  }
}

}  // namespace detail

// tabulate geometry + field shape functions of one signature at element_error's quadrature points
inline bool compile_error_signature(int kind, int maxp, int etype, const int *norder, const int *norie, const int *norif, ErrSigHost &S) {
  using namespace detail;
  S = ErrSigHost();
  S.etype = etype;
  S.space = kind <= 2 ? ES_H1 : (kind == 3 ? ES_HCURL : ES_L2);
  S.ncomp = kind == 4 ? 6 : 1;
  std::vector<double> xi;   // (3, nint)
  std::vector<double> val, der;
  const int INTEGRATION = 2;   // compute_error.F90:305
  if (etype == 1) {
    const HexaOrders o = HexaOrders::decode(norder);
    for (int e = 0; e < 12; e++) if (o.edge[e] < 1 || o.edge[e] > 9 || (norie[e] != 0 && norie[e] != 1)) { S.err = "bad edge order/orientation"; return false; }
    for (int f = 0; f < 6; f++) if (o.face[f][0] < 1 || o.face[f][1] < 1 || norif[f] < 0 || norif[f] > 7) { S.err = "bad face order/orientation"; return false; }
    for (int d = 0; d < 3; d++) if (o.mid[d] < 1) { S.err = "bad middle node order"; return false; }
    int pmax[3], nq[3];
    hexa_axis_max_order(norder, norif, pmax);
    Tables1D t[3];
    int ptab = 1;
    for (int d = 0; d < 3; d++) {
      nq[d] = std::min(pmax[d] + INTEGRATION, maxp) + 1;   // set_3D_int.F90:236-259
      if (nq[d] > MAXQ) { S.err = "order exceeds the 10-point Gauss table limit"; return false; }
      t[d] = make_tables(1, nq[d]);
      ptab = std::max(ptab, pmax[d]);
    }
    S.nint = nq[0] * nq[1] * nq[2];
    xi.resize(3 * (size_t)S.nint); S.w.resize(S.nint);
    for (int qz = 0, l = 0; qz < nq[2]; qz++)
      for (int qy = 0; qy < nq[1]; qy++)
        for (int qx = 0; qx < nq[0]; qx++, l++) {
          xi[3 * l] = t[0].x[qx]; xi[3 * l + 1] = t[1].x[qy]; xi[3 * l + 2] = t[2].x[qz];
          S.w[l] = t[0].w[qx] * t[1].w[qy] * t[2].w[qz];
        }
    const std::vector<TensorDof> hd = hexa_dofs_H1(norder, norie, norif);
    std::vector<TensorDof> fd;
    if (S.space == ES_HCURL) fd = hexa_dofs_Hcurl(norder, norie, norif);
    if (S.space == ES_L2) fd = hexa_dofs_L2(norder);
    S.nH = (int)hd.size(); S.nF = S.space == ES_H1 ? S.nH : (int)fd.size();
    S.tabH.assign((size_t)4 * S.nH * S.nint, 0.0);
    const int nfc = S.space == ES_HCURL ? 6 : 1;
    if (S.space != ES_H1) S.tabF.assign((size_t)nfc * S.nF * S.nint, 0.0);
    val.resize(3 * (size_t)std::max(S.nH, S.nF)); der.resize(val.size());
    for (int l = 0; l < S.nint; l++) {
      hexa_shape_at(ES_H1, hd, ptab, &xi[3 * l], val.data(), der.data());
      for (int k = 0; k < S.nH; k++) {
        S.tabH[((size_t)0 * S.nH + k) * S.nint + l] = val[k];
        for (int j = 0; j < 3; j++) S.tabH[((size_t)(1 + j) * S.nH + k) * S.nint + l] = der[3 * k + j];
      }
      if (S.space == ES_H1) continue;
      hexa_shape_at(S.space, fd, ptab, &xi[3 * l], val.data(), der.data());
      for (int k = 0; k < S.nF; k++) {
        if (S.space == ES_L2) S.tabF[(size_t)k * S.nint + l] = val[k];
        else for (int j = 0; j < 3; j++) {
          S.tabF[((size_t)j * S.nF + k) * S.nint + l] = val[3 * k + j];
          S.tabF[((size_t)(3 + j) * S.nF + k) * S.nint + l] = der[3 * k + j];
        }
      }
    }
    return true;
  }
  if (etype != 3) { S.err = "unknown element type (HP3D_MDLB = 1 and HP3D_MDLP = 3 are implemented)"; return false; }
  // ---- prism
  const PrismOrders o = PrismOrders::decode(norder);
  if (o.mid[0] < 1 || o.mid[1] < 1) { S.err = "bad prism middle node order"; return false; }
  int pmax[2];
  prism_axis_max_order(norder, norif, pmax);
  const int ordh = std::min(pmax[0] + INTEGRATION, maxp), ordz = std::min(pmax[1] + INTEGRATION, maxp);
  if (ordh > 9 || ordz + 1 > MAXQ || pmax[0] > TRI_MAXORD - 1) { S.err = "prism order exceeds the quadrature table limits"; return false; }
  const int nqt = TRI_RULE_NPTS[ordh - 1], nqz = ordz + 1;
  const double(*tpts)[3] = &TRI_RULE_PTS[TRI_RULE_OFF[ordh - 1]];
  const Tables1D tz = make_tables(1, nqz);
  S.nint = nqt * nqz;
  xi.resize(3 * (size_t)S.nint); S.w.resize(S.nint);
  for (int qz = 0, l = 0; qz < nqz; qz++)
    for (int qt = 0; qt < nqt; qt++, l++) { xi[3 * l] = tpts[qt][0]; xi[3 * l + 1] = tpts[qt][1]; xi[3 * l + 2] = tz.x[qz]; S.w[l] = tpts[qt][2] * tz.w[qz]; }
  TriList TG, F0, F1;
  const std::vector<PrismDof> hd = prism_dofs_H1(norder, norie, norif, TG);
  std::vector<PrismDof> fd;
  if (S.space == ES_HCURL) fd = prism_dofs_Hcurl(norder, norie, norif, F0, F1);
  if (S.space == ES_L2) fd = prism_dofs_L2(norder, F0);
  S.nH = (int)hd.size(); S.nF = S.space == ES_H1 ? S.nH : (int)fd.size();
  S.tabH.assign((size_t)4 * S.nH * S.nint, 0.0);
  const int nfc = S.space == ES_HCURL ? 6 : 1;
  if (S.space != ES_H1) S.tabF.assign((size_t)nfc * S.nF * S.nint, 0.0);
  val.resize(3 * (size_t)std::max(S.nH, S.nF)); der.resize(val.size());
  const int ptab = MAXN1D - 1;
  const TriList none;
  for (int l = 0; l < S.nint; l++) {
    prism_shape_at(ES_H1, hd, TG, none, ptab, &xi[3 * l], val.data(), der.data());
    for (int k = 0; k < S.nH; k++) {
      S.tabH[((size_t)0 * S.nH + k) * S.nint + l] = val[k];
      for (int j = 0; j < 3; j++) S.tabH[((size_t)(1 + j) * S.nH + k) * S.nint + l] = der[3 * k + j];
    }
    if (S.space == ES_H1) continue;
    prism_shape_at(S.space, fd, F0, F1, ptab, &xi[3 * l], val.data(), der.data());
    for (int k = 0; k < S.nF; k++) {
      if (S.space == ES_L2) S.tabF[(size_t)k * S.nint + l] = val[k];
      else for (int j = 0; j < 3; j++) {
        S.tabF[((size_t)j * S.nF + k) * S.nint + l] = val[3 * k + j];
        S.tabF[((size_t)(3 + j) * S.nF + k) * S.nint + l] = der[3 * k + j];
      }
    }
  }
  return true;
}

struct ErrArgs {
  const double *w, *tabH, *tabF;   // signature tables (device)
  int nint, nH, nF, space, ncomp;
  GeomParams gp;                   // problem kind + parameters of the built-in exact solution
  int nel;
  const double *xnod; long long xnod_ld;       // (3, nH) per element
  const double *zdof; long long szd;            // (ncomp, nF) per element, component fastest; stride in SCALARS
  const double *exact; long long sex;           // optional table: nvals scalars per point; stride in SCALARS (nullptr: built-in isol = 1)
  int l2proj;
  double *err, *rnorm; int *info;
  int want_points; double *xq; long long sxq;   // optional: physical coordinates of the points (3, nint)
};

// exact solution of the field variable (isol = 1), layout: H1 [u, grad u(3)] ; H(curl) [E(3), curl E(3)] ; L2 [E(3), H(3)]
__device__ inline void exact_field(const GeomParams &gp, const double x[3], double vr[6], double vi[6]) {
  double p, h[9], s[3], c[3];
  const double a = (gp.kind <= 2) ? 3.14159265358979323846 : gp.omega;
  for (int i = 0; i < 3; i++) dev_sincos(a * x[i], s[i], c[i]);
  p = s[0] * s[1] * s[2];
  const double g[3] = {a * c[0] * s[1] * s[2], a * c[1] * s[0] * s[2], a * c[2] * s[0] * s[1]};
  (void)h;
  for (int i = 0; i < 6; i++) { vr[i] = 0.0; vi[i] = 0.0; }
  if (gp.kind <= 2) { vr[0] = p; for (int j = 0; j < 3; j++) vr[1 + j] = g[j]; return; }
  const int ic = gp.icomp;
  // E = (1+i) p e_ic ; curl E = (1+i) grad p x e_ic
  double cr[3] = {0, 0, 0};
  const int b = (ic + 1) % 3, cc = (ic + 2) % 3;
  cr[b] = g[cc]; cr[cc] = -g[b];
  vr[ic] = p; vi[ic] = p;
  if (gp.kind == 3) { for (int i = 0; i < 3; i++) { vr[3 + i] = cr[i]; vi[3 + i] = cr[i]; } }
  else {   // H = curl E / (-i w mu) = (i / (w mu)) (1+i) cr = (-1 + i) cr / (w mu)
    const double f = 1.0 / (gp.omega * gp.mu);
    for (int i = 0; i < 3; i++) { vr[3 + i] = -f * cr[i]; vi[3 + i] = f * cr[i]; }
  }
}

template <bool CPLX>
__global__ void __launch_bounds__(256) elem_error_kernel(ErrArgs A) {
  constexpr int NS = CPLX ? 2 : 1;
  __shared__ double red[2][8];
  const int e = blockIdx.x, tid = threadIdx.x;
  const double *xn = A.xnod + (long long)e * A.xnod_ld;
  const double *zd = A.zdof ? A.zdof + (long long)e * A.szd * NS : nullptr;
  const long long HS = (long long)A.nH * A.nint, FS = (long long)A.nF * A.nint;
  double err = 0.0, nrm = 0.0;
  for (int l = tid; l < A.nint; l += blockDim.x) {
    double x[3] = {0, 0, 0}, J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // J[c + 3d] = dx_c/dxi_d
    for (int k = 0; k < A.nH; k++) {
      const double v = A.tabH[(long long)k * A.nint + l], d0 = A.tabH[HS + (long long)k * A.nint + l], d1 = A.tabH[2 * HS + (long long)k * A.nint + l],
                   d2 = A.tabH[3 * HS + (long long)k * A.nint + l];
#pragma unroll
      for (int c = 0; c < 3; c++) { const double xc = xn[3 * k + c]; x[c] += xc * v; J[c] += xc * d0; J[c + 3] += xc * d1; J[c + 6] += xc * d2; }
    }
    const double det = J[0] * J[4] * J[8] + J[1] * J[5] * J[6] + J[2] * J[3] * J[7] - J[2] * J[4] * J[6] - J[0] * J[5] * J[7] - J[1] * J[3] * J[8];
    if (!(det > 0.0)) A.info[e] = -1;
    double Ji[9];   // Ji[a + 3c] = dxi_a/dx_c
    Ji[0] = (J[4] * J[8] - J[5] * J[7]) / det; Ji[1] = (-J[1] * J[8] + J[2] * J[7]) / det; Ji[2] = (J[1] * J[5] - J[2] * J[4]) / det;
    Ji[3] = (J[5] * J[6] - J[3] * J[8]) / det; Ji[4] = (J[0] * J[8] - J[2] * J[6]) / det; Ji[5] = (-J[0] * J[5] + J[2] * J[3]) / det;
    Ji[6] = (J[3] * J[7] - J[4] * J[6]) / det; Ji[7] = (-J[0] * J[7] + J[1] * J[6]) / det; Ji[8] = (J[0] * J[4] - J[1] * J[3]) / det;
    if (A.want_points) for (int c = 0; c < 3; c++) A.xq[(long long)e * A.sxq + 3 * l + c] = x[c];
    if (!zd) continue;
    // ---- solution in master coordinates: up to 6 complex values
    double sr[6] = {0, 0, 0, 0, 0, 0}, si[6] = {0, 0, 0, 0, 0, 0};
    if (A.space == ES_H1) {          // [u, grad^(3)]
      for (int k = 0; k < A.nF; k++) {
        const double zr = zd[k * NS], zi = CPLX ? zd[k * NS + 1] : 0.0;
#pragma unroll
        for (int j = 0; j < 4; j++) { const double t = A.tabH[j * HS + (long long)k * A.nint + l]; sr[j] += zr * t; si[j] += zi * t; }
      }
    } else if (A.space == ES_HCURL) {  // [E^(3), curl^(3)]
      for (int k = 0; k < A.nF; k++) {
        const double zr = zd[k * NS], zi = CPLX ? zd[k * NS + 1] : 0.0;
#pragma unroll
        for (int j = 0; j < 6; j++) { const double t = A.tabF[j * FS + (long long)k * A.nint + l]; sr[j] += zr * t; si[j] += zi * t; }
      }
    } else {                            // L2: ncomp values
      for (int k = 0; k < A.nF; k++) {
        const double t = A.tabF[(long long)k * A.nint + l];
        for (int n = 0; n < A.ncomp; n++) { sr[n] += zd[(n + (long long)A.ncomp * k) * NS] * t; if (CPLX) si[n] += zd[(n + (long long)A.ncomp * k) * NS + 1] * t; }
      }
    }
    // ---- Piola maps (soleval.F90): gradient / H(curl) value by J^-T, curl by J/det, L2 by 1/det
    double ur[6], ui[6];
    if (A.space == ES_H1) {
      ur[0] = sr[0]; ui[0] = si[0];
      for (int j = 0; j < 3; j++) { double a = 0, b = 0; for (int i = 0; i < 3; i++) { a += sr[1 + i] * Ji[i + 3 * j]; b += si[1 + i] * Ji[i + 3 * j]; } ur[1 + j] = a; ui[1 + j] = b; }
    } else if (A.space == ES_HCURL) {
      for (int i = 0; i < 3; i++) {
        double a = 0, b = 0, c = 0, d = 0;
        for (int j = 0; j < 3; j++) { a += Ji[j + 3 * i] * sr[j]; b += Ji[j + 3 * i] * si[j]; c += J[i + 3 * j] * sr[3 + j]; d += J[i + 3 * j] * si[3 + j]; }
        ur[i] = a; ui[i] = b; ur[3 + i] = c / det; ui[3 + i] = d / det;
      }
    } else {
      for (int n = 0; n < 6; n++) { ur[n] = sr[n] / det; ui[n] = si[n] / det; }
    }
    // ---- exact solution
    double er[6], ei[6];
    const int nv = A.space == ES_H1 ? 4 : 6;
    if (A.exact) {
      const double *t = A.exact + ((long long)e * A.sex + (long long)l * nv) * NS;
      for (int j = 0; j < nv; j++) { er[j] = t[j * NS]; ei[j] = CPLX ? t[j * NS + 1] : 0.0; }
    } else exact_field(A.gp, x, er, ei);
    const double wt = A.w[l] * det;
    // values first / derivatives: the split only matters for l2proj
    const int v0 = A.space == ES_H1 ? 0 : 0, v1 = A.space == ES_H1 ? 1 : (A.space == ES_HCURL ? 3 : 6);   // value entries [v0, v1)
    for (int j = 0; j < nv; j++) {
      const bool is_val = j >= v0 && j < v1;
      if (!is_val && A.l2proj) continue;
      const double dr = er[j] - ur[j], di = ei[j] - ui[j];
      err += (dr * dr + di * di) * wt;
      nrm += (er[j] * er[j] + ei[j] * ei[j]) * wt;
    }
  }
  // block reduction (fixed order: deterministic)
  for (int o = 16; o; o >>= 1) { err += __shfl_xor_sync(0xffffffffu, err, o); nrm += __shfl_xor_sync(0xffffffffu, nrm, o); }
  if ((tid & 31) == 0) { red[0][tid >> 5] = err; red[1][tid >> 5] = nrm; }
  __syncthreads();
  if (tid == 0 && A.err) {
    double a = 0, b = 0;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); wv++) { a += red[0][wv]; b += red[1][wv]; }
    A.err[e] = a; A.rnorm[e] = b;
  }
}

}  // namespace hp3d
