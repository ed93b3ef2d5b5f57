// pbi.cuh -- batched H1 projection-based interpolation (interpolation half):
//   geometry dofs      update_gdof.F90:88-200,409-435 -> hpvert.F90:22, hpedge.F90:23, hpface_opt.F90:24, hpmdle_opt.F90:23
//   H1 Dirichlet dofs  update_Ddof.F90              -> dhpvert.F90:19, edge/dhpedgeH.F90:25, face/dhpfaceH_opt.F90:27
// One algorithm: the interpolated function g (the GMP map x(eta): 3 components, INTEGRATION = 0; or a Dirichlet datum
// u(x(eta)): NREQNH real / 2 NREQNH interleaved complex components, INTEGRATION = 1) enters through its vertex values and
// through its gradient in the reference coordinates eta of the GMP block, tabulated by the host at the points this file
// defines (the reference calls `hexa/prism(No,eta, x,dxdeta)` and `dirichlet` at the same points).  Node by node
// (vertices -> edges -> faces -> middle) the dofs of the node minimise the H1 seminorm in eta of
//        g - (interpolant of the lower-dimensional nodes) - sum_j dof_j phi_j
// over the node (tangential gradient on an edge, surface gradient on a face, full gradient inside), eta(xi) being the
// multilinear map through the element's vertex reference coordinates Etav (refgeom3D, geom3D.F90:235-305).
//
// Host: per signature (type, orders, orientations, INTEGRATION) the quadrature points of every node and the master gradients
// of all H1 functions at them are tabulated once.  Device: one CTA per (node, element):
//   A  a thread owns a point: Jacobian of eta(xi), tangent / normal, residual gradient R = dg/deta - sum_known dof_k grad phi_k,
//      projected test gradients; both scaled by sqrt(weight) and stored as rows of one matrix D = [test rows ; R rows]
//   B  G = D D_test^T  (64 x 64 register-tiled product: rows 0..n-1 the stiffness, rows n.. the load vectors) -- the reference's
//      DSFRK + load loop (hpface_opt.F90:203-237)
//   C  Cholesky of the stiffness with the load rows carried along (forward substitution for free), back substitution
//      (DPFTRF / DPFTRS, hpface_opt.F90:254-261; the edge routine's DGETRF solves the same SPD system).
#pragma once
#include "error_eval.cuh"

#include <algorithm>
#include <string>
#include <vector>

namespace hp3d {

constexpr int PBI_MAXNODE = 27;   // 8 + 12 + 6 + 1 nodes of a brick
constexpr int PBI_MAXCOMP = 12;

struct PbiNode { int kind, t0, n, nknown, p0, np, th0, nh; };   // kind 0 vertex, 1 edge, 2 face, 3 middle ; dofs [t0, t0+n) ; points [p0, p0+np)
                                                                // H(curl): [t0, t0+n) are E dofs, [th0, th0+nh) the node's H1 bubbles (multipliers)

enum PbiSpace { PBI_H1 = 0, PBI_HCURL = 1, PBI_HDIV = 2 };

struct PbiSigHost {
  int etype = 1, space = 0, nH = 0, nEF = 0, nrv = 8, nre = 12, nrf = 6, nnode = 27, npts = 0;   // nEF: H(curl) dofs of edges + faces
  PbiNode node[PBI_MAXNODE];
  std::vector<double> xi;     // (3, npts) master coordinates
  std::vector<double> wa;     // (npts) quadrature weights
  std::vector<double> tan;    // (6, npts) d xi / d t of the edge (first 3) or face (both) parametrisation
  std::vector<double> grad;   // [3][nH][npts] master gradients of the element's H1 functions
  std::vector<double> tabE;   // H(curl) only: [6][nEF][npts] master values (3) and curls (3) of the edge and face functions
  std::string err;
};

namespace detail {
static const int BR_EDGE_VERT[12][2] = {{1, 2}, {2, 3}, {4, 3}, {1, 4}, {5, 6}, {6, 7}, {8, 7}, {5, 8}, {1, 5}, {2, 6}, {3, 7}, {4, 8}};   // element_data.F90:67-70
static const int BR_FACE_VERT[6][4] = {{1, 2, 3, 4}, {5, 6, 7, 8}, {1, 2, 6, 5}, {2, 3, 7, 6}, {4, 3, 7, 8}, {1, 4, 8, 5}};                // :90-93
static const int BR_FACE_EDGE[6][4] = {{1, 2, 3, 4}, {5, 6, 7, 8}, {1, 10, 5, 9}, {2, 11, 6, 10}, {3, 11, 7, 12}, {4, 12, 8, 9}};
static const int PR_EDGE_VERT[9][2] = {{1, 2}, {2, 3}, {1, 3}, {4, 5}, {5, 6}, {4, 6}, {1, 4}, {2, 5}, {3, 6}};                              // :62-65
}  // namespace detail

// points + tables of one signature.  `tables = false` only fills the node descriptors and the points (size queries).
inline bool compile_pbi_signature(int etype, const int *norder, const int *norie, const int *norif, int integration, int maxp, bool tables,
                                  PbiSigHost &S, int space = PBI_H1) {
  using namespace detail;
  S = PbiSigHost();
  S.etype = etype; S.space = space;
  const bool brick = etype == 1;
  if (!brick && etype != 3) { S.err = "unknown element type (HP3D_MDLB = 1 and HP3D_MDLP = 3 are implemented)"; return false; }
  if (integration < 0 || integration > 2) { S.err = "INTEGRATION must be 0, 1 or 2"; return false; }
  S.nrv = brick ? 8 : 6; S.nre = brick ? 12 : 9; S.nrf = brick ? 6 : 5; S.nnode = S.nrv + S.nre + S.nrf + 1;
  const int nrv = S.nrv, nre = S.nre, nrf = S.nrf;
  for (int e = 0; e < nre; e++) if (norder[e] < 1 || norder[e] > 9 || (norie[e] != 0 && norie[e] != 1)) { S.err = "bad edge order/orientation"; return false; }
  for (int f = 0; f < nrf; f++) {
    const bool tri = !brick && f < 2;
    const int o = norder[nre + f];
    if (tri ? (o < 1 || o > 9 || norif[f] < 0 || norif[f] > 5) : (o / 10 < 1 || o % 10 < 1 || o > 99 || norif[f] < 0 || norif[f] > 7)) { S.err = "bad face order/orientation"; return false; }
  }
  // ---- dofs per node (ndof_nod, element_data.F90:808-870)
  int cnt[PBI_MAXNODE];
  for (int v = 0; v < nrv; v++) cnt[v] = 1;
  for (int e = 0; e < nre; e++) cnt[nrv + e] = norder[e] - 1;
  for (int f = 0; f < nrf; f++) {
    const int o = norder[nre + f];
    cnt[nrv + nre + f] = (!brick && f < 2) ? (o - 1) * (o - 2) / 2 : (o / 10 - 1) * (o % 10 - 1);
  }
  const int om = norder[nre + nrf];
  if (brick) { if (om / 100 < 1 || (om / 10) % 10 < 1 || om % 10 < 1) { S.err = "bad middle node order"; return false; }
    cnt[S.nnode - 1] = (om / 100 - 1) * ((om / 10) % 10 - 1) * (om % 10 - 1); }
  else { if (om / 10 < 1 || om % 10 < 1) { S.err = "bad prism middle node order"; return false; }
    cnt[S.nnode - 1] = (om / 10 - 1) * (om / 10 - 2) / 2 * (om % 10 - 1); }
  int off = 0;
  for (int i = 0; i < S.nnode; i++) {
    PbiNode &nd = S.node[i];
    nd.kind = i < nrv ? 0 : (i < nrv + nre ? 1 : (i < nrv + nre + nrf ? 2 : 3));
    nd.t0 = off; nd.n = cnt[i]; off += cnt[i];
    nd.p0 = 0; nd.np = 0; nd.nknown = 0; nd.th0 = 0; nd.nh = 0;
  }
  S.nH = off;
  for (int i = nrv; i < S.nnode; i++)
    S.node[i].nknown = S.node[i].kind == 1 ? nrv : (S.node[i].kind == 2 ? S.node[nrv + nre].t0 : S.node[S.nnode - 1].t0);
  if (space == PBI_HCURL) {   // E dofs of the edge and face nodes; the H1 bubbles of a face become its multipliers (dhpfaceE_opt.F90:358-372)
    int offE = 0;
    for (int i = 0; i < S.nnode; i++) {
      PbiNode &nd = S.node[i];
      nd.th0 = nd.t0; nd.nh = nd.kind == 2 ? nd.n : 0;
      int nE = 0;
      if (nd.kind == 1) nE = norder[i - nrv];
      else if (nd.kind == 2) {
        const int f = i - nrv - nre, o = norder[nre + f];
        nE = (!brick && f < 2) ? (o - 1) * o : (o / 10) * (o % 10 - 1) + (o / 10 - 1) * (o % 10);
      }
      nd.t0 = offE; nd.n = nE; offE += nE;
      nd.nknown = 0;
    }
    S.nEF = offE;
    for (int f = 0; f < nrf; f++) S.node[nrv + nre + f].nknown = S.node[nrv + nre].t0;   // all edge dofs
  }
  if (space == PBI_HDIV) {    // V dofs of the face nodes (dhpfaceV_opt.F90): no multipliers, nothing to subtract
    int offV = 0;
    for (int i = 0; i < S.nnode; i++) {
      PbiNode &nd = S.node[i];
      nd.th0 = 0; nd.nh = 0; nd.nknown = 0;
      int nV = 0;
      if (nd.kind == 2) {
        const int f = i - nrv - nre, o = norder[nre + f];
        nV = (!brick && f < 2) ? o * (o + 1) / 2 : (o / 10) * (o % 10);
      }
      nd.t0 = offV; nd.n = nV; offV += nV;
    }
    S.nEF = offV;
  }
  // ---- points
  auto cap = [&](int p) { return std::min(p + integration, maxp); };
  auto vert = [&](int v1) { return brick ? std::array<double, 3>{(double)VSIDE[v1 - 1][0], (double)VSIDE[v1 - 1][1], (double)VSIDE[v1 - 1][2]}
                                         : std::array<double, 3>{PR_COORD[v1 - 1][0], PR_COORD[v1 - 1][1], PR_COORD[v1 - 1][2]}; };
  auto push = [&](const double x[3], double w, const double t[6]) {
    for (int c = 0; c < 3; c++) S.xi.push_back(x[c]);
    S.wa.push_back(w);
    for (int c = 0; c < 6; c++) S.tan.push_back(t[c]);
  };
  for (int e = 0; e < nre; e++) {   // set_1Dint + edge_param (set_1D_int.F90:24-49, element_data.F90:508-549)
    PbiNode &nd = S.node[nrv + e];
    nd.p0 = (int)S.wa.size();
    if (nd.n <= 0) continue;
    const int nq = cap(norder[e]) + 1;
    if (nq > MAXQ) { S.err = "order exceeds the 10-point Gauss table limit"; return false; }
    const Tables1D g = make_tables(1, nq);
    const int *ev = brick ? BR_EDGE_VERT[e] : PR_EDGE_VERT[e];
    const std::array<double, 3> a = vert(ev[0]), b = vert(ev[1]);
    for (int l = 0; l < nq; l++) {
      double x[3], t[6] = {0, 0, 0, 0, 0, 0};
      for (int c = 0; c < 3; c++) { t[c] = b[c] - a[c]; x[c] = a[c] + g.x[l] * t[c]; }
      push(x, g.w[l], t);
    }
    nd.np = nq;
  }
  for (int f = 0; f < nrf; f++) {   // face_order + set_2Dint + face_param (element_data.F90:750-803,554-603, set_2D_int.F90:8-21,177-247)
    PbiNode &nd = S.node[nrv + nre + f];
    nd.p0 = (int)S.wa.size();
    if (nd.n <= 0) continue;
    const bool tri = !brick && f < 2;
    const int *fe = brick ? BR_FACE_EDGE[f] : PR_FACE_EDGE[f], *fv = brick ? BR_FACE_VERT[f] : PR_FACE_VERT[f];
    const std::array<double, 3> a = vert(fv[0]), b = vert(fv[1]), c3 = vert(fv[tri ? 2 : 3]);
    double t[6];
    for (int c = 0; c < 3; c++) { t[c] = b[c] - a[c]; t[3 + c] = c3[c] - a[c]; }
    std::vector<double> tp, tw;   // (2, n) face coordinates + weights
    if (tri) {
      int o = std::max(std::max(norder[fe[0] - 1], norder[fe[1] - 1]), std::max(norder[fe[2] - 1], norder[nre + f]));
      o = cap(o);
      if (o > 9) { S.err = "triangle rule order exceeds 9"; return false; }
      for (int l = 0; l < TRI_RULE_NPTS[o - 1]; l++) { const double *p = TRI_RULE_PTS[TRI_RULE_OFF[o - 1] + l]; tp.push_back(p[0]); tp.push_back(p[1]); tw.push_back(p[2]); }
    } else {
      const int h = norder[nre + f] / 10, v = norder[nre + f] % 10;
      const int nx = cap(std::max(std::max(norder[fe[0] - 1], norder[fe[2] - 1]), h)) + 1, ny = cap(std::max(std::max(norder[fe[1] - 1], norder[fe[3] - 1]), v)) + 1;
      if (nx > MAXQ || ny > MAXQ) { S.err = "order exceeds the 10-point Gauss table limit"; return false; }
      const Tables1D g1 = make_tables(1, nx), g2 = make_tables(1, ny);
      for (int l2 = 0; l2 < ny; l2++)
        for (int l1 = 0; l1 < nx; l1++) { tp.push_back(g1.x[l1]); tp.push_back(g2.x[l2]); tw.push_back(g1.w[l1] * g2.w[l2]); }
    }
    for (size_t l = 0; l < tw.size(); l++) {
      double x[3];
      for (int c = 0; c < 3; c++) x[c] = a[c] + tp[2 * l] * t[c] + tp[2 * l + 1] * t[3 + c];
      push(x, tw[l], t);
    }
    nd.np = (int)tw.size();
  }
  {   // set_3Dint with the orders as stored (hpmdle_opt.F90:119; set_3D_int.F90:155-259)
    PbiNode &nd = S.node[S.nnode - 1];
    nd.p0 = (int)S.wa.size();
    const int zero[6] = {0, 0, 0, 0, 0, 0};
    const double t0[6] = {0, 0, 0, 0, 0, 0};
    if (nd.n > 0 && space == PBI_H1) {
      if (brick) {
        int pmax[3], nq[3];
        hexa_axis_max_order(norder, zero, pmax);
        Tables1D g[3];
        for (int d = 0; d < 3; d++) { nq[d] = cap(pmax[d]) + 1; if (nq[d] > MAXQ) { S.err = "order exceeds the 10-point Gauss table limit"; return false; } g[d] = make_tables(1, nq[d]); }
        for (int qz = 0; qz < nq[2]; qz++)
          for (int qy = 0; qy < nq[1]; qy++)
            for (int qx = 0; qx < nq[0]; qx++) { const double x[3] = {g[0].x[qx], g[1].x[qy], g[2].x[qz]}; push(x, g[0].w[qx] * g[1].w[qy] * g[2].w[qz], t0); }
        nd.np = nq[0] * nq[1] * nq[2];
      } else {
        int pmax[2];
        prism_axis_max_order(norder, zero, pmax);
        const int oh = cap(pmax[0]), nz = cap(pmax[1]) + 1;
        if (oh > 9 || nz > MAXQ) { S.err = "prism order exceeds the quadrature table limits"; return false; }
        const Tables1D gz = make_tables(1, nz);
        const int nt = TRI_RULE_NPTS[oh - 1];
        for (int qz = 0; qz < nz; qz++)
          for (int qt = 0; qt < nt; qt++) { const double *p = TRI_RULE_PTS[TRI_RULE_OFF[oh - 1] + qt]; const double x[3] = {p[0], p[1], gz.x[qz]}; push(x, p[2] * gz.w[qz], t0); }
        nd.np = nt * nz;
      }
    }
  }
  S.npts = (int)S.wa.size();
  if (!tables) return true;
  // ---- master gradients of the element's H1 functions at all points (reference dof order)
  std::vector<double> val((size_t)3 * S.nH), der((size_t)3 * S.nH);
  S.grad.assign((size_t)3 * S.nH * S.npts, 0.0);
  if (brick) {
    const std::vector<TensorDof> hd = hexa_dofs_H1(norder, norie, norif);
    if ((int)hd.size() != S.nH) { S.err = "internal: H1 dof count mismatch"; return false; }
    int ptab = 1;
    for (int e = 0; e < 12; e++) ptab = std::max(ptab, norder[e]);
    for (int f = 0; f < 6; f++) ptab = std::max(ptab, std::max(norder[12 + f] / 10, norder[12 + f] % 10));
    ptab = std::max(ptab, std::max(om / 100, std::max((om / 10) % 10, om % 10)));
    for (int l = 0; l < S.npts; l++) {
      hexa_shape_at(ES_H1, hd, ptab, &S.xi[3 * l], val.data(), der.data());
      for (int k = 0; k < S.nH; k++)
        for (int j = 0; j < 3; j++) S.grad[((size_t)j * S.nH + k) * S.npts + l] = der[3 * k + j];
    }
    if (space == PBI_HDIV) {   // face functions: sign * H_normal(side) Q Q e_normal (values in planes 0..2, planes 3..5 unused)
      const std::vector<TensorDof> fd = hexa_dofs_Hdiv(norder, norif);
      if ((int)fd.size() < S.nEF) { S.err = "internal: H(div) dof count mismatch"; return false; }
      S.tabE.assign((size_t)6 * S.nEF * S.npts, 0.0);
      for (int l = 0; l < S.npts; l++) {
        ZVals Z[3];
        for (int a = 0; a < 3; a++) Z[a] = eval_z(ptab, S.xi[3 * l + a]);
        for (int k = 0; k < S.nEF; k++) {
          const TensorDof &d = fd[k];
          const int a = d.fam, b = (a + 1) % 3, c = (a + 2) % 3;
          S.tabE[((size_t)a * S.nEF + k) * S.npts + l] = d.sgn * Z[a].H[d.idx[a]] * Z[b].Q[d.idx[b]] * Z[c].Q[d.idx[c]];
        }
      }
    }
    if (space == PBI_HCURL) {
      const std::vector<TensorDof> fd = hexa_dofs_Hcurl(norder, norie, norif);
      if ((int)fd.size() < S.nEF) { S.err = "internal: H(curl) dof count mismatch"; return false; }
      std::vector<double> v3((size_t)3 * fd.size()), c3((size_t)3 * fd.size());
      S.tabE.assign((size_t)6 * S.nEF * S.npts, 0.0);
      for (int l = 0; l < S.npts; l++) {
        hexa_shape_at(ES_HCURL, fd, ptab, &S.xi[3 * l], v3.data(), c3.data());
        for (int k = 0; k < S.nEF; k++)
          for (int j = 0; j < 3; j++) { S.tabE[((size_t)j * S.nEF + k) * S.npts + l] = v3[3 * k + j]; S.tabE[((size_t)(3 + j) * S.nEF + k) * S.npts + l] = c3[3 * k + j]; }
      }
    }
  } else {
    TriList TG;
    const TriList none;
    const std::vector<PrismDof> hd = prism_dofs_H1(norder, norie, norif, TG);
    if ((int)hd.size() != S.nH) { S.err = "internal: H1 dof count mismatch"; return false; }
    for (int l = 0; l < S.npts; l++) {
      prism_shape_at(ES_H1, hd, TG, none, MAXN1D - 1, &S.xi[3 * l], val.data(), der.data());
      for (int k = 0; k < S.nH; k++)
        for (int j = 0; j < 3; j++) S.grad[((size_t)j * S.nH + k) * S.npts + l] = der[3 * k + j];
    }
    if (space == PBI_HDIV) {   // face functions only (hp3d_gpu_prism_shape, space 2)
      TriList TZ, TH;
      const std::vector<PrismDof> fd = prism_dofs_Hdiv_faces(norder, norif, TZ, TH);
      if ((int)fd.size() != S.nEF) { S.err = "internal: H(div) dof count mismatch"; return false; }
      S.tabE.assign((size_t)6 * S.nEF * S.npts, 0.0);
      for (int l = 0; l < S.npts; l++) {
        const TriVals v0 = eval_list(TZ, S.xi[3 * l], S.xi[3 * l + 1]), v1 = eval_list(TH, S.xi[3 * l], S.xi[3 * l + 1]);
        const ZVals Z = eval_z(MAXN1D - 1, S.xi[3 * l + 2]);
        for (int k = 0; k < S.nEF; k++) {
          const PrismDof &q = fd[k];
          const double *t = (q.list == 0 ? v0 : v1).at(q.t), sg = q.sgn;
          double V[3];
          if (q.list == 0) { V[0] = V[1] = 0.0; V[2] = sg * t[0] * Z.H[q.zi]; }
          else { V[0] = sg * t[1] * Z.Q[q.zi]; V[1] = -sg * t[0] * Z.Q[q.zi]; V[2] = 0.0; }
          for (int j = 0; j < 3; j++) S.tabE[((size_t)j * S.nEF + k) * S.npts + l] = V[j];
        }
      }
    }
    if (space == PBI_HCURL) {
      TriList F0, F1;
      const std::vector<PrismDof> fd = prism_dofs_Hcurl(norder, norie, norif, F0, F1);
      if ((int)fd.size() < S.nEF) { S.err = "internal: H(curl) dof count mismatch"; return false; }
      std::vector<double> v3((size_t)3 * fd.size()), c3((size_t)3 * fd.size());
      S.tabE.assign((size_t)6 * S.nEF * S.npts, 0.0);
      for (int l = 0; l < S.npts; l++) {
        prism_shape_at(ES_HCURL, fd, F0, F1, MAXN1D - 1, &S.xi[3 * l], v3.data(), c3.data());
        for (int k = 0; k < S.nEF; k++)
          for (int j = 0; j < 3; j++) { S.tabE[((size_t)j * S.nEF + k) * S.npts + l] = v3[3 * k + j]; S.tabE[((size_t)(3 + j) * S.nEF + k) * S.npts + l] = c3[3 * k + j]; }
      }
    }
  }
  return true;
}

struct PbiArgs {
  const double *wa, *tan, *grad;   // signature tables (device)
  const PbiNode *nodes;
  int nH, nrv, npts, ncomp;
  int node0;                       // this launch handles nodes node0 + blockIdx.x
  int nel;                         // elements of this signature group ...
  const int *elems;                // ... and their positions in the call's arrays
  const double *etav;              // (3, 8) per element
  const double *fgrad;             // (ncomp, 3, npts) per element, component fastest, stride fgrad_ld
  const double *fvert;             // (ncomp, 8) per element
  const unsigned *mask;            // per element (nullptr: every node)
  double *dof;                     // (ncomp, nH) per element, stride dof_ld
  long long fgrad_ld, dof_ld;
  double *ws; long long ws_stride; // workspace per CTA
  int g_in_smem;                   // large-node variant: the n x n system (+ load rows) lives in dynamic shared memory, only D in the workspace
  int *info;
};

// vertices: the value of g (hpvert.F90:22-58, dhpvert.F90:73-111)
__global__ void pbi_vertex_kernel(PbiArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.nel * A.nrv) return;
  const int e = A.elems[i / A.nrv], v = i % A.nrv;
  if (A.mask && !((A.mask[e] >> v) & 1u)) return;
  for (int c = 0; c < A.ncomp; c++) A.dof[(long long)e * A.dof_ld + (long long)v * A.ncomp + c] = A.fvert[((long long)e * 8 + v) * A.ncomp + c];
}

// MINB = resident CTAs per SM the register budget is cut for: 4 (64 registers) wins while the node systems are small and the kernel is
// latency-bound (p <= 5: +25..35 % measured), 2 (128 registers) for the GEMM-heavy middle nodes of higher orders (p = 7: 4 costs 9 %)
// SMALL: the node's D and G fit in shared memory (edges, faces up to p ~ 6): CTAs of 64 threads, up to 16 per SM, the products as plain
// per-entry dot products -- these launches are barrier-bound, not flop-bound, so fewer idle threads per barrier is what pays.
constexpr int PBI_SMALL_BYTES = 40 * 1024;
constexpr int PBI_GSMEM_BYTES = 36 * 1024;   // larger nodes: at least the system matrix (factorised under ~5 barriers per column) in shared memory
template <int MINB, bool SMALL>
__global__ void __launch_bounds__(SMALL ? 64 : 256, SMALL ? 16 : MINB) pbi_node_kernel(PbiArgs A) {
  extern __shared__ double pbi_dyn[];
  const int tid = threadIdx.x, inode = A.node0 + blockIdx.x;
  const PbiNode nd = A.nodes[inode];
  const int n = nd.n, np = nd.np, nc = A.ncomp, K3 = 3 * np, R = n + nc;
  if (n <= 0) return;
  double *D = SMALL ? pbi_dyn : A.ws + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * A.ws_stride;   // [R][K3]
  double *G = (!SMALL && A.g_in_smem) ? pbi_dyn : D + (long long)R * K3;                                  // [R][n]
  const long long HS = (long long)A.nH * A.npts;
  for (int ie = blockIdx.y; ie < A.nel; ie += gridDim.y) {
    const int e = A.elems[ie];
    if (A.mask && !((A.mask[e] >> inode) & 1u)) continue;
    const double *ev = A.etav + (long long)e * 24;
    double *dof = A.dof + (long long)e * A.dof_ld;
    const double *fg = A.fgrad + (long long)e * A.fgrad_ld;
    __syncthreads();   // the previous element's solve has finished with D / G
    // ---- A: one thread per point
    // a thread owns (point l, slice sl): with fewer points than threads the rows of D are dealt out over nsl slices per point (each
    // slice recomputes the point's geometry: 8 vertex gradients)
    const int nsl = max(1, (int)blockDim.x / np);
    for (int item = tid; item < np * nsl; item += blockDim.x) {
      const int l = item % np, sl = item / np;
      const int gl = nd.p0 + l;
      double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // J[c + 3m] = d eta_c / d xi_m
      for (int v = 0; v < A.nrv; v++) {
#pragma unroll
        for (int m = 0; m < 3; m++) {
          const double g = A.grad[m * HS + (long long)v * A.npts + gl];
#pragma unroll
          for (int c = 0; c < 3; c++) J[c + 3 * m] += ev[3 * v + c] * g;
        }
      }
      const double det = J[0] * J[4] * J[8] + J[1] * J[5] * J[6] + J[2] * J[3] * J[7] - J[2] * J[4] * J[6] - J[0] * J[5] * J[7] - J[1] * J[3] * J[8];
      if (!(det > 0.0)) A.info[e] = -1;
      double Ji[9];   // Ji[a + 3i] = d xi_a / d eta_i (geom.F90:57-113)
      Ji[0] = (J[4] * J[8] - J[5] * J[7]) / det; Ji[1] = (-J[1] * J[8] + J[2] * J[7]) / det; Ji[2] = (J[1] * J[5] - J[2] * J[4]) / det;
      Ji[3] = (J[5] * J[6] - J[3] * J[8]) / det; Ji[4] = (J[0] * J[8] - J[2] * J[6]) / det; Ji[5] = (-J[0] * J[5] + J[2] * J[3]) / det;
      Ji[6] = (J[3] * J[7] - J[4] * J[6]) / det; Ji[7] = (-J[0] * J[7] + J[1] * J[6]) / det; Ji[8] = (J[0] * J[4] - J[1] * J[3]) / det;
      double weight, dir[3] = {0, 0, 0};   // unit tangent (edge) / unit normal (face)
      if (nd.kind == 3) weight = A.wa[gl] * det;
      else {
        const double *t = A.tan + 6LL * gl;
        double d1[3], d2[3];
#pragma unroll
        for (int c = 0; c < 3; c++) { d1[c] = J[c] * t[0] + J[c + 3] * t[1] + J[c + 6] * t[2]; d2[c] = J[c] * t[3] + J[c + 3] * t[4] + J[c + 6] * t[5]; }
        if (nd.kind == 1) { dir[0] = d1[0]; dir[1] = d1[1]; dir[2] = d1[2]; }   // hpedge.F90:140-146
        else { dir[0] = d1[1] * d2[2] - d1[2] * d2[1]; dir[1] = d1[2] * d2[0] - d1[0] * d2[2]; dir[2] = d1[0] * d2[1] - d1[1] * d2[0]; }   // brefgeom3D
        const double bj = sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
        dir[0] /= bj; dir[1] /= bj; dir[2] /= bj;
        weight = A.wa[gl] * bj;
      }
      const double sw = sqrt(weight);
      // residual rows, one component at a time: R_c = d g_c / d eta minus the known part, accumulated in registers (the table
      // entries are re-read per component from L1; independent loads, so the loop pipelines)
      for (int c = sl; c < nc; c += nsl) {
        double r0 = fg[(long long)gl * 3 * nc + c], r1 = fg[(long long)gl * 3 * nc + nc + c], r2 = fg[(long long)gl * 3 * nc + 2 * nc + c];
#pragma unroll 4
        for (int k = 0; k < nd.nknown; k++) {
          const double g0 = A.grad[(long long)k * A.npts + gl], g1 = A.grad[HS + (long long)k * A.npts + gl], g2 = A.grad[2 * HS + (long long)k * A.npts + gl];
          const double z = dof[(long long)k * nc + c];
          r0 -= z * (g0 * Ji[0] + g1 * Ji[1] + g2 * Ji[2]); r1 -= z * (g0 * Ji[3] + g1 * Ji[4] + g2 * Ji[5]); r2 -= z * (g0 * Ji[6] + g1 * Ji[7] + g2 * Ji[8]);
        }
        double *Dr = D + (long long)(n + c) * K3 + 3 * l;
        Dr[0] = r0 * sw; Dr[1] = r1 * sw; Dr[2] = r2 * sw;
      }
#pragma unroll 2
      for (int j = sl; j < n; j += nsl) {
        const int k = nd.t0 + j;
        const double g0 = A.grad[(long long)k * A.npts + gl], g1 = A.grad[HS + (long long)k * A.npts + gl], g2 = A.grad[2 * HS + (long long)k * A.npts + gl];
        double dv[3] = {g0 * Ji[0] + g1 * Ji[1] + g2 * Ji[2], g0 * Ji[3] + g1 * Ji[4] + g2 * Ji[5], g0 * Ji[6] + g1 * Ji[7] + g2 * Ji[8]};
        const double pr = dv[0] * dir[0] + dv[1] * dir[1] + dv[2] * dir[2];
        if (nd.kind == 1) { dv[0] = pr * dir[0]; dv[1] = pr * dir[1]; dv[2] = pr * dir[2]; }           // hpedge.F90:188-189
        else if (nd.kind == 2) { dv[0] -= pr * dir[0]; dv[1] -= pr * dir[1]; dv[2] -= pr * dir[2]; }    // hpface_opt.F90:213-214
        for (int i = 0; i < 3; i++) D[(long long)j * K3 + 3 * l + i] = dv[i] * sw;
      }
    }
    __syncthreads();
    // ---- B: G[r][j] = sum_k D[r][k] D[j][k], r < R, j < n, tiles with j-tile <= r-tile (lower triangle + load rows)
    if constexpr (SMALL) {
      for (int q = tid; q < R * n; q += blockDim.x) {
        const int r = q / n, j = q % n;
        if (j > r) continue;
        const double *a = D + r * K3, *b = D + j * K3;
        double acc = 0.0;
        for (int k = 0; k < K3; k++) acc += a[k] * b[k];
        G[r * n + j] = acc;
      }
    } else {
      __shared__ double As[16][68], Bs[16][68];
      const int tx = tid & 15, ty = tid >> 4;
      for (int r0 = 0; r0 < R; r0 += 64)
        for (int j0 = 0; j0 <= r0 && j0 < n; j0 += 64) {
          double acc[4][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
          for (int k0 = 0; k0 < K3; k0 += 16) {
            {   // stage 64 rows x 16 k of both operands: thread -> (row = tid / 4, 4 consecutive k)
              const int row = tid >> 2, kk = (tid & 3) * 4;
  #pragma unroll
              for (int q = 0; q < 4; q++) {
                const int k = k0 + kk + q;
                As[kk + q][row] = (r0 + row < R && k < K3) ? D[(long long)(r0 + row) * K3 + k] : 0.0;
                Bs[kk + q][row] = (j0 + row < n && k < K3) ? D[(long long)(j0 + row) * K3 + k] : 0.0;
              }
            }
            __syncthreads();
  #pragma unroll
            for (int k = 0; k < 16; k++) {
              double a[4], b[4];
  #pragma unroll
              for (int q = 0; q < 4; q++) { a[q] = As[k][ty * 4 + q]; b[q] = Bs[k][tx * 4 + q]; }
  #pragma unroll
              for (int p = 0; p < 4; p++)
  #pragma unroll
                for (int q = 0; q < 4; q++) acc[p][q] += a[p] * b[q];
            }
            __syncthreads();
          }
  #pragma unroll
          for (int p = 0; p < 4; p++)
  #pragma unroll
            for (int q = 0; q < 4; q++) {
              const int r = r0 + ty * 4 + p, j = j0 + tx * 4 + q;
              if (r < R && j < n) G[(long long)r * n + j] = acc[p][q];
            }
        }
    }
    __syncthreads();
    // ---- C: right-looking Cholesky G = L L^T on the lower triangle; the load rows r >= n ride along (they end as y^T, L y = b)
    bool bad = false;
    for (int k = 0; k < n; k++) {
      const double d = G[(long long)k * n + k];
      if (!(d > 0.0)) { bad = true; break; }   // uniform: every thread reads the same value
      const double piv = sqrt(d);
      __syncthreads();
      for (int r = k + tid; r < R; r += blockDim.x) G[(long long)r * n + k] = (r == k) ? piv : G[(long long)r * n + k] / piv;
      __syncthreads();
      // trailing update: rows r > k, columns k < j <= min(r, n-1); a warp takes a row, lanes run along j
      for (int r = k + 1 + (tid >> 5); r < R; r += (blockDim.x >> 5)) {
        const double lrk = G[(long long)r * n + k];
        const int jmax = r < n ? r : n - 1;
        for (int j = k + 1 + (tid & 31); j <= jmax; j += 32) G[(long long)r * n + j] -= lrk * G[(long long)j * n + k];
      }
      __syncthreads();
    }
    if (bad) { if (tid == 0) A.info[e] = inode + 1; continue; }   // LAPACK-style: the node whose stiffness is not positive definite
    // back substitution L^T x = y, column oriented: x_k = y_k / L_kk, then y_j -= L_kj x_k for j < k
    for (int k = n - 1; k >= 0; k--) {
      const double lkk = G[(long long)k * n + k];
      __syncthreads();
      if (tid < nc) G[(long long)(n + tid) * n + k] /= lkk;
      __syncthreads();
      for (int q = tid; q < nc * k; q += blockDim.x) {
        const int c = q / k, j = q % k;
        G[(long long)(n + c) * n + j] -= G[(long long)k * n + j] * G[(long long)(n + c) * n + k];
      }
    }
    __syncthreads();
    for (int q = tid; q < nc * n; q += blockDim.x) { const int j = q / nc, c = q % nc; dof[(long long)(nd.t0 + j) * nc + c] = G[(long long)(n + c) * n + j]; }
  }
}

// ---- H(curl) Dirichlet dofs: edge/dhpedgeE.F90:24-391, face/dhpfaceE_opt.F90:26-549 -------------------------------------------------
// The datum enters pulled back to eta at the points of the signature (INTEGRATION = 1): E_eta = dxdeta^T E and
// curl_eta = det(dxdeta) dxdeta^-1 curl E (dhpfaceE_opt.F90:269-279).  An edge projects the tangential component in L2; a face
// minimises the normal component of curl_eta (E_eta - known edges - sum dof_j E_j) subject to orthogonality of the tangential
// residual to the surface gradients of the face's H1 bubbles: the saddle-point system [C B; B^T 0] of :395-414, solved by LU with
// partial pivoting like the reference's DGETRF (:439).
struct PbiEArgs {
  const double *wa, *tan, *grad, *tabE;
  const PbiNode *nodes;
  int nH, nEF, nrv, npts, ncomp, node0, nel, space;   // space: PBI_HCURL or PBI_HDIV (faces only, L2 projection of the normal component)
  const int *elems;
  const double *etav, *fval, *fcurl;   // fval / fcurl: (ncomp, 3, npts) per element, component fastest, stride f_ld
  const unsigned *mask;
  double *dof;                         // (ncomp, nEF) per element, stride dof_ld
  long long f_ld, dof_ld;
  double *ws; long long ws_stride;
  int g_in_smem;                       // large-node variant: W in dynamic shared memory
  int *info;
};

// out[(orow0 + r) * ldo + ocol0 + j] = sum_k D[(xrow0 + r) * K3 + k] * D[(yrow0 + j) * K3 + k], r < nx, j < ny (whole CTA, 256 threads)
__device__ inline void pbi_rows_product(const double *D, int K3, int xrow0, int nx, int yrow0, int ny, double *out, int ldo, int orow0, int ocol0,
                                        double (*As)[68], double (*Bs)[68]) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  for (int r0 = 0; r0 < nx; r0 += 64)
    for (int j0 = 0; j0 < ny; j0 += 64) {
      double acc[4][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
      for (int k0 = 0; k0 < K3; k0 += 16) {
        const int row = tid >> 2, kk = (tid & 3) * 4;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int k = k0 + kk + q;
          As[kk + q][row] = (r0 + row < nx && k < K3) ? D[(long long)(xrow0 + r0 + row) * K3 + k] : 0.0;
          Bs[kk + q][row] = (j0 + row < ny && k < K3) ? D[(long long)(yrow0 + j0 + row) * K3 + k] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; k++) {
          double a[4], b[4];
#pragma unroll
          for (int q = 0; q < 4; q++) { a[q] = As[k][ty * 4 + q]; b[q] = Bs[k][tx * 4 + q]; }
#pragma unroll
          for (int p = 0; p < 4; p++)
#pragma unroll
            for (int q = 0; q < 4; q++) acc[p][q] += a[p] * b[q];
        }
        __syncthreads();
      }
#pragma unroll
      for (int p = 0; p < 4; p++)
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int r = r0 + ty * 4 + p, j = j0 + tx * 4 + q;
          if (r < nx && j < ny) out[(long long)(orow0 + r) * ldo + ocol0 + j] = acc[p][q];
        }
    }
}

// out[(orow0 + r) * ldo + ocol0 + j] as above by per-entry dot products (small nodes whose D lives in shared memory)
__device__ inline void pbi_rows_product_small(const double *D, int K3, int xrow0, int nx, int yrow0, int ny, double *out, int ldo, int orow0, int ocol0) {
  for (int q = threadIdx.x; q < nx * ny; q += blockDim.x) {
    const int r = q / ny, j = q % ny;
    const double *a = D + (xrow0 + r) * K3, *b = D + (yrow0 + j) * K3;
    double acc = 0.0;
    for (int k = 0; k < K3; k++) acc += a[k] * b[k];
    out[(orow0 + r) * ldo + ocol0 + j] = acc;
  }
}

template <bool SMALL>
__global__ void __launch_bounds__(SMALL ? 64 : 256, SMALL ? 6 : 2) pbi_hcurl_kernel(PbiEArgs A) {
  extern __shared__ double pbi_dyn[];
  __shared__ double s_red[8];
  __shared__ int s_idx[8], s_piv;
  const int tid = threadIdx.x, inode = A.node0 + blockIdx.x;
  const PbiNode nd = A.nodes[inode];
  const int nE = nd.n, nHb = nd.nh, np = nd.np, nc = A.ncomp, K3 = 3 * np, nt = nE + nHb;
  if (nE <= 0) return;
  const bool hdiv = A.space == PBI_HDIV;
  const bool face = nd.kind == 2 && !hdiv;   // the saddle-point path of dhpfaceE_opt
  // rows of D: [CE (nE, faces only) | E (nE) | GH (nHb) | Rc (nc, faces only) | Rv (nc)]
  const int rE = face ? nE : 0, rG = rE + nE, rRc = rG + nHb, rRv = rRc + (face ? nc : 0), nrows = rRv + nc;
  double *D = SMALL ? pbi_dyn : A.ws + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * A.ws_stride;
  double *W = (!SMALL && A.g_in_smem) ? pbi_dyn : D + (long long)nrows * K3;   // [(nt + nc)][nt]: rows < nt the (symmetric) system, rows >= nt the load vectors
  const long long HS = (long long)A.nH * A.npts, ES = (long long)A.nEF * A.npts;
  for (int ie = blockIdx.y; ie < A.nel; ie += gridDim.y) {
    const int e = A.elems[ie];
    if (A.mask && !((A.mask[e] >> inode) & 1u)) continue;
    const double *ev = A.etav + (long long)e * 24;
    double *dof = A.dof + (long long)e * A.dof_ld;
    const double *fv = A.fval + (long long)e * A.f_ld, *fc = A.fcurl + (long long)e * A.f_ld;
    __syncthreads();
    // a thread owns (point l, slice sl): with fewer points than threads the rows of D are dealt out over nsl slices per point (each
    // slice recomputes the point's geometry: 8 vertex gradients)
    const int nsl = max(1, (int)blockDim.x / np);
    for (int item = tid; item < np * nsl; item += blockDim.x) {
      const int l = item % np, sl = item / np;
      const int gl = nd.p0 + l;
      double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int v = 0; v < A.nrv; v++) {
#pragma unroll
        for (int m = 0; m < 3; m++) {
          const double g = A.grad[m * HS + (long long)v * A.npts + gl];
#pragma unroll
          for (int c = 0; c < 3; c++) J[c + 3 * m] += ev[3 * v + c] * g;
        }
      }
      const double det = J[0] * J[4] * J[8] + J[1] * J[5] * J[6] + J[2] * J[3] * J[7] - J[2] * J[4] * J[6] - J[0] * J[5] * J[7] - J[1] * J[3] * J[8];
      if (!(det > 0.0)) A.info[e] = -1;
      double Ji[9];
      Ji[0] = (J[4] * J[8] - J[5] * J[7]) / det; Ji[1] = (-J[1] * J[8] + J[2] * J[7]) / det; Ji[2] = (J[1] * J[5] - J[2] * J[4]) / det;
      Ji[3] = (J[5] * J[6] - J[3] * J[8]) / det; Ji[4] = (J[0] * J[8] - J[2] * J[6]) / det; Ji[5] = (-J[0] * J[5] + J[2] * J[3]) / det;
      Ji[6] = (J[3] * J[7] - J[4] * J[6]) / det; Ji[7] = (-J[0] * J[7] + J[1] * J[6]) / det; Ji[8] = (J[0] * J[4] - J[1] * J[3]) / det;
      const double *t = A.tan + 6LL * gl;
      double d1[3], d2[3], dir[3];
#pragma unroll
      for (int c = 0; c < 3; c++) { d1[c] = J[c] * t[0] + J[c + 3] * t[1] + J[c + 6] * t[2]; d2[c] = J[c] * t[3] + J[c + 3] * t[4] + J[c + 6] * t[5]; }
      if (nd.kind != 2) { dir[0] = d1[0]; dir[1] = d1[1]; dir[2] = d1[2]; }
      else { dir[0] = d1[1] * d2[2] - d1[2] * d2[1]; dir[1] = d1[2] * d2[0] - d1[0] * d2[2]; dir[2] = d1[0] * d2[1] - d1[1] * d2[0]; }
      const double bj = sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
      dir[0] /= bj; dir[1] /= bj; dir[2] /= bj;
      const double sw = sqrt(A.wa[gl] * bj);
      // master -> eta: value u = Ji^T E^ (u_i = sum_a E^_a Ji[a + 3i]), curl cu = J C^ / det.  Residual rows one component at a time with
      // register accumulators (the edges' contributions removed, dhpfaceE_opt.F90:295-310); the table entries are re-read from L1
      for (int c = sl; c < nc; c += nsl) {
        const long long fo = (long long)gl * 3 * nc + c;
        double v0 = fv[fo], v1 = fv[fo + nc], v2 = fv[fo + 2 * nc], c0r = 0.0, c1r = 0.0, c2r = 0.0;
        if (face) { c0r = fc[fo]; c1r = fc[fo + nc]; c2r = fc[fo + 2 * nc]; }
#pragma unroll 2
        for (int k = 0; k < nd.nknown; k++) {
          const double e0 = A.tabE[(long long)k * A.npts + gl], e1 = A.tabE[ES + (long long)k * A.npts + gl], e2 = A.tabE[2 * ES + (long long)k * A.npts + gl];
          const double q0 = A.tabE[3 * ES + (long long)k * A.npts + gl], q1 = A.tabE[4 * ES + (long long)k * A.npts + gl], q2 = A.tabE[5 * ES + (long long)k * A.npts + gl];
          const double z = dof[(long long)k * nc + c], zd = z / det;
          v0 -= z * (e0 * Ji[0] + e1 * Ji[1] + e2 * Ji[2]); v1 -= z * (e0 * Ji[3] + e1 * Ji[4] + e2 * Ji[5]); v2 -= z * (e0 * Ji[6] + e1 * Ji[7] + e2 * Ji[8]);
          c0r -= zd * (J[0] * q0 + J[3] * q1 + J[6] * q2); c1r -= zd * (J[1] * q0 + J[4] * q1 + J[7] * q2); c2r -= zd * (J[2] * q0 + J[5] * q1 + J[8] * q2);
        }
        double *Dv = D + (long long)(rRv + c) * K3 + 3 * l;
        Dv[0] = v0 * sw; Dv[1] = v1 * sw; Dv[2] = v2 * sw;
        if (face) { double *Dc = D + (long long)(rRc + c) * K3 + 3 * l; Dc[0] = c0r * sw; Dc[1] = c1r * sw; Dc[2] = c2r * sw; }
      }
      for (int j = sl; j < nE; j += nsl) {
        const int k = nd.t0 + j;
        const double e0 = A.tabE[(long long)k * A.npts + gl], e1 = A.tabE[ES + (long long)k * A.npts + gl], e2 = A.tabE[2 * ES + (long long)k * A.npts + gl];
        double v[3] = {e0 * Ji[0] + e1 * Ji[1] + e2 * Ji[2], e0 * Ji[3] + e1 * Ji[4] + e2 * Ji[5], e0 * Ji[6] + e1 * Ji[7] + e2 * Ji[8]};
        if (hdiv) { v[0] = (J[0] * e0 + J[3] * e1 + J[6] * e2) / det; v[1] = (J[1] * e0 + J[4] * e1 + J[7] * e2) / det; v[2] = (J[2] * e0 + J[5] * e1 + J[8] * e2) / det; }   // dhpfaceV_opt.F90:251-253
        const double pr = v[0] * dir[0] + v[1] * dir[1] + v[2] * dir[2];
        if (!face) { v[0] = pr * dir[0]; v[1] = pr * dir[1]; v[2] = pr * dir[2]; }            // dhpedgeE.F90:223-224, dhpfaceV_opt.F90:256-257
        else { v[0] -= pr * dir[0]; v[1] -= pr * dir[1]; v[2] -= pr * dir[2]; }                 // dhpfaceE_opt.F90:337-338
        for (int i = 0; i < 3; i++) D[(long long)(rE + j) * K3 + 3 * l + i] = v[i] * sw;
        if (face) {
          const double c0 = A.tabE[3 * ES + (long long)k * A.npts + gl], c1 = A.tabE[4 * ES + (long long)k * A.npts + gl], c2 = A.tabE[5 * ES + (long long)k * A.npts + gl];
          const double cv[3] = {(J[0] * c0 + J[3] * c1 + J[6] * c2) / det, (J[1] * c0 + J[4] * c1 + J[7] * c2) / det, (J[2] * c0 + J[5] * c1 + J[8] * c2) / det};
          const double pc = cv[0] * dir[0] + cv[1] * dir[1] + cv[2] * dir[2];
          for (int i = 0; i < 3; i++) D[(long long)j * K3 + 3 * l + i] = pc * dir[i] * sw;     // :341-342
        }
      }
      for (int j = sl; j < nHb; j += nsl) {   // surface gradients of the face's H1 bubbles (:358-372)
        const int k = nd.th0 + j;
        const double g0 = A.grad[(long long)k * A.npts + gl], g1 = A.grad[HS + (long long)k * A.npts + gl], g2 = A.grad[2 * HS + (long long)k * A.npts + gl];
        double dv[3] = {g0 * Ji[0] + g1 * Ji[1] + g2 * Ji[2], g0 * Ji[3] + g1 * Ji[4] + g2 * Ji[5], g0 * Ji[6] + g1 * Ji[7] + g2 * Ji[8]};
        const double pr = dv[0] * dir[0] + dv[1] * dir[1] + dv[2] * dir[2];
        for (int i = 0; i < 3; i++) D[(long long)(rG + j) * K3 + 3 * l + i] = (dv[i] - pr * dir[i]) * sw;
      }
    }
    __syncthreads();
    // ---- the system and its load vectors
    auto product = [&](int xrow0, int nx, int yrow0, int ny, int orow0, int ocol0) {
      if constexpr (SMALL) pbi_rows_product_small(D, K3, xrow0, nx, yrow0, ny, W, nt, orow0, ocol0);
      else { __shared__ double As[16][68], Bs[16][68]; pbi_rows_product(D, K3, xrow0, nx, yrow0, ny, W, nt, orow0, ocol0, As, Bs); }
    };
    if (!face) {
      product(0, nE, 0, nE, 0, 0);             // mass matrix of the tangential (normal) component
      product(rRv, nc, 0, nE, nt, 0);
    } else {
      product(0, nE, 0, nE, 0, 0);             // curl-curl (DSYRK, :395)
      product(rRc, nc, 0, nE, nt, 0);
      if (nHb > 0) {
        product(rG, nHb, rE, nE, nE, 0);       // B^T (DGEMM, :399)
        product(rRv, nc, rG, nHb, nt, nE);
        __syncthreads();
        for (int q = tid; q < nHb * nE; q += blockDim.x) { const int j = q / nE, i = q % nE; W[(long long)i * nt + nE + j] = W[(long long)(nE + j) * nt + i]; }
        for (int q = tid; q < nHb * nHb; q += blockDim.x) W[(long long)(nE + q / nHb) * nt + nE + q % nHb] = 0.0;
      }
    }
    __syncthreads();
    // ---- LU with partial pivoting of M (nt x nt) with the load vectors as extra columns.  Memory W[r * nt + i] is read as the
    // column-major matrix Wc(i, r) = [M | b_1 .. b_nc] (M is symmetric), so that columns are contiguous.
    const int ncol = nt + nc;
    bool bad = false;
    for (int k = 0; k < nt; k++) {
      // pivot search in column k, rows i >= k
      double best = -1.0; int bi = k;
      for (int i = k + tid; i < nt; i += blockDim.x) { const double a = fabs(W[(long long)k * nt + i]); if (a > best) { best = a; bi = i; } }
      for (int o = 16; o; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if ((tid & 31) == 0) { s_red[tid >> 5] = best; s_idx[tid >> 5] = bi; }
      __syncthreads();
      if (tid == 0) {
        double b = s_red[0]; int ix = s_idx[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) if (s_red[w] > b || (s_red[w] == b && s_idx[w] < ix)) { b = s_red[w]; ix = s_idx[w]; }
        s_piv = (b > 0.0) ? ix : -1;
      }
      __syncthreads();
      const int p = s_piv;
      if (p < 0) { bad = true; break; }
      if (p != k) for (int r = tid; r < ncol; r += blockDim.x) { const double a = W[(long long)r * nt + k]; W[(long long)r * nt + k] = W[(long long)r * nt + p]; W[(long long)r * nt + p] = a; }
      __syncthreads();
      const double piv = W[(long long)k * nt + k];
      __syncthreads();
      for (int i = k + 1 + tid; i < nt; i += blockDim.x) W[(long long)k * nt + i] /= piv;   // multipliers
      __syncthreads();
      for (int r = k + 1 + (tid >> 5); r < ncol; r += (blockDim.x >> 5)) {                    // columns r > k, lanes along the rows
        const double ukr = W[(long long)r * nt + k];
        for (int i = k + 1 + (tid & 31); i < nt; i += 32) W[(long long)r * nt + i] -= W[(long long)k * nt + i] * ukr;
      }
      __syncthreads();
    }
    if (bad) { if (tid == 0) A.info[e] = inode + 1; continue; }
    // back substitution with U (upper triangle of Wc) for the nc load columns, column oriented
    for (int k = nt - 1; k >= 0; k--) {
      const double ukk = W[(long long)k * nt + k];
      __syncthreads();
      if (tid < nc) W[(long long)(nt + tid) * nt + k] /= ukk;
      __syncthreads();
      for (int q = tid; q < nc * k; q += blockDim.x) {
        const int c = q / k, i = q % k;
        W[(long long)(nt + c) * nt + i] -= W[(long long)k * nt + i] * W[(long long)(nt + c) * nt + k];
      }
    }
    __syncthreads();
    for (int q = tid; q < nc * nE; q += blockDim.x) { const int j = q / nc, c = q % nc; dof[(long long)(nd.t0 + j) * nc + c] = W[(long long)(nt + c) * nt + j]; }
  }
}

}  // namespace hp3d
