/*
 * pbi.c -- CPU restatement ("oracle") of hp3D's H1 projection-based interpolation (interpolation half):
 *   geometry dofs (update_gdof)         trunk/src/hpinterp/hpvert.F90:22, hpedge.F90:23, hpface_opt.F90:24, hpmdle_opt.F90:23
 *   H1 Dirichlet dofs (update_Ddof)     trunk/src/hpinterp/dhpvert.F90:19, edge/dhpedgeH.F90:25, face/dhpfaceH_opt.F90:27
 * The two families are the same algorithm: the interpolated function g (the GMP map x(eta), 3 components, or the
 * Dirichlet datum u(x(eta)), NREQNH components) is given by a callback returning its value and its gradient with respect
 * to the REFERENCE coordinates eta of the GMP block (dxdeta, resp. zdvalH * dxdeta, dhpfaceH_opt.F90:217-224); the
 * projections are done in eta (hpmdle_opt.F90:6-7), eta(xi) being the trilinear / linear-prism map through the element's
 * vertex reference coordinates Etav (refgeom3D, trunk/src/element/util/geom3D.F90:235-305).  The geometry routines
 * integrate with INTEGRATION = 0, the Dirichlet routines with INTEGRATION = 1 (dhpedgeH.F90:145, dhpfaceH_opt.F90:173).
 * TEST INFRASTRUCTURE ONLY (see hp3d_oracle.h).
 *
 * Pins (tests/test_pbi_oracle.py): the reference's test-suite holds no numeric vectors for src/hpinterp; what
 * trunk/test/poly_pois.F90 asserts through update_gdof + update_Ddof is polynomial reproduction, pinned here directly:
 * a function of the element's polynomial space is reproduced exactly (dofs evaluated back through shape3DH), a trilinear
 * map has zero higher-order dofs, and neighbours sharing a face obtain identical dofs for the shared entities.  Independently of the
 * oracle's quadrature tables, assembly and solvers, the interpolants satisfy their variational definition (Galerkin orthogonality node by
 * node in the H1 seminorm; for H(curl) faces both block rows of the saddle-point system) under numpy's own 12-point Gauss-Legendre rule.
 */
#include "dense.h"
#include "hp3d_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static const double BRICK_COORD[8][3] = {{0,0,0},{1,0,0},{1,1,0},{0,1,0},{0,0,1},{1,0,1},{1,1,1},{0,1,1}};
static const int BRICK_EDGE_TO_VERT[12][2] = {{1,2},{2,3},{4,3},{1,4},{5,6},{6,7},{8,7},{5,8},{1,5},{2,6},{3,7},{4,8}};
static const double PRISM_COORD[6][3] = {{0,0,0},{1,0,0},{0,1,0},{0,0,1},{1,0,1},{0,1,1}};
static const int PRISM_EDGE_TO_VERT[9][2] = {{1,2},{2,3},{1,3},{4,5},{5,6},{4,6},{1,4},{2,5},{3,6}};

/* element_data.F90:508-549 ; ie is 1-based */
void orc_edge_param(int et, int ie, double t, double xi[3], double dxidt[3]) {
  const double *x1, *x2;
  if (et == ORC_MDLP) { x1 = PRISM_COORD[PRISM_EDGE_TO_VERT[ie - 1][0] - 1]; x2 = PRISM_COORD[PRISM_EDGE_TO_VERT[ie - 1][1] - 1]; }
  else { x1 = BRICK_COORD[BRICK_EDGE_TO_VERT[ie - 1][0] - 1]; x2 = BRICK_COORD[BRICK_EDGE_TO_VERT[ie - 1][1] - 1]; }
  for (int c = 0; c < 3; c++) { dxidt[c] = x2[c] - x1[c]; xi[c] = x1[c] + t * dxidt[c]; }
}

/* refgeom3D, geom3D.F90:235-305: eta = sum_v Etav_v phi_v over the nrv VERTEX functions */
static void refgeom3D(const double *etav, const double *shapH, const double *gradH, int nrv, double eta[3], double detadxi[9],
                      double dxideta[9], double *rjac, int *iflag) {
  for (int c = 0; c < 3; c++) eta[c] = 0.0;
  for (int i = 0; i < 9; i++) detadxi[i] = 0.0;
  for (int k = 0; k < nrv; k++)
    for (int c = 0; c < 3; c++) {
      eta[c] += etav[c + 3 * k] * shapH[k];
      for (int m = 0; m < 3; m++) detadxi[c + 3 * m] += etav[c + 3 * k] * gradH[m + 3 * k];
    }
  orc_geom(detadxi, dxideta, rjac, iflag);
}

/* offsets of the nodes' H1 dofs inside the element's dof list (vertices, edges, faces, middle): celndof.F90:24 */
void orc_pbi_offsets(int et, const int *norder, int *off /*nrv+nre+nrf+1 entries + total*/) {
  const int nrv = orc_nvert(et), nre = orc_nedge(et), nrf = orc_nface(et);
  int n = 0, k = 0, h, e, v, q;
  for (int i = 0; i < nrv; i++) { off[k++] = n; n += 1; }
  for (int i = 0; i < nre; i++) { off[k++] = n; n += norder[i] - 1; }
  for (int i = 0; i < nrf; i++) { off[k++] = n; orc_ndof_nod_face(et, i + 1, norder[nre + i], &h, &e, &v, &q); n += h; }
  off[k++] = n; orc_ndof_nod_mid(et, norder[nre + nrf], &h, &e, &v, &q); n += h;
  off[k] = n;
}

/* One node of the element.  node: 0..nrv-1 vertices, then edges, faces, middle (the element's nodesl order).
 * dof: (ncomp, nrdofH) component fastest, full-element numbering; the lower-dimensional nodes' entries are read, the node's
 * own entries are written. */
int orc_pbi_node(int et, const int *norder, const int *norie, const int *norif, const double *etav, int ncomp, int integration,
                 int maxp, int node, orc_pbi_fn f, void *ctx, double *dof) {
  const int nrv = orc_nvert(et), nre = orc_nedge(et), nrf = orc_nface(et);
  int off[32];
  orc_pbi_offsets(et, norder, off);
  const int n = off[node + 1] - off[node], t0 = off[node];
  double *val = malloc(sizeof(double) * ncomp), *dval = malloc(sizeof(double) * 3 * ncomp);
  if (node < nrv) { /* hpvert.F90:22-58, dhpvert.F90:73-111 */
    f(etav + 3 * node, val, dval, ctx);
    for (int c = 0; c < ncomp; c++) dof[c + ncomp * t0] = val[c];
    free(val); free(dval);
    return 0;
  }
  if (n <= 0) { free(val); free(dval); return 0; }
  double *shapH = malloc(sizeof(double) * ORC_MAXBRICK_H), *gradH = malloc(sizeof(double) * 3 * ORC_MAXBRICK_H);
  double *aa = calloc((size_t)n * n, sizeof(double)), *bb = calloc((size_t)n * ncomp, sizeof(double));
  int nord1[19], zero6[6] = {0, 0, 0, 0, 0, 0}, info = 0;
  int nint = 0;
  double *pts = malloc(sizeof(double) * 3 * 1000), *wts = malloc(sizeof(double) * 1000);
  double *atest = NULL;
  const int kind = node < nrv + nre ? 1 : (node < nrv + nre + nrf ? 2 : 3);
  int nknown; /* dofs of the lower-dimensional nodes that are subtracted */
  if (kind == 1) { /* hpedge.F90:113-117 */
    const int ie = node - nrv + 1;
    orc_initiate_order(et, nord1);
    nord1[ie - 1] = norder[ie - 1];
    int na = norder[ie - 1] + integration; if (na > maxp) na = maxp;   /* set_1D_int.F90:41-43 */
    nint = na + 1;
    orc_gauss1(nint, pts, wts);
    nknown = nrv;
  } else if (kind == 2) { /* hpface_opt.F90:129-146 */
    const int jf = node - nrv - nre + 1;
    int nordf[5];
    orc_initiate_order(et, nord1);
    for (int i = 0; i < nre; i++) nord1[i] = norder[i];
    nord1[nre + jf - 1] = norder[nre + jf - 1];
    orc_face_order(et, jf, norder, nordf);
    nint = orc_set_2D_int(orc_face_is_tri(et, jf), nordf, 0, integration, maxp, pts, wts);
    nknown = off[nrv + nre];
  } else { /* hpmdle_opt.F90:119 */
    for (int i = 0; i < nre + nrf + 1; i++) nord1[i] = norder[i];
    nint = orc_set_3D_int(et, norder, zero6, integration, maxp, pts, wts);
    nknown = off[nrv + nre + nrf];
  }
  if (kind >= 2) atest = malloc(sizeof(double) * (size_t)n * 3 * nint);
  for (int l = 0; l < nint; l++) {
    double xi[3], dxidt[6], eta[3], detadxi[9], dxideta[9], rjac, weight = 0, rt[3] = {0, 0, 0}, rn[3] = {0, 0, 0};
    int iflag;
    if (kind == 1) orc_edge_param(et, node - nrv + 1, pts[l], xi, dxidt);
    else if (kind == 2) orc_face_param(et, node - nrv - nre + 1, pts + 2 * l, xi, dxidt);
    else for (int c = 0; c < 3; c++) xi[c] = pts[3 * l + c];
    int nrdofH = orc_shape3DH(et, xi, nord1, norie, norif, shapH, gradH);
    refgeom3D(etav, shapH, gradH, nrv, eta, detadxi, dxideta, &rjac, &iflag);
    if (iflag != 0) info = -1;
    if (kind == 1) { /* hpedge.F90:140-146 */
      double bjac = 0;
      for (int c = 0; c < 3; c++) { rt[c] = detadxi[c] * dxidt[0] + detadxi[c + 3] * dxidt[1] + detadxi[c + 6] * dxidt[2]; bjac += rt[c] * rt[c]; }
      bjac = sqrt(bjac);
      for (int c = 0; c < 3; c++) rt[c] /= bjac;
      weight = wts[l] * bjac;
    } else if (kind == 2) { /* brefgeom3D, geom3D.F90:343-393 */
      double d[6], bjac;
      for (int i = 0; i < 2; i++)
        for (int c = 0; c < 3; c++) d[c + 3 * i] = detadxi[c] * dxidt[3 * i] + detadxi[c + 3] * dxidt[3 * i + 1] + detadxi[c + 6] * dxidt[3 * i + 2];
      rn[0] = d[1] * d[5] - d[2] * d[4]; rn[1] = d[2] * d[3] - d[0] * d[5]; rn[2] = d[0] * d[4] - d[1] * d[3];
      bjac = sqrt(rn[0] * rn[0] + rn[1] * rn[1] + rn[2] * rn[2]);
      const int ns = orc_nsign_param(et, node - nrv - nre + 1);
      for (int c = 0; c < 3; c++) rn[c] = rn[c] * ns / bjac;
      weight = wts[l] * bjac;
    } else weight = wts[l] * rjac;
    f(eta, val, dval, ctx); /* dval(c, i) = d g_c / d eta_i, c fastest */
    /* remove the lower-dimensional nodes' contributions (hpedge.F90:164-176, hpface_opt.F90:186-198, hpmdle_opt.F90:156-167);
     * in the reduced-order element the known functions keep their positions, the node's own functions come last */
    const int own0 = nrdofH - n;
    for (int k = 0; k < own0; k++) {
      double du[3];
      for (int i = 0; i < 3; i++) du[i] = gradH[3 * k] * dxideta[3 * i] + gradH[1 + 3 * k] * dxideta[1 + 3 * i] + gradH[2 + 3 * k] * dxideta[2 + 3 * i];
      for (int i = 0; i < 3; i++)
        for (int c = 0; c < ncomp; c++) dval[c + ncomp * i] -= dof[c + ncomp * k] * du[i];
    }
    if (own0 != nknown) info = -2;
    for (int j = 0; j < n; j++) {
      const int kj = own0 + j;
      double dv[3], prod = 0;
      for (int i = 0; i < 3; i++) dv[i] = gradH[3 * kj] * dxideta[3 * i] + gradH[1 + 3 * kj] * dxideta[1 + 3 * i] + gradH[2 + 3 * kj] * dxideta[2 + 3 * i];
      if (kind == 1) { for (int i = 0; i < 3; i++) prod += dv[i] * rt[i]; for (int i = 0; i < 3; i++) dv[i] = prod * rt[i]; }
      if (kind == 2) { for (int i = 0; i < 3; i++) prod += dv[i] * rn[i]; for (int i = 0; i < 3; i++) dv[i] -= prod * rn[i]; }
      for (int c = 0; c < ncomp; c++)
        bb[j + n * c] += (dval[c] * dv[0] + dval[c + ncomp] * dv[1] + dval[c + 2 * ncomp] * dv[2]) * weight;
      if (kind == 1) { /* hpedge.F90:199-213: full gradient of the trial function against the projected test gradient */
        for (int i = 0; i < n; i++) {
          const int ki = own0 + i;
          double du[3];
          for (int m = 0; m < 3; m++) du[m] = gradH[3 * ki] * dxideta[3 * m] + gradH[1 + 3 * ki] * dxideta[1 + 3 * m] + gradH[2 + 3 * ki] * dxideta[2 + 3 * m];
          aa[j + n * i] += (dv[0] * du[0] + dv[1] * du[1] + dv[2] * du[2]) * weight;
        }
      } else { /* hpface_opt.F90:223, hpmdle_opt.F90:189 */
        const double sw = sqrt(weight);
        for (int i = 0; i < 3; i++) atest[j + (size_t)n * (3 * l + i)] = dv[i] * sw;
      }
    }
  }
  if (kind == 1) { /* dgetrf + dlaswp + 2 x dtrsm, hpedge.F90:240-258 */
    int *ipiv = malloc(sizeof(int) * n);
    if (orc_dgetrf(n, aa, n, ipiv) != 0) info = 1;
    else orc_dgetrs(n, ncomp, aa, n, ipiv, bb, n);
    free(ipiv);
  } else { /* DSFRK + DPFTRF + DPFTRS (hpface_opt.F90:234-261): the RFP storage is a layout, the algebra is syrk + potrf + potrs */
    orc_dsyrk_u('N', n, 3 * nint, 1.0, atest, n, 0.0, aa, n);
    if (orc_dpotrf_u(n, aa, n) != 0) info = 1;
    else { orc_dtrsm_u('T', n, ncomp, aa, n, bb, n); orc_dtrsm_u('N', n, ncomp, aa, n, bb, n); }
  }
  if (info == 0)
    for (int j = 0; j < n; j++)
      for (int c = 0; c < ncomp; c++) dof[c + ncomp * (t0 + j)] = bb[j + n * c];
  free(val); free(dval); free(shapH); free(gradH); free(aa); free(bb); free(pts); free(wts); free(atest);
  return info;
}

/* The element's nodes in the order update_gdof / update_Ddof visit them (vertices, then edges, then faces, then the middle
 * node: update_gdof.F90 loops by node type in that order so that each projection finds the lower-dimensional dofs ready).
 * mask bit i selects node i; unselected nodes keep the entries of `dof` they came with. */
int orc_pbi_element(int et, const int *norder, const int *norie, const int *norif, const double *etav, int ncomp, int integration,
                    int maxp, unsigned mask, orc_pbi_fn f, void *ctx, double *dof) {
  const int nn = orc_nvert(et) + orc_nedge(et) + orc_nface(et) + 1;
  for (int node = 0; node < nn; node++)
    if (mask & (1u << node)) {
      int rc = orc_pbi_node(et, norder, norie, norif, etav, ncomp, integration, maxp, node, f, ctx, dof);
      if (rc) return rc;
    }
  return 0;
}

/* a fixed smooth 3-component map for timing the restatement without a Python callback (bench of tests/bench_pbi.py) */
void orc_pbi_sample_fn(const double *eta, double *val, double *dval, void *ctx) {
  const double x = eta[0], y = eta[1], z = eta[2];
  (void)ctx;
  val[0] = x + 0.1 * sin(2.0 * y) * z; val[1] = y + 0.05 * x * x; val[2] = z + 0.1 * cos(x + y);
  dval[0] = 1.0;                dval[3] = 0.2 * cos(2.0 * y) * z; dval[6] = 0.1 * sin(2.0 * y);
  dval[1] = 0.1 * x;            dval[4] = 1.0;                    dval[7] = 0.0;
  dval[2] = -0.1 * sin(x + y);  dval[5] = -0.1 * sin(x + y);      dval[8] = 1.0;
}
/* element loop shaped like update_gdof.F90:409-435 (OpenMP over elements); dof stride = 3*nrdofH of each element's own order */
int orc_pbi_batch_sample(int nel, const int *etype, const int *norder, const int *norie, const int *norif, const double *etav,
                         int integration, int maxp, double *dof, long dof_ld, int nthreads) {
  int bad = 0;
#pragma omp parallel for schedule(dynamic) num_threads(nthreads) reduction(+ : bad)
  for (int e = 0; e < nel; e++)
    bad += orc_pbi_element(etype ? etype[e] : ORC_MDLB, norder + 19 * e, norie + 12 * e, norif + 6 * e, etav + 24 * e, 3, integration, maxp,
                           0x7ffffffu, orc_pbi_sample_fn, NULL, dof + dof_ld * e) != 0;
  return bad;
}

/* ------------------------------------------------------------------------------------------------------------------------
 * H(curl) Dirichlet dofs: edge/dhpedgeE.F90:24-391, face/dhpfaceE_opt.F90:26-549 (INTEGRATION = 1).
 * f(eta, E[ncomp*3], curlE[ncomp*3], dxdeta[9], ctx): the datum in PHYSICAL components (E[c + ncomp*j] = E_c,j, what `dirichlet`
 * returns as zvalE and the curl formed from zdvalE, dhpfaceE_opt.F90:249-251) and the GMP Jacobian dxdeta(j,i) = dxdeta[j + 3i];
 * the pullbacks to eta are done here as in the reference.  Nodes: 0..nre-1 edges, then faces.
 * dofE: (ncomp, nrdofE) component fastest, full-element numbering. */
void orc_pbi_offsets_E(int et, const int *norder, int *off /* nre+nrf entries + total of edges+faces */) {
  const int nre = orc_nedge(et), nrf = orc_nface(et);
  int n = 0, k = 0, h, e, v, q;
  for (int i = 0; i < nre; i++) { off[k++] = n; n += norder[i]; }
  for (int i = 0; i < nrf; i++) { off[k++] = n; orc_ndof_nod_face(et, i + 1, norder[nre + i], &h, &e, &v, &q); n += e; }
  off[k] = n;
}

int orc_pbi_hcurl_node(int et, const int *norder, const int *norie, const int *norif, const double *etav, int ncomp, int maxp, int node,
                       orc_pbi_fnE f, void *ctx, double *dofE) {
  const int nrv = orc_nvert(et), nre = orc_nedge(et), nrf = orc_nface(et), integration = 1;
  int offE[32], offH[32];
  orc_pbi_offsets_E(et, norder, offE);
  orc_pbi_offsets(et, norder, offH);
  const int nE = offE[node + 1] - offE[node], t0 = offE[node];
  if (nE <= 0) return 0;
  const int isface = node >= nre;
  const int nH = isface ? offH[nrv + node + 1] - offH[nrv + node] : 0;   /* H1 bubbles of the face: Lagrange multipliers */
  const int nt = nE + nH;
  double *shapH = malloc(sizeof(double) * ORC_MAXBRICK_H), *gradH = malloc(sizeof(double) * 3 * ORC_MAXBRICK_H);
  double *shapE = malloc(sizeof(double) * 3 * ORC_MAXBRICK_E), *curlE = malloc(sizeof(double) * 3 * ORC_MAXBRICK_E);
  double *aa = calloc((size_t)nt * nt, sizeof(double)), *bb = calloc((size_t)nt * ncomp, sizeof(double));
  double *val = malloc(sizeof(double) * 3 * ncomp), *crl = malloc(sizeof(double) * 3 * ncomp);
  double *veta = malloc(sizeof(double) * 3 * ncomp), *ceta = malloc(sizeof(double) * 3 * ncomp);
  double *pts = malloc(sizeof(double) * 3 * 1000), *wts = malloc(sizeof(double) * 1000);
  int nord1[19], nint, info = 0;
  double *a_e = NULL, *a_ce = NULL, *a_gh = NULL;
  orc_initiate_order(et, nord1);
  if (!isface) { /* dhpedgeE.F90:147-153 */
    nord1[node] = norder[node];
    int na = norder[node] + integration; if (na > maxp) na = maxp;
    nint = na + 1;
    orc_gauss1(nint, pts, wts);
  } else { /* dhpfaceE_opt.F90:168-186 */
    const int jf = node - nre + 1;
    int nordf[5];
    for (int i = 0; i < nre; i++) nord1[i] = norder[i];
    nord1[nre + jf - 1] = norder[nre + jf - 1];
    orc_face_order(et, jf, norder, nordf);
    nint = orc_set_2D_int(orc_face_is_tri(et, jf), nordf, 0, integration, maxp, pts, wts);
    a_e = malloc(sizeof(double) * (size_t)nE * 3 * nint); a_ce = malloc(sizeof(double) * (size_t)nE * 3 * nint);
    a_gh = malloc(sizeof(double) * (size_t)(nH > 0 ? nH : 1) * 3 * nint);
  }
  for (int l = 0; l < nint; l++) {
    double xi[3], dxidt[6], eta[3], detadxi[9], dxideta[9], rjac, weight, dir[3], dxdeta[9], detadx[9], rjx;
    int iflag;
    if (!isface) orc_edge_param(et, node + 1, pts[l], xi, dxidt);
    else orc_face_param(et, node - nre + 1, pts + 2 * l, xi, dxidt);
    int nrdofH = orc_shape3DH(et, xi, nord1, norie, norif, shapH, gradH);
    int nrdofE = orc_shape3DE(et, xi, nord1, norie, norif, shapE, curlE);
    refgeom3D(etav, shapH, gradH, nrv, eta, detadxi, dxideta, &rjac, &iflag);
    if (iflag != 0) info = -1;
    if (!isface) {
      double bjac = 0;
      for (int c = 0; c < 3; c++) { dir[c] = detadxi[c] * dxidt[0] + detadxi[c + 3] * dxidt[1] + detadxi[c + 6] * dxidt[2]; bjac += dir[c] * dir[c]; }
      bjac = sqrt(bjac);
      for (int c = 0; c < 3; c++) dir[c] /= bjac;
      weight = wts[l] * bjac;
    } else {
      double d[6], bjac;
      for (int i = 0; i < 2; i++)
        for (int c = 0; c < 3; c++) d[c + 3 * i] = detadxi[c] * dxidt[3 * i] + detadxi[c + 3] * dxidt[3 * i + 1] + detadxi[c + 6] * dxidt[3 * i + 2];
      dir[0] = d[1] * d[5] - d[2] * d[4]; dir[1] = d[2] * d[3] - d[0] * d[5]; dir[2] = d[0] * d[4] - d[1] * d[3];
      bjac = sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
      const int ns = orc_nsign_param(et, node - nre + 1);
      for (int c = 0; c < 3; c++) dir[c] = dir[c] * ns / bjac;
      weight = wts[l] * bjac;
    }
    f(eta, val, crl, dxdeta, ctx);
    orc_geom(dxdeta, detadx, &rjx, &iflag);   /* dhpfaceE_opt.F90:238 */
    /* pullbacks (dhpedgeE.F90:204-211, dhpfaceE_opt.F90:269-279): E_eta = dxdeta^T E, curl_eta = rjac * detadx * curl */
    for (int c = 0; c < ncomp; c++)
      for (int i = 0; i < 3; i++) {
        double a = 0, b = 0;
        for (int j = 0; j < 3; j++) { a += val[c + ncomp * j] * dxdeta[j + 3 * i]; b += crl[c + ncomp * j] * detadx[i + 3 * j] * rjx; }
        veta[c + ncomp * i] = a; ceta[c + ncomp * i] = b;
      }
    if (!isface) { /* dhpedgeE.F90:214-243 ; the edge's functions sit after one function of each preceding edge */
      for (int j = 0; j < nE; j++) {
        const int kj = node + j;
        double v[3], prod = 0;
        for (int i = 0; i < 3; i++) v[i] = shapE[3 * kj] * dxideta[3 * i] + shapE[1 + 3 * kj] * dxideta[1 + 3 * i] + shapE[2 + 3 * kj] * dxideta[2 + 3 * i];
        for (int i = 0; i < 3; i++) prod += v[i] * dir[i];
        for (int i = 0; i < 3; i++) v[i] = prod * dir[i];
        for (int c = 0; c < ncomp; c++) bb[j + nt * c] += (veta[c] * v[0] + veta[c + ncomp] * v[1] + veta[c + 2 * ncomp] * v[2]) * weight;
        for (int i2 = 0; i2 < nE; i2++) {
          const int ki = node + i2;
          double u[3];
          for (int i = 0; i < 3; i++) u[i] = shapE[3 * ki] * dxideta[3 * i] + shapE[1 + 3 * ki] * dxideta[1 + 3 * i] + shapE[2 + 3 * ki] * dxideta[2 + 3 * i];
          aa[j + nt * i2] += (v[0] * u[0] + v[1] * u[1] + v[2] * u[2]) * weight;
        }
      }
      continue;
    }
    /* face: remove the edges' contributions (dhpfaceE_opt.F90:295-310) */
    const int ownE = nrdofE - nE, ownH = nrdofH - nH;
    if (ownE != offE[nre]) info = -2;
    for (int k = 0; k < ownE; k++) {
      double u[3], cu[3];
      for (int i = 0; i < 3; i++) {
        u[i] = shapE[3 * k] * dxideta[3 * i] + shapE[1 + 3 * k] * dxideta[1 + 3 * i] + shapE[2 + 3 * k] * dxideta[2 + 3 * i];
        cu[i] = (detadxi[i] * curlE[3 * k] + detadxi[i + 3] * curlE[1 + 3 * k] + detadxi[i + 6] * curlE[2 + 3 * k]) / rjac;
      }
      for (int i = 0; i < 3; i++)
        for (int c = 0; c < ncomp; c++) { veta[c + ncomp * i] -= dofE[c + ncomp * k] * u[i]; ceta[c + ncomp * i] -= dofE[c + ncomp * k] * cu[i]; }
    }
    const double sw = sqrt(weight);
    for (int j = 0; j < nE; j++) { /* :325-356 */
      const int kj = ownE + j;
      double v[3], cv[3], prod = 0;
      for (int i = 0; i < 3; i++) {
        v[i] = shapE[3 * kj] * dxideta[3 * i] + shapE[1 + 3 * kj] * dxideta[1 + 3 * i] + shapE[2 + 3 * kj] * dxideta[2 + 3 * i];
        cv[i] = (detadxi[i] * curlE[3 * kj] + detadxi[i + 3] * curlE[1 + 3 * kj] + detadxi[i + 6] * curlE[2 + 3 * kj]) / rjac;
      }
      for (int i = 0; i < 3; i++) prod += v[i] * dir[i];
      for (int i = 0; i < 3; i++) v[i] -= prod * dir[i];
      prod = 0;
      for (int i = 0; i < 3; i++) prod += cv[i] * dir[i];
      for (int i = 0; i < 3; i++) cv[i] = prod * dir[i];
      for (int c = 0; c < ncomp; c++) bb[j + nt * c] += (cv[0] * ceta[c] + cv[1] * ceta[c + ncomp] + cv[2] * ceta[c + 2 * ncomp]) * weight;
      for (int i = 0; i < 3; i++) { a_e[j + (size_t)nE * (3 * l + i)] = v[i] * sw; a_ce[j + (size_t)nE * (3 * l + i)] = cv[i] * sw; }
    }
    for (int j = 0; j < nH; j++) { /* :358-372 */
      const int kj = ownH + j;
      double dv[3], prod = 0;
      for (int i = 0; i < 3; i++) dv[i] = gradH[3 * kj] * dxideta[3 * i] + gradH[1 + 3 * kj] * dxideta[1 + 3 * i] + gradH[2 + 3 * kj] * dxideta[2 + 3 * i];
      for (int i = 0; i < 3; i++) prod += dv[i] * dir[i];
      for (int i = 0; i < 3; i++) dv[i] -= prod * dir[i];
      for (int c = 0; c < ncomp; c++) bb[nE + j + nt * c] += (veta[c] * dv[0] + veta[c + ncomp] * dv[1] + veta[c + 2 * ncomp] * dv[2]) * weight;
      for (int i = 0; i < 3; i++) a_gh[j + (size_t)nH * (3 * l + i)] = dv[i] * sw;
    }
  }
  if (isface) { /* DSYRK + DGEMM + symmetric fill (:395-414): [curl-curl, E.grad ; (E.grad)^T, 0] */
    orc_dsyrk_u('N', nE, 3 * nint, 1.0, a_ce, nE, 0.0, aa, nt);
    for (int j = 0; j < nE; j++) for (int i = j + 1; i < nE; i++) aa[i + nt * j] = aa[j + nt * i];
    if (nH > 0) {
      orc_dgemm('N', 'T', nE, nH, 3 * nint, 1.0, a_e, nE, a_gh, nH, 0.0, aa + (size_t)nt * nE, nt);
      for (int j = 0; j < nH; j++) for (int i = 0; i < nE; i++) aa[nE + j + nt * i] = aa[i + nt * (nE + j)];
    }
  }
  { /* DGETRF + DLASWP + 2 x DTRSM (dhpedgeE.F90:279-309, dhpfaceE_opt.F90:439-458) */
    int *ipiv = malloc(sizeof(int) * nt);
    if (orc_dgetrf(nt, aa, nt, ipiv) != 0) info = info ? info : 1;
    else orc_dgetrs(nt, ncomp, aa, nt, ipiv, bb, nt);
    free(ipiv);
  }
  if (info == 0)
    for (int j = 0; j < nE; j++)
      for (int c = 0; c < ncomp; c++) dofE[c + ncomp * (t0 + j)] = bb[j + nt * c];
  free(shapH); free(gradH); free(shapE); free(curlE); free(aa); free(bb); free(val); free(crl); free(veta); free(ceta); free(pts); free(wts);
  free(a_e); free(a_ce); free(a_gh);
  return info;
}

/* edges first, then faces (update_Ddof.F90 visits the nodes by type in that order); bit i of mask = node i (edges, faces) */
int orc_pbi_hcurl_element(int et, const int *norder, const int *norie, const int *norif, const double *etav, int ncomp, int maxp, unsigned mask,
                          orc_pbi_fnE f, void *ctx, double *dofE) {
  const int nn = orc_nedge(et) + orc_nface(et);
  for (int node = 0; node < nn; node++)
    if (mask & (1u << node)) {
      int rc = orc_pbi_hcurl_node(et, norder, norie, norif, etav, ncomp, maxp, node, f, ctx, dofE);
      if (rc) return rc;
    }
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------------------
 * H(div) Dirichlet dofs: face/dhpfaceV_opt.F90:26-413 (INTEGRATION = 1): L2 projection of the normal component of the datum pulled
 * back to eta, V_eta = det(dxdeta) dxdeta^-1 V (:231-238), onto the face's H(div) functions mapped by the Piola transform of
 * eta(xi) (:251-257).  f as for H(curl) (the curl output is ignored).  Nodes: faces 0..nrf-1.  dofV: (ncomp, sum of face dofs). */
void orc_pbi_offsets_V(int et, const int *norder, int *off) {
  const int nre = orc_nedge(et), nrf = orc_nface(et);
  int n = 0, h, e, v, q;
  for (int i = 0; i < nrf; i++) { off[i] = n; orc_ndof_nod_face(et, i + 1, norder[nre + i], &h, &e, &v, &q); n += v; }
  off[nrf] = n;
}

int orc_pbi_hdiv_node(int et, const int *norder, const int *norie, const int *norif, const double *etav, int ncomp, int maxp, int iface0,
                      orc_pbi_fnE f, void *ctx, double *dofV) {
  const int nrv = orc_nvert(et), nre = orc_nedge(et), nrf = orc_nface(et), integration = 1, jf = iface0 + 1;
  int offV[8];
  orc_pbi_offsets_V(et, norder, offV);
  const int n = offV[iface0 + 1] - offV[iface0], t0 = offV[iface0];
  if (n <= 0) return 0;
  double *shapH = malloc(sizeof(double) * ORC_MAXBRICK_H), *gradH = malloc(sizeof(double) * 3 * ORC_MAXBRICK_H);
  double *shapV = malloc(sizeof(double) * 3 * ORC_MAXBRICK_E), *divV = malloc(sizeof(double) * ORC_MAXBRICK_E);
  double *aa = calloc((size_t)n * n, sizeof(double)), *bb = calloc((size_t)n * ncomp, sizeof(double));
  double *val = malloc(sizeof(double) * 3 * ncomp), *crl = malloc(sizeof(double) * 3 * ncomp), *veta = malloc(sizeof(double) * 3 * ncomp);
  double *pts = malloc(sizeof(double) * 3 * 1000), *wts = malloc(sizeof(double) * 1000);
  int nord1[19], nordi[19], nordf[5], info = 0;
  orc_initiate_order(et, nord1);                       /* :162-164 */
  memcpy(nordi, nord1, sizeof nord1);
  nordi[nre + jf - 1] = norder[nre + jf - 1];
  orc_face_order(et, jf, norder, nordf);
  const int nint = orc_set_2D_int(orc_face_is_tri(et, jf), nordf, 0, integration, maxp, pts, wts);
  double *atest = malloc(sizeof(double) * (size_t)n * 3 * nint);
  for (int l = 0; l < nint; l++) {
    double xi[3], dxidt[6], eta[3], detadxi[9], dxideta[9], rjac, d[6], rn[3], bjac, dxdeta[9], detadx[9], rjx;
    int iflag;
    orc_face_param(et, jf, pts + 2 * l, xi, dxidt);
    orc_shape3DH(et, xi, nord1, norie, norif, shapH, gradH);
    orc_shape3DV(et, xi, nordi, norif, shapV, divV);
    refgeom3D(etav, shapH, gradH, nrv, eta, detadxi, dxideta, &rjac, &iflag);
    if (iflag != 0) info = -1;
    for (int i = 0; i < 2; i++)
      for (int c = 0; c < 3; c++) d[c + 3 * i] = detadxi[c] * dxidt[3 * i] + detadxi[c + 3] * dxidt[3 * i + 1] + detadxi[c + 6] * dxidt[3 * i + 2];
    rn[0] = d[1] * d[5] - d[2] * d[4]; rn[1] = d[2] * d[3] - d[0] * d[5]; rn[2] = d[0] * d[4] - d[1] * d[3];
    bjac = sqrt(rn[0] * rn[0] + rn[1] * rn[1] + rn[2] * rn[2]);
    const int ns = orc_nsign_param(et, jf);
    for (int c = 0; c < 3; c++) rn[c] = rn[c] * ns / bjac;
    const double weight = wts[l] * bjac, sw = sqrt(weight);
    f(eta, val, crl, dxdeta, ctx);
    orc_geom(dxdeta, detadx, &rjx, &iflag);
    for (int c = 0; c < ncomp; c++)
      for (int i = 0; i < 3; i++) {
        double a = 0;
        for (int j = 0; j < 3; j++) a += detadx[i + 3 * j] * val[c + ncomp * j] * rjx;
        veta[c + ncomp * i] = a;
      }
    for (int j = 0; j < n; j++) { /* :247-266 ; one function of each preceding face comes first */
      const int kj = iface0 + j;
      double v[3], prod = 0;
      for (int i = 0; i < 3; i++) v[i] = (detadxi[i] * shapV[3 * kj] + detadxi[i + 3] * shapV[1 + 3 * kj] + detadxi[i + 6] * shapV[2 + 3 * kj]) / rjac;
      for (int i = 0; i < 3; i++) prod += v[i] * rn[i];
      for (int i = 0; i < 3; i++) v[i] = prod * rn[i];
      for (int c = 0; c < ncomp; c++) bb[j + n * c] += (veta[c] * v[0] + veta[c + ncomp] * v[1] + veta[c + 2 * ncomp] * v[2]) * weight;
      for (int i = 0; i < 3; i++) atest[j + (size_t)n * (3 * l + i)] = v[i] * sw;
    }
  }
  orc_dsyrk_u('N', n, 3 * nint, 1.0, atest, n, 0.0, aa, n);   /* DSFRK + DPFTRF + DPFTRS, :277-317 */
  if (orc_dpotrf_u(n, aa, n) != 0) info = info ? info : 1;
  else { orc_dtrsm_u('T', n, ncomp, aa, n, bb, n); orc_dtrsm_u('N', n, ncomp, aa, n, bb, n); }
  if (info == 0)
    for (int j = 0; j < n; j++)
      for (int c = 0; c < ncomp; c++) dofV[c + ncomp * (t0 + j)] = bb[j + n * c];
  free(shapH); free(gradH); free(shapV); free(divV); free(aa); free(bb); free(val); free(crl); free(veta); free(pts); free(wts); free(atest);
  return info;
}

int orc_pbi_hdiv_element(int et, const int *norder, const int *norie, const int *norif, const double *etav, int ncomp, int maxp, unsigned mask,
                         orc_pbi_fnE f, void *ctx, double *dofV) {
  for (int jf = 0; jf < orc_nface(et); jf++)
    if (mask & (1u << jf)) {
      int rc = orc_pbi_hdiv_node(et, norder, norie, norif, etav, ncomp, maxp, jf, f, ctx, dofV);
      if (rc) return rc;
    }
  return 0;
}
