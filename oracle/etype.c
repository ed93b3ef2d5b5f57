/*
 * oracle/etype.c -- element-type layer of the CPU oracle (TEST INFRASTRUCTURE ONLY): the `select case(ntype)`
 * dispatchers of the reference for the two element types on the hot path (brick MDLB=1, prism MDLP=3,
 * src/modules/node_types.F90:8-10), plus the prism branches of the quadrature / topology helpers.
 *
 * Follows (relative to /root/reference/trunk/src):
 *   element/shape_1/ContExactSequence.F90:397,468,538,600 (shape3DH/E/V/Q), broken/BrokenExactSequence.F90:401,481,561,641
 *   element/quadrature/set_3D_int.F90:261-288 (prism), set_2D_int.F90:195-204 (triangle)
 *   datstrs/find_order.F90:68-100 (find_order_loc), hpinterp/initiate_order.F90
 *   modules/element_data.F90:25-29,62-65,85-88,108-111 (prism tables), :554 face_param, :616 Nsign_param,
 *       :650 Face_type, :750 face_order, :808 ndof_nod ; element/util/celndof.F90:24
 *   problems/MAXWELL/ULTRAWEAK_DPG/elem/elem.F90:147 (compute_enriched_order)
 */
#include "hp3d_oracle.h"
#include "tri_rules.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAXI(a, b) ((a) > (b) ? (a) : (b))
#define MINI(a, b) ((a) < (b) ? (a) : (b))

int orc_nvert(int et) { return et == ORC_MDLP ? 6 : 8; }
int orc_nedge(int et) { return et == ORC_MDLP ? 9 : 12; }
int orc_nface(int et) { return et == ORC_MDLP ? 5 : 6; }
/* 0: quad (RECT), 1: triangle (TRIA) ; iface 1-based */
int orc_face_is_tri(int et, int iface) { return et == ORC_MDLP && iface <= 2; }

int orc_shape3DH(int et, const double xi[3], const int *nord, const int *norie, const int *norif, double *s, double *g) {
  return et == ORC_MDLP ? orc_shape3DH_pris(xi, nord, norie, norif, s, g) : orc_shape3DH_hexa(xi, nord, norie, norif, s, g);
}
int orc_shape3DE(int et, const double xi[3], const int *nord, const int *norie, const int *norif, double *s, double *c) {
  return et == ORC_MDLP ? orc_shape3DE_pris(xi, nord, norie, norif, s, c) : orc_shape3DE_hexa(xi, nord, norie, norif, s, c);
}
int orc_shape3DV(int et, const double xi[3], const int *nord, const int *norif, double *s, double *d) {
  return et == ORC_MDLP ? orc_shape3DV_pris(xi, nord, norif, s, d) : orc_shape3DV_hexa(xi, nord, norif, s, d);
}
int orc_shape3DQ(int et, const double xi[3], const int *nord, double *s) {
  return et == ORC_MDLP ? orc_shape3DQ_pris(xi, nord, s) : orc_shape3DQ_hexa(xi, nord, s);
}
int orc_shape3HH(int et, const double xi[3], int nordM, double *s, double *g) {
  return et == ORC_MDLP ? orc_shape3HH_pris(xi, nordM, s, g) : orc_shape3HH_hexa(xi, nordM, s, g);
}
int orc_shape3EE(int et, const double xi[3], int nordM, double *s, double *c) {
  return et == ORC_MDLP ? orc_shape3EE_pris(xi, nordM, s, c) : orc_shape3EE_hexa(xi, nordM, s, c);
}
int orc_shape3VV(int et, const double xi[3], int nordM, double *s, double *d) {
  return et == ORC_MDLP ? orc_shape3VV_pris(xi, nordM, s, d) : orc_shape3VV_hexa(xi, nordM, s, d);
}
int orc_shape3QQ(int et, const double xi[3], int nordM, double *s) {
  return et == ORC_MDLP ? orc_shape3QQ_pris(xi, nordM, s) : orc_shape3QQ_hexa(xi, nordM, s);
}

/* ---- dof counts */
void orc_ndof_nod_tria(int nord, int *h, int *e, int *v, int *q) {
  *h = (nord - 2) * (nord - 1) / 2; *e = (nord - 1) * nord; *v = nord * (nord + 1) / 2; *q = 0;
}
void orc_ndof_nod_pris(int nord, int *h, int *e, int *v, int *q) {
  int nx, nz;
  orc_decode(nord, &nx, &nz);
  *h = (nx - 2) * (nx - 1) / 2 * (nz - 1);
  *e = (nx - 1) * nx * (nz - 1) + (nx - 2) * (nx - 1) / 2 * nz;
  *v = (nx - 1) * nx * nz + nx * (nx + 1) / 2 * (nz - 1);
  *q = (nx + 1) * nx / 2 * nz;
}
void orc_ndof_nod_mid(int et, int nord, int *h, int *e, int *v, int *q) {
  if (et == ORC_MDLP) orc_ndof_nod_pris(nord, h, e, v, q); else orc_ndof_nod_hexa(nord, h, e, v, q);
}
void orc_ndof_nod_face(int et, int iface, int nord, int *h, int *e, int *v, int *q) {
  if (orc_face_is_tri(et, iface)) orc_ndof_nod_tria(nord, h, e, v, q); else orc_ndof_nod_quad(nord, h, e, v, q);
}
void orc_celndof(int et, const int *nord, int *H, int *E, int *V, int *Q) {
  if (et != ORC_MDLP) { orc_celndof_hexa(nord, H, E, V, Q); return; }
  int h = 6, e = 0, v = 0, q = 0, a, b, c, d;
  for (int i = 0; i < 9; i++) { h += nord[i] - 1; e += nord[i]; }
  for (int i = 9; i < 11; i++) { orc_ndof_nod_tria(nord[i], &a, &b, &c, &d); h += a; e += b; v += c; q += d; }
  for (int i = 11; i < 14; i++) { orc_ndof_nod_quad(nord[i], &a, &b, &c, &d); h += a; e += b; v += c; q += d; }
  orc_ndof_nod_pris(nord[14], &a, &b, &c, &d);
  h += a; e += b; v += c; q += d;
  *H = h; *E = e; *V = v; *Q = q;
}
/* enriched order of the middle node: elem.F90:67-73 ; interface-only middle order: elem_opt.F90:184-194 */
int orc_enriched_mid(int et, int nord_mid, int dp) { return nord_mid + dp * (et == ORC_MDLP ? 11 : 111); }
int orc_trace_mid(int et) { return et == ORC_MDLP ? 11 : 111; }
void orc_compute_enriched_order(int et, int nordP, int *norder) {
  if (et != ORC_MDLP) { orc_compute_enriched_order_hexa(nordP, norder); return; }
  int nb[2];
  orc_decod(nordP, 10, 2, nb);
  for (int i = 0; i < 6; i++) norder[i] = nb[0];
  for (int i = 6; i < 9; i++) norder[i] = nb[1];
  norder[9] = norder[10] = nb[0];
  norder[11] = norder[12] = norder[13] = nordP;
  norder[14] = nordP;
}
void orc_initiate_order(int et, int *norder) {
  if (et == ORC_MDLP) {
    for (int i = 0; i < 11; i++) norder[i] = 1;
    for (int i = 11; i < 15; i++) norder[i] = 11;
  } else {
    for (int i = 0; i < 12; i++) norder[i] = 1;
    for (int i = 12; i < 18; i++) norder[i] = 11;
    norder[18] = 111;
  }
}

/* ---- quadrature */
static const int NFAXES3[8] = {0, 1, 0, 1, 1, 0, 1, 0};
static int tri_rule(int nord, double *t2 /*(2,n)*/, double *w) {
  if (nord < 1 || nord > 9) { fprintf(stderr, "oracle: triangle rule order %d out of range (NSELECT has 9 entries)\n", nord); exit(1); }
  int n = TRI_RULE_NPTS[nord - 1], o = TRI_RULE_OFF[nord - 1];
  for (int l = 0; l < n; l++) { t2[2 * l] = TRI_RULE_PTS[o + l][0]; t2[2 * l + 1] = TRI_RULE_PTS[o + l][1]; w[l] = TRI_RULE_PTS[o + l][2]; }
  return n;
}
int orc_set_3D_int(int et, const int *norder, const int *norif, int integration, int maxp, double *xiloc, double *waloc) {
  if (et != ORC_MDLP) return orc_set_3D_int_hexa(norder, norif, integration, maxp, xiloc, waloc);
  int nl[15], nh = 0, nz = 0, hv[2];
  for (int i = 0; i < 15; i++) nl[i] = norder[i];
  for (int j = 0; j < 3; j++)   /* find_order_loc, prism branch */
    if (NFAXES3[norif[2 + j]] == 1) { int h, v; orc_decode(norder[11 + j], &h, &v); nl[11 + j] = v * 10 + h; }
  for (int i = 1; i <= 15; i++) {
    int o = nl[i - 1];
    if (i <= 6 || i == 10 || i == 11) nh = MAXI(nh, o);
    else if (i <= 9) nz = MAXI(nz, o);
    else { orc_decod(o, 10, 2, hv); nh = MAXI(nh, hv[0]); nz = MAXI(nz, hv[1]); }
  }
  nh = MINI(nh + integration, maxp); nz = MINI(nz + integration, maxp);
  double t2[2 * 80], wt[80], x3[10], w3[10];
  int nx = tri_rule(nh, t2, wt), n3 = nz + 1, l = 0;
  orc_gauss1(n3, x3, w3);
  for (int l2 = 0; l2 < n3; l2++)
    for (int l1 = 0; l1 < nx; l1++) {
      xiloc[3 * l] = t2[2 * l1]; xiloc[3 * l + 1] = t2[2 * l1 + 1]; xiloc[3 * l + 2] = x3[l2];
      waloc[l] = wt[l1] * w3[l2];
      l++;
    }
  return l;
}
int orc_set_2D_int(int is_tri, const int nordf[5], int norif, int integration, int maxp, double *tloc, double *wtloc) {
  if (!is_tri) return orc_set_2D_int_quad(nordf, norif, integration, maxp, tloc, wtloc);
  int nord = MAXI(MAXI(nordf[0], nordf[1]), MAXI(nordf[2], nordf[3]));
  nord = MINI(nord + integration, maxp);
  return tri_rule(nord, tloc, wtloc);
}

/* ---- master prism topology, 1-based tables as in element_data.F90 */
static const double PRISM_COORD[6][3] = {{0,0,0},{1,0,0},{0,1,0},{0,0,1},{1,0,1},{0,1,1}};
static const int PRISM_FACE_TO_VERT[5][4] = {{1,2,3,1},{4,5,6,4},{1,2,5,4},{2,3,6,5},{1,3,6,4}};
static const int PRISM_FACE_TO_EDGE[5][4] = {{1,2,3,1},{4,5,6,4},{1,8,4,7},{2,9,5,8},{3,9,6,7}};

void orc_face_param(int et, int iface, const double t[2], double xi[3], double dxidt[6]) {
  if (et != ORC_MDLP) { orc_face_param_hexa(iface, t, xi, dxidt); return; }
  int k = orc_face_is_tri(et, iface) ? 3 : 4;
  const double *x1 = PRISM_COORD[PRISM_FACE_TO_VERT[iface - 1][0] - 1];
  const double *x2 = PRISM_COORD[PRISM_FACE_TO_VERT[iface - 1][1] - 1];
  const double *x3 = PRISM_COORD[PRISM_FACE_TO_VERT[iface - 1][k - 1] - 1];
  for (int c = 0; c < 3; c++) {
    dxidt[c] = x2[c] - x1[c];
    dxidt[3 + c] = x3[c] - x1[c];
    xi[c] = x1[c] + t[0] * dxidt[c] + t[1] * dxidt[3 + c];
  }
}
int orc_nsign_param(int et, int iface) {
  if (et != ORC_MDLP) return orc_nsign_param_hexa(iface);
  return (iface == 1 || iface == 5) ? -1 : 1;
}
void orc_face_order(int et, int iface, const int *norder, int nordf[5]) {
  if (et != ORC_MDLP) { orc_face_order_hexa(iface, norder, nordf); return; }
  for (int i = 0; i < 5; i++) nordf[i] = 0;
  if (iface <= 2) {
    for (int i = 0; i < 3; i++) nordf[i] = norder[PRISM_FACE_TO_EDGE[iface - 1][i] - 1];
    nordf[3] = norder[9 + iface - 1];
  } else {
    for (int i = 0; i < 4; i++) nordf[i] = norder[PRISM_FACE_TO_EDGE[iface - 1][i] - 1];
    nordf[4] = norder[9 + iface - 1];
  }
}
