/*
 * oracle/quad_geom.c -- quadrature rules, geometry map, dof counting (TEST INFRASTRUCTURE ONLY).
 *
 * Follows (relative to /root/reference/trunk/src):
 *   element/quadrature/gauss_quadrature.F90:518-651 (1-D Gauss-Legendre nodes/weights as the reference
 *       hard-codes them, ~15 significant digits) and :749-750 (map to [0,1]: x=(1+x)/2, w=w/2)
 *   element/quadrature/set_3D_int.F90:155-259 (set_3Dint_aux, brick) ; set_2D_int.F90:121-236 (quad)
 *   datstrs/find_order.F90:68 (find_order_loc) ; modules/element_data.F90:246 (NFAXES)
 *   element/util/geom.F90:30 ; element/util/geom3D.F90:30,149
 *   modules/element_data.F90:554 (face_param), :616 (Nsign_param), :750 (face_order), :808 (ndof_nod)
 *   element/util/celndof.F90:24
 */
#include "hp3d_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

/* Gauss-Legendre on [-1,1]; only the non-negative half is listed (rule is symmetric).  The digits are
 * the ones the reference carries (gauss_quadrature.F90:518-651): parity needs the same truncation. */
static const double GX[10][5] = {
  {0.0},
  {.577350269189626},
  {.000000000000000, .774596669241483},
  {.339981043584856, .861136311594053},
  {.000000000000000, .538469310105683, .906179845938664},
  {.238619186083197, .661209386466265, .932469514203152},
  {.000000000000000, .405845151377397, .741531185599394, .949107912342759},
  {.183434642495650, .525532409916329, .796666477413627, .960289856497536},
  {.000000000000000, .324253423403809, .613371432700590, .836031107326636, .968160239507626},
  {.1488743389816312108848260, .4333953941292471907992659, .6794095682990244062343274,
   .8650633666889845107320967, .9739065285171717200779640}};
static const double GW[10][5] = {
  {2.000000000000000},
  {1.000000000000000},
  {.888888888888889, .555555555555556},
  {.652145154862546, .347854845137454},
  {.568888888888889, .478628670499366, .236926885056189},
  {.467913934572691, .360761573048139, .171324492379170},
  {.417959183673469, .381830050505119, .279705391489277, .129484966168870},
  {.362683783378362, .313706645877887, .222381034453374, .101228536290376},
  {.330239355001260, .312347077040003, .260610696402935, .180648160694857, .081274388361574},
  {.2955242247147528701738930, .2692667193099963550912269, .2190863625159820439955349,
   .1494513491505805931457763, .0666713443086881375935688}};

void orc_gauss1(int n, double *xi, double *w) {
  if (n < 1 || n > 10) { fprintf(stderr, "orc_gauss1: n=%d out of table\n", n); exit(1); }
  int half = n / 2, odd = n & 1;
  for (int j = 0; j < n; j++) {
    double x, wt;
    if (j < half) { int k = half - 1 - j + odd; x = -GX[n - 1][k]; wt = GW[n - 1][k]; }
    else          { int k = j - half;               x =  GX[n - 1][k]; wt = GW[n - 1][k]; }
    xi[j] = 0.5 * (1.0 + x);
    w[j] = 0.5 * wt;
  }
}

/* NFAXES(3,ori): have the face axes been swapped?  element_data.F90:246-249 */
static const int NFAXES3[8] = {0, 1, 0, 1, 1, 0, 1, 0};

void orc_find_order_loc_hexa(const int norder[19], const int norif[6], int nloc[19]) {
  for (int i = 0; i < 19; i++) nloc[i] = norder[i];
  for (int j = 0; j < 6; j++)
    if (NFAXES3[norif[j]] == 1) {
      int h, v;
      orc_decode(norder[12 + j], &h, &v);
      nloc[12 + j] = v * 10 + h;
    }
}

/* set_3D_int[_DPG] for the brick: set_3D_int.F90:47-62,112-127,203-259.  `maxp` is MAXP (Galerkin) or
 * MAXPP (DPG); `integration` is control::INTEGRATION (DPG sets it to NORD_ADD around the call). */
int orc_set_3D_int_hexa(const int norder[19], const int norif[6], int integration, int maxp, double *xiloc,
                        double *waloc) {
  int nl[19], nx = 0, ny = 0, nz = 0, hv[2], xyz[3];
  orc_find_order_loc_hexa(norder, norif, nl);
#define MAXI(a, b) ((a) > (b) ? (a) : (b))
  for (int i = 1; i <= 19; i++) {
    int o = nl[i - 1];
    switch (i) {
      case 1: case 3: case 5: case 7: nx = MAXI(nx, o); break;
      case 2: case 4: case 6: case 8: ny = MAXI(ny, o); break;
      case 9: case 10: case 11: case 12: nz = MAXI(nz, o); break;
      case 13: case 14: orc_decod(o, 10, 2, hv); nx = MAXI(nx, hv[0]); ny = MAXI(ny, hv[1]); break;
      case 15: case 17: orc_decod(o, 10, 2, hv); nx = MAXI(nx, hv[0]); nz = MAXI(nz, hv[1]); break;
      case 16: case 18: orc_decod(o, 10, 2, hv); ny = MAXI(ny, hv[0]); nz = MAXI(nz, hv[1]); break;
      case 19: orc_decod(o, 10, 3, xyz); nx = MAXI(nx, xyz[0]); ny = MAXI(ny, xyz[1]); nz = MAXI(nz, xyz[2]); break;
    }
  }
#define MINI(a, b) ((a) < (b) ? (a) : (b))
  nx = MINI(nx + integration, maxp); ny = MINI(ny + integration, maxp); nz = MINI(nz + integration, maxp);
  int n1 = nx + 1, n2 = ny + 1, n3 = nz + 1;
  double x1[10], w1[10], x2[10], w2[10], x3[10], w3[10];
  orc_gauss1(n1, x1, w1); orc_gauss1(n2, x2, w2); orc_gauss1(n3, x3, w3);
  int l = 0;
  for (int l3 = 0; l3 < n3; l3++)
    for (int l2 = 0; l2 < n2; l2++)
      for (int l1 = 0; l1 < n1; l1++) {
        xiloc[3 * l + 0] = x1[l1]; xiloc[3 * l + 1] = x2[l2]; xiloc[3 * l + 2] = x3[l3];
        waloc[l] = w1[l1] * w2[l2] * w3[l3];
        l++;
      }
  return l;
}

/* set_2D_int[_DPG] for a quad face: set_2D_int.F90:121-150 (swap for NFAXES), :177-236 */
int orc_set_2D_int_quad(const int nordf[5], int norif, int integration, int maxp, double *tloc, double *wtloc) {
  int nl[5], xy[2];
  for (int i = 0; i < 5; i++) nl[i] = nordf[i];
  if (NFAXES3[norif] == 1) { int h, v; orc_decode(nordf[4], &h, &v); nl[4] = v * 10 + h; }
  orc_decod(nl[4], 10, 2, xy);
  int nx = MAXI(MAXI(nl[0], nl[2]), xy[0]), ny = MAXI(MAXI(nl[1], nl[3]), xy[1]);
  nx = MINI(nx + integration, maxp); ny = MINI(ny + integration, maxp);
  int n1 = nx + 1, n2 = ny + 1, l = 0;
  double x1[10], w1[10], x2[10], w2[10];
  orc_gauss1(n1, x1, w1); orc_gauss1(n2, x2, w2);
  for (int l2 = 0; l2 < n2; l2++)
    for (int l1 = 0; l1 < n1; l1++) {
      tloc[2 * l] = x1[l1]; tloc[2 * l + 1] = x2[l2];
      wtloc[l] = w1[l1] * w2[l2];
      l++;
    }
  return l;
}

/* geom.F90:57-113 : Sarrus determinant + cofactor inverse; column-major 3x3 (a(i,j) = a[i+3j]) */
void orc_geom(const double J[9], double Ji[9], double *rjac, int *iflag) {
#define A(i, j) J[(i - 1) + 3 * (j - 1)]
#define B(i, j) Ji[(i - 1) + 3 * (j - 1)]
  *iflag = 0;
  double r = A(1,1)*A(2,2)*A(3,3) + A(2,1)*A(3,2)*A(1,3) + A(3,1)*A(1,2)*A(2,3)
           - A(3,1)*A(2,2)*A(1,3) - A(1,1)*A(3,2)*A(2,3) - A(2,1)*A(1,2)*A(3,3);
  *rjac = r;
  if (r < 0.0) *iflag = 1;
  B(1,1) = ( A(2,2)*A(3,3) - A(3,2)*A(2,3)) / r;
  B(2,1) = (-A(2,1)*A(3,3) + A(3,1)*A(2,3)) / r;
  B(3,1) = ( A(2,1)*A(3,2) - A(3,1)*A(2,2)) / r;
  B(1,2) = ( A(3,2)*A(1,3) - A(1,2)*A(3,3)) / r;
  B(2,2) = ( A(1,1)*A(3,3) - A(3,1)*A(1,3)) / r;
  B(3,2) = (-A(1,1)*A(3,2) + A(3,1)*A(1,2)) / r;
  B(1,3) = ( A(1,2)*A(2,3) - A(2,2)*A(1,3)) / r;
  B(2,3) = (-A(1,1)*A(2,3) + A(2,1)*A(1,3)) / r;
  B(3,3) = ( A(1,1)*A(2,2) - A(2,1)*A(1,2)) / r;
#undef A
#undef B
}

/* geom3D.F90:77-91 (isoparametric branch, EXGEOM=0) */
void orc_geom3D(const double *xnod, const double *shapH, const double *gradH, int nrdofH, double x[3],
                double dxdxi[9], double dxidx[9], double *rjac, int *iflag) {
  for (int i = 0; i < 3; i++) x[i] = 0.0;
  for (int i = 0; i < 9; i++) dxdxi[i] = 0.0;
  for (int k = 0; k < nrdofH; k++) {
    for (int c = 0; c < 3; c++) x[c] += xnod[3 * k + c] * shapH[k];
    for (int i = 0; i < 3; i++)
      for (int c = 0; c < 3; c++) dxdxi[c + 3 * i] += xnod[3 * k + c] * gradH[3 * k + i];
  }
  orc_geom(dxdxi, dxidx, rjac, iflag);
}

/* geom3D.F90:149-203 */
void orc_bgeom3D(const double *xnod, const double *shapH, const double *gradH, int nrdofH,
                 const double dxidt[6], int nsign, double x[3], double dxdxi[9], double dxidx[9], double *rjac,
                 double dxdt[6], double rn[3], double *bjac) {
  int iflag;
  orc_geom3D(xnod, shapH, gradH, nrdofH, x, dxdxi, dxidx, rjac, &iflag);
  if (iflag != 0) { fprintf(stderr, "orc_bgeom3D: negative Jacobian %e\n", *rjac); exit(1); }
  for (int i = 0; i < 6; i++) dxdt[i] = 0.0;
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 3; j++)
      for (int c = 0; c < 3; c++) dxdt[c + 3 * i] += dxdxi[c + 3 * j] * dxidt[j + 3 * i];
  const double *a = dxdt, *b = dxdt + 3;
  rn[0] = a[1] * b[2] - a[2] * b[1];
  rn[1] = a[2] * b[0] - a[0] * b[2];
  rn[2] = a[0] * b[1] - a[1] * b[0];
  *bjac = sqrt(rn[0] * rn[0] + rn[1] * rn[1] + rn[2] * rn[2]);
  for (int c = 0; c < 3; c++) rn[c] = rn[c] * nsign / (*bjac);
}

/* element_data.F90:31-35, :90-93, :106-109 */
static const double BRICK_COORD[8][3] = {{0,0,0},{1,0,0},{1,1,0},{0,1,0},{0,0,1},{1,0,1},{1,1,1},{0,1,1}};
static const int BRICK_FACE_TO_VERT[6][4] = {{1,2,3,4},{5,6,7,8},{1,2,6,5},{2,3,7,6},{4,3,7,8},{1,4,8,5}};
static const int BRICK_FACE_TO_EDGE[6][4] = {{1,2,3,4},{5,6,7,8},{1,10,5,9},{2,11,6,10},{3,11,7,12},{4,12,8,9}};

/* element_data.F90:554-603 ; iface is 1-based */
void orc_face_param_hexa(int iface, const double t[2], double xi[3], double dxidt[6]) {
  const double *x1 = BRICK_COORD[BRICK_FACE_TO_VERT[iface - 1][0] - 1];
  const double *x2 = BRICK_COORD[BRICK_FACE_TO_VERT[iface - 1][1] - 1];
  const double *x3 = BRICK_COORD[BRICK_FACE_TO_VERT[iface - 1][3] - 1];
  for (int c = 0; c < 3; c++) {
    dxidt[c] = x2[c] - x1[c];
    dxidt[3 + c] = x3[c] - x1[c];
    xi[c] = x1[c] + t[0] * dxidt[c] + t[1] * dxidt[3 + c];
  }
}
/* element_data.F90:616-640 */
int orc_nsign_param_hexa(int iface) { return (iface == 1 || iface == 5 || iface == 6) ? -1 : 1; }
/* element_data.F90:750-790 */
void orc_face_order_hexa(int iface, const int norder[19], int nordf[5]) {
  for (int i = 0; i < 4; i++) nordf[i] = norder[BRICK_FACE_TO_EDGE[iface - 1][i] - 1];
  nordf[4] = norder[12 + iface - 1];
}

/* element_data.F90:808-870 */
void orc_ndof_nod_quad(int nord, int *h, int *e, int *v, int *q) {
  int nx, ny;
  orc_decode(nord, &nx, &ny);
  *h = (nx - 1) * (ny - 1);
  *e = nx * (ny - 1) + (nx - 1) * ny;
  *v = nx * ny;
  *q = 0;
}
void orc_ndof_nod_hexa(int nord, int *h, int *e, int *v, int *q) {
  int naux, nx, ny, nz;
  orc_decode(nord, &naux, &nz);
  orc_decode(naux, &nx, &ny);
  *h = (nx - 1) * (ny - 1) * (nz - 1);
  *e = nx * (ny - 1) * (nz - 1) + (nx - 1) * ny * (nz - 1) + (nx - 1) * (ny - 1) * nz;
  *v = (nx - 1) * ny * nz + nx * (ny - 1) * nz + nx * ny * (nz - 1);
  *q = nx * ny * nz;
}
/* celndof.F90:24-100 (brick branch) */
void orc_celndof_hexa(const int nord[19], int *H, int *E, int *V, int *Q) {
  int h = 8, e = 0, v = 0, q = 0, a, b, c, d;
  for (int i = 0; i < 12; i++) { h += nord[i] - 1; e += nord[i]; }
  for (int i = 12; i < 18; i++) { orc_ndof_nod_quad(nord[i], &a, &b, &c, &d); h += a; e += b; v += c; q += d; }
  orc_ndof_nod_hexa(nord[18], &a, &b, &c, &d);
  h += a; e += b; v += c; q += d;
  *H = h; *E = e; *V = v; *Q = q;
}
/* MAXWELL/ULTRAWEAK_DPG/elem/elem.F90:147-190 (brick branch) */
void orc_compute_enriched_order_hexa(int nordP, int norder[19]) {
  int temp[2], nF[3], nB[3], xy[2];
  orc_decod(nordP, 10, 2, temp);
  nF[0] = temp[0]; nB[2] = temp[1];
  orc_decod(nF[0], 10, 2, xy);
  nB[0] = xy[0]; nB[1] = xy[1];
  int xz[2] = {nB[0], nB[2]}, yz[2] = {nB[1], nB[2]};
  orc_encod(xz, 10, 2, &nF[1]);
  orc_encod(yz, 10, 2, &nF[2]);
  for (int i = 0; i < 8; i++) norder[i] = (i % 2 == 0) ? nB[0] : nB[1];
  for (int i = 8; i < 12; i++) norder[i] = nB[2];
  norder[12] = norder[13] = nF[0];
  norder[14] = nF[1]; norder[15] = nF[2]; norder[16] = nF[1]; norder[17] = nF[2];
  norder[18] = nordP;
}
