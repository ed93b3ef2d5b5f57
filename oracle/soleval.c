/*
 * soleval.c -- CPU restatement ("oracle") of hp3D's solution evaluation and element error:
 *   soleval        trunk/src/element/util/soleval.F90:30-283   (x, dx/dxi and the solution of every family at a master point)
 *   element_error  trunk/src/element/util/compute_error.F90:226-579 (error and norm of one physical attribute over an element)
 * with the exact solutions of the four problem directories (isol = 1, "sin" solutions):
 *   POISSON u = sin(pi x)sin(pi y)sin(pi z)                            problems/POISSON/GALERKIN/common/exact.F90
 *   MAXWELL E = p e_ic, p = (1+i) sin(w x)sin(w y)sin(w z), H = curl E/(-i w mu)
 *           problems/MAXWELL/ULTRAWEAK_DPG/exact.F90:26-113, common/mfd_solutions.F90:80-100
 * SURVEY.md 8(f) row f4 (error evaluation half).  TEST INFRASTRUCTURE ONLY (see hp3d_oracle.h).
 *
 * Pins (tests/test_error_eval.py): a trilinear / polynomial field set through its dofs is evaluated exactly (error 0 against
 * a table of the same polynomial), the norm of the sin solution over the unit cube equals its closed form to quadrature
 * accuracy, and H(curl)/L2 Piola maps are checked against finite differences of the geometry map.
 */
#include "hp3d_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* soleval.F90:30.  zdof*: (ncomp, nrdof) column-major (component fastest), as solelm returns them; nc* = 0 skips a family.
 * Outputs (any may be NULL): zsolH[nc], zgradH[nc*3] (n + nc*j), zsolE/zcurlE[3*nc] (i + 3n), zsolV[3*nc], zdivV[nc], zsolQ[nc] */
int orc_soleval(int et, const double xi[3], const int *norder, const int *norie, const int *norif, const double *xnod, int ncH,
                const zdouble *zdofH, int ncE, const zdouble *zdofE, int ncV, const zdouble *zdofV, int ncQ, const zdouble *zdofQ,
                double x[3], double dxdxi[9], double *rjac_out, zdouble *zsolH, zdouble *zgradH, zdouble *zsolE, zdouble *zcurlE,
                zdouble *zsolV, zdouble *zdivV, zdouble *zsolQ) {
  static const int MAXD = 3 * ORC_MAXBRICK_E;
  double *shp = malloc(sizeof(double) * MAXD), *der = malloc(sizeof(double) * MAXD);
  double dxidx[9], rjac;
  int iflag;
  /* geometry map (soleval.F90:85-88) */
  int nH = orc_shape3DH(et, xi, norder, norie, norif, shp, der);
  orc_geom3D(xnod, shp, der, nH, x, dxdxi, dxidx, &rjac, &iflag);
  if (rjac_out) *rjac_out = rjac;
  if (ncH > 0 && zsolH) { /* H1: value, gradient mapped by J^-T (:113-127) */
    for (int n = 0; n < ncH; n++) { zsolH[n] = 0; for (int j = 0; j < 3; j++) zgradH[n + ncH * j] = 0; }
    for (int k = 0; k < nH; k++) {
      double gx[3] = {0, 0, 0};
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) gx[j] += der[i + 3 * k] * dxidx[i + 3 * j];
      for (int n = 0; n < ncH; n++) {
        zsolH[n] += zdofH[n + ncH * k] * shp[k];
        for (int j = 0; j < 3; j++) zgradH[n + ncH * j] += zdofH[n + ncH * k] * gx[j];
      }
    }
  }
  if (ncE > 0 && zsolE) { /* H(curl): E = J^-T E^, curl = J C^/det (:145-168) */
    int nE = orc_shape3DE(et, xi, norder, norie, norif, shp, der);
    for (int i = 0; i < 3 * ncE; i++) { zsolE[i] = 0; zcurlE[i] = 0; }
    for (int k = 0; k < nE; k++) {
      double ex[3] = {0, 0, 0}, cx[3] = {0, 0, 0};
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) { ex[i] += dxidx[j + 3 * i] * shp[j + 3 * k]; cx[i] += dxdxi[i + 3 * j] * der[j + 3 * k] / rjac; }
      for (int n = 0; n < ncE; n++)
        for (int i = 0; i < 3; i++) { zsolE[i + 3 * n] += zdofE[n + ncE * k] * ex[i]; zcurlE[i + 3 * n] += zdofE[n + ncE * k] * cx[i]; }
    }
  }
  if (ncV > 0 && zsolV) { /* H(div): V = J V^/det, div = div^/det (:186-208) */
    int nV = orc_shape3DV(et, xi, norder, norif, shp, der);
    for (int i = 0; i < 3 * ncV; i++) zsolV[i] = 0;
    for (int n = 0; n < ncV; n++) zdivV[n] = 0;
    for (int k = 0; k < nV; k++) {
      double vx[3] = {0, 0, 0};
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) vx[i] += dxdxi[i + 3 * j] * shp[j + 3 * k] / rjac;
      for (int n = 0; n < ncV; n++) {
        for (int i = 0; i < 3; i++) zsolV[i + 3 * n] += zdofV[n + ncV * k] * vx[i];
        zdivV[n] += zdofV[n + ncV * k] * (der[k] / rjac);
      }
    }
  }
  if (ncQ > 0 && zsolQ) { /* L2: q = q^/det (:226-238) */
    int nQ = orc_shape3DQ(et, xi, norder, shp);
    for (int n = 0; n < ncQ; n++) zsolQ[n] = 0;
    for (int k = 0; k < nQ; k++) for (int n = 0; n < ncQ; n++) zsolQ[n] += zdofQ[n + ncQ * k] * (shp[k] / rjac);
  }
  free(shp); free(der);
  return iflag;
}

static void sinpot(double a, zdouble c, const double x[3], zdouble *p, zdouble g[3], zdouble h[9]) {
  double s[3], co[3];
  for (int i = 0; i < 3; i++) { s[i] = sin(x[i] * a); co[i] = cos(x[i] * a); }
  *p = s[0] * s[1] * s[2] * c;
  g[0] = a * co[0] * s[1] * s[2] * c; g[1] = a * co[1] * s[0] * s[2] * c; g[2] = a * co[2] * s[0] * s[1] * c;
  const double a2 = a * a;
  h[0] = h[4] = h[8] = -a2 * s[0] * s[1] * s[2] * c;
  h[1] = h[3] = a2 * co[0] * co[1] * s[2] * c;
  h[2] = h[6] = a2 * co[0] * co[2] * s[1] * c;
  h[5] = h[7] = a2 * co[1] * co[2] * s[0] * c;
}

/* Number of exact values per quadrature point the field variable of `kind` is compared with (the layout of a caller table):
 * H1 field (kinds 1,2): [u, du/dx, du/dy, du/dz]; H(curl) field (kind 3): [E(3), curl E(3)]; L2 field (kind 4): [E(3), H(3)] */
int orc_error_nvals(int kind) { return kind <= 2 ? 4 : 6; }

/* the exact solution of the problem's FIELD variable at x (isol = 1), in the layout of orc_error_nvals */
void orc_exact_field(int kind, const orc_params *prm, const double x[3], zdouble *val) {
  zdouble p, g[3], h[9];
  if (kind <= 2) {
    sinpot(M_PI, 1.0, x, &p, g, h);
    val[0] = p; for (int j = 0; j < 3; j++) val[1 + j] = g[j];
    return;
  }
  sinpot(prm->omega, 1.0 + 1.0 * I, x, &p, g, h);
  const int ic = prm->icomp_exact - 1;
  zdouble E[3] = {0, 0, 0}, cE[3];
  E[ic] = p;
  /* curl (p e_ic): dE(i,j) = d_j E_i nonzero only for i = ic */
  zdouble dE[3][3] = {{0}};
  for (int j = 0; j < 3; j++) dE[ic][j] = g[j];
  cE[0] = dE[2][1] - dE[1][2]; cE[1] = dE[0][2] - dE[2][0]; cE[2] = dE[1][0] - dE[0][1];
  for (int i = 0; i < 3; i++) val[i] = E[i];
  if (kind == 3) for (int i = 0; i < 3; i++) val[3 + i] = cE[i];
  else for (int i = 0; i < 3; i++) val[3 + i] = cE[i] / (-I * prm->omega * prm->mu);   /* exact.F90:89 */
}

/* element_error (compute_error.F90:226) for the FIELD variable of problem `kind`: quadrature set_3Dint with INTEGRATION = 2
 * (:305-307), weight = wa*rjac, error / norm accumulated over components, values (+ gradient / curl unless l2proj).
 * zdof: (ncomp, nrdof) column-major dofs of that variable (kind 1,2: H1, 1 comp; 3: H(curl), 1; 4: L2, 6).
 * exact_tab: NULL (built-in isol = 1) or nint*nvals values per element in quadrature order. */
int orc_element_error(int et, int kind, const int *norder, const int *norie, const int *norif, const double *xnod, const zdouble *zdof,
                      const orc_params *prm, const zdouble *exact_tab, int l2proj, double *err, double *rnorm) {
  double *xiloc = malloc(sizeof(double) * 3 * 2000), *waloc = malloc(sizeof(double) * 2000);
  const int nint = orc_set_3D_int(et, norder, norif, 2, orc_get_maxp(), xiloc, waloc);
  const int nv = orc_error_nvals(kind);
  double e = 0, r = 0;
  int bad = 0;
  for (int l = 0; l < nint; l++) {
    double x[3], dxdxi[9], rjac;
    zdouble sH[1], gH[3], sE[3], cE[3], sQ[6], ex[6];
    int fl;
    if (kind <= 2) fl = orc_soleval(et, xiloc + 3 * l, norder, norie, norif, xnod, 1, zdof, 0, 0, 0, 0, 0, 0, x, dxdxi, &rjac, sH, gH, 0, 0, 0, 0, 0);
    else if (kind == 3) fl = orc_soleval(et, xiloc + 3 * l, norder, norie, norif, xnod, 0, 0, 1, zdof, 0, 0, 0, 0, x, dxdxi, &rjac, 0, 0, sE, cE, 0, 0, 0);
    else fl = orc_soleval(et, xiloc + 3 * l, norder, norie, norif, xnod, 0, 0, 0, 0, 0, 0, 6, zdof, x, dxdxi, &rjac, 0, 0, 0, 0, 0, 0, sQ);
    if (fl) bad = 1;
    if (exact_tab) memcpy(ex, exact_tab + (size_t)l * nv, sizeof(zdouble) * nv); else orc_exact_field(kind, prm, x, ex);
    const double w = waloc[l] * rjac;
    if (kind <= 2) {
      if (!l2proj) for (int j = 0; j < 3; j++) { double d = cabs(ex[1 + j] - gH[j]); e += d * d * w; r += cabs(ex[1 + j]) * cabs(ex[1 + j]) * w; }
      { double d = cabs(ex[0] - sH[0]); e += d * d * w; r += cabs(ex[0]) * cabs(ex[0]) * w; }
    } else if (kind == 3) {
      if (!l2proj) for (int j = 0; j < 3; j++) { double d = cabs(cE[j] - ex[3 + j]); e += d * d * w; r += cabs(ex[3 + j]) * cabs(ex[3 + j]) * w; }
      for (int j = 0; j < 3; j++) { double d = cabs(sE[j] - ex[j]); e += d * d * w; r += cabs(ex[j]) * cabs(ex[j]) * w; }
    } else {
      for (int j = 0; j < 6; j++) { double d = cabs(ex[j] - sQ[j]); e += d * d * w; r += cabs(ex[j]) * cabs(ex[j]) * w; }
    }
  }
  *err = e; *rnorm = r;
  free(xiloc); free(waloc);
  return bad ? -1 : nint;
}

/* quadrature points (physical coordinates) of element_error, for callers that tabulate their own exact solution */
int orc_error_points(int et, const int *norder, const int *norie, const int *norif, const double *xnod, double *xq /*(3,nint)*/) {
  double *xiloc = malloc(sizeof(double) * 3 * 2000), *waloc = malloc(sizeof(double) * 2000);
  const int nint = orc_set_3D_int(et, norder, norif, 2, orc_get_maxp(), xiloc, waloc);
  for (int l = 0; l < nint; l++) {
    double dxdxi[9], rjac;
    orc_soleval(et, xiloc + 3 * l, norder, norie, norif, xnod, 0, 0, 0, 0, 0, 0, 0, 0, xq + 3 * l, dxdxi, &rjac, 0, 0, 0, 0, 0, 0, 0);
  }
  free(xiloc); free(waloc);
  return nint;
}
