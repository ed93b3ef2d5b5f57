/*
 * oracle/elem.c -- CPU restatement of hp3D's element routines + static condensation
 * (TEST INFRASTRUCTURE ONLY; never linked into the product).
 *
 * Follows (relative to /root/reference/trunk):
 *   problems/POISSON/GALERKIN/elem_opt.F90:22-142          (DSYRK formulation)
 *   problems/POISSON/PRIMAL_DPG/elem_opt.F90:32-459        (Gram + DPFTRF/DTFSM/DSYRK)
 *   problems/MAXWELL/GALERKIN/elem_opt.F90:22-158          (2 x ZSYRK)
 *   problems/MAXWELL/ULTRAWEAK_DPG/elem/elem_opt.F90:25-957 (adjoint-graph Gram, ZPOTRF/ZTRTRS/ZHERK)
 *   src/modules/stc.F90:94 (stc_get_nrdof), :182 (stc_fwd_wrapper), :338 (stc_fwd_herm), :443 (stc_fwd_gen)
 * The user callbacks getf/get_permittivity (problem files, not library code) are replaced by the
 * built-in manufactured sources below (same form as common/mfd_solutions.F90:80-100, isol=1) or by a
 * per-quadrature-point table.
 */
#include "hp3d_oracle.h"
#include "dense.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MIDX(et) (orc_nedge(et) + orc_nface(et))
#define IDX(A, ld, i, j) (A)[(size_t)(i) + (size_t)(ld) * (size_t)(j)]
static void *xmalloc(size_t n) { void *p = calloc(n ? n : 1, 1); if (!p) { fprintf(stderr, "oracle: out of memory\n"); exit(1); } return p; }

void orc_params_default(orc_params *p) {
  memset(p, 0, sizeof *p);
  p->nord_add = 1; p->test_norm = 1; p->alpha_norm = 1.0;
  p->omega = 1.0; p->eps = 1.0; p->mu = 1.0; p->sigma = 0.0;
  p->eps_tensor[0] = p->eps_tensor[4] = p->eps_tensor[8] = 1.0;
  p->source = 1; p->icomp_exact = 1;
}

/* scalar potential p(x) = c*sin(a x)sin(a y)sin(a z) with gradient and Hessian (mfd_solutions.F90:80-100) */
static void sin_potential(double a, zdouble c, const double x[3], zdouble *p, zdouble g[3], zdouble h[9]) {
  double s[3], co[3];
  for (int i = 0; i < 3; i++) { s[i] = sin(x[i] * a); co[i] = cos(x[i] * a); }
  *p = s[0] * s[1] * s[2] * c;
  g[0] = a * co[0] * s[1] * s[2] * c; g[1] = a * co[1] * s[0] * s[2] * c; g[2] = a * co[2] * s[0] * s[1] * c;
  double a2 = a * a;
  h[0] = -a2 * s[0] * s[1] * s[2] * c; h[4] = h[0]; h[8] = h[0];
  h[1] = h[3] = a2 * co[0] * co[1] * s[2] * c;
  h[2] = h[6] = a2 * co[0] * co[2] * s[1] * c;
  h[5] = h[7] = a2 * co[1] * co[2] * s[0] * c;
}
/* Poisson source f = -Laplace(u), u = sin(pi x)sin(pi y)sin(pi z) (POISSON/GALERKIN/common/getf.F90) */
static double poisson_source(const orc_params *prm, const double x[3], int l) {
  if (prm->source == 9) return ((const double *)prm->source_table)[l];
  if (prm->source != 1) return 0.0;
  zdouble p, g[3], h[9];
  sin_potential(M_PI, 1.0, x, &p, g, h);
  return -creal(h[0] + h[4] + h[8]);
}
/* curl curl of E = p e_ic :  grad(d_ic p) - Laplace(p) e_ic */
static void curlcurl_ic(const zdouble h[9], int ic, zdouble cc[3]) {
  zdouble lap = h[0] + h[4] + h[8];
  for (int c = 0; c < 3; c++) cc[c] = h[c + 3 * ic];
  cc[ic] -= lap;
}
/* UW Maxwell: J = curl H - i w eps E with H = curl E/(-i w mu)  (ULTRAWEAK_DPG/getf.F90, exact.F90) */
static void maxwell_uw_source(const orc_params *prm, const double x[3], int l, zdouble J[3]) {
  J[0] = J[1] = J[2] = 0;
  if (prm->source == 9) { for (int c = 0; c < 3; c++) J[c] = ((const zdouble *)prm->source_table)[3 * l + c]; return; }
  if (prm->source != 1) return;
  zdouble p, g[3], h[9], cc[3];
  int ic = prm->icomp_exact - 1;
  sin_potential(prm->omega, 1.0 + 1.0 * I, x, &p, g, h);
  curlcurl_ic(h, ic, cc);
  for (int c = 0; c < 3; c++) J[c] = cc[c] / (-I * prm->omega * prm->mu);
  J[ic] -= I * prm->omega * prm->eps * p;
}
/* Galerkin Maxwell: -i w J = curl(1/mu curl E) - (w^2 eps - i w sigma) E  (MAXWELL/GALERKIN/common/getf.F90:50-73) */
static void maxwell_gal_source(const orc_params *prm, const double x[3], int l, zdouble J[3]) {
  J[0] = J[1] = J[2] = 0;
  if (prm->source == 9) { for (int c = 0; c < 3; c++) J[c] = ((const zdouble *)prm->source_table)[3 * l + c]; return; }
  if (prm->source != 1) return;
  zdouble p, g[3], h[9], cc[3];
  int ic = prm->icomp_exact - 1;
  sin_potential(prm->omega, 1.0 + 1.0 * I, x, &p, g, h);
  curlcurl_ic(h, ic, cc);
  zdouble zb = prm->omega * prm->omega * prm->eps - I * prm->omega * prm->sigma;
  for (int c = 0; c < 3; c++) J[c] = cc[c] / prm->mu;
  J[ic] -= zb * p;
  for (int c = 0; c < 3; c++) J[c] = J[c] / (-I * prm->omega);
}

/* Piola maps as written in the element routines: F = J^-T Fhat, curl F = J Chat / det */
static void pull_grad(const double *ghat, const double dxidx[9], double out[3]) {
  for (int c = 0; c < 3; c++)
    out[c] = ghat[0] * dxidx[0 + 3 * c] + ghat[1] * dxidx[1 + 3 * c] + ghat[2] * dxidx[2 + 3 * c];
}
static void push_curl(const double *chat, const double dxdxi[9], double rjac, double out[3]) {
  for (int c = 0; c < 3; c++)
    out[c] = (dxdxi[c + 0] * chat[0] + dxdxi[c + 3] * chat[1] + dxdxi[c + 6] * chat[2]) / rjac;
}
static void check_jac(int iflag, double rjac) {
  if (iflag) { fprintf(stderr, "oracle: negative Jacobian %e\n", rjac); exit(1); }
}

/* ======================================================================= POISSON / GALERKIN */
int orc_elem_poisson_galerkin_t(int et, const int norder[19], const int norie[12], const int norif[6], const double *xnod,
                              const orc_params *prm, double *Aloc, double *Bloc, int *n_out) {
  int nH, nE, nV, nQ;
  orc_celndof(et, norder, &nH, &nE, &nV, &nQ);
  double *xiloc = xmalloc(sizeof(double) * 3 * 1000), *waloc = xmalloc(sizeof(double) * 1000);
  int nint = orc_set_3D_int(et, norder, norif, 0, orc_get_maxp(), xiloc, waloc);
  int nda = 3 * nint;
  double *AT = xmalloc(sizeof(double) * nH * nda);
  double *shapH = xmalloc(sizeof(double) * nH), *gradH = xmalloc(sizeof(double) * 3 * nH);
  for (int k = 0; k < nH; k++) Bloc[k] = 0.0;
  for (int l = 0; l < nint; l++) {
    double x[3], J[9], Ji[9], rjac; int iflag;
    orc_shape3DH(et, xiloc + 3 * l, norder, norie, norif, shapH, gradH);
    orc_geom3D(xnod, shapH, gradH, nH, x, J, Ji, &rjac, &iflag);
    check_jac(iflag, rjac);
    double weight = rjac * waloc[l], fval = poisson_source(prm, x, l), sw = sqrt(weight);
    for (int k = 0; k < nH; k++) {
      double dq[3];
      pull_grad(gradH + 3 * k, Ji, dq);
      Bloc[k] += shapH[k] * fval * weight;
      for (int c = 0; c < 3; c++) IDX(AT, nH, k, 3 * l + c) = dq[c] * sw;
    }
  }
  orc_dsyrk_u('N', nH, nda, 1.0, AT, nH, 0.0, Aloc, nH);
  for (int k = 0; k < nH; k++) for (int i = k + 1; i < nH; i++) IDX(Aloc, nH, i, k) = IDX(Aloc, nH, k, i);
  free(xiloc); free(waloc); free(AT); free(shapH); free(gradH);
  *n_out = nH;
  return 0;
}

/* ======================================================================= MAXWELL / GALERKIN */
int orc_elem_maxwell_galerkin_t(int et, const int norder[19], const int norie[12], const int norif[6], const double *xnod,
                              const orc_params *prm, zdouble *Aloc, zdouble *Bloc, int *n_out) {
  int nH, nE, nV, nQ;
  orc_celndof(et, norder, &nH, &nE, &nV, &nQ);
  double *xiloc = xmalloc(sizeof(double) * 3 * 1000), *waloc = xmalloc(sizeof(double) * 1000);
  int nint = orc_set_3D_int(et, norder, norif, 0, orc_get_maxp(), xiloc, waloc);
  int nda = 3 * nint;
  zdouble *AT = xmalloc(sizeof(zdouble) * nE * nda), *MT = xmalloc(sizeof(zdouble) * nE * nda);
  double *shapH = xmalloc(sizeof(double) * nH), *gradH = xmalloc(sizeof(double) * 3 * nH);
  double *shapE = xmalloc(sizeof(double) * 3 * nE), *curlE = xmalloc(sizeof(double) * 3 * nE);
  for (int k = 0; k < nE; k++) Bloc[k] = 0.0;
  zdouble zb = prm->omega * prm->omega * prm->eps - I * prm->omega * prm->sigma;
  for (int l = 0; l < nint; l++) {
    double x[3], J[9], Ji[9], rjac; int iflag;
    orc_shape3DH(et, xiloc + 3 * l, norder, norie, norif, shapH, gradH);
    orc_shape3DE(et, xiloc + 3 * l, norder, norie, norif, shapE, curlE);
    orc_geom3D(xnod, shapH, gradH, nH, x, J, Ji, &rjac, &iflag);
    check_jac(iflag, rjac);
    double weight = rjac * waloc[l];
    zdouble zJ[3];
    maxwell_gal_source(prm, x, l, zJ);
    zdouble sm = csqrt(zb * weight);
    double sa = sqrt(weight / prm->mu);
    for (int k = 0; k < nE; k++) {
      double F[3], CF[3];
      pull_grad(shapE + 3 * k, Ji, F);
      push_curl(curlE + 3 * k, J, rjac, CF);
      zdouble za = F[0] * zJ[0] + F[1] * zJ[1] + F[2] * zJ[2];
      Bloc[k] -= I * prm->omega * za * weight;
      for (int c = 0; c < 3; c++) { IDX(MT, nE, k, 3 * l + c) = F[c] * sm; IDX(AT, nE, k, 3 * l + c) = CF[c] * sa; }
    }
  }
  orc_zsyrk_u('N', nE, nda, 1.0, AT, nE, 0.0, Aloc, nE);
  orc_zsyrk_u('N', nE, nda, -1.0, MT, nE, 1.0, Aloc, nE);
  for (int k = 0; k < nE; k++) for (int i = k + 1; i < nE; i++) IDX(Aloc, nE, i, k) = IDX(Aloc, nE, k, i);
  free(xiloc); free(waloc); free(AT); free(MT); free(shapH); free(gradH); free(shapE); free(curlE);
  *n_out = nE;
  return 0;
}

/* ======================================================================= POISSON / PRIMAL DPG */
int orc_elem_poisson_primal_dpg_t(int et, const int norder[19], const int norie[12], const int norif[6],
                                const double *xnod, const orc_params *prm, double *Aloc, double *Bloc, int *nH_out,
                                int *nVi_out) {
  int nH, nE, nV, nQ, nHH, nEE, nVV, nQQ, bH, bE, bV, bQ, norderP[19];
  int dp = prm->nord_add, nordP = orc_enriched_mid(et, norder[MIDX(et)], dp);
  orc_compute_enriched_order(et, nordP, norderP);
  orc_celndof(et, norder, &nH, &nE, &nV, &nQ);
  orc_celndof(et, norderP, &nHH, &nEE, &nVV, &nQQ);
  orc_ndof_nod_mid(et, norder[MIDX(et)], &bH, &bE, &bV, &bQ);
  int nVi = nV - bV, nTest = nHH, nTrial = nH + nVi, maxpp = orc_get_maxp() + 1;
  double *xiloc = xmalloc(sizeof(double) * 3 * 1000), *waloc = xmalloc(sizeof(double) * 1000);
  int nint = orc_set_3D_int(et, norder, norif, dp, maxpp, xiloc, waloc);
  int nda = 3 * nint;
  double *testH = xmalloc(sizeof(double) * nHH * nint), *testGH = xmalloc(sizeof(double) * nHH * nda);
  double *trialGH = xmalloc(sizeof(double) * nH * nda), *bload = xmalloc(sizeof(double) * nTest);
  double *shapH = xmalloc(sizeof(double) * nH), *gradH = xmalloc(sizeof(double) * 3 * nH);
  double *shapHH = xmalloc(sizeof(double) * nHH), *gradHH = xmalloc(sizeof(double) * 3 * nHH);
  double *shapV = xmalloc(sizeof(double) * 3 * nV), *divV = xmalloc(sizeof(double) * nV);
  for (int l = 0; l < nint; l++) {
    double x[3], J[9], Ji[9], rjac; int iflag;
    orc_shape3DH(et, xiloc + 3 * l, norder, norie, norif, shapH, gradH);
    orc_shape3HH(et, xiloc + 3 * l, nordP, shapHH, gradHH);
    orc_geom3D(xnod, shapH, gradH, nH, x, J, Ji, &rjac, &iflag);
    check_jac(iflag, rjac);
    double weight = rjac * waloc[l], sw = sqrt(weight), fval = poisson_source(prm, x, l);
    for (int k = 0; k < nHH; k++) {
      double dv[3];
      bload[k] += fval * shapHH[k] * weight;
      IDX(testH, nHH, k, l) = shapHH[k] * sw;
      pull_grad(gradHH + 3 * k, Ji, dv);
      for (int c = 0; c < 3; c++) IDX(testGH, nHH, k, 3 * l + c) = dv[c] * sw;
    }
    for (int k = 0; k < nH; k++) {
      double dpv[3];
      pull_grad(gradH + 3 * k, Ji, dpv);
      for (int c = 0; c < 3; c++) IDX(trialGH, nH, k, 3 * l + c) = dpv[c] * sw;
    }
  }
  /* Gram (upper) = (grad v, grad q) + (v, q); stiffness = (grad u, grad v) */
  double *gram = xmalloc(sizeof(double) * nTest * nTest);
  double *stiff = xmalloc(sizeof(double) * nTest * (nTrial + 1));
  orc_dsyrk_u('N', nHH, nda, 1.0, testGH, nHH, 0.0, gram, nTest);
  orc_dsyrk_u('N', nHH, nint, 1.0, testH, nHH, 1.0, gram, nTest);
  orc_dgemm('N', 'T', nHH, nH, nda, 1.0, testGH, nHH, trialGH, nH, 0.0, stiff, nTest);
  /* boundary: -<sigma.n, v> */
  double *stiffHV = xmalloc(sizeof(double) * nTest * (nVi ? nVi : 1));
  double *tH = xmalloc(sizeof(double) * nHH * 100), *tV = xmalloc(sizeof(double) * (nVi ? nVi : 1) * 100);
  int noff = 0;
  for (int ifc = 1; ifc <= orc_nface(et); ifc++) {
    int nordf[5], nord_ifc[19], fh, fe, fv, fq;
    double tloc[200], wtloc[100];
    int nsign = orc_nsign_param(et, ifc);
    orc_face_order(et, ifc, norder, nordf);
    int nintf = orc_set_2D_int(orc_face_is_tri(et, ifc), nordf, norif[ifc - 1], dp, maxpp, tloc, wtloc);
    orc_ndof_nod_face(et, ifc, norder[orc_nedge(et) + ifc - 1], &fh, &fe, &fv, &fq);
    orc_initiate_order(et, nord_ifc);                            /* initiate_order + edges (elem_opt.F90:322-324) */
    for (int i = 0; i < orc_nedge(et); i++) nord_ifc[i] = norder[i];
    nord_ifc[orc_nedge(et) + ifc - 1] = norder[orc_nedge(et) + ifc - 1];
    memset(tV, 0, sizeof(double) * (nVi ? nVi : 1) * 100);
    for (int l = 0; l < nintf; l++) {
      double xi[3], dxidt[6], x[3], J[9], Ji[9], rjac, dxdt[6], rn[3], bjac;
      orc_face_param(et, ifc, tloc + 2 * l, xi, dxidt);
      orc_shape3HH(et, xi, nordP, shapHH, gradHH);
      orc_shape3DV(et, xi, nord_ifc, norif, shapV, divV);
      orc_shape3DH(et, xi, norder, norie, norif, shapH, gradH);
      orc_bgeom3D(xnod, shapH, gradH, nH, dxidt, nsign, x, J, Ji, &rjac, dxdt, rn, &bjac);
      double weight = bjac * wtloc[l], sw = sqrt(weight);
      for (int k = 0; k < nHH; k++) IDX(tH, nHH, k, l) = shapHH[k] * sw;
      for (int i = 0; i < fv; i++) {
        int k = (ifc - 1) + i;
        double s[3];
        push_curl(shapV + 3 * k, J, rjac, s); /* same contravariant Piola: J V / det */
        double sn = s[0] * rn[0] + s[1] * rn[1] + s[2] * rn[2];
        IDX(tV, nVi, noff + i, l) = sn * sw;
      }
    }
    orc_dgemm('N', 'T', nHH, nVi, nintf, -1.0, tH, nHH, tV, nVi, (ifc == 1) ? 0.0 : 1.0, stiffHV, nTest);
    noff += fv;
  }
  for (int j = 0; j < nVi; j++) for (int i = 0; i < nTest; i++) IDX(stiff, nTest, i, nH + j) = IDX(stiffHV, nTest, i, j);
  for (int i = 0; i < nTest; i++) IDX(stiff, nTest, i, nTrial) = bload[i];
  int info = orc_dpotrf_u(nTest, gram, nTest);
  if (info) { fprintf(stderr, "oracle primal DPG: POTRF info=%d\n", info); return info; }
  orc_dtrsm_u('T', nTest, nTrial + 1, gram, nTest, stiff, nTest);
  double *ral = xmalloc(sizeof(double) * (nTrial + 1) * (nTrial + 1));
  orc_dsyrk_u('T', nTrial + 1, nTest, 1.0, stiff, nTest, 0.0, ral, nTrial + 1);
  for (int j = 0; j < nTrial; j++) {
    for (int i = 0; i < nTrial; i++) IDX(Aloc, nTrial, i, j) = (i <= j) ? IDX(ral, nTrial + 1, i, j) : IDX(ral, nTrial + 1, j, i);
    Bloc[j] = IDX(ral, nTrial + 1, j, nTrial);
  }
  free(xiloc); free(waloc); free(testH); free(testGH); free(trialGH); free(bload); free(shapH); free(gradH);
  free(shapHH); free(gradHH); free(shapV); free(divV); free(gram); free(stiff); free(stiffHV); free(tH); free(tV); free(ral);
  *nH_out = nH; *nVi_out = nVi;
  return 0;
}

/* ======================================================================= MAXWELL / ULTRAWEAK DPG */
int orc_elem_maxwell_uw_dpg_t(int et, const int norder[19], const int norie[12], const int norif[6], const double *xnod,
                            const orc_params *prm, zdouble *Aloc, zdouble *Bloc, int *nEi_out, int *nQ_out,
                            zdouble *gram_out, zdouble *stiff_out) {
  int nH, nE, nV, nQ, nHH, nEE, nVV, nQQ, bH, bE, bV, bQ, norderP[19];
  int dp = prm->nord_add, nordP = orc_enriched_mid(et, norder[MIDX(et)], dp), maxpp = orc_get_maxp() + 1;
  orc_compute_enriched_order(et, nordP, norderP);
  orc_celndof(et, norder, &nH, &nE, &nV, &nQ);
  orc_celndof(et, norderP, &nHH, &nEE, &nVV, &nQQ);
  orc_ndof_nod_mid(et, norder[MIDX(et)], &bH, &bE, &bV, &bQ);
  int nEi = nE - bE, nTest = 2 * nEE, nTrial = 2 * nEi + 6 * nQ;
  double *xiloc = xmalloc(sizeof(double) * 3 * 1000), *waloc = xmalloc(sizeof(double) * 1000);
  int nint = orc_set_3D_int(et, norder, norif, dp, maxpp, xiloc, waloc);
  int nda = 3 * nint, n3Q = 3 * nQ;
  double *test_rE = xmalloc(sizeof(double) * nEE * nda), *test_rCE = xmalloc(sizeof(double) * nEE * nda);
  double *trial_rQ = xmalloc(sizeof(double) * n3Q * nda);
  double *d_rEPS = xmalloc(sizeof(double) * 3 * nda), *d_iEPS = xmalloc(sizeof(double) * 3 * nda);
  double *d_rMU = xmalloc(sizeof(double) * 3 * nda), *d_iMU = xmalloc(sizeof(double) * 3 * nda);
  zdouble *bload = xmalloc(sizeof(zdouble) * nTest);
  double *shapH = xmalloc(sizeof(double) * nH), *gradH = xmalloc(sizeof(double) * 3 * nH);
  double *shapQ = xmalloc(sizeof(double) * nQ);
  double *shapEE = xmalloc(sizeof(double) * 3 * nEE), *curlEE = xmalloc(sizeof(double) * 3 * nEE);
  double *shapE = xmalloc(sizeof(double) * 3 * nE), *curlE = xmalloc(sizeof(double) * 3 * nE);
  /* ---- volume loop: elem_opt.F90:236-325 */
  for (int l = 0; l < nint; l++) {
    double x[3], J[9], Ji[9], rjac; int iflag;
    orc_shape3DH(et, xiloc + 3 * l, norder, norie, norif, shapH, gradH);
    orc_shape3DQ(et, xiloc + 3 * l, norder, shapQ);
    orc_shape3EE(et, xiloc + 3 * l, nordP, shapEE, curlEE);
    orc_geom3D(xnod, shapH, gradH, nH, x, J, Ji, &rjac, &iflag);
    check_jac(iflag, rjac);
    for (int j = 0; j < 3; j++)
      for (int i = 0; i < 3; i++) {
        zdouble za = (I * prm->omega * prm->eps) * prm->eps_tensor[i + 3 * j];
        zdouble zc = (I * prm->omega * prm->mu) * ((i == j) ? 1.0 : 0.0);
        IDX(d_rEPS, 3, i, 3 * l + j) = creal(za); IDX(d_iEPS, 3, i, 3 * l + j) = cimag(za);
        IDX(d_rMU, 3, i, 3 * l + j) = creal(zc);  IDX(d_iMU, 3, i, 3 * l + j) = cimag(zc);
      }
    double weight = rjac * waloc[l], sw = sqrt(weight);
    zdouble zJ[3];
    maxwell_uw_source(prm, x, l, zJ);
    for (int k = 0; k < nQ; k++) {
      double u = shapQ[k] / rjac;
      for (int c = 0; c < 3; c++) IDX(trial_rQ, n3Q, 3 * k + c, 3 * l + c) = u * sw;
    }
    for (int k = 0; k < nEE; k++) {
      double F[3], C[3];
      pull_grad(shapEE + 3 * k, Ji, F);
      push_curl(curlEE + 3 * k, J, rjac, C);
      bload[2 * k] += (F[0] * zJ[0] + F[1] * zJ[1] + F[2] * zJ[2]) * weight;
      for (int c = 0; c < 3; c++) { IDX(test_rE, nEE, k, 3 * l + c) = F[c] * sw; IDX(test_rCE, nEE, k, 3 * l + c) = C[c] * sw; }
    }
  }
  /* ---- stiffness blocks: elem_opt.F90:339-379 */
  double *st_rFE = xmalloc(sizeof(double) * nEE * n3Q), *st_iFE = xmalloc(sizeof(double) * nEE * n3Q);
  double *st_rFH = xmalloc(sizeof(double) * nEE * n3Q);
  double *st_rGH = xmalloc(sizeof(double) * nEE * n3Q), *st_iGH = xmalloc(sizeof(double) * nEE * n3Q);
  double *tmp_rQ = xmalloc(sizeof(double) * n3Q * nda), *tmp_iQ = xmalloc(sizeof(double) * n3Q * nda);
  for (int l = 0; l < nint; l++) {
    orc_dgemm('N', 'T', n3Q, 3, 3, 1.0, trial_rQ + (size_t)n3Q * 3 * l, n3Q, d_rEPS + 9 * l, 3, 0.0, tmp_rQ + (size_t)n3Q * 3 * l, n3Q);
    orc_dgemm('N', 'T', n3Q, 3, 3, 1.0, trial_rQ + (size_t)n3Q * 3 * l, n3Q, d_iEPS + 9 * l, 3, 0.0, tmp_iQ + (size_t)n3Q * 3 * l, n3Q);
  }
  orc_dgemm('N', 'T', nEE, n3Q, nda, -1.0, test_rE, nEE, tmp_rQ, n3Q, 0.0, st_rFE, nEE);
  orc_dgemm('N', 'T', nEE, n3Q, nda, -1.0, test_rE, nEE, tmp_iQ, n3Q, 0.0, st_iFE, nEE);
  orc_dgemm('N', 'T', nEE, n3Q, nda, 1.0, test_rCE, nEE, trial_rQ, n3Q, 0.0, st_rFH, nEE);
  for (int l = 0; l < nint; l++) {
    orc_dgemm('N', 'T', n3Q, 3, 3, 1.0, trial_rQ + (size_t)n3Q * 3 * l, n3Q, d_rMU + 9 * l, 3, 0.0, tmp_rQ + (size_t)n3Q * 3 * l, n3Q);
    orc_dgemm('N', 'T', n3Q, 3, 3, 1.0, trial_rQ + (size_t)n3Q * 3 * l, n3Q, d_iMU + 9 * l, 3, 0.0, tmp_iQ + (size_t)n3Q * 3 * l, n3Q);
  }
  orc_dgemm('N', 'T', nEE, n3Q, nda, 1.0, test_rE, nEE, tmp_rQ, n3Q, 0.0, st_rGH, nEE);
  orc_dgemm('N', 'T', nEE, n3Q, nda, 1.0, test_rE, nEE, tmp_iQ, n3Q, 0.0, st_iGH, nEE);
  free(trial_rQ); free(tmp_rQ); free(tmp_iQ);
  /* ---- Gram: elem_opt.F90:385-470 */
  double afac = (prm->test_norm == 2) ? 1.0 : prm->alpha_norm;
  double *gram_r = xmalloc(sizeof(double) * nEE * nEE);
  orc_dsyrk_u('N', nEE, nda, afac, test_rE, nEE, 0.0, gram_r, nEE);
  orc_dsyrk_u('N', nEE, nda, 1.0, test_rCE, nEE, 1.0, gram_r, nEE);
  zdouble *gFF = NULL, *gGG = NULL, *gFG = NULL;
  if (prm->test_norm != 2) {
    gFF = xmalloc(sizeof(zdouble) * nEE * nEE); gGG = xmalloc(sizeof(zdouble) * nEE * nEE);
    for (size_t i = 0; i < (size_t)nEE * nEE; i++) { gFF[i] = gram_r[i]; gGG[i] = gram_r[i]; }
    double *t_r = xmalloc(sizeof(double) * nEE * nda), *t_i = xmalloc(sizeof(double) * nEE * nda);
    zdouble *t_E = xmalloc(sizeof(zdouble) * nEE * nda);
    double *gram_i = NULL;
    for (int l = 0; l < nint; l++) {
      orc_dgemm('N', 'N', nEE, 3, 3, 1.0, test_rE + (size_t)nEE * 3 * l, nEE, d_rEPS + 9 * l, 3, 0.0, t_r + (size_t)nEE * 3 * l, nEE);
      orc_dgemm('N', 'N', nEE, 3, 3, 1.0, test_rE + (size_t)nEE * 3 * l, nEE, d_iEPS + 9 * l, 3, 0.0, t_i + (size_t)nEE * 3 * l, nEE);
    }
    for (size_t i = 0; i < (size_t)nEE * nda; i++) t_E[i] = t_r[i] + I * t_i[i];
    orc_zherk_u('N', nEE, nda, 1.0, t_E, nEE, 1.0, gFF, nEE);
    if (prm->test_norm != 3) {
      gram_i = xmalloc(sizeof(double) * nEE * nEE);
      orc_dgemm('N', 'T', nEE, nEE, nda, -1.0, t_r, nEE, test_rCE, nEE, 0.0, gram_r, nEE);
      orc_dgemm('N', 'T', nEE, nEE, nda, -1.0, t_i, nEE, test_rCE, nEE, 0.0, gram_i, nEE);
    }
    for (int l = 0; l < nint; l++) {
      orc_dgemm('N', 'N', nEE, 3, 3, 1.0, test_rE + (size_t)nEE * 3 * l, nEE, d_rMU + 9 * l, 3, 0.0, t_r + (size_t)nEE * 3 * l, nEE);
      orc_dgemm('N', 'N', nEE, 3, 3, 1.0, test_rE + (size_t)nEE * 3 * l, nEE, d_iMU + 9 * l, 3, 0.0, t_i + (size_t)nEE * 3 * l, nEE);
    }
    for (size_t i = 0; i < (size_t)nEE * nda; i++) t_E[i] = t_r[i] + I * t_i[i];
    orc_zherk_u('N', nEE, nda, 1.0, t_E, nEE, 1.0, gGG, nEE);
    if (prm->test_norm != 3) {
      orc_dgemm('N', 'T', nEE, nEE, nda, 1.0, test_rCE, nEE, t_r, nEE, 1.0, gram_r, nEE);
      orc_dgemm('N', 'T', nEE, nEE, nda, -1.0, test_rCE, nEE, t_i, nEE, 1.0, gram_i, nEE);
      gFG = xmalloc(sizeof(zdouble) * nEE * nEE);
      for (size_t i = 0; i < (size_t)nEE * nEE; i++) gFG[i] = gram_r[i] + I * gram_i[i];
      free(gram_i);
    }
    free(t_r); free(t_i); free(t_E);
  }
  free(d_rEPS); free(d_iEPS); free(d_rMU); free(d_iMU); free(test_rCE);
  /* ---- boundary integrals: elem_opt.F90:492-666 (impedance BC not supported there either: it stops) */
  double *st_rEE = xmalloc(sizeof(double) * nEE * nEi);
  double *t_rE = xmalloc(sizeof(double) * nEE * 300), *t_rnE = xmalloc(sizeof(double) * nEi * 300);
  int noffE = 0;
  for (int ifc = 1; ifc <= orc_nface(et); ifc++) {
    int nordf[5], nord_ifc[19], fh, fe, fv, fq;
    double tloc[200], wtloc[100];
    int nsign = orc_nsign_param(et, ifc);
    orc_face_order(et, ifc, norder, nordf);
    int nintf = orc_set_2D_int(orc_face_is_tri(et, ifc), nordf, norif[ifc - 1], dp, maxpp, tloc, wtloc);
    orc_ndof_nod_face(et, ifc, norder[orc_nedge(et) + ifc - 1], &fh, &fe, &fv, &fq);
    orc_initiate_order(et, nord_ifc);                            /* initiate_order + edges (elem_opt.F90:322-324) */
    for (int i = 0; i < orc_nedge(et); i++) nord_ifc[i] = norder[i];
    nord_ifc[orc_nedge(et) + ifc - 1] = norder[orc_nedge(et) + ifc - 1];
    memset(t_rnE, 0, sizeof(double) * nEi * 300);
    for (int l = 0; l < nintf; l++) {
      double xi[3], dxidt[6], x[3], J[9], Ji[9], rjac, dxdt[6], rn[3], bjac;
      orc_face_param(et, ifc, tloc + 2 * l, xi, dxidt);
      orc_shape3EE(et, xi, nordP, shapEE, curlEE);
      orc_shape3DH(et, xi, norder, norie, norif, shapH, gradH);
      orc_bgeom3D(xnod, shapH, gradH, nH, dxidt, nsign, x, J, Ji, &rjac, dxdt, rn, &bjac);
      double weight = bjac * wtloc[l], sw = sqrt(weight);
      int nE_ifc = orc_shape3DE(et, xi, nord_ifc, norie, norif, shapE, curlE);
      for (int k = 0; k < nEE; k++) {
        double F[3];
        pull_grad(shapEE + 3 * k, Ji, F);
        for (int c = 0; c < 3; c++) IDX(t_rE, nEE, k, 3 * l + c) = F[c] * sw;
      }
      int nedge_fn = nE_ifc - fe;
      for (int k = 0; k < nE_ifc; k++) {
        double E2[3], rxE[3];
        pull_grad(shapE + 3 * k, Ji, E2);
        rxE[0] = rn[1] * E2[2] - rn[2] * E2[1];
        rxE[1] = rn[2] * E2[0] - rn[0] * E2[2];
        rxE[2] = rn[0] * E2[1] - rn[1] * E2[0];
        int k2 = (k < nedge_fn) ? k : noffE + k;
        for (int c = 0; c < 3; c++) IDX(t_rnE, nEi, k2, 3 * l + c) = rxE[c] * sw;
      }
    }
    noffE += fe;
    orc_dgemm('N', 'T', nEE, nEi, 3 * nintf, 1.0, t_rE, nEE, t_rnE, nEi, (ifc == 1) ? 0.0 : 1.0, st_rEE, nEE);
  }
  free(t_rE); free(t_rnE); free(test_rE);
  /* ---- DPG system, interleaved ordering (blocks=.false.): elem_opt.F90:684-768 */
  int jE = 2 * nEi, jQ = 6 * nQ, ncol = nTrial + 1;
  zdouble *stiff = xmalloc(sizeof(zdouble) * nTest * ncol);
  for (int j = 0; j < nEi; j++)
    for (int i = 0; i < nEE; i++) {
      IDX(stiff, nTest, 2 * i, 2 * j + 1) = IDX(st_rEE, nEE, i, j);   /* <F, n x H> */
      IDX(stiff, nTest, 2 * i + 1, 2 * j) = IDX(st_rEE, nEE, i, j);   /* <G, n x E> */
    }
  for (int j = 0; j < nQ; j++)
    for (int i = 0; i < nEE; i++)
      for (int c = 0; c < 3; c++) {
        IDX(stiff, nTest, 2 * i, jE + 6 * j + c) = IDX(st_rFE, nEE, i, 3 * j + c) + I * IDX(st_iFE, nEE, i, 3 * j + c);
        IDX(stiff, nTest, 2 * i, jE + 6 * j + 3 + c) = IDX(st_rFH, nEE, i, 3 * j + c);
        IDX(stiff, nTest, 2 * i + 1, jE + 6 * j + c) = IDX(st_rFH, nEE, i, 3 * j + c);
        IDX(stiff, nTest, 2 * i + 1, jE + 6 * j + 3 + c) = IDX(st_rGH, nEE, i, 3 * j + c) + I * IDX(st_iGH, nEE, i, 3 * j + c);
      }
  for (int i = 0; i < nTest; i++) IDX(stiff, nTest, i, jE + jQ) = bload[i];
  free(st_rFE); free(st_iFE); free(st_rFH); free(st_rGH); free(st_iGH); free(st_rEE);
  /* ---- Gram assembly + Cholesky: elem_opt.F90:771-850 */
  zdouble *gram = xmalloc(sizeof(zdouble) * nTest * nTest);
  int info = 0;
  if (prm->test_norm == 2) { /* MATH_NORM: factor the real block first, then interleave the factor */
    info = orc_dpotrf_u(nEE, gram_r, nEE);
    for (int j = 0; j < nEE; j++) for (int i = 0; i <= j; i++) {
      IDX(gram, nTest, 2 * i, 2 * j) = IDX(gram_r, nEE, i, j); IDX(gram, nTest, 2 * i + 1, 2 * j + 1) = IDX(gram_r, nEE, i, j); }
  } else if (prm->test_norm == 3) { /* GRAPH_DIAG */
    info = orc_zpotrf_u(nEE, gFF, nEE);
    if (!info) info = orc_zpotrf_u(nEE, gGG, nEE);
    for (int j = 0; j < nEE; j++) for (int i = 0; i <= j; i++) {
      IDX(gram, nTest, 2 * i, 2 * j) = IDX(gFF, nEE, i, j); IDX(gram, nTest, 2 * i + 1, 2 * j + 1) = IDX(gGG, nEE, i, j); }
  } else { /* GRAPH_NORM */
    for (int j = 0; j < nEE; j++) for (int i = 0; i <= j; i++) {
      IDX(gram, nTest, 2 * i, 2 * j) = IDX(gFF, nEE, i, j);
      IDX(gram, nTest, 2 * i + 1, 2 * j + 1) = IDX(gGG, nEE, i, j);
      IDX(gram, nTest, 2 * i, 2 * j + 1) = IDX(gFG, nEE, i, j);
      IDX(gram, nTest, 2 * i + 1, 2 * j) = conj(IDX(gFG, nEE, j, i));
    }
    if (gram_out) memcpy(gram_out, gram, sizeof(zdouble) * nTest * nTest);
    info = orc_zpotrf_u(nTest, gram, nTest);
  }
  if (stiff_out) memcpy(stiff_out, stiff, sizeof(zdouble) * nTest * ncol);
  free(gram_r); free(gFF); free(gGG); free(gFG);
  if (info) { fprintf(stderr, "oracle UW DPG: POTRF info=%d\n", info); free(gram); free(stiff); return info; }
  /* ---- B~ = U^-H [B|l] ; [A|b] = B~^H B~ : elem_opt.F90:851-869 */
  orc_ztrsm_u('C', nTest, ncol, gram, nTest, stiff, nTest);
  free(gram);
  zdouble *zal = xmalloc(sizeof(zdouble) * ncol * ncol);
  orc_zherk_u('C', ncol, nTest, 1.0, stiff, nTest, 0.0, zal, ncol);
  free(stiff);
  for (int j = 0; j < nTrial; j++) {
    for (int i = 0; i < nTrial; i++) IDX(Aloc, nTrial, i, j) = (i <= j) ? IDX(zal, ncol, i, j) : conj(IDX(zal, ncol, j, i));
    Bloc[j] = IDX(zal, ncol, j, nTrial);
  }
  free(zal); free(xiloc); free(waloc); free(bload); free(shapH); free(gradH); free(shapQ); free(shapEE); free(curlEE);
  free(shapE); free(curlE);
  *nEi_out = nEi; *nQ_out = nQ;
  return 0;
}

/* ======================================================================= static condensation */
int orc_stc_fwd_real(int herm, int ni, int nb, double *Aii, double *Abi, double *Aib, double *Abb, double *Bi, double *Bb) {
  int info = 0;
  if (herm) { /* stc.F90:338-414 : Cholesky (RFP there), two POTRS, two GEMM */
    info = orc_dpotrf_u(nb, Abb, nb);
    if (info) return info;
    orc_dtrsm_u('T', nb, 1, Abb, nb, Bb, nb);  orc_dtrsm_u('N', nb, 1, Abb, nb, Bb, nb);
    orc_dtrsm_u('T', nb, ni, Abb, nb, Abi, nb); orc_dtrsm_u('N', nb, ni, Abb, nb, Abi, nb);
  } else {    /* stc.F90:443-507 : pivoted LU */
    int *piv = xmalloc(sizeof(int) * nb);
    info = orc_dgetrf(nb, Abb, nb, piv);
    if (info) { free(piv); return info; }
    orc_dgetrs(nb, 1, Abb, nb, piv, Bb, nb);
    orc_dgetrs(nb, ni, Abb, nb, piv, Abi, nb);
    free(piv);
  }
  orc_dgemm('N', 'N', ni, 1, nb, -1.0, Aib, ni, Bb, nb, 1.0, Bi, ni);
  orc_dgemm('N', 'N', ni, ni, nb, -1.0, Aib, ni, Abi, nb, 1.0, Aii, ni);
  return 0;
}
int orc_stc_fwd_cplx(int herm, int ni, int nb, zdouble *Aii, zdouble *Abi, zdouble *Aib, zdouble *Abb, zdouble *Bi, zdouble *Bb) {
  int info = 0;
  if (herm) {
    info = orc_zpotrf_u(nb, Abb, nb);
    if (info) return info;
    orc_ztrsm_u('C', nb, 1, Abb, nb, Bb, nb);  orc_ztrsm_u('N', nb, 1, Abb, nb, Bb, nb);
    orc_ztrsm_u('C', nb, ni, Abb, nb, Abi, nb); orc_ztrsm_u('N', nb, ni, Abb, nb, Abi, nb);
  } else {
    int *piv = xmalloc(sizeof(int) * nb);
    info = orc_zgetrf(nb, Abb, nb, piv);
    if (info) { free(piv); return info; }
    orc_zgetrs(nb, 1, Abb, nb, piv, Bb, nb);
    orc_zgetrs(nb, ni, Abb, nb, piv, Abi, nb);
    free(piv);
  }
  orc_zgemm('N', 'N', ni, 1, nb, -1.0, Aib, ni, Bb, nb, 1.0, Bi, ni);
  orc_zgemm('N', 'N', ni, ni, nb, -1.0, Aib, ni, Abi, nb, 1.0, Aii, ni);
  return 0;
}

/* stc_get_nrdof + the gather order of stc_fwd_wrapper (stc.F90:94-170, :226-261) for the four problems:
 *  1 POIS_GAL : one H1 variable              -> interface = non-middle-node dofs (already first)
 *  2 POIS_PDPG: H1 (PHYSAi=F) + H(div) trace (PHYSAi=T) -> [H1 interface | trace] then [H1 bubbles]
 *  3 MAXW_GAL : one H(curl) variable
 *  4 MAXW_UW  : H(curl) traces x2 (PHYSAi=T) + L2 x6 (all bubble) */
int orc_stc_partition_t(int et, int kind, const int norder[19], int *perm, int *ni_out, int *nb_out) {
  int nH, nE, nV, nQ, bH, bE, bV, bQ, ni = 0, nb = 0, n = 0;
  orc_celndof(et, norder, &nH, &nE, &nV, &nQ);
  orc_ndof_nod_mid(et, norder[MIDX(et)], &bH, &bE, &bV, &bQ);
  switch (kind) {
    case 1: ni = nH - bH; nb = bH; n = nH; for (int i = 0; i < n; i++) perm[i] = i; break;
    case 3: ni = nE - bE; nb = bE; n = nE; for (int i = 0; i < n; i++) perm[i] = i; break;
    case 2: {
      int iH = nH - bH, nVi = nV - bV, k = 0;
      ni = iH + nVi; nb = bH;
      for (int i = 0; i < iH; i++) perm[k++] = i;
      for (int i = 0; i < nVi; i++) perm[k++] = nH + i;
      for (int i = 0; i < bH; i++) perm[k++] = iH + i;
      break; }
    case 4: ni = 2 * (nE - bE); nb = 6 * nQ; n = ni + nb; for (int i = 0; i < n; i++) perm[i] = i; break;
    default: return -1;
  }
  *ni_out = ni; *nb_out = nb;
  return 0;
}

int orc_condensed_element_t(int et, int kind, const int norder[19], const int norie[12], const int norif[6], const double *xnod,
                          const orc_params *prm, void *Aii_o, void *Bi_o, void *AS_o, void *BS_o, int *ni_o, int *nb_o) {
  int ni, nb, info = 0;
  int *perm = xmalloc(sizeof(int) * 8192);
  if (orc_stc_partition_t(et, kind, norder, perm, &ni, &nb)) { free(perm); return -1; }
  int n = ni + nb, cplx = (kind >= 3), herm = (kind == 2 || kind == 4);
  size_t es = cplx ? sizeof(zdouble) : sizeof(double);
  void *A = xmalloc(es * n * n), *b = xmalloc(es * n);
  int a1, a2;
  switch (kind) {
    case 1: info = orc_elem_poisson_galerkin_t(et, norder, norie, norif, xnod, prm, A, b, &a1); break;
    case 2: info = orc_elem_poisson_primal_dpg_t(et, norder, norie, norif, xnod, prm, A, b, &a1, &a2); break;
    case 3: info = orc_elem_maxwell_galerkin_t(et, norder, norie, norif, xnod, prm, A, b, &a1); break;
    case 4: info = orc_elem_maxwell_uw_dpg_t(et, norder, norie, norif, xnod, prm, A, b, &a1, &a2, NULL, NULL); break;
  }
  if (info) { free(A); free(b); free(perm); return info; }
  void *Abb = xmalloc(es * nb * nb), *Aib = xmalloc(es * ni * nb);
#define GATHER(T)                                                                                          \
  do {                                                                                                     \
    T *Af = A, *bf = b, *Aii = Aii_o, *Bi = Bi_o, *AS = AS_o, *BS = BS_o, *Abb_ = Abb, *Aib_ = Aib;        \
    for (int j = 0; j < ni; j++) { for (int i = 0; i < ni; i++) IDX(Aii, ni, i, j) = IDX(Af, n, perm[i], perm[j]); \
                                   for (int i = 0; i < nb; i++) IDX(AS, nb, i, j) = IDX(Af, n, perm[ni + i], perm[j]); } \
    for (int j = 0; j < nb; j++) { for (int i = 0; i < ni; i++) IDX(Aib_, ni, i, j) = IDX(Af, n, perm[i], perm[ni + j]); \
                                   for (int i = 0; i < nb; i++) IDX(Abb_, nb, i, j) = IDX(Af, n, perm[ni + i], perm[ni + j]); } \
    for (int i = 0; i < ni; i++) Bi[i] = bf[perm[i]];                                                      \
    for (int i = 0; i < nb; i++) BS[i] = bf[perm[ni + i]];                                                 \
  } while (0)
  if (nb > 0) {
    if (cplx) { GATHER(zdouble); info = orc_stc_fwd_cplx(herm, ni, nb, Aii_o, AS_o, Aib, Abb, Bi_o, BS_o); }
    else      { GATHER(double);  info = orc_stc_fwd_real(herm, ni, nb, Aii_o, AS_o, Aib, Abb, Bi_o, BS_o); }
  } else {
    if (cplx) GATHER(zdouble); else GATHER(double);
  }
  free(A); free(b); free(Abb); free(Aib); free(perm);
  *ni_o = ni; *nb_o = nb;
  return info;
}

/* Element loop with the structure of par_mumps_sc.F90:318-357 (!$OMP DO SCHEDULE(DYNAMIC) over the
 * subdomain's elements, thread-private workspaces).  Outputs are packed per element with the given
 * strides (in scalars).  Returns the number of elements whose info != 0. */
int orc_condensed_batch_t(int kind, int nel, const int *etype, const int *norder, const int *norie, const int *norif, const double *xnod,
                        int xnod_stride, const orc_params *prm, void *Aii, void *Bi, void *ASchur, void *BSchur,
                        long sAii, long sBi, long sAS, long sBS, int *info, int nthreads) {
  int cplx = (kind >= 3), bad = 0;
  size_t es = cplx ? sizeof(zdouble) : sizeof(double);
#pragma omp parallel for schedule(dynamic) num_threads(nthreads) reduction(+ : bad)
  for (int e = 0; e < nel; e++) {
    int ni, nb;
    int r = orc_condensed_element_t(etype ? etype[e] : ORC_MDLB, kind, norder + 19 * e, norie + 12 * e, norif + 6 * e, xnod + (size_t)xnod_stride * e, prm,
                                  (char *)Aii + es * sAii * e, (char *)Bi + es * sBi * e, (char *)ASchur + es * sAS * e,
                                  (char *)BSchur + es * sBS * e, &ni, &nb);
    if (info) info[e] = r;
    if (r) bad++;
  }
  return bad;
}

/* ---- brick-only entry points kept for the existing callers */
int orc_elem_poisson_galerkin(const int norder[19], const int norie[12], const int norif[6], const double *xnod,
                              const orc_params *prm, double *Aloc, double *Bloc, int *n) {
  return orc_elem_poisson_galerkin_t(ORC_MDLB, norder, norie, norif, xnod, prm, Aloc, Bloc, n);
}
int orc_elem_poisson_primal_dpg(const int norder[19], const int norie[12], const int norif[6], const double *xnod,
                                const orc_params *prm, double *Aloc, double *Bloc, int *nH, int *nVi) {
  return orc_elem_poisson_primal_dpg_t(ORC_MDLB, norder, norie, norif, xnod, prm, Aloc, Bloc, nH, nVi);
}
int orc_elem_maxwell_galerkin(const int norder[19], const int norie[12], const int norif[6], const double *xnod,
                              const orc_params *prm, zdouble *Aloc, zdouble *Bloc, int *n) {
  return orc_elem_maxwell_galerkin_t(ORC_MDLB, norder, norie, norif, xnod, prm, Aloc, Bloc, n);
}
int orc_elem_maxwell_uw_dpg(const int norder[19], const int norie[12], const int norif[6], const double *xnod,
                            const orc_params *prm, zdouble *Aloc, zdouble *Bloc, int *nEi, int *nQ, zdouble *gram_out,
                            zdouble *stiff_out) {
  return orc_elem_maxwell_uw_dpg_t(ORC_MDLB, norder, norie, norif, xnod, prm, Aloc, Bloc, nEi, nQ, gram_out, stiff_out);
}
int orc_stc_partition(int kind, const int norder[19], int *perm, int *ni, int *nb) {
  return orc_stc_partition_t(ORC_MDLB, kind, norder, perm, ni, nb);
}
int orc_condensed_element(int kind, const int norder[19], const int norie[12], const int norif[6], const double *xnod,
                          const orc_params *prm, void *Aii, void *Bi, void *ASchur, void *BSchur, int *ni, int *nb) {
  return orc_condensed_element_t(ORC_MDLB, kind, norder, norie, norif, xnod, prm, Aii, Bi, ASchur, BSchur, ni, nb);
}
int orc_condensed_batch(int kind, int nel, const int *norder, const int *norie, const int *norif, const double *xnod,
                        int xnod_stride, const orc_params *prm, void *Aii, void *Bi, void *ASchur, void *BSchur,
                        long sAii, long sBi, long sAS, long sBS, int *info, int nthreads) {
  return orc_condensed_batch_t(kind, nel, NULL, norder, norie, norif, xnod, xnod_stride, prm, Aii, Bi, ASchur, BSchur, sAii,
                               sBi, sAS, sBS, info, nthreads);
}

/* ======================================================================= MAXWELL / ULTRAWEAK DPG, scalar-loop twin
 * Restatement of problems/MAXWELL/ULTRAWEAK_DPG/elem/elem_maxwell.F90:25-660 -- the reference's own SECOND formulation of the
 * element: per-point loops over test/trial functions that accumulate the load (:290-298), the field stiffness stiff_EQ_T
 * (:303-326), the Hermitian Gram matrix entry by entry in 2x2 blocks (:330-410, packed upper storage there), the trace pairings
 * stiff_EE_T (:445-560), then ZPPTRF / ZTPTRS / ZHERK (:588-624).  It shares NO dense kernel call pattern with the BLAS3
 * version above (no sqrt(weight) factorisation, no real/imaginary splitting): the test-suite asserts that both give the same Gram
 * matrix, enriched stiffness and element matrices for arbitrary complex permittivity tensors (SURVEY.md 7 step 0 iii). */
int orc_elem_maxwell_uw_scalar_t(int et, const int norder[19], const int norie[12], const int norif[6], const double *xnod,
                                 const orc_params *prm, zdouble *Aloc, zdouble *Bloc, zdouble *gram_out, zdouble *stiff_out) {
  int nH, nE, nV, nQ, nHH, nEE, nVV, nQQ, bH, bE, bV, bQ, norderP[19], norderi[19];
  int dp = prm->nord_add, nordP = orc_enriched_mid(et, norder[MIDX(et)], dp), maxpp = orc_get_maxp() + 1;
  orc_compute_enriched_order(et, nordP, norderP);
  orc_celndof(et, norder, &nH, &nE, &nV, &nQ);
  orc_celndof(et, norderP, &nHH, &nEE, &nVV, &nQQ);
  orc_ndof_nod_mid(et, norder[MIDX(et)], &bH, &bE, &bV, &bQ);
  const int nEi = nE - bE, nTest = 2 * nEE, nTrial = 2 * nEi + 6 * nQ, ncol = nTrial + 1, jE = 2 * nEi;
  /* norderi: the element's orders with the middle node forced to the lowest order (no interior functions; :176-189) */
  memcpy(norderi, norder, sizeof(int) * 19);
  norderi[MIDX(et)] = (et == ORC_MDLB) ? 111 : 11;
  const double afac = (prm->test_norm == 2) ? 1.0 : prm->alpha_norm;
  zdouble *gram = xmalloc(sizeof(zdouble) * nTest * nTest), *stiff = xmalloc(sizeof(zdouble) * nTest * ncol);
  double *xiloc = xmalloc(sizeof(double) * 3 * 1000), *waloc = xmalloc(sizeof(double) * 1000);
  double *shapH = xmalloc(sizeof(double) * nH), *gradH = xmalloc(sizeof(double) * 3 * nH), *shapQ = xmalloc(sizeof(double) * nQ);
  double *shapEE = xmalloc(sizeof(double) * 3 * nEE), *curlEE = xmalloc(sizeof(double) * 3 * nEE);
  double *shapE = xmalloc(sizeof(double) * 3 * nE), *curlE = xmalloc(sizeof(double) * 3 * nE);
  double *shapF = xmalloc(sizeof(double) * 3 * nEE), *curlF = xmalloc(sizeof(double) * 3 * nEE), *shapFi = xmalloc(sizeof(double) * 3 * nE);
  zdouble *epsTshapF = xmalloc(sizeof(zdouble) * 3 * nEE), *epscurlF = xmalloc(sizeof(zdouble) * 3 * nEE);
  /* ---- element integrals */
  const int nint = orc_set_3D_int(et, norder, norif, dp, maxpp, xiloc, waloc);
  for (int l = 0; l < nint; l++) {
    double x[3], J[9], Ji[9], rjac; int iflag;
    orc_shape3DH(et, xiloc + 3 * l, norder, norie, norif, shapH, gradH);
    orc_shape3DQ(et, xiloc + 3 * l, norder, shapQ);
    orc_shape3EE(et, xiloc + 3 * l, nordP, shapEE, curlEE);
    orc_geom3D(xnod, shapH, gradH, nH, x, J, Ji, &rjac, &iflag);
    check_jac(iflag, rjac);
    const double weight = rjac * waloc[l];
    zdouble zJ[3], za[9];
    maxwell_uw_source(prm, x, l, zJ);
    for (int i = 0; i < 9; i++) za[i] = (I * prm->omega * prm->eps) * prm->eps_tensor[i];   /* za(i,j) at [i + 3j] */
    const zdouble zc1 = I * prm->omega * prm->mu;
    for (int k = 0; k < nEE; k++) {
      pull_grad(shapEE + 3 * k, Ji, shapF + 3 * k);
      push_curl(curlEE + 3 * k, J, rjac, curlF + 3 * k);
      for (int c = 0; c < 3; c++) {   /* epsTshapF = za^H F ; epscurlF = za curl F */
        zdouble a = 0, b = 0;
        for (int d = 0; d < 3; d++) { a += conj(za[d + 3 * c]) * shapF[3 * k + d]; b += za[c + 3 * d] * curlF[3 * k + d]; }
        epsTshapF[3 * k + c] = a; epscurlF[3 * k + c] = b;
      }
    }
    for (int k1 = 0; k1 < nEE; k1++) {
      const double *fldF = shapF + 3 * k1, *crlF = curlF + 3 * k1;
      const zdouble *epsTfldF = epsTshapF + 3 * k1;
      IDX(stiff, nTest, 2 * k1, nTrial) += (fldF[0] * zJ[0] + fldF[1] * zJ[1] + fldF[2] * zJ[2]) * weight;
      for (int k2 = 0; k2 < nQ; k2++) {
        const int m = jE + 6 * k2;
        const double q = shapQ[k2] / rjac;
        for (int c = 0; c < 3; c++) {
          IDX(stiff, nTest, 2 * k1, m + c) -= q * conj(epsTfldF[c]) * weight;      /* -i w eps (E,F) */
          IDX(stiff, nTest, 2 * k1, m + 3 + c) += q * crlF[c] * weight;            /* (H, curl F)    */
          IDX(stiff, nTest, 2 * k1 + 1, m + c) += q * crlF[c] * weight;            /* (E, curl G)    */
          IDX(stiff, nTest, 2 * k1 + 1, m + 3 + c) += zc1 * q * fldF[c] * weight;  /* i w mu (H,G)   */
        }
      }
      for (int k2 = k1; k2 < nEE; k2++) {
        const double *fldE = shapF + 3 * k2, *crlE = curlF + 3 * k2;
        const zdouble *epsTfldE = epsTshapF + 3 * k2, *epscrlE = epscurlF + 3 * k2;
        const double FF = fldF[0] * fldE[0] + fldF[1] * fldE[1] + fldF[2] * fldE[2];
        const double CC = crlF[0] * crlE[0] + crlF[1] * crlE[1] + crlF[2] * crlE[2];
        zdouble zaux = 0, zcux = 0;
        if (prm->test_norm != 2) {
          zaux = conj(epsTfldF[0]) * epsTfldE[0] + conj(epsTfldF[1]) * epsTfldE[1] + conj(epsTfldF[2]) * epsTfldE[2];
          zcux = (creal(zc1) * creal(zc1) + cimag(zc1) * cimag(zc1)) * FF;
        }
        IDX(gram, nTest, 2 * k1, 2 * k2) += (zaux + afac * FF + CC) * weight;            /* G_11 */
        IDX(gram, nTest, 2 * k1 + 1, 2 * k2 + 1) += (zcux + afac * FF + CC) * weight;    /* G_22 */
        if (prm->test_norm != 1) continue;
        zaux = -(fldF[0] * epscrlE[0] + fldF[1] * epscrlE[1] + fldF[2] * epscrlE[2]);
        zcux = conj(zc1) * (crlF[0] * fldE[0] + crlF[1] * fldE[1] + crlF[2] * fldE[2]);
        IDX(gram, nTest, 2 * k1, 2 * k2 + 1) += (zaux + zcux) * weight;                  /* G_12 */
        if (k1 != k2) {
          zaux = -(crlF[0] * epsTfldE[0] + crlF[1] * epsTfldE[1] + crlF[2] * epsTfldE[2]);
          zcux = zc1 * (fldF[0] * crlE[0] + fldF[1] * crlE[1] + fldF[2] * crlE[2]);
          IDX(gram, nTest, 2 * k1 + 1, 2 * k2) += (zaux + zcux) * weight;                /* G_21 */
        }
      }
    }
  }
  /* ---- boundary integrals: every face, all interface functions at once (shape3DE with the middle order forced to 1) */
  for (int ifc = 1; ifc <= orc_nface(et); ifc++) {
    int nordf[5];
    double tloc[200], wtloc[100];
    const int nsign = orc_nsign_param(et, ifc);
    orc_face_order(et, ifc, norder, nordf);
    const int nintf = orc_set_2D_int(orc_face_is_tri(et, ifc), nordf, norif[ifc - 1], dp, maxpp, tloc, wtloc);
    for (int l = 0; l < nintf; l++) {
      double xi[3], dxidt[6], x[3], J[9], Ji[9], rjac, dxdt[6], rn[3], bjac;
      orc_face_param(et, ifc, tloc + 2 * l, xi, dxidt);
      orc_shape3EE(et, xi, nordP, shapEE, curlEE);
      orc_shape3DH(et, xi, norder, norie, norif, shapH, gradH);
      const int nEi_l = orc_shape3DE(et, xi, norderi, norie, norif, shapE, curlE);
      if (nEi_l != nEi) { fprintf(stderr, "oracle scalar twin: inconsistent NrdofEi %d vs %d\n", nEi_l, nEi); exit(1); }
      orc_bgeom3D(xnod, shapH, gradH, nH, dxidt, nsign, x, J, Ji, &rjac, dxdt, rn, &bjac);
      const double weight = bjac * wtloc[l];
      for (int k = 0; k < nEE; k++) pull_grad(shapEE + 3 * k, Ji, shapF + 3 * k);
      for (int k = 0; k < nEi; k++) pull_grad(shapE + 3 * k, Ji, shapFi + 3 * k);
      for (int k1 = 0; k1 < nEE; k1++) {
        const double *E1 = shapF + 3 * k1;
        for (int k2 = 0; k2 < nEi; k2++) {
          const double *E2 = shapFi + 3 * k2;
          const double rxE[3] = {rn[1] * E2[2] - rn[2] * E2[1], rn[2] * E2[0] - rn[0] * E2[2], rn[0] * E2[1] - rn[1] * E2[0]};
          const double v = (E1[0] * rxE[0] + E1[1] * rxE[1] + E1[2] * rxE[2]) * weight;
          IDX(stiff, nTest, 2 * k1, 2 * k2 + 1) += v;    /* <n x H^, F> */
          IDX(stiff, nTest, 2 * k1 + 1, 2 * k2) += v;    /* <n x E^, G> */
        }
      }
    }
  }
  if (gram_out) memcpy(gram_out, gram, sizeof(zdouble) * nTest * nTest);
  if (stiff_out) memcpy(stiff_out, stiff, sizeof(zdouble) * nTest * ncol);
  /* ---- G = U^H U, B~ = U^-H [B|l], [A|b] = B~^H B~ (ZPPTRF / ZTPTRS / ZHERK on packed storage in the reference) */
  int info = orc_zpotrf_u(nTest, gram, nTest);
  if (!info) {
    orc_ztrsm_u('C', nTest, ncol, gram, nTest, stiff, nTest);
    zdouble *zal = xmalloc(sizeof(zdouble) * ncol * ncol);
    orc_zherk_u('C', ncol, nTest, 1.0, stiff, nTest, 0.0, zal, ncol);
    for (int j = 0; j < nTrial; j++) {
      for (int i = 0; i < nTrial; i++) IDX(Aloc, nTrial, i, j) = (i <= j) ? IDX(zal, ncol, i, j) : conj(IDX(zal, ncol, j, i));
      Bloc[j] = IDX(zal, ncol, j, nTrial);
    }
    free(zal);
  }
  free(gram); free(stiff); free(xiloc); free(waloc); free(shapH); free(gradH); free(shapQ); free(shapEE); free(curlEE); free(shapE);
  free(curlE); free(shapF); free(curlF); free(shapFi); free(epsTshapF); free(epscurlF);
  return info;
}
