/*
 * oracle/shape.c -- CPU restatement of hp3D's hexahedron shape functions (TEST INFRASTRUCTURE ONLY).
 *
 * Follows (paths relative to /root/reference/trunk/src):
 *   utility/decod.F90:21, encod.F90:21, decode.F90:19, ij_to_packed.F90:15
 *   element/shape_1/Polynomials.F90:34 (PolyLegendre), :109 (PolyILegendre), :455 (HomLegendre), :497 (HomILegendre)
 *   element/shape_1/Ancillary.F90:37 (AncPhiE), :88 (AncEE), :171 (AncPhiQuad), :237 (AncEQuad), :307 (AncVQuad)
 *   element/shape_1/AffineCoordinates.F90:56 (AffineHexahedron)
 *   element/shape_1/BlendProject.F90:197 (BlendHexaV), :255 (BlendProjectHexaE), :365 (BlendProjectHexaF)
 *   element/shape_1/Orient.F90:9 (OrientE), :39 (OrientQuad)
 *   element/shape_1/Hexahedron.F90:33,240,468,634 (shape3D{H,E,V,Q}Hexa)
 *   element/shape_1/Segment.F90:30,138 ; broken/BrokenHexahedron.F90:31,139,286,422
 * The reference enumerates vertices/edges/faces by hand; here the same topology is table-driven.
 */
#include "hp3d_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ integer encodings */
void orc_decod(int nick, int mod, int n, int *narray) {
  /* decod.F90:31-37: the first slot keeps whatever is left (555 with n=2 -> 55,5) */
  int nick1 = nick;
  for (int i = 1; i <= n - 1; i++) {
    int nick2 = nick1 / mod;
    narray[n - i] = nick1 - nick2 * mod;
    nick1 = nick2;
  }
  narray[0] = nick1;
}
void orc_encod(const int *narray, int mod, int n, int *nick) {
  int v = narray[0];
  for (int i = 1; i < n; i++) v = v * mod + narray[i];
  *nick = v;
}
void orc_decode(int nick, int *j1, int *j2) { *j1 = nick / 10; *j2 = nick - (*j1) * 10; }
void orc_decode2(int nick, int *j1, int *j2) { *j1 = nick / 100; *j2 = nick - (*j1) * 100; }
void orc_ddecode(int nick, int *j1, int *j2, int *j3) {
  int nick2 = nick / 10;
  *j3 = nick - nick2 * 10;
  *j1 = nick2 / 10;
  *j2 = nick2 - (*j1) * 10;
}
int orc_ij_upper_to_packed(int i, int j) { return i + (j - 1) * j / 2; }
int orc_ij_lower_to_packed(int i, int j, int n) { return i + (j - 1) * (2 * n - j) / 2; }

static int g_maxp = 6;
void orc_set_maxp(int maxp) { g_maxp = maxp; }
int orc_get_maxp(void) { return g_maxp; }

/* ------------------------------------------------------------------ 1-D polynomials */
/* Polynomials.F90:34-62 : shifted scaled Legendre, i*P_i = (2i-1)(2x-t)P_{i-1} - (i-1)t^2 P_{i-2} */
void orc_poly_legendre(double x, double t, int nord, double *P) {
  P[0] = 1.0;
  double y = 0.0;
  if (nord >= 1) { y = 2.0 * x - t; P[1] = y; }
  if (nord >= 2) {
    double tt = t * t;
    for (int i = 2; i <= nord; i++) {
      P[i] = (2 * i - 1) * y * P[i - 1] - (i - 1) * tt * P[i - 2];
      P[i] = P[i] / i;
    }
  }
}
/* Polynomials.F90:109-147 : L_i=(P_i - t^2 P_{i-2})/(4i-2), dL_i/dx=P_{i-1}, dL_i/dt=-(P_{i-1}+tP_{i-2})/2 */
void orc_poly_ilegendre(double x, double t, int nord, int idec, double *L, double *P, double *R) {
  double ptemp[ORC_MAXN + 2];
  orc_poly_legendre(x, t, nord, ptemp);
  for (int i = 1; i <= nord - 1; i++) P[i] = ptemp[i];
  double tt = t * t;
  for (int i = 2; i <= nord; i++) {
    int ifact = 4 * i - 2;
    L[i] = (ptemp[i] - tt * ptemp[i - 2]) / ifact;
    if (!idec) R[i - 1] = -(ptemp[i - 1] + t * ptemp[i - 2]) / 2;
  }
}

/* an "affine pair" (s0,s1) with its gradients in R^3 */
typedef struct { double s[2]; double ds[2][3]; } apair;

static void cross3(const double a[3], const double b[3], double c[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

/* Polynomials.F90:455 HomLegendre */
static void hom_legendre(const apair *S, int nord, double *homP) {
  orc_poly_legendre(S->s[1], S->s[0] + S->s[1], nord, homP);
}
/* Polynomials.F90:497 HomILegendre == Ancillary.F90:37 AncPhiE ; L[2..nord], dL[2..nord][3] */
static void anc_phiE(const apair *S, int nord, int idec, double *L, double (*dL)[3]) {
  double hP[ORC_MAXN + 2], hR[ORC_MAXN + 2];
  if (nord < 2) return;
  if (idec) {
    orc_poly_ilegendre(S->s[1], 1.0, nord, 1, L, hP, hR);
    for (int i = 2; i <= nord; i++)
      for (int n = 0; n < 3; n++) dL[i][n] = hP[i - 1] * S->ds[1][n];
  } else {
    orc_poly_ilegendre(S->s[1], S->s[0] + S->s[1], nord, 0, L, hP, hR);
    double ds01[3];
    for (int n = 0; n < 3; n++) ds01[n] = S->ds[0][n] + S->ds[1][n];
    for (int i = 2; i <= nord; i++)
      for (int n = 0; n < 3; n++) dL[i][n] = hP[i - 1] * S->ds[1][n] + hR[i - 1] * ds01[n];
  }
}
/* Ancillary.F90:88 AncEE ; EE[0..nord-1][3], curlEE[0..nord-1][3] */
static void anc_EE(const apair *S, int nord, int idec, double (*EE)[3], double (*cEE)[3]) {
  double homP[ORC_MAXN + 2];
  if (nord < 1) return;
  hom_legendre(S, nord - 1, homP);
  if (idec) {
    for (int i = 0; i <= nord - 1; i++)
      for (int n = 0; n < 3; n++) { EE[i][n] = homP[i] * S->ds[1][n]; cEE[i][n] = 0.0; }
  } else {
    double whi[3], cwhi[3];
    for (int n = 0; n < 3; n++) whi[n] = S->s[0] * S->ds[1][n] - S->s[1] * S->ds[0][n];
    cross3(S->ds[0], S->ds[1], cwhi);
    for (int i = 0; i <= nord - 1; i++)
      for (int n = 0; n < 3; n++) { EE[i][n] = homP[i] * whi[n]; cEE[i][n] = (i + 2) * homP[i] * cwhi[n]; }
  }
}
#define NQ (ORC_MAXN + 1)
/* Ancillary.F90:171 AncPhiQuad ; phi[i][j], i in 2..n0, j in 2..n1 */
static void anc_phi_quad(const apair ST[2], const int nord[2], const int idec[2], double (*phi)[NQ],
                         double (*dphi)[NQ][3]) {
  double pS[NQ], pT[NQ], dS[NQ][3], dT[NQ][3];
  anc_phiE(&ST[0], nord[0], idec[0], pS, dS);
  anc_phiE(&ST[1], nord[1], idec[1], pT, dT);
  for (int j = 2; j <= nord[1]; j++)
    for (int i = 2; i <= nord[0]; i++) {
      phi[i][j] = pS[i] * pT[j];
      for (int n = 0; n < 3; n++) dphi[i][j][n] = pS[i] * dT[j][n] + pT[j] * dS[i][n];
    }
}
/* Ancillary.F90:237 AncEQuad ; E[i][j], i in 0..n0-1, j in 2..n1 */
static void anc_E_quad(const apair ST[2], const int nord[2], const int idec[2], double (*E)[NQ][3],
                       double (*cE)[NQ][3]) {
  double EES[NQ][3], cEES[NQ][3], pT[NQ], dT[NQ][3], x[3];
  anc_EE(&ST[0], nord[0], idec[0], EES, cEES);
  anc_phiE(&ST[1], nord[1], idec[1], pT, dT);
  for (int j = 2; j <= nord[1]; j++)
    for (int i = 0; i <= nord[0] - 1; i++) {
      cross3(dT[j], EES[i], x);
      for (int n = 0; n < 3; n++) { E[i][j][n] = EES[i][n] * pT[j]; cE[i][j][n] = cEES[i][n] * pT[j] + x[n]; }
    }
}
/* Ancillary.F90:307 AncVQuad ; V[i][j], i in 0..n0-1, j in 0..n1-1 */
static void anc_V_quad(const apair ST[2], const int nord[2], const int idec[2], double (*V)[NQ][3],
                       double (*dV)[NQ]) {
  double EES[NQ][3], cEES[NQ][3], EET[NQ][3], cEET[NQ][3];
  anc_EE(&ST[0], nord[0], idec[0], EES, cEES);
  anc_EE(&ST[1], nord[1], idec[1], EET, cEET);
  for (int j = 0; j <= nord[1] - 1; j++)
    for (int i = 0; i <= nord[0] - 1; i++) {
      cross3(EES[i], EET[j], V[i][j]);
      if (idec[0] && idec[1]) dV[i][j] = 0.0;
      else {
        double p1 = 0, p2 = 0;
        for (int n = 0; n < 3; n++) { p1 += EET[j][n] * cEES[i][n]; p2 += EES[i][n] * cEET[j][n]; }
        dV[i][j] = p1 - p2;
      }
    }
}

/* ------------------------------------------------------------------ master hexahedron topology
 * element_data.F90:31-35 (BRICK_COORD), :67-70 (edges), :90-93 (faces); blending/projection pairs as
 * enumerated in BlendProject.F90:197-560.  Axes 0,1,2 = x,y,z ; side 0 -> mu0=1-xi, side 1 -> mu1=xi. */
static const int VERT_SIDE[8][3] = {{0,0,0},{1,0,0},{1,1,0},{0,1,0},{0,0,1},{1,0,1},{1,1,1},{0,1,1}};
/* edge: {axis, blendAxisA, sideA, blendAxisB, sideB} */
static const int EDGE_DEF[12][5] = {
  {0, 1,0, 2,0}, {1, 0,1, 2,0}, {0, 1,1, 2,0}, {1, 0,0, 2,0},
  {0, 1,0, 2,1}, {1, 0,1, 2,1}, {0, 1,1, 2,1}, {1, 0,0, 2,1},
  {2, 0,0, 1,0}, {2, 0,1, 1,0}, {2, 0,1, 1,1}, {2, 0,0, 1,1}};
/* face: {blendAxis, side, sAxis, tAxis} */
static const int FACE_DEF[6][4] = {{2,0, 0,1}, {2,1, 0,1}, {1,0, 0,2}, {0,1, 1,2}, {1,1, 0,2}, {0,0, 1,2}};
/* Orient.F90:58-99 : per orientation, are the pairs swapped / is S flipped / is T flipped */
static const int OQ_SWAP[8]  = {0,1,0,1,1,0,1,0};
static const int OQ_FLIPS[8] = {0,0,1,1,0,1,1,0};
static const int OQ_FLIPT[8] = {0,1,1,0,0,0,1,1};

/* AffineCoordinates.F90:56 */
static void affine_hexa(const double xi[3], apair Mu[3]) {
  for (int d = 0; d < 3; d++) {
    Mu[d].s[0] = 1.0 - xi[d]; Mu[d].s[1] = xi[d];
    for (int n = 0; n < 3; n++) { Mu[d].ds[0][n] = 0.0; Mu[d].ds[1][n] = 0.0; }
    Mu[d].ds[0][d] = -1.0; Mu[d].ds[1][d] = 1.0;
  }
}
static void flip_pair(const apair *in, int flip, apair *out) {
  int a = flip ? 1 : 0, b = flip ? 0 : 1;
  out->s[0] = in->s[a]; out->s[1] = in->s[b];
  for (int n = 0; n < 3; n++) { out->ds[0][n] = in->ds[a][n]; out->ds[1][n] = in->ds[b][n]; }
}
/* Orient.F90:9 */
static void orient_edge(const apair *in, int nori, apair *out) {
  if (nori != 0 && nori != 1) { fprintf(stderr, "orient_edge: invalid orientation %d\n", nori); exit(1); }
  flip_pair(in, nori, out);
}
/* Orient.F90:39 */
static void orient_quad(const apair ST[2], int nori, apair G[2]) {
  if (nori < 0 || nori > 7) { fprintf(stderr, "orient_quad: invalid orientation %d\n", nori); exit(1); }
  int p0 = OQ_SWAP[nori] ? 1 : 0, p1 = OQ_SWAP[nori] ? 0 : 1;
  flip_pair(&ST[p0], OQ_FLIPS[nori], &G[0]);
  flip_pair(&ST[p1], OQ_FLIPT[nori], &G[1]);
}

/* ------------------------------------------------------------------ H1 : Hexahedron.F90:33-156 */
int orc_shape3DH_hexa(const double xi[3], const int nord[19], const int norie[12], const int norif[6],
                      double *shapH, double *gradH) {
  apair Mu[3];
  affine_hexa(xi, Mu);
  int m = 0;
  const int one2[2] = {1, 1};
  /* vertices (BlendHexaV) */
  for (int v = 0; v < 8; v++) {
    double b[3], db[3][3];
    for (int d = 0; d < 3; d++) {
      b[d] = Mu[d].s[VERT_SIDE[v][d]];
      for (int n = 0; n < 3; n++) db[d][n] = Mu[d].ds[VERT_SIDE[v][d]][n];
    }
    shapH[m] = b[0] * b[1] * b[2];
    for (int n = 0; n < 3; n++)
      gradH[3 * m + n] = b[0] * b[1] * db[2][n] + b[0] * db[1][n] * b[2] + db[0][n] * b[1] * b[2];
    m++;
  }
  /* edges */
  for (int e = 0; e < 12; e++) {
    if (nord[e] - 1 <= 0) continue;
    const int *E = EDGE_DEF[e];
    double b1 = Mu[E[1]].s[E[2]], b2 = Mu[E[3]].s[E[4]];
    const double *db1 = Mu[E[1]].ds[E[2]], *db2 = Mu[E[3]].ds[E[4]];
    apair G;
    orient_edge(&Mu[E[0]], norie[e], &G);
    double phi[NQ], dphi[NQ][3];
    anc_phiE(&G, nord[e], 1, phi, dphi);
    for (int i = 2; i <= nord[e]; i++) {
      shapH[m] = b1 * b2 * phi[i];
      for (int n = 0; n < 3; n++)
        gradH[3 * m + n] = b1 * b2 * dphi[i][n] + b1 * db2[n] * phi[i] + db1[n] * b2 * phi[i];
      m++;
    }
  }
  /* faces */
  double phiQ[NQ][NQ], dphiQ[NQ][NQ][3];
  for (int f = 0; f < 6; f++) {
    int nf[2];
    orc_decod(nord[12 + f], 10, 2, nf);
    if ((nf[0] - 1) * (nf[1] - 1) <= 0) continue;
    const int *F = FACE_DEF[f];
    double b = Mu[F[0]].s[F[1]];
    const double *db = Mu[F[0]].ds[F[1]];
    apair ST[2] = {Mu[F[2]], Mu[F[3]]}, G[2];
    orient_quad(ST, norif[f], G);
    anc_phi_quad(G, nf, one2, phiQ, dphiQ);
    for (int j = 2; j <= nf[1]; j++)
      for (int i = 2; i <= nf[0]; i++) {
        shapH[m] = b * phiQ[i][j];
        for (int n = 0; n < 3; n++) gradH[3 * m + n] = b * dphiQ[i][j][n] + db[n] * phiQ[i][j];
        m++;
      }
  }
  /* interior */
  int nb[3];
  orc_decod(nord[18], 10, 3, nb);
  if ((nb[0] - 1) * (nb[1] - 1) * (nb[2] - 1) > 0) {
    apair ST[2] = {Mu[0], Mu[1]};
    double phi[NQ], dphi[NQ][3];
    anc_phi_quad(ST, nb, one2, phiQ, dphiQ);
    anc_phiE(&Mu[2], nb[2], 1, phi, dphi);
    for (int k = 2; k <= nb[2]; k++)
      for (int j = 2; j <= nb[1]; j++)
        for (int i = 2; i <= nb[0]; i++) {
          shapH[m] = phiQ[i][j] * phi[k];
          for (int n = 0; n < 3; n++) gradH[3 * m + n] = phiQ[i][j] * dphi[k][n] + dphiQ[i][j][n] * phi[k];
          m++;
        }
  }
  return m;
}

/* ------------------------------------------------------------------ H(curl) : Hexahedron.F90:240-388 */
int orc_shape3DE_hexa(const double xi[3], const int nord[19], const int norie[12], const int norif[6],
                      double *shapE, double *curlE) {
  apair Mu[3];
  affine_hexa(xi, Mu);
  int m = 0;
  const int one2[2] = {1, 1};
  double EQ[NQ][NQ][3], cEQ[NQ][NQ][3];
  /* edges */
  for (int e = 0; e < 12; e++) {
    if (nord[e] <= 0) continue;
    const int *E = EDGE_DEF[e];
    double b1 = Mu[E[1]].s[E[2]], b2 = Mu[E[3]].s[E[4]];
    const double *db1 = Mu[E[1]].ds[E[2]], *db2 = Mu[E[3]].ds[E[4]];
    apair G;
    orient_edge(&Mu[E[0]], norie[e], &G);
    double EE[NQ][3], cEE[NQ][3];
    anc_EE(&G, nord[e], 1, EE, cEE);
    for (int i = 0; i <= nord[e] - 1; i++) {
      double dt[3], ct[3];
      for (int n = 0; n < 3; n++) dt[n] = b1 * db2[n] + db1[n] * b2;
      cross3(dt, EE[i], ct);
      for (int n = 0; n < 3; n++) { shapE[3 * m + n] = b1 * b2 * EE[i][n]; curlE[3 * m + n] = ct[n]; }
      m++;
    }
  }
  /* faces: two families; outer loop always along the 2nd (oriented) face axis */
  for (int f = 0; f < 6; f++) {
    int nf[2];
    orc_decod(nord[12 + f], 10, 2, nf);
    const int *F = FACE_DEF[f];
    double b = Mu[F[0]].s[F[1]];
    const double *db = Mu[F[0]].ds[F[1]];
    apair ST[2] = {Mu[F[2]], Mu[F[3]]}, G[2];
    orient_quad(ST, norif[f], G);
    for (int fam = 0; fam < 2; fam++) {
      int a = fam, bb = 1 - fam; /* E-type axis a, phi-type axis bb (0-based face axes) */
      if (nf[a] * (nf[bb] - 1) <= 0) continue;
      apair Gab[2] = {G[a], G[bb]};
      int nab[2] = {nf[a], nf[bb]};
      anc_E_quad(Gab, nab, one2, EQ, cEQ);
      int lo[2], hi[2];
      lo[a] = 0; hi[a] = nf[a] - 1; lo[bb] = 2; hi[bb] = nf[bb];
      for (int jg = lo[1]; jg <= hi[1]; jg++)
        for (int ig = lo[0]; ig <= hi[0]; ig++) {
          int g[2] = {ig, jg};
          int i = g[a], j = g[bb];
          double ct[3];
          cross3(db, EQ[i][j], ct);
          for (int n = 0; n < 3; n++) {
            shapE[3 * m + n] = b * EQ[i][j][n];
            curlE[3 * m + n] = b * cEQ[i][j][n] + ct[n];
          }
          m++;
        }
    }
  }
  /* interior: three families (a,b,c) = cyclic shifts of (x,y,z) */
  int nb[3];
  orc_decod(nord[18], 10, 3, nb);
  for (int fam = 0; fam < 3; fam++) {
    int a = fam, b = (fam + 1) % 3, c = (fam + 2) % 3;
    if (nb[a] * (nb[b] - 1) * (nb[c] - 1) <= 0) continue;
    apair ST[2] = {Mu[a], Mu[b]};
    int nab[2] = {nb[a], nb[b]};
    double phi[NQ], dphi[NQ][3];
    anc_E_quad(ST, nab, one2, EQ, cEQ);
    anc_phiE(&Mu[c], nb[c], 1, phi, dphi);
    int lo[3], hi[3];
    lo[a] = 0; hi[a] = nb[a] - 1; lo[b] = 2; hi[b] = nb[b]; lo[c] = 2; hi[c] = nb[c];
    for (int kg = lo[2]; kg <= hi[2]; kg++)
      for (int jg = lo[1]; jg <= hi[1]; jg++)
        for (int ig = lo[0]; ig <= hi[0]; ig++) {
          int g[3] = {ig, jg, kg};
          int i = g[a], j = g[b], k = g[c];
          double ct[3];
          cross3(dphi[k], EQ[i][j], ct);
          for (int n = 0; n < 3; n++) {
            shapE[3 * m + n] = EQ[i][j][n] * phi[k];
            curlE[3 * m + n] = phi[k] * cEQ[i][j][n] + ct[n];
          }
          m++;
        }
  }
  return m;
}

/* ------------------------------------------------------------------ H(div) : Hexahedron.F90:468-569 */
int orc_shape3DV_hexa(const double xi[3], const int nord[19], const int norif[6], double *shapV, double *divV) {
  apair Mu[3];
  affine_hexa(xi, Mu);
  int m = 0;
  const int one2[2] = {1, 1};
  double VQ[NQ][NQ][3], dVQ[NQ][NQ];
  for (int f = 0; f < 6; f++) {
    int nf[2];
    orc_decod(nord[12 + f], 10, 2, nf);
    const int *F = FACE_DEF[f];
    double b = Mu[F[0]].s[F[1]];
    const double *db = Mu[F[0]].ds[F[1]];
    apair ST[2] = {Mu[F[2]], Mu[F[3]]}, G[2];
    orient_quad(ST, norif[f], G);
    if (nf[0] * nf[1] <= 0) continue;
    anc_V_quad(G, nf, one2, VQ, dVQ);
    for (int j = 0; j <= nf[1] - 1; j++)
      for (int i = 0; i <= nf[0] - 1; i++) {
        double d = 0;
        for (int n = 0; n < 3; n++) { shapV[3 * m + n] = b * VQ[i][j][n]; d += db[n] * VQ[i][j][n]; }
        divV[m] = d;
        m++;
      }
  }
  int nb[3];
  orc_decod(nord[18], 10, 3, nb);
  for (int fam = 0; fam < 3; fam++) {
    int a = fam, b = (fam + 1) % 3, c = (fam + 2) % 3;
    if (nb[a] * nb[b] * (nb[c] - 1) <= 0) continue;
    apair ST[2] = {Mu[a], Mu[b]};
    int nab[2] = {nb[a], nb[b]};
    double phi[NQ], dphi[NQ][3];
    anc_V_quad(ST, nab, one2, VQ, dVQ);
    anc_phiE(&Mu[c], nb[c], 1, phi, dphi);
    int lo[3], hi[3];
    lo[a] = 0; hi[a] = nb[a] - 1; lo[b] = 0; hi[b] = nb[b] - 1; lo[c] = 2; hi[c] = nb[c];
    for (int kg = lo[2]; kg <= hi[2]; kg++)
      for (int jg = lo[1]; jg <= hi[1]; jg++)
        for (int ig = lo[0]; ig <= hi[0]; ig++) {
          int g[3] = {ig, jg, kg};
          int i = g[a], j = g[b], k = g[c];
          double d = 0;
          for (int n = 0; n < 3; n++) { shapV[3 * m + n] = phi[k] * VQ[i][j][n]; d += dphi[k][n] * VQ[i][j][n]; }
          divV[m] = d;
          m++;
        }
  }
  return m;
}

/* ------------------------------------------------------------------ L2 : Hexahedron.F90:634-678 */
int orc_shape3DQ_hexa(const double xi[3], const int nord[19], double *shapQ) {
  apair Mu[3];
  affine_hexa(xi, Mu);
  int nb[3], m = 0;
  orc_decod(nord[18], 10, 3, nb);
  if (nb[0] * nb[1] * nb[2] <= 0) return 0;
  double hP[3][NQ];
  for (int d = 0; d < 3; d++) hom_legendre(&Mu[d], nb[d] - 1, hP[d]);
  for (int k = 0; k <= nb[2] - 1; k++)
    for (int j = 0; j <= nb[1] - 1; j++)
      for (int i = 0; i <= nb[0] - 1; i++) shapQ[m++] = hP[0][i] * hP[1][j] * hP[2][k];
  return m;
}

/* ------------------------------------------------------------------ broken (enriched) functions
 * Segment.F90:30-100 (shape1DHSeg: [1-x, x, L_2..L_p], derivs [-1, 1, P_1..P_{p-1}]),
 * Segment.F90:138-190 (shape1DQSeg: [P_0..P_{p-1}]); BrokenSegment.F90 forwards to these. */
int orc_shape1HH(double xi, int nord, double *shapH, double *gradH) {
  double mu0 = 1.0 - xi, mu1 = xi;
  shapH[0] = mu0; gradH[0] = -1.0;
  shapH[1] = mu1; gradH[1] = 1.0;
  int m = 2;
  if (nord - 1 > 0) {
    double L[NQ], P[NQ], R[NQ];
    orc_poly_ilegendre(mu1, 1.0, nord, 1, L, P, R);
    for (int i = 2; i <= nord; i++) { shapH[m] = L[i]; gradH[m] = P[i - 1] * 1.0; m++; }
  }
  return m;
}
int orc_shape1QQ(double xi, int nord, double *shapQ) {
  double mu0 = 1.0 - xi, mu1 = xi;
  if (nord <= 0) return 0;
  double hP[NQ];
  orc_poly_legendre(mu1, mu0 + mu1, nord - 1, hP);
  for (int i = 0; i <= nord - 1; i++) shapQ[i] = hP[i];
  return nord;
}
/* BrokenHexahedron.F90:31-110 */
int orc_shape3HH_hexa(const double xi[3], int nordM, double *shapH, double *gradH) {
  int nb[3], nd[3], m = 0;
  double s[3][NQ + 1], d[3][NQ + 1];
  orc_decod(nordM, 10, 3, nb);
  for (int a = 0; a < 3; a++) nd[a] = orc_shape1HH(xi[a], nb[a], s[a], d[a]);
  for (int k = 0; k < nd[2]; k++)
    for (int j = 0; j < nd[1]; j++)
      for (int i = 0; i < nd[0]; i++) {
        shapH[m] = s[0][i] * s[1][j] * s[2][k];
        gradH[3 * m + 0] = d[0][i] * s[1][j] * s[2][k];
        gradH[3 * m + 1] = s[0][i] * d[1][j] * s[2][k];
        gradH[3 * m + 2] = s[0][i] * s[1][j] * d[2][k];
        m++;
      }
  return m;
}
/* BrokenHexahedron.F90:139-260 */
int orc_shape3EE_hexa(const double xi[3], int nordM, double *shapE, double *curlE) {
  int nb[3], nh[3], nq[3], m = 0;
  double s[3][NQ + 1], d[3][NQ + 1], q[3][NQ + 1];
  orc_decod(nordM, 10, 3, nb);
  for (int a = 0; a < 3; a++) { nh[a] = orc_shape1HH(xi[a], nb[a], s[a], d[a]); nq[a] = orc_shape1QQ(xi[a], nb[a], q[a]); }
  /* x family */
  for (int k = 0; k < nh[2]; k++)
    for (int j = 0; j < nh[1]; j++)
      for (int i = 0; i < nq[0]; i++) {
        double *E = shapE + 3 * m, *C = curlE + 3 * m;
        E[0] = q[0][i] * s[1][j] * s[2][k]; E[1] = 0.0; E[2] = 0.0;
        C[0] = 0.0; C[1] = q[0][i] * s[1][j] * d[2][k]; C[2] = -q[0][i] * d[1][j] * s[2][k];
        m++;
      }
  /* y family */
  for (int k = 0; k < nh[2]; k++)
    for (int j = 0; j < nq[1]; j++)
      for (int i = 0; i < nh[0]; i++) {
        double *E = shapE + 3 * m, *C = curlE + 3 * m;
        E[0] = 0.0; E[1] = s[0][i] * q[1][j] * s[2][k]; E[2] = 0.0;
        C[0] = -s[0][i] * q[1][j] * d[2][k]; C[1] = 0.0; C[2] = d[0][i] * q[1][j] * s[2][k];
        m++;
      }
  /* z family */
  for (int k = 0; k < nq[2]; k++)
    for (int j = 0; j < nh[1]; j++)
      for (int i = 0; i < nh[0]; i++) {
        double *E = shapE + 3 * m, *C = curlE + 3 * m;
        E[0] = 0.0; E[1] = 0.0; E[2] = s[0][i] * s[1][j] * q[2][k];
        C[0] = s[0][i] * d[1][j] * q[2][k]; C[1] = -d[0][i] * s[1][j] * q[2][k]; C[2] = 0.0;
        m++;
      }
  return m;
}
/* BrokenHexahedron.F90:286-400 */
int orc_shape3VV_hexa(const double xi[3], int nordM, double *shapV, double *divV) {
  int nb[3], nh[3], nq[3], m = 0;
  double s[3][NQ + 1], d[3][NQ + 1], q[3][NQ + 1];
  orc_decod(nordM, 10, 3, nb);
  for (int a = 0; a < 3; a++) { nh[a] = orc_shape1HH(xi[a], nb[a], s[a], d[a]); nq[a] = orc_shape1QQ(xi[a], nb[a], q[a]); }
  for (int k = 0; k < nq[2]; k++)
    for (int j = 0; j < nq[1]; j++)
      for (int i = 0; i < nh[0]; i++) {
        double *V = shapV + 3 * m;
        V[0] = s[0][i] * q[1][j] * q[2][k]; V[1] = 0.0; V[2] = 0.0;
        divV[m] = d[0][i] * q[1][j] * q[2][k];
        m++;
      }
  for (int k = 0; k < nq[2]; k++)
    for (int j = 0; j < nh[1]; j++)
      for (int i = 0; i < nq[0]; i++) {
        double *V = shapV + 3 * m;
        V[0] = 0.0; V[1] = q[0][i] * s[1][j] * q[2][k]; V[2] = 0.0;
        divV[m] = q[0][i] * d[1][j] * q[2][k];
        m++;
      }
  for (int k = 0; k < nh[2]; k++)
    for (int j = 0; j < nq[1]; j++)
      for (int i = 0; i < nq[0]; i++) {
        double *V = shapV + 3 * m;
        V[0] = 0.0; V[1] = 0.0; V[2] = q[0][i] * q[1][j] * s[2][k];
        divV[m] = q[0][i] * q[1][j] * d[2][k];
        m++;
      }
  return m;
}
/* BrokenHexahedron.F90:422-440 */
int orc_shape3QQ_hexa(const double xi[3], int nordM, double *shapQ) {
  int norder[19];
  for (int i = 0; i < 12; i++) norder[i] = 1;
  for (int i = 12; i < 18; i++) norder[i] = 11;
  norder[18] = nordM;
  return orc_shape3DQ_hexa(xi, norder, shapQ);
}

/* triangle + prism shape functions share the static helpers above */
#include "shape_prism.c"
