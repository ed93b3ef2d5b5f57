/*
 * oracle/shape_prism.c -- CPU restatement of hp3D's triangle and triangular-prism shape functions
 * (TEST INFRASTRUCTURE ONLY).  Textually included at the end of shape.c (it shares that file's static
 * helpers: apair, anc_phiE, anc_EE, anc_phi_quad, anc_E_quad, anc_V_quad, orient_edge, orient_quad).
 *
 * Follows (paths relative to /root/reference/trunk/src/element/shape_1):
 *   Polynomials.F90:197 (PolyJacobi), :303 (PolyIJacobi), :530 (HomJacobi), :560 (HomIJacobi)
 *   Ancillary.F90:397 (AncPhiTri), :473 (AncETri), :553 (AncVTri)
 *   AffineCoordinates.F90:42 (AffineTriangle), :91 (AffinePrism)
 *   BlendProject.F90:136 (BlendTriV), :162 (ProjectTriE), :562 (BlendPrisV), :598 (BlendProjectPrisME),
 *                    :650 (BlendProjectPrisQE), :688 (BlendProjectPrisTF), :722 (ProjectPrisQF)
 *   Orient.F90:119 (OrientTri)
 *   Triangle.F90:30,140,270,330 (shape2D{H,E,V,Q}Tri) -- evaluated here through the N=3 routines with zero
 *       third gradient components: the first two components (and the third curl component) are bit-identical
 *   Prism.F90:38,358,760,1040 (shape3D{H,E,V,Q}Pris)
 *   broken/BrokenTriangle.F90, broken/BrokenPrism.F90 (enriched test spaces: orientation 0, uniform order)
 */

typedef struct { double s[3]; double ds[3][3]; } atri; /* three affine coordinates with gradients */

/* Polynomials.F90:197-262 PolyJacobi : P[a][i] = P_i^{alpha(a)}(x;t), alpha(a) = minalpha + 2a, a,i in 0..nord */
static void poly_jacobi(double x, double t, int nord, int minalpha, double P[NQ + 1][NQ + 1]) {
  int alpha[NQ + 1];
  for (int a = 0; a <= nord; a++) { alpha[a] = minalpha + 2 * a; P[a][0] = 1.0; }
  double y = 0.0;
  if (nord >= 1) {
    y = 2.0 * x - t;
    for (int a = 0; a <= nord - 1; a++) P[a][1] = y + alpha[a] * x;
  }
  if (nord >= 2) {
    double tt = t * t;
    int ni = -1;
    for (int a = 0; a <= nord - 2; a++) {
      int al = alpha[a], aa = al * al;
      ni++;
      for (int i = 2; i <= nord - ni; i++) {
        int ai = 2 * i * (i + al) * (2 * i + al - 2);
        int bi = 2 * i + al - 1;
        int ci = (2 * i + al) * (2 * i + al - 2);
        int di = 2 * (i + al - 1) * (i - 1) * (2 * i + al);
        P[a][i] = bi * (ci * y + aa * t) * P[a][i - 1] - di * tt * P[a][i - 2];
        P[a][i] = P[a][i] / ai;
      }
    }
  }
}
/* Polynomials.F90:303-400 PolyIJacobi : L[a][i] (a,i in 1..nord), P[a][0..nord-1], R[a][0..nord-1] ; alpha(a)=minalpha+2(a-1) */
static void poly_ijacobi(double x, double t, int nord, int minalpha, int idec, double L[NQ + 1][NQ + 1],
                         double P[NQ + 1][NQ + 1], double R[NQ + 1][NQ + 1]) {
  double pt[NQ + 1][NQ + 1]; /* pt[a-1][i] = ptemp(a,i) */
  poly_jacobi(x, t, nord, minalpha, pt);
  for (int a = 1; a <= nord; a++) {
    for (int i = 0; i <= nord - 1; i++) P[a][i] = pt[a - 1][i];
    L[a][1] = x;
    if (!idec) R[a][0] = 0.0;
  }
  if (nord >= 2) {
    double tt = t * t;
    int ni = -1;
    for (int a = 1; a <= nord - 1; a++) {
      int al = minalpha + 2 * (a - 1);
      ni++;
      for (int i = 2; i <= nord - ni; i++) {
        int tia = i + i + al, tiam1 = tia - 1, tiam2 = tia - 2;
        double ai = (double)(i + al) / (tiam1 * tia);
        double bi = (double)al / (tiam2 * tia);
        double ci = (i - 1.0) / (tiam2 * tiam1);
        L[a][i] = ai * pt[a - 1][i] + bi * t * pt[a - 1][i - 1] - ci * tt * pt[a - 1][i - 2];
        if (!idec) {
          R[a][i - 1] = -(i - 1) * (pt[a - 1][i - 1] + t * pt[a - 1][i - 2]);
          R[a][i - 1] = R[a][i - 1] / tiam2;
        }
      }
    }
  }
}
/* Polynomials.F90:560-620 HomIJacobi : L[a][i], dL[a][i][3], a in 1..nord, i in 1..nord-(a-1) */
static void hom_ijacobi(const apair *S, int nord, int minalpha, int idec, double L[NQ + 1][NQ + 1],
                        double dL[NQ + 1][NQ + 1][3]) {
  double hP[NQ + 1][NQ + 1], hR[NQ + 1][NQ + 1];
  if (idec) {
    poly_ijacobi(S->s[1], 1.0, nord, minalpha, 1, L, hP, hR);
    int ni = -1;
    for (int a = 1; a <= nord; a++) {
      ni++;
      for (int i = 1; i <= nord - ni; i++)
        for (int n = 0; n < 3; n++) dL[a][i][n] = hP[a][i - 1] * S->ds[1][n];
    }
  } else {
    poly_ijacobi(S->s[1], S->s[0] + S->s[1], nord, minalpha, 0, L, hP, hR);
    double ds01[3];
    for (int n = 0; n < 3; n++) ds01[n] = S->ds[0][n] + S->ds[1][n];
    int ni = -1;
    for (int a = 1; a <= nord; a++) {
      ni++;
      for (int i = 1; i <= nord - ni; i++)
        for (int n = 0; n < 3; n++) dL[a][i][n] = hP[a][i - 1] * S->ds[1][n] + hR[a][i - 1] * ds01[n];
    }
  }
}
static void tri_sl(const atri *S, apair *sL) { /* (s0+s1, s2) with gradients: Ancillary.F90:431-434 */
  sL->s[0] = S->s[0] + S->s[1]; sL->s[1] = S->s[2];
  for (int n = 0; n < 3; n++) { sL->ds[0][n] = S->ds[0][n] + S->ds[1][n]; sL->ds[1][n] = S->ds[2][n]; }
}
static void tri_pair(const atri *S, int a, int b, apair *out) {
  out->s[0] = S->s[a]; out->s[1] = S->s[b];
  for (int n = 0; n < 3; n++) { out->ds[0][n] = S->ds[a][n]; out->ds[1][n] = S->ds[b][n]; }
}
static void tri_perm(const atri *S, const int p[3], atri *G) {
  for (int k = 0; k < 3; k++) { G->s[k] = S->s[p[k]]; for (int n = 0; n < 3; n++) G->ds[k][n] = S->ds[p[k]][n]; }
}
/* Orient.F90:119-160 OrientTri */
static const int OT_PERM[6][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}, {0, 2, 1}, {1, 0, 2}, {2, 1, 0}};
static void orient_tri(const atri *S, int nori, atri *G) {
  if (nori < 0 || nori > 5) { fprintf(stderr, "orient_tri: invalid orientation %d\n", nori); exit(1); }
  tri_perm(S, OT_PERM[nori], G);
}
/* Ancillary.F90:397-450 AncPhiTri : phi[i][j], i in 2..nord-1, j in 1..nord-2, i+j <= nord */
static void anc_phi_tri(const atri *S, int nord, int idec, double phi[NQ][NQ], double dphi[NQ][NQ][3]) {
  if (nord < 3) return;
  double pE[NQ], dE[NQ][3], hL[NQ + 1][NQ + 1], dhL[NQ + 1][NQ + 1][3];
  apair s01, sL;
  tri_pair(S, 0, 1, &s01);
  anc_phiE(&s01, nord - 1, 0, pE, dE);
  tri_sl(S, &sL);
  hom_ijacobi(&sL, nord - 2, 4, idec, hL, dhL); /* hL[i-1][j] = homLal(i,j), alpha = 2i */
  for (int nij = 3; nij <= nord; nij++)
    for (int i = 2; i <= nij - 1; i++) {
      int j = nij - i;
      phi[i][j] = pE[i] * hL[i - 1][j];
      for (int n = 0; n < 3; n++) dphi[i][j][n] = hL[i - 1][j] * dE[i][n] + pE[i] * dhL[i - 1][j][n];
    }
}
/* Ancillary.F90:473-530 AncETri : E[i][j], i in 0..nord-2, j in 1..nord-1, i+j <= nord-1 */
static void anc_E_tri(const atri *S, int nord, int idec, double E[NQ][NQ][3], double cE[NQ][NQ][3]) {
  if (nord < 2) return;
  double EE[NQ][3], cEE[NQ][3], hL[NQ + 1][NQ + 1], dhL[NQ + 1][NQ + 1][3], x[3];
  apair s01, sL;
  tri_pair(S, 0, 1, &s01);
  anc_EE(&s01, nord - 1, 0, EE, cEE);
  tri_sl(S, &sL);
  hom_ijacobi(&sL, nord - 1, 1, idec, hL, dhL); /* hL[i+1][j] = homLal(i,j), alpha = 2i+1 */
  for (int nij = 1; nij <= nord - 1; nij++)
    for (int i = 0; i <= nij - 1; i++) {
      int j = nij - i;
      cross3(dhL[i + 1][j], EE[i], x);
      for (int n = 0; n < 3; n++) { E[i][j][n] = EE[i][n] * hL[i + 1][j]; cE[i][j][n] = hL[i + 1][j] * cEE[i][n] + x[n]; }
    }
}
/* Ancillary.F90:553-620 AncVTri : V[i][j], i,j in 0..nord-1, i+j <= nord-1 */
static void anc_V_tri(const atri *S, int nord, int idec, double V[NQ][NQ][3], double dV[NQ][NQ]) {
  if (nord < 1) return;
  double hP[NQ + 2], hPal[NQ + 1][NQ + 1], V00[3];
  apair s01;
  tri_pair(S, 0, 1, &s01);
  hom_legendre(&s01, nord - 1, hP);
  poly_jacobi(S->s[2], S->s[0] + S->s[1] + S->s[2], nord - 1, 1, hPal); /* HomJacobi((s0+s1, s2)) */
  if (idec) {
    cross3(S->ds[1], S->ds[2], V00);
    for (int nij = 0; nij <= nord - 1; nij++)
      for (int i = 0; i <= nij; i++) {
        int j = nij - i;
        for (int n = 0; n < 3; n++) V[i][j][n] = hP[i] * hPal[i][j] * V00[n];
        dV[i][j] = 0.0;
      }
  } else {
    double c01[3], c12[3], c20[3];
    cross3(S->ds[0], S->ds[1], c01); cross3(S->ds[1], S->ds[2], c12); cross3(S->ds[2], S->ds[0], c20);
    for (int n = 0; n < 3; n++) V00[n] = S->s[0] * c12[n] + S->s[1] * c20[n] + S->s[2] * c01[n];
    double triple = S->ds[0][0] * c12[0] + S->ds[0][1] * c12[1] + S->ds[0][2] * c12[2];
    for (int nij = 0; nij <= nord - 1; nij++)
      for (int i = 0; i <= nij; i++) {
        int j = nij - i;
        double psi = hP[i] * hPal[i][j];
        for (int n = 0; n < 3; n++) V[i][j][n] = psi * V00[n];
        dV[i][j] = (nij + 3) * psi * triple;
      }
  }
}

/* ------------------------------------------------------------------ triangle (2-D, embedded in R^3) */
static void affine_tri(const double x[2], atri *Nu) { /* AffineCoordinates.F90:42 */
  Nu->s[0] = 1.0 - x[0] - x[1]; Nu->s[1] = x[0]; Nu->s[2] = x[1];
  memset(Nu->ds, 0, sizeof Nu->ds);
  Nu->ds[0][0] = -1.0; Nu->ds[1][0] = 1.0;
  Nu->ds[0][1] = -1.0; Nu->ds[2][1] = 1.0;
}
static const int TRI_EDGE[3][2] = {{0, 1}, {1, 2}, {0, 2}}; /* ProjectTriE, BlendProject.F90:162-195 */

/* Triangle.F90:30 shape2DHTri ; nord[4] = 3 edges + face ; gradH (2,n) */
int orc_shape2DH_tri(const double x[2], const int nord[4], const int norie[3], double *shapH, double *gradH) {
  atri Nu;
  affine_tri(x, &Nu);
  int m = 0;
  for (int v = 0; v < 3; v++) { shapH[m] = Nu.s[v]; gradH[2 * m] = Nu.ds[v][0]; gradH[2 * m + 1] = Nu.ds[v][1]; m++; }
  for (int e = 0; e < 3; e++) {
    if (nord[e] - 1 <= 0) continue;
    apair P, G;
    double phi[NQ], dphi[NQ][3];
    tri_pair(&Nu, TRI_EDGE[e][0], TRI_EDGE[e][1], &P);
    orient_edge(&P, norie[e], &G);
    anc_phiE(&G, nord[e], 0, phi, dphi);
    for (int i = 2; i <= nord[e]; i++) { shapH[m] = phi[i]; gradH[2 * m] = dphi[i][0]; gradH[2 * m + 1] = dphi[i][1]; m++; }
  }
  int nf = nord[3];
  if ((nf - 1) * (nf - 2) / 2 > 0) {
    double phi[NQ][NQ], dphi[NQ][NQ][3];
    anc_phi_tri(&Nu, nf, 1, phi, dphi);
    for (int nij = 3; nij <= nf; nij++)
      for (int i = 2; i <= nij - 1; i++) {
        int j = nij - i;
        shapH[m] = phi[i][j]; gradH[2 * m] = dphi[i][j][0]; gradH[2 * m + 1] = dphi[i][j][1]; m++;
      }
  }
  return m;
}
/* Triangle.F90:140 shape2DETri ; shapE (2,n), curlE n */
int orc_shape2DE_tri(const double x[2], const int nord[4], const int norie[3], double *shapE, double *curlE) {
  atri Nu;
  affine_tri(x, &Nu);
  int m = 0;
  for (int e = 0; e < 3; e++) {
    if (nord[e] <= 0) continue;
    apair P, G;
    double EE[NQ][3], cEE[NQ][3];
    tri_pair(&Nu, TRI_EDGE[e][0], TRI_EDGE[e][1], &P);
    orient_edge(&P, norie[e], &G);
    anc_EE(&G, nord[e], 0, EE, cEE);
    for (int i = 0; i <= nord[e] - 1; i++) { shapE[2 * m] = EE[i][0]; shapE[2 * m + 1] = EE[i][1]; curlE[m] = cEE[i][2]; m++; }
  }
  int nf = nord[3];
  if (nf * (nf - 1) / 2 > 0) {
    int famctr = m;
    for (int fam = 0; fam < 2; fam++) {
      int mm = famctr + fam - 1, p[3] = {fam % 3, (fam + 1) % 3, (fam + 2) % 3};
      atri G;
      double E[NQ][NQ][3], cE[NQ][NQ][3];
      tri_perm(&Nu, p, &G);
      anc_E_tri(&G, nf, 1, E, cE);
      for (int nij = 1; nij <= nf - 1; nij++)
        for (int i = 0; i <= nij - 1; i++) {
          int j = nij - i;
          mm += 2;
          shapE[2 * (mm - 1)] = E[i][j][0]; shapE[2 * (mm - 1) + 1] = E[i][j][1]; curlE[mm - 1] = cE[i][j][2];
          if (mm > m) m = mm;
        }
    }
  }
  return m;
}
/* Triangle.F90:270 shape2DVTri : rotated H(curl) */
int orc_shape2DV_tri(const double x[2], const int nord[4], const int norie[3], double *shapV, double *divV) {
  int n = orc_shape2DE_tri(x, nord, norie, shapV, divV);
  for (int m = 0; m < n; m++) { double e1 = shapV[2 * m], e2 = shapV[2 * m + 1]; shapV[2 * m] = e2; shapV[2 * m + 1] = -e1; }
  return n;
}
/* Triangle.F90:330 shape2DQTri */
int orc_shape2DQ_tri(const double x[2], int nordf, double *shapQ) {
  atri Nu;
  affine_tri(x, &Nu);
  int m = 0;
  if ((nordf + 1) * nordf / 2 <= 0) return 0;
  double hP[NQ + 2], hPal[NQ + 1][NQ + 1];
  apair s01;
  tri_pair(&Nu, 0, 1, &s01);
  hom_legendre(&s01, nordf - 1, hP);
  poly_jacobi(Nu.s[2], Nu.s[0] + Nu.s[1] + Nu.s[2], nordf - 1, 1, hPal);
  for (int nij = 0; nij <= nordf - 1; nij++)
    for (int i = 0; i <= nij; i++) { int j = nij - i; shapQ[m++] = hP[i] * hPal[i][j]; }
  return m;
}

/* ------------------------------------------------------------------ prism */
static void affine_prism(const double x[3], apair *Mu, atri *Nu) { /* AffineCoordinates.F90:91 */
  Nu->s[0] = 1.0 - x[0] - x[1]; Nu->s[1] = x[0]; Nu->s[2] = x[1];
  memset(Nu->ds, 0, sizeof Nu->ds);
  Nu->ds[0][0] = -1.0; Nu->ds[1][0] = 1.0;
  Nu->ds[0][1] = -1.0; Nu->ds[2][1] = 1.0;
  Mu->s[0] = 1.0 - x[2]; Mu->s[1] = x[2];
  memset(Mu->ds, 0, sizeof Mu->ds);
  Mu->ds[0][2] = -1.0; Mu->ds[1][2] = 1.0;
}
/* mixed edges e=1..6: blend Mu(e>3), project on the Nu pair TRI_EDGE[(e-1)%3]   (BlendProjectPrisME)
 * quad  edges e=7..9: blend Nu(e-7), project on Mu                               (BlendProjectPrisQE)
 * triangle faces f=1,2: blend Mu(f-1)                                            (BlendProjectPrisTF)
 * quad faces f=3..5: (S,T) = (Nu pair TRI_EDGE[f-3], Mu), IdecQF = (false,true)  (ProjectPrisQF) */

/* Prism.F90:38 shape3DHPris ; nord[15] = 9 edges, 2 tri faces, 3 quad faces, middle */
int orc_shape3DH_pris(const double x[3], const int nord[15], const int norie[9], const int norif[5], double *shapH,
                      double *gradH) {
  apair Mu; atri Nu;
  affine_prism(x, &Mu, &Nu);
  int m = 0;
  for (int v = 0; v < 6; v++) {
    int a = v % 3, b = v / 3;
    shapH[m] = Nu.s[a] * Mu.s[b];
    for (int n = 0; n < 3; n++) gradH[3 * m + n] = Nu.ds[a][n] * Mu.s[b] + Nu.s[a] * Mu.ds[b][n];
    m++;
  }
  for (int e = 0; e < 6; e++) {
    if (nord[e] - 1 <= 0) continue;
    int b = e / 3;
    apair P, G;
    double phi[NQ], dphi[NQ][3];
    tri_pair(&Nu, TRI_EDGE[e % 3][0], TRI_EDGE[e % 3][1], &P);
    orient_edge(&P, norie[e], &G);
    anc_phiE(&G, nord[e], 0, phi, dphi);
    for (int i = 2; i <= nord[e]; i++) {
      shapH[m] = phi[i] * Mu.s[b];
      for (int n = 0; n < 3; n++) gradH[3 * m + n] = dphi[i][n] * Mu.s[b] + phi[i] * Mu.ds[b][n];
      m++;
    }
  }
  for (int e = 0; e < 3; e++) {
    if (nord[6 + e] - 1 <= 0) continue;
    apair G;
    double phi[NQ], dphi[NQ][3];
    orient_edge(&Mu, norie[6 + e], &G);
    anc_phiE(&G, nord[6 + e], 1, phi, dphi);
    for (int i = 2; i <= nord[6 + e]; i++) {
      shapH[m] = phi[i] * Nu.s[e];
      for (int n = 0; n < 3; n++) gradH[3 * m + n] = dphi[i][n] * Nu.s[e] + phi[i] * Nu.ds[e][n];
      m++;
    }
  }
  double phiT[NQ][NQ], dphiT[NQ][NQ][3];
  for (int f = 0; f < 2; f++) {
    int nf = nord[9 + f];
    if ((nf - 1) * (nf - 2) / 2 <= 0) continue;
    atri G;
    orient_tri(&Nu, norif[f], &G);
    anc_phi_tri(&G, nf, 1, phiT, dphiT);
    for (int nij = 3; nij <= nf; nij++)
      for (int i = 2; i <= nij - 1; i++) {
        int j = nij - i;
        shapH[m] = phiT[i][j] * Mu.s[f];
        for (int n = 0; n < 3; n++) gradH[3 * m + n] = dphiT[i][j][n] * Mu.s[f] + phiT[i][j] * Mu.ds[f][n];
        m++;
      }
  }
  double phiQ[NQ][NQ], dphiQ[NQ][NQ][3];
  for (int f = 0; f < 3; f++) {
    int nf[2];
    orc_decod(nord[11 + f], 10, 2, nf);
    if ((nf[0] - 1) * (nf[1] - 1) <= 0) continue;
    apair ST[2], G[2];
    tri_pair(&Nu, TRI_EDGE[f][0], TRI_EDGE[f][1], &ST[0]);
    ST[1] = Mu;
    const int idecST[2] = {0, 1};
    const int o = norif[2 + f];
    orient_quad(ST, o, G);
    int gidec[2] = {idecST[OQ_SWAP[o] ? 1 : 0], idecST[OQ_SWAP[o] ? 0 : 1]};
    anc_phi_quad(G, nf, gidec, phiQ, dphiQ);
    for (int j = 2; j <= nf[1]; j++)
      for (int i = 2; i <= nf[0]; i++) {
        shapH[m] = phiQ[i][j];
        for (int n = 0; n < 3; n++) gradH[3 * m + n] = dphiQ[i][j][n];
        m++;
      }
  }
  int nb[2];
  orc_decod(nord[14], 10, 2, nb);
  if ((nb[0] - 1) * (nb[0] - 2) * (nb[1] - 1) / 2 > 0) {
    double phi[NQ], dphi[NQ][3];
    anc_phi_tri(&Nu, nb[0], 1, phiT, dphiT);
    anc_phiE(&Mu, nb[1], 1, phi, dphi);
    for (int k = 2; k <= nb[1]; k++)
      for (int nij = 3; nij <= nb[0]; nij++)
        for (int i = 2; i <= nij - 1; i++) {
          int j = nij - i;
          shapH[m] = phiT[i][j] * phi[k];
          for (int n = 0; n < 3; n++) gradH[3 * m + n] = dphiT[i][j][n] * phi[k] + phiT[i][j] * dphi[k][n];
          m++;
        }
  }
  return m;
}

/* Prism.F90:358 shape3DEPris */
int orc_shape3DE_pris(const double x[3], const int nord[15], const int norie[9], const int norif[5], double *shapE,
                      double *curlE) {
  apair Mu; atri Nu;
  affine_prism(x, &Mu, &Nu);
  int m = 0;
  double ct[3];
  for (int e = 0; e < 6; e++) {
    if (nord[e] <= 0) continue;
    int b = e / 3;
    apair P, G;
    double EE[NQ][3], cEE[NQ][3];
    tri_pair(&Nu, TRI_EDGE[e % 3][0], TRI_EDGE[e % 3][1], &P);
    orient_edge(&P, norie[e], &G);
    anc_EE(&G, nord[e], 0, EE, cEE);
    for (int i = 0; i <= nord[e] - 1; i++) {
      cross3(Mu.ds[b], EE[i], ct);
      for (int n = 0; n < 3; n++) { shapE[3 * m + n] = Mu.s[b] * EE[i][n]; curlE[3 * m + n] = Mu.s[b] * cEE[i][n] + ct[n]; }
      m++;
    }
  }
  for (int e = 0; e < 3; e++) {
    if (nord[6 + e] <= 0) continue;
    apair G;
    double EE[NQ][3], cEE[NQ][3];
    orient_edge(&Mu, norie[6 + e], &G);
    anc_EE(&G, nord[6 + e], 1, EE, cEE);
    for (int i = 0; i <= nord[6 + e] - 1; i++) {
      cross3(Nu.ds[e], EE[i], ct);
      for (int n = 0; n < 3; n++) { shapE[3 * m + n] = Nu.s[e] * EE[i][n]; curlE[3 * m + n] = ct[n]; }
      m++;
    }
  }
  double ET[NQ][NQ][3], cET[NQ][NQ][3];
  for (int f = 0; f < 2; f++) {
    int nf = nord[9 + f];
    if (nf * (nf - 1) / 2 <= 0) continue;
    atri G, Gp;
    orient_tri(&Nu, norif[f], &G);
    int famctr = m;
    for (int fam = 0; fam < 2; fam++) {
      int mm = famctr + fam - 1, p[3] = {fam % 3, (fam + 1) % 3, (fam + 2) % 3};
      tri_perm(&G, p, &Gp);
      anc_E_tri(&Gp, nf, 1, ET, cET);
      for (int nij = 1; nij <= nf - 1; nij++)
        for (int i = 0; i <= nij - 1; i++) {
          int j = nij - i;
          mm += 2;
          cross3(Mu.ds[f], ET[i][j], ct);
          for (int n = 0; n < 3; n++) {
            shapE[3 * (mm - 1) + n] = ET[i][j][n] * Mu.s[f];
            curlE[3 * (mm - 1) + n] = Mu.s[f] * cET[i][j][n] + ct[n];
          }
          if (mm > m) m = mm;
        }
    }
  }
  double EQ[NQ][NQ][3], cEQ[NQ][NQ][3];
  for (int f = 0; f < 3; f++) {
    int nf[2];
    orc_decod(nord[11 + f], 10, 2, nf);
    apair ST[2], G[2];
    tri_pair(&Nu, TRI_EDGE[f][0], TRI_EDGE[f][1], &ST[0]);
    ST[1] = Mu;
    const int idecST[2] = {0, 1};
    const int o = norif[2 + f];
    orient_quad(ST, o, G);
    int gidec[2] = {idecST[OQ_SWAP[o] ? 1 : 0], idecST[OQ_SWAP[o] ? 0 : 1]};
    for (int fam = 0; fam < 2; fam++) {
      int a = fam, bb = 1 - fam;
      if (nf[a] * (nf[bb] - 1) <= 0) continue;
      apair Gab[2] = {G[a], G[bb]};
      int nab[2] = {nf[a], nf[bb]}, idab[2] = {gidec[a], gidec[bb]};
      anc_E_quad(Gab, nab, idab, EQ, cEQ);
      int lo[2], hi[2];
      lo[a] = 0; hi[a] = nf[a] - 1; lo[bb] = 2; hi[bb] = nf[bb];
      for (int jg = lo[1]; jg <= hi[1]; jg++)
        for (int ig = lo[0]; ig <= hi[0]; ig++) {
          int g[2] = {ig, jg};
          int i = g[a], j = g[bb];
          for (int n = 0; n < 3; n++) { shapE[3 * m + n] = EQ[i][j][n]; curlE[3 * m + n] = cEQ[i][j][n]; }
          m++;
        }
    }
  }
  int nb[2];
  orc_decod(nord[14], 10, 2, nb);
  if (nb[0] * (nb[0] - 1) * (nb[1] - 1) / 2 > 0) { /* families 1,2: triangle-type */
    double phi[NQ], dphi[NQ][3];
    anc_phiE(&Mu, nb[1], 1, phi, dphi);
    int famctr = m;
    for (int fam = 0; fam < 2; fam++) {
      int mm = famctr + fam - 1, p[3] = {fam % 3, (fam + 1) % 3, (fam + 2) % 3};
      atri Gp;
      tri_perm(&Nu, p, &Gp);
      anc_E_tri(&Gp, nb[0], 1, ET, cET);
      for (int k = 2; k <= nb[1]; k++)
        for (int nij = 1; nij <= nb[0] - 1; nij++)
          for (int i = 0; i <= nij - 1; i++) {
            int j = nij - i;
            mm += 2;
            cross3(dphi[k], ET[i][j], ct);
            for (int n = 0; n < 3; n++) {
              shapE[3 * (mm - 1) + n] = ET[i][j][n] * phi[k];
              curlE[3 * (mm - 1) + n] = phi[k] * cET[i][j][n] + ct[n];
            }
            if (mm > m) m = mm;
          }
    }
  }
  if ((nb[0] - 1) * (nb[0] - 2) * nb[1] / 2 > 0) { /* family 3: quadrilateral-type */
    double phiT[NQ][NQ], dphiT[NQ][NQ][3], EE[NQ][3], cEE[NQ][3];
    anc_phi_tri(&Nu, nb[0], 1, phiT, dphiT);
    anc_EE(&Mu, nb[1], 1, EE, cEE);
    for (int k = 0; k <= nb[1] - 1; k++)
      for (int nij = 3; nij <= nb[0]; nij++)
        for (int i = 2; i <= nij - 1; i++) {
          int j = nij - i;
          cross3(dphiT[i][j], EE[k], ct);
          for (int n = 0; n < 3; n++) { shapE[3 * m + n] = phiT[i][j] * EE[k][n]; curlE[3 * m + n] = ct[n]; }
          m++;
        }
  }
  return m;
}

/* Prism.F90:760 shape3DVPris */
int orc_shape3DV_pris(const double x[3], const int nord[15], const int norif[5], double *shapV, double *divV) {
  apair Mu; atri Nu;
  affine_prism(x, &Mu, &Nu);
  int m = 0;
  double VT[NQ][NQ][3], dVT[NQ][NQ];
  for (int f = 0; f < 2; f++) {
    int nf = nord[9 + f];
    if ((nf + 1) * nf / 2 <= 0) continue;
    atri G;
    orient_tri(&Nu, norif[f], &G);
    anc_V_tri(&G, nf, 1, VT, dVT);
    for (int nij = 0; nij <= nf - 1; nij++)
      for (int i = 0; i <= nij; i++) {
        int j = nij - i;
        double d = 0;
        for (int n = 0; n < 3; n++) { shapV[3 * m + n] = Mu.s[f] * VT[i][j][n]; d += Mu.ds[f][n] * VT[i][j][n]; }
        divV[m] = d;
        m++;
      }
  }
  double VQ[NQ][NQ][3], dVQ[NQ][NQ];
  for (int f = 0; f < 3; f++) {
    int nf[2];
    orc_decod(nord[11 + f], 10, 2, nf);
    if (nf[0] * nf[1] <= 0) continue;
    apair ST[2], G[2];
    tri_pair(&Nu, TRI_EDGE[f][0], TRI_EDGE[f][1], &ST[0]);
    ST[1] = Mu;
    const int idecST[2] = {0, 1};
    const int o = norif[2 + f];
    orient_quad(ST, o, G);
    int gidec[2] = {idecST[OQ_SWAP[o] ? 1 : 0], idecST[OQ_SWAP[o] ? 0 : 1]};
    anc_V_quad(G, nf, gidec, VQ, dVQ);
    for (int j = 0; j <= nf[1] - 1; j++)
      for (int i = 0; i <= nf[0] - 1; i++) {
        for (int n = 0; n < 3; n++) shapV[3 * m + n] = VQ[i][j][n];
        divV[m] = dVQ[i][j];
        m++;
      }
  }
  int nb[2];
  orc_decod(nord[14], 10, 2, nb);
  if (nb[0] * (nb[0] - 1) * nb[1] / 2 > 0) { /* families 1,2 */
    double ET[NQ][NQ][3], cET[NQ][NQ][3], EE[NQ][3], cEE[NQ][3];
    anc_EE(&Mu, nb[1], 1, EE, cEE);
    int famctr = m;
    for (int fam = 0; fam < 2; fam++) {
      int mm = famctr + fam - 1, p[3] = {fam % 3, (fam + 1) % 3, (fam + 2) % 3};
      atri Gp;
      tri_perm(&Nu, p, &Gp);
      anc_E_tri(&Gp, nb[0], 1, ET, cET);
      for (int k = 0; k <= nb[1] - 1; k++)
        for (int nij = 1; nij <= nb[0] - 1; nij++)
          for (int i = 0; i <= nij - 1; i++) {
            int j = nij - i;
            mm += 2;
            cross3(ET[i][j], EE[k], shapV + 3 * (mm - 1));
            divV[mm - 1] = EE[k][0] * cET[i][j][0] + EE[k][1] * cET[i][j][1] + EE[k][2] * cET[i][j][2];
            if (mm > m) m = mm;
          }
    }
  }
  if ((nb[0] + 1) * nb[0] * (nb[1] - 1) / 2 > 0) { /* family 3 */
    double phi[NQ], dphi[NQ][3];
    anc_V_tri(&Nu, nb[0], 1, VT, dVT);
    anc_phiE(&Mu, nb[1], 1, phi, dphi);
    for (int k = 2; k <= nb[1]; k++)
      for (int nij = 0; nij <= nb[0] - 1; nij++)
        for (int i = 0; i <= nij; i++) {
          int j = nij - i;
          double d = 0;
          for (int n = 0; n < 3; n++) { shapV[3 * m + n] = phi[k] * VT[i][j][n]; d += dphi[k][n] * VT[i][j][n]; }
          divV[m] = d;
          m++;
        }
  }
  return m;
}

/* Prism.F90:1040 shape3DQPris */
int orc_shape3DQ_pris(const double x[3], const int nord[15], double *shapQ) {
  apair Mu; atri Nu;
  affine_prism(x, &Mu, &Nu);
  int nb[2], m = 0;
  orc_decod(nord[14], 10, 2, nb);
  double hP[NQ + 2], hPal[NQ + 1][NQ + 1], hPz[NQ + 2];
  apair s01;
  tri_pair(&Nu, 0, 1, &s01);
  hom_legendre(&s01, nb[0] - 1, hP);
  poly_jacobi(Nu.s[2], Nu.s[0] + Nu.s[1] + Nu.s[2], nb[0] - 1, 1, hPal);
  hom_legendre(&Mu, nb[1] - 1, hPz);
  for (int k = 0; k <= nb[1] - 1; k++)
    for (int nij = 0; nij <= nb[0] - 1; nij++)
      for (int i = 0; i <= nij; i++) { int j = nij - i; shapQ[m++] = hP[i] * hPal[i][j] * hPz[k]; }
  return m;
}

/* ------------------------------------------------------------------ broken prism: broken/BrokenPrism.F90:31,100,190,290 */
#define TRI_MAXN 128
int orc_shape3HH_pris(const double xi[3], int nordM, double *shapH, double *gradH) {
  int nb[2], m = 0;
  orc_decod(nordM, 10, 2, nb);
  const int no4[4] = {nb[0], nb[0], nb[0], nb[0]}, z3[3] = {0, 0, 0};
  double h12[TRI_MAXN], d12[2 * TRI_MAXN], h3[NQ], d3[NQ];
  int n12 = orc_shape2DH_tri(xi, no4, z3, h12, d12);
  int n3 = orc_shape1HH(xi[2], nb[1], h3, d3);
  for (int i3 = 0; i3 < n3; i3++)
    for (int i = 0; i < n12; i++) {
      shapH[m] = h12[i] * h3[i3];
      gradH[3 * m] = d12[2 * i] * h3[i3]; gradH[3 * m + 1] = d12[2 * i + 1] * h3[i3]; gradH[3 * m + 2] = h12[i] * d3[i3];
      m++;
    }
  return m;
}
int orc_shape3EE_pris(const double xi[3], int nordM, double *shapE, double *curlE) {
  int nb[2], m = 0;
  orc_decod(nordM, 10, 2, nb);
  const int no4[4] = {nb[0], nb[0], nb[0], nb[0]}, z3[3] = {0, 0, 0};
  double e12[2 * TRI_MAXN], c12[TRI_MAXN], h12[TRI_MAXN], d12[2 * TRI_MAXN], h3[NQ], d3[NQ], q3[NQ];
  int nE12 = orc_shape2DE_tri(xi, no4, z3, e12, c12);
  int nH3 = orc_shape1HH(xi[2], nb[1], h3, d3);
  int nH12 = orc_shape2DH_tri(xi, no4, z3, h12, d12);
  int nQ3 = orc_shape1QQ(xi[2], nb[1], q3);
  for (int i3 = 0; i3 < nH3; i3++)
    for (int i = 0; i < nE12; i++) {
      shapE[3 * m] = e12[2 * i] * h3[i3]; shapE[3 * m + 1] = e12[2 * i + 1] * h3[i3]; shapE[3 * m + 2] = 0.0;
      curlE[3 * m] = -e12[2 * i + 1] * d3[i3]; curlE[3 * m + 1] = e12[2 * i] * d3[i3]; curlE[3 * m + 2] = c12[i] * h3[i3];
      m++;
    }
  for (int i3 = 0; i3 < nQ3; i3++)
    for (int i = 0; i < nH12; i++) {
      shapE[3 * m] = 0.0; shapE[3 * m + 1] = 0.0; shapE[3 * m + 2] = h12[i] * q3[i3];
      curlE[3 * m] = d12[2 * i + 1] * q3[i3]; curlE[3 * m + 1] = -d12[2 * i] * q3[i3]; curlE[3 * m + 2] = 0.0;
      m++;
    }
  return m;
}
int orc_shape3VV_pris(const double xi[3], int nordM, double *shapV, double *divV) {
  int nb[2], m = 0;
  orc_decod(nordM, 10, 2, nb);
  const int no4[4] = {nb[0], nb[0], nb[0], nb[0]}, z3[3] = {0, 0, 0};
  double v12[2 * TRI_MAXN], dv12[TRI_MAXN], q12[TRI_MAXN], h3[NQ], d3[NQ], q3[NQ];
  int nV12 = orc_shape2DV_tri(xi, no4, z3, v12, dv12);
  int nQ3 = orc_shape1QQ(xi[2], nb[1], q3);
  int nQ12 = orc_shape2DQ_tri(xi, nb[0], q12);
  int nH3 = orc_shape1HH(xi[2], nb[1], h3, d3);
  for (int i3 = 0; i3 < nQ3; i3++)
    for (int i = 0; i < nV12; i++) {
      shapV[3 * m] = v12[2 * i] * q3[i3]; shapV[3 * m + 1] = v12[2 * i + 1] * q3[i3]; shapV[3 * m + 2] = 0.0;
      divV[m] = dv12[i] * q3[i3];
      m++;
    }
  for (int i3 = 0; i3 < nH3; i3++)
    for (int i = 0; i < nQ12; i++) {
      shapV[3 * m] = 0.0; shapV[3 * m + 1] = 0.0; shapV[3 * m + 2] = q12[i] * h3[i3];
      divV[m] = q12[i] * d3[i3];
      m++;
    }
  return m;
}
int orc_shape3QQ_pris(const double xi[3], int nordM, double *shapQ) {
  int nord[15];
  for (int i = 0; i < 11; i++) nord[i] = 1;
  nord[11] = nord[12] = nord[13] = 11;
  nord[14] = nordM;
  return orc_shape3DQ_pris(xi, nord, shapQ);
}
