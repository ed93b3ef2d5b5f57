// tables.hpp -- host-side 1-D ingredients of the hot path: Gauss-Legendre rules on [0,1] and the 1-D shape
// function tables H = [1-x, x, L_2..L_p], dH = [-1, 1, P_1..P_{p-1}], Q = [P_0..P_{p-1}] evaluated at them.
//
// Reference behaviour being reproduced:
//   * rules: src/element/quadrature/gauss_quadrature.F90:518-651 carries the nodes/weights on [-1,1] as decimal
//     literals with 15 digits after the point (n<=9) and maps them with x=(1+x)/2, w=w/2 (:749-750).  We compute
//     the nodes by Newton iteration in extended precision and round to the same 15 decimals, which reproduces
//     those literals exactly (checked against the oracle's table in tests/test_tables.py); n=10 is kept at full
//     double precision as in the reference.
//   * polynomials: src/element/shape_1/Polynomials.F90:34-62 (three-term recurrence of the shifted Legendre
//     polynomials, same operation order) and :109-147 (L_i = (P_i - P_{i-2})/(4i-2), dL_i/dx = P_{i-1}).
#pragma once
#include <cmath>
#include <vector>

namespace hp3d {

constexpr int MAXN1D = 10;  // Gauss table limit of the reference: <= 10 points / functions per axis

inline void gauss01(int n, double *x, double *w) {
  for (int j = 0; j < n; j++) {
    // j-th root (ascending) of the Legendre polynomial P_n on [-1,1]
    long double z = -cosl(3.14159265358979323846264338327950288L * (j + 0.75L) / (n + 0.5L));
    long double pp = 1.0L;
    for (int it = 0; it < 100; it++) {
      long double p1 = 1.0L, p2 = 0.0L;
      for (int k = 1; k <= n; k++) { long double p3 = p2; p2 = p1; p1 = ((2.0L * k - 1.0L) * z * p2 - (k - 1.0L) * p3) / k; }
      pp = n * (z * p1 - p2) / (z * z - 1.0L);
      long double dz = p1 / pp;
      z -= dz;
      if (fabsl(dz) < 1e-19L) break;
    }
    long double wt = 2.0L / ((1.0L - z * z) * pp * pp);
    if (2 * j + 1 == n) z = 0.0L;
    double xd = (double)z, wd = (double)wt;
    if (n <= 9) {  // the reference's literals: 15 decimals
      xd = (double)(roundl(z * 1e15L) / 1e15L);
      wd = (double)(roundl(wt * 1e15L) / 1e15L);
    }
    x[j] = 0.5 * (1.0 + xd);
    w[j] = 0.5 * wd;
  }
}

// shifted Legendre P_0..P_nord at x in [0,1]  (Polynomials.F90:50-62 with t=1)
inline void legendre01(double x, int nord, double *P) {
  P[0] = 1.0;
  double y = 0.0;
  if (nord >= 1) { y = 2.0 * x - 1.0; P[1] = y; }
  for (int i = 2; i <= nord; i++) {
    P[i] = (2 * i - 1) * y * P[i - 1] - (i - 1) * P[i - 2];
    P[i] = P[i] / i;
  }
}

// 1-D tables of order p at the points xs[0..nq): H[i*nq+q] (i=0..p), dH likewise, Q[i*nq+q] (i=0..p-1)
struct Tables1D {
  int p = 0, nq = 0;
  std::vector<double> x, w, H, dH, Q;
};

inline void eval_tables_1d(int p, int nq, const double *xs, double *H, double *dH, double *Q) {
  double P[MAXN1D + 2];
  for (int q = 0; q < nq; q++) {
    const double x = xs[q];
    legendre01(x, p, P);
    H[0 * nq + q] = 1.0 - x; dH[0 * nq + q] = -1.0;
    H[1 * nq + q] = x;       dH[1 * nq + q] = 1.0;
    for (int i = 2; i <= p; i++) {
      H[i * nq + q] = (P[i] - P[i - 2]) / (4 * i - 2);
      dH[i * nq + q] = P[i - 1];
    }
    for (int i = 0; i < p; i++) Q[i * nq + q] = P[i];
  }
}

inline Tables1D make_tables(int p, int nq) {
  Tables1D t;
  t.p = p; t.nq = nq;
  t.x.resize(nq); t.w.resize(nq);
  gauss01(nq, t.x.data(), t.w.data());
  t.H.assign((size_t)(p + 1) * nq, 0.0); t.dH.assign((size_t)(p + 1) * nq, 0.0); t.Q.assign((size_t)(p > 0 ? p : 1) * nq, 0.0);
  eval_tables_1d(p, nq, t.x.data(), t.H.data(), t.dH.data(), t.Q.data());
  return t;
}

}  // namespace hp3d
