// tri_space.hpp -- host-side evaluation of hp3D's TRIANGLE shape-function ingredients, the 2-D factors of the
// prism's tensor structure (north-star subsystem 1, prism branch).
//
// Every prism shape function of the reference (src/element/shape_1/Prism.F90:38,358,760,1040 and
// broken/BrokenPrism.F90) is   sign * T(x,y) * Z(z)   where Z is one of the 1-D tables of tables.hpp
// (H = [1-z, z, L_2..], dH, Q = [P_0..]) and T is a function on the master triangle built from the affine
// coordinates nu_0 = 1-x-y, nu_1 = x, nu_2 = y (AffineCoordinates.F90:42,91):
//   scalar  VERT a            nu_a
//           EDGE (a,b), i     [L_i](nu_b ; nu_a+nu_b)            scaled integrated Legendre  (Ancillary.F90:37, Polynomials.F90:109,497)
//           FACE (s0,s1,s2),i,j   [L_i](s1; s0+s1) [L^{2i}_j](s2; 1)   integrated Jacobi        (Ancillary.F90:397, Polynomials.F90:303,560)
//           L2   (s0,s1,s2),i,j   [P_i](s1; s0+s1) [P^{2i+1}_j](s2; s0+s1+s2)                   (Triangle.F90:330, Ancillary.F90:553)
//   vector  EDGE (a,b), i     [P_i](nu_b; nu_a+nu_b) (nu_a grad nu_b - nu_b grad nu_a)          (Ancillary.F90:88)
//           FACE (s0,s1,s2),i,j   [L^{2i+1}_j](s2; 1) * EDGE(s0,s1),i                            (Ancillary.F90:473)
// The reference carries hand-derived gradient / curl formulas through every routine; here the polynomials are
// evaluated on dual numbers (value + 2-D gradient), so gradients and curls follow from the values by the chain rule.
#pragma once
#include <cmath>
#include <vector>

namespace hp3d {

struct D2 {  // value and gradient with respect to the master-triangle coordinates
  double v, x, y;
};
inline D2 operator+(D2 a, D2 b) { return {a.v + b.v, a.x + b.x, a.y + b.y}; }
inline D2 operator-(D2 a, D2 b) { return {a.v - b.v, a.x - b.x, a.y - b.y}; }
inline D2 operator*(D2 a, D2 b) { return {a.v * b.v, a.x * b.v + a.v * b.x, a.y * b.v + a.v * b.y}; }
inline D2 operator*(double s, D2 a) { return {s * a.v, s * a.x, s * a.y}; }
inline D2 d2const(double c) { return {c, 0.0, 0.0}; }

constexpr int TRI_MAXORD = 10;

// scaled Legendre P_0..P_n (x;t)       i P_i = (2i-1)(2x-t) P_{i-1} - (i-1) t^2 P_{i-2}     (Polynomials.F90:50-62)
inline void d2_legendre(D2 x, D2 t, int n, D2 *P) {
  P[0] = d2const(1.0);
  if (n < 1) return;
  const D2 y = 2.0 * x - t, tt = t * t;
  P[1] = y;
  for (int i = 2; i <= n; i++) P[i] = (1.0 / i) * ((double)(2 * i - 1) * (y * P[i - 1]) - (double)(i - 1) * (tt * P[i - 2]));
}
// scaled integrated Legendre L_2..L_n (x;t) = (P_i - t^2 P_{i-2}) / (4i-2)                   (Polynomials.F90:133-144)
inline void d2_ilegendre(D2 x, D2 t, int n, D2 *L) {
  D2 P[TRI_MAXORD + 2];
  d2_legendre(x, t, n, P);
  const D2 tt = t * t;
  for (int i = 2; i <= n; i++) L[i] = (1.0 / (4 * i - 2)) * (P[i] - tt * P[i - 2]);
}
// scaled Jacobi P^alpha_0..n (x;t)                                                            (Polynomials.F90:197-262)
inline void d2_jacobi(D2 x, D2 t, int n, int al, D2 *P) {
  P[0] = d2const(1.0);
  if (n < 1) return;
  const D2 y = 2.0 * x - t, tt = t * t;
  P[1] = y + (double)al * x;
  for (int i = 2; i <= n; i++) {
    const double ai = 2.0 * i * (i + al) * (2 * i + al - 2), bi = 2 * i + al - 1, ci = (double)(2 * i + al) * (2 * i + al - 2),
                 di = 2.0 * (i + al - 1) * (i - 1) * (2 * i + al);
    P[i] = (1.0 / ai) * (bi * ((ci * y + (double)(al * al) * t) * P[i - 1]) - di * (tt * P[i - 2]));
  }
}
// scaled integrated Jacobi L^alpha_1..n (x;t)                                                 (Polynomials.F90:303-400)
inline void d2_ijacobi(D2 x, D2 t, int n, int al, D2 *L) {
  D2 P[TRI_MAXORD + 2];
  d2_jacobi(x, t, n, al, P);
  const D2 tt = t * t;
  L[1] = x;
  for (int i = 2; i <= n; i++) {
    const double tia = 2 * i + al, ai = (i + al) / ((tia - 1) * tia), bi = al / ((tia - 2) * tia), ci = (i - 1.0) / ((tia - 2) * (tia - 1));
    L[i] = ai * P[i] + bi * (t * P[i - 1]) - ci * (tt * P[i - 2]);
  }
}

enum TriKind { TK_ONE = 0, TK_VERT, TK_EDGE, TK_FACE, TK_L2, TK_VEDGE, TK_VFACE };
struct TriFn {
  int kind;
  int s[3];  // indices of the affine coordinates playing the roles (s0, s1, s2)
  int i, j;
  bool operator==(const TriFn &o) const { return kind == o.kind && s[0] == o.s[0] && s[1] == o.s[1] && s[2] == o.s[2] && i == o.i && j == o.j; }
};
inline TriFn tri_fn(int kind, int s0, int s1, int s2, int i, int j) { TriFn f; f.kind = kind; f.s[0] = s0; f.s[1] = s1; f.s[2] = s2; f.i = i; f.j = j; return f; }

inline void tri_affine(double x, double y, D2 nu[3]) {
  nu[0] = {1.0 - x - y, -1.0, -1.0};
  nu[1] = {x, 1.0, 0.0};
  nu[2] = {y, 0.0, 1.0};
}
// Whitney-type edge function  [P_i](s1; s0+s1) (s0 grad s1 - s1 grad s0)  as two dual components
inline void tri_vedge(D2 s0, D2 s1, int i, D2 E[2]) {
  D2 P[TRI_MAXORD + 2];
  d2_legendre(s1, s0 + s1, i, P);
  const D2 W0 = s1.x * s0 - s0.x * s1, W1 = s1.y * s0 - s0.y * s1;   // gradients of affine coordinates are constants
  E[0] = P[i] * W0; E[1] = P[i] * W1;
}
// out[0..2]: scalar kinds -> (value, d/dx, d/dy) ; vector kinds -> (E_x, E_y, curl = dE_y/dx - dE_x/dy)
inline void tri_eval(const TriFn &f, double x, double y, double out[3]) {
  D2 nu[3];
  tri_affine(x, y, nu);
  const D2 s0 = nu[f.s[0]], s1 = nu[f.s[1]], s2 = nu[f.s[2]];
  D2 L[TRI_MAXORD + 2];
  switch (f.kind) {
    case TK_ONE: out[0] = 1.0; out[1] = out[2] = 0.0; return;
    case TK_VERT: out[0] = s0.v; out[1] = s0.x; out[2] = s0.y; return;
    case TK_EDGE: { d2_ilegendre(s1, s0 + s1, f.i, L); out[0] = L[f.i].v; out[1] = L[f.i].x; out[2] = L[f.i].y; return; }
    case TK_FACE: {
      D2 LJ[TRI_MAXORD + 2];
      d2_ilegendre(s1, s0 + s1, f.i, L);
      d2_ijacobi(s2, d2const(1.0), f.j, 2 * f.i, LJ);
      const D2 r = L[f.i] * LJ[f.j];
      out[0] = r.v; out[1] = r.x; out[2] = r.y; return;
    }
    case TK_L2: {
      D2 P[TRI_MAXORD + 2], PJ[TRI_MAXORD + 2];
      d2_legendre(s1, s0 + s1, f.i, P);
      d2_jacobi(s2, s0 + s1 + s2, f.j, 2 * f.i + 1, PJ);
      const D2 r = P[f.i] * PJ[f.j];
      out[0] = r.v; out[1] = r.x; out[2] = r.y; return;
    }
    case TK_VEDGE: { D2 E[2]; tri_vedge(s0, s1, f.i, E); out[0] = E[0].v; out[1] = E[1].v; out[2] = E[1].x - E[0].y; return; }
    case TK_VFACE: {
      D2 E[2], LJ[TRI_MAXORD + 2];
      tri_vedge(s0, s1, f.i, E);
      d2_ijacobi(s2, d2const(1.0), f.j, 2 * f.i + 1, LJ);
      const D2 e0 = LJ[f.j] * E[0], e1 = LJ[f.j] * E[1];
      out[0] = e0.v; out[1] = e1.v; out[2] = e1.x - e0.y; return;
    }
  }
  out[0] = out[1] = out[2] = 0.0;
}

// a deduplicated list of triangle functions = the "2-D axis" of a prism family
struct TriList {
  std::vector<TriFn> f;
  int add(const TriFn &t) {
    for (size_t k = 0; k < f.size(); k++) if (f[k] == t) return (int)k;
    f.push_back(t);
    return (int)f.size() - 1;
  }
  int size() const { return (int)f.size(); }
};

}  // namespace hp3d
