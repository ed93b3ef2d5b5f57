// forms_prism.hpp -- form compiler for TRIANGULAR PRISMS (element type MDLP): same weak forms and the same dense-phase
// layout as forms.hpp, but a family is a (triangle list) x (1-D z table) grid instead of three 1-D axes:
//   FamilyDesc.n = {nT, 1, nZ},  FamilyDesc.tab = {offset of the list's block in the T-table pool, -, z table type}
//   TermDesc.dA / dB = component of the triangle table (0..2)  [scalar lists: value, d/dx, d/dy ; vector lists: E_x, E_y, curl]
//   SlotDesc.zA / zB = z table type (T_DH when the factor is differentiated in z)
// Quadrature: Dunavant rule NSELECT(p_xy [+dp]) x Gauss(p_z [+dp] + 1), triangle point fastest
// (src/element/quadrature/set_3D_int.F90:261-288).  Trace pairings <n x E^, F> and <sigma^.n, v> are metric-free (see
// forms.hpp) and are integrated here, once per signature, with the reference's own face rules
// (set_2D_int.F90:195-236; face parametrisation element_data.F90:554-603, normal sign :616-626).
#pragma once
#include "forms.hpp"
#include "prism_space.hpp"
#include "tri_rules.hpp"

namespace hp3d {
namespace detail {

enum PFamKind { PF_SCALAR = 0, PF_EH = 1, PF_EV = 2, PF_L2 = 3, PF_UNIT = 4 };
struct PFam { int id = -1, kind = 0, nT = 0, nZ = 0, off = 0; };   // off: first grid index inside its space
struct CompRef { int tc, zd, sgn; };                                // triangle-table component, z derivative?, sign ; tc < 0: absent

inline CompRef pf_val(int kind, int c) {
  switch (kind) {
    case PF_EH: if (c == 0) return {0, 0, 1}; if (c == 1) return {1, 0, 1}; return {-1, 0, 0};
    case PF_EV: if (c == 2) return {0, 0, 1}; return {-1, 0, 0};
    default: return {-1, 0, 0};
  }
}
inline CompRef pf_curl(int kind, int c) {
  switch (kind) {
    case PF_EH: if (c == 0) return {1, 1, -1}; if (c == 1) return {0, 1, 1}; return {2, 0, 1};
    case PF_EV: if (c == 0) return {2, 0, 1}; if (c == 1) return {1, 0, -1}; return {-1, 0, 0};
    default: return {-1, 0, 0};
  }
}
inline CompRef pf_grad(int c) { if (c == 0) return {1, 0, 1}; if (c == 1) return {2, 0, 1}; return {0, 1, 1}; }

// evaluate a triangle list at the nqt rule points into the pool: block [3][nT][nqt]; returns the pool offset
inline int add_tri_table(SigHost &S, const TriList &T, int nqt, const double (*pts)[3]) {
  const int off = (int)S.ttab.size(), nT = T.size();
  S.ttab.resize((size_t)off + 3 * (size_t)nT * nqt, 0.0);
  for (int t = 0; t < nT; t++)
    for (int q = 0; q < nqt; q++) {
      double o[3];
      tri_eval(T.f[t], pts[q][0], pts[q][1], o);
      for (int c = 0; c < 3; c++) S.ttab[(size_t)off + ((size_t)c * nT + t) * nqt + q] = o[c];
    }
  return off;
}
inline PFam add_pfam(SigHost &S, int kind, const TriList &T, int nZ, int ztab, int nqt, const double (*pts)[3]) {
  PFam f;
  f.kind = kind; f.nT = T.size(); f.nZ = nZ;
  FamilyDesc d; d.n[0] = f.nT; d.n[1] = 1; d.n[2] = nZ; d.tab[0] = add_tri_table(S, T, nqt, pts); d.tab[1] = T_ONE; d.tab[2] = ztab;
  S.fam.push_back(d);
  f.id = (int)S.fam.size() - 1;
  return f;
}
// grid map of a conforming space onto a family grid: entry (t + nT*z) -> +-(target+1)
template <class TargetFn>
inline int add_pgrid_map(SigHost &S, const std::vector<PrismDof> &dofs, int list, const PFam &F, TargetFn target) {
  const int off = (int)S.maps.size();
  S.maps.resize(off + (size_t)F.nT * F.nZ, 0);
  for (size_t k = 0; k < dofs.size(); k++) {
    const PrismDof &d = dofs[k];
    if (d.list != list) continue;
    S.maps[off + d.t + F.nT * d.zi] = d.sgn * (target((int)k) + 1);
  }
  return off;
}

// (grad A, grad B) weighted by D
inline void add_grad_grad(BlockBuilder &b, double c0) {
  for (int ca = 0; ca < 3; ca++)
    for (int cb = 0; cb < 3; cb++) {
      const CompRef A = pf_grad(ca), B = pf_grad(cb);
      b.addp(A.tc, A.zd, B.tc, B.zd, F_D + sym_idx(ca, cb), 1.0, c0, 0.0);
    }
}
// c_mass * (A, B)_D + c_curl * (curl A, curl B)_C into both channels with weights (m0, m1) / (k0, k1)
inline void add_hcurl_pair(BlockBuilder &b, int kindA, int kindB, double m0, double m1, double k0, double k1) {
  for (int ca = 0; ca < 3; ca++)
    for (int cb = 0; cb < 3; cb++) {
      const CompRef A = pf_val(kindA, ca), B = pf_val(kindB, cb);
      if (A.tc >= 0 && B.tc >= 0) b.addp(A.tc, A.zd, B.tc, B.zd, F_D + sym_idx(ca, cb), A.sgn * B.sgn, m0, m1);
      const CompRef CA = pf_curl(kindA, ca), CB = pf_curl(kindB, cb);
      if (CA.tc >= 0 && CB.tc >= 0) b.addp(CA.tc, CA.zd, CB.tc, CB.zd, F_C + sym_idx(ca, cb), CA.sgn * CB.sgn, k0, k1);
    }
}

// ---- pointwise evaluation through the T x Z decomposition (host-side face quadrature, tests)
struct TriVals { std::vector<double> v; int nT = 0; const double *at(int t) const { return v.data() + 3 * (size_t)t; } };
inline TriVals eval_list(const TriList &T, double x, double y) {
  TriVals r; r.nT = T.size(); r.v.resize(3 * (size_t)r.nT);
  for (int t = 0; t < r.nT; t++) tri_eval(T.f[t], x, y, r.v.data() + 3 * (size_t)t);
  return r;
}
struct ZVals { double H[MAXN1D + 2], dH[MAXN1D + 2], Q[MAXN1D + 2]; };
inline ZVals eval_z(int p, double z) { ZVals r; eval_tables_1d(p, 1, &z, r.H, r.dH, r.Q); return r; }

struct FaceRule { std::vector<double> xi, w; int n = 0; double nvec[3]; };   // 3-D points on the face, weights, nsign*(a1 x a2)
inline FaceRule prism_face_rule(const int norder[15], const int norif[5], int f, int integration, int cap) {
  FaceRule R;
  const bool tri = f < 2;
  const double *x1 = PR_COORD[PR_FACE_VERT[f][0] - 1], *x2 = PR_COORD[PR_FACE_VERT[f][1] - 1], *x3 = PR_COORD[PR_FACE_VERT[f][tri ? 2 : 3] - 1];
  double a1[3], a2[3];
  for (int c = 0; c < 3; c++) { a1[c] = x2[c] - x1[c]; a2[c] = x3[c] - x1[c]; }
  R.nvec[0] = PR_NSIGN[f] * (a1[1] * a2[2] - a1[2] * a2[1]);
  R.nvec[1] = PR_NSIGN[f] * (a1[2] * a2[0] - a1[0] * a2[2]);
  R.nvec[2] = PR_NSIGN[f] * (a1[0] * a2[1] - a1[1] * a2[0]);
  std::vector<double> t1, t2;
  if (tri) {
    int nord = norder[9 + f];
    for (int i = 0; i < 3; i++) nord = std::max(nord, norder[PR_FACE_EDGE[f][i] - 1]);
    nord = std::min(nord + integration, cap);
    const int n = TRI_RULE_NPTS[nord - 1], o = TRI_RULE_OFF[nord - 1];
    for (int l = 0; l < n; l++) { t1.push_back(TRI_RULE_PTS[o + l][0]); t2.push_back(TRI_RULE_PTS[o + l][1]); R.w.push_back(TRI_RULE_PTS[o + l][2]); }
  } else {
    int h = norder[9 + f] / 10, v = norder[9 + f] % 10;
    if (QSWAP_ORDER[norif[f]]) std::swap(h, v);
    int nx = std::max(std::max(norder[PR_FACE_EDGE[f][0] - 1], norder[PR_FACE_EDGE[f][2] - 1]), h);
    int ny = std::max(std::max(norder[PR_FACE_EDGE[f][1] - 1], norder[PR_FACE_EDGE[f][3] - 1]), v);
    nx = std::min(nx + integration, cap) + 1; ny = std::min(ny + integration, cap) + 1;
    const Tables1D tx = make_tables(1, nx), ty = make_tables(1, ny);
    for (int l2 = 0; l2 < ny; l2++)
      for (int l1 = 0; l1 < nx; l1++) { t1.push_back(tx.x[l1]); t2.push_back(ty.x[l2]); R.w.push_back(tx.w[l1] * ty.w[l2]); }
  }
  R.n = (int)R.w.size();
  R.xi.resize(3 * (size_t)R.n);
  for (int l = 0; l < R.n; l++)
    for (int c = 0; c < 3; c++) R.xi[3 * l + c] = x1[c] + t1[l] * a1[c] + t2[l] * a2[c];
  return R;
}

}  // namespace detail

// value (3 components) of conforming prism H(curl) dof d at a point, given the evaluated lists
inline void prism_hcurl_value(const PrismDof &d, const detail::TriVals &TV, const detail::TriVals &TS, const detail::ZVals &Z, double E[3]) {
  if (d.list == 0) { const double *t = TV.at(d.t); const double z = Z.H[d.zi] * d.sgn; E[0] = t[0] * z; E[1] = t[1] * z; E[2] = 0.0; }
  else { const double *t = TS.at(d.t); E[0] = E[1] = 0.0; E[2] = t[0] * Z.Q[d.zi] * d.sgn; }
}

// sizes_only: stop once the dof counts, quadrature size and padded dense extents are known (no blocks, no trace pairings)
inline bool compile_signature_prism(const FormParams &P, const int norder[19], const int norie[12], const int norif[6], SigHost &S, bool sizes_only = false) {
  using namespace detail;
  S = SigHost();
  S.kind = P.kind; S.etype = 3;
  memcpy(S.norder, norder, sizeof S.norder); memcpy(S.norie, norie, sizeof S.norie); memcpy(S.norif, norif, sizeof S.norif);
  const PrismOrders o = PrismOrders::decode(norder);
  for (int e = 0; e < 9; e++) if (o.edge[e] < 1 || o.edge[e] > 9 || (norie[e] != 0 && norie[e] != 1)) { S.err = "bad prism edge order/orientation"; return false; }
  for (int f = 0; f < 2; f++) if (o.tface[f] < 1 || o.tface[f] > 9 || norif[f] < 0 || norif[f] > 5) { S.err = "bad prism triangle-face order/orientation"; return false; }
  for (int f = 0; f < 3; f++) if (o.qface[f][0] < 1 || o.qface[f][1] < 1 || norif[2 + f] < 0 || norif[2 + f] > 7) { S.err = "bad prism quad-face order/orientation"; return false; }
  if (o.mid[0] < 1 || o.mid[1] < 1) { S.err = "bad prism middle node order"; return false; }
  const bool dpg = (P.kind == 2 || P.kind == 4);
  const int dp = dpg ? P.nord_add : 0;
  int pmax[2];
  prism_axis_max_order(norder, norif, pmax);
  const int cap = dpg ? P.maxp + 1 : P.maxp;
  const int ordh = std::min(pmax[0] + dp, cap), ordz = std::min(pmax[1] + dp, cap);
  const int pe = o.mid[0] + dp, pze = o.mid[1] + dp;           // enriched test order
  if (ordh > 9 || ordz + 1 > MAXQ || pe > TRI_MAXORD - 1 || pze > MAXQ - 1 || pmax[0] > TRI_MAXORD - 1) { S.err = "prism order exceeds the quadrature table limits"; return false; }
  const int nqt = TRI_RULE_NPTS[ordh - 1], nqz = ordz + 1;
  const double(*tpts)[3] = &TRI_RULE_PTS[TRI_RULE_OFF[ordh - 1]];
  S.nq[0] = nqt; S.nq[1] = 1; S.nq[2] = nqz;
  S.nint = nqt * nqz;
  // ---- z tables (axis 2 of the 12-table block; axes 0,1 unused) and weights: wq[0..nqt) triangle, wq[MAXQ'..] z
  const int ptab = std::max(pmax[1], pze);
  S.tab.assign((size_t)12 * TABSZ, 0.0);
  {
    Tables1D t = make_tables(ptab, nqz);
    std::copy(t.H.begin(), t.H.end(), S.tab.begin() + (2 * 4 + T_H) * TABSZ);
    std::copy(t.dH.begin(), t.dH.end(), S.tab.begin() + (2 * 4 + T_DH) * TABSZ);
    std::copy(t.Q.begin(), t.Q.end(), S.tab.begin() + (2 * 4 + T_Q) * TABSZ);
    for (int q = 0; q < nqz; q++) S.tab[(2 * 4 + T_ONE) * TABSZ + q] = 1.0;
    S.wq.assign((size_t)nqt + MAXQ, 0.0);
    for (int q = 0; q < nqt; q++) S.wq[q] = tpts[q][2];
    for (int q = 0; q < nqz; q++) S.wq[nqt + q] = t.w[q];
  }
  // ---- geometry dofs
  TriList TG;
  const std::vector<PrismDof> hd = prism_dofs_H1(norder, norie, norif, TG);
  S.nH = (int)hd.size();
  for (const PrismDof &d : hd) S.hdof.push_back(d.t | (d.zi << 8) | ((d.sgn < 0) << 24));
  int bH, bE, bV, bQ;
  prism_mid_counts(o.mid, bH, bE, bV, bQ);
  int norderi[15];
  memcpy(norderi, norder, sizeof norderi);
  norderi[14] = 11;   // trace variables: middle-node order forced to 11 (elem_opt.F90:190)
  TriList TOne; TOne.add(tri_fn(TK_ONE, 0, 1, 2, 0, 0));
  const std::complex<double> I(0, 1);
  // geometry list first in the pool (the geometry kernel reads it)
  S.geo_toff = add_tri_table(S, TG, nqt, tpts);
  S.geo_nT = TG.size();
  const PFam unit = add_pfam(S, PF_UNIT, TOne, 1, T_ONE, nqt, tpts);

  if (P.kind == 4) {
    // =============================================================== ultraweak Maxwell
    if (P.test_norm < 1 || P.test_norm > 3) { S.err = "unknown test norm"; return false; }
    TriList TV, TS, TQ, TEt, THt;
    const std::vector<PrismDof> ed = prism_dofs_Hcurl(norderi, norie, norif, TV, TS);
    const std::vector<PrismDof> qd = prism_dofs_L2(norder, TQ);
    const int nEi = (int)ed.size(), nQ = (int)qd.size();
    tri_list_Hcurl(pe, TEt); tri_list_H1(pe, THt);
    PFam tf[2];
    tf[0] = add_pfam(S, PF_EH, TEt, pze + 1, T_H, nqt, tpts);
    tf[1] = add_pfam(S, PF_EV, THt, pze, T_Q, nqt, tpts);
    tf[0].off = 0; tf[1].off = tf[0].nT * tf[0].nZ;
    const int nEE = tf[1].off + tf[1].nT * tf[1].nZ;
    const PFam fq = add_pfam(S, PF_L2, TQ, o.mid[1], T_Q, nqt, tpts);
    S.cplx = true; S.dpg = true; S.ntest = 2 * nEE; S.ni = 2 * nEi; S.nb = 6 * nQ;
    DenseDims &D = S.dims;
    const bool rs = rs_applicable(P);   // real-structured dense phase (forms.hpp)
    D.cplx = !rs; D.rs = rs; D.nload = (rs ? 2 : 1) * P.nrhs; D.dpg = true; D.n = S.ntest; D.nb = S.nb; D.ni = S.ni; D.finish();
    if (sizes_only) return true;
    const int rowB = D.np, rowI = D.np + D.nbp, rowL = rowI + D.nil - D.nload;   // load row(s): last padded interface rows, independent of ni
    const std::complex<double> za = I * P.omega * P.eps, zc = I * P.omega * P.mu;
    const double aF = (P.test_norm == 2) ? 1.0 : P.alpha_norm + std::norm(za);
    const double aG = (P.test_norm == 2) ? 1.0 : P.alpha_norm + std::norm(zc);
    for (int a = 0; a < 2; a++)
      for (int a2 = 0; a2 <= a; a2++) {
        BlockBuilder b(S, tf[a].id, tf[a2].id, channel(0, 0, tf[a].off, tf[a2].off), channel(0, 0, nEE + tf[a].off, nEE + tf[a2].off));
        if (P.tensor && P.test_norm != 2) {
          // permittivity tensor (forms.hpp): FF = alpha D + w^2 eps^2 TD, GG = aG D, written as aG D (1,1) + [(alpha - aG) D + w^2 eps^2 TD] (1,0)
          const double kap = P.omega * P.omega * P.eps * P.eps;
          for (int ca = 0; ca < 3; ca++)
            for (int cb = 0; cb < 3; cb++) {
              const CompRef A = pf_val(tf[a].kind, ca), B = pf_val(tf[a2].kind, cb);
              if (A.tc >= 0 && B.tc >= 0) {
                b.addp(A.tc, A.zd, B.tc, B.zd, F_D + sym_idx(ca, cb), A.sgn * B.sgn * aG, 1.0, 1.0);
                b.addp(A.tc, A.zd, B.tc, B.zd, F_D + sym_idx(ca, cb), A.sgn * B.sgn * (P.alpha_norm - aG), 1.0, 0.0);
                b.addp(A.tc, A.zd, B.tc, B.zd, F_TD + sym_idx(ca, cb), A.sgn * B.sgn * kap, 1.0, 0.0);
              }
              const CompRef CA = pf_curl(tf[a].kind, ca), CB = pf_curl(tf[a2].kind, cb);
              if (CA.tc >= 0 && CB.tc >= 0) b.addp(CA.tc, CA.zd, CB.tc, CB.zd, F_C + sym_idx(ca, cb), CA.sgn * CB.sgn, 1.0, 1.0);
            }
        } else
          add_hcurl_pair(b, tf[a].kind, tf[a2].kind, aF, aG, 1.0, 1.0);
        b.finish();
      }
    // imaginary part of the FF Gram blocks for a complex permittivity tensor: F_r^T (w^2 eps^2 S) F_c, S antisymmetric
    if (P.tensor && P.test_norm != 2 && !P.tensor_real())
      for (int a = 0; a < 2; a++)
        for (int a2 = 0; a2 <= a; a2++) {
          BlockBuilder b(S, tf[a].id, tf[a2].id, channel(0, 1, tf[a].off, tf[a2].off), no_channel());
          const double kap = P.omega * P.omega * P.eps * P.eps;
          for (int ca = 0; ca < 3; ca++)
            for (int cb = 0; cb < 3; cb++) {
              if (ca == cb) continue;
              const CompRef A = pf_val(tf[a].kind, ca), B = pf_val(tf[a2].kind, cb);
              if (A.tc < 0 || B.tc < 0) continue;
              const int hi = std::max(ca, cb), lo = std::min(ca, cb);
              b.addp(A.tc, A.zd, B.tc, B.zd, F_TS + (hi == 1 ? 0 : lo + 1), A.sgn * B.sgn * kap * (ca > cb ? 1.0 : -1.0), 1.0, 0.0);
            }
          b.finish();
        }
    if (P.test_norm == 1)
      for (int a = 0; a < 2; a++)        // G row family
        for (int a2 = 0; a2 < 2; a2++) { // F column family
          BlockBuilder b(S, tf[a].id, tf[a2].id, channel(0, 0, nEE + tf[a].off, tf[a2].off), rs ? no_channel() : channel(0, 1, nEE + tf[a].off, tf[a2].off));
          const std::complex<double> m1 = -std::conj(za), m2 = zc;
          const double m1r = rs ? rs_real(m1, 1, 0, false) : m1.real(), m1i = rs ? 0.0 : m1.imag();
          const double m2r = rs ? rs_real(m2, 1, 0, false) : m2.real(), m2i = rs ? 0.0 : m2.imag();
          for (int c = 0; c < 3; c++) {
            const CompRef cg = pf_curl(tf[a].kind, c), vf = pf_val(tf[a2].kind, c);
            if (P.tensor) {   // -(curl G)^T za^H F through the fields T1R / T1I (curl component c, value component c2), see forms.hpp
              const std::complex<double> m1t = m1 * std::complex<double>(0.0, -1.0);
              for (int c2 = 0; c2 < 3; c2++) {
                const CompRef v2 = pf_val(tf[a2].kind, c2);
                if (cg.tc < 0 || v2.tc < 0) continue;
                b.addp(cg.tc, cg.zd, v2.tc, v2.zd, F_T1R + 3 * c + c2, cg.sgn * v2.sgn, m1r, m1i);
                if (!P.tensor_real()) b.addp(cg.tc, cg.zd, v2.tc, v2.zd, F_T1I + 3 * c + c2, cg.sgn * v2.sgn, m1t.real(), m1t.imag());
              }
            } else if (cg.tc >= 0 && vf.tc >= 0) b.addp(cg.tc, cg.zd, vf.tc, vf.zd, F_W, cg.sgn * vf.sgn, m1r, m1i);
            const CompRef vg = pf_val(tf[a].kind, c), cf = pf_curl(tf[a2].kind, c);
            if (vg.tc >= 0 && cf.tc >= 0) b.addp(vg.tc, vg.zd, cf.tc, cf.zd, F_W, vg.sgn * cf.sgn, m2r, m2i);
          }
          b.finish();
        }
    int mapE[3], mapH[3];
    for (int c = 0; c < 3; c++) {
      mapE[c] = (int)S.maps.size();
      for (int j = 0; j < nQ; j++) S.maps.push_back(rowB + 6 * j + c + 1);
      mapH[c] = (int)S.maps.size();
      for (int j = 0; j < nQ; j++) S.maps.push_back(rowB + 6 * j + 3 + c + 1);
    }
    for (int c = 0; c < 3; c++)
      for (int a = 0; a < 2; a++) {
        {  // B(F_i, E_jc) = -za (E_c, F_i)
          BlockBuilder b(S, fq.id, tf[a].id, channel(0, 0, 0, tf[a].off, mapE[c]), rs ? no_channel() : channel(0, 1, 0, tf[a].off, mapE[c]));
          for (int c2 = 0; c2 < 3; c2++) {   // -(za E_c e_c, F) = -sum_c2 za(c2,c) F_c2: the identity tensor keeps c2 = c only
            const std::complex<double> m = P.tensor ? -std::conj(za * P.epst[c2 + 3 * c]) : (c2 == c ? -std::conj(za) : std::complex<double>(0.0, 0.0));
            if (m == std::complex<double>(0.0, 0.0)) continue;
            const double mr = rs ? rs_real(m, 0, 0, true) : m.real(), mi = rs ? 0.0 : m.imag();
            for (int d = 0; d < 3; d++) { const CompRef v = pf_val(tf[a].kind, d); if (v.tc >= 0) b.addp(0, 0, v.tc, v.zd, F_WJI + 3 * d + c2, v.sgn, mr, mi); }
          }
          b.finish();
        }
        {  // B(F_i, H_jc) = B(G_i, E_jc) = (H_c, curl F_i)
          BlockBuilder b(S, fq.id, tf[a].id, channel(0, 0, 0, tf[a].off, mapH[c]), channel(0, 0, 0, nEE + tf[a].off, mapE[c]));
          const double c0 = rs ? rs_real(1.0, 1, 0, true) : 1.0, c1 = rs ? rs_real(1.0, 0, 1, true) : 1.0;
          for (int d = 0; d < 3; d++) { const CompRef v = pf_curl(tf[a].kind, d); if (v.tc >= 0) b.addp(0, 0, v.tc, v.zd, F_WJD + 3 * c + d, v.sgn, c0, c1); }
          b.finish();
        }
        {  // B(G_i, H_jc) = zc (H_c, G_i)
          const std::complex<double> m = std::conj(zc);
          BlockBuilder b(S, fq.id, tf[a].id, channel(0, 0, 0, nEE + tf[a].off, mapH[c]), rs ? no_channel() : channel(0, 1, 0, nEE + tf[a].off, mapH[c]));
          const double mr = rs ? rs_real(m, 1, 1, true) : m.real(), mi = rs ? 0.0 : m.imag();
          for (int d = 0; d < 3; d++) { const CompRef v = pf_val(tf[a].kind, d); if (v.tc >= 0) b.addp(0, 0, v.tc, v.zd, F_WJI + 3 * d + c, v.sgn, mr, mi); }
          b.finish();
        }
      }
    for (int a = 0; a < 2; a++) {   // load
      BlockBuilder b(S, unit.id, tf[a].id, channel(0, 0, rowL, tf[a].off), rs ? channel(0, 0, rowL + 1, tf[a].off) : channel(0, 1, rowL, tf[a].off));
      b.load();
      for (int d = 0; d < 3; d++) {
        const CompRef v = pf_val(tf[a].kind, d);
        if (v.tc < 0) continue;
        b.addp(0, 0, v.tc, v.zd, F_SRC + 2 * d, v.sgn, 1.0, 0.0);
        b.addp(0, 0, v.tc, v.zd, F_SRC + 2 * d + 1, v.sgn, 0.0, -1.0);
      }
      b.finish();
    }
    // trace pairings P(k,j) = sum_faces int n^.(E^_j x F^_k) dS^
    S.crow.resize(2 * nEi);
    S.CW.assign((size_t)2 * nEi * D.np, 0.0);
    for (int j = 0; j < nEi; j++) { S.crow[2 * j] = rowI + 2 * j; S.crow[2 * j + 1] = rowI + 2 * j + 1; }
    std::vector<double> Ft((size_t)3 * nEE);
    for (int f = 0; f < 5; f++) {
      const FaceRule R = prism_face_rule(norder, norif, f, dp, cap);
      for (int l = 0; l < R.n; l++) {
        const double x = R.xi[3 * l], y = R.xi[3 * l + 1], z = R.xi[3 * l + 2], w = R.w[l];
        const TriVals vTV = eval_list(TV, x, y), vTS = eval_list(TS, x, y), vEt = eval_list(TEt, x, y), vHt = eval_list(THt, x, y);
        const ZVals Z = eval_z(std::max(ptab, 1), z);
        for (int k = 0; k < tf[0].nZ; k++)
          for (int t = 0; t < tf[0].nT; t++) { double *F = &Ft[3 * (size_t)(t + tf[0].nT * k)]; F[0] = vEt.at(t)[0] * Z.H[k]; F[1] = vEt.at(t)[1] * Z.H[k]; F[2] = 0.0; }
        for (int k = 0; k < tf[1].nZ; k++)
          for (int t = 0; t < tf[1].nT; t++) { double *F = &Ft[3 * (size_t)(tf[1].off + t + tf[1].nT * k)]; F[0] = F[1] = 0.0; F[2] = vHt.at(t)[0] * Z.Q[k]; }
        for (int j = 0; j < nEi; j++) {
          const PrismDof &dj = ed[j];
          if (dj.ent == 2 && dj.ent_idx != f) continue;   // face functions of the other faces are not generated (norder_ifc, elem_opt.F90:539-541)
          double E[3];
          prism_hcurl_value(dj, vTV, vTS, Z, E);
          // n.(E x F) = (n x E).F
          const double nx[3] = {R.nvec[1] * E[2] - R.nvec[2] * E[1], R.nvec[2] * E[0] - R.nvec[0] * E[2], R.nvec[0] * E[1] - R.nvec[1] * E[0]};
          if (nx[0] == 0.0 && nx[1] == 0.0 && nx[2] == 0.0) continue;
          double *rowH = &S.CW[(size_t)(2 * j + 1) * D.np], *rowE = &S.CW[(size_t)(2 * j) * D.np + nEE];
          for (int k = 0; k < nEE; k++) {
            const double v = w * (nx[0] * Ft[3 * k] + nx[1] * Ft[3 * k + 1] + nx[2] * Ft[3 * k + 2]);
            rowH[k] += rs ? rs_real(v, 1, 0, true) : v; rowE[k] += rs ? rs_real(v, 0, 1, true) : v;
          }
        }
      }
    }
  } else if (P.kind == 1) {
    // =============================================================== Poisson Galerkin
    const int iH = S.nH - bH;
    const PFam fu = add_pfam(S, PF_SCALAR, TG, pmax[1] + 1, T_H, nqt, tpts);
    S.cplx = false; S.dpg = false; S.ntest = 0; S.ni = iH; S.nb = bH;
    DenseDims &D = S.dims;
    D.cplx = false; D.dpg = false; D.n = 0; D.nb = S.nb; D.ni = S.ni; D.finish();
    if (sizes_only) return true;
    const int mapU = add_pgrid_map(S, hd, 0, fu, [&](int k) { return k < iH ? D.nbp + k : k - iH; });
    {
      BlockBuilder b(S, fu.id, fu.id, channel(1, 0, 0, 0, mapU, mapU), no_channel());
      add_grad_grad(b, 1.0);
      b.finish();
    }
    {
      BlockBuilder b(S, unit.id, fu.id, channel(1, 0, D.nbp + D.nil - 1, 0, -1, mapU), no_channel());
      b.addp(0, 0, 0, 0, F_SRC, 1.0, 1.0, 0.0);
      b.finish();
    }
  } else if (P.kind == 3) {
    // =============================================================== Maxwell Galerkin (complex symmetric, pivoted-LU condensation)
    TriList TV, TS;
    const std::vector<PrismDof> ed = prism_dofs_Hcurl(norder, norie, norif, TV, TS);
    const int nE = (int)ed.size(), iE = nE - bE;
    S.cplx = true; S.dpg = false; S.gen_stc = true; S.ntest = 0; S.ni = iE; S.nb = bE;
    DenseDims &D = S.dims;
    // lossless medium (sigma = 0): A is REAL symmetric and only the load is complex -> real LU on a real matrix with the load as
    // two real columns (Re, Im); the output kernel interleaves them (same idea as the ultraweak real form, without phases)
    const bool rsg = P.real_struct && P.sigma == 0.0;
    D.cplx = !rsg; D.rs = rsg; D.nload = rsg ? 2 : 1; D.dpg = false; D.n = 0; D.nb = S.nb; D.ni = S.ni; D.finish();
    if (sizes_only) return true;
    PFam fe[2];
    fe[0] = add_pfam(S, PF_EH, TV, pmax[1] + 1, T_H, nqt, tpts);
    fe[1] = add_pfam(S, PF_EV, TS, std::max(pmax[1], 1), T_Q, nqt, tpts);
    int mapE[2];
    for (int a = 0; a < 2; a++) mapE[a] = add_pgrid_map(S, ed, a, fe[a], [&](int k) { return k < iE ? D.nbp + k : k - iE; });
    const std::complex<double> zb(P.omega * P.omega * P.eps, -P.omega * P.sigma);
    for (int a = 0; a < 2; a++)
      for (int a2 = 0; a2 < 2; a2++) {
        if (fe[a].nT == 0 || fe[a2].nT == 0) continue;
        BlockBuilder b(S, fe[a].id, fe[a2].id, channel(1, 0, 0, 0, mapE[a], mapE[a2]), rsg ? no_channel() : channel(1, 1, 0, 0, mapE[a], mapE[a2]));
        add_hcurl_pair(b, fe[a].kind, fe[a2].kind, -zb.real(), rsg ? 0.0 : -zb.imag(), 1.0 / P.mu, 0.0);
        b.finish();
      }
    for (int a = 0; a < 2; a++) {
      if (fe[a].nT == 0) continue;
      BlockBuilder b(S, fe[a].id, unit.id, channel(1, 0, 0, D.nbp + D.nil - D.nload, mapE[a]), rsg ? channel(1, 0, 0, D.nbp + D.nil - 1, mapE[a]) : channel(1, 1, 0, D.nbp + D.nil - 1, mapE[a]));
      for (int d = 0; d < 3; d++) {
        const CompRef v = pf_val(fe[a].kind, d);
        if (v.tc < 0) continue;
        b.addp(v.tc, v.zd, 0, 0, F_SRC + 2 * d, v.sgn, 1.0, 0.0);
        b.addp(v.tc, v.zd, 0, 0, F_SRC + 2 * d + 1, v.sgn, 0.0, 1.0);
      }
      b.finish();
    }
  } else if (P.kind == 2) {
    // =============================================================== Poisson primal DPG
    TriList TZ, TH, THt;
    const std::vector<PrismDof> vd = prism_dofs_Hdiv_faces(norderi, norif, TZ, TH);
    const int nVi = (int)vd.size(), iH = S.nH - bH;
    tri_list_H1(pe, THt);
    const PFam ft = add_pfam(S, PF_SCALAR, THt, pze + 1, T_H, nqt, tpts);
    const int nHH = ft.nT * ft.nZ;
    const PFam fu = add_pfam(S, PF_SCALAR, TG, pmax[1] + 1, T_H, nqt, tpts);
    S.cplx = false; S.dpg = true; S.ntest = nHH; S.ni = iH + nVi; S.nb = bH;
    DenseDims &D = S.dims;
    D.cplx = false; D.dpg = true; D.nload = P.nrhs; D.n = nHH; D.nb = S.nb; D.ni = S.ni; D.finish();
    if (sizes_only) return true;
    const int rowB = D.np, rowI = D.np + D.nbp, rowL = rowI + D.nil - D.nload;   // load rows: the last padded interface rows, independent of ni
    const int mapU = add_pgrid_map(S, hd, 0, fu, [&](int k) { return k < iH ? rowI + k : rowB + (k - iH); });
    {
      BlockBuilder b(S, ft.id, ft.id, channel(0, 0, 0, 0), no_channel());
      add_grad_grad(b, 1.0);
      b.addp(0, 0, 0, 0, F_WDET, 1.0, 1.0, 0.0);
      b.finish();
    }
    {
      BlockBuilder b(S, fu.id, ft.id, channel(0, 0, 0, 0, mapU), no_channel());
      add_grad_grad(b, 1.0);
      b.finish();
    }
    {
      BlockBuilder b(S, unit.id, ft.id, channel(0, 0, rowL, 0), no_channel());
      b.load();
      b.addp(0, 0, 0, 0, F_SRC, 1.0, 1.0, 0.0);
      b.finish();
    }
    // -<sigma^.n, v>
    S.crow.resize(nVi);
    S.CW.assign((size_t)nVi * D.np, 0.0);
    for (int j = 0; j < nVi; j++) S.crow[j] = rowI + iH + j;
    std::vector<double> vt((size_t)nHH);
    for (int f = 0; f < 5; f++) {
      const FaceRule R = prism_face_rule(norder, norif, f, dp, cap);
      for (int l = 0; l < R.n; l++) {
        const double x = R.xi[3 * l], y = R.xi[3 * l + 1], z = R.xi[3 * l + 2], w = R.w[l];
        const TriVals vTZ = eval_list(TZ, x, y), vTH = eval_list(TH, x, y), vHt = eval_list(THt, x, y);
        const ZVals Z = eval_z(std::max(ptab, 1), z);
        for (int k = 0; k < ft.nZ; k++) for (int t = 0; t < ft.nT; t++) vt[t + (size_t)ft.nT * k] = vHt.at(t)[0] * Z.H[k];
        for (int j = 0; j < nVi; j++) {
          const PrismDof &dj = vd[j];
          if (dj.ent_idx != f) continue;
          double V[3];
          if (dj.list == 0) { V[0] = V[1] = 0.0; V[2] = vTZ.at(dj.t)[0] * Z.H[dj.zi] * dj.sgn; }
          else { const double *t = vTH.at(dj.t); const double q = Z.Q[dj.zi] * dj.sgn; V[0] = t[1] * q; V[1] = -t[0] * q; V[2] = 0.0; }
          const double vn = w * (V[0] * R.nvec[0] + V[1] * R.nvec[1] + V[2] * R.nvec[2]);
          if (vn == 0.0) continue;
          double *row = &S.CW[(size_t)j * D.np];
          for (int k = 0; k < nHH; k++) row[k] -= vn * vt[k];
        }
      }
    }
  } else {
    S.err = "problem kind not implemented for prisms";
    return false;
  }
  // ---- launch geometry of the prism kernel: smem = z tables (4*TABSZ) | Z [4][NMAX][NMAX] | G [nqt*nqz] | U [ns][NMAX][nTB]
  size_t u = 0;
  int items = 1, nm = 1;
  for (const BlockDesc &B : S.block) {
    const FamilyDesc &fa = S.fam[B.famA], &fb = S.fam[B.famB];
    u = std::max(u, (size_t)B.ns * fb.n[0]);   // x NMAX rows below
    items = std::max(items, std::max(fa.n[2], nqz) * fb.n[0]);
    nm = std::max(nm, std::max(std::max(fa.n[2], fb.n[2]), nqz));
    if (B.ns > 5) { S.err = "a block of the form has more than 5 z-slots (TP_SMAX)"; return false; }
  }
  S.nmax = nm <= 4 ? 4 : nm <= 6 ? 6 : nm <= 8 ? 8 : 10;
  S.threads = std::min(384, std::max(64, (items + 31) / 32 * 32));
  u *= S.nmax;
  S.smem_u_off = (size_t)4 * TABSZ + (size_t)4 * S.nmax * S.nmax + (((size_t)nqt * nqz + 1) & ~(size_t)1);
  S.smem_bytes = (S.smem_u_off + u) * sizeof(double);
  if (S.smem_bytes > 200 * 1024) { S.err = "integration kernel needs more than 200 KB of shared memory"; return false; }
  return true;
}

}  // namespace hp3d
