// forms.hpp -- host-side "form compiler": turns one element signature (problem kind + node orders + orientations)
// into the tables, tensor-product blocks, index maps and element-independent rows the GPU kernels execute.
//
// Weak forms follow the reference's element routines (paths relative to trunk/problems):
//   POISSON/GALERKIN/elem_opt.F90:110-137      A = (grad u, grad v),             b = (f, v)
//   POISSON/PRIMAL_DPG/elem_opt.F90:219-269,361-385   B = [(grad u, grad v) | -<sigma.n, v>], l = (f,v), G = (v,q)+(grad v,grad q)
//   MAXWELL/GALERKIN/elem_opt.F90:120-149      A = (1/mu curl E, curl F) - ((w^2 eps - i w sigma) E, F), b = -i w (J,F)
//   MAXWELL/ULTRAWEAK_DPG/elem/elem_opt.F90:338-470,643-649,719-768  (ultraweak Maxwell, adjoint-graph test norm)
// Trace terms: <n x E, F> and <sigma.n, v> are metric-free (the Piola maps cancel the surface Jacobian pointwise),
// so they are integrated once per signature here on the host with the SAME face rules the reference uses
// (src/element/quadrature/set_2D_int.F90:121-236) and shipped to the GPU as constant rows.
//
// Internal dof layout of the dense phase (dense_pipeline.cuh): trial = [bubble | interface | load]; inside each
// group the reference's own order is kept (stc gather order, src/modules/stc.F90:226-261), so the output maps
// are identities and orientation signs are applied here, at integration time.
#pragma once
#include <cstdlib>
#include "dense_pipeline.cuh"
#include "hexa_space.hpp"
#include "integ_kernels.cuh"
#include "tables.hpp"

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <string>
#include <vector>

namespace hp3d {

// mirror of hp3d_params (include/hp3d_gpu.h) restricted to what the forms need
struct FormParams {
  int kind = 0, nord_add = 1, maxp = 6, test_norm = 1;
  double alpha_norm = 1, omega = 1, eps = 1, mu = 1, sigma = 0;
  int source = 1, icomp = 0;
  // Ultraweak Maxwell with real eps, mu and no conductivity: with T = diag(i^ph) (ph = 0 for E-type trial dofs and the test
  // functions F, 1 for H-type dofs and G), the Gram matrix D^H G D is REAL and the enriched stiffness D^H B T is purely
  // imaginary, so A = T A~ T^H with A~ = R G~^-1 R^T real: the whole dense phase runs in real arithmetic (a quarter of the
  // flops of the reference's ZPOTRF/ZTRTRS/ZHERK) and the phases come back in the output kernels (formats.cuh).
  bool real_struct = true;
  // constant permittivity tensor of ultraweak Maxwell (get_permittivity: za = i w eps * eps_t, elem_opt.F90:260-266); entry (i,j) at [i + 3j]
  int nrhs = 1;   // NR_RHS: load vectors per element (DPG problems; the extra ones come from source tables)
  bool tensor = false;
  std::complex<double> epst[9];
  bool tensor_real() const { for (int i = 0; i < 9; i++) if (epst[i].imag() != 0.0) return false; return true; }
};

// element_data.F90:106-109 (0-based): edges of face f: [0],[2] run along the face's first axis, [1],[3] along the second
static const int FACE_EDGES[6][4] = {{0, 1, 2, 3}, {4, 5, 6, 7}, {0, 9, 4, 8}, {1, 10, 5, 9}, {2, 10, 6, 11}, {3, 11, 7, 8}};

struct SigHost {
  int kind = 0;
  int etype = 1;               // HP3D_MDLB / HP3D_MDLP
  std::vector<double> ttab;    // prism: pool of triangle-function tables, one block [3 comps][nT][nqt] per list
  int geo_toff = 0, geo_nT = 0; // prism: pool offset / size of the geometry (H1) list
  int norder[19], norie[12], norif[6];
  int nq[3] = {0, 0, 0}, nint = 0;
  int nH = 0;                 // geometry dofs
  int ni = 0, nb = 0, ntest = 0;
  bool cplx = false, dpg = false;
  bool gen_stc = false;        // condensation by pivoted LU (stc_fwd_gen) instead of Cholesky
  DenseDims dims;
  std::vector<double> tab, wq;
  std::vector<int> hdof;
  std::vector<FamilyDesc> fam;
  std::vector<TermDesc> term;
  std::vector<SlotDesc> slot;
  std::vector<BlockDesc> block;
  std::vector<WorkItem> work;
  std::vector<int> maps;
  std::vector<int> crow;       // constant rows of W
  std::vector<double> CW;      // [crow.size()][np]
  int nmax = 4, threads = 64;
  size_t smem_u_off = 0, smem_bytes = 0;
  size_t smem_f_off = 0, smem_t1_off = 0;   // hexahedron kernel: weight-field staging and x-contracted terms
  size_t smem_q_off = 0;                    // hexahedron kernel: y-factor products of the terms, double-buffered
  std::string err;
};

namespace detail {

inline ChannelDesc no_channel() { ChannelDesc c; c.mat = -1; c.plane = 0; c.row0 = c.col0 = 0; c.rowmap = c.colmap = -1; return c; }
inline ChannelDesc channel(int mat, int plane, int row0, int col0, int rowmap = -1, int colmap = -1) {
  ChannelDesc c; c.mat = mat; c.plane = plane; c.row0 = row0; c.col0 = col0; c.rowmap = rowmap; c.colmap = colmap; return c;
}

struct BlockBuilder {
  SigHost &S;
  BlockDesc B;
  BlockBuilder(SigHost &s, int famA, int famB, ChannelDesc c0, ChannelDesc c1) : S(s) {
    B.famA = famA; B.famB = famB; B.t0 = (int)S.term.size(); B.nt = 0; B.s0 = (int)S.slot.size(); B.ns = 0;
    B.ch[0] = c0; B.ch[1] = c1; B.is_load = 0;
  }
  BlockBuilder &load() { B.is_load = 1; return *this; }   // this block is a load vector (NR_RHS > 1 integrates it once per right-hand side)
  // add   coef * (c0, c1) * integral( dA-derivative of A  *  dB-derivative of B  *  field )
  void add(int dA, int dB, int field, double coef, double c0, double c1) {
    if (coef == 0.0 || (c0 == 0.0 && c1 == 0.0)) return;
    const FamilyDesc &fa = S.fam[B.famA], &fb = S.fam[B.famB];
    const int zA = dA == 2 ? (int)T_DH : fa.tab[2], zB = dB == 2 ? (int)T_DH : fb.tab[2];
    int s = -1;
    for (int i = 0; i < B.ns; i++) {
      const SlotDesc &sl = S.slot[B.s0 + i];
      if (sl.zA == zA && sl.zB == zB && sl.c[0] == c0 && sl.c[1] == c1) { s = i; break; }
    }
    if (s < 0) { SlotDesc sl; sl.zA = zA; sl.zB = zB; sl.c[0] = c0; sl.c[1] = c1; S.slot.push_back(sl); s = B.ns++; }
    TermDesc t; t.dA = dA; t.dB = dB; t.field = field; t.slot = s; t.coef = coef;
    S.term.push_back(t); B.nt++;
  }
  // prism variant: the A / B factor is (component tcA of the triangle table) x (z table, differentiated if zdA)
  void addp(int tcA, int zdA, int tcB, int zdB, int field, double coef, double c0, double c1) {
    if (coef == 0.0 || (c0 == 0.0 && c1 == 0.0)) return;
    const FamilyDesc &fa = S.fam[B.famA], &fb = S.fam[B.famB];
    const int zA = zdA ? (int)T_DH : fa.tab[2], zB = zdB ? (int)T_DH : fb.tab[2];
    int s = -1;
    for (int i = 0; i < B.ns; i++) {
      const SlotDesc &sl = S.slot[B.s0 + i];
      if (sl.zA == zA && sl.zB == zB && sl.c[0] == c0 && sl.c[1] == c1) { s = i; break; }
    }
    if (s < 0) { SlotDesc sl; sl.zA = zA; sl.zB = zB; sl.c[0] = c0; sl.c[1] = c1; S.slot.push_back(sl); s = B.ns++; }
    TermDesc t; t.dA = tcA; t.dB = tcB; t.field = field; t.slot = s; t.coef = coef;
    S.term.push_back(t); B.nt++;
  }
  void finish() {
    if (B.nt == 0) { S.slot.resize(B.s0); return; }
    // the kernels walk the terms slot by slot
    std::stable_sort(S.term.begin() + B.t0, S.term.begin() + B.t0 + B.nt, [](const TermDesc &a, const TermDesc &b) { return a.slot < b.slot; });
    const int bi = (int)S.block.size();
    S.block.push_back(B);
    const FamilyDesc &fa = S.fam[B.famA];
    // one work item (CTA) per first-axis index of the A family; the kernels loop over the second axis themselves
    for (int i = 0; i < fa.n[0]; i++) { WorkItem w; w.block = (short)bi; w.iA = (short)i; w.jA = 0; w.pad = 0; S.work.push_back(w); }
  }
};

// Real-structured storage of the dense phase's input W = [G ; B^H] (rows: test dofs of the Gram, then trial dofs; columns: test
// dofs): W~[r,c] = kappa * conj(i^pr) * i^pc * W[r,c] with kappa = 1 on Gram rows and i on trial rows; real by construction.
inline double rs_real(std::complex<double> w, int pr, int pc, bool trial_row) {
  const std::complex<double> I(0, 1);
  const std::complex<double> z = w * (pr ? -I : std::complex<double>(1, 0)) * (pc ? I : std::complex<double>(1, 0)) * (trial_row ? I : std::complex<double>(1, 0));
  return z.real();
}
inline bool rs_applicable(const FormParams &P) { return P.kind == 4 && P.real_struct && P.sigma == 0.0 && (!P.tensor || P.tensor_real()); }

inline int add_family(SigHost &S, int n0, int n1, int n2, int t0, int t1, int t2) {
  FamilyDesc f; f.n[0] = n0; f.n[1] = n1; f.n[2] = n2; f.tab[0] = t0; f.tab[1] = t1; f.tab[2] = t2;
  S.fam.push_back(f);
  return (int)S.fam.size() - 1;
}
// curl of a reference H(curl) function of family a:  component (a+1)%3 = +d/d(a+2),  component (a+2)%3 = -d/d(a+1)
struct CurlComp { int comp, dax, sgn; };
inline void curl_comps(int a, CurlComp cc[2]) {
  cc[0].comp = (a + 1) % 3; cc[0].dax = (a + 2) % 3; cc[0].sgn = 1;
  cc[1].comp = (a + 2) % 3; cc[1].dax = (a + 1) % 3; cc[1].sgn = -1;
}
// map over the canonical grid (n0 x n1 x n2, first index fastest) of a conforming space:
// entry = +-(target+1) for grid functions that exist in the space, 0 otherwise
template <class TargetFn>
inline int add_grid_map(SigHost &S, const std::vector<TensorDof> &dofs, int fam, const int n[3], TargetFn target) {
  const int off = (int)S.maps.size();
  S.maps.resize(off + (size_t)n[0] * n[1] * n[2], 0);
  for (size_t k = 0; k < dofs.size(); k++) {
    const TensorDof &d = dofs[k];
    if (d.fam != fam) continue;
    const int g = d.idx[0] + n[0] * (d.idx[1] + n[1] * d.idx[2]);
    S.maps[off + g] = d.sgn * (target((int)k) + 1);
  }
  return off;
}

// 1-D pairing sum_q w_q TA[iA](x_q) TB[iB](x_q) with an nq-point rule; tables of order 9 evaluated on the fly
struct Pairing1D {
  Tables1D t[MAXN1D + 1];
  Pairing1D() { for (int n = 1; n <= MAXN1D; n++) t[n] = make_tables(MAXN1D - 1, n); }
  double val(int nq, int kindA, int iA, int kindB, int iB) const {
    const Tables1D &T = t[nq];
    const double *A = (kindA == KH ? T.H.data() : T.Q.data()) + (size_t)iA * nq;
    const double *B = (kindB == KH ? T.H.data() : T.Q.data()) + (size_t)iB * nq;
    double s = 0;
    for (int q = 0; q < nq; q++) s += T.w[q] * A[q] * B[q];
    return s;
  }
};
inline double end_value(int kind, int i, int side) {  // 1-D basis function at x = 0 / 1
  if (kind == KH) return i == side ? 1.0 : 0.0;
  return side ? 1.0 : (double)parity_sign(i);
}
// points per face axis, set_2D_int[_DPG] for a quad (set_2D_int.F90:121-150,218-236)
inline void face_rule(const int norder[19], const int norif[6], int f, int integration, int cap, int nqf[2]) {
  const HexaOrders o = HexaOrders::decode(norder);
  int h = o.face[f][0], v = o.face[f][1];
  if (QSWAP_ORDER[norif[f]]) std::swap(h, v);
  int nx = std::max(std::max(o.edge[FACE_EDGES[f][0]], o.edge[FACE_EDGES[f][2]]), h);
  int ny = std::max(std::max(o.edge[FACE_EDGES[f][1]], o.edge[FACE_EDGES[f][3]]), v);
  nqf[0] = std::min(nx + integration, cap) + 1;
  nqf[1] = std::min(ny + integration, cap) + 1;
}

}  // namespace detail

// Build everything for one signature.  Returns false (with S.err) for unsupported input.
// sizes_only: stop once the dof counts, quadrature size and padded dense extents are known (no blocks, no trace pairings)
inline bool compile_signature(const FormParams &P, const int norder[19], const int norie[12], const int norif[6], SigHost &S, bool sizes_only = false) {
  using namespace detail;
  S = SigHost();
  S.kind = P.kind;
  memcpy(S.norder, norder, sizeof S.norder); memcpy(S.norie, norie, sizeof S.norie); memcpy(S.norif, norif, sizeof S.norif);
  const HexaOrders o = HexaOrders::decode(norder);
  for (int e = 0; e < 12; e++) if (o.edge[e] < 1 || o.edge[e] > 9 || (norie[e] != 0 && norie[e] != 1)) { S.err = "bad edge order/orientation"; return false; }
  for (int f = 0; f < 6; f++) if (o.face[f][0] < 1 || o.face[f][1] < 1 || norif[f] < 0 || norif[f] > 7) { S.err = "bad face order/orientation"; return false; }
  for (int d = 0; d < 3; d++) if (o.mid[d] < 1) { S.err = "bad middle node order"; return false; }
  const bool dpg = (P.kind == 2 || P.kind == 4);
  const int dp = dpg ? P.nord_add : 0;
  int pmax[3], pe[3], ptab[3];
  hexa_axis_max_order(norder, norif, pmax);
  const int cap = dpg ? P.maxp + 1 : P.maxp;   // MAXPP for set_3D_int_DPG, MAXP otherwise (set_3D_int.F90:47-62,112-127)
  for (int d = 0; d < 3; d++) {
    pe[d] = o.mid[d] + dp;
    S.nq[d] = std::min(pmax[d] + dp, cap) + 1;
    ptab[d] = std::max(pmax[d], pe[d]);
    if (S.nq[d] > MAXQ || ptab[d] > MAXQ - 1) { S.err = "order exceeds the 10-point Gauss table limit"; return false; }
  }
  S.nint = S.nq[0] * S.nq[1] * S.nq[2];
  // ---- 1-D tables
  S.tab.assign((size_t)12 * TABSZ, 0.0);
  S.wq.assign((size_t)3 * MAXQ, 0.0);
  for (int d = 0; d < 3; d++) {
    Tables1D t = make_tables(ptab[d], S.nq[d]);
    for (int q = 0; q < S.nq[d]; q++) S.wq[d * MAXQ + q] = t.w[q];
    std::copy(t.H.begin(), t.H.end(), S.tab.begin() + (d * 4 + T_H) * TABSZ);
    std::copy(t.dH.begin(), t.dH.end(), S.tab.begin() + (d * 4 + T_DH) * TABSZ);
    std::copy(t.Q.begin(), t.Q.end(), S.tab.begin() + (d * 4 + T_Q) * TABSZ);
    for (int q = 0; q < S.nq[d]; q++) S.tab[(d * 4 + T_ONE) * TABSZ + q] = 1.0;
  }
  // ---- geometry dofs
  const std::vector<TensorDof> hd = hexa_dofs_H1(norder, norie, norif);
  S.nH = (int)hd.size();
  for (const TensorDof &d : hd) S.hdof.push_back(d.idx[0] | (d.idx[1] << 8) | (d.idx[2] << 16) | ((d.sgn < 0) << 24));
  int bH, bE, bV, bQ;
  hexa_mid_counts(o.mid, bH, bE, bV, bQ);
  int norderi[19];
  memcpy(norderi, norder, sizeof norderi);
  norderi[18] = 111;   // trace variables: middle-node order forced to 1 (elem_opt.F90:187)
  static const Pairing1D pair1d;
  const int unit = add_family(S, 1, 1, 1, T_ONE, T_ONE, T_ONE);
  const std::complex<double> I(0, 1);

  if (P.kind == 4) {
    // =============================================================== ultraweak Maxwell (complex, DPG)
    if (P.test_norm < 1 || P.test_norm > 3) { S.err = "unknown test norm"; return false; }
    const std::vector<TensorDof> ed = hexa_dofs_Hcurl(norderi, norie, norif);
    const int nEi = (int)ed.size(), nQ = bQ;
    int tf[3], offE[3], nEE = 0;
    for (int a = 0; a < 3; a++) {
      int n[3], t[3];
      for (int d = 0; d < 3; d++) { n[d] = d == a ? pe[d] : pe[d] + 1; t[d] = d == a ? T_Q : T_H; }
      tf[a] = add_family(S, n[0], n[1], n[2], t[0], t[1], t[2]);
      offE[a] = nEE; nEE += n[0] * n[1] * n[2];
    }
    const int fq = add_family(S, o.mid[0], o.mid[1], o.mid[2], T_Q, T_Q, T_Q);
    S.cplx = true; S.dpg = true; S.ntest = 2 * nEE; S.ni = 2 * nEi; S.nb = 6 * nQ;
    DenseDims &D = S.dims;
    const bool rs = rs_applicable(P);
    D.cplx = !rs; D.rs = rs; D.nload = (rs ? 2 : 1) * P.nrhs; D.dpg = true; D.n = S.ntest; D.nb = S.nb; D.ni = S.ni; D.finish();
    if (sizes_only) return true;
    const int rowB = D.np, rowI = D.np + D.nbp, rowL = rowI + D.nil - D.nload;   // load row(s): last padded interface rows, independent of ni
    const std::complex<double> za = I * P.omega * P.eps, zc = I * P.omega * P.mu;
    const double aF = (P.test_norm == 2) ? 1.0 : P.alpha_norm + std::norm(za);
    const double aG = (P.test_norm == 2) ? 1.0 : P.alpha_norm + std::norm(zc);
    // Gram diagonal blocks FF (channel 0) and GG (channel 1), lower block-triangle
    for (int a = 0; a < 3; a++)
      for (int a2 = 0; a2 <= a; a2++) {
        BlockBuilder b(S, tf[a], tf[a2], channel(0, 0, offE[a], offE[a2]), channel(0, 0, nEE + offE[a], nEE + offE[a2]));
        if (P.tensor && P.test_norm != 2) {
          // (za^H F_c, za^H F_r) = w^2 eps^2 F_r^T (eps_t eps_t^H) F_c: the real part R here, the imaginary part S in a block of its own below
          // FF: alpha D + w^2 eps^2 TD, GG: aG D -- written as aG D (1,1) + [(alpha - aG) D + w^2 eps^2 TD] (1,0) so that the mass terms
          // share the z-slots (tab,tab,1,1) of the curl-curl terms and (tab,tab,1,0): five slots per block (TP_SMAX)
          b.add(-1, -1, F_D + sym_idx(a, a2), aG, 1.0, 1.0);
          b.add(-1, -1, F_D + sym_idx(a, a2), P.alpha_norm - aG, 1.0, 0.0);
          b.add(-1, -1, F_TD + sym_idx(a, a2), P.omega * P.omega * P.eps * P.eps, 1.0, 0.0);
        } else
          b.add(-1, -1, F_D + sym_idx(a, a2), 1.0, aF, aG);
        CurlComp ca[2], cb[2];
        curl_comps(a, ca); curl_comps(a2, cb);
        for (int i = 0; i < 2; i++)
          for (int j = 0; j < 2; j++) b.add(ca[i].dax, cb[j].dax, F_C + sym_idx(ca[i].comp, cb[j].comp), ca[i].sgn * cb[j].sgn, 1.0, 1.0);
        b.finish();
      }
    // imaginary part of the FF Gram blocks for a complex permittivity tensor: F_r^T (w^2 eps^2 S) F_c, S antisymmetric (zero on a == a2)
    if (P.tensor && P.test_norm != 2 && !P.tensor_real())
      for (int a = 1; a < 3; a++)
        for (int a2 = 0; a2 < a; a2++) {
          BlockBuilder b(S, tf[a], tf[a2], channel(0, 1, offE[a], offE[a2]), no_channel());
          b.add(-1, -1, F_TS + (a == 1 ? 0 : a2 + 1), P.omega * P.omega * P.eps * P.eps, 1.0, 0.0);
          b.finish();
        }
    // Gram cross block: W[G_j][F_i] = conj(gFG(i,j)) = -conj(za) (F_i, curl G_j) + zc (curl F_i, G_j)   (metric-free: weight w)
    // permittivity tensor: -(curl G_j)^T za^H F_i = i w eps (curl^ G_j)^T [w (J^T er^T J^-T) - i w (J^T ei^T J^-T)] F^_i: every curl
    // component b of the row family meets the column family through the fields T1R / T1I (b, a2), also for a == a2
    if (P.test_norm == 1)
      for (int a = 0; a < 3; a++)      // G row family
        for (int a2 = 0; a2 < 3; a2++) {  // F column family
          if (a == a2 && !P.tensor) continue;
          BlockBuilder b(S, tf[a], tf[a2], channel(0, 0, nEE + offE[a], offE[a2]), rs ? no_channel() : channel(0, 1, nEE + offE[a], offE[a2]));
          CurlComp cg[2], cf[2];
          curl_comps(a, cg); curl_comps(a2, cf);
          const std::complex<double> m1 = -std::conj(za), m2 = zc;
          for (int i = 0; i < 2; i++) {
            if (P.tensor) {
              b.add(cg[i].dax, -1, F_T1R + 3 * cg[i].comp + a2, cg[i].sgn, rs ? rs_real(m1, 1, 0, false) : m1.real(), rs ? 0.0 : m1.imag());
              const std::complex<double> m1i = m1 * std::complex<double>(0.0, -1.0);
              if (!P.tensor_real()) b.add(cg[i].dax, -1, F_T1I + 3 * cg[i].comp + a2, cg[i].sgn, m1i.real(), m1i.imag());
            } else if (cg[i].comp == a2) b.add(cg[i].dax, -1, F_W, cg[i].sgn, rs ? rs_real(m1, 1, 0, false) : m1.real(), rs ? 0.0 : m1.imag());
            if (cf[i].comp == a) b.add(-1, cf[i].dax, F_W, cf[i].sgn, rs ? rs_real(m2, 1, 0, false) : m2.real(), rs ? 0.0 : m2.imag());
          }
          b.finish();
        }
    // enriched stiffness, field columns (rows of W = trial dofs 6*j+comp, reference bubble order elem_opt.F90:738-760)
    int mapE[3], mapH[3];
    for (int c = 0; c < 3; c++) {
      mapE[c] = (int)S.maps.size();
      for (int j = 0; j < nQ; j++) S.maps.push_back(rowB + 6 * j + c + 1);
      mapH[c] = (int)S.maps.size();
      for (int j = 0; j < nQ; j++) S.maps.push_back(rowB + 6 * j + 3 + c + 1);
    }
    for (int c = 0; c < 3; c++)
      for (int a = 0; a < 3; a++) {
        {  // B(F_i, E_jc) = -za (E_c, F_i)  ->  W = conj
          BlockBuilder b(S, fq, tf[a], channel(0, 0, 0, offE[a], mapE[c]), rs ? no_channel() : channel(0, 1, 0, offE[a], mapE[c]));
          for (int d = 0; d < 3; d++) {   // -(za E_c e_c, F) = -sum_d za(d,c) F_d: the identity tensor keeps d = c only
            const std::complex<double> m = P.tensor ? -std::conj(za * P.epst[d + 3 * c]) : (d == c ? -std::conj(za) : std::complex<double>(0.0, 0.0));
            if (m == std::complex<double>(0.0, 0.0)) continue;
            b.add(-1, -1, F_WJI + 3 * a + d, 1.0, rs ? rs_real(m, 0, 0, true) : m.real(), rs ? 0.0 : m.imag());
          }
          b.finish();
        }
        {  // B(F_i, H_jc) = B(G_i, E_jc) = (H_c, curl F_i)   real
          BlockBuilder b(S, fq, tf[a], channel(0, 0, 0, offE[a], mapH[c]), channel(0, 0, 0, nEE + offE[a], mapE[c]));
          CurlComp cf[2];
          curl_comps(a, cf);
          for (int i = 0; i < 2; i++) b.add(-1, cf[i].dax, F_WJD + 3 * c + cf[i].comp, cf[i].sgn, rs ? rs_real(1.0, 1, 0, true) : 1.0, rs ? rs_real(1.0, 0, 1, true) : 1.0);
          b.finish();
        }
        {  // B(G_i, H_jc) = zc (H_c, G_i)  ->  W = conj
          const std::complex<double> m = std::conj(zc);
          BlockBuilder b(S, fq, tf[a], channel(0, 0, 0, nEE + offE[a], mapH[c]), rs ? no_channel() : channel(0, 1, 0, nEE + offE[a], mapH[c]));
          b.add(-1, -1, F_WJI + 3 * a + c, 1.0, rs ? rs_real(m, 1, 1, true) : m.real(), rs ? 0.0 : m.imag());
          b.finish();
        }
      }
    // load: l(F_i) = (J, F_i) ; W[load][F_i] = conj
    for (int a = 0; a < 3; a++) {
      // real-structured: Re and Im of the stored load row (= conj(l)) go to the two load rows of the single plane
      BlockBuilder b(S, unit, tf[a], channel(0, 0, rowL, offE[a]), rs ? channel(0, 0, rowL + 1, offE[a]) : channel(0, 1, rowL, offE[a]));
      b.load();
      b.add(-1, -1, F_SRC + 2 * a, 1.0, 1.0, 0.0);
      b.add(-1, -1, F_SRC + 2 * a + 1, 1.0, 0.0, -1.0);
      b.finish();
    }
    // trace pairings  P(k,j) = sum_faces int n.(E_j x F_k):  B(F_k, H^_j) = B(G_k, E^_j) = P(k,j)
    S.crow.resize(2 * nEi);
    S.CW.assign((size_t)2 * nEi * D.np, 0.0);
    for (int j = 0; j < nEi; j++) { S.crow[2 * j] = rowI + 2 * j; S.crow[2 * j + 1] = rowI + 2 * j + 1; }
    for (int f = 0; f < 6; f++) {
      const int c = FAX[f][0], side = FAX[f][1], a = (c + 1) % 3, b = (c + 2) % 3;
      const double nsign = side ? 1.0 : -1.0;
      int nqf2[2], nqa, nqb;
      face_rule(norder, norif, f, dp, cap, nqf2);
      // face axes (FAX[f][2], FAX[f][3]) are ascending, and so are (a,b) up to order: match them
      nqa = (FAX[f][2] == a) ? nqf2[0] : nqf2[1];
      nqb = (FAX[f][2] == b) ? nqf2[0] : nqf2[1];
      for (int j = 0; j < nEi; j++) {
        const TensorDof &dj = ed[j];
        if (dj.fam == c) continue;
        const double vj = end_value(kind_of(SP_HCURL, dj.fam, c), dj.idx[c], side);
        if (vj == 0.0) continue;
        const int fk = (dj.fam == a) ? b : a;                 // the only test family that pairs with it
        const double s = nsign * ((dj.fam == a) ? 1.0 : -1.0) * dj.sgn * vj;
        const FamilyDesc &F = S.fam[tf[fk]];
        for (int k2 = 0; k2 < F.n[2]; k2++)
          for (int k1 = 0; k1 < F.n[1]; k1++)
            for (int k0 = 0; k0 < F.n[0]; k0++) {
              const int ik[3] = {k0, k1, k2};
              const double vk = end_value(kind_of(SP_HCURL, fk, c), ik[c], side);
              if (vk == 0.0) continue;
              const double Ia = pair1d.val(nqa, kind_of(SP_HCURL, dj.fam, a), dj.idx[a], kind_of(SP_HCURL, fk, a), ik[a]);
              const double Ib = pair1d.val(nqb, kind_of(SP_HCURL, dj.fam, b), dj.idx[b], kind_of(SP_HCURL, fk, b), ik[b]);
              const int k = offE[fk] + k0 + F.n[0] * (k1 + F.n[1] * k2);
              const double v = s * vk * Ia * Ib;
              S.CW[(size_t)(2 * j + 1) * D.np + k] += rs ? rs_real(v, 1, 0, true) : v;        // row H^_j, column F_k
              S.CW[(size_t)(2 * j) * D.np + nEE + k] += rs ? rs_real(v, 0, 1, true) : v;      // row E^_j, column G_k
            }
      }
    }
  } else if (P.kind == 2) {
    // =============================================================== Poisson primal DPG (real)
    const std::vector<TensorDof> vd = hexa_dofs_Hdiv(norderi, norif);
    const int nVi = (int)vd.size(), iH = S.nH - bH;
    const int ft = add_family(S, pe[0] + 1, pe[1] + 1, pe[2] + 1, T_H, T_H, T_H);
    const int nHH = (pe[0] + 1) * (pe[1] + 1) * (pe[2] + 1);
    const int ng[3] = {pmax[0] + 1, pmax[1] + 1, pmax[2] + 1};
    const int fu = add_family(S, ng[0], ng[1], ng[2], T_H, T_H, T_H);
    S.cplx = false; S.dpg = true; S.ntest = nHH; S.ni = iH + nVi; S.nb = bH;
    DenseDims &D = S.dims;
    D.cplx = false; D.dpg = true; D.nload = P.nrhs; D.n = nHH; D.nb = S.nb; D.ni = S.ni; D.finish();
    if (sizes_only) return true;
    const int rowB = D.np, rowI = D.np + D.nbp, rowL = rowI + D.nil - D.nload;   // load rows: the last padded interface rows, independent of ni
    const int mapU = add_grid_map(S, hd, -1, ng, [&](int k) { return k < iH ? rowI + k : rowB + (k - iH); });
    {  // Gram (v,q) + (grad v, grad q)
      BlockBuilder b(S, ft, ft, channel(0, 0, 0, 0), no_channel());
      for (int d = 0; d < 3; d++) for (int d2 = 0; d2 < 3; d2++) b.add(d, d2, F_D + sym_idx(d, d2), 1.0, 1.0, 0.0);
      b.add(-1, -1, F_WDET, 1.0, 1.0, 0.0);
      b.finish();
    }
    {  // (grad u, grad v)
      BlockBuilder b(S, fu, ft, channel(0, 0, 0, 0, mapU), no_channel());
      for (int d = 0; d < 3; d++) for (int d2 = 0; d2 < 3; d2++) b.add(d, d2, F_D + sym_idx(d, d2), 1.0, 1.0, 0.0);
      b.finish();
    }
    {  // load
      BlockBuilder b(S, unit, ft, channel(0, 0, rowL, 0), no_channel());
      b.load();
      b.add(-1, -1, F_SRC, 1.0, 1.0, 0.0);
      b.finish();
    }
    // -<sigma.n, v>
    S.crow.resize(nVi);
    S.CW.assign((size_t)nVi * D.np, 0.0);
    for (int j = 0; j < nVi; j++) S.crow[j] = rowI + iH + j;
    for (int f = 0; f < 6; f++) {
      const int c = FAX[f][0], side = FAX[f][1], a = FAX[f][2], b = FAX[f][3];
      const double nsign = side ? 1.0 : -1.0;
      int nqf2[2];
      face_rule(norder, norif, f, dp, cap, nqf2);
      for (int j = 0; j < nVi; j++) {
        const TensorDof &dj = vd[j];
        if (dj.fam != c) continue;
        const double vj = end_value(KH, dj.idx[c], side);
        if (vj == 0.0) continue;
        const FamilyDesc &F = S.fam[ft];
        for (int k2 = 0; k2 < F.n[2]; k2++)
          for (int k1 = 0; k1 < F.n[1]; k1++)
            for (int k0 = 0; k0 < F.n[0]; k0++) {
              const int ik[3] = {k0, k1, k2};
              const double vk = end_value(KH, ik[c], side);
              if (vk == 0.0) continue;
              const double Ia = pair1d.val(nqf2[0], KQ, dj.idx[a], KH, ik[a]);
              const double Ib = pair1d.val(nqf2[1], KQ, dj.idx[b], KH, ik[b]);
              S.CW[(size_t)j * D.np + (k0 + F.n[0] * (k1 + F.n[1] * k2))] -= nsign * dj.sgn * vj * vk * Ia * Ib;
            }
      }
    }
  } else if (P.kind == 1) {
    // =============================================================== Poisson Galerkin (real)
    const int iH = S.nH - bH;
    const int ng[3] = {pmax[0] + 1, pmax[1] + 1, pmax[2] + 1};
    const int fu = add_family(S, ng[0], ng[1], ng[2], T_H, T_H, T_H);
    S.cplx = false; S.dpg = false; S.ntest = 0; S.ni = iH; S.nb = bH;
    DenseDims &D = S.dims;
    D.cplx = false; D.dpg = false; D.n = 0; D.nb = S.nb; D.ni = S.ni; D.finish();
    if (sizes_only) return true;
    const int mapU = add_grid_map(S, hd, -1, ng, [&](int k) { return k < iH ? D.nbp + k : k - iH; });
    {
      BlockBuilder b(S, fu, fu, channel(1, 0, 0, 0, mapU, mapU), no_channel());
      for (int d = 0; d < 3; d++) for (int d2 = 0; d2 < 3; d2++) b.add(d, d2, F_D + sym_idx(d, d2), 1.0, 1.0, 0.0);
      b.finish();
    }
    {
      BlockBuilder b(S, unit, fu, channel(1, 0, D.nbp + D.nil - 1, 0, -1, mapU), no_channel());
      b.add(-1, -1, F_SRC, 1.0, 1.0, 0.0);
      b.finish();
    }
  } else if (P.kind == 3) {
    // =============================================================== Maxwell Galerkin (complex symmetric, bilinear)
    // A = (1/mu curl E, curl F) - ((w^2 eps - i w sigma) E, F),  b = -i w (J, F): no conjugation (ZSYRK in the reference),
    // bubble block indefinite -> pivoted-LU condensation (HERM_STC = .false.)
    const std::vector<TensorDof> ed = hexa_dofs_Hcurl(norder, norie, norif);
    const int nE = (int)ed.size(), iE = nE - bE;
    S.cplx = true; S.dpg = false; S.gen_stc = true; S.ntest = 0; S.ni = iE; S.nb = bE;
    DenseDims &D = S.dims;
    // lossless medium (sigma = 0): A is REAL symmetric and only the load is complex -> real LU on a real matrix with the load as
    // two real columns (Re, Im); the output kernel interleaves them (same idea as the ultraweak real form, without phases)
    const bool rsg = P.real_struct && P.sigma == 0.0;
    D.cplx = !rsg; D.rs = rsg; D.nload = rsg ? 2 : 1; D.dpg = false; D.n = 0; D.nb = S.nb; D.ni = S.ni; D.finish();
    if (sizes_only) return true;
    int fe[3], mapE[3];
    for (int a = 0; a < 3; a++) {
      int n[3], t[3];
      for (int d = 0; d < 3; d++) { n[d] = d == a ? pmax[d] : pmax[d] + 1; t[d] = d == a ? T_Q : T_H; }
      fe[a] = add_family(S, n[0], n[1], n[2], t[0], t[1], t[2]);
      mapE[a] = add_grid_map(S, ed, a, n, [&](int k) { return k < iE ? D.nbp + k : k - iE; });
    }
    const std::complex<double> zb(P.omega * P.omega * P.eps, -P.omega * P.sigma);
    for (int a = 0; a < 3; a++)
      for (int a2 = 0; a2 < 3; a2++) {
        BlockBuilder b(S, fe[a], fe[a2], channel(1, 0, 0, 0, mapE[a], mapE[a2]), rsg ? no_channel() : channel(1, 1, 0, 0, mapE[a], mapE[a2]));
        b.add(-1, -1, F_D + sym_idx(a, a2), 1.0, -zb.real(), rsg ? 0.0 : -zb.imag());
        CurlComp ca[2], cb[2];
        curl_comps(a, ca); curl_comps(a2, cb);
        for (int i = 0; i < 2; i++)
          for (int j = 0; j < 2; j++) b.add(ca[i].dax, cb[j].dax, F_C + sym_idx(ca[i].comp, cb[j].comp), ca[i].sgn * cb[j].sgn, 1.0 / P.mu, 0.0);
        b.finish();
      }
    for (int a = 0; a < 3; a++) {  // load as a COLUMN (the LU condensation eliminates rows)
      BlockBuilder b(S, fe[a], unit, channel(1, 0, 0, D.nbp + D.nil - D.nload, mapE[a]), rsg ? channel(1, 0, 0, D.nbp + D.nil - 1, mapE[a]) : channel(1, 1, 0, D.nbp + D.nil - 1, mapE[a]));
      b.add(-1, -1, F_SRC + 2 * a, 1.0, 1.0, 0.0);
      b.add(-1, -1, F_SRC + 2 * a + 1, 1.0, 0.0, 1.0);
      b.finish();
    }
  } else {
    S.err = "problem kind not implemented on the GPU yet";
    return false;
  }
  // ---- launch geometry of the tp3 kernel
  int items = 1, nm = 1, ntmax = 1, nMmax = 1;
  for (const BlockDesc &B : S.block) {
    const FamilyDesc &fa = S.fam[B.famA], &fb = S.fam[B.famB];
    items = std::max(items, S.nq[2] * fb.n[0] * fb.n[1]);   // y contraction: one thread per (qz, iB, jB)
    nm = std::max(nm, std::max(std::max(fa.n[2], fb.n[2]), std::max(S.nq[1], S.nq[2])));
    ntmax = std::max(ntmax, B.nt);
    if (B.ns > 5) { S.err = "a block of the form has more than 5 z-slots (TP_SMAX)"; return false; }
  }
  S.nmax = nm <= 4 ? 4 : nm <= 6 ? 6 : nm <= 8 ? 8 : 10;
  size_t t1 = 0, u = 0, q = 0;
  for (const BlockDesc &B : S.block) {
    const FamilyDesc &fa = S.fam[B.famA], &fb = S.fam[B.famB];
    t1 = std::max(t1, (size_t)B.nt * fb.n[0] * S.nq[2] * (S.nmax + 2));   // T1 [t][qz][iB][qy padded to NMAX+2]
    u = std::max(u, (size_t)B.ns * fb.n[0] * fb.n[1] * S.nmax);          // U [ns][NMAX][nij]
    q = std::max(q, (size_t)B.nt * fb.n[1] * (S.nmax + 2));              // Q [t][jB][qy padded]
    nMmax = std::max(nMmax, fa.n[2] * ((S.nmax + 7) / 8));               // m-tiles of the z contraction
  }
  // CTA size: one thread per y-contraction item, rounded up to whole warps and then to a multiple of the m-tile count of the
  // z contraction (every warp owns one m-tile there), at most 448
  int warps = (std::max(64, items) + 31) / 32;
  if (nMmax <= 14) warps = std::min(14 / nMmax * nMmax, (warps + nMmax - 1) / nMmax * nMmax);
  if (const char *w = getenv("HP3D_TP3_WARPS")) warps = atoi(w);   // experiment hook
  S.threads = 32 * std::max(2, std::min(14, warps));
  // dynamic smem of tp3_kernel: tables | Z | F [ntmax][fs] | T1 | U | Q [2][...]  (all offsets even: 16-byte aligned for the TMA bulk copies)
  S.smem_f_off = (size_t)12 * TABSZ + (size_t)4 * S.nmax * S.nmax;   // tables | z tables re-laid out [4][NMAX][NMAX]
  S.smem_t1_off = S.smem_f_off + (size_t)ntmax * wf_stride(S.nint);
  S.smem_u_off = S.smem_t1_off + ((t1 + 1) & ~(size_t)1);
  S.smem_q_off = S.smem_u_off + ((u + 1) & ~(size_t)1);
  S.smem_bytes = (S.smem_q_off + 2 * q) * sizeof(double);
  if (S.smem_bytes > 200 * 1024) { S.err = "integration kernel needs more than 200 KB of shared memory"; return false; }
  return true;
}

}  // namespace hp3d
