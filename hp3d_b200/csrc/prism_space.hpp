// prism_space.hpp -- host-side description of hp3D's triangular-prism shape functions as SIGNED PRODUCTS
//        sign * T(x,y) * Z(z)   [times a direction for the vector spaces]
// of a triangle function T (tri_space.hpp) and a 1-D factor Z from the H / Q tables (tables.hpp), enumerated in
// the reference's dof order (src/element/shape_1/Prism.F90:38 H1, :358 H(curl), :760 H(div), :1040 L2; blending and
// projection pairs BlendProject.F90:562-800; orientations Orient.F90:9,39,119; topology element_data.F90:25-29,
// 62-65,85-88,108-111).  The enriched (broken) test spaces are full T x Z grids (broken/BrokenPrism.F90:31-290).
//
// Vector structure of the two H(curl) "directions" of a prism:
//   horizontal  E = (T_1 Z, T_2 Z, 0)        curl E = (-T_2 Z', T_1 Z', (curl T) Z)         T: triangle H(curl) function
//   vertical    E = (0, 0, T Z)              curl E = (dT/dy Z, -dT/dx Z, 0)                T: triangle H1 function, Z from Q
#pragma once
#include "hexa_space.hpp"
#include "tables.hpp"
#include "tri_space.hpp"

namespace hp3d {

static const int PR_TRI_EDGE[3][2] = {{0, 1}, {1, 2}, {0, 2}};                      // ProjectTriE / BlendProjectPrisME
static const int PR_OT[6][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}, {0, 2, 1}, {1, 0, 2}, {2, 1, 0}};  // OrientTri
static const int PR_OT_PARITY[6] = {1, 1, 1, -1, -1, -1};
// master prism: vertices (element_data.F90:25-29), faces -> vertices (1-based, :85-88)
static const double PR_COORD[6][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {0, 1, 1}};
static const int PR_FACE_VERT[5][4] = {{1, 2, 3, 1}, {4, 5, 6, 4}, {1, 2, 5, 4}, {2, 3, 6, 5}, {1, 3, 6, 4}};
static const int PR_FACE_EDGE[5][4] = {{1, 2, 3, 1}, {4, 5, 6, 4}, {1, 8, 4, 7}, {2, 9, 5, 8}, {3, 9, 6, 7}};
static const int PR_NSIGN[5] = {-1, 1, 1, 1, -1};                                    // Nsign_param, element_data.F90:616-626

struct PrismOrders {
  int edge[9];      // 6 mixed (triangle) edges, 3 vertical edges
  int tface[2];     // triangle faces
  int qface[3][2];  // quad faces, digits in the face's own (oriented) frame
  int mid[2];       // (p_xy, p_z)
  static PrismOrders decode(const int norder[15]) {
    PrismOrders o;
    for (int e = 0; e < 9; e++) o.edge[e] = norder[e];
    o.tface[0] = norder[9]; o.tface[1] = norder[10];
    for (int f = 0; f < 3; f++) { o.qface[f][0] = norder[11 + f] / 10; o.qface[f][1] = norder[11 + f] % 10; }
    o.mid[0] = norder[14] / 10; o.mid[1] = norder[14] % 10;
    return o;
  }
};

// which node of the element owns the dof: 0 vertex, 1 edge, 2 face, 3 middle ; idx = 0-based entity number
struct PrismDof {
  short t;            // index into the T list (list 0: main list of the space, list 1: the "vertical" scalar list of H(curl))
  signed char list;   // 0 / 1
  signed char sgn;
  unsigned char zk;   // KH / KQ
  unsigned char zi;
  unsigned char ent, ent_idx;
};
inline PrismDof pdof(int t, int list, int sgn, int zk, int zi, int ent, int ent_idx) {
  PrismDof d; d.t = (short)t; d.list = (signed char)list; d.sgn = (signed char)sgn; d.zk = (unsigned char)zk; d.zi = (unsigned char)zi;
  d.ent = (unsigned char)ent; d.ent_idx = (unsigned char)ent_idx; return d;
}
inline TriFn tf_vert(int a) { return tri_fn(TK_VERT, a, (a + 1) % 3, (a + 2) % 3, 0, 0); }
inline TriFn tf_edge(int e, int i) { return tri_fn(TK_EDGE, PR_TRI_EDGE[e][0], PR_TRI_EDGE[e][1], 3 - PR_TRI_EDGE[e][0] - PR_TRI_EDGE[e][1], i, 0); }
inline TriFn tf_vedge(int e, int i) { return tri_fn(TK_VEDGE, PR_TRI_EDGE[e][0], PR_TRI_EDGE[e][1], 3 - PR_TRI_EDGE[e][0] - PR_TRI_EDGE[e][1], i, 0); }

// oriented local frame of quad face f (0..2): is local axis 0 the triangle-edge (horizontal) axis? reversal flags
struct PQuadFrame { bool ax0_is_tri; int rev[2]; };
inline PQuadFrame pquad_frame(int orient) {
  PQuadFrame fr;
  fr.ax0_is_tri = !QSWAP[orient];   // (S,T) = (Nu pair, Mu): ProjectPrisQF, BlendProject.F90:722-760
  fr.rev[0] = QREV0[orient]; fr.rev[1] = QREV1[orient];
  return fr;
}

// ---- H1                                                                                      [Prism.F90:105-280]
inline std::vector<PrismDof> prism_dofs_H1(const int norder[15], const int norie[9], const int norif[5], TriList &T) {
  const PrismOrders o = PrismOrders::decode(norder);
  std::vector<PrismDof> out;
  for (int v = 0; v < 6; v++) out.push_back(pdof(T.add(tf_vert(v % 3)), 0, 1, KH, v / 3, 0, v));
  for (int e = 0; e < 6; e++)
    for (int i = 2; i <= o.edge[e]; i++) out.push_back(pdof(T.add(tf_edge(e % 3, i)), 0, norie[e] ? parity_sign(i) : 1, KH, e / 3, 1, e));
  for (int e = 0; e < 3; e++)
    for (int i = 2; i <= o.edge[6 + e]; i++) out.push_back(pdof(T.add(tf_vert(e)), 0, norie[6 + e] ? parity_sign(i) : 1, KH, i, 1, 6 + e));
  for (int f = 0; f < 2; f++) {
    const int *p = PR_OT[norif[f]];
    for (int nij = 3; nij <= o.tface[f]; nij++)
      for (int i = 2; i <= nij - 1; i++) out.push_back(pdof(T.add(tri_fn(TK_FACE, p[0], p[1], p[2], i, nij - i)), 0, 1, KH, f, 2, f));
  }
  for (int f = 0; f < 3; f++) {
    const PQuadFrame fr = pquad_frame(norif[2 + f]);
    const int n0 = o.qface[f][0], n1 = o.qface[f][1];
    for (int j = 2; j <= n1; j++)
      for (int i = 2; i <= n0; i++) {
        const int s = (fr.rev[0] ? parity_sign(i) : 1) * (fr.rev[1] ? parity_sign(j) : 1);
        const int it = fr.ax0_is_tri ? i : j, iz = fr.ax0_is_tri ? j : i;
        out.push_back(pdof(T.add(tf_edge(f, it)), 0, s, KH, iz, 2, 2 + f));
      }
  }
  if ((o.mid[0] - 1) * (o.mid[0] - 2) * (o.mid[1] - 1) / 2 > 0)
    for (int k = 2; k <= o.mid[1]; k++)
      for (int nij = 3; nij <= o.mid[0]; nij++)
        for (int i = 2; i <= nij - 1; i++) out.push_back(pdof(T.add(tri_fn(TK_FACE, 0, 1, 2, i, nij - i)), 0, 1, KH, k, 3, 0));
  return out;
}

// ---- H(curl): TV = triangle H(curl) functions (horizontal dofs, list 0), TS = triangle H1 functions (vertical, list 1)
//                                                                                              [Prism.F90:430-700]
inline std::vector<PrismDof> prism_dofs_Hcurl(const int norder[15], const int norie[9], const int norif[5], TriList &TV, TriList &TS) {
  const PrismOrders o = PrismOrders::decode(norder);
  std::vector<PrismDof> out;
  for (int e = 0; e < 6; e++)
    for (int i = 0; i <= o.edge[e] - 1; i++) out.push_back(pdof(TV.add(tf_vedge(e % 3, i)), 0, norie[e] ? -parity_sign(i) : 1, KH, e / 3, 1, e));
  for (int e = 0; e < 3; e++)
    for (int i = 0; i <= o.edge[6 + e] - 1; i++) out.push_back(pdof(TS.add(tf_vert(e)), 1, norie[6 + e] ? -parity_sign(i) : 1, KQ, i, 1, 6 + e));
  for (int f = 0; f < 2; f++) {
    const int nf = o.tface[f];
    if (nf * (nf - 1) / 2 <= 0) continue;
    const int *p = PR_OT[norif[f]];
    const size_t base = out.size();
    int cnt = 0;
    for (int nij = 1; nij <= nf - 1; nij++) for (int i = 0; i <= nij - 1; i++) cnt++;
    out.resize(base + 2 * (size_t)cnt);
    for (int fam = 0; fam < 2; fam++) {   // families interleaved: m = famctr + fam - 1, then m += 2
      int c = 0;
      for (int nij = 1; nij <= nf - 1; nij++)
        for (int i = 0; i <= nij - 1; i++, c++)
          out[base + 2 * c + fam] = pdof(TV.add(tri_fn(TK_VFACE, p[fam % 3], p[(fam + 1) % 3], p[(fam + 2) % 3], i, nij - i)), 0, 1, KH, f, 2, f);
    }
  }
  for (int f = 0; f < 3; f++) {
    const PQuadFrame fr = pquad_frame(norif[2 + f]);
    for (int fam = 0; fam < 2; fam++) {
      const int a = fam, b = 1 - fam;   // local axis carrying the Whitney factor / the H1 factor
      const int na = o.qface[f][a], nb = o.qface[f][b];
      if (na * (nb - 1) <= 0) continue;
      int lo[2], hi[2];
      lo[a] = 0; hi[a] = na - 1; lo[b] = 2; hi[b] = nb;
      const bool a_is_tri = (a == 0) == fr.ax0_is_tri;
      for (int jg = lo[1]; jg <= hi[1]; jg++)
        for (int ig = lo[0]; ig <= hi[0]; ig++) {
          const int g[2] = {ig, jg};
          const int i = g[a], j = g[b];
          const int s = (fr.rev[a] ? -parity_sign(i) : 1) * (fr.rev[b] ? parity_sign(j) : 1);
          if (a_is_tri) out.push_back(pdof(TV.add(tf_vedge(f, i)), 0, s, KH, j, 2, 2 + f));
          else out.push_back(pdof(TS.add(tf_edge(f, j)), 1, s, KQ, i, 2, 2 + f));
        }
    }
  }
  const int p = o.mid[0], pz = o.mid[1];
  if (p * (p - 1) * (pz - 1) / 2 > 0) {   // bubble families 1,2 (triangle type), interleaved
    const size_t base = out.size();
    int cnt = 0;
    for (int k = 2; k <= pz; k++) for (int nij = 1; nij <= p - 1; nij++) for (int i = 0; i <= nij - 1; i++) cnt++;
    out.resize(base + 2 * (size_t)cnt);
    for (int fam = 0; fam < 2; fam++) {
      int c = 0;
      for (int k = 2; k <= pz; k++)
        for (int nij = 1; nij <= p - 1; nij++)
          for (int i = 0; i <= nij - 1; i++, c++)
            out[base + 2 * c + fam] = pdof(TV.add(tri_fn(TK_VFACE, fam % 3, (fam + 1) % 3, (fam + 2) % 3, i, nij - i)), 0, 1, KH, k, 3, 0);
    }
  }
  if ((p - 1) * (p - 2) * pz / 2 > 0)      // bubble family 3 (quadrilateral type)
    for (int k = 0; k <= pz - 1; k++)
      for (int nij = 3; nij <= p; nij++)
        for (int i = 2; i <= nij - 1; i++) out.push_back(pdof(TS.add(tri_fn(TK_FACE, 0, 1, 2, i, nij - i)), 1, 1, KQ, k, 3, 0));
  return out;
}

// ---- L2 : k outer, (i+j, i) inner                                                           [Prism.F90:1040-1131]
inline std::vector<PrismDof> prism_dofs_L2(const int norder[15], TriList &T) {
  const PrismOrders o = PrismOrders::decode(norder);
  std::vector<PrismDof> out;
  for (int k = 0; k <= o.mid[1] - 1; k++)
    for (int nij = 0; nij <= o.mid[0] - 1; nij++)
      for (int i = 0; i <= nij; i++) out.push_back(pdof(T.add(tri_fn(TK_L2, 0, 1, 2, i, nij - i)), 0, 1, KQ, k, 3, 0));
  return out;
}

// ---- H(div), face functions only (normal traces; the middle-node order of a trace variable is forced to 11)
//   triangle faces: V = sign * T Z e_z with T an L2-type function of the oriented coordinates (list 0)        [Prism.F90:800-830]
//   quad faces    : V = sign * (T_2, -T_1, 0) Z with T a triangle-edge Whitney function (list 1), Z from Q      [Prism.F90:832-860]
inline std::vector<PrismDof> prism_dofs_Hdiv_faces(const int norder[15], const int norif[5], TriList &TZ, TriList &TH) {
  const PrismOrders o = PrismOrders::decode(norder);
  std::vector<PrismDof> out;
  for (int f = 0; f < 2; f++) {
    const int *p = PR_OT[norif[f]];
    for (int nij = 0; nij <= o.tface[f] - 1; nij++)
      for (int i = 0; i <= nij; i++) out.push_back(pdof(TZ.add(tri_fn(TK_L2, p[0], p[1], p[2], i, nij - i)), 0, PR_OT_PARITY[norif[f]], KH, f, 2, f));
  }
  for (int f = 0; f < 3; f++) {
    const PQuadFrame fr = pquad_frame(norif[2 + f]);
    const int n0 = o.qface[f][0], n1 = o.qface[f][1];
    if (n0 * n1 <= 0) continue;
    for (int j = 0; j <= n1 - 1; j++)
      for (int i = 0; i <= n0 - 1; i++) {
        int s = (fr.rev[0] ? -parity_sign(i) : 1) * (fr.rev[1] ? -parity_sign(j) : 1);
        const int it = fr.ax0_is_tri ? i : j, iz = fr.ax0_is_tri ? j : i;
        if (!fr.ax0_is_tri) s = -s;     // e_z x E = -(E x e_z)
        out.push_back(pdof(TH.add(tf_vedge(f, it)), 1, s, KQ, iz, 2, 2 + f));
      }
  }
  return out;
}

// ---- broken spaces: triangle lists of uniform order p, orientation 0 (broken/BrokenTriangle.F90, Triangle.F90:30,140)
inline void tri_list_H1(int p, TriList &T) {
  for (int a = 0; a < 3; a++) T.add(tf_vert(a));
  for (int e = 0; e < 3; e++) for (int i = 2; i <= p; i++) T.add(tf_edge(e, i));
  for (int nij = 3; nij <= p; nij++) for (int i = 2; i <= nij - 1; i++) T.add(tri_fn(TK_FACE, 0, 1, 2, i, nij - i));
}
inline void tri_list_Hcurl(int p, TriList &T) {
  for (int e = 0; e < 3; e++) for (int i = 0; i <= p - 1; i++) T.add(tf_vedge(e, i));
  for (int nij = 1; nij <= p - 1; nij++)
    for (int i = 0; i <= nij - 1; i++)
      for (int fam = 0; fam < 2; fam++) T.add(tri_fn(TK_VFACE, fam % 3, (fam + 1) % 3, (fam + 2) % 3, i, nij - i));
}

// dof counts of the middle node (ndof_nod, element_data.F90:838-846)
inline void prism_mid_counts(const int mid[2], int &h, int &e, int &v, int &q) {
  const int x = mid[0], z = mid[1];
  h = (x - 2) * (x - 1) / 2 * (z - 1);
  e = (x - 1) * x * (z - 1) + (x - 2) * (x - 1) / 2 * z;
  v = (x - 1) * x * z + x * (x + 1) / 2 * (z - 1);
  q = (x + 1) * x / 2 * z;
}
// orders the quadrature sees (set_3D_int.F90:261-277 with find_order_loc): max over triangle-type / z-type nodes
inline void prism_axis_max_order(const int norder[15], const int norif[5], int pmax[2]) {
  const PrismOrders o = PrismOrders::decode(norder);
  pmax[0] = pmax[1] = 0;
  auto up = [&](int ax, int p) { if (p > pmax[ax]) pmax[ax] = p; };
  for (int e = 0; e < 6; e++) up(0, o.edge[e]);
  for (int e = 6; e < 9; e++) up(1, o.edge[e]);
  up(0, o.tface[0]); up(0, o.tface[1]);
  for (int f = 0; f < 3; f++) {
    int h = o.qface[f][0], v = o.qface[f][1];
    if (QSWAP_ORDER[norif[2 + f]]) { int t = h; h = v; v = t; }
    up(0, h); up(1, v);
  }
  up(0, o.mid[0]); up(1, o.mid[1]);
}

// pointwise value of a 1-D table entry (for the host-side face quadrature of the trace pairings)
inline double eval_1d(int kind, int deriv, int i, double z) {
  double H[MAXN1D + 2], dH[MAXN1D + 2], Q[MAXN1D + 2];
  const int p = i < 1 ? 1 : (kind == KQ ? i + 1 : i);
  eval_tables_1d(p, 1, &z, H, dH, Q);
  if (kind == KQ) return Q[i];
  return deriv ? dH[i] : H[i];
}

}  // namespace hp3d
