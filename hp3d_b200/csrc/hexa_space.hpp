// hexa_space.hpp -- host-side description of hp3D's hexahedral shape functions as SIGNED TENSOR PRODUCTS
// of 1-D factors (north-star subsystem 1).
//
// On the master hexahedron every orientation-embedded H1 / H(curl) / H(div) / L2 shape function of the
// reference (src/element/shape_1/Hexahedron.F90:33,240,468,634 with Orient.F90:9,39 and
// BlendProject.F90:197-560) is, exactly,
//        sign * T_x[ix](xi_1) * T_y[iy](xi_2) * T_z[iz](xi_3)      (times the unit vector e_fam for vector spaces)
// where each 1-D factor is taken from one of two tables per axis
//        "H" = [1-x, x, L_2(x), ..., L_p(x)]   (Segment.F90:30-100 ; L_i = integrated Legendre, Polynomials.F90:109)
//        "Q" = [P_0(x), ..., P_{p-1}(x)]       (Segment.F90:138-190 ; shifted Legendre, Polynomials.F90:34)
// and an edge / face orientation only flips the argument x -> 1-x of a factor (a sign (-1)^i for L_i, P_i,
// one more -1 for the Whitney factor of an H(curl)/H(div) edge factor) or swaps the two face axes.
// This file enumerates the functions in the reference's dof order and records (family, ix, iy, iz, sign) for
// each one; the GPU kernels then only ever see 1-D tables + these integer maps.
//
// The enriched ("broken") test spaces of DPG (broken/BrokenHexahedron.F90:31,139,286) are full tensor grids
// with i fastest and no orientation, so they need no map at all.
#pragma once
#include <array>
#include <cstdlib>
#include <vector>

namespace hp3d {

// ---- topology of the master hexahedron (src/modules/element_data.F90:31-35,67-70,90-93), 0-based axes.
// vertex v sits at side VSIDE[v][d] (0: xi_d = 0, 1: xi_d = 1) of axis d
static const int VSIDE[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
// edge e runs along axis EAX[e][0]; it lies on side EAX[e][2] of axis EAX[e][1] and side EAX[e][4] of axis EAX[e][3]
static const int EAX[12][5] = {{0, 1, 0, 2, 0}, {1, 0, 1, 2, 0}, {0, 1, 1, 2, 0}, {1, 0, 0, 2, 0}, {0, 1, 0, 2, 1}, {1, 0, 1, 2, 1},
                               {0, 1, 1, 2, 1}, {1, 0, 0, 2, 1}, {2, 0, 0, 1, 0}, {2, 0, 1, 1, 0}, {2, 0, 1, 1, 1}, {2, 0, 0, 1, 1}};
// face f is the side FAX[f][1] of axis FAX[f][0]; its own (unoriented) axes are (FAX[f][2], FAX[f][3])
static const int FAX[6][4] = {{2, 0, 0, 1}, {2, 1, 0, 1}, {1, 0, 0, 2}, {0, 1, 1, 2}, {1, 1, 0, 2}, {0, 0, 1, 2}};
// quad orientations 0..7 (Orient.F90:58-99): local axis pair swapped? first local axis reversed? second reversed?
static const int QSWAP[8] = {0, 1, 0, 1, 1, 0, 1, 0};
static const int QREV0[8] = {0, 0, 1, 1, 0, 1, 1, 0};
static const int QREV1[8] = {0, 1, 1, 0, 0, 0, 1, 1};
// orientations for which the face order digits are swapped when seen in the element frame
// (NFAXES(3,:), element_data.F90:246-249 ; used by find_order_loc, src/datstrs/find_order.F90:68-90)
static const int QSWAP_ORDER[8] = {0, 1, 0, 1, 1, 0, 1, 0};

enum Kind1D { KH = 0, KQ = 1 };  // which 1-D table a factor comes from

struct TensorDof {
  signed char fam;   // vector direction (0,1,2) ; -1 for scalar spaces
  signed char sgn;   // +1 / -1
  unsigned char idx[3];  // index into the H or Q table of each axis (which table: see kind_of below)
};

// table kind of axis d for a function of family `fam` in the given space
//   H1: all H ; L2: all Q ; H(curl): Q along fam, H elsewhere ; H(div): H along fam, Q elsewhere
enum SpaceKind { SP_H1 = 0, SP_HCURL = 1, SP_HDIV = 2, SP_L2 = 3 };
inline int kind_of(SpaceKind sp, int fam, int d) {
  switch (sp) {
    case SP_H1: return KH;
    case SP_L2: return KQ;
    case SP_HCURL: return d == fam ? KQ : KH;
    default: return d == fam ? KH : KQ;
  }
}

inline int parity_sign(int i) { return (i & 1) ? -1 : 1; }

struct HexaOrders {
  int edge[12];
  int face[6][2];  // as stored at the face node: orders along the face's OWN (oriented) first / second axis
  int mid[3];
  static HexaOrders decode(const int norder[19]) {  // decimal encoding, MODORDER=10 (utility/decod.F90:21)
    HexaOrders o;
    for (int e = 0; e < 12; e++) o.edge[e] = norder[e];
    for (int f = 0; f < 6; f++) { o.face[f][0] = norder[12 + f] / 10; o.face[f][1] = norder[12 + f] % 10; }
    o.mid[0] = norder[18] / 100; o.mid[1] = (norder[18] / 10) % 10; o.mid[2] = norder[18] % 10;
    return o;
  }
};

// oriented local axes of face f: global axis + "reversed" flag of the first and second local axis
struct FaceFrame { int ax[2]; int rev[2]; int normal_ax, side; };
inline FaceFrame face_frame(int f, int orient) {
  FaceFrame fr;
  const int s = FAX[f][2], t = FAX[f][3];
  fr.ax[0] = QSWAP[orient] ? t : s;
  fr.ax[1] = QSWAP[orient] ? s : t;
  fr.rev[0] = QREV0[orient];
  fr.rev[1] = QREV1[orient];
  fr.normal_ax = FAX[f][0];
  fr.side = FAX[f][1];
  return fr;
}

inline TensorDof make_dof(int fam, int sgn, int i0, int i1, int i2) {
  TensorDof d;
  d.fam = (signed char)fam; d.sgn = (signed char)sgn;
  d.idx[0] = (unsigned char)i0; d.idx[1] = (unsigned char)i1; d.idx[2] = (unsigned char)i2;
  return d;
}

// ---- H1 : vertices, edges (i=2..p), faces (j outer, i inner), interior (k,j,i)    [Hexahedron.F90:73-156]
inline std::vector<TensorDof> hexa_dofs_H1(const int norder[19], const int norie[12], const int norif[6]) {
  const HexaOrders o = HexaOrders::decode(norder);
  std::vector<TensorDof> out;
  for (int v = 0; v < 8; v++) out.push_back(make_dof(-1, 1, VSIDE[v][0], VSIDE[v][1], VSIDE[v][2]));
  for (int e = 0; e < 12; e++)
    for (int i = 2; i <= o.edge[e]; i++) {
      int id[3];
      id[EAX[e][0]] = i; id[EAX[e][1]] = EAX[e][2]; id[EAX[e][3]] = EAX[e][4];
      out.push_back(make_dof(-1, norie[e] ? parity_sign(i) : 1, id[0], id[1], id[2]));
    }
  for (int f = 0; f < 6; f++) {
    const FaceFrame fr = face_frame(f, norif[f]);
    for (int j = 2; j <= o.face[f][1]; j++)
      for (int i = 2; i <= o.face[f][0]; i++) {
        int id[3];
        id[fr.ax[0]] = i; id[fr.ax[1]] = j; id[fr.normal_ax] = fr.side;
        int s = (fr.rev[0] ? parity_sign(i) : 1) * (fr.rev[1] ? parity_sign(j) : 1);
        out.push_back(make_dof(-1, s, id[0], id[1], id[2]));
      }
  }
  for (int k = 2; k <= o.mid[2]; k++)
    for (int j = 2; j <= o.mid[1]; j++)
      for (int i = 2; i <= o.mid[0]; i++) out.push_back(make_dof(-1, 1, i, j, k));
  return out;
}

// ---- H(curl) : edges (i=0..p-1), faces (2 families), interior (3 families)        [Hexahedron.F90:282-388]
inline std::vector<TensorDof> hexa_dofs_Hcurl(const int norder[19], const int norie[12], const int norif[6]) {
  const HexaOrders o = HexaOrders::decode(norder);
  std::vector<TensorDof> out;
  for (int e = 0; e < 12; e++)
    for (int i = 0; i <= o.edge[e] - 1; i++) {
      int id[3];
      const int a = EAX[e][0];
      id[a] = i; id[EAX[e][1]] = EAX[e][2]; id[EAX[e][3]] = EAX[e][4];
      out.push_back(make_dof(a, norie[e] ? -parity_sign(i) : 1, id[0], id[1], id[2]));
    }
  for (int f = 0; f < 6; f++) {
    const FaceFrame fr = face_frame(f, norif[f]);
    for (int fam = 0; fam < 2; fam++) {
      const int la = fam, lb = 1 - fam;  // local axis carrying the Whitney (Q) factor / the H factor
      if (o.face[f][la] * (o.face[f][lb] - 1) <= 0) continue;
      int lo[2], hi[2];
      lo[la] = 0; hi[la] = o.face[f][la] - 1; lo[lb] = 2; hi[lb] = o.face[f][lb];
      for (int jg = lo[1]; jg <= hi[1]; jg++)      // outer loop: second local axis, whichever family
        for (int ig = lo[0]; ig <= hi[0]; ig++) {
          const int g[2] = {ig, jg};
          int id[3];
          id[fr.ax[0]] = ig; id[fr.ax[1]] = jg; id[fr.normal_ax] = fr.side;
          int s = (fr.rev[la] ? -parity_sign(g[la]) : 1) * (fr.rev[lb] ? parity_sign(g[lb]) : 1);
          out.push_back(make_dof(fr.ax[la], s, id[0], id[1], id[2]));
        }
    }
  }
  for (int fam = 0; fam < 3; fam++) {
    const int a = fam, b = (fam + 1) % 3, c = (fam + 2) % 3;
    if (o.mid[a] * (o.mid[b] - 1) * (o.mid[c] - 1) <= 0) continue;
    int lo[3], hi[3];
    lo[a] = 0; hi[a] = o.mid[a] - 1; lo[b] = 2; hi[b] = o.mid[b]; lo[c] = 2; hi[c] = o.mid[c];
    for (int kg = lo[2]; kg <= hi[2]; kg++)
      for (int jg = lo[1]; jg <= hi[1]; jg++)
        for (int ig = lo[0]; ig <= hi[0]; ig++) out.push_back(make_dof(a, 1, ig, jg, kg));
  }
  return out;
}

// ---- H(div) : faces (j outer, i inner), interior (3 families)                      [Hexahedron.F90:504-569]
inline std::vector<TensorDof> hexa_dofs_Hdiv(const int norder[19], const int norif[6]) {
  const HexaOrders o = HexaOrders::decode(norder);
  std::vector<TensorDof> out;
  for (int f = 0; f < 6; f++) {
    const FaceFrame fr = face_frame(f, norif[f]);
    if (o.face[f][0] * o.face[f][1] <= 0) continue;
    // grad(s1 of local axis 0) x grad(s1 of local axis 1) = +- e_normal
    const int a0 = fr.ax[0], a1 = fr.ax[1], c = fr.normal_ax;
    const int levi = ((a0 + 1) % 3 == a1) ? 1 : -1;  // e_a0 x e_a1 = levi * e_c
    for (int j = 0; j <= o.face[f][1] - 1; j++)
      for (int i = 0; i <= o.face[f][0] - 1; i++) {
        int id[3];
        id[a0] = i; id[a1] = j; id[c] = fr.side;
        int s = levi * (fr.rev[0] ? -parity_sign(i) : 1) * (fr.rev[1] ? -parity_sign(j) : 1);
        out.push_back(make_dof(c, s, id[0], id[1], id[2]));
      }
  }
  for (int fam = 0; fam < 3; fam++) {
    const int a = fam, b = (fam + 1) % 3, c = (fam + 2) % 3;
    if (o.mid[a] * o.mid[b] * (o.mid[c] - 1) <= 0) continue;
    int lo[3], hi[3];
    lo[a] = 0; hi[a] = o.mid[a] - 1; lo[b] = 0; hi[b] = o.mid[b] - 1; lo[c] = 2; hi[c] = o.mid[c];
    for (int kg = lo[2]; kg <= hi[2]; kg++)
      for (int jg = lo[1]; jg <= hi[1]; jg++)
        for (int ig = lo[0]; ig <= hi[0]; ig++) out.push_back(make_dof(c, 1, ig, jg, kg));
  }
  return out;
}

// ---- L2 : P_i P_j P_k, i fastest                                                   [Hexahedron.F90:663-678]
inline std::vector<TensorDof> hexa_dofs_L2(const int norder[19]) {
  const HexaOrders o = HexaOrders::decode(norder);
  std::vector<TensorDof> out;
  for (int k = 0; k < o.mid[2]; k++)
    for (int j = 0; j < o.mid[1]; j++)
      for (int i = 0; i < o.mid[0]; i++) out.push_back(make_dof(-1, 1, i, j, k));
  return out;
}

// number of functions owned by the middle node (element_data.F90:808-870, ndof_nod for the brick)
inline void hexa_mid_counts(const int mid[3], int &h, int &e, int &v, int &q) {
  const int x = mid[0], y = mid[1], z = mid[2];
  h = (x - 1) * (y - 1) * (z - 1);
  e = x * (y - 1) * (z - 1) + (x - 1) * y * (z - 1) + (x - 1) * (y - 1) * z;
  v = (x - 1) * y * z + x * (y - 1) * z + x * y * (z - 1);
  q = x * y * z;
}

// orders seen in the ELEMENT frame per axis = what set_3D_int uses (set_3D_int.F90:203-235 with
// find_order_loc): the largest order met along each axis over edges, faces and the middle node.
inline void hexa_axis_max_order(const int norder[19], const int norif[6], int pmax[3]) {
  const HexaOrders o = HexaOrders::decode(norder);
  pmax[0] = pmax[1] = pmax[2] = 0;
  auto up = [&](int ax, int p) { if (p > pmax[ax]) pmax[ax] = p; };
  for (int e = 0; e < 12; e++) up(EAX[e][0], o.edge[e]);
  for (int f = 0; f < 6; f++) {
    // digits stored in the face's oriented frame; the element frame sees them swapped for QSWAP_ORDER
    int h = o.face[f][0], v = o.face[f][1];
    if (QSWAP_ORDER[norif[f]]) { int t = h; h = v; v = t; }
    up(FAX[f][2], h);
    up(FAX[f][3], v);
  }
  for (int d = 0; d < 3; d++) up(d, o.mid[d]);
}

// enriched test order of DPG: nordP = middle-node order + NORD_ADD*111 (MAXWELL/ULTRAWEAK_DPG/elem/elem.F90:67-73)
inline void hexa_enriched_mid(const int norder[19], int dp, int pe[3]) {
  const HexaOrders o = HexaOrders::decode(norder);
  for (int d = 0; d < 3; d++) pe[d] = o.mid[d] + dp;
}

}  // namespace hp3d
