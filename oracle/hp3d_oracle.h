/*
 * hp3d_oracle.h -- CPU restatement ("oracle") of hp3D's element-local hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Pinning status: the reference (Fortran 90 + PETSc/MUMPS) cannot be built in this image (no
 * Fortran compiler), and its test-suite holds no numeric golden vectors for this path.  The oracle is
 * pinned on the reference's own *analytic* known-answer tests instead (tests/test_oracle_*.py):
 *   - integer encodings: trunk/test/decode.F90, encod_decod.F90, ij_to_packed.F90 (exact);
 *   - trunk/test/poly_pois.F90:101  (u=xyz reproduced to 1e-15 through elem + stc + solve);
 *   - trunk/test/poly_maxw.F90:101  (polynomial E reproduced to 1e-14, complex);
 *   - BLAS3 `elem_opt` == scalar-loop `elem_*` twins, exact-sequence identities;
 *   - celem.c (constraints/compression): identity case, C^T A C algebra, u = xyz on a mesh with hanging nodes;
 *   - soleval.c (soleval/element_error): closed-form norms of the manufactured solutions, exactly reproduced polynomials;
 *   - pbi.c (projection-based interpolation): polynomial reproduction (what poly_pois.F90 asserts through update_gdof/update_Ddof),
 *     zero higher-order dofs of a trilinear map, identical dofs on entities shared by neighbours.
 * DPG element matrices, Cholesky condensation and p>=3 are "parity unpinned" by the reference's own
 * tests; the oracle adds the self-consistency pins listed above.
 *
 * Every function cites the reference file:line (relative to /root/reference/trunk) it follows.
 * All matrices are column-major (Fortran order); complex = interleaved (re,im) = complex(8).
 */
#ifndef HP3D_ORACLE_H
#define HP3D_ORACLE_H
#include <complex.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef double _Complex zdouble;

#define ORC_MAXP 9                 /* highest order incl. enrichment (Gauss table has <=10 points) */
#define ORC_MAXN (ORC_MAXP + 1)
#define ORC_MAXBRICK_H ((ORC_MAXP + 1) * (ORC_MAXP + 1) * (ORC_MAXP + 1))
#define ORC_MAXBRICK_E (3 * ORC_MAXP * (ORC_MAXP + 1) * (ORC_MAXP + 1))

/* ---- integer encodings: src/utility/decod.F90:21, encod.F90:21, decode.F90:19, ij_to_packed.F90:15 */
void orc_decod(int nick, int mod, int n, int *narray);
void orc_encod(const int *narray, int mod, int n, int *nick);
void orc_decode(int nick, int *j1, int *j2);
void orc_decode2(int nick, int *j1, int *j2);
void orc_ddecode(int nick, int *j1, int *j2, int *j3);
int orc_ij_upper_to_packed(int i, int j);
int orc_ij_lower_to_packed(int i, int j, int n);

/* ---- limits (src/modules/parameters.F90:17 MAXP, parametersDPG.F90:14 MAXPP=MAXP+1) */
void orc_set_maxp(int maxp);
int orc_get_maxp(void);

/* ---- 1-D polynomials: src/element/shape_1/Polynomials.F90:34,109,455,497 */
void orc_poly_legendre(double x, double t, int nord, double *P /*0..nord*/);
void orc_poly_ilegendre(double x, double t, int nord, int idec, double *L /*[2..nord]*/,
                        double *P /*[1..nord-1]*/, double *R /*[1..nord-1]*/);

/* ---- Gauss rule on [0,1]: src/element/quadrature/gauss_quadrature.F90:518-651,749-750 */
void orc_gauss1(int n, double *xi /*n*/, double *w /*n*/);

/* ---- hexahedron shape functions: src/element/shape_1/Hexahedron.F90:33,240,468,634 */
int orc_shape3DH_hexa(const double xi[3], const int nord[19], const int norie[12], const int norif[6],
                      double *shapH, double *gradH /*(3,n)*/);
int orc_shape3DE_hexa(const double xi[3], const int nord[19], const int norie[12], const int norif[6],
                      double *shapE /*(3,n)*/, double *curlE /*(3,n)*/);
int orc_shape3DV_hexa(const double xi[3], const int nord[19], const int norif[6], double *shapV /*(3,n)*/,
                      double *divV);
int orc_shape3DQ_hexa(const double xi[3], const int nord[19], double *shapQ);
/* broken (enriched test) functions: src/element/shape_1/broken/BrokenHexahedron.F90:31,139,286,422 */
int orc_shape1HH(double xi, int nord, double *shapH, double *gradH);
int orc_shape1QQ(double xi, int nord, double *shapQ);
int orc_shape3HH_hexa(const double xi[3], int nordM, double *shapH, double *gradH);
int orc_shape3EE_hexa(const double xi[3], int nordM, double *shapE, double *curlE);
int orc_shape3VV_hexa(const double xi[3], int nordM, double *shapV, double *divV);
int orc_shape3QQ_hexa(const double xi[3], int nordM, double *shapQ);

/* ---- dof counts: src/modules/element_data.F90:808 (ndof_nod), src/element/util/celndof.F90:24 */
void orc_ndof_nod_quad(int nord, int *h, int *e, int *v, int *q);
void orc_ndof_nod_hexa(int nord, int *h, int *e, int *v, int *q);
void orc_celndof_hexa(const int nord[19], int *h, int *e, int *v, int *q);
void orc_compute_enriched_order_hexa(int nordP, int norder[19]); /* MAXWELL/ULTRAWEAK_DPG/elem/elem.F90:147 */

/* ---- quadrature: src/element/quadrature/set_3D_int.F90:47,112,155 ; set_2D_int.F90:121,177 ;
 *      src/datstrs/find_order.F90:68 (find_order_loc) */
void orc_find_order_loc_hexa(const int norder[19], const int norif[6], int nloc[19]);
int orc_set_3D_int_hexa(const int norder[19], const int norif[6], int integration, int maxp,
                        double *xiloc /*(3,nint)*/, double *waloc);
int orc_set_2D_int_quad(const int nordf[5], int norif, int integration, int maxp, double *tloc /*(2,nint)*/,
                        double *wtloc);

/* ---- geometry: src/element/util/geom.F90:30, geom3D.F90:30,149 ; element_data.F90:554,616,750 */
void orc_geom(const double dxdxi[9], double dxidx[9], double *rjac, int *iflag);
void orc_geom3D(const double *xnod, const double *shapH, const double *gradH, int nrdofH, double x[3],
                double dxdxi[9], double dxidx[9], double *rjac, int *iflag);
void orc_bgeom3D(const double *xnod, const double *shapH, const double *gradH, int nrdofH,
                 const double dxidt[6], int nsign, double x[3], double dxdxi[9], double dxidx[9], double *rjac,
                 double dxdt[6], double rn[3], double *bjac);
void orc_face_param_hexa(int iface, const double t[2], double xi[3], double dxidt[6]);
int orc_nsign_param_hexa(int iface);
void orc_face_order_hexa(int iface, const int norder[19], int nordf[5]);

/* ---- element types (src/modules/node_types.F90:8-10) and the prism / triangle branch (shape_prism.c, etype.c) */
#define ORC_MDLB 1
#define ORC_MDLP 3
int orc_nvert(int et); int orc_nedge(int et); int orc_nface(int et); int orc_face_is_tri(int et, int iface);
/* triangle: src/element/shape_1/Triangle.F90:30,140,270,330 ; nord[4] = 3 edges + face */
int orc_shape2DH_tri(const double x[2], const int nord[4], const int norie[3], double *shapH, double *gradH /*(2,n)*/);
int orc_shape2DE_tri(const double x[2], const int nord[4], const int norie[3], double *shapE /*(2,n)*/, double *curlE);
int orc_shape2DV_tri(const double x[2], const int nord[4], const int norie[3], double *shapV /*(2,n)*/, double *divV);
int orc_shape2DQ_tri(const double x[2], int nordf, double *shapQ);
/* prism: src/element/shape_1/Prism.F90:38,358,760,1040 ; nord[15] = 9 edges, 2 triangle faces, 3 quad faces, middle */
int orc_shape3DH_pris(const double x[3], const int nord[15], const int norie[9], const int norif[5], double *shapH, double *gradH);
int orc_shape3DE_pris(const double x[3], const int nord[15], const int norie[9], const int norif[5], double *shapE, double *curlE);
int orc_shape3DV_pris(const double x[3], const int nord[15], const int norif[5], double *shapV, double *divV);
int orc_shape3DQ_pris(const double x[3], const int nord[15], double *shapQ);
/* broken prism: src/element/shape_1/broken/BrokenPrism.F90 ; nordM = 10*p_tri + p_z */
int orc_shape3HH_pris(const double xi[3], int nordM, double *shapH, double *gradH);
int orc_shape3EE_pris(const double xi[3], int nordM, double *shapE, double *curlE);
int orc_shape3VV_pris(const double xi[3], int nordM, double *shapV, double *divV);
int orc_shape3QQ_pris(const double xi[3], int nordM, double *shapQ);
/* select case(ntype) dispatchers: ContExactSequence.F90:397-634, broken/BrokenExactSequence.F90:401-691 */
int orc_shape3DH(int et, const double xi[3], const int *nord, const int *norie, const int *norif, double *s, double *g);
int orc_shape3DE(int et, const double xi[3], const int *nord, const int *norie, const int *norif, double *s, double *c);
int orc_shape3DV(int et, const double xi[3], const int *nord, const int *norif, double *s, double *d);
int orc_shape3DQ(int et, const double xi[3], const int *nord, double *s);
int orc_shape3HH(int et, const double xi[3], int nordM, double *s, double *g);
int orc_shape3EE(int et, const double xi[3], int nordM, double *s, double *c);
int orc_shape3VV(int et, const double xi[3], int nordM, double *s, double *d);
int orc_shape3QQ(int et, const double xi[3], int nordM, double *s);
void orc_ndof_nod_tria(int nord, int *h, int *e, int *v, int *q);
void orc_ndof_nod_pris(int nord, int *h, int *e, int *v, int *q);
void orc_ndof_nod_mid(int et, int nord, int *h, int *e, int *v, int *q);
void orc_ndof_nod_face(int et, int iface, int nord, int *h, int *e, int *v, int *q);
void orc_celndof(int et, const int *nord, int *H, int *E, int *V, int *Q);
int orc_enriched_mid(int et, int nord_mid, int dp);
int orc_trace_mid(int et);
void orc_compute_enriched_order(int et, int nordP, int *norder);
void orc_initiate_order(int et, int *norder);
int orc_set_3D_int(int et, const int *norder, const int *norif, int integration, int maxp, double *xiloc, double *waloc);
int orc_set_2D_int(int is_tri, const int nordf[5], int norif, int integration, int maxp, double *tloc, double *wtloc);
void orc_face_param(int et, int iface, const double t[2], double xi[3], double dxidt[6]);
int orc_nsign_param(int et, int iface);
void orc_face_order(int et, int iface, const int *norder, int nordf[5]);

/* ---- dense kernels (BLAS/LAPACK subset the path calls; see dense.c).  If orc_dense_use_blas() finds
 *      an OpenBLAS it forwards to it, else runs the built-in textbook loops. */
int orc_dense_use_blas(const char *libpath); /* returns 1 if loaded */
void orc_dense_set_threads(int n);

/* ---- problem parameters shared by the element routines */
typedef struct {
  int nord_add;       /* parametersDPG.F90 NORD_ADD (enrichment dp)                 */
  int test_norm;      /* MAXWELL/UW commonParam: 1=GRAPH_NORM (adjoint graph), 2=MATH_NORM, 3=GRAPH_DIAG */
  double alpha_norm;  /* ALPHA_NORM                                                  */
  double omega, eps, mu, sigma;
  zdouble eps_tensor[9]; /* get_permittivity (column-major 3x3); identity by default */
  int source;         /* 0: none, 1: manufactured "sin" (isol=1), 2: polynomial (isol=2), 9: table */
  int icomp_exact;    /* ICOMP_EXACT (1..3) for Maxwell manufactured solutions       */
  const void *source_table; /* optional per-quadrature-point source values (source==9) */
} orc_params;
void orc_params_default(orc_params *p);

/* ---- element routines (BLAS3 formulation).  Outputs are dense column-major blocks sized exactly.
 *  POISSON/GALERKIN/elem_opt.F90:22 */
int orc_elem_poisson_galerkin(const int norder[19], const int norie[12], const int norif[6], const double *xnod,
                              const orc_params *prm, double *Aloc /*(n,n)*/, double *Bloc /*n*/, int *n);
/*  POISSON/PRIMAL_DPG/elem_opt.F90:32 : trial = [H1 (nH) | H(div) trace (nVi)] */
int orc_elem_poisson_primal_dpg(const int norder[19], const int norie[12], const int norif[6],
                                const double *xnod, const orc_params *prm, double *Aloc /*(nt,nt)*/,
                                double *Bloc /*nt*/, int *nH, int *nVi);
/*  MAXWELL/GALERKIN/elem_opt.F90:22 */
int orc_elem_maxwell_galerkin(const int norder[19], const int norie[12], const int norif[6], const double *xnod,
                              const orc_params *prm, zdouble *Aloc, zdouble *Bloc, int *n);
/*  MAXWELL/ULTRAWEAK_DPG/elem/elem_opt.F90:25 : trial = [2*nEi trace | 6*nQ field]; also returns
 *  the enriched stiffness/Gram when the pointers are non-NULL (for kernel-level parity tests).   */
int orc_elem_maxwell_uw_dpg(const int norder[19], const int norie[12], const int norif[6], const double *xnod,
                            const orc_params *prm, zdouble *Aloc /*(nt,nt)*/, zdouble *Bloc /*nt*/, int *nEi,
                            int *nQ, zdouble *gram_out /*(nTest,nTest) upper*/,
                            zdouble *stiff_out /*(nTest,nt+1)*/);

/* ---- static condensation: src/modules/stc.F90:338 (herm), :443 (gen).  A is (ni+nb)^2 with the
 *  interface dofs first; on exit Aii/Bi hold the condensed system, ASchur=(nb,ni), BSchur=nb. */
int orc_stc_fwd_real(int herm, int ni, int nb, double *Aii, double *Abi, double *Aib, double *Abb, double *Bi,
                     double *Bb);
int orc_stc_fwd_cplx(int herm, int ni, int nb, zdouble *Aii, zdouble *Abi, zdouble *Aib, zdouble *Abb,
                     zdouble *Bi, zdouble *Bb);

/* ---- whole unit of work (elem + stc_fwd_wrapper) for one element; problem_kind as in include/hp3d_gpu.h */
int orc_condensed_element(int problem_kind, const int norder[19], const int norie[12], const int norif[6],
                          const double *xnod, const orc_params *prm, void *Aii, void *Bi, void *ASchur,
                          void *BSchur, int *ni, int *nb);
/* interface/bubble permutation of the full local matrix (stc.F90:226-261): perm[k] = local dof index */
int orc_stc_partition(int problem_kind, const int norder[19], int *perm, int *ni, int *nb);

/* OpenMP element loop shaped like par_mumps_sc.F90:318-357 (used for the CPU baseline) */
int orc_condensed_batch(int problem_kind, int nel, const int *norder, const int *norie, const int *norif,
                        const double *xnod, int xnod_stride, const orc_params *prm, void *Aii, void *Bi,
                        void *ASchur, void *BSchur, long sAii, long sBi, long sAS, long sBS, int *info,
                        int nthreads);

/* element-type aware variants (et = ORC_MDLB / ORC_MDLP; descriptor arrays keep the 19/12/6 layout, a prism uses
 * the first 15/9/5 entries) */
int orc_elem_poisson_galerkin_t(int et, const int norder[19], const int norie[12], const int norif[6], const double *xnod,
                                const orc_params *prm, double *Aloc, double *Bloc, int *n);
int orc_elem_poisson_primal_dpg_t(int et, const int norder[19], const int norie[12], const int norif[6], const double *xnod,
                                  const orc_params *prm, double *Aloc, double *Bloc, int *nH, int *nVi);
int orc_elem_maxwell_galerkin_t(int et, const int norder[19], const int norie[12], const int norif[6], const double *xnod,
                                const orc_params *prm, zdouble *Aloc, zdouble *Bloc, int *n);
int orc_elem_maxwell_uw_dpg_t(int et, const int norder[19], const int norie[12], const int norif[6], const double *xnod,
                              const orc_params *prm, zdouble *Aloc, zdouble *Bloc, int *nEi, int *nQ, zdouble *gram_out,
                              zdouble *stiff_out);
/* the reference's scalar-loop twin of the same element (elem_maxwell.F90): independent second formulation, see elem.c */
int orc_elem_maxwell_uw_scalar_t(int et, const int norder[19], const int norie[12], const int norif[6], const double *xnod,
                                 const orc_params *prm, zdouble *Aloc, zdouble *Bloc, zdouble *gram_out, zdouble *stiff_out);
int orc_stc_partition_t(int et, int problem_kind, const int norder[19], int *perm, int *ni, int *nb);
int orc_condensed_element_t(int et, int problem_kind, const int norder[19], const int norie[12], const int norif[6],
                            const double *xnod, const orc_params *prm, void *Aii, void *Bi, void *ASchur, void *BSchur,
                            int *ni, int *nb);
int orc_condensed_batch_t(int problem_kind, int nel, const int *etype, const int *norder, const int *norie, const int *norif,
                          const double *xnod, int xnod_stride, const orc_params *prm, void *Aii, void *Bi, void *ASchur,
                          void *BSchur, long sAii, long sBi, long sAS, long sBS, int *info, int nthreads);

/* ---- after the element: constrained-approximation transform, Dirichlet lift, compression, COO fill (celem.c).
 * Physics table of the problem (src/modules/physics.F90): D_TYPE 0 CONTIN, 1 TANGEN, 2 NORMAL, 3 DISCON. */
#define ORC_MAXPHYS 8
typedef struct orc_physics {
  int nphys;                 /* NR_PHYSA */
  int dtype[ORC_MAXPHYS];    /* D_TYPE   */
  int ncomp[ORC_MAXPHYS];    /* NR_COMP  */
  int adres[ORC_MAXPHYS];    /* ADRES: offset of the variable's first component within its family (0-based) */
  int active[ORC_MAXPHYS];   /* itest(i) = jtrial(i) = 1 and the variable has interface dofs */
  int nrvar[3];              /* NRHVAR, NREVAR, NRVVAR */
} orc_physics;
/* celem_systemI.F90:543-785 */
int orc_celem_modify(const orc_physics *ph, const int nrdofl[3], const int *const nrcon[3], const int *const nac[3],
                     const double *const constr[3], int nacdim, const int nrdofm_f[3], int ni, const zdouble *A,
                     const zdouble *b, const int *idbc, const zdouble *zdofd, int nrdofc, const int *nextract, int isym,
                     zdouble *zbload, zdouble *zastif, zdouble *zamod_out);
/* par_mumps_sc.F90:419-448 */
void orc_coo_fill(int ndof, const int *lcon, const zdouble *ztemp, const zdouble *zload, zdouble *a_loc, int *irn, int *jcn,
                  zdouble *rhs);

/* ---- solution evaluation and element error (soleval.c): soleval.F90:30, compute_error.F90:226 */
int orc_soleval(int et, const double xi[3], const int *norder, const int *norie, const int *norif, const double *xnod, int ncH,
                const zdouble *zdofH, int ncE, const zdouble *zdofE, int ncV, const zdouble *zdofV, int ncQ, const zdouble *zdofQ,
                double x[3], double dxdxi[9], double *rjac, zdouble *zsolH, zdouble *zgradH, zdouble *zsolE, zdouble *zcurlE,
                zdouble *zsolV, zdouble *zdivV, zdouble *zsolQ);
int orc_error_nvals(int kind);
void orc_exact_field(int kind, const orc_params *prm, const double x[3], zdouble *val);
int orc_element_error(int et, int kind, const int *norder, const int *norie, const int *norif, const double *xnod, const zdouble *zdof,
                      const orc_params *prm, const zdouble *exact_tab, int l2proj, double *err, double *rnorm);
int orc_error_points(int et, const int *norder, const int *norie, const int *norif, const double *xnod, double *xq);

/* ---- H1 projection-based interpolation (pbi.c): hpvert/hpedge/hpface_opt/hpmdle_opt (geometry dofs, INTEGRATION = 0) and
 * dhpvert/dhpedgeH/dhpfaceH_opt (H1 Dirichlet dofs, INTEGRATION = 1); SURVEY 8f row f4, interpolation half.
 * f(eta, val[ncomp], dval[ncomp*3], ctx): the interpolated function and its gradient in the reference coordinates eta
 * (dval[c + ncomp*i] = d g_c / d eta_i).  etav (3, nrv): reference coordinates of the element's vertices.
 * dof (ncomp, nrdofH), component fastest; nodes are numbered vertices, edges, faces, middle (bit i of mask = node i). */
typedef void (*orc_pbi_fn)(const double *eta, double *val, double *dval, void *ctx);
void orc_edge_param(int et, int ie, double t, double xi[3], double dxidt[3]);
void orc_pbi_offsets(int et, const int *norder, int *off);
int orc_pbi_node(int et, const int *norder, const int *norie, const int *norif, const double *etav, int ncomp, int integration,
                 int maxp, int node, orc_pbi_fn f, void *ctx, double *dof);
void orc_pbi_sample_fn(const double *eta, double *val, double *dval, void *ctx);
int orc_pbi_batch_sample(int nel, const int *etype, const int *norder, const int *norie, const int *norif, const double *etav,
                         int integration, int maxp, double *dof, long dof_ld, int nthreads);
int orc_pbi_element(int et, const int *norder, const int *norie, const int *norif, const double *etav, int ncomp, int integration,
                    int maxp, unsigned mask, orc_pbi_fn f, void *ctx, double *dof);

/* H(curl) Dirichlet dofs: edge/dhpedgeE.F90:24, face/dhpfaceE_opt.F90:26 (INTEGRATION = 1).  f(eta, E[ncomp*3], curlE[ncomp*3],
 * dxdeta[9], ctx) returns the datum in physical components (E[c + ncomp*j]) and the GMP Jacobian; nodes = edges, then faces. */
typedef void (*orc_pbi_fnE)(const double *eta, double *E, double *curlE, double *dxdeta, void *ctx);
void orc_pbi_offsets_E(int et, const int *norder, int *off);
int orc_pbi_hcurl_node(int et, const int *norder, const int *norie, const int *norif, const double *etav, int ncomp, int maxp, int node,
                       orc_pbi_fnE f, void *ctx, double *dofE);
int orc_pbi_hcurl_element(int et, const int *norder, const int *norie, const int *norif, const double *etav, int ncomp, int maxp, unsigned mask,
                          orc_pbi_fnE f, void *ctx, double *dofE);

/* H(div) Dirichlet dofs: face/dhpfaceV_opt.F90:26 (callback as for H(curl), the curl output is ignored); nodes = faces */
void orc_pbi_offsets_V(int et, const int *norder, int *off);
int orc_pbi_hdiv_node(int et, const int *norder, const int *norie, const int *norif, const double *etav, int ncomp, int maxp, int iface0,
                      orc_pbi_fnE f, void *ctx, double *dofV);
int orc_pbi_hdiv_element(int et, const int *norder, const int *norie, const int *norif, const double *etav, int ncomp, int maxp, unsigned mask,
                         orc_pbi_fnE f, void *ctx, double *dofV);

#ifdef __cplusplus
}
#endif
#endif
