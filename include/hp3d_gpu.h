/*
 * hp3d_gpu.h -- C ABI of the B200 element engine: the drop-in boundary for hp3D's element-local hot path.
 *
 * What it replaces (paths relative to the reference's trunk/):
 *   subroutine elem(Mdle, Itest,Itrial)          problems/<PROB>/elem.F90:20  (fills ALOC/BLOC, src/modules/assembly.F90:36-37)
 *   subroutine stc_fwd_wrapper(Iel,Mdle)         src/modules/stc.F90:182      (condenses ALOC/BLOC, stores CLOC(Iel)%ASchur/BSchur)
 * for ALL elements of a subdomain at once (the !$OMP DO element loop of src/solver/par_mumps/par_mumps_sc.F90:347-357
 * becomes "one batched call, then copy-out inside the unchanged loop").  Everything above that boundary (LCON, MUMPS;
 * celem_systemI's constraint transform / compression too, unless the fused hp3d_gpu_celem_batch is used) and below it on the
 * host (find_order, find_orient, nodcor, logic: the per-element descriptors) stays in the Fortran code.  INTEGRATION.md
 * shows the ISO_C_BINDING stub.
 *
 * Conventions: plain C, no C++/torch types.  All matrices are column-major; complex values are interleaved
 * (re,im) doubles == Fortran complex(8) (HP3D_COMPLEX=1, src/common/hp3d/typedefs.h:2-6).  Device, not host,
 * pointers are never exposed except through the *_dev entry points.  Every function returns 0 on success or a
 * negative HP3D_E* code; per-element LAPACK-style info (0 ok, >0 first non-positive pivot, -1 negative
 * Jacobian) is returned in `info[]` so the Fortran shim can print + stop like the reference does
 * (stc.F90:371-374, elem_opt.F90:846-849, geom3D.F90:92-109).
 */
#ifndef HP3D_GPU_H
#define HP3D_GPU_H
#ifdef __cplusplus
extern "C" {
#endif

#define HP3D_OK 0
#define HP3D_EINVAL (-1)     /* bad argument / unsupported configuration                       */
#define HP3D_ENODEV (-2)     /* no CUDA device / CUDA runtime error (see hp3d_gpu_last_error)  */
#define HP3D_ENOMEM (-3)
#define HP3D_ENOPLAN (-4)

/* problem_kind: which `elem` plugin of the reference is being replaced */
#define HP3D_POIS_GAL 1  /* problems/POISSON/GALERKIN/elem_opt.F90:22      real,    stc LU       */
#define HP3D_POIS_PDPG 2 /* problems/POISSON/PRIMAL_DPG/elem_opt.F90:32    real,    stc Cholesky */
#define HP3D_MAXW_GAL 3  /* problems/MAXWELL/GALERKIN/elem_opt.F90:22      complex, stc LU       */
#define HP3D_MAXW_UW 4   /* problems/MAXWELL/ULTRAWEAK_DPG/elem/elem_opt.F90:25  complex, stc Cholesky */

/* element types (src/modules/node_types.F90:8-10): hexahedron (brick) and triangular prism.  Per-element descriptor arrays
 * always use the brick layout (19 orders, 12 edge orientations, 6 face orientations); a prism fills the first 15 / 9 / 5
 * entries (9 edges: 6 triangle edges then 3 vertical; 2 triangle faces then 3 quad faces; middle node 10*p_xy + p_z). */
#define HP3D_MDLB 1
#define HP3D_MDLP 3

/* test norms of the ultraweak Maxwell problem (problems/MAXWELL/ULTRAWEAK_DPG/modules/commonParam.F90) */
#define HP3D_GRAPH_NORM 1
#define HP3D_MATH_NORM 2
#define HP3D_GRAPH_DIAG 3

/* source term ("getf" user callback of the reference) */
#define HP3D_SRC_ZERO 0
#define HP3D_SRC_SIN 1   /* manufactured sin solution, common/mfd_solutions.F90:80-100 (isol=1)          */
#define HP3D_SRC_TABLE 9 /* caller-supplied values at the quadrature points (see hp3d_gpu_quad_points) */

typedef struct hp3d_params {
  int nord_add;          /* parametersDPG NORD_ADD (enrichment dp); ignored by Galerkin problems       */
  int maxp;              /* parameters MAXP (src/modules/parameters.F90:17); MAXPP = maxp+1            */
  int test_norm;         /* HP3D_GRAPH_NORM ...                                                        */
  double alpha_norm;     /* ALPHA_NORM                                                                 */
  double omega, eps, mu, sigma; /* OMEGA, EPSILON, MU, SIGMA                                           */
  double eps_tensor[18]; /* get_permittivity: 3x3 complex, column-major, interleaved (identity default)*/
  int source;            /* HP3D_SRC_*                                                                 */
  int icomp_exact;       /* ICOMP_EXACT 1..3 (Maxwell manufactured solutions)                          */
  int store_schur;       /* STORE_STC: also return ASchur/BSchur (stc.F90:273-277)                     */
  int real_reduction;    /* 1 (default): ultraweak Maxwell with real eps, mu and sigma = 0 is computed through its
                            REAL form (A = T A~ T^H, T = diag(i^k): a quarter of the flops of the reference's
                            ZPOTRF/ZTRTRS/ZHERK, same result); 0: always the general complex kernels            */
  int aii_packed;        /* 1: Aii of the Hermitian (DPG) problems returns as its LOWER triangle in LAPACK packed
                            column-major storage, AP(i + (j-1)(2 ni - j)/2) = A(i,j), i >= j (ZTRTTP('L') of the Hermitian
                            Aii stc_fwd_herm returns, stc.F90:407-414): half the bytes over PCIe; the other
                            triangle is its conjugate mirror (hp3d_gpu_hermitian_unpack_batch).
                            2: the caller still receives the FULL ni x ni block, but only the lower triangle crosses
                            PCIe (as 64-column block trapezoids the copy engine places at their final position, 55 % of
                            the bytes at ni = 600); the upper blocks are mirrored by the library's host threads
                            (HP3D_HOST_THREADS; default: the CPUs the process is bound to minus one, at most 8) while the device works on the next chunks.
                            0 (default): the full block crosses PCIe                                              */
  int nr_rhs;            /* NR_RHS (src/modules/parameters.F90): number of load vectors carried through the condensation
                            (stc.F90:223-257: Bi(ni,NR_RHS), CLOC%BSchur(nb,NR_RHS)).  1 (default) for every problem of the
                            reference.  > 1: DPG problems through hp3d_gpu_elem_batch / _elem_batch_cloc / _cloc_bwd_batch /
                            _stc_bwd_batch only, sources through HP3D_SRC_TABLE: element e holds nr_rhs consecutive
                            tables (each of the element's own nint points, hp3d_gpu_sig_dims) at source_qp + e*source_ld; Bi, BSchur, xi, xb then hold nr_rhs columns per element
                            (column q at offset q*ni resp. q*nb of the element's block)                               */
} hp3d_params;

void hp3d_gpu_params_default(hp3d_params *p);

/* Select the device, create streams / workspaces lazily.  complex_mode mirrors HP3D_COMPLEX. */
int hp3d_gpu_init(int device);
int hp3d_gpu_finalize(void);
const char *hp3d_gpu_last_error(void);

/* Fix the problem + its parameters; returns a plan handle >= 0.  Plans are cheap; tables for each
 * (order, orientation) signature met later are built on first use and cached inside the plan. */
int hp3d_gpu_plan(int problem_kind, const hp3d_params *prm);
int hp3d_gpu_plan_destroy(int plan);

/* Sizes for one element signature: ni/nb as stc_get_nrdof (stc.F90:94), nint = # volume quadrature points,
 * nrdofH = # geometry dofs (columns of xnod actually read). */
int hp3d_gpu_sizes(int plan, const int *norder /*19*/, int *ni, int *nb, int *nint, int *nrdofH);   /* brick */
int hp3d_gpu_sizes_t(int plan, int etype, const int *norder /*19*/, int *ni, int *nb, int *nint, int *nrdofH);

/* Host-only: dims[8] = {ntest, ni, nb, nint, nrdofH, np, nbp, nip} of one element signature (np/nbp/nip: the padded
 * extents of the dense phase) -- what a caller needs to weight elements for load balancing (Zoltan OBJ_WEIGHT). */
int hp3d_gpu_sig_dims(int plan, int etype, const int *norder, const int *norient_edge, const int *norient_face, int *dims);

/* The batched unit of work.  For e = 0..nel-1 (elements may differ in order and orientation; they are grouped
 * by signature internally):
 *   etype[e]             element type (HP3D_MDLB or HP3D_MDLP; NULL = all bricks)
 *   norder[19*e..]       find_order            (src/datstrs/find_order.F90:5)
 *   norient_edge[12*e..] , norient_face[6*e..]  find_orient (find_orient.F90:8)
 *   xnod[xnod_ld*e..]    nodcor: geometry dofs, (3, nrdofH) column-major (src/constrs/nodcor.F90:19)
 *   source_qp            HP3D_SRC_TABLE only: element e at source_qp + e*source_ld doubles: nint values f(x_q)
 *                        (real problems) or 3*nint complex values J(x_q) (Maxwell), quadrature order of set_3D_int
 * Outputs, element e at offset e*stride (strides in SCALARS of the problem's value type; pass the sizes of the
 * largest element), each block column-major with its own exact leading dimension:
 *   Aii (ni x ni), Bi (ni)              condensed system  == ALOC/BLOC after stc_fwd_wrapper
 *   ASchur (nb x ni), BSchur (nb)       == CLOC(iel)%ASchur / %BSchur          (may be NULL if !store_schur)
 *   ni_out[e], nb_out[e], info[e]
 */
int hp3d_gpu_elem_batch(int plan, int nel, const int *etype, const int *norder, const int *norient_edge,
                        const int *norient_face, const double *xnod, int xnod_ld, const void *source_qp,
                        long long source_ld, void *Aii, long long sAii, void *Bi, long long sBi, void *ASchur,
                        long long sAS, void *BSchur, long long sBS, int *ni_out, int *nb_out, int *info);

/* Back-substitution WITHOUT stored Schur factors (the "recompute" option the reference leaves unimplemented, stc.F90:279-281;
 * stc_bwd_wrapper stc.F90:529): the element matrices and their condensation are recomputed on the device and only
 *   xb = BSchur - ASchur * xi      (nb values per element; stc_bwd, stc.F90:661-677)
 * returns to the host, instead of storing / shipping nb x ni factors per element (7.2 MB at p=5 ultraweak Maxwell).
 * xi: interface dofs of element e at xi + e*sxi scalars, in the order of the rows of Aii; xb likewise (order of ASchur rows). */
int hp3d_gpu_elem_bwd_batch(int plan, int nel, const int *etype, const int *norder, const int *norient_edge,
                            const int *norient_face, const double *xnod, int xnod_ld, const void *source_qp, long long source_ld,
                            const void *xi, long long sxi, void *xb, long long sxb, int *nb_out, int *info);

/* DPG element residual (error indicator), problems/MAXWELL/ULTRAWEAK_DPG/elem/elem_residual_maxwell.F90:246-552 and
 * POISSON/PRIMAL_DPG/elem_residual.F90 (called from residual.F90:48-57):  resid[e] = || l - B u ||^2 in the dual of the
 * test norm = (G^-1 (l - B u), l - B u), for the element solution u = (xi | xb) in the dof order of hp3d_gpu_elem_batch's
 * condensed system (interface dofs) and Schur factors (bubble dofs).  DPG problem kinds only. */
int hp3d_gpu_elem_residual_batch(int plan, int nel, const int *etype, const int *norder, const int *norient_edge,
                                 const int *norient_face, const double *xnod, int xnod_ld, const void *source_qp,
                                 long long source_ld, const void *xi, long long sxi, const void *xb, long long sxb, double *resid,
                                 int *info);

/* ------------------------------------------------------------------------------------------------------------------
 * what celem_systemI does AFTER elem + stc_fwd_wrapper, fused into the batched call --
 * constrained-approximation transform ZAMOD = C^T A C, ZBMOD = C^T b, Dirichlet lift, compression to Zbload / Zastif
 * (src/constrs/celem_systemI.F90:543-785) and, optionally, the COO triplets of the distributed MUMPS interface
 * (src/solver/par_mumps/par_mumps_sc.F90:419-448).  Only the compressed system travels to the host.
 *
 * Physics table of the problem (src/modules/physics.F90), needed to address the modified-element dofs. */
#define HP3D_MAXPHYS 8
typedef struct hp3d_physics {
  int nphys;                  /* NR_PHYSA                                                                     */
  int dtype[HP3D_MAXPHYS];    /* D_TYPE: 0 CONTIN, 1 TANGEN, 2 NORMAL, 3 DISCON                                */
  int ncomp[HP3D_MAXPHYS];    /* NR_COMP                                                                       */
  int adres[HP3D_MAXPHYS];    /* ADRES (offset of the variable's first component within its family, 0-based)   */
  int nrvar[3];               /* NRHVAR, NREVAR, NRVVAR                                                        */
} hp3d_physics;
/* the table of problem_kind as its input/physics file defines it */
int hp3d_gpu_physics_default(int problem_kind, hp3d_physics *ph);

/* Host-only: turn the output of `logic` (src/constrs/logic.F90) for ONE element into the flat per-modified-dof lists
 * hp3d_gpu_celem_batch takes.  nrcon?/nac?/constr? are the Fortran arrays nrconH(MAXbrickH), nacH(NACDIM,MAXbrickH),
 * constrH(NACDIM,MAXbrickH) (and E, V) with nacdim = NACDIM; nrdofl = (nrdoflHi, nrdoflEi, nrdoflVi), nrdofm_f =
 * (nrdofmH, nrdofmE, nrdofmV) as celem_systemI computes them (:104-236).  Modified dof ll (1..Nrdofm) receives the entries
 * cidx/cval[cptr[ll-1] .. cptr[ll]-1]: cidx = 1-based row of the condensed element system (order of the rows of Aii, i.e.
 * physics-blocked as stc.F90:305-323), in the order celem_systemI's loops meet them (so sums are formed in the same order).
 * Returns the number of entries written (<= cap) or a negative error. */
long long hp3d_gpu_celem_pack(const hp3d_physics *ph, const int *nrdofl, const int *nrconH, const int *nacH, const double *constrH,
                              const int *nrconE, const int *nacE, const double *constrE, const int *nrconV, const int *nacV,
                              const double *constrV, int nacdim, const int *nrdofm_f, long long *cptr /*[Nrdofm+1]*/, int *cidx,
                              double *cval, long long cap);

/* elem + stc_fwd_wrapper + (celem_systemI.F90:543-785) for nel elements.  Descriptors as hp3d_gpu_elem_batch.  Per element e:
 *   mptr[e]..mptr[e+1]   its modified dofs g = mptr[e] + ll - 1   (mptr[nel+1]: prefix sums of Nrdofm)
 *   cptr[g]..cptr[g+1]   entries (cidx 1-based row of Aii, cval) of modified dof g   (cptr[mptr[nel]+1], ABSOLUTE offsets);
 *                        cptr = cidx = cval = NULL: regular mesh, modified dof ll == element dof ll (Nrdofm = ni)
 *   idbc[g], zdofd[g]    IDBC / ZDOFD (value type of the problem; NR_RHS = 1)
 *   xptr[e]..xptr[e+1]   its compressed dofs (prefix sums of Nrdofc);  nextract[] = NEXTRACT (1-based ll), lcon[] = LCON
 *                        (global dof numbers, only read when irn/jcn are requested)
 *   isym_flag            ISYM_FLAG: 1 symmetric packed k=(l1-1)l1/2+l2 with (ZAMOD(k1,k2)+ZAMOD(k2,k1))/2, 2 row-major, 3 column-major
 * Outputs: zbload[xptr[e] + l1-1]; zastif at aptr[e] (SCALARS; aptr[nel+1] prefix sums of Nrdofc^2 or Nrdofc(Nrdofc+1)/2,
 * i.e. exactly the layout of mumps_par%A_loc when elements are appended one after the other); irn/jcn (may be NULL; isym_flag
 * 2 or 3) at the same positions: irn = LCON(row), jcn = LCON(column).  ASchur/BSchur as hp3d_gpu_elem_batch (may be NULL). */
int hp3d_gpu_celem_batch(int plan, int nel, const int *etype, const int *norder, const int *norient_edge, const int *norient_face,
                         const double *xnod, int xnod_ld, const void *source_qp, long long source_ld, const long long *mptr,
                         const long long *cptr, const int *cidx, const double *cval, const int *idbc, const void *zdofd,
                         const long long *xptr, const int *nextract, const int *lcon, int isym_flag, const long long *aptr,
                         void *zbload, void *zastif, int *irn, int *jcn, void *ASchur, long long sAS, void *BSchur, long long sBS,
                         int *ni_out, int *nb_out, int *info);

/* ------------------------------------------------------------------------------------------------------------------
 * SURVEY 8(f) row f4, error-evaluation half: soleval + element_error (src/element/util/soleval.F90:30,
 * src/element/util/compute_error.F90:226-579) for the FIELD variable of the plan's problem, batched over elements:
 *   err[e] = sum_l wa_l rjac_l |u_exact - u_h|^2 ,  rnorm[e] = sum_l wa_l rjac_l |u_exact|^2
 * over the points of set_3Dint with INTEGRATION = 2 (compute_error.F90:305-307), accumulated over components, values plus --
 * unless l2proj (L2PROJ) -- the gradient (H1) / curl (H(curl)).  Field variable and the layout of one point's exact values:
 *   HP3D_POIS_GAL, HP3D_POIS_PDPG  H1, 1 component        [u, du/dx, du/dy, du/dz]
 *   HP3D_MAXW_GAL                  H(curl), 1 component   [E(3), curl E(3)]
 *   HP3D_MAXW_UW                   L2, 6 components       [E(3), H(3)]
 * zdof: the variable's dofs of element e at zdof + e*szdof scalars, (ncomp, nrdof) column-major = solelm's zdofH / zdofE /
 * zdofQ rows of that variable (interface dofs first, then the middle node's: [xi ; xb] of the condensed solve).
 * exact_qp: NULL = the built-in manufactured solution (isol = 1: exact.F90 of the problem directory, mfd_solutions.F90:80-100),
 * else element e's exact values at exact_qp + e*exact_ld scalars, nvals per point in quadrature order (evaluate the problem's
 * `exact` at the points hp3d_gpu_error_points returns).  info[e] = -1 for a non-positive Jacobian. */
int hp3d_gpu_elem_error_batch(int plan, int nel, const int *etype, const int *norder, const int *norient_edge, const int *norient_face,
                              const double *xnod, int xnod_ld, const void *zdof, long long szdof, const void *exact_qp, long long exact_ld,
                              int l2proj, double *err, double *rnorm, int *info);
/* physical coordinates (3, nint) of element_error's quadrature points per element (xq may be NULL to query nint_out only) */
int hp3d_gpu_error_points(int plan, int nel, const int *etype, const int *norder, const int *norient_edge, const int *norient_face,
                          const double *xnod, int xnod_ld, double *xq, long long sxq, int *nint_out);

/* Upper bound on the number of elements processed per internal chunk by hp3d_gpu_elem_batch (0 = automatic: as many
 * as fit in device memory, but at least four chunks for large groups so that result copies overlap compute). */
int hp3d_gpu_set_chunk(int max_elements);

/* Physical coordinates of the volume quadrature points (3, nint) per element, for callers that evaluate
 * their own getf() on the host and pass the values back through source_qp. */
int hp3d_gpu_quad_points(int plan, int nel, const int *etype, const int *norder, const int *norient_edge,
                         const int *norient_face, const double *xnod, int xnod_ld, double *xq, long long sxq);

/* Back-substitution of the bubble dofs after the global solve (stc_bwd, src/modules/stc.F90:661-677):
 *   xb = BSchur - ASchur * xi      for a batch of elements with identical (ni, nb). */
int hp3d_gpu_stc_bwd_batch(int complex_mode, int nel, int ni, int nb, const void *ASchur, long long sAS,
                           const void *BSchur, long long sBS, const void *xi, long long sxi, void *xb, long long sxb);

/* ------------------------------------------------------------------------------------------------------------------
 * Device-resident CLOC: stc_fwd_wrapper with STORE_STC = .true. keeps CLOC(iel)%ASchur (nb x ni) and %BSchur (nb) of every
 * element until the back-substitution (src/modules/stc.F90:45-58,273-277); stc_bwd_wrapper (:529-677) reads them again after
 * the global solve.  Nothing on the host looks at them in between, so they stay in HBM under the caller's element index
 * (7.2 of the 13.0 MB an ultraweak Maxwell p=5 element produces never cross PCIe):
 *   hp3d_gpu_cloc_create     a store for `plan`, holding at most limit_bytes of HBM (0 = 60 % of what is free now);
 *                            returns a handle >= 0
 *   hp3d_gpu_elem_batch_cloc hp3d_gpu_elem_batch, with the Schur factors of element e filed under iel[e] (NULL: e) instead
 *                            of returning; an element met again (same iel) overwrites its factors.  Elements beyond the
 *                            byte limit are SPILLED: only their descriptors are kept (on the host) and the
 *                            back-substitution recomputes them (hp3d_gpu_elem_bwd_batch's path) -- same results
 *   hp3d_gpu_celem_batch_cloc  likewise for the fused hp3d_gpu_celem_batch
 *   hp3d_gpu_cloc_bwd_batch  stc_bwd for nel elements named by iel[]:  xb = BSchur - ASchur xi  (stc.F90:661-677);
 *                            xi at xi + e*sxi scalars (ni values, order of the rows of Aii), xb likewise (nb values)
 *   hp3d_gpu_cloc_fetch      one element's factors to the host (tests, debugging); ASchur / BSchur may be NULL;
 *                            returns 1 if the element is spilled (nothing written), 0 if resident
 *   hp3d_gpu_cloc_stats      stats[4] = {resident elements, spilled elements, bytes held, byte limit}
 *   hp3d_gpu_cloc_clear      forget every element (between refinement steps); hp3d_gpu_cloc_destroy also frees the handle */
int hp3d_gpu_cloc_create(int plan, long long limit_bytes);
int hp3d_gpu_cloc_clear(int cloc);
int hp3d_gpu_cloc_destroy(int cloc);
int hp3d_gpu_cloc_stats(int cloc, long long *stats /*4*/);
int hp3d_gpu_elem_batch_cloc(int plan, int cloc, int nel, const long long *iel, const int *etype, const int *norder,
                             const int *norient_edge, const int *norient_face, const double *xnod, int xnod_ld,
                             const void *source_qp, long long source_ld, void *Aii, long long sAii, void *Bi, long long sBi,
                             int *ni_out, int *nb_out, int *info);
int hp3d_gpu_celem_batch_cloc(int plan, int cloc, int nel, const long long *iel, const int *etype, const int *norder,
                              const int *norient_edge, const int *norient_face, const double *xnod, int xnod_ld,
                              const void *source_qp, long long source_ld, const long long *mptr, const long long *cptr,
                              const int *cidx, const double *cval, const int *idbc, const void *zdofd, const long long *xptr,
                              const int *nextract, const int *lcon, int isym_flag, const long long *aptr, void *zbload,
                              void *zastif, int *irn, int *jcn, int *ni_out, int *nb_out, int *info);
int hp3d_gpu_cloc_bwd_batch(int cloc, int nel, const long long *iel, const void *xi, long long sxi, void *xb, long long sxb,
                            int *nb_out, int *info);
int hp3d_gpu_cloc_fetch(int cloc, long long iel, void *ASchur, void *BSchur, int *ni, int *nb);

/* Host only (threads = 0: all hardware threads): the full Hermitian (complex_mode) / symmetric matrices from packed lower
 * triangles (hp3d_params.aii_packed): A(i,j) = AP(i + (j-1)(2 ni - j)/2) for i >= j, A(j,i) = conj(A(i,j)); element e reads
 * AP + e*sAP scalars and writes A + e*sA scalars (ni x ni column-major, ni_e[e] or `ni` if ni_e == NULL).  In place is NOT
 * supported.  This is the one host-side step the Fortran shim adds before copying into ALOC (ZTPTTR + mirror). */
int hp3d_gpu_hermitian_unpack_batch(int complex_mode, int nel, int ni, const int *ni_e, const void *AP, long long sAP, void *A,
                                    long long sA, int threads);

/* Throughput driver used by bench.py: runs the hot path `reps` times over `nel` RESIDENT elements (geometry dofs
 * already in HBM, condensed outputs left in HBM), timed with CUDA events on the launching stream.
 *   lanes      1: chunks run back to back on one stream (stage times ms_integ / ms_dense are then meaningful);
 *              2..4: chunks rotate over that many buffer sets / streams, as hp3d_gpu_elem_batch runs them (it uses four)
 *   ms_total   device time of all reps;  ms_integ / ms_dense: the part spent in integration / in the dense phase
 *   launches   kernels launched inside the timed region
 * (The end-to-end number with host buffers is measured by calling hp3d_gpu_elem_batch itself.) */
int hp3d_gpu_bench(int plan, int nel, const int *norder, const int *norient_edge, const int *norient_face,
                   const double *xnod, int xnod_ld, int reps, int max_chunk, int lanes, double *ms_total,
                   double *ms_integ, double *ms_dense, long long *launches);

/* same, with per-element element types (NULL = all bricks) */
int hp3d_gpu_bench_t(int plan, int nel, const int *etype, const int *norder, const int *norient_edge, const int *norient_face,
                     const double *xnod, int xnod_ld, int reps, int max_chunk, int lanes, double *ms_total, double *ms_integ,
                     double *ms_dense, long long *launches);

/* FP64 tensor-pipe (DMMA, mma.sync m8n8k4 f64) issue-rate probe on the current device: the measured roofline denominator
 * of the dense phase (about 0.1 s; tflops = 2*8*8*4 flops per warp instruction / best of four timed launches). */
int hp3d_gpu_fp64_peak_probe(double *tflops, double *ms_best);

/* Page-locked host memory for the caller's result arrays (so that the D2H copies of hp3d_gpu_elem_batch are
 * asynchronous DMA transfers that overlap the next chunk's kernels).  Pageable buffers work too, only slower. */
void *hp3d_gpu_host_alloc(long long bytes);
void hp3d_gpu_host_free(void *p);

/* ---- H1 projection-based interpolation (interpolation half): geometry dofs and H1 Dirichlet dofs.
 * Replaces, for all elements of a subdomain at once, the node-by-node calls of
 *   update_gdof  (src/hpinterp/update_gdof.F90:88-200,409-435): hpvert.F90:22, hpedge.F90:23, hpface_opt.F90:24, hpmdle_opt.F90:23
 *   update_Ddof  (src/hpinterp/update_Ddof.F90):                 dhpvert.F90:19, edge/dhpedgeH.F90:25, face/dhpfaceH_opt.F90:27
 * The interpolated function g has `ncomp` REAL components (the GMP map x(eta): 3; a complex Dirichlet datum: re/im interleaved,
 * 2 NREQNH) and is projected in the reference coordinates eta of the GMP block; eta(xi) is the multilinear map through
 *   etav   (3, 8) per element   reference coordinates of the element's vertices (refel's xsub; a prism uses the first 6).
 * Nodes are numbered as in the element's nodesl: vertices, edges, faces, middle node (27 for a brick, 21 for a prism);
 * bit i of mask[e] selects node i (mask == NULL: every node, update_gdof; for update_Ddof: the nodes with a Dirichlet flag).
 *
 * hp3d_gpu_pbi_points (host only, no GPU needed) returns where the host has to evaluate g:
 *   xi      (3, npts[e]) master coordinates, stride xi_ld per element (NULL: sizes only); the points of the edges (set_1Dint +
 *           edge_param), then of the faces (set_2Dint + face_param), then of the middle node (set_3Dint), each with
 *           INTEGRATION = integration (0 in update_gdof, 1 in update_Ddof) and orders capped at maxp (MAXP)
 *   nodes   (4, 27) per element: first dof, number of dofs, first point, number of points of every node (vertices have
 *           no points; a node without dofs has none either, as the reference returns before integrating)
 * hp3d_gpu_pbi_h1_batch:
 *   fvert   (ncomp, 8) per element         g at the vertices (hpvert / dhpvert)
 *   fgrad   (ncomp, 3, npts) per element   dg_c/deta_i at the points, component fastest (dxdeta(1:3,1:3) of `hexa/prism(No,eta,..)`;
 *           zdvalH * dxdeta for Dirichlet data, dhpfaceH_opt.F90:217-224), stride fgrad_ld doubles
 *   dof     (ncomp, nrdofH) per element, component fastest, reference dof order, stride dof_ld; IN: the dofs of the nodes that
 *           are not selected (they enter the projections of the higher-dimensional nodes), OUT: the selected nodes' dofs
 *   info    per element: 0, -1 (Jacobian of eta(xi) not positive), i > 0 (stiffness of node i not positive definite) */
int hp3d_gpu_pbi_points(int nel, const int *etype, const int *norder, const int *norient_edge, const int *norient_face, int integration,
                        int maxp, double *xi, long long xi_ld, int *npts, int *nrdofH, int *nodes);
int hp3d_gpu_pbi_h1_batch(int nel, const int *etype, const int *norder, const int *norient_edge, const int *norient_face, int integration,
                          int maxp, const double *etav, int ncomp, const double *fvert, const double *fgrad, long long fgrad_ld,
                          const unsigned *mask, double *dof, long long dof_ld, int *info);

/* H(curl) Dirichlet dofs of update_Ddof: edge/dhpedgeE.F90:24 and face/dhpfaceE_opt.F90:26 for all elements at once (INTEGRATION = 1;
 * middle nodes carry no Dirichlet data).  Same node numbering / mask bits as above (vertex and middle bits are ignored).
 *   hp3d_gpu_pbi_hcurl_points: points of the edge and face nodes that own H(curl) dofs (an order-1 edge has one), nrdofE = number of
 *           edge + face dofs of the element (they come first in the reference's H(curl) dof order), nodes (4, 27) as above
 *   fval    (ncomp, 3, npts) per element: the datum pulled back to eta, E_eta(i) = sum_j E_j dxdeta(j,i)        (dhpfaceE_opt.F90:273-274)
 *   fcurl   (ncomp, 3, npts) per element: curl_eta(i) = det(dxdeta) sum_j dxdeta^-1(i,j) (curl E)_j             (:275-277); edges ignore it
 *           both component fastest, stride f_ld doubles; ncomp REAL components (complex data: re/im interleaved)
 *   dof     (ncomp, nrdofE) per element, stride dof_ld; IN: dofs of unselected edges, OUT: selected nodes' dofs
 *   info    per element: 0, -1 (Jacobian not positive), i > 0 (singular system at node i) */
int hp3d_gpu_pbi_hcurl_points(int nel, const int *etype, const int *norder, const int *norient_edge, const int *norient_face, int maxp,
                              double *xi, long long xi_ld, int *npts, int *nrdofE, int *nodes);
int hp3d_gpu_pbi_hcurl_batch(int nel, const int *etype, const int *norder, const int *norient_edge, const int *norient_face, int maxp,
                             const double *etav, int ncomp, const double *fval, const double *fcurl, long long f_ld, const unsigned *mask,
                             double *dof, long long dof_ld, int *info);

/* H(div) Dirichlet dofs of update_Ddof: face/dhpfaceV_opt.F90:26 (INTEGRATION = 1): L2 projection of the normal component on every
 * selected face (bits of the face nodes in mask; the others are ignored).
 *   fval    (ncomp, 3, npts) per element: V_eta(i) = det(dxdeta) sum_j dxdeta^-1(i,j) V_j at hp3d_gpu_pbi_hdiv_points  (:231-238)
 *   dof     (ncomp, nrdofV) per element, nrdofV = number of face dofs (they come first in the reference's H(div) dof order) */
int hp3d_gpu_pbi_hdiv_points(int nel, const int *etype, const int *norder, const int *norient_edge, const int *norient_face, int maxp,
                             double *xi, long long xi_ld, int *npts, int *nrdofV, int *nodes);
int hp3d_gpu_pbi_hdiv_batch(int nel, const int *etype, const int *norder, const int *norient_edge, const int *norient_face, int maxp,
                            const double *etav, int ncomp, const double *fval, long long f_ld, const unsigned *mask, double *dof,
                            long long dof_ld, int *info);

/* Device memory the cached signature tables of the interpolation entry points may hold (default 8 GiB; one H1 signature is 2.6 MB at
 * p=5, an hp mesh has thousands).  When a batch call starts above the limit the tables are dropped and rebuilt on demand. */
int hp3d_gpu_pbi_cache_limit(long long bytes);

/* ---- host-only introspection (no GPU needed): the signed tensor-product description of the shape functions.
 * space: 0 H1, 1 H(curl), 2 H(div), 3 L2.  For dof k (reference order, src/element/shape_1/Hexahedron.F90):
 *   fam[k] vector direction (0..2, -1 scalar), idx[3k..3k+2] 1-D table index per axis, sgn[k] = +-1.
 * Returns the number of dofs (or a negative error); arrays may be NULL to query the count. */
int hp3d_gpu_dof_map(int space, const int *norder, const int *norient_edge, const int *norient_face, int cap, int *fam,
                     int *idx, int *sgn);
/* Prism analogue (host only): values of the shape functions of one prism at the master point xi[3], evaluated through the
 * product's (triangle function) x (1-D table) decomposition, reference dof order (src/element/shape_1/Prism.F90).
 *   space 0: val[3k] = phi_k, der[3k..] = grad ; 1: val = E_k, der = curl E_k ; 2: face functions only, val = V_k, der[3k] = div ;
 *   3: val[3k] = q_k.  Returns the number of functions; val == NULL queries the count. */
int hp3d_gpu_prism_shape(int space, const int *norder, const int *norient_edge, const int *norient_face, const double *xi, int cap,
                         double *val, double *der);
/* 1-D Gauss rule on [0,1] (nq points) and the tables H[(p+1) x nq], dH[(p+1) x nq], Q[p x nq] evaluated at it */
int hp3d_gpu_tables_1d(int p, int nq, double *x, double *w, double *H, double *dH, double *Q);

/* Test hook: integrate ONE element of a DPG plan and return the dense phase's raw input buffer W
 * (planes x R x np doubles, row-major, see hp3d_b200/csrc/dense_pipeline.cuh) with dims[8] =
 * {np, nbp, nip, n, nb, ni, R, planes}; nip = the interface rows that carry data (a multiple of 32; the load rows are the last
 * of them), R = np + nbp + pad64(nip).  For non-DPG plans the buffer is Am (planes x M x M), dims[0] = 0. */
int hp3d_gpu_integrate_debug(int plan, const int *norder, const int *norient_edge, const int *norient_face,
                             const double *xnod, const void *source_qp, double *W, long long cap_doubles, int *dims);

int hp3d_gpu_integrate_debug_t(int plan, int etype, const int *norder, const int *norient_edge, const int *norient_face,
                               const double *xnod, const void *source_qp, double *W, long long cap_doubles, int *dims);

/* Test hook: run only the dense phase (DPG normal equations + static condensation) on caller-provided
 * Gram / enriched stiffness matrices.  G: (n x n) Hermitian, upper triangle read; Bm: n x (nb+ni+1), columns
 * ordered [bubble | interface | load].  cplx selects real(8)/complex(8). */
/* Test hook (host only): the chunk sizes hp3d_gpu_elem_batch uses for `ntot` elements of one dense class with at most `cap`
 * elements per chunk on `nlanes` lanes (ramped start, geometric taper); returns the number of chunks. */
int hp3d_gpu_chunk_plan_debug(long long ntot, int cap, int nlanes, int max_chunk, long long *sizes, int cap_sizes);

int hp3d_gpu_dense_debug(int cplx, int nel, int n, int nb, int ni, const void *G, const void *Bm, void *Aii,
                         void *Bi, void *ASchur, void *BSchur, int *info);

#ifdef __cplusplus
}
#endif
#endif
