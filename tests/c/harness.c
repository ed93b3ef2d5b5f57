/* harness.c -- plain C (gcc) client of libhp3d_gpu.so: what the reference-side binding sees.
 *   harness layout            prints sizeof / offsetof of hp3d_params and hp3d_physics (compared with the Fortran bind(C)
 *                             derived types of integration/hp3d_gpu_mod.F90 by tests/test_c_harness.py; no GPU needed)
 *   harness replay <p> <nel>  replays the call sequence of INTEGRATION.md section 2 against the raw .so on a GPU: init, plan,
 *                             sizes, elem_batch (host factors), cloc_create, elem_batch_cloc, cloc_bwd_batch, and checks
 *                             xb(cloc) == BSchur - ASchur xi from the host factors, Aii Hermitian, info == 0
 * The library is loaded with dlopen so that the layout mode runs where no CUDA driver exists. */
#include <complex.h>
#include <dlfcn.h>
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/hp3d_gpu.h"

#define SYM(name) __typeof__(&name) f_##name = (__typeof__(&name))dlsym(h, #name); if (!f_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 3; }

static int layout(void) {
  printf("{\"sizeof_params\": %zu, \"params\": {", sizeof(hp3d_params));
#define F(f) printf("\"%s\": [%zu, %zu], ", #f, offsetof(hp3d_params, f), sizeof(((hp3d_params *)0)->f));
  F(nord_add) F(maxp) F(test_norm) F(alpha_norm) F(omega) F(eps) F(mu) F(sigma) F(eps_tensor) F(source) F(icomp_exact) F(store_schur) F(real_reduction) F(aii_packed)
#undef F
  printf("\"nr_rhs\": [%zu, %zu]}, ", offsetof(hp3d_params, nr_rhs), sizeof(((hp3d_params *)0)->nr_rhs));
  printf("\"sizeof_physics\": %zu, \"physics\": {", sizeof(hp3d_physics));
#define F(f) printf("\"%s\": [%zu, %zu], ", #f, offsetof(hp3d_physics, f), sizeof(((hp3d_physics *)0)->f));
  F(nphys) F(dtype) F(ncomp) F(adres)
#undef F
  printf("\"nrvar\": [%zu, %zu]}}\n", offsetof(hp3d_physics, nrvar), sizeof(((hp3d_physics *)0)->nrvar));
  return 0;
}

int main(int argc, char **argv) {
  if (argc >= 2 && !strcmp(argv[1], "layout")) return layout();
  if (argc < 5 || strcmp(argv[1], "replay")) { fprintf(stderr, "usage: harness layout | harness replay <lib.so> <p> <nel>\n"); return 2; }
  void *h = dlopen(argv[2], RTLD_NOW);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 3; }
  const int p = atoi(argv[3]), nel = atoi(argv[4]);
  SYM(hp3d_gpu_params_default) SYM(hp3d_gpu_init) SYM(hp3d_gpu_finalize) SYM(hp3d_gpu_last_error) SYM(hp3d_gpu_plan) SYM(hp3d_gpu_plan_destroy)
  SYM(hp3d_gpu_sizes_t) SYM(hp3d_gpu_elem_batch) SYM(hp3d_gpu_cloc_create) SYM(hp3d_gpu_elem_batch_cloc) SYM(hp3d_gpu_cloc_bwd_batch)
  SYM(hp3d_gpu_cloc_stats) SYM(hp3d_gpu_cloc_destroy) SYM(hp3d_gpu_host_alloc) SYM(hp3d_gpu_host_free)
#define CK(x) do { int rc_ = (x); if (rc_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #x, rc_, f_hp3d_gpu_last_error()); return 4; } } while (0)
  CK(f_hp3d_gpu_init(0));
  hp3d_params prm;
  f_hp3d_gpu_params_default(&prm);
  prm.omega = 6.283185307179586; prm.maxp = 6;
  const int plan = f_hp3d_gpu_plan(HP3D_MAXW_UW, &prm);
  if (plan < 0) { fprintf(stderr, "plan: %s\n", f_hp3d_gpu_last_error()); return 4; }
  /* find_order / find_orient of a uniform order-p brick; nodcor of a sheared unit cube (vertex dofs only carry geometry) */
  int *norder = calloc(19 * nel, sizeof(int)), *nedge = calloc(12 * nel, sizeof(int)), *nface = calloc(6 * nel, sizeof(int)), *etype = malloc(nel * sizeof(int));
  for (int e = 0; e < nel; e++) {
    etype[e] = HP3D_MDLB;
    for (int i = 0; i < 12; i++) norder[19 * e + i] = p;
    for (int i = 12; i < 18; i++) norder[19 * e + i] = 11 * p;
    norder[19 * e + 18] = 111 * p;
  }
  int ni, nb, nint, nH;
  CK(f_hp3d_gpu_sizes_t(plan, HP3D_MDLB, norder, &ni, &nb, &nint, &nH));
  const int xld = 3 * nH;
  double *xnod = calloc((size_t)xld * nel, sizeof(double));
  static const double V[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
  for (int e = 0; e < nel; e++)
    for (int v = 0; v < 8; v++) {
      const double s = 0.05 * (e + 1);
      xnod[(size_t)xld * e + 3 * v + 0] = 0.5 * (V[v][0] + s * V[v][1]);
      xnod[(size_t)xld * e + 3 * v + 1] = 0.5 * (V[v][1] + 0.3 * s * V[v][2]);
      xnod[(size_t)xld * e + 3 * v + 2] = 0.5 * V[v][2];
    }
  const long long sA = (long long)ni * ni, sB = ni, sAS = (long long)nb * ni, sBS = nb;
  double complex *Aii = f_hp3d_gpu_host_alloc(16 * sA * nel), *Bi = malloc(16 * sB * nel), *AS = malloc(16 * sAS * nel), *BS = malloc(16 * sBS * nel);
  double complex *Aii2 = malloc(16 * sA * nel), *Bi2 = malloc(16 * sB * nel);
  int *nio = malloc(nel * sizeof(int)), *nbo = malloc(nel * sizeof(int)), *info = malloc(nel * sizeof(int));
  CK(f_hp3d_gpu_elem_batch(plan, nel, etype, norder, nedge, nface, xnod, xld, NULL, 0, Aii, sA, Bi, sB, AS, sAS, BS, sBS, nio, nbo, info));
  for (int e = 0; e < nel; e++) if (info[e] != 0 || nio[e] != ni || nbo[e] != nb) { fprintf(stderr, "element %d: info %d ni %d nb %d\n", e, info[e], nio[e], nbo[e]); return 5; }
  const int cloc = f_hp3d_gpu_cloc_create(plan, 0);
  if (cloc < 0) { fprintf(stderr, "cloc: %s\n", f_hp3d_gpu_last_error()); return 4; }
  long long *iel = malloc(nel * sizeof(long long));
  for (int e = 0; e < nel; e++) iel[e] = 1000 + 7 * e;
  CK(f_hp3d_gpu_elem_batch_cloc(plan, cloc, nel, iel, etype, norder, nedge, nface, xnod, xld, NULL, 0, Aii2, sA, Bi2, sB, nio, nbo, info));
  double herm = 0, same = 0, nrm = 0;
  for (int e = 0; e < nel; e++)
    for (int c = 0; c < ni; c++)
      for (int r = 0; r < ni; r++) {
        const double complex a = Aii[sA * e + r + (long long)ni * c], b = Aii[sA * e + c + (long long)ni * r];
        herm = fmax(herm, cabs(a - conj(b))); nrm = fmax(nrm, cabs(a));
        same = fmax(same, cabs(a - Aii2[sA * e + r + (long long)ni * c]));
      }
  double complex *xi = malloc(16 * sB * nel), *xb = malloc(16 * sBS * nel);
  for (long long i = 0; i < sB * nel; i++) xi[i] = cos(0.37 * i) + I * sin(0.11 * i);
  CK(f_hp3d_gpu_cloc_bwd_batch(cloc, nel, iel, xi, sB, xb, sBS, nbo, info));
  double err = 0, xn = 0;
  for (int e = 0; e < nel; e++)
    for (int r = 0; r < nb; r++) {
      double complex s = BS[sBS * e + r];
      for (int c = 0; c < ni; c++) s -= AS[sAS * e + r + (long long)nb * c] * xi[sB * e + c];
      err = fmax(err, cabs(s - xb[sBS * e + r])); xn = fmax(xn, cabs(s));
    }
  long long st[4];
  CK(f_hp3d_gpu_cloc_stats(cloc, st));
  printf("{\"ni\": %d, \"nb\": %d, \"nint\": %d, \"hermitian_defect\": %.3e, \"cloc_vs_host_aii\": %.3e, \"amax\": %.3e, \"bwd_err\": %.3e, \"bwd_max\": %.3e, \"resident\": %lld, \"spilled\": %lld}\n",
         ni, nb, nint, herm, same, nrm, err, xn, st[0], st[1]);
  CK(f_hp3d_gpu_cloc_destroy(cloc));
  CK(f_hp3d_gpu_plan_destroy(plan));
  CK(f_hp3d_gpu_finalize());
  f_hp3d_gpu_host_free(Aii);
  return (herm == 0.0 && same == 0.0 && err <= 1e-12 * (1.0 + xn)) ? 0 : 6;
}
