/*
 * celem.c -- CPU restatement ("oracle") of the part of hp3D's celem_systemI that FOLLOWS elem + stc_fwd_wrapper:
 *   constrained-approximation transform of the condensed element system, Dirichlet lift, compression
 *   (trunk/src/constrs/celem_systemI.F90:543-785), and the COO fill of the distributed MUMPS interface
 *   (trunk/src/solver/par_mumps/par_mumps_sc.F90:419-448).  SURVEY.md 8(f), row f1.
 *
 * TEST INFRASTRUCTURE ONLY (see hp3d_oracle.h): used by tests/ to check hp3d_gpu_celem_batch.
 * Parity pins: the reference's tests hold no vectors for this routine; tests/test_celem_oracle.py pins the
 * restatement on (a) the identity case (no constraints => Zastif is a permutation of ALOC), (b) the algebraic identity
 * ZAMOD = C^T A C, ZBMOD = C^T b - ZAMOD z_D, and (c) a 1-irregular p=1 mesh on which u = xyz must be reproduced
 * through the constrained assembly (test/poly_pois.F90's criterion on a mesh with hanging nodes).
 *
 * Arithmetic is done in complex(8); the real build of the reference is the same code with zero imaginary parts
 * (every operation below maps real inputs to real outputs exactly).
 * Index arrays hold the reference's 1-based values.
 */
#include "hp3d_oracle.h"

#include <stdlib.h>
#include <string.h>

/* celem_systemI.F90:543-785.
 *   ph           physics table (PHYSA / D_TYPE / NR_COMP / ADRES, src/modules/physics.F90), active = itest = jtrial
 *   nrdofl[3]    nrdoflHi, nrdoflEi, nrdoflVi: single-component element dofs without the middle node (:132)
 *   nrcon/nac/constr[3]  output of `logic` per family: dof k is a combination of nrcon[k] modified-element dofs
 *                nac[kp + nacdim*k] (1-based, single component) with coefficients constr[kp + nacdim*k]
 *   nrdofm_f[3]  nrdofmH, nrdofmE, nrdofmV (expanded mode: all components), Nrdofm = their sum
 *   A (ni x ni), b (ni)   condensed element system after stc_fwd_wrapper, physics-blocked as stc.F90:305-323 extracts it:
 *                ALOC(i,j)%array(kk,c) = A[off(i)+kk-1 + ni*(off(j)+c-1)],  off(i) = sum_{j<i} Nrdofs(j)
 *   idbc/zdofd   Dirichlet flags/values per modified dof (:262-470); nextract: compressed -> modified dof (:311)
 *   isym         ISYM_FLAG 1 (symmetric packed), 2 (row-major), 3 (column-major) (:748-779)
 * Outputs: zbload[Nrdofc], zastif; zamod (Nrdofm x Nrdofm column-major) if non-NULL. */
int orc_celem_modify(const orc_physics *ph, const int nrdofl[3], const int *const nrcon[3], const int *const nac[3],
                     const double *const constr[3], int nacdim, const int nrdofm_f[3], int ni, const zdouble *A,
                     const zdouble *b, const int *idbc, const zdouble *zdofd, int nrdofc, const int *nextract, int isym,
                     zdouble *zbload, zdouble *zastif, zdouble *zamod_out) {
  const int nrdofm = nrdofm_f[0] + nrdofm_f[1] + nrdofm_f[2];
  /* Nrdofs(i) (:116-133, PHYSAi branch) and the block offsets of the condensed system */
  int nrdofs[ORC_MAXPHYS], off[ORC_MAXPHYS], tot = 0;
  for (int i = 0; i < ph->nphys; i++) {
    nrdofs[i] = (ph->dtype[i] <= 2 && ph->active[i]) ? nrdofl[ph->dtype[i]] * ph->ncomp[i] : 0;
    off[i] = tot;
    tot += nrdofs[i];
  }
  if (tot != ni) return -1;
  const int fam_base[3] = {0, nrdofm_f[0], nrdofm_f[0] + nrdofm_f[1]};   /* 0, nrdofmH, nrdofmHE */
  zdouble *zbmod = calloc((size_t)nrdofm > 0 ? nrdofm : 1, sizeof(zdouble));
  zdouble *aaux = calloc((size_t)nrdofm * ni + 1, sizeof(zdouble));   /* AAUX(iphys2)%array(ll, c) = aaux[ll + nrdofm*(off(iphys2)+c)] */
  zdouble *zamod = calloc((size_t)nrdofm * nrdofm + 1, sizeof(zdouble));
  if (!zbmod || !aaux || !zamod) { free(zbmod); free(aaux); free(zamod); return -2; }
  /* rows: AAUX = C^T ALOC, ZBMOD = C^T BLOC   (:553-647) */
  for (int p1 = 0; p1 < ph->nphys; p1++) {
    if (!ph->active[p1] || ph->dtype[p1] > 2) continue;
    const int f = ph->dtype[p1], nvar = ph->nrvar[f];
    for (int k = 1; k <= nrdofl[f]; k++)
      for (int kp = 1; kp <= nrcon[f][k - 1]; kp++) {
        const int l = nac[f][(kp - 1) + nacdim * (k - 1)];
        const double c = constr[f][(kp - 1) + nacdim * (k - 1)];
        for (int ivar = 1; ivar <= ph->ncomp[p1]; ivar++) {
          const int ll = fam_base[f] + (l - 1) * nvar + ph->adres[p1] + ivar;   /* 1-based */
          const int kk = (k - 1) * ph->ncomp[p1] + ivar;
          zbmod[ll - 1] = zbmod[ll - 1] + b[off[p1] + kk - 1] * c;
          for (int p2 = 0; p2 < ph->nphys; p2++) {
            if (!ph->active[p2] || ph->dtype[p2] > 2) continue;
            for (int cc = 0; cc < nrdofs[p2]; cc++) {
              zdouble *t = &aaux[(size_t)(ll - 1) + (size_t)nrdofm * (off[p2] + cc)];
              *t = *t + A[(size_t)(off[p1] + kk - 1) + (size_t)ni * (off[p2] + cc)] * c;
            }
          }
        }
      }
  }
  /* columns: ZAMOD = AAUX C   (:650-716) */
  for (int p = 0; p < ph->nphys; p++) {
    if (!ph->active[p] || ph->dtype[p] > 2) continue;
    const int f = ph->dtype[p], nvar = ph->nrvar[f];
    for (int k = 1; k <= nrdofl[f]; k++)
      for (int kp = 1; kp <= nrcon[f][k - 1]; kp++) {
        const int l = nac[f][(kp - 1) + nacdim * (k - 1)];
        const double c = constr[f][(kp - 1) + nacdim * (k - 1)];
        for (int ivar = 1; ivar <= ph->ncomp[p]; ivar++) {
          const int ll = fam_base[f] + (l - 1) * nvar + ph->adres[p] + ivar;
          const int kk = (k - 1) * ph->ncomp[p] + ivar;
          for (int r = 0; r < nrdofm; r++)
            zamod[(size_t)r + (size_t)nrdofm * (ll - 1)] =
                zamod[(size_t)r + (size_t)nrdofm * (ll - 1)] + aaux[(size_t)r + (size_t)nrdofm * (off[p] + kk - 1)] * c;
        }
      }
  }
  /* Dirichlet lift (:720-731) */
  for (int k2 = 0; k2 < nrdofm; k2++)
    if (idbc[k2] == 1)
      for (int k1 = 0; k1 < nrdofm; k1++) zbmod[k1] = zbmod[k1] - zamod[(size_t)k1 + (size_t)nrdofm * k2] * zdofd[k2];
  /* compression (:735-781) */
  for (int l1 = 1; l1 <= nrdofc; l1++) {
    const int k1 = nextract[l1 - 1];
    zbload[l1 - 1] = zbmod[k1 - 1];
    switch (isym) {
      case 1:
        for (int l2 = 1; l2 <= l1; l2++) {
          const int k2 = nextract[l2 - 1];
          const long k = (long)(l1 - 1) * l1 / 2 + l2;
          zastif[k - 1] = (zamod[(size_t)(k1 - 1) + (size_t)nrdofm * (k2 - 1)] + zamod[(size_t)(k2 - 1) + (size_t)nrdofm * (k1 - 1)]) / 2.0;
        }
        break;
      case 2:
        for (int l2 = 1; l2 <= nrdofc; l2++) {
          const int k2 = nextract[l2 - 1];
          zastif[(size_t)(l1 - 1) * nrdofc + l2 - 1] = zamod[(size_t)(k1 - 1) + (size_t)nrdofm * (k2 - 1)];
        }
        break;
      default:
        for (int l2 = 1; l2 <= nrdofc; l2++) {
          const int k2 = nextract[l2 - 1];
          zastif[(size_t)(l2 - 1) * nrdofc + l1 - 1] = zamod[(size_t)(k1 - 1) + (size_t)nrdofm * (k2 - 1)];
        }
    }
  }
  if (zamod_out) memcpy(zamod_out, zamod, sizeof(zdouble) * (size_t)nrdofm * nrdofm);
  free(zbmod); free(aaux); free(zamod);
  return 0;
}

/* par_mumps_sc.F90:419-448: the element's ndof^2 triplets in the order the element loop appends them
 * (k = (k1-1)*ndof + k2 ; IRN = LCON(k1), JCN = LCON(k2), A_loc = ZTEMP(k)) and the load accumulation RHS(LCON(k1)) += ZLOAD(k1). */
void orc_coo_fill(int ndof, const int *lcon, const zdouble *ztemp, const zdouble *zload, zdouble *a_loc, int *irn, int *jcn,
                  zdouble *rhs /* global vector, 1-based dof numbers */) {
  for (int k1 = 1; k1 <= ndof; k1++) rhs[lcon[k1 - 1] - 1] = rhs[lcon[k1 - 1] - 1] + zload[k1 - 1];
  long nnz = 0;
  for (int k1 = 1; k1 <= ndof; k1++)
    for (int k2 = 1; k2 <= ndof; k2++) {
      const long k = (long)(k1 - 1) * ndof + k2;
      a_loc[nnz] = ztemp[k - 1];
      irn[nnz] = lcon[k1 - 1];
      jcn[nnz] = lcon[k2 - 1];
      nnz++;
    }
}
