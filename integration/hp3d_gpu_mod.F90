!-----------------------------------------------------------------------------------------------------------------------
!  hp3d_gpu_mod.F90 -- ISO_C_BINDING interface of libhp3d_gpu.so (include/hp3d_gpu.h) for hp3D.
!
!  What a maintainer adds as trunk/src/modules/hp3d_gpu.F90.  It is NOT compiled in this repository (the build image has no
!  Fortran compiler); every interface below mirrors one prototype of include/hp3d_gpu.h argument by argument, and
!  INTEGRATION.md shows the call sites in par_mumps_sc / celem_systemI / compute_error / residual.
!  Conventions: arrays are passed as they are in hp3D (column-major, complex(8) = interleaved doubles); VTYPE arrays go
!  through c_loc(); index VALUES stay 1-based (nac, NEXTRACT, LCON), prefix arrays (mptr, cptr, xptr, aptr) are 0-based offsets.
!-----------------------------------------------------------------------------------------------------------------------
#include "typedefs.h"
module hp3d_gpu
   use, intrinsic :: iso_c_binding
   implicit none
   integer(c_int), parameter :: HP3D_POIS_GAL=1, HP3D_POIS_PDPG=2, HP3D_MAXW_GAL=3, HP3D_MAXW_UW=4
   integer(c_int), parameter :: HP3D_MDLB=1, HP3D_MDLP=3          ! == node_types MDLB / MDLP (src/modules/node_types.F90:8-10)
   integer(c_int), parameter :: HP3D_SRC_ZERO=0, HP3D_SRC_SIN=1, HP3D_SRC_TABLE=9
   integer(c_int), parameter :: HP3D_MAXPHYS=8
!
   type, bind(C) :: hp3d_params                 ! struct hp3d_params
      integer(c_int) :: nord_add, maxp, test_norm
      real(c_double) :: alpha_norm, omega, eps, mu, sigma
      real(c_double) :: eps_tensor(18)
      integer(c_int) :: source, icomp_exact, store_schur, real_reduction, aii_packed, nr_rhs
   end type
   type, bind(C) :: hp3d_physics                ! struct hp3d_physics (src/modules/physics.F90: D_TYPE, NR_COMP, ADRES, NR?VAR)
      integer(c_int) :: nphys
      integer(c_int) :: dtype(HP3D_MAXPHYS), ncomp(HP3D_MAXPHYS), adres(HP3D_MAXPHYS)
      integer(c_int) :: nrvar(3)
   end type
!
   interface
      subroutine hp3d_gpu_params_default(p) bind(C)
         import; type(hp3d_params) :: p
      end subroutine
      integer(c_int) function hp3d_gpu_init(device) bind(C)
         import; integer(c_int), value :: device
      end function
      integer(c_int) function hp3d_gpu_finalize() bind(C)
         import
      end function
      function hp3d_gpu_last_error() bind(C) result(msg)
         import; type(c_ptr) :: msg
      end function
      integer(c_int) function hp3d_gpu_plan(kind, prm) bind(C)
         import; integer(c_int), value :: kind; type(hp3d_params) :: prm
      end function
      integer(c_int) function hp3d_gpu_plan_destroy(plan) bind(C)
         import; integer(c_int), value :: plan
      end function
      integer(c_int) function hp3d_gpu_sizes_t(plan, etype, norder, ni, nb, nint, nrdofH) bind(C)
         import; integer(c_int), value :: plan, etype; integer(c_int) :: norder(19), ni, nb, nint, nrdofH
      end function
      type(c_ptr) function hp3d_gpu_host_alloc(bytes) bind(C)           ! pinned memory for the result arrays
         import; integer(c_long_long), value :: bytes
      end function
      subroutine hp3d_gpu_host_free(p) bind(C)
         import; type(c_ptr), value :: p
      end subroutine
!
!  ...elem + stc_fwd_wrapper for all elements of a subdomain (celem_systemI.F90:523-534)
      integer(c_int) function hp3d_gpu_elem_batch(plan, nel, etype, norder, norient_edge, norient_face, xnod, xnod_ld,     &
                          source_qp, source_ld, Aii, sAii, Bi, sBi, ASchur, sAS, BSchur, sBS, ni_out, nb_out, info) bind(C)
         import
         integer(c_int), value :: plan, nel, xnod_ld
         integer(c_long_long), value :: source_ld, sAii, sBi, sAS, sBS
         integer(c_int) :: etype(*), norder(19,*), norient_edge(12,*), norient_face(6,*), ni_out(*), nb_out(*), info(*)
         real(c_double) :: xnod(xnod_ld,*)
         type(c_ptr), value :: source_qp, Aii, Bi, ASchur, BSchur
      end function
!
!  ...the same + celem_systemI.F90:543-785 (constraints, Dirichlet lift, compression) + par_mumps_sc.F90:433-448 (IRN/JCN)
      integer(c_long_long) function hp3d_gpu_celem_pack(ph, nrdofl, nrconH, nacH, constrH, nrconE, nacE, constrE,            &
                          nrconV, nacV, constrV, nacdim, nrdofm_f, cptr, cidx, cval, cap) bind(C)
         import
         type(hp3d_physics) :: ph
         integer(c_int) :: nrdofl(3), nrconH(*), nacH(*), nrconE(*), nacE(*), nrconV(*), nacV(*), nrdofm_f(3), cidx(*)
         real(c_double) :: constrH(*), constrE(*), constrV(*), cval(*)
         integer(c_int), value :: nacdim
         integer(c_long_long) :: cptr(*)
         integer(c_long_long), value :: cap
      end function
      integer(c_int) function hp3d_gpu_physics_default(kind, ph) bind(C)
         import; integer(c_int), value :: kind; type(hp3d_physics) :: ph
      end function
      integer(c_int) function hp3d_gpu_celem_batch(plan, nel, etype, norder, norient_edge, norient_face, xnod, xnod_ld,      &
                          source_qp, source_ld, mptr, cptr, cidx, cval, idbc, zdofd, xptr, nextract, lcon, isym_flag, aptr,   &
                          zbload, zastif, irn, jcn, ASchur, sAS, BSchur, sBS, ni_out, nb_out, info) bind(C)
         import
         integer(c_int), value :: plan, nel, xnod_ld, isym_flag
         integer(c_long_long), value :: source_ld, sAS, sBS
         integer(c_int) :: etype(*), norder(19,*), norient_edge(12,*), norient_face(6,*), cidx(*), idbc(*), nextract(*), lcon(*)
         integer(c_int) :: ni_out(*), nb_out(*), info(*)
         integer(c_long_long) :: mptr(*), cptr(*), xptr(*), aptr(*)
         real(c_double) :: xnod(xnod_ld,*), cval(*)
         type(c_ptr), value :: source_qp, zdofd, zbload, zastif, irn, jcn, ASchur, BSchur     ! irn/jcn: c_loc(IRN_loc) or c_null_ptr
      end function
!
!  ...stc_bwd without stored factors (stc.F90:279-281,661-677) and the DPG element residual (residual.F90:48-57)
      integer(c_int) function hp3d_gpu_elem_bwd_batch(plan, nel, etype, norder, norient_edge, norient_face, xnod, xnod_ld,   &
                          source_qp, source_ld, xi, sxi, xb, sxb, nb_out, info) bind(C)
         import
         integer(c_int), value :: plan, nel, xnod_ld
         integer(c_long_long), value :: source_ld, sxi, sxb
         integer(c_int) :: etype(*), norder(19,*), norient_edge(12,*), norient_face(6,*), nb_out(*), info(*)
         real(c_double) :: xnod(xnod_ld,*)
         type(c_ptr), value :: source_qp, xi, xb
      end function
      integer(c_int) function hp3d_gpu_elem_residual_batch(plan, nel, etype, norder, norient_edge, norient_face, xnod,       &
                          xnod_ld, source_qp, source_ld, xi, sxi, xb, sxb, resid, info) bind(C)
         import
         integer(c_int), value :: plan, nel, xnod_ld
         integer(c_long_long), value :: source_ld, sxi, sxb
         integer(c_int) :: etype(*), norder(19,*), norient_edge(12,*), norient_face(6,*), info(*)
         real(c_double) :: xnod(xnod_ld,*), resid(*)
         type(c_ptr), value :: source_qp, xi, xb
      end function
!
!  ...element_error (compute_error.F90:226) for the field variable; exact_qp = c_null_ptr selects the built-in solution
      integer(c_int) function hp3d_gpu_elem_error_batch(plan, nel, etype, norder, norient_edge, norient_face, xnod, xnod_ld, &
                          zdof, szdof, exact_qp, exact_ld, l2proj, err, rnorm, info) bind(C)
         import
         integer(c_int), value :: plan, nel, xnod_ld, l2proj
         integer(c_long_long), value :: szdof, exact_ld
         integer(c_int) :: etype(*), norder(19,*), norient_edge(12,*), norient_face(6,*), info(*)
         real(c_double) :: xnod(xnod_ld,*), err(*), rnorm(*)
         type(c_ptr), value :: zdof, exact_qp
      end function
      integer(c_int) function hp3d_gpu_error_points(plan, nel, etype, norder, norient_edge, norient_face, xnod, xnod_ld,      &
                          xq, sxq, nint_out) bind(C)
         import
         integer(c_int), value :: plan, nel, xnod_ld
         integer(c_long_long), value :: sxq
         integer(c_int) :: etype(*), norder(19,*), norient_edge(12,*), norient_face(6,*), nint_out(*)
         real(c_double) :: xnod(xnod_ld,*)
         type(c_ptr), value :: xq
      end function
      integer(c_int) function hp3d_gpu_quad_points(plan, nel, etype, norder, norient_edge, norient_face, xnod, xnod_ld,       &
                          xq, sxq) bind(C)
         import
         integer(c_int), value :: plan, nel, xnod_ld
         integer(c_long_long), value :: sxq
         integer(c_int) :: etype(*), norder(19,*), norient_edge(12,*), norient_face(6,*)
         real(c_double) :: xnod(xnod_ld,*), xq(*)
      end function
      integer(c_int) function hp3d_gpu_pbi_points(nel, etype, norder, norient_edge, norient_face, integration, maxp,          &
                          xi, xi_ld, npts, nrdofH, nodes) bind(C)
         import
         integer(c_int), value :: nel, integration, maxp
         integer(c_long_long), value :: xi_ld
         integer(c_int) :: etype(*), norder(19,*), norient_edge(12,*), norient_face(6,*)
         type(c_ptr), value :: xi, npts, nrdofH, nodes
      end function
      integer(c_int) function hp3d_gpu_pbi_h1_batch(nel, etype, norder, norient_edge, norient_face, integration, maxp,        &
                          etav, ncomp, fvert, fgrad, fgrad_ld, mask, dof, dof_ld, info) bind(C)
         import
         integer(c_int), value :: nel, integration, maxp, ncomp
         integer(c_long_long), value :: fgrad_ld, dof_ld
         integer(c_int) :: etype(*), norder(19,*), norient_edge(12,*), norient_face(6,*), info(*)
         real(c_double) :: etav(3,8,*), fvert(ncomp,8,*), fgrad(fgrad_ld,*), dof(dof_ld,*)
         type(c_ptr), value :: mask
      end function
      integer(c_int) function hp3d_gpu_pbi_hcurl_points(nel, etype, norder, norient_edge, norient_face, maxp,                 &
                          xi, xi_ld, npts, nrdofE, nodes) bind(C)
         import
         integer(c_int), value :: nel, maxp
         integer(c_long_long), value :: xi_ld
         integer(c_int) :: etype(*), norder(19,*), norient_edge(12,*), norient_face(6,*)
         type(c_ptr), value :: xi, npts, nrdofE, nodes
      end function
      integer(c_int) function hp3d_gpu_pbi_hcurl_batch(nel, etype, norder, norient_edge, norient_face, maxp,                  &
                          etav, ncomp, fval, fcurl, f_ld, mask, dof, dof_ld, info) bind(C)
         import
         integer(c_int), value :: nel, maxp, ncomp
         integer(c_long_long), value :: f_ld, dof_ld
         integer(c_int) :: etype(*), norder(19,*), norient_edge(12,*), norient_face(6,*), info(*)
         real(c_double) :: etav(3,8,*), fval(f_ld,*), fcurl(f_ld,*), dof(dof_ld,*)
         type(c_ptr), value :: mask
      end function
      integer(c_int) function hp3d_gpu_pbi_hdiv_points(nel, etype, norder, norient_edge, norient_face, maxp,                  &
                          xi, xi_ld, npts, nrdofV, nodes) bind(C)
         import
         integer(c_int), value :: nel, maxp
         integer(c_long_long), value :: xi_ld
         integer(c_int) :: etype(*), norder(19,*), norient_edge(12,*), norient_face(6,*)
         type(c_ptr), value :: xi, npts, nrdofV, nodes
      end function
      integer(c_int) function hp3d_gpu_pbi_hdiv_batch(nel, etype, norder, norient_edge, norient_face, maxp,                   &
                          etav, ncomp, fval, f_ld, mask, dof, dof_ld, info) bind(C)
         import
         integer(c_int), value :: nel, maxp, ncomp
         integer(c_long_long), value :: f_ld, dof_ld
         integer(c_int) :: etype(*), norder(19,*), norient_edge(12,*), norient_face(6,*), info(*)
         real(c_double) :: etav(3,8,*), fval(f_ld,*), dof(dof_ld,*)
         type(c_ptr), value :: mask
      end function
!
!  ...device-resident CLOC (stc.F90:45-58,273-277 with STORE_STC = .true.): the Schur factors stay in HBM under the element index
!     Iel of the subdomain; stc_bwd (stc.F90:661-677) runs on them after the global solve
      integer(c_int) function hp3d_gpu_cloc_create(plan, limit_bytes) bind(C)
         import; integer(c_int), value :: plan; integer(c_long_long), value :: limit_bytes
      end function
      integer(c_int) function hp3d_gpu_cloc_clear(cloc) bind(C)
         import; integer(c_int), value :: cloc
      end function
      integer(c_int) function hp3d_gpu_cloc_destroy(cloc) bind(C)
         import; integer(c_int), value :: cloc
      end function
      integer(c_int) function hp3d_gpu_cloc_stats(cloc, stats) bind(C)
         import; integer(c_int), value :: cloc; integer(c_long_long) :: stats(4)
      end function
      integer(c_int) function hp3d_gpu_elem_batch_cloc(plan, cloc, nel, iel, etype, norder, norient_edge, norient_face, xnod, &
                          xnod_ld, source_qp, source_ld, Aii, sAii, Bi, sBi, ni_out, nb_out, info) bind(C)
         import
         integer(c_int), value :: plan, cloc, nel, xnod_ld
         integer(c_long_long), value :: source_ld, sAii, sBi
         integer(c_long_long) :: iel(*)
         integer(c_int) :: etype(*), norder(19,*), norient_edge(12,*), norient_face(6,*), ni_out(*), nb_out(*), info(*)
         real(c_double) :: xnod(xnod_ld,*)
         type(c_ptr), value :: source_qp, Aii, Bi
      end function
      integer(c_int) function hp3d_gpu_celem_batch_cloc(plan, cloc, nel, iel, etype, norder, norient_edge, norient_face,      &
                          xnod, xnod_ld, source_qp, source_ld, mptr, cptr, cidx, cval, idbc, zdofd, xptr, nextract, lcon,     &
                          isym_flag, aptr, zbload, zastif, irn, jcn, ni_out, nb_out, info) bind(C)
         import
         integer(c_int), value :: plan, cloc, nel, xnod_ld, isym_flag
         integer(c_long_long), value :: source_ld
         integer(c_long_long) :: iel(*), mptr(*), cptr(*), xptr(*), aptr(*)
         integer(c_int) :: etype(*), norder(19,*), norient_edge(12,*), norient_face(6,*), cidx(*), idbc(*), nextract(*), lcon(*)
         integer(c_int) :: ni_out(*), nb_out(*), info(*)
         real(c_double) :: xnod(xnod_ld,*), cval(*)
         type(c_ptr), value :: source_qp, zdofd, zbload, zastif, irn, jcn
      end function
      integer(c_int) function hp3d_gpu_cloc_bwd_batch(cloc, nel, iel, xi, sxi, xb, sxb, nb_out, info) bind(C)
         import
         integer(c_int), value :: cloc, nel
         integer(c_long_long), value :: sxi, sxb
         integer(c_long_long) :: iel(*)
         integer(c_int) :: nb_out(*), info(*)
         type(c_ptr), value :: xi, xb
      end function
!
!  ...full Hermitian blocks from the packed lower triangles of hp3d_params%aii_packed = 1 (host only; ZTPTTR + conjugate mirror)
      integer(c_int) function hp3d_gpu_hermitian_unpack_batch(complex_mode, nel, ni, ni_e, AP, sAP, A, sA, threads) bind(C)
         import
         integer(c_int), value :: complex_mode, nel, ni, threads
         integer(c_long_long), value :: sAP, sA
         type(c_ptr), value :: ni_e, AP, A
      end function
   end interface
!
contains
!
!  ...print the library's message and stop, the way the reference reacts to LAPACK info /= 0 (stc.F90:371-374)
   subroutine hp3d_gpu_check(ierr, where)
      integer(c_int),   intent(in) :: ierr
      character(len=*), intent(in) :: where
      character(kind=c_char), pointer :: msg(:)
      integer :: n
      if (ierr .eq. 0) return
      call c_f_pointer(hp3d_gpu_last_error(), msg, [512])
      n = 1
      do while (n .lt. 512 .and. msg(n) .ne. c_null_char)
         n = n + 1
      enddo
      write(*,*) where, ': hp3d_gpu error ', ierr, ': ', msg(1:n-1)
      stop
   end subroutine hp3d_gpu_check
!
end module hp3d_gpu
