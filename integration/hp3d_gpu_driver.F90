!-----------------------------------------------------------------------------------------------------------------------
!  hp3d_gpu_driver.F90 -- the Fortran side of the drop-in: what a maintainer adds next to hp3d_gpu_mod.F90
!  (trunk/src/modules/hp3d_gpu_driver.F90).  NOT compiled in this repository (no Fortran compiler in the build image);
!  tests/test_c_harness.py replays the same call sequence against the raw .so from C and checks the struct layouts.
!
!  One batched call per subdomain replaces the pair  elem + stc_fwd_wrapper  of every element
!  (src/constrs/celem_systemI.F90:523,534) inside the element loop of the solver interfaces
!  (src/solver/par_mumps/par_mumps_sc.F90:318-357, mumps_sc, pardiso_sc, par_nested, petsc_solve):
!
!     call hp3d_gpu_condense_subdomain(ierr)            before the !$OMP PARALLEL of STEP 2   (par_mumps_sc.F90:318)
!     ... unchanged element loop ...  celem_systemI:
!           if (HP3D_GPU_ON) then
!              call hp3d_gpu_scatter_to_aloc(iel)        instead of  call elem(...) ; call stc_fwd_wrapper(...)
!           else ...
!     call hp3d_gpu_stc_bwd_subdomain(...)               instead of the stc_bwd_wrapper loop (stc.F90:529-677) after the solve
!
!  STORE_STC = .true. (stc.F90:273-277): the Schur factors CLOC(iel)%ASchur / %BSchur are not copied to the host at all; they
!  stay in device memory under the index iel (hp3d_gpu_cloc_*), where the back-substitution reads them.
!-----------------------------------------------------------------------------------------------------------------------
#include "typedefs.h"
module hp3d_gpu_driver
   use, intrinsic :: iso_c_binding
   use hp3d_gpu
   use parameters , only: MAXbrickH, NR_RHS
   use physics    , only: NR_PHYSA
   use data_structure3D, only: NODES, NRELES_SUBD, ELEM_SUBD          ! src/modules/data_structure3D.F90:26,420-432
   use assembly   , only: ALOC, BLOC
   use stc        , only: stc_get_nrdof
   use mpi_param  , only: RANK
   implicit none
!
   logical :: HP3D_GPU_ON = .false.
   integer(c_int), save :: GPU_PLAN = -1, GPU_CLOC = -1, GPU_KIND = 0
!
!..per-subdomain descriptors and results (allocated by hp3d_gpu_condense_subdomain)
   integer(c_int), allocatable, target, save :: g_etype(:), g_norder(:,:), g_nedge(:,:), g_nface(:,:)
   integer(c_int), allocatable, target, save :: g_ni(:), g_nb(:), g_info(:)
   integer(c_long_long), allocatable, target, save :: g_iel(:)
   real(c_double), allocatable, target, save :: g_xnod(:,:,:)
!..results live in page-locked memory (hp3d_gpu_host_alloc) so that the copies overlap the kernels;
!  value type of the LIBRARY: real(8) for the Poisson kinds, complex(8) for the Maxwell kinds -- independent of HP3D_COMPLEX
   type(c_ptr), save :: p_Aii = c_null_ptr, p_Bi = c_null_ptr
   real(c_double),    pointer, save :: r_Aii(:,:), r_Bi(:,:)
   complex(c_double_complex), pointer, save :: z_Aii(:,:), z_Bi(:,:)
   integer(c_long_long), save :: sAii = 0, sBi = 0
   integer, save :: ni_max = 0
!
contains
!
!-----------------------------------------------------------------------------------------------------------------------
!  once per run: bind MPI rank r to GPU mod(r, ngpu) (hp3D runs one rank per GPU: the distr_mesh partition is the GPU
!  partition, par_mesh.F90:39-82) and fix the problem.  Kind = HP3D_POIS_GAL ... HP3D_MAXW_UW; prm from the problem's
!  parameter modules (OMEGA, EPSILON, MU, SIGMA, ALPHA_NORM, TEST_NORM, NORD_ADD, MAXP, ICOMP_EXACT).
   subroutine hp3d_gpu_setup(Kind, Prm, Ngpu)
      integer(c_int)   , intent(in)    :: Kind, Ngpu
      type(hp3d_params), intent(inout) :: Prm
      call hp3d_gpu_check(hp3d_gpu_init(int(mod(RANK, Ngpu), c_int)), 'hp3d_gpu_setup: init')
      Prm%store_schur = 1          ! STORE_STC: the factors are formed; they stay on the device (cloc) unless fetched
      Prm%aii_packed  = 0          ! full Aii blocks (1: packed lower triangle, 2: half over PCIe + host mirror by the library)
      Prm%nr_rhs      = int(NR_RHS, c_int)   ! > 1: DPG problems with HP3D_SRC_TABLE sources (one table per load)
      GPU_PLAN = hp3d_gpu_plan(Kind, Prm)
      if (GPU_PLAN .lt. 0) call hp3d_gpu_check(GPU_PLAN, 'hp3d_gpu_setup: plan')
      GPU_CLOC = hp3d_gpu_cloc_create(GPU_PLAN, 0_c_long_long)     ! 0: up to 60 % of the free device memory; the rest spills to recompute
      if (GPU_CLOC .lt. 0) call hp3d_gpu_check(GPU_CLOC, 'hp3d_gpu_setup: cloc')
      GPU_KIND = Kind
      HP3D_GPU_ON = .true.
   end subroutine hp3d_gpu_setup
!
!-----------------------------------------------------------------------------------------------------------------------
!  elem + stc_fwd_wrapper for all elements of the subdomain (one call); results wait in g_Aii / g_Bi for celem_systemI
   subroutine hp3d_gpu_condense_subdomain(Ierr)
      integer, intent(out) :: Ierr
      integer :: iel, mdle, n1, n2, n3, n4
      integer(c_int) :: norder(19), ni, nb, nint, nrdofH
      integer(c_long_long) :: nbytes, es
!
      Ierr = 0
      if (allocated(g_etype)) deallocate(g_etype, g_norder, g_nedge, g_nface, g_ni, g_nb, g_info, g_iel, g_xnod)
      allocate(g_etype(NRELES_SUBD), g_norder(19,NRELES_SUBD), g_nedge(12,NRELES_SUBD), g_nface(6,NRELES_SUBD))
      allocate(g_ni(NRELES_SUBD), g_nb(NRELES_SUBD), g_info(NRELES_SUBD), g_iel(NRELES_SUBD))
      allocate(g_xnod(3,MAXbrickH,NRELES_SUBD))
      g_norder = 0; g_nedge = 0; g_nface = 0; g_xnod = 0.d0
!
!  ...the descriptors the element routine itself would compute (elem_opt.F90:173-206); independent per element
!$OMP PARALLEL DO PRIVATE(mdle) SCHEDULE(DYNAMIC)
      do iel = 1,NRELES_SUBD
         mdle = ELEM_SUBD(iel)
         g_etype(iel) = NODES(mdle)%ntype                               ! MDLB = 1 / MDLP = 3 pass through unchanged
         call find_order (mdle, g_norder(:,iel))                         ! src/datstrs/find_order.F90:5
         call find_orient(mdle, g_nedge(:,iel), g_nface(:,iel))          ! src/datstrs/find_orient.F90:8
         call nodcor     (mdle, g_xnod(:,:,iel))                         ! src/constrs/nodcor.F90:19
         g_iel(iel) = int(iel, c_long_long)                              ! CLOC index (stc.F90:45-58)
      enddo
!$OMP END PARALLEL DO
!
!  ...result strides: the largest element of the subdomain (host-only size queries; safe from any thread)
      ni_max = 0
      do iel = 1,NRELES_SUBD
         call hp3d_gpu_check(hp3d_gpu_sizes_t(GPU_PLAN, g_etype(iel), g_norder(:,iel), ni, nb, nint, nrdofH), 'sizes')
         ni_max = max(ni_max, int(ni))
      enddo
      sAii = int(ni_max, c_long_long)**2; sBi = int(ni_max, c_long_long)*NR_RHS      ! Bi(ni,NR_RHS) per element
      es = 8_c_long_long; if (GPU_KIND .ge. HP3D_MAXW_GAL) es = 16_c_long_long
      if (c_associated(p_Aii)) then; call hp3d_gpu_host_free(p_Aii); call hp3d_gpu_host_free(p_Bi); endif
      nbytes = es*sAii*NRELES_SUBD; p_Aii = hp3d_gpu_host_alloc(max(nbytes, 8_c_long_long))
      nbytes = es*sBi *NRELES_SUBD; p_Bi  = hp3d_gpu_host_alloc(max(nbytes, 8_c_long_long))
      if (.not. c_associated(p_Aii) .or. .not. c_associated(p_Bi)) then; Ierr = -3; return; endif
      if (GPU_KIND .ge. HP3D_MAXW_GAL) then
         call c_f_pointer(p_Aii, z_Aii, [int(sAii), NRELES_SUBD]); call c_f_pointer(p_Bi, z_Bi, [int(sBi), NRELES_SUBD])
      else
         call c_f_pointer(p_Aii, r_Aii, [int(sAii), NRELES_SUBD]); call c_f_pointer(p_Bi, r_Bi, [int(sBi), NRELES_SUBD])
      endif
!
!  ...one batched call: Aii / Bi to the host, ASchur / BSchur into the device-resident store under g_iel
      Ierr = hp3d_gpu_elem_batch_cloc(GPU_PLAN, GPU_CLOC, int(NRELES_SUBD, c_int), g_iel, g_etype, g_norder, g_nedge, g_nface,  &
                                      g_xnod, int(3*MAXbrickH, c_int), c_null_ptr, 0_c_long_long, p_Aii, sAii, p_Bi, sBi,       &
                                      g_ni, g_nb, g_info)
      call hp3d_gpu_check(int(Ierr, c_int), 'hp3d_gpu_condense_subdomain')
!  ...reference behaviour for LAPACK info /= 0 (stc.F90:371-374) and non-positive Jacobians (geom3D.F90:92-109): print + stop
      if (any(g_info(1:NRELES_SUBD) .ne. 0)) then
         iel = maxloc(abs(g_info(1:NRELES_SUBD)), 1)
         write(*,*) 'hp3d_gpu_condense_subdomain: Mdle,info = ', ELEM_SUBD(iel), g_info(iel)
         stop
      endif
   end subroutine hp3d_gpu_condense_subdomain
!
!-----------------------------------------------------------------------------------------------------------------------
!  ALOC / BLOC <- the condensed blocks of element Iel, exactly what stc_fwd_wrapper leaves there (stc.F90:287-305):
!  variable i owns nrdofi(i) consecutive rows of Aii (stc_get_nrdof, stc.F90:94), in variable order; the bubble rows are gone.
!  HP3D_COMPLEX builds of the REAL problems (typedefs.h:2-6: VTYPE = complex(8) for Poisson too): the library's real(8) results
!  are widened on assignment.  Called from celem_systemI inside the unchanged OpenMP element loop (thread safe: reads only).
   subroutine hp3d_gpu_scatter_to_aloc(Iel)
      integer, intent(in) :: Iel
      integer :: nrdofi(NR_PHYSA), nrdofb(NR_PHYSA)
      integer :: i, j, ii, ji, ki, kj, ni, r, c, q
!
      call stc_get_nrdof(ELEM_SUBD(Iel), nrdofi, nrdofb)
      ni = sum(nrdofi(1:NR_PHYSA))
      if (ni .ne. g_ni(Iel)) then
         write(*,*) 'hp3d_gpu_scatter_to_aloc: ni mismatch ', ni, g_ni(Iel); stop
      endif
      kj = 0
      do j = 1,NR_PHYSA
         ji = nrdofi(j)
         ki = 0
         do i = 1,NR_PHYSA
            ii = nrdofi(i)
            if (ii .gt. 0 .and. ji .gt. 0) then
               do c = 1,ji
                  do r = 1,ii
!                 ...Aii is column-major with leading dimension ni (NOT ni_max): entry (ki+r, kj+c)
                     if (GPU_KIND .ge. HP3D_MAXW_GAL) then
                        ALOC(i,j)%array(r,c) = z_Aii((ki+r) + (kj+c-1)*ni, Iel)
                     else
                        ALOC(i,j)%array(r,c) = r_Aii((ki+r) + (kj+c-1)*ni, Iel)
                     endif
                  enddo
               enddo
            endif
            ki = ki + ii
         enddo
         if (ji .gt. 0) then
!        ...Bi(ni,NR_RHS), column-major with leading dimension ni: load q of dof kj+r at (q-1)*ni + kj + r
            do q = 1,NR_RHS
               if (GPU_KIND .ge. HP3D_MAXW_GAL) then
                  BLOC(j)%array(1:ji,q) = z_Bi((q-1)*ni+kj+1:(q-1)*ni+kj+ji, Iel)
               else
                  BLOC(j)%array(1:ji,q) = r_Bi((q-1)*ni+kj+1:(q-1)*ni+kj+ji, Iel)
               endif
            enddo
         endif
         kj = kj + ji
      enddo
   end subroutine hp3d_gpu_scatter_to_aloc
!
!-----------------------------------------------------------------------------------------------------------------------
!  stc_bwd for the whole subdomain (stc_bwd_wrapper, stc.F90:529-677): Xi(1:ni,iel) = interface solution of element iel in the
!  row order of Aii (what solout gathers, src/solver/frontal/interf/solout.F90), Xb(1:nb,iel) = BSchur - ASchur * xi
!  (NR_RHS > 1: NR_RHS columns per element, column q at offset (q-1)*ni resp. (q-1)*nb; Ldxi >= ni*NR_RHS, Ldxb >= nb*NR_RHS).
!  Elements whose factors did not fit into the store were spilled at condensation time and are recomputed here -- same result.
   subroutine hp3d_gpu_stc_bwd_subdomain(Xi, Ldxi, Xb, Ldxb, Ierr)
      integer, intent(in)  :: Ldxi, Ldxb
      VTYPE, target, intent(in)  :: Xi(Ldxi, NRELES_SUBD)
      VTYPE, target, intent(out) :: Xb(Ldxb, NRELES_SUBD)
      integer, intent(out) :: Ierr
      real(c_double), allocatable, target :: xr(:,:), yr(:,:)
!
#if HP3D_COMPLEX
      if (GPU_KIND .lt. HP3D_MAXW_GAL) then
!     ...complex build of a real problem: the library works on the real parts (the imaginary parts are zero by construction)
         allocate(xr(Ldxi,NRELES_SUBD), yr(Ldxb,NRELES_SUBD)); xr = real(Xi, c_double)
         Ierr = hp3d_gpu_cloc_bwd_batch(GPU_CLOC, int(NRELES_SUBD,c_int), g_iel, c_loc(xr), int(Ldxi,c_long_long), c_loc(yr),    &
                                        int(Ldxb,c_long_long), g_nb, g_info)
         Xb = yr
         deallocate(xr, yr)
      else
         Ierr = hp3d_gpu_cloc_bwd_batch(GPU_CLOC, int(NRELES_SUBD,c_int), g_iel, c_loc(Xi), int(Ldxi,c_long_long), c_loc(Xb),    &
                                        int(Ldxb,c_long_long), g_nb, g_info)
      endif
#else
      Ierr = hp3d_gpu_cloc_bwd_batch(GPU_CLOC, int(NRELES_SUBD,c_int), g_iel, c_loc(Xi), int(Ldxi,c_long_long), c_loc(Xb),       &
                                     int(Ldxb,c_long_long), g_nb, g_info)
#endif
      call hp3d_gpu_check(int(Ierr, c_int), 'hp3d_gpu_stc_bwd_subdomain')
   end subroutine hp3d_gpu_stc_bwd_subdomain
!
!-----------------------------------------------------------------------------------------------------------------------
!  after a mesh refinement the element indices change: forget the stored factors (stc_dealloc, stc.F90:68-70)
   subroutine hp3d_gpu_stc_dealloc()
      if (GPU_CLOC .ge. 0) call hp3d_gpu_check(hp3d_gpu_cloc_clear(GPU_CLOC), 'hp3d_gpu_stc_dealloc')
   end subroutine hp3d_gpu_stc_dealloc
!
   subroutine hp3d_gpu_shutdown()
      if (c_associated(p_Aii)) then; call hp3d_gpu_host_free(p_Aii); call hp3d_gpu_host_free(p_Bi); endif
      p_Aii = c_null_ptr; p_Bi = c_null_ptr
      if (HP3D_GPU_ON) call hp3d_gpu_check(hp3d_gpu_finalize(), 'hp3d_gpu_shutdown')
      HP3D_GPU_ON = .false.; GPU_PLAN = -1; GPU_CLOC = -1
   end subroutine hp3d_gpu_shutdown
!
end module hp3d_gpu_driver
