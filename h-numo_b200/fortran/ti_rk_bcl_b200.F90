! ti_rk_bcl_b200.F90 -- drop-in replacement of the reference's src/ti_rk_bcl.F90 (same external subroutine name and
! dummy arguments, src/ti_rk_bcl.F90:9,32-34) that runs the step on the B200 through libhnumo_b200.so.
!
! Build: replace ti_rk_bcl.o by this file + hnumo_b200_iface.F90 in src/Makefile and add -lhnumo_b200 -lcudart to the
! link line.  numo3d.in, the p4est mesh, the initial conditions and all output stay untouched.
! NOT compiled in the build container (no Fortran compiler / MPI / p4est there) -- see INTEGRATION.md.
!
! Two modes:
!   default      : upload -> one step -> download every call: bit-compatible drop-in (state lives on the host between
!                  calls exactly as in the reference), costs 2 x 8 B x (6 nlayers + 4) npoin of PCIe traffic per step;
!   HNUMO_RESIDENT (cpp macro): the state stays on the GPU; the caller (mod_time_loop.F90:219, before diagnostics /
!                  restart output) calls hnumo_b200_sync_host(q_df, qb_df, qprime_df) when it needs host copies.
module hnumo_b200_state
    use iso_c_binding
    use hnumo_b200_iface
    implicit none
    type(c_ptr), save :: handle = c_null_ptr
    logical, save :: resident_valid = .false.
    real(c_double), allocatable, target, save :: elem_metrics(:,:), face_geom(:,:), ssprk_a_c(:,:)
    ! general quadrilaterals (gmsh meshes, curved elements): the geometry per point (include/hnumo_b200.h, ABI 3)
    logical, save :: general_quads = .false.
    real(c_double), allocatable, target, save :: point_metrics_q(:,:), point_metrics(:,:), face_geom_q(:,:,:), face_geom_n(:,:,:), coord_c(:,:)
    integer(c_int32_t), allocatable, target, save :: face_c(:,:), nbh_proc_c(:), num_send_recv_c(:), nbh_send_recv_c(:)
contains

    subroutine hnumo_b200_setup()
        ! everything ti_rk_bcl reads through `use` (src/ti_rk_bcl.F90:19-28 and the modules below it)
        use mpi
        use mod_grid, only: nelem, npoin, npoin_q, nface, face, face_type, coord
        use mod_basis, only: ngl, nq, psiq, dpsiq, wnq, wgl, dpsi
        use mod_input, only: nlayers, kstages, dt, dt_btp, botfr, cd_mlswe, method_visc, visc_mlswe, ad_mlswe, max_shear_dz
        use mod_constants, only: gravity
        use mod_metrics, only: ksiq_x, ksiq_y, etaq_x, etaq_y, jacq, ksi_x, ksi_y, eta_x, eta_y, jac, massinv
        use mod_face, only: normal_vector_q, jac_faceq, normal_vector, jac_face
        use mod_initial, only: pbprime_df, coriolis_df, tau_wind_df, zbot_df, alpha_mlswe, ssprk_a, ssprk_beta, &
                               N_btp
        use mod_parallel, only: num_nbh, nbh_proc, num_send_recv, nbh_send_recv
        use mod_mpi_utilities, only: irank, numproc
        integer :: e, f, i, j, iq, n, ip, ierr, ntot, ndev, local_rank, node_comm
        integer(c_int) :: rc
        real(c_double) :: tol, ref
        character(kind=c_char) :: id(128)

        ! Affine meshes (bricks, parallelograms: one metric set per element, one normal / edge Jacobian per face) run the optimised
        ! kernels on elem_metrics / face_geom.  Anything else with conforming faces -- curved or gmsh quads -- is handed over
        ! point by point (point_metrics_q ... coord) and runs the library's run-time-size kernels.  AMR meshes have
        ! non-conforming faces (face_type 21/12, src/mod_grid.F90:79): refused.
        general_quads = .false.
        do f = 1, nface
            if (face_type(f) /= 1 .and. face_type(f) /= 2) stop "hnumo_b200: non-conforming or unknown face_type (only 1, 2)"
            do iq = 2, nq
                if (maxval(abs(normal_vector_q(1:2,iq,1,f) - normal_vector_q(1:2,1,1,f))) > 1.0e-12) general_quads = .true.
                if (abs(jac_faceq(iq,1,f)/wnq(iq) - jac_faceq(1,1,f)/wnq(1)) > 1.0e-10*abs(jac_faceq(1,1,f)/wnq(1))) general_quads = .true.
            end do
        end do
        do e = 1, nelem
            ref = abs(ksiq_x(1,1,1,e)) + abs(ksiq_y(1,1,1,e)) + abs(etaq_x(1,1,1,e)) + abs(etaq_y(1,1,1,e))
            tol = 1.0e-10*ref
            do j = 1, nq
                do i = 1, nq
                    if (abs(ksiq_x(i,j,1,e) - ksiq_x(1,1,1,e)) > tol .or. abs(ksiq_y(i,j,1,e) - ksiq_y(1,1,1,e)) > tol .or. &
                        abs(etaq_x(i,j,1,e) - etaq_x(1,1,1,e)) > tol .or. abs(etaq_y(i,j,1,e) - etaq_y(1,1,1,e)) > tol) &
                        general_quads = .true.
                end do
            end do
        end do
        if (general_quads) then
            allocate(point_metrics_q(5,npoin_q), point_metrics(5,npoin), face_geom_q(3,nq,nface), face_geom_n(3,ngl,nface), coord_c(2,npoin))
            do e = 1, nelem
                do j = 1, nq
                    do i = 1, nq
                        ip = (e-1)*nq*nq + (j-1)*nq + i          ! intma_dg_quad, src/mod_grid.F90:242-249
                        point_metrics_q(:,ip) = (/ ksiq_x(i,j,1,e), ksiq_y(i,j,1,e), etaq_x(i,j,1,e), etaq_y(i,j,1,e), jacq(i,j,1,e) /)
                    end do
                end do
                do j = 1, ngl
                    do i = 1, ngl
                        ip = (e-1)*ngl*ngl + (j-1)*ngl + i       ! intma_dg, src/mod_grid.F90:230-237
                        point_metrics(:,ip) = (/ ksi_x(i,j,1,e), ksi_y(i,j,1,e), eta_x(i,j,1,e), eta_y(i,j,1,e), jac(i,j,1,e) /)
                    end do
                end do
            end do
            do f = 1, nface
                do iq = 1, nq
                    face_geom_q(1:2,iq,f) = normal_vector_q(1:2,iq,1,f); face_geom_q(3,iq,f) = jac_faceq(iq,1,f)
                end do
                do n = 1, ngl
                    face_geom_n(1:2,n,f) = normal_vector(1:2,n,1,f); face_geom_n(3,n,f) = jac_face(n,1,f)
                end do
            end do
            coord_c(1:2,1:npoin) = coord(1:2,1:npoin)
        end if

        allocate(elem_metrics(5,nelem), face_geom(3,nface), face_c(8,nface), ssprk_a_c(kstages,3))
        do e = 1, nelem      ! bricks are affine: one metric set per element (include/hnumo_b200.h, elem_metrics)
            elem_metrics(1,e) = ksiq_x(1,1,1,e); elem_metrics(2,e) = ksiq_y(1,1,1,e)
            elem_metrics(3,e) = etaq_x(1,1,1,e); elem_metrics(4,e) = etaq_y(1,1,1,e)
            elem_metrics(5,e) = jacq(1,1,1,e) / (wnq(1)*wnq(1))
        end do
        do f = 1, nface
            face_geom(1:2,f) = normal_vector_q(1:2,1,1,f)
            face_geom(3,f)   = jac_faceq(1,1,f) / wnq(1)
            face_c(:,f) = int(face(:,f), c_int32_t)
        end do
        ssprk_a_c = ssprk_a(1:kstages,1:3)
        ntot = 0
        if (num_nbh > 0) ntot = sum(num_send_recv(1:num_nbh))
        allocate(nbh_proc_c(max(num_nbh,1)), num_send_recv_c(max(num_nbh,1)), nbh_send_recv_c(max(ntot,1)))
        if (num_nbh > 0) then
            nbh_proc_c(1:num_nbh) = nbh_proc(1:num_nbh); num_send_recv_c(1:num_nbh) = num_send_recv(1:num_nbh)
            nbh_send_recv_c(1:ntot) = nbh_send_recv(1:ntot)
        end if

        ! GPU of this rank: node-local rank modulo the number of devices of the node
        call mpi_comm_split_type(mpi_comm_world, MPI_COMM_TYPE_SHARED, 0, MPI_INFO_NULL, node_comm, ierr)
        call mpi_comm_rank(node_comm, local_rank, ierr)
        call mpi_comm_free(node_comm, ierr)
        ndev = hnumo_device_count()
        if (ndev < 1) stop "hnumo_b200: no CUDA device on this node (the library has no CPU fallback)"

        ! The reference's module arrays are plain allocatables (no TARGET: src/mod_basis.F90:53-55, src/mod_metrics.F90:40-42,
        ! src/mod_initial.F90:64-79), so c_loc cannot be applied to them here.  They are passed to a contained procedure whose
        ! dummies are TARGET assumed-size arrays (sequence association); the addresses are taken there and are used only
        ! during that call -- hnumo_init copies everything it needs and retains no host pointer.
        call init_with(psiq, dpsiq, wnq, wgl, dpsi, pbprime_df, massinv, coriolis_df, tau_wind_df, zbot_df, alpha_mlswe, &
                       ssprk_beta, rc)
        if (rc /= 0) stop "hnumo_init failed"
        if (numproc > 1) then               ! replaces mod_mpi_communicator_create: NCCL id broadcast over MPI
            if (irank == 0) then
                if (hnumo_comm_get_unique_id(id) /= 0) stop "hnumo_comm_get_unique_id failed"
            end if
            call mpi_bcast(id, 128, MPI_CHARACTER, 0, mpi_comm_world, ierr)
            if (hnumo_comm_init(handle, id) /= 0) stop "hnumo_comm_init failed"
        end if

    contains

        subroutine init_with(psiq_a, dpsiq_a, wnq_a, wgl_a, dpsi_a, pbprime_a, massinv_a, coriolis_a, tauw_a, zbot_a, &
                             alpha_a, beta_a, rc_out)
            real(c_double), target, intent(in) :: psiq_a(*), dpsiq_a(*), wnq_a(*), wgl_a(*), dpsi_a(*), pbprime_a(*), &
                                                  massinv_a(*), coriolis_a(*), tauw_a(*), zbot_a(*), alpha_a(*), beta_a(*)
            integer(c_int), intent(out) :: rc_out
            type(hnumo_desc_t) :: d
            d%abi_version = HNUMO_ABI_VERSION
            d%nelem = nelem; d%ngl = ngl; d%nq = nq; d%nlayers = nlayers; d%nface = nface
            d%kstages = kstages; d%N_btp = N_btp; d%dt = dt; d%dt_btp = dt_btp
            d%botfr = botfr; d%method_visc = method_visc
            d%gravity = gravity; d%cd_mlswe = cd_mlswe; d%visc_mlswe = visc_mlswe; d%ad_mlswe = ad_mlswe
            d%psiq = c_loc(psiq_a); d%dpsiq = c_loc(dpsiq_a); d%wnq = c_loc(wnq_a); d%wgl = c_loc(wgl_a); d%dpsi = c_loc(dpsi_a)
            d%face = c_loc(face_c); d%elem_metrics = c_loc(elem_metrics); d%face_geom = c_loc(face_geom)
            d%point_metrics_q = c_null_ptr; d%point_metrics = c_null_ptr; d%face_geom_q = c_null_ptr; d%face_geom_n = c_null_ptr
            d%coord = c_null_ptr
            if (general_quads) then
                d%point_metrics_q = c_loc(point_metrics_q); d%point_metrics = c_loc(point_metrics)
                d%face_geom_q = c_loc(face_geom_q); d%face_geom_n = c_loc(face_geom_n); d%coord = c_loc(coord_c)
            end if
            d%pbprime_df = c_loc(pbprime_a); d%massinv = c_loc(massinv_a); d%coriolis_df = c_loc(coriolis_a)
            d%tau_wind_df = c_loc(tauw_a); d%zbot_df = c_loc(zbot_a); d%alpha_mlswe = c_loc(alpha_a)
            d%ssprk_a = c_loc(ssprk_a_c); d%ssprk_beta = c_loc(beta_a)
            d%rank = irank; d%nranks = numproc; d%num_nbh = num_nbh
            d%nbh_proc = c_null_ptr; d%num_send_recv = c_null_ptr; d%nbh_send_recv = c_null_ptr
            if (num_nbh > 0) then
                d%nbh_proc = c_loc(nbh_proc_c); d%num_send_recv = c_loc(num_send_recv_c); d%nbh_send_recv = c_loc(nbh_send_recv_c)
            end if
            d%device = mod(local_rank, ndev) + 1
            d%stage_kernel_variant = 0
            d%max_shear_dz = max_shear_dz
            rc_out = hnumo_init(d, handle)
        end subroutine init_with
    end subroutine hnumo_b200_setup

    subroutine hnumo_b200_sync_host(q_df, qb_df, qprime_df)
        real(c_double), intent(inout) :: q_df(*), qb_df(*), qprime_df(*)
        if (hnumo_download_state(handle, q_df, qb_df, qprime_df) /= 0) stop "hnumo_download_state failed"
    end subroutine hnumo_b200_sync_host

    ! Resident mode: the numbers print_diagnostics_mlswe prints (src/print_diagnostics.F90:59-128), reduced on the device and
    ! then over the ranks exactly as the reference reduces them (mass: sum; max/min; Courant numbers: max; dx, dy: min).
    ! No download of the state is needed on steps that only print this table (mod_time_loop.F90:184,251).
    subroutine hnumo_b200_diagnostics(mass_conserv, qmax_layers, qmin_layers, qbmax, qbmin, cfl_vector, min_dx_vec)
        use mpi
        use mod_input, only: nlayers
        real(c_double), intent(out) :: mass_conserv(nlayers), qmax_layers(5,nlayers), qmin_layers(5,nlayers)
        real(c_double), intent(out) :: qbmax(4), qbmin(4), cfl_vector(2), min_dx_vec(2)
        real(c_double) :: d(11*nlayers + 12), g(11*nlayers + 12)
        integer :: k, ierr, n
        n = 11*nlayers + 12
        if (hnumo_diagnostics(handle, d, int(n, c_int64_t)) /= n) stop "hnumo_diagnostics failed"
        do k = 1, nlayers
            call mpi_reduce(d(11*(k-1)+1), g(11*(k-1)+1), 1, MPI_DOUBLE_PRECISION, mpi_sum, 0, mpi_comm_world, ierr)
            call mpi_reduce(d(11*(k-1)+2), g(11*(k-1)+2), 5, MPI_DOUBLE_PRECISION, mpi_max, 0, mpi_comm_world, ierr)
            call mpi_reduce(d(11*(k-1)+7), g(11*(k-1)+7), 5, MPI_DOUBLE_PRECISION, mpi_min, 0, mpi_comm_world, ierr)
            mass_conserv(k) = g(11*(k-1)+1)
            qmax_layers(:,k) = g(11*(k-1)+2 : 11*(k-1)+6)     ! h, u, v, dp, interface elevation (diagnostics.F90:24-45)
            qmin_layers(:,k) = g(11*(k-1)+7 : 11*(k-1)+11)
        end do
        k = 11*nlayers
        call mpi_reduce(d(k+1), g(k+1), 4, MPI_DOUBLE_PRECISION, mpi_max, 0, mpi_comm_world, ierr)
        call mpi_reduce(d(k+5), g(k+5), 4, MPI_DOUBLE_PRECISION, mpi_min, 0, mpi_comm_world, ierr)
        call mpi_reduce(d(k+9), g(k+9), 2, MPI_DOUBLE_PRECISION, mpi_max, 0, mpi_comm_world, ierr)
        call mpi_reduce(d(k+11), g(k+11), 2, MPI_DOUBLE_PRECISION, mpi_min, 0, mpi_comm_world, ierr)
        qbmax = g(k+1:k+4); qbmin = g(k+5:k+8); cfl_vector = g(k+9:k+10); min_dx_vec = g(k+11:k+12)
    end subroutine hnumo_b200_diagnostics
end module hnumo_b200_state

subroutine ti_rk_bcl(q_df, qb_df, qprime_df)
    use iso_c_binding
    use hnumo_b200_iface
    use hnumo_b200_state
    use mod_grid, only: npoin
    use mod_input, only: nlayers
    implicit none
    real, dimension(4,npoin), intent(inout) :: qb_df
    real, dimension(3,npoin,nlayers), intent(inout) :: q_df
    real, dimension(3,npoin,nlayers), intent(inout) :: qprime_df
    integer(c_int) :: rc

    if (.not. c_associated(handle)) call hnumo_b200_setup()
#ifdef HNUMO_RESIDENT
    if (.not. resident_valid) then
        if (hnumo_upload_state(handle, q_df, qb_df, qprime_df) /= 0) stop "hnumo_upload_state failed"
        resident_valid = .true.
    end if
    rc = hnumo_step(handle, 1_c_int32_t)
#else
    rc = hnumo_ti_rk_bcl(handle, q_df, qb_df, qprime_df)
#endif
    ! same failure behaviour as the reference (src/mod_splitting.F90:74-77,228-231)
    if (rc == 1) stop "Negative mass in thickness at some points"
    if (rc /= 0) stop "hnumo_b200: CUDA/NCCL error in ti_rk_bcl"
end subroutine ti_rk_bcl
