! hnumo_b200_iface.F90 -- ISO_C_BINDING interfaces of libhnumo_b200.so (include/hnumo_b200.h).
!
! New code for the h-NUMO driver (the reference has no bind(C) anywhere; its only Fortran->C precedent is the
! implicit-interface p4est glue, src/p4est.c:60-74, src/mod_p4est.F90:174).  Compile this file with the reference's
! own flags (-fdefault-real-8 makes `real` = real(c_double), config.user:8) and link with -lhnumo_b200.
! NOT compiled in the build container (no Fortran compiler there); kept in sync with the header by
! tests/test_host_cpu.py::test_fortran_interface_matches_header.
module hnumo_b200_iface
    use iso_c_binding
    implicit none

    integer(c_int32_t), parameter :: HNUMO_ABI_VERSION = 3

    ! mirrors hnumo_desc_t field by field
    type, bind(C) :: hnumo_desc_t
        integer(c_int32_t) :: abi_version
        integer(c_int32_t) :: nelem, ngl, nq, nlayers, nface
        integer(c_int32_t) :: kstages, N_btp
        real(c_double)     :: dt, dt_btp
        integer(c_int32_t) :: botfr, method_visc
        real(c_double)     :: gravity, cd_mlswe, visc_mlswe, ad_mlswe
        type(c_ptr) :: psiq, dpsiq, wnq, wgl, dpsi
        type(c_ptr) :: face
        type(c_ptr) :: elem_metrics
        type(c_ptr) :: face_geom
        type(c_ptr) :: pbprime_df, massinv, coriolis_df, tau_wind_df, zbot_df, alpha_mlswe
        type(c_ptr) :: ssprk_a, ssprk_beta
        integer(c_int32_t) :: rank, nranks, num_nbh
        type(c_ptr) :: nbh_proc, num_send_recv, nbh_send_recv
        integer(c_int32_t) :: device
        integer(c_int32_t) :: stage_kernel_variant
        real(c_double)     :: max_shear_dz
        type(c_ptr) :: point_metrics_q, point_metrics, face_geom_q, face_geom_n
        type(c_ptr) :: coord
    end type hnumo_desc_t

    interface
        integer(c_int) function hnumo_init(desc, handle) bind(C, name="hnumo_init")
            import :: c_int, c_ptr, hnumo_desc_t
            type(hnumo_desc_t), intent(in) :: desc
            type(c_ptr), intent(out) :: handle
        end function
        integer(c_int) function hnumo_device_count() bind(C, name="hnumo_device_count")
            import :: c_int
        end function
        integer(c_int) function hnumo_finalize(handle) bind(C, name="hnumo_finalize")
            import :: c_int, c_ptr
            type(c_ptr), value :: handle
        end function
        type(c_ptr) function hnumo_last_error() bind(C, name="hnumo_last_error")
            import :: c_ptr
        end function
        integer(c_int) function hnumo_upload_state(handle, q_df, qb_df, qprime_df) bind(C, name="hnumo_upload_state")
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: handle
            real(c_double), intent(in) :: q_df(*), qb_df(*), qprime_df(*)
        end function
        integer(c_int) function hnumo_download_state(handle, q_df, qb_df, qprime_df) bind(C, name="hnumo_download_state")
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: handle
            real(c_double), intent(inout) :: q_df(*), qb_df(*), qprime_df(*)
        end function
        integer(c_int) function hnumo_step(handle, nsteps) bind(C, name="hnumo_step")
            import :: c_int, c_ptr, c_int32_t
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: nsteps
        end function
        integer(c_int) function hnumo_ti_rk_bcl(handle, q_df, qb_df, qprime_df) bind(C, name="hnumo_ti_rk_bcl")
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: handle
            real(c_double), intent(inout) :: q_df(*), qb_df(*), qprime_df(*)
        end function
        integer(c_int) function hnumo_comm_get_unique_id(id128) bind(C, name="hnumo_comm_get_unique_id")
            import :: c_int, c_char
            character(kind=c_char), intent(out) :: id128(128)
        end function
        integer(c_int) function hnumo_comm_init(handle, id128) bind(C, name="hnumo_comm_init")
            import :: c_int, c_ptr, c_char
            type(c_ptr), value :: handle
            character(kind=c_char), intent(in) :: id128(128)
        end function
        ! device-side diagnostics: replaces diagnostics / compute_conserved / the max-min and Courant part of
        ! print_diagnostics_mlswe for the resident state (out: 11*nlayers + 12 values, see include/hnumo_b200.h)
        integer(c_int64_t) function hnumo_diagnostics(handle, out, capacity) bind(C, name="hnumo_diagnostics")
            import :: c_int64_t, c_ptr, c_double
            type(c_ptr), value :: handle
            real(c_double), intent(out) :: out(*)
            integer(c_int64_t), value :: capacity
        end function
        integer(c_int) function hnumo_timing(handle, out8, reset) bind(C, name="hnumo_timing")
            import :: c_int, c_ptr, c_double, c_int32_t
            type(c_ptr), value :: handle
            real(c_double), intent(out) :: out8(8)
            integer(c_int32_t), value :: reset
        end function
    end interface
end module hnumo_b200_iface
