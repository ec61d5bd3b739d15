/* ============================================================================
 * hnumo_b200.h -- C-ABI of the B200-native h-NUMO hot path (libhnumo_b200.so).
 *
 * The library replaces exactly one call of the reference Fortran program:
 *
 *     call ti_rk_bcl(q0_df_mlswe, qb0_df_mlswe, qprime0_df)     src/mod_time_loop.F90:209
 *
 * i.e. one baroclinic predictor-corrector step of the multilayer shallow-water
 * equations including both barotropic SSPRK substep loops (src/ti_rk_bcl.F90:9-87,
 * src/mod_rk_mlswe.F90:19-151 and everything they call).  The Fortran driver keeps
 * reading numo3d.in, building the p4est mesh, metrics and initial conditions, and
 * hands the static arrays over once (hnumo_init).  State stays resident on the GPU
 * between steps.  The reference-side binding (ISO_C_BINDING) is in INTEGRATION.md and
 * h-numo_b200/fortran/.
 *
 * Conventions
 *   - plain C, no torch / CUDA types; all pointers are HOST pointers unless noted
 *   - arrays are column-major with the reference's shapes; integer tables keep the
 *     reference's 1-based element numbers (converted at init)
 *   - the library copies what it needs during the call and never retains host pointers
 *   - return value: 0 ok; <0 CUDA/NCCL/usage error (hnumo_last_error gives text);
 *     >0 physics error: 1 = negative layer thickness (reference `stop`,
 *     src/mod_splitting.F90:74-77,228-231)
 *   - one host thread per handle, calls in program order (the reference is one MPI rank
 *     = one thread = one GPU)
 * ========================================================================== */
#ifndef HNUMO_B200_H
#define HNUMO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HNUMO_ABI_VERSION 3
#define HNUMO_MAX_LAYERS 20 /* lakeAtrest supports 2..20 layers, src/initial_conditions.F90:130-169 */
#define HNUMO_MAX_NGL 9     /* nop <= 8 (BASELINE config 5) */

typedef struct hnumo_handle_s* hnumo_handle_t;

/* Everything the hot path reads from the reference's modules, in the reference's own terms. */
typedef struct hnumo_desc {
    int32_t abi_version;
    /* sizes: mod_grid nelem,nface; mod_basis ngl,nq; mod_input nlayers */
    int32_t nelem, ngl, nq, nlayers, nface;
    /* time stepping: mod_input kstages, dt; mod_initial N_btp and the re-derived dt_btp
     * (src/mod_initial.F90:176-177) */
    int32_t kstages, N_btp;
    double dt, dt_btp;
    /* physics switches: mod_input botfr, cd_mlswe, method_visc, visc_mlswe, ad_mlswe; gravity as reset by the
     * test case (9.806, src/initial_conditions.F90:97).  method_visc == 1 selects the LDG viscosity with the flux variable at the
     * quadrature points (src/mod_laplacian_quad.F90:125-223,252-355; run-time-size kernels), any other value the nodal form
     * (:32-121,227-248); both are skipped when visc_mlswe == 0.  ad_mlswe > 0 switches the vertical
     * shear stress between the layers on (src/mod_create_rhs_mlswe.F90:146-279) and needs max_shear_dz > 0 (last member). */
    int32_t botfr, method_visc;
    double gravity, cd_mlswe, visc_mlswe, ad_mlswe;
    /* 1-D operators of mod_basis (src/mod_basis.F90:157-160): psiq(ngl,nq), dpsiq(ngl,nq), wnq(nq), wgl(ngl),
     * dpsi(ngl,ngl) with dpsi(i,j) = d/dx l_i (x_j) */
    const double* psiq;
    const double* dpsiq;
    const double* wnq;
    const double* wgl;
    const double* dpsi;
    /* mesh: face(8,nface) of mod_grid (src/p4est.c:1590-1700): rows 5,6 local face of left/right element,
     * row 7 left element, row 8 right element (>0), 0 = processor boundary, -4 free slip, -2 no slip */
    const int32_t* face;
    /* per-element affine geometry (bricks, parallelograms; general quadrilaterals: see point_metrics_q at the end): elem_metrics(5,nelem) = ksiq_x, ksiq_y, etaq_x, etaq_y of
     * mod_metrics at quadrature point (1,1,1,e) and |J| = jacq(1,1,1,e)/(wnq(1)*wnq(1)) */
    const double* elem_metrics;
    /* per-face geometry: face_geom(3,nface) = normal_vector_q(1:2,1,1,iface) (outward from the left element)
     * and the edge Jacobian jac_faceq(1,1,iface)/wnq(1)   (src/create_normals_quad.F90:136-211) */
    const double* face_geom;
    /* nodal statics of mod_initial / mod_metrics, all (npoin) unless noted */
    const double* pbprime_df;
    const double* massinv;
    const double* coriolis_df;
    const double* tau_wind_df; /* (2,npoin) */
    const double* zbot_df;
    const double* alpha_mlswe; /* (nlayers) */
    /* SSPRK tables of mod_initial: ssprk_a(kstages,3), ssprk_beta(kstages)  (src/mod_initial_mlswe.F90:652-678) */
    const double* ssprk_a;
    const double* ssprk_beta;
    /* partition description of mod_parallel (src/p4est.c:1340-1420); all zero / NULL for a single rank.
     * nbh_proc(num_nbh) 1-based neighbour ranks, num_send_recv(num_nbh) faces per neighbour,
     * nbh_send_recv(sum) 1-based local face numbers in the agreed exchange order */
    int32_t rank, nranks, num_nbh;
    const int32_t* nbh_proc;
    const int32_t* num_send_recv;
    const int32_t* nbh_send_recv;
    /* 0 = use the current CUDA device */
    int32_t device;
    /* 0 = default: element-record stage kernel (one contiguous record per element; one warp per element at nop 3 and 4, one
     * block of 128 threads per element at nop 5..8; falls back to 1 for other orders or inexact integration),
     * 1 = simple reference-form kernel (any order; bisecting aid) */
    int32_t stage_kernel_variant;
    /* mod_input max_shear_dz: upper bound of the shear-layer thickness of the vertical shear stress (only read when ad_mlswe > 0;
     * added with ABI version 2) */
    double max_shear_dz;
    /* General (non-affine) quadrilaterals -- gmsh meshes, curved or skewed elements (src/metrics.F90, src/metrics_quad.F90,
     * src/create_normals.F90, src/create_normals_quad.F90): the geometry per point, in the reference's own numbers.  All NULL on
     * affine meshes (bricks, parallelograms), where elem_metrics / face_geom above describe an element by five numbers and a face by
     * three; when point_metrics_q is given the other four must be given too, elem_metrics and face_geom are not read, and the step
     * runs the run-time-size kernels (the optimised kernels assume one Jacobian per element).  Added with ABI version 3.
     *   point_metrics_q(5,npoin_q) = ksiq_x, ksiq_y, etaq_x, etaq_y, jacq   of mod_metrics at every quadrature point (jacq with its
     *                                quadrature weights, as the reference stores it)
     *   point_metrics(5,npoin)     = ksi_x, ksi_y, eta_x, eta_y, jac        at every node (jac with its LGL weights)
     *   face_geom_q(3,nq,nface)    = normal_vector_q(1:2,iq,1,iface), jac_faceq(iq,1,iface)   (outward from the left element)
     *   face_geom_n(3,ngl,nface)   = normal_vector(1:2,n,1,iface), jac_face(n,1,iface)
     *   coord(2,npoin)             = node coordinates of mod_grid (read by the Courant numbers of hnumo_diagnostics only) */
    const double* point_metrics_q;
    const double* point_metrics;
    const double* face_geom_q;
    const double* face_geom_n;
    const double* coord;
} hnumo_desc_t;

/* ---- life cycle ----------------------------------------------------------------------------- */
/* replaces: module set-up consumed by ti_rk_bcl through `use` (src/ti_rk_bcl.F90:19-28) */
int hnumo_init(const hnumo_desc_t* desc, hnumo_handle_t* out);
/* number of CUDA devices this process can use (0 when there is none: the library has no CPU fallback) */
int hnumo_device_count(void);
int hnumo_finalize(hnumo_handle_t h);
const char* hnumo_last_error(void);

/* ---- state: the three dummy arguments of ti_rk_bcl (src/ti_rk_bcl.F90:32-34) ---------------- */
/* q_df(3,npoin,nlayers), qb_df(4,npoin), qprime_df(3,npoin,nlayers) */
int hnumo_upload_state(hnumo_handle_t h, const double* q_df, const double* qb_df, const double* qprime_df);
int hnumo_download_state(hnumo_handle_t h, double* q_df, double* qb_df, double* qprime_df);

/* ---- the hot path ---------------------------------------------------------------------------- */
/* nsteps calls of ti_rk_bcl on the device-resident state (src/mod_time_loop.F90:198-211) */
int hnumo_step(hnumo_handle_t h, int32_t nsteps);
/* drop-in with the original signature: upload -> one step -> download */
int hnumo_ti_rk_bcl(hnumo_handle_t h, double* q_df, double* qb_df, double* qprime_df);

/* ---- per-phase entries (tests; same semantics as the reference routines named) --------------- */
/* btp_bcl_coeffs_qdf on the current qprime_df, with dpprime_visc = qprime_df(1,:,:)
 * (src/ti_rk_bcl.F90:43-50, src/mod_barotropic_terms.F90:219-409) */
int hnumo_btp_bcl_coeffs(hnumo_handle_t h);
/* ti_barotropic_ssprk_mlswe(qb_df, qprime_df) on the resident state (src/mod_rk_mlswe.F90:19-151) */
int hnumo_btp_substeps(hnumo_handle_t h);
/* create_rhs_btp(rhs, qb_df, qprime_df): rhs(3,npoin) to host (src/mod_rhs_btp.F90:28-59).  The time-average
 * accumulators are not touched. */
int hnumo_rhs_btp(hnumo_handle_t h, double* rhs);
/* layer_mass_rhs(dp_advec, qprime_df, qprime_df_face) on the resident qprime_df with the time averages of the last
 * hnumo_btp_substeps: dp_advec(npoin,nlayers) to host; sum_layer_mass_flux(_face) are refreshed
 * (src/mod_create_rhs_mlswe.F90:53-78,822-877,922-1034) */
int hnumo_layer_mass_rhs(hnumo_handle_t h, double* dp_advec);
/* layer_momentum_rhs(rhs_mom, qprime_df, q_df, qprime_df_face) on the resident state with dpprime_visc = qprime_df(1,:,:), the
 * coefficients of the last hnumo_btp_bcl_coeffs and the time averages of the last hnumo_btp_substeps: rhs_mom(2,npoin,nlayers)
 * to host (src/mod_create_rhs_mlswe.F90:28-51,281-820, src/mod_laplacian_quad.F90:227-248) */
int hnumo_layer_momentum_rhs(hnumo_handle_t h, double* rhs_mom);
/* rhs_layer_shear_stress(rhs_stress, q_df) on the resident q_df (needs ad_mlswe > 0): rhs_stress(2,npoin,nlayers) to host, without
 * the inverse mass matrix, as the reference routine returns it (src/mod_create_rhs_mlswe.F90:146-279) */
int hnumo_layer_shear_stress(hnumo_handle_t h, double* rhs_stress);
/* face-halo exchange of nv nodal fields, the device replacement of create_nbhs_face_df / send_bound_dg_general_df
 * (src/create_rhs_dynamics_flux.F90:104-182, src/send_receive_bound.F90:272-327): nodal(nv,npoin) in, halo(nv,ngl,nhalo) out with
 * the processor faces in the order of nbh_send_recv ("side 2 := the neighbour's side 1").  Collective over the ranks of the
 * partition.  Returns the number of processor faces of this rank or <0. */
int64_t hnumo_halo_exchange(hnumo_handle_t h, const double* nodal, int32_t nv, double* halo);
/* copy a named mod_variables work array to the host in the reference's layout; face arrays are indexed by the
 * face numbers of desc->face.  Names: Q_uu_dp Q_uv_dp Q_vv_dp H_bcl Q_uu_dp_edge Q_uv_dp_edge Q_vv_dp_edge
 * H_bcl_edge grad_zbot_quad (note: entries of grad z_bot that are pure differentiation noise, |sum| <= 1e-13 sum|terms|, i.e. the
 * derivative of a constant, are stored as exact zeros -- the one deliberate deviation from the reference's set-up arithmetic;
 * build with -DHN_NO_GZ_FLUSH to keep them) ope_ave H_ave Qu_ave Qv_ave Quv_ave ope2_ave btp_mass_flux_ave uvb_ave tau_bot_ave ope2_ave_df
 * uvb_ave_df graduvb_ave uvb_face_ave btp_mass_flux_face_ave ope_face_ave ope2_face_ave H_face_ave Qu_face_ave
 * Qv_face_ave one_plus_eta_edge_2_ave btp_dpp_graduv pbprime_visc.  Returns the element count or <0. */
int64_t hnumo_get_array(hnumo_handle_t h, const char* name, double* out, int64_t capacity);

/* ---- multi-GPU: replaces mod_mpi_communicator / send_receive_bound (src/create_rhs_communicator.F90) ---- */
/* 128-byte NCCL unique id; call on one rank, broadcast by the caller's own means (MPI_Bcast in the Fortran
 * driver, torch.distributed in the Python harness) */
int hnumo_comm_get_unique_id(void* id128);
int hnumo_comm_init(hnumo_handle_t h, const void* id128);

/* ---- device-side diagnostics ------------------------------------------------
 * Replaces, for the resident state, what the reference computes on the host at every output step:
 *   diagnostics (src/diagnostics.F90:24-45: h = alpha_k/g dp, u, v, dp, interface elevation),
 *   compute_conserved (src/compute_conserved.F90:29-41: layer mass = sum wjac_df h),
 *   print_diagnostics_mlswe (src/print_diagnostics.F90:59-128: max/min per layer and of qb(1:4,:)),
 *   courant_cube_mlswe (src/courant.F90:34-126: CFL_B as written with qb(3:4,:) = p_b*(ubar,vbar), CFL with the layer velocities).
 * out[k*11 + 0] = mass of layer k, [k*11 + 1..5] = max of h,u,v,dp,elevation, [k*11 + 6..10] = min of the same;
 * then qb max (4), qb min (4), CFL_B, CFL, min dx, min dy.  Values are those of this rank's elements: reduce over ranks like the
 * reference's mpi_reduce (sum for the mass, max, min).  Returns 11*nlayers + 12 or <0. */
int64_t hnumo_diagnostics(hnumo_handle_t h, double* out, int64_t capacity);

/* ---- text snapshots and restart (host only, no GPU needed) -------------------
 * hnumo_snapshot_write: the `mlswe####` text files of src/diagnostics.F90:73-91 (nlayers, npoin, dt, dt_btp, then one d23.16
 *   value per line: coordinates, pb, pb*ub, pb*vb, h, u, v, interface elevations, zbot) from arrays in the reference layouts
 *   (what hnumo_download_state returns).  coord (2,npoin) may be NULL.
 * hnumo_snapshot_info: header of such a file.
 * hnumo_snapshot_read_restart: load_data_mlswe + restart_mlswe (src/mod_restart.F90:15-66,88-160): reads the file and rebuilds
 *   q_df, qb_df, qprime_df for hnumo_upload_state with the statics of the restarting run.
 * Return 0, -6 I/O error, -7 malformed file, -8 nlayers/npoin mismatch. */
int hnumo_snapshot_write(const char* path, int32_t nlayers, int64_t npoin, double dt, double dt_btp, const double* coord,
                         const double* q_df, const double* qb_df, const double* zbot_df, const double* alpha_mlswe, double gravity);
int hnumo_snapshot_info(const char* path, int32_t* nlayers, int64_t* npoin, double* dt, double* dt_btp);
int hnumo_snapshot_read_restart(const char* path, int32_t nlayers, int64_t npoin, const double* pbprime_df, const double* alpha_mlswe,
                                double gravity, double* q_df, double* qb_df, double* qprime_df, double* coord_out);

/* NetCDF snapshots (classic 64-bit-offset format, written and read without a NetCDF library): the file of
 * src/diagnostics_nc.F90:98-165 -- dimensions time (unlimited), npoin, nlayers, zi; variables dt, dt_btp (time), x, y, pb, pbub, pbvb
 * (npoin), h, u, v (nlayers, npoin), eta (zi, npoin) with the reference's names and attributes.  pbprime_df(npoin) gives
 * eta(:,1) = pb / pbprime - 1.  hnumo_snapshot_read_nc_restart rebuilds q_df, qb_df, qprime_df from such a file the way restart_mlswe
 * does from the text snapshot.  Same return codes as the text functions. */
int hnumo_snapshot_write_nc(const char* path, int32_t nlayers, int64_t npoin, double dt, double dt_btp, const double* coord,
                            const double* q_df, const double* qb_df, const double* zbot_df, const double* alpha_mlswe,
                            const double* pbprime_df, double gravity);
int hnumo_snapshot_read_nc_restart(const char* path, int32_t nlayers, int64_t npoin, const double* pbprime_df, const double* alpha_mlswe,
                                   double gravity, double* q_df, double* qb_df, double* qprime_df, double* coord_out);

/* ---- measurement ------------------------------------------------------------------------------ */
/* out[0] = GPU ms spent in barotropic stages since the last reset (CUDA events on the compute stream),
 * out[1] = number of barotropic stages, out[2] = GPU ms in whole steps, out[3] = steps,
 * out[4] = kernel launches issued, out[5..7] reserved.  reset != 0 clears the counters after reading. */
int hnumo_timing(hnumo_handle_t h, double* out8, int32_t reset);
/* tuning switches, all optional (unknown keys are rejected with -2):
 *   "use_graph"            -1 auto (default) / 0 off / 1 on: CUDA-graph replay of the barotropic substep loop (only taken on
 *                          partitions without processor faces; bitwise the same result)
 *   "overlap"              1 (default) / 0: boundary elements + halo exchange on a second stream, overlapped with the interior
 *   "stage_kernel_variant" as hnumo_desc_t.stage_kernel_variant (effective if the records were allocated at init)
 *   "pair_prefetch", "pair_pf_dist"   L2 prefetch bits / distance of the stage kernel (sweeps)
 *   "layer_warp"           bit mask of the warp-per-element layer kernels (1 coefficients, 2 layer mass, 4 consistency, 8 laplacian,
 *                          16 momentum volume, 32 momentum faces + update; default 47); cleared bits run the block-per-element forms
 *   "mom_volume_batched"              1 (default): layer momentum volume term with all layers of an element in flight (nop 4, 2-3 layers); 0: layer by layer */
int hnumo_set_option(hnumo_handle_t h, const char* key, double value);

#ifdef __cplusplus
}
#endif
#endif /* HNUMO_B200_H */
