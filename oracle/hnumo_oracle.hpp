// ============================================================================
// hnumo_oracle.hpp -- TEST INFRASTRUCTURE ONLY (CPU oracle).
//
// CPU restatement (C++17, FP64) of the h-NUMO multilayer shallow-water hot path
// (everything executed inside `ti_rk_bcl`, reference src/mod_time_loop.F90:209)
// plus the one-time set-up the Fortran driver performs before the time loop.
//
// PARITY STATUS: pinned only weakly.  The reference cannot be built in this
// container (no gfortran/MPI/p4est/NetCDF).  The oracle is checked against the
// reference's own golden `CI/bump/ref_mlswe_FIN.txt` (per-layer max/min of
// h,u,v,ssh after 108 steps; see tests/test_oracle_golden.py for which digits
// agree) and against invariants the reference documents (mass conservation,
// lake at rest).  Field-level parity with the Fortran build is UNPINNED.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.  The product (h-numo_b200/) never does.
//
// Every routine cites the reference file:line it follows.  Arrays keep the
// reference's names, shapes and (column-major) index order, 0-based.
// ============================================================================
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace orc {

// Column-major N-d array with Fortran index order (first index fastest), 0-based.
struct Arr {
    std::vector<double> v;
    int d[5] = {1, 1, 1, 1, 1};
    Arr() {}
    void alloc(int d0, int d1 = 1, int d2 = 1, int d3 = 1, int d4 = 1) {
        d[0] = d0; d[1] = d1; d[2] = d2; d[3] = d3; d[4] = d4;
        v.assign((size_t)d0 * d1 * d2 * d3 * d4, 0.0);
    }
    size_t size() const { return v.size(); }
    void zero() { std::fill(v.begin(), v.end(), 0.0); }
    double* data() { return v.data(); }
    const double* data() const { return v.data(); }
    inline double& operator()(int i) { return v[i]; }
    inline double& operator()(int i, int j) { return v[i + (size_t)d[0] * j]; }
    inline double& operator()(int i, int j, int k) { return v[i + (size_t)d[0] * (j + (size_t)d[1] * k)]; }
    inline double& operator()(int i, int j, int k, int l) {
        return v[i + (size_t)d[0] * (j + (size_t)d[1] * (k + (size_t)d[2] * l))];
    }
    inline double& operator()(int i, int j, int k, int l, int m) {
        return v[i + (size_t)d[0] * (j + (size_t)d[1] * (k + (size_t)d[2] * (l + (size_t)d[3] * m)))];
    }
    inline double operator()(int i) const { return v[i]; }
    inline double operator()(int i, int j) const { return v[i + (size_t)d[0] * j]; }
    inline double operator()(int i, int j, int k) const { return v[i + (size_t)d[0] * (j + (size_t)d[1] * k)]; }
    inline double operator()(int i, int j, int k, int l) const {
        return v[i + (size_t)d[0] * (j + (size_t)d[1] * (k + (size_t)d[2] * l))];
    }
    inline double operator()(int i, int j, int k, int l, int m) const {
        return v[i + (size_t)d[0] * (j + (size_t)d[1] * (k + (size_t)d[2] * (l + (size_t)d[3] * m)))];
    }
};

enum TestCase { TC_BUMP = 0, TC_LAKE = 1, TC_DOUBLE_GYRE = 2, TC_DOUBLE_GYRE_SYNTH = 3 };

// Mirrors the namelist keys of numo3d.in that the hot path reads
// (reference src/mod_input.F90:320-380).
struct Config {
    int nelx, nely, nop, nlayers;
    double xdims[2], ydims[2];
    int x_boundary[2], y_boundary[2];  // 4 = free slip, 2 = no slip
    double dt, dt_btp;
    int kstages;
    int botfr;
    double cd_mlswe;
    int method_visc;
    double visc_mlswe;
    double f0, beta;
    int test_case;
    int dg_integ_exact;
    // Synthetic double gyre (test_case 3): explicit interface depths (nlayers+1, negative
    // downwards, z[0]=0) and alpha per layer; perturbation amplitude on dp'.
    double synth_z[33];
    double synth_alpha[32];
    double synth_perturb;
    // 1: replace the numerically differentiated metrics (which carry ~1e-14 relative round-off that differs from
    // point to point, exactly like the reference's metrics.F90) by their per-element / per-face constant values
    // (first quadrature point), i.e. the affine-brick geometry the CUDA library is given.  0 = as the reference.
    int affine_metrics;
    // vertical shear stress between the layers (mod_input ad_mlswe, max_shear_dz; mod_create_rhs_mlswe.F90:146-279); 0 = off
    double ad_mlswe, max_shear_dz;
    // Test aid: the reference's metric, normal and operator-table code is written for general (isoparametric) quadrilaterals
    // (gmsh meshes; metrics.F90, metrics_quad.F90, create_normals(_quad).F90), but its shipped cases are bricks.  mesh_warp != 0
    // moves the nodes of the brick by  dx = w hx sin(pi kx X) cos(pi ky Y),  dy = w hy cos(pi kx X) sin(pi ky Y)
    // (X, Y in [0,1], hx, hy element sizes, kx = max(1, nelx/2), ky = max(1, nely/2)): curved, non-affine elements, conforming
    // (the displacement is a function of position), the domain boundary stays where it is.  affine_metrics must be 0.
    double mesh_warp;
};

struct Oracle {
    Config cfg;
    // sizes
    int ngl, nq, npts, nelem, npoin, npoin_q, nface, nl;
    int N_btp, kstages;
    double dt, dt_btp, gravity;
    // basis (mod_basis)
    std::vector<double> xgl, wgl, xnq, wnq;
    Arr psi, dpsi, psiq, dpsiq;  // psi(ngl,ngl) dpsi(ngl,ngl) psiq(ngl,nq) dpsiq(ngl,nq)
    // grid
    Arr coord;                   // (2,npoin)
    std::vector<int> face;       // (8,nface) Fortran 1-based content kept (face(5..8))
    std::vector<int> fnodeL, fnodeR;  // (ngl,nface) 0-based global node of imapl/imapr
    // Threading aid (not in the reference): the faces of every element in ascending face number, side 0 = the element is
    // the face's left element, 1 = right.  The reference scatters face by face into the nodes of both elements; the
    // threaded loops evaluate the faces in parallel and then gather per element in the same face order, so every node
    // receives the same additions in the same order -> bit-identical to the serial face loop.
    std::vector<int> ef_ptr, ef_face, ef_side;
    void build_elem_faces();
    // metrics
    Arr ksi_x, ksi_y, eta_x, eta_y, jac;         // (npoin)
    Arr ksiq_x, ksiq_y, etaq_x, etaq_y, jacq;    // (npoin_q)
    Arr massinv;                                 // (npoin)
    Arr normal_vector, jac_face;      // (2,ngl,nface), (ngl,nface)
    Arr normal_vector_q, jac_faceq;   // (2,nq,nface), (nq,nface)
    // dense operator tables (Tensor_product.F90)
    Arr psih, dpsidx, dpsidy, wjac;          // (npts,npoin_q) ... (npoin_q)
    Arr psih_df, dpsidx_df, dpsidy_df, wjac_df;
    std::vector<int> indexq, index_df;       // (npts,npoin_q), (npts,npoin)
    // statics (mod_initial)
    Arr alpha_mlswe, pbprime, pbprime_df, pbprime_face, one_over_pbprime, one_over_pbprime_face,
        pbprime_edge, one_over_pbprime_edge, one_over_pbprime_df, one_over_pbprime_df_face, pbprime_df_face,
        tau_wind, tau_wind_df, coriolis_df, coriolis_quad, coeff_pbpert_L, coeff_pbpert_R, coeff_pbub_LR,
        coeff_mass_pbub_L, coeff_mass_pbub_R, coeff_mass_pbpert_LR, zbot, zbot_df, zbot_face, grad_zbot_quad,
        fdt_bcl, fdt2_bcl, a_bcl, b_bcl, ssprk_a, ssprk_beta, z_interface;
    // state
    Arr q_df, qb_df, qprime_df;  // (3,npoin,nl) (4,npoin) (3,npoin,nl)
    // mod_variables work arrays
    Arr Q_uu_dp, Q_uv_dp, Q_vv_dp, H_bcl, Q_uu_dp_edge, Q_uv_dp_edge, Q_vv_dp_edge, H_bcl_edge;
    Arr ope_ave, H_ave, Qu_ave, Qv_ave, Quv_ave, ope2_ave, btp_mass_flux_ave, uvb_ave, ope2_ave_df, uvb_face_ave,
        btp_mass_flux_face_ave, ope_face_ave, H_face_ave, Qu_face_ave, Qv_face_ave, Quv_face_ave, tau_wind_ave,
        tau_bot_ave, one_plus_eta_edge_2_ave, uvb_ave_df, ope2_face_ave;
    Arr dpprime_visc, pbprime_visc, btp_dpp_graduv, dpp_graduv, graduv_dpp_face, btp_graduv_dpp_face,
        graduvb_face_ave, graduvb_ave;
    Arr sum_layer_mass_flux, sum_layer_mass_flux_face;
    int error_flag = 0;  // 1 = negative layer thickness (mod_splitting.F90:74-77)
    long stage_count = 0;
    double btp_seconds = 0.0;  // wall time spent in ti_barotropic_ssprk_mlswe (for the CPU baseline)

    std::map<std::string, Arr*> reg;

    explicit Oracle(const Config& c);
    // set-up
    void build_basis();
    void build_grid();
    void build_metrics();
    void build_faces();
    void make_metrics_affine();
    void build_tensor_tables();
    void build_initial();
    void allocate_variables();
    // hot path
    void ti_rk_bcl();
    void ti_barotropic_ssprk_mlswe(Arr& qb, const Arr& qprime);
    void create_rhs_btp(Arr& rhs, const Arr& qb, const Arr& qprime);
    void btp_extract_df(Arr& qb_df_face, const Arr& qb);
    void create_rhs_btp_volume_qdf(Arr& rhs, const Arr& qb, const Arr& qprime);
    void creat_btp_fluxes_qdf(Arr& rhs, const Arr& qb_df_face);
    void btp_create_laplacian(Arr& rhs_lap, const Arr& qb);
    void compute_gradient_uv(Arr& grad_uv, const double* uv, int ld);  // uv(2,npoin) slice with leading dim ld
    void btp_mom_boundary_df(Arr& qb);
    void btp_bcl_coeffs_qdf(const Arr& qprime_df_face, const Arr& qprime);
    void extract_qprime_df_face(Arr& qprime_df_face, const Arr& qprime);
    void extract_dprime_df_face(Arr& qprime_df_face, const Arr& qprime);  // writes component 0 only
    void momentum_mass(Arr& q, Arr& qprime_df_face, Arr& qprime, const Arr& qb);
    void thickness(Arr& qprime, Arr& q, const Arr& qb, Arr& qprime_df_face);
    void momentum(Arr& q, Arr& qprime, const Arr& qb, const Arr& qprime_df_face);
    void layer_mass_rhs(Arr& dp_advec, const Arr& qprime, const Arr& qprime_df_face);
    void apply_consistency(Arr& q);
    void rhs_momentum(Arr& rhs_mom, const Arr& qprime, const Arr& q, const Arr& qprime_df_face);
    void bcl_create_laplacian(Arr& rhs_lap);
    void layer_momentum_volume(Arr& rhs_mom, const Arr& qprime, const Arr& q);
    void apply_layers_fluxes(Arr& rhs_mom, const Arr& qprime_df_face);
    void layer_mom_boundary_df(Arr& q);
    // method_visc == 1 (oracle_visc_q.cpp)
    Arr dpprime_visc_q;   // (npoin_q, nl)
    void interpolate_dpp();
    void compute_gradient_uv_q(Arr& grad_uv, const Arr& uv);
    void visc_flux_faces(Arr& ff, const Arr& flux);
    void compute_laplacian_quad(Arr& lap_q, const Arr& grad_dpuvp);
    void create_rhs_laplacian_flux_quad(Arr& rhs, const Arr& gradq_face);
    void btp_create_laplacian_v2(Arr& rhs_lap, const Arr& qprime, const Arr& qb);
    void bcl_create_laplacian_v2(Arr& rhs_lap, const Arr& qprime);
    void rhs_layer_shear_stress(Arr& rhs_stress, const Arr& q);
    void velocity_df(Arr& q, const Arr& qb);
    void add_shear_stress(Arr& q_df_temp, const Arr& q, const Arr& qb);
    void extract_velocity(Arr& uv, const Arr& q, const Arr& qb);
    void evaluate_bcl(Arr& qprime_df_face, Arr& q, Arr& qprime, const Arr& qb);
    void evaluate_bcl_v1(Arr& q, Arr& qprime, const Arr& qb);
    // diagnostics (diagnostics.F90:24-45, compute_conserved.F90)
    void diagnostics(Arr& qout) const;
    double layer_mass(const Arr& qout, int k) const;
};

}  // namespace orc
