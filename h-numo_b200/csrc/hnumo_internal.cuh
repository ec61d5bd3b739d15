// hnumo_internal.cuh -- device-resident solver state and shared device helpers.
//
// Data layout in HBM: structure-of-arrays "planes", element-blocked.
//   nodal plane   P[e*npts + m*ngl + n]          (same point order as the reference's intma_dg, mod_grid.F90:230)
//   quad plane    Q[e*nq2  + j*nq  + i]          (intma_dg_quad, mod_grid.F90:242)
//   slot plane    S[(e*4+s)*nq + iq]             one entry per element side; a face's data lives at its OWNER
//                                                slot = the side of the face's left element (p4est.c:1693-1704)
//   slots s: 0: eta=-1 (numa local face 3), 1: eta=+1 (4), 2: ksi=-1 (5), 3: ksi=+1 (6)
// Faces are never stored as a list on the device: every element evaluates all four of its faces from the
// canonical left-element perspective (gather only, no atomics, deterministic).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/hnumo_b200.h"

#define HN_MAXL HNUMO_MAX_LAYERS
#define HN_MAXNGL HNUMO_MAX_NGL
#define HN_MAXNQ (2 * HNUMO_MAX_NGL - 1)

namespace hn {

// neighbour codes in nbr[]
enum { NBR_HALO = -1, NBR_NOSLIP = -2, NBR_FREESLIP = -4 };

// Small, kernel-argument-sized view of the static mesh/operator data.
struct Mesh {
    int nelem, ngl, nq, npts, nq2, nl;
    int npoin, npoin_q, nslots;
    const int* nbr;      // [nslots] neighbour element (>=0), or NBR_* code
    const int* nbslot;   // [nslots] neighbour's slot (interior), halo face index (NBR_HALO)
    const double* fgeom; // [nslots*3] canonical (left-element) unit normal nx,ny and edge jacobian
    const double* em;    // [nelem*5] ksi_x ksi_y eta_x eta_y |J|
    // general (non-affine) quadrilaterals: geometry per point, nullptr on affine meshes (the kernels read them through met_q / met_n /
    // fg_q / fg_n of hnumo_dev.cuh, which fall back to the per-element numbers above)
    const double* mq;    // [5][npoin_q] ksiq_x ksiq_y etaq_x etaq_y jacq/(w_i w_j)      (mod_metrics, src/metrics_quad.F90)
    const double* mn;    // [5][npoin]   ksi_x ksi_y eta_x eta_y jac/(w_i w_j)           (src/metrics.F90)
    const double* fgq;   // [3][nslots*nq]  normal_vector_q(1:2), jac_faceq/w_iq at the face quadrature points (canonical left normal)
    const double* fgn;   // [3][nslots*ngl] normal_vector(1:2), jac_face/w_n at the face nodes
    const double* coord; // [2][npoin] node coordinates (mod_grid coord), read by the Courant diagnostics only
};

// 1-D operator tables (constant memory copies are in hnumo_ops.cu)
struct alignas(16) Ops {
    double A[HN_MAXNGL * HN_MAXNQ];   // psiq(n,i)   -> A[n + ngl*i]
    double B[HN_MAXNGL * HN_MAXNQ];   // dpsiq(n,i)
    double D[HN_MAXNGL * HN_MAXNGL];  // dpsi(k,n) = l_k'(x_n) -> D[k + ngl*n]
    double wq[HN_MAXNQ];
    double wg[HN_MAXNGL];
    double xg[HN_MAXNGL];             // LGL nodes (recomputed from ngl: the descriptor carries weights and matrices only)
    // transposed copies (output index fastest) for the line contractions of stage_pair.cuh: consecutive entries feed
    // independent accumulators, and one 128-bit uniform load brings two of them
    alignas(16) double AT[HN_MAXNGL * HN_MAXNQ];   // AT[i + nq*n] = A[n + ngl*i]
    alignas(16) double BT[HN_MAXNGL * HN_MAXNQ];
    alignas(16) double DT[HN_MAXNGL * HN_MAXNGL];  // DT[n + ngl*k] = D[k + ngl*n]
};

struct Planes {  // convenience: contiguous set of planes
    double* p = nullptr;
    size_t stride = 0;  // elements per plane
    int n = 0;
    __host__ __device__ double* operator[](int i) const { return p + (size_t)i * stride; }
};

struct Solver {
    hnumo_desc_t desc;  // scalars only are valid after init (pointers are cleared)
    Mesh mesh;
    Ops ops;
    int device = 0;
    cudaStream_t stream = nullptr, comm_stream = nullptr;
    // scalars
    int nelem, ngl, nq, npts, nq2, nl, nface, npoin, npoin_q, nslots;
    int kstages, N_btp, botfr, has_visc;
    double dt, dt_btp, g, cd, visc;
    double ad = 0.0, max_shear_dz = 0.0;   // vertical shear stress between the layers (0 = off)
    double ssprk_a[5][3], ssprk_beta[5], alpha[HN_MAXL];
    int variant = 0;
    // host copies of the face table (for reference-layout output) : face f -> owner slot, right slot (or -1)
    std::vector<int> face_owner_slot, face_right_slot, face_er;
    // connectivity
    int *d_nbr = nullptr, *d_nbslot = nullptr;
    double *d_fgeom = nullptr, *d_em = nullptr;
    bool general = false;   // general quadrilaterals: per-point geometry (Mesh::mq, mn, fgq, fgn, coord)
    double *d_mq = nullptr, *d_mn = nullptr, *d_fgq = nullptr, *d_fgn = nullptr, *d_coord = nullptr;
    // static nodal planes
    double *pbprime_df, *oop_df, *massinv, *coriolis_df, *tauw_df /*2*/, *zbot_df, *a_bcl, *b_bcl, *fdt2;
    // static quad planes (derived exactly like the reference set-up derives them)
    double *pbprime_q, *oop_q, *coriolis_q, *tauw_q /*2*/, *gradzb_q /*2*/;
    // static slot planes [nslots*nq] (valid at owner slots): Riemann coefficients etc.
    double *cL, *cR, *cLR, *lam, *oop_edge, *pbf_l, *pbf_r, *zbf_l, *zbf_r, *pbl, *pbr;
    // static slot-node plane [nslots*ngl]: pbprime_df at the neighbour's face nodes (pbprime_df_face(2,..))
    double* pbn;
    // state
    Planes qb;      // 3: pbpert, pbub, pbvb   (pb = pbpert + pbprime_df)
    Planes q;       // 3*nl: [var*nl + k]
    Planes qprime;  // 3*nl
    // work (baroclinic)
    Planes qbp, q2, qprime2, qprime3, dpv, qdp_tmp, dpprime2;
    // barotropic work
    Planes qb0, qb2w;
    Planes trace[2];  // 7 planes each of [ (nslots + nhalo) * ngl ]
    // btp<-bcl coefficients
    double *Quu, *Quv, *Qvv, *Hbcl;                  // quad planes
    double *Quu_e, *Quv_e, *Qvv_e, *Hbcl_e;          // slot planes
    Planes dpp_graduv;                               // 4*nl nodal planes [v*nl+k]
    Planes btp_dpp_graduv;                           // 4 nodal
    double* pbprime_visc;                            // nodal
    // accumulators (reduced set) ...
    Planes acc_n;   // nodal: 0 ope2_df, 1 ub_df, 2 vb_df, 3 S_pbpert, 4 S_mx, 5 S_my, 6..9 graduvb
    Planes acc_q;   // quad : 0 Qu, 1 Qv, 2 Quv, 3 ope2, 4 ub, 5 vb, (6,7 tau_bot when botfr==2)
    Planes acc_f;   // slot : 0 quu,1 quv,2 qvu,3 qvv,4 ope2_L,5 ope2_R,6 ope_e2,7 uL,8 uR,9 vL,10 vR
    // ... and the full set of reference averages produced by the finalize kernel
    Planes ave_q;   // 0 ope,1 H,2 Qu,3 Qv,4 Quv,5 ope2,6 mfx,7 mfy,8 ub,9 vb,10 tbx,11 tby
    Planes ave_f;   // 0 mfx,1 mfy,2 H,3 quu,4 quv,5 qvu,6 qvv,7 ope_L,8 ope_R,9 ope2_L,10 ope2_R,11 ope_e2,12 uL,13 uR,14 vL,15 vR
    Planes ave_n;   // 0 ope2_df,1 ub_df,2 vb_df,3..6 graduvb
    // baroclinic work
    Planes slmf_q;  // 2 quad: sum_layer_mass_flux
    Planes slmf_f;  // 2 slot: sum_layer_mass_flux_face
    Planes rhs_mom, rhs_visc;  // 2*nl nodal planes each
    Planes rhs_full;           // 2*nl, only when ad > 0
    // method_visc == 1 (visc_q.cuh): S_c, P at the quadrature points; face values of the flux variable in slot planes of width nq
    int visc_q = 0;
    double* vq_P = nullptr;
    Planes vq_S;               // 4 quad planes
    Planes trq[2];             // 4 planes of (nslots + nhalo) * nq: barotropic flux variable, ping-pong over the stages
    Planes trq_l;              // 4*nl planes: layers
    double* stage_buf = nullptr;  // AoS staging for upload/download
    // halo copies of neighbour nodal traces on processor faces: [plane][nhalo*ngl]
    Planes h_q, h_dp, h_dpv, h_dpg, h_gub, h_stat;
    void* local_group = nullptr;
    int* d_flag = nullptr;  // physics error flag
    // halo
    int nhalo = 0;  // processor-boundary faces
    std::vector<int> halo_slot;  // halo index -> local slot
    int* d_halo_slot = nullptr;
    std::vector<int> nbh_rank, nbh_count, nbh_offset;
    double *d_send = nullptr, *d_recv = nullptr;
    size_t halo_capacity = 0;  // doubles per buffer
    void* nccl_comm = nullptr;
    // timing
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // interior stage done / boundary stage + exchange done (btp_solve_pair)
    // pipelined drop-in call (hnumo_ti_rk_bcl): host<->device copies on copy_stream overlapped with the step
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_pipe[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int pipe_wait_q = 0;               // bcl_step: q_df is still on its way (staged in q2), transpose it before the first layer kernel
    double* pipe_qb_host = nullptr;    // bcl_step: download qb_df as soon as the corrector's barotropic solve is done
    std::vector<std::pair<void*, size_t>> pinned_host;   // caller arrays page-locked by the library (cudaHostRegister)
    void* graph_exec = nullptr;   // CUDA graph of one cycle of the substep loop (see btp_solve_pair)
    int graph_key = -1, use_graph = -1;   // use_graph: -1 auto (on when the partition has no processor faces), 0 off, 1 on
    std::vector<cudaEvent_t> ev_pool;  // pairs (start, stop) around every barotropic stage loop / whole step
    std::vector<int> ev_kind;          // 0 = barotropic stage loop, 1 = whole step
    size_t ev_used = 0;
    double ms_btp = 0, ms_step = 0, ms_btp_last = 0, ms_step_last = 0;
    long n_stages = 0, n_steps = 0, n_launches = 0;
    // record layout of the element-pair stage kernel (stage_pair.cuh): one record per element, face sums, traces
    double *p_rec = nullptr, *p_accf = nullptr, *p_tr[2] = {nullptr, nullptr};
    int pair_ne = 1, pair_warps = 4, pair_prefetch = 6, pair_pf_dist = 0, pair_units_per_wave = 0;
    // halo exchange overlapped with interior work: the elements that own a processor face advance on comm_stream
    // (high priority) followed by pack + send/recv, every other element advances on `stream` at the same time
    int* d_belems = nullptr;
    int n_belem = 0, overlap = 1;
    double *d_diag_partial = nullptr, *d_diag_res = nullptr;   // device-side diagnostics (diag.cuh), allocated on first use
    int num_sms = 148;
    // warp-per-element layer kernels (layer_warp.cuh) where instantiated, bit per kernel: 1 coeffs, 2 layer mass, 4 consistency,
    // 8 laplacian, 16 momentum volume, 32 momentum faces + update; 0 = block-per-element kernels (bcl_kernels.cuh).
    // Default 47: the momentum volume kernel is faster in its block-per-element form (1.37 vs 1.57 ms per launch at 62 500
    // elements: the warp form interpolates every layer twice), see profiles/r2_layer_kernels.md
    int layer_warp = 47;
    int mom_volume_batched = 1;   // momentum volume term with all layers of an element in flight (nop 4, 2 or 3 layers)
    std::vector<void*> allocs;
};

// allocation helpers (host side)
double* dalloc(Solver& S, size_t n);
Planes palloc(Solver& S, int nplanes, size_t stride);

// phases (host drivers)
int btp_bcl_coeffs(Solver& S, const Planes& qprime, const Planes& dpv);
int btp_solve(Solver& S, Planes& qb, const Planes& qprime);
int rhs_btp_only(Solver& S, Planes& qb, const Planes& qprime, double* d_rhs_out);
int bcl_step(Solver& S);
void upload_ops(const Ops& ops);
int halo_exchange(Solver& S, const double* const* planes, int nplanes, double* halo_out, size_t plane_stride);

#define HN_CUDA(call)                                                          \
    do {                                                                       \
        cudaError_t _e = (call);                                               \
        if (_e != cudaSuccess) { hn::set_error(#call, cudaGetErrorString(_e)); return -1; } \
    } while (0)
void set_error(const char* what, const char* detail);

}  // namespace hn
