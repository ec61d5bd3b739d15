// stage_pair.cuh -- barotropic SSPRK stage kernel, "element pair per warp" (default variant).
//
// Same operator as k_btp_stage_simple / k_btp_stage_fused (reference src/mod_rhs_btp.F90:28-370,
// src/mod_barotropic_terms.F90:25-97,165-217, src/mod_laplacian_quad.F90:32-121,357-519, src/mod_rk_mlswe.F90:87-114).
//
// Shape (history and measurements: profiles/r1_fused_kernel_summary.md, r1_record_kernel_summary.md,
// r1_stage_kernel_experiments.md):
//   * everything an element needs for a stage lives in ONE contiguous record (state, statics, running sums, face
//     coefficients): one base pointer per element, compile-time offsets, no per-plane address arithmetic;
//   * default at nop 3/4: one warp advances one element (NE = 1), 4 warps per block, 16 warps per SM (13.7 kB of shared
//     memory per warp, 128 registers); the loads of a phase are issued one phase ahead and held in registers, the statics
//     of the whole element are requested when the warp starts; every warp prefetches the head and the quadrature statics
//     of the record a quarter wave ahead into L2 (cp.async.bulk.prefetch.L2);
//   * NE = 2 (two elements per warp, every shared-memory word a double2) is kept as an option: fewer instructions per
//     element but half the resident warps -- measured slower;
//   * nop 8: the same source with BLK = true, one element per block of 128 threads, phases separated by __syncthreads,
//     independent line jobs of a phase on different warps, 96 registers (5 blocks per SM);
//   * the LDG face flux is evaluated by the lane that gathers the traces (no staging of gradient traces), and the two
//     scatter passes are one iteration each with the psiq/dpsiq parts chained in registers;
//   * running sums are fire-and-forget RED.E.ADD.F64 (one add per address and launch).
// The kernel is memory bound by design (1.3 flop/B); tensor cores do not apply (FP64, 5x9 / 9x17 operators, FP64 pipe
// 35 % busy at both orders).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "btp_kernels.cuh"
#include "line_ops.cuh"

namespace hn {


// forcing-sparsity flags of an element (header word 22): a zero field is not read
enum { PF_GZ = 1, PF_TWX = 2, PF_TWY = 4, PF_COR = 8 };

template <int G, int Q>
struct PairRec {
    static constexpr int NP = G * G, NQ2 = Q * Q, PER = G * Q;
    // header (doubles): 0 ksx 1 ksy 2 etx 3 ety 4 J 5 - | 6..17 fgeom[4][3] | 18,19 int nbr[4] | 20,21 int trec[4] |
    //                   22 int flags, int - | 23 -
    static constexpr int HDR = 24;
    static constexpr int QBSZ = pr_pad2(3 * NP);
    static constexpr int O_QB = HDR;
    // nodal statics: 0 pbprime 1 1/pbprime 2 massinv 3 qp_dp 4 qp_u 5 qp_v 6 pbprime_visc 7..10 btp_dpp_graduv
    static constexpr int NST_NF = 11;
    static constexpr int O_NST = O_QB + QBSZ;
    static constexpr int O_ACCN = O_NST + pr_pad2(NST_NF * NP);   // 0 ope2_df 1 ub 2 vb 3 S_pbpert 4 S_mx 5 S_my
    // quadrature statics: 0 H_bcl 1 Quu 2 Quv 3 Qvv 4 coriolis 5 tau_x; rarely-read tail of the record: tau_y dzb/dx dzb/dy
    // (1/pbprime at the quadrature points is not stored: it is 1/(pb - pbpert) of the interpolated state)
    static constexpr int QST_NF = 6, QST_RARE = 3;
    static constexpr int O_QST = O_ACCN + pr_pad2(6 * NP);
    static constexpr int O_ACCQ = O_QST + pr_pad2(QST_NF * NQ2);  // 0 Qu 1 Qv 2 Quv 3 ope2 4 ub 5 vb (tail: 6,7 tau_bot, botfr==2)
    // face statics per side: cL cR cLR lam oop_edge Quu_e Quv_e Qvv_e Hbcl_e 1/pbl 1/pbr (copy of the owner's values)
    static constexpr int FSIDE = pr_pad2(11 * Q);
    static constexpr int O_FST = O_ACCQ + pr_pad2(6 * NQ2);
    // the neighbour's pbprime at the face nodes.  (The neighbour's LDG statics are not needed: a trace record carries the flux
    // variable q = pbprime_visc grad(ub,vb) + btp_dpp_graduv of its element, which is all the face flux reads of the other side.)
    static constexpr int VSIDE = pr_pad2(G);
    static constexpr int O_VST = O_FST + 4 * FSIDE;
    static constexpr int O_Q0 = O_VST + 4 * VSIDE, O_Q2 = O_Q0 + QBSZ;         // SSPRK work states
    static constexpr int O_QSTR = O_Q2 + QBSZ;                                   // rare quadrature statics 7..9
    static constexpr int O_ACCQR = O_QSTR + pr_pad2(QST_RARE * NQ2);             // rare quadrature sums 6,7
    static constexpr int REC = O_ACCQR + pr_pad2(2 * NQ2);
    static constexpr int ASIDE = pr_pad2(11 * Q);   // face sums, separate array [slot][ASIDE]
    static constexpr int TSIDE = pr_pad2(7 * G);    // trace record [slot][7][G]: pbpert mx my q0..q3 (LDG flux variable)
    // shared memory per warp, in units of NE doubles.  The strides SX, ST, TM are padded so that in the line phases
    // (lane = Q*f + i, or lane = G*f + m) the word index is congruent to the lane number modulo 16: no bank conflicts
    // (profiles/smem_conflicts.py).
    static constexpr int SX = LineGeom<G, Q>::SX;   // stride of the quadrature-point arrays
    static constexpr int TM = LineGeom<G, Q>::TM;   // row stride of the pass-1 arrays T[f][m][i]
    static constexpr int ST = LineGeom<G, Q>::ST;
    static constexpr int S_NOD = 0;                 // 0 dpp 1 mx 2 my 3 pb 4 pp 5 up 6 vp 7 u 8 v ; later 4..7 = LDG flux variable
    static constexpr int S_X = S_NOD + 9 * NP;      // 8 quadrature-point arrays; later rhs, face traces, face fluxes
    // Face phases 7b-7d of the warp form (nop 3, 4): row and side strides are chosen so that the shared-memory word index stays
    // congruent to the lane number modulo 16 across the four sides -- the natural strides (8 Q per side, Q per row) cost one
    // bank-conflict replay per half-warp that straddles a side (62 of 980 data-pipe wavefronts per element with the gradient-line
    // lanes below; profiles/tools/bank_model.py).  The block form (nop 5-8) keeps the natural strides.
    static constexpr bool WARPF = (G <= 5);
    static constexpr int TR7 = WARPF ? 10 : Q;                                        // interpolated traces T[s][L/R x var][iq]: row stride (even, = 2 mod 4)
    static constexpr int TS7 = WARPF ? LineGeom<G, Q>::pick(8 * TR7, Q) : 8 * Q;      // ... side stride (= Q mod 16)
    static constexpr int FR7 = WARPF ? 10 : Q;                                        // face fluxes X_FF[s][field][iq]
    static constexpr int FS7 = WARPF ? LineGeom<G, Q>::pick(3 * FR7, Q) : 3 * Q;
    static constexpr int X_RHS = 0, X_FL = 3 * NP, X_FR = X_FL + 16 * G + (WARPF ? 2 : 0), X_LF = X_FR + 16 * G, X_FF = X_LF + 8 * G;
    static constexpr int S_T = S_X + 8 * SX;        // pass-1 results; later interpolated traces, projected face fluxes
    static constexpr int T_SZ = 8 * ST;
    static constexpr int S_L = S_T + T_SZ;          // LDG: 0..3 gradient lines / laplacian lines, 4..7 G, 8..11 Z
    static constexpr int S_TOTAL = S_L + 12 * NP;
    static_assert(X_FF + 4 * FS7 <= 8 * SX, "face work does not fit in X");
    static_assert(4 * TS7 <= T_SZ && 7 * ST <= T_SZ, "T too small");
    static size_t smem_bytes(int ne, int warps) { return (size_t)warps * ((size_t)S_TOTAL * ne + (size_t)HDR * ne) * sizeof(double); }
};

struct PairArgs {
    int nelem, nslots;
    double* rec;
    double* accf;
    const double* tr_in;
    double* tr_out;
    double a1, a2, a3, dtt, g, cd_g, cd_alpha, visc;   // cd_g = cd/g (botfr 1), cd_alpha = cd/alpha_bottom (botfr 2)
    int botfr, load_q0, load_q2, store_q0, store_q2;
    int prefetch, pf_dist;   // L2 prefetch: bit 0: own record tail at start; bits 1,2: head (header, state, nodal statics and sums) and
                             // quadrature statics of the record pf_dist units ahead
    // element subsets of a launch (halo exchange overlapped with interior work):
    //   part 0: every element; part 1: the `count` elements of `elist` (those with a processor face);
    //   part 2: every element that has no processor face (warps of the others leave once their header has arrived)
    int part, count;
    const int* elist;
    // rhs_only (per-phase test entry hnumo_rhs_btp): with a1 = a2 = a3 = 0 and dtt = 1 the "new state" the kernel stores in
    // the records IS create_rhs_btp of the packed state (mod_rhs_btp.F90:28-59); the flag only switches the wall projection
    // off.  The caller unpacks the records afterwards; the next solve packs them afresh.
    int rhs_only;
    int configure_only;   // host: set the per-device function attributes and return (hnumo_init, before any graph capture)
};


// phase time stamps of a sample of warps (debug builds only: -DHN_PAIR_TIMING)
#ifdef HN_PAIR_TIMING
#define HN_PT_SLOTS 16
#define HN_PT_UNITS 4096
__device__ long long g_pair_timing[HN_PT_UNITS * HN_PT_SLOTS];
#define PR_STAMP(k) do { if (lane == 0 && (unit % 61) == 0 && unit / 61 < HN_PT_UNITS) g_pair_timing[(unit / 61) * HN_PT_SLOTS + (k)] = clock64(); } while (0)
#else
#define PR_STAMP(k) do { } while (0)
#endif

// reciprocal: hardware seed + two Newton steps (<= 1 ulp); the IEEE division sequence costs ~20 issue slots and a
// slow-path branch per use, and there are 14 of them per element and stage
__device__ __forceinline__ double pr_rcp(double a) {
#ifdef HN_EXACT_DIV
    return 1.0 / a;
#else
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
    double e = fma(-a, x, 1.0);
    x = fma(x, e, x);
    e = fma(-a, x, 1.0);
    x = fma(x, e, x);
    return x;
#endif
}

// running sums: fire-and-forget FP64 reduction at the L2 (RED.E.ADD.F64).  Every address receives exactly one add per
// launch, so the result is bitwise the one of load-add-store; no register is tied up by the old value and no load
// has to be waited for.
#if defined(HN_PAIR_DIAG) && HN_PAIR_DIAG == 1
__device__ __forceinline__ void pr_red(double* p, double v) { *p = v; }          // timing diagnostics only (wrong sums)
#elif defined(HN_PAIR_DIAG) && HN_PAIR_DIAG == 2
__device__ __forceinline__ void pr_red(double* p, double v) { }
#else
__device__ __forceinline__ void pr_red(double* p, double v) { atomicAdd(p, v); }
#endif

// (Running sums by ONE asynchronous bulk reduction per sum array -- cp.reduce.async.bulk .add.f64 from a shared-memory staging
//  area -- were measured slower than the RED instructions, 1.567 vs 1.533 ms per stage at 500x500 elements: the pointwise phase
//  waits for the memory system, not for the issue of the REDs.  profiles/r1_stage_kernel_experiments.md; the code path was removed
//  in round 2.)

__device__ __forceinline__ void pr_prefetch_l2(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

template <int G>
__device__ __forceinline__ int pr_face_node(int s, int n) {
    return s == 0 ? n : s == 1 ? (G - 1) * G + n : s == 2 ? n * G : n * G + G - 1;
}

// BOTFR: 0 none, 1 linear bottom drag, 2 quadratic bottom drag (mod_rhs_btp.F90:149-169) -- compile time, so that the
// pointwise phases are single basic blocks the scheduler can interleave
//
// BLK = false: W warps per block, each warp advances NE elements (lane = thread within the warp, phases separated by
//              __syncwarp).  Orders up to nop 4 (every per-lane item list fits in 32 lanes).
// BLK = true : one element per block of NT = 32*W threads, same phases with "lane" = thread within the block and
//              __syncthreads between them: high orders (nop 8: 81 nodes, 289 quadrature points, 68 face points per element).
#ifndef HN_BLK_MAXNREG
#define HN_BLK_MAXNREG 96    // registers per thread of the block-per-element form: 96 -> 5 blocks of 128 threads per SM (no spills; 0.494 of the roofline vs 0.464 with 128 registers and 4 blocks)
#endif
#ifndef HN_WARP_MAXNREG
#define HN_WARP_MAXNREG 128  // registers per thread of the warp-per-element form: 128 -> 16 warps per SM (96 -> 20 warps with 11.3 kB of shared
                             // memory per warp was measured equal: 1.353 vs 1.360 ms per stage at 500x500 elements, profiles/r2_stage_kernel_experiments.md)
#endif
#ifndef HN_BLK_W
#define HN_BLK_W 4           // warps per element of the block-per-element form at nop 8 (5: 1.308 ms, 6: 1.38-1.54 ms, 4: 1.231 ms per stage at 250x250)
#endif
template <int NT> __device__ __forceinline__ void pr_sync() { if (NT == 32) __syncwarp(); else __syncthreads(); }
template <int G, int Q, int NE, int W, bool VISC, int BOTFR, bool BLK = false>
__global__ void __launch_bounds__(32 * W) __maxnreg__(BLK ? HN_BLK_MAXNREG : NE == 2 ? (W == 4 ? 224 : 168) : HN_WARP_MAXNREG) k_btp_stage_pair(const PairArgs a) {
    using R = PairRec<G, Q>;
    using V = PV<NE>;
    typedef typename V::T VT;
    constexpr int NP = R::NP, NQ2 = R::NQ2, SX = R::SX, ST = R::ST, TM = R::TM;
    constexpr int NT = BLK ? 32 * W : 32;
    constexpr int NQIT = (NQ2 + NT - 1) / NT, NFIT = (4 * Q + NT - 1) / NT;
    static_assert(NP <= NT && 4 * G <= NT && 3 * Q <= NT && 12 <= NT && R::HDR <= NT, "polynomial order too high for this lane mapping");
    static_assert(!BLK || NE == 1, "block-per-element mode advances one element");
    // BLK: the independent jobs of a line phase are given to different warps of the block (different schedulers) instead of
    // one after the other to the lowest lanes: job lane offsets (0 in warp mode)
    constexpr int J2B = BLK ? 4 * G : 0, J2C = BLK ? 64 : 0;   // phase 2: bottom-layer rows, LDG gradient lines
    constexpr int J5A = BLK ? 64 : 0;                           // phase 5: the B.Fe + A.S contraction
    constexpr int J6B = BLK ? 32 : 0, J6C = BLK ? 32 + 4 * G : 0;   // phase 6: LDG laplacian lines, face traces
    // The 4 G lines of a gradient phase: in the warp form the row lines (along ksi) take the first half-warp and the column lines
    // (along eta) the second -- a row lane and a column lane of one half-warp hit the same bank with different words otherwise
    constexpr bool GSPLIT = !BLK && 2 * G <= 16;
    auto grad_lane = [](int lj) { return GSPLIT ? (lj >= 0 && lj < 32 && (lj & 15) < 2 * G) : (lj >= 0 && lj < 4 * G); };
    // (splitting scatter pass 2 into its B.TB and A.TA halves on two warps was measured neutral: 1.435 vs 1.422 ms)
    static_assert(!BLK || (J2B + 3 * G <= J2C && J2C + 4 * G <= NT && J5A + 3 * Q <= NT && 3 * Q <= J5A && 3 * G <= J6B && J6C + 4 * G <= NT),
                  "job lane ranges");
    extern __shared__ __align__(16) double sm_all[];
    const int lane = BLK ? (int)threadIdx.x : (int)(threadIdx.x & 31), warp = BLK ? 0 : (int)(threadIdx.x >> 5);
    const int unit = BLK ? (int)blockIdx.x : (int)(blockIdx.x * W + warp);
    if (unit * NE >= a.count) return;
    PR_STAMP(0);
    double* smw = sm_all + (size_t)warp * ((size_t)R::S_TOTAL * NE + (size_t)R::HDR * NE);
    double* hdr = smw;                                   // [NE][HDR], plain doubles
    VT* sv = reinterpret_cast<VT*>(smw + R::HDR * NE);
    VT* nod = sv + R::S_NOD;
    VT* X = sv + R::S_X;
    VT* T = sv + R::S_T;
    VT* Lr = sv + R::S_L;

    int e[NE];
    bool ok[NE];
    double* rec[NE];
    PR_FORC {
        e[c] = unit * NE + c;
        ok[c] = e[c] < a.count;
        if (!ok[c]) e[c] = a.count - 1;
        if (a.part == 1) e[c] = a.elist[e[c]];
        rec[c] = a.rec + (size_t)e[c] * R::REC;
    }
    constexpr bool botfr = BOTFR != 0;

    // ---- 1. nodal loads + nodal sums (mod_rk_mlswe.F90:90-92)
    double pbp[NE], pv[NE], bd[4][NE];   // kept by lane I < NP for the LDG flux variable and the update
    double dpp[NE], mx[NE], my[NE], oop[NE], pp[NE], up[NE], vp[NE];
    if (lane < NP) {
        const int I = lane;
        PR_FORC {
            const double* r = rec[c];
            dpp[c] = r[R::O_QB + I]; mx[c] = r[R::O_QB + NP + I]; my[c] = r[R::O_QB + 2 * NP + I];
            pbp[c] = r[R::O_NST + I]; oop[c] = r[R::O_NST + NP + I];
            if (botfr) { pp[c] = r[R::O_NST + 3 * NP + I]; up[c] = r[R::O_NST + 4 * NP + I]; vp[c] = r[R::O_NST + 5 * NP + I]; }
            if (VISC) {
                pv[c] = r[R::O_NST + 6 * NP + I];
#pragma unroll
                for (int k = 0; k < 4; ++k) bd[k][c] = r[R::O_NST + (7 + k) * NP + I];
            }
        }
    }
    double qs[NQIT][R::QST_NF][NE];
    // quadrature-point statics of all blocks of 32 points (6 per point and element): issued with the first loads, they
    // are consumed three phases later.  Lanes past the last point re-read it (branch-free phases; they never store sums).
    // (Requested one phase ahead instead -- after pass 1 -- the kernel is 1.2 % slower: 1.245 vs 1.231 ms per stage at 500x500 elements.)
#pragma unroll
    for (int it = 0; it < NQIT; ++it) {
        const int q = min(it * NT + lane, NQ2 - 1);
        PR_FORC {
            const double* rs = rec[c] + R::O_QST + q;
#pragma unroll
            for (int k = 0; k < R::QST_NF; ++k) qs[it][k][c] = rs[k * NQ2];
        }
    }
    // ---- 0. header -> shared memory (issued after the nodal loads so that the two round trips overlap); the tail of
    //          the record -> L2 with one bulk prefetch per element
    {
        double hv[NE];
        if (lane < R::HDR) { PR_FORC hv[c] = rec[c][lane]; }
        if (a.prefetch && lane < NE) {
            const double* p = (lane == 0) ? rec[0] : rec[NE - 1];
            if (a.prefetch & 1) {
                pr_prefetch_l2(p + R::O_QST, (uint32_t)((R::O_Q0 - R::O_QST) * sizeof(double)));
                if (a.load_q0) pr_prefetch_l2(p + R::O_Q0, (uint32_t)(R::QBSZ * sizeof(double)));
                if (a.load_q2) pr_prefetch_l2(p + R::O_Q2, (uint32_t)(R::QBSZ * sizeof(double)));
            }
            // the head of the record (header, state, nodal statics and sums) of the unit that follows one wave later
            if ((a.prefetch & 2) && a.part != 1 && (long)(unit + a.pf_dist) * NE + lane < a.nelem)
                pr_prefetch_l2(p + (size_t)a.pf_dist * NE * R::REC, (uint32_t)(R::O_QST * sizeof(double)));
            // (also prefetching its face coefficients, its quadrature running sums or its trace records was measured slower:
            //  1.44-1.51 against 1.43 ms per stage at 500x500 elements, profiles/r1_stage_kernel_experiments.md)
            // ... and its quadrature statics (bit 2)
            if ((a.prefetch & 4) && a.part != 1 && (long)(unit + a.pf_dist) * NE + lane < a.nelem)
                pr_prefetch_l2(p + (size_t)a.pf_dist * NE * R::REC + R::O_QST, (uint32_t)((R::O_ACCQ - R::O_QST) * sizeof(double)));
        }
        if (lane < R::HDR) { PR_FORC hdr[c * R::HDR + lane] = hv[c]; }
        if (a.part == 2) {   // interior launch: elements with a processor face belong to the boundary launch
            bool any = false;
            PR_FORC {
                bool hb = false;
                if (lane == 18 || lane == 19) {
                    const int lo = __double2loint(hv[c]), hi_ = __double2hiint(hv[c]);
                    hb = (lo == NBR_HALO) || (hi_ == NBR_HALO);
                }
                if (BLK ? __syncthreads_or(hb) : (__ballot_sync(0xffffffffu, hb) != 0u)) ok[c] = false;
                any = any || ok[c];
            }
            if (!any) return;
        }
    }
    if (lane < NP) {
        const int I = lane;
        double pb[NE], u[NE], v[NE];
        PR_FORC {
            pb[c] = dpp[c] + pbp[c];
            const double rpb = pr_rcp(pb[c]);
            u[c] = mx[c] * rpb; v[c] = my[c] * rpb;
            const double t = 1.0 + dpp[c] * oop[c];
            if (ok[c]) {
                double* r = rec[c];
                pr_red(r + R::O_ACCN + I, t * t); pr_red(r + R::O_ACCN + NP + I, u[c]); pr_red(r + R::O_ACCN + 2 * NP + I, v[c]);
                pr_red(r + R::O_ACCN + 3 * NP + I, dpp[c]); pr_red(r + R::O_ACCN + 4 * NP + I, mx[c]); pr_red(r + R::O_ACCN + 5 * NP + I, my[c]);
            }
        }
        V::st(nod + 0 * NP + I, dpp); V::st(nod + 1 * NP + I, mx); V::st(nod + 2 * NP + I, my); V::st(nod + 3 * NP + I, pb);
        if (botfr) { V::st(nod + 4 * NP + I, pp); V::st(nod + 5 * NP + I, up); V::st(nod + 6 * NP + I, vp); }
        V::st(nod + 7 * NP + I, u); V::st(nod + 8 * NP + I, v);
    }
    pr_sync<NT>();
    PR_STAMP(1);
    // per-element geometry and flags (registers)
    double ksx[NE], ksy[NE], etx[NE], ety[NE], J[NE];
    int flags[NE];
    PR_FORC {
        const double* h = hdr + c * R::HDR;
        ksx[c] = h[0]; ksy[c] = h[1]; etx[c] = h[2]; ety[c] = h[3]; J[c] = h[4];
        flags[c] = reinterpret_cast<const int*>(h + 22)[0];
    }
    // the rarely-read forcing fields (tau_y, grad z_bot) of the record pf_dist units ahead -> L2 when this element has them
    // (neighbouring elements are alike): without it every quadrature point pays two dependent DRAM loads in phase 4
    if ((a.prefetch & 4) && a.part != 1 && lane < NE) {
        PR_FORC {
            if (lane == c && (flags[c] & (PF_GZ | PF_TWY)) && (long)(unit + a.pf_dist) * NE + c < a.nelem)
                pr_prefetch_l2(rec[c] + (size_t)a.pf_dist * NE * R::REC + R::O_QSTR, (uint32_t)(pr_pad2(R::QST_RARE * NQ2) * sizeof(double)));
        }
    }
    // neighbour traces and owned face sums -> L2 (the header has just told us where they are): one 128-byte line per lane
    if (a.prefetch & 1) {
        constexpr int TL = (R::TSIDE * 8 + 127) / 128 + 1, AL = (R::ASIDE * 8 + 127) / 128 + 1;
        PR_FORC {
            const int* hi = reinterpret_cast<const int*>(hdr + c * R::HDR + 18);
            if (lane < 4 * TL) {
                const int s = lane / TL, ln = lane - s * TL, tr = hi[4 + s];
                const char* p = reinterpret_cast<const char*>(a.tr_in + (size_t)(tr >= 0 ? tr : 0) * R::TSIDE);
                const char* q = reinterpret_cast<const char*>((uintptr_t)p & ~(uintptr_t)127) + ln * 128;
                if (tr >= 0 && q < p + R::TSIDE * 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
            }
            if (lane < 4 * AL) {
                const int s = lane / AL, ln = lane - s * AL, nb = hi[s];
                const char* p = reinterpret_cast<const char*>(a.accf + ((size_t)e[c] * 4 + s) * R::ASIDE);
                const char* q = reinterpret_cast<const char*>((uintptr_t)p & ~(uintptr_t)127) + ln * 128;
                if ((nb < 0 || e[c] < nb) && q < p + R::ASIDE * 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
            }
        }
    }
    PR_STAMP(12);
    // ---- 2. sum factorisation, pass 1: one nodal row per lane -> T[f][m][i]; LDG gradient lines -> Lr[0..3]
    {
        if (lane < 4 * G) {
            const int f = lane / G, m = lane - f * G;
            pl_n2q<NE, G, Q, false, 1, 1>(nod + f * NP + m * G, T + f * ST + m * TM);
        }
        if (botfr && lane >= J2B && lane < J2B + 3 * G) {
            const int lj = lane - J2B, f = 4 + lj / G, m = lj % G;
            pl_n2q<NE, G, Q, false, 1, 1>(nod + f * NP + m * G, T + f * ST + m * TM);
        }
        if (VISC && grad_lane(lane - J2C)) {
            const int lj = lane - J2C, kind = GSPLIT ? lj >> 4 : lj / (2 * G), r = GSPLIT ? (lj & 15) : lj - kind * 2 * G, f = r / G, l = r - f * G;
            const int stride = kind ? G : 1, off = kind ? l : l * G;
            pl_grad<NE, G, false>(nod + (7 + f) * NP + off, Lr + (2 * kind + f) * NP + off, stride);
        }
    }
    pr_sync<NT>();
    PR_STAMP(2);
    // ---- 3. pass 2: one quadrature column per lane -> X[f][j][i]; LDG: G = grad(ub,vb), flux variable, weighted metric terms
    {
        const int nlines = (botfr ? 7 : 4) * Q;
#pragma unroll 1
        for (int it = lane; it < nlines; it += NT) {
            const int f = it / Q, i = it - f * Q;
            pl_n2q<NE, G, Q, false, TM, Q>(T + f * ST + i, X + f * SX + i);
        }
        if (VISC && lane < NP) {
            const int I = lane, m = I / G, n = I - m * G;
            double dku[NE], dkv[NE], deu[NE], dev[NE];
            V::ld(Lr + 0 * NP + I, dku); V::ld(Lr + 1 * NP + I, dkv); V::ld(Lr + 2 * NP + I, deu); V::ld(Lr + 3 * NP + I, dev);
            const double w0 = c_ops.wg[n] * c_ops.wg[m];
            double g0[NE], g1[NE], g2[NE], g3[NE], q0[NE], q1[NE], q2[NE], q3[NE], z0[NE], z1[NE], z2[NE], z3[NE];
            PR_FORC {
                g0[c] = ksx[c] * dku[c] + etx[c] * deu[c]; g1[c] = ksy[c] * dku[c] + ety[c] * deu[c];
                g2[c] = ksx[c] * dkv[c] + etx[c] * dev[c]; g3[c] = ksy[c] * dkv[c] + ety[c] * dev[c];
                q0[c] = pv[c] * g0[c] + bd[0][c]; q1[c] = pv[c] * g1[c] + bd[1][c];
                q2[c] = pv[c] * g2[c] + bd[2][c]; q3[c] = pv[c] * g3[c] + bd[3][c];
                const double w = w0 * J[c];
                z0[c] = w * (ksx[c] * q0[c] + ksy[c] * q1[c]); z1[c] = w * (ksx[c] * q2[c] + ksy[c] * q3[c]);
                z2[c] = w * (etx[c] * q0[c] + ety[c] * q1[c]); z3[c] = w * (etx[c] * q2[c] + ety[c] * q3[c]);
            }
            V::st(nod + 4 * NP + I, q0); V::st(nod + 5 * NP + I, q1); V::st(nod + 6 * NP + I, q2); V::st(nod + 7 * NP + I, q3);
            V::st(Lr + 8 * NP + I, z0); V::st(Lr + 9 * NP + I, z1); V::st(Lr + 10 * NP + I, z2); V::st(Lr + 11 * NP + I, z3);
        }
    }
    pr_sync<NT>();
    PR_STAMP(3);
    // ---- 4. pointwise physics at the quadrature points (mod_rhs_btp.F90:136-192); the fluxes overwrite X in place
#pragma unroll
    for (int it = 0; it < NQIT; ++it) {
        const int q = min(it * NT + lane, NQ2 - 1);
        const bool qvalid = (it + 1) * NT <= NQ2 || it * NT + lane < NQ2;
        {
            const int j = q / Q, i = q - j * Q;
            const double w0 = c_ops.wq[i] * c_ops.wq[j];
            double dpp[NE], udp[NE], vdp[NE], dp[NE], pp[NE], up[NE], vp[NE];
            V::ld(X + 0 * SX + q, dpp); V::ld(X + 1 * SX + q, udp); V::ld(X + 2 * SX + q, vdp); V::ld(X + 3 * SX + q, dp);
            if (botfr) { V::ld(X + 4 * SX + q, pp); V::ld(X + 5 * SX + q, up); V::ld(X + 6 * SX + q, vp); }
            double o0[NE], o1[NE], o2[NE], o3[NE], o4[NE], o5[NE], o6[NE], o7[NE];
            PR_FORC {
                const double wq = w0 * J[c];
                const double rdp = pr_rcp(dp[c]);
                const double ub = udp[c] * rdp, vb = vdp[c] * rdp;
                double tb_u = 0.0, tb_v = 0.0;
                if (botfr) {
                    const double ubot = up[c] + ub, vbot = vp[c] + vb;
                    const double spd = (BOTFR == 1) ? a.cd_g * pp[c] : a.cd_alpha * sqrt(ubot * ubot + vbot * vbot);
                    tb_u = spd * ubot; tb_v = spd * vbot;
                }
                const double pbq = dp[c] - dpp[c];             // pbprime at the point: pb = pbpert + pbprime, all interpolated
                const double s_oop = pbq > 0.0 ? pr_rcp(pbq) : 0.0;   // guarded as in initial_conditions.F90:344-348
                const double s_H = qs[it][0][c], s_uu = qs[it][1][c], s_uv = qs[it][2][c], s_vv = qs[it][3][c];
                const double fcor = qs[it][4][c], s_twx = qs[it][5][c];
                // rare forcing fields are read on demand (predicated loads, no branch)
                const double* rs = rec[c] + R::O_QSTR + q;
                const double s_twy = (flags[c] & PF_TWY) ? rs[0] : 0.0;
                const double s_gzx = (flags[c] & PF_GZ) ? rs[NQ2] : 0.0, s_gzy = (flags[c] & PF_GZ) ? rs[2 * NQ2] : 0.0;
                const double sc_x = fcor * vdp[c] + a.g * (s_twx - tb_u) - a.g * dp[c] * s_gzx;
                const double sc_y = -fcor * udp[c] + a.g * (s_twy - tb_v) - a.g * dp[c] * s_gzy;
                const double ope = 1.0 + dpp[c] * s_oop;
                const double ope2 = ope * ope;
                const double Hq = ope2 * s_H;
                const double qu = ub * udp[c] + ope * s_uu;
                const double quv = ub * vdp[c] + ope * s_uv;
                const double qv = vb * vdp[c] + ope * s_vv;
                if (qvalid && ok[c]) {
                    double* ra = rec[c] + R::O_ACCQ + q;
                    pr_red(ra, qu); pr_red(ra + NQ2, qv); pr_red(ra + 2 * NQ2, quv);
                    pr_red(ra + 3 * NQ2, ope2); pr_red(ra + 4 * NQ2, ub); pr_red(ra + 5 * NQ2, vb);
                    if (BOTFR == 2) { double* rr_ = rec[c] + R::O_ACCQR + q; pr_red(rr_, tb_u); pr_red(rr_ + NQ2, tb_v); }
                }
                const double Fx2 = Hq + qu, Fy3 = Hq + qv;
                o0[c] = wq * (ksx[c] * udp[c] + ksy[c] * vdp[c]);   // Fk1
                o1[c] = wq * (etx[c] * udp[c] + ety[c] * vdp[c]);   // Fe1
                o2[c] = wq * sc_x;                                  // S2
                o3[c] = wq * (ksx[c] * Fx2 + ksy[c] * quv);         // Fk2
                o4[c] = wq * (etx[c] * Fx2 + ety[c] * quv);         // Fe2
                o5[c] = wq * sc_y;                                  // S3
                o6[c] = wq * (ksx[c] * quv + ksy[c] * Fy3);         // Fk3
                o7[c] = wq * (etx[c] * quv + ety[c] * Fy3);         // Fe3
            }
            // array order in X: Fk1 Fk2 Fk3 | Fe1 Fe2 Fe3 | S2 S3  (field f of a kind at a constant offset: conflict-free scatter)
            // (lanes past the last point work on a copy of it: within one warp their loads precede every store, across the
            //  warps of a block they would race with the in-place update of that point)
            if (NT == 32 || qvalid) {
                V::st(X + 0 * SX + q, o0); V::st(X + 3 * SX + q, o1); V::st(X + 6 * SX + q, o2); V::st(X + 1 * SX + q, o3);
                V::st(X + 4 * SX + q, o4); V::st(X + 7 * SX + q, o5); V::st(X + 2 * SX + q, o6); V::st(X + 5 * SX + q, o7);
            }
        }
    }
    // neighbour traces (state and LDG flux variable) and the neighbour's pbprime of face node (s,n): requested after scatter pass 1
    // (right behind the 18 REDs of the pointwise phase they queued up in the LSU: -0.5 % of the stage time when issued one phase later)
    double tn[7][NE], pbnp[NE];
    const int lf = lane - J6C;                     // face node owned by this lane (0 <= lf < 4G)
    const bool face_lane = lf >= 0 && lf < 4 * G;
    const int fs = face_lane ? lf / G : 0, fn = face_lane ? lf - fs * G : 0;
    pr_sync<NT>();
    PR_STAMP(4);
    // ---- 5. scatter pass 1 (contraction over j): lane (f,i) -> TB_f = A.Fk_f, TA_f = B.Fe_f + A.S_f, as T[f][m][i], T[3+f][m][i]
    if (lane < 3 * Q) {
        const int f = lane / Q, i = lane - f * Q;
        double tb[G][NE];
        pl_q2n_acc<NE, G, Q, false, Q, true>(X + f * SX + i, tb);
#pragma unroll
        for (int m = 0; m < G; ++m) V::st(T + f * ST + m * TM + i, tb[m]);
    }
    if (lane >= J5A && lane < J5A + 3 * Q) {
        const int lj = lane - J5A, f = lj / Q, i = lj - f * Q;
        double ta[G][NE];
        pl_q2n_acc<NE, G, Q, true, Q, true>(X + (3 + f) * SX + i, ta);
        if (f > 0) pl_q2n_acc<NE, G, Q, false, Q, false>(X + (5 + f) * SX + i, ta);
#pragma unroll
        for (int m = 0; m < G; ++m) V::st(T + (3 + f) * ST + m * TM + i, ta[m]);
    }
    pr_sync<NT>();
    PR_STAMP(5);
    if (face_lane) {
        PR_FORC {
            const int tr = reinterpret_cast<const int*>(hdr + c * R::HDR + 18)[4 + fs];
            if (tr >= 0) {
                const double* t = a.tr_in + (size_t)tr * R::TSIDE + fn;
#pragma unroll
                for (int k = 0; k < (VISC ? 7 : 3); ++k) tn[k][c] = t[k * G];
            }
            pbnp[c] = rec[c][R::O_VST + fs * R::VSIDE + fn];
        }
    }


    // ---- 6. scatter pass 2 (contraction over i): lane (f,m) -> rhs[f][m][n] = B.TB_f + A.TA_f
    if (lane < 3 * G) {
        const int f = lane / G, m = lane - f * G;
        double r[G][NE];
        pl_q2n_acc<NE, G, Q, true, 1, true>(T + f * ST + m * TM, r);
        pl_q2n_acc<NE, G, Q, false, 1, false>(T + (3 + f) * ST + m * TM, r);
#pragma unroll
        for (int n = 0; n < G; ++n) V::st(X + R::X_RHS + f * NP + m * G + n, r[n]);
    }
    // LDG volume lines (btp_compute_laplacian): lapX[c][m'][n'] = sum_n D(n',n) Zxi_c[m'][n]; lapE[c][m'][n'] = sum_m D(m',m) Zeta_c[m][n']
    if (VISC && grad_lane(lane - J6B)) {
        const int lj = lane - J6B, kind = GSPLIT ? lj >> 4 : lj / (2 * G), r = GSPLIT ? (lj & 15) : lj - kind * 2 * G, cc = r / G, l = r - cc * G;
        const int stride = kind ? G : 1, off = kind ? l : l * G;
        pl_grad<NE, G, true>(Lr + (8 + 2 * kind + cc) * NP + off, Lr + (2 * kind + cc) * NP + off, stride);
    }
    // ---- 7a. face traces: own and neighbour state in canonical (left, right) order; LDG face flux at the face nodes,
    //          as written (btp_extract_df; mod_laplacian_quad.F90:85-98,427-519)
    if (face_lane) {
        const int s = fs, n = fn, I = pr_face_node<G>(s, n);
        double ow0[NE], ow1[NE], ow2[NE], pbo[NE];
        V::ld(nod + 0 * NP + I, ow0); V::ld(nod + 1 * NP + I, ow1); V::ld(nod + 2 * NP + I, ow2); V::ld(nod + 3 * NP + I, pbo);
        double qo[4][NE];
        if (VISC) {
#pragma unroll
            for (int k = 0; k < 4; ++k) V::ld(nod + (4 + k) * NP + I, qo[k]);
        }
        double L0[NE], L1[NE], L2[NE], L3[NE], R0[NE], R1[NE], R2[NE], R3[NE], lfu[NE], lfv[NE];
        const double wgn = c_ops.wg[n];
        PR_FORC {
            const double* h = hdr + c * R::HDR;
            const int* hi = reinterpret_cast<const int*>(h + 18);
            const int nb = hi[s], tr = hi[4 + s];
            const double nx = h[6 + s * 3 + 0], ny = h[6 + s * 3 + 1], nlen = h[6 + s * 3 + 2];
            const bool left = (nb < 0) || (e[c] < nb);
            double n0, n1, n2;
            if (tr >= 0) { n0 = tn[0][c]; n1 = tn[1][c]; n2 = tn[2][c]; }
            else {
                n0 = ow0[c]; n1 = ow1[c]; n2 = ow2[c];
                if (nb == NBR_FREESLIP) { const double un = nx * ow1[c] + ny * ow2[c]; n1 = ow1[c] - 2.0 * un * nx; n2 = ow2[c] - 2.0 * un * ny; }
                else if (nb == NBR_NOSLIP) { n1 = -ow1[c]; n2 = -ow2[c]; }
            }
            const double pbn = n0 + pbnp[c];
            L0[c] = left ? pbo[c] : pbn; R0[c] = left ? pbn : pbo[c];
            L1[c] = left ? ow0[c] : n0;  R1[c] = left ? n0 : ow0[c];
            L2[c] = left ? ow1[c] : n1;  R2[c] = left ? n1 : ow1[c];
            L3[c] = left ? ow2[c] : n2;  R3[c] = left ? n2 : ow2[c];
            if (VISC) {
                // the other side's flux variable: published by the neighbour, or -- at a wall -- the ghost of the own one: gradient and
                // statics of the ghost are the mirror images of the own ones (mod_laplacian_quad.F90:85-98), and the mirror is linear
                double fo_[4], fn_[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { fo_[k] = qo[k][c]; fn_[k] = (tr >= 0) ? tn[3 + k][c] : qo[k][c]; }
                if (tr < 0 && nb == NBR_FREESLIP) {
                    double un = qo[0][c] * nx + qo[1][c] * ny;
                    fn_[0] = qo[0][c] - 2.0 * un * nx; fn_[1] = qo[1][c] - 2.0 * un * ny;
                    un = qo[2][c] * nx + qo[3][c] * ny;
                    fn_[2] = qo[2][c] - 2.0 * un * nx; fn_[3] = qo[3][c] - 2.0 * un * ny;
                }
                double fl[4], fr[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { fl[k] = left ? fo_[k] : fn_[k]; fr[k] = left ? fn_[k] : fo_[k]; }
                const double qu0 = 0.5 * fl[0] + 0.5 * fr[0], qu1 = 0.5 * fl[1] + 0.5 * fr[1];
                const double qv0 = 0.5 * fl[2] + 0.5 * fr[2], qv1 = 0.5 * fl[3] + 0.5 * fr[3];
                const double wq = wgn * nlen;
                const double flux_qu = (qu0 - fl[0] * nx) + (qu1 - fl[1] * ny);
                const double flux_qv = (qv0 - fl[2] * nx) + (qv1 - fl[3] * ny);
                const double sgn = left ? wq : -wq;
                lfu[c] = sgn * flux_qu; lfv[c] = sgn * flux_qv;
            }
        }
        VT* FL = X + R::X_FL + lf;   // [side][var][s][n]: the word index follows the lane number
        VT* FR = X + R::X_FR + lf;
        V::st(FL, L0); V::st(FL + 4 * G, L1); V::st(FL + 8 * G, L2); V::st(FL + 12 * G, L3);
        V::st(FR, R0); V::st(FR + 4 * G, R1); V::st(FR + 8 * G, R2); V::st(FR + 12 * G, R3);
        if (VISC) { V::st(X + R::X_LF + (s * 2 + 0) * G + n, lfu); V::st(X + R::X_LF + (s * 2 + 1) * G + n, lfv); }
    }
    // face coefficients of all face quadrature points (11 per point and element), loaded one phase ahead
    double fc[NFIT][11][NE];
#pragma unroll
    for (int it = 0; it < NFIT; ++it) {
        const int p = min(it * NT + lane, 4 * Q - 1);
        {
            const int s = p / Q, iq = p - s * Q;
            PR_FORC {
                // a face's coefficients live once, in the record of its left (owner) element: the other side reads them
                // there (its neighbour is in flight at the same time, so this is an L2 hit, not a second DRAM read)
                const int* hi = reinterpret_cast<const int*>(hdr + c * R::HDR + 18);
                const int nb = hi[s];
                const bool left = (nb < 0) || (e[c] < nb);
                const double* cf = left ? rec[c] + R::O_FST + s * R::FSIDE + iq
                                        : a.rec + (size_t)nb * R::REC + R::O_FST + (hi[4 + s] - 4 * nb) * R::FSIDE + iq;
#pragma unroll
                for (int k = 0; k < 9; ++k) fc[it][k][c] = cf[k * Q];
                fc[it][9][c] = left ? cf[9 * Q] : 0.0;
                fc[it][10][c] = left ? cf[10 * Q] : 0.0;
            }
        }
    }
    pr_sync<NT>();
    PR_STAMP(6);
    // ---- 7b. interpolate the traces to the face quadrature points: one (side, L/R, variable) line per lane
    if (NT == 32 || lane < 32) {
        const int s = lane >> 3, side = (lane >> 2) & 1, var = lane & 3;   // var: 0 pb 1 pbpert 2 mx 3 my
        pl_n2q<NE, G, Q, false, 1, 1>(X + (side ? R::X_FR : R::X_FL) + (var * 4 + s) * G, T + s * R::TS7 + (side * 4 + var) * R::TR7);
    }
    // update operands of node I, one phase ahead
    double mi[NE], q0v[3][NE], q2v[3][NE];
    if (lane < NP) {
        PR_FORC {
            const double* r = rec[c];
            mi[c] = r[R::O_NST + 2 * NP + lane];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                q0v[k][c] = a.load_q0 ? r[R::O_Q0 + k * NP + lane] : 0.0;
                q2v[k][c] = a.load_q2 ? r[R::O_Q2 + k * NP + lane] : 0.0;
            }
        }
    }
    pr_sync<NT>();
    PR_STAMP(7);
    // ---- 7c. face fluxes, canonical left perspective (mod_rhs_btp.F90:237-330)
#pragma unroll
    for (int it = 0; it < NFIT; ++it) {
        const int p = min(it * NT + lane, 4 * Q - 1);
        const bool pvalid = (it + 1) * NT <= 4 * Q || it * NT + lane < 4 * Q;
        {
            const int s = p / Q, iq = p - s * Q;
            constexpr int TR7 = R::TR7;
            const VT* Lp = T + s * R::TS7 + iq;
            const VT* Rp = Lp + 4 * TR7;
            double pbL[NE], ppL[NE], mxL[NE], myL[NE], pbR[NE], ppR[NE], mxR[NE], myR[NE];
            V::ld(Lp, pbL); V::ld(Lp + TR7, ppL); V::ld(Lp + 2 * TR7, mxL); V::ld(Lp + 3 * TR7, myL);
            V::ld(Rp, pbR); V::ld(Rp + TR7, ppR); V::ld(Rp + 2 * TR7, mxR); V::ld(Rp + 3 * TR7, myR);
            const double wq0 = c_ops.wq[iq];
            double f0[NE], f1[NE], f2[NE];
            PR_FORC {
                const double* h = hdr + c * R::HDR;
                const int nb = reinterpret_cast<const int*>(h + 18)[s];
                const bool left = (nb < 0) || (e[c] < nb);
                const double nxl = h[6 + s * 3 + 0], nyl = h[6 + s * 3 + 1], nlen = h[6 + s * 3 + 2];
                const double cL = fc[it][0][c], cR = fc[it][1][c], cLR = fc[it][2][c], lam = fc[it][3][c];
                const double s_oope = fc[it][4][c], s_uue = fc[it][5][c], s_uve = fc[it][6][c], s_vve = fc[it][7][c], s_He = fc[it][8][c];
                const double pU_L = nxl * mxL[c] + nyl * myL[c];
                const double pU_R = -nxl * mxR[c] - nyl * myR[c];
                const double pbpert_edge = cL * ppL[c] + cR * ppR[c] + cLR * (pU_L + pU_R);
                const double ope_e = 1.0 + pbpert_edge * s_oope;
                const double fex = cR * mxL[c] + cL * mxR[c] + lam * (nxl * ppL[c] - nxl * ppR[c]);
                const double fey = cR * myL[c] + cL * myR[c] + lam * (nyl * ppL[c] - nyl * ppR[c]);
                const double rl = pr_rcp(pbL[c]), rr = pr_rcp(pbR[c]);
                const double ul = mxL[c] * rl, ur = mxR[c] * rr, vl = myL[c] * rl, vr = myR[c] * rr;
                const double quu = 0.5 * (ul * mxL[c] + ur * mxR[c]) + ope_e * s_uue;
                const double quv = 0.5 * (vl * mxL[c] + vr * mxR[c]) + ope_e * s_uve;
                const double qvu = 0.5 * (ul * myL[c] + ur * myR[c]) + ope_e * s_uve;
                const double qvv = 0.5 * (vl * myL[c] + vr * myR[c]) + ope_e * s_vve;
                const double e2 = ope_e * ope_e;
                const double Hf = e2 * s_He;
                if (left && pvalid && ok[c]) {
                    const double ol = 1.0 + ppL[c] * fc[it][9][c], orr = 1.0 + ppR[c] * fc[it][10][c];
                    double* af = a.accf + ((size_t)e[c] * 4 + s) * R::ASIDE + iq;
                    pr_red(af, quu); pr_red(af + Q, quv); pr_red(af + 2 * Q, qvu); pr_red(af + 3 * Q, qvv);
                    pr_red(af + 4 * Q, ol * ol); pr_red(af + 5 * Q, orr * orr); pr_red(af + 6 * Q, e2);
                    pr_red(af + 7 * Q, ul); pr_red(af + 8 * Q, ur); pr_red(af + 9 * Q, vl); pr_red(af + 10 * Q, vr);
                }
                const double wq = wq0 * nlen;
                const double dispu = 0.5 * lam * (mxR[c] - mxL[c]), dispv = 0.5 * lam * (myR[c] - myL[c]);
                const double flux_x = nxl * quu + nyl * quv - dispu;
                const double flux_y = nxl * qvu + nyl * qvv - dispv;
                const double flux = nxl * fex + nyl * fey;
                const double sgn = left ? -wq : wq;
                f0[c] = sgn * flux; f1[c] = sgn * (nxl * Hf + flux_x); f2[c] = sgn * (nyl * Hf + flux_y);
            }
            VT* ff = X + R::X_FF + s * R::FS7 + iq;
            V::st(ff, f0); V::st(ff + R::FR7, f1); V::st(ff + 2 * R::FR7, f2);
        }
    }
    pr_sync<NT>();
    PR_STAMP(8);
    // ---- 7d. project the face fluxes onto the face nodes: one (side, field) line per lane -> T[0 .. 12G)
    if (lane < 12) {
        double pr[G][NE];
        pl_q2n_acc<NE, G, Q, false, 1, true>(X + R::X_FF + (lane / 3) * R::FS7 + (lane % 3) * R::FR7, pr);
#pragma unroll
        for (int n = 0; n < G; ++n) V::st(T + lane * G + n, pr[n]);
    }
    pr_sync<NT>();
    PR_STAMP(9);
    // ---- 8. gather per node, mass matrix, viscosity, SSPRK update, wall projection (mod_rk_mlswe.F90:97-108)
    if (lane < NP) {
        const int I = lane, m = I / G, n = I - m * G;
        double r0[NE], r1[NE], r2[NE], l0[NE], l1[NE];
        V::ld(X + R::X_RHS + 0 * NP + I, r0); V::ld(X + R::X_RHS + 1 * NP + I, r1); V::ld(X + R::X_RHS + 2 * NP + I, r2);

        if (VISC) {
            double t0[NE], t1[NE], t2[NE], t3[NE];
            V::ld(Lr + 0 * NP + I, t0); V::ld(Lr + 1 * NP + I, t1); V::ld(Lr + 2 * NP + I, t2); V::ld(Lr + 3 * NP + I, t3);
            PR_FORC { l0[c] = -(t0[c] + t2[c]); l1[c] = -(t1[c] + t3[c]); }
        }
        // a node lies on at most one eta-side (s = 0, 1) and one ksi-side (s = 2, 3): two gathers with a per-lane side instead of
        // four predicated ones (same additions in the same order: sides in ascending order)
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const int s = (pass == 0) ? ((m == 0) ? 0 : (m == G - 1) ? 1 : -1) : ((n == 0) ? 2 : (n == G - 1) ? 3 : -1);
            if (s < 0) continue;
            const int nf = (pass == 0) ? n : m;
            double p0[NE], p1[NE], p2[NE];
            V::ld(T + (s * 3 + 0) * G + nf, p0); V::ld(T + (s * 3 + 1) * G + nf, p1); V::ld(T + (s * 3 + 2) * G + nf, p2);
            PR_FORC { r0[c] += p0[c]; r1[c] += p1[c]; r2[c] += p2[c]; }
            if (VISC) {
                double a0[NE], a1[NE];
                V::ld(X + R::X_LF + (s * 2 + 0) * G + nf, a0); V::ld(X + R::X_LF + (s * 2 + 1) * G + nf, a1);
                PR_FORC { l0[c] += a0[c]; l1[c] += a1[c]; }
            }
        }
        double q1a[NE], q1b[NE], q1c[NE];
        V::ld(nod + 0 * NP + I, q1a); V::ld(nod + 1 * NP + I, q1b); V::ld(nod + 2 * NP + I, q1c);
        double nw0[NE], nw1[NE], nw2[NE], un_[NE], vn_[NE];
        PR_FORC {
            double rr[3] = {mi[c] * r0[c], mi[c] * r1[c], mi[c] * r2[c]};
            if (VISC) { rr[1] = rr[1] + a.visc * mi[c] * l0[c]; rr[2] = rr[2] + a.visc * mi[c] * l1[c]; }
            const double q1[3] = {q1a[c], q1b[c], q1c[c]};
            double qn[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double q0 = a.load_q0 ? q0v[k][c] : q1[k];
                qn[k] = a.a1 * q0 + a.a2 * q1[k] + a.a3 * q2v[k][c] + a.dtt * rr[k];
            }
            const double* h = hdr + c * R::HDR;
            const int* hi = reinterpret_cast<const int*>(h + 18);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const bool on = (s == 0) ? (m == 0) : (s == 1) ? (m == G - 1) : (s == 2) ? (n == 0) : (n == G - 1);
                if (!on) continue;
                const int nb = a.rhs_only ? 0 : hi[s];
                if (nb == NBR_FREESLIP) {
                    const double nx = h[6 + s * 3 + 0], ny = h[6 + s * 3 + 1];
                    const double unl = qn[1] * nx + qn[2] * ny;
                    qn[1] = qn[1] - unl * nx; qn[2] = qn[2] - unl * ny;
                } else if (nb == NBR_NOSLIP) { qn[1] = 0.0; qn[2] = 0.0; }
            }
            if (ok[c]) {
                double* r = rec[c];
                if (a.store_q0) { r[R::O_Q0 + I] = q1[0]; r[R::O_Q0 + NP + I] = q1[1]; r[R::O_Q0 + 2 * NP + I] = q1[2]; }
                r[R::O_QB + I] = qn[0]; r[R::O_QB + NP + I] = qn[1]; r[R::O_QB + 2 * NP + I] = qn[2];
                if (a.store_q2) { r[R::O_Q2 + I] = qn[0]; r[R::O_Q2 + NP + I] = qn[1]; r[R::O_Q2 + 2 * NP + I] = qn[2]; }
            }
            const double rpb = pr_rcp(qn[0] + pbp[c]);
            nw0[c] = qn[0]; nw1[c] = qn[1]; nw2[c] = qn[2];
            un_[c] = qn[1] * rpb; vn_[c] = qn[2] * rpb;
        }
        V::st(nod + 0 * NP + I, nw0); V::st(nod + 1 * NP + I, nw1); V::st(nod + 2 * NP + I, nw2);
        V::st(nod + 7 * NP + I, un_); V::st(nod + 8 * NP + I, vn_);
    }
    pr_sync<NT>();
    PR_STAMP(10);
    // ---- 9. traces of the new state (+ LDG gradient) for the next stage
    if (VISC) {
        if (grad_lane(lane)) {
            const int kind = GSPLIT ? lane >> 4 : lane / (2 * G), r = GSPLIT ? (lane & 15) : lane - kind * 2 * G, f = r / G, l = r - f * G;
            const int stride = kind ? G : 1, off = kind ? l : l * G;
            pl_grad<NE, G, false>(nod + (7 + f) * NP + off, Lr + (2 * kind + f) * NP + off, stride);
        }
        pr_sync<NT>();
        PR_STAMP(11);
        // flux variable of the new state at every node (the lanes that hold pbprime_visc and btp_dpp_graduv) -> Lr[4..7]
        if (lane < NP) {
            const int I = lane;
            double dku[NE], dkv[NE], deu[NE], dev[NE], q0[NE], q1[NE], q2[NE], q3[NE];
            V::ld(Lr + 0 * NP + I, dku); V::ld(Lr + 1 * NP + I, dkv); V::ld(Lr + 2 * NP + I, deu); V::ld(Lr + 3 * NP + I, dev);
            PR_FORC {
                const double g0 = ksx[c] * dku[c] + etx[c] * deu[c], g1 = ksy[c] * dku[c] + ety[c] * deu[c];
                const double g2 = ksx[c] * dkv[c] + etx[c] * dev[c], g3 = ksy[c] * dkv[c] + ety[c] * dev[c];
                q0[c] = pv[c] * g0 + bd[0][c]; q1[c] = pv[c] * g1 + bd[1][c];
                q2[c] = pv[c] * g2 + bd[2][c]; q3[c] = pv[c] * g3 + bd[3][c];
            }
            V::st(Lr + 4 * NP + I, q0); V::st(Lr + 5 * NP + I, q1); V::st(Lr + 6 * NP + I, q2); V::st(Lr + 7 * NP + I, q3);
        }
        pr_sync<NT>();
    }
    if (face_lane) {
        const int s = fs, n = fn, I = pr_face_node<G>(s, n);
        double t0[NE], t1[NE], t2[NE], q0[NE], q1[NE], q2[NE], q3[NE];
        V::ld(nod + 0 * NP + I, t0); V::ld(nod + 1 * NP + I, t1); V::ld(nod + 2 * NP + I, t2);
        if (VISC) { V::ld(Lr + 4 * NP + I, q0); V::ld(Lr + 5 * NP + I, q1); V::ld(Lr + 6 * NP + I, q2); V::ld(Lr + 7 * NP + I, q3); }
        PR_FORC {
            if (ok[c]) {
                double* to = a.tr_out + ((size_t)e[c] * 4 + s) * R::TSIDE + n;
                to[0] = t0[c]; to[G] = t1[c]; to[2 * G] = t2[c];
                if (VISC) { to[3 * G] = q0[c]; to[4 * G] = q1[c]; to[5 * G] = q2[c]; to[6 * G] = q3[c]; }
            }
        }
    }
    PR_STAMP(15);
}

// ---- runtime view of the record layout (pack kernels, host) ------------------------------------------------------
struct PairDims {
    int G, Q, NP, NQ2, HDR, O_QB, O_Q0, O_Q2, O_NST, O_ACCN, O_QST, O_ACCQ, O_FST, O_VST, O_QSTR, O_ACCQR, REC, FSIDE, VSIDE, ASIDE, TSIDE;
};
template <int G, int Q>
inline PairDims make_pairdims_t() {
    using R = PairRec<G, Q>;
    PairDims d;
    d.G = G; d.Q = Q; d.NP = R::NP; d.NQ2 = R::NQ2; d.HDR = R::HDR; d.O_QB = R::O_QB; d.O_Q0 = R::O_Q0; d.O_Q2 = R::O_Q2; d.O_NST = R::O_NST;
    d.O_ACCN = R::O_ACCN; d.O_QST = R::O_QST; d.O_ACCQ = R::O_ACCQ; d.O_FST = R::O_FST; d.O_VST = R::O_VST; d.REC = R::REC;
    d.O_QSTR = R::O_QSTR; d.O_ACCQR = R::O_ACCQR;
    d.FSIDE = R::FSIDE; d.VSIDE = R::VSIDE; d.ASIDE = R::ASIDE; d.TSIDE = R::TSIDE;
    return d;
}
// nop 3, 4: warp per element; nop 5..8: block per element (exact integration: nq = 2 nop + 1)
inline bool stage_pair_supported(const Solver& S) { return S.ngl >= 4 && S.ngl <= 9 && S.nq == 2 * S.ngl - 1; }
inline PairDims make_pairdims(int G, int Q) {
    switch (G) {
        case 4: return make_pairdims_t<4, 7>();
        case 5: return make_pairdims_t<5, 9>();
        case 6: return make_pairdims_t<6, 11>();
        case 7: return make_pairdims_t<7, 13>();
        case 8: return make_pairdims_t<8, 15>();
        default: return make_pairdims_t<9, 17>();
    }
}

struct PairPackArgs {
    Mesh M;
    PairDims D;
    const double* qb[3];
    const double* nstp[11];
    const double* qstp[9];
    const double* fstp[11];   // the last two (pbl, pbr) are stored as reciprocals
    const double* bdg[4];
    const double* pbv;
    const double* pbn;
    double *rec, *tr;
    int has_visc;
};
// planes -> records, once per substep loop (block per element): header, state, statics, initial traces.
// The running sums (ACCN, ACCQ, face sums) and the SSPRK work states are zeroed.
// (G_, Q_ > 0: compile-time sizes -- the index arithmetic of the copy loops is division-heavy; 0: run-time sizes)
template <int G_, int Q_>
__global__ void k_pair_pack(PairPackArgs a) {
    extern __shared__ double sm[];
    __shared__ int s_flags;
    const PairDims& D = a.D;
    const int G = G_ ? G_ : D.G, Q = Q_ ? Q_ : D.Q, NP = G * G, NQ2 = Q * Q;
    const int e = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const size_t nbase = (size_t)e * NP, qbase = (size_t)e * NQ2;
    double* r = a.rec + (size_t)e * D.REC;
    double* u = sm;
    double* v = sm + NP;
    if (tid == 0) s_flags = 0;
    __syncthreads();
    // header
    if (tid < 5) r[tid] = a.M.em[e * 5 + tid];
    if (tid == 5 || tid == 23) r[tid] = 0.0;
    if (tid >= 6 && tid < 18) r[tid] = a.M.fgeom[(size_t)e * 12 + (tid - 6)];
    if (tid < 4) {
        const int nb = a.M.nbr[e * 4 + tid], nbs = a.M.nbslot[e * 4 + tid];
        int* hi = reinterpret_cast<int*>(r + 18);
        hi[tid] = nb;
        hi[4 + tid] = (nb >= 0) ? nb * 4 + nbs : (nb == NBR_HALO) ? a.M.nslots + nbs : -1;
    }
    // state, work states, nodal statics, nodal sums
    for (int t = tid; t < 3 * NP; t += nt) { r[D.O_QB + t] = a.qb[t / NP][nbase + t % NP]; r[D.O_Q0 + t] = 0.0; r[D.O_Q2 + t] = 0.0; }
    for (int t = tid; t < 11 * NP; t += nt) {
        const double* p = a.nstp[t / NP];
        r[D.O_NST + t] = p ? p[nbase + t % NP] : 0.0;
    }
    for (int t = tid; t < 6 * NP; t += nt) r[D.O_ACCN + t] = 0.0;
    // quadrature statics (+ forcing sparsity flags), quadrature sums
    int fl = 0;
    for (int t = tid; t < 9 * NQ2; t += nt) {
        const int f = t / NQ2;
        const double x = a.qstp[f][qbase + t % NQ2];
        r[(f < 6 ? D.O_QST : D.O_QSTR - 6 * NQ2) + t] = x;
        if (x != 0.0) fl |= (f == 4) ? PF_COR : (f == 5) ? PF_TWX : (f == 6) ? PF_TWY : (f >= 7) ? PF_GZ : 0;
    }
    if (fl) atomicOr(&s_flags, fl);
    for (int t = tid; t < 6 * NQ2; t += nt) r[D.O_ACCQ + t] = 0.0;
    for (int t = tid; t < 2 * NQ2; t += nt) r[D.O_ACCQR + t] = 0.0;
    // face statics: a copy of the owner's coefficients for every side
    for (int t = tid; t < 4 * 11 * Q; t += nt) {
        int s = t / (11 * Q), rr = t - s * 11 * Q, f = rr / Q, iq = rr - f * Q;
        int slot = e * 4 + s, nb = a.M.nbr[slot], nbs = a.M.nbslot[slot];
        bool left = (nb < 0) || (e < nb);
        double x = 0.0;
        if (left) {   // the right element reads the owner's copy
            x = a.fstp[f][(size_t)slot * Q + iq];
            if (f >= 9) x = 1.0 / x;
        }
        r[D.O_FST + s * D.FSIDE + rr] = x;
    }
    for (int t = tid; t < NP; t += nt) {
        double pb = a.qb[0][nbase + t] + a.nstp[0][nbase + t];
        u[t] = a.qb[1][nbase + t] / pb; v[t] = a.qb[2][nbase + t] / pb;
    }
    __syncthreads();
    if (tid == 0) { int* hi = reinterpret_cast<int*>(r + 22); hi[0] = s_flags; hi[1] = 0; }
    for (int t = tid; t < 4 * G; t += nt) {
        int s = t / G, n = t - s * G, slot = e * 4 + s;
        int I = face_node(s, n, G);
        r[D.O_VST + s * D.VSIDE + n] = a.pbn[(size_t)slot * G + n];
        double so[5] = {0, 0, 0, 0, 0};   // own LDG statics at the face node: btp_dpp_graduv, pbprime_visc
        if (a.has_visc) {
            for (int k = 0; k < 4; ++k) so[k] = a.bdg[k][nbase + I];
            so[4] = a.pbv[nbase + I];
        }
        // traces of the initial state of the loop (the stage kernel publishes the later ones)
        double* tr = a.tr + ((size_t)e * 4 + s) * D.TSIDE + n;
        for (int k = 0; k < 3; ++k) tr[k * G] = a.qb[k][nbase + I];
        double g4[4] = {0, 0, 0, 0};
        if (a.has_visc) {
            const double ksx = a.M.em[e * 5 + 0], ksy = a.M.em[e * 5 + 1], etx = a.M.em[e * 5 + 2], ety = a.M.em[e * 5 + 3];
            int m = I / G, nn = I - m * G;
            double dku = 0, deu = 0, dkv = 0, dev = 0;
            for (int k = 0; k < G; ++k) {
                dku += c_ops.D[k + G * nn] * u[m * G + k]; deu += c_ops.D[k + G * m] * u[k * G + nn];
                dkv += c_ops.D[k + G * nn] * v[m * G + k]; dev += c_ops.D[k + G * m] * v[k * G + nn];
            }
            g4[0] = ksx * dku + etx * deu; g4[1] = ksy * dku + ety * deu; g4[2] = ksx * dkv + etx * dev; g4[3] = ksy * dkv + ety * dev;
        }
        for (int k = 0; k < 4; ++k) tr[(3 + k) * G] = so[4] * g4[k] + so[k];   // LDG flux variable, as the stage kernel publishes it
    }
}
// state records -> planes after the loop
__global__ void k_pair_unpack_qb(int nelem, PairDims D, const double* rec, double* q0, double* q1, double* q2) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)nelem * D.NP) return;
    size_t e = t / D.NP; int I = (int)(t - e * D.NP);
    const double* r = rec + e * D.REC + D.O_QB;
    q0[t] = r[I]; q1[t] = r[D.NP + I]; q2[t] = r[2 * D.NP + I];
}
// traces of the nodal sums S_pbpert, S_mx, S_my (ACCN fields 3..5) for k_btp_finalize
__global__ void k_pair_sum_traces(int nelem, PairDims D, const double* rec, double* tr) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)nelem * 4 * D.G) return;
    size_t e = t / (4 * D.G); int r = (int)(t - e * 4 * D.G), s = r / D.G, n = r - s * D.G;
    int I = face_node(s, n, D.G);
    double* to = tr + (e * 4 + s) * D.TSIDE + n;
    for (int k = 0; k < 3; ++k) to[k * D.G] = rec[e * D.REC + D.O_ACCN + (3 + k) * D.NP + I];
}

template <int G, int Q, int NE, int W, bool VISC, int BOTFR, bool BLK = false>
static int launch_pair_k(Solver& S, const PairArgs& a) {
    using R = PairRec<G, Q>;
    const size_t smem = R::smem_bytes(NE, BLK ? 1 : W);
    auto kern = k_btp_stage_pair<G, Q, NE, W, VISC, BOTFR, BLK>;
    // function attributes are per device: one configuration record per (instantiation, device), set up under a lock
    // (several Solvers on different GPUs, or the threaded in-process ranks of the tests, may arrive here together)
    constexpr int MAXDEV = 64;
    static std::mutex mtx;
    static bool configured[MAXDEV] = {};
    static int upw[MAXDEV] = {};
    const int dv = (S.device >= 0 && S.device < MAXDEV) ? S.device : 0;
    int units_per_wave;
    {
        std::lock_guard<std::mutex> lk(mtx);
        if (!configured[dv]) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
                set_error("cudaFuncSetAttribute", "shared memory opt-in failed"); return -1;
            }
            cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 32 * W, smem);
            upw[dv] = nb * (BLK ? 1 : W) * S.num_sms;
            configured[dv] = true;
            if (getenv("HNUMO_DEBUG"))
                fprintf(stderr, "[hnumo] device %d element-record stage kernel G=%d NE=%d W=%d blk=%d visc=%d botfr=%d: %zu B smem/block, %d blocks/SM (%d warps)\n",
                        dv, G, NE, W, (int)BLK, (int)VISC, BOTFR, smem, nb, nb * W);
        }
        units_per_wave = upw[dv];
    }
    if (a.configure_only) return 0;
    PairArgs b = a;
    // L2 prefetch distance in units (warps, or blocks in block-per-element mode): a quarter / half of the resident wave
    // (measured: 296-1480 units are equivalent at nop 4, 185-370 blocks best at nop 8; profiles/r1_stage_kernel_experiments.md)
    if (b.pf_dist <= 0) b.pf_dist = BLK ? units_per_wave / 2 : units_per_wave / 4;
    if (b.part != 1) b.count = S.nelem;
    if (b.count <= 0) return 0;
    const int units = (b.count + NE - 1) / NE;
    const int blocks = BLK ? units : (units + W - 1) / W;
    kern<<<blocks, 32 * W, smem, b.part == 1 ? S.comm_stream : S.stream>>>(b);
    S.n_launches++;
    return 0;
}
template <int G, int Q, int NE, int W, bool BLK = false>
static int launch_pair_w(Solver& S, const PairArgs& a) {
    const int bf = S.botfr == 1 ? 1 : S.botfr == 2 ? 2 : 0;
    if (S.has_visc) {
        if (bf == 1) return launch_pair_k<G, Q, NE, W, true, 1, BLK>(S, a);
        if (bf == 2) return launch_pair_k<G, Q, NE, W, true, 2, BLK>(S, a);
        return launch_pair_k<G, Q, NE, W, true, 0, BLK>(S, a);
    }
    if (bf == 1) return launch_pair_k<G, Q, NE, W, false, 1, BLK>(S, a);
    if (bf == 2) return launch_pair_k<G, Q, NE, W, false, 2, BLK>(S, a);
    return launch_pair_k<G, Q, NE, W, false, 0, BLK>(S, a);
}
template <int G, int Q>
static int launch_pair_t(Solver& S, const PairArgs& a) {
    // (NE = 2, two elements per warp with double2 shared-memory words, was measured slower at every size -- half the
    //  resident warps -- and is no longer instantiated; the kernel source keeps the NE parameter)
    return launch_pair_w<G, Q, 1, 4>(S, a);
}
inline int launch_stage_pair(Solver& S, const PairArgs& a) {
    if (S.ngl == 5 && S.nq == 9) return launch_pair_t<5, 9>(S, a);
    if (S.ngl == 4 && S.nq == 7) return launch_pair_t<4, 7>(S, a);
    // nop 8: one element per block of 128 threads (81 nodes, 119 pass-2 lines, 68 face points fit; 289 quadrature points in 3 sweeps)
    if (S.ngl == 9 && S.nq == 17) return launch_pair_w<9, 17, 1, HN_BLK_W, true>(S, a);   // (96 threads per element measured slower: 0.41 vs 0.47)
    if (S.ngl == 8 && S.nq == 15) return launch_pair_w<8, 15, 1, 4, true>(S, a);
    if (S.ngl == 7 && S.nq == 13) return launch_pair_w<7, 13, 1, 4, true>(S, a);
    if (S.ngl == 6 && S.nq == 11) return launch_pair_w<6, 11, 1, 4, true>(S, a);
    return -1;
}

}  // namespace hn
