// stage_tma.cuh -- barotropic SSPRK stage kernel with per-element RECORD layout and TMA bulk-copy staging.
//
// Same operator as k_btp_stage_simple / k_btp_stage_fused (reference src/mod_rhs_btp.F90:28-370,
// src/mod_barotropic_terms.F90:25-97,165-217, src/mod_laplacian_quad.F90:32-121,357-519, src/mod_rk_mlswe.F90:87-114).
//
// B200 design
//   * Everything a stage touches lives in per-element records (16-byte aligned, contiguous): state, nodal statics,
//     quadrature statics, face coefficients, accumulators, traces.  The records are built once per substep loop
//     (k_rec_pack) from the plane layout the baroclinic kernels use, and the state is unpacked after the loop.
//   * A block of 4 warps owns one element at a time (persistent loop, elements strided over the grid).  One thread
//     brings the element's ~21 kB of inputs into shared memory with cp.async.bulk (TMA engine, mbarrier completion);
//     accumulators are updated IN shared memory and go back with bulk stores.  No global load/store instruction and
//     no per-plane address arithmetic is left in the compute phases; several blocks per SM overlap one block's
//     transfer with the others' arithmetic.
//   * Forcing (Coriolis, wind stress, bottom slope) is interpolated from its 4 nodal fields in-kernel instead of
//     being read from 5 quadrature-point planes; the LDG gradient sums are derived after the loop (k_btp_finalize).
//   * Contractions are "one line per lane" with compile-time constant-bank matrix operands, as in stage_fused.cuh;
//     lines of one phase are packed 32 per warp, different warps take different roles in the same phase.
#pragma once
#include "stage_fused.cuh"

namespace hn {

__host__ __device__ constexpr int pad2(int n) { return (n + 1) & ~1; }

template <int G, int Q>
struct RecLayout {
    static constexpr int NP = G * G, NQ2 = Q * Q, PER = G * Q;
    // ---- global records (doubles; every size is even => 16-byte aligned records)
    static constexpr int GEOC = 22;                 // ksx ksy etx ety J - | fgeom[4][3] | int nbr[4], nbslot[4]
    static constexpr int QB = pad2(3 * NP);         // pbpert, pb*ub, pb*vb
    static constexpr int NST_NF = 14;               // 0 pbprime 1 massinv 2 qp_dp 3 qp_u 4 qp_v 5 pbv 6..9 bdg 10 f 11 taux 12 tauy 13 zbot
    static constexpr int NST = pad2(NST_NF * NP);
    static constexpr int ACCN = pad2(6 * NP);       // 0 ope2_df 1 ub 2 vb 3 S_pbpert 4 S_mx 5 S_my
    static constexpr int QST_NF = 5;                // 0 oop_q 1 Hbcl 2 Quu 3 Quv 4 Qvv
    static constexpr int QST = pad2(QST_NF * NQ2);
    static constexpr int ACCQ = pad2(8 * NQ2);      // 0 Qu 1 Qv 2 Quv 3 ope2 4 ub 5 vb (6,7 tau_bot when botfr==2)
    static constexpr int FSIDE = pad2(11 * Q);      // cL cR cLR lam oop_edge Quu_e Quv_e Qvv_e Hbcl_e pbl pbr
    static constexpr int FST = 4 * FSIDE;
    static constexpr int ASIDE = pad2(11 * Q);      // acc_f[0..10]
    static constexpr int ACCF = 4 * ASIDE;
    static constexpr int VSIDE = pad2(6 * G);       // neighbour's bdg0..3, pbv, pbprime at the face nodes (ghosts resolved)
    static constexpr int VST = 4 * VSIDE;
    static constexpr int TSIDE = pad2(7 * G);       // pbpert mx my G0..G3 at the face nodes
    static constexpr int TR = 4 * TSIDE;
    // ---- shared memory (doubles)
    static constexpr int cmax(int a, int b) { return a > b ? a : b; }
    static constexpr int S_GEOC = 0;
    static constexpr int S_QB = S_GEOC + GEOC;
    static constexpr int S_NST = S_QB + QB;
    static constexpr int S_ACCN = S_NST + NST;
    static constexpr int S_Q0 = S_ACCN + ACCN;
    static constexpr int S_Q2 = S_Q0 + QB;
    static constexpr int S_QST = S_Q2 + QB;
    static constexpr int S_FST = S_QST + QST;
    static constexpr int S_ACCF = S_FST + FST;
    static constexpr int S_VST = S_ACCF + ACCF;
    static constexpr int S_TR = S_VST + VST;       // incoming neighbour traces, later the outgoing traces
    static constexpr int S_NW = S_TR + TR;         // pb, u, v
    static constexpr int S_T = S_NW + pad2(3 * NP);
    static constexpr int T_SZ = pad2(11 * PER);    // pass-1 results: 0..9 psiq rows, 10 zbot dpsiq rows; later 8 scatter arrays
    static constexpr int S_U = S_T + T_SZ;         // zbot B-columns U[j][n]
    static constexpr int S_X = S_U + pad2(PER);
    static constexpr int X_SZ = 12 * NQ2;          // quadrature-point fields, later the weighted fluxes
    static constexpr int X_FQV = 0;                // (after scatter pass 1) 4*2*4*Q interpolated traces
    static constexpr int X_FF = 32 * Q;            // 4*3*Q face fluxes
    static constexpr int X_PROJ = X_FF + 12 * Q;   // 4*3*G projected face fluxes
    static constexpr int S_L = S_X + X_SZ;         // LDG work 16*NP
    static constexpr int S_OWN = S_L + 16 * NP;    // own traces [s][8][G]: dpp mx my G0..3 pb
    static constexpr int S_NBT = S_OWN + 32 * G;
    static constexpr int S_OWNV = S_NBT + 32 * G;  // own viscosity statics [s][5][G]
    static constexpr int S_LF = S_OWNV + 20 * G;   // LDG face flux [s][2][G]
    static constexpr int S_R = S_LF + 8 * G;       // rhs parts 6*NP
    static constexpr int S_BAR = S_R + pad2(6 * NP);
    static constexpr int S_ACCQ = S_BAR + 2;       // last: 6 or 8 planes
    static int smem_doubles(int naccq) { return S_ACCQ + pad2(naccq * NQ2); }
};

struct TmaArgs {
    int nelem, nslots;
    const double *geoc, *nst, *qst, *fst, *vst;
    double *qb, *q0, *q2, *accn, *accq, *accf;
    const double* tr_in;
    double* tr_out;
    const int4* nbx;   // per element side: bits 0..29 trace record of the neighbour (0x3fffffff: wall), bit 30: this element owns the face
    double a1, a2, a3, dtt, g, cd, alpha_bot, visc;
    int botfr, has_visc, load_q0, load_q2, store_q0, store_q2;
};

// ---- PTX wrappers (cp.async.bulk + mbarrier, sm_90+) ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(void* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded spin: a byte-count mismatch must surface as a launch failure, never as a hung GPU
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 26)) __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// gradient lines of the nodal fields u, v: Lr[0],Lr[1] = d/dksi (u,v), Lr[2],Lr[3] = d/deta (u,v)
template <int G>
__device__ __forceinline__ void ldg_gradient_lines2(const double* u, const double* v, double* Lr, int lane) {
    constexpr int NP = G * G;
    for (int it = lane; it < 4 * G; it += 32) {
        int kind = it / (2 * G), r = it - kind * 2 * G, f = r / G, l = r - f * G;
        const double* src = f ? v : u;
        double in[G], out[G];
        if (kind == 0) {
#pragma unroll
            for (int k = 0; k < G; ++k) in[k] = src[l * G + k];
            line_grad<G>(in, out);
#pragma unroll
            for (int n = 0; n < G; ++n) Lr[f * NP + l * G + n] = out[n];
        } else {
#pragma unroll
            for (int k = 0; k < G; ++k) in[k] = src[k * G + l];
            line_grad<G>(in, out);
#pragma unroll
            for (int m = 0; m < G; ++m) Lr[(2 + f) * NP + m * G + l] = out[m];
        }
    }
}

// lines handled by warps [w0, w0+nw): packed 32 per warp
#define HN_LINES(l, L, w0, nw) for (int l = (warp - (w0)) * 32 + lane; (warp >= (w0)) && (warp < (w0) + (nw)) && l < (L); l += (nw) * 32)

template <int G, int Q>
__global__ void __launch_bounds__(128, 5) k_btp_stage_tma(const TmaArgs a, const int naccq) {
    using RL = RecLayout<G, Q>;
    constexpr int NP = RL::NP, NQ2 = RL::NQ2, PER = RL::PER;
    extern __shared__ __align__(128) double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* geo = sm + RL::S_GEOC;
    const int* conn = reinterpret_cast<const int*>(geo + 18);
    double* qb = sm + RL::S_QB;
    double* nst = sm + RL::S_NST;
    double* accn = sm + RL::S_ACCN;
    double* q0s = sm + RL::S_Q0;
    double* q2s = sm + RL::S_Q2;
    double* qst = sm + RL::S_QST;
    double* accq = sm + RL::S_ACCQ;
    double* fst = sm + RL::S_FST;
    double* accf = sm + RL::S_ACCF;
    double* vst = sm + RL::S_VST;
    double* trs = sm + RL::S_TR;
    double* pbw = sm + RL::S_NW;
    double* uw = pbw + NP;
    double* vw = uw + NP;
    double* T = sm + RL::S_T;
    double* U = sm + RL::S_U;
    double* X = sm + RL::S_X;
    double* Lr = sm + RL::S_L;
    double* own = sm + RL::S_OWN;
    double* nbt = sm + RL::S_NBT;
    double* ownv = sm + RL::S_OWNV;
    double* lf = sm + RL::S_LF;
    double* R = sm + RL::S_R;
    void* bar = sm + RL::S_BAR;
    const bool visc = a.has_visc != 0;
    const int botfr = a.botfr;
    const uint32_t accq_bytes = (uint32_t)(naccq * NQ2 * sizeof(double) + 15) & ~15u;

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    int4 nbx_cur = make_int4(0, 0, 0, 0);
    if (tid == 0 && (int)blockIdx.x < a.nelem) nbx_cur = a.nbx[blockIdx.x];
    uint32_t parity = 0;

    for (int e = blockIdx.x; e < a.nelem; e += gridDim.x, parity ^= 1u) {
        // ================= producer: one thread brings the element's records in =================================
        if (tid == 0) {
            bulk_wait_read0();  // the previous element's bulk stores have finished reading shared memory
            const int nb4[4] = {nbx_cur.x, nbx_cur.y, nbx_cur.z, nbx_cur.w};
            uint32_t bytes = (RL::GEOC + RL::QB + RL::NST + RL::ACCN + RL::QST + RL::FST + RL::VST) * 8u + accq_bytes;
            if (a.load_q0) bytes += RL::QB * 8u;
            if (a.load_q2) bytes += RL::QB * 8u;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                if (nb4[s] & 0x40000000) bytes += RL::ASIDE * 8u;
                if ((nb4[s] & 0x3fffffff) != 0x3fffffff) bytes += RL::TSIDE * 8u;
            }
            mbar_expect_tx(bar, bytes);
            const size_t E = (size_t)e;
            bulk_g2s(geo, a.geoc + E * RL::GEOC, RL::GEOC * 8u, bar);
            bulk_g2s(qb, a.qb + E * RL::QB, RL::QB * 8u, bar);
            bulk_g2s(nst, a.nst + E * RL::NST, RL::NST * 8u, bar);
            bulk_g2s(accn, a.accn + E * RL::ACCN, RL::ACCN * 8u, bar);
            if (a.load_q0) bulk_g2s(q0s, a.q0 + E * RL::QB, RL::QB * 8u, bar);
            if (a.load_q2) bulk_g2s(q2s, a.q2 + E * RL::QB, RL::QB * 8u, bar);
            bulk_g2s(qst, a.qst + E * RL::QST, RL::QST * 8u, bar);
            bulk_g2s(accq, a.accq + E * RL::ACCQ, accq_bytes, bar);
            bulk_g2s(fst, a.fst + E * RL::FST, RL::FST * 8u, bar);
            bulk_g2s(vst, a.vst + E * RL::VST, RL::VST * 8u, bar);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                if (nb4[s] & 0x40000000) bulk_g2s(accf + s * RL::ASIDE, a.accf + (E * 4 + s) * RL::ASIDE, RL::ASIDE * 8u, bar);
                const int rec = nb4[s] & 0x3fffffff;
                if (rec != 0x3fffffff) bulk_g2s(trs + s * RL::TSIDE, a.tr_in + (size_t)rec * RL::TSIDE, RL::TSIDE * 8u, bar);
            }
            const long en = (long)e + gridDim.x;
            if (en < a.nelem) nbx_cur = a.nbx[en];  // connectivity of the next element, one iteration ahead
        }
        if (warp == 0) mbar_wait(bar, parity);   // one polling warp; the others sleep in the hardware barrier
        __syncthreads();

        const double ksx = geo[0], ksy = geo[1], etx = geo[2], ety = geo[3], J = geo[4];
        const double* fg = geo + 6;

        // ---- P1. nodal prep + nodal sums (mod_rk_mlswe.F90:90-92)
        if (warp == 0) {
            for (int I = lane; I < NP; I += 32) {
                double dpp = qb[I], mx = qb[NP + I], my = qb[2 * NP + I], pbp = nst[I];
                double pb = dpp + pbp;
                double rpb = 1.0 / pb;
                double u = mx * rpb, v = my * rpb;
                pbw[I] = pb; uw[I] = u; vw[I] = v;
                double oop = pbp > 0.0 ? 1.0 / pbp : 0.0;
                double t = 1.0 + dpp * oop;
                accn[I] += t * t; accn[NP + I] += u; accn[2 * NP + I] += v;
                accn[3 * NP + I] += dpp; accn[4 * NP + I] += mx; accn[5 * NP + I] += my;
            }
        }
        __syncthreads();
        // ---- P2. sum factorisation pass 1.  T sets: 0 dpp 1 mx 2 my 3 pb 4 pp 5 up 6 vp 7 f 8 taux 9 tauy (psiq rows),
        //          10 zbot (dpsiq rows); U: zbot dpsiq columns.
        {
            const int nset = botfr ? 10 : 7;
            HN_LINES(l, nset * G, 0, 2) {
                int k = l / G, m = l - k * G;
                int f = (botfr || k < 4) ? k : k + 3;
                const double* src = (f < 3) ? qb + f * NP : (f == 3) ? pbw : (f < 7) ? nst + (f - 2) * NP : nst + (f + 3) * NP;
                double in[G], out[Q];
#pragma unroll
                for (int n = 0; n < G; ++n) in[n] = src[m * G + n];
                line_n2q<G, Q, false>(in, out);
#pragma unroll
                for (int i = 0; i < Q; ++i) T[f * PER + m * Q + i] = out[i];
            }
            HN_LINES(l, 2 * G, 2, 1) {
                const double* zb = nst + 13 * NP;
                double in[G], out[Q];
                if (l < G) {
#pragma unroll
                    for (int n = 0; n < G; ++n) in[n] = zb[l * G + n];
                    line_n2q<G, Q, true>(in, out);
#pragma unroll
                    for (int i = 0; i < Q; ++i) T[10 * PER + l * Q + i] = out[i];
                } else {
                    const int n = l - G;
#pragma unroll
                    for (int m = 0; m < G; ++m) in[m] = zb[m * G + n];
                    line_n2q<G, Q, true>(in, out);
#pragma unroll
                    for (int j = 0; j < Q; ++j) U[j * G + n] = out[j];
                }
            }
            if (visc && warp == 3) ldg_gradient_lines2<G>(uw, vw, Lr, lane);
        }
        __syncthreads();
        // ---- P3. pass 2 -> X sets: 0 dpp 1 udp 2 vdp 3 dp 4 pp 5 up 6 vp 7 f 8 taux 9 tauy 10 dzb/dksi 11 dzb/deta
        {
            const int nset = botfr ? 12 : 9;
            HN_LINES(l, nset * Q, 0, 4) {
                int k = l / Q, c = l - k * Q;
                int f = (botfr || k < 4) ? k : k + 3;
                double in[G], out[Q];
                if (f < 11) {
                    const double* src = T + f * PER;                    // X[10] from the dpsiq rows T[10]
#pragma unroll
                    for (int m = 0; m < G; ++m) in[m] = src[m * Q + c];
                    line_n2q<G, Q, false>(in, out);
#pragma unroll
                    for (int j = 0; j < Q; ++j) X[f * NQ2 + j * Q + c] = out[j];
                } else {                                                // X[11][j][i] = sum_n psiq(n,i) U[j][n]
#pragma unroll
                    for (int n = 0; n < G; ++n) in[n] = U[c * G + n];
                    line_n2q<G, Q, false>(in, out);
#pragma unroll
                    for (int i = 0; i < Q; ++i) X[11 * NQ2 + c * Q + i] = out[i];
                }
            }
        }
        __syncthreads();
        // ---- P4. pointwise physics at the quadrature points (mod_rhs_btp.F90:136-192), fluxes overwrite X in place;
        //          warp 3: LDG auxiliary variable at the nodes (mod_laplacian_quad.F90:50-56,357-390)
        if (warp < 3) {
            for (int q = tid; q < NQ2; q += 96) {
                int j = q / Q, i = q - j * Q;
                double dpp = X[0 * NQ2 + q], udp = X[1 * NQ2 + q], vdp = X[2 * NQ2 + q], dp = X[3 * NQ2 + q];
                double wq = c_ops.wq[i] * c_ops.wq[j] * J;
                double rdp = 1.0 / dp;
                double ub = udp * rdp, vb = vdp * rdp;
                double tb_u = 0.0, tb_v = 0.0;
                if (botfr) {
                    double pp = X[4 * NQ2 + q], up = X[5 * NQ2 + q], vp = X[6 * NQ2 + q];
                    double ubot = up + ub, vbot = vp + vb;
                    double spd = (botfr == 1) ? (a.cd / a.g) * pp : (a.cd / a.alpha_bot) * sqrt(ubot * ubot + vbot * vbot);
                    tb_u = spd * ubot; tb_v = spd * vbot;
                }
                double fcor = X[7 * NQ2 + q], twx = X[8 * NQ2 + q], twy = X[9 * NQ2 + q];
                double zk = X[10 * NQ2 + q], ze = X[11 * NQ2 + q];
                double gzx = ksx * zk + etx * ze, gzy = ksy * zk + ety * ze;
                double sc_x = fcor * vdp + a.g * (twx - tb_u) - a.g * dp * gzx;
                double sc_y = -fcor * udp + a.g * (twy - tb_v) - a.g * dp * gzy;
                double ope = 1.0 + dpp * qst[q];
                double ope2 = ope * ope;
                double Hq = ope2 * qst[NQ2 + q];
                double qu = ub * udp + ope * qst[2 * NQ2 + q];
                double quv = ub * vdp + ope * qst[3 * NQ2 + q];
                double qv = vb * vdp + ope * qst[4 * NQ2 + q];
                accq[q] += qu; accq[NQ2 + q] += qv; accq[2 * NQ2 + q] += quv; accq[3 * NQ2 + q] += ope2;
                accq[4 * NQ2 + q] += ub; accq[5 * NQ2 + q] += vb;
                if (botfr == 2) { accq[6 * NQ2 + q] += tb_u; accq[7 * NQ2 + q] += tb_v; }
                double Fx2 = Hq + qu, Fy3 = Hq + qv;
                X[0 * NQ2 + q] = wq * (ksx * udp + ksy * vdp);   // Fk1
                X[1 * NQ2 + q] = wq * (etx * udp + ety * vdp);   // Fe1
                X[2 * NQ2 + q] = wq * sc_x;                      // S2
                X[3 * NQ2 + q] = wq * (ksx * Fx2 + ksy * quv);   // Fk2
                X[4 * NQ2 + q] = wq * (etx * Fx2 + ety * quv);   // Fe2
                X[5 * NQ2 + q] = wq * sc_y;                      // S3
                X[6 * NQ2 + q] = wq * (ksx * quv + ksy * Fy3);   // Fk3
                X[7 * NQ2 + q] = wq * (etx * quv + ety * Fy3);   // Fe3
            }
        } else if (visc) {
            for (int I = lane; I < NP; I += 32) {
                int m = I / G, n = I - m * G;
                double dku = Lr[0 * NP + I], dkv = Lr[1 * NP + I], deu = Lr[2 * NP + I], dev = Lr[3 * NP + I];
                double g0 = ksx * dku + etx * deu, g1 = ksy * dku + ety * deu, g2 = ksx * dkv + etx * dev, g3 = ksy * dkv + ety * dev;
                Lr[4 * NP + I] = g0; Lr[5 * NP + I] = g1; Lr[6 * NP + I] = g2; Lr[7 * NP + I] = g3;
                double pv = nst[5 * NP + I];
                double q0 = pv * g0 + nst[6 * NP + I], q1 = pv * g1 + nst[7 * NP + I];
                double q2 = pv * g2 + nst[8 * NP + I], q3 = pv * g3 + nst[9 * NP + I];
                double w = c_ops.wg[n] * c_ops.wg[m] * J;
                Lr[8 * NP + I] = w * (ksx * q0 + ksy * q1); Lr[9 * NP + I] = w * (ksx * q2 + ksy * q3);
                Lr[10 * NP + I] = w * (etx * q0 + ety * q1); Lr[11 * NP + I] = w * (etx * q2 + ety * q3);
            }
        }
        __syncthreads();
        // ---- P5. scatter pass 1 (contraction over j): psiq arrays Fk1 S2 Fk2 S3 Fk3 on warps 0-1, dpsiq arrays Fe1 Fe2 Fe3 on warp 2;
        //          warp 3: face traces, own and neighbour (btp_extract_df), LDG gradient traces
        {
            HN_LINES(l, 5 * Q, 0, 2) {
                int k = l / Q, i = l - k * Q;
                int arr = (k == 0) ? 0 : (k == 1) ? 2 : (k == 2) ? 3 : (k == 3) ? 5 : 6;
                double in[Q], out[G];
#pragma unroll
                for (int j = 0; j < Q; ++j) in[j] = X[arr * NQ2 + j * Q + i];
                line_q2n<G, Q, false>(in, out);
#pragma unroll
                for (int m = 0; m < G; ++m) T[arr * PER + m * Q + i] = out[m];
            }
            HN_LINES(l, 3 * Q, 2, 1) {
                int k = l / Q, i = l - k * Q;
                int arr = (k == 0) ? 1 : (k == 1) ? 4 : 7;
                double in[Q], out[G];
#pragma unroll
                for (int j = 0; j < Q; ++j) in[j] = X[arr * NQ2 + j * Q + i];
                line_q2n<G, Q, true>(in, out);
#pragma unroll
                for (int m = 0; m < G; ++m) T[arr * PER + m * Q + i] = out[m];
            }
            HN_LINES(it, 4 * G, 3, 1) {
                int s = it / G, n = it - s * G;
                const int nb = conn[s];
                const double nx = fg[s * 3 + 0], ny = fg[s * 3 + 1];
                const int I = face_node(s, n, G);
                const double ow0 = qb[I], ow1 = qb[NP + I], ow2 = qb[2 * NP + I];
                double n0, n1, n2;
                const double* tn = trs + s * RL::TSIDE + n;
                if (nb >= 0 || nb == NBR_HALO) { n0 = tn[0]; n1 = tn[G]; n2 = tn[2 * G]; }
                else {
                    n0 = ow0; n1 = ow1; n2 = ow2;
                    if (nb == NBR_FREESLIP) { double un = nx * ow1 + ny * ow2; n1 = ow1 - 2.0 * un * nx; n2 = ow2 - 2.0 * un * ny; }
                    else if (nb == NBR_NOSLIP) { n1 = -ow1; n2 = -ow2; }
                }
                double* po = own + (s * 8) * G + n;
                double* pn = nbt + (s * 8) * G + n;
                po[0] = ow0; po[G] = ow1; po[2 * G] = ow2; po[7 * G] = pbw[I];
                pn[0] = n0; pn[G] = n1; pn[2 * G] = n2; pn[7 * G] = n0 + vst[s * RL::VSIDE + 5 * G + n];
                if (visc) {
                    double go[4] = {Lr[4 * NP + I], Lr[5 * NP + I], Lr[6 * NP + I], Lr[7 * NP + I]}, gn[4];
                    if (nb >= 0 || nb == NBR_HALO) {
#pragma unroll
                        for (int v = 0; v < 4; ++v) gn[v] = tn[(3 + v) * G];
                    } else {
#pragma unroll
                        for (int v = 0; v < 4; ++v) gn[v] = go[v];
                        if (nb == NBR_FREESLIP) reflect4(go, nx, ny, gn);
                    }
#pragma unroll
                    for (int v = 0; v < 4; ++v) { po[(3 + v) * G] = go[v]; pn[(3 + v) * G] = gn[v]; }
#pragma unroll
                    for (int v = 0; v < 4; ++v) ownv[(s * 5 + v) * G + n] = nst[(6 + v) * NP + I];
                    ownv[(s * 5 + 4) * G + n] = nst[5 * NP + I];
                }
            }
        }
        __syncthreads();
        // ---- P6. warp 0/1: scatter pass 2 (contraction over i) -> R; warp 2: traces -> face quadrature points;
        //          warp 3: LDG volume lines and LDG face flux
        {
            HN_LINES(it, 3 * G, 0, 1) {   // dpsiq part: sum_i B(n,i) Fk_f[m][i]
                int f = it / G, m = it - f * G;
                double in[Q], out[G];
#pragma unroll
                for (int i = 0; i < Q; ++i) in[i] = T[(3 * f) * PER + m * Q + i];
                line_q2n<G, Q, true>(in, out);
#pragma unroll
                for (int n = 0; n < G; ++n) R[f * NP + m * G + n] = out[n];
            }
            HN_LINES(it, 3 * G, 1, 1) {   // psiq part: sum_i A(n,i) (Fe_f + S_f)[m][i]
                int f = it / G, m = it - f * G;
                double in[Q], out[G];
#pragma unroll
                for (int i = 0; i < Q; ++i) {
                    double v = T[(3 * f + 1) * PER + m * Q + i];
                    if (f > 0) v += T[(3 * f - 1) * PER + m * Q + i];
                    in[i] = v;
                }
                line_q2n<G, Q, false>(in, out);
#pragma unroll
                for (int n = 0; n < G; ++n) R[(3 + f) * NP + m * G + n] = out[n];
            }
            HN_LINES(it, 32, 2, 1) {      // (side, L/R, variable) lines; var: 0 pb 1 dpp 2 mx 3 my
                int s = it >> 3, side = (it >> 2) & 1, var = it & 3;
                const bool left = (conn[s] < 0) || (e < conn[s]);
                const double* src = ((side == 0) == left ? own : nbt) + (s * 8 + (var == 0 ? 7 : var - 1)) * G;
                double in[G], out[Q];
#pragma unroll
                for (int n = 0; n < G; ++n) in[n] = src[n];
                line_n2q<G, Q, false>(in, out);
#pragma unroll
                for (int iq = 0; iq < Q; ++iq) X[RL::X_FQV + it * Q + iq] = out[iq];
            }
            if (visc && warp == 3) {
                for (int it = lane; it < 4 * G; it += 32) {
                    int kind = it / (2 * G), r = it - kind * 2 * G, c = r / G, l = r - c * G;
                    double in[G], out[G];
                    if (kind == 0) {
#pragma unroll
                        for (int n = 0; n < G; ++n) in[n] = Lr[(8 + c) * NP + l * G + n];
                        line_gradT<G>(in, out);
#pragma unroll
                        for (int n = 0; n < G; ++n) Lr[(12 + c) * NP + l * G + n] = out[n];
                    } else {
#pragma unroll
                        for (int m = 0; m < G; ++m) in[m] = Lr[(10 + c) * NP + m * G + l];
                        line_gradT<G>(in, out);
#pragma unroll
                        for (int m = 0; m < G; ++m) Lr[(14 + c) * NP + m * G + l] = out[m];
                    }
                }
                // LDG face flux at the face nodes, as written (mod_laplacian_quad.F90:427-519)
                for (int it = lane; it < 4 * G; it += 32) {
                    int s = it / G, n = it - s * G;
                    const bool left = (conn[s] < 0) || (e < conn[s]);
                    const double nx = fg[s * 3 + 0], ny = fg[s * 3 + 1], nlen = fg[s * 3 + 2];
                    const double* go = own + (s * 8 + 3) * G + n;
                    const double* gn = nbt + (s * 8 + 3) * G + n;
                    const double* so = ownv + s * 5 * G + n;
                    const double* sn = vst + s * RL::VSIDE + n;
                    double fo[4], fn[4];
#pragma unroll
                    for (int v = 0; v < 4; ++v) { fo[v] = so[4 * G] * go[v * G] + so[v * G]; fn[v] = sn[4 * G] * gn[v * G] + sn[v * G]; }
                    const double* fl = left ? fo : fn;
                    const double* fr = left ? fn : fo;
                    double qu0 = 0.5 * fl[0] + 0.5 * fr[0], qu1 = 0.5 * fl[1] + 0.5 * fr[1];
                    double qv0 = 0.5 * fl[2] + 0.5 * fr[2], qv1 = 0.5 * fl[3] + 0.5 * fr[3];
                    double wq = c_ops.wg[n] * nlen;
                    double flux_qu = (qu0 - fl[0] * nx) + (qu1 - fl[1] * ny);
                    double flux_qv = (qv0 - fl[2] * nx) + (qv1 - fl[3] * ny);
                    double sgn = left ? wq : -wq;
                    lf[(s * 2 + 0) * G + n] = sgn * flux_qu;
                    lf[(s * 2 + 1) * G + n] = sgn * flux_qv;
                }
            }
        }
        __syncthreads();
        // ---- P7. face fluxes, canonical left perspective (mod_rhs_btp.F90:237-330): 4Q points over warps 0-1
        {
            const double* fqv = X + RL::X_FQV;
            double* ff = X + RL::X_FF;
            constexpr int HALF = (4 * Q + 1) / 2;
            if (warp < 2) {
                for (int k = lane; k < HALF; k += 32) {
                    const int it = warp * HALF + k;
                    if (it >= 4 * Q) break;
                    int s = it / Q, iq = it - s * Q;
                    const bool left = (conn[s] < 0) || (e < conn[s]);
                    const double nxl = fg[s * 3 + 0], nyl = fg[s * 3 + 1], nlen = fg[s * 3 + 2];
                    const double* Lp = fqv + (s * 8) * Q + iq;
                    const double* Rp = fqv + (s * 8 + 4) * Q + iq;
                    double pbL = Lp[0], ppL = Lp[Q], mxL = Lp[2 * Q], myL = Lp[3 * Q];
                    double pbR = Rp[0], ppR = Rp[Q], mxR = Rp[2 * Q], myR = Rp[3 * Q];
                    const double* fc = fst + s * RL::FSIDE + iq;
                    double cL = fc[0], cR = fc[Q], cLR = fc[2 * Q], lam = fc[3 * Q];
                    double pU_L = nxl * mxL + nyl * myL;
                    double pU_R = -nxl * mxR - nyl * myR;
                    double pbpert_edge = cL * ppL + cR * ppR + cLR * (pU_L + pU_R);
                    double ope_e = 1.0 + pbpert_edge * fc[4 * Q];
                    double fex = cR * mxL + cL * mxR + lam * (nxl * ppL - nxl * ppR);
                    double fey = cR * myL + cL * myR + lam * (nyl * ppL - nyl * ppR);
                    double rl = 1.0 / pbL, rr = 1.0 / pbR;
                    double ul = mxL * rl, ur = mxR * rr, vl = myL * rl, vr = myR * rr;
                    double quu = 0.5 * (ul * mxL + ur * mxR) + ope_e * fc[5 * Q];
                    double quv = 0.5 * (vl * mxL + vr * mxR) + ope_e * fc[6 * Q];
                    double qvu = 0.5 * (ul * myL + ur * myR) + ope_e * fc[6 * Q];
                    double qvv = 0.5 * (vl * myL + vr * myR) + ope_e * fc[7 * Q];
                    double e2 = ope_e * ope_e;
                    double Hf = e2 * fc[8 * Q];
                    if (left) {
                        double ol = 1.0 + (ppL / fc[9 * Q]), orr = 1.0 + (ppR / fc[10 * Q]);
                        double* af = accf + s * RL::ASIDE + iq;
                        af[0] += quu; af[Q] += quv; af[2 * Q] += qvu; af[3 * Q] += qvv;
                        af[4 * Q] += ol * ol; af[5 * Q] += orr * orr; af[6 * Q] += e2;
                        af[7 * Q] += ul; af[8 * Q] += ur; af[9 * Q] += vl; af[10 * Q] += vr;
                    }
                    double wq = c_ops.wq[iq] * nlen;
                    double dispu = 0.5 * lam * (mxR - mxL), dispv = 0.5 * lam * (myR - myL);
                    double flux_x = nxl * quu + nyl * quv - dispu;
                    double flux_y = nxl * qvu + nyl * qvv - dispv;
                    double flux = nxl * fex + nyl * fey;
                    double sgn = left ? -wq : wq;
                    ff[(s * 3 + 0) * Q + iq] = sgn * flux;
                    ff[(s * 3 + 1) * Q + iq] = sgn * (nxl * Hf + flux_x);
                    ff[(s * 3 + 2) * Q + iq] = sgn * (nyl * Hf + flux_y);
                }
            }
        }
        __syncthreads();
        // ---- P8-P11 on warp 3: projection of the face fluxes, update, traces of the new state
        if (warp == 3) {
            const double* ff = X + RL::X_FF;
            double* proj = X + RL::X_PROJ;
            for (int it = lane; it < 12; it += 32) {
                double in[Q], out[G];
#pragma unroll
                for (int iq = 0; iq < Q; ++iq) in[iq] = ff[it * Q + iq];
                line_q2n<G, Q, false>(in, out);
#pragma unroll
                for (int n = 0; n < G; ++n) proj[it * G + n] = out[n];
            }
            __syncwarp();
            // gather per node, mass matrix, viscosity, SSPRK update, wall projection (mod_rk_mlswe.F90:97-108)
            for (int I = lane; I < NP; I += 32) {
                int m = I / G, n = I - m * G;
                double r0 = R[0 * NP + I] + R[3 * NP + I], r1 = R[1 * NP + I] + R[4 * NP + I], r2 = R[2 * NP + I] + R[5 * NP + I];
                double l0 = 0.0, l1 = 0.0;
                if (visc) { l0 = -(Lr[12 * NP + I] + Lr[14 * NP + I]); l1 = -(Lr[13 * NP + I] + Lr[15 * NP + I]); }
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    bool on = (s == 0) ? (m == 0) : (s == 1) ? (m == G - 1) : (s == 2) ? (n == 0) : (n == G - 1);
                    if (!on) continue;
                    int nf = (s < 2) ? n : m;
                    r0 += proj[(s * 3 + 0) * G + nf]; r1 += proj[(s * 3 + 1) * G + nf]; r2 += proj[(s * 3 + 2) * G + nf];
                    if (visc) { l0 += lf[(s * 2 + 0) * G + nf]; l1 += lf[(s * 2 + 1) * G + nf]; }
                }
                double mi = nst[NP + I];
                r0 = mi * r0; r1 = mi * r1; r2 = mi * r2;
                if (visc) { r1 = r1 + a.visc * mi * l0; r2 = r2 + a.visc * mi * l1; }
                double q1[3] = {qb[I], qb[NP + I], qb[2 * NP + I]};
                double q0[3], q2[3] = {0.0, 0.0, 0.0};
#pragma unroll
                for (int v = 0; v < 3; ++v) q0[v] = a.load_q0 ? q0s[v * NP + I] : q1[v];
                if (a.load_q2) {
#pragma unroll
                    for (int v = 0; v < 3; ++v) q2[v] = q2s[v * NP + I];
                }
                double rr[3] = {r0, r1, r2}, qn[3];
#pragma unroll
                for (int v = 0; v < 3; ++v) qn[v] = a.a1 * q0[v] + a.a2 * q1[v] + a.a3 * q2[v] + a.dtt * rr[v];
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    bool on = (s == 0) ? (m == 0) : (s == 1) ? (m == G - 1) : (s == 2) ? (n == 0) : (n == G - 1);
                    if (!on) continue;
                    int nb = conn[s];
                    if (nb == NBR_FREESLIP) {
                        double nx = fg[s * 3 + 0], ny = fg[s * 3 + 1];
                        double unl = qn[1] * nx + qn[2] * ny;
                        qn[1] = qn[1] - unl * nx; qn[2] = qn[2] - unl * ny;
                    } else if (nb == NBR_NOSLIP) { qn[1] = 0.0; qn[2] = 0.0; }
                }
                if (a.store_q0) {
#pragma unroll
                    for (int v = 0; v < 3; ++v) q0s[v * NP + I] = q1[v];
                }
#pragma unroll
                for (int v = 0; v < 3; ++v) qb[v * NP + I] = qn[v];
                if (a.store_q2) {
#pragma unroll
                    for (int v = 0; v < 3; ++v) q2s[v * NP + I] = qn[v];
                }
                double pbn_ = qn[0] + nst[I];
                uw[I] = qn[1] / pbn_; vw[I] = qn[2] / pbn_;
            }
            __syncwarp();
            if (visc) { ldg_gradient_lines2<G>(uw, vw, Lr, lane); __syncwarp(); }
            for (int it = lane; it < 4 * G; it += 32) {
                int s = it / G, n = it - s * G;
                int I = face_node(s, n, G);
                double* to = trs + s * RL::TSIDE + n;
                to[0] = qb[I]; to[G] = qb[NP + I]; to[2 * G] = qb[2 * NP + I];
                if (visc) {
                    double dku = Lr[0 * NP + I], dkv = Lr[1 * NP + I], deu = Lr[2 * NP + I], dev = Lr[3 * NP + I];
                    to[3 * G] = ksx * dku + etx * deu;
                    to[4 * G] = ksy * dku + ety * deu;
                    to[5 * G] = ksx * dkv + etx * dev;
                    to[6 * G] = ksy * dkv + ety * dev;
                } else { to[3 * G] = 0.0; to[4 * G] = 0.0; to[5 * G] = 0.0; to[6 * G] = 0.0; }
            }
        }
        // ================= results go back: generic-proxy writes -> async proxy, then one thread stores ===========
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            const size_t E = (size_t)e;
            const int own4[4] = {conn[0], conn[1], conn[2], conn[3]};
            bulk_s2g(a.qb + E * RL::QB, qb, RL::QB * 8u);
            bulk_s2g(a.accn + E * RL::ACCN, accn, RL::ACCN * 8u);
            bulk_s2g(a.accq + E * RL::ACCQ, accq, accq_bytes);
#pragma unroll
            for (int s = 0; s < 4; ++s)
                if (own4[s] < 0 || e < own4[s]) bulk_s2g(a.accf + (E * 4 + s) * RL::ASIDE, accf + s * RL::ASIDE, RL::ASIDE * 8u);
            bulk_s2g(a.tr_out + E * RL::TR, trs, RL::TR * 8u);
            if (a.store_q0) bulk_s2g(a.q0 + E * RL::QB, q0s, RL::QB * 8u);
            if (a.store_q2) bulk_s2g(a.q2 + E * RL::QB, q2s, RL::QB * 8u);
            bulk_commit();
        }
    }
    if (tid == 0) bulk_wait0();
}

// ---- runtime view of the record sizes (pack kernels, host) ---------------------------------------------------
struct RecDims {
    int G, Q, NP, NQ2, GEOC, QB, NST, ACCN, QST, ACCQ, FSIDE, ASIDE, VSIDE, TSIDE;
};
inline RecDims make_recdims(int G, int Q) {
    RecDims d;
    d.G = G; d.Q = Q; d.NP = G * G; d.NQ2 = Q * Q; d.GEOC = 22; d.QB = pad2(3 * d.NP); d.NST = pad2(14 * d.NP); d.ACCN = pad2(6 * d.NP);
    d.QST = pad2(5 * d.NQ2); d.ACCQ = pad2(8 * d.NQ2); d.FSIDE = pad2(11 * Q); d.ASIDE = pad2(11 * Q); d.VSIDE = pad2(6 * G); d.TSIDE = pad2(7 * G);
    return d;
}

// static per-element records: geometry + connectivity, and the producer's neighbour table
__global__ void k_rec_static(Mesh M, RecDims D, double* geoc, int4* nbx) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M.nelem) return;
    double* g = geoc + (size_t)e * D.GEOC;
    for (int k = 0; k < 5; ++k) g[k] = M.em[e * 5 + k];
    g[5] = 0.0;
    for (int k = 0; k < 12; ++k) g[6 + k] = M.fgeom[(size_t)e * 12 + k];
    int* c = reinterpret_cast<int*>(g + 18);
    int x[4];
    for (int s = 0; s < 4; ++s) {
        int nb = M.nbr[e * 4 + s], nbs = M.nbslot[e * 4 + s];
        c[s] = nb; c[4 + s] = nbs;
        int rec = (nb >= 0) ? nb * 4 + nbs : (nb == NBR_HALO) ? M.nslots + nbs : 0x3fffffff;
        bool left = (nb < 0) || (e < nb);
        x[s] = rec | (left ? 0x40000000 : 0);
    }
    nbx[e] = make_int4(x[0], x[1], x[2], x[3]);
}

struct RecPackArgs {
    Mesh M;
    RecDims D;
    const double* qb[3];
    const double* nstp[14];
    const double* qstp[5];
    const double* fstp[11];
    const double* bdg[4];
    const double* pbv;
    const double* pbn;
    const double* hstat;  // halo copies of (bdg0..3, pbv) traces
    size_t hstat_stride;
    double *r_qb, *r_nst, *r_qst, *r_fst, *r_vst, *r_tr;
    int has_visc;
};
// planes -> records, once per substep loop (block per element)
__global__ void k_rec_pack(RecPackArgs a) {
    extern __shared__ double sm[];
    const RecDims& D = a.D;
    const int G = D.G, Q = D.Q, NP = D.NP, NQ2 = D.NQ2;
    const int e = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const size_t nbase = (size_t)e * NP, qbase = (size_t)e * NQ2;
    double* u = sm;
    double* v = sm + NP;
    for (int t = tid; t < 3 * NP; t += nt) a.r_qb[(size_t)e * D.QB + t] = a.qb[t / NP][nbase + t % NP];
    for (int t = tid; t < 14 * NP; t += nt) {
        const double* p = a.nstp[t / NP];
        a.r_nst[(size_t)e * D.NST + t] = p ? p[nbase + t % NP] : 0.0;
    }
    for (int t = tid; t < 5 * NQ2; t += nt) a.r_qst[(size_t)e * D.QST + t] = a.qstp[t / NQ2][qbase + t % NQ2];
    for (int t = tid; t < 4 * 11 * Q; t += nt) {
        int s = t / (11 * Q), r = t - s * 11 * Q, f = r / Q, iq = r - f * Q;
        int slot = e * 4 + s, nb = a.M.nbr[slot], nbs = a.M.nbslot[slot];
        bool left = (nb < 0) || (e < nb);
        int oslot = left ? slot : nb * 4 + nbs;
        a.r_fst[((size_t)e * 4 + s) * D.FSIDE + r] = a.fstp[f][(size_t)oslot * Q + iq];
    }
    for (int t = tid; t < NP; t += nt) {
        double pb = a.qb[0][nbase + t] + a.nstp[0][nbase + t];
        u[t] = a.qb[1][nbase + t] / pb; v[t] = a.qb[2][nbase + t] / pb;
    }
    __syncthreads();
    for (int t = tid; t < 4 * G; t += nt) {
        int s = t / G, n = t - s * G, slot = e * 4 + s, nb = a.M.nbr[slot], nbs = a.M.nbslot[slot];
        int I = face_node(s, n, G);
        double nx = a.M.fgeom[slot * 3 + 0], ny = a.M.fgeom[slot * 3 + 1];
        double* vs = a.r_vst + ((size_t)e * 4 + s) * D.VSIDE + n;
        double so[5] = {0, 0, 0, 0, 0}, sn[5] = {0, 0, 0, 0, 0};
        if (a.has_visc) {
            for (int k = 0; k < 4; ++k) so[k] = a.bdg[k][nbase + I];
            so[4] = a.pbv[nbase + I];
            if (nb >= 0) {
                size_t In = (size_t)nb * NP + face_node(nbs, n, G);
                for (int k = 0; k < 4; ++k) sn[k] = a.bdg[k][In];
                sn[4] = a.pbv[In];
            } else if (nb == NBR_HALO) {
                for (int k = 0; k < 5; ++k) sn[k] = a.hstat[k * a.hstat_stride + (size_t)nbs * G + n];
            } else {
                for (int k = 0; k < 5; ++k) sn[k] = so[k];
                if (nb == NBR_FREESLIP) reflect4(so, nx, ny, sn);
            }
        }
        for (int k = 0; k < 5; ++k) vs[k * G] = sn[k];
        vs[5 * G] = a.pbn[(size_t)slot * G + n];
        // traces of the initial state of the loop (the stage kernel publishes the later ones)
        double* tr = a.r_tr + ((size_t)e * 4 + s) * D.TSIDE + n;
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
        for (int k = 0; k < 4; ++k) tr[(3 + k) * G] = g4[k];
    }
}
// state records -> planes after the loop
__global__ void k_rec_unpack_qb(int nelem, RecDims D, const double* r_qb, double* q0, double* q1, double* q2) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)nelem * D.NP) return;
    size_t e = t / D.NP; int I = (int)(t - e * D.NP);
    const double* r = r_qb + e * D.QB;
    q0[t] = r[I]; q1[t] = r[D.NP + I]; q2[t] = r[2 * D.NP + I];
}
// traces of the nodal sums S_pbpert, S_mx, S_my (accn fields 3..5) for k_btp_finalize
__global__ void k_rec_sum_traces(int nelem, RecDims D, const double* r_accn, double* r_tr) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)nelem * 4 * D.G) return;
    size_t e = t / (4 * D.G); int r = (int)(t - e * 4 * D.G), s = r / D.G, n = r - s * D.G;
    int I = face_node(s, n, D.G);
    double* tr = r_tr + (e * 4 + s) * D.TSIDE + n;
    for (int k = 0; k < 3; ++k) tr[k * D.G] = r_accn[e * D.ACCN + (3 + k) * D.NP + I];
}
// halo: gather the trace records of the processor faces into the send buffer (the receive side lands in the
// halo region of the trace buffer directly)
__global__ void k_pack_trace_records(const double* tr, const int* halo_slot, int nhalo, int tside, double* send) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)nhalo * tside) return;
    size_t h = t / tside; int j = (int)(t - h * tside);
    send[t] = tr[(size_t)halo_slot[h] * tside + j];
}

inline bool stage_tma_supported(const Solver& S) { return (S.ngl == 5 && S.nq == 9) || (S.ngl == 4 && S.nq == 7); }

template <int G, int Q>
static int launch_tma_t(Solver& S, const TmaArgs& a, int naccq) {
    using RL = RecLayout<G, Q>;
    size_t smem = (size_t)RL::smem_doubles(naccq) * sizeof(double);
    static int blocks_per_sm = 0;
    static size_t configured_smem = 0;
    if (configured_smem != smem) {
        if (cudaFuncSetAttribute(k_btp_stage_tma<G, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            set_error("cudaFuncSetAttribute", "shared memory opt-in failed"); return -1;
        }
        cudaFuncSetAttribute(k_btp_stage_tma<G, Q>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_btp_stage_tma<G, Q>, 128, smem) != cudaSuccess || blocks_per_sm < 1) {
            set_error("cudaOccupancy", "stage kernel does not fit on an SM"); return -1;
        }
        configured_smem = smem;
    }
    int bps = S.tma_blocks_per_sm > 0 ? std::min(S.tma_blocks_per_sm, blocks_per_sm) : blocks_per_sm;
    int grid = std::min(S.nelem, S.num_sms * bps);
    k_btp_stage_tma<G, Q><<<grid, 128, smem, S.stream>>>(a, naccq);
    S.n_launches++;
    return 0;
}
inline int launch_stage_tma(Solver& S, const TmaArgs& a, int naccq) {
    if (S.ngl == 5 && S.nq == 9) return launch_tma_t<5, 9>(S, a, naccq);
    if (S.ngl == 4 && S.nq == 7) return launch_tma_t<4, 7>(S, a, naccq);
    return -1;
}

}  // namespace hn
