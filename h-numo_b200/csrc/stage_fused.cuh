// stage_fused.cuh -- optimised fused barotropic stage kernel: ONE WARP PER ELEMENT, compile-time polynomial order.
//
// Same operator as k_btp_stage_simple (btp_kernels.cuh; reference src/mod_rhs_btp.F90:28-370,
// src/mod_laplacian_quad.F90:32-121,357-519, src/mod_rk_mlswe.F90:87-114), reorganised for the B200 SM:
//   * no block-wide barriers: a warp owns an element, phases are separated by __syncwarp() only;
//   * every 1-D contraction of the sum factorisation is done "one line per lane": a lane loads one line of
//     NGL (or NQ) values from shared memory and produces the whole output line with fully unrolled DFMAs whose
//     matrix operand is a compile-time constant-bank address (c[bank][imm]) -- no matrix loads, no index arithmetic;
//   * volume, face, LDG-viscosity, SSPRK update, wall projection, time-average sums and the trace epilogue are one
//     kernel: each state/coefficient/accumulator word crosses HBM once per stage.
// The kernel is bandwidth bound by design (about 1.3 flop/B); tensor cores are not applicable (FP64, 5x9 operators).
#pragma once
#include "btp_kernels.cuh"

namespace hn {

template <int G, int Q>
struct FusedLayout {
    static constexpr int NP = G * G, NQ2 = Q * Q, PER = G * Q;
    static constexpr int cmax(int a, int b) { return a > b ? a : b; }
    // nodal fields: 0 dpp 1 mx 2 my 3 pb 4 pp 5 up 6 vp 7 u 8 v
    static constexpr int NOD = 0;
    static constexpr int NOD_SZ = 9 * NP;
    // X: quadrature-point work (qv/fq), later rhs parts + face traces
    static constexpr int X = NOD + NOD_SZ;
    static constexpr int X_RHS = 0;                       // 6*NP   rhsP[3], rhsR[3]
    static constexpr int X_OWN = 6 * NP;                  // 4*8*G  own traces   [s][v][n], v: 0 dpp 1 mx 2 my 3..6 G 7 pb
    static constexpr int X_NBT = X_OWN + 32 * G;          // 4*8*G  neighbour traces
    static constexpr int X_NBV = X_NBT + 32 * G;          // 4*5*G  neighbour viscosity statics (own ones are re-read from L1)
    static constexpr int X_LF = X_NBV + 20 * G;           // 4*2*G  LDG face flux
    static constexpr int X_SZ = cmax(8 * NQ2, X_LF + 8 * G);
    // T: pass-1 intermediates, later face quadrature data
    static constexpr int T = X + X_SZ;
    static constexpr int T_FQV = 0;                       // 4*2*4*Q interpolated traces [s][side][var][iq]
    static constexpr int T_FF = 32 * Q;                   // 4*3*Q   face fluxes
    static constexpr int T_PROJ = T_FQV;                  // 4*3*G   projected face fluxes (the interpolated traces are dead by then)
    static constexpr int T_SZ = cmax(8 * PER, T_FF + 12 * Q);
    // L: LDG work: 0..3 dxi(u),dxi(v),det(u),det(v), later lapX(2),lapE(2) | 4..7 G | 8..11 Zxi(2),Zeta(2)
    static constexpr int L = T + T_SZ;
    static constexpr int L_SZ = 12 * NP;
    static constexpr int TOTAL = L + L_SZ;
};

// out[i] = sum_n M(n,i) in[n]   (nodes -> quadrature points), M = psiq or dpsiq in the constant bank
template <int G, int Q, bool DERIV>
__device__ __forceinline__ void line_n2q(const double (&in)[G], double (&out)[Q]) {
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        double s = 0.0;
#pragma unroll
        for (int n = 0; n < G; ++n) s = fma(DERIV ? c_ops.B[n + G * i] : c_ops.A[n + G * i], in[n], s);
        out[i] = s;
    }
}
// out[n] = sum_i M(n,i) in[i]   (quadrature points -> nodes, the transposed operator of the weak form)
template <int G, int Q, bool DERIV>
__device__ __forceinline__ void line_q2n(const double (&in)[Q], double (&out)[G]) {
#pragma unroll
    for (int n = 0; n < G; ++n) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < Q; ++i) s = fma(DERIV ? c_ops.B[n + G * i] : c_ops.A[n + G * i], in[i], s);
        out[n] = s;
    }
}
// collocation derivative along a nodal line: out[n] = sum_k D(k,n) in[k]
template <int G>
__device__ __forceinline__ void line_grad(const double (&in)[G], double (&out)[G]) {
#pragma unroll
    for (int n = 0; n < G; ++n) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < G; ++k) s = fma(c_ops.D[k + G * n], in[k], s);
        out[n] = s;
    }
}
// transposed: out[n'] = sum_n D(n',n) in[n]
template <int G>
__device__ __forceinline__ void line_gradT(const double (&in)[G], double (&out)[G]) {
#pragma unroll
    for (int np = 0; np < G; ++np) {
        double s = 0.0;
#pragma unroll
        for (int n = 0; n < G; ++n) s = fma(c_ops.D[np + G * n], in[n], s);
        out[np] = s;
    }
}

// gradient lines of the nodal fields u (nod[7]) and v (nod[8]): Lr[0],Lr[1] = d/dksi, Lr[2],Lr[3] = d/deta
template <int G>
__device__ __forceinline__ void ldg_gradient_lines(const double* nod, double* Lr, int lane) {
    constexpr int NP = G * G;
    for (int it = lane; it < 4 * G; it += 32) {
        int kind = it / (2 * G), r = it - kind * 2 * G, f = r / G, l = r - f * G;
        const double* src = nod + (7 + f) * NP;
        double in[G], out[G];
        if (kind == 0) {
#pragma unroll
            for (int k = 0; k < G; ++k) in[k] = src[l * G + k];  // row m = l
            line_grad<G>(in, out);
#pragma unroll
            for (int n = 0; n < G; ++n) Lr[f * NP + l * G + n] = out[n];
        } else {
#pragma unroll
            for (int k = 0; k < G; ++k) in[k] = src[k * G + l];  // column n = l
            line_grad<G>(in, out);
#pragma unroll
            for (int m = 0; m < G; ++m) Lr[(2 + f) * NP + m * G + l] = out[m];
        }
    }
}


// L2 prefetch of the chunk [first, first+n) doubles of a plane: one 128-byte line per lane (n*8 <= 32*128)
__device__ __forceinline__ void pf_chunk(const double* base, size_t first, int n, int lane) {
    if (base == nullptr) return;
    const char* lo = (const char*)(base + first);
    const char* hi = lo + (size_t)n * sizeof(double);
    const char* p = (const char*)((uintptr_t)lo & ~(uintptr_t)127) + (size_t)lane * 128;
    if (p < hi) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// Software prefetch into L2 of everything the block `pf_blocks` launches ahead will read (the kernel is otherwise
// bound by the latency of its ~14 dependent DRAM round trips per element, see profiles/r1_fused_kernel_summary.md).
template <int G, int Q>
__device__ __forceinline__ void stage_prefetch(const StageArgs& a, int e0, int ne, int warp, int nwarps, int lane) {
    constexpr int NP = G * G, NQ2 = Q * Q;
    const size_t n0 = (size_t)e0 * NP, q0 = (size_t)e0 * NQ2, s0 = (size_t)e0 * 4 * Q, t0 = (size_t)e0 * 4 * G;
    const int nn = ne * NP, nq = ne * NQ2, ns = ne * 4 * Q, nt = ne * 4 * G;
    for (int set = warp; set < 4; set += nwarps) {
    if (set == 0) {
#pragma unroll
        for (int v = 0; v < 3; ++v) pf_chunk(a.qb[v], n0, nn, lane);
        pf_chunk(a.pbprime_df, n0, nn, lane); pf_chunk(a.oop_df, n0, nn, lane); pf_chunk(a.massinv, n0, nn, lane);
#pragma unroll
        for (int v = 0; v < 6; ++v) pf_chunk(a.acc_n[v], n0, nn, lane);
        if (a.botfr) { pf_chunk(a.qp_dp, n0, nn, lane); pf_chunk(a.qp_u, n0, nn, lane); pf_chunk(a.qp_v, n0, nn, lane); }
        if (a.has_visc) {
            pf_chunk(a.pbv, n0, nn, lane);
#pragma unroll
            for (int v = 0; v < 4; ++v) pf_chunk(a.bdg[v], n0, nn, lane);
            if (a.acc_graduvb) {
#pragma unroll
                for (int v = 6; v < 10; ++v) pf_chunk(a.acc_n[v], n0, nn, lane);
            }
        }
        if (a.load_q0) {
#pragma unroll
            for (int v = 0; v < 3; ++v) pf_chunk(a.qb0[v], n0, nn, lane);
        }
        if (a.load_q2) {
#pragma unroll
            for (int v = 0; v < 3; ++v) pf_chunk(a.qb2[v], n0, nn, lane);
        }
    } else if (set == 1) {
        pf_chunk(a.coriolis_q, q0, nq, lane); pf_chunk(a.tauwx_q, q0, nq, lane); pf_chunk(a.tauwy_q, q0, nq, lane);
        pf_chunk(a.gzx_q, q0, nq, lane); pf_chunk(a.gzy_q, q0, nq, lane); pf_chunk(a.oop_q, q0, nq, lane);
        pf_chunk(a.Hbcl, q0, nq, lane); pf_chunk(a.Quu, q0, nq, lane); pf_chunk(a.Quv, q0, nq, lane); pf_chunk(a.Qvv, q0, nq, lane);
    } else if (set == 2) {
#pragma unroll
        for (int v = 0; v < 6; ++v) pf_chunk(a.acc_q[v], q0, nq, lane);
        if (a.botfr == 2) { pf_chunk(a.acc_q[6], q0, nq, lane); pf_chunk(a.acc_q[7], q0, nq, lane); }
        const int ntr = a.has_visc ? 7 : 3;
        for (int v = 0; v < ntr; ++v) pf_chunk(a.tr_in + (size_t)v * a.trstride, t0, nt, lane);
        pf_chunk(a.pbn, t0, nt, lane);
    } else {
        pf_chunk(a.cL, s0, ns, lane); pf_chunk(a.cR, s0, ns, lane); pf_chunk(a.cLR, s0, ns, lane); pf_chunk(a.lam, s0, ns, lane);
        pf_chunk(a.oop_edge, s0, ns, lane); pf_chunk(a.Quu_e, s0, ns, lane); pf_chunk(a.Quv_e, s0, ns, lane);
        pf_chunk(a.Qvv_e, s0, ns, lane); pf_chunk(a.Hbcl_e, s0, ns, lane); pf_chunk(a.pbl, s0, ns, lane); pf_chunk(a.pbr, s0, ns, lane);
#pragma unroll
        for (int v = 0; v < 11; ++v) pf_chunk(a.acc_f[v], s0, ns, lane);
    }
    }
}

template <int G, int Q>
__global__ void __maxnreg__(112) k_btp_stage_fused(StageArgs a) {
    using LAY = FusedLayout<G, Q>;
    constexpr int NP = LAY::NP, NQ2 = LAY::NQ2, PER = LAY::PER;
    extern __shared__ double sm_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = blockIdx.x * (blockDim.x >> 5) + warp;
    if (a.pf_blocks != 0 && !a.rhs_only) {   // > 0: blocks ahead;  < 0: this block's own later-phase data
        const int wpb = blockDim.x >> 5;
        const long e0 = ((long)blockIdx.x + (a.pf_blocks > 0 ? a.pf_blocks : 0)) * wpb;
        if (e0 < a.M.nelem) stage_prefetch<G, Q>(a, (int)e0, min(wpb, a.M.nelem - (int)e0), warp, wpb, lane);
    }
    if (e >= a.M.nelem) return;
    double* sm = sm_all + (size_t)warp * LAY::TOTAL;
    double* nod = sm + LAY::NOD;
    double* X = sm + LAY::X;
    double* T = sm + LAY::T;
    double* Lr = sm + LAY::L;
    const double ksx = a.M.em[e * 5 + 0], ksy = a.M.em[e * 5 + 1], etx = a.M.em[e * 5 + 2], ety = a.M.em[e * 5 + 3], J = a.M.em[e * 5 + 4];
    const size_t nbase = (size_t)e * NP, qbase = (size_t)e * NQ2;
    const bool acc = !a.rhs_only;
    const bool visc = a.has_visc != 0;
    const int NFI = a.botfr ? 7 : 4;

    // ---- 1. nodal loads + nodal sums (mod_rk_mlswe.F90:90-92)
    for (int I = lane; I < NP; I += 32) {
        double dpp = a.qb[0][nbase + I], mx = a.qb[1][nbase + I], my = a.qb[2][nbase + I];
        double pb = dpp + a.pbprime_df[nbase + I];
        nod[0 * NP + I] = dpp; nod[1 * NP + I] = mx; nod[2 * NP + I] = my; nod[3 * NP + I] = pb;
        if (a.botfr) { nod[4 * NP + I] = a.qp_dp[nbase + I]; nod[5 * NP + I] = a.qp_u[nbase + I]; nod[6 * NP + I] = a.qp_v[nbase + I]; }
        double rpb = 1.0 / pb;
        double u = mx * rpb, v = my * rpb;
        nod[7 * NP + I] = u; nod[8 * NP + I] = v;
        if (acc) {
            // all accumulator loads are issued before the first store (the stores may alias for the compiler)
            double t = 1.0 + dpp * a.oop_df[nbase + I];
            double c0 = a.acc_n[0][nbase + I], c1 = a.acc_n[1][nbase + I], c2 = a.acc_n[2][nbase + I];
            double c3 = a.acc_n[3][nbase + I], c4 = a.acc_n[4][nbase + I], c5 = a.acc_n[5][nbase + I];
            a.acc_n[0][nbase + I] = c0 + t * t; a.acc_n[1][nbase + I] = c1 + u; a.acc_n[2][nbase + I] = c2 + v;
            a.acc_n[3][nbase + I] = c3 + dpp; a.acc_n[4][nbase + I] = c4 + mx; a.acc_n[5][nbase + I] = c5 + my;
        }
    }
    __syncwarp();
    // ---- 2. sum factorisation, pass 1: one nodal row per lane -> T[f][m][i]
    for (int it = lane; it < NFI * G; it += 32) {
        int f = it / G, m = it - f * G;
        double in[G], out[Q];
#pragma unroll
        for (int n = 0; n < G; ++n) in[n] = nod[f * NP + m * G + n];
        line_n2q<G, Q, false>(in, out);
#pragma unroll
        for (int i = 0; i < Q; ++i) T[f * PER + m * Q + i] = out[i];
    }
    if (visc) ldg_gradient_lines<G>(nod, Lr, lane);
    __syncwarp();
    // ---- 3. pass 2: one quadrature column per lane -> X[f][j][i]
    for (int it = lane; it < NFI * Q; it += 32) {
        int f = it / Q, i = it - f * Q;
        double in[G], out[Q];
#pragma unroll
        for (int m = 0; m < G; ++m) in[m] = T[f * PER + m * Q + i];
        line_n2q<G, Q, false>(in, out);
#pragma unroll
        for (int j = 0; j < Q; ++j) X[f * NQ2 + j * Q + i] = out[j];
    }
    // LDG: G = grad(ub,vb) at the nodes, flux variable qq and its weighted metric combinations
    if (visc) {
        for (int I = lane; I < NP; I += 32) {
            int m = I / G, n = I - m * G;
            double dku = Lr[0 * NP + I], dkv = Lr[1 * NP + I], deu = Lr[2 * NP + I], dev = Lr[3 * NP + I];
            double g0 = ksx * dku + etx * deu, g1 = ksy * dku + ety * deu, g2 = ksx * dkv + etx * dev, g3 = ksy * dkv + ety * dev;
            Lr[4 * NP + I] = g0; Lr[5 * NP + I] = g1; Lr[6 * NP + I] = g2; Lr[7 * NP + I] = g3;
            double pv = a.pbv[nbase + I];
            double b0 = a.bdg[0][nbase + I], b1 = a.bdg[1][nbase + I], b2 = a.bdg[2][nbase + I], b3 = a.bdg[3][nbase + I];
            if (acc && a.acc_graduvb) {
                double c0 = a.acc_n[6][nbase + I], c1 = a.acc_n[7][nbase + I], c2 = a.acc_n[8][nbase + I], c3 = a.acc_n[9][nbase + I];
                a.acc_n[6][nbase + I] = c0 + g0; a.acc_n[7][nbase + I] = c1 + g1; a.acc_n[8][nbase + I] = c2 + g2; a.acc_n[9][nbase + I] = c3 + g3;
            }
            double q0 = pv * g0 + b0, q1 = pv * g1 + b1;
            double q2 = pv * g2 + b2, q3 = pv * g3 + b3;
            double w = c_ops.wg[n] * c_ops.wg[m] * J;
            Lr[8 * NP + I] = w * (ksx * q0 + ksy * q1); Lr[9 * NP + I] = w * (ksx * q2 + ksy * q3);
            Lr[10 * NP + I] = w * (etx * q0 + ety * q1); Lr[11 * NP + I] = w * (etx * q2 + ety * q3);
        }
    }
    __syncwarp();
    // ---- 4. pointwise physics at the quadrature points (mod_rhs_btp.F90:136-192); fq overwrites qv in place
    for (int q = lane; q < NQ2; q += 32) {
        int j = q / Q, i = q - j * Q;
        size_t Iq = qbase + q;
        double dpp = X[0 * NQ2 + q], udp = X[1 * NQ2 + q], vdp = X[2 * NQ2 + q], dp = X[3 * NQ2 + q];
        double wq = c_ops.wq[i] * c_ops.wq[j] * J;
        double rdp = 1.0 / dp;
        double ub = udp * rdp, vb = vdp * rdp;
        double tb_u = 0.0, tb_v = 0.0;
        if (a.botfr) {
            double pp = X[4 * NQ2 + q], up = X[5 * NQ2 + q], vp = X[6 * NQ2 + q];
            double ubot = up + ub, vbot = vp + vb;
            double spd = (a.botfr == 1) ? (a.cd / a.g) * pp : (a.cd / a.alpha_bot) * sqrt(ubot * ubot + vbot * vbot);
            tb_u = spd * ubot; tb_v = spd * vbot;
        }
        double fcor = a.coriolis_q[Iq];
        double s_twx = a.tauwx_q[Iq], s_twy = a.tauwy_q[Iq], s_gzx = a.gzx_q[Iq], s_gzy = a.gzy_q[Iq], s_oop = a.oop_q[Iq];
        double s_H = a.Hbcl[Iq], s_uu = a.Quu[Iq], s_uv = a.Quv[Iq], s_vv = a.Qvv[Iq];
        double c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0;
        if (acc) {
            c0 = a.acc_q[0][Iq]; c1 = a.acc_q[1][Iq]; c2 = a.acc_q[2][Iq]; c3 = a.acc_q[3][Iq]; c4 = a.acc_q[4][Iq]; c5 = a.acc_q[5][Iq];
            if (a.botfr == 2) { c6 = a.acc_q[6][Iq]; c7 = a.acc_q[7][Iq]; }
        }
        double sc_x = fcor * vdp + a.g * (s_twx - tb_u) - a.g * dp * s_gzx;
        double sc_y = -fcor * udp + a.g * (s_twy - tb_v) - a.g * dp * s_gzy;
        double ope = 1.0 + dpp * s_oop;
        double ope2 = ope * ope;
        double Hq = ope2 * s_H;
        double qu = ub * udp + ope * s_uu;
        double quv = ub * vdp + ope * s_uv;
        double qv = vb * vdp + ope * s_vv;
        if (acc) {
            a.acc_q[0][Iq] = c0 + qu; a.acc_q[1][Iq] = c1 + qv; a.acc_q[2][Iq] = c2 + quv; a.acc_q[3][Iq] = c3 + ope2;
            a.acc_q[4][Iq] = c4 + ub; a.acc_q[5][Iq] = c5 + vb;
            if (a.botfr == 2) { a.acc_q[6][Iq] = c6 + tb_u; a.acc_q[7][Iq] = c7 + tb_v; }
        }
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
    __syncwarp();
    // ---- 5. scatter pass 1: contraction over j, one quadrature column per lane -> T[arr][m][i]
    {
        // arrays contracted with psiq: Fk1 S2 Fk2 S3 Fk3 ; with dpsiq: Fe1 Fe2 Fe3
        for (int it = lane; it < 5 * Q; it += 32) {
            int k = it / Q, i = it - k * Q;
            int arr = (k == 0) ? 0 : (k == 1) ? 2 : (k == 2) ? 3 : (k == 3) ? 5 : 6;
            double in[Q], out[G];
#pragma unroll
            for (int j = 0; j < Q; ++j) in[j] = X[arr * NQ2 + j * Q + i];
            line_q2n<G, Q, false>(in, out);
#pragma unroll
            for (int m = 0; m < G; ++m) T[arr * PER + m * Q + i] = out[m];
        }
        for (int it = lane; it < 3 * Q; it += 32) {
            int k = it / Q, i = it - k * Q;
            int arr = (k == 0) ? 1 : (k == 1) ? 4 : 7;
            double in[Q], out[G];
#pragma unroll
            for (int j = 0; j < Q; ++j) in[j] = X[arr * NQ2 + j * Q + i];
            line_q2n<G, Q, true>(in, out);
#pragma unroll
            for (int m = 0; m < G; ++m) T[arr * PER + m * Q + i] = out[m];
        }
    }
    __syncwarp();
    // ---- 6. scatter pass 2: contraction over i, one nodal row per lane -> X_RHS: rhsP[f], rhsR[f]
    {
        double* R = X + LAY::X_RHS;
        for (int it = lane; it < 3 * G; it += 32) {  // dpsiq part: sum_i B(n,i) Fk_f[m][i]
            int f = it / G, m = it - f * G;
            double in[Q], out[G];
#pragma unroll
            for (int i = 0; i < Q; ++i) in[i] = T[(3 * f) * PER + m * Q + i];
            line_q2n<G, Q, true>(in, out);
#pragma unroll
            for (int n = 0; n < G; ++n) R[f * NP + m * G + n] = out[n];
        }
        for (int it = lane; it < 3 * G; it += 32) {  // psiq part: sum_i A(n,i) (Fe_f + S_f)[m][i]
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
    }
    // LDG volume lines (btp_compute_laplacian): lapX[c][m'][n'] = sum_n D(n',n) Zxi_c[m'][n] ; lapE[c][m'][n'] = sum_m D(m',m) Zeta_c[m][n']
    if (visc) {
        for (int it = lane; it < 4 * G; it += 32) {
            int kind = it / (2 * G), r = it - kind * 2 * G, c = r / G, l = r - c * G;
            double in[G], out[G];
            if (kind == 0) {
#pragma unroll
                for (int n = 0; n < G; ++n) in[n] = Lr[(8 + c) * NP + l * G + n];
                line_gradT<G>(in, out);
#pragma unroll
                for (int n = 0; n < G; ++n) Lr[(0 + c) * NP + l * G + n] = out[n];
            } else {
#pragma unroll
                for (int m = 0; m < G; ++m) in[m] = Lr[(10 + c) * NP + m * G + l];
                line_gradT<G>(in, out);
#pragma unroll
                for (int m = 0; m < G; ++m) Lr[(2 + c) * NP + m * G + l] = out[m];
            }
        }
    }
    // ---- 7a. face traces: own and neighbour (btp_extract_df), LDG gradient traces, viscosity statics
    {
        double* own = X + LAY::X_OWN; double* nbt = X + LAY::X_NBT; double* nbv = X + LAY::X_NBV;
        for (int it = lane; it < 4 * G; it += 32) {
            int s = it / G, n = it - s * G;
            int slot = e * 4 + s, nb = a.M.nbr[slot], nbs = a.M.nbslot[slot];
            double nx = a.M.fgeom[slot * 3 + 0], ny = a.M.fgeom[slot * 3 + 1];
            int I = face_node(s, n, G);
            double ow[3] = {nod[0 * NP + I], nod[1 * NP + I], nod[2 * NP + I]}, nbv3[3];
            neighbour_state(a, e, s, n, nb, nbs, nx, ny, ow, nbv3);
            double* po = own + (s * 8) * G + n;
            double* pn = nbt + (s * 8) * G + n;
#pragma unroll
            for (int v = 0; v < 3; ++v) { po[v * G] = ow[v]; pn[v * G] = nbv3[v]; }
            po[7 * G] = nod[3 * NP + I];
            pn[7 * G] = nbv3[0] + a.pbn[(size_t)slot * G + n];
            if (visc) {
                double go[4] = {Lr[4 * NP + I], Lr[5 * NP + I], Lr[6 * NP + I], Lr[7 * NP + I]}, gn[4];
                double so[5] = {a.bdg[0][nbase + I], a.bdg[1][nbase + I], a.bdg[2][nbase + I], a.bdg[3][nbase + I], a.pbv[nbase + I]}, sn[5];
                if (nb >= 0) {
                    size_t base = ((size_t)nb * 4 + nbs) * G + n;
#pragma unroll
                    for (int v = 0; v < 4; ++v) gn[v] = a.tr_in[(TR_G + v) * a.trstride + base];
                    size_t In = (size_t)nb * NP + face_node(nbs, n, G);
#pragma unroll
                    for (int v = 0; v < 4; ++v) sn[v] = a.bdg[v][In];
                    sn[4] = a.pbv[In];
                } else if (nb == NBR_HALO) {
                    size_t base = ((size_t)a.M.nslots + nbs) * G + n;
#pragma unroll
                    for (int v = 0; v < 4; ++v) gn[v] = a.tr_in[(TR_G + v) * a.trstride + base];
#pragma unroll
                    for (int v = 0; v < 5; ++v) sn[v] = a.hstat[v * a.hstat_stride + (size_t)nbs * G + n];
                } else {
#pragma unroll
                    for (int v = 0; v < 4; ++v) { gn[v] = go[v]; sn[v] = so[v]; }
                    sn[4] = so[4];
                    if (nb == NBR_FREESLIP) { reflect4(go, nx, ny, gn); reflect4(so, nx, ny, sn); }
                }
#pragma unroll
                for (int v = 0; v < 4; ++v) { po[(3 + v) * G] = go[v]; pn[(3 + v) * G] = gn[v]; }
#pragma unroll
                for (int v = 0; v < 5; ++v) nbv[(s * 5 + v) * G + n] = sn[v];
            }
        }
    }
    __syncwarp();
    // ---- 7b. interpolate the traces to the face quadrature points: one (slot, side, variable) line per lane
    {
        const double* own = X + LAY::X_OWN; const double* nbt = X + LAY::X_NBT;
        double* fqv = T + LAY::T_FQV;
        for (int it = lane; it < 32; it += 32) {
            int s = it >> 3, side = (it >> 2) & 1, var = it & 3;  // var: 0 pb 1 dpp 2 mx 3 my
            int slot = e * 4 + s, nb = a.M.nbr[slot];
            bool left = (nb < 0) || (e < nb);
            const double* src = ((side == 0) == left ? own : nbt) + (s * 8 + (var == 0 ? 7 : var - 1)) * G;
            double in[G], out[Q];
#pragma unroll
            for (int n = 0; n < G; ++n) in[n] = src[n];
            line_n2q<G, Q, false>(in, out);
#pragma unroll
            for (int iq = 0; iq < Q; ++iq) fqv[it * Q + iq] = out[iq];
        }
    }
    __syncwarp();
    // ---- 7c. face fluxes, canonical left perspective (mod_rhs_btp.F90:237-330)
    {
        const double* fqv = T + LAY::T_FQV;
        double* ff = T + LAY::T_FF;
        for (int it = lane; it < 4 * Q; it += 32) {
            int s = it / Q, iq = it - s * Q;
            int slot = e * 4 + s, nb = a.M.nbr[slot], nbs = a.M.nbslot[slot];
            bool left = (nb < 0) || (e < nb);
            int oslot = left ? slot : nb * 4 + nbs;
            double nxl = a.M.fgeom[slot * 3 + 0], nyl = a.M.fgeom[slot * 3 + 1], nlen = a.M.fgeom[slot * 3 + 2];
            const double* L = fqv + (s * 8) * Q + iq;
            const double* R = fqv + (s * 8 + 4) * Q + iq;
            double pbL = L[0], ppL = L[Q], mxL = L[2 * Q], myL = L[3 * Q];
            double pbR = R[0], ppR = R[Q], mxR = R[2 * Q], myR = R[3 * Q];
            size_t fo = (size_t)oslot * Q + iq;
            double cL = a.cL[fo], cR = a.cR[fo], cLR = a.cLR[fo], lam = a.lam[fo];
            const double s_oope = a.oop_edge[fo], s_uue = a.Quu_e[fo], s_uve = a.Quv_e[fo], s_vve = a.Qvv_e[fo], s_He = a.Hbcl_e[fo];
            const bool upd = acc && left;
            double fa[11], s_pbl = 1.0, s_pbr = 1.0;
            if (upd) {
                s_pbl = a.pbl[fo]; s_pbr = a.pbr[fo];
#pragma unroll
                for (int v = 0; v < 11; ++v) fa[v] = a.acc_f[v][fo];
            }
            double pU_L = nxl * mxL + nyl * myL;
            double pU_R = -nxl * mxR - nyl * myR;
            double pbpert_edge = cL * ppL + cR * ppR + cLR * (pU_L + pU_R);
            double ope_e = 1.0 + pbpert_edge * s_oope;
            double fex = cR * mxL + cL * mxR + lam * (nxl * ppL - nxl * ppR);
            double fey = cR * myL + cL * myR + lam * (nyl * ppL - nyl * ppR);
            double rl = 1.0 / pbL, rr = 1.0 / pbR;
            double ul = mxL * rl, ur = mxR * rr, vl = myL * rl, vr = myR * rr;
            double quu = 0.5 * (ul * mxL + ur * mxR) + ope_e * s_uue;
            double quv = 0.5 * (vl * mxL + vr * mxR) + ope_e * s_uve;
            double qvu = 0.5 * (ul * myL + ur * myR) + ope_e * s_uve;
            double qvv = 0.5 * (vl * myL + vr * myR) + ope_e * s_vve;
            double e2 = ope_e * ope_e;
            double Hf = e2 * s_He;
            if (upd) {
                double ol = 1.0 + (ppL / s_pbl), orr = 1.0 + (ppR / s_pbr);
                a.acc_f[0][fo] = fa[0] + quu; a.acc_f[1][fo] = fa[1] + quv; a.acc_f[2][fo] = fa[2] + qvu; a.acc_f[3][fo] = fa[3] + qvv;
                a.acc_f[4][fo] = fa[4] + ol * ol; a.acc_f[5][fo] = fa[5] + orr * orr; a.acc_f[6][fo] = fa[6] + e2;
                a.acc_f[7][fo] = fa[7] + ul; a.acc_f[8][fo] = fa[8] + ur; a.acc_f[9][fo] = fa[9] + vl; a.acc_f[10][fo] = fa[10] + vr;
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
        // ---- 7e. LDG face flux at the face nodes, as written (mod_laplacian_quad.F90:427-519)
        if (visc) {
            const double* own = X + LAY::X_OWN; const double* nbt = X + LAY::X_NBT; const double* nbv = X + LAY::X_NBV;
            double* lf = X + LAY::X_LF;
            for (int it = lane; it < 4 * G; it += 32) {
                int s = it / G, n = it - s * G;
                int slot = e * 4 + s, nb = a.M.nbr[slot];
                bool left = (nb < 0) || (e < nb);
                double nx = a.M.fgeom[slot * 3 + 0], ny = a.M.fgeom[slot * 3 + 1], nlen = a.M.fgeom[slot * 3 + 2];
                const double* go = own + (s * 8 + 3) * G + n;
                const double* gn = nbt + (s * 8 + 3) * G + n;
                const double* sn = nbv + s * 5 * G + n;
                const int I = face_node(s, n, G);
                const double pvo = a.pbv[nbase + I];
                double fo_[4], fn_[4];
#pragma unroll
                for (int v = 0; v < 4; ++v) { fo_[v] = pvo * go[v * G] + a.bdg[v][nbase + I]; fn_[v] = sn[4 * G] * gn[v * G] + sn[v * G]; }
                const double* fl = left ? fo_ : fn_;
                const double* fr = left ? fn_ : fo_;
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
    __syncwarp();
    // ---- 7d. project the face fluxes onto the face nodes: one (slot, field) line per lane
    {
        const double* ff = T + LAY::T_FF;
        double* proj = T + LAY::T_PROJ;
        for (int it = lane; it < 12; it += 32) {
            double in[Q], out[G];
#pragma unroll
            for (int iq = 0; iq < Q; ++iq) in[iq] = ff[it * Q + iq];
            line_q2n<G, Q, false>(in, out);
#pragma unroll
            for (int n = 0; n < G; ++n) proj[it * G + n] = out[n];
        }
    }
    __syncwarp();
    // ---- 8. gather per node, mass matrix, viscosity, SSPRK update, wall projection (mod_rk_mlswe.F90:97-108)
    {
        const double* R = X + LAY::X_RHS;
        const double* proj = T + LAY::T_PROJ;
        const double* lf = X + LAY::X_LF;
        for (int I = lane; I < NP; I += 32) {
            int m = I / G, n = I - m * G;
            double r0 = R[0 * NP + I] + R[3 * NP + I], r1 = R[1 * NP + I] + R[4 * NP + I], r2 = R[2 * NP + I] + R[5 * NP + I];
            double l0 = 0.0, l1 = 0.0;
            if (visc) { l0 = -(Lr[0 * NP + I] + Lr[2 * NP + I]); l1 = -(Lr[1 * NP + I] + Lr[3 * NP + I]); }
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                bool on = (s == 0) ? (m == 0) : (s == 1) ? (m == G - 1) : (s == 2) ? (n == 0) : (n == G - 1);
                if (!on) continue;
                int nf = (s < 2) ? n : m;
                r0 += proj[(s * 3 + 0) * G + nf]; r1 += proj[(s * 3 + 1) * G + nf]; r2 += proj[(s * 3 + 2) * G + nf];
                if (visc) { l0 += lf[(s * 2 + 0) * G + nf]; l1 += lf[(s * 2 + 1) * G + nf]; }
            }
            double mi = a.massinv[nbase + I];
            r0 = mi * r0; r1 = mi * r1; r2 = mi * r2;
            if (visc) { r1 = r1 + a.visc * mi * l0; r2 = r2 + a.visc * mi * l1; }
            if (a.rhs_only) {
                a.rhs_out[0][nbase + I] = r0; a.rhs_out[1][nbase + I] = r1; a.rhs_out[2][nbase + I] = r2;
            } else {
                double q1[3] = {nod[0 * NP + I], nod[1 * NP + I], nod[2 * NP + I]};
                double q0[3], q2[3] = {0.0, 0.0, 0.0};
#pragma unroll
                for (int v = 0; v < 3; ++v) q0[v] = a.load_q0 ? a.qb0[v][nbase + I] : q1[v];
                if (a.load_q2) {
#pragma unroll
                    for (int v = 0; v < 3; ++v) q2[v] = a.qb2[v][nbase + I];
                }
                double rr[3] = {r0, r1, r2}, qn[3];
#pragma unroll
                for (int v = 0; v < 3; ++v) qn[v] = a.a1 * q0[v] + a.a2 * q1[v] + a.a3 * q2[v] + a.dtt * rr[v];
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    bool on = (s == 0) ? (m == 0) : (s == 1) ? (m == G - 1) : (s == 2) ? (n == 0) : (n == G - 1);
                    if (!on) continue;
                    int slot = e * 4 + s, nb = a.M.nbr[slot];
                    if (nb == NBR_FREESLIP) {
                        double nx = a.M.fgeom[slot * 3 + 0], ny = a.M.fgeom[slot * 3 + 1];
                        double unl = qn[1] * nx + qn[2] * ny;
                        qn[1] = qn[1] - unl * nx; qn[2] = qn[2] - unl * ny;
                    } else if (nb == NBR_NOSLIP) { qn[1] = 0.0; qn[2] = 0.0; }
                }
                if (a.store_q0) {
#pragma unroll
                    for (int v = 0; v < 3; ++v) a.qb0[v][nbase + I] = q1[v];
                }
#pragma unroll
                for (int v = 0; v < 3; ++v) a.qb[v][nbase + I] = qn[v];
                if (a.store_q2) {
#pragma unroll
                    for (int v = 0; v < 3; ++v) a.qb2[v][nbase + I] = qn[v];
                }
                double pbn_ = qn[0] + a.pbprime_df[nbase + I];
                nod[0 * NP + I] = qn[0]; nod[1 * NP + I] = qn[1]; nod[2 * NP + I] = qn[2];
                nod[7 * NP + I] = qn[1] / pbn_; nod[8 * NP + I] = qn[2] / pbn_;
            }
        }
    }
    if (a.rhs_only) return;
    __syncwarp();
    // ---- 9. traces of the new state (+ LDG gradient) for the next stage
    if (visc) { ldg_gradient_lines<G>(nod, Lr, lane); __syncwarp(); }
    for (int it = lane; it < 4 * G; it += 32) {
        int s = it / G, n = it - s * G;
        int I = face_node(s, n, G);
        size_t base = ((size_t)e * 4 + s) * G + n;
        a.tr_out[TR_PBPERT * a.trstride + base] = nod[0 * NP + I];
        a.tr_out[TR_MX * a.trstride + base] = nod[1 * NP + I];
        a.tr_out[TR_MY * a.trstride + base] = nod[2 * NP + I];
        if (visc) {
            double dku = Lr[0 * NP + I], dkv = Lr[1 * NP + I], deu = Lr[2 * NP + I], dev = Lr[3 * NP + I];
            a.tr_out[(TR_G + 0) * a.trstride + base] = ksx * dku + etx * deu;
            a.tr_out[(TR_G + 1) * a.trstride + base] = ksy * dku + ety * deu;
            a.tr_out[(TR_G + 2) * a.trstride + base] = ksx * dkv + etx * dev;
            a.tr_out[(TR_G + 3) * a.trstride + base] = ksy * dkv + ety * dev;
        }
    }
}

inline void upload_fused_ops(const Ops&, int, int) {}
inline bool stage_fused_supported(const Solver& S) { return (S.ngl == 5 && S.nq == 9) || (S.ngl == 4 && S.nq == 7); }

template <int G, int Q>
static int launch_fused_t(Solver& S, const StageArgs& a) {
    using LAY = FusedLayout<G, Q>;
    const int warps = 3;   // 6 blocks of 3 warps per SM (shared memory bound)
    size_t smem = (size_t)warps * LAY::TOTAL * sizeof(double);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_btp_stage_fused<G, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            set_error("cudaFuncSetAttribute", "shared memory opt-in failed"); return -1;
        }
        configured = true;
    }
    int blocks = (S.nelem + warps - 1) / warps;
    k_btp_stage_fused<G, Q><<<blocks, warps * 32, smem, S.stream>>>(a);
    S.n_launches++;
    return 0;
}
inline int launch_stage_fused(Solver& S, const StageArgs& a) {
    if (S.ngl == 5 && S.nq == 9) return launch_fused_t<5, 9>(S, a);
    if (S.ngl == 4 && S.nq == 7) return launch_fused_t<4, 7>(S, a);
    return -1;
}

}  // namespace hn
