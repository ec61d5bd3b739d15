// layer_warp.cuh -- baroclinic (layer) kernels, one WARP per element.
//
// Same operators and the same point-wise arithmetic as the block-per-element kernels of bcl_kernels.cuh (which stay as the
// any-order / any-layer-count path and as the bisecting aid), re-expressed like the barotropic stage kernel:
//   * one warp advances one element, 4 independent warps per block: the phases of an element are separated by
//     __syncwarp only (the block-per-element form spent 3-8 issue slots per instruction waiting at __syncthreads with
//     14-36 % of the warp slots occupied: profiles/r2_layer_kernels_ncu.md);
//   * every 1-D contraction is "one line per lane" with the operator entry a uniform-register operand (line_ops.cuh):
//     about a third of the instructions of the index-computing loops of sf_pass1 / sf_eval / sf_scatter;
//   * a lane owns up to NQIT = ceil(nq^2/32) quadrature points and NFIT = ceil(4 nq/32) face quadrature points; values
//     that must survive the layer loop are per-point registers, everything else goes through the warp's shared-memory tiles.
// Instantiated for the orders with a warp-per-element stage kernel (nop 3 and 4); other orders run bcl_kernels.cuh.
#pragma once
#include "bcl_kernels.cuh"
#include "line_ops.cuh"

#ifndef LW_MINB_A
#define LW_MINB_A 4   // resident blocks per SM the register allocation of the coefficient / mass / consistency / laplacian kernels aims at
#endif
#ifndef LW_MINB_L
#define LW_MINB_L 4   // ... of the layer laplacian kernel
#endif
#ifndef LW_MINB_F
#define LW_MINB_F 4   // ... of the momentum face kernel (128 registers, 16 B of spills: 583 against 700 us per launch at 62 500 elements with 3 blocks
                      //     and 158 registers; the other kernels are slower with 3 blocks and equal with 5)
#endif
namespace hn {

constexpr int LW_WARPS = 4;   // warps (elements) per block

template <int G, int Q>
struct LW {
    using LG = LineGeom<G, Q>;
    static constexpr int NP = G * G, NQ2 = Q * Q, SX = LG::SX, ST = LG::ST, TM = LG::TM;
    static constexpr int NQIT = (NQ2 + 31) / 32, NFIT = (4 * Q + 31) / 32;
    static_assert(NP <= 32 && 4 * G <= 32, "warp-per-element layer kernels: order too high");
};

// nodal index of face node n on side s (compile-time G)
template <int G>
__device__ __forceinline__ int lw_face_node(int s, int n) { return s == 0 ? n : s == 1 ? (G - 1) * G + n : s == 2 ? n * G : n * G + G - 1; }

// ---- warp-level sum factorisation (all pointers: this warp's shared memory) -------------------------------------------------
// F nodal fields nod[f][NP] -> values at the quadrature points X[f][SX]; T[f][ST] is scratch
template <int G, int Q, int F>
__device__ __forceinline__ void lw_interp(const double* nod, double* T, double* X, int lane) {
    using W = LW<G, Q>;
#pragma unroll 1
    for (int job = lane; job < F * G; job += 32) {
        const int f = job / G, m = job - f * G;
        pl_n2q<1, G, Q, false, 1, 1>(nod + f * W::NP + m * G, T + f * W::ST + m * W::TM);
    }
    __syncwarp();
#pragma unroll 1
    for (int job = lane; job < F * Q; job += 32) {
        const int f = job / Q, i = job - f * Q;
        pl_n2q<1, G, Q, false, W::TM, Q>(T + f * W::ST + i, X + f * W::SX + i);
    }
    __syncwarp();
}
// reference-space gradient of F nodal fields at the quadrature points: Xk[f] = d/dksi, Xe[f] = d/deta; T[2F][ST] scratch
template <int G, int Q, int F>
__device__ __forceinline__ void lw_grad_q(const double* nod, double* T, double* Xk, double* Xe, int lane) {
    using W = LW<G, Q>;
#pragma unroll 1
    for (int job = lane; job < 2 * F * G; job += 32) {
        const int kind = job / (F * G), r = job - kind * F * G, f = r / G, m = r - f * G;
        if (kind == 0) pl_n2q<1, G, Q, true, 1, 1>(nod + f * W::NP + m * G, T + f * W::ST + m * W::TM);         // B in ksi
        else pl_n2q<1, G, Q, false, 1, 1>(nod + f * W::NP + m * G, T + (F + f) * W::ST + m * W::TM);             // A in ksi
    }
    __syncwarp();
#pragma unroll 1
    for (int job = lane; job < 2 * F * Q; job += 32) {
        const int kind = job / (F * Q), r = job - kind * F * Q, f = r / Q, i = r - f * Q;
        if (kind == 0) pl_n2q<1, G, Q, false, W::TM, Q>(T + f * W::ST + i, Xk + f * W::SX + i);                   // A in eta of the B-pass
        else pl_n2q<1, G, Q, true, W::TM, Q>(T + (F + f) * W::ST + i, Xe + f * W::SX + i);                        // B in eta of the A-pass
    }
    __syncwarp();
}
// weak-form scatter: out[f][m][n] = sum_ij  dpsi/dksi Fk[f] + dpsi/deta Fe[f] (+ psi S[f]); Fk, Fe, S are [f][SX] with weights
// and metric factors applied; T[2F][ST] scratch
template <int G, int Q, int F, bool HAS_S>
__device__ __forceinline__ void lw_scatter(const double* Fk, const double* Fe, const double* S, double* T, double* out, int lane) {
    using W = LW<G, Q>;
#pragma unroll 1
    for (int job = lane; job < 2 * F * Q; job += 32) {
        const int kind = job / (F * Q), r = job - kind * F * Q, f = r / Q, i = r - f * Q;
        double t[G][1];
        if (kind == 0) pl_q2n_acc<1, G, Q, false, Q, true>(Fk + f * W::SX + i, t);          // TB = A.Fk over j
        else {
            pl_q2n_acc<1, G, Q, true, Q, true>(Fe + f * W::SX + i, t);                      // TA = B.Fe (+ A.S) over j
            if (HAS_S) pl_q2n_acc<1, G, Q, false, Q, false>(S + f * W::SX + i, t);
        }
#pragma unroll
        for (int m = 0; m < G; ++m) T[(kind * F + f) * W::ST + m * W::TM + i] = t[m][0];
    }
    __syncwarp();
#pragma unroll 1
    for (int job = lane; job < F * G; job += 32) {
        const int f = job / G, m = job - f * G;
        double r[G][1];
        pl_q2n_acc<1, G, Q, true, 1, true>(T + f * W::ST + m * W::TM, r);                   // B.TB over i
        pl_q2n_acc<1, G, Q, false, 1, false>(T + (F + f) * W::ST + m * W::TM, r);           // + A.TA
#pragma unroll
        for (int n = 0; n < G; ++n) out[f * W::NP + m * G + n] = r[n][0];
    }
    __syncwarp();
}
// NL lines of G face-node values -> Q face quadrature points each: src[line][G] -> dst[line][Q]
template <int G, int Q>
__device__ __forceinline__ void lw_face_interp(const double* src, double* dst, int nlines, int lane) {
#pragma unroll 1
    for (int job = lane; job < nlines; job += 32) pl_n2q<1, G, Q, false, 1, 1>(src + job * G, dst + job * Q);
    __syncwarp();
}
// projection of face-quadrature data onto the face nodes: src[line][Q] -> dst[line][G]  (sum_iq psiq(n,iq) src[iq])
template <int G, int Q>
__device__ __forceinline__ void lw_face_project(const double* src, double* dst, int nlines, int lane) {
#pragma unroll 1
    for (int job = lane; job < nlines; job += 32) {
        double r[G][1];
        pl_q2n_acc<1, G, Q, false, 1, true>(src + job * Q, r);
#pragma unroll
        for (int n = 0; n < G; ++n) dst[job * G + n] = r[n][0];
    }
    __syncwarp();
}
// collocation gradient of NF nodal fields: out[(2*f + c)][NP], c = 0: d/dksi, 1: d/deta
template <int G, int NF>
__device__ __forceinline__ void lw_grad_n(const double* nod, int np, double* out, int lane) {
#pragma unroll 1
    for (int job = lane; job < 2 * NF * G; job += 32) {
        const int kind = job / (NF * G), r = job - kind * NF * G, f = r / G, l = r - f * G;
        const int stride = kind ? G : 1, off = kind ? l : l * G;
        pl_grad<1, G, false>(nod + f * np + off, out + (2 * f + kind) * np + off, stride);
    }
    __syncwarp();
}

// neighbour's nodal trace of (dp', u', v') of layer k at the face nodes, ghosts resolved (extract_qprime_df_face,
// mod_layer_terms.F90:354-415): lanes < 4G write own[(s*3+v)*G+n] and nbq[(s*3+v)*G+n]
template <int G>
__device__ __forceinline__ void lw_qprime_traces(const Mesh& M, const double* qprime, size_t nstride, const double* hq, size_t hstride, int nl, int k,
                                                 int e, const double* nod, double* own, double* nbq, int lane) {
    constexpr int NP = G * G;
    if (lane < 4 * G) {
        const int s = lane / G, n = lane - s * G, slot = e * 4 + s, nb = M.nbr[slot], nbs = M.nbslot[slot];
        const int I = lw_face_node<G>(s, n);
        double ow[3] = {nod[I], nod[NP + I], nod[2 * NP + I]}, nv[3];
#pragma unroll
        for (int v = 0; v < 3; ++v)
            nv[v] = nb_nodal(M, qprime + (size_t)(v * nl + k) * nstride, hq + (size_t)(v * nl + k) * hstride, nb, nbs, n, ow[v]);
        if (nb == NBR_FREESLIP) {
            const double nx = M.fgeom[slot * 3], ny = M.fgeom[slot * 3 + 1];
            const double un = ow[1] * nx + ow[2] * ny;
            nv[1] = ow[1] - 2.0 * un * nx; nv[2] = ow[2] - 2.0 * un * ny;
        } else if (nb == NBR_NOSLIP) { nv[1] = -ow[1]; nv[2] = -ow[2]; }
#pragma unroll
        for (int v = 0; v < 3; ++v) { own[(s * 3 + v) * G + n] = ow[v]; nbq[(s * 3 + v) * G + n] = nv[v]; }
    }
}

// ================================================================================================================================
// btp_bcl_coeffs_qdf (mod_barotropic_terms.F90:219-409)
template <int G, int Q, int NL_>
__global__ void __launch_bounds__(32 * LW_WARPS, LW_MINB_A) k_bcl_coeffs_w(CoeffArgs a) {
    using W = LW<G, Q>;
    constexpr int NP = W::NP, NQ2 = W::NQ2, SX = W::SX, ST = W::ST;
    constexpr int PER_WARP = 3 * NP + 4 * ST + 3 * SX + 24 * G + 24 * Q + 4 * NP;
    extern __shared__ double sm_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = blockIdx.x * LW_WARPS + warp;
    if (e >= a.M.nelem) return;
    const int nl = NL_ ? NL_ : a.M.nl;
    double* sm = sm_all + (size_t)warp * PER_WARP;
    double* nod = sm;                 // [3][NP]
    double* T = nod + 3 * NP;         // [4][ST]
    double* X = T + 4 * ST;           // [3][SX]
    double* own = X + 3 * SX;         // [4][3][G]
    double* nbq = own + 12 * G;       // [4][3][G]
    double* fq = nbq + 12 * G;        // [24][Q]: own lines 0..11, neighbour lines 12..23
    double* gr = fq + 24 * Q;         // [4][NP] collocation gradients of u', v'
    const size_t nbase = (size_t)e * NP, qbase = (size_t)e * NQ2;
    const double ksx = a.M.em[e * 5 + 0], ksy = a.M.em[e * 5 + 1], etx = a.M.em[e * 5 + 2], ety = a.M.em[e * 5 + 3];
    double Quu[W::NQIT], Quv[W::NQIT], Qvv[W::NQIT], H[W::NQIT], pprime[W::NQIT];
    double eu[W::NFIT], euv[W::NFIT], ev[W::NFIT], eH[W::NFIT], ppl[W::NFIT], ppr[W::NFIT];
#pragma unroll
    for (int it = 0; it < W::NQIT; ++it) { Quu[it] = 0; Quv[it] = 0; Qvv[it] = 0; H[it] = 0; pprime[it] = 0; }
#pragma unroll
    for (int it = 0; it < W::NFIT; ++it) { eu[it] = 0; euv[it] = 0; ev[it] = 0; eH[it] = 0; ppl[it] = 0; ppr[it] = 0; }
    double bsum[4] = {0, 0, 0, 0}, pvs = 0;
#pragma unroll 1
    for (int k = 0; k < nl; ++k) {
        __syncwarp();
        if (lane < NP) {
#pragma unroll
            for (int v = 0; v < 3; ++v) nod[v * NP + lane] = a.qprime[(size_t)(v * nl + k) * a.nstride + nbase + lane];
        }
        __syncwarp();
        lw_qprime_traces<G>(a.M, a.qprime, a.nstride, a.hq, a.hstride, nl, k, e, nod, own, nbq, lane);
        lw_interp<G, Q, 3>(nod, T, X, lane);
#pragma unroll
        for (int it = 0; it < W::NQIT; ++it) {
            const int q = min(it * 32 + lane, NQ2 - 1);
            const double q0 = X[q], q1 = X[SX + q], q2 = X[2 * SX + q];
            Quu[it] = Quu[it] + q1 * (q1 * q0);
            Quv[it] = Quv[it] + q2 * (q1 * q0);
            Qvv[it] = Qvv[it] + q2 * (q2 * q0);
            const double pn = pprime[it] + q0;
            H[it] = H[it] + 0.5 * a.alpha[k] * (pn * pn - pprime[it] * pprime[it]);
            pprime[it] = pn;
        }
        // owned faces: edge coefficients from left/right traces (mod_barotropic_terms.F90:306-337)
        lw_face_interp<G, Q>(own, fq, 24, lane);   // own and nbq are contiguous: 24 lines
#pragma unroll
        for (int it = 0; it < W::NFIT; ++it) {
            const int p = min(it * 32 + lane, 4 * Q - 1), s = p / Q, iq = p - s * Q;
            const double* o = fq + (s * 3) * Q + iq;
            const double* r = fq + (12 + s * 3) * Q + iq;
            const double ql0 = o[0], ql1 = o[Q], ql2 = o[2 * Q], qr0 = r[0], qr1 = r[Q], qr2 = r[2 * Q];
            eu[it] = eu[it] + 0.5 * ((ql1 * ql1 * ql0) + (qr1 * qr1 * qr0));
            euv[it] = euv[it] + 0.5 * ((ql2 * ql1 * ql0) + (qr2 * qr1 * qr0));
            ev[it] = ev[it] + 0.5 * ((ql2 * ql2 * ql0) + (qr2 * qr2 * qr0));
            const double pl = ppl[it] + ql0, pr = ppr[it] + qr0;
            const double left_dp = 0.5 * a.alpha[k] * (pl * pl - ppl[it] * ppl[it]);
            const double right_dp = 0.5 * a.alpha[k] * (pr * pr - ppr[it] * ppr[it]);
            eH[it] = eH[it] + 0.5 * (left_dp + right_dp);
            ppl[it] = pl; ppr[it] = pr;
        }
        // viscosity auxiliaries at the nodes (mod_barotropic_terms.F90:287-304)
        if (a.has_visc) {
            lw_grad_n<G, 2>(nod + NP, NP, gr, lane);
            if (lane < NP) {
                double gv[4];
                gv[0] = ksx * gr[lane] + etx * gr[NP + lane]; gv[1] = ksy * gr[lane] + ety * gr[NP + lane];
                gv[2] = ksx * gr[2 * NP + lane] + etx * gr[3 * NP + lane]; gv[3] = ksy * gr[2 * NP + lane] + ety * gr[3 * NP + lane];
                const double d = a.dpv[(size_t)k * a.nstride + nbase + lane];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const double t = d * gv[v];
                    a.dpp_graduv[(size_t)(v * nl + k) * a.nstride + nbase + lane] = t;
                    bsum[v] = bsum[v] + t;
                }
                pvs = pvs + d;
            }
        }
    }
#pragma unroll
    for (int it = 0; it < W::NQIT; ++it) {
        const int q = it * 32 + lane;
        if (q < NQ2) { a.Quu[qbase + q] = Quu[it]; a.Quv[qbase + q] = Quv[it]; a.Qvv[qbase + q] = Qvv[it]; a.Hbcl[qbase + q] = H[it]; }
    }
#pragma unroll
    for (int it = 0; it < W::NFIT; ++it) {
        const int p = it * 32 + lane;
        if (p < 4 * Q) {
            const int s = p / Q, iq = p - s * Q, slot = e * 4 + s, nb = a.M.nbr[slot];
            if ((nb < 0) || (e < nb)) {
                const size_t fo = (size_t)slot * Q + iq;
                a.Quu_e[fo] = eu[it]; a.Quv_e[fo] = euv[it]; a.Qvv_e[fo] = ev[it]; a.Hbcl_e[fo] = eH[it];
            }
        }
    }
    if (a.has_visc && lane < NP) {
#pragma unroll
        for (int v = 0; v < 4; ++v) a.btp_dpp_graduv[(size_t)v * a.nstride + nbase + lane] = bsum[v];
        a.pbprime_visc[nbase + lane] = pvs;
    }
}
template <int G, int Q>
constexpr size_t lw_coeffs_smem() { return (size_t)LW_WARPS * (3 * G * G + 4 * LW<G, Q>::ST + 3 * LW<G, Q>::SX + 24 * G + 24 * Q + 4 * G * G) * sizeof(double); }


// ================================================================================================================================
// layer_mass_rhs + update of q_df(1) (mod_create_rhs_mlswe.F90:53-78,822-877,922-1034; mod_splitting.F90:58-78 / 217-232)
template <int G, int Q>
__host__ __device__ constexpr int lw_mass_doubles() { return 3 * G * G + 2 * LW<G, Q>::ST + 3 * LW<G, Q>::SX + 24 * G + 24 * Q + G * G + 4 * Q + 4 * G; }
template <int G, int Q>
constexpr size_t lw_mass_smem() { return (size_t)LW_WARPS * lw_mass_doubles<G, Q>() * sizeof(double); }
template <int G, int Q, int NL_>
__global__ void __launch_bounds__(32 * LW_WARPS, LW_MINB_A) k_layer_mass_w(MassArgs a) {
    using W = LW<G, Q>;
    constexpr int NP = W::NP, NQ2 = W::NQ2, SX = W::SX, ST = W::ST;
    extern __shared__ double sm_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = blockIdx.x * LW_WARPS + warp;
    if (e >= a.M.nelem) return;
    const int nl = NL_ ? NL_ : a.M.nl;
    double* sm = sm_all + (size_t)warp * lw_mass_doubles<G, Q>();
    double* nod = sm;                 // [3][NP]
    double* T = nod + 3 * NP;         // [2][ST]  (the interpolation of 3 fields uses X[2] as its third scratch row: see below)
    double* X = T + 2 * ST;           // [3][SX]
    double* own = X + 3 * SX;         // [4][3][G]
    double* nbq = own + 12 * G;       // [4][3][G]
    double* fq = nbq + 12 * G;        // [24][Q]
    double* adv = fq + 24 * Q;        // [NP]
    double* ff = adv + NP;            // [4][Q]
    double* fp = ff + 4 * Q;          // [4][G]
    const size_t nbase = (size_t)e * NP, qbase = (size_t)e * NQ2;
    const double ksx = a.M.em[e * 5 + 0], ksy = a.M.em[e * 5 + 1], etx = a.M.em[e * 5 + 2], ety = a.M.em[e * 5 + 3], J = a.M.em[e * 5 + 4];
    double sx[W::NQIT], sy[W::NQIT], ope_a[W::NQIT], ub_a[W::NQIT], vb_a[W::NQIT];
#pragma unroll
    for (int it = 0; it < W::NQIT; ++it) {
        const int q = min(it * 32 + lane, NQ2 - 1);
        sx[it] = 0; sy[it] = 0;
        ope_a[it] = a.ave_q[0][qbase + q]; ub_a[it] = a.ave_q[8][qbase + q]; vb_a[it] = a.ave_q[9][qbase + q];
    }
    double sfx[W::NFIT], sfy[W::NFIT], ul_a[W::NFIT], ur_a[W::NFIT], vl_a[W::NFIT], vr_a[W::NFIT], ol_a[W::NFIT], or_a[W::NFIT];
#pragma unroll
    for (int it = 0; it < W::NFIT; ++it) {
        const int p = min(it * 32 + lane, 4 * Q - 1), s = p / Q, iq = p - s * Q, slot = e * 4 + s, nb = a.M.nbr[slot], nbs = a.M.nbslot[slot];
        const bool left = (nb < 0) || (e < nb);
        const size_t fo = (size_t)(left ? slot : nb * 4 + nbs) * Q + iq;
        sfx[it] = 0; sfy[it] = 0;
        ul_a[it] = a.ave_f[12][fo]; ur_a[it] = a.ave_f[13][fo]; vl_a[it] = a.ave_f[14][fo]; vr_a[it] = a.ave_f[15][fo];
        ol_a[it] = a.ave_f[7][fo]; or_a[it] = a.ave_f[8][fo];
    }
    const double mi = lane < NP ? a.massinv[nbase + lane] : 0.0;
#pragma unroll 1
    for (int k = 0; k < nl; ++k) {
        __syncwarp();
        if (lane < NP) {
#pragma unroll
            for (int v = 0; v < 3; ++v) nod[v * NP + lane] = a.qprime[(size_t)(v * nl + k) * a.nstride + nbase + lane];
        }
        __syncwarp();
        lw_qprime_traces<G>(a.M, a.qprime, a.nstride, a.hq, a.hstride, nl, k, e, nod, own, nbq, lane);
        lw_interp<G, Q, 2>(nod, T, X, lane);
        lw_interp<G, Q, 1>(nod + 2 * NP, T, X + 2 * SX, lane);
#pragma unroll
        for (int it = 0; it < W::NQIT; ++it) {
            const int q = min(it * 32 + lane, NQ2 - 1), j = q / Q, i = q - j * Q;
            const double q0 = X[q], q1 = X[SX + q], q2 = X[2 * SX + q];
            const double dp_temp = q0 * ope_a[it];
            const double udp = (q1 + ub_a[it]) * dp_temp;
            const double vdp = (q2 + vb_a[it]) * dp_temp;
            sx[it] = sx[it] + udp; sy[it] = sy[it] + vdp;
            const double wq = c_ops.wq[i] * c_ops.wq[j] * J;
            if (it * 32 + lane < NQ2) { X[q] = wq * (ksx * udp + ksy * vdp); X[SX + q] = wq * (etx * udp + ety * vdp); }
        }
        lw_face_interp<G, Q>(own, fq, 24, lane);
#pragma unroll
        for (int it = 0; it < W::NFIT; ++it) {
            const int p = min(it * 32 + lane, 4 * Q - 1), s = p / Q, iq = p - s * Q, slot = e * 4 + s, nb = a.M.nbr[slot];
            const bool left = (nb < 0) || (e < nb);
            const double nxl = a.M.fgeom[slot * 3], nyl = a.M.fgeom[slot * 3 + 1], nlen = a.M.fgeom[slot * 3 + 2];
            const double* o = fq + (s * 3) * Q + iq;
            const double* r = fq + (12 + s * 3) * Q + iq;
            const double ql0 = left ? o[0] : r[0], ql1 = left ? o[Q] : r[Q], ql2 = left ? o[2 * Q] : r[2 * Q];
            const double qr0 = left ? r[0] : o[0], qr1 = left ? r[Q] : o[Q], qr2 = left ? r[2 * Q] : o[2 * Q];
            const double uu = 0.5 * ((ql1 + ul_a[it]) + (qr1 + ur_a[it]));
            const double vv = 0.5 * ((ql2 + vl_a[it]) + (qr2 + vr_a[it]));
            const double dpl = ol_a[it] * ql0, dpr = or_a[it] * qr0;
            const double fu = (uu * nxl > 0.0) ? uu * dpl : uu * dpr;
            const double fv = (vv * nyl > 0.0) ? vv * dpl : vv * dpr;
            if (left) { sfx[it] = sfx[it] + fu; sfy[it] = sfy[it] + fv; }
            const double flux = nxl * fu + nyl * fv;
            if (it * 32 + lane < 4 * Q) ff[p] = (left ? -1.0 : 1.0) * (c_ops.wq[iq] * nlen) * flux;
        }
        __syncwarp();
        lw_scatter<G, Q, 1, false>(X, X + SX, nullptr, T, adv, lane);
        lw_face_project<G, Q>(ff, fp, 4, lane);
        if (lane < NP) {
            const int m = lane / G, n = lane - m * G;
            double r = adv[lane];
            if (m == 0) r += fp[0 * G + n];
            if (m == G - 1) r += fp[1 * G + n];
            if (n == 0) r += fp[2 * G + m];
            if (n == G - 1) r += fp[3 * G + m];
            const double dpa = mi * r;
            const size_t Iqd = (size_t)k * a.nstride + nbase + lane;
            const double v = a.qdp_in[Iqd] + a.dt * dpa;
            a.qdp[Iqd] = v;
            if (v < 0.0) *a.flag = 1;
        }
    }
#pragma unroll
    for (int it = 0; it < W::NQIT; ++it) {
        const int q = it * 32 + lane;
        if (q < NQ2) { a.slmf_q[0][qbase + q] = sx[it]; a.slmf_q[1][qbase + q] = sy[it]; }
    }
#pragma unroll
    for (int it = 0; it < W::NFIT; ++it) {
        const int p = it * 32 + lane;
        if (p < 4 * Q) {
            const int s = p / Q, iq = p - s * Q, slot = e * 4 + s, nb = a.M.nbr[slot];
            if ((nb < 0) || (e < nb)) { a.slmf_f[0][(size_t)slot * Q + iq] = sfx[it]; a.slmf_f[1][(size_t)slot * Q + iq] = sfy[it]; }
        }
    }
}

// ================================================================================================================================
// apply_consistency (mod_splitting.F90:324-366, mod_layer_terms.F90:57-137, mod_create_rhs_mlswe.F90:80-101,879-920,1036-1115)
template <int G, int Q>
__host__ __device__ constexpr int lw_cons_doubles() { return G * G + 2 * LW<G, Q>::ST + 2 * LW<G, Q>::SX + 8 * G + 8 * Q + G * G + 4 * Q + 4 * G; }
template <int G, int Q>
constexpr size_t lw_cons_smem() { return (size_t)LW_WARPS * lw_cons_doubles<G, Q>() * sizeof(double); }
template <int G, int Q, int NL_>
__global__ void __launch_bounds__(32 * LW_WARPS, LW_MINB_A) k_consistency_w(ConsArgs a) {
    using W = LW<G, Q>;
    constexpr int NP = W::NP, NQ2 = W::NQ2, SX = W::SX, ST = W::ST;
    extern __shared__ double sm_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = blockIdx.x * LW_WARPS + warp;
    if (e >= a.M.nelem) return;
    const int nl = NL_ ? NL_ : a.M.nl;
    double* sm = sm_all + (size_t)warp * lw_cons_doubles<G, Q>();
    double* nod = sm;                 // [NP] dpprime_df of the layer
    double* T = nod + NP;             // [2][ST]
    double* X = T + 2 * ST;           // [2][SX]
    double* own = X + 2 * SX;         // [4][G]
    double* nbd = own + 4 * G;        // [4][G]
    double* fq = nbd + 4 * G;         // [8][Q]
    double* adv = fq + 8 * Q;         // [NP]
    double* ff = adv + NP;            // [4][Q]
    double* fp = ff + 4 * Q;          // [4][G]
    const size_t nbase = (size_t)e * NP, qbase = (size_t)e * NQ2;
    const double ksx = a.M.em[e * 5 + 0], ksy = a.M.em[e * 5 + 1], etx = a.M.em[e * 5 + 2], ety = a.M.em[e * 5 + 3], J = a.M.em[e * 5 + 4];
    // dpprime_df = q_df(1) / (sum_k q_df(1) / pbprime_df)      (mod_splitting.F90:350-353): the divisor of this node and of the
    // neighbour's node across each face
    double ope_own = 1.0, mi = 0.0;
    if (lane < NP) {
        double s = 0.0;
        for (int k = 0; k < nl; ++k) s += a.qdp_in[(size_t)k * a.nstride + nbase + lane];
        ope_own = s / a.pbprime_df[nbase + lane];
        mi = a.massinv[nbase + lane];
    }
    double ope_nb = 1.0;
    int f_nb = 0, f_nbs = 0, f_s = 0, f_n = 0;
    if (lane < 4 * G) {
        f_s = lane / G; f_n = lane - f_s * G;
        const int slot = e * 4 + f_s;
        f_nb = a.M.nbr[slot]; f_nbs = a.M.nbslot[slot];
        if (f_nb >= 0 || f_nb == NBR_HALO) {
            double sum = 0.0;
            for (int k = 0; k < nl; ++k) sum += nb_nodal(a.M, a.qdp_in + (size_t)k * a.nstride, a.hdp + (size_t)k * a.hstride, f_nb, f_nbs, f_n, 0.0);
            ope_nb = sum / a.pbn[(size_t)slot * G + f_n];
        }
    }
    double pbq[W::NQIT], d0q[W::NQIT], d1q[W::NQIT];
#pragma unroll
    for (int it = 0; it < W::NQIT; ++it) {
        const int q = min(it * 32 + lane, NQ2 - 1);
        pbq[it] = a.pbprime_q[qbase + q];
        d0q[it] = a.ave_q[6][qbase + q] - a.slmf_q[0][qbase + q];
        d1q[it] = a.ave_q[7][qbase + q] - a.slmf_q[1][qbase + q];
    }
    double pfl[W::NFIT], pfr[W::NFIT], d0f[W::NFIT], d1f[W::NFIT];
#pragma unroll
    for (int it = 0; it < W::NFIT; ++it) {
        const int p = min(it * 32 + lane, 4 * Q - 1), s = p / Q, iq = p - s * Q, slot = e * 4 + s, nb = a.M.nbr[slot], nbs = a.M.nbslot[slot];
        const bool left = (nb < 0) || (e < nb);
        const size_t fo = (size_t)(left ? slot : nb * 4 + nbs) * Q + iq;
        pfl[it] = a.pbf_l[fo]; pfr[it] = a.pbf_r[fo];
        d0f[it] = a.ave_f[0][fo] - a.slmf_f[0][fo]; d1f[it] = a.ave_f[1][fo] - a.slmf_f[1][fo];
    }
#pragma unroll 1
    for (int k = 0; k < nl; ++k) {
        __syncwarp();
        double qin = 0.0;
        if (lane < NP) { qin = a.qdp_in[(size_t)k * a.nstride + nbase + lane]; nod[lane] = qin / ope_own; }
        __syncwarp();
        if (lane < 4 * G) {
            const double ow = nod[lw_face_node<G>(f_s, f_n)];
            double nv = ow;
            if (f_nb >= 0 || f_nb == NBR_HALO) nv = nb_nodal(a.M, a.qdp_in + (size_t)k * a.nstride, a.hdp + (size_t)k * a.hstride, f_nb, f_nbs, f_n, 0.0) / ope_nb;
            own[f_s * G + f_n] = ow; nbd[f_s * G + f_n] = nv;
        }
        lw_interp<G, Q, 1>(nod, T, X, lane);
#pragma unroll
        for (int it = 0; it < W::NQIT; ++it) {
            const int q = min(it * 32 + lane, NQ2 - 1), j = q / Q, i = q - j * Q;
            const double dp = X[q];
            const double weight = dp / pbq[it];
            const double udp = weight * d0q[it];
            const double vdp = weight * d1q[it];
            const double wq = c_ops.wq[i] * c_ops.wq[j] * J;
            if (it * 32 + lane < NQ2) { X[q] = wq * (ksx * udp + ksy * vdp); X[SX + q] = wq * (etx * udp + ety * vdp); }
        }
        lw_face_interp<G, Q>(own, fq, 8, lane);
#pragma unroll
        for (int it = 0; it < W::NFIT; ++it) {
            const int p = min(it * 32 + lane, 4 * Q - 1), s = p / Q, iq = p - s * Q, slot = e * 4 + s, nb = a.M.nbr[slot];
            const bool left = (nb < 0) || (e < nb);
            const double nxl = a.M.fgeom[slot * 3], nyl = a.M.fgeom[slot * 3 + 1], nlen = a.M.fgeom[slot * 3 + 2];
            const double qo = fq[s * Q + iq], qn = fq[(4 + s) * Q + iq];
            const double qprime_l = left ? qo : qn, qprime_r = left ? qn : qo;
            const double wl = qprime_l / pfl[it], wr = qprime_r / pfr[it];
            const double fu = ((wl * d0f[it]) * nxl > 0.0) ? wl * d0f[it] : wr * d0f[it];
            const double fv = ((wl * d1f[it]) * nyl > 0.0) ? wl * d1f[it] : wr * d1f[it];
            const double flux = nxl * fu + nyl * fv;
            if (it * 32 + lane < 4 * Q) ff[p] = (left ? -1.0 : 1.0) * (c_ops.wq[iq] * nlen) * flux;
        }
        __syncwarp();
        lw_scatter<G, Q, 1, false>(X, X + SX, nullptr, T, adv, lane);
        lw_face_project<G, Q>(ff, fp, 4, lane);
        if (lane < NP) {
            const int m = lane / G, n = lane - m * G;
            double r = adv[lane];
            if (m == 0) r += fp[0 * G + n];
            if (m == G - 1) r += fp[1 * G + n];
            if (n == 0) r += fp[2 * G + m];
            if (n == G - 1) r += fp[3 * G + m];
            a.qdp_out[(size_t)k * a.nstride + nbase + lane] = qin + a.dt * mi * r;
        }
    }
}

// ================================================================================================================================
// bcl_create_laplacian (mod_laplacian_quad.F90:227-248,392-425,521-611) -> rhs_visc planes [2*nl]
template <int G, int Q>
__host__ __device__ constexpr int lw_lap_doubles() { return 8 * G * G + 8 * G; }
template <int G, int Q>
constexpr size_t lw_lap_smem() { return (size_t)LW_WARPS * lw_lap_doubles<G, Q>() * sizeof(double); }
template <int G, int Q, int NL_>
__global__ void __launch_bounds__(32 * LW_WARPS, LW_MINB_L) k_bcl_laplacian_w(LapArgs a) {
    constexpr int NP = G * G;
    extern __shared__ double sm_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = blockIdx.x * LW_WARPS + warp;
    if (e >= a.M.nelem) return;
    const int nl = NL_ ? NL_ : a.M.nl;
    double* sm = sm_all + (size_t)warp * lw_lap_doubles<G, Q>();
    double* gub = sm;                 // [4][NP]
    double* qq = gub + 4 * NP;        // [4][NP]
    double* lf = qq + 4 * NP;         // [4][2][G]
    const size_t nbase = (size_t)e * NP;
    const double ksx = a.M.em[e * 5 + 0], ksy = a.M.em[e * 5 + 1], etx = a.M.em[e * 5 + 2], ety = a.M.em[e * 5 + 3], J = a.M.em[e * 5 + 4];
    const int m = lane < NP ? lane / G : 0, n = lane < NP ? lane - m * G : 0;
    // operator entries of this node (lane-dependent indices: read once)
    double wk1[G], wk2[G];
#pragma unroll
    for (int kk = 0; kk < G; ++kk) {
        wk1[kk] = c_ops.wg[kk] * c_ops.wg[m] * J * c_ops.D[n + G * kk];
        wk2[kk] = c_ops.wg[n] * c_ops.wg[kk] * J * c_ops.D[m + G * kk];
    }
    double mi = 0.0;
    if (lane < NP) {
#pragma unroll
        for (int v = 0; v < 4; ++v) gub[v * NP + lane] = a.graduvb[v][nbase + lane];
        mi = a.massinv[nbase + lane];
    }
    // face node of this lane: geometry, neighbour and the neighbour's averaged barotropic gradient (layer independent)
    const int fs = lane < 4 * G ? lane / G : 0, fn = lane < 4 * G ? lane - fs * G : 0;
    const int slot = e * 4 + fs, nb = a.M.nbr[slot], nbs = a.M.nbslot[slot];
    const bool left = (nb < 0) || (e < nb);
    const double nx = a.M.fgeom[slot * 3], ny = a.M.fgeom[slot * 3 + 1], nlen = a.M.fgeom[slot * 3 + 2];
    const int If = lw_face_node<G>(fs, fn);
    double go[4] = {0, 0, 0, 0}, gn[4] = {0, 0, 0, 0};
    if (lane < 4 * G) {
#pragma unroll
        for (int v = 0; v < 4; ++v) go[v] = a.graduvb[v][nbase + If];
        if (nb >= 0 || nb == NBR_HALO) {
#pragma unroll
            for (int v = 0; v < 4; ++v) gn[v] = nb_nodal(a.M, a.graduvb[v], a.h_gub + (size_t)v * a.hstride, nb, nbs, fn, 0.0);
        } else {
#pragma unroll
            for (int v = 0; v < 4; ++v) gn[v] = go[v];
            if (nb == NBR_FREESLIP) {
                double un = go[0] * nx + go[1] * ny; gn[0] = go[0] - 2.0 * un * nx; gn[1] = go[1] - 2.0 * un * ny;
                un = go[2] * nx + go[3] * ny; gn[2] = go[2] - 2.0 * un * nx; gn[3] = go[3] - 2.0 * un * ny;
            }
        }
    }
#pragma unroll 1
    for (int k = 0; k < nl; ++k) {
        __syncwarp();
        if (lane < NP) {
            const double d = a.dpv[(size_t)k * a.nstride + nbase + lane];
#pragma unroll
            for (int v = 0; v < 4; ++v) qq[v * NP + lane] = d * gub[v * NP + lane] + a.dpp_graduv[(size_t)(v * nl + k) * a.nstride + nbase + lane];
        }
        // face flux at the face nodes (mod_laplacian_quad.F90:521-611)
        if (lane < 4 * G) {
            double so[5], sn[5];
#pragma unroll
            for (int v = 0; v < 4; ++v) so[v] = a.dpp_graduv[(size_t)(v * nl + k) * a.nstride + nbase + If];
            so[4] = a.dpv[(size_t)k * a.nstride + nbase + If];
            if (nb >= 0 || nb == NBR_HALO) {
#pragma unroll
                for (int v = 0; v < 4; ++v)
                    sn[v] = nb_nodal(a.M, a.dpp_graduv + (size_t)(v * nl + k) * a.nstride, a.h_dpg + (size_t)(v * nl + k) * a.hstride, nb, nbs, fn, 0.0);
                sn[4] = nb_nodal(a.M, a.dpv + (size_t)k * a.nstride, a.h_dpv + (size_t)k * a.hstride, nb, nbs, fn, 0.0);
            } else {
#pragma unroll
                for (int v = 0; v < 5; ++v) sn[v] = so[v];
                if (nb == NBR_FREESLIP) {
                    double un = so[0] * nx + so[1] * ny; sn[0] = so[0] - 2.0 * un * nx; sn[1] = so[1] - 2.0 * un * ny;
                    un = so[2] * nx + so[3] * ny; sn[2] = so[2] - 2.0 * un * nx; sn[3] = so[3] - 2.0 * un * ny;
                }
            }
            double fl[4], fr[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const double fo_ = so[4] * go[v] + so[v], fn_ = sn[4] * gn[v] + sn[v];
                fl[v] = left ? fo_ : fn_; fr[v] = left ? fn_ : fo_;
            }
            const double qu0 = 0.5 * fl[0] + 0.5 * fr[0], qu1 = 0.5 * fl[1] + 0.5 * fr[1];
            const double qv0 = 0.5 * fl[2] + 0.5 * fr[2], qv1 = 0.5 * fl[3] + 0.5 * fr[3];
            const double wq = c_ops.wg[fn] * nlen;
            const double flux_qu = (qu0 - fl[0] * nx) + (qu1 - fl[1] * ny);
            const double flux_qv = (qv0 - fl[2] * nx) + (qv1 - fl[3] * ny);
            const double sgn = left ? 1.0 : -1.0;
            lf[(fs * 2 + 0) * G + fn] = sgn * wq * flux_qu;
            lf[(fs * 2 + 1) * G + fn] = sgn * wq * flux_qv;
        }
        __syncwarp();
        if (lane < NP) {
            double l0 = 0.0, l1 = 0.0;
#pragma unroll
            for (int kk = 0; kk < G; ++kk) {
                const int I1 = m * G + kk, I2 = kk * G + n;
                l0 -= wk1[kk] * (ksx * qq[I1] + ksy * qq[NP + I1]) + wk2[kk] * (etx * qq[I2] + ety * qq[NP + I2]);
                l1 -= wk1[kk] * (ksx * qq[2 * NP + I1] + ksy * qq[3 * NP + I1]) + wk2[kk] * (etx * qq[2 * NP + I2] + ety * qq[3 * NP + I2]);
            }
            if (m == 0) { l0 += lf[(0 * 2 + 0) * G + n]; l1 += lf[(0 * 2 + 1) * G + n]; }
            if (m == G - 1) { l0 += lf[(1 * 2 + 0) * G + n]; l1 += lf[(1 * 2 + 1) * G + n]; }
            if (n == 0) { l0 += lf[(2 * 2 + 0) * G + m]; l1 += lf[(2 * 2 + 1) * G + m]; }
            if (n == G - 1) { l0 += lf[(3 * 2 + 0) * G + m]; l1 += lf[(3 * 2 + 1) * G + m]; }
            a.rhs_visc[(size_t)(0 * nl + k) * a.nstride + nbase + lane] = a.visc * mi * l0;
            a.rhs_visc[(size_t)(1 * nl + k) * a.nstride + nbase + lane] = a.visc * mi * l1;
        }
    }
}


// ================================================================================================================================
// create_rhs_dynamics_volume_layers (mod_create_rhs_mlswe.F90:281-456) -> rhs_mom planes [2*nl] (not yet multiplied by massinv).
// The reference (and the block-per-element kernel) keeps a dozen per-layer values of every quadrature point across two layer
// loops.  Here the first loop only accumulates the six sums over the layers; the second loop interpolates the layer again
// and recomputes its values -- the same expressions on the same inputs, hence the same numbers -- so that the per-point state
// that survives a layer is independent of the number of layers.
template <int G, int Q>
__host__ __device__ constexpr int lw_mvol_doubles(int nl) { return (nl + 1) * G * G + 5 * G * G + 5 * LW<G, Q>::ST + 6 * LW<G, Q>::SX + 2 * G * G; }
template <int G, int Q, int NL_>
__global__ void __launch_bounds__(32 * LW_WARPS, 3) k_mom_volume_w(MomVolArgs a) {
    using W = LW<G, Q>;
    constexpr int NP = W::NP, NQ2 = W::NQ2, SX = W::SX, ST = W::ST, NQIT = W::NQIT;
    extern __shared__ double sm_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = blockIdx.x * LW_WARPS + warp;
    if (e >= a.M.nelem) return;
    const int nl = NL_ ? NL_ : a.M.nl;
    double* sm = sm_all + (size_t)warp * lw_mvol_doubles<G, Q>(nl);
    double* zl = sm;                      // [nl+1][NP] interface elevations z_elv(:,k)
    double* nod = zl + (nl + 1) * NP;     // [5][NP]: dp', u', v', udp, vdp of the current layer
    double* T = nod + 5 * NP;             // [5][ST]
    double* X = T + 5 * ST;               // [6][SX]
    double* out = X + 6 * SX;             // [2][NP]
    const size_t nbase = (size_t)e * NP, qbase = (size_t)e * NQ2;
    const double ksx = a.M.em[e * 5 + 0], ksy = a.M.em[e * 5 + 1], etx = a.M.em[e * 5 + 2], ety = a.M.em[e * 5 + 3], J = a.M.em[e * 5 + 4];
    const double eps1 = 1.0e-20;
    const double Pstress = (a.g / a.alpha[0]) * 50.0, Pbstress = (a.g / a.alpha[nl - 1]) * 10.0;
    // interface elevations at the nodes, bottom up (mod_create_rhs_mlswe.F90:320-325)
    if (lane < NP) {
        double zcur = a.zbot_df[nbase + lane];
        zl[nl * NP + lane] = zcur;
        const double sq = sqrt(a.ope2_df[nbase + lane]);
        for (int k = nl - 1; k >= 0; --k) {
            zcur = zcur + (a.alpha[k] / a.g) * (sq * a.qprime[(size_t)(0 * nl + k) * a.nstride + nbase + lane]);
            zl[k * NP + lane] = zcur;
        }
    }
    __syncwarp();
    // ---- first layer loop: sums over the layers
    double s_uu[NQIT], s_uv[NQIT], s_vv[NQIT], s_tu[NQIT], s_tv[NQIT], s_H[NQIT], qpl0[NQIT], qpl1[NQIT], qpl2[NQIT];
    {
        double pcur[NQIT];
#pragma unroll
        for (int it = 0; it < NQIT; ++it) { s_uu[it] = 0; s_uv[it] = 0; s_vv[it] = 0; s_tu[it] = 0; s_tv[it] = 0; s_H[it] = 0; pcur[it] = 0; qpl0[it] = 0; qpl1[it] = 0; qpl2[it] = 0; }
#pragma unroll 1
        for (int k = 0; k < nl; ++k) {
            __syncwarp();
            if (lane < NP) {
#pragma unroll
                for (int v = 0; v < 3; ++v) nod[v * NP + lane] = a.qprime[(size_t)(v * nl + k) * a.nstride + nbase + lane];
                nod[3 * NP + lane] = a.q[(size_t)(1 * nl + k) * a.nstride + nbase + lane];
                nod[4 * NP + lane] = a.q[(size_t)(2 * nl + k) * a.nstride + nbase + lane];
            }
            __syncwarp();
            lw_interp<G, Q, 5>(nod, T, X, lane);
#pragma unroll
            for (int it = 0; it < NQIT; ++it) {
                const int q = min(it * 32 + lane, NQ2 - 1);
                const size_t Iq = qbase + q;
                const double q0 = X[q], q1 = X[SX + q], q2 = X[2 * SX + q], tu = X[3 * SX + q], tv = X[4 * SX + q];
                const double sq_ope2 = sqrt(a.ave_q[5][Iq]), ope_a = a.ave_q[0][Iq], ub_a = a.ave_q[8][Iq], vb_a = a.ave_q[9][Iq];
                const double pn = pcur[it] + sq_ope2 * q0;
                s_H[it] += 0.5 * a.alpha[k] * (pn * pn - pcur[it] * pcur[it]);
                pcur[it] = pn;
                const double dp = q0 * ope_a, u = q1 + ub_a, v = q2 + vb_a;
                s_uu[it] += dp * u * u; s_vv[it] += dp * v * v; s_uv[it] += u * v * dp;
                s_tu[it] += fabs(tu) + eps1; s_tv[it] += fabs(tv) + eps1;
                if (k == nl - 1) { qpl0[it] = q0; qpl1[it] = q1; qpl2[it] = q2; }
            }
        }
    }
    // deficits and weights of the point (mod_create_rhs_mlswe.F90:396-414)
    double uu_def[NQIT], uv_def[NQIT], vv_def[NQIT], oosu[NQIT], oosv[NQIT], wH[NQIT];
#pragma unroll
    for (int it = 0; it < NQIT; ++it) {
        const size_t Iq = qbase + min(it * 32 + lane, NQ2 - 1);
        uu_def[it] = a.ave_q[2][Iq] - s_uu[it]; uv_def[it] = a.ave_q[4][Iq] - s_uv[it]; vv_def[it] = a.ave_q[3][Iq] - s_vv[it];
        oosu[it] = 1.0 / s_tu[it]; oosv[it] = 1.0 / s_tv[it];
        wH[it] = 1.0;
        if (s_H[it] > 0.0) wH[it] = a.ave_q[1][Iq] / s_H[it];
    }
    // ---- second layer loop: fluxes and sources of every layer, scattered to the nodes
    double gz1[NQIT], gz2[NQIT], pcur[NQIT], ppt[NQIT];
    {   // grad z of the top interface
        __syncwarp();
        lw_grad_q<G, Q, 1>(zl, T, X, X + SX, lane);
#pragma unroll
        for (int it = 0; it < NQIT; ++it) {
            const int q = min(it * 32 + lane, NQ2 - 1);
            const double dks = X[q], det = X[SX + q];
            gz1[it] = ksx * dks + etx * det; gz2[it] = ksy * dks + ety * det;
            pcur[it] = 0.0; ppt[it] = 0.0;
        }
    }
#pragma unroll 1
    for (int k = 0; k < nl; ++k) {
        __syncwarp();
        if (lane < NP) {
#pragma unroll
            for (int v = 0; v < 3; ++v) nod[v * NP + lane] = a.qprime[(size_t)(v * nl + k) * a.nstride + nbase + lane];
            nod[3 * NP + lane] = a.q[(size_t)(1 * nl + k) * a.nstride + nbase + lane];
            nod[4 * NP + lane] = a.q[(size_t)(2 * nl + k) * a.nstride + nbase + lane];
        }
        __syncwarp();
        lw_interp<G, Q, 5>(nod, T, X, lane);
        double q0[NQIT], q1[NQIT], q2[NQIT], tu[NQIT], tv[NQIT];
#pragma unroll
        for (int it = 0; it < NQIT; ++it) {
            const int q = min(it * 32 + lane, NQ2 - 1);
            q0[it] = X[q]; q1[it] = X[SX + q]; q2[it] = X[2 * SX + q]; tu[it] = X[3 * SX + q]; tv[it] = X[4 * SX + q];
        }
        __syncwarp();
        lw_grad_q<G, Q, 1>(zl + (k + 1) * NP, T, X, X + SX, lane);   // grad z of the lower interface of the layer
        double gn1[NQIT], gn2[NQIT];
#pragma unroll
        for (int it = 0; it < NQIT; ++it) {
            const int q = min(it * 32 + lane, NQ2 - 1);
            const double dks = X[q], det = X[SX + q];
            gn1[it] = ksx * dks + etx * det; gn2[it] = ksy * dks + ety * det;
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < NQIT; ++it) {
            const int q = min(it * 32 + lane, NQ2 - 1), j = q / Q, i = q - j * Q;
            const size_t Iq = qbase + q;
            const double sq_ope2 = sqrt(a.ave_q[5][Iq]), ope_a = a.ave_q[0][Iq], ub_a = a.ave_q[8][Iq], vb_a = a.ave_q[9][Iq];
            const double p_k = pcur[it], p_k1 = pcur[it] + sq_ope2 * q0[it];
            const double H_tmp = 0.5 * a.alpha[k] * (p_k1 * p_k1 - p_k * p_k);
            const double dp = q0[it] * ope_a, u = q1[it] + ub_a, v = q2[it] + vb_a;
            const double u_udp = dp * u * u, v_vdp = dp * v * v, u_vdp1 = u * v * dp, u_vdp2 = v * u * dp;
            const double temp_uu = fabs(tu[it]) + eps1, temp_vv = fabs(tv[it]) + eps1;
            // hazard 1 (mod_create_rhs_mlswe.F90:382): as written for nl<=3 (qp of the LAST layer indexed by the layer number),
            // intent (dp'_k) for nl>3
            const double inc = (nl <= 3) ? (k == 0 ? qpl0[it] : k == 1 ? qpl1[it] : qpl2[it]) : q0[it];
            const double ppn = ppt[it] + inc;
            double wgt = temp_uu * oosu[it];
            const double var_uu = u_udp + wgt * uu_def[it];
            const double var_uv = u_vdp1 + wgt * uv_def[it];
            wgt = temp_vv * oosv[it];
            const double var_vu = u_vdp2 + wgt * uv_def[it];
            const double var_vv = v_vdp + wgt * vv_def[it];
            const double Hq = H_tmp * wH[it];
            const double pbq = a.pbprime_q[Iq], twx = a.tauwx_q[Iq], twy = a.tauwy_q[Iq], tbx = a.ave_q[10][Iq], tby = a.ave_q[11][Iq];
            const double temp1 = (fmin(ppn, Pstress) - fmin(ppt[it], Pstress)) / Pstress;
            double tempbot = fmin(Pbstress, pbq - ppn) - fmin(Pbstress, pbq - ppt[it]);
            tempbot = tempbot / Pbstress;
            const double source_x = a.g * (temp1 * twx - tempbot * tbx + p_k * gz1[it] - p_k1 * gn1[it]);
            const double source_y = a.g * (temp1 * twy - tempbot * tby + p_k * gz2[it] - p_k1 * gn2[it]);
            ppt[it] = ppn; pcur[it] = p_k1; gz1[it] = gn1[it]; gz2[it] = gn2[it];
            const double wq = c_ops.wq[i] * c_ops.wq[j] * J;
            const double Fx1 = Hq + var_uu, Fy1 = var_uv, Fx2 = var_vu, Fy2 = Hq + var_vv;
            if (it * 32 + lane < NQ2) {
                X[0 * SX + q] = wq * source_x;
                X[1 * SX + q] = wq * source_y;
                X[2 * SX + q] = wq * (ksx * Fx1 + ksy * Fy1);
                X[3 * SX + q] = wq * (ksx * Fx2 + ksy * Fy2);
                X[4 * SX + q] = wq * (etx * Fx1 + ety * Fy1);
                X[5 * SX + q] = wq * (etx * Fx2 + ety * Fy2);
            }
        }
        __syncwarp();
        lw_scatter<G, Q, 2, true>(X + 2 * SX, X + 4 * SX, X, T, out, lane);
        if (lane < NP) {
            a.rhs_mom[(size_t)(0 * nl + k) * a.nstride + nbase + lane] = out[lane];
            a.rhs_mom[(size_t)(1 * nl + k) * a.nstride + nbase + lane] = out[NP + lane];
        }
    }
}


// ================================================================================================================================
// Apply_layers_fluxes + momentum update + Coriolis rotation + wall projection + velocity reconciliation
// (mod_create_rhs_mlswe.F90:458-820, mod_splitting.F90:94-287, mod_layer_terms.F90:198-320,529-584)
template <int G, int Q>
__host__ __device__ constexpr int lw_mface_doubles(int nl) { return 24 * nl * G + 24 * nl * Q + 8 * nl * Q + 8 * nl * G; }
template <int G, int Q, int NL_>
__global__ void __launch_bounds__(32 * LW_WARPS, LW_MINB_F) k_mom_faces_update_w(MomFaceArgs a) {
    using W = LW<G, Q>;
    constexpr int NP = W::NP;
    constexpr int LMAX = NL_ ? NL_ : HN_MAXL;
    extern __shared__ double sm_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = blockIdx.x * LW_WARPS + warp;
    if (e >= a.M.nelem) return;
    const int nl = NL_ ? NL_ : a.M.nl;
    const int tid = lane, ngl = G, nq = Q;
    double* sm = sm_all + (size_t)warp * lw_mface_doubles<G, Q>(nl);
    double* ownq = sm;                          // [4][3][nl][G]
    double* nbq = ownq + 12 * nl * G;           // [4][3][nl][G]
    double* fq = nbq + 12 * nl * G;             // [24 nl][Q]: the same lines at the face quadrature points
    double* ff = fq + 24 * nl * Q;              // [4][nl][2][Q]
    double* fp = ff + 8 * nl * Q;               // [4][nl][2][G]
    const size_t nbase = (size_t)e * NP;
    const double eps1 = 1.0e-20;
#pragma unroll 1
    for (int t = lane; t < 4 * G * nl; t += 32) {
        const int k = t / (4 * G), r = t - k * 4 * G, s = r / G, n = r - s * G;
        const int slot = e * 4 + s, nb = a.M.nbr[slot], nbs = a.M.nbslot[slot];
        const int I = lw_face_node<G>(s, n);
        double ow[3], nv[3];
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            ow[v] = a.qprime[(size_t)(v * nl + k) * a.nstride + nbase + I];
            nv[v] = nb_nodal(a.M, a.qprime + (size_t)(v * nl + k) * a.nstride, a.hq + (size_t)(v * nl + k) * a.hstride, nb, nbs, n, ow[v]);
        }
        if (nb == NBR_FREESLIP) {
            const double nx = a.M.fgeom[slot * 3], ny = a.M.fgeom[slot * 3 + 1];
            const double un = ow[1] * nx + ow[2] * ny;
            nv[1] = ow[1] - 2.0 * un * nx; nv[2] = ow[2] - 2.0 * un * ny;
        } else if (nb == NBR_NOSLIP) { nv[1] = -ow[1]; nv[2] = -ow[2]; }
#pragma unroll
        for (int v = 0; v < 3; ++v) { ownq[((s * 3 + v) * nl + k) * G + n] = ow[v]; nbq[((s * 3 + v) * nl + k) * G + n] = nv[v]; }
    }
    __syncwarp();
    lw_face_interp<G, Q>(ownq, fq, 24 * nl, lane);
#pragma unroll 1
    for (int it = 0; it < W::NFIT; ++it) {
        if (it * 32 + lane >= 4 * Q) break;
        const int p = it * 32 + lane, s = p / Q, iq = p - s * Q, slot = e * 4 + s, nb = a.M.nbr[slot], nbs = a.M.nbslot[slot];
        const bool left = (nb < 0) || (e < nb);
        const int oslot = left ? slot : nb * 4 + nbs;
        const size_t fo = (size_t)oslot * Q + iq;
        const double nxl = a.M.fgeom[slot * 3], nyl = a.M.fgeom[slot * 3 + 1], nlen = a.M.fgeom[slot * 3 + 2];
        const int tl_off = left ? 0 : 12 * nl, tr_off = left ? 12 * nl : 0;
        double ql0[LMAX], qr0[LMAX];
        double udpl[LMAX], udpr[LMAX], vdpl[LMAX], vdpr[LMAX];
        double uf0[LMAX], uf1[LMAX], vf0[LMAX], vf1[LMAX], HL[LMAX], HR[LMAX];
        double qbl0 = a.ave_f[7][fo], qbl1 = a.ave_f[12][fo], qbl2 = a.ave_f[14][fo];
        double qbr0 = a.ave_f[8][fo], qbr1 = a.ave_f[13][fo], qbr2 = a.ave_f[15][fo];
        for (int k = 0; k < nl; ++k) {
            double ql[3], qr[3];
#pragma unroll
            for (int v = 0; v < 3; ++v) { ql[v] = fq[(tl_off + (s * 3 + v) * nl + k) * Q + iq]; qr[v] = fq[(tr_off + (s * 3 + v) * nl + k) * Q + iq]; }
            ql0[k] = ql[0]; qr0[k] = qr[0];
            double dpl = qbl0 * ql[0], dpr = qbr0 * qr[0];
            double ul = ql[1] + qbl1, ur = qr[1] + qbr1, vl = ql[2] + qbl2, vr = qr[2] + qbr2;
            double uu = 0.5 * (ul + ur), vv = 0.5 * (vl + vr);
            udpl[k] = ul * dpl; udpr[k] = ur * dpr; vdpl[k] = vl * dpl; vdpr[k] = vr * dpr;
            if (uu * nxl > 0.0) { uf0[k] = uu * (ul * dpl); vf0[k] = uu * (vl * dpl); }
            else { uf0[k] = uu * (ur * dpr); vf0[k] = uu * (vr * dpr); }
            if (vv * nyl > 0.0) { uf1[k] = vv * (ul * dpl); vf1[k] = vv * (vl * dpl); }
            else { uf1[k] = vv * (ur * dpr); vf1[k] = vv * (vr * dpr); }
        }
        double su0 = 0, su1 = 0, sv0 = 0, sv1 = 0;
        for (int k = 0; k < nl; ++k) { su0 += uf0[k]; su1 += uf1[k]; sv0 += vf0[k]; sv1 += vf1[k]; }
        double uu_def = a.ave_f[3][fo] - su0, uv_def = a.ave_f[4][fo] - su1;
        double vu_def = a.ave_f[5][fo] - sv0, vv_def = a.ave_f[6][fo] - sv1;
        double sl = 0, sr = 0;
        for (int k = 0; k < nl; ++k) { sl += fabs(udpl[k]) + eps1; sr += fabs(udpr[k]) + eps1; }
        double oosl = 1.0 / sl, oosr = 1.0 / sr;
        for (int k = 0; k < nl; ++k) {
            double w = (uu_def * nxl > 0.0) ? fabs(udpl[k]) * oosl : fabs(udpr[k]) * oosr;
            uf0[k] = uf0[k] + w * uu_def;
            w = (uv_def * nyl > 0.0) ? fabs(udpl[k]) * oosl : fabs(udpr[k]) * oosr;
            uf1[k] = uf1[k] + w * uv_def;
        }
        sl = 0; sr = 0;
        for (int k = 0; k < nl; ++k) { sl += fabs(vdpl[k]) + eps1; sr += fabs(vdpr[k]) + eps1; }
        oosl = 1.0 / sl; oosr = 1.0 / sr;
        for (int k = 0; k < nl; ++k) {
            double w = (vu_def * nxl > 0.0) ? fabs(vdpl[k]) * oosl : fabs(vdpr[k]) * oosr;
            vf0[k] = vf0[k] + w * vu_def;
            w = (vv_def * nyl > 0.0) ? fabs(vdpl[k]) * oosl : fabs(vdpr[k]) * oosr;
            vf1[k] = vf1[k] + w * vv_def;
        }
        // pressure forcing H_face (mod_create_rhs_mlswe.F90:627-773)
        double pfl[LMAX + 1], pfr[LMAX + 1], zfl[LMAX + 1], zfr[LMAX + 1];
        double pep[LMAX + 1], pem[LMAX + 1], zep[LMAX + 1], zem[LMAX + 1];
        double ope_l = sqrt(a.ave_f[9][fo]), ope_r = sqrt(a.ave_f[10][fo]);
        pfl[0] = 0.0; pfr[0] = 0.0;
        for (int k = 0; k < nl; ++k) { pfl[k + 1] = pfl[k] + ope_l * ql0[k]; pfr[k + 1] = pfr[k] + ope_r * qr0[k]; }
        double ope_e = sqrt(a.ave_f[11][fo]);
        zfl[nl] = a.zbf_l[fo]; zfr[nl] = a.zbf_r[fo]; zep[nl] = a.zbf_l[fo]; zem[nl] = a.zbf_r[fo];
        for (int k = nl - 1; k >= 0; --k) {
            double aog = a.alpha[k] / a.g;
            zfl[k] = zfl[k + 1] + aog * (ope_l * ql0[k]);
            zfr[k] = zfr[k + 1] + aog * (ope_r * qr0[k]);
            zep[k] = zep[k + 1] + aog * (ope_e * ql0[k]);
            zem[k] = zem[k + 1] + aog * (ope_e * qr0[k]);
        }
        pep[0] = 0.0; pem[0] = 0.0;
        pep[1] = ope_e * ql0[0]; pem[1] = ope_e * qr0[0];
        for (int k = 1; k < nl; ++k) { pep[k + 1] = pep[k] + ope_e * ql0[k]; pem[k + 1] = pem[k] + ope_e * qr0[k]; }
        for (int k = 0; k < nl; ++k) {
            double H_r_plus = 0.5 * a.alpha[k] * (pep[k + 1] * pep[k + 1] - pep[k] * pep[k]);
            double H_r_minus = 0.0;
            for (int kt = 0; kt < nl; ++kt) {
                double zt = fmin(zem[kt], zep[k]), zb = fmax(zem[kt + 1], zep[k + 1]);
                if (zt - zb > 0.0) {
                    double goa = a.g / a.alpha[kt];
                    double pb_ = pem[kt + 1] - goa * (zb - zem[kt + 1]);
                    double pt_ = pem[kt + 1] - goa * (zt - zem[kt + 1]);
                    H_r_minus = H_r_minus + 0.5 * a.alpha[kt] * (pb_ * pb_ - pt_ * pt_);
                }
            }
            HL[k] = 0.5 * (H_r_plus + H_r_minus);
            H_r_minus = 0.5 * a.alpha[k] * (pem[k + 1] * pem[k + 1] - pem[k] * pem[k]);
            H_r_plus = 0.0;
            for (int kt = 0; kt < nl; ++kt) {
                double zt = fmin(zep[kt], zem[k]), zb = fmax(zep[kt + 1], zem[k + 1]);
                if (zt - zb > 0.0) {
                    double goa = a.g / a.alpha[kt];
                    double pb_ = pep[kt + 1] - goa * (zb - zep[kt + 1]);
                    double pt_ = pep[kt + 1] - goa * (zt - zep[kt + 1]);
                    H_r_plus = H_r_plus + 0.5 * a.alpha[kt] * (pb_ * pb_ - pt_ * pt_);
                }
            }
            HR[k] = 0.5 * (H_r_plus + H_r_minus);
        }
        if (nb == NBR_FREESLIP) {
            double p2l = 0.0, p2r = 0.0;
            for (int k = 0; k < nl; ++k) {
                HL[k] = 0.5 * a.alpha[k] * (pfl[k + 1] * pfl[k + 1] - p2l * p2l); p2l = pfl[k + 1];
                HR[k] = 0.5 * a.alpha[k] * (pfr[k + 1] * pfr[k + 1] - p2r * p2r); p2r = pfr[k + 1];
            }
        } else {
            for (int k = 0; k < nl - 1; ++k) {
                double goa = a.g / a.alpha[k];
                double p_inc1 = goa * (zfl[k + 1] - zep[k + 1]);
                double H_corr1 = 0.5 * a.alpha[k] * ((pfl[k + 1] + p_inc1) * (pfl[k + 1] + p_inc1) - pfl[k + 1] * pfl[k + 1]);
                HL[k] = HL[k] - H_corr1; HL[k + 1] = HL[k + 1] + H_corr1;
                double p_inc2 = goa * (zfr[k + 1] - zem[k + 1]);
                double H_corr2 = 0.5 * a.alpha[k] * ((pfr[k + 1] + p_inc2) * (pfr[k + 1] + p_inc2) - pfr[k + 1] * pfr[k + 1]);
                HR[k] = HR[k] - H_corr2; HR[k + 1] = HR[k + 1] + H_corr2;
            }
        }
        double Hfa = a.ave_f[2][fo];
        double accl = 0.0, accr = 0.0;
        for (int k = 0; k < nl; ++k) { accl += HL[k]; accr += HR[k]; }
        double wl = (accl > 0.0) ? Hfa / accl : 1.0, wr = (accr > 0.0) ? Hfa / accr : 1.0;
        double wq = c_ops.wq[iq] * nlen;
        for (int k = 0; k < nl; ++k) {
            double Hs = left ? HL[k] * wl : HR[k] * wr;
            double flux_x = nxl * uf0[k] + nyl * uf1[k];
            double flux_y = nxl * vf0[k] + nyl * vf1[k];
            double sgn = left ? -1.0 : 1.0;
            ff[((s * nl + k) * 2 + 0) * Q + iq] = sgn * wq * (nxl * Hs + flux_x);
            ff[((s * nl + k) * 2 + 1) * Q + iq] = sgn * wq * (nyl * Hs + flux_y);
        }
    }
    __syncwarp();
    lw_face_project<G, Q>(ff, fp, 8 * nl, lane);
    if (lane < NP) {
        const int m = lane / G, n = lane - m * G;
        double qd[LMAX], qx[LMAX], qy[LMAX];
        double mi = a.massinv[nbase + tid];
        double f2 = a.fdt2[nbase + tid], ab = a.a_bcl[nbase + tid], bb = a.b_bcl[nbase + tid];
        for (int k = 0; k < nl; ++k) {
            double r0 = a.rhs_mom[(size_t)(0 * nl + k) * a.nstride + nbase + tid];
            double r1 = a.rhs_mom[(size_t)(1 * nl + k) * a.nstride + nbase + tid];
            if (m == 0) { r0 += fp[((0 * nl + k) * 2 + 0) * G + n]; r1 += fp[((0 * nl + k) * 2 + 1) * G + n]; }
            if (m == G - 1) { r0 += fp[((1 * nl + k) * 2 + 0) * G + n]; r1 += fp[((1 * nl + k) * 2 + 1) * G + n]; }
            if (n == 0) { r0 += fp[((2 * nl + k) * 2 + 0) * G + m]; r1 += fp[((2 * nl + k) * 2 + 1) * G + m]; }
            if (n == G - 1) { r0 += fp[((3 * nl + k) * 2 + 0) * G + m]; r1 += fp[((3 * nl + k) * 2 + 1) * G + m]; }
            r0 = mi * r0 + a.rhs_visc[(size_t)(0 * nl + k) * a.nstride + nbase + tid];
            r1 = mi * r1 + a.rhs_visc[(size_t)(1 * nl + k) * a.nstride + nbase + tid];
            if (a.rhs_only) {   // per-phase entry hnumo_layer_momentum_rhs: rhs_mom(1:2,I,k) of layer_momentum_rhs, nothing else
                a.rhs_out[(size_t)(0 * nl + k) * a.nstride + nbase + tid] = r0;
                a.rhs_out[(size_t)(1 * nl + k) * a.nstride + nbase + tid] = r1;
                qd[k] = 1.0; qx[k] = 0.0; qy[k] = 0.0;
                continue;
            }
            double dpk = a.q[(size_t)(0 * nl + k) * a.nstride + nbase + tid];
            double mxo = a.q_in[(size_t)(1 * nl + k) * a.nstride + nbase + tid], myo = a.q_in[(size_t)(2 * nl + k) * a.nstride + nbase + tid];
            double t1 = mxo + a.dt * r0, t2 = myo + a.dt * r1;
            double tempu = t1 + f2 * myo, tempv = t2 - f2 * mxo;
            double mxn = ab * tempu + bb * tempv, myn = -bb * tempu + ab * tempv;
            // wall projection (layer_mom_boundary_df)
            for (int s = 0; s < 4; ++s) {
                bool on = (s == 0) ? (m == 0) : (s == 1) ? (m == ngl - 1) : (s == 2) ? (n == 0) : (n == ngl - 1);
                if (!on) continue;
                int slot = e * 4 + s, nb = a.M.nbr[slot];
                if (nb == NBR_FREESLIP) {
                    double nx = a.M.fgeom[slot * 3], ny = a.M.fgeom[slot * 3 + 1];
                    double up = mxn * nx + myn * ny;
                    mxn = mxn - up * nx; myn = myn - up * ny;
                } else if (nb == NBR_NOSLIP) { mxn = 0.0; myn = 0.0; }
            }
            qd[k] = dpk; qx[k] = mxn; qy[k] = myn;
        }
        if (a.rhs_only) return;
        // evaluate_bcl / evaluate_bcl_v1: two passes of extract_velocity
        double pb = a.qb[0][nbase + tid] + a.pbprime_df[nbase + tid];
        double mbx = a.qb[1][nbase + tid], mby = a.qb[2][nbase + tid];
        double uk[LMAX], vk[LMAX];
        for (int pass = 0; pass < 2; ++pass) {
            double ubar = 0.0, vbar = 0.0;
            for (int k = 0; k < nl; ++k) { uk[k] = qx[k] / qd[k]; vk[k] = qy[k] / qd[k]; }
            for (int k = 0; k < nl; ++k) { ubar = ubar + uk[k] * qd[k]; vbar = vbar + vk[k] * qd[k]; }
            if (pb > 0.0) {
                ubar = ubar / pb; vbar = vbar / pb;
                for (int k = 0; k < nl; ++k) { uk[k] = uk[k] - ubar + mbx / pb; vk[k] = vk[k] - vbar + mby / pb; }
            } else {
                for (int k = 0; k < nl; ++k) { uk[k] = 0.0; vk[k] = 0.0; }
            }
            if (pass == 0) for (int k = 0; k < nl; ++k) { qx[k] = uk[k] * qd[k]; qy[k] = vk[k] * qd[k]; }
        }
        double ope = 0.0;
        for (int k = 0; k < nl; ++k) ope = ope + qd[k];
        ope = ope / a.pbprime_df[nbase + tid];
        for (int k = 0; k < nl; ++k) {
            a.q[(size_t)(1 * nl + k) * a.nstride + nbase + tid] = qx[k];
            a.q[(size_t)(2 * nl + k) * a.nstride + nbase + tid] = qy[k];
            if (a.full_prime) a.qprime_out[(size_t)(0 * nl + k) * a.nstride + nbase + tid] = qd[k] / ope;
            a.qprime_out[(size_t)(1 * nl + k) * a.nstride + nbase + tid] = uk[k] - mbx / pb;
            a.qprime_out[(size_t)(2 * nl + k) * a.nstride + nbase + tid] = vk[k] - mby / pb;
        }
    }
}

}  // namespace hn
