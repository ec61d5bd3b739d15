// stage_rec.cuh -- barotropic SSPRK stage kernel, record layout + TMA staging, "one OUTPUT per thread" contractions.
//
// Same operator and the same global records as stage_tma.cuh (reference src/mod_rhs_btp.F90:28-370,
// src/mod_barotropic_terms.F90:25-97,165-217, src/mod_laplacian_quad.F90:32-121,357-519, src/mod_rk_mlswe.F90:87-114).
//
// Why a second mapping: on sm_100a an FP64 FMA cannot take a constant-bank operand; every matrix entry costs an
// LDCU (uniform load) next to its DFMA, so the "one line per lane" contraction of stage_fused/stage_tma needs two
// instructions per multiply-add, leaves lanes idle when a phase has fewer than 32 lines, and serialises the element
// on single-warp phases.  Here every thread owns ONE output of a 1-D contraction and keeps its column of the operator
// in registers (the column index, tid % nq, is the same for the interpolation along ksi, along eta and on the faces),
// inputs are read with 128-bit shared-memory loads, and the quadrature-point values never leave registers between
// the second interpolation pass and the pointwise physics.
// Warp roles: warps 0-2 run the volume pipeline (interpolate -> physics -> weak-form scatter), warp 3 runs the face
// pipeline (traces, ghosts, Riemann fluxes, LDG face flux) concurrently; they meet for the nodal update.
#pragma once
#include "stage_tma.cuh"

namespace hn {

template <int G, int Q>
struct RecSmem {
    using RL = RecLayout<G, Q>;
    static constexpr int NP = G * G, NQ2 = Q * Q;
    static constexpr int TP = pad2(G);       // row pitch of T[f][i][m]   (m contiguous)
    static constexpr int XP = pad2(Q);       // row pitch of X[arr][i][j] (j contiguous)
    static constexpr int PP = pad2(Q);       // row pitch of P[arr][m][i] (i contiguous)
    static constexpr int cmax(int a, int b) { return a > b ? a : b; }
    // ---- staged records (same order as RecLayout)
    static constexpr int S_GEOC = 0;
    static constexpr int S_QB = S_GEOC + RL::GEOC;
    static constexpr int S_NST = S_QB + RL::QB;
    static constexpr int S_ACCN = S_NST + RL::NST;
    static constexpr int S_Q0 = S_ACCN + RL::ACCN;
    static constexpr int S_Q2 = S_Q0 + RL::QB;
    static constexpr int S_QST = S_Q2 + RL::QB;
    static constexpr int S_FST = S_QST + RL::QST;
    static constexpr int S_ACCF = S_FST + RL::FST;
    static constexpr int S_VST = S_ACCF + RL::ACCF;
    static constexpr int S_TR = S_VST + RL::VST;
    // ---- operators: A[n + G*i], B, D[k + G*n], wq, wg, and the transposed copies At[n*XP + i], Bt
    static constexpr int S_OPS = S_TR + RL::TR;
    static constexpr int OPS_SZ = pad2(2 * G * Q + G * G + Q + G);
    static constexpr int S_AT = S_OPS + OPS_SZ;
    static constexpr int S_BT = S_AT + G * XP;
    // ---- work
    static constexpr int S_NW = S_BT + G * XP;              // pb, u, v
    static constexpr int S_T = S_NW + pad2(3 * NP);         // 12 sets [Q][TP]; later P: 8 arrays [G][PP]
    static constexpr int T_SZ = cmax(12 * Q * TP, 8 * G * PP);
    static constexpr int S_X = S_T + T_SZ;                  // 8 arrays [Q][XP]
    static constexpr int S_L = S_X + 8 * Q * XP;            // LDG: 0..3 G, 4..7 Zksi(2) Zeta(2), 8..9 lap
    static constexpr int S_OWN = S_L + 10 * NP;             // own face state [s][n][4]: pb pbpert mx my
    static constexpr int S_NBT = S_OWN + 16 * G;
    static constexpr int S_OWNG = S_NBT + 16 * G;           // own LDG gradient traces [s][4][G]
    static constexpr int S_NBG = S_OWNG + 16 * G;
    static constexpr int S_OWNV = S_NBG + 16 * G;           // own viscosity statics [s][5][G]
    static constexpr int S_LF = S_OWNV + 20 * G;            // LDG face flux [s][2][G]
    static constexpr int S_FF = S_LF + 8 * G;               // face fluxes [s][3][XP]
    static constexpr int S_PROJ = S_FF + 12 * XP;           // projected face fluxes [s][3][G]
    static constexpr int S_R = S_PROJ + 12 * G;             // rhs [3][NP]
    static constexpr int S_BAR = S_R + pad2(3 * NP);
    static constexpr int S_ACCQ = S_BAR + 2;                // last: 6 or 8 planes
    static int smem_doubles(int naccq) { return S_ACCQ + pad2(naccq * NQ2); }
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

template <int G, int Q>
__global__ void __launch_bounds__(128, 5) k_btp_stage_rec(const TmaArgs a, const int naccq) {
    using RL = RecLayout<G, Q>;
    using SL = RecSmem<G, Q>;
    constexpr int NP = RL::NP, NQ2 = RL::NQ2, TP = SL::TP, XP = SL::XP, PP = SL::PP;
    constexpr int NV = 96;  // threads of the volume pipeline (warps 0-2)
    static_assert(G * Q + 2 * NP <= NV && NQ2 <= NV && 2 * G * Q <= NV && 3 * NP <= NV, "polynomial order too high for this mapping");
    static_assert(4 * Q - 32 <= NV - NQ2, "face points do not fit");
    extern __shared__ __align__(128) double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* geo = sm + SL::S_GEOC;
    const int* conn = reinterpret_cast<const int*>(geo + 18);
    double* qb = sm + SL::S_QB;
    double* nst = sm + SL::S_NST;
    double* accn = sm + SL::S_ACCN;
    double* q0s = sm + SL::S_Q0;
    double* q2s = sm + SL::S_Q2;
    double* qst = sm + SL::S_QST;
    double* accq = sm + SL::S_ACCQ;
    double* fst = sm + SL::S_FST;
    double* accf = sm + SL::S_ACCF;
    double* vst = sm + SL::S_VST;
    double* trs = sm + SL::S_TR;
    double* opA = sm + SL::S_OPS;
    double* opB = opA + G * Q;
    double* opD = opB + G * Q;
    double* opWq = opD + G * G;
    double* opWg = opWq + Q;
    double* opAt = sm + SL::S_AT;
    double* opBt = sm + SL::S_BT;
    double* pbw = sm + SL::S_NW;
    double* uw = pbw + NP;
    double* vw = uw + NP;
    double* T = sm + SL::S_T;
    double* P = sm + SL::S_T;
    double* X = sm + SL::S_X;
    double* Lr = sm + SL::S_L;
    double* ownS = sm + SL::S_OWN;
    double* nbtS = sm + SL::S_NBT;
    double* ownG = sm + SL::S_OWNG;
    double* nbG = sm + SL::S_NBG;
    double* ownv = sm + SL::S_OWNV;
    double* lf = sm + SL::S_LF;
    double* ff = sm + SL::S_FF;
    double* proj = sm + SL::S_PROJ;
    double* R = sm + SL::S_R;
    void* bar = sm + SL::S_BAR;
    const bool visc = a.has_visc != 0;
    const int botfr = a.botfr;
    const uint32_t accq_bytes = (uint32_t)(naccq * NQ2 * sizeof(double) + 15) & ~15u;

    // operators to shared memory (once per block), mbarrier
    for (int t = tid; t < G * Q; t += 128) {
        double va = c_ops.A[t], vb = c_ops.B[t];
        opA[t] = va; opB[t] = vb;
        int n = t % G, i = t / G;
        opAt[n * XP + i] = va; opBt[n * XP + i] = vb;
    }
    for (int t = tid; t < G * G; t += 128) opD[t] = c_ops.D[t];
    for (int t = tid; t < Q; t += 128) opWq[t] = c_ops.wq[t];
    for (int t = tid; t < G; t += 128) opWg[t] = c_ops.wg[t];
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // this thread's column of psiq / dpsiq: index tid % Q on the volume warps, lane % Q on the face warp
    const int qi = (warp == 3 ? lane : tid) % Q;
    double Aq[G], Bq[G];
#pragma unroll
    for (int n = 0; n < G; ++n) { Aq[n] = opA[n + G * qi]; Bq[n] = opB[n + G * qi]; }

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

        // ---- A. nodal prep + nodal sums (mod_rk_mlswe.F90:90-92)
        if (tid < NP) {
            const int I = tid;
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
        __syncthreads();
        // ---- B. volume warps: interpolation pass 1 (threads < G*Q), LDG auxiliary variable (next 2*NP threads);
        //         face warp: state traces, own and neighbour (btp_extract_df)
        if (tid < G * Q) {
            // T[f][i][m] = sum_n psiq(n,i) N_f[m][n];   i = tid % Q = qi, m = tid / Q
            const int m = tid / Q;
            double* dst = T + qi * TP + m;
            auto pass1 = [&](const double* src, int f, bool deriv) {
                double s = 0.0;
#pragma unroll
                for (int n = 0; n < G; ++n) s = fma(deriv ? Bq[n] : Aq[n], src[m * G + n], s);
                dst[f * Q * TP] = s;
            };
            pass1(qb, 0, false); pass1(qb + NP, 1, false); pass1(qb + 2 * NP, 2, false); pass1(pbw, 3, false);
            if (botfr) { pass1(nst + 2 * NP, 4, false); pass1(nst + 3 * NP, 5, false); pass1(nst + 4 * NP, 6, false); }
            pass1(nst + 10 * NP, 7, false); pass1(nst + 11 * NP, 8, false); pass1(nst + 12 * NP, 9, false);
            pass1(nst + 13 * NP, 10, false);   // zbot, psiq rows  (for d/deta)
            pass1(nst + 13 * NP, 11, true);    // zbot, dpsiq rows (for d/dksi)
        } else if (tid < G * Q + 2 * NP) {
            if (visc) {
                // G = grad(ub or vb) at a node, flux variable and its weighted metric combinations
                const int idx = tid - G * Q, f = idx / NP, I = idx - f * NP, m = I / G, n = I - m * G;
                const double* src = f ? vw : uw;
                double dk = 0.0, de = 0.0;
#pragma unroll
                for (int k = 0; k < G; ++k) {
                    dk = fma(opD[k + G * n], src[m * G + k], dk);
                    de = fma(opD[k + G * m], src[k * G + n], de);
                }
                double ga = ksx * dk + etx * de, gb = ksy * dk + ety * de;
                Lr[(2 * f) * NP + I] = ga; Lr[(2 * f + 1) * NP + I] = gb;
                double pv = nst[5 * NP + I];
                double qa = pv * ga + nst[(6 + 2 * f) * NP + I], qbb = pv * gb + nst[(7 + 2 * f) * NP + I];
                double w = opWg[n] * opWg[m] * J;
                Lr[(4 + f) * NP + I] = w * (ksx * qa + ksy * qbb);
                Lr[(6 + f) * NP + I] = w * (etx * qa + ety * qbb);
            }
        } else if (warp == 3) {
            for (int it = lane; it < 4 * G; it += 32) {
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
                double* po = ownS + (s * G + n) * 4;
                double* pn = nbtS + (s * G + n) * 4;
                po[0] = pbw[I]; po[1] = ow0; po[2] = ow1; po[3] = ow2;
                pn[0] = n0 + vst[s * RL::VSIDE + 5 * G + n]; pn[1] = n0; pn[2] = n1; pn[3] = n2;
            }
        }
        __syncthreads();
        // ---- C. volume warps: interpolation pass 2 + pointwise physics (mod_rhs_btp.F90:136-192) per quadrature point,
        //         + the face points that do not fit the face warp;  face warp: LDG traces, LDG face flux, face fluxes
        auto face_point = [&](int s, int iq, const double (&Af)[G]) {
            // traces -> face quadrature point, canonical left perspective (mod_rhs_btp.F90:237-330)
            const bool left = (conn[s] < 0) || (e < conn[s]);
            const double* Ls = (left ? ownS : nbtS) + s * G * 4;
            const double* Rs = (left ? nbtS : ownS) + s * G * 4;
            double pbL = 0, ppL = 0, mxL = 0, myL = 0, pbR = 0, ppR = 0, mxR = 0, myR = 0;
#pragma unroll
            for (int n = 0; n < G; ++n) {
                const double2 l01 = *reinterpret_cast<const double2*>(Ls + n * 4), l23 = *reinterpret_cast<const double2*>(Ls + n * 4 + 2);
                const double2 r01 = *reinterpret_cast<const double2*>(Rs + n * 4), r23 = *reinterpret_cast<const double2*>(Rs + n * 4 + 2);
                pbL = fma(Af[n], l01.x, pbL); ppL = fma(Af[n], l01.y, ppL); mxL = fma(Af[n], l23.x, mxL); myL = fma(Af[n], l23.y, myL);
                pbR = fma(Af[n], r01.x, pbR); ppR = fma(Af[n], r01.y, ppR); mxR = fma(Af[n], r23.x, mxR); myR = fma(Af[n], r23.y, myR);
            }
            const double nxl = fg[s * 3 + 0], nyl = fg[s * 3 + 1], nlen = fg[s * 3 + 2];
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
            double wq = opWq[iq] * nlen;
            double dispu = 0.5 * lam * (mxR - mxL), dispv = 0.5 * lam * (myR - myL);
            double flux_x = nxl * quu + nyl * quv - dispu;
            double flux_y = nxl * qvu + nyl * qvv - dispv;
            double flux = nxl * fex + nyl * fey;
            double sgn = left ? -wq : wq;
            ff[(s * 3 + 0) * XP + iq] = sgn * flux;
            ff[(s * 3 + 1) * XP + iq] = sgn * (nxl * Hf + flux_x);
            ff[(s * 3 + 2) * XP + iq] = sgn * (nyl * Hf + flux_y);
        };
        if (tid < NQ2) {
            const int i = tid / Q;          // j = tid % Q = qi
            const int q = qi * Q + i;       // index in the records (intma_dg_quad order: j slow, i fast)
            const double* Ti = T + i * TP;
            auto pass2 = [&](int f, bool deriv) {
                double s = 0.0;
                const double* src = Ti + f * Q * TP;
#pragma unroll
                for (int m2 = 0; m2 < TP; m2 += 2) {
                    const double2 tv = *reinterpret_cast<const double2*>(src + m2);
                    if (m2 < G) s = fma(deriv ? Bq[m2] : Aq[m2], tv.x, s);
                    if (m2 + 1 < G) s = fma(deriv ? Bq[m2 + 1] : Aq[m2 + 1], tv.y, s);
                }
                return s;
            };
            double dpp = pass2(0, false), udp = pass2(1, false), vdp = pass2(2, false), dp = pass2(3, false);
            double wq = opWq[i] * opWq[qi] * J;
            double rdp = 1.0 / dp;
            double ub = udp * rdp, vb = vdp * rdp;
            double tb_u = 0.0, tb_v = 0.0;
            if (botfr) {
                double pp = pass2(4, false), up = pass2(5, false), vp = pass2(6, false);
                double ubot = up + ub, vbot = vp + vb;
                double spd = (botfr == 1) ? (a.cd / a.g) * pp : (a.cd / a.alpha_bot) * sqrt(ubot * ubot + vbot * vbot);
                tb_u = spd * ubot; tb_v = spd * vbot;
            }
            double fcor = pass2(7, false), twx = pass2(8, false), twy = pass2(9, false);
            double ze = pass2(10, true), zk = pass2(11, false);
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
            double* Xq = X + i * XP + qi;   // X[arr][i][j]
            Xq[0 * Q * XP] = wq * (ksx * udp + ksy * vdp);   // Fk1
            Xq[1 * Q * XP] = wq * (etx * udp + ety * vdp);   // Fe1
            Xq[2 * Q * XP] = wq * sc_x;                      // S2
            Xq[3 * Q * XP] = wq * (ksx * Fx2 + ksy * quv);   // Fk2
            Xq[4 * Q * XP] = wq * (etx * Fx2 + ety * quv);   // Fe2
            Xq[5 * Q * XP] = wq * sc_y;                      // S3
            Xq[6 * Q * XP] = wq * (ksx * quv + ksy * Fy3);   // Fk3
            Xq[7 * Q * XP] = wq * (etx * quv + ety * Fy3);   // Fe3
        } else if (tid < NV) {
            // face points beyond the 32 the face warp handles
            const int p = 32 + (tid - NQ2);
            if (p < 4 * Q) {
                const int s = p / Q, iq = p - s * Q;
                double Af[G];
#pragma unroll
                for (int n = 0; n < G; ++n) Af[n] = opA[n + G * iq];
                face_point(s, iq, Af);
            }
        } else {
            // face warp
            if (visc) {
                for (int it = lane; it < 4 * G; it += 32) {
                    int s = it / G, n = it - s * G;
                    const int nb = conn[s];
                    const double nx = fg[s * 3 + 0], ny = fg[s * 3 + 1], nlen = fg[s * 3 + 2];
                    const int I = face_node(s, n, G);
                    double go[4] = {Lr[0 * NP + I], Lr[1 * NP + I], Lr[2 * NP + I], Lr[3 * NP + I]}, gn[4];
                    if (nb >= 0 || nb == NBR_HALO) {
                        const double* tn = trs + s * RL::TSIDE + n;
#pragma unroll
                        for (int v = 0; v < 4; ++v) gn[v] = tn[(3 + v) * G];
                    } else {
#pragma unroll
                        for (int v = 0; v < 4; ++v) gn[v] = go[v];
                        if (nb == NBR_FREESLIP) reflect4(go, nx, ny, gn);
                    }
                    // LDG face flux at the face node, as written (mod_laplacian_quad.F90:427-519)
                    const bool left = (nb < 0) || (e < nb);
                    const double* sn = vst + s * RL::VSIDE + n;
                    double fo[4], fn[4];
                    const double pvo = nst[5 * NP + I];
#pragma unroll
                    for (int v = 0; v < 4; ++v) { fo[v] = pvo * go[v] + nst[(6 + v) * NP + I]; fn[v] = sn[4 * G] * gn[v] + sn[v * G]; }
                    const double* fl = left ? fo : fn;
                    const double* fr = left ? fn : fo;
                    double qu0 = 0.5 * fl[0] + 0.5 * fr[0], qu1 = 0.5 * fl[1] + 0.5 * fr[1];
                    double qv0 = 0.5 * fl[2] + 0.5 * fr[2], qv1 = 0.5 * fl[3] + 0.5 * fr[3];
                    double wq = opWg[n] * nlen;
                    double flux_qu = (qu0 - fl[0] * nx) + (qu1 - fl[1] * ny);
                    double flux_qv = (qv0 - fl[2] * nx) + (qv1 - fl[3] * ny);
                    double sgn = left ? wq : -wq;
                    lf[(s * 2 + 0) * G + n] = sgn * flux_qu;
                    lf[(s * 2 + 1) * G + n] = sgn * flux_qv;
                }
            }
            if (lane < 4 * Q) {
                const int s = lane / Q;   // iq = lane % Q = qi
                face_point(s, qi, Aq);
            }
        }
        if (warp < 3) {
            named_bar_sync(1, NV);
            // ---- D. weak-form scatter, contraction over j:  P[arr][m][i] = sum_j M(m,j) X[arr][i][j]
            //         threads < G*Q: psiq arrays Fk1 S2 Fk2 S3 Fk3;  next G*Q threads: dpsiq arrays Fe1 Fe2 Fe3
            if (tid < 2 * G * Q) {
                const bool deriv = tid >= G * Q;
                const int r = deriv ? tid - G * Q : tid, m = r % G, i = r / G;
                double Mrow[XP];
                const double* mt = (deriv ? opBt : opAt) + m * XP;
#pragma unroll
                for (int j = 0; j < XP; j += 2) { const double2 v2 = *reinterpret_cast<const double2*>(mt + j); Mrow[j] = v2.x; Mrow[j + 1] = v2.y; }
                auto scat1 = [&](int arr) {
                    const double* src = X + arr * Q * XP + i * XP;
                    double s1 = 0.0;
#pragma unroll
                    for (int j = 0; j < XP; j += 2) {
                        const double2 xv = *reinterpret_cast<const double2*>(src + j);
                        if (j < Q) s1 = fma(Mrow[j], xv.x, s1);
                        if (j + 1 < Q) s1 = fma(Mrow[j + 1], xv.y, s1);
                    }
                    P[arr * G * PP + m * PP + i] = s1;
                };
                if (!deriv) { scat1(0); scat1(2); scat1(3); scat1(5); scat1(6); }
                else { scat1(1); scat1(4); scat1(7); }
            }
            named_bar_sync(1, NV);
            // ---- E. contraction over i:  R[f][m][n] = sum_i dpsiq(n,i) P[Fk_f][m][i] + psiq(n,i) (P[Fe_f] + P[S_f])[m][i]
            if (tid < 3 * NP) {
                const int n = tid % G, fm = tid / G, f = fm / G, m = fm - f * G;
                const double* at = opAt + n * XP;
                const double* bt = opBt + n * XP;
                const double* pk = P + (3 * f) * G * PP + m * PP;
                const double* pe = P + (3 * f + 1) * G * PP + m * PP;
                const double* ps = P + (f > 0 ? 3 * f - 1 : 0) * G * PP + m * PP;
                const double sfac = f > 0 ? 1.0 : 0.0;
                double sB = 0.0, sA = 0.0;
#pragma unroll
                for (int i = 0; i < PP; i += 2) {
                    const double2 a2 = *reinterpret_cast<const double2*>(at + i), b2 = *reinterpret_cast<const double2*>(bt + i);
                    const double2 k2 = *reinterpret_cast<const double2*>(pk + i), e2 = *reinterpret_cast<const double2*>(pe + i);
                    const double2 s2 = *reinterpret_cast<const double2*>(ps + i);
                    if (i < Q) { sB = fma(b2.x, k2.x, sB); sA = fma(a2.x, e2.x + sfac * s2.x, sA); }
                    if (i + 1 < Q) { sB = fma(b2.y, k2.y, sB); sA = fma(a2.y, e2.y + sfac * s2.y, sA); }
                }
                R[f * NP + m * G + n] = sB + sA;
            }
        }
        __syncthreads();
        // ---- E2. warps 0-1: LDG volume term (btp_compute_laplacian);  warps 2-3: face fluxes projected on the face nodes
        if (tid < 2 * NP) {
            if (visc) {
                const int c = tid / NP, I = tid - c * NP, m = I / G, n = I - m * G;
                const double* zk = Lr + (4 + c) * NP;
                const double* ze = Lr + (6 + c) * NP;
                double lx = 0.0, le = 0.0;
#pragma unroll
                for (int k = 0; k < G; ++k) {
                    lx = fma(opD[n + G * k], zk[m * G + k], lx);
                    le = fma(opD[m + G * k], ze[k * G + n], le);
                }
                Lr[(8 + c) * NP + I] = -(lx + le);
            }
        } else if (tid >= 64 && tid < 64 + 12 * G) {
            const int idx = tid - 64, sf = idx / G, n = idx - sf * G;
            const double* at = opAt + n * XP;
            const double* src = ff + sf * XP;
            double s1 = 0.0;
#pragma unroll
            for (int iq = 0; iq < XP; iq += 2) {
                const double2 a2 = *reinterpret_cast<const double2*>(at + iq), f2 = *reinterpret_cast<const double2*>(src + iq);
                if (iq < Q) s1 = fma(a2.x, f2.x, s1);
                if (iq + 1 < Q) s1 = fma(a2.y, f2.y, s1);
            }
            proj[sf * G + n] = s1;
        }
        __syncthreads();
        // ---- U. gather per node, mass matrix, viscosity, SSPRK update, wall projection (mod_rk_mlswe.F90:97-108)
        if (tid < NP) {
            const int I = tid, m = I / G, n = I - m * G;
            double r0 = R[I], r1 = R[NP + I], r2 = R[2 * NP + I];
            double l0 = 0.0, l1 = 0.0;
            if (visc) { l0 = Lr[8 * NP + I]; l1 = Lr[9 * NP + I]; }
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
            if (visc) {
                double pbn_ = qn[0] + nst[I];
                uw[I] = qn[1] / pbn_; vw[I] = qn[2] / pbn_;
            }
        }
        __syncthreads();
        // ---- G2. LDG gradient of the new state (only its face traces are needed by the next stage)
        if (visc) {
            if (tid < 2 * NP) {
                const int f = tid / NP, I = tid - f * NP, m = I / G, n = I - m * G;
                const double* src = f ? vw : uw;
                double dk = 0.0, de = 0.0;
#pragma unroll
                for (int k = 0; k < G; ++k) {
                    dk = fma(opD[k + G * n], src[m * G + k], dk);
                    de = fma(opD[k + G * m], src[k * G + n], de);
                }
                Lr[(2 * f) * NP + I] = ksx * dk + etx * de; Lr[(2 * f + 1) * NP + I] = ksy * dk + ety * de;
            }
            __syncthreads();
        }
        // ---- TR. traces of the new state for the next stage
        if (tid < 4 * G) {
            int s = tid / G, n = tid - s * G;
            int I = face_node(s, n, G);
            double* to = trs + s * RL::TSIDE + n;
            to[0] = qb[I]; to[G] = qb[NP + I]; to[2 * G] = qb[2 * NP + I];
#pragma unroll
            for (int v = 0; v < 4; ++v) to[(3 + v) * G] = visc ? Lr[v * NP + I] : 0.0;
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

template <int G, int Q>
static int launch_rec_t(Solver& S, const TmaArgs& a, int naccq) {
    using SL = RecSmem<G, Q>;
    size_t smem = (size_t)SL::smem_doubles(naccq) * sizeof(double);
    static int blocks_per_sm = 0;
    static size_t configured_smem = 0;
    if (configured_smem != smem) {
        if (cudaFuncSetAttribute(k_btp_stage_rec<G, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            set_error("cudaFuncSetAttribute", "shared memory opt-in failed"); return -1;
        }
        cudaFuncSetAttribute(k_btp_stage_rec<G, Q>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_btp_stage_rec<G, Q>, 128, smem) != cudaSuccess || blocks_per_sm < 1) {
            set_error("cudaOccupancy", "stage kernel does not fit on an SM"); return -1;
        }
        configured_smem = smem;
    }
    int bps = S.tma_blocks_per_sm > 0 ? std::min(S.tma_blocks_per_sm, blocks_per_sm) : blocks_per_sm;
    int grid = std::min(S.nelem, S.num_sms * bps);
    k_btp_stage_rec<G, Q><<<grid, 128, smem, S.stream>>>(a, naccq);
    S.n_launches++;
    return 0;
}
inline int launch_stage_rec(Solver& S, const TmaArgs& a, int naccq) {
    if (S.ngl == 5 && S.nq == 9) return launch_rec_t<5, 9>(S, a, naccq);
    if (S.ngl == 4 && S.nq == 7) return launch_rec_t<4, 7>(S, a, naccq);
    return -1;
}

}  // namespace hn
