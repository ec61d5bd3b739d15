// btp_kernels.cuh -- barotropic substep loop on the device.
//
// Replaces (reference file:line)
//   ti_barotropic_ssprk_mlswe          src/mod_rk_mlswe.F90:19-151
//   create_rhs_btp                     src/mod_rhs_btp.F90:28-59
//     btp_extract_df                   src/mod_barotropic_terms.F90:25-97      (face traces + ghost state)
//     create_rhs_btp_volume_qdf        src/mod_rhs_btp.F90:102-209             (volume integral)
//     creat_btp_fluxes_qdf             src/mod_rhs_btp.F90:211-370             (face fluxes)
//     btp_create_laplacian             src/mod_laplacian_quad.F90:32-121,357-390,427-519 (LDG viscosity)
//   SSPRK update + btp_mom_boundary_df src/mod_rk_mlswe.F90:97-114, src/mod_barotropic_terms.F90:165-217
//
// One fused kernel per SSPRK stage.  Each element (one thread block in this "simple" variant) computes its
// volume term and all four face terms from the canonical left-element perspective, applies the stage update and
// the wall projection, and publishes the face traces (state + LDG gradient) of the NEW state for its neighbours
// in a ping-pong trace buffer -- the same buffer is the halo send buffer on processor boundaries.
//
// Time averages: only the sums that are nonlinear in the state are accumulated per stage (6 per quadrature
// point, 11 per face quadrature point, 10 per node); the linear ones are reconstructed after the loop by
// k_btp_finalize from nodal sums (SURVEY.md section 7 "accumulator diet").
#pragma once
#include "hnumo_dev.cuh"
#include "visc_q.cuh"

namespace hn {

enum { TR_PBPERT = 0, TR_MX = 1, TR_MY = 2, TR_G = 3, TR_NV = 7 };

struct StageArgs {
    Mesh M;
    double* qb[3];
    double* qb0[3];
    double* qb2[3];
    const double* tr_in;
    double* tr_out;
    size_t trstride;
    const double *pbprime_df, *oop_df, *massinv;
    const double *oop_q, *coriolis_q, *tauwx_q, *tauwy_q, *gzx_q, *gzy_q;
    const double *Quu, *Quv, *Qvv, *Hbcl, *Quu_e, *Quv_e, *Qvv_e, *Hbcl_e;
    const double *cL, *cR, *cLR, *lam, *oop_edge, *pbl, *pbr, *pbn;
    const double *qp_dp, *qp_u, *qp_v;
    const double* bdg[4];
    const double* pbv;
    const double* hstat;  // halo copies of (bdg[0..3], pbv) traces: [5][nhalo*ngl]
    size_t hstat_stride;
    double* acc_n[10];
    double* acc_q[8];
    double* acc_f[11];
    double* rhs_out[3];
    double a1, a2, a3, dtt, g, cd, alpha_bot, visc;
    int botfr, has_visc, load_q0, load_q2, store_q0, store_q2, rhs_only;
    int acc_graduvb;  // 1: accumulate graduvb per stage (reference form); 0: derived after the loop from the nodal velocity sums
    // method_visc == 1 (visc_q.cuh): flux variable at the quadrature points F = S + P grad(ub,vb), face values in slot planes
    int visc_q;
    const double* vqP;
    const double* vqS[4];
    const double* trq_in;
    double* trq_out;
    size_t trq_stride;
};

// value of trace variable v at face node n of (e,s) as seen from the neighbour's side ("side 2" of
// btp_extract_df when e is the left element).  own[] holds this element's own face values.
__device__ __forceinline__ void neighbour_state(const StageArgs& a, int e, int s, int n, int nb, int nbs, double nx,
                                                double ny, const double own[3], double out[3]) {
    const int ngl = a.M.ngl;
    if (nb >= 0 || nb == NBR_HALO) {
        size_t base = (nb >= 0) ? ((size_t)nb * 4 + nbs) * ngl + n : ((size_t)a.M.nslots + nbs) * ngl + n;
        out[0] = a.tr_in[TR_PBPERT * a.trstride + base];
        out[1] = a.tr_in[TR_MX * a.trstride + base];
        out[2] = a.tr_in[TR_MY * a.trstride + base];
    } else {
        out[0] = own[0]; out[1] = own[1]; out[2] = own[2];
        if (nb == NBR_FREESLIP) {
            double un = nx * own[1] + ny * own[2];
            out[1] = own[1] - 2.0 * un * nx;
            out[2] = own[2] - 2.0 * un * ny;
        } else if (nb == NBR_NOSLIP) {
            out[1] = -own[1]; out[2] = -own[2];
        }
    }
}

// reflect a 4-vector of gradients (G1,G2),(G3,G4) about the normal (free-slip ghost, mod_laplacian_quad.F90:85-98)
__device__ __forceinline__ void reflect4(const double in[4], double nx, double ny, double out[4]) {
    double un = in[0] * nx + in[1] * ny;
    out[0] = in[0] - 2.0 * un * nx; out[1] = in[1] - 2.0 * un * ny;
    un = in[2] * nx + in[3] * ny;
    out[2] = in[2] - 2.0 * un * nx; out[3] = in[3] - 2.0 * un * ny;
}

// shared memory plan of k_btp_stage_simple (doubles)
struct StageSmem {
    int ops, nod, tmp, fq, tP, tR, rhs, lap, own, nbt, nbv, ownv, ff, lf, vF, lfq, total;
    __host__ __device__ StageSmem(int ngl, int nq) {
        int npts = ngl * ngl, nq2 = nq * nq, per = ngl * nq;
        int o = 0;
        ops = o; o += 2 * per + npts + nq + ngl;
        nod = o; o += 13 * npts;  // 0 pb 1 dpp 2 mx 3 my 4 pp 5 up 6 vp 7 u 8 v 9..12 G
        tmp = o; o += 7 * per;
        fq = o; o += 8 * nq2;     // 0 Fk1 1 Fe1 2 S2 3 Fk2 4 Fe2 5 S3 6 Fk3 7 Fe3
        tP = o; o += 3 * per;
        tR = o; o += 3 * per;
        rhs = o; o += 3 * npts;
        lap = o; o += 2 * npts;
        own = o; o += 4 * 7 * ngl;   // own traces [s][v][n]
        nbt = o; o += 4 * 7 * ngl;   // neighbour traces
        nbv = o; o += 4 * 5 * ngl;   // neighbour viscosity statics
        ownv = o; o += 4 * 5 * ngl;  // own viscosity statics
        ff = o; o += 4 * 3 * nq;     // face flux at quadrature points [s][f][iq]
        lf = o; o += 4 * 2 * ngl;    // LDG face flux at face nodes [s][c][n]
        vF = o; o += 4 * nq2;        // method_visc 1: weighted flux variable Fk_u Fk_v | Fe_u Fe_v
        lfq = o; o += 4 * 2 * nq;    // method_visc 1: LDG face flux at face quadrature points [s][c][iq]
        total = o;
    }
};

__global__ void k_btp_stage_simple(StageArgs a) {
    extern __shared__ double sm[];
    const int ngl = a.M.ngl, nq = a.M.nq, npts = a.M.npts, nq2 = a.M.nq2, per = ngl * nq;
    const int e = blockIdx.x, tid = threadIdx.x;
    const StageSmem L(ngl, nq);
    SOps o = load_sops(sm + L.ops, ngl, nq);
    double* nod = sm + L.nod;
    double* tmp = sm + L.tmp;
    double* fq = sm + L.fq;
    double* rhs = sm + L.rhs;
    double* lap = sm + L.lap;
    double* own = sm + L.own;
    double* nbt = sm + L.nbt;
    double* nbv = sm + L.nbv;
    double* ownv = sm + L.ownv;
    double* ff = sm + L.ff;
    double* lf = sm + L.lf;
    double* vF = sm + L.vF;
    double* lfq = sm + L.lfq;
    const bool nodal_visc = a.has_visc && !a.visc_q, quad_visc = a.has_visc && a.visc_q;
    // metric terms and normals are read per point (met_q, met_n, fg_q, fg_n: per element / per side on affine meshes)
    const size_t nbase = (size_t)e * npts, qbase = (size_t)e * nq2;
    const bool acc = !a.rhs_only;

    // ---- 1. nodal loads, nodal accumulators (mod_rk_mlswe.F90:90-92)
    if (tid < npts) {
        double dpp = a.qb[0][nbase + tid], mx = a.qb[1][nbase + tid], my = a.qb[2][nbase + tid];
        double pb = dpp + a.pbprime_df[nbase + tid];
        nod[0 * npts + tid] = pb; nod[1 * npts + tid] = dpp; nod[2 * npts + tid] = mx; nod[3 * npts + tid] = my;
        if (a.botfr) {
            nod[4 * npts + tid] = a.qp_dp[nbase + tid]; nod[5 * npts + tid] = a.qp_u[nbase + tid];
            nod[6 * npts + tid] = a.qp_v[nbase + tid];
        }
        double u = mx / pb, v = my / pb;
        nod[7 * npts + tid] = u; nod[8 * npts + tid] = v;
        if (acc) {
            double t = 1.0 + dpp * a.oop_df[nbase + tid];
            a.acc_n[0][nbase + tid] += t * t;
            a.acc_n[1][nbase + tid] += u;
            a.acc_n[2][nbase + tid] += v;
            a.acc_n[3][nbase + tid] += dpp;
            a.acc_n[4][nbase + tid] += mx;
            a.acc_n[5][nbase + tid] += my;
        }
    }
    __syncthreads();
    // ---- 2. LDG auxiliary variable G = grad(ub,vb) at the nodes (mod_laplacian_quad.F90:50-56)
    if (nodal_visc && tid < npts) {
        int m = tid / ngl, n = tid - m * ngl;
        const Met mt = met_n(a.M, e, tid);
        const double ksx = mt.ksx, ksy = mt.ksy, etx = mt.etx, ety = mt.ety;
        double dk, de;
        nodal_grad(o, ngl, nod + 7 * npts, n, m, dk, de);
        double g0 = ksx * dk + etx * de, g1 = ksy * dk + ety * de;
        nodal_grad(o, ngl, nod + 8 * npts, n, m, dk, de);
        double g2 = ksx * dk + etx * de, g3 = ksy * dk + ety * de;
        nod[9 * npts + tid] = g0; nod[10 * npts + tid] = g1; nod[11 * npts + tid] = g2; nod[12 * npts + tid] = g3;
        if (acc) {
            a.acc_n[6][nbase + tid] += g0; a.acc_n[7][nbase + tid] += g1;
            a.acc_n[8][nbase + tid] += g2; a.acc_n[9][nbase + tid] += g3;
        }
    }
    // ---- 3. interpolate pb,dpp,mx,my (+ bottom-layer primes) to the quadrature points
    const int NF = a.botfr ? 7 : 4;
    sf_pass1(o, ngl, nq, NF, nod, npts, tmp, nullptr);
    if (quad_visc) sf_pass1(o, ngl, nq, 2, nod + 7 * npts, npts, sm + L.tR, sm + L.tP);   // ub, vb: A-pass in tR, B-pass in tP
    __syncthreads();
    // ---- 4. pointwise physics (mod_rhs_btp.F90:136-192)
    if (tid < nq2) {
        int j = tid / nq, i = tid - j * nq;
        double dp = sf_eval(o, ngl, nq, tmp, 0, i, j), dpp = sf_eval(o, ngl, nq, tmp, 1, i, j);
        double udp = sf_eval(o, ngl, nq, tmp, 2, i, j), vdp = sf_eval(o, ngl, nq, tmp, 3, i, j);
        const Met mt = met_q(a.M, e, tid);
        const double ksx = mt.ksx, ksy = mt.ksy, etx = mt.etx, ety = mt.ety;
        double wq = o.wq[i] * o.wq[j] * mt.J;
        double ub = udp / dp, vb = vdp / dp;
        double tb_u = 0.0, tb_v = 0.0;
        if (a.botfr) {
            double pp = sf_eval(o, ngl, nq, tmp, 4, i, j), up = sf_eval(o, ngl, nq, tmp, 5, i, j),
                   vp = sf_eval(o, ngl, nq, tmp, 6, i, j);
            double ubot = up + ub, vbot = vp + vb;
            double spd = (a.botfr == 1) ? (a.cd / a.g) * pp : (a.cd / a.alpha_bot) * sqrt(ubot * ubot + vbot * vbot);
            tb_u = spd * ubot; tb_v = spd * vbot;
        }
        size_t Iq = qbase + tid;
        double fcor = a.coriolis_q[Iq];
        double sc_x = fcor * vdp + a.g * (a.tauwx_q[Iq] - tb_u) - a.g * dp * a.gzx_q[Iq];
        double sc_y = -fcor * udp + a.g * (a.tauwy_q[Iq] - tb_v) - a.g * dp * a.gzy_q[Iq];
        double ope = 1.0 + dpp * a.oop_q[Iq];
        double Hq = (ope * ope) * a.Hbcl[Iq];
        double qu = ub * udp + ope * a.Quu[Iq];
        double quv = ub * vdp + ope * a.Quv[Iq];
        double qv = vb * vdp + ope * a.Qvv[Iq];
        if (acc) {
            a.acc_q[0][Iq] += qu; a.acc_q[1][Iq] += qv; a.acc_q[2][Iq] += quv; a.acc_q[3][Iq] += ope * ope;
            a.acc_q[4][Iq] += ub; a.acc_q[5][Iq] += vb;
            if (a.botfr == 2) { a.acc_q[6][Iq] += tb_u; a.acc_q[7][Iq] += tb_v; }
        }
        // weak form: rhs_I += wq (psi S + dpsi/dx Fx + dpsi/dy Fy); dpsi/dx = ksi_x d/dksi + eta_x d/deta
        double Fx1 = udp, Fy1 = vdp, Fx2 = Hq + qu, Fy2 = quv, Fx3 = quv, Fy3 = Hq + qv;
        fq[0 * nq2 + tid] = wq * (ksx * Fx1 + ksy * Fy1); fq[1 * nq2 + tid] = wq * (etx * Fx1 + ety * Fy1);
        fq[2 * nq2 + tid] = wq * sc_x;
        fq[3 * nq2 + tid] = wq * (ksx * Fx2 + ksy * Fy2); fq[4 * nq2 + tid] = wq * (etx * Fx2 + ety * Fy2);
        fq[5 * nq2 + tid] = wq * sc_y;
        fq[6 * nq2 + tid] = wq * (ksx * Fx3 + ksy * Fy3); fq[7 * nq2 + tid] = wq * (etx * Fx3 + ety * Fy3);
        if (quad_visc) {   // btp_create_laplacian_v2 + compute_laplacian_quad (mod_laplacian_quad.F90:125-159,613-642)
            const double uk = sf_eval(o, ngl, nq, sm + L.tP, 0, i, j), ue = sf_eval_B(o, ngl, nq, sm + L.tR, 0, i, j);
            const double vk = sf_eval(o, ngl, nq, sm + L.tP, 1, i, j), ve = sf_eval_B(o, ngl, nq, sm + L.tR, 1, i, j);
            const double pq = a.vqP[Iq];
            const double F0 = a.vqS[0][Iq] + pq * (ksx * uk + etx * ue), F1 = a.vqS[1][Iq] + pq * (ksy * uk + ety * ue);
            const double F2 = a.vqS[2][Iq] + pq * (ksx * vk + etx * ve), F3 = a.vqS[3][Iq] + pq * (ksy * vk + ety * ve);
            vF[0 * nq2 + tid] = -wq * (ksx * F0 + ksy * F1); vF[1 * nq2 + tid] = -wq * (ksx * F2 + ksy * F3);
            vF[2 * nq2 + tid] = -wq * (etx * F0 + ety * F1); vF[3 * nq2 + tid] = -wq * (etx * F2 + ety * F3);
        }
    }
    __syncthreads();
    // ---- 5. scatter to the nodes (field 1 has no source term)
    sf_scatter(o, ngl, nq, 1, nullptr, fq + 0 * nq2, fq + 1 * nq2, nq2, sm + L.tP, sm + L.tR, rhs, npts, false);
    sf_scatter(o, ngl, nq, 1, fq + 2 * nq2, fq + 3 * nq2, fq + 4 * nq2, nq2, sm + L.tP, sm + L.tR, rhs + npts, npts, false);
    sf_scatter(o, ngl, nq, 1, fq + 5 * nq2, fq + 6 * nq2, fq + 7 * nq2, nq2, sm + L.tP, sm + L.tR, rhs + 2 * npts, npts, false);
    if (quad_visc) sf_scatter(o, ngl, nq, 2, nullptr, vF, vF + 2 * nq2, nq2, sm + L.tP, sm + L.tR, lap, npts, false);

    // ---- 6. face traces: own and neighbour (btp_extract_df) + LDG gradient traces and viscosity statics
    if (tid < 4 * ngl) {
        int s = tid / ngl, n = tid - s * ngl;
        int slot = e * 4 + s, nb = a.M.nbr[slot], nbs = a.M.nbslot[slot];
        const FGeo fgn_ = fg_n(a.M, slot, n);   // nodal normal (normal_vector): ghost states are built at the face nodes
        double nx = fgn_.nx, ny = fgn_.ny;
        int I = face_node(s, n, ngl);
        double ow[3] = {nod[1 * npts + I], nod[2 * npts + I], nod[3 * npts + I]}, nbv3[3];
        neighbour_state(a, e, s, n, nb, nbs, nx, ny, ow, nbv3);
        double* po = own + (s * 7) * ngl + n;
        double* pn = nbt + (s * 7) * ngl + n;
        for (int v = 0; v < 3; ++v) { po[v * ngl] = ow[v]; pn[v * ngl] = nbv3[v]; }
        if (nodal_visc) {
            double go[4] = {nod[9 * npts + I], nod[10 * npts + I], nod[11 * npts + I], nod[12 * npts + I]}, gn[4];
            double so[5] = {a.bdg[0][nbase + I], a.bdg[1][nbase + I], a.bdg[2][nbase + I], a.bdg[3][nbase + I],
                            a.pbv[nbase + I]}, sn[5];
            if (nb >= 0) {
                size_t base = ((size_t)nb * 4 + nbs) * ngl + n;
                for (int v = 0; v < 4; ++v) gn[v] = a.tr_in[(TR_G + v) * a.trstride + base];
                size_t In = (size_t)nb * npts + face_node(nbs, n, ngl);
                for (int v = 0; v < 4; ++v) sn[v] = a.bdg[v][In];
                sn[4] = a.pbv[In];
            } else if (nb == NBR_HALO) {
                size_t base = ((size_t)a.M.nslots + nbs) * ngl + n;
                for (int v = 0; v < 4; ++v) gn[v] = a.tr_in[(TR_G + v) * a.trstride + base];
                for (int v = 0; v < 5; ++v) sn[v] = a.hstat[v * a.hstat_stride + (size_t)nbs * ngl + n];
            } else {
                for (int v = 0; v < 4; ++v) { gn[v] = go[v]; sn[v] = so[v]; }
                sn[4] = so[4];
                if (nb == NBR_FREESLIP) { reflect4(go, nx, ny, gn); reflect4(so, nx, ny, sn); }
            }
            for (int v = 0; v < 4; ++v) { po[(3 + v) * ngl] = go[v]; pn[(3 + v) * ngl] = gn[v]; }
            for (int v = 0; v < 5; ++v) { ownv[(s * 5 + v) * ngl + n] = so[v]; nbv[(s * 5 + v) * ngl + n] = sn[v]; }
        }
    }
    __syncthreads();
    // ---- 7. face fluxes at the face quadrature points, canonical left perspective (mod_rhs_btp.F90:237-330)
    if (tid < 4 * nq) {
        int s = tid / nq, iq = tid - s * nq;
        int slot = e * 4 + s, nb = a.M.nbr[slot], nbs = a.M.nbslot[slot];
        bool left = (nb < 0) || (e < nb);
        int oslot = left ? slot : nb * 4 + nbs;
        const FGeo fgq_ = fg_q(a.M, slot, iq);  // normal_vector_q, jac_faceq
        double nxl = fgq_.nx, nyl = fgq_.ny, nlen = fgq_.len;
        double nxr = -nxl, nyr = -nyl;
        const double* tl = (left ? own : nbt) + s * 7 * ngl;
        const double* tr = (left ? nbt : own) + s * 7 * ngl;
        // pb on either side: pbpert + pbprime_df at that side's node
        double qbl[4] = {0, 0, 0, 0}, qbr[4] = {0, 0, 0, 0};
        for (int n = 0; n < ngl; ++n) {
            double hi = o.A[n + ngl * iq];
            double pbo = nod[face_node(s, n, ngl)];                                     // own pb = pbpert + pbprime_df
            double pbnb = nbt[s * 7 * ngl + n] + a.pbn[(size_t)slot * ngl + n];         // neighbour's pb
            qbl[0] += hi * (left ? pbo : pbnb); qbr[0] += hi * (left ? pbnb : pbo);
            qbl[1] += hi * tl[0 * ngl + n]; qbr[1] += hi * tr[0 * ngl + n];
            qbl[2] += hi * tl[1 * ngl + n]; qbr[2] += hi * tr[1 * ngl + n];
            qbl[3] += hi * tl[2 * ngl + n]; qbr[3] += hi * tr[2 * ngl + n];
        }
        size_t fo = (size_t)oslot * nq + iq;
        double cL = a.cL[fo], cR = a.cR[fo], cLR = a.cLR[fo], lam = a.lam[fo];
        double pU_L = nxl * qbl[2] + nyl * qbl[3];
        double pU_R = nxr * qbr[2] + nyr * qbr[3];
        double pbpert_edge = cL * qbl[1] + cR * qbr[1] + cLR * (pU_L + pU_R);
        double ope_e = 1.0 + pbpert_edge * a.oop_edge[fo];
        // coeff_mass_pbub_L == coeff_pbpert_R and coeff_mass_pbub_R == coeff_pbpert_L (mod_initial_mlswe.F90:382-396)
        double fex = cR * qbl[2] + cL * qbr[2] + lam * (nxl * qbl[1] + nxr * qbr[1]);
        double fey = cR * qbl[3] + cL * qbr[3] + lam * (nyl * qbl[1] + nyr * qbr[1]);
        double ul = qbl[2] / qbl[0], ur = qbr[2] / qbr[0], vl = qbl[3] / qbl[0], vr = qbr[3] / qbr[0];
        double quu = 0.5 * (ul * qbl[2] + ur * qbr[2]) + ope_e * a.Quu_e[fo];
        double quv = 0.5 * (vl * qbl[2] + vr * qbr[2]) + ope_e * a.Quv_e[fo];
        double qvu = 0.5 * (ul * qbl[3] + ur * qbr[3]) + ope_e * a.Quv_e[fo];
        double qvv = 0.5 * (vl * qbl[3] + vr * qbr[3]) + ope_e * a.Qvv_e[fo];
        double Hf = (ope_e * ope_e) * a.Hbcl_e[fo];
        if (acc && left) {
            double ol = 1.0 + (qbl[1] / a.pbl[fo]), orr = 1.0 + (qbr[1] / a.pbr[fo]);
            a.acc_f[0][fo] += quu; a.acc_f[1][fo] += quv; a.acc_f[2][fo] += qvu; a.acc_f[3][fo] += qvv;
            a.acc_f[4][fo] += ol * ol; a.acc_f[5][fo] += orr * orr; a.acc_f[6][fo] += ope_e * ope_e;
            a.acc_f[7][fo] += ul; a.acc_f[8][fo] += ur; a.acc_f[9][fo] += vl; a.acc_f[10][fo] += vr;
        }
        double wq = o.wq[iq] * nlen;
        double dispu = 0.5 * lam * (qbr[2] - qbl[2]), dispv = 0.5 * lam * (qbr[3] - qbl[3]);
        double flux_x = nxl * quu + nyl * quv - dispu;
        double flux_y = nxl * qvu + nyl * qvv - dispv;
        double flux = nxl * fex + nyl * fey;
        double sgn = left ? -1.0 : 1.0;
        ff[(s * 3 + 0) * nq + iq] = sgn * wq * flux;
        ff[(s * 3 + 1) * nq + iq] = sgn * wq * (nxl * Hf + flux_x);
        ff[(s * 3 + 2) * nq + iq] = sgn * wq * (nyl * Hf + flux_y);
    }
    // ---- 7b. LDG face flux at the face nodes, as written (mod_laplacian_quad.F90:427-519)
    if (quad_visc) {   // create_rhs_laplacian_flux_quad (mod_laplacian_quad.F90:644-722) on the published face values
        for (int t = tid; t < 4 * nq; t += blockDim.x) {
            const int s = t / nq, iq = t - s * nq;
            double fl[2];
            vq_face_flux(a.M, a.trq_in, a.trq_stride, e, s, iq, o.wq[iq], fl);
            lfq[(s * 2 + 0) * nq + iq] = fl[0]; lfq[(s * 2 + 1) * nq + iq] = fl[1];
        }
    }
    if (nodal_visc && tid >= 4 * nq && tid < 4 * nq + 4 * ngl) {
        int t = tid - 4 * nq, s = t / ngl, n = t - s * ngl;
        int slot = e * 4 + s, nb = a.M.nbr[slot];
        bool left = (nb < 0) || (e < nb);
        const FGeo fgn_ = fg_n(a.M, slot, n);   // normal_vector, jac_face
        double nx = fgn_.nx, ny = fgn_.ny, nlen = fgn_.len;
        const double* gl = (left ? own : nbt) + (s * 7 + 3) * ngl + n;
        const double* gr = (left ? nbt : own) + (s * 7 + 3) * ngl + n;
        const double* sl = (left ? ownv : nbv) + s * 5 * ngl + n;
        const double* sr = (left ? nbv : ownv) + s * 5 * ngl + n;
        double fl[4], fr[4];
        for (int v = 0; v < 4; ++v) {
            fl[v] = sl[4 * ngl] * gl[v * ngl] + sl[v * ngl];
            fr[v] = sr[4 * ngl] * gr[v * ngl] + sr[v * ngl];
        }
        double qu0 = 0.5 * fl[0] + 0.5 * fr[0], qu1 = 0.5 * fl[1] + 0.5 * fr[1];
        double qv0 = 0.5 * fl[2] + 0.5 * fr[2], qv1 = 0.5 * fl[3] + 0.5 * fr[3];
        double wq = o.wg[n] * nlen;
        double flux_qu = (qu0 - fl[0] * nx) + (qu1 - fl[1] * ny);
        double flux_qv = (qv0 - fl[2] * nx) + (qv1 - fl[3] * ny);
        double sgn = left ? 1.0 : -1.0;
        lf[(s * 2 + 0) * ngl + n] = sgn * wq * flux_qu;
        lf[(s * 2 + 1) * ngl + n] = sgn * wq * flux_qv;
    }
    // ---- 7c. LDG volume term (btp_compute_laplacian, mod_laplacian_quad.F90:357-390); collocation weak form
    if (nodal_visc && tid < npts) {
        // stage qq = pbprime_visc*G + btp_dpp_graduv in tmp (reuse): tmp[c][node], c=0..3
        double pv = a.pbv[nbase + tid];
        for (int v = 0; v < 4; ++v) tmp[v * npts + tid] = pv * nod[(9 + v) * npts + tid] + a.bdg[v][nbase + tid];
    }
    __syncthreads();
    if (nodal_visc && tid < npts) {
        int m = tid / ngl, n = tid - m * ngl;
        double l0 = 0.0, l1 = 0.0;
        for (int k = 0; k < ngl; ++k) {
            // derivative of basis (n,m) at node (k,m): D(n,k) ; at node (n,k): D(m,k)
            int I1 = m * ngl + k, I2 = k * ngl + n;
            const Met m1 = met_n(a.M, e, I1), m2 = met_n(a.M, e, I2);   // metric terms at the collocation point of the sum
            double wk1 = o.wg[k] * o.wg[m] * m1.J * o.D[n + ngl * k];
            double wk2 = o.wg[n] * o.wg[k] * m2.J * o.D[m + ngl * k];
            l0 -= wk1 * (m1.ksx * tmp[0 * npts + I1] + m1.ksy * tmp[1 * npts + I1]) + wk2 * (m2.etx * tmp[0 * npts + I2] + m2.ety * tmp[1 * npts + I2]);
            l1 -= wk1 * (m1.ksx * tmp[2 * npts + I1] + m1.ksy * tmp[3 * npts + I1]) + wk2 * (m2.etx * tmp[2 * npts + I2] + m2.ety * tmp[3 * npts + I2]);
        }
        lap[tid] = l0; lap[npts + tid] = l1;
    }
    __syncthreads();
    // ---- 8. gather face contributions per node, mass matrix, viscosity, SSPRK update, wall projection
    if (tid < npts) {
        int m = tid / ngl, n = tid - m * ngl;
        double r0 = rhs[tid], r1 = rhs[npts + tid], r2 = rhs[2 * npts + tid];
        double l0 = a.has_visc ? lap[tid] : 0.0, l1 = a.has_visc ? lap[npts + tid] : 0.0;
        for (int s = 0; s < 4; ++s) {
            int nf;
            if (s == 0) { if (m != 0) continue; nf = n; }
            else if (s == 1) { if (m != ngl - 1) continue; nf = n; }
            else if (s == 2) { if (n != 0) continue; nf = m; }
            else { if (n != ngl - 1) continue; nf = m; }
            double p0 = 0.0, p1 = 0.0, p2 = 0.0;
            for (int iq = 0; iq < nq; ++iq) {
                double hi = o.A[nf + ngl * iq];
                p0 += hi * ff[(s * 3 + 0) * nq + iq]; p1 += hi * ff[(s * 3 + 1) * nq + iq]; p2 += hi * ff[(s * 3 + 2) * nq + iq];
            }
            r0 += p0; r1 += p1; r2 += p2;
            if (nodal_visc) { l0 += lf[(s * 2 + 0) * ngl + nf]; l1 += lf[(s * 2 + 1) * ngl + nf]; }
            if (quad_visc) {
                double v0 = 0.0, v1 = 0.0;
                for (int iq = 0; iq < nq; ++iq) {
                    double hi = o.A[nf + ngl * iq];
                    v0 += hi * lfq[(s * 2 + 0) * nq + iq]; v1 += hi * lfq[(s * 2 + 1) * nq + iq];
                }
                l0 += v0; l1 += v1;
            }
        }
        double mi = a.massinv[nbase + tid];
        r0 = mi * r0; r1 = mi * r1; r2 = mi * r2;
        if (a.has_visc) { r1 = r1 + a.visc * mi * l0; r2 = r2 + a.visc * mi * l1; }
        if (a.rhs_only) {
            a.rhs_out[0][nbase + tid] = r0; a.rhs_out[1][nbase + tid] = r1; a.rhs_out[2][nbase + tid] = r2;
        } else {
            double q1[3] = {nod[1 * npts + tid], nod[2 * npts + tid], nod[3 * npts + tid]};
            double q0[3], q2[3] = {0.0, 0.0, 0.0};
            for (int v = 0; v < 3; ++v) q0[v] = a.load_q0 ? a.qb0[v][nbase + tid] : q1[v];
            if (a.load_q2) for (int v = 0; v < 3; ++v) q2[v] = a.qb2[v][nbase + tid];
            double rr[3] = {r0, r1, r2}, qn[3];
            for (int v = 0; v < 3; ++v) qn[v] = a.a1 * q0[v] + a.a2 * q1[v] + a.a3 * q2[v] + a.dtt * rr[v];
            // wall projection (btp_mom_boundary_df)
            for (int s = 0; s < 4; ++s) {
                bool on = (s == 0) ? (m == 0) : (s == 1) ? (m == ngl - 1) : (s == 2) ? (n == 0) : (n == ngl - 1);
                if (!on) continue;
                int slot = e * 4 + s, nb = a.M.nbr[slot];
                if (nb == NBR_FREESLIP) {
                    const FGeo fgn_ = fg_n(a.M, slot, s < 2 ? n : m);
                    double nx = fgn_.nx, ny = fgn_.ny;
                    double unl = qn[1] * nx + qn[2] * ny;
                    qn[1] = qn[1] - unl * nx; qn[2] = qn[2] - unl * ny;
                } else if (nb == NBR_NOSLIP) { qn[1] = 0.0; qn[2] = 0.0; }
            }
            if (a.store_q0) for (int v = 0; v < 3; ++v) a.qb0[v][nbase + tid] = q1[v];
            for (int v = 0; v < 3; ++v) a.qb[v][nbase + tid] = qn[v];
            if (a.store_q2) for (int v = 0; v < 3; ++v) a.qb2[v][nbase + tid] = qn[v];
            // stage the new state for the trace epilogue
            double pbn_ = qn[0] + a.pbprime_df[nbase + tid];
            nod[1 * npts + tid] = qn[0]; nod[2 * npts + tid] = qn[1]; nod[3 * npts + tid] = qn[2];
            nod[7 * npts + tid] = qn[1] / pbn_; nod[8 * npts + tid] = qn[2] / pbn_;
        }
    }
    if (a.rhs_only) return;
    __syncthreads();
    // ---- 9. publish the traces of the new state (and of its LDG gradient) for the next stage
    if (tid < 4 * ngl) {
        int s = tid / ngl, n = tid - s * ngl;
        int I = face_node(s, n, ngl);
        size_t base = ((size_t)e * 4 + s) * ngl + n;
        a.tr_out[TR_PBPERT * a.trstride + base] = nod[1 * npts + I];
        a.tr_out[TR_MX * a.trstride + base] = nod[2 * npts + I];
        a.tr_out[TR_MY * a.trstride + base] = nod[3 * npts + I];
        if (nodal_visc) {
            int m = I / ngl, nn = I - m * ngl;
            const Met mt = met_n(a.M, e, I);
            const double ksx = mt.ksx, ksy = mt.ksy, etx = mt.etx, ety = mt.ety;
            double dk, de;
            nodal_grad(o, ngl, nod + 7 * npts, nn, m, dk, de);
            a.tr_out[(TR_G + 0) * a.trstride + base] = ksx * dk + etx * de;
            a.tr_out[(TR_G + 1) * a.trstride + base] = ksy * dk + ety * de;
            nodal_grad(o, ngl, nod + 8 * npts, nn, m, dk, de);
            a.tr_out[(TR_G + 2) * a.trstride + base] = ksx * dk + etx * de;
            a.tr_out[(TR_G + 3) * a.trstride + base] = ksy * dk + ety * de;
        }
    }
    if (quad_visc) {   // face values of the flux variable of the new state
        for (int t = tid; t < 4 * nq; t += blockDim.x) {
            const int s = t / nq, iq = t - s * nq;
            double F[4];
            vq_face_point(a.M, e, o, ngl, nq, nod + 7 * npts, nod + 8 * npts, s, iq, a.vqP, a.vqS, qbase, F);
            for (int c = 0; c < 4; ++c) a.trq_out[c * a.trq_stride + ((size_t)e * 4 + s) * nq + iq] = F[c];
        }
    }
}

// traces of an existing state (start of the substep loop), also used for the nodal sums before k_btp_finalize.
// mode 0: planes are (pbpert,mx,my) and G is computed from them;  mode 1: planes are 7 nodal planes copied verbatim.
struct PrimeArgs {
    Mesh M;
    const double* in[7];
    const double* pbprime_df;
    double* tr_out;
    size_t trstride;
    int has_visc, mode;
};
__global__ void k_btp_prime_traces(PrimeArgs a) {
    extern __shared__ double sm[];
    const int ngl = a.M.ngl, nq = a.M.nq, npts = a.M.npts;
    const int e = blockIdx.x, tid = threadIdx.x;
    SOps o = load_sops(sm, ngl, nq);
    double* u = sm + sops_doubles(ngl, nq);
    double* v = u + npts;
    const size_t nbase = (size_t)e * npts;
    if (a.mode == 0 && a.has_visc && tid < npts) {
        double pb = a.in[0][nbase + tid] + a.pbprime_df[nbase + tid];
        u[tid] = a.in[1][nbase + tid] / pb; v[tid] = a.in[2][nbase + tid] / pb;
    }
    __syncthreads();
    if (tid < 4 * ngl) {
        int s = tid / ngl, n = tid - s * ngl, I = face_node(s, n, ngl);
        size_t base = ((size_t)e * 4 + s) * ngl + n;
        int nv = (a.mode == 0) ? 3 : 7;
        for (int k = 0; k < nv; ++k) a.tr_out[k * a.trstride + base] = a.in[k][nbase + I];
        if (a.mode == 0 && a.has_visc) {
            const Met mt = met_n(a.M, e, I);
            const double ksx = mt.ksx, ksy = mt.ksy, etx = mt.etx, ety = mt.ety;
            int m = I / ngl, nn = I - m * ngl;
            double dk, de;
            nodal_grad(o, ngl, u, nn, m, dk, de);
            a.tr_out[(TR_G + 0) * a.trstride + base] = ksx * dk + etx * de;
            a.tr_out[(TR_G + 1) * a.trstride + base] = ksy * dk + ety * de;
            nodal_grad(o, ngl, v, nn, m, dk, de);
            a.tr_out[(TR_G + 2) * a.trstride + base] = ksx * dk + etx * de;
            a.tr_out[(TR_G + 3) * a.trstride + base] = ksy * dk + ety * de;
        }
    }
}

// Reconstruct the full set of reference time averages (mod_rk_mlswe.F90:124-149) from the reduced sums.
// The sums are addressed as  acc_n[v][e*en + node],  acc_q[v][e*eq + point],  acc_f[v][slot*ef + iq]  and the traces
// of the nodal sums as  tr[v*tr_vs + rec*tr_rs + n]  so that both the plane layout (en = npts, eq = nq2, ef = nq,
// tr_vs = plane stride, tr_rs = ngl) and the record layout of stage_tma.cuh can be read.
struct FinalizeArgs {
    Mesh M;
    const double* acc_n[10];
    const double* acc_q[8];
    const double* acc_f[11];
    size_t en, eq, ef;
    const double* tr;  // traces of the nodal sums: 0 S_pbpert 1 S_mx 2 S_my
    size_t tr_vs, tr_rs;
    double* ave_q[12];
    double* ave_f[16];
    double* ave_n[7];
    const double *oop_q, *Hbcl, *Hbcl_e, *cL, *cR, *lam, *pbl, *pbr;
    const double *qp_dp, *qp_u, *qp_v;
    double S, N_inv, cd_over_g;
    int botfr;
    int derive_graduvb;  // 1: graduvb_ave = grad(uvb_ave_df) (the gradient is linear); 0: acc_n[6..9] hold the per-stage sums
};
template <int G_, int Q_>
__global__ void k_btp_finalize(FinalizeArgs a) {
    extern __shared__ double sm[];
    const int ngl = G_ ? G_ : a.M.ngl, nq = Q_ ? Q_ : a.M.nq, npts = ngl * ngl, nq2 = nq * nq, per = ngl * nq;
    const int e = blockIdx.x, tid = threadIdx.x;
    SOps o = load_sops(sm, ngl, nq);
    double* nod = sm + sops_doubles(ngl, nq);  // 0 S_pbpert 1 S_mx 2 S_my 3 pp 4 up 5 vp 6 ub_ave 7 vb_ave
    double* tmp = nod + 8 * npts;
    const size_t nbase = (size_t)e * npts, qbase = (size_t)e * nq2;
    const size_t ab_n = (size_t)e * a.en, ab_q = (size_t)e * a.eq;
    if (tid < npts) {
        nod[tid] = a.acc_n[3][ab_n + tid]; nod[npts + tid] = a.acc_n[4][ab_n + tid]; nod[2 * npts + tid] = a.acc_n[5][ab_n + tid];
        if (a.botfr == 1) {
            nod[3 * npts + tid] = a.qp_dp[nbase + tid]; nod[4 * npts + tid] = a.qp_u[nbase + tid]; nod[5 * npts + tid] = a.qp_v[nbase + tid];
        }
        double ua = a.N_inv * a.acc_n[1][ab_n + tid], va = a.N_inv * a.acc_n[2][ab_n + tid];
        nod[6 * npts + tid] = ua; nod[7 * npts + tid] = va;
        a.ave_n[0][nbase + tid] = a.N_inv * a.acc_n[0][ab_n + tid];
        a.ave_n[1][nbase + tid] = ua;
        a.ave_n[2][nbase + tid] = va;
        if (!a.derive_graduvb)
            for (int v = 0; v < 4; ++v) a.ave_n[3 + v][nbase + tid] = a.N_inv * a.acc_n[6 + v][ab_n + tid];
    }
    __syncthreads();
    if (a.derive_graduvb && tid < npts) {
        const Met mt = met_n(a.M, e, tid);
        const double ksx = mt.ksx, ksy = mt.ksy, etx = mt.etx, ety = mt.ety;
        int m = tid / ngl, n = tid - m * ngl;
        double dk, de;
        nodal_grad(o, ngl, nod + 6 * npts, n, m, dk, de);
        a.ave_n[3][nbase + tid] = ksx * dk + etx * de; a.ave_n[4][nbase + tid] = ksy * dk + ety * de;
        nodal_grad(o, ngl, nod + 7 * npts, n, m, dk, de);
        a.ave_n[5][nbase + tid] = ksx * dk + etx * de; a.ave_n[6][nbase + tid] = ksy * dk + ety * de;
    }
    sf_pass1(o, ngl, nq, a.botfr == 1 ? 6 : 3, nod, npts, tmp, nullptr);
    __syncthreads();
    if (tid < nq2) {
        int j = tid / nq, i = tid - j * nq;
        size_t Iq = qbase + tid, Aq = ab_q + tid;
        double sp = sf_eval(o, ngl, nq, tmp, 0, i, j), smx = sf_eval(o, ngl, nq, tmp, 1, i, j), smy = sf_eval(o, ngl, nq, tmp, 2, i, j);
        double ope2 = a.acc_q[3][Aq];
        a.ave_q[0][Iq] = a.N_inv * (a.S + sp * a.oop_q[Iq]);
        a.ave_q[1][Iq] = a.N_inv * (a.Hbcl[Iq] * ope2);
        a.ave_q[2][Iq] = a.N_inv * a.acc_q[0][Aq];
        a.ave_q[3][Iq] = a.N_inv * a.acc_q[1][Aq];
        a.ave_q[4][Iq] = a.N_inv * a.acc_q[2][Aq];
        a.ave_q[5][Iq] = a.N_inv * ope2;
        a.ave_q[6][Iq] = a.N_inv * smx;
        a.ave_q[7][Iq] = a.N_inv * smy;
        double ubs = a.acc_q[4][Aq], vbs = a.acc_q[5][Aq];
        a.ave_q[8][Iq] = a.N_inv * ubs;
        a.ave_q[9][Iq] = a.N_inv * vbs;
        double tbx = 0.0, tby = 0.0;
        if (a.botfr == 1) {
            double pp = sf_eval(o, ngl, nq, tmp, 3, i, j), up = sf_eval(o, ngl, nq, tmp, 4, i, j), vp = sf_eval(o, ngl, nq, tmp, 5, i, j);
            double spd = a.cd_over_g * pp;
            tbx = spd * (a.S * up + ubs); tby = spd * (a.S * vp + vbs);
        } else if (a.botfr == 2) { tbx = a.acc_q[6][Aq]; tby = a.acc_q[7][Aq]; }
        a.ave_q[10][Iq] = a.N_inv * tbx;
        a.ave_q[11][Iq] = a.N_inv * tby;
    }
    // faces owned by this element
    if (tid < 4 * nq) {
        int s = tid / nq, iq = tid - s * nq;
        int slot = e * 4 + s, nb = a.M.nbr[slot], nbs = a.M.nbslot[slot];
        bool left = (nb < 0) || (e < nb);
        if (left) {
            const FGeo fgq_ = fg_q(a.M, slot, iq);
            double nxl = fgq_.nx, nyl = fgq_.ny;
            double sl[3] = {0, 0, 0}, sr[3] = {0, 0, 0};
            for (int n = 0; n < ngl; ++n) {
                double hi = o.A[n + ngl * iq];
                int I = face_node(s, n, ngl);
                double ow[3] = {nod[I], nod[npts + I], nod[2 * npts + I]}, nv[3];
                if (nb >= 0 || nb == NBR_HALO) {
                    size_t rec = (nb >= 0) ? ((size_t)nb * 4 + nbs) : ((size_t)a.M.nslots + nbs);
                    for (int v = 0; v < 3; ++v) nv[v] = a.tr[v * a.tr_vs + rec * a.tr_rs + n];
                } else {
                    nv[0] = ow[0]; nv[1] = ow[1]; nv[2] = ow[2];
                    if (nb == NBR_FREESLIP) {   // the ghost is built at the face nodes with the nodal normal, then interpolated
                        const FGeo fgn_ = fg_n(a.M, slot, n);
                        double un = fgn_.nx * ow[1] + fgn_.ny * ow[2]; nv[1] = ow[1] - 2.0 * un * fgn_.nx; nv[2] = ow[2] - 2.0 * un * fgn_.ny;
                    }
                    else if (nb == NBR_NOSLIP) { nv[1] = -ow[1]; nv[2] = -ow[2]; }
                }
                for (int v = 0; v < 3; ++v) { sl[v] += hi * ow[v]; sr[v] += hi * nv[v]; }
            }
            size_t fo = (size_t)slot * nq + iq, fa = (size_t)slot * a.ef + iq;
            double cL = a.cL[fo], cR = a.cR[fo], lam = a.lam[fo];
            a.ave_f[0][fo] = a.N_inv * (cR * sl[1] + cL * sr[1] + lam * (nxl * sl[0] - nxl * sr[0]));
            a.ave_f[1][fo] = a.N_inv * (cR * sl[2] + cL * sr[2] + lam * (nyl * sl[0] - nyl * sr[0]));
            double e2 = a.acc_f[6][fa];
            a.ave_f[2][fo] = a.N_inv * (a.Hbcl_e[fo] * e2);
            for (int v = 0; v < 4; ++v) a.ave_f[3 + v][fo] = a.N_inv * a.acc_f[v][fa];
            a.ave_f[7][fo] = a.N_inv * (a.S + sl[0] / a.pbl[fo]);
            a.ave_f[8][fo] = a.N_inv * (a.S + sr[0] / a.pbr[fo]);
            a.ave_f[9][fo] = a.N_inv * a.acc_f[4][fa];
            a.ave_f[10][fo] = a.N_inv * a.acc_f[5][fa];
            a.ave_f[11][fo] = a.N_inv * e2;
            for (int v = 0; v < 4; ++v) a.ave_f[12 + v][fo] = a.N_inv * a.acc_f[7 + v][fa];
        }
    }
}

}  // namespace hn
