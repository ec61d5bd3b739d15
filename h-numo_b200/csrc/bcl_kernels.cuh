// bcl_kernels.cuh -- baroclinic (layer) kernels of one ti_rk_bcl step.
//
// Replaces (reference file:line)
//   btp_bcl_coeffs_qdf                 src/mod_barotropic_terms.F90:219-409
//   layer_mass_rhs                     src/mod_create_rhs_mlswe.F90:53-78,822-877,922-1034
//   apply_consistency                  src/mod_splitting.F90:324-366, src/mod_layer_terms.F90:57-137,
//                                      src/mod_create_rhs_mlswe.F90:80-101,879-920,1036-1115
//   bcl_create_laplacian               src/mod_laplacian_quad.F90:227-248,392-425,521-611
//   layer_momentum_rhs                 src/mod_create_rhs_mlswe.F90:28-51,281-456,458-820
//   momentum / momentum_mass updates   src/mod_splitting.F90:94-180,182-287
//   layer_mom_boundary_df, evaluate_bcl(_v1), extract_velocity   src/mod_layer_terms.F90:198-320,529-584
//
// Face arrays of the reference (qprime_df_face, graduv_dpp_face, graduvb_face_ave, mass_deficit_mass_face) are never
// materialised: every kernel gathers the neighbour's nodal values at the shared face nodes (or the ghost state, or
// the halo copy on processor boundaries) and evaluates the face from the canonical left-element perspective.
// One thread block per element ("simple" variant; these kernels run 2x per baroclinic step, < 6 % of the work).
#pragma once
#include "hnumo_dev.cuh"

namespace hn {

// neighbour's nodal value of a plane at face node n of slot (e,s); `own` is returned for walls.
__device__ __forceinline__ double nb_nodal(const Mesh& M, const double* plane, const double* hplane, int nb, int nbs, int n,
                                           double own) {
    if (nb >= 0) return plane[(size_t)nb * M.npts + face_node(nbs, n, M.ngl)];
    if (nb == NBR_HALO) return hplane[(size_t)nbs * M.ngl + n];
    return own;
}

// --------------------------------------------------------------------------------------------------------------
struct CoeffArgs {
    Mesh M;
    const double* qprime;  // [3*nl] nodal planes, index v*nl+k
    const double* dpv;     // [nl]
    size_t nstride;
    const double* hq;      // halo copy of qprime traces [3*nl][nhalo*ngl]
    size_t hstride;
    double *Quu, *Quv, *Qvv, *Hbcl, *Quu_e, *Quv_e, *Qvv_e, *Hbcl_e;
    double* dpp_graduv;      // [4*nl] nodal planes
    double* btp_dpp_graduv;  // [4]
    double* pbprime_visc;
    double alpha[HN_MAXL];
    int has_visc;
};
template <int G_, int Q_, int NL_>
__global__ void k_bcl_coeffs(CoeffArgs a) {
    extern __shared__ double sm[];
    const int ngl = G_ ? G_ : a.M.ngl, nq = Q_ ? Q_ : a.M.nq, npts = ngl * ngl, nq2 = nq * nq, nl = NL_ ? NL_ : a.M.nl, per = ngl * nq;
    constexpr int LMAX = NL_ ? NL_ : HN_MAXL;   // per-thread layer arrays stay in registers when the layer count is a compile-time constant
    const int e = blockIdx.x, tid = threadIdx.x;
    SOps o = load_sops(sm, ngl, nq);
    double* nod = sm + sops_doubles(ngl, nq);  // [3][npts]
    double* tmp = nod + 3 * npts;              // [3][per]
    double* nbq = tmp + 3 * per;               // [4][3][ngl]
    const size_t nbase = (size_t)e * npts, qbase = (size_t)e * nq2;
    double Quu = 0, Quv = 0, Qvv = 0, H = 0, pprime = 0;             // quad thread accumulators
    double eu = 0, euv = 0, ev = 0, eH = 0, ppl = 0, ppr = 0;         // face-quad thread accumulators
    double bsum[4] = {0, 0, 0, 0}, pvs = 0;                           // nodal thread accumulators
    // metric terms of this thread's quadrature point (phases `tid < nq2`) and of its node (phases `tid < npts`); per element on affine meshes
    const Met mq_ = met_q(a.M, e, tid < nq2 ? tid : 0), mn_ = met_n(a.M, e, tid < npts ? tid : 0);
    for (int k = 0; k < nl; ++k) {
        __syncthreads();
        if (tid < npts)
            for (int v = 0; v < 3; ++v) nod[v * npts + tid] = a.qprime[(size_t)(v * nl + k) * a.nstride + nbase + tid];
        __syncthreads();
        if (tid < 4 * ngl) {
            int s = tid / ngl, n = tid - s * ngl, slot = e * 4 + s, nb = a.M.nbr[slot], nbs = a.M.nbslot[slot];
            int I = face_node(s, n, ngl);
            double ow[3] = {nod[I], nod[npts + I], nod[2 * npts + I]}, nv[3];
            for (int v = 0; v < 3; ++v)
                nv[v] = nb_nodal(a.M, a.qprime + (size_t)(v * nl + k) * a.nstride, a.hq + (size_t)(v * nl + k) * a.hstride, nb, nbs, n, ow[v]);
            if (nb == NBR_FREESLIP) {
                const FGeo fgn_ = fg_n(a.M, slot, n);
                double nx = fgn_.nx, ny = fgn_.ny;
                double un = ow[1] * nx + ow[2] * ny;
                nv[1] = ow[1] - 2.0 * un * nx; nv[2] = ow[2] - 2.0 * un * ny;
            } else if (nb == NBR_NOSLIP) { nv[1] = -ow[1]; nv[2] = -ow[2]; }
            for (int v = 0; v < 3; ++v) nbq[(s * 3 + v) * ngl + n] = nv[v];
        }
        sf_pass1(o, ngl, nq, 3, nod, npts, tmp, nullptr);
        __syncthreads();
        if (tid < nq2) {
            int j = tid / nq, i = tid - j * nq;
            double q0 = sf_eval(o, ngl, nq, tmp, 0, i, j), q1 = sf_eval(o, ngl, nq, tmp, 1, i, j), q2 = sf_eval(o, ngl, nq, tmp, 2, i, j);
            Quu = Quu + q1 * (q1 * q0);
            Quv = Quv + q2 * (q1 * q0);
            Qvv = Qvv + q2 * (q2 * q0);
            double pn = pprime + q0;
            H = H + 0.5 * a.alpha[k] * (pn * pn - pprime * pprime);
            pprime = pn;
        }
        // owned faces: edge coefficients from left/right traces (mod_barotropic_terms.F90:306-337)
        if (tid < 4 * nq) {
            int s = tid / nq, iq = tid - s * nq, slot = e * 4 + s, nb = a.M.nbr[slot];
            bool left = (nb < 0) || (e < nb);
            if (left) {
                double ql[3] = {0, 0, 0}, qr[3] = {0, 0, 0};
                for (int n = 0; n < ngl; ++n) {
                    double hi = o.A[n + ngl * iq];
                    int I = face_node(s, n, ngl);
                    for (int v = 0; v < 3; ++v) { ql[v] += hi * nod[v * npts + I]; qr[v] += hi * nbq[(s * 3 + v) * ngl + n]; }
                }
                eu = eu + 0.5 * ((ql[1] * ql[1] * ql[0]) + (qr[1] * qr[1] * qr[0]));
                euv = euv + 0.5 * ((ql[2] * ql[1] * ql[0]) + (qr[2] * qr[1] * qr[0]));
                ev = ev + 0.5 * ((ql[2] * ql[2] * ql[0]) + (qr[2] * qr[2] * qr[0]));
                double pl = ppl + ql[0], pr = ppr + qr[0];
                double left_dp = 0.5 * a.alpha[k] * (pl * pl - ppl * ppl);
                double right_dp = 0.5 * a.alpha[k] * (pr * pr - ppr * ppr);
                eH = eH + 0.5 * (left_dp + right_dp);
                ppl = pl; ppr = pr;
            }
        }
        // viscosity auxiliaries at the nodes (mod_barotropic_terms.F90:287-304)
        if (a.has_visc && tid < npts) {
            int m = tid / ngl, n = tid - m * ngl;
            double dk, de, gv[4];
            nodal_grad(o, ngl, nod + npts, n, m, dk, de);
            gv[0] = mn_.ksx * dk + mn_.etx * de; gv[1] = mn_.ksy * dk + mn_.ety * de;
            nodal_grad(o, ngl, nod + 2 * npts, n, m, dk, de);
            gv[2] = mn_.ksx * dk + mn_.etx * de; gv[3] = mn_.ksy * dk + mn_.ety * de;
            double d = a.dpv[(size_t)k * a.nstride + nbase + tid];
            for (int v = 0; v < 4; ++v) {
                double t = d * gv[v];
                a.dpp_graduv[(size_t)(v * nl + k) * a.nstride + nbase + tid] = t;
                bsum[v] = bsum[v] + t;
            }
            pvs = pvs + d;
        }
    }
    if (tid < nq2) { a.Quu[qbase + tid] = Quu; a.Quv[qbase + tid] = Quv; a.Qvv[qbase + tid] = Qvv; a.Hbcl[qbase + tid] = H; }
    if (tid < 4 * nq) {
        int s = tid / nq, iq = tid - s * nq, slot = e * 4 + s, nb = a.M.nbr[slot];
        if ((nb < 0) || (e < nb)) {
            size_t fo = (size_t)slot * nq + iq;
            a.Quu_e[fo] = eu; a.Quv_e[fo] = euv; a.Qvv_e[fo] = ev; a.Hbcl_e[fo] = eH;
        }
    }
    if (a.has_visc && tid < npts) {
        for (int v = 0; v < 4; ++v) a.btp_dpp_graduv[(size_t)v * a.nstride + nbase + tid] = bsum[v];
        a.pbprime_visc[nbase + tid] = pvs;
    }
}

// --------------------------------------------------------------------------------------------------------------
struct MassArgs {
    Mesh M;
    const double* qprime; size_t nstride;
    const double* hq; size_t hstride;
    const double* ave_q[12];
    const double* ave_f[16];
    const double* qdp_in; // [nl] nodal planes: q_df(1,:,k) before the update
    double* qdp;          // [nl] nodal planes: q_df(1,:,k) after it (may alias qdp_in)
    double* slmf_q[2];
    double* slmf_f[2];
    const double* massinv;
    int* flag;
    double dt;
};
template <int G_, int Q_, int NL_>
__global__ void k_layer_mass(MassArgs a) {
    extern __shared__ double sm[];
    const int ngl = G_ ? G_ : a.M.ngl, nq = Q_ ? Q_ : a.M.nq, npts = ngl * ngl, nq2 = nq * nq, nl = NL_ ? NL_ : a.M.nl, per = ngl * nq;
    constexpr int LMAX = NL_ ? NL_ : HN_MAXL;   // per-thread layer arrays stay in registers when the layer count is a compile-time constant
    const int e = blockIdx.x, tid = threadIdx.x;
    SOps o = load_sops(sm, ngl, nq);
    double* nod = sm + sops_doubles(ngl, nq);
    double* tmp = nod + 3 * npts;
    double* nbq = tmp + 3 * per;       // [4][3][ngl]
    double* fq = nbq + 12 * ngl;       // [2][nq2]
    double* tP = fq + 2 * nq2;
    double* tR = tP + per;
    double* adv = tR + per;            // [npts]
    double* ff = adv + npts;           // [4][nq]
    const size_t nbase = (size_t)e * npts, qbase = (size_t)e * nq2;
    // metric terms of this thread's quadrature point (phases `tid < nq2`) and of its node (phases `tid < npts`); per element on affine meshes
    const Met mq_ = met_q(a.M, e, tid < nq2 ? tid : 0), mn_ = met_n(a.M, e, tid < npts ? tid : 0);
    double sx = 0, sy = 0, sfx = 0, sfy = 0;
    for (int k = 0; k < nl; ++k) {
        __syncthreads();
        if (tid < npts)
            for (int v = 0; v < 3; ++v) nod[v * npts + tid] = a.qprime[(size_t)(v * nl + k) * a.nstride + nbase + tid];
        __syncthreads();
        if (tid < 4 * ngl) {
            int s = tid / ngl, n = tid - s * ngl, slot = e * 4 + s, nb = a.M.nbr[slot], nbs = a.M.nbslot[slot];
            int I = face_node(s, n, ngl);
            double ow[3] = {nod[I], nod[npts + I], nod[2 * npts + I]}, nv[3];
            for (int v = 0; v < 3; ++v)
                nv[v] = nb_nodal(a.M, a.qprime + (size_t)(v * nl + k) * a.nstride, a.hq + (size_t)(v * nl + k) * a.hstride, nb, nbs, n, ow[v]);
            if (nb == NBR_FREESLIP) {
                const FGeo fgn_ = fg_n(a.M, slot, n);
                double nx = fgn_.nx, ny = fgn_.ny;
                double un = ow[1] * nx + ow[2] * ny;
                nv[1] = ow[1] - 2.0 * un * nx; nv[2] = ow[2] - 2.0 * un * ny;
            } else if (nb == NBR_NOSLIP) { nv[1] = -ow[1]; nv[2] = -ow[2]; }
            for (int v = 0; v < 3; ++v) nbq[(s * 3 + v) * ngl + n] = nv[v];
        }
        sf_pass1(o, ngl, nq, 3, nod, npts, tmp, nullptr);
        __syncthreads();
        if (tid < nq2) {
            int j = tid / nq, i = tid - j * nq;
            size_t Iq = qbase + tid;
            double q0 = sf_eval(o, ngl, nq, tmp, 0, i, j), q1 = sf_eval(o, ngl, nq, tmp, 1, i, j), q2 = sf_eval(o, ngl, nq, tmp, 2, i, j);
            double dp_temp = q0 * a.ave_q[0][Iq];
            double udp = (q1 + a.ave_q[8][Iq]) * dp_temp;
            double vdp = (q2 + a.ave_q[9][Iq]) * dp_temp;
            sx = sx + udp; sy = sy + vdp;
            double wq = o.wq[i] * o.wq[j] * mq_.J;
            fq[tid] = wq * (mq_.ksx * udp + mq_.ksy * vdp);
            fq[nq2 + tid] = wq * (mq_.etx * udp + mq_.ety * vdp);
        }
        if (tid < 4 * nq) {
            int s = tid / nq, iq = tid - s * nq, slot = e * 4 + s, nb = a.M.nbr[slot], nbs = a.M.nbslot[slot];
            bool left = (nb < 0) || (e < nb);
            int oslot = left ? slot : nb * 4 + nbs;
            size_t fo = (size_t)oslot * nq + iq;
            const FGeo fgq_ = fg_q(a.M, slot, iq);
            double nxl = fgq_.nx, nyl = fgq_.ny, nlen = fgq_.len;
            double qo[3] = {0, 0, 0}, qn[3] = {0, 0, 0};
            for (int n = 0; n < ngl; ++n) {
                double hi = o.A[n + ngl * iq];
                int I = face_node(s, n, ngl);
                for (int v = 0; v < 3; ++v) { qo[v] += hi * nod[v * npts + I]; qn[v] += hi * nbq[(s * 3 + v) * ngl + n]; }
            }
            const double* ql = left ? qo : qn;
            const double* qr = left ? qn : qo;
            double uu = 0.5 * ((ql[1] + a.ave_f[12][fo]) + (qr[1] + a.ave_f[13][fo]));
            double vv = 0.5 * ((ql[2] + a.ave_f[14][fo]) + (qr[2] + a.ave_f[15][fo]));
            double dpl = a.ave_f[7][fo] * ql[0], dpr = a.ave_f[8][fo] * qr[0];
            double fu = (uu * nxl > 0.0) ? uu * dpl : uu * dpr;
            double fv = (vv * nyl > 0.0) ? vv * dpl : vv * dpr;
            if (left) { sfx = sfx + fu; sfy = sfy + fv; }
            double flux = nxl * fu + nyl * fv;
            ff[s * nq + iq] = (left ? -1.0 : 1.0) * (o.wq[iq] * nlen) * flux;
        }
        __syncthreads();
        sf_scatter(o, ngl, nq, 1, nullptr, fq, fq + nq2, nq2, tP, tR, adv, npts, false);
        if (tid < npts) {
            int m = tid / ngl, n = tid - m * ngl;
            double r = adv[tid];
            for (int s = 0; s < 4; ++s) {
                int nf;
                if (s == 0) { if (m != 0) continue; nf = n; }
                else if (s == 1) { if (m != ngl - 1) continue; nf = n; }
                else if (s == 2) { if (n != 0) continue; nf = m; }
                else { if (n != ngl - 1) continue; nf = m; }
                double p = 0.0;
                for (int iq = 0; iq < nq; ++iq) p += o.A[nf + ngl * iq] * ff[s * nq + iq];
                r += p;
            }
            double dpa = a.massinv[nbase + tid] * r;
            const size_t Iqd = (size_t)k * a.nstride + nbase + tid;
            double v = a.qdp_in[Iqd] + a.dt * dpa;
            a.qdp[Iqd] = v;
            if (v < 0.0) *a.flag = 1;
        }
    }
    if (tid < nq2) { a.slmf_q[0][qbase + tid] = sx; a.slmf_q[1][qbase + tid] = sy; }
    if (tid < 4 * nq) {
        int s = tid / nq, iq = tid - s * nq, slot = e * 4 + s, nb = a.M.nbr[slot];
        if ((nb < 0) || (e < nb)) { a.slmf_f[0][(size_t)slot * nq + iq] = sfx; a.slmf_f[1][(size_t)slot * nq + iq] = sfy; }
    }
}

// --------------------------------------------------------------------------------------------------------------
struct ConsArgs {
    Mesh M;
    const double* qdp_in;   // [nl] nodal planes (after the advective update)
    double* qdp_out;        // [nl]
    size_t nstride;
    const double* hdp;      // halo copy of qdp_in traces [nl][nhalo*ngl]
    size_t hstride;
    const double *pbprime_df, *pbn, *pbprime_q, *pbf_l, *pbf_r, *massinv;
    const double* ave_q[12];
    const double* ave_f[16];
    const double* slmf_q[2];
    const double* slmf_f[2];
    double dt;
};
template <int G_, int Q_, int NL_>
__global__ void k_consistency(ConsArgs a) {
    extern __shared__ double sm[];
    const int ngl = G_ ? G_ : a.M.ngl, nq = Q_ ? Q_ : a.M.nq, npts = ngl * ngl, nq2 = nq * nq, nl = NL_ ? NL_ : a.M.nl, per = ngl * nq;
    constexpr int LMAX = NL_ ? NL_ : HN_MAXL;   // per-thread layer arrays stay in registers when the layer count is a compile-time constant
    const int e = blockIdx.x, tid = threadIdx.x;
    SOps o = load_sops(sm, ngl, nq);
    double* dpp = sm + sops_doubles(ngl, nq);  // [nl][npts] own dpprime_df
    double* nbd = dpp + nl * npts;             // [4][nl][ngl] neighbour dpprime_df at the face nodes
    double* tmp = nbd + 4 * nl * ngl;          // [per]
    double* fq = tmp + per;                    // [2][nq2]
    double* tP = fq + 2 * nq2;
    double* tR = tP + per;
    double* adv = tR + per;
    double* ff = adv + npts;                   // [4][nq]
    const size_t nbase = (size_t)e * npts, qbase = (size_t)e * nq2;
    // metric terms of this thread's quadrature point (phases `tid < nq2`) and of its node (phases `tid < npts`); per element on affine meshes
    const Met mq_ = met_q(a.M, e, tid < nq2 ? tid : 0), mn_ = met_n(a.M, e, tid < npts ? tid : 0);
    // dpprime_df = q_df(1)/ (sum_k q_df(1) / pbprime_df)      (mod_splitting.F90:350-353)
    if (tid < npts) {
        double s = 0.0;
        for (int k = 0; k < nl; ++k) s += a.qdp_in[(size_t)k * a.nstride + nbase + tid];
        double ope = s / a.pbprime_df[nbase + tid];
        for (int k = 0; k < nl; ++k) dpp[k * npts + tid] = a.qdp_in[(size_t)k * a.nstride + nbase + tid] / ope;
    }
    __syncthreads();
    if (tid < 4 * ngl) {
        int s = tid / ngl, n = tid - s * ngl, slot = e * 4 + s, nb = a.M.nbr[slot], nbs = a.M.nbslot[slot];
        int I = face_node(s, n, ngl);
        if (nb >= 0 || nb == NBR_HALO) {
            double sum = 0.0, vals[LMAX];
            for (int k = 0; k < nl; ++k) {
                vals[k] = nb_nodal(a.M, a.qdp_in + (size_t)k * a.nstride, a.hdp + (size_t)k * a.hstride, nb, nbs, n, 0.0);
                sum += vals[k];
            }
            double ope = sum / a.pbn[(size_t)slot * ngl + n];
            for (int k = 0; k < nl; ++k) nbd[(s * nl + k) * ngl + n] = vals[k] / ope;
        } else {
            for (int k = 0; k < nl; ++k) nbd[(s * nl + k) * ngl + n] = dpp[k * npts + I];
        }
    }
    for (int k = 0; k < nl; ++k) {
        __syncthreads();
        sf_pass1(o, ngl, nq, 1, dpp + k * npts, npts, tmp, nullptr);
        __syncthreads();
        if (tid < nq2) {
            int j = tid / nq, i = tid - j * nq;
            size_t Iq = qbase + tid;
            double dp = sf_eval(o, ngl, nq, tmp, 0, i, j);
            double weight = dp / a.pbprime_q[Iq];
            double udp = weight * (a.ave_q[6][Iq] - a.slmf_q[0][Iq]);
            double vdp = weight * (a.ave_q[7][Iq] - a.slmf_q[1][Iq]);
            double wq = o.wq[i] * o.wq[j] * mq_.J;
            fq[tid] = wq * (mq_.ksx * udp + mq_.ksy * vdp);
            fq[nq2 + tid] = wq * (mq_.etx * udp + mq_.ety * vdp);
        }
        if (tid < 4 * nq) {
            int s = tid / nq, iq = tid - s * nq, slot = e * 4 + s, nb = a.M.nbr[slot], nbs = a.M.nbslot[slot];
            bool left = (nb < 0) || (e < nb);
            int oslot = left ? slot : nb * 4 + nbs;
            size_t fo = (size_t)oslot * nq + iq;
            const FGeo fgq_ = fg_q(a.M, slot, iq);
            double nxl = fgq_.nx, nyl = fgq_.ny, nlen = fgq_.len;
            double qo = 0.0, qn = 0.0;
            for (int n = 0; n < ngl; ++n) {
                double hi = o.A[n + ngl * iq];
                qo += hi * dpp[k * npts + face_node(s, n, ngl)];
                qn += hi * nbd[(s * nl + k) * ngl + n];
            }
            double qprime_l = left ? qo : qn, qprime_r = left ? qn : qo;
            double wl = qprime_l / a.pbf_l[fo], wr = qprime_r / a.pbf_r[fo];
            double d0 = a.ave_f[0][fo] - a.slmf_f[0][fo], d1 = a.ave_f[1][fo] - a.slmf_f[1][fo];
            double fu = ((wl * d0) * nxl > 0.0) ? wl * d0 : wr * d0;
            double fv = ((wl * d1) * nyl > 0.0) ? wl * d1 : wr * d1;
            double flux = nxl * fu + nyl * fv;
            ff[s * nq + iq] = (left ? -1.0 : 1.0) * (o.wq[iq] * nlen) * flux;
        }
        __syncthreads();
        sf_scatter(o, ngl, nq, 1, nullptr, fq, fq + nq2, nq2, tP, tR, adv, npts, false);
        if (tid < npts) {
            int m = tid / ngl, n = tid - m * ngl;
            double r = adv[tid];
            for (int s = 0; s < 4; ++s) {
                int nf;
                if (s == 0) { if (m != 0) continue; nf = n; }
                else if (s == 1) { if (m != ngl - 1) continue; nf = n; }
                else if (s == 2) { if (n != 0) continue; nf = m; }
                else { if (n != ngl - 1) continue; nf = m; }
                double p = 0.0;
                for (int iq = 0; iq < nq; ++iq) p += o.A[nf + ngl * iq] * ff[s * nq + iq];
                r += p;
            }
            size_t I = (size_t)k * a.nstride + nbase + tid;
            a.qdp_out[I] = a.qdp_in[I] + a.dt * a.massinv[nbase + tid] * r;
        }
    }
}

// --------------------------------------------------------------------------------------------------------------
// bcl_create_laplacian -> rhs_visc planes [2*nl]
struct LapArgs {
    Mesh M;
    const double* dpv; const double* dpp_graduv; size_t nstride;
    const double* graduvb[4];           // ave_n[3..6]
    const double* h_dpv;                // halo [nl][nhalo*ngl]
    const double* h_dpg;                // halo [4*nl][nhalo*ngl]
    const double* h_gub;                // halo [4][nhalo*ngl]
    size_t hstride;
    const double* massinv;
    double* rhs_visc;                   // [2*nl]
    double visc;
};
template <int G_, int Q_, int NL_>
__global__ void k_bcl_laplacian(LapArgs a) {
    extern __shared__ double sm[];
    const int ngl = G_ ? G_ : a.M.ngl, nq = Q_ ? Q_ : a.M.nq, npts = ngl * ngl, nl = NL_ ? NL_ : a.M.nl;
    constexpr int LMAX = NL_ ? NL_ : HN_MAXL;
    const int e = blockIdx.x, tid = threadIdx.x;
    SOps o = load_sops(sm, ngl, nq);
    double* gub = sm + sops_doubles(ngl, nq);  // [4][npts]
    double* qq = gub + 4 * npts;               // [4][npts]
    double* lf = qq + 4 * npts;                // [4][2][ngl]
    const size_t nbase = (size_t)e * npts;
    if (tid < npts) for (int v = 0; v < 4; ++v) gub[v * npts + tid] = a.graduvb[v][nbase + tid];
    for (int k = 0; k < nl; ++k) {
        __syncthreads();
        if (tid < npts) {
            double d = a.dpv[(size_t)k * a.nstride + nbase + tid];
            for (int v = 0; v < 4; ++v) qq[v * npts + tid] = d * gub[v * npts + tid] + a.dpp_graduv[(size_t)(v * nl + k) * a.nstride + nbase + tid];
        }
        // face flux at the face nodes (mod_laplacian_quad.F90:521-611)
        if (tid >= npts && tid < npts + 4 * ngl) {
            int t = tid - npts, s = t / ngl, n = t - s * ngl, slot = e * 4 + s, nb = a.M.nbr[slot], nbs = a.M.nbslot[slot];
            bool left = (nb < 0) || (e < nb);
            const FGeo fgn_ = fg_n(a.M, slot, n);
            double nx = fgn_.nx, ny = fgn_.ny, nlen = fgn_.len;
            int I = face_node(s, n, ngl);
            double go[4], gn[4], so[5], sn[5];
            for (int v = 0; v < 4; ++v) {
                go[v] = a.graduvb[v][nbase + I];
                so[v] = a.dpp_graduv[(size_t)(v * nl + k) * a.nstride + nbase + I];
            }
            so[4] = a.dpv[(size_t)k * a.nstride + nbase + I];
            if (nb >= 0 || nb == NBR_HALO) {
                for (int v = 0; v < 4; ++v) {
                    gn[v] = nb_nodal(a.M, a.graduvb[v], a.h_gub + (size_t)v * a.hstride, nb, nbs, n, 0.0);
                    sn[v] = nb_nodal(a.M, a.dpp_graduv + (size_t)(v * nl + k) * a.nstride, a.h_dpg + (size_t)(v * nl + k) * a.hstride, nb, nbs, n, 0.0);
                }
                sn[4] = nb_nodal(a.M, a.dpv + (size_t)k * a.nstride, a.h_dpv + (size_t)k * a.hstride, nb, nbs, n, 0.0);
            } else {
                for (int v = 0; v < 4; ++v) { gn[v] = go[v]; sn[v] = so[v]; }
                sn[4] = so[4];
                if (nb == NBR_FREESLIP) {
                    double un = go[0] * nx + go[1] * ny; gn[0] = go[0] - 2.0 * un * nx; gn[1] = go[1] - 2.0 * un * ny;
                    un = go[2] * nx + go[3] * ny; gn[2] = go[2] - 2.0 * un * nx; gn[3] = go[3] - 2.0 * un * ny;
                    un = so[0] * nx + so[1] * ny; sn[0] = so[0] - 2.0 * un * nx; sn[1] = so[1] - 2.0 * un * ny;
                    un = so[2] * nx + so[3] * ny; sn[2] = so[2] - 2.0 * un * nx; sn[3] = so[3] - 2.0 * un * ny;
                }
            }
            const double *gl = left ? go : gn, *gr = left ? gn : go, *sl = left ? so : sn, *sr = left ? sn : so;
            double fl[4], fr[4];
            for (int v = 0; v < 4; ++v) { fl[v] = sl[4] * gl[v] + sl[v]; fr[v] = sr[4] * gr[v] + sr[v]; }
            double qu0 = 0.5 * fl[0] + 0.5 * fr[0], qu1 = 0.5 * fl[1] + 0.5 * fr[1];
            double qv0 = 0.5 * fl[2] + 0.5 * fr[2], qv1 = 0.5 * fl[3] + 0.5 * fr[3];
            double wq = o.wg[n] * nlen;
            double flux_qu = (qu0 - fl[0] * nx) + (qu1 - fl[1] * ny);
            double flux_qv = (qv0 - fl[2] * nx) + (qv1 - fl[3] * ny);
            double sgn = left ? 1.0 : -1.0;
            lf[(s * 2 + 0) * ngl + n] = sgn * wq * flux_qu;
            lf[(s * 2 + 1) * ngl + n] = sgn * wq * flux_qv;
        }
        __syncthreads();
        if (tid < npts) {
            int m = tid / ngl, n = tid - m * ngl;
            double l0 = 0.0, l1 = 0.0;
            for (int kk = 0; kk < ngl; ++kk) {
                int I1 = m * ngl + kk, I2 = kk * ngl + n;
                const Met m1 = met_n(a.M, e, I1), m2 = met_n(a.M, e, I2);   // metric terms at the collocation point of the sum
                double wk1 = o.wg[kk] * o.wg[m] * m1.J * o.D[n + ngl * kk];
                double wk2 = o.wg[n] * o.wg[kk] * m2.J * o.D[m + ngl * kk];
                l0 -= wk1 * (m1.ksx * qq[I1] + m1.ksy * qq[npts + I1]) + wk2 * (m2.etx * qq[I2] + m2.ety * qq[npts + I2]);
                l1 -= wk1 * (m1.ksx * qq[2 * npts + I1] + m1.ksy * qq[3 * npts + I1]) + wk2 * (m2.etx * qq[2 * npts + I2] + m2.ety * qq[3 * npts + I2]);
            }
            if (m == 0) { l0 += lf[(0 * 2 + 0) * ngl + n]; l1 += lf[(0 * 2 + 1) * ngl + n]; }
            if (m == ngl - 1) { l0 += lf[(1 * 2 + 0) * ngl + n]; l1 += lf[(1 * 2 + 1) * ngl + n]; }
            if (n == 0) { l0 += lf[(2 * 2 + 0) * ngl + m]; l1 += lf[(2 * 2 + 1) * ngl + m]; }
            if (n == ngl - 1) { l0 += lf[(3 * 2 + 0) * ngl + m]; l1 += lf[(3 * 2 + 1) * ngl + m]; }
            double mi = a.massinv[nbase + tid];
            a.rhs_visc[(size_t)(0 * nl + k) * a.nstride + nbase + tid] = a.visc * mi * l0;
            a.rhs_visc[(size_t)(1 * nl + k) * a.nstride + nbase + tid] = a.visc * mi * l1;
        }
    }
}

// --------------------------------------------------------------------------------------------------------------
// create_rhs_dynamics_volume_layers -> rhs_mom planes [2*nl] (volume part, not yet multiplied by massinv)
struct MomVolArgs {
    Mesh M;
    const double* qprime; const double* q;  // [3*nl] each
    size_t nstride;
    const double* ave_q[12];
    const double* ope2_df;  // ave_n[0]
    const double *zbot_df, *tauwx_q, *tauwy_q, *pbprime_q;
    double* rhs_mom;        // [2*nl]
    double alpha[HN_MAXL];
    double g;
};
template <int G_, int Q_, int NL_>
__global__ void k_mom_volume(MomVolArgs a) {
    extern __shared__ double sm[];
    const int ngl = G_ ? G_ : a.M.ngl, nq = Q_ ? Q_ : a.M.nq, npts = ngl * ngl, nq2 = nq * nq, nl = NL_ ? NL_ : a.M.nl, per = ngl * nq;
    constexpr int LMAX = NL_ ? NL_ : HN_MAXL;   // per-thread layer arrays stay in registers when the layer count is a compile-time constant
    const int e = blockIdx.x, tid = threadIdx.x;
    SOps o = load_sops(sm, ngl, nq);
    double* nod = sm + sops_doubles(ngl, nq);  // [5][npts]: dp',u',v', udp, vdp of the current layer; also z levels
    double* tA = nod + 5 * npts;               // [5][per]
    double* tB = tA + 5 * per;                 // [per]
    double* fq = tB + per;                     // [6][nq2]
    double* tP = fq + 6 * nq2;                 // [2][per]
    double* tR = tP + 2 * per;                 // [2][per]
    double* out = tR + 2 * per;                // [2][npts]
    const size_t nbase = (size_t)e * npts, qbase = (size_t)e * nq2;
    // metric terms of this thread's quadrature point (phases `tid < nq2`) and of its node (phases `tid < npts`); per element on affine meshes
    const Met mq_ = met_q(a.M, e, tid < nq2 ? tid : 0), mn_ = met_n(a.M, e, tid < npts ? tid : 0);
    const double eps1 = 1.0e-20;
    const double Pstress = (a.g / a.alpha[0]) * 50.0, Pbstress = (a.g / a.alpha[nl - 1]) * 10.0;
    // per quadrature point, per layer
    double p_tmp[LMAX + 1], H_tmp[LMAX], u_udp[LMAX], v_vdp[LMAX], u_vdp1[LMAX], u_vdp2[LMAX],
        temp_uu[LMAX], temp_vv[LMAX], gradz1[LMAX + 1], gradz2[LMAX + 1], dpq[LMAX];
    double qp_last[3] = {0, 0, 0};
    const int j = tid / nq, i = tid - j * nq;
    const bool qa = tid < nq2;
    double sq_ope2 = 0, ope_a = 0, ub_a = 0, vb_a = 0;
    if (qa) { sq_ope2 = sqrt(a.ave_q[5][qbase + tid]); ope_a = a.ave_q[0][qbase + tid]; ub_a = a.ave_q[8][qbase + tid]; vb_a = a.ave_q[9][qbase + tid]; }
    p_tmp[0] = 0.0;
    // gradients of the interface elevations z_elv(:,k), k = nl .. 0 (mod_create_rhs_mlswe.F90:320-325,362-367)
    {
        double zcur = 0.0;  // nodal thread: running z
        for (int k = nl; k >= 0; --k) {
            __syncthreads();
            if (tid < npts) {
                if (k == nl) zcur = a.zbot_df[nbase + tid];
                else zcur = zcur + (a.alpha[k] / a.g) * (sqrt(a.ope2_df[nbase + tid]) * a.qprime[(size_t)(0 * nl + k) * a.nstride + nbase + tid]);
                nod[tid] = zcur;
            }
            __syncthreads();
            sf_pass1(o, ngl, nq, 1, nod, npts, tA, tB);
            __syncthreads();
            if (qa) {
                double dks = sf_eval(o, ngl, nq, tB, 0, i, j);    // d/dksi: B in first direction, A in second
                double det = sf_eval_B(o, ngl, nq, tA, 0, i, j);  // d/deta
                gradz1[k] = mq_.ksx * dks + mq_.etx * det;
                gradz2[k] = mq_.ksy * dks + mq_.ety * det;
            }
        }
    }
    for (int k = 0; k < nl; ++k) {
        __syncthreads();
        if (tid < npts) {
            for (int v = 0; v < 3; ++v) nod[v * npts + tid] = a.qprime[(size_t)(v * nl + k) * a.nstride + nbase + tid];
            nod[3 * npts + tid] = a.q[(size_t)(1 * nl + k) * a.nstride + nbase + tid];
            nod[4 * npts + tid] = a.q[(size_t)(2 * nl + k) * a.nstride + nbase + tid];
        }
        __syncthreads();
        sf_pass1(o, ngl, nq, 5, nod, npts, tA, nullptr);
        __syncthreads();
        if (qa) {
            double q0 = sf_eval(o, ngl, nq, tA, 0, i, j), q1 = sf_eval(o, ngl, nq, tA, 1, i, j), q2 = sf_eval(o, ngl, nq, tA, 2, i, j);
            double tu = sf_eval(o, ngl, nq, tA, 3, i, j), tv = sf_eval(o, ngl, nq, tA, 4, i, j);
            qp_last[0] = q0; qp_last[1] = q1; qp_last[2] = q2;
            dpq[k] = q0;
            p_tmp[k + 1] = p_tmp[k] + sq_ope2 * q0;
            H_tmp[k] = 0.5 * a.alpha[k] * (p_tmp[k + 1] * p_tmp[k + 1] - p_tmp[k] * p_tmp[k]);
            double dp = q0 * ope_a, u = q1 + ub_a, v = q2 + vb_a;
            u_udp[k] = dp * u * u; v_vdp[k] = dp * v * v; u_vdp1[k] = u * v * dp; u_vdp2[k] = v * u * dp;
            temp_uu[k] = fabs(tu) + eps1; temp_vv[k] = fabs(tv) + eps1;
        }
    }
    double s_uu = 0, s_uv = 0, s_vv = 0, s_tu = 0, s_tv = 0, s_H = 0;
    double uu_def = 0, uv_def = 0, vv_def = 0, oosu = 0, oosv = 0, wq = 0, pbq = 0, twx = 0, twy = 0, tbx = 0, tby = 0, Hav = 0;
    if (qa) {
        for (int k = 0; k < nl; ++k) { s_uu += u_udp[k]; s_uv += u_vdp1[k]; s_vv += v_vdp[k]; s_tu += temp_uu[k]; s_tv += temp_vv[k]; s_H += H_tmp[k]; }
        size_t Iq = qbase + tid;
        uu_def = a.ave_q[2][Iq] - s_uu; uv_def = a.ave_q[4][Iq] - s_uv; vv_def = a.ave_q[3][Iq] - s_vv;
        oosu = 1.0 / s_tu; oosv = 1.0 / s_tv;
        wq = o.wq[i] * o.wq[j] * mq_.J;
        pbq = a.pbprime_q[Iq]; twx = a.tauwx_q[Iq]; twy = a.tauwy_q[Iq]; tbx = a.ave_q[10][Iq]; tby = a.ave_q[11][Iq]; Hav = a.ave_q[1][Iq];
    }
    double ppt = 0.0;  // pprime_temp(k)
    for (int k = 0; k < nl; ++k) {
        __syncthreads();
        if (qa) {
            // hazard 1 (mod_create_rhs_mlswe.F90:382): as written for nl<=3, intent (dp'_k) for nl>3
            double inc = (nl <= 3) ? qp_last[k] : dpq[k];
            double ppn = ppt + inc;
            double wgt = temp_uu[k] * oosu;
            double var_uu = u_udp[k] + wgt * uu_def;
            double var_uv = u_vdp1[k] + wgt * uv_def;
            wgt = temp_vv[k] * oosv;
            double var_vu = u_vdp2[k] + wgt * uv_def;
            double var_vv = v_vdp[k] + wgt * vv_def;
            double weight = 1.0;
            if (s_H > 0.0) weight = Hav / s_H;
            double Hq = H_tmp[k] * weight;
            double temp1 = (fmin(ppn, Pstress) - fmin(ppt, Pstress)) / Pstress;
            double tempbot = fmin(Pbstress, pbq - ppn) - fmin(Pbstress, pbq - ppt);
            tempbot = tempbot / Pbstress;
            double source_x = a.g * (temp1 * twx - tempbot * tbx + p_tmp[k] * gradz1[k] - p_tmp[k + 1] * gradz1[k + 1]);
            double source_y = a.g * (temp1 * twy - tempbot * tby + p_tmp[k] * gradz2[k] - p_tmp[k + 1] * gradz2[k + 1]);
            ppt = ppn;
            double Fx1 = Hq + var_uu, Fy1 = var_uv, Fx2 = var_vu, Fy2 = Hq + var_vv;
            fq[0 * nq2 + tid] = wq * source_x;
            fq[1 * nq2 + tid] = wq * source_y;
            fq[2 * nq2 + tid] = wq * (mq_.ksx * Fx1 + mq_.ksy * Fy1);
            fq[3 * nq2 + tid] = wq * (mq_.ksx * Fx2 + mq_.ksy * Fy2);
            fq[4 * nq2 + tid] = wq * (mq_.etx * Fx1 + mq_.ety * Fy1);
            fq[5 * nq2 + tid] = wq * (mq_.etx * Fx2 + mq_.ety * Fy2);
        }
        __syncthreads();
        sf_scatter(o, ngl, nq, 2, fq, fq + 2 * nq2, fq + 4 * nq2, nq2, tP, tR, out, npts, false);
        if (tid < npts) {
            a.rhs_mom[(size_t)(0 * nl + k) * a.nstride + nbase + tid] = out[tid];
            a.rhs_mom[(size_t)(1 * nl + k) * a.nstride + nbase + tid] = out[npts + tid];
        }
    }
}

// Same operator with ALL layers of an element in flight (compile-time layer count, small): one interpolation sweep for the nl+1
// interface elevations, one for the 5 nl layer fields, one weak-form scatter for the 2 nl momentum components -- 8 block barriers per
// element instead of 11 nl + 3, and every sweep keeps all threads busy.  Per-field arithmetic is that of k_mom_volume (the results
// agree to the last bits: the compiler may contract the multiply-adds of the two kernels differently).
template <int G_, int Q_, int NL_>
constexpr size_t mom_volume_b_doubles() { return (size_t)(2 * G_ * Q_ + G_ * G_ + Q_ + G_) + 5 * NL_ * G_ * G_ + 5 * NL_ * G_ * Q_ + (NL_ + 1) * G_ * Q_ + 6 * NL_ * Q_ * Q_ + 4 * NL_ * G_ * Q_ + 2 * NL_ * G_ * G_; }
#ifndef HN_MVB_MAXNREG
#define HN_MVB_MAXNREG 128   // 5 blocks of 96 threads per SM (56 B of spills): 1.12 ms per launch at 62 500 elements against 1.23 ms at 154 registers, 1.16 ms at 112
#endif
template <int G_, int Q_, int NL_>
__global__ void __maxnreg__(HN_MVB_MAXNREG) k_mom_volume_b(MomVolArgs a) {
    static_assert(G_ > 0 && Q_ > 0 && NL_ > 0, "compile-time sizes");
    extern __shared__ double sm[];
    constexpr int ngl = G_, nq = Q_, npts = ngl * ngl, nq2 = nq * nq, nl = NL_, per = ngl * nq;
    const int e = blockIdx.x, tid = threadIdx.x;
    SOps o = load_sops(sm, ngl, nq);
    double* nod = sm + sops_doubles(ngl, nq);  // [5 nl][npts]: first the nl+1 interface elevations, then dp',u',v',udp,vdp of every layer
    double* tA = nod + 5 * nl * npts;          // [5 nl][per]
    double* tB = tA + 5 * nl * per;            // [nl+1][per]
    double* fq = tB + (nl + 1) * per;          // [3][2 nl][nq2]: sources | ksi-fluxes | eta-fluxes of (layer, component)
    double* tP = fq + 6 * nl * nq2;            // [2 nl][per]
    double* tR = tP + 2 * nl * per;            // [2 nl][per]
    double* out = tR + 2 * nl * per;           // [2 nl][npts]
    const size_t nbase = (size_t)e * npts, qbase = (size_t)e * nq2;
    const Met mq_ = met_q(a.M, e, tid < nq2 ? tid : 0);
    const double eps1 = 1.0e-20;
    const double Pstress = (a.g / a.alpha[0]) * 50.0, Pbstress = (a.g / a.alpha[nl - 1]) * 10.0;
    double p_tmp[nl + 1], H_tmp[nl], u_udp[nl], v_vdp[nl], u_vdp1[nl], u_vdp2[nl], temp_uu[nl], temp_vv[nl], gradz1[nl + 1], gradz2[nl + 1], dpq[nl];
    double qp_last[3] = {0, 0, 0};
    const int j = tid / nq, i = tid - j * nq;
    const bool qa = tid < nq2;
    double sq_ope2 = 0, ope_a = 0, ub_a = 0, vb_a = 0;
    if (qa) { sq_ope2 = sqrt(a.ave_q[5][qbase + tid]); ope_a = a.ave_q[0][qbase + tid]; ub_a = a.ave_q[8][qbase + tid]; vb_a = a.ave_q[9][qbase + tid]; }
    p_tmp[0] = 0.0;
    __syncthreads();
    // interface elevations z_elv(:,k), k = nl .. 0 (mod_create_rhs_mlswe.F90:320-325,362-367)
    if (tid < npts) {
        double zcur = 0.0;
#pragma unroll
        for (int k = nl; k >= 0; --k) {
            if (k == nl) zcur = a.zbot_df[nbase + tid];
            else zcur = zcur + (a.alpha[k] / a.g) * (sqrt(a.ope2_df[nbase + tid]) * a.qprime[(size_t)(0 * nl + k) * a.nstride + nbase + tid]);
            nod[k * npts + tid] = zcur;
        }
    }
    __syncthreads();
    sf_pass1(o, ngl, nq, nl + 1, nod, npts, tA, tB);
    __syncthreads();
    if (qa) {
#pragma unroll
        for (int k = 0; k <= nl; ++k) {
            double dks = sf_eval(o, ngl, nq, tB, k, i, j);
            double det = sf_eval_B(o, ngl, nq, tA, k, i, j);
            gradz1[k] = mq_.ksx * dks + mq_.etx * det;
            gradz2[k] = mq_.ksy * dks + mq_.ety * det;
        }
    }
    __syncthreads();
    if (tid < npts) {
#pragma unroll
        for (int k = 0; k < nl; ++k) {
            for (int v = 0; v < 3; ++v) nod[(k * 5 + v) * npts + tid] = a.qprime[(size_t)(v * nl + k) * a.nstride + nbase + tid];
            nod[(k * 5 + 3) * npts + tid] = a.q[(size_t)(1 * nl + k) * a.nstride + nbase + tid];
            nod[(k * 5 + 4) * npts + tid] = a.q[(size_t)(2 * nl + k) * a.nstride + nbase + tid];
        }
    }
    __syncthreads();
    sf_pass1(o, ngl, nq, 5 * nl, nod, npts, tA, nullptr);
    __syncthreads();
    if (qa) {
#pragma unroll
        for (int k = 0; k < nl; ++k) {
            double q0 = sf_eval(o, ngl, nq, tA, k * 5 + 0, i, j), q1 = sf_eval(o, ngl, nq, tA, k * 5 + 1, i, j), q2 = sf_eval(o, ngl, nq, tA, k * 5 + 2, i, j);
            double tu = sf_eval(o, ngl, nq, tA, k * 5 + 3, i, j), tv = sf_eval(o, ngl, nq, tA, k * 5 + 4, i, j);
            qp_last[0] = q0; qp_last[1] = q1; qp_last[2] = q2;
            dpq[k] = q0;
            p_tmp[k + 1] = p_tmp[k] + sq_ope2 * q0;
            H_tmp[k] = 0.5 * a.alpha[k] * (p_tmp[k + 1] * p_tmp[k + 1] - p_tmp[k] * p_tmp[k]);
            double dp = q0 * ope_a, u = q1 + ub_a, v = q2 + vb_a;
            u_udp[k] = dp * u * u; v_vdp[k] = dp * v * v; u_vdp1[k] = u * v * dp; u_vdp2[k] = v * u * dp;
            temp_uu[k] = fabs(tu) + eps1; temp_vv[k] = fabs(tv) + eps1;
        }
        double s_uu = 0, s_uv = 0, s_vv = 0, s_tu = 0, s_tv = 0, s_H = 0;
        for (int k = 0; k < nl; ++k) { s_uu += u_udp[k]; s_uv += u_vdp1[k]; s_vv += v_vdp[k]; s_tu += temp_uu[k]; s_tv += temp_vv[k]; s_H += H_tmp[k]; }
        const size_t Iq = qbase + tid;
        const double uu_def = a.ave_q[2][Iq] - s_uu, uv_def = a.ave_q[4][Iq] - s_uv, vv_def = a.ave_q[3][Iq] - s_vv;
        const double oosu = 1.0 / s_tu, oosv = 1.0 / s_tv;
        const double wq = o.wq[i] * o.wq[j] * mq_.J;
        const double pbq = a.pbprime_q[Iq], twx = a.tauwx_q[Iq], twy = a.tauwy_q[Iq], tbx = a.ave_q[10][Iq], tby = a.ave_q[11][Iq], Hav = a.ave_q[1][Iq];
        double ppt = 0.0;  // pprime_temp(k)
#pragma unroll
        for (int k = 0; k < nl; ++k) {
            // hazard 1 (mod_create_rhs_mlswe.F90:382): as written for nl<=3, intent (dp'_k) for nl>3
            double inc = (nl <= 3) ? qp_last[k < 3 ? k : 0] : dpq[k];
            double ppn = ppt + inc;
            double wgt = temp_uu[k] * oosu;
            double var_uu = u_udp[k] + wgt * uu_def;
            double var_uv = u_vdp1[k] + wgt * uv_def;
            wgt = temp_vv[k] * oosv;
            double var_vu = u_vdp2[k] + wgt * uv_def;
            double var_vv = v_vdp[k] + wgt * vv_def;
            double weight = 1.0;
            if (s_H > 0.0) weight = Hav / s_H;
            double Hq = H_tmp[k] * weight;
            double temp1 = (fmin(ppn, Pstress) - fmin(ppt, Pstress)) / Pstress;
            double tempbot = fmin(Pbstress, pbq - ppn) - fmin(Pbstress, pbq - ppt);
            tempbot = tempbot / Pbstress;
            double source_x = a.g * (temp1 * twx - tempbot * tbx + p_tmp[k] * gradz1[k] - p_tmp[k + 1] * gradz1[k + 1]);
            double source_y = a.g * (temp1 * twy - tempbot * tby + p_tmp[k] * gradz2[k] - p_tmp[k + 1] * gradz2[k + 1]);
            ppt = ppn;
            double Fx1 = Hq + var_uu, Fy1 = var_uv, Fx2 = var_vu, Fy2 = Hq + var_vv;
            fq[(0 * nl + 2 * k + 0) * nq2 + tid] = wq * source_x;
            fq[(0 * nl + 2 * k + 1) * nq2 + tid] = wq * source_y;
            fq[(2 * nl + 2 * k + 0) * nq2 + tid] = wq * (mq_.ksx * Fx1 + mq_.ksy * Fy1);
            fq[(2 * nl + 2 * k + 1) * nq2 + tid] = wq * (mq_.ksx * Fx2 + mq_.ksy * Fy2);
            fq[(4 * nl + 2 * k + 0) * nq2 + tid] = wq * (mq_.etx * Fx1 + mq_.ety * Fy1);
            fq[(4 * nl + 2 * k + 1) * nq2 + tid] = wq * (mq_.etx * Fx2 + mq_.ety * Fy2);
        }
    }
    __syncthreads();
    sf_scatter(o, ngl, nq, 2 * nl, fq, fq + 2 * nl * nq2, fq + 4 * nl * nq2, nq2, tP, tR, out, npts, false);
    for (int t = tid; t < 2 * nl * npts; t += blockDim.x) {
        const int f = t / npts, k = f >> 1, c = f & 1;
        a.rhs_mom[(size_t)(c * nl + k) * a.nstride + nbase + (t - f * npts)] = out[t];
    }
}

// --------------------------------------------------------------------------------------------------------------
// Apply_layers_fluxes + momentum update + Coriolis rotation + wall projection + velocity reconciliation
struct MomFaceArgs {
    Mesh M;
    const double* qprime;   // [3*nl] input primes (for the traces)
    const double* q_in;     // [3*nl] q_df before the step: momentum planes are read here
    double* q;              // [3*nl] q_df after it: thickness planes are read (already updated), momentum planes written (may alias q_in)
    double* qprime_out;     // [3*nl]
    size_t nstride;
    const double* hq; size_t hstride;
    const double* qb[3];    // pbpert, mx, my of the barotropic state used by evaluate_bcl
    const double* pbprime_df;
    const double* ave_f[16];
    const double *zbf_l, *zbf_r;
    const double *rhs_mom, *rhs_visc;  // [2*nl]
    const double *massinv, *a_bcl, *b_bcl, *fdt2;
    double alpha[HN_MAXL];
    double g, dt;
    int full_prime;  // 1: evaluate_bcl (predictor), 0: evaluate_bcl_v1 (corrector)
    int rhs_only;    // 1: write rhs_mom = massinv (volume + faces) + rhs_visc to rhs_out [2*nl] and stop (hnumo_layer_momentum_rhs)
    double* rhs_out;
};
template <int G_, int Q_, int NL_>
__global__ void k_mom_faces_update(MomFaceArgs a) {
    extern __shared__ double sm[];
    const int ngl = G_ ? G_ : a.M.ngl, nq = Q_ ? Q_ : a.M.nq, npts = ngl * ngl, nl = NL_ ? NL_ : a.M.nl;
    constexpr int LMAX = NL_ ? NL_ : HN_MAXL;
    const int e = blockIdx.x, tid = threadIdx.x;
    SOps o = load_sops(sm, ngl, nq);
    double* ownq = sm + sops_doubles(ngl, nq);  // [4][3][nl][ngl]
    double* nbq = ownq + 12 * nl * ngl;         // [4][3][nl][ngl]
    double* ff = nbq + 12 * nl * ngl;           // [4][nl][2][nq]
    const size_t nbase = (size_t)e * npts;
    const double eps1 = 1.0e-20;
    for (int t = tid; t < 4 * ngl * nl; t += blockDim.x) {
        int k = t / (4 * ngl), r = t - k * 4 * ngl, s = r / ngl, n = r - s * ngl;
        int slot = e * 4 + s, nb = a.M.nbr[slot], nbs = a.M.nbslot[slot];
        int I = face_node(s, n, ngl);
        double ow[3], nv[3];
        for (int v = 0; v < 3; ++v) {
            ow[v] = a.qprime[(size_t)(v * nl + k) * a.nstride + nbase + I];
            nv[v] = nb_nodal(a.M, a.qprime + (size_t)(v * nl + k) * a.nstride, a.hq + (size_t)(v * nl + k) * a.hstride, nb, nbs, n, ow[v]);
        }
        if (nb == NBR_FREESLIP) {
            const FGeo fgn_ = fg_n(a.M, slot, n);
            double nx = fgn_.nx, ny = fgn_.ny;
            double un = ow[1] * nx + ow[2] * ny;
            nv[1] = ow[1] - 2.0 * un * nx; nv[2] = ow[2] - 2.0 * un * ny;
        } else if (nb == NBR_NOSLIP) { nv[1] = -ow[1]; nv[2] = -ow[2]; }
        for (int v = 0; v < 3; ++v) { ownq[((s * 3 + v) * nl + k) * ngl + n] = ow[v]; nbq[((s * 3 + v) * nl + k) * ngl + n] = nv[v]; }
    }
    __syncthreads();
    if (tid < 4 * nq) {
        int s = tid / nq, iq = tid - s * nq, slot = e * 4 + s, nb = a.M.nbr[slot], nbs = a.M.nbslot[slot];
        bool left = (nb < 0) || (e < nb);
        int oslot = left ? slot : nb * 4 + nbs;
        size_t fo = (size_t)oslot * nq + iq;
        const FGeo fgq_ = fg_q(a.M, slot, iq);
            double nxl = fgq_.nx, nyl = fgq_.ny, nlen = fgq_.len;
        const double* tl = left ? ownq : nbq;
        const double* tr = left ? nbq : ownq;
        double ql0[LMAX], qr0[LMAX];
        double udpl[LMAX], udpr[LMAX], vdpl[LMAX], vdpr[LMAX];
        double uf0[LMAX], uf1[LMAX], vf0[LMAX], vf1[LMAX], HL[LMAX], HR[LMAX];
        double qbl0 = a.ave_f[7][fo], qbl1 = a.ave_f[12][fo], qbl2 = a.ave_f[14][fo];
        double qbr0 = a.ave_f[8][fo], qbr1 = a.ave_f[13][fo], qbr2 = a.ave_f[15][fo];
        for (int k = 0; k < nl; ++k) {
            double ql[3] = {0, 0, 0}, qr[3] = {0, 0, 0};
            for (int n = 0; n < ngl; ++n) {
                double hi = o.A[n + ngl * iq];
                for (int v = 0; v < 3; ++v) { ql[v] += hi * tl[((s * 3 + v) * nl + k) * ngl + n]; qr[v] += hi * tr[((s * 3 + v) * nl + k) * ngl + n]; }
            }
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
        double wq = o.wq[iq] * nlen;
        for (int k = 0; k < nl; ++k) {
            double Hs = left ? HL[k] * wl : HR[k] * wr;
            double flux_x = nxl * uf0[k] + nyl * uf1[k];
            double flux_y = nxl * vf0[k] + nyl * vf1[k];
            double sgn = left ? -1.0 : 1.0;
            ff[((s * nl + k) * 2 + 0) * nq + iq] = sgn * wq * (nxl * Hs + flux_x);
            ff[((s * nl + k) * 2 + 1) * nq + iq] = sgn * wq * (nyl * Hs + flux_y);
        }
    }
    __syncthreads();
    if (tid < npts) {
        int m = tid / ngl, n = tid - m * ngl;
        double qd[LMAX], qx[LMAX], qy[LMAX];
        double mi = a.massinv[nbase + tid];
        double f2 = a.fdt2[nbase + tid], ab = a.a_bcl[nbase + tid], bb = a.b_bcl[nbase + tid];
        for (int k = 0; k < nl; ++k) {
            double r0 = a.rhs_mom[(size_t)(0 * nl + k) * a.nstride + nbase + tid];
            double r1 = a.rhs_mom[(size_t)(1 * nl + k) * a.nstride + nbase + tid];
            for (int s = 0; s < 4; ++s) {
                int nf;
                if (s == 0) { if (m != 0) continue; nf = n; }
                else if (s == 1) { if (m != ngl - 1) continue; nf = n; }
                else if (s == 2) { if (n != 0) continue; nf = m; }
                else { if (n != ngl - 1) continue; nf = m; }
                double p0 = 0.0, p1 = 0.0;
                for (int iq = 0; iq < nq; ++iq) {
                    double hi = o.A[nf + ngl * iq];
                    p0 += hi * ff[((s * nl + k) * 2 + 0) * nq + iq]; p1 += hi * ff[((s * nl + k) * 2 + 1) * nq + iq];
                }
                r0 += p0; r1 += p1;
            }
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
                    const FGeo fgn_ = fg_n(a.M, slot, s < 2 ? n : m);
                    double nx = fgn_.nx, ny = fgn_.ny;
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

// --------------------------------------------------------------------------------------------------------------
// Vertical shear stress between the layers (ad_mlswe > 0) + the momentum update it sits in
// (mod_splitting.F90:139-179 / 247-286, mod_create_rhs_mlswe.F90:146-279, mod_layer_terms.F90:139-196).
// Runs instead of the update epilogue of k_mom_faces_update (which is launched with rhs_only = 1 and leaves the complete
// rhs_mom in `rhs`): per element
//   A. nodes:  q_df_temp = q + dt rhs_mom; q_df3 = Coriolis rotation of it; velocity_df (velocities reconciled with the
//              barotropic velocity, momentum rebuilt);
//   B. quadrature points: dp, udp, vdp of every layer interpolated; the tridiagonal system of the implicit shear stress
//              as written (sub-diagonal -coeff, super-diagonal -coeff1, right-hand side u = udp/dp); tau at the interfaces;
//   C. nodes:  rhs_stress = sum_q wq psi tau_q (no mass matrix inside), q_df_temp += dt massinv rhs_stress, then the
//              rotation, the wall projection and evaluate_bcl(_v1) exactly like the epilogue of k_mom_faces_update.
// Two values the reference leaves undefined are set to the evident intent (parity hazards 2 and 3): the stress
// through the bottom interface tau(nlayers+1) = 0, and the corrector (`momentum`) uses q_df3 like the predictor does.
// One block per element, run-time sizes (optional physics: not on the benchmark path).
struct ShearArgs {
    Mesh M;
    const double* q_in;     // [3*nl] q_df before the step: momentum planes are read
    double* q;              // [3*nl] thickness planes read (already updated), momentum planes written (may alias q_in)
    double* qprime_out;     // [3*nl]
    size_t nstride;
    const double* qb[3];    // pbpert, mx, my of the barotropic state
    const double* pbprime_df;
    const double* rhs;      // [2*nl] rhs_mom = massinv (volume + faces) + rhs_visc
    const double *massinv, *a_bcl, *b_bcl, *fdt2, *coriolis_q;
    double alpha0, g, dt, ad, max_shear_dz;
    int full_prime;
    // per-phase entry hnumo_layer_shear_stress: rhs_layer_shear_stress of q itself -> stress_out [2*nl], nothing else
    int stress_only;
    double* stress_out;
};
__global__ void k_shear_update(ShearArgs a) {
    extern __shared__ double sm[];
    const int ngl = a.M.ngl, nq = a.M.nq, npts = ngl * ngl, nq2 = nq * nq, nl = a.M.nl;
    constexpr int LMAX = HN_MAXL;
    const int e = blockIdx.x, tid = threadIdx.x;
    SOps o = load_sops(sm, ngl, nq);
    double* s3 = sm + sops_doubles(ngl, nq);   // [3][nl][npts]: dp, udp, vdp of q_df3 after velocity_df
    double* st = s3 + 3 * nl * npts;           // [2][nl][npts]: q_df_temp
    double* tq = st + 2 * nl * npts;           // [2][nl][nq2]: wq * gravity * (tau(k) - tau(k+1))
    const size_t nbase = (size_t)e * npts, qbase = (size_t)e * nq2;
    __syncthreads();
    if (a.stress_only) {
        for (int t = tid; t < 3 * nl * npts; t += blockDim.x) { const int vk = t / npts; s3[t] = a.q[(size_t)vk * a.nstride + nbase + (t - vk * npts)]; }
    } else
    for (int t = tid; t < npts; t += blockDim.x) {
        const double f2 = a.fdt2[nbase + t], ab = a.a_bcl[nbase + t], bb = a.b_bcl[nbase + t];
        const double pb = a.qb[0][nbase + t] + a.pbprime_df[nbase + t];
        const double mbx = a.qb[1][nbase + t], mby = a.qb[2][nbase + t];
        double dpk[LMAX], uk[LMAX], vk[LMAX];
        double ubar = 0.0, vbar = 0.0;
        for (int k = 0; k < nl; ++k) {
            const double mxo = a.q_in[(size_t)(1 * nl + k) * a.nstride + nbase + t], myo = a.q_in[(size_t)(2 * nl + k) * a.nstride + nbase + t];
            const double t1 = mxo + a.dt * a.rhs[(size_t)(0 * nl + k) * a.nstride + nbase + t];
            const double t2 = myo + a.dt * a.rhs[(size_t)(1 * nl + k) * a.nstride + nbase + t];
            st[(0 * nl + k) * npts + t] = t1; st[(1 * nl + k) * npts + t] = t2;
            const double tempu = t1 + f2 * myo, tempv = t2 - f2 * mxo;
            const double m3x = ab * tempu + bb * tempv, m3y = -bb * tempu + ab * tempv;
            dpk[k] = a.q[(size_t)(0 * nl + k) * a.nstride + nbase + t];
            uk[k] = m3x / dpk[k]; vk[k] = m3y / dpk[k];
        }
        for (int k = 0; k < nl; ++k) { ubar = ubar + uk[k] * dpk[k]; vbar = vbar + vk[k] * dpk[k]; }
        if (pb > 0.0) {
            ubar = ubar / pb; vbar = vbar / pb;
            for (int k = 0; k < nl; ++k) { uk[k] = uk[k] - ubar + mbx / pb; vk[k] = vk[k] - vbar + mby / pb; }
        } else {
            for (int k = 0; k < nl; ++k) { uk[k] = 0.0; vk[k] = 0.0; }
        }
        for (int k = 0; k < nl; ++k) {
            s3[(0 * nl + k) * npts + t] = dpk[k]; s3[(1 * nl + k) * npts + t] = uk[k] * dpk[k]; s3[(2 * nl + k) * npts + t] = vk[k] * dpk[k];
        }
    }
    __syncthreads();
    for (int t = tid; t < nq2; t += blockDim.x) {
        const int j = t / nq, i = t - j * nq;
        double dp[LMAX], u[LMAX], v[LMAX], b[LMAX], tu[LMAX + 1], tv[LMAX + 1];
        for (int k = 0; k < nl; ++k) {
            double d0 = 0.0, d1 = 0.0, d2 = 0.0;
            for (int m = 0; m < ngl; ++m)
                for (int n = 0; n < ngl; ++n) {
                    const double hi = o.A[n + ngl * i] * o.A[m + ngl * j];
                    d0 = d0 + hi * s3[(0 * nl + k) * npts + m * ngl + n];
                    d1 = d1 + hi * s3[(1 * nl + k) * npts + m * ngl + n];
                    d2 = d2 + hi * s3[(2 * nl + k) * npts + m * ngl + n];
                }
            dp[k] = d0; u[k] = d1 / d0; v[k] = d2 / d0;
        }
        const double coeff = fmax(sqrt(0.5 * a.coriolis_q[qbase + t] * a.ad) / a.alpha0, a.ad / (a.alpha0 * a.max_shear_dz));
        const double coeff1 = a.g * a.dt * coeff;
        // a(k) = -coeff (k > 1), c(k) = -coeff1 (k < nl), b(k) = dp(k) + 2 coeff1, b(1) and b(nl) = dp + coeff1
        for (int k = 0; k < nl; ++k) b[k] = dp[k] + 2.0 * coeff1;
        b[0] = dp[0] + coeff1;
        b[nl - 1] = dp[nl - 1] + coeff1;
        for (int k = 1; k < nl; ++k) {
            const double mult = -coeff / b[k - 1];
            const double ckm = (k - 1 == nl - 1) ? 0.0 : -coeff1;
            b[k] = b[k] - mult * ckm;
            u[k] = u[k] - mult * u[k - 1];
            v[k] = v[k] - mult * v[k - 1];
        }
        u[nl - 1] = u[nl - 1] / b[nl - 1];
        v[nl - 1] = v[nl - 1] / b[nl - 1];
        for (int k = nl - 2; k >= 0; --k) {
            u[k] = (u[k] - (-coeff1) * u[k + 1]) / b[k];
            v[k] = (v[k] - (-coeff1) * v[k + 1]) / b[k];
        }
        tu[0] = 0.0; tv[0] = 0.0;
        for (int k = 1; k < nl; ++k) { tu[k] = coeff * (u[k - 1] - u[k]); tv[k] = coeff * (v[k - 1] - v[k]); }
        tu[nl] = 0.0; tv[nl] = 0.0;
        const double wq = o.wq[i] * o.wq[j] * met_q(a.M, e, t).J;
        for (int k = 0; k < nl; ++k) {
            tq[(0 * nl + k) * nq2 + t] = wq * (a.g * (tu[k] - tu[k + 1]));
            tq[(1 * nl + k) * nq2 + t] = wq * (a.g * (tv[k] - tv[k + 1]));
        }
    }
    __syncthreads();
    for (int t = tid; t < npts; t += blockDim.x) {
        const int m = t / ngl, n = t - m * ngl;
        const double mi = a.massinv[nbase + t];
        const double f2 = a.fdt2[nbase + t], ab = a.a_bcl[nbase + t], bb = a.b_bcl[nbase + t];
        double qd[LMAX], qx[LMAX], qy[LMAX];
        for (int k = 0; k < nl; ++k) {
            double r0 = 0.0, r1 = 0.0;
            for (int j = 0; j < nq; ++j)
                for (int i = 0; i < nq; ++i) {
                    const double hi = o.A[n + ngl * i] * o.A[m + ngl * j];
                    r0 = r0 + hi * tq[(0 * nl + k) * nq2 + j * nq + i];
                    r1 = r1 + hi * tq[(1 * nl + k) * nq2 + j * nq + i];
                }
            if (a.stress_only) {
                a.stress_out[(size_t)(0 * nl + k) * a.nstride + nbase + t] = r0;
                a.stress_out[(size_t)(1 * nl + k) * a.nstride + nbase + t] = r1;
                qd[k] = 1.0; qx[k] = 0.0; qy[k] = 0.0;
                continue;
            }
            const double t1 = st[(0 * nl + k) * npts + t] + a.dt * (mi * r0);
            const double t2 = st[(1 * nl + k) * npts + t] + a.dt * (mi * r1);
            const double mxo = a.q_in[(size_t)(1 * nl + k) * a.nstride + nbase + t], myo = a.q_in[(size_t)(2 * nl + k) * a.nstride + nbase + t];
            const double tempu = t1 + f2 * myo, tempv = t2 - f2 * mxo;
            double mxn = ab * tempu + bb * tempv, myn = -bb * tempu + ab * tempv;
            for (int s = 0; s < 4; ++s) {   // wall projection (layer_mom_boundary_df)
                const bool on = (s == 0) ? (m == 0) : (s == 1) ? (m == ngl - 1) : (s == 2) ? (n == 0) : (n == ngl - 1);
                if (!on) continue;
                const int slot = e * 4 + s, nb = a.M.nbr[slot];
                if (nb == NBR_FREESLIP) {
                    const FGeo fgn_ = fg_n(a.M, slot, s < 2 ? n : m);
                    const double nx = fgn_.nx, ny = fgn_.ny;
                    const double up = mxn * nx + myn * ny;
                    mxn = mxn - up * nx; myn = myn - up * ny;
                } else if (nb == NBR_NOSLIP) { mxn = 0.0; myn = 0.0; }
            }
            qd[k] = s3[(0 * nl + k) * npts + t]; qx[k] = mxn; qy[k] = myn;
        }
        if (a.stress_only) continue;
        // evaluate_bcl / evaluate_bcl_v1: two passes of extract_velocity
        const double pb = a.qb[0][nbase + t] + a.pbprime_df[nbase + t];
        const double mbx = a.qb[1][nbase + t], mby = a.qb[2][nbase + t];
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
        ope = ope / a.pbprime_df[nbase + t];
        for (int k = 0; k < nl; ++k) {
            a.q[(size_t)(1 * nl + k) * a.nstride + nbase + t] = qx[k];
            a.q[(size_t)(2 * nl + k) * a.nstride + nbase + t] = qy[k];
            if (a.full_prime) a.qprime_out[(size_t)(0 * nl + k) * a.nstride + nbase + t] = qd[k] / ope;
            a.qprime_out[(size_t)(1 * nl + k) * a.nstride + nbase + t] = uk[k] - mbx / pb;
            a.qprime_out[(size_t)(2 * nl + k) * a.nstride + nbase + t] = vk[k] - mby / pb;
        }
    }
}

// --------------------------------------------------------------------------------------------------------------
// small pointwise kernels
// out = 0.5*(x + y)   (ti_rk_bcl.F90:64-66)
__global__ void k_average(double* out, const double* x, const double* y, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = 0.5 * (x[i] + y[i]);
}
// thickness epilogue (mod_splitting.F90:83-87, ti_rk_bcl.F90:78-79): dpprime2 = q_dp/ope ; qprime2.dp = 0.5*(qprime.dp + dpprime2)
// (dpprime2 may be the same planes as qprime_dp: every point reads its old value before it writes the new one, and no
//  point touches another point's entries -- the __restrict__ qualifiers only tell the compiler that the loads of one layer
//  need not wait for the stores of the previous one)
__global__ void k_thickness_finish(const double* __restrict__ qdp, const double* __restrict__ pbprime_df, const double* __restrict__ qprime_dp,
                                   double* __restrict__ dpprime2, double* __restrict__ qprime2_dp, int nl, size_t nstride, size_t npoin) {
    size_t I = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= npoin) return;
    double s = 0.0;
    for (int k = 0; k < nl; ++k) s += qdp[(size_t)k * nstride + I];
    double ope = s / pbprime_df[I];
    // four layers at a time: all loads of the group are issued before its first store
    for (int k0 = 0; k0 < nl; k0 += 4) {
        double qd[4], old[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = min(k0 + j, nl - 1);
            qd[j] = qdp[(size_t)k * nstride + I]; old[j] = qprime_dp[(size_t)k * nstride + I];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (k0 + j < nl) {
                const double d = qd[j] / ope;
                dpprime2[(size_t)(k0 + j) * nstride + I] = d;
                qprime2_dp[(size_t)(k0 + j) * nstride + I] = 0.5 * (old[j] + d);
            }
        }
    }
}

}  // namespace hn
