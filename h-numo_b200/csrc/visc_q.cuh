// visc_q.cuh -- quadrature-point LDG viscosity, method_visc == 1.
//
// Replaces (reference file:line)
//   interpolate_dpp                    src/mod_layer_terms.F90:25-55
//   compute_gradient_uv_q              src/mod_barotropic_terms.F90:445-477
//   btp_create_laplacian_v2            src/mod_laplacian_quad.F90:125-223   (with k_btp_stage_simple, btp_kernels.cuh)
//   bcl_create_laplacian_v2            src/mod_laplacian_quad.F90:252-355
//   compute_laplacian_quad             src/mod_laplacian_quad.F90:613-642
//   create_rhs_laplacian_flux_quad     src/mod_laplacian_quad.F90:644-722
//   create_communicator_quad, bcl_create_communicator(.,4,nlayers,nq)   src/create_rhs_communicator.F90:82-134
//
// Barotropic flux variable.  The reference rebuilds, in every barotropic stage,
//     F_c(Iq) = sum_k dpprime_visc_q(Iq,k) * grad_c(u'_k + ub)(Iq),      c = (du/dx, du/dy, dv/dx, dv/dy),
// looping over all layers.  The gradient is linear and u'_k, dpprime_visc_q do not change inside the substep loop, so
//     F_c = S_c + P * grad_c(ub),   S_c = sum_k dpprime_visc_q(k) grad_c(u'_k),   P = sum_k dpprime_visc_q(k):
// k_viscq_coeffs builds S_c and P once per loop (with the other btp<-bcl coefficients) and a stage costs one gradient,
// independent of the number of layers.  Same numbers up to summation order.
//
// Face values.  A face quadrature point coincides with a volume quadrature point of either element (imapl_q / imapr_q), so
// every element publishes F at its 4 nq boundary points in slot planes [c][(slot (+ halo)) * nq + iq]; both elements of a
// face -- and the neighbour rank, through one halo message -- read the SAME numbers, which keeps the central flux
// antisymmetric to the last bit.  Ghosts: copy of the left value, mirrored about the normal on free-slip walls.
//
// Run-time sizes, one block per element (optional physics, not on the benchmark path).
#pragma once
#include "hnumo_dev.cuh"

namespace hn {

// reference-space derivatives of a nodal field at quadrature point (i,j), direct double sum
__device__ __forceinline__ void vq_grad_point(const SOps& o, int ngl, const double* f, int i, int j, double& dks, double& det) {
    dks = 0.0; det = 0.0;
    for (int m = 0; m < ngl; ++m) {
        const double am = o.A[m + ngl * j], bm = o.B[m + ngl * j];
        double rb = 0.0, ra = 0.0;
        for (int n = 0; n < ngl; ++n) { rb += o.B[n + ngl * i] * f[m * ngl + n]; ra += o.A[n + ngl * i] * f[m * ngl + n]; }
        dks += am * rb; det += bm * ra;
    }
}
__device__ __forceinline__ double vq_interp_point(const SOps& o, int ngl, const double* f, int i, int j) {
    double v = 0.0;
    for (int m = 0; m < ngl; ++m) {
        double r = 0.0;
        for (int n = 0; n < ngl; ++n) r += o.A[n + ngl * i] * f[m * ngl + n];
        v += o.A[m + ngl * j] * r;
    }
    return v;
}

// ---- S_c, P at the quadrature points (once per substep loop) -----------------------------------------------------------
struct VqCoeffArgs {
    Mesh M;
    const double* qprime;   // [3*nl]: u', v' are planes nl..3nl-1
    const double* dpv;      // [nl] dpprime_visc
    size_t nstride;
    double* P;              // [npoin_q]
    double* S[4];           // [npoin_q]
};
__global__ void k_viscq_coeffs(VqCoeffArgs a) {
    extern __shared__ double sm[];
    const int ngl = a.M.ngl, nq = a.M.nq, npts = ngl * ngl, nq2 = nq * nq, per = ngl * nq, nl = a.M.nl;
    const int e = blockIdx.x, tid = threadIdx.x;
    SOps o = load_sops(sm, ngl, nq);
    double* nod = sm + sops_doubles(ngl, nq);   // [3][npts]
    double* tA = nod + 3 * npts;                // [3][per]
    double* tB = tA + 3 * per;                  // [3][per]
    const size_t nbase = (size_t)e * npts, qbase = (size_t)e * nq2;
    const Met mt = met_q(a.M, e, tid < nq2 ? tid : 0);   // metric terms at this thread's quadrature point
    const double ksx = mt.ksx, ksy = mt.ksy, etx = mt.etx, ety = mt.ety;
    double P = 0.0, S0 = 0.0, S1 = 0.0, S2 = 0.0, S3 = 0.0;
    for (int k = 0; k < nl; ++k) {
        __syncthreads();
        for (int t = tid; t < npts; t += blockDim.x) {
            nod[t] = a.dpv[(size_t)k * a.nstride + nbase + t];
            nod[npts + t] = a.qprime[(size_t)(1 * nl + k) * a.nstride + nbase + t];
            nod[2 * npts + t] = a.qprime[(size_t)(2 * nl + k) * a.nstride + nbase + t];
        }
        __syncthreads();
        sf_pass1(o, ngl, nq, 3, nod, npts, tA, tB);
        __syncthreads();
        if (tid < nq2) {
            const int j = tid / nq, i = tid - j * nq;
            const double d = sf_eval(o, ngl, nq, tA, 0, i, j);
            const double uk = sf_eval(o, ngl, nq, tB, 1, i, j), ue = sf_eval_B(o, ngl, nq, tA, 1, i, j);
            const double vk = sf_eval(o, ngl, nq, tB, 2, i, j), ve = sf_eval_B(o, ngl, nq, tA, 2, i, j);
            P += d;
            S0 += d * (ksx * uk + etx * ue); S1 += d * (ksy * uk + ety * ue);
            S2 += d * (ksx * vk + etx * ve); S3 += d * (ksy * vk + ety * ve);
        }
    }
    if (tid < nq2) {
        a.P[qbase + tid] = P;
        a.S[0][qbase + tid] = S0; a.S[1][qbase + tid] = S1; a.S[2][qbase + tid] = S2; a.S[3][qbase + tid] = S3;
    }
}

// F_c = S_c + P grad_c(u,v) at boundary quadrature point iq of side s; u, v: nodal velocities of the element (shared memory)
__device__ __forceinline__ void vq_face_point(const Mesh& M, int e, const SOps& o, int ngl, int nq, const double* u, const double* v, int s, int iq,
                                              const double* P, const double* const* S, size_t qbase, double out[4]) {
    const int q = face_quad(s, iq, nq), j = q / nq, i = q - j * nq;
    const Met mt = met_q(M, e, q);
    const double ksx = mt.ksx, ksy = mt.ksy, etx = mt.etx, ety = mt.ety;
    double uk, ue, vk, ve;
    vq_grad_point(o, ngl, u, i, j, uk, ue);
    vq_grad_point(o, ngl, v, i, j, vk, ve);
    const double p = P[qbase + q];
    out[0] = S[0][qbase + q] + p * (ksx * uk + etx * ue); out[1] = S[1][qbase + q] + p * (ksy * uk + ety * ue);
    out[2] = S[2][qbase + q] + p * (ksx * vk + etx * ve); out[3] = S[3][qbase + q] + p * (ksy * vk + ety * ve);
}

// face values of a barotropic state (start of the substep loop; afterwards the stage kernel publishes them)
struct VqPrimeArgs {
    Mesh M;
    const double* qb[3];
    const double* pbprime_df;
    const double* P;
    const double* S[4];
    double* trq;            // [4][(nslots + nhalo) * nq]
    size_t trq_stride;
};
__global__ void k_viscq_prime(VqPrimeArgs a) {
    extern __shared__ double sm[];
    const int ngl = a.M.ngl, nq = a.M.nq, npts = ngl * ngl, nq2 = nq * nq;
    const int e = blockIdx.x, tid = threadIdx.x;
    SOps o = load_sops(sm, ngl, nq);
    double* u = sm + sops_doubles(ngl, nq);
    double* v = u + npts;
    const size_t nbase = (size_t)e * npts, qbase = (size_t)e * nq2;
    for (int t = tid; t < npts; t += blockDim.x) {
        const double pb = a.qb[0][nbase + t] + a.pbprime_df[nbase + t];
        u[t] = a.qb[1][nbase + t] / pb; v[t] = a.qb[2][nbase + t] / pb;
    }
    __syncthreads();
    for (int t = tid; t < 4 * nq; t += blockDim.x) {
        const int s = t / nq, iq = t - s * nq;
        double F[4];
        vq_face_point(a.M, e, o, ngl, nq, u, v, s, iq, a.P, a.S, qbase, F);
        for (int c = 0; c < 4; ++c) a.trq[c * a.trq_stride + ((size_t)e * 4 + s) * nq + iq] = F[c];
    }
}

// own and neighbour face values of slot (e,s) at face point iq from the slot planes; planes pl[c] = base + c*cstride.
// Returns the LDG face flux of create_rhs_laplacian_flux_quad (as written, beta = 0.5) times the face weight, signed for
// element e (left: +, right: -).
__device__ __forceinline__ void vq_face_flux(const Mesh& M, const double* base, size_t cstride, int e, int s, int iq, double wq_iq, double out[2]) {
    const int nq = M.nq, slot = e * 4 + s, nb = M.nbr[slot], nbs = M.nbslot[slot];
    const FGeo fgq_ = fg_q(M, slot, iq);
    const double nx = fgq_.nx, ny = fgq_.ny, nlen = fgq_.len;
    double own[4], nbv[4];
    for (int c = 0; c < 4; ++c) own[c] = base[c * cstride + (size_t)slot * nq + iq];
    if (nb >= 0) { for (int c = 0; c < 4; ++c) nbv[c] = base[c * cstride + ((size_t)nb * 4 + nbs) * nq + iq]; }
    else if (nb == NBR_HALO) { for (int c = 0; c < 4; ++c) nbv[c] = base[c * cstride + ((size_t)M.nslots + nbs) * nq + iq]; }
    else {
        for (int c = 0; c < 4; ++c) nbv[c] = own[c];
        if (nb == NBR_FREESLIP) {
            double un = own[0] * nx + own[1] * ny;
            nbv[0] = own[0] - 2.0 * un * nx; nbv[1] = own[1] - 2.0 * un * ny;
            un = own[2] * nx + own[3] * ny;
            nbv[2] = own[2] - 2.0 * un * nx; nbv[3] = own[3] - 2.0 * un * ny;
        }
    }
    const bool left = (nb < 0) || (e < nb);
    const double* fl = left ? own : nbv;
    const double* fr = left ? nbv : own;
    const double qu0 = 0.5 * fl[0] + 0.5 * fr[0], qu1 = 0.5 * fl[1] + 0.5 * fr[1];
    const double qv0 = 0.5 * fl[2] + 0.5 * fr[2], qv1 = 0.5 * fl[3] + 0.5 * fr[3];
    const double flux_qu = (qu0 - fl[0] * nx) + (qu1 - fl[1] * ny);
    const double flux_qv = (qv0 - fl[2] * nx) + (qv1 - fl[3] * ny);
    const double w = (left ? 1.0 : -1.0) * (wq_iq * nlen);
    out[0] = w * flux_qu; out[1] = w * flux_qv;
}

// ---- layers: bcl_create_laplacian_v2 -----------------------------------------------------------------------------------
struct VqLayerArgs {
    Mesh M;
    const double* qprime;     // [3*nl]
    const double* dpv;        // [nl] dpprime_visc (dpprime_visc_q is its interpolation)
    size_t nstride;
    const double *ub_df, *vb_df;   // uvb_ave_df
    double* trq;              // [4*nl] slot planes [(c*nl + k)][(nslots + nhalo) * nq]
    size_t trq_stride;
    const double* massinv;
    double* rhs_visc;         // [2*nl]
    double visc;
};
// pass 1: face values of dpprime_visc_q(k) grad(u'_k + uvb_ave_df) for every layer
__global__ void k_bcl_lapq_traces(VqLayerArgs a) {
    extern __shared__ double sm[];
    const int ngl = a.M.ngl, nq = a.M.nq, npts = ngl * ngl, nl = a.M.nl;
    const int e = blockIdx.x, tid = threadIdx.x;
    SOps o = load_sops(sm, ngl, nq);
    double* nod = sm + sops_doubles(ngl, nq);   // [3][npts]: dpv, U, V
    const size_t nbase = (size_t)e * npts;
    for (int k = 0; k < nl; ++k) {
        __syncthreads();
        for (int t = tid; t < npts; t += blockDim.x) {
            nod[t] = a.dpv[(size_t)k * a.nstride + nbase + t];
            nod[npts + t] = a.qprime[(size_t)(1 * nl + k) * a.nstride + nbase + t] + a.ub_df[nbase + t];
            nod[2 * npts + t] = a.qprime[(size_t)(2 * nl + k) * a.nstride + nbase + t] + a.vb_df[nbase + t];
        }
        __syncthreads();
        for (int t = tid; t < 4 * nq; t += blockDim.x) {
            const int s = t / nq, iq = t - s * nq, q = face_quad(s, iq, nq), j = q / nq, i = q - j * nq;
            const Met mt = met_q(a.M, e, q);
            const double ksx = mt.ksx, ksy = mt.ksy, etx = mt.etx, ety = mt.ety;
            const double d = vq_interp_point(o, ngl, nod, i, j);
            double uk, ue, vk, ve;
            vq_grad_point(o, ngl, nod + npts, i, j, uk, ue);
            vq_grad_point(o, ngl, nod + 2 * npts, i, j, vk, ve);
            const size_t at = ((size_t)e * 4 + s) * nq + iq;
            a.trq[(size_t)(0 * nl + k) * a.trq_stride + at] = d * (ksx * uk + etx * ue);
            a.trq[(size_t)(1 * nl + k) * a.trq_stride + at] = d * (ksy * uk + ety * ue);
            a.trq[(size_t)(2 * nl + k) * a.trq_stride + at] = d * (ksx * vk + etx * ve);
            a.trq[(size_t)(3 * nl + k) * a.trq_stride + at] = d * (ksy * vk + ety * ve);
        }
    }
}
// pass 2: volume term + face flux, rhs_visc = visc massinv (lap + faces)
__global__ void k_bcl_lapq_apply(VqLayerArgs a) {
    extern __shared__ double sm[];
    const int ngl = a.M.ngl, nq = a.M.nq, npts = ngl * ngl, nq2 = nq * nq, per = ngl * nq, nl = a.M.nl;
    const int e = blockIdx.x, tid = threadIdx.x;
    SOps o = load_sops(sm, ngl, nq);
    double* nod = sm + sops_doubles(ngl, nq);   // [3][npts]
    double* tA = nod + 3 * npts;                // [3][per]; later scatter scratch
    double* tB = tA + 3 * per;                  // [3][per]
    double* vF = tB + 3 * per;                  // [4][nq2]: Fk_u Fk_v | Fe_u Fe_v (weighted, negative)
    double* lap = vF + 4 * nq2;                 // [2][npts]
    double* lfq = lap + 2 * npts;               // [4][2][nq]
    const size_t nbase = (size_t)e * npts;
    const Met mt = met_q(a.M, e, tid < nq2 ? tid : 0);   // metric terms at this thread's quadrature point
    const double ksx = mt.ksx, ksy = mt.ksy, etx = mt.etx, ety = mt.ety, J = mt.J;
    for (int k = 0; k < nl; ++k) {
        __syncthreads();
        for (int t = tid; t < npts; t += blockDim.x) {
            nod[t] = a.dpv[(size_t)k * a.nstride + nbase + t];
            nod[npts + t] = a.qprime[(size_t)(1 * nl + k) * a.nstride + nbase + t] + a.ub_df[nbase + t];
            nod[2 * npts + t] = a.qprime[(size_t)(2 * nl + k) * a.nstride + nbase + t] + a.vb_df[nbase + t];
        }
        __syncthreads();
        sf_pass1(o, ngl, nq, 3, nod, npts, tA, tB);
        for (int t = tid; t < 4 * nq; t += blockDim.x) {
            const int s = t / nq, iq = t - s * nq;
            double fl[2];
            vq_face_flux(a.M, a.trq + (size_t)k * a.trq_stride, (size_t)nl * a.trq_stride, e, s, iq, o.wq[iq], fl);
            lfq[(s * 2 + 0) * nq + iq] = fl[0]; lfq[(s * 2 + 1) * nq + iq] = fl[1];
        }
        __syncthreads();
        if (tid < nq2) {
            const int j = tid / nq, i = tid - j * nq;
            const double d = sf_eval(o, ngl, nq, tA, 0, i, j);
            const double uk = sf_eval(o, ngl, nq, tB, 1, i, j), ue = sf_eval_B(o, ngl, nq, tA, 1, i, j);
            const double vk = sf_eval(o, ngl, nq, tB, 2, i, j), ve = sf_eval_B(o, ngl, nq, tA, 2, i, j);
            const double F0 = d * (ksx * uk + etx * ue), F1 = d * (ksy * uk + ety * ue);
            const double F2 = d * (ksx * vk + etx * ve), F3 = d * (ksy * vk + ety * ve);
            const double wq = o.wq[i] * o.wq[j] * J;
            vF[0 * nq2 + tid] = -wq * (ksx * F0 + ksy * F1); vF[1 * nq2 + tid] = -wq * (ksx * F2 + ksy * F3);
            vF[2 * nq2 + tid] = -wq * (etx * F0 + ety * F1); vF[3 * nq2 + tid] = -wq * (etx * F2 + ety * F3);
        }
        __syncthreads();
        sf_scatter(o, ngl, nq, 2, nullptr, vF, vF + 2 * nq2, nq2, tA, tB, lap, npts, false);
        for (int t = tid; t < npts; t += blockDim.x) {
            const int m = t / ngl, n = t - m * ngl;
            double l0 = lap[t], l1 = lap[npts + t];
            for (int s = 0; s < 4; ++s) {
                const bool on = (s == 0) ? (m == 0) : (s == 1) ? (m == ngl - 1) : (s == 2) ? (n == 0) : (n == ngl - 1);
                if (!on) continue;
                const int nf = (s < 2) ? n : m;
                double p0 = 0.0, p1 = 0.0;
                for (int iq = 0; iq < nq; ++iq) {
                    const double hi = o.A[nf + ngl * iq];
                    p0 += hi * lfq[(s * 2 + 0) * nq + iq]; p1 += hi * lfq[(s * 2 + 1) * nq + iq];
                }
                l0 += p0; l1 += p1;
            }
            const double mi = a.massinv[nbase + t];
            a.rhs_visc[(size_t)(0 * nl + k) * a.nstride + nbase + t] = a.visc * mi * l0;
            a.rhs_visc[(size_t)(1 * nl + k) * a.nstride + nbase + t] = a.visc * mi * l1;
        }
    }
}

}  // namespace hn
