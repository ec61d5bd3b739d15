// diag.cuh -- device-side diagnostics of the resident state: the numbers the reference prints every
// output step are reduced on the GPU, so that a run needs hnumo_download_state only at output cadence.
//   layer fields h, u, v, dp, interface elevation          src/diagnostics.F90:24-45
//   layer mass  sum_I wjac_df(I) h(I)                       src/compute_conserved.F90:29-41 (psih_df is the identity at the nodes)
//   max / min per layer and of qb(1:4,:)                     src/print_diagnostics.F90:59-128
//   Courant numbers on the sub-cells of the LGL grid         src/courant.F90:34-126
// As written in the reference, the barotropic Courant number is formed with qb(3:4,:) = p_b*ubar, p_b*vbar (momentum, not
// velocity: courant.F90:88-89 with the qb_df slice passed by mod_time_loop.F90:184); it is restated as written.
// One deviation: the reference divides by the running minimum of dx, dy over the cells visited so far (courant.F90:96-99);
// here the minimum over all local cells is taken first (identical after the first element on bricks).
#pragma once
#include "hnumo_dev.cuh"

namespace hn {

// per layer: 0 mass | 1..5 max h u v dp elev | 6..10 min h u v dp elev ; tail: qb max(4) qb min(4) cfl_b cfl min_dx min_dy
enum { DIAG_PER_LAYER = 11, DIAG_TAIL = 12 };

struct DiagArgs {
    Mesh M;
    const double* q;        // planes [v*nl + k]
    size_t nstride;
    const double* qb[3];    // pbpert, pbub, pbvb
    const double* pbprime_df;
    const double* zbot_df;
    const double* massinv;
    double alpha_over_g[HN_MAXL];
    double dt, dt_btp;
    double* partial;        // [nvals][nelem]
    int nvals;
};

__device__ __forceinline__ double dg_warp_red(double v, int op) {   // op 0 sum, 1 max, 2 min
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_xor_sync(0xffffffffu, v, o);
        v = op == 0 ? v + w : op == 1 ? fmax(v, w) : fmin(v, w);
    }
    return v;
}
// block reduction of one value (all threads call); the result is valid in thread 0.  red: [warps] scratch.
__device__ __forceinline__ double dg_block_red(double v, int op, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = dg_warp_red(v, op);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double w = lane < nw ? red[lane] : (op == 0 ? 0.0 : op == 1 ? -1.0e300 : 1.0e300);
        v = dg_warp_red(w, op);
    }
    return v;
}

// one block per element: nodal thread (m, n), blockDim >= npts
__global__ void k_diag_partial(DiagArgs a) {
    __shared__ double red[32];
    extern __shared__ double sm[];   // [nl][2][npts] layer velocities u, v; [2][npts] barotropic momentum
    const int ngl = a.M.ngl, npts = a.M.npts, nl = a.M.nl;
    const int e = blockIdx.x, tid = threadIdx.x;
    const bool act = tid < npts;
    const size_t I = (size_t)e * npts + (act ? tid : 0);
    double* out = a.partial + e;             // value j of this element at out[j * nelem]
    const size_t ne = (size_t)a.M.nelem;
    const double big = 1.0e300;
    const double wj = act ? 1.0 / a.massinv[I] : 0.0;
    double elev = act ? a.zbot_df[I] : 0.0;
    for (int k = nl - 1; k >= 0; --k) {
        double dp = 1.0, h = 0.0, u = 0.0, v = 0.0;
        if (act) {
            dp = a.q[(size_t)(0 * nl + k) * a.nstride + I];
            h = a.alpha_over_g[k] * dp;
            u = a.q[(size_t)(1 * nl + k) * a.nstride + I] / dp;
            v = a.q[(size_t)(2 * nl + k) * a.nstride + I] / dp;
            elev += h;
            sm[(k * 2 + 0) * npts + tid] = u; sm[(k * 2 + 1) * npts + tid] = v;
        }
        const double vals[5] = {h, u, v, dp, elev};
        double r = dg_block_red(act ? wj * h : 0.0, 0, red);
        if (tid == 0) out[(size_t)(k * DIAG_PER_LAYER) * ne] = r;
#pragma unroll
        for (int f = 0; f < 5; ++f) {
            r = dg_block_red(act ? vals[f] : -big, 1, red);
            if (tid == 0) out[(size_t)(k * DIAG_PER_LAYER + 1 + f) * ne] = r;
            r = dg_block_red(act ? vals[f] : big, 2, red);
            if (tid == 0) out[(size_t)(k * DIAG_PER_LAYER + 6 + f) * ne] = r;
        }
    }
    double* tail = out + (size_t)(nl * DIAG_PER_LAYER) * ne;
    {
        double qbv[4] = {0, 0, 0, 0};
        if (act) {
            qbv[1] = a.qb[0][I]; qbv[0] = qbv[1] + a.pbprime_df[I]; qbv[2] = a.qb[1][I]; qbv[3] = a.qb[2][I];
            sm[(nl * 2 + 0) * npts + tid] = qbv[2]; sm[(nl * 2 + 1) * npts + tid] = qbv[3];
        }
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            double r = dg_block_red(act ? qbv[f] : -big, 1, red);
            if (tid == 0) tail[(size_t)f * ne] = r;
            r = dg_block_red(act ? qbv[f] : big, 2, red);
            if (tid == 0) tail[(size_t)(4 + f) * ne] = r;
        }
    }
    __syncthreads();
    // sub-cells (i, j), i, j < ngl-1: mean of the 4 corner values; cell sizes from the affine element map
    const int m = tid / ngl, n = tid - m * ngl;
    const bool cell = act && m < ngl - 1 && n < ngl - 1;
    double dx = big, dy = big, cb_x = 0.0, cb_y = 0.0, c_x = 0.0, c_y = 0.0;
    if (cell) {
        const int c0 = tid, c1 = tid + 1, c2 = tid + ngl, c3 = tid + ngl + 1;
        if (a.M.coord) {   // general quadrilaterals: extent of the four corner nodes, as courant_cube_mlswe takes it from coord (courant.F90:83-97)
            const double* cx = a.M.coord + (size_t)e * npts; const double* cy = cx + (size_t)a.M.npoin;
            dx = fmax(fmax(cx[c0], cx[c1]), fmax(cx[c2], cx[c3])) - fmin(fmin(cx[c0], cx[c1]), fmin(cx[c2], cx[c3]));
            dy = fmax(fmax(cy[c0], cy[c1]), fmax(cy[c2], cy[c3])) - fmin(fmin(cy[c0], cy[c1]), fmin(cy[c2], cy[c3]));
        } else {
            const double ksx = a.M.em[e * 5 + 0], ksy = a.M.em[e * 5 + 1], etx = a.M.em[e * 5 + 2], ety = a.M.em[e * 5 + 3];
            const double det = ksx * ety - ksy * etx;
            const double dks = c_ops.xg[n + 1] - c_ops.xg[n], det_ = c_ops.xg[m + 1] - c_ops.xg[m];
            dx = (fabs(ety) * dks + fabs(ksy) * det_) / fabs(det);
            dy = (fabs(etx) * dks + fabs(ksx) * det_) / fabs(det);
        }
        const double* bu = sm + (nl * 2 + 0) * npts; const double* bv = sm + (nl * 2 + 1) * npts;
        cb_x = fabs(bu[c0] / 4 + bu[c1] / 4 + bu[c2] / 4 + bu[c3] / 4);
        cb_y = fabs(bv[c0] / 4 + bv[c1] / 4 + bv[c2] / 4 + bv[c3] / 4);
        for (int k = 0; k < nl; ++k) {
            const double* uk = sm + (k * 2 + 0) * npts; const double* vk = sm + (k * 2 + 1) * npts;
            c_x = fmax(c_x, fabs(uk[c0] / 4 + uk[c1] / 4 + uk[c2] / 4 + uk[c3] / 4));
            c_y = fmax(c_y, fabs(vk[c0] / 4 + vk[c1] / 4 + vk[c2] / 4 + vk[c3] / 4));
        }
    }
    double r;
    r = dg_block_red(cb_x, 1, red); if (tid == 0) tail[(size_t)8 * ne] = r;    // max |mean p_b u| (divided by min dx in the second pass)
    r = dg_block_red(cb_y, 1, red); if (tid == 0) tail[(size_t)9 * ne] = r;
    r = dg_block_red(dx, 2, red);   if (tid == 0) tail[(size_t)10 * ne] = r;
    r = dg_block_red(dy, 2, red);   if (tid == 0) tail[(size_t)11 * ne] = r;
    r = dg_block_red(c_x, 1, red);  if (tid == 0) tail[(size_t)12 * ne] = r;
    r = dg_block_red(c_y, 1, red);  if (tid == 0) tail[(size_t)13 * ne] = r;
}

// second pass: block j reduces value j over the elements in a fixed order (deterministic); the Courant numbers are formed on
// the host from the six reduced tail values
__global__ void k_diag_final(const double* partial, int nelem, int nl, double* res) {
    __shared__ double red[32];
    const int tid = threadIdx.x, j = blockIdx.x;
    int op;
    if (j < nl * DIAG_PER_LAYER) { const int f = j % DIAG_PER_LAYER; op = f == 0 ? 0 : f <= 5 ? 1 : 2; }
    else { const int t = j - nl * DIAG_PER_LAYER; op = t < 4 ? 1 : t < 8 ? 2 : (t == 10 || t == 11) ? 2 : 1; }
    double v = op == 0 ? 0.0 : op == 1 ? -1.0e300 : 1.0e300;
    const double* p = partial + (size_t)j * nelem;
    for (int e = tid; e < nelem; e += blockDim.x) {
        const double w = p[e];
        v = op == 0 ? v + w : op == 1 ? fmax(v, w) : fmin(v, w);
    }
    v = dg_block_red(v, op, red);
    if (tid == 0) res[j] = v;
}

}  // namespace hn
