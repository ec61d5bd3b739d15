// hnumo_dev.cuh -- device helpers shared by all kernels (one translation unit).
#pragma once
#include "hnumo_internal.cuh"

namespace hn {

__constant__ Ops c_ops;

// nodal index of face node n on slot s; quad index of face quadrature point iq on slot s
__device__ __forceinline__ int face_node(int s, int n, int ngl) {
    return s == 0 ? n : s == 1 ? (ngl - 1) * ngl + n : s == 2 ? n * ngl : n * ngl + ngl - 1;
}
__device__ __forceinline__ int face_quad(int s, int iq, int nq) {
    return s == 0 ? iq : s == 1 ? (nq - 1) * nq + iq : s == 2 ? iq * nq : iq * nq + nq - 1;
}

// ---- geometry at a point.  Affine elements (bricks, parallelograms): five numbers per element and three per side.  General
// quadrilaterals (isoparametric, src/metrics.F90, src/metrics_quad.F90, src/create_normals(_quad).F90): the same numbers per
// quadrature point / node / face point.  Every use of the metric terms in the run-time-size kernels is pointwise, so the two cases
// differ only in where the numbers come from.
struct Met { double ksx, ksy, etx, ety, J; };
struct FGeo { double nx, ny, len; };
__device__ __forceinline__ Met met_elem(const Mesh& M, int e) {
    const double* p = M.em + (size_t)e * 5;
    Met m; m.ksx = p[0]; m.ksy = p[1]; m.etx = p[2]; m.ety = p[3]; m.J = p[4];
    return m;
}
// quadrature point q = j*nq + i of element e
__device__ __forceinline__ Met met_q(const Mesh& M, int e, int q) {
    if (!M.mq) return met_elem(M, e);
    const size_t I = (size_t)e * M.nq2 + q, S = (size_t)M.npoin_q;
    Met m; m.ksx = M.mq[I]; m.ksy = M.mq[S + I]; m.etx = M.mq[2 * S + I]; m.ety = M.mq[3 * S + I]; m.J = M.mq[4 * S + I];
    return m;
}
// node I = m*ngl + n of element e
__device__ __forceinline__ Met met_n(const Mesh& M, int e, int I) {
    if (!M.mn) return met_elem(M, e);
    const size_t P = (size_t)e * M.npts + I, S = (size_t)M.npoin;
    Met m; m.ksx = M.mn[P]; m.ksy = M.mn[S + P]; m.etx = M.mn[2 * S + P]; m.ety = M.mn[3 * S + P]; m.J = M.mn[4 * S + P];
    return m;
}
// canonical (left-element) unit normal and edge Jacobian at face quadrature point iq / face node n of element side `slot`
__device__ __forceinline__ FGeo fg_q(const Mesh& M, int slot, int iq) {
    FGeo g;
    if (!M.fgq) { g.nx = M.fgeom[slot * 3 + 0]; g.ny = M.fgeom[slot * 3 + 1]; g.len = M.fgeom[slot * 3 + 2]; return g; }
    const size_t I = (size_t)slot * M.nq + iq, S = (size_t)M.nslots * M.nq;
    g.nx = M.fgq[I]; g.ny = M.fgq[S + I]; g.len = M.fgq[2 * S + I];
    return g;
}
__device__ __forceinline__ FGeo fg_n(const Mesh& M, int slot, int n) {
    FGeo g;
    if (!M.fgn) { g.nx = M.fgeom[slot * 3 + 0]; g.ny = M.fgeom[slot * 3 + 1]; g.len = M.fgeom[slot * 3 + 2]; return g; }
    const size_t I = (size_t)slot * M.ngl + n, S = (size_t)M.nslots * M.ngl;
    g.nx = M.fgn[I]; g.ny = M.fgn[S + I]; g.len = M.fgn[2 * S + I];
    return g;
}

// Operator tables staged in shared memory by every "simple" kernel.
struct SOps {
    const double *A, *B, *D, *wq, *wg;
};
__device__ __forceinline__ int sops_doubles(int ngl, int nq) { return 2 * ngl * nq + ngl * ngl + nq + ngl; }
__device__ __forceinline__ SOps load_sops(double* s, int ngl, int nq) {
    double* A = s; double* B = A + ngl * nq; double* D = B + ngl * nq; double* wq = D + ngl * ngl; double* wg = wq + nq;
    for (int t = threadIdx.x; t < ngl * nq; t += blockDim.x) { A[t] = c_ops.A[t]; B[t] = c_ops.B[t]; }
    for (int t = threadIdx.x; t < ngl * ngl; t += blockDim.x) D[t] = c_ops.D[t];
    for (int t = threadIdx.x; t < nq; t += blockDim.x) wq[t] = c_ops.wq[t];
    for (int t = threadIdx.x; t < ngl; t += blockDim.x) wg[t] = c_ops.wg[t];
    SOps o; o.A = A; o.B = B; o.D = D; o.wq = wq; o.wg = wg;
    return o;
}

// first sum-factorisation pass for NF nodal fields held in shared memory:
//   tA[f][m][i] = sum_n A(n,i) nod[f][m][n]      (and tB with B if tB != nullptr)
__device__ __forceinline__ void sf_pass1(const SOps& o, int ngl, int nq, int NF, const double* nod, int nod_stride, double* tA,
                                double* tB) {
    const int per = ngl * nq;
    for (int t = threadIdx.x; t < NF * per; t += blockDim.x) {
        int f = t / per, r = t - f * per, m = r / nq, i = r - m * nq;
        const double* row = nod + f * nod_stride + m * ngl;
        double a = 0.0, b = 0.0;
        for (int n = 0; n < ngl; ++n) {
            a += o.A[n + ngl * i] * row[n];
            if (tB) b += o.B[n + ngl * i] * row[n];
        }
        tA[t] = a;
        if (tB) tB[t] = b;
    }
}
// second pass: value of field f at quadrature point (i,j)
__device__ __forceinline__ double sf_eval(const SOps& o, int ngl, int nq, const double* tA, int f, int i, int j) {
    const double* c = tA + f * ngl * nq + i;
    double v = 0.0;
    for (int m = 0; m < ngl; ++m) v += o.A[m + ngl * j] * c[m * nq];
    return v;
}
// with derivative in eta: sum_m B(m,j) tA[f][m][i]
__device__ __forceinline__ double sf_eval_B(const SOps& o, int ngl, int nq, const double* tA, int f, int i, int j) {
    const double* c = tA + f * ngl * nq + i;
    double v = 0.0;
    for (int m = 0; m < ngl; ++m) v += o.B[m + ngl * j] * c[m * nq];
    return v;
}

// Weak-form scatter of quadrature-point data to the nodes of one element, sum-factorised:
//   out[f][m][n] (+)= sum_ij  psi_nm(i,j) S[f][ij] + dpsi/dksi Fk[f][ij] + dpsi/deta Fe[f][ij]
// S, Fk, Fe already contain the quadrature weight and metric factors.  Any of S may be null.
// tP,tR: scratch [NF][ngl*nq].  Caller must __syncthreads() before (inputs ready) -- this routine
// syncs internally between the passes and after the result is written.
__device__ __forceinline__ void sf_scatter(const SOps& o, int ngl, int nq, int NF, const double* S, const double* Fk,
                                  const double* Fe, int qstride, double* tP, double* tR, double* out, int out_stride,
                                  bool accumulate) {
    const int per = ngl * nq;
    for (int t = threadIdx.x; t < NF * per; t += blockDim.x) {
        int f = t / per, r = t - f * per, m = r / nq, i = r - m * nq;
        double P = 0.0, R = 0.0;
        for (int j = 0; j < nq; ++j) {
            double a = o.A[m + ngl * j], b = o.B[m + ngl * j];
            P += a * Fk[f * qstride + j * nq + i];
            R += b * Fe[f * qstride + j * nq + i];
            if (S) R += a * S[f * qstride + j * nq + i];
        }
        tP[t] = P; tR[t] = R;
    }
    __syncthreads();
    const int npts = ngl * ngl;
    for (int t = threadIdx.x; t < NF * npts; t += blockDim.x) {
        int f = t / npts, r = t - f * npts, m = r / ngl, n = r - m * ngl;
        double v = 0.0;
        for (int i = 0; i < nq; ++i)
            v += o.B[n + ngl * i] * tP[f * per + m * nq + i] + o.A[n + ngl * i] * tR[f * per + m * nq + i];
        if (accumulate) out[f * out_stride + r] += v; else out[f * out_stride + r] = v;
    }
    __syncthreads();
}

// nodal (collocation) gradient of a field held in shared memory at node (n,m):
//   d/dksi = sum_k D(k,n) f[m][k],  d/deta = sum_k D(k,m) f[k][n]      (compute_gradient_uv, mod_barotropic_terms.F90:411)
__device__ __forceinline__ void nodal_grad(const SOps& o, int ngl, const double* f, int n, int m, double& dk, double& de) {
    dk = 0.0; de = 0.0;
    for (int k = 0; k < ngl; ++k) {
        dk += o.D[k + ngl * n] * f[m * ngl + k];
        de += o.D[k + ngl * m] * f[k * ngl + n];
    }
}

}  // namespace hn
