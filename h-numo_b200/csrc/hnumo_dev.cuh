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
