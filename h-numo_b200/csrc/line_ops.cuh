// line_ops.cuh -- 1-D line contractions of the sum-factorised operators, shared by the barotropic stage kernel
// (stage_pair.cuh) and the warp-per-element layer kernels (layer_warp.cuh).
//
// "One line per lane": a lane owns one line of an element (a nodal row, a column of quadrature points, a face) and
// contracts it with a 1-D operator whose entries are uniform-register operands (constant memory, compile-time indices).
// Loop order: contraction index outside, OUTPUT index inside -- consecutive FMAs belong to independent accumulators
// (an FP64 FMA has ~8 cycles of dependent-issue latency; the chain-serial order costs 4x the pipe time).
#pragma once
#include "hnumo_dev.cuh"

namespace hn {

__host__ __device__ constexpr int pr_pad2(int n) { return (n + 1) & ~1; }

// shared-memory strides of the per-warp tiles.  SX, ST, TM are padded so that in the line phases (lane = Q*f + i, or
// lane = G*f + m) the word index is congruent to the lane number modulo 16: no bank conflicts (profiles/smem_conflicts.py).
template <int G, int Q>
struct LineGeom {
    static constexpr int NP = G * G, NQ2 = Q * Q;
    static constexpr int pick(int lo, int mod) { return lo + ((mod - lo) % 16 + 16) % 16; }
    static constexpr int SX = pick(NQ2, Q);            // stride of the quadrature-point arrays X[f][j][i]
    static constexpr int TM = (Q <= 9) ? 11 : Q + 2;   // row stride of the pass-1 arrays T[f][m][i]
    static constexpr int ST = pick((G - 1) * TM + Q, Q);
};

// shared-memory vector of NE doubles
template <int NE> struct PV;
template <> struct PV<1> {
    typedef double T;
    static __device__ __forceinline__ void ld(const T* p, double (&o)[1]) { o[0] = *p; }
    static __device__ __forceinline__ void st(T* p, const double (&i)[1]) { *p = i[0]; }
};
template <> struct PV<2> {
    typedef double2 T;
    static __device__ __forceinline__ void ld(const T* p, double (&o)[2]) { double2 t = *p; o[0] = t.x; o[1] = t.y; }
    static __device__ __forceinline__ void st(T* p, const double (&i)[2]) { *p = make_double2(i[0], i[1]); }
};

#define PR_FORC _Pragma("unroll") for (int c = 0; c < NE; ++c)

// ---- line contractions on NE elements at once.  The operator entry is a uniform-register operand shared by the NE
// FMAs.  Loop order: contraction index outside, OUTPUT index inside -- consecutive FMAs belong to independent
// accumulators (an FP64 FMA has ~8 cycles of dependent-issue latency; the chain-serial order costs 4x the pipe time).
// out[i] = sum_n M(n,i) in[n], nodes -> quadrature points
// High orders (nop 7, 8: 8x15 and 9x17 operators): the fully unrolled contraction of one line is 150 FMAs with as many
// constant operands, and a stage kernel has a dozen of them inlined -- more code than the instruction cache holds (ncu, round 1:
// 1.07 issue slots per instruction lost to no_instruction).  There the contraction index is a real loop: the line value is
// loaded inside it and the operator row is addressed with the (warp-uniform) loop counter.
#ifndef HN_ROLL_FROM_G
#define HN_ROLL_FROM_G 99   // measured SLOWER at nop 8 (stage 5.70 vs 4.89 ms at 250 000 elements): kept for experiments only
#endif
template <int NE, int G, int Q, bool DERIV, int SS, int DS>
__device__ __forceinline__ void pl_n2q(const typename PV<NE>::T* src, typename PV<NE>::T* dst) {
    if constexpr (G >= HN_ROLL_FROM_G) {
        double s[Q][NE];
#pragma unroll
        for (int i = 0; i < Q; ++i) { PR_FORC s[i][c] = 0.0; }
        const double* op = DERIV ? c_ops.BT : c_ops.AT;
#pragma unroll 1
        for (int n = 0; n < G; ++n) {
            double v[NE];
            PV<NE>::ld(src + n * SS, v);
#pragma unroll
            for (int i = 0; i < Q; ++i) { PR_FORC s[i][c] = fma(op[i + Q * n], v[c], s[i][c]); }
        }
#pragma unroll
        for (int i = 0; i < Q; ++i) PV<NE>::st(dst + i * DS, s[i]);
        return;
    }
    double in[G][NE], s[Q][NE];
#pragma unroll
    for (int n = 0; n < G; ++n) PV<NE>::ld(src + n * SS, in[n]);
#pragma unroll
    for (int n = 0; n < G; ++n) {
#pragma unroll
        for (int i = 0; i < Q; ++i) {
            const double mv = DERIV ? c_ops.BT[i + Q * n] : c_ops.AT[i + Q * n];
            if (n == 0) { PR_FORC s[i][c] = mv * in[0][c]; }
            else { PR_FORC s[i][c] = fma(mv, in[n][c], s[i][c]); }
        }
    }
#pragma unroll
    for (int i = 0; i < Q; ++i) PV<NE>::st(dst + i * DS, s[i]);
}
// acc[n] (+)= sum_i M(n,i) in[i], quadrature points -> nodes (weak-form transpose)
template <int NE, int G, int Q, bool DERIV, int SS, bool FIRST>
__device__ __forceinline__ void pl_q2n_acc(const typename PV<NE>::T* src, double (&acc)[G][NE]) {
    if constexpr (G >= HN_ROLL_FROM_G) {
        if (FIRST) {
#pragma unroll
            for (int n = 0; n < G; ++n) { PR_FORC acc[n][c] = 0.0; }
        }
        const double* op = DERIV ? c_ops.B : c_ops.A;
#pragma unroll 1
        for (int i = 0; i < Q; ++i) {
            double v[NE];
            PV<NE>::ld(src + i * SS, v);
#pragma unroll
            for (int n = 0; n < G; ++n) { PR_FORC acc[n][c] = fma(op[n + G * i], v[c], acc[n][c]); }
        }
        return;
    }
    double in[Q][NE];
#pragma unroll
    for (int i = 0; i < Q; ++i) PV<NE>::ld(src + i * SS, in[i]);
#pragma unroll
    for (int i = 0; i < Q; ++i) {
#pragma unroll
        for (int n = 0; n < G; ++n) {
            const double mv = DERIV ? c_ops.B[n + G * i] : c_ops.A[n + G * i];
            if (FIRST && i == 0) { PR_FORC acc[n][c] = mv * in[0][c]; }
            else { PR_FORC acc[n][c] = fma(mv, in[i][c], acc[n][c]); }
        }
    }
}
// collocation derivative along a nodal line (runtime stride): out[n] = sum_k D(k,n) in[k]  or transposed D(n,k)
template <int NE, int G, bool TRANSP>
__device__ __forceinline__ void pl_grad(const typename PV<NE>::T* src, typename PV<NE>::T* dst, int stride) {
    double in[G][NE], s[G][NE];
#pragma unroll
    for (int k = 0; k < G; ++k) PV<NE>::ld(src + k * stride, in[k]);
#pragma unroll
    for (int k = 0; k < G; ++k) {
#pragma unroll
        for (int n = 0; n < G; ++n) {
            const double mv = TRANSP ? c_ops.D[n + G * k] : c_ops.DT[n + G * k];
            if (k == 0) { PR_FORC s[n][c] = mv * in[0][c]; }
            else { PR_FORC s[n][c] = fma(mv, in[k][c], s[n][c]); }
        }
    }
#pragma unroll
    for (int n = 0; n < G; ++n) PV<NE>::st(dst + n * stride, s[n]);
}

}  // namespace hn
