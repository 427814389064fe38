// Scan (semiseparable) form of the B0 family: cell-integrated Matern-1/2 features of
// gridded_kronecker_structure.py:1325-1374 without ever forming a dense feature column.  DESIGN.md section 10,
// algebra pinned by oracle/b0_scan.py + tests/test_b0_scan_form.py.
//
// For x with containing cell c (c = -1 left of the mesh, c = K-1 right of it; e = c + 1 indexes the K + 1 "extended"
// cells; cells i = 0..K-2, knots t_0..t_{K-1}):
//     phi_i(x) = fL(x) GL[e][i]  (i < c),   fC(x)  (i == c),   fR(x) GR[e][i]  (i > c)
//     fL = s2 l exp(-(x - t_c) / l),   fR = s2 l exp(-(t_{c+1} - x) / l),   fC = 2 s2 l - fL - fR
//     GL[e][i] = exp(-(t_c - t_{i+1}) / l) - exp(-(t_c - t_i) / l),   GR[e][i] = exp(-(t_i - t_{c+1}) / l) - exp(-(t_{i+1} - t_{c+1}) / l)
// so that   mu = sum_{X,Y} f1^X f2^Y T^{XY}[e1][e2],  T^{XY} = G1^X A G2^Y^T      (X, Y in {L, C, R}, G^C[e] = unit row c)
//           p_d = sum_{X,Y} f^X f^Y W^{XY}[e],        W^{XY}[e] = (G^X P_d G^Y^T)[e][e]   (q_d likewise from Q_d).
// GL and GR are rank-one decay (semiseparable) matrices, so the tables are first-order recurrences along the modes on the
// grid side (k_b0s_scan, O(M) per mode); the per-point work is O(1).
#pragma once
#include "obs.cuh"
#include "obs_binned.cuh"

namespace vggp {

constexpr int B0S_L = 0, B0S_C = 1, B0S_R = 2;
// order of the 6 distinct entries of the symmetric 3 x 3 block W: LL, LC, LR, CC, CR, RR
__device__ __forceinline__ int b0s_sym_index(int X, int Y) {
    const int a = X < Y ? X : Y, b = X < Y ? Y : X;
    return a == 0 ? b : (a == 1 ? 2 + b : 5);
}

// GL, GR and their lengthscale derivatives for dimension blockIdx.y: (K + 1) x (K - 1) row-major each.
// grid (ceil((K+1)(K-1) / 256), D)
struct B0sGArgs {
    const float* knots[VGGP_MAX_D];
    int K[VGGP_MAX_D];
    double* GL[VGGP_MAX_D];
    double* GR[VGGP_MAX_D];
    double* dGL[VGGP_MAX_D];
    double* dGR[VGGP_MAX_D];
    const double* theta;
};

__global__ void __launch_bounds__(256) k_b0s_G(const __grid_constant__ B0sGArgs a) {
    const int d = blockIdx.y;
    const int K = a.K[d], M = K - 1, E = K + 1;
    const i64 idx = (i64)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (i64)E * M) return;
    const int e = (int)(idx / M), i = (int)(idx % M), c = e - 1;
    const float* t = a.knots[d];
    const double l = a.theta[d];
    const double ti = (double)t[i], ti1 = (double)t[i + 1];
    double gl = 0.0, gr = 0.0, dgl = 0.0, dgr = 0.0;
    if (i < c) {
        const double tc = (double)t[c < K - 1 ? c : K - 1];
        const double a1 = tc - ti1, a2 = tc - ti;                 // 0 <= a1 < a2
        const double e1 = exp(-a1 / l), e2 = exp(-a2 / l);
        gl = e1 - e2;
        dgl = (a1 * e1 - a2 * e2) / (l * l);
    } else if (i > c) {
        const double tc1 = (double)t[c + 1 > 0 ? c + 1 : 0];
        const double a1 = ti - tc1, a2 = ti1 - tc1;               // 0 <= a1 < a2
        const double e1 = exp(-a1 / l), e2 = exp(-a2 / l);
        gr = e1 - e2;
        dgr = (a1 * e1 - a2 * e2) / (l * l);
    }
    a.GL[d][idx] = gl;
    a.GR[d][idx] = gr;
    a.dGL[d][idx] = dgl;
    a.dGR[d][idx] = dgr;
}

// Per-cell decay factors of dimension blockIdx.y: eps_c = exp(-h_c / l), gam_c = 1 - eps_c (via expm1), d eps_c / d l.
// GL and GR are rank-one decay (semiseparable) matrices: moving the position by one cell multiplies everything
// accumulated so far by eps_c and adds gam_c times the new cell, so a product G^X v is a first-order recurrence.
// out[d] = [eps (M) | gam (M) | deps (M)].  grid (ceil(Mmax / 256), D)
struct B0sEpsArgs {
    const float* knots[VGGP_MAX_D];
    int K[VGGP_MAX_D];
    double* out[VGGP_MAX_D];
    const double* theta;
};
__global__ void __launch_bounds__(256) k_b0s_eps(const __grid_constant__ B0sEpsArgs a) {
    const int d = blockIdx.y, M = a.K[d] - 1;
    const int c = (int)blockIdx.x * 256 + threadIdx.x;
    if (c >= M) return;
    const double l = a.theta[d];
    const double h = (double)a.knots[d][c + 1] - (double)a.knots[d][c];
    const double e = exp(-h / l);
    a.out[d][c] = e;
    a.out[d][M + c] = -expm1(-h / l);
    a.out[d][2 * M + c] = e * h / (l * l);
}

// L- and R-transform of every fibre of a tensor along one mode, one thread per fibre, two sequential sweeps:
//   L[e+1] = eps_{e-1} L[e] + gam_{e-1} v_{e-1}  (e = 1..M, L[0] = L[1] = 0),   R[e-1] = gam_{e-1} v_{e-1} + eps_{e-1} R[e]  (e = M..1, R[M] = R[M+1] = 0)
// i.e. dstL = G^L v, dstR = G^R v with E = M + 2 entries per fibre.  Element i of fibre f of the source is at
// src[f_hi * s_hi + f_lo * s_lo + i * s_mode] with f = f_hi * n_lo + f_lo (same decomposition for the destination).
struct B0sScanArgs {
    const double* src;
    double* dstL;
    double* dstR;
    const double* eps;       // [eps | gam | deps] of the mode's dimension
    int M;
    i64 n_fibres, n_lo;
    i64 s_hi, s_lo, s_mode;  // source strides
    i64 d_hi, d_lo, d_mode;  // destination strides
    double* tanL;            // optional: d dstL / d l and d dstR / d l (same layout as the destination), else nullptr
    double* tanR;
};
__global__ void __launch_bounds__(128) k_b0s_scan(const __grid_constant__ B0sScanArgs a) {
    const i64 f = (i64)blockIdx.x * 128 + threadIdx.x;
    if (f >= a.n_fibres) return;
    const i64 fh = f / a.n_lo, fl = f - fh * a.n_lo;
    const double* __restrict__ v = a.src + fh * a.s_hi + fl * a.s_lo;
    double* __restrict__ L = a.dstL + fh * a.d_hi + fl * a.d_lo;
    double* __restrict__ R = a.dstR + fh * a.d_hi + fl * a.d_lo;
    const double* __restrict__ eps = a.eps;
    const double* __restrict__ gam = a.eps + a.M;
    const int M = a.M;
    const double* __restrict__ deps = a.eps + 2 * a.M;
    double* __restrict__ TL = a.tanL ? a.tanL + fh * a.d_hi + fl * a.d_lo : nullptr;
    double* __restrict__ TR = a.tanR ? a.tanR + fh * a.d_hi + fl * a.d_lo : nullptr;
    // tangent: d acc' = eps d acc + deps (acc - v)   (gam = 1 - eps)
    double acc = 0.0, tan = 0.0;
    L[0] = 0.0;
    L[a.d_mode] = 0.0;
    if (TL) { TL[0] = 0.0; TL[a.d_mode] = 0.0; }
    for (int e = 1; e <= M; ++e) {
        const double vv = v[(i64)(e - 1) * a.s_mode];
        tan = fma(eps[e - 1], tan, deps[e - 1] * (acc - vv));
        acc = fma(eps[e - 1], acc, gam[e - 1] * vv);
        L[(i64)(e + 1) * a.d_mode] = acc;
        if (TL) TL[(i64)(e + 1) * a.d_mode] = tan;
    }
    acc = 0.0;
    tan = 0.0;
    R[(i64)(M + 1) * a.d_mode] = 0.0;
    R[(i64)M * a.d_mode] = 0.0;
    if (TR) { TR[(i64)(M + 1) * a.d_mode] = 0.0; TR[(i64)M * a.d_mode] = 0.0; }
    for (int e = M; e >= 1; --e) {
        const double vv = v[(i64)(e - 1) * a.s_mode];
        tan = fma(eps[e - 1], tan, deps[e - 1] * (acc - vv));
        acc = fma(eps[e - 1], acc, gam[e - 1] * vv);
        R[(i64)(e - 1) * a.d_mode] = acc;
        if (TR) TR[(i64)(e - 1) * a.d_mode] = tan;
    }
}

// Adjoint of the three transforms along one mode, one thread per fibre:
//   dv[i] = sum_e GL[e][i] gL[e] + GR[e][i] gR[e]  +  gC[i + 1]           (gC may be nullptr)
// two sequential sweeps: a[e] = gL[e] + eps_{e-1} a[e+1] (e = M..1), dv[e-1] += gam_{e-1} a[e+1];
//                        b[e] = gR[e] + eps_{e-1} b[e-1] (e = 1..M), dv[e-1] += gam_{e-1} b[e-1].
// gL / gC / gR share the layout (g_hi, g_lo, g_mode) with E = M + 2 entries per fibre; dv has M entries per fibre.
struct B0sAdjArgs {
    const double* gL;
    const double* gC;
    const double* gR;
    double* dv;
    const double* eps;
    int M;
    i64 n_fibres, n_lo;
    i64 g_hi, g_lo, g_mode;
    i64 v_hi, v_lo, v_mode;
};
__global__ void __launch_bounds__(128) k_b0s_scan_adj(const __grid_constant__ B0sAdjArgs a) {
    const i64 f = (i64)blockIdx.x * 128 + threadIdx.x;
    if (f >= a.n_fibres) return;
    const i64 fh = f / a.n_lo, fl = f - fh * a.n_lo;
    const i64 go = fh * a.g_hi + fl * a.g_lo;
    const double* __restrict__ gL = a.gL + go;
    const double* __restrict__ gR = a.gR + go;
    double* __restrict__ dv = a.dv + fh * a.v_hi + fl * a.v_lo;
    const double* __restrict__ eps = a.eps;
    const double* __restrict__ gam = a.eps + a.M;
    const int M = a.M;
    double carry = gL[(i64)(M + 1) * a.g_mode];               // a[M+1]
    for (int e = M; e >= 1; --e) {
        dv[(i64)(e - 1) * a.v_mode] = gam[e - 1] * carry + (a.gC ? a.gC[go + (i64)e * a.g_mode] : 0.0);
        carry = fma(eps[e - 1], carry, gL[(i64)e * a.g_mode]);
    }
    carry = gR[0];                                            // b[0]
    for (int e = 1; e <= M; ++e) {
        dv[(i64)(e - 1) * a.v_mode] += gam[e - 1] * carry;
        carry = fma(eps[e - 1], carry, gR[(i64)e * a.g_mode]);
    }
}

// ---- segmented forms of the two sweeps kernels above (the ones the library launches) --------------------------------------
// One thread per fibre runs 2 M dependent steps on 4 - 5 CTAs (0.1 - 0.35 ms per launch at M = 511).  Every sweep is a
// first-order affine recurrence  state' = eps_i state + add_i  whose multiplier does not depend on the fibre, so a fibre is
// cut into S = ceil(M / B0S_SEG) segments, one thread each:
//   pass 1   the segment's composite map from a zero state (B, with the tangent also T) -> shared memory; the multiplier
//            product of a segment, a dual number (A, dA) with the tangent, is the same for every fibre: once per CTA;
//   carry    every thread folds the maps of the segments before it (<= S - 1 steps);
//   pass 2   the segment again from its true carry-in, writing the outputs.
// The thread keeps its B0S_SEG source values in registers (all loads in flight before the first dependent step), the
// coefficient rows [eps | gam | deps] sit in shared memory with one double of skew per segment (lanes of different
// segments hit different banks).  2 B0S_SEG + S dependent steps per sweep instead of M; F = B0S_SEG_THREADS / S fibres per CTA.
// Thread mapping: fibres fastest when neighbouring fibres are contiguous in memory (column fibres), segments fastest when
// the mode itself is contiguous (row fibres) -- either way a warp's accesses fall into few lines.
constexpr int B0S_SEG = 16;
constexpr int B0S_SEG_THREADS = 128;
__host__ __device__ __forceinline__ int b0s_skew(int i) { return i + i / B0S_SEG; }
inline size_t b0s_seg_smem_bytes(int M, bool tan) {
    const int S = (M + B0S_SEG - 1) / B0S_SEG;
    (void)tan;
    return sizeof(double) * ((size_t)3 * (M + S + 1) + 2 * (size_t)S + 2 * (size_t)B0S_SEG_THREADS);
}

// Several independent transforms in one launch (blockIdx.y selects the entry): the 22 sweeps of a step are 128-CTA kernels
// of ~20 us each, too small to fill the GPU one at a time.
constexpr int B0S_MAX_BATCH = 6;
struct B0sSegGeo { int S, F, f_fast, blocks; };
struct B0sScanBatch { B0sScanArgs a[B0S_MAX_BATCH]; B0sSegGeo g[B0S_MAX_BATCH]; };
struct B0sAdjBatch { B0sAdjArgs a[B0S_MAX_BATCH]; B0sSegGeo g[B0S_MAX_BATCH]; };

template <bool TAN, int dir>
__device__ __forceinline__ void b0s_seg_sweep(const B0sScanArgs& a, const double (&vv)[B0S_SEG], const double* ceps, const double* cgam,
                                          const double* cdeps, const double* cA, const double* cdA, double* maps, int S, int fl,
                                          int s, int slot, int i0, int len, bool live, i64 dbase) {
    const int M = a.M;
    // pass 1: composite map of the segment from a zero state
    double acc = 0.0, tan = 0.0;
#pragma unroll
    for (int kk = 0; kk < B0S_SEG; ++kk) {
        const int k = dir ? B0S_SEG - 1 - kk : kk;
        if (k < len) {
            const double e = ceps[b0s_skew(i0 + k)];
            if (TAN) tan = fma(e, tan, cdeps[b0s_skew(i0 + k)] * (acc - vv[k]));
            acc = fma(e, acc, cgam[b0s_skew(i0 + k)] * vv[k]);
        }
    }
    if (dir) __syncthreads();             // the maps of the L sweep have been consumed
    if (live) {
        maps[slot] = acc;
        if (TAN) maps[B0S_SEG_THREADS + slot] = tan;
    }
    __syncthreads();
    // carry-in: fold the segments before this one in sweep order
    double cin = 0.0, tin = 0.0;
    if (live) {
        const int qa = dir ? S - 1 : 0, qstep = dir ? -1 : 1;
        for (int q = qa; q != s; q += qstep) {
            if (TAN) tin = fma(cA[q], tin, fma(cdA[q], cin, maps[B0S_SEG_THREADS + fl * S + q]));
            cin = fma(cA[q], cin, maps[fl * S + q]);
        }
    }
    // pass 2: the segment from its true carry-in
    double* __restrict__ O = (dir ? a.dstR : a.dstL) + dbase;
    double* __restrict__ TO = TAN ? (dir ? a.tanR : a.tanL) + dbase : nullptr;
    acc = cin;
    tan = tin;
    if (live && s == (dir ? S - 1 : 0)) {  // the two zero entries at the start of the sweep
        const i64 z0 = dir ? (i64)(M + 1) * a.d_mode : 0, z1 = dir ? (i64)M * a.d_mode : a.d_mode;
        O[z0] = 0.0; O[z1] = 0.0;
        if (TAN) { TO[z0] = 0.0; TO[z1] = 0.0; }
    }
#pragma unroll
    for (int kk = 0; kk < B0S_SEG; ++kk) {
        const int k = dir ? B0S_SEG - 1 - kk : kk;
        if (k < len) {
            const double e = ceps[b0s_skew(i0 + k)];
            if (TAN) tan = fma(e, tan, cdeps[b0s_skew(i0 + k)] * (acc - vv[k]));
            acc = fma(e, acc, cgam[b0s_skew(i0 + k)] * vv[k]);
            const i64 o = (i64)(dir ? i0 + k : i0 + k + 2) * a.d_mode;
            O[o] = acc;
            if (TAN) TO[o] = tan;
        }
    }
}

template <bool TAN>
__global__ void __launch_bounds__(B0S_SEG_THREADS) k_b0s_scan_seg(const __grid_constant__ B0sScanBatch bt) {
    const B0sScanArgs& a = bt.a[blockIdx.y];
    const int S = bt.g[blockIdx.y].S, F = bt.g[blockIdx.y].F, f_fast = bt.g[blockIdx.y].f_fast;
    if ((int)blockIdx.x >= bt.g[blockIdx.y].blocks) return;
    extern __shared__ double sm[];
    const int M = a.M, P = M + S + 1;
    double* ceps = sm;
    double* cgam = ceps + P;
    double* cdeps = cgam + P;
    double* cA = cdeps + P;                   // dual number (A, dA) of the multiplier product of every segment: fibre independent
    double* cdA = cA + S;
    double* maps = cdA + S;                   // [B | T][fl * S + s]
    for (int i = threadIdx.x; i < M; i += B0S_SEG_THREADS) {
        ceps[b0s_skew(i)] = a.eps[i];
        cgam[b0s_skew(i)] = a.eps[M + i];
        if (TAN) cdeps[b0s_skew(i)] = a.eps[2 * M + i];
    }
    const int t = threadIdx.x;
    const int fl = f_fast ? t % F : t / S, s = f_fast ? t / F : t % S;
    const i64 f = (i64)blockIdx.x * F + fl;
    const bool live = fl < F && s < S && f < a.n_fibres;
    const i64 fh = live ? f / a.n_lo : 0, fo = live ? f - fh * a.n_lo : 0;
    const double* __restrict__ v = a.src + fh * a.s_hi + fo * a.s_lo;
    const i64 dbase = fh * a.d_hi + fo * a.d_lo;
    const int i0 = s * B0S_SEG;
    const int len = live ? (M - i0 < B0S_SEG ? M - i0 : B0S_SEG) : 0;
    double vv[B0S_SEG];
#pragma unroll
    for (int k = 0; k < B0S_SEG; ++k) vv[k] = k < len ? v[(i64)(i0 + k) * a.s_mode] : 0.0;
    const int slot = fl * S + s;
    __syncthreads();
    if (t < S) {                              // the product is commutative: one (A, dA) per segment serves both sweeps
        double A = 1.0, dA = 0.0;
        const int q0 = t * B0S_SEG, ql = M - q0 < B0S_SEG ? M - q0 : B0S_SEG;
        for (int k = 0; k < ql; ++k) {
            const double e = ceps[b0s_skew(q0 + k)];
            if (TAN) dA = fma(e, dA, cdeps[b0s_skew(q0 + k)] * A);
            A *= e;
        }
        cA[t] = A;
        if (TAN) cdA[t] = dA;
    }
    b0s_seg_sweep<TAN, 0>(a, vv, ceps, cgam, cdeps, cA, cdA, maps, S, fl, s, slot, i0, len, live, dbase);     // L sweep (ascending)
    b0s_seg_sweep<TAN, 1>(a, vv, ceps, cgam, cdeps, cA, cdA, maps, S, fl, s, slot, i0, len, live, dbase);     // R sweep (descending)
}

template <int dir>
__device__ __forceinline__ void b0s_seg_adj_sweep(const B0sAdjArgs& a, double (&g)[B0S_SEG], double (&out)[B0S_SEG], const double* ceps,
                                                  const double* cgam, const double* cA, double* maps, int S, int fl, int s, int slot,
                                                  int i0, int len, bool live, i64 go) {
    const int M = a.M;
    if (dir) {
#pragma unroll
        for (int k = 0; k < B0S_SEG; ++k) g[k] = k < len ? a.gR[go + (i64)(i0 + k + 1) * a.g_mode] : 0.0;
    }
    double c = 0.0;
#pragma unroll
    for (int kk = 0; kk < B0S_SEG; ++kk) {
        const int k = dir ? kk : B0S_SEG - 1 - kk;
        if (k < len) c = fma(ceps[b0s_skew(i0 + k)], c, g[k]);
    }
    if (dir) __syncthreads();
    if (live) maps[slot] = c;
    __syncthreads();
    double cin = 0.0;
    if (live) {
        if (!dir) {
            cin = a.gL[go + (i64)(M + 1) * a.g_mode];
            for (int q = S - 1; q > s; --q) cin = fma(cA[q], cin, maps[fl * S + q]);
        } else {
            cin = a.gR[go];
            for (int q = 0; q < s; ++q) cin = fma(cA[q], cin, maps[fl * S + q]);
        }
    }
    c = cin;
#pragma unroll
    for (int kk = 0; kk < B0S_SEG; ++kk) {
        const int k = dir ? kk : B0S_SEG - 1 - kk;
        if (k < len) {
            out[k] = fma(cgam[b0s_skew(i0 + k)], c, out[k]);
            c = fma(ceps[b0s_skew(i0 + k)], c, g[k]);
        }
    }
}

// Segmented adjoint sweeps (same decomposition): sweep 1 descending, c <- eps_i c + gL[i + 1] from c = gL[M + 1],
// dv[i] = gam_i c + gC[i + 1]; sweep 2 ascending, c <- eps_i c + gR[i + 1] from c = gR[0], dv[i] += gam_i c
// (c = the state BEFORE the step in both).  dv stays in registers between the sweeps.
__global__ void __launch_bounds__(B0S_SEG_THREADS) k_b0s_scan_adj_seg(const __grid_constant__ B0sAdjBatch bt) {
    const B0sAdjArgs& a = bt.a[blockIdx.y];
    const int S = bt.g[blockIdx.y].S, F = bt.g[blockIdx.y].F, f_fast = bt.g[blockIdx.y].f_fast;
    if ((int)blockIdx.x >= bt.g[blockIdx.y].blocks) return;
    extern __shared__ double sm[];
    const int M = a.M, P = M + S + 1;
    double* ceps = sm;
    double* cgam = ceps + P;
    double* cA = cgam + P;                    // multiplier product of every segment (fibre independent), S entries (of P)
    double* maps = cA + P;
    for (int i = threadIdx.x; i < M; i += B0S_SEG_THREADS) {
        ceps[b0s_skew(i)] = a.eps[i];
        cgam[b0s_skew(i)] = a.eps[M + i];
    }
    const int t = threadIdx.x;
    const int fl = f_fast ? t % F : t / S, s = f_fast ? t / F : t % S;
    const i64 f = (i64)blockIdx.x * F + fl;
    const bool live = fl < F && s < S && f < a.n_fibres;
    const i64 fh = live ? f / a.n_lo : 0, fo = live ? f - fh * a.n_lo : 0;
    const i64 go = fh * a.g_hi + fo * a.g_lo;
    double* __restrict__ dv = a.dv + fh * a.v_hi + fo * a.v_lo;
    const int i0 = s * B0S_SEG;
    const int len = live ? (M - i0 < B0S_SEG ? M - i0 : B0S_SEG) : 0;
    double g[B0S_SEG], out[B0S_SEG];
#pragma unroll
    for (int k = 0; k < B0S_SEG; ++k) g[k] = k < len ? a.gL[go + (i64)(i0 + k + 1) * a.g_mode] : 0.0;
#pragma unroll
    for (int k = 0; k < B0S_SEG; ++k) out[k] = (k < len && a.gC) ? a.gC[go + (i64)(i0 + k + 1) * a.g_mode] : 0.0;
    __syncthreads();
    if (t < S) {
        double A = 1.0;
        const int q0 = t * B0S_SEG, ql = M - q0 < B0S_SEG ? M - q0 : B0S_SEG;
        for (int k = 0; k < ql; ++k) A *= ceps[b0s_skew(q0 + k)];
        cA[t] = A;
    }
    const int slot = fl * S + s;
    b0s_seg_adj_sweep<0>(a, g, out, ceps, cgam, cA, maps, S, fl, s, slot, i0, len, live, go);     // descending sweep over gL
    b0s_seg_adj_sweep<1>(a, g, out, ceps, cgam, cA, maps, S, fl, s, slot, i0, len, live, go);     // ascending sweep over gR
#pragma unroll
    for (int k = 0; k < B0S_SEG; ++k) if (k < len) dv[(i64)(i0 + k) * a.v_mode] = out[k];
}

// Per-cell quadratic-form tables of dimension blockIdx.y, one warp per extended cell e:
//   W[mat][s][e], mat 0 = P, 1 = Q, s = LL, LC, LR, CC, CR, RR, from V^X = G^X Mat (E x M) and the rows of G^Y.
// grid (ceil((K+1) / 8), D), 256 threads
template <typename T>
struct B0sWArgs {
    int K[VGGP_MAX_D];
    const double* GL[VGGP_MAX_D];
    const double* GR[VGGP_MAX_D];
    const double* V[VGGP_MAX_D][4];      // VP_L, VP_R, VQ_L, VQ_R
    const double* Mat[VGGP_MAX_D][2];    // P_d, Q_d (M x M)
    T* W[VGGP_MAX_D];                    // [2][6][E]
};

template <typename T>
__global__ void __launch_bounds__(256) k_b0s_W(const __grid_constant__ B0sWArgs<T> a) {
    const int d = blockIdx.y;
    const int K = a.K[d], M = K - 1, E = K + 1;
    const int lane = threadIdx.x & 31;
    const int e = (int)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (e >= E) return;
    const int c = e - 1;
    const bool real = c >= 0 && c < M;
    const double* gl = a.GL[d] + (i64)e * M;
    const double* gr = a.GR[d] + (i64)e * M;
#pragma unroll
    for (int mat = 0; mat < 2; ++mat) {
        const double* vl = a.V[d][2 * mat] + (i64)e * M;
        const double* vr = a.V[d][2 * mat + 1] + (i64)e * M;
        double ll = 0.0, lr = 0.0, rr = 0.0;
        for (int i = lane; i < M; i += 32) {
            ll = fma(vl[i], gl[i], ll);
            lr = fma(vl[i], gr[i], lr);
            rr = fma(vr[i], gr[i], rr);
        }
        ll = warp_sum(ll);
        lr = warp_sum(lr);
        rr = warp_sum(rr);
        if (lane == 0) {
            T* w = a.W[d] + (i64)mat * 6 * E;
            w[0 * E + e] = (T)ll;
            w[1 * E + e] = (T)(real ? vl[c] : 0.0);
            w[2 * E + e] = (T)lr;
            w[3 * E + e] = (T)(real ? a.Mat[d][mat][(i64)c * M + c] : 0.0);
            w[4 * E + e] = (T)(real ? vr[c] : 0.0);
            w[5 * E + e] = (T)rr;
        }
    }
}

// 2-D mean tables T[X][Y][e1][e2] (obs dtype) from the four dense corner products and the embedded one-sided products.
//   TT[2 x + y] = G1^x U^y (E1 x E2), x, y in {L, R};  B[x] = G1^x A (E1 x M2);  U[y] = A G2^y^T (M1 x E2);  A (M1 x M2)
// grid (ceil(E1 E2 / 256)), 256 threads
template <typename T>
struct B0sT2Args {
    int E1, E2, M1, M2;
    const double* TT[4];
    const double* B[2];
    const double* U[2];
    const double* A;
    T* Tt;                               // [3][3][E1][E2]
};

template <typename T>
__global__ void __launch_bounds__(256) k_b0s_T2(const __grid_constant__ B0sT2Args<T> a) {
    const i64 idx = (i64)blockIdx.x * 256 + threadIdx.x;
    const i64 EE = (i64)a.E1 * a.E2;
    if (idx >= EE) return;
    const int e1 = (int)(idx / a.E2), e2 = (int)(idx % a.E2);
    const int c1 = e1 - 1, c2 = e2 - 1;
    const bool r1 = c1 >= 0 && c1 < a.M1, r2 = c2 >= 0 && c2 < a.M2;
#pragma unroll
    for (int X = 0; X < 3; ++X)
#pragma unroll
        for (int Y = 0; Y < 3; ++Y) {
            double v;
            const int x = X == B0S_L ? 0 : 1, y = Y == B0S_L ? 0 : 1;
            if (X != B0S_C && Y != B0S_C) v = a.TT[2 * x + y][idx];
            else if (X != B0S_C) v = r2 ? a.B[x][(i64)e1 * a.M2 + c2] : 0.0;
            else if (Y != B0S_C) v = r1 ? a.U[y][(i64)c1 * a.E2 + e2] : 0.0;
            else v = (r1 && r2) ? a.A[(i64)c1 * a.M2 + c2] : 0.0;
            a.Tt[(i64)(3 * X + Y) * EE + idx] = (T)v;
        }
}

// 1-D mean tables T[X][e]: one warp per extended cell.  grid (ceil(E / 8)), 256 threads
template <typename T>
__global__ void __launch_bounds__(256) k_b0s_T1(int K, const double* __restrict__ GL, const double* __restrict__ GR,
                                                const double* __restrict__ A, T* __restrict__ Tt) {
    const int M = K - 1, E = K + 1;
    const int lane = threadIdx.x & 31;
    const int e = (int)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (e >= E) return;
    double l = 0.0, r = 0.0;
    for (int i = lane; i < M; i += 32) {
        l = fma(GL[(i64)e * M + i], A[i], l);
        r = fma(GR[(i64)e * M + i], A[i], r);
    }
    l = warp_sum(l);
    r = warp_sum(r);
    if (lane == 0) {
        const int c = e - 1;
        Tt[0 * E + e] = (T)l;
        Tt[1 * E + e] = (T)((c >= 0 && c < M) ? A[c] : 0.0);
        Tt[2 * E + e] = (T)r;
    }
}

// ---- per-point side ---------------------------------------------------------------------------------------
// extended cell and the three local features (and, optionally, their lengthscale derivatives) of one coordinate
template <typename T>
__device__ __forceinline__ int b0s_local(const float* __restrict__ t, int K, T x, T l, T s2, T (&f)[3], T* df) {
    const int idx = lower_bound_knots<T>(t, K, x);           // knots < x: x in (t_{idx-1}, t_idx]
    const int c = idx - 1;
    const T tc = (T)t[c > 0 ? (c < K - 1 ? c : K - 1) : 0];
    const T tc1 = (T)t[c + 1 < K - 1 ? c + 1 : K - 1];
    const bool hasL = c >= 0, hasR = c <= K - 2, real = hasL && hasR;
    const T zL = (x - tc) / l, zR = (tc1 - x) / l;
    const T eL = hasL ? exp(-zL) : (T)0, eR = hasR ? exp(-zR) : (T)0;
    f[B0S_L] = s2 * l * eL;
    f[B0S_R] = s2 * l * eR;
    f[B0S_C] = real ? (T)2 * s2 * l - f[B0S_L] - f[B0S_R] : (T)0;
    if (df) {
        df[B0S_L] = hasL ? s2 * eL * ((T)1 + zL) : (T)0;
        df[B0S_R] = hasR ? s2 * eR * ((T)1 + zR) : (T)0;
        df[B0S_C] = real ? (T)2 * s2 - df[B0S_L] - df[B0S_R] : (T)0;
    }
    return c + 1;
}

template <typename T, int D>
struct B0sPointTables {
    MeshView mesh[D];
    int E[D];
    const T* Tt;             // D = 1: [3][E1];  D = 2: [3][3][E1][E2]
    const T* W[D];           // [2][6][E_d]
    const double* theta;     // l[D], s2[D], noise
};

// mean, prod p, prod q at one point; fq[d][X] receives the local features, tm[d][X] = d mu / d f_d^X when `tm` != nullptr
template <typename T, int D>
__device__ __forceinline__ void b0s_eval(const B0sPointTables<T, D>& a, const int (&e)[D], const T (&f)[D][3], T& mu,
                                         T (&p)[D], T (&q)[D], T (*tm)[3], T (*zp)[3], T (*zq)[3]) {
    if (D == 1) {
        mu = (T)0;
#pragma unroll
        for (int X = 0; X < 3; ++X) {
            const T tv = a.Tt[(i64)X * a.E[0] + e[0]];
            mu = fma(f[0][X], tv, mu);
            if (tm) tm[0][X] = tv;
        }
    } else {
        const i64 EE = (i64)a.E[0] * a.E[1], at = (i64)e[0] * a.E[1] + e[1];
        T t1[3] = {(T)0, (T)0, (T)0}, t2[3] = {(T)0, (T)0, (T)0};
#pragma unroll
        for (int X = 0; X < 3; ++X)
#pragma unroll
            for (int Y = 0; Y < 3; ++Y) {
                const T tv = a.Tt[(i64)(3 * X + Y) * EE + at];
                t1[X] = fma(f[D - 1][Y], tv, t1[X]);
                t2[Y] = fma(f[0][X], tv, t2[Y]);
            }
        mu = f[0][0] * t1[0] + f[0][1] * t1[1] + f[0][2] * t1[2];
        if (tm) {
#pragma unroll
            for (int X = 0; X < 3; ++X) { tm[0][X] = t1[X]; tm[D - 1][X] = t2[X]; }
        }
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const int E = a.E[d];
#pragma unroll
        for (int mat = 0; mat < 2; ++mat) {
            const T* w = a.W[d] + (i64)mat * 6 * E + e[d];
            const T ll = w[0], lc = w[E], lr = w[2 * E], cc = w[3 * E], cr = w[4 * E], rr = w[5 * E];
            // z^X = sum_Y W^{XY} f^Y  (half the gradient of the quadratic form)
            const T z0 = ll * f[d][0] + lc * f[d][1] + lr * f[d][2];
            const T z1 = lc * f[d][0] + cc * f[d][1] + cr * f[d][2];
            const T z2 = lr * f[d][0] + cr * f[d][1] + rr * f[d][2];
            const T v = f[d][0] * z0 + f[d][1] * z1 + f[d][2] * z2;
            if (mat == 0) { p[d] = v; if (zp) { zp[d][0] = z0; zp[d][1] = z1; zp[d][2] = z2; } }
            else { q[d] = v; if (zq) { zq[d][0] = z0; zq[d][1] = z1; zq[d][2] = z2; } }
        }
    }
}

// Point prediction for the B0 family (kronecker_structure.py:199-230 restricted to the marginals), O(1) per point.
template <typename T, int D>
struct B0sPredictArgs {
    B0sPointTables<T, D> tab;
    const T* x[D];
    i64 n;
    T* mean;
    T* var;
};

template <typename T, int D>
__global__ void __launch_bounds__(256) k_predict_b0s(const __grid_constant__ B0sPredictArgs<T, D> a) {
    T kff = (T)1;
#pragma unroll
    for (int d = 0; d < D; ++d) kff *= (T)a.tab.theta[D + d];
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (i64)gridDim.x * blockDim.x) {
        int e[D];
        T f[D][3], p[D], q[D], mu;
#pragma unroll
        for (int d = 0; d < D; ++d)
            e[d] = b0s_local<T>(a.tab.mesh[d].t, a.tab.mesh[d].K, a.x[d][i], (T)a.tab.theta[d], (T)a.tab.theta[D + d], f[d], nullptr);
        b0s_eval<T, D>(a.tab, e, f, mu, p, q, nullptr, nullptr, nullptr);
        T pp = p[0], qq = q[0];
#pragma unroll
        for (int d = 1; d < D; ++d) { pp *= p[d]; qq *= q[d]; }
        a.mean[i] = mu;
        a.var[i] = kff - pp + qq;
    }
}


// ---------------------------------------------------------------------------------------------------------
// K1 for the B0 family in scan form: fused per-observation forward + backward over the binned layout of binplan.hpp
// with EXTENDED cells (e_d = 0..K_d: observations outside the mesh belong to the virtual end cells and do contribute).
// Same task / run structure as k_obs_b1_binned: one lane walks one run = (a share of) the observations of one extended
// cell, the 32 runs of a warp task have the same padded length, the per-cell tables are loaded once per run.
// Raw sums per cell (1 / noise, ell_scale applied on the grid side, like k_obs_b0):
//   GT[X][Y][e1][e2]  += r f1^X f2^Y                      (adjoint of the mean tables; D = 1: GT[X][e])
//   GW[d][mat][s][e_d] += o_d f_d^X f_d^Y                 o = prod_{e != d} p_e (mat 0) or q_e (mat 1), s = LL, LC, LR, CC, CR, RR
//   E += r^2 - prod p + prod q;   G_l[d] += sum_X gf_d^X df_d^X / dl;   G_s[d] += sum_X gf_d^X f_d^X
//        gf_d^X = r dmu/df_d^X + o^p_d zP_d^X - o^q_d zQ_d^X           (the part of the theta gradient that goes through the
//                                                                      local features; the table part is added by the adjoint stage)
// ---------------------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ void b0s_decode_cell(uint32_t cell, const int (&E)[D], int (&e)[D]) {
#pragma unroll
    for (int d = D - 1; d >= 0; --d) {
        e[d] = (int)(cell % (uint32_t)E[d]);
        cell /= (uint32_t)E[d];
    }
}

// flat extended-cell key of every observation (row-major over e_1..e_D); every observation has one
template <typename T, int D>
struct B0sKeyArgs {
    const T* x[D];
    i64 n;
    MeshView mesh[D];
};
template <typename T, int D>
__global__ void __launch_bounds__(256) k_b0s_keys(const __grid_constant__ B0sKeyArgs<T, D> a, uint32_t* __restrict__ keys,
                                                  uint32_t* __restrict__ idx) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (i64)gridDim.x * blockDim.x) {
        uint32_t key = 0;
#pragma unroll
        for (int d = 0; d < D; ++d)
            key = key * (uint32_t)(a.mesh[d].K + 1) + (uint32_t)lower_bound_knots<T>(a.mesh[d].t, a.mesh[d].K, a.x[d][i]);
        keys[i] = key;
        idx[i] = (uint32_t)i;
    }
}

// fill the binned data blocks (same layout as k_bin_gather); padding slots get a finite coordinate of the run's cell
template <typename T, int D>
struct B0sGatherArgs {
    const T* x[D];
    const T* y;
    const uint32_t* perm;
    const float* knots[D];
    int K[D];
    unsigned char* buf;
    i64 off_task_off, off_task_R, off_run_cell, off_run_n, off_run_start, off_data;
    int n_tasks;
};
template <typename T, int D>
__global__ void __launch_bounds__(256) k_b0s_gather(const __grid_constant__ B0sGatherArgs<T, D> a) {
    const i64* task_off = reinterpret_cast<const i64*>(a.buf + a.off_task_off);
    const int* task_R = reinterpret_cast<const int*>(a.buf + a.off_task_R);
    const uint32_t* run_cell = reinterpret_cast<const uint32_t*>(a.buf + a.off_run_cell);
    const int* run_n = reinterpret_cast<const int*>(a.buf + a.off_run_n);
    const uint32_t* run_start = reinterpret_cast<const uint32_t*>(a.buf + a.off_run_start);
    T* data = reinterpret_cast<T*>(a.buf + a.off_data);
    int E[D];
#pragma unroll
    for (int d = 0; d < D; ++d) E[d] = a.K[d] + 1;
    for (int task = blockIdx.x; task < a.n_tasks; task += gridDim.x) {
        const i64 elems = (i64)32 * task_R[task] * (D + 1);
        T* dst = data + task_off[task];
        for (i64 el = threadIdx.x; el < elems; el += blockDim.x) {
            const BinSlot sl = bin_slot_of(el, D);
            const i64 slot = (i64)task * 32 + sl.lane;
            const uint32_t cell = run_cell[slot];
            T v = (T)0;
            if (sl.j < run_n[slot]) {
                const i64 src = (i64)a.perm[(i64)run_start[slot] + sl.j];
                v = (sl.arr < D) ? a.x[sl.arr < D ? sl.arr : 0][src] : a.y[src];
            } else if (sl.arr < D) {
                int e[D];
                b0s_decode_cell<D>(cell != BIN_EMPTY ? cell : 0u, E, e);
                const int c = e[sl.arr] - 1, K = a.K[sl.arr];
                v = (T)a.knots[sl.arr][c > 0 ? (c < K - 1 ? c : K - 1) : 0];
            }
            dst[el] = v;
        }
    }
}

// exp of a non-positive argument in the per-observation loop: float32 observations take the hardware exponential
// (ex2.approx: ~ 2 ulp, far inside the 1e-3 tolerance of float32 configurations), float64 the library function
__device__ __forceinline__ double b0s_exp(double x) { return exp(x); }
__device__ __forceinline__ float b0s_exp(float x) {
#ifdef VGGP_EMUL
    return expf(x);
#else
    return __expf(x);
#endif
}

template <typename T, int D>
struct B0sObsArgs {
    B0sPointTables<T, D> tab;
    const unsigned char* buf;        // binned buffer
    i64 off_task_off, off_task_R, off_run_cell, off_run_n, off_data;
    int n_tasks;
    T* GT;                           // D = 1: [3][E1];  D = 2: [3][3][E1][E2]   (zero on entry)
    T* GW[D];                        // [2][6][E_d]                               (zero on entry)
    double* gs;                      // gbuf scalars: [E, n, -, G_l[0..1], G_s[0..1]]
    double n_real;
    unsigned int* counter;
};

constexpr int B0S_THREADS = 128;

template <typename T, int D>
__global__ void __launch_bounds__(B0S_THREADS, (sizeof(T) == 4 ? 4 : 2)) k_obs_b0s(const __grid_constant__ B0sObsArgs<T, D> a) {
    __shared__ double red[32];
    constexpr int NT = D == 1 ? 3 : 9;
    const int lane = threadIdx.x & 31;
    double accE = 0.0, accGl[D], accGs[D];
    T l[D], s2[D], rl[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        accGl[d] = 0.0; accGs[d] = 0.0;
        l[d] = (T)a.tab.theta[d];
        rl[d] = (T)(1.0 / a.tab.theta[d]);      // one reciprocal per kernel instead of two divisions per observation and dimension
        s2[d] = (T)a.tab.theta[D + d];
    }
    const i64 EE = D == 1 ? (i64)a.tab.E[0] : (i64)a.tab.E[0] * a.tab.E[D - 1];
    bool first = true;                  // first task = the warp's global index, later ones from the counter (see bin_next_task)
    const unsigned int nwarps_all = gridDim.x * (blockDim.x >> 5);
    for (;;) {
        unsigned int task = 0;
        if (first) {
            task = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
            first = false;
        } else {
            if (lane == 0) task = atomicAdd(a.counter, 1u) + nwarps_all;
            task = __shfl_sync(0xffffffffu, task, 0);
        }
        if (task >= (unsigned int)a.n_tasks) break;
        const i64 slot = (i64)task * 32 + lane;
        const uint32_t cell = __ldg(reinterpret_cast<const uint32_t*>(a.buf + a.off_run_cell) + slot);
        const int nrun = __ldg(reinterpret_cast<const int*>(a.buf + a.off_run_n) + slot);
        const int groups = __ldg(reinterpret_cast<const int*>(a.buf + a.off_task_R) + task) >> 2;
        const T* base = reinterpret_cast<const T*>(a.buf + a.off_data)
                        + __ldg(reinterpret_cast<const i64*>(a.buf + a.off_task_off) + task) + lane * 4;
        const bool valid = cell != BIN_EMPTY;
        int e[D];
        b0s_decode_cell<D>(valid ? cell : 0u, a.tab.E, e);
        const i64 at = D == 1 ? (i64)e[0] : (i64)e[0] * a.tab.E[D - 1] + e[D - 1];
        // per-cell constants: mean tables, quadratic-form tables, the two knots the local features decay from
        T Tc[NT], Wc[D][2][6], tc[D], tc1[D];
        bool hasL[D], hasR[D];
#pragma unroll
        for (int k = 0; k < NT; ++k) Tc[k] = a.tab.Tt[(i64)k * EE + at];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const int E = a.tab.E[d], K = a.tab.mesh[d].K, c = e[d] - 1;
#pragma unroll
            for (int mat = 0; mat < 2; ++mat)
#pragma unroll
                for (int s = 0; s < 6; ++s) Wc[d][mat][s] = a.tab.W[d][(i64)(mat * 6 + s) * E + e[d]];
            tc[d] = (T)a.tab.mesh[d].t[c > 0 ? (c < K - 1 ? c : K - 1) : 0];
            tc1[d] = (T)a.tab.mesh[d].t[c + 1 < K - 1 ? c + 1 : K - 1];
            hasL[d] = c >= 0;
            hasR[d] = c <= K - 2;
        }
        T gT[NT], gW[D][2][6];
#pragma unroll
        for (int k = 0; k < NT; ++k) gT[k] = (T)0;
#pragma unroll
        for (int d = 0; d < D; ++d)
#pragma unroll
            for (int mat = 0; mat < 2; ++mat)
#pragma unroll
                for (int s = 0; s < 6; ++s) gW[d][mat][s] = (T)0;
        T accEr = (T)0, gl[D], gsv[D];
#pragma unroll
        for (int d = 0; d < D; ++d) { gl[d] = (T)0; gsv[d] = (T)0; }

#pragma unroll 1
        for (int gi = 0; gi < groups; ++gi) {
            T xg[D][4], yg[4];
            bin_load_group<T, D>(base + (i64)gi * ((D + 1) * 128), xg, yg);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (4 * gi + j >= nrun) continue;                   // padding slots of this lane
                T f[D][3], df[D][3];
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const T zL = (xg[d][j] - tc[d]) * rl[d], zR = (tc1[d] - xg[d][j]) * rl[d];
                    const T eL = hasL[d] ? b0s_exp(-zL) : (T)0, eR = hasR[d] ? b0s_exp(-zR) : (T)0;
                    const bool real = hasL[d] && hasR[d];
                    f[d][0] = s2[d] * l[d] * eL;
                    f[d][2] = s2[d] * l[d] * eR;
                    f[d][1] = real ? (T)2 * s2[d] * l[d] - f[d][0] - f[d][2] : (T)0;
                    df[d][0] = hasL[d] ? s2[d] * eL * ((T)1 + zL) : (T)0;
                    df[d][2] = hasR[d] ? s2[d] * eR * ((T)1 + zR) : (T)0;
                    df[d][1] = real ? (T)2 * s2[d] - df[d][0] - df[d][2] : (T)0;
                }
                // mean and d mu / d f
                T mu = (T)0, tm[D][3];
                if (D == 1) {
#pragma unroll
                    for (int X = 0; X < 3; ++X) { tm[0][X] = Tc[X]; mu = fma(f[0][X], Tc[X], mu); }
                } else {
#pragma unroll
                    for (int X = 0; X < 3; ++X) { tm[0][X] = (T)0; tm[D - 1][X] = (T)0; }
#pragma unroll
                    for (int X = 0; X < 3; ++X)
#pragma unroll
                        for (int Y = 0; Y < 3; ++Y) {
                            tm[0][X] = fma(f[D - 1][Y], Tc[3 * X + Y], tm[0][X]);
                            tm[D - 1][Y] = fma(f[0][X], Tc[3 * X + Y], tm[D - 1][Y]);
                        }
#pragma unroll
                    for (int X = 0; X < 3; ++X) mu = fma(f[0][X], tm[0][X], mu);
                }
                // quadratic forms and their half gradients
                T pq[D][2], z[D][2][3];
#pragma unroll
                for (int d = 0; d < D; ++d)
#pragma unroll
                    for (int mat = 0; mat < 2; ++mat) {
                        const T* w = Wc[d][mat];
                        z[d][mat][0] = w[0] * f[d][0] + w[1] * f[d][1] + w[2] * f[d][2];
                        z[d][mat][1] = w[1] * f[d][0] + w[3] * f[d][1] + w[4] * f[d][2];
                        z[d][mat][2] = w[2] * f[d][0] + w[4] * f[d][1] + w[5] * f[d][2];
                        pq[d][mat] = f[d][0] * z[d][mat][0] + f[d][1] * z[d][mat][1] + f[d][2] * z[d][mat][2];
                    }
                const T r = yg[j] - mu;
                T pp = pq[0][0], qq = pq[0][1];
#pragma unroll
                for (int d = 1; d < D; ++d) { pp *= pq[d][0]; qq *= pq[d][1]; }
                accEr += fma(r, r, qq - pp);
                if (D == 1) {
#pragma unroll
                    for (int X = 0; X < 3; ++X) gT[X] = fma(r, f[0][X], gT[X]);
                } else {
#pragma unroll
                    for (int X = 0; X < 3; ++X) {
                        const T rf = r * f[0][X];
#pragma unroll
                        for (int Y = 0; Y < 3; ++Y) gT[3 * X + Y] = fma(rf, f[D - 1][Y], gT[3 * X + Y]);
                    }
                }
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    T o[2] = {(T)1, (T)1};
#pragma unroll
                    for (int e2 = 0; e2 < D; ++e2)
                        if (e2 != d) { o[0] *= pq[e2][0]; o[1] *= pq[e2][1]; }
                    const T ff[6] = {f[d][0] * f[d][0], f[d][0] * f[d][1], f[d][0] * f[d][2],
                                     f[d][1] * f[d][1], f[d][1] * f[d][2], f[d][2] * f[d][2]};
#pragma unroll
                    for (int mat = 0; mat < 2; ++mat)
#pragma unroll
                        for (int s = 0; s < 6; ++s) gW[d][mat][s] = fma(o[mat], ff[s], gW[d][mat][s]);
#pragma unroll
                    for (int X = 0; X < 3; ++X) {
                        const T gf = r * tm[d][X] + o[0] * z[d][0][X] - o[1] * z[d][1][X];
                        gl[d] = fma(gf, df[d][X], gl[d]);
                        gsv[d] = fma(gf, f[d][X], gsv[d]);
                    }
                }
            }
        }
        if (valid) {
#pragma unroll
            for (int k = 0; k < NT; ++k) atomicAdd(a.GT + (i64)k * EE + at, gT[k]);
#pragma unroll
            for (int d = 0; d < D; ++d)
#pragma unroll
                for (int mat = 0; mat < 2; ++mat)
#pragma unroll
                    for (int s = 0; s < 6; ++s) atomicAdd(a.GW[d] + (i64)(mat * 6 + s) * a.tab.E[d] + e[d], gW[d][mat][s]);
            accE += (double)accEr;
#pragma unroll
            for (int d = 0; d < D; ++d) { accGl[d] += (double)gl[d]; accGs[d] += (double)gsv[d]; }
        }
    }
    double eb = block_sum(accE, red);
    if (threadIdx.x == 0) atomicAdd(a.gs + 0, eb);
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const double g1 = block_sum(accGl[d], red);
        const double g2 = block_sum(accGs[d], red);
        if (threadIdx.x == 0) {
            atomicAdd(a.gs + 3 + d, g1);
            atomicAdd(a.gs + 5 + d, g2);
        }
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) a.gs[1] = a.n_real;
}

// ---- adjoint stage: raw per-cell sums -> the gbuf blocks the B0 grid backward consumes -----------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_b0s_to_double(const T* __restrict__ src, double* __restrict__ dst, i64 n) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) dst[i] = (double)src[i];
}
template <typename T>
__global__ void __launch_bounds__(256) k_b0s_from_double(const double* __restrict__ src, T* __restrict__ dst, i64 n) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) dst[i] = (T)src[i];
}
__device__ __forceinline__ double b0s_gw(const double* gw6, int X, int Y) { return gw6[b0s_sym_index(X, Y)]; }

// One dimension, one matrix (mat 0 = P, 1 = Q):  S^X[e][i] = sum_Y gw^{XY}[e] G^Y[e][i]   (X = L, C, R; G^C[e] = unit row e-1)
// grid (ceil(E M / 256)), 256 threads
template <typename T>
__global__ void __launch_bounds__(256) k_b0s_S(int K, const double* __restrict__ GL, const double* __restrict__ GR,
                                               const T* __restrict__ GW /* [6][E] of this matrix */, double* __restrict__ SL,
                                               double* __restrict__ SC, double* __restrict__ SR) {
    const int M = K - 1, E = K + 1;
    const i64 idx = (i64)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (i64)E * M) return;
    const int e = (int)(idx / M), i = (int)(idx % M);
    double gw6[6];
#pragma unroll
    for (int s = 0; s < 6; ++s) gw6[s] = (double)GW[(i64)s * E + e];
    const double g[3] = {GL[idx], (i == e - 1) ? 1.0 : 0.0, GR[idx]};
    double out[3];
#pragma unroll
    for (int X = 0; X < 3; ++X) out[X] = b0s_gw(gw6, X, 0) * g[0] + b0s_gw(gw6, X, 1) * g[1] + b0s_gw(gw6, X, 2) * g[2];
    SL[idx] = out[0];
    SC[idx] = out[1];
    SR[idx] = out[2];
}

// 1-D d alpha: galpha[i] = sum_X sum_e G^X[e][i] GT^X[e]  (one thread per i)
template <typename T>
__global__ void __launch_bounds__(256) k_b0s_galpha1(int K, const double* __restrict__ GL, const double* __restrict__ GR,
                                                     const double* __restrict__ GT /* [3][E] */, T* __restrict__ galpha) {
    const int M = K - 1, E = K + 1;
    const int i = (int)blockIdx.x * 256 + threadIdx.x;
    if (i >= M) return;
    double acc = GT[(i64)B0S_C * E + i + 1];
    for (int e = 0; e < E; ++e)
        acc += GL[(i64)e * M + i] * GT[(i64)B0S_L * E + e] + GR[(i64)e * M + i] * GT[(i64)B0S_R * E + e];
    galpha[i] = (T)acc;
}

// Table part of the lengthscale gradient that goes through the quadratic-form tables (and, in 1-D, the mean tables):
//   out += sum_e sum_{X in L,R} sum_Y  gwP^{XY}[e] <VP^Y[e,:], dG^X[e,:]>  -  gwQ^{XY}[e] <VQ^Y[e,:], dG^X[e,:]>
//          (+ 1-D:  GT^X[e] <A, dG^X[e,:]>),                      V^C[e,:] = row e-1 of the matrix
// one warp per extended cell e.  grid (ceil(E / 8)), 256 threads
template <typename T>
__global__ void __launch_bounds__(256) k_b0s_gl_rows(int K, const T* __restrict__ GW /* [2][6][E] */,
                                                     const double* __restrict__ VPL, const double* __restrict__ VPR,
                                                     const double* __restrict__ VQL, const double* __restrict__ VQR,
                                                     const double* __restrict__ P, const double* __restrict__ Q,
                                                     const double* __restrict__ dGL, const double* __restrict__ dGR,
                                                     const double* __restrict__ GT1 /* 1-D: [3][E] float64, else nullptr */,
                                                     const double* __restrict__ A1, double* __restrict__ out) {
    __shared__ double red[32];
    const int M = K - 1, E = K + 1;
    const int lane = threadIdx.x & 31;
    const int e = (int)blockIdx.x * 8 + (threadIdx.x >> 5);
    double acc = 0.0;
    if (e < E) {
        const int c = e - 1;
        const bool real = c >= 0 && c < M;
        double gp[6], gq[6];
#pragma unroll
        for (int s = 0; s < 6; ++s) { gp[s] = (double)GW[(i64)s * E + e]; gq[s] = (double)GW[(i64)(6 + s) * E + e]; }
        const i64 row = (i64)e * M;
        for (int i = lane; i < M; i += 32) {
            const double vp[3] = {VPL[row + i], real ? P[(i64)c * M + i] : 0.0, VPR[row + i]};
            const double vq[3] = {VQL[row + i], real ? Q[(i64)c * M + i] : 0.0, VQR[row + i]};
            const double dg[2] = {dGL[row + i], dGR[row + i]};
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int X = k == 0 ? B0S_L : B0S_R;
                double v = 0.0;
#pragma unroll
                for (int Y = 0; Y < 3; ++Y) v += b0s_gw(gp, X, Y) * vp[Y] - b0s_gw(gq, X, Y) * vq[Y];
                if (GT1) v += GT1[(i64)X * E + e] * A1[i];
                acc = fma(v, dg[k], acc);
            }
        }
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(out, acc);
}

// out += sum_i a[i] b[i]
__global__ void __launch_bounds__(256) k_b0s_dot1(const double* __restrict__ x, const double* __restrict__ y, i64 n,
                                                  double* __restrict__ out) {
    __shared__ double red[32];
    double acc = 0.0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) acc = fma(x[i], y[i], acc);
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(out, acc);
}

// several dot products in one launch: out[e] += sum_i x_e[i] y_e[i], blockIdx.y = e
struct B0sDotBatch { const double* x[B0S_MAX_BATCH]; const double* y[B0S_MAX_BATCH]; double* out[B0S_MAX_BATCH]; i64 n; };
__global__ void __launch_bounds__(256) k_b0s_dot_batch(const __grid_constant__ B0sDotBatch b) {
    __shared__ double red[32];
    const double* __restrict__ x = b.x[blockIdx.y];
    const double* __restrict__ y = b.y[blockIdx.y];
    double acc = 0.0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < b.n; i += (i64)gridDim.x * blockDim.x) acc = fma(x[i], y[i], acc);
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(b.out[blockIdx.y], acc);
}

}  // namespace vggp
