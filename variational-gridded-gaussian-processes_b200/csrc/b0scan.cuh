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
// The tables are dense float64 products on the grid side (k_gemm_one), the per-point work is O(1).
#pragma once
#include "obs.cuh"

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

}  // namespace vggp
