// Per-observation kernels: B1 interpolation stencil (bit-exact restatement of bspline.py:23-77, 92-94),
// dense feature matrices (tests / small predictions) and the fused ELBO forward+backward over observations.
#pragma once
#include "common.cuh"

namespace vggp {

struct MeshView {
    const float* t;      // knots (float32, exactly as uploaded)
    int K;
    float t0, inv_h;     // arithmetic guess of the cell: floor((x - t0) * inv_h)
    int nearly_uniform;  // guess is within +-2 cells of the truth (checked at plan creation)
};

// c = clamp(searchsorted(mesh, x, right=False) - 1, 0, K-2);  inside = mesh[0] <= x <= mesh[K-1].
// Comparisons are done in T against the float32 knots promoted to T (exact), like torch does.
template <typename T>
__device__ __forceinline__ int find_cell(const float* __restrict__ t, int K, float t0, float inv_h,
                                         int nearly_uniform, T x, bool& inside) {
    inside = (x >= (T)t[0]) && (x <= (T)t[K - 1]);
    int c;
    if (nearly_uniform) {
        float gf = ((float)x - t0) * inv_h;
        gf = fminf(fmaxf(gf, 0.0f), (float)(K - 2));     // NaN -> 0
        c = (int)gf;
#pragma unroll
        for (int it = 0; it < 3; ++it)
            if (c > 0 && x <= (T)t[c]) --c;
#pragma unroll
        for (int it = 0; it < 3; ++it)
            if (c < K - 2 && x > (T)t[c + 1]) ++c;
    } else {
        // lower_bound: first index with t[idx] >= x
        int lo = 0, hi = K;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((T)t[mid] < x) lo = mid + 1; else hi = mid;
        }
        c = min(max(lo - 1, 0), K - 2);
    }
    return c;
}

// Interpolation weights of the two hats overlapping cell c.  The denominator is the float32 knot
// difference promoted to T (0-dim float32 arithmetic in the reference), the numerators are computed in T;
// true IEEE subtraction and division (no reciprocal, no fma contraction is possible here).
template <typename T>
__device__ __forceinline__ void b1_weights(const float* __restrict__ t, int c, T x, bool inside, T& w_lo, T& w_hi) {
    const float tl = t[c], th = t[c + 1];
    const T h = (T)(th - tl);
    const T wh = (x - (T)tl) / h;
    const T wl = ((T)th - x) / h;
    w_hi = inside ? wh : (T)0;
    w_lo = inside ? wl : (T)0;
}

template <typename T>
__global__ void __launch_bounds__(256) k_b1_stencil(MeshView mv, const T* __restrict__ x, i64 n,
                                                    int32_t* __restrict__ c_out, T* __restrict__ w_lo,
                                                    T* __restrict__ w_hi) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const T xv = x[i];
        bool inside;
        const int c = find_cell<T>(mv.t, mv.K, mv.t0, mv.inv_h, mv.nearly_uniform, xv, inside);
        T wl, wh;
        b1_weights<T>(mv.t, c, xv, inside, wl, wh);
        c_out[i] = inside ? c : -1;
        w_lo[i] = wl;
        w_hi[i] = wh;
    }
}

// Dense (M_d, n) B1 feature matrix (zero-filled first by the caller).
template <typename T>
__global__ void __launch_bounds__(256) k_b1_dense(MeshView mv, const T* __restrict__ x, i64 n, T* __restrict__ phi) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const T xv = x[i];
        bool inside;
        const int c = find_cell<T>(mv.t, mv.K, mv.t0, mv.inv_h, mv.nearly_uniform, xv, inside);
        T wl, wh;
        b1_weights<T>(mv.t, c, xv, inside, wl, wh);
        if (inside) {
            phi[(i64)c * n + i] = wl;
            phi[(i64)(c + 1) * n + i] = wh;
        }
    }
}

// B0 cell-integrated Matern-1/2 feature of cell k at x (gridded_kronecker_structure.py:1325-1374), and its
// derivative w.r.t. the lengthscale.  idx = searchsorted(mesh, x, right=False).
template <typename T>
__device__ __forceinline__ T b0_feature(const float* __restrict__ t, int k, int idx, T x, T l, T s2) {
    const T a = (T)t[k], b = (T)t[k + 1];
    const T e1 = l * exp(-fabs(x - a) / l);
    const T e2 = l * exp(-fabs(x - b) / l);
    const int sgn = idx - k - 1;                // indicator = -sign(sgn)
    T v;
    if (sgn == 0) v = (T)2 * l - (e1 + e2);
    else if (sgn > 0) v = -(e1 - e2);
    else v = (e1 - e2);
    return v * s2;
}

template <typename T>
__device__ __forceinline__ int lower_bound_knots(const float* __restrict__ t, int K, T x) {
    int lo = 0, hi = K;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((T)t[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Dense (K-1, n) B0 feature matrix.  grid (ceil(n/256), K-1)
template <typename T>
__global__ void __launch_bounds__(256) k_b0_dense(MeshView mv, const T* __restrict__ x, i64 n, double l, double s2,
                                                  T* __restrict__ phi) {
    const int k = blockIdx.y;
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T xv = x[i];
    const int idx = lower_bound_knots<T>(mv.t, mv.K, xv);
    phi[(i64)k * n + i] = b0_feature<T>(mv.t, k, idx, xv, (T)l, (T)s2);
}

// ---------------------------------------------------------------------------------------------------------
// K1: fused per-observation ELBO forward + backward, B1 (ASVGP) family.
//
// Per observation: cell + weights per dimension, mu = <kron phi_d, alpha> (2^D-point gather), p_d, q_d from the
// main/first off diagonals of P_d, Q_d, then in the same pass the reverse-mode contributions
//   g_alpha[corner] += w_corner * r,  bp_d += (prod_{e!=d} p_e) w (x) w,  bq_d += (prod_{e!=d} q_e) w (x) w,
//   E += r^2 - prod p + prod q,  n += 1.
// All hyper-parameter dependent factors (1/noise, ell_scale, kff) are applied on the grid side.
// ---------------------------------------------------------------------------------------------------------
template <typename T, int D>
struct ObsArgs {
    const T* x[D];
    const T* y;
    i64 n;
    MeshView mesh[D];
    i64 stride[D];
    int band_off[D];     // offset of dim d inside band tables: [pd | po | qd | qo] each K[d] long
    int band_total;      // 4 * sum K
    int knot_off[D];
    int knot_total;
    const T* alpha;
    const T* band;
    T* galpha;
    T* gband;
    double* gs;
};

template <typename T, int D>
__global__ void __launch_bounds__(256) k_obs_b1_v1(const __grid_constant__ ObsArgs<T, D> a) {
    extern __shared__ __align__(16) unsigned char smraw[];
    T* s_band = reinterpret_cast<T*>(smraw);                 // band_total
    T* s_gband = s_band + a.band_total;                       // band_total
    float* s_knots = reinterpret_cast<float*>(s_gband + a.band_total);   // knot_total
    __shared__ double red[32];

    for (int i = threadIdx.x; i < a.band_total; i += blockDim.x) {
        s_band[i] = a.band[i];
        s_gband[i] = (T)0;
    }
#pragma unroll
    for (int d = 0; d < D; ++d)
        for (int i = threadIdx.x; i < a.mesh[d].K; i += blockDim.x) s_knots[a.knot_off[d] + i] = a.mesh[d].t[i];
    __syncthreads();

    T accE = (T)0;
    i64 cnt = 0, cnt_in = 0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (i64)gridDim.x * blockDim.x) {
        int c[D];
        T wl[D], wh[D];
        bool all_in = true;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const T xv = a.x[d][i];
            bool inside;
            const float* t = s_knots + a.knot_off[d];
            c[d] = find_cell<T>(t, a.mesh[d].K, a.mesh[d].t0, a.mesh[d].inv_h, a.mesh[d].nearly_uniform, xv, inside);
            b1_weights<T>(t, c[d], xv, inside, wl[d], wh[d]);
            all_in = all_in && inside;
        }
        const T yv = a.y[i];
        T p[D], q[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const int K = a.mesh[d].K;
            const T* b = s_band + a.band_off[d];
            const T ll = wl[d] * wl[d], lh = wl[d] * wh[d], hh = wh[d] * wh[d];
            p[d] = ll * b[c[d]] + (T)2 * lh * b[K + c[d]] + hh * b[c[d] + 1];
            q[d] = ll * b[2 * K + c[d]] + (T)2 * lh * b[3 * K + c[d]] + hh * b[2 * K + c[d] + 1];
        }
        T pp = (T)1, qq = (T)1;
#pragma unroll
        for (int d = 0; d < D; ++d) { pp *= p[d]; qq *= q[d]; }
        i64 base = 0;
#pragma unroll
        for (int d = 0; d < D; ++d) base += (i64)c[d] * a.stride[d];
        T mu = (T)0;
        T wc[1 << D];
        if (all_in) {
#pragma unroll
            for (int corner = 0; corner < (1 << D); ++corner) {
                T w = (T)1;
                i64 off = base;
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const int hi = (corner >> (D - 1 - d)) & 1;
                    w *= hi ? wh[d] : wl[d];
                    off += hi ? a.stride[d] : 0;
                }
                wc[corner] = w;
                mu += w * __ldg(a.alpha + off);
            }
        }
        const T r = yv - mu;
        accE += r * r - pp + qq;
        cnt += 1;
        if (all_in) {
            cnt_in += 1;
#pragma unroll
            for (int corner = 0; corner < (1 << D); ++corner) {
                i64 off = base;
#pragma unroll
                for (int d = 0; d < D; ++d) off += ((corner >> (D - 1 - d)) & 1) ? a.stride[d] : 0;
                atomicAdd(a.galpha + off, wc[corner] * r);
            }
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const int K = a.mesh[d].K;
                T op = (T)1, oq = (T)1;
#pragma unroll
                for (int e = 0; e < D; ++e)
                    if (e != d) { op *= p[e]; oq *= q[e]; }
                T* gb = s_gband + a.band_off[d];
                const T ll = wl[d] * wl[d], lh = wl[d] * wh[d], hh = wh[d] * wh[d];
                atomicAdd(gb + c[d], op * ll);
                atomicAdd(gb + K + c[d], op * lh);
                atomicAdd(gb + c[d] + 1, op * hh);
                atomicAdd(gb + 2 * K + c[d], oq * ll);
                atomicAdd(gb + 3 * K + c[d], oq * lh);
                atomicAdd(gb + 2 * K + c[d] + 1, oq * hh);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < a.band_total; i += blockDim.x) {
        const T v = s_gband[i];
        if (v != (T)0) atomicAdd(a.gband + i, v);
    }
    double e = block_sum((double)accE, red);
    double c1 = block_sum((double)cnt, red);
    double c2 = block_sum((double)cnt_in, red);
    if (threadIdx.x == 0) {
        atomicAdd(a.gs + 0, e);
        atomicAdd(a.gs + 1, c1);
        atomicAdd(a.gs + 2, c2);
    }
}

}  // namespace vggp
