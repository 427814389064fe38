// Evaluation metrics of the reference (src/utils/evaluationmetrics.py:6-54: MSE, MAE, RMSE, R^2) as one fused
// reduction, and the same reduction fused with the point prediction so that the predictive mean never goes to HBM
// (SURVEY.md section 8f row 2).  Both kernels emit four raw float64 sums; the host mirror finishes the formulas:
//     out = { sum (t - p)^2,  sum |t - p|,  sum (t - t_0),  sum (t - t_0)^2 }          t_0 = the first target
//     MSE = out0 / n, MAE = out1 / n, RMSE = sqrt(MSE), R^2 = 1 - out0 / (out3 - out2^2 / n)
// (the total sum of squares is taken about the pivot t_0 so that a large mean does not cancel in the one-pass form).
// HBM-bound: 2 (or D + 1) values read per element, warp-shuffle + shared-memory block reduction, one float64 atomic per
// block and output.
#pragma once
#include "obs.cuh"

namespace vggp {

__device__ __forceinline__ void metrics_flush(double sse, double sae, double s1, double s2, double* red, double* out) {
    sse = block_sum(sse, red);
    sae = block_sum(sae, red);
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) {
        atomicAdd(out + 0, sse);
        atomicAdd(out + 1, sae);
        atomicAdd(out + 2, s1);
        atomicAdd(out + 3, s2);
    }
}

// out[4] must be zero on entry
template <typename T>
__global__ void __launch_bounds__(256) k_metrics(const T* __restrict__ truth, const T* __restrict__ pred, i64 n,
                                                 double* __restrict__ out) {
    __shared__ double red[32];
    double sse = 0.0, sae = 0.0, s1 = 0.0, s2 = 0.0;
    const double pivot = (double)truth[0];
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const double t = (double)truth[i];
        const double e = t - (double)pred[i];
        sse = fma(e, e, sse);
        sae += fabs(e);
        const double c = t - pivot;
        s1 += c;
        s2 = fma(c, c, s2);
    }
    metrics_flush(sse, sae, s1, s2, red, out);
}

// Prediction of the posterior mean at the test points (same arithmetic as k_predict_b1) fused with the metrics of
// (y, mean).  out[4] must be zero on entry.
template <typename T, int D>
struct PredictMetricsArgs {
    PredictArgs<T, D> p;     // mean / var pointers unused
    const T* y;
    double* out;
};

template <typename T, int D>
__global__ void __launch_bounds__(256) k_predict_metrics(const __grid_constant__ PredictMetricsArgs<T, D> m) {
    __shared__ double red[32];
    const PredictArgs<T, D>& a = m.p;
    double sse = 0.0, sae = 0.0, s1 = 0.0, s2 = 0.0;
    const double pivot = (double)m.y[0];
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (i64)gridDim.x * blockDim.x) {
        int c[D];
        T w[D];
        bool all_in = true;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const T xv = a.x[d][i];
            bool inside;
            c[d] = find_cell<T>(a.mesh[d].t, a.mesh[d].K, a.mesh[d].t0, a.mesh[d].inv_h, a.mesh[d].nearly_uniform, xv, inside);
            all_in = all_in && inside;
            const int n = a.mesh[d].K;
            const T* tb = a.tab + (a.tab_off[d] + c[d]);
            w[d] = div_by_cached_rcp(xv - (T)a.mesh[d].t[c[d]], tb[6 * n], tb[7 * n]);
        }
        T mu = (T)0;
        if (all_in) {
            int base = 0;
#pragma unroll
            for (int d = 0; d < D; ++d) base += c[d] * a.stride[d];
#pragma unroll
            for (int corner = 0; corner < (1 << D); ++corner) {
                T wt = (T)1;
                int off = base;
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const bool hi = (corner >> (D - 1 - d)) & 1;
                    wt *= hi ? w[d] : ((T)1 - w[d]);
                    off += hi ? a.stride[d] : 0;
                }
                mu += wt * a.alpha[off];
            }
        }
        const double t = (double)m.y[i];
        const double e = t - (double)mu;
        sse = fma(e, e, sse);
        sae += fabs(e);
        const double cc = t - pivot;
        s1 += cc;
        s2 = fma(cc, cc, s2);
    }
    metrics_flush(sse, sae, s1, s2, red, m.out);
}


// ---------------------------------------------------------------------------------------------------------
// Min-max scaling of the reference's data preparation (src/utils/dataprocessors.py:3-44), SURVEY.md section 8f row 4:
//   min_max_scaling:  (x - min) / (max - min)        min_max_inverse:  x * (max - min) + min
// evaluated in the tensor's own dtype with separately rounded operations (no fma contraction), i.e. bit for bit what
// torch computes; min and max stay on the device (no host synchronisation between the reduction and the scaling).
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_min(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const T w = __shfl_xor_sync(0xffffffffu, v, o); v = w < v ? w : v; }
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const T w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
    return v;
}

// partial[2 b], partial[2 b + 1] = min, max over the elements block b visits.  n > 0.
template <typename T>
__global__ void __launch_bounds__(256) k_minmax_partial(const T* __restrict__ x, i64 n, T* __restrict__ partial) {
    __shared__ T smin[8], smax[8];
    T lo = x[0], hi = x[0];
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const T v = x[i];
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
    }
    lo = warp_min(lo);
    hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { lo = smin[w] < lo ? smin[w] : lo; hi = smax[w] > hi ? smax[w] : hi; }
        partial[2 * blockIdx.x] = lo;
        partial[2 * blockIdx.x + 1] = hi;
    }
}
// out[0] = min, out[1] = max over `blocks` partial pairs (one block)
template <typename T>
__global__ void __launch_bounds__(256) k_minmax_final(const T* __restrict__ partial, int blocks, T* __restrict__ out) {
    __shared__ T smin[8], smax[8];
    T lo = partial[0], hi = partial[1];
    for (int b = threadIdx.x; b < blocks; b += 256) {
        lo = partial[2 * b] < lo ? partial[2 * b] : lo;
        hi = partial[2 * b + 1] > hi ? partial[2 * b + 1] : hi;
    }
    lo = warp_min(lo);
    hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { lo = smin[w] < lo ? smin[w] : lo; hi = smax[w] > hi ? smax[w] : hi; }
        out[0] = lo;
        out[1] = hi;
    }
}

__device__ __forceinline__ float rn_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double rn_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float rn_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double rn_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float rn_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double rn_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float rn_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double rn_div(double a, double b) { return __ddiv_rn(a, b); }

// inverse == 0: y = (x - mm[0]) / (mm[1] - mm[0]);  inverse != 0: y = x * (mm[1] - mm[0]) + mm[0]
template <typename T>
__global__ void __launch_bounds__(256) k_minmax_scale(const T* __restrict__ x, i64 n, const T* __restrict__ mm, int inverse,
                                                      T* __restrict__ y) {
    const T lo = mm[0], span = rn_sub(mm[1], mm[0]);
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        y[i] = inverse ? rn_add(rn_mul(x[i], span), lo) : rn_div(rn_sub(x[i], lo), span);
}

// ---------------------------------------------------------------------------------------------------------
// Synthetic satellite-track observations generated on the device: the geometry of the
// reference's `generate_track` (src/utils/dataloaders.py:290-377; notebook call trajectory_gradient = 2) -- `passes`
// ascending passes x1 = o_j + t / g, x2 = t followed by as many descending ones (x2 = 1 - t), offsets o_j = j / passes, x1
// wrapped into [0, 1) -- in acquisition order (pass-major), a smooth field plus noise as targets, and for D = 3 the
// acquisition time as third coordinate.  Every observation is a pure function of its GLOBAL index (counter-based hash), so a
// rank generates exactly its shard [lo, hi) of the same data set whatever the sharding.  All arithmetic is float64 with
// separately rounded operations (no FMA contraction): bit for bit the torch expressions of bench.py's make_tracks_torch.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double track_hash_uniform(i64 idx, i64 salt) {
    unsigned long long h = (unsigned long long)idx * 6364136223846793005ull
                           + (1442695040888963407ull + (unsigned long long)salt * 7046029254386353131ull);
    h ^= h >> 29;
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 32;
    return (double)((h >> 11) & ((1ull << 40) - 1ull)) / 1099511627776.0;      // 2^40
}

template <typename T, int D>
struct TrackArgs {
    T* x[D];
    T* y;
    i64 lo, hi, n_total;
    i64 seed;
    int passes;
    double gradient;
};

template <typename T, int D>
__global__ void __launch_bounds__(256) k_generate_tracks(const __grid_constant__ TrackArgs<T, D> a) {
    i64 per_pass = a.n_total / (2 * (i64)a.passes);
    if (per_pass < 1) per_pass = 1;
    const double noise_c = 0.05 * 1.7320508075688772;          // 0.05 * sqrt(3): sum of 4 uniforms has variance 1 / 3
    for (i64 i = a.lo + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < a.hi; i += (i64)gridDim.x * blockDim.x) {
        i64 j = i / per_pass;
        if (j > 2 * (i64)a.passes - 1) j = 2 * (i64)a.passes - 1;
        const i64 k = i - j * per_pass;
        const double jitter = track_hash_uniform(i, 2 * a.seed + 1);
        double t = __ddiv_rn(__dadd_rn((double)k, jitter), (double)per_pass);
        t = fmin(fmax(t, 0.0), 1.0);
        const bool asc = j < a.passes;
        const double off = __ddiv_rn((double)(j % a.passes), (double)a.passes);
        double x1 = __dadd_rn(off, __ddiv_rn(t, a.gradient));
        x1 = __dsub_rn(x1, floor(x1));
        const double x2 = asc ? t : __dsub_rn(1.0, t);
        double u = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) u = __dadd_rn(u, track_hash_uniform(i, 2 * a.seed + 10 + q));
        u = __dsub_rn(u, 2.0);
        double f = __dadd_rn(sin(__dmul_rn(5.0, x1)), cos(__dmul_rn(7.0, x2)));
        f = __dadd_rn(f, __dmul_rn(0.5, sin(__dmul_rn(15.0, x1))));
        f = __dadd_rn(f, __dmul_rn(0.5, cos(__dmul_rn(12.0, x2))));
        double yv = __dadd_rn(f, __dmul_rn(noise_c, u));
        const i64 o = i - a.lo;
        a.x[0][o] = (T)x1;
        if (D >= 2) a.x[D >= 2 ? 1 : 0][o] = (T)x2;
        if (D == 3) {
            double x3 = __ddiv_rn(__dadd_rn((double)i, track_hash_uniform(i, 2 * a.seed + 31)), (double)a.n_total);
            x3 = fmin(fmax(x3, 0.0), 1.0);
            yv = __dadd_rn(yv, __dmul_rn(__dmul_rn(0.3, sin(__dmul_rn(4.0, x3))), cos(__dmul_rn(3.0, x1))));
            a.x[D - 1][o] = (T)x3;
        }
        a.y[o] = (T)yv;
    }
}

}  // namespace vggp
