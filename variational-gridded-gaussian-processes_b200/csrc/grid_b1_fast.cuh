// k_fibre_pass_fast: the fibre engine of grid_b1.cuh specialised for M_d <= 512 (every BASELINE.json grid) and tiles of
// 8 fibres.  Same tasks, same results; what changes is the instruction count -- the generic kernel is issue-bound (ncu +
// clock64 stamps, round 2: ~1500 instructions per warp in the recurrence phase alone, two thirds of them selects,
// predicates and index arithmetic):
//   * every lane owns exactly 16 consecutive elements of a fibre (lane stride 17 in shared memory), whatever n <= 512 is;
//     elements past n carry the IDENTITY of the recurrences (x = 0, rl = ru = 1), so the sweeps are straight-line code
//     without a single select or branch;
//   * a thread's elements are an arithmetic progression in both the global and the shared address (element step 32 .. 256,
//     a multiple of 16), so all index arithmetic is two additions per element;
//   * the generator loads, the tile loads and (dL tasks) the band loads are all issued before the first shared-memory store.
#pragma once
#include "grid_b1.cuh"

namespace vggp {

constexpr int FF_S = 16;                     // elements per lane
constexpr int FF_LS = 17;                    // lane stride in shared memory
constexpr int FF_PITCH = 32 * FF_LS + 2;     // 546: consecutive fibres start 2 banks (of 8 bytes) apart
constexpr int FF_F = 8;                      // fibres per tile

__host__ __device__ inline size_t ff_smem_bytes(bool aux) {
    return sizeof(double) * ((size_t)3 * FF_PITCH + 2 * 512 + (size_t)(aux ? 2 : 1) * FF_F * FF_PITCH);
}
__device__ __forceinline__ int ff_p(int i) { return i + (i >> 4); }

// one warp, one fibre (see fp_fibre): X holds x_i on entry and y_i = (P x)_i on exit; pd / rl / rup are padded with the
// identity (0, 1, 1) past n and rup[p(i)] = ru[i - 1] (0 for i = 0)
__device__ __forceinline__ void ff_fibre(const double* __restrict__ pd, const double* __restrict__ rl,
                                         const double* __restrict__ rup, double* __restrict__ X, int lane, int n) {
    const int p0 = lane * FF_LS;
    double sv[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) sv[j] = pd[p0 + j] * X[p0 + j];
    double lA = 1.0, lB = 0.0, uA = 1.0, uB = 0.0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const double r = rl[p0 + j];
        lB = fma(r, lB, r * sv[j]);
        lA *= r;
    }
#pragma unroll
    for (int j = 15; j >= 0; --j) {
        const double r = rup[p0 + j];
        uB = fma(r, uB, r * sv[j]);
        uA *= r;
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double Ap = __shfl_sync(0xffffffffu, lA, (lane - o) & 31), Bp = __shfl_sync(0xffffffffu, lB, (lane - o) & 31);
        const double Aq = __shfl_sync(0xffffffffu, uA, (lane + o) & 31), Bq = __shfl_sync(0xffffffffu, uB, (lane + o) & 31);
        if (lane >= o) { lB = fma(lA, Bp, lB); lA *= Ap; }
        if (lane + o < 32) { uB = fma(uA, Bq, uB); uA *= Aq; }
    }
    double l = __shfl_sync(0xffffffffu, lB, (lane - 1) & 31);
    double u = __shfl_sync(0xffffffffu, uB, (lane + 1) & 31);
    if (lane == 0) l = 0.0;
    if (lane == 31) u = 0.0;
    double uu[16];
#pragma unroll
    for (int j = 15; j >= 0; --j) {
        uu[j] = u;
        const double r = rup[p0 + j];
        u = fma(r, u, r * sv[j]);
    }
    const int i0 = lane * FF_S;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const double r = rl[p0 + j];
        if (i0 + j < n) X[p0 + j] = sv[j] + l + uu[j];
        l = fma(r, l, r * sv[j]);
    }
}

// acc[dl + 1][i] += sum over the fibres [f0, f1) of Y[f][i] * Cx[f - f0][i + dl]
__device__ __forceinline__ void ff_band_dots(const double* __restrict__ Y, const double* __restrict__ Cx, int f0, int f1, int n,
                                             double* __restrict__ acc) {
    for (int i = threadIdx.x; i < n; i += FP_THREADS) {
        const int p = ff_p(i);
        const int pm = (i > 0) ? ff_p(i - 1) : p, pp = (i + 1 < n) ? ff_p(i + 1) : p;
        double am = 0.0, a0 = 0.0, ap = 0.0;
        for (int f = f0; f < f1; ++f) {
            const double y = Y[f * FF_PITCH + p];
            const double* c = Cx + (f - f0) * FF_PITCH;
            am = fma(y, c[pm], am);
            a0 = fma(y, c[p], a0);
            ap = fma(y, c[pp], ap);
        }
        if (i > 0) atomicAdd(acc + i, am);
        atomicAdd(acc + n + i, a0);
        if (i + 1 < n) atomicAdd(acc + 2 * n + i, ap);
    }
}

template <typename T>
__global__ void __launch_bounds__(FP_THREADS, 2) k_fibre_pass_fast(const __grid_constant__ FpPass P) {
    extern __shared__ __align__(128) unsigned char smraw[];
    __shared__ double red[32];
    int ti = 0;
    while (ti + 1 < P.ntasks && (int)blockIdx.x >= P.t[ti + 1].tile0) ++ti;
    const FpTask& tk = P.t[ti];
    const int tile = (int)blockIdx.x - tk.tile0;
    if (tile >= tk.ntiles) return;
    if (tk.kind == FP_QROW) { fp_qrow<T>(P, tk, tile); return; }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = tk.n, d = tk.d, kind = tk.kind;
    double* pd = reinterpret_cast<double*>(smraw);
    double* rl = pd + FF_PITCH;
    double* rup = rl + FF_PITCH;
    double* bqs = rup + FF_PITCH;                    // DL: 2 cQ [bq_diag (512) | bq_off (512)]
    double* X = bqs + 2 * 512;
    double* Cx = X + FF_F * FF_PITCH;                // kinds with auxiliary fibres only
    const double noise = P.theta[2 * P.D];
    const double cg = P.ell_scale / noise;
    const double cQ = -P.ell_scale / (2.0 * noise);
    const bool contiguous = (tk.inner == 1);
    const int nsrc = (kind == FP_GA) ? FF_F / 2 : FF_F;               // source fibres of a tile
    const int sh = (kind == FP_GA) ? 2 : 3;
    const i64 fib0 = (i64)tile * nsrc;
    const int nf = (int)min((i64)nsrc, tk.nfib - fib0);
    // ---- this thread's elements: fibre f, elements i0, i0 + di, ... (< n): arithmetic progressions everywhere
    int f, i0, di;
    i64 a0, da;
    if (contiguous) {                               // FP_THREADS / nsrc consecutive threads walk one fibre
        di = FP_THREADS >> sh;
        f = tid >> (8 - sh);
        i0 = tid & (di - 1);
        a0 = (fib0 + f) * (i64)n + i0;
        da = di;
    } else {                                        // nsrc consecutive threads take the nsrc fibres of one element index
        di = FP_THREADS >> sh;
        f = tid & (nsrc - 1);
        i0 = tid >> sh;
        const i64 fg = fib0 + f;
        const unsigned inner = (unsigned)tk.inner;
        const unsigned o = (unsigned)fg / inner, r = (unsigned)fg - o * inner;      // fg, inner < 2^31
        a0 = ((i64)o * n + i0) * (i64)inner + r;
        da = (i64)di * (i64)inner;
    }
    const bool live = f < nf;
    const int p0 = ff_p(i0), dp = di + (di >> 4);
    const int xo = f * FF_PITCH + p0;               // shared offset of (f, i0)
    constexpr int MAXU = 16;                        // n / di <= 512 / 32
    // ---- issue every global load of the prologue, then store
    const double* __restrict__ gen = P.gen[d];
    double g0[2], g1[2], g2[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int i = tid + u * FP_THREADS;
        g0[u] = (i < n) ? gen[i] : 0.0;                              // pd
        g1[u] = (i < n) ? gen[2 * n + i] : 1.0;                      // rl (0 at n - 1)
        g2[u] = (i < n) ? (i > 0 ? gen[n + i - 1] : 0.0) : 1.0;      // ru shifted by one
    }
    double v[MAXU];
    switch (kind) {
        case FP_R: {
            const double* __restrict__ L = tk.s0;
            const i64 k = fib0 + f;
#pragma unroll
            for (int u = 0; u < MAXU; ++u) {
                const int i = i0 + u * di;
                v[u] = (live && i < n && (i64)i >= k) ? L[(i64)i * n + k] : 0.0;
            }
        } break;
        case FP_PROD: case FP_ALPHA: case FP_DM: {
            const double* __restrict__ src = tk.s0;
#pragma unroll
            for (int u = 0; u < MAXU; ++u) v[u] = (live && i0 + u * di < n) ? src[a0 + u * da] : 0.0;
        } break;
        case FP_DL: {
            const double* __restrict__ R = tk.s0;
            const i64 k = fib0 + f;
#pragma unroll
            for (int u = 0; u < MAXU; ++u) {
                const int i = i0 + u * di;
                v[u] = (live && i < n) ? R[(i64)i * n + k] : 0.0;
            }
        } break;
        default: break;
    }
    if (kind == FP_GA || kind == FP_GAONLY) {
        const T* __restrict__ ga = reinterpret_cast<const T*>(tk.t0);
        const double* __restrict__ m = tk.s0;
        const double* __restrict__ al = tk.s1;
        const bool both = (kind == FP_GA);
        // two batches of 8 elements x 3 loads (register budget)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            double gv[8], mv[8], av[8];
#pragma unroll
            for (int u8 = 0; u8 < 8; ++u8) {
                const int u = h * 8 + u8;
                const bool ok = live && i0 + u * di < n;
                gv[u8] = ok ? (double)ga[a0 + u * da] : 0.0;
                mv[u8] = ok ? m[a0 + u * da] : 0.0;
                av[u8] = ok ? al[a0 + u * da] : 0.0;
            }
#pragma unroll
            for (int u8 = 0; u8 < 8; ++u8) {
                const int u = h * 8 + u8;
                if (i0 + u * di < 512) {               // all 512 slots are written: zeros past n and for absent fibres
                    const double gg = cg * gv[u8], hh = gg - 0.5 * mv[u8];
                    if (both) {
                        X[xo + u * dp] = gg;
                        X[xo + nsrc * FF_PITCH + u * dp] = hh;
                    } else {
                        X[xo + u * dp] = hh;
                    }
                    Cx[xo + u * dp] = av[u8];
                }
            }
        }
    } else if (kind == FP_DL) {
        const T* __restrict__ bnd = reinterpret_cast<const T*>(tk.t0) + 2 * n;        // [bq_diag | bq_off]
        T bq[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = tid + u * FP_THREADS;
            bq[u] = (e < 2 * n) ? bnd[e] : (T)0;
        }
#pragma unroll
        for (int u = 0; u < MAXU; ++u)
            if (i0 + u * di < 512) Cx[xo + u * dp] = v[u];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = tid + u * FP_THREADS;
            if (e < 2 * n) bqs[(e < n) ? e : 512 + (e - n)] = 2.0 * cQ * (double)bq[u];
        }
    } else {
#pragma unroll
        for (int u = 0; u < MAXU; ++u)
            if (i0 + u * di < 512) X[xo + u * dp] = v[u];
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int i = tid + u * FP_THREADS;       // all 512 element slots: identity past n
        const int p = ff_p(i);
        pd[p] = g0[u]; rl[p] = g1[u]; rup[p] = g2[u];
    }
    __syncthreads();
    if (kind == FP_DL) {
        // column k of dR_d = 2 cQ tridiag(bq) R_d from the staged column of R_d
        const double* c = Cx + f * FF_PITCH;
#pragma unroll
        for (int u = 0; u < MAXU; ++u) {
            const int i = i0 + u * di;
            if (i < 512) {
                double r = 0.0;
                if (i < n) {
                    const int p = p0 + u * dp;
                    r = bqs[i] * c[p];
                    if (i > 0) r = fma(bqs[512 + i - 1], c[ff_p(i - 1)], r);
                    if (i + 1 < n) r = fma(bqs[512 + i], c[ff_p(i + 1)], r);
                }
                X[xo + u * dp] = r;
            }
        }
        __syncthreads();
    }
    // ---- recurrences: one warp per fibre
    ff_fibre(pd, rl, rup, X + warp * FF_PITCH, lane, n);
    __syncthreads();
    // ---- epilogues
    switch (kind) {
        case FP_R: {
            double* __restrict__ R = tk.o0;
            const i64 k = fib0 + f;
#pragma unroll
            for (int u = 0; u < MAXU; ++u) {
                const int i = i0 + u * di;
                if (live && i < n) R[(i64)i * n + k] = X[xo + u * dp];
            }
        } break;
        case FP_PROD: {
            double* __restrict__ dst = tk.o0;
#pragma unroll
            for (int u = 0; u < MAXU; ++u)
                if (live && i0 + u * di < n) dst[a0 + u * da] = X[xo + u * dp];
        } break;
        case FP_DM: case FP_ALPHA: {
            double* __restrict__ dst = tk.o0;
            const double* __restrict__ al = tk.s1;            // DM: alpha; ALPHA: m
            T* __restrict__ aT = reinterpret_cast<T*>(tk.t1);
            double w[MAXU];
#pragma unroll
            for (int u = 0; u < MAXU; ++u) w[u] = (live && i0 + u * di < n) ? al[a0 + u * da] : 0.0;
            double dot = 0.0;
#pragma unroll
            for (int u = 0; u < MAXU; ++u) {
                if (live && i0 + u * di < n) {
                    const double y = X[xo + u * dp];
                    if (kind == FP_DM) {
                        dst[a0 + u * da] = y - w[u];
                    } else {
                        dst[a0 + u * da] = y;
                        aT[a0 + u * da] = (T)y;
                        dot = fma(y, w[u], dot);
                    }
                }
            }
            if (kind == FP_ALPHA) {
                dot = block_sum(dot, red);
                if (tid == 0) atomicAdd(P.sc + SC_MALPHA, dot);
            }
        } break;
        case FP_GA: {
            double* __restrict__ dst = tk.o0;
#pragma unroll
            for (int u = 0; u < MAXU; ++u) {
                if (live && i0 + u * di < n) {
                    const double y = X[xo + u * dp];
                    dst[a0 + u * da] = tk.direct ? y - Cx[xo + u * dp] : y;      // Cx holds alpha of these fibres
                }
            }
            ff_band_dots(X, Cx, nsrc, nsrc + nf, n, P.acc[d]);
        } break;
        case FP_GAONLY:
            ff_band_dots(X, Cx, 0, nf, n, P.acc[d]);
            break;
        case FP_DL: {
            double* __restrict__ dL = tk.o0;
            const double* __restrict__ L = tk.s1;
            double trO = 1.0;
            for (int e2 = 0; e2 < P.D; ++e2)
                if (e2 != d) trO *= P.sc[SC_TR + e2];
            const double ratio = (double)P.M / (double)n;
            const i64 k = fib0 + f;
            const double Lkk = (live) ? L[k * n + k] : 1.0;
#pragma unroll
            for (int u = 0; u < MAXU; ++u) {
                const int i = i0 + u * di;
                if (live && i < n) {
                    double vv = 0.0;
                    if ((i64)i >= k) {
                        vv = X[xo + u * dp] - trO * Cx[xo + u * dp];
                        if ((i64)i == k) vv += ratio / Lkk;
                    }
                    dL[(i64)i * n + k] = vv;
                }
            }
            ff_band_dots(X, Cx, 0, nf, n, P.acc[d]);
        } break;
        default: break;
    }
}

}  // namespace vggp
