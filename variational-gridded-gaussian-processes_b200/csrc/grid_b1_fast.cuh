// k_fibre_pass_fast: the fibre engine of grid_b1.cuh specialised for M_d <= 512 (every BASELINE.json grid) and tiles of
// 8 fibres.  Same tasks, same results; what changes is the instruction count -- the generic kernel is issue-bound (ncu +
// clock64 stamps, round 2: ~1500 instructions per warp in the recurrence phase alone, two thirds of them selects,
// predicates and index arithmetic):
//   * every lane owns exactly 16 consecutive elements of a fibre (lane stride 17 in shared memory), whatever n <= 512 is;
//     elements past n carry the IDENTITY of the recurrences (x = 0, rl = ru = 1), so the sweeps are straight-line code
//     without a single select or branch;
//   * a thread's elements are an arithmetic progression in both the global and the shared address (element step 32 .. 256,
//     a multiple of 16), so all index arithmetic is two additions per element;
//   * the generator loads, the tile loads and (dL tasks) the band loads are all issued before the first shared-memory store;
//   * FIBRE PACKING (mode-product kinds, FpTask::pk): for M_d <= 256 a 512-slot row holds 2^pk real fibres side by side, each in
//     its own window of npad = 512 >> pk slots.  The recurrences need no change: the generators are staged periodically and
//     rl[n - 1] = 0 = ru[-1], so nothing carries from one fibre into the next, and identity slots between fibres carry zeros.
//     A 3-D 256 x 256 x 64 grid swept 512 slots for 64 live elements (8 x the work, 8 x the CTAs) before this.
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
template <bool FULLWIN = false>
__device__ __forceinline__ void ff_fibre(const double* __restrict__ pd, const double* __restrict__ rl,
                                         const double* __restrict__ rup, double* __restrict__ X, int lane, int n, int mask) {
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
        if (FULLWIN || ((i0 + j) & mask) < n) X[p0 + j] = sv[j] + l + uu[j];
        l = fma(r, l, r * sv[j]);
    }
}

// acc[dl + 1][i] += sum over the rows [f0, f1) and the 2^lg-slot windows of a row of Y[f][w + i] * Cx[f - f0][w + i + dl].
// `scr` (3 x 512 doubles: the generator arrays, free after the recurrences) holds the per-window partial sums so that one
// atomic per (dl, i) leaves the CTA however many fibres a row packs.
__device__ __forceinline__ void ff_band_dots(const double* __restrict__ Y, const double* __restrict__ Cx, int f0, int f1, int n,
                                             int lg, double* __restrict__ scr, double* __restrict__ acc, double* __restrict__ det) {
    const int mask = (1 << lg) - 1;
    const bool packed = lg < 9;
    for (int s_ = threadIdx.x; s_ < 512; s_ += FP_THREADS) {
        const int i = s_ & mask;
        double am = 0.0, a0 = 0.0, ap = 0.0;
        if (i < n) {
            const int p = ff_p(s_);
            const int pm = (i > 0) ? ff_p(s_ - 1) : p, pp = (i + 1 < n) ? ff_p(s_ + 1) : p;
            for (int f = f0; f < f1; ++f) {
                const double y = Y[f * FF_PITCH + p];
                const double* c = Cx + (f - f0) * FF_PITCH;
                am = fma(y, c[pm], am);
                a0 = fma(y, c[p], a0);
                ap = fma(y, c[pp], ap);
            }
            if (!packed) {
                if (det) {          // deterministic mode: this CTA's partials, summed in tile order by k_fp_det_reduce
                    det[FP_DET_BAND + i] = am; det[FP_DET_BAND + 512 + i] = a0; det[FP_DET_BAND + 1024 + i] = ap;
                } else {
                    if (i > 0) atomicAdd(acc + i, am);
                    atomicAdd(acc + n + i, a0);
                    if (i + 1 < n) atomicAdd(acc + 2 * n + i, ap);
                }
            }
        }
        if (packed) { scr[s_] = am; scr[512 + s_] = a0; scr[1024 + s_] = ap; }
    }
    if (!packed) return;
    __syncthreads();
    for (int k = 0; k < 3; ++k)
        for (int i = threadIdx.x; i < n; i += FP_THREADS) {
            if ((k == 0 && i == 0) || (k == 2 && i + 1 >= n)) continue;
            double v = 0.0;
            for (int w = 0; w < 512; w += mask + 1) v += scr[k * 512 + w + i];
            if (det) det[FP_DET_BAND + k * 512 + i] = v;
            else atomicAdd(acc + k * n + i, v);
        }
}

// SIMPLE: every slot of the tile is a live element -- n fills its window exactly (n == npad: 64, 128, 256, 512) and the tile
// holds all its fibres -- and the per-stage address step fits 32 bits.  Validity is then `slot < 512` and the address an
// arithmetic progression, which removes most of the integer work that dominated the big passes (ncu, round 2, 3-D backward
// pass: 37 M warp instructions of which 13 % were float64 arithmetic and over half index / predicate arithmetic).
template <typename T, bool SIMPLE, int NU>
__device__ __forceinline__ void ff_pass_body(const FpPass& P, const FpTask& tk, const int tile, unsigned char* smraw, double* red) {
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
    const int nsrc = (kind == FP_GA) ? FF_F / 2 : FF_F;               // source rows of a tile
    const int sh = (kind == FP_GA) ? 2 : 3;
    // fibre packing: a row holds 2^pk real fibres in windows of npad = 512 >> pk slots (pk = 0: one fibre per row)
    const int pk = tk.pk, lg = 9 - pk, npad = 1 << lg, mask = npad - 1;
    const i64 fib0 = (i64)tile * (nsrc << pk);                        // first real fibre of the tile
    const int nf = (int)min((i64)(nsrc << pk), tk.nfib - fib0);       // real fibres of the tile
    // rows that hold at least one of them (contiguous: fibre q in row q >> pk; strided: in row q mod nsrc)
    const int nrows = contiguous ? (nf + (1 << pk) - 1) >> pk : min(nf, nsrc);
    // ---- this thread's slots: row f, slots s0, s0 + ds, ... (< 512): arithmetic progressions in the shared AND the global
    // address, up to a correction per window for non-power-of-two n in the contiguous case:
    //   slot(u) = s0 + u ds,   element = slot & mask,   window = slot >> lg,
    //   global address = a0 + u da - window * gap,   real fibre = fid0 + window * fstep
    int f, s0, ds, gap, fstep;
    i64 a0, da, fid0;
    if (contiguous) {                               // FP_THREADS / nsrc consecutive threads walk one row
        ds = FP_THREADS >> sh;
        f = tid >> (8 - sh);
        s0 = tid & (ds - 1);
        fid0 = fib0 + ((i64)f << pk);
        a0 = fid0 * (i64)n + s0;
        da = ds;
        gap = npad - n;
        fstep = 1;
    } else {                                        // nsrc 2^pk consecutive threads take the tile's fibres at one element index
        const int q = tid & ((nsrc << pk) - 1);     // real fibre of the tile; consecutive q are adjacent in memory
        f = q & (nsrc - 1);
        const int e0 = tid >> (sh + pk);
        ds = FP_THREADS >> (sh + pk);
        s0 = (q >> sh) * npad + e0;
        fid0 = fib0 + q;
        const unsigned inner = (unsigned)tk.inner;
        const unsigned o = (unsigned)fid0 / inner, r = (unsigned)fid0 - o * inner;      // fid0, inner < 2^31
        a0 = ((i64)o * n + e0) * (i64)inner + r;
        da = (i64)ds * (i64)inner;
        gap = 0;
        fstep = 0;
    }
    const int xrow = f * FF_PITCH;
    const i64 nfib = tk.nfib;
    const int da32 = (int)da;
    const int so0 = xrow + ff_p(s0), dso = ds + (ds >> 4);      // SIMPLE: ds is a multiple of 16, so the padded position is affine in u
#define FF_SLOT(u) (s0 + (u) * ds)
#define FF_IN(u) (SIMPLE ? ((u) < NU) : (FF_SLOT(u) < 512))
#define FF_OK(u) (SIMPLE ? ((u) < NU) : (FF_SLOT(u) < 512 && (FF_SLOT(u) & mask) < n && fid0 + (i64)((FF_SLOT(u) >> lg) * fstep) < nfib))
#define FF_GA(u) (SIMPLE ? a0 + (i64)((u) * da32) : a0 + (u) * da - (i64)((FF_SLOT(u) >> lg) * gap))
#define FF_SO(u) (SIMPLE ? so0 + (u) * dso : xrow + ff_p(FF_SLOT(u)))
    // R / DL tasks (pk == 0, strided in the factor): column k = fib0 + f, rows i0 + u di
    const int i0 = s0, di = ds;
    const bool live = fib0 + f < nfib;
    constexpr int MAXU = 16;                        // n / di <= 512 / 32
    // ---- issue every global load of the prologue, then store
    const double* __restrict__ gen = P.gen[d];
    double g0[2], g1[2], g2[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int i = (tid + u * FP_THREADS) & mask;                // periodic over the windows of a row
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
            for (int u = 0; u < MAXU; ++u) v[u] = FF_OK(u) ? src[FF_GA(u)] : 0.0;
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
                const bool ok = FF_OK(u);
                gv[u8] = ok ? (double)ga[FF_GA(u)] : 0.0;
                mv[u8] = ok ? m[FF_GA(u)] : 0.0;
                av[u8] = ok ? al[FF_GA(u)] : 0.0;
            }
#pragma unroll
            for (int u8 = 0; u8 < 8; ++u8) {
                const int u = h * 8 + u8;
                if (FF_IN(u)) {                        // all 512 slots are written: zeros past n and for absent fibres
                    const double gg = cg * gv[u8], hh = gg - 0.5 * mv[u8];
                    if (both) {
                        X[FF_SO(u)] = gg;
                        X[FF_SO(u) + nsrc * FF_PITCH] = hh;
                    } else {
                        X[FF_SO(u)] = hh;
                    }
                    Cx[FF_SO(u)] = av[u8];
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
            if (FF_IN(u)) Cx[FF_SO(u)] = v[u];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = tid + u * FP_THREADS;
            if (e < 2 * n) bqs[(e < n) ? e : 512 + (e - n)] = 2.0 * cQ * (double)bq[u];
        }
    } else {
#pragma unroll
        for (int u = 0; u < MAXU; ++u)
            if (FF_IN(u)) X[FF_SO(u)] = v[u];
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
                    const int p = ff_p(i);
                    r = bqs[i] * c[p];
                    if (i > 0) r = fma(bqs[512 + i - 1], c[ff_p(i - 1)], r);
                    if (i + 1 < n) r = fma(bqs[512 + i], c[ff_p(i + 1)], r);
                }
                X[FF_SO(u)] = r;
            }
        }
        __syncthreads();
    }
    // ---- recurrences: one warp per fibre
    ff_fibre<SIMPLE>(pd, rl, rup, X + warp * FF_PITCH, lane, n, mask);
    __syncthreads();
    // ---- epilogues
    switch (kind) {
        case FP_R: {
            double* __restrict__ R = tk.o0;
            const i64 k = fib0 + f;
#pragma unroll
            for (int u = 0; u < MAXU; ++u) {
                const int i = i0 + u * di;
                if (live && i < n) R[(i64)i * n + k] = X[FF_SO(u)];
            }
        } break;
        case FP_PROD: {
            double* __restrict__ dst = tk.o0;
#pragma unroll
            for (int u = 0; u < MAXU; ++u)
                if (FF_OK(u)) dst[FF_GA(u)] = X[FF_SO(u)];
        } break;
        case FP_DM: case FP_ALPHA: {
            double* __restrict__ dst = tk.o0;
            const double* __restrict__ al = tk.s1;            // DM: alpha; ALPHA: m
            T* __restrict__ aT = reinterpret_cast<T*>(tk.t1);
            double w[MAXU];
#pragma unroll
            for (int u = 0; u < MAXU; ++u) w[u] = FF_OK(u) ? al[FF_GA(u)] : 0.0;
            double dot = 0.0;
#pragma unroll
            for (int u = 0; u < MAXU; ++u) {
                if (FF_OK(u)) {
                    const double y = X[FF_SO(u)];
                    if (kind == FP_DM) {
                        dst[FF_GA(u)] = y - w[u];
                    } else {
                        dst[FF_GA(u)] = y;
                        aT[FF_GA(u)] = (T)y;
                        dot = fma(y, w[u], dot);
                    }
                }
            }
            if (kind == FP_ALPHA) {
                dot = block_sum(dot, red);
                if (tid == 0) {
                    if (P.det) P.det[(i64)blockIdx.x * FP_DET_SLOT + FP_DET_MALPHA] = dot;
                    else atomicAdd(P.sc + SC_MALPHA, dot);
                }
            }
        } break;
        case FP_GA: {
            double* __restrict__ dst = tk.o0;
#pragma unroll
            for (int u = 0; u < MAXU; ++u) {
                if (FF_OK(u)) {
                    const double y = X[FF_SO(u)];
                    dst[FF_GA(u)] = tk.direct ? y - Cx[FF_SO(u)] : y;            // Cx holds alpha of these fibres
                }
            }
            ff_band_dots(X, Cx, nsrc, nsrc + nrows, n, lg, pd, P.acc[d], P.det ? P.det + (i64)blockIdx.x * FP_DET_SLOT : nullptr);       // the generator arrays are scratch by now
        } break;
        case FP_GAONLY:
            ff_band_dots(X, Cx, 0, nrows, n, lg, pd, P.acc[d], P.det ? P.det + (i64)blockIdx.x * FP_DET_SLOT : nullptr);
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
                        vv = X[FF_SO(u)] - trO * Cx[FF_SO(u)];
                        if ((i64)i == k) vv += ratio / Lkk;
                    }
                    dL[(i64)i * n + k] = vv;
                }
            }
            ff_band_dots(X, Cx, 0, nf, n, 9, pd, P.acc[d], P.det ? P.det + (i64)blockIdx.x * FP_DET_SLOT : nullptr);
        } break;
        default: break;
    }
#undef FF_SLOT
#undef FF_IN
#undef FF_OK
#undef FF_GA
#undef FF_SO
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
    // uniform over the CTA
    const int kind = tk.kind;
    bool simple = false;
    if (kind == FP_PROD || kind == FP_ALPHA || kind == FP_DM || kind == FP_GA || kind == FP_GAONLY) {
        const int nsrc = (kind == FP_GA) ? FF_F / 2 : FF_F;
        const int npad = 512 >> tk.pk;
        const i64 per_tile = (i64)nsrc << tk.pk;
        const i64 step = (tk.inner == 1) ? (i64)(FP_THREADS >> (kind == FP_GA ? 2 : 3))
                                         : (i64)(FP_THREADS >> ((kind == FP_GA ? 2 : 3) + tk.pk)) * tk.inner;
        const int ds = (tk.inner == 1) ? (FP_THREADS >> (kind == FP_GA ? 2 : 3)) : (FP_THREADS >> ((kind == FP_GA ? 2 : 3) + tk.pk));
        simple = tk.n == npad && (i64)(tile + 1) * per_tile <= tk.nfib && step * 16 < ((i64)1 << 31) && (ds & 15) == 0 &&
                 (kind != FP_GA || tk.inner == 1);
    }
    // live elements per thread: 512 / ds = 8 for the GA kind (4 source rows, contiguous), 16 otherwise
    if (simple && kind == FP_GA) ff_pass_body<T, true, 8>(P, tk, tile, smraw, red);
    else if (simple) ff_pass_body<T, true, 16>(P, tk, tile, smraw, red);
    else ff_pass_body<T, false, 16>(P, tk, tile, smraw, red);
}

// Deterministic mode: sum the per-CTA partials of the pass that just ran, tiles in order, into the accumulators the atomics
// would have hit.  grid (D, FP_DET_YB): CTA (d, y) walks the tasks of dimension d in task order (two tasks of a pass may feed
// the same band); a band element is owned by exactly one thread of one CTA, the scalar sums by CTA (d, 0).
constexpr int FP_DET_YB = 6;        // 6 x 256 threads = one thread per band element at n = 512
__global__ void __launch_bounds__(256) k_fp_det_reduce(const __grid_constant__ FpPass P) {
    const int d = blockIdx.x;
    for (int ti = 0; ti < P.ntasks; ++ti) {
        const FpTask& tk = P.t[ti];
        if (tk.d != d) continue;
        const int n = tk.n, kind = tk.kind;
        const double* base = P.det + (i64)tk.tile0 * FP_DET_SLOT;
        if (kind == FP_GA || kind == FP_GAONLY || kind == FP_DL) {
            for (int e = blockIdx.y * 256 + threadIdx.x; e < 3 * n; e += 256 * FP_DET_YB) {
                const int k = e / n, i = e - k * n;
                if ((k == 0 && i == 0) || (k == 2 && i + 1 >= n)) continue;
                const double* src = base + FP_DET_BAND + k * 512 + i;
                double v = 0.0;
                int t = 0;
                for (; t + 16 <= tk.ntiles; t += 16) {      // sixteen loads in flight, added in tile order
                    double q[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) q[u] = src[(i64)(t + u) * FP_DET_SLOT];
#pragma unroll
                    for (int u = 0; u < 16; ++u) v += q[u];
                }
                for (; t < tk.ntiles; ++t) v += src[(i64)t * FP_DET_SLOT];
                P.acc[d][e] += v;
            }
        } else if (blockIdx.y != 0) {
            continue;
        } else if (kind == FP_ALPHA) {
            if (threadIdx.x == 0) {
                double v = 0.0;
                for (int t = 0; t < tk.ntiles; ++t) v += base[(i64)t * FP_DET_SLOT + FP_DET_MALPHA];
                P.sc[SC_MALPHA] += v;
            }
        } else if (kind == FP_QROW) {
            // rows i = t, t + 256, ... per thread, then a fixed-shape tree: the order depends on n only
            __shared__ double sh[2][256];
            double tr = 0.0, ld = 0.0;
            for (int i = threadIdx.x; i < n; i += 256) {
                const double* q = base + (i64)(i / FP_WARPS) * FP_DET_SLOT + (i % FP_WARPS);
                tr += q[FP_DET_TR];
                ld += q[FP_DET_LOGDET];
            }
            sh[0][threadIdx.x] = tr; sh[1][threadIdx.x] = ld;
            __syncthreads();
            for (int o = 128; o > 0; o >>= 1) {
                if ((int)threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; }
                __syncthreads();
            }
            if (threadIdx.x == 0) { P.sc[SC_TR + d] += sh[0][0]; P.sc[SC_LOGDETS + d] += sh[1][0]; }
            __syncthreads();
        }
    }
}

}  // namespace vggp
