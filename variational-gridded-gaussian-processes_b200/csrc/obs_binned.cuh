// K1 "binned": fused per-observation ELBO forward + backward for the B1 (ASVGP) family over the binned layout of
// binplan.hpp.  Same outputs as k_obs_b1 (obs.cuh) into the same gbuf; what changes is who owns a cell:
//
//   * one lane walks one RUN = (a share of) the observations of ONE grid cell, and the 32 runs of a warp task have
//     the same padded length, so the warp enters its cells, streams, and flushes in lock step -- no per-observation
//     cell test, no divergent cell switch;
//   * the variance terms are accumulated as MOMENTS of the hat weights, sum_n prod_d a_{n,d}^{k_d} (k_d = 0..2),
//     instead of evaluating p_d, q_d per observation: every band sum of P_d and Q_d and sum_n (prod q - prod p) is a
//     fixed linear combination of these 3^D numbers with the per-cell band polynomials as coefficients, applied once
//     per run at the flush.  Straight-line cost per observation in 2-D: 28 FP32 instructions (8 weights, 4 mean,
//     5 residual sums, 1 r^2, 10 moments) against ~50 in k_obs_b1;
//   * padding slots hold x_d = lower knot of the run's cell (weight a_d = 0 exactly: no moment is touched) and the
//     residual of slots past the run's length is forced to 0, so padding contributes nothing;
//   * observations outside every mesh are not streamed at all: their only contribution, sum y^2, is a constant of the
//     data set computed at packing time.
//
// The per-lane arithmetic is __host__ __device__ so that tests/host_emul can execute exactly this code on the CPU
// (test infrastructure only: libvggp.so exports no host path).
#pragma once
#include <stdint.h>
#include <math.h>
#include "binplan.hpp"

#if defined(__CUDACC__)
#define VGGP_HD __host__ __device__ __forceinline__
#else
#define VGGP_HD inline
#endif

namespace vggp {

template <int D> struct Pow3 { static constexpr int v = 3 * Pow3<D - 1>::v; };
template <> struct Pow3<0> { static constexpr int v = 1; };

VGGP_HD float bin_fma(float a, float b, float c) { return fmaf(a, b, c); }
VGGP_HD double bin_fma(double a, double b, double c) { return fma(a, b, c); }

// correctly rounded u / h from the correctly rounded reciprocal (same scheme as div_by_cached_rcp in obs.cuh)
VGGP_HD float bin_div(float u, float h, float rh) {
    const float q0 = u * rh;
    const float rem = fmaf(-q0, h, u);
    return fmaf(rem, rh, q0);
}
VGGP_HD double bin_div(double u, double h, double rh) { (void)rh; return u / h; }

template <typename T>
VGGP_HD T bin_ldg(const T* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// Addressing of the per-cell tables, the knots and the gradient buffer (a subset of PackedArgs in obs.cuh).
template <int D>
struct BinGeom {
    int K[D];            // knots per dimension (cells: K - 1)
    int stride[D];       // row-major strides of alpha
    int band_off[D];     // offset of dim d inside the gradient band block: [bp_d | bp_o | bq_d | bq_o] each K long
    int tab_off[D];      // offset of dim d inside the cell tables: [pe0 pe1 pe2 qe0 qe1 qe2 h rh] each K long
    int knot_off[D];     // offset of dim d inside the knot block
};

template <typename T, int D>
struct BinLane {
    int c[D];
    T tlo[D], h[D], rh[D];
    T am[1 << D];                // alpha at the cell corners, monomial basis
    T gm[1 << D];                // sum r prod_{d in S} a_d
    T mom[Pow3<D>::v];           // sum prod_d a_d^{k_d}, index sum_d k_d 3^(D-1-d); entry 0 (the count) is not accumulated
    T e;                         // sum r^2
};

// flat cell id (row-major over cells, dimension 0 slowest) -> per-dimension cell index
template <int D>
VGGP_HD void bin_decode_cell(uint32_t cell, const int (&K)[D], int (&c)[D]) {
#pragma unroll
    for (int d = D - 1; d >= 0; --d) {
        const uint32_t nc = (uint32_t)(K[d] - 1);
        c[d] = (int)(cell % nc);
        cell /= nc;
    }
}

// Enter cell c: knots and knot spacing from the staged tables, alpha corners from L2; clear the sums.
template <typename T, int D>
VGGP_HD void bin_lane_enter(const BinGeom<D>& g, BinLane<T, D>& s, const int (&c)[D], const T* tab,
                            const float* knots, const T* alpha) {
    int base = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const int n = g.K[d];
        const T* tb = tab + (g.tab_off[d] + c[d]);
        s.c[d] = c[d];
        s.tlo[d] = (T)knots[g.knot_off[d] + c[d]];
        s.h[d] = tb[6 * n];          // (T)(float32 knot difference), reference semantics
        s.rh[d] = tb[7 * n];         // correctly rounded 1 / h
        base += c[d] * g.stride[d];
    }
    const T* al = alpha + base;
#pragma unroll
    for (int i = 0; i < (1 << D); ++i) {
        int off = 0;
#pragma unroll
        for (int d = 0; d < D; ++d) off += (i & (1 << (D - 1 - d))) ? g.stride[d] : 0;
        s.am[i] = bin_ldg(al + off);
        s.gm[i] = (T)0;
    }
    // corner values -> monomial coefficients: per dimension (lo, hi) -> (lo, hi - lo)
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const int bit = 1 << (D - 1 - d);
#pragma unroll
        for (int i = 0; i < (1 << D); ++i)
            if (i & bit) s.am[i] -= s.am[i ^ bit];
    }
#pragma unroll
    for (int i = 0; i < Pow3<D>::v; ++i) s.mom[i] = (T)0;
    s.e = (T)0;
}

// One observation of the run.  `live` is false for the padding slots past the run's length (their x is the cell's
// lower knot, so a_d = 0 and only the forced-zero residual needs the flag).
template <typename T, int D>
VGGP_HD void bin_lane_obs(BinLane<T, D>& s, const T (&x)[D], T y, bool live) {
    T w[D], w2[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        w[d] = bin_div(x[d] - s.tlo[d], s.h[d], s.rh[d]);
        w2[d] = w[d] * w[d];
    }
    // monomials prod_{d in S} a_d, S indexed by bits (dimension 0 = most significant bit)
    T mono[1 << D];
    mono[0] = (T)1;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const int bit = 1 << (D - 1 - d);
#pragma unroll
        for (int i = 0; i < (1 << D); ++i)
            if ((i & bit) && !(i & (bit - 1))) mono[i] = (i ^ bit) ? mono[i ^ bit] * w[d] : w[d];
    }
    T mu = s.am[0];
#pragma unroll
    for (int i = 1; i < (1 << D); ++i) mu = bin_fma(s.am[i], mono[i], mu);
    const T r = live ? y - mu : (T)0;
    s.gm[0] += r;
#pragma unroll
    for (int i = 1; i < (1 << D); ++i) s.gm[i] = bin_fma(r, mono[i], s.gm[i]);
    s.e = bin_fma(r, r, s.e);
    // moments: products over the trailing dimensions 1..D-1 are materialised (tp), dimension 0 is fused into the sums
    constexpr int LEN = Pow3<D - 1>::v;
    T tp[LEN];
    tp[0] = (T)1;
    {
        int len = 1;
#pragma unroll
        for (int d = D - 1; d >= 1; --d) {
#pragma unroll
            for (int j = 0; j < LEN; ++j)
                if (j < len) {
                    tp[len + j] = (j == 0) ? w[d] : w[d] * tp[j];
                    tp[2 * len + j] = (j == 0) ? w2[d] : w2[d] * tp[j];
                }
            len *= 3;
        }
    }
#pragma unroll
    for (int j = 1; j < LEN; ++j) s.mom[j] += tp[j];
    s.mom[LEN] += w[0];
    s.mom[2 * LEN] += w2[0];
#pragma unroll
    for (int j = 1; j < LEN; ++j) {
        s.mom[LEN + j] = bin_fma(w[0], tp[j], s.mom[LEN + j]);
        s.mom[2 * LEN + j] = bin_fma(w2[0], tp[j], s.mom[2 * LEN + j]);
    }
}

// Leave the cell: moments -> band sums of P_d, Q_d (monomial -> (diag, off, next diag) form), residual sums -> corner
// form, all added to the gradient buffer through `add` (RED atomics on the device).  Returns the run's contribution to
// sum_n (r^2 - prod p + prod q).
template <typename T, int D, typename Adder>
VGGP_HD T bin_lane_flush(const BinGeom<D>& g, BinLane<T, D>& s, int nrun, const T* tab, T* galpha, T* gband,
                         const Adder& add) {
    // per dimension (m0, m1) -> (m0 - m1, m1)
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const int bit = 1 << (D - 1 - d);
#pragma unroll
        for (int i = 0; i < (1 << D); ++i)
            if (!(i & bit)) s.gm[i] -= s.gm[i | bit];
    }
    int base = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) base += s.c[d] * g.stride[d];
    T* ga = galpha + base;
#pragma unroll
    for (int i = 0; i < (1 << D); ++i) {
        int off = 0;
#pragma unroll
        for (int d = 0; d < D; ++d) off += (i & (1 << (D - 1 - d))) ? g.stride[d] : 0;
        add(ga + off, s.gm[i]);
    }
    T pe[D][3], qe[D][3];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const int n = g.K[d];
        const T* tb = tab + (g.tab_off[d] + s.c[d]);
        pe[d][0] = tb[0]; pe[d][1] = tb[n]; pe[d][2] = tb[2 * n];
        qe[d][0] = tb[3 * n]; qe[d][1] = tb[4 * n]; qe[d][2] = tb[5 * n];
    }
    s.mom[0] = (T)nrun;
    T accV = (T)0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        T bp[3] = {(T)0, (T)0, (T)0}, bq[3] = {(T)0, (T)0, (T)0};
#pragma unroll
        for (int idx = 0; idx < Pow3<D>::v; ++idx) {
            T cp = (T)1, cq = (T)1;
            int kd = 0, rem = idx;
#pragma unroll
            for (int e = D - 1; e >= 0; --e) {
                const int k = rem % 3;
                rem /= 3;
                if (e == d) kd = k;
                else { cp *= pe[e][k]; cq *= qe[e][k]; }
            }
            bp[kd] = bin_fma(cp, s.mom[idx], bp[kd]);
            bq[kd] = bin_fma(cq, s.mom[idx], bq[kd]);
        }
        if (d == 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) accV += qe[0][k] * bq[k] - pe[0][k] * bp[k];
        }
        const int n = g.K[d];
        T* gb = gband + (g.band_off[d] + s.c[d]);
        // sums of w (1-a)^2, w (1-a) a, w a^2 from the moments s0, s1, s2
        add(gb, bp[0] - (T)2 * bp[1] + bp[2]);
        add(gb + n, bp[1] - bp[2]);
        add(gb + 1, bp[2]);
        add(gb + 2 * n, bq[0] - (T)2 * bq[1] + bq[2]);
        add(gb + 3 * n, bq[1] - bq[2]);
        add(gb + 2 * n + 1, bq[2]);
    }
    return s.e + accV;
}

// Where element `e` of a task's data block comes from.  A task stores R / 4 groups; group g holds, for each of the
// D + 1 streamed arrays, 128 values: 4 consecutive observations of each of the 32 lanes (lane-major), so that one
// 16-byte load per lane and array fetches observations 4 g .. 4 g + 3 of its run and a warp reads (D + 1) x 512
// contiguous bytes per group.
struct BinSlot { int arr, lane, j; };
VGGP_HD BinSlot bin_slot_of(int64_t e, int D) {
    const int per_group = (D + 1) * 128;
    const int64_t g = e / per_group;
    const int rem = (int)(e - g * per_group);
    BinSlot s;
    s.arr = rem >> 7;
    s.lane = (rem & 127) >> 2;
    s.j = (int)(g * 4) + (rem & 3);
    return s;
}
VGGP_HD int64_t bin_elem_of(int arr, int lane, int j, int D) {
    return (int64_t)(j >> 2) * ((D + 1) * 128) + arr * 128 + lane * 4 + (j & 3);
}

}  // namespace vggp

// ---------------------------------------------------------------------------------------------------------
// Device side
// ---------------------------------------------------------------------------------------------------------
#if (defined(__CUDACC__) || defined(VGGP_EMUL)) && !defined(VGGP_HOST_EMUL)   // VGGP_EMUL: tests/host_emul SIMT emulator
#include "obs.cuh"

namespace vggp {

struct AtomicAdder {
    template <typename T>
    __device__ __forceinline__ void operator()(T* p, T v) const { atomicAdd(p, v); }
};

template <typename T, int D>
struct BinnedArgs {
    BinGeom<D> geo;
    const unsigned char* buf;        // binned buffer (binplan.hpp: bin_offsets)
    i64 off_task_off, off_task_R, off_run_cell, off_run_n, off_data;
    int n_tasks;
    int knots_byte_off;
    const unsigned char* tables;     // [cell tables (T) | pad16 | knots (float) | pad16] in global memory: touched once
                                     // per run (enter / flush), never in the per-observation loop, so they stay in L2
    const T* alpha;
    T* galpha;
    T* gband;                        // band sums: replica r of the block starts at gband + r * band_rep_stride
    int n_rep;                       // >= 1 replicas of the band block; a CTA adds into replica blockIdx.x % n_rep.  The band
    i64 band_rep_stride;             // block is tiny (4 sum K_d values), so without replicas every flush of every warp hits
                                     // the same few L2 lines and the atomic unit serialises them (ncu, round 2: the kernel
                                     // time grew linearly with the number of runs); k_band_reduce sums the replicas
    double* gs;
    double n_real;
    unsigned int* counter;           // work-stealing counter over tasks (zeroed before the launch)
};

constexpr int BIN_THREADS = 128;
constexpr int BIN_WARPS = BIN_THREADS / 32;
constexpr int BIN_STAGES = 3;        // TMA ring: stages per warp
constexpr int BIN_GPS = 2;           // groups (of 4 observations per lane) per stage

template <typename T, int D>
constexpr int bin_min_blocks() { return sizeof(T) == 4 ? (D == 3 ? 4 : 6) : (D == 1 ? 4 : 2); }

// one group = 4 consecutive observations of this lane's run: (D + 1) 16-byte loads, 128 values apart
template <typename T, int D>
__device__ __forceinline__ void bin_load_group(const T* p, T (&x)[D][4], T (&y)[4]) {
#pragma unroll
    for (int d = 0; d < D; ++d) load4<T>(p + d * 128, x[d]);
    load4<T>(p + D * 128, y);
}

// the same group out of a shared-memory stage filled by a bulk copy (16-byte LDS, conflict-free: lane l reads bytes
// [16 l, 16 l + 16) of every 128-value row)
__device__ __forceinline__ void lds4(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void lds4(const double* p, double (&v)[4]) {
    const double2 a = *reinterpret_cast<const double2*>(p);
    const double2 b = *(reinterpret_cast<const double2*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <typename T, int D>
__device__ __forceinline__ void bin_read_stage(const T* p, T (&x)[D][4], T (&y)[4]) {
#pragma unroll
    for (int d = 0; d < D; ++d) lds4(p + d * 128, x[d]);
    lds4(p + D * 128, y);
}

// `left` = observations of the run not yet consumed (slots past it are padding)
template <typename T, int D>
__device__ __forceinline__ void bin_group(BinLane<T, D>& s, const T (&x)[D][4], const T (&y)[4], int left) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        T xx[D];
#pragma unroll
        for (int d = 0; d < D; ++d) xx[d] = x[d][j];
        bin_lane_obs<T, D>(s, xx, y[j], j < left);
    }
}

// what a warp needs to know about its task
template <typename T, int D>
struct BinTask {
    const T* base;      // first value of this lane's first group
    int groups, nrun;
    bool valid;
    i64 slot;           // 32 task + lane (deterministic mode: where this run's record goes)
    uint32_t cell;      // flat cell id of the run (0 for an empty slot)
};

// cell constants of the task's run (tables and alpha corners from L2).  The LDG kernel calls it AFTER it has issued the loads of
// the run's first group, so that the two round trips overlap instead of queueing behind each other at every task switch.
template <typename T, int D>
__device__ __forceinline__ void bin_task_enter(const BinnedArgs<T, D>& a, BinLane<T, D>& s, const BinTask<T, D>& t) {
    int c[D];
    bin_decode_cell<D>(t.cell, a.geo.K, c);
    bin_lane_enter<T, D>(a.geo, s, c, reinterpret_cast<const T*>(a.tables),
                         reinterpret_cast<const float*>(a.tables + a.knots_byte_off), a.alpha);
}

// `first`: the warp's first task is its own global index -- no atomic.  (All resident warps asking one counter for their first
// task at kernel start serialise on a single L2 address: ~3500 same-address atomics, the ~12 us that separated the kernel from
// the roofline at every problem size.)  Later tasks come from the counter, offset by the number of warps of the launch.
template <typename T, int D, bool ENTER = true>
__device__ __forceinline__ bool bin_next_task(const BinnedArgs<T, D>& a, int lane, BinLane<T, D>& s, BinTask<T, D>& t, bool& first) {
    const unsigned int nwarps = gridDim.x * (blockDim.x >> 5);
    unsigned int task = 0;
    if (first) {
        task = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        first = false;
    } else {
        if (lane == 0) task = atomicAdd(a.counter, 1u) + nwarps;
        task = __shfl_sync(0xffffffffu, task, 0);
    }
    if (task >= (unsigned int)a.n_tasks) return false;
    const i64 slot = (i64)task * 32 + lane;
    t.slot = slot;
    const uint32_t cell = __ldg(reinterpret_cast<const uint32_t*>(a.buf + a.off_run_cell) + slot);
    t.nrun = __ldg(reinterpret_cast<const int*>(a.buf + a.off_run_n) + slot);
    t.groups = __ldg(reinterpret_cast<const int*>(a.buf + a.off_task_R) + task) >> 2;
    t.base = reinterpret_cast<const T*>(a.buf + a.off_data) + __ldg(reinterpret_cast<const i64*>(a.buf + a.off_task_off) + task) + lane * 4;
    t.valid = cell != BIN_EMPTY;
    t.cell = t.valid ? cell : 0u;
    if (ENTER) bin_task_enter<T, D>(a, s, t);
    return true;
}

template <typename T, int D>
__device__ __forceinline__ void bin_finish(const BinnedArgs<T, D>& a, double etot, double* red) {
    const double e = block_sum(etot, red);
    if (threadIdx.x == 0) {
        double extra = 0.0;
        if (blockIdx.x == 0) {
            extra = *reinterpret_cast<const double*>(a.buf);     // sum y^2 of the observations outside the mesh
            a.gs[1] = a.n_real;                                  // single writer; summed over ranks by the all-reduce
        }
        atomicAdd(a.gs + 0, e + extra);
    }
}

// Variant 0: the stream is read with coalesced 16-byte LDG.128 (evict-first) into two register buffers in ping-pong.
template <typename T, int D>
__global__ void __launch_bounds__(BIN_THREADS, (bin_min_blocks<T, D>()))
k_obs_b1_binned(const __grid_constant__ BinnedArgs<T, D> a) {
    __shared__ double red[32];
    const int lane = threadIdx.x & 31;
    const AtomicAdder add;
    double etot = 0.0;
    BinLane<T, D> s;
    BinTask<T, D> t;
    T* const gband = a.gband + (i64)(blockIdx.x % (unsigned)a.n_rep) * a.band_rep_stride;
    // persistent warps: tasks are ordered longest first and handed out from a global counter (LPT scheduling)
    bool first = true;
    while (bin_next_task<T, D, false>(a, lane, s, t, first)) {
        // the loads of the next group of 4 observations are in flight while the current one is processed (no buffer
        // rotation: the loop body handles two groups)
        T xa[D][4], ya[4], xb[D][4], yb[4];
        bin_load_group<T, D>(t.base, xa, ya);
        bin_task_enter<T, D>(a, s, t);
#pragma unroll 1
        for (int gi = 0; gi < t.groups; gi += 2) {
            if (gi + 1 < t.groups) bin_load_group<T, D>(t.base + (i64)(gi + 1) * ((D + 1) * 128), xb, yb);
            bin_group<T, D>(s, xa, ya, t.nrun - 4 * gi);
            if (gi + 2 < t.groups) bin_load_group<T, D>(t.base + (i64)(gi + 2) * ((D + 1) * 128), xa, ya);
            if (gi + 1 < t.groups) bin_group<T, D>(s, xb, yb, t.nrun - 4 * (gi + 1));
        }
        if (t.valid)
            etot += (double)bin_lane_flush<T, D>(a.geo, s, t.nrun, reinterpret_cast<const T*>(a.tables), a.galpha, gband, add);
    }
    bin_finish<T, D>(a, etot, red);
}

// Variant 1: the stream is staged through shared memory by TMA.  Every warp owns a ring of BIN_STAGES stages of
// BIN_GPS groups ((D + 1) x 512 B per group for float32) each; lane 0 issues one cp.async.bulk per stage,
// BIN_STAGES - 1 stages ahead, each completing on the stage's mbarrier; the lanes wait on the barrier's phase and read
// their 16 bytes per array and group with LDS.128 into registers.  The __syncwarp after the last read of a stage hands
// it back to the producer, which refills it at the top of the next iteration.  Nothing but the ring lives in shared
// memory, so the bytes in flight per SM (warps x (BIN_STAGES - 1) x BIN_GPS x 1.5 KB) do not depend on registers.
template <typename T, int D>
__global__ void __launch_bounds__(BIN_THREADS, (bin_min_blocks<T, D>()))
k_obs_b1_binned_tma(const __grid_constant__ BinnedArgs<T, D> a) {
    constexpr int GROUP_ELEMS = (D + 1) * 128;
    constexpr int STAGE_ELEMS = BIN_GPS * GROUP_ELEMS;
    extern __shared__ __align__(128) unsigned char smraw[];
    __shared__ __align__(8) uint64_t full[BIN_WARPS][BIN_STAGES];
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T* ring = reinterpret_cast<T*>(smraw) + (size_t)warp * BIN_STAGES * STAGE_ELEMS;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < BIN_STAGES; ++i) mbar_init(&full[warp][i], 1);
        fence_mbar_init();
    }
    __syncwarp();
    const AtomicAdder add;
    double etot = 0.0;
    BinLane<T, D> s;
    BinTask<T, D> t;
    T* const gband = a.gband + (i64)(blockIdx.x % (unsigned)a.n_rep) * a.band_rep_stride;
    unsigned int k0 = 0;                 // stages this warp has consumed so far: slot = k % STAGES, phase = (k / STAGES) & 1
    bool first = true;
    while (bin_next_task<T, D>(a, lane, s, t, first)) {
        const T* src = t.base - lane * 4;                    // the task's first group
        const int nst = (t.groups + BIN_GPS - 1) / BIN_GPS;  // stages of this task (the last one may be partial)
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < BIN_STAGES - 1; ++i)
                if (i < nst) {
                    const unsigned int st = (k0 + i) % BIN_STAGES;
                    const uint32_t bytes = (uint32_t)(min(BIN_GPS, t.groups - i * BIN_GPS) * GROUP_ELEMS * (int)sizeof(T));
                    mbar_expect_tx(&full[warp][st], bytes);
                    bulk_g2s(ring + (size_t)st * STAGE_ELEMS, src + (i64)i * STAGE_ELEMS, bytes, &full[warp][st]);
                }
        }
#pragma unroll 1
        for (int si = 0; si < nst; ++si) {
            const unsigned int k = k0 + si, st = k % BIN_STAGES;
            const int nx = si + BIN_STAGES - 1;
            if (lane == 0 && nx < nst) {
                // refill the slot consumed in the previous iteration (all lanes passed its __syncwarp)
                const unsigned int sn = (k + BIN_STAGES - 1) % BIN_STAGES;
                const uint32_t bytes = (uint32_t)(min(BIN_GPS, t.groups - nx * BIN_GPS) * GROUP_ELEMS * (int)sizeof(T));
                mbar_expect_tx(&full[warp][sn], bytes);
                bulk_g2s(ring + (size_t)sn * STAGE_ELEMS, src + (i64)nx * STAGE_ELEMS, bytes, &full[warp][sn]);
            }
            mbar_wait(&full[warp][st], (k / BIN_STAGES) & 1u);
            const T* sp = ring + (size_t)st * STAGE_ELEMS + lane * 4;
#pragma unroll
            for (int g = 0; g < BIN_GPS; ++g) {
                const int gi = si * BIN_GPS + g;
                T x[D][4], y[4];
                if (gi < t.groups) bin_read_stage<T, D>(sp + g * GROUP_ELEMS, x, y);
                if (g == BIN_GPS - 1) __syncwarp();
                if (gi < t.groups) bin_group<T, D>(s, x, y, t.nrun - 4 * gi);
            }
        }
        k0 += (unsigned int)nst;
        if (t.valid)
            etot += (double)bin_lane_flush<T, D>(a.geo, s, t.nrun, reinterpret_cast<const T*>(a.tables), a.galpha, gband, add);
    }
    bin_finish<T, D>(a, etot, red);
}

template <typename T, int D>
constexpr size_t bin_tma_smem_bytes() { return (size_t)BIN_WARPS * BIN_STAGES * BIN_GPS * (D + 1) * 128 * sizeof(T); }

// Sum of the band replicas -> the band block of gbuf (overwritten); the replicas and the work-stealing counter are cleared
// for the next launch (both start zeroed at plan creation), so the per-observation call needs no memset of its own for them.
// grid (ceil(n / 16)), 256 threads = 16 elements x 16 replica groups; every thread has all its loads in flight at once.
template <typename T>
__global__ void __launch_bounds__(256) k_band_reduce(T* __restrict__ rep, int n_rep, i64 stride, int n, T* __restrict__ out,
                                                     unsigned int* __restrict__ counter) {
    __shared__ T part[16][17];
    if (blockIdx.x == 0 && threadIdx.x == 0) *counter = 0u;      // the task counter of the kernel that just ran: ready for the next launch
    const int el = threadIdx.x & 15, grp = threadIdx.x >> 4;
    const int e = (int)blockIdx.x * 16 + el;
    T acc = (T)0;
    if (e < n) {
        // all loads of a batch before the first store: as far as the compiler knows the stores may alias the later loads
        for (int r0 = grp; r0 < n_rep; r0 += 16 * 8) {
            T v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = r0 + 16 * u;
                v[u] = (r < n_rep) ? rep[(i64)r * stride + e] : (T)0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = r0 + 16 * u;
                acc += v[u];
                if (r < n_rep) rep[(i64)r * stride + e] = (T)0;
            }
        }
    }
    part[grp][el] = acc;
    __syncthreads();
    if (grp == 0 && e < n) {
        T t = (T)0;
#pragma unroll
        for (int k = 0; k < 16; ++k) t += part[k][el];
        out[e] = t;
    }
}

// ---- deterministic mode (vggp_set_deterministic) ----------------------------------------------------------
// Floating-point atomics make the sums above depend on the order in which warps finish.  In deterministic mode a run does
// not add anything: it writes its flush values as a RECORD (in the fixed order of the `add` calls of bin_lane_flush, then the
// run's scalar term), and the reductions below sum the records in an order that depends only on the binned layout -- runs
// sorted by cell, slot order inside a cell -- so that two launches over the same buffer agree bit for bit whatever
// the grid size, the work stealing or the machine load.
template <int D> struct BinRec { static constexpr int v = (1 << D) + 6 * D + 1; };

template <typename T>
struct RecordAdder {
    T* rec;
    mutable int k;
    __device__ __forceinline__ void operator()(T*, T v) const { rec[k++] = v; }
};

template <typename T, int D>
__global__ void __launch_bounds__(BIN_THREADS, (bin_min_blocks<T, D>()))
k_obs_b1_binned_det(const __grid_constant__ BinnedArgs<T, D> a, T* __restrict__ rec) {
    constexpr int R = BinRec<D>::v;
    const int lane = threadIdx.x & 31;
    BinLane<T, D> s;
    BinTask<T, D> t;
    bool first = true;
    while (bin_next_task<T, D>(a, lane, s, t, first)) {
        T xa[D][4], ya[4], xb[D][4], yb[4];
        bin_load_group<T, D>(t.base, xa, ya);
#pragma unroll 1
        for (int gi = 0; gi < t.groups; gi += 2) {
            if (gi + 1 < t.groups) bin_load_group<T, D>(t.base + (i64)(gi + 1) * ((D + 1) * 128), xb, yb);
            bin_group<T, D>(s, xa, ya, t.nrun - 4 * gi);
            if (gi + 2 < t.groups) bin_load_group<T, D>(t.base + (i64)(gi + 2) * ((D + 1) * 128), xa, ya);
            if (gi + 1 < t.groups) bin_group<T, D>(s, xb, yb, t.nrun - 4 * (gi + 1));
        }
        if (t.valid) {
            T* r = rec + t.slot * R;
            const RecordAdder<T> add{r, 0};
            const T e = bin_lane_flush<T, D>(a.geo, s, t.nrun, reinterpret_cast<const T*>(a.tables), a.galpha, a.gband, add);
            r[R - 1] = e;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) a.gs[1] = a.n_real;
}

// sort keys of the run slots: the cell of the run (a stable sort keeps the slot order inside a cell: an order fixed by the
// layout); empty slots get the key `ncells` and go to the end
__global__ void __launch_bounds__(256) k_det_keys(const uint32_t* __restrict__ run_cell, uint32_t ncells,
                                                  i64 nslots, uint32_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nslots) return;
    const uint32_t c = run_cell[i];
    keys[i] = c != BIN_EMPTY ? c : ncells;
    idx[i] = (uint32_t)i;
}

// [cell_first, cell_end) = positions in the sorted slot list of the runs of a cell (both zero-filled before: no runs)
__global__ void __launch_bounds__(256) k_det_mark(const uint32_t* __restrict__ run_cell, const uint32_t* __restrict__ sorted_slot,
                                                  i64 nslots, uint32_t* __restrict__ cell_first, uint32_t* __restrict__ cell_end) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nslots) return;
    const uint32_t c = run_cell[sorted_slot[i]];
    if (c == BIN_EMPTY) return;
    const uint32_t prev = i > 0 ? run_cell[sorted_slot[i - 1]] : BIN_EMPTY;
    const uint32_t next = i + 1 < nslots ? run_cell[sorted_slot[i + 1]] : BIN_EMPTY;
    if (prev != c) cell_first[c] = (uint32_t)i;
    if (next != c) cell_end[c] = (uint32_t)(i + 1);
}

struct DetIndex {
    const uint32_t* sorted_slot;
    const uint32_t* cell_first;
    const uint32_t* cell_end;
};

// d alpha: one thread per node sums the (up to 2^D) cells it is a corner of, cells in corner order, runs in stream order
template <typename T, int D>
__global__ void __launch_bounds__(256) k_det_alpha(const BinGeom<D> geo, const DetIndex ix, const T* __restrict__ rec,
                                                   T* __restrict__ galpha, i64 M) {
    constexpr int R = BinRec<D>::v;
    const i64 node = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= M) return;
    int idx[D];
    {
        i64 rem = node;
#pragma unroll
        for (int d = 0; d < D; ++d) { idx[d] = (int)(rem / geo.stride[d]); rem -= (i64)idx[d] * geo.stride[d]; }
    }
    T acc = (T)0;
#pragma unroll
    for (int b = 0; b < (1 << D); ++b) {
        bool ok = true;
        i64 flat = 0;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const int c = idx[d] - ((b >> (D - 1 - d)) & 1);
            ok = ok && c >= 0 && c <= geo.K[d] - 2;
            flat = flat * (geo.K[d] - 1) + c;
        }
        if (!ok) continue;
        const uint32_t e = ix.cell_end[flat];
        for (uint32_t i = ix.cell_first[flat]; i < e; ++i) acc += rec[(i64)ix.sorted_slot[i] * R + b];
    }
    galpha[node] = acc;
}

// fixed-shape tree sum over a block of 256 threads of NV values per thread; result in thread 0
template <typename A, int NV>
__device__ __forceinline__ void det_block_sum(A (&v)[NV], A (*sh)[256]) {
#pragma unroll
    for (int j = 0; j < NV; ++j) sh[j][threadIdx.x] = v[j];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
#pragma unroll
            for (int j = 0; j < NV; ++j) sh[j][threadIdx.x] += sh[j][threadIdx.x + o];
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = sh[j][0];
    __syncthreads();
}

// band sums: CTA (d, c) sums the six band values of every run of the hyperplane of cells with index c in dimension d
// (thread t takes the hyperplane's cells t, t + 256, ... in row-major order) -> S[(hoff_d + c) * 6 + j]
template <typename T, int D>
__global__ void __launch_bounds__(256) k_det_band(const BinGeom<D> geo, const DetIndex ix, const T* __restrict__ rec,
                                                  T* __restrict__ S) {
    constexpr int R = BinRec<D>::v;
    __shared__ T sh[6][256];
    int d = 0, c = (int)blockIdx.x, hoff = 0;
    while (d + 1 < D && c >= geo.K[d] - 1) { c -= geo.K[d] - 1; hoff += geo.K[d] - 1; ++d; }
    i64 plane = 1;
#pragma unroll
    for (int e = 0; e < D; ++e)
        if (e != d) plane *= geo.K[e] - 1;
    T v[6] = {(T)0, (T)0, (T)0, (T)0, (T)0, (T)0};
    for (i64 q = threadIdx.x; q < plane; q += 256) {
        // q enumerates the other dimensions row-major; insert c at dimension d
        i64 rem = q, flat = 0, mul = 1;
#pragma unroll
        for (int e = D - 1; e >= 0; --e) {
            int ce;
            if (e == d) ce = c;
            else { ce = (int)(rem % (geo.K[e] - 1)); rem /= geo.K[e] - 1; }
            flat += (i64)ce * mul;
            mul *= geo.K[e] - 1;
        }
        const uint32_t en = ix.cell_end[flat];
        for (uint32_t i = ix.cell_first[flat]; i < en; ++i) {
            const T* r = rec + (i64)ix.sorted_slot[i] * R + (1 << D) + 6 * d;
#pragma unroll
            for (int j = 0; j < 6; ++j) v[j] += r[j];
        }
    }
    det_block_sum<T, 6>(v, sh);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < 6; ++j) S[(i64)(hoff + c) * 6 + j] = v[j];
    }
}

// scalar term: CTA b sums the records of a fixed contiguous range of slots (thread t takes slots t, t + 256, ... of the range;
// fixed-shape tree) -> E[b]; the ranges and the order in which k_det_final adds E[0 .. DET_EBLOCKS) depend on nslots only
constexpr int DET_EBLOCKS = 128;
template <typename T, int D>
__global__ void __launch_bounds__(256) k_det_escal(const T* __restrict__ rec, const uint32_t* __restrict__ run_cell, i64 nslots,
                                                   double* __restrict__ E) {
    constexpr int R = BinRec<D>::v;
    __shared__ double sh[1][256];
    const i64 chunk = (nslots + DET_EBLOCKS - 1) / DET_EBLOCKS;
    const i64 lo = (i64)blockIdx.x * chunk, hi = lo + chunk < nslots ? lo + chunk : nslots;
    double e[1] = {0.0};
    for (i64 i = lo + threadIdx.x; i < hi; i += 256)
        if (run_cell[i] != BIN_EMPTY) e[0] += (double)rec[i * R + R - 1];
    det_block_sum<double, 1>(e, sh);
    if (threadIdx.x == 0) E[blockIdx.x] = e[0];
}

// last step, one CTA: the band block from the hyperplane sums (same destinations as the adds of bin_lane_flush), the scalar
// term from the partial sums of k_det_escal, the sum y^2 of the observations outside the mesh, and the reset of the task counter
template <typename T, int D>
__global__ void __launch_bounds__(256) k_det_final(const BinGeom<D> geo, const T* __restrict__ S, const double* __restrict__ E,
                                                   T* __restrict__ gband, double* __restrict__ gs,
                                                   const unsigned char* __restrict__ buf, unsigned int* __restrict__ counter) {
    int hoff = 0;
    for (int d = 0; d < D; ++d) {
        const int n = geo.K[d], nc = n - 1;
        T* gb = gband + geo.band_off[d];
        for (int i = threadIdx.x; i < n; i += 256) {
            const T* lo = S + (i64)(hoff + i) * 6;          // cell i (i < nc)
            const T* hi = S + (i64)(hoff + i - 1) * 6;      // cell i - 1 (i > 0)
            T pd = (T)0, qd = (T)0;
            if (i < nc) { pd = lo[0]; qd = lo[3]; gb[n + i] = lo[1]; gb[3 * n + i] = lo[4]; }
            if (i > 0) { pd += hi[2]; qd += hi[5]; }
            gb[i] = pd;
            gb[2 * n + i] = qd;
        }
        hoff += nc;
    }
    if (threadIdx.x == 0) {
        double e = 0.0;
        for (int b = 0; b < DET_EBLOCKS; ++b) e += E[b];
        gs[0] = e + *reinterpret_cast<const double*>(buf);
        *counter = 0u;
    }
}

// ---- packing (one-time setup) ---------------------------------------------------------------------------

// observations per flat cell id (keys[i] = ncells for observations outside the mesh)
__global__ void __launch_bounds__(256) k_bin_histogram(const uint32_t* __restrict__ keys, i64 n, uint32_t* __restrict__ count) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        atomicAdd(count + keys[i], 1u);
}

template <typename T, int D>
struct BinGatherArgs {
    const T* x[D];
    const T* y;
    const uint32_t* perm;            // cell-sorted order: stream position -> input index
    const float* knots[D];           // device knot arrays
    int K[D];
    unsigned char* buf;
    i64 off_task_off, off_task_R, off_run_cell, off_run_n, off_run_start, off_data;
    int n_tasks;
};

// one CTA per task (grid-stride): fill the task's data block, coalesced writes, gathered reads
template <typename T, int D>
__global__ void __launch_bounds__(256) k_bin_gather(const __grid_constant__ BinGatherArgs<T, D> a) {
    const i64* task_off = reinterpret_cast<const i64*>(a.buf + a.off_task_off);
    const int* task_R = reinterpret_cast<const int*>(a.buf + a.off_task_R);
    const uint32_t* run_cell = reinterpret_cast<const uint32_t*>(a.buf + a.off_run_cell);
    const int* run_n = reinterpret_cast<const int*>(a.buf + a.off_run_n);
    const uint32_t* run_start = reinterpret_cast<const uint32_t*>(a.buf + a.off_run_start);
    T* data = reinterpret_cast<T*>(a.buf + a.off_data);
    for (int task = blockIdx.x; task < a.n_tasks; task += gridDim.x) {
        const i64 elems = (i64)32 * task_R[task] * (D + 1);
        T* dst = data + task_off[task];
        for (i64 e = threadIdx.x; e < elems; e += blockDim.x) {
            const BinSlot sl = bin_slot_of(e, D);
            const i64 slot = (i64)task * 32 + sl.lane;
            const uint32_t cell = run_cell[slot];
            T v;
            if (sl.j < run_n[slot]) {
                const i64 src = (i64)a.perm[(i64)run_start[slot] + sl.j];
                v = (sl.arr < D) ? a.x[sl.arr < D ? sl.arr : 0][src] : a.y[src];
            } else if (sl.arr < D) {
                int c[D];
                bin_decode_cell<D>(cell != BIN_EMPTY ? cell : 0u, a.K, c);
                v = (T)a.knots[sl.arr][c[sl.arr]];           // padding: weight 0 in every dimension
            } else {
                v = (T)0;
            }
            dst[e] = v;
        }
    }
}

// sum of y^2 over the observations outside the mesh: stream positions [first, n) of the cell-sorted order
template <typename T>
__global__ void __launch_bounds__(256) k_bin_sum_y2(const T* __restrict__ y, const uint32_t* __restrict__ perm, i64 first,
                                                    i64 n, double* __restrict__ out) {
    __shared__ double red[32];
    double acc = 0.0;
    for (i64 i = first + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const double v = (double)y[perm[i]];
        acc += v * v;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(out, acc);
}

}  // namespace vggp
#endif
