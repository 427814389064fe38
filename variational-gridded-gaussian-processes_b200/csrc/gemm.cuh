// Grouped / batched / strided float64 GEMM for the grid-side algebra (Kronecker mode-n products, factor
// products, Cholesky trailing updates, triangular-inverse recursion).
//
// One launch = a group of independent problems described by GemmDesc records in device memory; blockIdx.z
// walks (problem, batch entry, k-split).  The inner loop is the FP64 tensor-core instruction
// mma.sync.aligned.m8n8k4.f64 (tcgen05 has no f64 kind); a SIMT inner loop is kept only to cross-check it
// in the tests.
#pragma once
#include "common.cuh"

namespace vggp {

struct GemmDesc {
    const double* A;
    const double* B;
    double* C;
    int m, n, k;
    i64 rsA, csA, rsB, csB, rsC, csC;
    int kinner;          // 0: plain k addressing; else k = ko*kinner + ki with ko using koA/koB
    i64 koA, koB;
    double alpha, beta;
    int batch;
    i64 bsA, bsB, bsC;
    int splitk;          // > 1: partial sums are atomically added into C (C must already hold beta*C)
    int lower_only;      // skip output tiles strictly above the diagonal
    int tri_b;           // 1: B(k,n) = 0 for k < n (lower-triangular B); 2: B(k,n) = 0 for k > n (its transpose)
    int a_mfast, b_kfast;  // which index is contiguous in memory (coalescing of the tile loads)
    int fast;              // eligible for the pipelined cp.async kernel (unit stride along one index of A, B and C rows)
    int a_kcontig, b_ncontig, a_vec2, b_vec2;
    int zstart;          // first blockIdx.z of this problem (filled by gemm_finalize_group)
    int tiles_m, tiles_n;
    // k-split without atomics (automatic splits): every CTA parks its 64 x 64 partial in `ws`, the last one to arrive at a tile
    // (counter `cnt`, left at zero again) adds the partials in split order and stores alpha * sum + beta * C.  No destination
    // clear, no float64 atomics, and the result does not depend on the arrival order.  Null: the atomic form above.
    double* ws;          // [batch][tiles_m][tiles_n][splitk][16][256]
    int* cnt;            // [batch][tiles_m][tiles_n]
};

constexpr int GBM = 64, GBN = 64, GBK = 16;
constexpr int GEMM_THREADS_MMA = 256, GEMM_THREADS_SIMT = 256;
constexpr int GEMM_NI = 2;      // DMMA path: 8 warps as 2 (m) x 4 (n), each 32 x 16 = 4 x 2 tiles of m8n8

#ifndef VGGP_EMUL   // tests/host_emul supplies a CPU stand-in
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
#endif

__device__ __forceinline__ void gemm_store(const GemmDesc& d, double* __restrict__ Cb, int row, int col, double v) {
    if (row >= d.m || col >= d.n) return;
    double* p = Cb + (i64)row * d.rsC + (i64)col * d.csC;
    if (d.splitk > 1) {
        atomicAdd(p, d.alpha * v);
    } else {
        double r = d.alpha * v;
        if (d.beta != 0.0) r += d.beta * (*p);
        *p = r;
    }
}

// Fix-up of a k-split tile (see GemmDesc::ws): v[16] = this thread's accumulators in a fixed enumeration.  Returns false for
// every CTA but the last one of the tile, which leaves with v = sum over the splits in split order.
__device__ __forceinline__ bool gemm_fixup(const GemmDesc& d, int b, int tile_m, int tile_n, int split, double (&v)[16]) {
    __shared__ int last;
    const int tid = threadIdx.x;
    const i64 tile_lin = ((i64)b * d.tiles_m + tile_m) * d.tiles_n + tile_n;
    double* base = d.ws + tile_lin * d.splitk * 4096;
    double* slot = base + (i64)split * 4096;
#pragma unroll
    for (int i = 0; i < 16; ++i) slot[i * 256 + tid] = v[i];
    __threadfence();
    __syncthreads();
    if (tid == 0) last = (atomicAdd(d.cnt + tile_lin, 1) == d.splitk - 1) ? 1 : 0;
    __syncthreads();
    if (!last) return false;
    __threadfence();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.0;
    for (int sp = 0; sp < d.splitk; ++sp) {
        const volatile double* q = base + (i64)sp * 4096;
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += q[i * 256 + tid];
    }
    if (tid == 0) d.cnt[tile_lin] = 0;
    return true;
}
__device__ __forceinline__ void gemm_store_plain(const GemmDesc& d, double* __restrict__ Cb, int row, int col, double v) {
    if (row >= d.m || col >= d.n) return;
    double* p = Cb + (i64)row * d.rsC + (i64)col * d.csC;
    double r = d.alpha * v;
    if (d.beta != 0.0) r += d.beta * (*p);
    *p = r;
}

template <bool MMA>
__device__ __forceinline__ void gemm_body(const GemmDesc& d, int zz, double (*As)[GBK + 4], double (*Bs)[GBN + 4]) {
    const int tid = threadIdx.x;
    constexpr int nthreads = MMA ? GEMM_THREADS_MMA : GEMM_THREADS_SIMT;
    const int tile_m = blockIdx.y, tile_n = blockIdx.x;
    if (tile_m >= d.tiles_m || tile_n >= d.tiles_n) return;
    const int m0 = tile_m * GBM, n0 = tile_n * GBN;
    if (d.lower_only && n0 >= m0 + GBM) return;
    const int split = zz % d.splitk;
    const int b = zz / d.splitk;
    int kbeg = 0, kend = d.k;
    if (d.splitk > 1) {
        const int kchunk = ((d.k + d.splitk - 1) / d.splitk + GBK - 1) / GBK * GBK;
        kbeg = split * kchunk;
        kend = min(d.k, kbeg + kchunk);
        if (kbeg >= kend) { if (!d.ws) return; kbeg = kend; }     // fix-up form: every split of a tile must arrive
    }
    if (d.tri_b == 1) kbeg = max(kbeg, n0 / GBK * GBK);          // triangular operand: skip the all-zero k range
    if (d.tri_b == 2) kend = min(kend, n0 + GBN);
    if (kbeg >= kend && d.splitk > 1) { if (!d.ws) return; kbeg = kend; }
    const double* __restrict__ Ab = d.A + (i64)b * d.bsA;
    const double* __restrict__ Bb = d.B + (i64)b * d.bsB;
    double* __restrict__ Cb = d.C + (i64)b * d.bsC;

    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp >> 2) * 32, wn = (warp & 3) * 16;   // MMA: 8 warps as 2 x 4, 32 x 16 each
    const int ty = tid >> 4, tx = tid & 15;                  // SIMT: 16 x 16 threads, 4 x 4 each

    double acc[4][4][MMA ? 2 : 1];      // MMA path uses [4][GEMM_NI][2]
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < (MMA ? 2 : 1); ++c) acc[i][j][c] = 0.0;

    for (int k0 = kbeg; k0 < kend; k0 += GBK) {
        // ---- stage the A (GBM x GBK) and B (GBK x GBN) tiles (zero-filled at the edges) ----
        for (int e = tid; e < GBM * GBK; e += nthreads) {
            int mm, kk;
            if (d.a_mfast) { mm = e % GBM; kk = e / GBM; } else { kk = e % GBK; mm = e / GBK; }
            const int gm = m0 + mm, gk = k0 + kk;
            double v = 0.0;
            if (gm < d.m && gk < kend) {
                i64 off = (i64)gm * d.rsA;
                off += d.kinner ? (i64)(gk / d.kinner) * d.koA + (i64)(gk % d.kinner) * d.csA : (i64)gk * d.csA;
                v = __ldg(Ab + off);
            }
            As[mm][kk] = v;
        }
        for (int e = tid; e < GBK * GBN; e += nthreads) {
            int nn, kk;
            if (d.b_kfast) { kk = e % GBK; nn = e / GBK; } else { nn = e % GBN; kk = e / GBN; }
            const int gn = n0 + nn, gk = k0 + kk;
            double v = 0.0;
            if (gn < d.n && gk < kend) {
                i64 off = (i64)gn * d.csB;
                off += d.kinner ? (i64)(gk / d.kinner) * d.koB + (i64)(gk % d.kinner) * d.rsB : (i64)gk * d.rsB;
                v = __ldg(Bb + off);
            }
            Bs[kk][nn] = v;
        }
        __syncthreads();
        if constexpr (MMA) {
#pragma unroll
            for (int ks = 0; ks < GBK / 4; ++ks) {
                double a[4], bb[GEMM_NI];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) a[mi] = As[wm + mi * 8 + g][ks * 4 + t];
#pragma unroll
                for (int ni = 0; ni < GEMM_NI; ++ni) bb[ni] = Bs[ks * 4 + t][wn + ni * 8 + g];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < GEMM_NI; ++ni) dmma_m8n8k4(acc[mi][ni][0], acc[mi][ni][1], a[mi], bb[ni]);
            }
        } else {
#pragma unroll
            for (int kk = 0; kk < GBK; ++kk) {
                double a[4], bb[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = As[ty * 4 + i][kk];
#pragma unroll
                for (int j = 0; j < 4; ++j) bb[j] = Bs[kk][tx * 4 + j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j][0] = fma(a[i], bb[j], acc[i][j][0]);
            }
        }
        __syncthreads();
    }

    if (d.ws && d.splitk > 1) {
        double v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if constexpr (MMA) v[i] = acc[i >> 2][(i >> 1) & 1][i & 1];
            else v[i] = acc[i >> 2][i & 3][0];
        }
        if (!gemm_fixup(d, b, tile_m, tile_n, split, v)) return;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MMA) gemm_store_plain(d, Cb, m0 + wm + (i >> 2) * 8 + g, n0 + wn + ((i >> 1) & 1) * 8 + t * 2 + (i & 1), v[i]);
            else gemm_store_plain(d, Cb, m0 + ty * 4 + (i >> 2), n0 + tx * 4 + (i & 3), v[i]);
        }
        return;
    }
    if constexpr (MMA) {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < GEMM_NI; ++ni) {
                const int row = m0 + wm + mi * 8 + g;
                const int col = n0 + wn + ni * 8 + t * 2;
                gemm_store(d, Cb, row, col, acc[mi][ni][0]);
                gemm_store(d, Cb, row, col + 1, acc[mi][ni][1]);
            }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) gemm_store(d, Cb, m0 + ty * 4 + i, n0 + tx * 4 + j, acc[i][j][0]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Pipelined path: 64 x 64 x 16 tiles, GF_NSTAGE cp.async stages (the round-1 kernel had two: with a k-step of 16 the DMMA work
// of one stage is shorter than an L2 round trip, so the pipe idled at every stage switch), 8 warps of 32 x 16 (4 x 2 DMMA
// m8n8k4 tiles each); two warps per scheduler are needed to keep the DMMA pipe busy (ncu: 40 % with one warp per scheduler).
// Operand tiles keep their memory orientation in shared memory (rows along the contiguous index) with row
// pitches chosen so that the DMMA fragment reads are bank-conflict free:
//   k-contiguous operand : [64][20]  (element (mn, k) at mn * 20 + k)
//   mn-contiguous operand: [16][68]  (element (mn, k) at k * 68 + mn)
// ---------------------------------------------------------------------------------------------------------
constexpr int GF_STAGE = 2560;     // doubles per stage: A 1280 + B 1280
constexpr int GF_NSTAGE = 4;       // 80 KB of dynamic shared memory per CTA: two CTAs per SM (three stages / three CTAs measured 8 % slower)
constexpr size_t GEMM_SMEM_BYTES = sizeof(double) * GF_NSTAGE * GF_STAGE;

#ifndef VGGP_EMUL   // tests/host_emul supplies CPU stand-ins
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(src_bytes));
}
// predicate forms on a shared-window address: `ignore` = write zeros, read nothing
typedef uint32_t smaddr_t;
__device__ __forceinline__ smaddr_t sm_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16_p(smaddr_t dst, const void* src, bool ignore) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %2, 0;\n cp.async.cg.shared.global [%0], [%1], 16, p;\n}\n" ::"r"(dst), "l"(src), "r"((int)ignore));
}
__device__ __forceinline__ void cp_async8_p(smaddr_t dst, const void* src, bool ignore) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %2, 0;\n cp.async.ca.shared.global [%0], [%1], 8, p;\n}\n" ::"r"(dst), "l"(src), "r"((int)ignore));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
#endif

// Staging of one operand tile, two 16-byte slots per thread and stage.  `kc`: the k index is the contiguous one.
//   kc  : 64 rows (mn) x 16 (k),  global (mn, k) at base + mn * ld + k,  shared [64][20]
//   !kc : 16 rows (k)  x 64 (mn), global (mn, k) at base + k * ld + mn,  shared [16][68]
// Everything that does not depend on the k-step -- source pointer at kbeg, shared-memory offset, the row / column validity --
// is worked out once per slot before the main loop (the first version recomputed it for every stage: 12 instructions per
// DMMA in the ncu capture of round 2); a stage then costs a clamp, the copy and nothing else.
struct GemmSlot {
    const double* src;      // element at k = kbeg (may point outside the operand when the slot is never valid)
    int dst;                // offset inside the stage's operand tile
    int kpos;               // k offset of the slot inside a stage (kc: column, !kc: row)
    int fixed;              // kc: 1 if the row exists; !kc: number of valid elements along mn (0..2)
};
__device__ __forceinline__ GemmSlot gemm_slot(const double* __restrict__ base, i64 ld, bool kc, int mn0, int kbeg, int mn_lim,
                                              int tid, int it) {
    const int c = tid + it * 256;
    GemmSlot s;
    if (kc) {
        const int row = c >> 3, col = (c & 7) << 1;
        s.dst = row * 20 + col;
        s.kpos = col;
        s.fixed = (mn0 + row < mn_lim) ? 1 : 0;
        s.src = base + (i64)(mn0 + row) * ld + (kbeg + col);
    } else {
        const int row = c >> 5, col = (c & 31) << 1;
        s.dst = row * 68 + col;
        s.kpos = row;
        int v = mn_lim - (mn0 + col);
        s.fixed = v < 0 ? 0 : (v > 2 ? 2 : v);
        s.src = base + (i64)(kbeg + row) * ld + (mn0 + col);
    }
    return s;
}
// copy the slot for the stage that starts at k = kbeg + koff (koff a multiple of GBK); `kstep` = ld for !kc, 1 for kc
__device__ __forceinline__ void gemm_stage_slot(double* sm, const GemmSlot& s, bool kc, bool vec2, i64 kstep, int koff, int krem,
                                                const double* __restrict__ base) {
    // krem = k_lim - (kbeg + koff): k values left from the start of this stage
    int valid;
    if (kc) { valid = s.fixed ? krem - s.kpos : 0; valid = valid < 0 ? 0 : (valid > 2 ? 2 : valid); }
    else valid = (s.kpos < krem) ? s.fixed : 0;
    const double* src = valid ? s.src + (i64)koff * kstep : base;
    double* dst = sm + s.dst;
    if (vec2) {
        cp_async16(dst, src, valid * 8);
    } else {
        cp_async8(dst, src, valid > 0 ? 8 : 0);
        cp_async8(dst + 1, valid > 1 ? src + 1 : base, valid > 1 ? 8 : 0);
    }
}

// The same slot for a stage that lies entirely inside the k range ("lean" form).  Rows past the m / n edge are CLAMPED to the
// last row instead of zero-filled: they only feed output rows / columns that are never stored, so no predicate and no byte
// count is left -- per copy one address (32-bit element offset from the operand base, advanced by a uniform step per
// stage) and the LDGSTS.  Only the k edge needs zeros; the one partial stage of a k range keeps the general form.
struct GemmLean {
    int o0, o1;             // element offsets of the copy (16 bytes at o0 when aligned, else 8 bytes at o0 and at o1)
    smaddr_t dst;           // destination inside stage 0
};
__device__ __forceinline__ GemmLean gemm_lean(i64 ld, bool kc, bool vec2, int mn0, int kbeg, int mn_lim, int tid, int it,
                                              double* sm_operand) {
    const int c = tid + it * 256;
    GemmLean l;
    if (kc) {
        const int row = c >> 3, col = (c & 7) << 1;
        const int rr = min(mn0 + row, mn_lim - 1);
        l.o0 = (int)((i64)rr * ld) + kbeg + col;
        l.o1 = l.o0 + 1;
        l.dst = sm_addr(sm_operand + row * 20 + col);
    } else {
        const int row = c >> 5, col = (c & 31) << 1;
        const int rowoff = (int)((i64)(kbeg + row) * ld);
        // aligned pairs stay aligned: with 16-byte copies the edge is even (gemm_lean_ok), so the last pair starts at lim - 2
        l.o0 = rowoff + min(mn0 + col, vec2 ? mn_lim - 2 : mn_lim - 1);
        l.o1 = rowoff + min(mn0 + col + 1, mn_lim - 1);
        l.dst = sm_addr(sm_operand + row * 68 + col);
    }
    return l;
}
// is the lean form usable for this operand tile?  (offsets fit 32 bits; an aligned 16-byte copy never straddles the edge)
__device__ __forceinline__ bool gemm_lean_ok(i64 ld, bool kc, bool vec2, int mn0, int mn_lim, int kend) {
    const i64 span = kc ? (i64)mn_lim * ld + kend : (i64)(kend + GBK) * ld + mn_lim;
    if (span >= ((i64)1 << 31) || mn_lim < 1) return false;
    if (vec2 && !kc && mn_lim < mn0 + 64 && (mn_lim & 1)) return false;
    if (vec2 && !kc && (mn_lim <= mn0 || mn_lim < 2)) return false;
    return true;
}
__device__ __forceinline__ void gemm_stage_lean(const double* __restrict__ base, int stage_bytes, GemmLean& l, bool vec2, int inc) {
    const smaddr_t dst = l.dst + stage_bytes;
    if (vec2) {
        cp_async16_p(dst, base + l.o0, false);
    } else {
        cp_async8_p(dst, base + l.o0, false);
        cp_async8_p(dst + 8, base + l.o1, false);
    }
    l.o0 += inc;
    l.o1 += inc;
}

__device__ __forceinline__ void gemm_fast_body(const GemmDesc& d, int zz, double* sm) {
    const int tid = threadIdx.x;
    const int tile_m = blockIdx.y, tile_n = blockIdx.x;
    if (tile_m >= d.tiles_m || tile_n >= d.tiles_n) return;
    const int m0 = tile_m * GBM, n0 = tile_n * GBN;
    if (d.lower_only && n0 >= m0 + GBM) return;
    const int split = zz % d.splitk;
    const int b = zz / d.splitk;
    int kbeg = 0, kend = d.k;
    if (d.splitk > 1) {
        const int kchunk = ((d.k + d.splitk - 1) / d.splitk + GBK - 1) / GBK * GBK;
        kbeg = split * kchunk;
        kend = min(d.k, kbeg + kchunk);
        if (kbeg >= kend) { if (!d.ws) return; kbeg = kend; }     // fix-up form: every split of a tile must arrive
    }
    if (d.tri_b == 1) kbeg = max(kbeg, n0 / GBK * GBK);          // triangular operand: skip the all-zero k range
    if (d.tri_b == 2) kend = min(kend, n0 + GBN);
    if (kbeg >= kend && d.splitk > 1) { if (!d.ws) return; kbeg = kend; }
    const double* __restrict__ Ab = d.A + (i64)b * d.bsA;
    const double* __restrict__ Bb = d.B + (i64)b * d.bsB;
    double* __restrict__ Cb = d.C + (i64)b * d.bsC;
    const bool akc = d.a_kcontig != 0, bnc = d.b_ncontig != 0;
    const i64 lda = akc ? d.rsA : d.csA;
    const i64 ldb = bnc ? d.rsB : d.csB;
    const int sAm = akc ? 20 : 1, sAk = akc ? 1 : 68;
    const int sBk = bnc ? 68 : 1, sBn = bnc ? 1 : 20;

    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp >> 2) * 32, wn = (warp & 3) * 16;
    const int aoff = (wm + g) * sAm + t * sAk;
    const int boff = t * sBk + (wn + g) * sBn;

    double acc[4][GEMM_NI][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < GEMM_NI; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    const int nk = (kend - kbeg + GBK - 1) / GBK;
    const bool av2 = d.a_vec2 != 0, bv2 = d.b_vec2 != 0;
    const GemmSlot sa0 = gemm_slot(Ab, lda, akc, m0, kbeg, d.m, tid, 0), sa1 = gemm_slot(Ab, lda, akc, m0, kbeg, d.m, tid, 1);
    const GemmSlot sb0 = gemm_slot(Bb, ldb, !bnc, n0, kbeg, d.n, tid, 0), sb1 = gemm_slot(Bb, ldb, !bnc, n0, kbeg, d.n, tid, 1);
    const i64 ksa = akc ? 1 : lda, ksb = !bnc ? 1 : ldb;
    // Full stages (all but possibly the last one) go through the lean form (SASS of round 2: 250 of the 330 instructions of a
    // k-step were staging, issued by every warp right after the barrier while the tensor pipe idled).  Stages are issued in
    // order, so the offsets just advance.
    const bool lean = gemm_lean_ok(lda, akc, av2, m0, d.m, kend) && gemm_lean_ok(ldb, !bnc, bv2, n0, d.n, kend);
    const int nfull = lean ? (kend - kbeg) / GBK : 0;       // stages that lie entirely inside the k range
    GemmLean la0 = gemm_lean(lda, akc, av2, m0, kbeg, d.m, tid, 0, sm), la1 = gemm_lean(lda, akc, av2, m0, kbeg, d.m, tid, 1, sm);
    GemmLean lb0 = gemm_lean(ldb, !bnc, bv2, n0, kbeg, d.n, tid, 0, sm + 1280), lb1 = gemm_lean(ldb, !bnc, bv2, n0, kbeg, d.n, tid, 1, sm + 1280);
    const int inca = (int)(ksa * GBK), incb = (int)(ksb * GBK);
    auto stage = [&](int st) {
        double* nx = sm + (st % GF_NSTAGE) * GF_STAGE;
        if (st < nfull) {
            const int sb = (st % GF_NSTAGE) * (GF_STAGE * (int)sizeof(double));
            gemm_stage_lean(Ab, sb, la0, av2, inca);
            gemm_stage_lean(Ab, sb, la1, av2, inca);
            gemm_stage_lean(Bb, sb, lb0, bv2, incb);
            gemm_stage_lean(Bb, sb, lb1, bv2, incb);
        } else {
            const int koff = st * GBK, krem = kend - kbeg - koff;
            gemm_stage_slot(nx, sa0, akc, av2, ksa, koff, krem, Ab);
            gemm_stage_slot(nx, sa1, akc, av2, ksa, koff, krem, Ab);
            gemm_stage_slot(nx + 1280, sb0, !bnc, bv2, ksb, koff, krem, Bb);
            gemm_stage_slot(nx + 1280, sb1, !bnc, bv2, ksb, koff, krem, Bb);
        }
    };
    // prologue: GF_NSTAGE - 1 stages in flight (one commit group per stage, empty groups past the end keep the count uniform)
#pragma unroll 1
    for (int s = 0; s < GF_NSTAGE - 1; ++s) {
        if (s < nk) stage(s);
        cp_async_commit();
    }
    for (int it = 0; it < nk; ++it) {
        cp_async_wait<GF_NSTAGE - 2>();          // stage `it` has landed (for this thread's copies) ...
        __syncthreads();                         // ... and for everybody's; everybody has also left stage it - 1
        if (it + GF_NSTAGE - 1 < nk) stage(it + GF_NSTAGE - 1);      // refill the slot that stage it - 1 occupied
        cp_async_commit();
        const double* As = sm + (it % GF_NSTAGE) * GF_STAGE;
        const double* Bs = As + 1280;
        // fragments of k-step ks + 1 are loaded before the DMMAs of k-step ks are issued (register double buffer): the
        // shared-memory latency sits under the tensor instructions instead of in front of them
        double a[2][4], bb[2][GEMM_NI];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) a[0][mi] = As[aoff + mi * 8 * sAm];
#pragma unroll
        for (int ni = 0; ni < GEMM_NI; ++ni) bb[0][ni] = Bs[boff + ni * 8 * sBn];
#pragma unroll
        for (int ks = 0; ks < GBK / 4; ++ks) {
            const int cur = ks & 1, nxt = cur ^ 1;
            if (ks + 1 < GBK / 4) {
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) a[nxt][mi] = As[aoff + mi * 8 * sAm + (ks + 1) * 4 * sAk];
#pragma unroll
                for (int ni = 0; ni < GEMM_NI; ++ni) bb[nxt][ni] = Bs[boff + (ks + 1) * 4 * sBk + ni * 8 * sBn];
            }
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < GEMM_NI; ++ni) dmma_m8n8k4(acc[mi][ni][0], acc[mi][ni][1], a[cur][mi], bb[cur][ni]);
        }
    }
    cp_async_wait<0>();
    if (d.ws && d.splitk > 1) {
        double v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = acc[i >> 2][(i >> 1) & 1][i & 1];
        if (!gemm_fixup(d, b, tile_m, tile_n, split, v)) return;
#pragma unroll
        for (int i = 0; i < 16; ++i)
            gemm_store_plain(d, Cb, m0 + wm + (i >> 2) * 8 + g, n0 + wn + ((i >> 1) & 1) * 8 + t * 2 + (i & 1), v[i]);
        return;
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < GEMM_NI; ++ni) {
            const int row = m0 + wm + mi * 8 + g;
            const int col = n0 + wn + ni * 8 + t * 2;
            gemm_store(d, Cb, row, col, acc[mi][ni][0]);
            gemm_store(d, Cb, row, col + 1, acc[mi][ni][1]);
        }
}

// Group launch: descriptors live in device memory (built once at plan creation).
template <bool MMA>
__global__ void __launch_bounds__(MMA ? GEMM_THREADS_MMA : GEMM_THREADS_SIMT)
k_gemm_group(const GemmDesc* __restrict__ descs, int ndesc) {
    extern __shared__ double sm[];          // GEMM_SMEM_BYTES
    double* smem = sm;
    __shared__ GemmDesc sd;
    const int z = blockIdx.z;
    int p = 0;
    for (int i = 1; i < ndesc; ++i)
        if (z >= descs[i].zstart) p = i;
    if (threadIdx.x == 0) sd = descs[p];
    __syncthreads();
    if (MMA && sd.fast) {
        gemm_fast_body(sd, z - sd.zstart, smem);
    } else {
        double (*As)[GBK + 4] = reinterpret_cast<double (*)[GBK + 4]>(smem);
        double (*Bs)[GBN + 4] = reinterpret_cast<double (*)[GBN + 4]>(smem + GBM * (GBK + 4));
        gemm_body<MMA>(sd, z - sd.zstart, As, Bs);
    }
}

// Single ad-hoc problem passed by value (tests, vggp_gemm_f64, vggp_mode_product).
template <bool MMA>
__global__ void __launch_bounds__(MMA ? GEMM_THREADS_MMA : GEMM_THREADS_SIMT)
k_gemm_one(const __grid_constant__ GemmDesc d) {
    extern __shared__ double sm[];          // GEMM_SMEM_BYTES
    double* smem = sm;
    if (MMA && d.fast) {
        gemm_fast_body(d, blockIdx.z, smem);
    } else {
        double (*As)[GBK + 4] = reinterpret_cast<double (*)[GBK + 4]>(smem);
        double (*Bs)[GBN + 4] = reinterpret_cast<double (*)[GBN + 4]>(smem + GBM * (GBK + 4));
        gemm_body<MMA>(d, blockIdx.z, As, Bs);
    }
}

// ---- host side -------------------------------------------------------------------------------------------
struct GemmGroupDims {
    int gx, gy, gz;
};

inline void gemm_desc_defaults(GemmDesc& d) {
    memset(&d, 0, sizeof(d));
    d.alpha = 1.0;
    d.beta = 0.0;
    d.batch = 1;
    d.splitk = 1;
    d.ws = nullptr;
    d.cnt = nullptr;
}

// Fill the derived fields of a group of descriptors (zstart, tile counts, coalescing hints).
inline GemmGroupDims gemm_finalize_group(GemmDesc* descs, int ndesc) {
    GemmGroupDims g = {0, 0, 0};
    int z = 0;
    for (int i = 0; i < ndesc; ++i) {
        GemmDesc& d = descs[i];
        d.tiles_m = (d.m + GBM - 1) / GBM;
        d.tiles_n = (d.n + GBN - 1) / GBN;
        d.a_mfast = (d.rsA == 1 && d.csA != 1) ? 1 : 0;
        d.b_kfast = (d.rsB == 1 && d.csB != 1) ? 1 : 0;
        if (d.splitk < 1) d.splitk = 1;
        if (d.batch < 1) d.batch = 1;
        // pipelined kernel: A and B have unit stride along one index, plain k addressing, k >= one tile
        const bool a_unit = (d.csA == 1) || (d.rsA == 1), b_unit = (d.csB == 1) || (d.rsB == 1);
        d.a_kcontig = (d.csA == 1) ? 1 : 0;
        d.b_ncontig = (d.csB == 1) ? 1 : 0;
        d.fast = (a_unit && b_unit && d.kinner == 0 && d.k >= 1) ? 1 : 0;
        const i64 lda = d.a_kcontig ? d.rsA : d.csA, ldb = d.b_ncontig ? d.rsB : d.csB;
        d.a_vec2 = (lda % 2 == 0 && d.bsA % 2 == 0 && ((uintptr_t)d.A % 16) == 0) ? 1 : 0;
        d.b_vec2 = (ldb % 2 == 0 && d.bsB % 2 == 0 && ((uintptr_t)d.B % 16) == 0) ? 1 : 0;
        d.zstart = z;
        z += d.batch * d.splitk;
        if (d.tiles_n > g.gx) g.gx = d.tiles_n;
        if (d.tiles_m > g.gy) g.gy = d.tiles_m;
    }
    g.gz = z;
    return g;
}

}  // namespace vggp
