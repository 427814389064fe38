// k_allreduce_gbuf: the one collective of a step -- sum of the per-observation gradient buffer over the ranks of one
// NVSwitch node -- as a single kernel over peer memory, launched on the stream of the step between the per-observation
// kernel and the grid-side backward ([d alpha (M) | band sums | 8 float64 scalars], 1 MiB at 512^2).
//
// The buffers are symmetric allocations (same size on every rank, mapped into every process; the host side uses
// torch.distributed._symmetric_memory for the allocation and the handle exchange only).  Two data paths:
//   multicast (NVLS)   rank r reduces slice r of the buffer IN THE SWITCH (multimem.ld_reduce on the multicast address) and
//                      broadcasts it (multimem.st): every element crosses each NVLink once in and once out, two barriers;
//   peer loads         without multicast support: rank r reads slice r from every peer, sums in registers, writes it to every
//                      peer (two-shot all-reduce over plain P2P loads / stores).
// Barriers: one signal word per (block, peer) in a small symmetric pad; a rank publishes a monotone sequence number with
// st.release.sys and polls its own pad with ld.acquire.sys -- no reset, no ABA.  The sequence number lives in DEVICE memory
// (one call counter per block behind the signal words of the local pad, advanced by the kernel itself), so the launch has no
// per-call argument and can be captured into a CUDA graph and replayed: a whole sharded step -- forward, per-observation
// kernel, this collective, backward -- is then one graph launch.  Every poll loop is bounded; a timeout sets an error flag
// instead of hanging the device.
#pragma once
#include "common.cuh"

namespace vggp {

constexpr int AR_MAX_RANKS = 8;
constexpr int AR_BLOCKS = 64;              // most blocks a call uses (the launch picks 8..64 from the buffer size: ar_blocks)
constexpr int AR_THREADS = 512;
constexpr int AR_UNROLL = 4;               // 16-byte units a thread keeps in flight
constexpr int AR_PAD_WORDS = AR_BLOCKS * AR_MAX_RANKS + AR_BLOCKS;
constexpr unsigned int AR_SPIN_LIMIT = 1u << 24;      // ~ 0.1 - 1 s of polling, then give up

struct ArArgs {
    void* mc;                              // multicast address of the buffer, or null
    void* buf[AR_MAX_RANKS];               // unicast address of every rank's buffer in this process (buf[rank] = local)
    unsigned int* pad[AR_MAX_RANKS];       // every rank's signal pad: AR_BLOCKS x AR_MAX_RANKS signal words, then (local
                                           // use only) AR_BLOCKS call counters; AR_PAD_WORDS words, zeroed once at setup
    int rank, world;
    i64 n_obs;                             // values of the observation dtype in the first block
    int obs_f32;
    i64 scal_off;                          // byte offset of the float64 scalars
    int n_scal;
    int* err;                              // device flag: set to 1 on a barrier timeout
};

#ifndef VGGP_EMUL
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 mc_ld_reduce_f32x4(const void* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st_f32x4(void* p, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float mc_ld_reduce_f32(const void* p) {
    float v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st_f32(void* p, float v) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ double mc_ld_reduce_f64(const void* p) {
    double v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st_f64(void* p, double v) {
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// peer loads that must not be served from a stale L1 line: system-scope relaxed loads
__device__ __forceinline__ float4 ld_sys_f32x4(const void* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double2 ld_sys_f64x2(const void* p) {
    double2 v;
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_sys_f32(const void* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_sys_f64(const void* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// all ranks' block b meet: publish `seq` in every peer's pad, wait until every peer has published it in ours
__device__ __forceinline__ void ar_barrier(const ArArgs& a, unsigned int seq) {
    __syncthreads();
    if ((int)threadIdx.x < a.world) {
        const int peer = threadIdx.x;
        __threadfence_system();
        st_release_sys(a.pad[peer] + blockIdx.x * AR_MAX_RANKS + a.rank, seq);
        const unsigned int* mine = a.pad[a.rank] + blockIdx.x * AR_MAX_RANKS + peer;
        unsigned int spins = 0;
        // sequence numbers are compared modulo 2^32 (the counter only ever moves forward)
        while ((int)(ld_acquire_sys(mine) - seq) < 0) {
            if (++spins > AR_SPIN_LIMIT) { *a.err = 1; break; }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(AR_THREADS) k_allreduce_gbuf(const __grid_constant__ ArArgs a) {
    // call number of this block (same on every rank: all ranks make the same sequence of calls); barrier numbers 2 c + 1, 2 c + 2
    unsigned int* counter = a.pad[a.rank] + AR_BLOCKS * AR_MAX_RANKS + blockIdx.x;
    const unsigned int seq = 2u * *counter + 1u;
    ar_barrier(a, seq);                    // every rank's buffer is complete (its producers ran earlier on its stream)
    const int W = a.world, r = a.rank;
    const i64 tid = (i64)blockIdx.x * AR_THREADS + threadIdx.x, nthr = (i64)gridDim.x * AR_THREADS;
    const i64 esz = a.obs_f32 ? 4 : 8;
    // the observation block in units of 16 bytes; rank r owns units [lo, hi).  A thread keeps AR_UNROLL units in flight: the
    // loads of a batch are all issued before the first store (a reduction in the switch is a round trip of a few microseconds)
    const i64 units = a.n_obs * esz / 16;
    const i64 lo = units * r / W, hi = units * (r + 1) / W;
    if (a.mc) {
        unsigned char* mc = reinterpret_cast<unsigned char*>(a.mc);
        if (a.obs_f32) {
            for (i64 u = lo + tid; u < hi; u += AR_UNROLL * nthr) {
                float4 v[AR_UNROLL];
#pragma unroll
                for (int k = 0; k < AR_UNROLL; ++k) if (u + k * nthr < hi) v[k] = mc_ld_reduce_f32x4(mc + 16 * (u + k * nthr));
#pragma unroll
                for (int k = 0; k < AR_UNROLL; ++k) if (u + k * nthr < hi) mc_st_f32x4(mc + 16 * (u + k * nthr), v[k]);
            }
        } else {
            for (i64 u = lo + tid; u < hi; u += AR_UNROLL * nthr) {
                double x[AR_UNROLL], y[AR_UNROLL];
#pragma unroll
                for (int k = 0; k < AR_UNROLL; ++k) if (u + k * nthr < hi) {
                    x[k] = mc_ld_reduce_f64(mc + 16 * (u + k * nthr));
                    y[k] = mc_ld_reduce_f64(mc + 16 * (u + k * nthr) + 8);
                }
#pragma unroll
                for (int k = 0; k < AR_UNROLL; ++k) if (u + k * nthr < hi) {
                    mc_st_f64(mc + 16 * (u + k * nthr), x[k]);
                    mc_st_f64(mc + 16 * (u + k * nthr) + 8, y[k]);
                }
            }
        }
        if (r == 0 && blockIdx.x == 0) {           // tail of the observation block and the float64 scalars
            const i64 done = units * 16 / esz;
            for (i64 e = done + threadIdx.x; e < a.n_obs; e += AR_THREADS) {
                if (a.obs_f32) mc_st_f32(mc + 4 * e, mc_ld_reduce_f32(mc + 4 * e));
                else mc_st_f64(mc + 8 * e, mc_ld_reduce_f64(mc + 8 * e));
            }
            if ((int)threadIdx.x < a.n_scal) {
                unsigned char* q = mc + a.scal_off + 8 * threadIdx.x;
                mc_st_f64(q, mc_ld_reduce_f64(q));
            }
        }
    } else {
        if (a.obs_f32) {
            for (i64 u = lo + tid; u < hi; u += nthr) {
                float4 v[AR_MAX_RANKS];
#pragma unroll
                for (int p = 0; p < AR_MAX_RANKS; ++p)
                    if (p < W) v[p] = ld_sys_f32x4(reinterpret_cast<const unsigned char*>(a.buf[p]) + 16 * u);
                float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int p = 0; p < AR_MAX_RANKS; ++p) if (p < W) { s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w; }
                for (int p = 0; p < W; ++p) *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(a.buf[p]) + 16 * u) = s;
            }
        } else {
            for (i64 u = lo + tid; u < hi; u += nthr) {
                double2 v[AR_MAX_RANKS];
#pragma unroll
                for (int p = 0; p < AR_MAX_RANKS; ++p)
                    if (p < W) v[p] = ld_sys_f64x2(reinterpret_cast<const unsigned char*>(a.buf[p]) + 16 * u);
                double2 s = make_double2(0.0, 0.0);
#pragma unroll
                for (int p = 0; p < AR_MAX_RANKS; ++p) if (p < W) { s.x += v[p].x; s.y += v[p].y; }
                for (int p = 0; p < W; ++p) *reinterpret_cast<double2*>(reinterpret_cast<unsigned char*>(a.buf[p]) + 16 * u) = s;
            }
        }
        if (r == 0 && blockIdx.x == 0) {
            const i64 done = units * 16 / esz;
            for (i64 e = done + threadIdx.x; e < a.n_obs; e += AR_THREADS) {
                if (a.obs_f32) {
                    float s = 0.f;
                    for (int p = 0; p < W; ++p) s += ld_sys_f32(reinterpret_cast<const float*>(a.buf[p]) + e);
                    for (int p = 0; p < W; ++p) *(reinterpret_cast<float*>(a.buf[p]) + e) = s;
                } else {
                    double s = 0.0;
                    for (int p = 0; p < W; ++p) s += ld_sys_f64(reinterpret_cast<const double*>(a.buf[p]) + e);
                    for (int p = 0; p < W; ++p) *(reinterpret_cast<double*>(a.buf[p]) + e) = s;
                }
            }
            if ((int)threadIdx.x < a.n_scal) {
                double s = 0.0;
                for (int p = 0; p < W; ++p)
                    s += ld_sys_f64(reinterpret_cast<const unsigned char*>(a.buf[p]) + a.scal_off + 8 * threadIdx.x);
                for (int p = 0; p < W; ++p)
                    *reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(a.buf[p]) + a.scal_off + 8 * threadIdx.x) = s;
            }
        }
    }
    __threadfence_system();
    ar_barrier(a, seq + 1);                // every slice has been written everywhere
    if (threadIdx.x == 0) *counter = (seq + 1u) >> 1;
}

// blocks of a call: enough threads for AR_UNROLL units each, 8..AR_BLOCKS; a function of the sizes only (identical on all ranks)
inline int ar_blocks(i64 n_obs, int esz, int world) {
    const i64 per_rank = n_obs * esz / 16 / world;
    const i64 b = (per_rank + (i64)AR_THREADS * AR_UNROLL - 1) / ((i64)AR_THREADS * AR_UNROLL);
    return (int)(b < 8 ? 8 : (b > AR_BLOCKS ? AR_BLOCKS : b));
}
#endif  // VGGP_EMUL

}  // namespace vggp
