// TEST INFRASTRUCTURE ONLY -- a minimal SIMT emulator that lets g++ compile and run the library's CUDA kernels on the
// CPU (found as <cuda_runtime.h> through -I when tests/host_emul is built; nvcc never sees this file).
//
// One kernel launch = blocks run one after the other; the threads of a block are ucontext fibres on ONE OS thread,
// switched cooperatively at every synchronising primitive (__syncthreads, __syncwarp, warp shuffles, mbarrier waits,
// atomics).  `__shared__` is `thread_local` (= one copy, shared by all fibres; dynamic shared memory is a
// thread_local array defined in emul_device.cpp).  Bulk async copies (TMA) are queued and completed LATE, at random
// scheduling points, so that reading a stage before its barrier completed, or refilling it before it was consumed,
// corrupts the result.  This checks indexing, protocols and arithmetic of device code; it says nothing about memory
// ordering or performance.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <algorithm>
#include <functional>
#include <vector>

#define VGGP_EMUL 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ thread_local
#define __grid_constant__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

using std::min;
using std::max;

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }

// ---- host runtime API: "device" memory is host memory, streams are synchronous ----
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
struct cudaDeviceProp { int multiProcessorCount; };
template <typename P>
inline cudaError_t cudaMalloc(P** p, size_t bytes) {
    void* q = nullptr;
    if (posix_memalign(&q, 256, bytes ? bytes : 256)) return 2;
    memset(q, 0xCD, bytes);                      // poison: reading memory the library never wrote shows up
    *p = reinterpret_cast<P*>(q);
    return cudaSuccess;
}
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemset2DAsync(void* d, size_t pitch, int v, size_t width, size_t height, cudaStream_t = nullptr) {
    for (size_t r = 0; r < height; ++r) memset(static_cast<unsigned char*>(d) + r * pitch, v, width);
    return cudaSuccess;
}
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
// events and extra streams: no time passes on the emulator, ordering is program order
typedef void* cudaEvent_t;
enum { cudaEventDisableTiming = 2, cudaStreamNonBlocking = 1 };
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
constexpr unsigned cudaEventRecordExternal = 1;
inline cudaError_t cudaEventRecordWithFlags(cudaEvent_t, cudaStream_t = nullptr, unsigned = 0) { return cudaSuccess; }
enum cudaStreamCaptureStatus { cudaStreamCaptureStatusNone = 0, cudaStreamCaptureStatusActive = 1, cudaStreamCaptureStatusInvalidated = 2 };
inline cudaError_t cudaStreamIsCapturing(cudaStream_t, cudaStreamCaptureStatus* s) { *s = cudaStreamCaptureStatusNone; return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->multiProcessorCount = 3; return cudaSuccess; }
template <typename F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int bytes) { return bytes <= 227 * 1024 ? cudaSuccess : 1; }
template <typename F>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 2; return cudaSuccess; }

struct uint3 { unsigned int x, y, z; };
struct dim3 {
    unsigned int x, y, z;
    dim3(unsigned int a = 1, unsigned int b = 1, unsigned int c = 1) : x(a), y(b), z(c) {}
};
struct float4 { float x, y, z, w; } __attribute__((aligned(16)));
struct double2 { double x, y; } __attribute__((aligned(16)));

namespace cuda_emul {
extern uint3 threadIdx_, blockIdx_;
extern dim3 blockDim_, gridDim_;
void yield();                                       // give the other fibres of the block a turn
void sync_block();                                  // __syncthreads
void sync_warp();                                   // all live lanes of the calling warp
uint64_t* warp_slots();                             // 32 exchange slots of the calling warp
int lane_id();
void launch(dim3 grid, dim3 block, const std::function<void()>& kernel_body);
void bulk_copy_async(void* dst, const void* src, uint32_t bytes, uint64_t* bar);
void cp_async_enqueue(void* dst, const void* src, int bytes, int src_bytes);   // per-thread cp.async queue
void cp_async_commit_group();
void cp_async_wait_group(int newest_groups_left_pending);
unsigned long long yields();
}  // namespace cuda_emul

#define threadIdx (cuda_emul::threadIdx_)
#define blockIdx (cuda_emul::blockIdx_)
#define blockDim (cuda_emul::blockDim_)
#define gridDim (cuda_emul::gridDim_)

inline void __syncthreads() { cuda_emul::sync_block(); }
inline void __syncwarp(unsigned int = 0xffffffffu) { cuda_emul::sync_warp(); }
inline void __threadfence() { cuda_emul::yield(); }      // memory is sequentially consistent here; a fence is a scheduling point

template <typename T>
inline T __shfl_sync(unsigned int, T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle of <= 8 bytes");
    uint64_t* s = cuda_emul::warp_slots();
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    s[cuda_emul::lane_id()] = raw;
    cuda_emul::sync_warp();
    raw = s[src & 31];
    cuda_emul::sync_warp();
    T out;
    memcpy(&out, &raw, sizeof(T));
    return out;
}
template <typename T>
inline T __shfl_xor_sync(unsigned int m, T v, int mask) { return __shfl_sync(m, v, cuda_emul::lane_id() ^ mask); }

// atomics: fibres are cooperative, so a plain read-modify-write is atomic; yield afterwards to shuffle the order
template <typename T>
inline T atomicAdd(T* p, T v) { const T o = *p; *p = o + v; cuda_emul::yield(); return o; }
template <typename T>
inline T atomicMax(T* p, T v) { const T o = *p; if (v > o) *p = v; cuda_emul::yield(); return o; }

template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __ldcs(const T* p) { return *p; }
// round-to-nearest arithmetic intrinsics: plain operators (the emulator is compiled with -ffp-contract=off)
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline unsigned int __umulhi(unsigned int a, unsigned int b) { return (unsigned int)(((uint64_t)a * (uint64_t)b) >> 32); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __double2hiint(double v) { int64_t b; memcpy(&b, &v, 8); return (int)(b >> 32); }
inline double __hiloint2double(int hi, int lo) { const int64_t b = ((int64_t)hi << 32) | (uint32_t)lo; double v; memcpy(&v, &b, 8); return v; }
inline int __float_as_int(float v) { int i; memcpy(&i, &v, 4); return i; }
inline long long __double_as_longlong(double v) { long long i; memcpy(&i, &v, 8); return i; }
inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
inline double __longlong_as_double(long long v) { double f; memcpy(&f, &v, 8); return f; }

// CPU stand-ins for the PTX wrappers of csrc/obs.cuh (mbarrier + bulk async copy).  The barrier word keeps the parity
// of the phase in progress in bit 0 and the transaction bytes announced by expect_tx above bit 8; a queued copy
// completes (memcpy + parity flip) at a random later scheduling point and must match the announced byte count.
namespace vggp {
inline void mbar_init(uint64_t* bar, int) { *bar = 0; }
inline void fence_mbar_init() {}
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    if (*bar >> 8) { fprintf(stderr, "cuda_emul: mbarrier re-armed while a copy is pending\n"); abort(); }
    *bar = (*bar & 1ull) | ((uint64_t)bytes << 8);
}
inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    if ((*bar >> 8) != bytes || (bytes & 15u) || ((uintptr_t)dst & 15u) || ((uintptr_t)src & 15u)) {
        fprintf(stderr, "cuda_emul: bulk copy of %u bytes: size / alignment / expect_tx mismatch\n", bytes);
        abort();
    }
    cuda_emul::bulk_copy_async(dst, src, bytes, bar);
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
    while ((*bar & 1ull) == (uint64_t)parity) cuda_emul::yield();
}
}  // namespace vggp

// CPU stand-ins for the PTX wrappers of csrc/gemm.cuh.
namespace vggp {
// mma.sync.aligned.m8n8k4.row.col.f64: lane l holds A[l/4][l%4], B[l%4][l/4], C[l/4][2 (l%4) + {0,1}]
inline void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    uint64_t* s = cuda_emul::warp_slots();
    const int lane = cuda_emul::lane_id();
    double A[32], B[32];
    memcpy(&s[lane], &a, 8);
    cuda_emul::sync_warp();
    memcpy(A, s, sizeof(A));
    cuda_emul::sync_warp();
    memcpy(&s[lane], &b, 8);
    cuda_emul::sync_warp();
    memcpy(B, s, sizeof(B));
    cuda_emul::sync_warp();
    const int row = lane >> 2, col = (lane & 3) * 2;
    for (int k = 0; k < 4; ++k) {
        c0 = fma(A[row * 4 + k], B[col * 4 + k], c0);
        c1 = fma(A[row * 4 + k], B[(col + 1) * 4 + k], c1);
    }
}
// cp.async: the copy is DEFERRED until the group it belongs to is waited for (a missing wait reads stale data)
inline void cp_async16(void* dst, const void* src, int src_bytes) { cuda_emul::cp_async_enqueue(dst, src, 16, src_bytes); }
inline void cp_async8(void* dst, const void* src, int src_bytes) { cuda_emul::cp_async_enqueue(dst, src, 8, src_bytes); }
typedef char* smaddr_t;
inline smaddr_t sm_addr(const void* p) { return (char*)const_cast<void*>(p); }
inline void cp_async16_p(smaddr_t dst, const void* src, bool ignore) { cuda_emul::cp_async_enqueue(dst, src, 16, ignore ? 0 : 16); }
inline void cp_async8_p(smaddr_t dst, const void* src, bool ignore) { cuda_emul::cp_async_enqueue(dst, src, 8, ignore ? 0 : 8); }
inline void cp_async_commit() { cuda_emul::cp_async_commit_group(); }
template <int N>
inline void cp_async_wait() { cuda_emul::cp_async_wait_group(N); }
}  // namespace vggp
