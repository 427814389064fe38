// Diagnostic (not product): instruction latencies / issue rates that the grid-side cost model depends on, measured on the
// device at hand.  nvcc -arch=sm_100a -O3 -o tools/microbench/lat tools/microbench/lat.cu ; run under gpurun.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma_dep(double* out, double a, double b, int n, long long* cyc) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = fma(x, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ffma_dep(float* out, float a, float b, int n, long long* cyc) {
    float x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = fmaf(x, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP>
__global__ void k_dfma_ilp(double* out, double a, double b, int n, long long* cyc) {
    double x[ILP];
    for (int k = 0; k < ILP; ++k) x[k] = out[threadIdx.x + k];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds_chase(int* out, int n, long long* cyc) {
    __shared__ int nxt[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) nxt[i] = (i * 17 + 5) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) p = nxt[p];
    long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ldg_chase(const int* __restrict__ nxt, int* out, int n, long long* cyc) {
    int p = threadIdx.x * 32;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) p = nxt[p];
    long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_fill(int* nxt, int n, int stride) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) nxt[i] = (int)(((long long)i + stride) % n);
}
int main() {
    double* d; float* f; int* io; long long* cyc; int* nxt;
    cudaMalloc(&d, 8 * 1024); cudaMemset(d, 0, 8 * 1024);
    cudaMalloc(&f, 4 * 1024); cudaMemset(f, 0, 4 * 1024);
    cudaMalloc(&io, 4 * 1024); cudaMalloc(&cyc, 64);
    const int NW = 1 << 24;          // 64 MB of ints: L2 resident after the first sweep
    cudaMalloc(&nxt, 4ull * NW);
    long long h;
    const int n = 4096;
    for (int rep = 0; rep < 2; ++rep) {
        k_dfma_dep<<<1, 32>>>(d, 1.0000001, 1e-9, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("DFMA dependent chain, 1 warp            : %.1f cycles / op\n", (double)h / n);
        k_ffma_dep<<<1, 32>>>(f, 1.0000001f, 1e-9f, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("FFMA dependent chain, 1 warp            : %.1f cycles / op\n", (double)h / n);
        k_dfma_ilp<8><<<1, 32>>>(d, 1.0000001, 1e-9, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("DFMA 8 independent chains, 1 warp       : %.2f cycles / op (issue interval)\n", (double)h / n / 8);
        k_dfma_ilp<8><<<1, 128>>>(d, 1.0000001, 1e-9, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("DFMA 8 chains x 4 warps (1 / scheduler)  : %.2f cycles / op per warp\n", (double)h / n / 8);
        k_dfma_ilp<8><<<1, 512>>>(d, 1.0000001, 1e-9, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("DFMA 8 chains x 16 warps (4 / scheduler) : %.2f cycles / op per warp -> %.1f DFMA lanes / clk / SM\n", (double)h / n / 8, 16.0 * 32 / ((double)h / n / 8));
        k_lds_chase<<<1, 32>>>(io, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("LDS dependent chase                      : %.1f cycles / load\n", (double)h / n);
        k_fill<<<1024, 256>>>(nxt, NW, 4099 * 32);
        k_ldg_chase<<<1, 32>>>(nxt, io, 1024, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        k_ldg_chase<<<1, 32>>>(nxt, io, 1024, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("LDG dependent chase (L2 resident, 64 MB) : %.0f cycles / load\n", (double)h / 1024);
    }
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
