// TEST INFRASTRUCTURE ONLY -- stand-in for cub::DeviceRadixSort::SortPairs under the SIMT emulator (stable sort).
#pragma once
#include <cuda_runtime.h>
#include <numeric>
namespace cub {
struct DeviceRadixSort {
    template <typename K, typename V>
    static cudaError_t SortPairs(void* temp, size_t& temp_bytes, const K* keys_in, K* keys_out, const V* vals_in, V* vals_out,
                                 int n, int begin_bit = 0, int end_bit = sizeof(K) * 8, cudaStream_t = nullptr) {
        if (!temp) { temp_bytes = 16; return cudaSuccess; }
        std::vector<int> order((size_t)n);
        std::iota(order.begin(), order.end(), 0);
        const K mask = end_bit >= (int)sizeof(K) * 8 ? ~K(0) : ((K(1) << end_bit) - 1);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            return ((keys_in[a] & mask) >> begin_bit) < ((keys_in[b] & mask) >> begin_bit);
        });
        for (int i = 0; i < n; ++i) { keys_out[i] = keys_in[order[(size_t)i]]; vals_out[i] = vals_in[order[(size_t)i]]; }
        return cudaSuccess;
    }
};
}  // namespace cub
