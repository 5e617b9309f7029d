// stand-in for <cub/cub.cuh> (CPU emulator build only): the two device-wide primitives the build
// path calls, with cub's two-phase calling convention (null temp storage = size query)
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>
namespace cub {
struct DeviceRadixSort {
    template <class K, class V>
    static cudaError_t SortPairs(void* tmp, size_t& tmp_bytes, const K* keys_in, K* keys_out, const V* vals_in,
                                 V* vals_out, long long n, int begin_bit = 0, int end_bit = sizeof(K) * 8,
                                 cudaStream_t = nullptr) {
        if (!tmp) { tmp_bytes = 16; return 0; }
        std::vector<long long> order((size_t)n);
        std::iota(order.begin(), order.end(), 0ll);
        const int nb = end_bit - begin_bit;
        auto key = [&](long long i) {
            unsigned long long k = (unsigned long long)(typename std::make_unsigned<K>::type)keys_in[i];
            k >>= begin_bit;
            return nb >= 64 ? k : (k & ((1ull << nb) - 1ull));
        };
        std::stable_sort(order.begin(), order.end(), [&](long long a, long long b) { return key(a) < key(b); });
        std::vector<K> ko((size_t)n);
        std::vector<V> vo((size_t)n);
        for (long long i = 0; i < n; i++) { ko[(size_t)i] = keys_in[order[(size_t)i]]; vo[(size_t)i] = vals_in[order[(size_t)i]]; }
        std::copy(ko.begin(), ko.end(), keys_out);
        std::copy(vo.begin(), vo.end(), vals_out);
        return 0;
    }
};
struct DeviceScan {
    template <class In, class Out>
    static cudaError_t ExclusiveSum(void* tmp, size_t& tmp_bytes, const In* in, Out* out, long long n,
                                    cudaStream_t = nullptr) {
        if (!tmp) { tmp_bytes = 16; return 0; }
        Out acc = 0;
        for (long long i = 0; i < n; i++) { const Out v = (Out)in[i]; out[i] = acc; acc += v; }
        return 0;
    }
};
}  // namespace cub
