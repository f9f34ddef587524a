// TEST INFRASTRUCTURE: host stand-ins for the two device-wide CUB primitives the engine calls (see ../cuda_runtime.h).
#pragma once

#include <algorithm>
#include <cstddef>
#include <numeric>
#include <vector>

#include <cuda_runtime.h>

namespace cub {

struct DeviceScan {
  template <class In, class Out>
  static cudaError_t ExclusiveSum(void *tmp, size_t &tmp_bytes, In in, Out out, long long n, cudaStream_t = nullptr) {
    if (!tmp) { tmp_bytes = 16; return cudaSuccess; }
    auto acc = decltype(*out + *out)(0);
    for (long long i = 0; i < n; ++i) {
      const auto v = in[i];
      out[i] = acc;
      acc += v;
    }
    return cudaSuccess;
  }
};

struct DeviceRadixSort {
  template <class K, class V>
  static cudaError_t SortPairs(void *tmp, size_t &tmp_bytes, const K *keys_in, K *keys_out, const V *vals_in, V *vals_out,
                               long long n, int begin_bit = 0, int end_bit = sizeof(K) * 8, cudaStream_t = nullptr) {
    if (!tmp) { tmp_bytes = 16; return cudaSuccess; }
    std::vector<long long> order(n);
    std::iota(order.begin(), order.end(), 0LL);
    const unsigned long long mask = end_bit - begin_bit >= 64 ? ~0ULL : ((1ULL << (end_bit - begin_bit)) - 1ULL);
    auto key = [&](long long i) { return ((unsigned long long)keys_in[i] >> begin_bit) & mask; };
    std::stable_sort(order.begin(), order.end(), [&](long long a, long long b) { return key(a) < key(b); });
    for (long long i = 0; i < n; ++i) { keys_out[i] = keys_in[order[i]]; vals_out[i] = vals_in[order[i]]; }
    return cudaSuccess;
  }
};

}  // namespace cub
