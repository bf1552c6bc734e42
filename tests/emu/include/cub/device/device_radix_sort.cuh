// TEST INFRASTRUCTURE ONLY (see ../../cuda_runtime.h): cub::DeviceRadixSort::SortPairs as a stable host sort on the
// key bits [begin_bit, end_bit), which is all rf_order.cuh asks of the library.
#pragma once
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <vector>

#include <cuda_runtime.h>

namespace cub {
struct DeviceRadixSort {
  template <class K, class V>
  static cudaError_t SortPairs(void* tmp, size_t& tmp_bytes, const K* kin, K* kout, const V* vin, V* vout, int n, int begin_bit = 0,
                               int end_bit = (int)sizeof(K) * 8, cudaStream_t = nullptr) {
    if (!tmp) { tmp_bytes = 1; return cudaSuccess; }
    const int nb = end_bit - begin_bit;
    const unsigned long long mask = nb >= 64 ? ~0ull : ((1ull << nb) - 1ull);
    std::vector<int> idx((size_t)n);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
      return (((unsigned long long)kin[a] >> begin_bit) & mask) < (((unsigned long long)kin[b] >> begin_bit) & mask);
    });
    for (int i = 0; i < n; i++) { kout[i] = kin[idx[(size_t)i]]; vout[i] = vin[idx[(size_t)i]]; }
    return cudaSuccess;
  }
};
}  // namespace cub
