// Thin wrappers over CUB device-wide primitives (radix sort, scan, select) used by the one-time
// structure build (DOF numbering, boundary, sparsity pattern). Library code, not a hot kernel.
#pragma once

#include <cub/cub.cuh>

#include "common.cuh"

namespace tfelcu {

inline int bits_for(size_t n) {   // bits needed to represent values in [0, n)
  int b = 1;
  while (b < 64 && (n >> b)) ++b;
  return b;
}

template <typename T>
__global__ void k_iota(T* __restrict__ out, size_t n) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (T)i;
}

inline void iota_u32(Ctx* ctx, uint32_t* out, size_t n) {
  if (!n) return;
  k_iota<uint32_t><<<grid_for(n, 256), 256, 0, ctx->stream>>>(out, n);
  ctx->count(); TFELCU_LAUNCH_CHECK();
}

template <typename In, typename Out>
inline void exclusive_sum(Ctx* ctx, const In* in, Out* out, size_t n) {
  if (!n) return;
  size_t bytes = 0;
  TFELCU_CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int64_t)n, ctx->stream));
  DevBuf<char> tmp(bytes);
  TFELCU_CK(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, (int64_t)n, ctx->stream));
  ctx->count(2);
  ctx->sync();
}

template <typename In, typename Out>
inline void inclusive_sum(Ctx* ctx, const In* in, Out* out, size_t n) {
  if (!n) return;
  size_t bytes = 0;
  TFELCU_CK(cub::DeviceScan::InclusiveSum(nullptr, bytes, in, out, (int64_t)n, ctx->stream));
  DevBuf<char> tmp(bytes);
  TFELCU_CK(cub::DeviceScan::InclusiveSum(tmp.p, bytes, in, out, (int64_t)n, ctx->stream));
  ctx->count(2);
  ctx->sync();
}

// stable LSD radix sort of (key, value) pairs on key bits [begin_bit, end_bit)
template <typename K, typename V>
inline void sort_pairs(Ctx* ctx, const K* key_in, K* key_out, const V* val_in, V* val_out, size_t n,
                       int begin_bit, int end_bit) {
  if (!n) return;
  if (end_bit > (int)sizeof(K) * 8) end_bit = (int)sizeof(K) * 8;
  size_t bytes = 0;
  TFELCU_CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, key_in, key_out, val_in, val_out, (int64_t)n,
                                            begin_bit, end_bit, ctx->stream));
  DevBuf<char> tmp(bytes);
  TFELCU_CK(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, key_in, key_out, val_in, val_out, (int64_t)n,
                                            begin_bit, end_bit, ctx->stream));
  ctx->count((uint64_t)((end_bit - begin_bit + 7) / 8) * 2 + 1);
  ctx->sync();
}

inline void sort_pairs_u64_u32(Ctx* ctx, const uint64_t* key_in, uint64_t* key_out, const uint32_t* val_in,
                               uint32_t* val_out, size_t n, int begin_bit, int end_bit) {
  sort_pairs<uint64_t, uint32_t>(ctx, key_in, key_out, val_in, val_out, n, begin_bit, end_bit);
}

template <typename K>
inline void sort_keys(Ctx* ctx, const K* key_in, K* key_out, size_t n, int begin_bit, int end_bit) {
  if (!n) return;
  if (end_bit > (int)sizeof(K) * 8) end_bit = (int)sizeof(K) * 8;
  size_t bytes = 0;
  TFELCU_CK(cub::DeviceRadixSort::SortKeys(nullptr, bytes, key_in, key_out, (int64_t)n, begin_bit, end_bit, ctx->stream));
  DevBuf<char> tmp(bytes);
  TFELCU_CK(cub::DeviceRadixSort::SortKeys(tmp.p, bytes, key_in, key_out, (int64_t)n, begin_bit, end_bit, ctx->stream));
  ctx->count((uint64_t)((end_bit - begin_bit + 7) / 8) * 2 + 1);
  ctx->sync();
}

// out = unique consecutive keys of in; returns the number kept
template <typename K>
inline size_t unique_keys(Ctx* ctx, const K* in, K* out, size_t n) {
  if (!n) return 0;
  DevBuf<int64_t> d_count(1);
  size_t bytes = 0;
  TFELCU_CK(cub::DeviceSelect::Unique(nullptr, bytes, in, out, d_count.p, (int64_t)n, ctx->stream));
  DevBuf<char> tmp(bytes);
  TFELCU_CK(cub::DeviceSelect::Unique(tmp.p, bytes, in, out, d_count.p, (int64_t)n, ctx->stream));
  ctx->count(2);
  int64_t h = 0;
  ctx->download(&h, d_count.p, 1);
  return (size_t)h;
}

}  // namespace tfelcu
