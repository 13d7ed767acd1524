// FP64 CSR sparse matrix-vector product (what PETSc's MatMult does for the reference, solver.cpp:169).
// The operator is the reference's own CSR (int32 indices, rows / columns ascending, solver.cpp:71-96).
//
//  k_spmv_stream  (default): persistent CTAs stream contiguous [val | colind] slabs of the matrix into shared
//                 memory with TMA bulk copies (cp.async.bulk + mbarrier, 3-stage ring), form the products
//                 val * x[col] with entry-strided (coalesced, conflict-free) accesses, then reduce each row
//                 sequentially in ascending column order -- the same order and rounding as a scalar CSR loop.
//  k_spmv_vector  : a sub-warp of 2..32 lanes per row, shuffle reduction; used for validation and for matrices
//                 whose rows do not fit a slab.
// Both can fuse the dot product x . (A x) of conjugate gradients (per-CTA partial sums, fixed order).
//
// Algorithmic bytes per product: 12 nnz + 4 (n + 1) + 8 n (x once) + 8 n (y).
#include "solver.cuh"

namespace tfelcu {

constexpr int kStreamThreads = 256;
constexpr int kStreamStages = 3;
constexpr int kTileNnz = 2048;          // capacity of one slab (entries)
constexpr int kTilePad = 8;             // slack for 16-byte aligned bulk copies
constexpr int kRowWeight = 4;           // a row counts as len + kRowWeight when slabs are cut

// ------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk copy global -> shared (SASS: UBLKCP / SYNCS)
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// slab boundaries: slab b starts at the first row r with rowptr[r] + kRowWeight * r >= b * S
// ------------------------------------------------------------------------------------------

__global__ void k_slab_rows(const int32_t* __restrict__ rowptr, size_t n, long long S, int n_blocks,
                            int32_t* __restrict__ block_row) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > n_blocks) return;
  if (b == n_blocks) { block_row[b] = (int32_t)n; return; }
  const long long target = (long long)b * S;
  size_t lo = 0, hi = n;
  while (lo < hi) {
    const size_t mid = lo + (hi - lo) / 2;
    if ((long long)rowptr[mid] + (long long)kRowWeight * (long long)mid < target) lo = mid + 1; else hi = mid;
  }
  block_row[b] = (int32_t)lo;
}

__global__ void k_max_row_len(const int32_t* __restrict__ rowptr, size_t n, int* __restrict__ out) {
  const size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r < n) atomicMax(out, rowptr[r + 1] - rowptr[r]);
}

void spmv_plan(tfelcu_solver* s) {
  Ctx* ctx = s->ctx;
  SpmvPlan& P = s->plan;
  const double avg = s->n ? (double)s->nnz / (double)s->n : 1.0;
  P.lanes = avg <= 2.5 ? 2 : avg <= 5 ? 4 : avg <= 12 ? 8 : avg <= 24 ? 16 : 32;
  DevBuf<int> mx(1);
  ctx->zero(mx.p, 1);
  k_max_row_len<<<grid_for(s->n, 256), 256, 0, ctx->stream>>>(s->rowptr, s->n, mx.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  int max_len = 0; ctx->download(&max_len, mx.p, 1);
  const long long S = (long long)kTileNnz - max_len - kRowWeight;
  int want = s->params.spmv_kernel;
  if (want == TFELCU_SPMV_AUTO) want = TFELCU_SPMV_TMA_STREAM;
  if (want == TFELCU_SPMV_TMA_STREAM && (S < kTileNnz / 2 || s->n == 0)) want = TFELCU_SPMV_VECTOR;   // very long rows
  P.kernel = want;
  if (want != TFELCU_SPMV_TMA_STREAM) return;
  const long long total = (long long)s->nnz + (long long)kRowWeight * (long long)s->n;
  P.n_blocks = (int)((total + S - 1) / S);
  if (P.n_blocks < 1) P.n_blocks = 1;
  P.tile_nnz = kTileNnz; P.tile_rows = kTileNnz / kRowWeight;
  P.block_row.alloc((size_t)P.n_blocks + 1);
  k_slab_rows<<<grid_for((size_t)P.n_blocks + 1, 256), 256, 0, ctx->stream>>>(s->rowptr, s->n, S, P.n_blocks, P.block_row.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------
// TMA-streamed kernel
// ------------------------------------------------------------------------------------------

struct StreamSmem {
  double val[kStreamStages][kTileNnz + kTilePad];
  int32_t col[kStreamStages][kTileNnz + kTilePad];
  uint64_t full[kStreamStages];
  int32_t meta[kStreamStages][4];   // r0, r1, aligned first entry, entries copied
  double red[kStreamThreads / 32];
};

template <bool DOT>
__global__ void __launch_bounds__(kStreamThreads, 2)
k_spmv_stream(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind, const double* __restrict__ val,
              const int32_t* __restrict__ block_row, int n_blocks, const double* __restrict__ x,
              double* __restrict__ y, double* __restrict__ dot_partial, const KrylovScalars* __restrict__ scal) {
  extern __shared__ __align__(128) unsigned char raw[];
  StreamSmem& S = *reinterpret_cast<StreamSmem*>(raw);
  if (scal && scal->done) {
    if (DOT && threadIdx.x == 0) dot_partial[blockIdx.x] = 0.0;
    return;
  }
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < kStreamStages; ++s) mbar_init(&S.full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  // producer (thread 0): one slab = two bulk copies landing on the stage's mbarrier
  auto issue = [&](int b, int stage) {
    const int r0 = block_row[b], r1 = block_row[b + 1];
    const int e0 = rowptr[r0], e1 = rowptr[r1];
    const int a0 = e0 & ~3, a1 = (e1 + 3) & ~3;
    const int cnt = a1 - a0;
    S.meta[stage][0] = r0; S.meta[stage][1] = r1; S.meta[stage][2] = a0; S.meta[stage][3] = cnt;
    if (cnt > 0) {
      mbar_expect_tx(&S.full[stage], (uint32_t)cnt * 12u);
      bulk_g2s(S.val[stage], val + a0, (uint32_t)cnt * 8u, &S.full[stage]);
      bulk_g2s(S.col[stage], colind + a0, (uint32_t)cnt * 4u, &S.full[stage]);
    } else {
      mbar_expect_tx(&S.full[stage], 0u);
    }
  };

  const int first = blockIdx.x, step = gridDim.x;
  if (tid == 0)
    for (int s = 0; s < kStreamStages; ++s)
      if (first + s * step < n_blocks) issue(first + s * step, s);
  __syncthreads();   // meta of the first slabs is visible

  double dot = 0.0;
  int it = 0;
  for (int b = first; b < n_blocks; b += step, ++it) {
    const int stage = it % kStreamStages;
    const uint32_t parity = (uint32_t)(it / kStreamStages) & 1u;
    mbar_wait(&S.full[stage], parity);
    const int r0 = S.meta[stage][0], r1 = S.meta[stage][1], a0 = S.meta[stage][2], cnt = S.meta[stage][3];
    double* vs = S.val[stage];
    const int32_t* cs = S.col[stage];
    // products, entry-strided: every gather is independent of the others
    for (int e = tid; e < cnt; e += kStreamThreads) {
      const int ge = a0 + e;
      // entries before the first row / after the last row of the slab belong to neighbours: skip the gather
      const int32_t c = cs[e];
      vs[e] = (ge >= rowptr[r0] && ge < rowptr[r1]) ? vs[e] * __ldg(x + c) : 0.0;
    }
    __syncthreads();
    // one thread per row, ascending column order
    for (int r = r0 + tid; r < r1; r += kStreamThreads) {
      const int rs = rowptr[r] - a0, re = rowptr[r + 1] - a0;
      double sum = 0.0;
      for (int e = rs; e < re; ++e) sum += vs[e];
      y[r] = sum;
      if (DOT) dot += __ldg(x + r) * sum;
    }
    fence_proxy_async();   // generic-proxy accesses of this stage are ordered before the next bulk copy into it
    __syncthreads();
    if (tid == 0) {
      const int nb = b + kStreamStages * step;
      if (nb < n_blocks) issue(nb, stage);
    }
    __syncthreads();       // meta written by thread 0 for a later iteration of this stage
  }
  if (DOT) {
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_down_sync(0xffffffffu, dot, o);
    if ((tid & 31) == 0) S.red[tid >> 5] = dot;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < kStreamThreads / 32; ++w) t += S.red[w];
      dot_partial[blockIdx.x] = t;
    }
  }
}

// ------------------------------------------------------------------------------------------
// sub-warp per row kernel
// ------------------------------------------------------------------------------------------

template <int LANES, bool DOT>
__global__ void __launch_bounds__(256)
k_spmv_vector(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind, const double* __restrict__ val,
              size_t n, const double* __restrict__ x, double* __restrict__ y, double* __restrict__ dot_partial,
              const KrylovScalars* __restrict__ scal) {
  __shared__ double red[8];
  if (scal && scal->done) {
    if (DOT && threadIdx.x == 0) dot_partial[blockIdx.x] = 0.0;
    return;
  }
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t row = t / LANES;
  const int lane = (int)(t % LANES);
  double sum = 0.0;
  if (row < n) {
    const int rs = rowptr[row], re = rowptr[row + 1];
    for (int e = rs + lane; e < re; e += LANES) sum += val[e] * __ldg(x + colind[e]);
  }
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o, LANES);
  double dot = 0.0;
  if (row < n && lane == 0) {
    y[row] = sum;
    if (DOT) dot = __ldg(x + row) * sum;
  }
  if (DOT) {
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_down_sync(0xffffffffu, dot, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < 8; ++w) s += red[w];
      dot_partial[blockIdx.x] = s;
    }
  }
}

static int stream_grid(tfelcu_solver* s) {
  int g = s->ctx->sm_count * 2;
  if (g > s->plan.n_blocks) g = s->plan.n_blocks;
  return g < 1 ? 1 : g;
}

int spmv_partial_count(tfelcu_solver* s) {
  if (s->plan.kernel == TFELCU_SPMV_TMA_STREAM) return stream_grid(s);
  return (int)grid_for(s->n * (size_t)s->plan.lanes, 256);
}

int spmv_launch(tfelcu_solver* s, const double* d_x, double* d_y, double* dot_partial, const KrylovScalars* scal) {
  Ctx* ctx = s->ctx;
  if (s->plan.kernel == TFELCU_SPMV_TMA_STREAM) {
    const int grid = stream_grid(s);
    const size_t smem = sizeof(StreamSmem);
    static bool configured = false;
    if (!configured) {
      TFELCU_CK(cudaFuncSetAttribute(k_spmv_stream<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      TFELCU_CK(cudaFuncSetAttribute(k_spmv_stream<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = true;
    }
    if (dot_partial)
      k_spmv_stream<true><<<grid, kStreamThreads, smem, ctx->stream>>>(s->rowptr, s->colind, s->val, s->plan.block_row.p,
                                                                      s->plan.n_blocks, d_x, d_y, dot_partial, scal);
    else
      k_spmv_stream<false><<<grid, kStreamThreads, smem, ctx->stream>>>(s->rowptr, s->colind, s->val, s->plan.block_row.p,
                                                                       s->plan.n_blocks, d_x, d_y, nullptr, scal);
    ctx->count(); TFELCU_LAUNCH_CHECK();
    return grid;
  }
  const unsigned grid = grid_for(s->n * (size_t)s->plan.lanes, 256);
#define TFELCU_VEC(L)                                                                                             \
  do {                                                                                                            \
    if (dot_partial)                                                                                              \
      k_spmv_vector<L, true><<<grid, 256, 0, ctx->stream>>>(s->rowptr, s->colind, s->val, s->n, d_x, d_y, dot_partial, scal); \
    else                                                                                                          \
      k_spmv_vector<L, false><<<grid, 256, 0, ctx->stream>>>(s->rowptr, s->colind, s->val, s->n, d_x, d_y, nullptr, scal);    \
  } while (0)
  switch (s->plan.lanes) {
    case 2: TFELCU_VEC(2); break;
    case 4: TFELCU_VEC(4); break;
    case 8: TFELCU_VEC(8); break;
    case 16: TFELCU_VEC(16); break;
    default: TFELCU_VEC(32); break;
  }
#undef TFELCU_VEC
  ctx->count(); TFELCU_LAUNCH_CHECK();
  return (int)grid;
}

void solver_spmv(tfelcu_solver* s, const double* d_x, double* d_y) { spmv_launch(s, d_x, d_y, nullptr, nullptr); }

}  // namespace tfelcu
