// FP64 CSR sparse matrix-vector product (what PETSc's MatMult does for the reference, solver.cpp:169).
// The operator is the reference's own CSR (int32 indices, rows / columns ascending, solver.cpp:71-96).
//
//  k_spmv_stream  (default): persistent CTAs; a producer warp streams contiguous [val | colind | rowptr] slabs of
//                 the matrix into a shared-memory ring with TMA bulk copies (cp.async.bulk + mbarrier); consumer
//                 threads own one row each (LANES == 1): they read their entries from shared memory, issue all
//                 gathers x[col] of a group of 8 entries before the first use, and sum in ascending column order --
//                 the same order as a scalar CSR loop. Consecutive rows of a finite element matrix have
//                 consecutive columns, so the gathers of a warp coalesce. Warps run decoupled: the only
//                 synchronisation is the full / empty mbarrier pair of a ring stage.
//  k_spmv_vector  : a sub-warp of 2..32 lanes per row, shuffle reduction; used for validation and for matrices
//                 whose rows do not fit a slab.
// Both can fuse the dot product x . (A x) of conjugate gradients (per-CTA partial sums, fixed order).
//
// Algorithmic bytes per product: 12 nnz + 4 (n + 1) + 8 n (x once) + 8 n (y).
#include "tma.cuh"
#include "solver.cuh"
#include "sortutil.cuh"

#include <cstdlib>

namespace tfelcu {

constexpr int kTilePad = 16;            // slack for 16-byte aligned bulk copies (8 entries of 2 bytes on either side)

// ------------------------------------------------------------------------------------------
// slabs: consecutive rows whose weight sum_r max(len_r, unit) stays below the tile capacity; with
// unit = tile_nnz / tile_rows a slab has at most tile_rows rows and tile_nnz entries.
// Slab b starts at the first row whose exclusive weight prefix reaches b * S.
// ------------------------------------------------------------------------------------------

__global__ void k_row_weights(const int32_t* __restrict__ rowptr, size_t n, int unit, unsigned long long* __restrict__ w) {
  const size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int len = rowptr[r + 1] - rowptr[r];
  w[r] = (unsigned long long)(len > unit ? len : unit);
}

__global__ void k_slab_desc(const unsigned long long* __restrict__ cum, const int32_t* __restrict__ rowptr, size_t n,
                            unsigned long long S, int n_blocks, int4* __restrict__ desc) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  int r[2];
  for (int k = 0; k < 2; ++k) {
    const unsigned long long target = (unsigned long long)(b + k) * S;
    size_t lo = 0, hi = n;
    while (lo < hi) {
      const size_t mid = lo + (hi - lo) / 2;
      if (cum[mid] < target) lo = mid + 1; else hi = mid;
    }
    r[k] = (int)lo;
  }
  if (b == n_blocks - 1) r[1] = (int)n;
  desc[b] = make_int4(r[0], r[1], rowptr[r[0]], rowptr[r[1]]);
}

__global__ void k_max_row_len(const int32_t* __restrict__ rowptr, size_t n, int* __restrict__ out) {
  const size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r < n) atomicMax(out, rowptr[r + 1] - rowptr[r]);
}

__global__ void k_col_delta16(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind, size_t n,
                              int16_t* __restrict__ delta, int* __restrict__ far) {
  const size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  for (int e = rowptr[r]; e < rowptr[r + 1]; ++e) {
    const long long d = (long long)colind[e] - (long long)r;
    if (d < -32767 || d > 32767) { atomicExch(far, 1); delta[e] = 0; }
    else delta[e] = (int16_t)d;
  }
}

// tile shapes of the streamed kernel: (entries, consumer threads, ring stages, CTAs per SM)
struct StreamCfg { int tile_nnz, consumers, stages, ctas_per_sm; };
static const StreamCfg kCfgs[] = {
    {1024, 128, 3, 5},   // 0: default (measured best inside CG on configs[1]: 0.261 ms, 6.68 TB/s)
    {2048, 256, 4, 2},   // 1
    {1024, 128, 4, 4},   // 2
    {2048, 256, 2, 4},   // 3
    {512, 64, 4, 8},     // 4
};
constexpr int kNumCfgs = (int)(sizeof(kCfgs) / sizeof(kCfgs[0]));

void spmv_configure(tfelcu_solver* s);
int spmv_launch_vector(tfelcu_solver* s, const double* d_x, double* d_y, double* dot_partial, const KrylovScalars* scal);

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
  P.cfg = 0;
  if (const char* e = std::getenv("TFELCU_SPMV_CFG")) {   // tuning knob
    const int c = std::atoi(e);
    if (c >= 0 && c < kNumCfgs) P.cfg = c;
  }
  const StreamCfg& C = kCfgs[P.cfg];
  P.row_lanes = avg > 16.0 ? 4 : 1;
  const int tile_rows = C.consumers / P.row_lanes;
  const int unit = C.tile_nnz / tile_rows;
  const long long S = (long long)C.tile_nnz - kTilePad - (max_len > unit ? max_len : unit);
  int want = s->params.spmv_kernel;
  if (want == TFELCU_SPMV_AUTO) want = TFELCU_SPMV_TMA_STREAM;
  if (want == TFELCU_SPMV_TMA_STREAM && (S < C.tile_nnz / 2 || s->n == 0)) want = TFELCU_SPMV_VECTOR;   // very long rows
  P.kernel = want;
  if (want != TFELCU_SPMV_TMA_STREAM) return;
  DevBuf<unsigned long long> w(s->n), cum(s->n);
  k_row_weights<<<grid_for(s->n, 256), 256, 0, ctx->stream>>>(s->rowptr, s->n, unit, w.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  exclusive_sum(ctx, w.p, cum.p, s->n);
  unsigned long long last_cum = 0, last_w = 0;
  ctx->download(&last_cum, cum.p + (s->n - 1), 1);
  ctx->download(&last_w, w.p + (s->n - 1), 1);
  const unsigned long long total = last_cum + last_w;
  P.n_blocks = (int)((total + (unsigned long long)S - 1) / (unsigned long long)S);
  if (P.n_blocks < 1) P.n_blocks = 1;
  P.tile_nnz = C.tile_nnz; P.tile_rows = tile_rows;
  P.desc.alloc((size_t)P.n_blocks * 4);
  k_slab_desc<<<grid_for((size_t)P.n_blocks, 256), 256, 0, ctx->stream>>>(cum.p, s->rowptr, s->n, (unsigned long long)S,
                                                                          P.n_blocks, reinterpret_cast<int4*>(P.desc.p));
  ctx->count(); TFELCU_LAUNCH_CHECK();
  // 16-bit column offsets when every entry is close enough to the diagonal (finite element matrices in reference
  // numbering are banded: +-(n + 2) for the P1 square mesh)
  P.use_delta16 = false;
  if (!std::getenv("TFELCU_NO_DELTA16") && s->nnz) {
    DevBuf<int> far(1);
    ctx->zero(far.p, 1);
    P.delta16.alloc(s->nnz + 16);
    k_col_delta16<<<grid_for(s->n, 256), 256, 0, ctx->stream>>>(s->rowptr, s->colind, s->n, P.delta16.p, far.p);
    ctx->count(); TFELCU_LAUNCH_CHECK();
    int h_far = 0; ctx->download(&h_far, far.p, 1);
    P.use_delta16 = (h_far == 0);
    if (!P.use_delta16) P.delta16.release();
  }
  ctx->sync();
  spmv_configure(s);
}

// ------------------------------------------------------------------------------------------
// TMA-streamed kernel
// ------------------------------------------------------------------------------------------

// "no entry" in the unrolled gather below; local column indices themselves may be negative (halo entries owned by lower
// ranks sit in front of the owned part of a distributed vector, dist.cu)
constexpr int kNoColumn = -2147483647 - 1;

// COL_T = int32_t: column indices as stored; int16_t: column - row (solver-internal copy, built when every stored
// entry of the operator lies within +-32767 of the diagonal -- 2 bytes less per entry to stream)
template <int TILE_NNZ, int CONSUMERS, int STAGES, typename COL_T>
struct StreamSmem {
  double val[STAGES][TILE_NNZ + kTilePad];
  COL_T col[STAGES][TILE_NNZ + kTilePad];
  int32_t rp[STAGES][CONSUMERS + 12];
  int32_t meta[STAGES][4];   // r0, r1, first copied entry a0, first copied row r0a
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  double red[CONSUMERS / 32];
};

template <bool DOT, int LANES, int TILE_NNZ, int CONSUMERS, int STAGES, int CTAS, typename COL_T>
__global__ void __launch_bounds__(CONSUMERS + 32, CTAS)
k_spmv_stream(const int32_t* __restrict__ rowptr, const COL_T* __restrict__ colind, const double* __restrict__ val,
              const int4* __restrict__ desc, int n_blocks, const double* __restrict__ x,
              double* __restrict__ y, double* __restrict__ dot_partial, const KrylovScalars* __restrict__ scal) {
  extern __shared__ __align__(128) unsigned char raw[];
  using Smem = StreamSmem<TILE_NNZ, CONSUMERS, STAGES, COL_T>;
  constexpr bool kDelta = sizeof(COL_T) == 2;
  constexpr int kAlign = kDelta ? 8 : 4;   // entries per 16 bytes of the column stream
  Smem& S = *reinterpret_cast<Smem*>(raw);
  if (scal && scal->done) {
    if (DOT && threadIdx.x == 0) dot_partial[blockIdx.x] = 0.0;
    return;
  }
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], CONSUMERS); }
    mbar_fence_init();
  }
  __syncthreads();
  const int first = blockIdx.x, step = gridDim.x;

  if (tid >= CONSUMERS) {
    // ---- producer warp: lane 0 issues, one slab descriptor prefetched ahead --------------------
    if (tid == CONSUMERS) {
      int4 d = first < n_blocks ? __ldg(desc + first) : make_int4(0, 0, 0, 0);
      int it = 0;
      for (int b = first; b < n_blocks; b += step, ++it) {
        const int stage = it % STAGES;
        const int4 cur = d;
        if (b + step < n_blocks) d = __ldg(desc + b + step);
        if (it >= STAGES) mbar_wait(&S.empty[stage], (uint32_t)(it / STAGES - 1) & 1u);
        const int r0 = cur.x, r1 = cur.y, e0 = cur.z, e1 = cur.w;
        const int a0 = e0 & ~(kAlign - 1), a1 = (e1 + kAlign - 1) & ~(kAlign - 1);
        const int cnt = a1 - a0;
        const int r0a = r0 & ~3;
        const int rcnt = ((r1 + 1 + 3) & ~3) - r0a;
        int32_t* m = S.meta[stage];
        m[0] = r0; m[1] = r1; m[2] = a0; m[3] = r0a;
        mbar_expect_tx(&S.full[stage], (uint32_t)cnt * (8u + (uint32_t)sizeof(COL_T)) + (uint32_t)rcnt * 4u);
        if (cnt > 0) {
          bulk_g2s(S.val[stage], val + a0, (uint32_t)cnt * 8u, &S.full[stage]);
          bulk_g2s(S.col[stage], colind + a0, (uint32_t)cnt * (uint32_t)sizeof(COL_T), &S.full[stage]);
        }
        bulk_g2s(S.rp[stage], rowptr + r0a, (uint32_t)rcnt * 4u, &S.full[stage]);
      }
    }
    return;
  }

  // ---- consumers: LANES threads per row ---------------------------------------------------------
  double dot = 0.0;
  int it = 0;
  const int row_in_slab = tid / LANES, lane = tid % LANES;
  for (int b = first; b < n_blocks; b += step, ++it) {
    const int stage = it % STAGES;
    mbar_wait(&S.full[stage], (uint32_t)(it / STAGES) & 1u);
    const int32_t* m = S.meta[stage];
    const int r0 = m[0], r1 = m[1], a0 = m[2], r0a = m[3];
    const double* vs = S.val[stage];
    const COL_T* cs = S.col[stage];
    const int32_t* rp = S.rp[stage];
    const int r = r0 + row_in_slab;
    const bool has = r < r1;
    double sum = 0.0;
    if (has) {
      const int rs = rp[r - r0a] - a0, re = rp[r + 1 - r0a] - a0;
      constexpr int U = 8;
      for (int e = rs + lane; e < re; e += U * LANES) {
        int c[U]; double v[U], xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int eu = e + u * LANES;
          const bool ok = eu < re;
          c[u] = ok ? (kDelta ? r + (int)cs[eu] : (int)cs[eu]) : kNoColumn;
          v[u] = ok ? vs[eu] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) xv[u] = c[u] != kNoColumn ? __ldg(x + c[u]) : 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u) sum += v[u] * xv[u];
      }
    }
    mbar_arrive(&S.empty[stage]);   // shared-memory reads of this stage are done
    if (LANES > 1) {
#pragma unroll
      for (int o = LANES / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    if (has && lane == 0) {
      y[r] = sum;
      if (DOT) dot += __ldg(x + r) * sum;
    }
  }
  if (DOT) {
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_down_sync(0xffffffffu, dot, o);
    if ((tid & 31) == 0) S.red[tid >> 5] = dot;
    asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory");
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < CONSUMERS / 32; ++w) t += S.red[w];
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

static int stream_grid(tfelcu_solver* s, int n_slabs) {
  int g = s->ctx->sm_count * kCfgs[s->plan.cfg].ctas_per_sm;
  if (g > n_slabs) g = n_slabs;
  return g < 1 ? 1 : g;
}
static int stream_grid(tfelcu_solver* s) { return stream_grid(s, s->plan.n_blocks); }

int spmv_partial_count(tfelcu_solver* s) {
  if (s->plan.kernel == TFELCU_SPMV_TMA_STREAM) return stream_grid(s);
  return (int)grid_for(s->n * (size_t)s->plan.lanes, 256);
}

// grid < 0: only opt the kernel in to its dynamic shared memory size (done at plan time, outside any graph capture)
template <bool DOT, int LANES, int TILE_NNZ, int CONSUMERS, int STAGES, int CTAS, typename COL_T>
static void launch_stream_t(tfelcu_solver* s, int grid, const COL_T* col, const double* d_x, double* d_y, double* dot_partial,
                            const KrylovScalars* scal) {
  auto kern = k_spmv_stream<DOT, LANES, TILE_NNZ, CONSUMERS, STAGES, CTAS, COL_T>;
  const size_t smem = sizeof(StreamSmem<TILE_NNZ, CONSUMERS, STAGES, COL_T>);
  if (grid < 0) {
    TFELCU_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return;
  }
  kern<<<grid, CONSUMERS + 32, smem, s->ctx->stream>>>(s->rowptr, col, s->val,
                                                      reinterpret_cast<const int4*>(s->plan.desc.p) + s->plan.part_begin,
                                                      s->plan.part_count, d_x, d_y, dot_partial, scal);
}

template <bool DOT, int LANES, int TILE_NNZ, int CONSUMERS, int STAGES, int CTAS>
static void launch_stream(tfelcu_solver* s, int grid, const double* d_x, double* d_y, double* dot_partial,
                          const KrylovScalars* scal) {
  if (grid < 0) {   // configure both column formats
    launch_stream_t<DOT, LANES, TILE_NNZ, CONSUMERS, STAGES, CTAS, int32_t>(s, grid, nullptr, d_x, d_y, dot_partial, scal);
    launch_stream_t<DOT, LANES, TILE_NNZ, CONSUMERS, STAGES, CTAS, int16_t>(s, grid, nullptr, d_x, d_y, dot_partial, scal);
    return;
  }
  if (s->plan.use_delta16)
    launch_stream_t<DOT, LANES, TILE_NNZ, CONSUMERS, STAGES, CTAS, int16_t>(s, grid, s->plan.delta16.p, d_x, d_y, dot_partial, scal);
  else
    launch_stream_t<DOT, LANES, TILE_NNZ, CONSUMERS, STAGES, CTAS, int32_t>(s, grid, s->colind, d_x, d_y, dot_partial, scal);
}

template <int TILE_NNZ, int CONSUMERS, int STAGES, int CTAS>
static void launch_stream_cfg(tfelcu_solver* s, int grid, const double* d_x, double* d_y, double* dot_partial,
                              const KrylovScalars* scal) {
  if (grid < 0) {
    launch_stream<true, 4, TILE_NNZ, CONSUMERS, STAGES, CTAS>(s, grid, d_x, d_y, dot_partial, scal);
    launch_stream<false, 4, TILE_NNZ, CONSUMERS, STAGES, CTAS>(s, grid, d_x, d_y, dot_partial, scal);
    launch_stream<true, 1, TILE_NNZ, CONSUMERS, STAGES, CTAS>(s, grid, d_x, d_y, dot_partial, scal);
    launch_stream<false, 1, TILE_NNZ, CONSUMERS, STAGES, CTAS>(s, grid, d_x, d_y, dot_partial, scal);
    return;
  }
  if (s->plan.row_lanes == 4) {
    if (dot_partial) launch_stream<true, 4, TILE_NNZ, CONSUMERS, STAGES, CTAS>(s, grid, d_x, d_y, dot_partial, scal);
    else launch_stream<false, 4, TILE_NNZ, CONSUMERS, STAGES, CTAS>(s, grid, d_x, d_y, dot_partial, scal);
  } else {
    if (dot_partial) launch_stream<true, 1, TILE_NNZ, CONSUMERS, STAGES, CTAS>(s, grid, d_x, d_y, dot_partial, scal);
    else launch_stream<false, 1, TILE_NNZ, CONSUMERS, STAGES, CTAS>(s, grid, d_x, d_y, dot_partial, scal);
  }
}

static void stream_dispatch(tfelcu_solver* s, int grid, const double* d_x, double* d_y, double* dot_partial,
                            const KrylovScalars* scal);

int spmv_launch(tfelcu_solver* s, const double* d_x, double* d_y, double* dot_partial, const KrylovScalars* scal) {
  Ctx* ctx = s->ctx;
  if (s->plan.kernel == TFELCU_SPMV_TMA_STREAM) {
    const int grid = stream_grid(s);
    s->plan.part_begin = 0; s->plan.part_count = s->plan.n_blocks;
    stream_dispatch(s, grid, d_x, d_y, dot_partial, scal);
    ctx->count(); TFELCU_LAUNCH_CHECK();
    return grid;
  }
  return spmv_launch_vector(s, d_x, d_y, dot_partial, scal);
}

// part 1: the slabs whose rows reference owned columns only (can run while the halo travels); part 2: the others.
// Returns the number of dot partials written (0 when the part is empty and nothing was launched).
int spmv_launch_part(tfelcu_solver* s, int part, const double* d_x, double* d_y, double* dot_partial, const KrylovScalars* scal) {
  Ctx* ctx = s->ctx;
  TFELCU_REQUIRE(s->plan.kernel == TFELCU_SPMV_TMA_STREAM && s->plan.n_interior >= 0, "SpMV plan has no interior / boundary split");
  const int begin = part == 1 ? 0 : s->plan.n_interior;
  const int count = part == 1 ? s->plan.n_interior : s->plan.n_blocks - s->plan.n_interior;
  if (count <= 0) return 0;
  const int grid = stream_grid(s, count);
  s->plan.part_begin = begin; s->plan.part_count = count;
  stream_dispatch(s, grid, d_x, d_y, dot_partial, scal);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  return grid;
}

// slab b references a column another rank owns
__global__ void k_slab_is_boundary(const int4* __restrict__ desc, int n_blocks, const int32_t* __restrict__ colind, int n_local,
                                   int* __restrict__ flag) {
  const int b = blockIdx.x;
  if (b >= n_blocks) return;
  const int4 d = desc[b];
  int any = 0;
  for (int e = d.z + threadIdx.x; e < d.w; e += blockDim.x) any |= (colind[e] < 0 || colind[e] >= n_local) ? 1 : 0;
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) flag[b] = any;
}

// reorder the slab descriptors: interior slabs first, boundary slabs last (stable); sets plan.n_interior
void spmv_split_boundary(tfelcu_solver* s, size_t n_local) {
  Ctx* ctx = s->ctx;
  SpmvPlan& P = s->plan;
  P.n_interior = -1;
  if (P.kernel != TFELCU_SPMV_TMA_STREAM || P.n_blocks <= 0) return;
  DevBuf<int> flag((size_t)P.n_blocks);
  k_slab_is_boundary<<<P.n_blocks, 128, 0, ctx->stream>>>(reinterpret_cast<const int4*>(P.desc.p), P.n_blocks, s->colind, (int)n_local, flag.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  std::vector<int> h_flag((size_t)P.n_blocks);
  std::vector<int32_t> h_desc((size_t)P.n_blocks * 4), out((size_t)P.n_blocks * 4);
  ctx->download(h_flag.data(), flag.p, h_flag.size());
  ctx->download(h_desc.data(), P.desc.p, h_desc.size());
  size_t at = 0;
  for (int pass = 0; pass < 2; ++pass) {
    for (int b = 0; b < P.n_blocks; ++b)
      if ((h_flag[b] != 0) == (pass == 1)) { for (int k = 0; k < 4; ++k) out[at * 4 + k] = h_desc[(size_t)b * 4 + k]; ++at; }
    if (pass == 0) P.n_interior = (int)at;
  }
  ctx->upload(P.desc.p, out.data(), out.size());
  ctx->sync();
}

static void stream_dispatch(tfelcu_solver* s, int grid, const double* d_x, double* d_y, double* dot_partial,
                            const KrylovScalars* scal) {
    switch (s->plan.cfg) {
      case 1: launch_stream_cfg<2048, 256, 4, 2>(s, grid, d_x, d_y, dot_partial, scal); break;
      case 2: launch_stream_cfg<1024, 128, 4, 4>(s, grid, d_x, d_y, dot_partial, scal); break;
      case 3: launch_stream_cfg<2048, 256, 2, 4>(s, grid, d_x, d_y, dot_partial, scal); break;
      case 4: launch_stream_cfg<512, 64, 4, 8>(s, grid, d_x, d_y, dot_partial, scal); break;
      default: launch_stream_cfg<1024, 128, 3, 5>(s, grid, d_x, d_y, dot_partial, scal); break;
    }
}

void spmv_configure(tfelcu_solver* s) {
  if (s->plan.kernel == TFELCU_SPMV_TMA_STREAM) stream_dispatch(s, -1, nullptr, nullptr, nullptr, nullptr);
}

int spmv_launch_vector(tfelcu_solver* s, const double* d_x, double* d_y, double* dot_partial, const KrylovScalars* scal) {
  Ctx* ctx = s->ctx;
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
