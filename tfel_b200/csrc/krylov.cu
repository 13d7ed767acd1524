// solver::cuda::gmres_* -- restarted GMRES and the level-scheduled ILU(0) preconditioner.
// Replaces what PETSc does for the reference behind solver::petsc::gmres_ilu (src/core/solver.hpp:139-163:
// KSPGMRES with modified Gram-Schmidt, left preconditioning, PCILU in natural ordering, x0 = 0, convergence
// on the preconditioned residual:  ||M^-1 r|| <= max(rtol ||M^-1 b||, atol), divergence at dtol ||M^-1 b||).
//
// ILU(0): the factors live on the pattern of A (the reference's CSR, columns ascending). Rows are grouped in
// dependency levels (level(i) = 1 + max level of the rows i needs); the numeric factorisation and the two
// triangular solves walk the levels, one thread per row, each row summed in ascending column order -- the same
// operation order as the sequential IKJ algorithm, so the factors are bitwise those of a scalar loop.
#include "solver.cuh"
#include "sortutil.cuh"

#include <algorithm>
#include <cstdlib>

namespace tfelcu {

constexpr int kVecT = 256;

static int vgrid(tfelcu_solver* s) {
  size_t want = (s->n + kVecT - 1) / kVecT;
  const size_t cap = (size_t)s->ctx->sm_count * 4;
  if (want > cap) want = cap;
  return want < 1 ? 1 : (int)want;
}

// ------------------------------------------------------------------------------------------
// ILU(0): symbolic (diagonal positions, levels)
// ------------------------------------------------------------------------------------------

__global__ void k_diag_pos(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind, size_t n,
                           int32_t* __restrict__ diag, int* __restrict__ missing) {
  const size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  int lo = rowptr[r], hi = rowptr[r + 1];
  const int end = hi;
  while (lo < hi) {
    const int mid = lo + ((hi - lo) >> 1);   // lo + hi overflows int32 beyond 2^30 stored entries
    if (colind[mid] < (int32_t)r) lo = mid + 1; else hi = mid;
  }
  if (lo >= end || colind[lo] != (int32_t)r) { atomicExch(missing, 1); diag[r] = -1; }
  else diag[r] = lo;
}

// level of every row in one pass: CTAs take tickets so that rows start in ascending (LOWER) or descending order;
// a row spins until the rows it depends on -- all of which started earlier -- have published their level.
template <bool LOWER>
__global__ void __launch_bounds__(128)
k_levels(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind, const int32_t* __restrict__ diag,
         size_t n, unsigned int* __restrict__ ticket, volatile int* __restrict__ level, int* __restrict__ stuck) {
  __shared__ unsigned int blk;
  if (threadIdx.x == 0) blk = atomicAdd(ticket, 1u);
  __syncthreads();
  const size_t t = (size_t)blk * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const size_t r = LOWER ? t : n - 1 - t;
  const int e0 = LOWER ? rowptr[r] : diag[r] + 1;
  const int e1 = LOWER ? diag[r] : rowptr[r + 1];
  int lv = 0;
  for (int e = e0; e < e1; ++e) {
    const int c = colind[e];
    int l;
    const long long t0 = clock64();
    while ((l = level[c]) < 0) {
      if (clock64() - t0 > 40000000000ll) { atomicExch(stuck, 1); l = 0; break; }   // never expected: do not hang the GPU
      __nanosleep(20);
    }
    lv = l + 1 > lv ? l + 1 : lv;
  }
  __threadfence();
  level[r] = lv;
}

__global__ void k_max_int(const int* __restrict__ a, size_t n, int* __restrict__ out) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) atomicMax(out, a[i]);
}

__global__ void k_lower_bounds_i32(const uint32_t* __restrict__ sorted, size_t n, int n_targets, int32_t* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > n_targets) return;
  size_t lo = 0, hi = n;
  while (lo < hi) {
    const size_t mid = lo + (hi - lo) / 2;
    if (sorted[mid] < (uint32_t)t) lo = mid + 1; else hi = mid;
  }
  out[t] = (int32_t)lo;
}

// rows grouped by level (ascending row id inside a level) + start of every level
static void schedule(Ctx* ctx, const int* d_level, size_t n, DevBuf<int32_t>& order, std::vector<int32_t>& level_ptr) {
  DevBuf<int> mx(1);
  ctx->zero(mx.p, 1);
  k_max_int<<<grid_for(n, 256), 256, 0, ctx->stream>>>(d_level, n, mx.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  int n_levels = 0; ctx->download(&n_levels, mx.p, 1);
  n_levels += 1;
  DevBuf<uint32_t> ids(n), keys_sorted(n);
  iota_u32(ctx, ids.p, n);
  order.alloc(n);
  sort_pairs<uint32_t, uint32_t>(ctx, reinterpret_cast<const uint32_t*>(d_level), keys_sorted.p, ids.p,
                                 reinterpret_cast<uint32_t*>(order.p), n, 0, bits_for((size_t)n_levels + 1));
  DevBuf<int32_t> ptr((size_t)n_levels + 1);
  k_lower_bounds_i32<<<grid_for((size_t)n_levels + 1, 256), 256, 0, ctx->stream>>>(keys_sorted.p, n, n_levels, ptr.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  level_ptr.resize((size_t)n_levels + 1);
  ctx->download(level_ptr.data(), ptr.p, (size_t)n_levels + 1);
}

// ------------------------------------------------------------------------------------------
// ILU(0): numeric factorisation and triangular solves, one launch per group of levels
// ------------------------------------------------------------------------------------------

// IKJ elimination of one row (the scalar algorithm of a sequential ILU(0)): for k < i in the pattern of row i,
// l = a_ik / u_kk, then a_ij -= l u_kj for the j > k present in BOTH rows.
__device__ __forceinline__ void ilu0_row(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                                         const int32_t* __restrict__ diag, double* __restrict__ lu, int i) {
  const int rs = rowptr[i], re = rowptr[i + 1], di = diag[i];
  for (int e = rs; e < di; ++e) {
    const int k = colind[e];
    const double l = lu[e] / __ldcg(lu + diag[k]);
    lu[e] = l;
    if (l == 0.0) continue;
    int ei = e + 1;
    const int ke = rowptr[k + 1];
    for (int ek = diag[k] + 1; ek < ke; ++ek) {
      const int c = colind[ek];
      while (ei < re && colind[ei] < c) ++ei;
      if (ei < re && colind[ei] == c) lu[ei] -= l * __ldcg(lu + ek);
    }
  }
}

__global__ void k_ilu0_level(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                             const int32_t* __restrict__ diag, const int32_t* __restrict__ order, int begin, int end,
                             double* __restrict__ lu) {
  const int t = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (t < end) ilu0_row(rowptr, colind, diag, lu, order[t]);
}

// several consecutive small levels in one CTA (levels separated by a CTA barrier)
__global__ void __launch_bounds__(1024)
k_ilu0_levels_cta(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind, const int32_t* __restrict__ diag,
                  const int32_t* __restrict__ order, const int32_t* __restrict__ level_ptr, int l0, int l1,
                  double* __restrict__ lu) {
  for (int l = l0; l < l1; ++l) {
    const int t = level_ptr[l] + threadIdx.x;
    if (t < level_ptr[l + 1]) ilu0_row(rowptr, colind, diag, lu, order[t]);
    __threadfence_block();
    __syncthreads();
  }
}

template <bool LOWER>
__device__ __forceinline__ void tri_row(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                                        const int32_t* __restrict__ diag, const double* __restrict__ lu,
                                        const double* in, double* out, int i) {
  // forward: x_i = b_i - sum_{j<i} l_ij x_j ; backward: x_i = (y_i - sum_{j>i} u_ij x_j) / u_ii
  double acc = in[i];
  if (LOWER) {
    for (int e = rowptr[i]; e < diag[i]; ++e) acc -= lu[e] * __ldcg(out + colind[e]);
    out[i] = acc;
  } else {
    for (int e = diag[i] + 1; e < rowptr[i + 1]; ++e) acc -= lu[e] * __ldcg(out + colind[e]);
    out[i] = acc / lu[diag[i]];
  }
}

template <bool LOWER>
__global__ void k_tri_level(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                            const int32_t* __restrict__ diag, const double* __restrict__ lu,
                            const int32_t* __restrict__ order, int begin, int end, const double* in,
                            double* out, const KrylovScalars* __restrict__ S) {
  if (S && S->done) return;
  const int t = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (t < end) tri_row<LOWER>(rowptr, colind, diag, lu, in, out, order[t]);
}

// Several consecutive small levels in one CTA. A row is shared by `lpr` lanes (its entries strided over them, partial
// sums combined with shuffles): a level then costs a few dependent loads instead of one per entry of its longest row.
// The values of the rows computed inside the run live in SHARED memory (indexed by schedule position), rows computed
// by earlier launches are read from global memory. Metadata and the first entries of a lane in the next level (row,
// columns, factors, schedule positions of the columns: nothing that depends on the solution) are requested one level
// ahead of the barrier.
template <bool LOWER>
__global__ void __launch_bounds__(1024)
k_tri_levels_cta(const int32_t* __restrict__ colind, const int32_t* __restrict__ diag, const double* __restrict__ lu,
                 const int4* __restrict__ meta, const int32_t* __restrict__ sched_pos, const int32_t* __restrict__ level_ptr,
                 int l0, int l1, int lpr, const double* in, double* out, const KrylovScalars* __restrict__ S) {
  extern __shared__ double xs[];   // [rows of the run]
  if (S && S->done) return;
  constexpr int PF = 2;            // entries per lane kept in registers ahead of the barrier
  const int slot = (int)threadIdx.x / lpr, lane = (int)threadIdx.x % lpr;
  const int base = level_ptr[l0], run_rows = level_ptr[l1] - base;
  int lp_next = level_ptr[l0 + 1];
  int4 m = make_int4(-1, 0, 0, 0);
  int t_cur = -1;
  double acc = 0.0, ud = 1.0;
  int c[PF], ps[PF]; double v[PF];
  auto fetch_meta = [&](int first, int end) {
    const int t = first + slot;
    t_cur = t;
    m = t < end ? __ldg(meta + t) : make_int4(-1, 0, 0, 0);
  };
  auto fetch_entries = [&]() {
#pragma unroll
    for (int u = 0; u < PF; ++u) { c[u] = -1; v[u] = 0.0; ps[u] = -1; }
    if (m.x < 0) return;
    if (lane == 0) { acc = in[m.x]; if (!LOWER) ud = lu[diag[m.x]]; }
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int e = m.y + lane + u * lpr;
      if (e < m.z) { c[u] = __ldg(colind + e); v[u] = lu[e]; ps[u] = __ldg(sched_pos + c[u]) - base; }
    }
  };
  auto value_of = [&](int col, int p) -> double {   // shared memory when the column's row is computed in this run
    return (p >= 0 && p < run_rows) ? xs[p] : out[col];
  };
  fetch_meta(base, lp_next);
  fetch_entries();
  for (int l = l0; l < l1; ++l) {
    const int4 cur = m;
    const int cur_t = t_cur;
    const double cur_acc = acc, cur_ud = ud;
    int cc[PF], cp[PF]; double cv[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) { cc[u] = c[u]; cv[u] = v[u]; cp[u] = ps[u]; }
    const int nf = lp_next;
    int ne = nf;
    if (l + 1 < l1) { ne = level_ptr[l + 2]; fetch_meta(nf, ne); }
    double part = 0.0;
    if (cur.x >= 0) {
#pragma unroll
      for (int u = 0; u < PF; ++u) if (cc[u] >= 0) part += cv[u] * value_of(cc[u], cp[u]);
      for (int e = cur.y + lane + PF * lpr; e < cur.z; e += lpr) {
        const int col = colind[e];
        part += lu[e] * value_of(col, sched_pos[col] - base);
      }
    }
    for (int o = lpr >> 1; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (cur.x >= 0 && lane == 0) {
      const double a = cur_acc - part;
      const double r = LOWER ? a : a / cur_ud;
      xs[cur_t - base] = r;
      out[cur.x] = r;
    }
    if (l + 1 < l1) fetch_entries();
    lp_next = ne;
    __syncthreads();
  }
}

__global__ void k_sched_pos(const int32_t* __restrict__ order, size_t n, int32_t* __restrict__ pos) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t < n) pos[order[t]] = (int32_t)t;
}

__global__ void k_sched_meta(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ diag, const int32_t* __restrict__ order,
                             size_t n, int lower, int4* __restrict__ meta) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int i = order[t];
  meta[t] = lower ? make_int4(i, rowptr[i], diag[i], 0) : make_int4(i, diag[i] + 1, rowptr[i + 1], 0);
}

// launch plan of a level schedule: runs of consecutive levels of at most 1024 rows each go to one CTA

constexpr int kRunRows = 24 * 1024;   // 192 KB of shared memory
constexpr int kRunLevelRows = 256;    // levels up to this many rows are walked inside one CTA (>= 4 lanes per row)

static std::vector<LevelRun> plan_runs(const std::vector<int32_t>& level_ptr) {
  std::vector<LevelRun> runs;
  const int nl = (int)level_ptr.size() - 1;
  int l = 0;
  while (l < nl) {
    const int sz = level_ptr[l + 1] - level_ptr[l];
    if (sz <= kRunLevelRows) {
      int e = l + 1;
      // the rows of a run live in the shared memory of its CTA: at most kRunRows of them
      while (e < nl && level_ptr[e + 1] - level_ptr[e] <= kRunLevelRows && level_ptr[e + 1] - level_ptr[l] <= kRunRows) ++e;
      runs.push_back({l, e, true});
      l = e;
    } else {
      runs.push_back({l, l + 1, false});
      ++l;
    }
  }
  return runs;
}

// diagonal block of a row-distributed operator: entries whose (local) column is an owned row
__global__ void k_owned_flag(const int32_t* __restrict__ colind, size_t nnz, int32_t n, uint32_t* __restrict__ keep) {
  const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (e < nnz) keep[e] = (colind[e] >= 0 && colind[e] < n) ? 1u : 0u;   // halo columns: negative (lower ranks) or >= n
}

__global__ void k_block_rowptr(const int32_t* __restrict__ rowptr, const uint32_t* __restrict__ pos, size_t n, size_t nnz,
                               int32_t total, int32_t* __restrict__ out) {
  const size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r > n) return;
  out[r] = (r == n || (size_t)rowptr[r] >= nnz) ? total : (int32_t)pos[rowptr[r]];
}

__global__ void k_block_compact(const int32_t* __restrict__ colind, const double* __restrict__ val, const uint32_t* __restrict__ keep,
                                const uint32_t* __restrict__ pos, size_t nnz, int32_t* __restrict__ out_col, double* __restrict__ out_val) {
  const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (e < nnz && keep[e]) { out_col[pos[e]] = colind[e]; out_val[pos[e]] = val[e]; }
}

void ilu0_setup(tfelcu_solver* s, bool symbolic) {
  Ctx* ctx = s->ctx;
  const size_t n = s->n;
  TFELCU_REQUIRE(n < 0x7fffffffull, "ILU: too many rows");
  if (!s->ilu || s->ilu->levels != s->params.ilu_levels) symbolic = true;
  if (symbolic) s->ilu.reset(new Ilu0());
  Ilu0& I = *s->ilu;
  I.levels = s->params.ilu_levels;
  I.rowptr = s->rowptr; I.colind = s->colind; I.val = s->val; I.nnz = s->nnz;
  if (s->dist) {
    // row-distributed operator: ILU of the DIAGONAL BLOCK of this rank (block-Jacobi over the ranks with an
    // incomplete factorisation inside). Owned columns are the local indices < n and keep their ascending order.
    DevBuf<uint32_t> keep(s->nnz), pos(s->nnz);
    if (s->nnz) {
      k_owned_flag<<<grid_for(s->nnz, 256), 256, 0, ctx->stream>>>(s->colind, s->nnz, (int32_t)n, keep.p);
      ctx->count(); TFELCU_LAUNCH_CHECK();
      exclusive_sum<uint32_t, uint32_t>(ctx, keep.p, pos.p, s->nnz);
    }
    uint32_t last_pos = 0, last_keep = 0;
    if (s->nnz) { ctx->download(&last_pos, pos.p + (s->nnz - 1), 1); ctx->download(&last_keep, keep.p + (s->nnz - 1), 1); }
    I.nnz = (size_t)last_pos + last_keep;
    if (symbolic) { I.own_rowptr.alloc(n + 1); I.own_colind.alloc(I.nnz); I.own_val.alloc(I.nnz); }
    TFELCU_REQUIRE(I.own_val.n == I.nnz, "ILU: the pattern changed under a values-only refresh");
    k_block_rowptr<<<grid_for(n + 1, 256), 256, 0, ctx->stream>>>(s->rowptr, pos.p, n, s->nnz, (int32_t)I.nnz, I.own_rowptr.p);
    ctx->count(); TFELCU_LAUNCH_CHECK();
    if (s->nnz) {
      k_block_compact<<<grid_for(s->nnz, 256), 256, 0, ctx->stream>>>(s->colind, s->val, keep.p, pos.p, s->nnz,
                                                                     I.own_colind.p, I.own_val.p);
      ctx->count(); TFELCU_LAUNCH_CHECK();
    }
    ctx->sync();
    I.rowptr = I.own_rowptr.p; I.colind = I.own_colind.p; I.val = I.own_val.p;
  }
  // default: ILU(levels) without level scheduling (iluk.cu). TFELCU_ILU=levels keeps the launch-per-level ILU(0).
  const char* mode = std::getenv("TFELCU_ILU");
  I.syncfree = !(mode && std::string(mode) == "levels" && I.levels == 0);
  if (I.syncfree) { iluk_factorize(s, symbolic); return; }
  I.diag.alloc(n);
  DevBuf<int> missing(1);
  ctx->zero(missing.p, 1);
  k_diag_pos<<<grid_for(n, 256), 256, 0, ctx->stream>>>(I.rowptr, I.colind, n, I.diag.p, missing.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  int h_missing = 0; ctx->download(&h_missing, missing.p, 1);
  TFELCU_REQUIRE(!h_missing, "ILU(0): the matrix has a row without a stored diagonal entry");

  DevBuf<int> level(n);
  DevBuf<unsigned int> ticket(1);
  for (int pass = 0; pass < 2; ++pass) {
    TFELCU_CK(cudaMemsetAsync(level.p, 0xff, n * sizeof(int), ctx->stream));
    ctx->zero(ticket.p, 1);
    if (pass == 0) k_levels<true><<<grid_for(n, 128), 128, 0, ctx->stream>>>(I.rowptr, I.colind, I.diag.p, n, ticket.p, level.p, missing.p);
    else k_levels<false><<<grid_for(n, 128), 128, 0, ctx->stream>>>(I.rowptr, I.colind, I.diag.p, n, ticket.p, level.p, missing.p);
    ctx->count(); TFELCU_LAUNCH_CHECK();
    ctx->download(&h_missing, missing.p, 1);
    TFELCU_REQUIRE(!h_missing, "ILU(0): level analysis did not terminate");
    if (pass == 0) schedule(ctx, level.p, n, I.order_l, I.level_ptr_l);
    else schedule(ctx, level.p, n, I.order_u, I.level_ptr_u);
  }
  I.meta_l.alloc(4 * n); I.meta_u.alloc(4 * n);
  I.pos_l.alloc(n); I.pos_u.alloc(n);
  k_sched_pos<<<grid_for(n, 256), 256, 0, ctx->stream>>>(I.order_l.p, n, I.pos_l.p);
  k_sched_pos<<<grid_for(n, 256), 256, 0, ctx->stream>>>(I.order_u.p, n, I.pos_u.p);
  ctx->count(2);
  TFELCU_CK(cudaFuncSetAttribute(k_tri_levels_cta<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRunRows * (int)sizeof(double)));
  TFELCU_CK(cudaFuncSetAttribute(k_tri_levels_cta<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRunRows * (int)sizeof(double)));
  k_sched_meta<<<grid_for(n, 256), 256, 0, ctx->stream>>>(I.rowptr, I.diag.p, I.order_l.p, n, 1, reinterpret_cast<int4*>(I.meta_l.p));
  k_sched_meta<<<grid_for(n, 256), 256, 0, ctx->stream>>>(I.rowptr, I.diag.p, I.order_u.p, n, 0, reinterpret_cast<int4*>(I.meta_u.p));
  ctx->count(2); TFELCU_LAUNCH_CHECK();
  if (std::getenv("TFELCU_TRACE")) {
    const auto rl = plan_runs(I.level_ptr_l), ru = plan_runs(I.level_ptr_u);
    std::fprintf(stderr, "[tfelcu trace] ILU(0): n %zu nnz %zu, L levels %zu (runs %zu), U levels %zu (runs %zu)\n", n, I.nnz,
                 I.level_ptr_l.size() - 1, rl.size(), I.level_ptr_u.size() - 1, ru.size());
  }
  I.d_level_ptr_l.alloc(I.level_ptr_l.size());
  I.d_level_ptr_u.alloc(I.level_ptr_u.size());
  ctx->upload(I.d_level_ptr_l.p, I.level_ptr_l.data(), I.level_ptr_l.size());
  ctx->upload(I.d_level_ptr_u.p, I.level_ptr_u.data(), I.level_ptr_u.size());

  // numeric factorisation on a copy of the values, level by level of the L schedule
  I.lu.alloc(I.nnz);
  TFELCU_CK(cudaMemcpyAsync(I.lu.p, I.val, I.nnz * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  for (const LevelRun& r : plan_runs(I.level_ptr_l)) {
    if (r.cta) {
      k_ilu0_levels_cta<<<1, 1024, 0, ctx->stream>>>(I.rowptr, I.colind, I.diag.p, I.order_l.p, I.d_level_ptr_l.p,
                                                    r.l0, r.l1, I.lu.p);
    } else {
      const int b = I.level_ptr_l[r.l0], e = I.level_ptr_l[r.l1];
      k_ilu0_level<<<grid_for((size_t)(e - b), 128), 128, 0, ctx->stream>>>(I.rowptr, I.colind, I.diag.p, I.order_l.p, b, e, I.lu.p);
    }
    ctx->count(); TFELCU_LAUNCH_CHECK();
  }
  ctx->sync();
}

static void tri_launches(tfelcu_solver* s, const double* d_in, double* d_out, bool lower) {
  Ctx* ctx = s->ctx;
  Ilu0& I = *s->ilu;
  const std::vector<int32_t>& lp = lower ? I.level_ptr_l : I.level_ptr_u;
  const int32_t* order = lower ? I.order_l.p : I.order_u.p;
  const int32_t* d_lp = lower ? I.d_level_ptr_l.p : I.d_level_ptr_u.p;
  const KrylovScalars* S = s->scal.p;
  std::vector<LevelRun>& runs = lower ? I.runs_l : I.runs_u;
  if (runs.empty()) runs = plan_runs(lp);
  for (const LevelRun& r : runs) {
    if (r.cta) {
      const int4* meta = reinterpret_cast<const int4*>(lower ? I.meta_l.p : I.meta_u.p);
      const int32_t* pos = lower ? I.pos_l.p : I.pos_u.p;
      const size_t smem = (size_t)(lp[r.l1] - lp[r.l0]) * sizeof(double);
      int widest = 1;
      for (int l = r.l0; l < r.l1; ++l) widest = std::max(widest, lp[l + 1] - lp[l]);
      int lpr = 16;                                  // lanes per row: as many as 1024 threads allow, at most 16
      while (lpr > 1 && widest * lpr > 1024) lpr >>= 1;
      int threads = ((widest * lpr + 31) / 32) * 32;
      if (threads < 32) threads = 32;
      if (lower) k_tri_levels_cta<true><<<1, threads, smem, ctx->stream>>>(I.colind, I.diag.p, I.lu.p, meta, pos, d_lp, r.l0, r.l1, lpr, d_in, d_out, S);
      else k_tri_levels_cta<false><<<1, threads, smem, ctx->stream>>>(I.colind, I.diag.p, I.lu.p, meta, pos, d_lp, r.l0, r.l1, lpr, d_in, d_out, S);
    } else {
      const int b = lp[r.l0], e = lp[r.l1];
      const unsigned g = grid_for((size_t)(e - b), 128);
      if (lower) k_tri_level<true><<<g, 128, 0, ctx->stream>>>(I.rowptr, I.colind, I.diag.p, I.lu.p, order, b, e, d_in, d_out, S);
      else k_tri_level<false><<<g, 128, 0, ctx->stream>>>(I.rowptr, I.colind, I.diag.p, I.lu.p, order, b, e, d_in, d_out, S);
    }
    ctx->count();
  }
  TFELCU_LAUNCH_CHECK();
}

// out = U^-1 L^-1 in. The backward solve runs in place on the forward result.
void ilu0_apply(tfelcu_solver* s, const double* d_in, double* d_out) {
  if (s->ilu->syncfree) { iluk_apply(s, d_in, d_out); return; }
  tri_launches(s, d_in, d_out, true);
  tri_launches(s, d_out, d_out, false);
}

// ------------------------------------------------------------------------------------------
// GMRES(m), modified Gram-Schmidt, left preconditioning
// ------------------------------------------------------------------------------------------

struct GmresDev {      // device scalars of one restart cycle
  double target, dtol_abs, bnorm, rnorm;
  int done;            // 0 running, 1 converged, -1 diverged, -2 breakdown (happy or not)
  int k;               // columns of the current cycle
  int its, pad;
};

__device__ __forceinline__ double blk_sum(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < kVecT / 32; ++w) t += red[w];
  __syncthreads();
  return t;
}

// Sum of per-CTA partials in a fixed order (lane-strided slices combined by a butterfly: every lane, every CTA and every
// run gets the same bits). Must be called by all threads of warp 0 of the CTA; the other warps get the value through
// shared memory in cta_sum. A serial sum by every thread costs ~10 us per kernel at 592 partials.
__device__ __forceinline__ double warp_sum_fixed(const double* __restrict__ p, int n) {
  double t = 0.0;
  for (int i = threadIdx.x & 31; i < n; i += 32) t += p[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

__device__ __forceinline__ double cta_sum(const double* __restrict__ p, int n, double* sh) {
  if (threadIdx.x < 32) {
    const double t = warp_sum_fixed(p, n);
    if (threadIdx.x == 0) sh[0] = t;
  }
  __syncthreads();
  const double v = sh[0];
  __syncthreads();
  return v;
}

// partial sums of a . b
__global__ void __launch_bounds__(kVecT)
k_dot(const double* __restrict__ a, const double* __restrict__ b, size_t n, double* __restrict__ partial) {
  __shared__ double red[kVecT / 32];
  double v = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) v += a[i] * b[i];
  const double t = blk_sum(v, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// t = b - q  (q = A x)
__global__ void k_sub(const double* __restrict__ b, const double* __restrict__ q, size_t n, double* __restrict__ t) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) t[i] = b[i] - q[i];
}

__global__ void k_jacobi_apply(const double* __restrict__ dinv, const double* __restrict__ in, size_t n, double* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = dinv ? dinv[i] * in[i] : in[i];
}

// ||M^-1 b|| -> tolerances
__global__ void k_gmres_init(const double* __restrict__ partial, int np, double rtol, double atol, double dtol,
                             GmresDev* __restrict__ G) {
  if (blockIdx.x || threadIdx.x >= 32) return;
  const double bn = sqrt(warp_sum_fixed(partial, np));
  if (threadIdx.x) return;
  G->bnorm = bn; G->target = fmax(rtol * bn, atol); G->dtol_abs = dtol * bn; G->rnorm = bn;
  G->done = 0; G->k = 0; G->its = 0;
}

// start of a cycle: rnorm = ||v0||; test; g = rnorm e_0
__global__ void k_gmres_cycle_start(const double* __restrict__ partial, int np, GmresDev* __restrict__ G,
                                    double* __restrict__ g, int m) {
  if (blockIdx.x || threadIdx.x >= 32) return;
  const double rn = sqrt(warp_sum_fixed(partial, np));
  if (threadIdx.x) return;
  G->rnorm = rn; G->k = 0;
  if (!(rn == rn)) G->done = -2;
  else if (rn <= G->target) G->done = 1;
  else if (rn > G->dtol_abs) G->done = -1;
  for (int i = 0; i <= m; ++i) g[i] = 0.0;
  g[0] = rn;
}

// v *= 1 / sqrt(sum partial)   (normalisation by a norm whose square is in partial sums)
__global__ void __launch_bounds__(kVecT)
k_scale_by_norm(double* __restrict__ v, size_t n, const double* __restrict__ partial, int np, const GmresDev* __restrict__ G) {
  __shared__ double bc[1];
  if (G->done) return;
  const double nr = sqrt(cta_sum(partial, np, bc));
  if (nr == 0.0) return;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) v[i] /= nr;
}

// One modified Gram-Schmidt step, fused: apply the previous projection (h_prev = sum of partial_prev, w -= h_prev v_prev)
// and start the next dot product (partial_next of w . v_next; v_next == w gives the final norm).
// Every CTA owns a fixed slice of the vectors, so it only needs its own updated entries.
__global__ void __launch_bounds__(kVecT)
k_mgs_step(double* __restrict__ w, const double* __restrict__ v_prev, const double* __restrict__ partial_prev, int np,
           double* __restrict__ h_out, const double* __restrict__ v_next, int next_is_w, size_t n,
           double* __restrict__ partial_next, const GmresDev* __restrict__ G) {
  __shared__ double red[kVecT / 32];
  __shared__ double bc[1];
  if (G->done) return;
  double h = 0.0;
  if (v_prev) {
    h = cta_sum(partial_prev, np, bc);
    if (blockIdx.x == 0 && threadIdx.x == 0) *h_out = h;
  }
  double acc = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double wi = w[i];
    if (v_prev) { wi -= h * v_prev[i]; w[i] = wi; }
    acc += wi * (next_is_w ? wi : v_next[i]);
  }
  const double t = blk_sum(acc, red);
  if (threadIdx.x == 0) partial_next[blockIdx.x] = t;
}

// The Hessenberg matrix is stored packed: column k holds its k + 2 entries at k (k + 3) / 2 (restart 1000: 4 MB instead of
// 8; restart = maxits = 20000: 1.6 GB instead of 3.2)
__host__ __device__ __forceinline__ size_t hess_col(size_t k) { return k * (k + 3) / 2; }

// Givens rotations of column k of the Hessenberg matrix + residual estimate + stopping test (one thread).
// H is stored column-major with leading dimension m + 1.
__global__ void k_gmres_givens(double* __restrict__ H, double* __restrict__ cs, double* __restrict__ sn, double* __restrict__ g,
                               const double* __restrict__ partial_norm, int np, int k, int m, uint32_t maxits,
                               GmresDev* __restrict__ G) {
  if (blockIdx.x || threadIdx.x >= 32 || G->done) return;
  double* Hk = H + hess_col((size_t)k);
  const double hn = sqrt(warp_sum_fixed(partial_norm, np));
  if (threadIdx.x) return;
  Hk[k + 1] = hn;
  for (int j = 0; j < k; ++j) {
    const double a = Hk[j], b = Hk[j + 1];
    Hk[j] = cs[j] * a + sn[j] * b;
    Hk[j + 1] = -sn[j] * a + cs[j] * b;
  }
  const double a = Hk[k], b = Hk[k + 1];
  const double rr = hypot(a, b);
  cs[k] = rr == 0.0 ? 1.0 : a / rr; sn[k] = rr == 0.0 ? 0.0 : b / rr;
  Hk[k] = rr; Hk[k + 1] = 0.0;
  g[k + 1] = -sn[k] * g[k]; g[k] = cs[k] * g[k];
  G->its += 1; G->k = k + 1;
  const double rn = fabs(g[k + 1]);
  G->rnorm = rn;
  if (!(rn == rn)) G->done = -2;
  else if (rn <= G->target) G->done = 1;
  else if (rn > G->dtol_abs) G->done = -1;
  else if (hn == 0.0) G->done = -2;
  else if ((uint32_t)G->its >= maxits) G->done = 2;   // out of iterations
}

// y = R^-1 g (back substitution, one CTA), k columns
__global__ void __launch_bounds__(256)
k_gmres_backsolve(const double* __restrict__ H, const double* __restrict__ g, int k, int m, double* __restrict__ y) {
  __shared__ double red[256 / 32];
  for (int i = k - 1; i >= 0; --i) {
    double acc = 0.0;
    for (int j = i + 1 + threadIdx.x; j < k; j += 256) acc += H[hess_col((size_t)j) + i] * y[j];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += red[w];
      y[i] = (g[i] - t) / H[hess_col((size_t)i) + i];
    }
    __syncthreads();
  }
}

// x += sum_j y_j v_j, j ascending
__global__ void k_gmres_update_x(const double* __restrict__ V, const double* __restrict__ y, int k, size_t n, size_t ld,
                                 double* __restrict__ x) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double xi = x[i];
    for (int j = 0; j < k; ++j) xi += y[j] * V[(size_t)j * ld + i];
    x[i] = xi;
  }
}

static void precond(tfelcu_solver* s, const double* d_in, double* d_out) {
  Ctx* ctx = s->ctx;
  switch (s->params.precond) {
    case TFELCU_PC_ILU0: ilu0_apply(s, d_in, d_out); break;
    case TFELCU_PC_BLOCK_JACOBI: block_jacobi_apply(s, d_in, d_out); break;
    default:
      k_jacobi_apply<<<vgrid(s), kVecT, 0, ctx->stream>>>(s->params.precond == TFELCU_PC_JACOBI ? s->dinv.p : nullptr, d_in, s->n, d_out);
      ctx->count(); TFELCU_LAUNCH_CHECK();
  }
}

// Row-distributed operators (one process per GPU): every basis vector carries its halo entries behind the owned ones
// (it is the input of the next product), a halo exchange precedes each product, and every dot product / norm is reduced
// over the ranks -- in rank order, so that all ranks hold the same bits and take the same branch of every test.
void solve_gmres(tfelcu_solver* s, tfelcu_solver_report* rep) {
  Ctx* ctx = s->ctx;
  const bool dist = (bool)s->dist;
  const size_t n = s->n;
  const size_t n_halo = dist ? s->dist->n_halo : 0;
  const size_t front = dist ? s->dist->front_pad : 0;    // every basis vector has its lower halo in front of it
  const int vg = vgrid(s);
  const uint64_t launches0 = ctx->launches;
  uint64_t n_spmv = 0;
  uint32_t maxits = s->params.maxits;
  // restart length: the dictionary's, bounded by maxits and by the memory the Krylov basis may take
  long long m = s->params.restart;
  if (m > (long long)maxits) m = maxits;
  if (m < 1) m = 1;
  size_t free_b = 0, total_b = 0;
  TFELCU_CK(cudaMemGetInfo(&free_b, &total_b));
  const size_t ld = (front + n + n_halo + 15) & ~(size_t)15;
  if (!s->basis.n || s->basis.n < (size_t)(m + 1) * ld) {
    const size_t have = s->basis.bytes();
    long long fit = (long long)((free_b + have) * 0.8 / (8.0 * (double)(ld ? ld : 1))) - 1;
    if (dist) {   // every rank must run the same number of Arnoldi steps per cycle: agree on the smallest fit
      DevBuf<long long> f(1);
      ctx->upload(f.p, &fit, 1);
      TFELCU_NCCL(ncclAllReduce(f.p, f.p, 1, ncclInt64, ncclMin, ctx->comm, ctx->stream));
      ctx->download(&fit, f.p, 1);
    }
    if (m > fit) m = fit;
    TFELCU_REQUIRE(m >= 1, "GMRES: not enough device memory for a Krylov basis");
    s->basis.release();
    s->basis.alloc((size_t)(m + 1) * ld, front);
    s->gmres_m = (int)m;
  } else if (s->gmres_m > 0 && m > s->gmres_m) m = s->gmres_m;
  const int M = (int)m;
  s->restart_used = (uint32_t)M;
  const long long wanted = std::min<long long>((long long)s->params.restart, (long long)maxits);
  if ((long long)M < wanted && ctx->rank == 0)
    std::fprintf(stderr, "[tfelcu] warning: GMRES restart %lld does not fit in device memory, using restart %d\n", wanted, M);
  TFELCU_REQUIRE((long long)M * 4 >= wanted || M >= 30, "GMRES: the Krylov basis that fits in device memory is far shorter than the requested restart");
  s->hess.alloc(hess_col((size_t)M) + 4 * (size_t)(M + 1));
  double* H = s->hess.p;
  double* cs = H + hess_col((size_t)M);
  double* sn = cs + (M + 1);
  double* g = sn + (M + 1);
  double* y = g + (M + 1);
  DevBuf<GmresDev> Gd(1);
  GmresDev* G = Gd.p;
  double* V = s->basis.p;
  double* t = s->r.p;          // work vector
  double* P[2] = {s->partial.p, s->partial.p + s->n_partial};   // two alternating partial-sum buffers
  // what the consumers of a sum read: the per-CTA partials themselves, or the one value reduced over the ranks
  const double* R[2] = {P[0], P[1]};
  int RN[2] = {vg, vg};
  auto reduced = [&](int which) {
    if (!dist) return;
    const double* ptrs[1] = {P[which]}; const int ns[1] = {vg};
    R[which] = dist_reduce(s, ptrs, ns, 1, 5 + which);
    RN[which] = 1;
  };
  auto product = [&](double* v_in, double* out) {   // out = A v_in
    if (dist) dist_halo_exchange(s, v_in, nullptr);
    spmv_launch(s, v_in, out, nullptr, nullptr); ++n_spmv;
  };

  ctx->zero(s->x.p - front, front + n + n_halo);             // x0 = 0 (solver.cpp:165-166)
  ctx->zero(s->scal.p, 1);
  precond(s, s->b.p, s->q.p);
  k_dot<<<vg, kVecT, 0, ctx->stream>>>(s->q.p, s->q.p, n, P[0]);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  reduced(0);
  k_gmres_init<<<1, 32, 0, ctx->stream>>>(R[0], RN[0], s->params.rtol, s->params.atol, s->params.dtol, G);
  ctx->count(); TFELCU_LAUNCH_CHECK();

  GmresDev h;
  std::memset(&h, 0, sizeof(h));
  bool first_cycle = true;
  const int poll = s->params.check_every > 0 ? s->params.check_every : 8;
  while (true) {
    // v0 = M^-1 (b - A x)
    if (first_cycle) {
      TFELCU_CK(cudaMemcpyAsync(V, s->q.p, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
      product(s->x.p, s->q.p);
      k_sub<<<vg, kVecT, 0, ctx->stream>>>(s->b.p, s->q.p, n, t);
      ctx->count(); TFELCU_LAUNCH_CHECK();
      precond(s, t, V);
    }
    first_cycle = false;
    k_dot<<<vg, kVecT, 0, ctx->stream>>>(V, V, n, P[0]);
    ctx->count(); TFELCU_LAUNCH_CHECK();
    reduced(0);
    k_gmres_cycle_start<<<1, 32, 0, ctx->stream>>>(R[0], RN[0], G, g, M);
    ctx->count(); TFELCU_LAUNCH_CHECK();
    ctx->download(&h, G, 1);
    if (h.done) break;
    k_scale_by_norm<<<vg, kVecT, 0, ctx->stream>>>(V, n, R[0], RN[0], G);
    ctx->count(); TFELCU_LAUNCH_CHECK();
    int k = 0;
    for (; k < M; ++k) {
      double* w = V + (size_t)(k + 1) * ld;
      product(V + (size_t)k * ld, t);
      precond(s, t, w);
      double* Hk = H + hess_col((size_t)k);
      // MGS: k + 2 fused steps; step j projects out v_{j-1} and starts the product with v_j (the last: with w)
      for (int j = 0; j <= k + 1; ++j) {
        const double* v_prev = j > 0 ? V + (size_t)(j - 1) * ld : nullptr;
        const bool last = (j == k + 1);
        const int prev = (j & 1) ? 0 : 1, next = (j & 1) ? 1 : 0;
        k_mgs_step<<<vg, kVecT, 0, ctx->stream>>>(w, v_prev, R[prev], RN[prev], j > 0 ? Hk + (j - 1) : nullptr,
                                                 last ? w : V + (size_t)j * ld, last ? 1 : 0, n, P[next], G);
        ctx->count();
        reduced(next);
      }
      TFELCU_LAUNCH_CHECK();
      const int fin = ((k + 1) & 1) ? 1 : 0;
      k_gmres_givens<<<1, 32, 0, ctx->stream>>>(H, cs, sn, g, R[fin], RN[fin], k, M, maxits, G);
      ctx->count(); TFELCU_LAUNCH_CHECK();
      k_scale_by_norm<<<vg, kVecT, 0, ctx->stream>>>(w, n, R[fin], RN[fin], G);   // skipped on device once done
      ctx->count(); TFELCU_LAUNCH_CHECK();
      // the stopping test lives on the device (every kernel returns at once when G->done is set): the host only polls
      // every few Arnoldi steps, so that launches stay queued ahead of the GPU
      if ((k + 1) % poll == 0 || k + 1 == M) {
        ctx->download(&h, G, 1);
        if (h.done) { ++k; break; }
      }
    }
    ctx->download(&h, G, 1);
    const int cols = h.k;
    if (cols > 0) {
      k_gmres_backsolve<<<1, 256, 0, ctx->stream>>>(H, g, cols, M, y);
      ctx->count(); TFELCU_LAUNCH_CHECK();
      k_gmres_update_x<<<vg, kVecT, 0, ctx->stream>>>(V, y, cols, n, ld, s->x.p);
      ctx->count(); TFELCU_LAUNCH_CHECK();
    }
    if (h.done) break;
    if ((uint32_t)h.its >= maxits) break;
  }
  if (s->params.precond == TFELCU_PC_ILU) iluk_check(s);
  rep->its = (uint32_t)h.its;
  rep->converged = h.done == 1 ? 1 : (h.done == 2 || h.done == 0 ? 0 : h.done);
  rep->rnorm = h.rnorm;
  rep->bnorm = h.bnorm;
  rep->n_spmv = n_spmv;
  rep->n_launches = ctx->launches - launches0;
}

}  // namespace tfelcu
