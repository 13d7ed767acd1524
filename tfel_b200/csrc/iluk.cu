// ILU(k) in natural ordering -- what PETSc does for the reference behind solver::petsc::gmres_ilu
// (src/core/solver.hpp:162-163: PCSetType(pc, PCILU); PCFactorSetLevels(pc, ilufill)), entirely on the device and
// without a launch per dependency level.
//
//  symbolic   level-of-fill pattern with the sum rule lev(i, j) = min_k lev(i, k) + lev(k, j) + 1, entries above
//             `levels` dropped. The sequential row merge is replaced by one independent search per row (incomplete
//             fill-path theorem, Hysom & Pothen): (i, j) is an entry iff the graph of A has a path i -> ... -> j of at
//             most levels + 1 edges whose intermediate vertices are all smaller than min(i, j). One warp per row runs
//             `levels` rounds of a label-correcting search that only expands through vertices w < i and carries, for
//             every vertex reached, the smallest achievable "largest intermediate vertex" B(v): j > i is an entry when
//             it is reached at all, j < i when B(j) < j. Visited set = open-addressing hash table in shared memory.
//             Two passes (count, then write the columns of every row sorted ascending); no transposed pattern, no atomics
//             on global memory, result independent of scheduling (bit-identical to the sequential merge).
//  numeric    IKJ elimination, one warp per row, rows started in ascending order (ticket); a row waits for the rows of
//             its L part on a per-row flag (release / acquire). The operations on every entry happen in ascending k:
//             the factors are bitwise those of the sequential algorithm.
//  solves     L and U sweeps without level scheduling: G lanes per row, rows started in dependency order (ticket), the
//             output vector is pre-filled with a NaN sentinel and a lane simply re-reads x[col] until the value is
//             there. One launch per sweep whatever the depth of the dependency graph.
//
// All waits are bounded: a wait that exceeds the limit raises a flag that the host turns into an error instead of
// hanging the GPU.
#include "solver.cuh"
#include "sortutil.cuh"

#include <algorithm>
#include <cstdlib>

namespace tfelcu {

namespace {

constexpr unsigned long long kSentinel = 0xffffffffffffffffull;   // a NaN no arithmetic produces; memset(0xff)
constexpr long long kWaitLimit = 6000000000ll;                    // clock64 ticks (~3 s): never expected

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const void* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_f64(double* p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_i32(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_i32(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ------------------------------------------------------------------------------------------
// symbolic phase
// ------------------------------------------------------------------------------------------

constexpr int kSymWarps = 4;        // warps (= rows in flight) per CTA
constexpr int kIntMax = 0x7fffffff;

struct SymTable {
  int* keys;      // [cap] vertex, -1 = empty
  int* cur;       // [cap] B(v) after the previous round (kIntMax: not reached yet)
  int* nw;        // [cap] B(v) being built by the current round; later the list of selected columns
  unsigned* chg;  // [cap / 32] slots whose label improved in the previous round
  int* count;     // entries in the table
};

__device__ __forceinline__ unsigned sym_hash(int v, int shift) { return ((unsigned)v * 2654435761u) >> shift; }

// insert v with label lab (min-combine); false when the table is full
__device__ __forceinline__ bool sym_insert(const SymTable& T, int cap, int shift, int v, int lab) {
  unsigned h = sym_hash(v, shift);
  for (int probe = 0; probe < cap; ++probe) {
    const int k = ((volatile int*)T.keys)[h];
    if (k == v) { atomicMin(&T.nw[h], lab); return true; }
    if (k == -1) {
      const int old = atomicCAS(&T.keys[h], -1, v);
      if (old == -1) { atomicAdd(T.count, 1); atomicMin(&T.nw[h], lab); return true; }
      if (old == v) { atomicMin(&T.nw[h], lab); return true; }
    }
    h = (h + 1) & (unsigned)(cap - 1);
  }
  return false;
}

// nw -> cur for every slot; remembers which slots improved
__device__ __forceinline__ void sym_commit(const SymTable& T, int cap, int lane) {
  for (int s0 = 0; s0 < cap; s0 += 32) {
    const int s = s0 + lane;
    const int a = T.nw[s], b = T.cur[s];
    const bool better = a < b;
    if (better) T.cur[s] = a;
    const unsigned m = __ballot_sync(0xffffffffu, better);
    if (lane == 0) T.chg[s0 >> 5] = m;
  }
  __syncwarp();
}

// PASS 0: rowlen[i] = entries of row i of the ILU(levels) pattern. PASS 1: columns written ascending at fptr[i],
// position of the diagonal in fdiag[i].
template <int PASS>
__global__ void __launch_bounds__(kSymWarps * 32)
k_iluk_symbolic(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind, int n, int levels, int cap, int shift,
                long long* __restrict__ rowlen, const long long* __restrict__ fptr, int32_t* __restrict__ fcol,
                long long* __restrict__ fdiag, int* __restrict__ overflow) {
  extern __shared__ int sym_sh[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp = 3 * cap + cap / 32 + 32;
  int* base = sym_sh + warp * per_warp;
  SymTable T;
  T.keys = base; T.cur = base + cap; T.nw = base + 2 * cap;
  T.chg = reinterpret_cast<unsigned*>(base + 3 * cap);
  T.count = base + 3 * cap + cap / 32;
  const int limit = cap - cap / 4;
  for (long long row = (long long)blockIdx.x * kSymWarps + warp; row < n; row += (long long)gridDim.x * kSymWarps) {
    const int i = (int)row;
    for (int s = lane; s < cap; s += 32) { T.keys[s] = -1; T.cur[s] = kIntMax; T.nw[s] = kIntMax; }
    if (lane == 0) *T.count = 0;
    __syncwarp();
    bool full = false;
    // paths of one edge: no intermediate vertex
    for (int e = rowptr[i] + lane; e < rowptr[i + 1]; e += 32) {
      const int v = colind[e];
      if (v != i && !sym_insert(T, cap, shift, v, -1)) full = true;
    }
    __syncwarp();
    sym_commit(T, cap, lane);
    for (int d = 2; d <= levels + 1; ++d) {
      if (__any_sync(0xffffffffu, full) || *(volatile int*)T.count > limit) { full = true; break; }
      // expand every vertex w < i whose label improved in the previous round; 32 of them are fetched together
      for (int s0 = 0; s0 < cap; s0 += 32) {
        const unsigned word = T.chg[s0 >> 5];
        if (!word) continue;
        int w = -1, lab = 0, b = 0, e = 0;
        if ((word >> lane) & 1u) {
          w = T.keys[s0 + lane];
          if (w < i) { lab = max(T.cur[s0 + lane], w); b = rowptr[w]; e = rowptr[w + 1]; } else w = -1;
        }
        unsigned todo = __ballot_sync(0xffffffffu, w >= 0);
        while (todo) {
          const int src = __ffs(todo) - 1;
          todo &= todo - 1;
          const int bb = __shfl_sync(0xffffffffu, b, src), ee = __shfl_sync(0xffffffffu, e, src);
          const int ll = __shfl_sync(0xffffffffu, lab, src);
          for (int p = bb + lane; p < ee; p += 32) {
            const int v = colind[p];
            if (v != i && !sym_insert(T, cap, shift, v, ll)) full = true;
          }
        }
      }
      __syncwarp();
      sym_commit(T, cap, lane);
    }
    if (__any_sync(0xffffffffu, full) || *(volatile int*)T.count > limit) {
      if (lane == 0) atomicExch(overflow, 1);
      if (PASS == 0 && lane == 0) rowlen[i] = 1;
      __syncwarp();
      continue;
    }
    // entries of the row: the diagonal, every j > i reached, every j < i reached with all intermediates below j
    if (PASS == 0) {
      int cnt = 0;
      for (int s = lane; s < cap; s += 32) {
        const int v = T.keys[s];
        if (v >= 0 && (v > i || T.cur[s] < v)) ++cnt;
      }
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      if (lane == 0) rowlen[i] = cnt + 1;
    } else {
      int m = 0;                                     // selected so far (uniform)
      for (int s0 = 0; s0 < cap; s0 += 32) {
        const int s = s0 + lane;
        const int v = T.keys[s];
        const bool sel = v >= 0 && (v > i || T.cur[s] < v);
        const unsigned mask = __ballot_sync(0xffffffffu, sel);
        // nw[] of slots below s0 + 32 is no longer needed: the list grows behind the scan (m <= s0 + 32 always)
        __syncwarp();
        if (sel) T.nw[m + __popc(mask & ((1u << lane) - 1u))] = v;
        m += __popc(mask);
      }
      if (lane == 0) T.nw[m] = i;
      m += 1;
      int P = 32;
      while (P < m) P <<= 1;
      __syncwarp();
      for (int s = m + lane; s < P; s += 32) T.nw[s] = kIntMax;
      __syncwarp();
      // bitonic sort of nw[0, P)
      for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int t = lane; t < P; t += 32) {
            const int u = t ^ j;
            if (u > t) {
              const int a = T.nw[t], c = T.nw[u];
              const bool up = (t & k) == 0;
              if ((a > c) == up) { T.nw[t] = c; T.nw[u] = a; }
            }
          }
          __syncwarp();
        }
      }
      const long long out = fptr[i];
      for (int t = lane; t < m; t += 32) {
        const int v = T.nw[t];
        fcol[out + t] = v;
        if (v == i) fdiag[i] = out + t;
      }
      __syncwarp();
    }
  }
}

__global__ void k_widen_rowptr(const int32_t* __restrict__ rowptr, size_t n1, long long* __restrict__ out) {
  const size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r < n1) out[r] = rowptr[r];
}

__global__ void k_diag_pos64(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind, size_t n,
                             long long* __restrict__ diag, int* __restrict__ missing) {
  const size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  int lo = rowptr[r], hi = rowptr[r + 1];
  const int end = hi;
  while (lo < hi) {
    const int mid = lo + ((hi - lo) >> 1);
    if (colind[mid] < (int32_t)r) lo = mid + 1; else hi = mid;
  }
  if (lo >= end || colind[lo] != (int32_t)r) { atomicExch(missing, 1); diag[r] = -1; }
  else diag[r] = lo;
}

// ------------------------------------------------------------------------------------------
// dependency levels -> the order in which rows are started
// ------------------------------------------------------------------------------------------
// Rows started in index order keep only a window of consecutive rows in flight, and in a lexicographic FE numbering
// consecutive rows form a dependency chain (row i needs row i - 1): the window holds a few grid lines, i.e. a few
// independent chains, however many rows it spans (measured: 45 ms per sweep at 592 k rows). Started in LEVEL order
// (level(i) = 1 + max level of the rows i needs; rows of one level are independent) the window holds whole wavefronts.
// A row's dependencies have smaller levels, hence earlier tickets: they are resident or done when it waits for them.

template <bool LOWER>
__global__ void __launch_bounds__(128)
k_levels64(const long long* __restrict__ fptr, const int32_t* __restrict__ fcol, const long long* __restrict__ fdiag,
           size_t n, unsigned int* __restrict__ ticket, int* level, int* __restrict__ stuck) {
  __shared__ unsigned int blk;
  if (threadIdx.x == 0) blk = atomicAdd(ticket, 1u);
  __syncthreads();
  const size_t t = (size_t)blk * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const size_t r = LOWER ? t : n - 1 - t;
  const long long e0 = LOWER ? fptr[r] : fdiag[r] + 1;
  const long long e1 = LOWER ? fdiag[r] : fptr[r + 1];
  int lv = 0;
  for (long long e = e0; e < e1; ++e) {
    const int c = fcol[e];
    int l;
    const long long t0 = clock64();
    while ((l = ld_acquire_i32(level + c)) < 0) {
      if (clock64() - t0 > kWaitLimit) { atomicExch(stuck, 1); l = 0; break; }
      __nanosleep(20);
    }
    lv = l + 1 > lv ? l + 1 : lv;
  }
  st_release_i32(level + r, lv);
}

__global__ void k_max_level(const int* __restrict__ a, size_t n, int* __restrict__ out) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) atomicMax(out, a[i]);
}

// rows sorted by level (ascending row index inside a level); returns the number of levels
template <bool LOWER>
int level_order(Ctx* ctx, Ilu0& I, size_t n, DevBuf<int32_t>& order) {
  DevBuf<int> level(n), mx(1);
  TFELCU_CK(cudaMemsetAsync(level.p, 0xff, n * sizeof(int), ctx->stream));
  ctx->zero(I.ticket.p, 1);
  ctx->zero(mx.p, 1);
  k_levels64<LOWER><<<grid_for(n, 128), 128, 0, ctx->stream>>>(I.fptr.p, I.fcol, I.fdiag.p, n, I.ticket.p, level.p, I.stuck.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  k_max_level<<<grid_for(n, 256), 256, 0, ctx->stream>>>(level.p, n, mx.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  int h_stuck = 0, n_levels = 0;
  ctx->download(&h_stuck, I.stuck.p, 1);
  TFELCU_REQUIRE(!h_stuck, "ILU: level analysis did not terminate");
  ctx->download(&n_levels, mx.p, 1);
  n_levels += 1;
  DevBuf<uint32_t> ids(n), keys_sorted(n);
  iota_u32(ctx, ids.p, n);
  order.alloc(n);
  sort_pairs<uint32_t, uint32_t>(ctx, reinterpret_cast<const uint32_t*>(level.p), keys_sorted.p, ids.p,
                                 reinterpret_cast<uint32_t*>(order.p), n, 0, bits_for((size_t)n_levels + 1));
  return n_levels;
}

// ------------------------------------------------------------------------------------------
// numeric phase
// ------------------------------------------------------------------------------------------

constexpr int kNumWarps = 8;
constexpr int kNumRowCap = 448;     // rows up to this many entries are eliminated in shared memory

// position of column c in cols[lo, hi) (ascending), or -1
__device__ __forceinline__ int find_col(const int32_t* cols, int lo, int hi, int c) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cols[mid] < c) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kNumWarps * 32)
k_iluk_numeric(const int32_t* __restrict__ a_rowptr, const int32_t* __restrict__ a_col, const double* __restrict__ a_val,
               const long long* __restrict__ fptr, const int32_t* __restrict__ fcol, const long long* __restrict__ fdiag,
               const int32_t* __restrict__ order, int n, double* lu, unsigned int* __restrict__ ticket, int* rowdone,
               int* __restrict__ status) {
  __shared__ int32_t sh_col[kNumWarps][kNumRowCap];
  __shared__ double sh_w[kNumWarps][kNumRowCap];
  __shared__ unsigned int blk;
  if (threadIdx.x == 0) blk = atomicAdd(ticket, 1u);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long pos = (long long)blk * kNumWarps + warp;
  if (pos >= n) return;
  const int i = order[pos];
  const long long rs = fptr[i];
  const int len = (int)(fptr[i + 1] - rs);
  const int dpos = (int)(fdiag[i] - rs);
  const bool in_shared = len <= kNumRowCap;
  const int32_t* cols;
  volatile double* w;
  if (in_shared) {
    for (int t = lane; t < len; t += 32) { sh_col[warp][t] = fcol[rs + t]; sh_w[warp][t] = 0.0; }
    cols = sh_col[warp]; w = sh_w[warp];
  } else {
    for (int t = lane; t < len; t += 32) lu[rs + t] = 0.0;
    cols = fcol + rs; w = lu + rs;
  }
  __syncwarp();
  // the row of A (a subset of the pattern)
  for (int e = a_rowptr[i] + lane; e < a_rowptr[i + 1]; e += 32) {
    const int c = a_col[e];
    const int p = find_col(cols, 0, len, c);
    if (p < len && cols[p] == c) w[p] = a_val[e];
  }
  __syncwarp();
  for (int e = 0; e < dpos; ++e) {
    const int k = cols[e];
    const long long t0 = clock64();
    while (ld_acquire_i32(rowdone + k) == 0) {
      if (clock64() - t0 > kWaitLimit) { atomicExch(status, -1); break; }
      __nanosleep(40);
    }
    const long long kd = fdiag[k], ke = fptr[k + 1];
    const double l = w[e] / __ldcg(lu + kd);
    __syncwarp();
    if (lane == 0) w[e] = l;
    if (l == 0.0) continue;
    for (long long ek = kd + 1 + lane; ek < ke; ek += 32) {
      const int c = __ldg(fcol + ek);
      const int p = find_col(cols, e + 1, len, c);
      if (p < len && cols[p] == c) w[p] = __dsub_rn(w[p], __dmul_rn(l, __ldcg(lu + ek)));   // no contraction: the sequential bits
    }
    __syncwarp();
  }
  __syncwarp();
  if (in_shared)
    for (int t = lane; t < len; t += 32) __stcg(lu + rs + t, w[t]);
  if (lane == 0 && w[dpos] == 0.0) atomicCAS(status, 0, i + 1);
  __threadfence();
  __syncwarp();
  if (lane == 0) st_release_i32(rowdone + i, 1);
}

// ------------------------------------------------------------------------------------------
// triangular sweeps
// ------------------------------------------------------------------------------------------

constexpr int kTriThreads = 256;

// forward (LOWER): out_i = in_i - sum_{j<i} l_ij out_j ; backward: out_i = (in_i - sum_{j>i} u_ij out_j) / u_ii.
// `out` holds the sentinel wherever the value is not known yet.
//
// A row first waits at a GATE: one lane re-reads the value of the dependency closest to the row (the last column of its
// L part / the first column of its U part: in an FE numbering the row that finishes last) with a pause between reads;
// only when that value is there do all lanes collect theirs. Without the gate every lane of every row in flight (~300 k
// threads, most of them dozens of dependency levels ahead of the front) re-reads L2 without pause and the rows on the
// critical path queue behind that traffic (measured: 17 us per dependency level).
// Entries are accumulated in a fixed order whatever the arrival order of the values: bitwise reproducible.
template <bool LOWER, int G>
__global__ void __launch_bounds__(kTriThreads)
k_tri_syncfree(const long long* __restrict__ fptr, const int32_t* __restrict__ fcol, const long long* __restrict__ fdiag,
               const double* __restrict__ lu, const int32_t* __restrict__ order, int n, const double* __restrict__ in,
               double* out, unsigned int* __restrict__ ticket, int* __restrict__ stuck, int gate_sleep_ns,
               const KrylovScalars* __restrict__ S) {
  __shared__ unsigned int blk;
  if (S && S->done) return;
  if (threadIdx.x == 0) blk = atomicAdd(ticket, 1u);
  __syncthreads();
  constexpr int kRows = kTriThreads / G;
  const int g = threadIdx.x / G, lane = threadIdx.x % G;
  const long long t = (long long)blk * kRows + g;
  if (t >= n) return;
  const int i = __ldg(order + t);
  const long long d = fdiag[i];
  const long long b = LOWER ? fptr[i] : d + 1;
  const long long e = LOWER ? d : fptr[i + 1];
  const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
  double acc = 0.0;
  double rhs = 0.0, piv = 1.0;
  if (lane == 0) { rhs = in[i]; if (!LOWER) piv = lu[d]; }
  bool bad = false;
  bool first = true;
  // the columns and factors of four entries per lane are requested before the first wait (they do not depend on the
  // solution), and so are the four values -- re-read after the gate if they were not there yet
  for (long long p = b + lane; p - lane < e; p += 4 * G) {
    int c[4]; double l[4]; unsigned long long x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long q = p + (long long)u * G;
      const bool in_row = q < e;
      c[u] = in_row ? __ldg(fcol + q) : -1;
      l[u] = in_row ? __ldg(lu + q) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) x[u] = c[u] >= 0 ? ld_relaxed_u64(out + c[u]) : 0ull;
    if (first && gate_sleep_ns >= 0) {
      if (lane == 0) {
        const int gc = __ldg(fcol + (LOWER ? e - 1 : b));
        const long long t0 = clock64();
        int spins = 0;
        while (ld_relaxed_u64(out + gc) == kSentinel) {
          if (gate_sleep_ns > 0) __nanosleep(gate_sleep_ns);
          if ((++spins & 63) == 0 && (clock64() - t0 > kWaitLimit || *(volatile int*)stuck)) { bad = true; break; }
        }
      }
      __syncwarp(gmask);
    }
    first = false;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (c[u] < 0 || x[u] != kSentinel) continue;
      const long long t0 = clock64();
      int spins = 0;
      while ((x[u] = ld_relaxed_u64(out + c[u])) == kSentinel) {
        if ((++spins & 63) == 0 && (clock64() - t0 > kWaitLimit || *(volatile int*)stuck)) { bad = true; break; }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += l[u] * __longlong_as_double((long long)x[u]);
  }
  if (bad) atomicExch(stuck, 1);
#pragma unroll
  for (int o = G >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(gmask, acc, o);
  if (lane == 0) {
    double r = rhs - acc;
    if (!LOWER) r = r / piv;
    if (__double_as_longlong(r) == (long long)kSentinel) r = __longlong_as_double(0x7ff8000000000000ll);
    st_relaxed_f64(out + i, r);
  }
}

template <bool LOWER>
void launch_tri(Ctx* ctx, Ilu0& I, int n, int G, const double* in, double* out, const KrylovScalars* S) {
  ctx->zero(I.ticket.p, 1);
  TFELCU_CK(cudaMemsetAsync(out, 0xff, (size_t)n * sizeof(double), ctx->stream));
  const long long* fptr = I.fptr.p; const int32_t* fcol = I.fcol; const long long* fdiag = I.fdiag.p; const double* lu = I.lu.p;
  unsigned int* tk = I.ticket.p; int* st = I.stuck.p;
  const int sleep_ns = I.gate_sleep_ns;
  const size_t smem = I.tri_smem;   // > 0: fewer CTAs per SM, i.e. fewer rows in flight
  const int32_t* order = LOWER ? I.order_l.p : I.order_u.p;
#define TFELCU_TRI(GG)                                                                                                      \
  k_tri_syncfree<LOWER, GG><<<grid_for((size_t)n, kTriThreads / GG), kTriThreads, smem, ctx->stream>>>(fptr, fcol, fdiag, lu, order, \
                                                                                                   n, in, out, tk, st, sleep_ns, S)
  switch (G) {
    case 4: TFELCU_TRI(4); break;
    case 8: TFELCU_TRI(8); break;
    case 16: TFELCU_TRI(16); break;
    default: TFELCU_TRI(32); break;
  }
#undef TFELCU_TRI
  ctx->count(); TFELCU_LAUNCH_CHECK();
}

template <bool LOWER, int G>
void tri_attr(size_t smem) {
  TFELCU_CK(cudaFuncSetAttribute(k_tri_syncfree<LOWER, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
}

int lanes_for(double avg_len) {
  int g = 4;
  while (g < 32 && (double)g * 1.5 < avg_len) g <<= 1;
  return g;
}

}  // namespace

// symbolic + numeric factorisation of I.rowptr / I.colind / I.val (n rows) with level-of-fill `levels`
void iluk_factorize(tfelcu_solver* s, bool symbolic) {
  Ctx* ctx = s->ctx;
  Ilu0& I = *s->ilu;
  const size_t n = s->n;
  const int levels = I.levels;
  I.t_symbolic_s = 0.0;
  if (symbolic) {
    cudaEvent_t e0, e1;
    TFELCU_CK(cudaEventCreate(&e0)); TFELCU_CK(cudaEventCreate(&e1));
    TFELCU_CK(cudaEventRecord(e0, ctx->stream));
    I.fptr.alloc(n + 1);
    I.fdiag.alloc(n);
    I.ticket.alloc(1); I.stuck.alloc(1); I.rowdone.alloc(n);
    ctx->zero(I.stuck.p, 1);
    if (levels == 0) {
      // the pattern of A itself
      k_widen_rowptr<<<grid_for(n + 1, 256), 256, 0, ctx->stream>>>(I.rowptr, n + 1, I.fptr.p);
      ctx->count(); TFELCU_LAUNCH_CHECK();
      DevBuf<int> missing(1);
      ctx->zero(missing.p, 1);
      k_diag_pos64<<<grid_for(n, 256), 256, 0, ctx->stream>>>(I.rowptr, I.colind, n, I.fdiag.p, missing.p);
      ctx->count(); TFELCU_LAUNCH_CHECK();
      int h_missing = 0; ctx->download(&h_missing, missing.p, 1);
      TFELCU_REQUIRE(!h_missing, "ILU: the matrix has a row without a stored diagonal entry");
      I.fcol = I.colind; I.fnnz = I.nnz;
      I.fcol_store.release();
    } else {
      DevBuf<int> overflow(1);
      DevBuf<long long> rowlen(n + 1);
      const int caps[3] = {1024, 4096, 16384};
      int cap = 0, shift = 0;
      size_t smem = 0;
      bool ok = false;
      for (int attempt = 0; attempt < 3 && !ok; ++attempt) {
        cap = caps[attempt];
        shift = 32; for (int c = cap; c > 1; c >>= 1) --shift;
        smem = (size_t)kSymWarps * (3 * (size_t)cap + cap / 32 + 32) * sizeof(int);
        if (smem > 200 * 1024) break;
        TFELCU_CK(cudaFuncSetAttribute(k_iluk_symbolic<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TFELCU_CK(cudaFuncSetAttribute(k_iluk_symbolic<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->zero(overflow.p, 1);
        ctx->zero(rowlen.p, n + 1);
        const unsigned grid = (unsigned)std::min<size_t>((n + kSymWarps - 1) / kSymWarps, (size_t)ctx->sm_count * 8);
        k_iluk_symbolic<0><<<grid, kSymWarps * 32, smem, ctx->stream>>>(I.rowptr, I.colind, (int)n, levels, cap, shift, rowlen.p,
                                                                       nullptr, nullptr, nullptr, overflow.p);
        ctx->count(); TFELCU_LAUNCH_CHECK();
        int h_over = 0; ctx->download(&h_over, overflow.p, 1);
        ok = !h_over;
      }
      TFELCU_REQUIRE(ok, "ILU(k): the level-of-fill neighbourhood of a row exceeds the search table (reduce ilufill)");
      exclusive_sum<long long, long long>(ctx, rowlen.p, I.fptr.p, n + 1);
      long long total = 0; ctx->download(&total, I.fptr.p + n, 1);
      I.fnnz = (size_t)total;
      I.fcol_store.alloc(I.fnnz);
      I.fcol = I.fcol_store.p;
      const unsigned grid = (unsigned)std::min<size_t>((n + kSymWarps - 1) / kSymWarps, (size_t)ctx->sm_count * 8);
      k_iluk_symbolic<1><<<grid, kSymWarps * 32, smem, ctx->stream>>>(I.rowptr, I.colind, (int)n, levels, cap, shift, nullptr,
                                                                     I.fptr.p, I.fcol_store.p, I.fdiag.p, overflow.p);
      ctx->count(); TFELCU_LAUNCH_CHECK();
      int h_over = 0; ctx->download(&h_over, overflow.p, 1);
      TFELCU_REQUIRE(!h_over, "ILU(k): symbolic phase overflowed in its second pass");
    }
    I.n_levels_l = level_order<true>(ctx, I, n, I.order_l);
    I.n_levels_u = level_order<false>(ctx, I, n, I.order_u);
    I.lu.alloc(I.fnnz);
    I.ytmp.alloc(n);
    const double avg = n ? 0.5 * (double)I.fnnz / (double)n : 1.0;
    I.lanes = lanes_for(avg);
    if (const char* g = std::getenv("TFELCU_TRI_LANES")) { const int v = std::atoi(g); if (v == 4 || v == 8 || v == 16 || v == 32) I.lanes = v; }
    I.gate_sleep_ns = 0;     // plain spinning at the gate measured 4-7 % faster than 100-400 ns of nanosleep (profiles/r02_tri_sweep.log)
    if (const char* g = std::getenv("TFELCU_TRI_GATE_NS")) I.gate_sleep_ns = std::atoi(g);   // < 0: no gate
    I.tri_smem = 0;
    if (const char* g = std::getenv("TFELCU_TRI_CTAS")) {   // CTAs per SM of the sweeps (default: as many as fit, 8)
      const int v = std::atoi(g);
      if (v >= 1 && v < 8) I.tri_smem = (size_t)(220 * 1024 / v) & ~(size_t)1023;
    }
    if (I.tri_smem) {
      tri_attr<true, 4>(I.tri_smem); tri_attr<true, 8>(I.tri_smem); tri_attr<true, 16>(I.tri_smem); tri_attr<true, 32>(I.tri_smem);
      tri_attr<false, 4>(I.tri_smem); tri_attr<false, 8>(I.tri_smem); tri_attr<false, 16>(I.tri_smem); tri_attr<false, 32>(I.tri_smem);
    }
    TFELCU_CK(cudaEventRecord(e1, ctx->stream));
    TFELCU_CK(cudaEventSynchronize(e1));
    float ms = 0.f; TFELCU_CK(cudaEventElapsedTime(&ms, e0, e1));
    I.t_symbolic_s = ms * 1e-3;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (std::getenv("TFELCU_TRACE"))
      std::fprintf(stderr, "[tfelcu trace] ILU(%d): n %zu, nnz(A) %zu, nnz(LU) %zu (x%.2f), %d lanes per row, dependency levels L %d U %d\n", levels, n, I.nnz,
                   I.fnnz, I.nnz ? (double)I.fnnz / (double)I.nnz : 0.0, I.lanes, I.n_levels_l, I.n_levels_u);
  }
  if (levels == 0) I.fcol = I.colind;   // the arrays of A may have moved since the symbolic phase (same pattern)
  // numeric
  ctx->zero(I.ticket.p, 1);
  ctx->zero(I.rowdone.p, n);
  DevBuf<int> status(1);
  ctx->zero(status.p, 1);
  k_iluk_numeric<<<grid_for(n, kNumWarps), kNumWarps * 32, 0, ctx->stream>>>(I.rowptr, I.colind, I.val, I.fptr.p, I.fcol, I.fdiag.p,
                                                                            I.order_l.p, (int)n, I.lu.p, I.ticket.p, I.rowdone.p, status.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  int h_status = 0; ctx->download(&h_status, status.p, 1);
  TFELCU_REQUIRE(h_status >= 0, "ILU: numeric factorisation did not terminate");
  I.zero_pivot_row = h_status > 0 ? h_status - 1 : -1;
}

void iluk_apply(tfelcu_solver* s, const double* d_in, double* d_out) {
  Ctx* ctx = s->ctx;
  Ilu0& I = *s->ilu;
  const KrylovScalars* S = s->scal.p;
  launch_tri<true>(ctx, I, (int)s->n, I.lanes, d_in, I.ytmp.p, S);
  launch_tri<false>(ctx, I, (int)s->n, I.lanes, I.ytmp.p, d_out, S);
}

void iluk_check(tfelcu_solver* s) {
  if (!s->ilu || !s->ilu->stuck.p) return;
  int h = 0;
  s->ctx->download(&h, s->ilu->stuck.p, 1);
  if (h) {
    s->ctx->zero(s->ilu->stuck.p, 1);
    fail("ILU: a triangular sweep waited too long for a row it depends on");
  }
}

}  // namespace tfelcu
