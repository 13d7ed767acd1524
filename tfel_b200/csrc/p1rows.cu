// Owner-computes rows for constant-coefficient forms on P1 triangles (configs[0] and [1]): the row kernel of assemble.cu
// (bilinear_form.hpp:64-109 restricted to the cells that touch a row, in ascending cell order) with
//   * the element matrix row in closed form (p1form.cuh) instead of the contraction of a pre-integrated tensor staged in
//     shared memory: 178 -> 88 instructions per (row, cell);
//   * one 8-byte record per (row, cell) -- two vertex references, three CSR slots, local index, valid bit -- in a blocked
//     ELL layout (record k of the 128 rows of a CTA is one 1 KB line; 16-byte records where the fields do not fit);
//   * per-block vertex tables: the coordinates of the ~390 distinct vertices a block refers to are fetched once into
//     shared memory (cp.async), the records index them with 9 bits;
//   * a CTA head that is one load level deep (uniform block layout + one descriptor), see k_assemble_p1_ell;
//   * a constant-coefficient linear form without derivatives on the same space riding along (linear_form.hpp:20-142:
//     b_i += |K| sum_q w_q f psi_i): the row thread has the cell volumes anyway, the second pass over the mesh disappears.
// Same stage / write-out as k_assemble_rows: every stored value is written once, coalesced; no atomics; the additions to
// an entry happen in ascending cell order (bitwise reproducible, bitwise equal to k_assemble_rows).
// configs[1] (n = 4096), A and b: 1.345 ms (round 1, two generic kernels) -> 0.494 ms; profiles/r02_ncu_p1rows_tile.md.
// Environment knobs (measurement only): TFELCU_NO_COMPACT, TFELCU_NO_VTAB, TFELCU_NO_UNIFORM, TFELCU_NO_ISO select the
// fallback layouts / kernels; TFELCU_P1STREAM=1 the streamed kernel (TFELCU_P1_STAGES, TFELCU_P1_CTAS: its ring / residency).
#include "p1form.cuh"
#include "tma.cuh"
#include "sortutil.cuh"

#include <cstdlib>

namespace tfelcu {

namespace {

constexpr int kEllRows = 128;

__global__ void k_ell_degree(const uint32_t* __restrict__ adj_ptr, size_t row_begin, size_t row_end, unsigned int* __restrict__ deg) {
  const size_t r = row_begin + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  unsigned d = 0;
  if (r < row_end) d = adj_ptr[r + 1] - adj_ptr[r];
  __shared__ unsigned int mx;
  if (threadIdx.x == 0) mx = 0;
  __syncthreads();
  atomicMax(&mx, d);
  __syncthreads();
  if (threadIdx.x == 0) deg[blockIdx.x] = mx;
}

// rec[(ptr[block] + k) * 128 + t] = record k of row block * 128 + t, w = 0xffffffff where the row has fewer cells
__global__ void __launch_bounds__(kEllRows)
k_ell_fill(const uint32_t* __restrict__ adj_ptr, size_t a_begin, size_t row_begin, size_t row_end, const uint4* __restrict__ rec,
           const unsigned int* __restrict__ ptr, const unsigned int* __restrict__ deg, uint4* __restrict__ ell) {
  const size_t r = row_begin + blockIdx.x * (size_t)kEllRows + threadIdx.x;
  const unsigned D = deg[blockIdx.x];
  uint4* out = ell + (size_t)ptr[blockIdx.x] * kEllRows + threadIdx.x;
  uint32_t a0 = 0, a1 = 0;
  if (r < row_end) { a0 = adj_ptr[r]; a1 = adj_ptr[r + 1]; }
  for (unsigned k = 0; k < D; ++k) {
    uint4 q = make_uint4(0u, 0u, 0u, 0xffffffffu);
    if (a0 + k < a1) q = rec[(size_t)(a0 + k) - a_begin];
    out[(size_t)k * kEllRows] = q;
  }
}

// compact record (8 bytes) when dof == vertex id and the numbering is banded: the row's own vertex is known, the other
// two vertices of the cell are signed offsets from it (21 bits each), slots 6 bits each:
//   bits 0-20 offset a | 21-41 offset b | 42-59 slots of (own, a, b) | 60-61 local dof i | 63 valid   (a < b: the local indices != i)
// (the slots in the order the kernel produces the three entries, so that no shift depends on i)
__device__ __host__ __forceinline__ bool compact_fits(long long da, long long db, uint32_t w) {
  const long long lim = 1ll << 20;
  return da > -lim && da < lim && db > -lim && db < lim && (w & 0xffu) < 64u && ((w >> 8) & 0xffu) < 64u && ((w >> 16) & 0xffu) < 64u;
}

__global__ void k_ell_check_compact(const uint32_t* __restrict__ adj_ptr, size_t a_begin, size_t row_begin, size_t row_end,
                                    const uint4* __restrict__ rec, int* __restrict__ bad) {
  const size_t r = row_begin + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r >= row_end) return;
  if (r > 0x7ffffff0ull) { atomicExch(bad, 1); return; }   // the kernel forms vertex indices in 32-bit signed arithmetic
  for (uint32_t a = adj_ptr[r]; a < adj_ptr[r + 1]; ++a) {
    const uint4 q = rec[(size_t)a - a_begin];
    const int i = (int)(q.w >> 24);
    const uint32_t v[3] = {q.x, q.y, q.z};
    const int ia = i == 0 ? 1 : 0, ib = i == 2 ? 1 : 2;
    if (v[i] != (uint32_t)r || !compact_fits((long long)v[ia] - (long long)r, (long long)v[ib] - (long long)r, q.w)) { atomicExch(bad, 1); return; }
  }
}

__global__ void __launch_bounds__(kEllRows)
k_ell_fill_compact(const uint32_t* __restrict__ adj_ptr, size_t a_begin, size_t row_begin, size_t row_end, const uint4* __restrict__ rec,
                   const unsigned int* __restrict__ ptr, const unsigned int* __restrict__ deg, unsigned long long* __restrict__ ell) {
  const size_t r = row_begin + blockIdx.x * (size_t)kEllRows + threadIdx.x;
  const unsigned D = deg[blockIdx.x];
  unsigned long long* out = ell + (size_t)ptr[blockIdx.x] * kEllRows + threadIdx.x;
  uint32_t a0 = 0, a1 = 0;
  if (r < row_end) { a0 = adj_ptr[r]; a1 = adj_ptr[r + 1]; }
  for (unsigned k = 0; k < D; ++k) {
    unsigned long long c = 0ull;
    if (a0 + k < a1) {
      const uint4 q = rec[(size_t)(a0 + k) - a_begin];
      const int i = (int)(q.w >> 24);
      const uint32_t v[3] = {q.x, q.y, q.z};
      const int ia = i == 0 ? 1 : 0, ib = i == 2 ? 1 : 2;
      const unsigned long long da = (unsigned long long)((long long)v[ia] - (long long)r) & 0x1fffffull;
      const unsigned long long db = (unsigned long long)((long long)v[ib] - (long long)r) & 0x1fffffull;
      c = da | (db << 21) | ((unsigned long long)((q.w >> (8 * i)) & 0x3fu) << 42) | ((unsigned long long)((q.w >> (8 * ia)) & 0x3fu) << 48) |
          ((unsigned long long)((q.w >> (8 * ib)) & 0x3fu) << 54) | ((unsigned long long)i << 60) | (1ull << 63);
    }
    out[(size_t)k * kEllRows] = c;
  }
}

// Vertex table of a block of 128 rows: the ascending list of the vertices its records refer to (the two other corners of
// every incident cell; about 390 on a structured mesh). The kernel loads the table's coordinates into shared memory once
// -- one gather per distinct vertex instead of two per (row, cell) -- and the records then hold 9-bit indices into it.
// PASS 0: count (and refuse more than kVtCap); PASS 1: write the list and rewrite the records in place.
constexpr int kVtCap = 512;          // 9-bit indices; 8 KB of coordinates per CTA

// k_ell_vertex_sets: one CTA per block sorts the vertex ids of its records (shared-memory bitonic sort over NSORT >= 2 D 128
// keys), writes the distinct ones to scratch[block * kVtCap ...] and their number to raw_cnt (padded number to vt_cnt).
// k_ell_vertex_tables: copies a block's list to its place in vt_list (padded with its last id to the allotted length) and
// rewrites the block's records in place: vertex offsets -> indices into the list.
template <int NSORT>
__global__ void __launch_bounds__(kEllRows)
k_ell_vertex_sets(size_t row_begin, size_t row_end, const unsigned int* __restrict__ ptr, const unsigned int* __restrict__ deg,
                  const unsigned long long* __restrict__ ell, unsigned int* __restrict__ raw_cnt, unsigned int* __restrict__ vt_cnt,
                  uint32_t* __restrict__ scratch, int* __restrict__ bad) {
  constexpr int kChunk = NSORT / kEllRows;
  __shared__ uint32_t key[NSORT];
  __shared__ unsigned int chunk[kEllRows];
  __shared__ unsigned int n_unique;
  const size_t r = row_begin + blockIdx.x * (size_t)kEllRows + threadIdx.x;
  const unsigned D = deg[blockIdx.x];
  const unsigned long long* rec = ell + (size_t)ptr[blockIdx.x] * kEllRows + threadIdx.x;
  if (2u * D * kEllRows > (unsigned)NSORT) { if (threadIdx.x == 0) { atomicExch(bad, 1); raw_cnt[blockIdx.x] = 0; vt_cnt[blockIdx.x] = 0; } return; }
  for (int e = threadIdx.x; e < NSORT; e += kEllRows) key[e] = 0xffffffffu;
  __syncthreads();
  const long long rl = (long long)r;
  for (unsigned k = 0; k < D; ++k) {
    const unsigned long long c = rec[(size_t)k * kEllRows];
    if (!(c >> 63) || r >= row_end) continue;
    key[(2 * k) * kEllRows + threadIdx.x] = (uint32_t)(rl + (((long long)(c << 43)) >> 43));
    key[(2 * k + 1) * kEllRows + threadIdx.x] = (uint32_t)(rl + (((long long)(c << 22)) >> 43));
  }
  __syncthreads();
  for (int k = 2; k <= NSORT; k <<= 1)                        // bitonic sort, ascending (padding last)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < NSORT; t += kEllRows) {
        const int u = t ^ j;
        if (u > t) {
          const uint32_t a = key[t], b = key[u];
          const bool up = (t & k) == 0;
          if ((a > b) == up) { key[t] = b; key[u] = a; }
        }
      }
      __syncthreads();
    }
  // distinct values: one thread per chunk of consecutive keys, ranks by a scan over the chunk counts
  auto distinct = [&](int e) { return key[e] != 0xffffffffu && (e == 0 || key[e - 1] != key[e]); };
  {
    unsigned c = 0;
    for (int e = threadIdx.x * kChunk; e < (threadIdx.x + 1) * kChunk; ++e) c += distinct(e) ? 1u : 0u;
    chunk[threadIdx.x] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned acc = 0;
    for (int t = 0; t < kEllRows; ++t) { const unsigned c = chunk[t]; chunk[t] = acc; acc += c; }
    n_unique = acc;
  }
  __syncthreads();
  const unsigned nu = n_unique;
  if (threadIdx.x == 0) {
    raw_cnt[blockIdx.x] = nu;
    vt_cnt[blockIdx.x] = (nu + 3u) & ~3u;                     // 16-byte granules (bulk copies of the streamed kernel)
    if (((nu + 3u) & ~3u) > (unsigned)kVtCap) atomicExch(bad, 1);
  }
  if (nu > (unsigned)kVtCap) return;
  uint32_t* out = scratch + (size_t)blockIdx.x * kVtCap;
  unsigned at = chunk[threadIdx.x];
  for (int e = threadIdx.x * kChunk; e < (threadIdx.x + 1) * kChunk; ++e) if (distinct(e)) out[at++] = key[e];
}

__global__ void __launch_bounds__(kEllRows)
k_ell_vertex_tables(size_t row_begin, size_t row_end, const unsigned int* __restrict__ ptr, const unsigned int* __restrict__ deg,
                    unsigned long long* __restrict__ ell, const unsigned int* __restrict__ raw_cnt, const unsigned int* __restrict__ vt_ptr,
                    const uint32_t* __restrict__ scratch, uint32_t* __restrict__ vt_list) {
  __shared__ uint32_t key[kVtCap];
  const size_t r = row_begin + blockIdx.x * (size_t)kEllRows + threadIdx.x;
  const unsigned D = deg[blockIdx.x];
  const unsigned nu = raw_cnt[blockIdx.x];
  unsigned long long* rec = ell + (size_t)ptr[blockIdx.x] * kEllRows + threadIdx.x;
  for (unsigned e = threadIdx.x; e < nu; e += kEllRows) key[e] = scratch[(size_t)blockIdx.x * kVtCap + e];
  __syncthreads();
  uint32_t* out = vt_list + vt_ptr[blockIdx.x];
  const unsigned allotted = vt_ptr[blockIdx.x + 1] - vt_ptr[blockIdx.x];     // >= nu rounded up (uniform tables: the largest one)
  for (unsigned e = threadIdx.x; e < allotted; e += kEllRows) out[e] = nu ? key[e < nu ? e : nu - 1] : 0u;   // padding: a valid id
  auto find = [&](uint32_t v) {
    unsigned lo = 0, hi = nu;
    while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (key[mid] < v) lo = mid + 1; else hi = mid; }
    return (unsigned long long)lo;
  };
  const long long rl = (long long)r;
  for (unsigned k = 0; k < D; ++k) {
    const unsigned long long c = rec[(size_t)k * kEllRows];
    if (!(c >> 63) || r >= row_end) continue;
    const uint32_t va = (uint32_t)(rl + (((long long)(c << 43)) >> 43)), vb = (uint32_t)(rl + (((long long)(c << 22)) >> 43));
    rec[(size_t)k * kEllRows] = (c & ~((1ull << 42) - 1ull)) | find(va) | (find(vb) << 21);
  }
}

struct P1Rows {
  const double* vertices;
  const uint8_t* dir_row; const int32_t* rowptr; const int32_t* colind; double* val; int init_mode;
  const void* ell; const unsigned int* ell_ptr; const unsigned int* ell_deg;
  size_t row_begin, row_end, local_off; uint32_t off_te;
  double* b;            // fused linear form (local rows), or nullptr
  int b_init;           // 1: its clear() is pending: write the entries
  const unsigned int* vt_ptr; const uint32_t* vt_list;   // LOCALV: vertex tables of the blocks
  int stage_doubles;    // LOCALV: the coordinates of the block's table follow the stage at this (even) offset
  unsigned int uni_deg, uni_vt;   // LOCALV, both > 0: record lines / table entries per block are uniform
  const int4* desc;     // LOCALV: two int4 per block: (first record line, records per row, table offset, table entries),
                        //         (first CSR entry of the block, entries of the block, -, -)
};

// 1 / x: hardware seed (rcp.approx.ftz.f64, ~20 bits) and two Newton steps: within 1 ulp of the rounded quotient
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = r * (2.0 - x * r);
  r = r * (2.0 - x * r);
  return r;
}

// row i of the element matrix times |K| from the three corners; STIFF: the form has only gradient x gradient terms
template <bool STIFF>
__device__ __forceinline__ void p1_row_from_corners(const P1Form& F, double2 p0, double2 p1, double2 p2, int i, double& vol,
                                                    double (&K)[3]) {
  const double J00 = p1.x - p0.x, J10 = p1.y - p0.y, J01 = p2.x - p0.x, J11 = p2.y - p0.y;
  const double det = J00 * J11 - J10 * J01;
  vol = 0.5 * fabs(det);
  const double rdet = fast_rcp(det);
  double jmt[2][2];
  jmt[0][0] = J11 * rdet; jmt[0][1] = -J10 * rdet; jmt[1][0] = -J01 * rdet; jmt[1][1] = J00 * rdet;
  if (STIFF) {
    const double gi0 = i == 0 ? -1.0 : (i == 1 ? 1.0 : 0.0), gi1 = i == 0 ? -1.0 : (i == 2 ? 1.0 : 0.0);
    const double t0 = jmt[0][0] * gi0 + jmt[0][1] * gi1, t1 = jmt[1][0] * gi0 + jmt[1][1] * gi1;
    const double s = F.wsum * vol;
    const double u0 = (t0 * F.C[0][0] + t1 * F.C[1][0]) * s, u1 = (t0 * F.C[0][1] + t1 * F.C[1][1]) * s;
    K[1] = u0 * jmt[0][0] + u1 * jmt[1][0];
    K[2] = u0 * jmt[0][1] + u1 * jmt[1][1];
    K[0] = -(K[1] + K[2]);
  } else {
    p1_element_row(F, jmt, i, K);
    K[0] *= vol; K[1] *= vol; K[2] *= vol;
  }
}

// SHAPE: 0 any constant-coefficient form, 1 gradient x gradient terms only, 2 isotropic (P1Form::iso; compact records)
template <bool FUSE_B, int SHAPE, bool COMPACT, bool LOCALV = false, int MINB = (SHAPE == 2 ? (LOCALV ? 10 : 8) : 1), int AHEAD = ((LOCALV && SHAPE == 2) ? 6 : 8)>
__global__ void __launch_bounds__(kEllRows, MINB)
k_assemble_p1_ell(const __grid_constant__ P1Form F, const P1Rows B) {
  static_assert(COMPACT || !LOCALV, "vertex tables go with the compact records");
  constexpr bool STIFF = SHAPE == 1;
  extern __shared__ double stage[];
  const size_t r0 = B.row_begin + (size_t)blockIdx.x * kEllRows;
  size_t r1 = r0 + kEllRows; if (r1 > B.row_end) r1 = B.row_end;
  const size_t to_local = B.local_off - B.row_begin;
  const size_t r = r0 + threadIdx.x;
  const double2* v2 = reinterpret_cast<const double2*>(B.vertices);
  // The head of a CTA's life is a chain of dependent loads (ncu, source page: half of all stall samples sat in front of
  // the first barrier, one load level at a time). With the vertex tables everything a block needs to start comes from
  // ONE 32-byte descriptor; the loads that depend on nothing (row pointer, Dirichlet flag) go out with it, the records
  // and the table ids right behind it, the table's coordinates go global -> shared without a register round trip
  // (cp.async), and the stage is cleared while all of that is in flight.
  int4 d0 = make_int4(0, 0, 0, 0), d1 = d0;
  if (LOCALV) {
    d1 = __ldg(B.desc + 2 * (size_t)blockIdx.x + 1);
    // uniform layout: where the block's records and table are follows from its index -- they are requested without
    // waiting for the descriptor, which then only carries the extent of the CSR segment
    if (B.uni_deg) d0 = make_int4((int)(blockIdx.x * B.uni_deg), (int)B.uni_deg, (int)(blockIdx.x * B.uni_vt), (int)B.uni_vt);
    else d0 = __ldg(B.desc + 2 * (size_t)blockIdx.x);
  }
  int rs = 0;
  bool dir = false;
  if (r < r1) { rs = B.rowptr[r + to_local]; dir = B.dir_row[(size_t)B.off_te + r] != 0; }
  const int seg0 = LOCALV ? d1.x : B.rowptr[r0 + to_local];
  const int seg_len = LOCALV ? d1.y : B.rowptr[r1 + to_local] - seg0;
  const unsigned D = LOCALV ? (unsigned)d0.y : B.ell_deg[blockIdx.x];
  if (!LOCALV) for (int e = threadIdx.x; e < seg_len; e += kEllRows) stage[e] = 0.0;
  const double bc0 = F.b_c[0], bc1 = F.b_c[1], bc2 = F.b_c[2];
  double bsum = 0.0;
  double* row = stage + (rs - seg0);
  if (COMPACT) {
    const unsigned long long* rec = reinterpret_cast<const unsigned long long*>(B.ell) +
                                    (size_t)(LOCALV ? (unsigned)d0.x : B.ell_ptr[blockIdx.x]) * kEllRows + threadIdx.x;
    // The record stream is the DRAM stream of this kernel: the first eight records of a row (all of them on every mesh
    // seen so far) are requested up front, 64 bytes in flight per thread (with one record ahead the loop waited on the
    // long scoreboard: 7.5 stall cycles per issue, 3.2 TB/s); the two corner gathers of cell k + 1 are in flight while
    // cell k is computed.
    // Instruction diet of the step (ncu: the kernel is issue bound, 118 SASS instructions per step before): vertex
    // indices in 32-bit arithmetic from the row's own index, the two corner pairs in alternating registers (no moves),
    // the slots in production order behind a precomputed shared-memory address of the row, the valid bit tested on
    // the high word, one linear-form weight when the three are equal.
    const int rc = (int)(r < r1 ? r : r1 - 1);                               // rows past the end read a valid vertex
    constexpr int kAhead = AHEAD;
    unsigned long long qa[kAhead];
#pragma unroll
    for (int k = 0; k < kAhead; ++k) qa[k] = (unsigned)k < D ? __ldg(rec + (size_t)k * kEllRows) : 0ull;
    const double2 pr = __ldg(v2 + rc);                                       // the row's own vertex: once per row
    auto idx_a = [rc](unsigned long long c) { return rc + (((int)((unsigned)c << 11)) >> 11); };            // sign-extended
    auto idx_b = [rc](unsigned long long c) { return rc + (((int)((unsigned)(c >> 21) << 11)) >> 11); };    // 21-bit fields
    const unsigned row_addr = (unsigned)__cvta_generic_to_shared(row);
    auto add_slot = [row_addr](unsigned slot, double v) {
      const unsigned addr = row_addr + slot * 8u;
      double t;
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(t) : "r"(addr));
      t += v;
      asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(t) : "memory");
    };
    const bool b_uni = F.b_uniform != 0;
    // the diagonal entry takes a contribution from every cell of the row: summed in a register (same order, same
    // rounding as in place) and stored once
    double own_sum = 0.0; unsigned own_slot = 0xffffffffu;
    double2 xa = make_double2(0.0, 0.0), xb = xa, ya = xa, yb = xa;
    const double2* sv = reinterpret_cast<const double2*>(stage + B.stage_doubles);
    if (LOCALV) {
      // the coordinates of every vertex the block's records refer to: one coalesced read of the table, one gather each
      double2* svw = reinterpret_cast<double2*>(stage + B.stage_doubles);
      const unsigned nv = (unsigned)d0.w;
      const uint32_t* ids = B.vt_list + (unsigned)d0.z;
      uint32_t id[kVtCap / kEllRows];
#pragma unroll
      for (int u = 0; u < kVtCap / kEllRows; ++u) { const unsigned t = threadIdx.x + u * kEllRows; id[u] = t < nv ? __ldg(ids + t) : 0u; }
      for (int e = threadIdx.x; e < seg_len; e += kEllRows) stage[e] = 0.0;       // while records and ids are in flight
#pragma unroll
      for (int u = 0; u < kVtCap / kEllRows; ++u) { const unsigned t = threadIdx.x + u * kEllRows; if (t < nv) cp_async_16(svw + t, v2 + id[u]); }
      asm volatile("cp.async.wait_all;" ::: "memory");
    } else {
      xa = __ldg(v2 + idx_a(qa[0])); xb = __ldg(v2 + idx_b(qa[0]));
    }
    __syncthreads();
    // pa, pb: the corners of this cell (requested one step ago); na, nb: where the next cell's corners are requested
    auto step = [&](unsigned long long cur, unsigned long long nxt, const double2& qa_, const double2& qb_, double2& na, double2& nb) {
      if (!LOCALV) { na = __ldg(v2 + idx_a(nxt)); nb = __ldg(v2 + idx_b(nxt)); }   // padding records: offsets 0
      const unsigned hi = (unsigned)(cur >> 32);
      if ((int)hi >= 0) return;
      const double2 pa = LOCALV ? sv[(unsigned)cur & 0x1ffu] : qa_, pb = LOCALV ? sv[(unsigned)(cur >> 21) & 0x1ffu] : qb_;
      const int i = (int)((hi >> 28) & 3u);
      const unsigned s_own = (hi >> 10) & 0x3fu, s_a = (hi >> 16) & 0x3fu, s_b = (hi >> 22) & 0x3fu;
      if (SHAPE == 2) {
        // edges opposite to the row's vertex, to a and to b; K_ij = (c W / 4|K|) e_i . e_j + c00 |K| Mhat_ij
        const double ox = pb.x - pa.x, oy = pb.y - pa.y, ax = pr.x - pb.x, ay = pr.y - pb.y, bx = pa.x - pr.x, by = pa.y - pr.y;
        const double adet = fabs(bx * ay - by * ax);
        const double vol = 0.5 * adet;
        if (FUSE_B) bsum += (b_uni ? bc0 : (i == 0 ? bc0 : (i == 1 ? bc1 : bc2))) * vol;
        if (dir) return;
        const double sk = F.iso_k * fast_rcp(adet);
        double k_own = sk * (ox * ox + oy * oy), k_a = sk * (ox * ax + oy * ay), k_b = sk * (ox * bx + oy * by);
        if (F.has00) { k_own += F.iso_md * vol; k_a += F.iso_mo * vol; k_b += F.iso_mo * vol; }
        own_sum += k_own; own_slot = s_own; add_slot(s_a, k_a); add_slot(s_b, k_b);
        return;
      }
      // corners in local order: i = 0: (r, a, b), 1: (a, r, b), 2: (a, b, r)
      const double2 p0 = i == 0 ? pr : pa, p1 = i == 0 ? pa : (i == 1 ? pr : pb), p2 = i == 2 ? pr : pb;
      double vol, K[3];
      if (dir) {                                                  // identity rows take no contributions (bilinear_form.hpp:183-186)
        if (FUSE_B) {
          const double det = (p1.x - p0.x) * (p2.y - p0.y) - (p1.y - p0.y) * (p2.x - p0.x);
          bsum += (i == 0 ? bc0 : (i == 1 ? bc1 : bc2)) * (0.5 * fabs(det));
        }
        return;
      }
      p1_row_from_corners<STIFF>(F, p0, p1, p2, i, vol, K);
      if (FUSE_B) bsum += (i == 0 ? bc0 : (i == 1 ? bc1 : bc2)) * vol;
      // own = local i, a = the smaller of the other two local indices, b = the larger
      own_sum += i == 0 ? K[0] : (i == 1 ? K[1] : K[2]); own_slot = s_own;
      add_slot(s_a, i == 0 ? K[1] : K[0]);
      add_slot(s_b, i == 2 ? K[1] : K[2]);
    };
#pragma unroll
    for (int k = 0; k < kAhead; ++k) {
      if ((unsigned)k < D) {
        const unsigned long long nxt = k + 1 < kAhead ? qa[k + 1 < kAhead ? k + 1 : 0]
                                                      : ((unsigned)(k + 1) < D ? __ldg(rec + (size_t)(k + 1) * kEllRows) : 0ull);
        if (k & 1) step(qa[k], nxt, ya, yb, xa, xb); else step(qa[k], nxt, xa, xb, ya, yb);
      }
    }
    for (unsigned k = kAhead; k < D; ++k) {                                  // kAhead is even: the current corners are in x
      step(__ldg(rec + (size_t)k * kEllRows), k + 1 < D ? __ldg(rec + (size_t)(k + 1) * kEllRows) : 0ull, xa, xb, ya, yb);
      xa = ya; xb = yb;
    }
    if (own_slot != 0xffffffffu) add_slot(own_slot, own_sum);
  } else {
    const uint4* rec = reinterpret_cast<const uint4*>(B.ell) + (size_t)B.ell_ptr[blockIdx.x] * kEllRows + threadIdx.x;
    uint4 q = D ? __ldg(rec) : make_uint4(0u, 0u, 0u, 0xffffffffu);
    __syncthreads();
    for (unsigned k = 0; k < D; ++k) {
      const uint4 cur = q;
      if (k + 1 < D) q = __ldg(rec + (size_t)(k + 1) * kEllRows);
      if (cur.w == 0xffffffffu) continue;
      const double2 p0 = __ldg(v2 + cur.x), p1 = __ldg(v2 + cur.y), p2 = __ldg(v2 + cur.z);
      const int i = (int)(cur.w >> 24);
      double vol, K[3];
      if (dir) {
        if (FUSE_B) {
          const double det = (p1.x - p0.x) * (p2.y - p0.y) - (p1.y - p0.y) * (p2.x - p0.x);
          bsum += (i == 0 ? bc0 : (i == 1 ? bc1 : bc2)) * (0.5 * fabs(det));
        }
        continue;
      }
      p1_row_from_corners<STIFF>(F, p0, p1, p2, i, vol, K);
      if (FUSE_B) bsum += (i == 0 ? bc0 : (i == 1 ? bc1 : bc2)) * vol;
      row[cur.w & 0xffu] += K[0];
      row[(cur.w >> 8) & 0xffu] += K[1];
      row[(cur.w >> 16) & 0xffu] += K[2];
    }
  }
  if (r < r1) {
    if (dir && B.init_mode) {                                   // clear(): identity row (bilinear_form.hpp:159-165)
      const int re = B.rowptr[r + to_local + 1];
      const int32_t gcol = (int32_t)(B.off_te + r);
      for (int e = rs; e < re; ++e) if (B.colind[e] == gcol) row[e - rs] = 1.0;
    }
    if (FUSE_B) { double* out = B.b + B.local_off + (r - B.row_begin); *out = B.b_init ? bsum : *out + bsum; }
  }
  __syncthreads();
  if (B.init_mode) for (int e = threadIdx.x; e < seg_len; e += kEllRows) B.val[(size_t)seg0 + e] = stage[e];
  else for (int e = threadIdx.x; e < seg_len; e += kEllRows) B.val[(size_t)seg0 + e] += stage[e];
}

}  // namespace

// P1 x P1 on triangles, one block, one row segment, packed records available
bool p1rows_applicable(const tfelcu_matrix* m) {
  if (m->test.n != 1 || m->trial.n != 1 || m->rows.n != 1) return false;
  const tfelcu_fes* te = m->test.fes[0]; const tfelcu_fes* tr = m->trial.fes[0];
  if (te->mesh->dim != 2 || te->fe != TFELCU_FE_P1 || tr->fe != TFELCU_FE_P1) return false;
  return true;
}

// ------------------------------------------------------------------------------------------------------------------
// The same rows as a STREAM (default where the vertex tables exist and no block has more than kStreamDeg records per
// row). k_assemble_p1_ell is bound by the chain of dependent DRAM latencies of a CTA's short life -- block metadata ->
// records / table -> coordinates -> compute -> write-out, about 7 us per 128 rows with 8-12 CTAs per SM in flight
// (fewer instructions, -25 %, or shared-memory coordinates alone did not move it: profiles/r02_asm_sweep_*.jsonl).
// Here CTAs are persistent and a producer warp runs that chain AHEAD of the consumers: per block it issues three bulk
// copies (TMA, cp.async.bulk -> mbarrier) -- the block's records, its vertex table, the coordinates of its own 128
// vertices --, loads row pointers and Dirichlet flags, and, one block behind, gathers the table's coordinates into shared
// memory. The four consumer warps (one row per thread) only touch shared memory until they write the finished CSR
// segment, coalesced. Same arithmetic in the same order as k_assemble_p1_ell: bitwise the same values.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kStreamDeg = 8;        // records per row a stage holds (rows of P1 triangles: 6 in the interior)
constexpr int kStreamStages = 3;

struct P1Stream {
  int n_blocks, deg_cap, vt_cap, acc_doubles;   // shared-memory extents: records per row, table entries, CSR segment
  const int4* desc;                              // per block: first record line, records per row, table offset, table entries
};

__host__ __device__ inline size_t p1_stream_stage_bytes(int deg_cap, int vt_cap) {
  return (size_t)deg_cap * kEllRows * 8 + (size_t)vt_cap * 4 + (size_t)vt_cap * 16 + (size_t)kEllRows * 16 + 132 * 4;
}

__global__ void k_ell_stream_desc(const unsigned int* __restrict__ ell_ptr, const unsigned int* __restrict__ ell_deg,
                                  const unsigned int* __restrict__ vt_ptr, const int32_t* __restrict__ rowptr, size_t row_begin,
                                  size_t row_end, size_t local_off, size_t n_blocks, int4* __restrict__ desc) {
  const size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  const size_t r0 = row_begin + b * kEllRows;
  const size_t r1 = r0 + kEllRows < row_end ? r0 + kEllRows : row_end;
  const int seg0 = rowptr[r0 - row_begin + local_off];
  desc[2 * b] = make_int4((int)ell_ptr[b], (int)ell_deg[b], (int)vt_ptr[b], (int)(vt_ptr[b + 1] - vt_ptr[b]));
  desc[2 * b + 1] = make_int4(seg0, rowptr[r1 - row_begin + local_off] - seg0, 0, 0);
}

template <bool FUSE_B, int SHAPE, int S>
__global__ void __launch_bounds__(kEllRows + 32)
k_assemble_p1_stream(const __grid_constant__ P1Form F, const P1Rows B, const P1Stream T) {
  constexpr bool STIFF = SHAPE == 1;
  extern __shared__ __align__(128) unsigned char raw[];
  __shared__ uint64_t bulk_done[S], full[S], empty[S];
  const size_t stage_bytes = p1_stream_stage_bytes(T.deg_cap, T.vt_cap);
  // stage layout: records | table (ids) | table (coordinates) | own coordinates | row pointers
  auto st_rec = [&](int s) { return reinterpret_cast<unsigned long long*>(raw + (size_t)s * stage_bytes); };
  auto st_ids = [&](int s) { return reinterpret_cast<uint32_t*>(raw + (size_t)s * stage_bytes + (size_t)T.deg_cap * kEllRows * 8); };
  auto st_co = [&](int s) { return reinterpret_cast<double2*>(reinterpret_cast<unsigned char*>(st_ids(s)) + (size_t)T.vt_cap * 4); };
  auto st_own = [&](int s) { return st_co(s) + T.vt_cap; };
  auto st_rp = [&](int s) { return reinterpret_cast<int32_t*>(st_own(s) + kEllRows); };
  double* acc0 = reinterpret_cast<double*>(raw + (size_t)S * stage_bytes);   // two CSR-segment images, used alternately
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&bulk_done[s], 1); mbar_init(&full[s], 32); mbar_init(&empty[s], kEllRows); }
    mbar_fence_init();
  }
  __syncthreads();
  const int first = blockIdx.x, stride = gridDim.x;
  const size_t to_local = B.local_off - B.row_begin;
  const double2* v2 = reinterpret_cast<const double2*>(B.vertices);

  if (tid >= kEllRows) {
    // ---- producer warp: nothing here waits for a load -- bulk copies and cp.async only, completion through mbarriers ----
    const int lane = tid - kEllRows;
    int4 d = first < T.n_blocks ? __ldg(T.desc + 2 * (size_t)first) : make_int4(0, 0, 0, 0);
    int4 prev = d;
    int it = 0;
    auto finish = [&](int jt, int4 dj) {      // block jt: table ids -> coordinates, then the stage goes to the consumers
      const int s = jt % S;
      mbar_wait_or_trap(&bulk_done[s], (uint32_t)(jt / S) & 1u);      // issued one block ago
      const uint32_t* ids = st_ids(s);
      double2* co = st_co(s);
      for (int t = lane; t < dj.w; t += 32) cp_async_16(co + t, v2 + ids[t]);
      cp_async_mbar_arrive(&full[s]);        // fires when these and the row pointers of the block (issued earlier) have landed
    };
    for (int b = first; b < T.n_blocks; b += stride, ++it) {
      const int s = it % S;
      const int4 cur = d;
      if (b + stride < T.n_blocks) d = __ldg(T.desc + 2 * (size_t)(b + stride));
      if (it >= S) mbar_wait_or_trap(&empty[s], (uint32_t)(it / S - 1) & 1u);
      const size_t r0 = B.row_begin + (size_t)b * kEllRows;
      const int nrows = (int)((B.row_end - r0) < (size_t)kEllRows ? (B.row_end - r0) : (size_t)kEllRows);
      if (lane == 0) {
        const uint32_t bytes = (uint32_t)cur.y * kEllRows * 8u + (uint32_t)cur.w * 4u + (uint32_t)nrows * 16u;
        mbar_expect_tx(&bulk_done[s], bytes);
        if (cur.y) bulk_g2s(st_rec(s), reinterpret_cast<const unsigned long long*>(B.ell) + (size_t)cur.x * kEllRows, (uint32_t)cur.y * kEllRows * 8u, &bulk_done[s]);
        if (cur.w) bulk_g2s(st_ids(s), B.vt_list + cur.z, (uint32_t)cur.w * 4u, &bulk_done[s]);
        bulk_g2s(st_own(s), v2 + r0, (uint32_t)nrows * 16u, &bulk_done[s]);
      }
      int32_t* rp = st_rp(s);                 // nrows + 1 row pointers
      for (int t = lane; t <= nrows; t += 32) cp_async_4(rp + t, B.rowptr + r0 + to_local + t);
      if (it > 0) finish(it - 1, prev);
      prev = cur;
    }
    if (it > 0) finish(it - 1, prev);
    return;
  }

  // ---- consumers: one row per thread -------------------------------------------------------------------------------
  const double bc0 = F.b_c[0], bc1 = F.b_c[1], bc2 = F.b_c[2];
  const bool b_uni = F.b_uniform != 0;
  int it = 0;
  // what a consumer reads from global memory itself -- records per row of the block, its row's Dirichlet flag -- is
  // requested one block ahead
  auto flag_of = [&](int blk) {
    const size_t rr = B.row_begin + (size_t)blk * kEllRows + tid;
    return (blk < T.n_blocks && rr < B.row_end) ? __ldg(B.dir_row + (size_t)B.off_te + rr) : (uint8_t)0;
  };
  int next_deg = first < T.n_blocks ? __ldg(&T.desc[2 * (size_t)first].y) : 0;
  uint8_t next_flag = flag_of(first);
  for (int b = first; b < T.n_blocks; b += stride, ++it) {
    const int s = it % S;
    const int D = next_deg;
    const uint8_t dflag = next_flag;
    if (b + stride < T.n_blocks) next_deg = __ldg(&T.desc[2 * (size_t)(b + stride)].y);
    next_flag = flag_of(b + stride);
    const size_t r = B.row_begin + (size_t)b * kEllRows + tid;
    const size_t left = B.row_end - (B.row_begin + (size_t)b * kEllRows);
    const int nrows = (int)(left < (size_t)kEllRows ? left : (size_t)kEllRows);
    const bool live = tid < nrows;
    mbar_wait_or_trap(&full[s], (uint32_t)(it / S) & 1u);
    mbar_wait_or_trap(&bulk_done[s], (uint32_t)(it / S) & 1u);       // (complete long ago: the producer saw it before the gathers)
    const int32_t* rp = st_rp(s);
    const int seg0 = rp[0], seg_len = rp[nrows] - seg0;
    const int rs = live ? rp[tid] : seg0 + seg_len, re = live ? rp[tid + 1] : seg0 + seg_len;
    const bool dir = dflag != 0;
    double* acc = acc0 + (size_t)(it & 1) * T.acc_doubles;
    for (int e = rs; e < re; ++e) acc[e - seg0] = 0.0;               // every thread clears and fills its own row only
    const unsigned row_addr = smem_u32(acc + (rs - seg0));
    auto add_slot = [row_addr](unsigned slot, double v) {
      const unsigned addr = row_addr + slot * 8u;
      double t;
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(t) : "r"(addr));
      t += v;
      asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(t) : "memory");
    };
    const unsigned long long* rec = st_rec(s) + tid;
    const double2* co = st_co(s);
    const double2 pr = st_own(s)[tid];
    double bsum = 0.0;
    double own_sum = 0.0; unsigned own_slot = 0xffffffffu;          // the diagonal entry: register sum, one store
    for (int k = 0; k < D; ++k) {
      const unsigned long long cur = rec[(size_t)k * kEllRows];
      const unsigned hi = (unsigned)(cur >> 32);
      if ((int)hi >= 0) continue;
      const double2 pa = co[(unsigned)cur & 0x1ffu], pb = co[(unsigned)(cur >> 21) & 0x1ffu];
      const int i = (int)((hi >> 28) & 3u);
      const unsigned s_own = (hi >> 10) & 0x3fu, s_a = (hi >> 16) & 0x3fu, s_b = (hi >> 22) & 0x3fu;
      if (SHAPE == 2) {
        const double ox = pb.x - pa.x, oy = pb.y - pa.y, ax = pr.x - pb.x, ay = pr.y - pb.y, bx = pa.x - pr.x, by = pa.y - pr.y;
        const double adet = fabs(bx * ay - by * ax);
        const double vol = 0.5 * adet;
        if (FUSE_B) bsum += (b_uni ? bc0 : (i == 0 ? bc0 : (i == 1 ? bc1 : bc2))) * vol;
        if (dir) continue;
        const double sk = F.iso_k * fast_rcp(adet);
        double k_own = sk * (ox * ox + oy * oy), k_a = sk * (ox * ax + oy * ay), k_b = sk * (ox * bx + oy * by);
        if (F.has00) { k_own += F.iso_md * vol; k_a += F.iso_mo * vol; k_b += F.iso_mo * vol; }
        own_sum += k_own; own_slot = s_own; add_slot(s_a, k_a); add_slot(s_b, k_b);
        continue;
      }
      const double2 p0 = i == 0 ? pr : pa, p1 = i == 0 ? pa : (i == 1 ? pr : pb), p2 = i == 2 ? pr : pb;
      double vol, K[3];
      if (dir) {
        if (FUSE_B) {
          const double det = (p1.x - p0.x) * (p2.y - p0.y) - (p1.y - p0.y) * (p2.x - p0.x);
          bsum += (i == 0 ? bc0 : (i == 1 ? bc1 : bc2)) * (0.5 * fabs(det));
        }
        continue;
      }
      p1_row_from_corners<STIFF>(F, p0, p1, p2, i, vol, K);
      if (FUSE_B) bsum += (i == 0 ? bc0 : (i == 1 ? bc1 : bc2)) * vol;
      own_sum += i == 0 ? K[0] : (i == 1 ? K[1] : K[2]); own_slot = s_own;
      add_slot(s_a, i == 0 ? K[1] : K[0]);
      add_slot(s_b, i == 2 ? K[1] : K[2]);
    }
    mbar_arrive(&empty[s]);                                           // the stage may be refilled
    if (own_slot != 0xffffffffu) add_slot(own_slot, own_sum);
    if (live) {
      if (dir && B.init_mode) {                                       // clear(): identity row (bilinear_form.hpp:159-165)
        const int32_t gcol = (int32_t)(B.off_te + r);
        for (int e = rs; e < re; ++e) if (B.colind[e] == gcol) acc[e - seg0] = 1.0;
      }
      if (FUSE_B) { double* out = B.b + B.local_off + (r - B.row_begin); *out = B.b_init ? bsum : *out + bsum; }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kEllRows) : "memory");      // consumers only: the segment image is complete
    if (B.init_mode) for (int e = tid; e < seg_len; e += kEllRows) B.val[(size_t)seg0 + e] = acc[e];
    else for (int e = tid; e < seg_len; e += kEllRows) B.val[(size_t)seg0 + e] += acc[e];
    // no second barrier: the next block uses the other image, and the block after that starts behind its own barrier
  }
}

static bool build_ell(tfelcu_matrix* m) {
  if (m->has_ell) return m->ell.n != 0 || m->rows.n_local() == 0;
  m->has_ell = true;
  Ctx* ctx = m->ctx;
  if (!m->has_slots && !m->slots_failed) {
    try { matrix_build_slots(m); } catch (const Refusal&) { m->has_slots = false; m->slots_failed = true; }
  }
  if (!m->has_slots) return false;                // rows with more than 256 entries
  matrix_build_recs(m);
  if (m->recs.empty() || !m->recs[0].n) return m->rows.n_local() == 0;
  tfelcu_fes* te = m->test.fes[0];
  const RowMap& R = m->rows;
  const size_t rb = R.g_begin[0] - m->test.offset[0], re = R.g_end[0] - m->test.offset[0];
  const size_t n_blocks = (re - rb + kEllRows - 1) / kEllRows;
  if (!n_blocks) return true;
  // identity rows hold one entry: the kernel writes it at the start of the row's stage -- true for the reference's pattern
  m->ell_deg.alloc(n_blocks + 1); m->ell_ptr.alloc(n_blocks + 1);
  ctx->zero(m->ell_deg.p, n_blocks + 1);
  k_ell_degree<<<(unsigned)n_blocks, kEllRows, 0, ctx->stream>>>(te->adj_ptr.p, rb, re, m->ell_deg.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  // nearly uniform degrees (structured meshes: 6 almost everywhere): pad every block to the largest one, so that a CTA
  // finds its record lines at block * degree without reading a pointer first
  m->ell_uni_deg = 0;
  {
    std::vector<unsigned int> h_deg(n_blocks + 1, 0u);
    ctx->download(h_deg.data(), m->ell_deg.p, n_blocks);
    unsigned long long sum = 0; unsigned int mx = 0;
    for (size_t b = 0; b < n_blocks; ++b) { sum += h_deg[b]; mx = std::max(mx, h_deg[b]); }
    if (mx > 0 && (double)mx * (double)n_blocks <= 1.08 * (double)sum && !std::getenv("TFELCU_NO_UNIFORM")) {
      for (size_t b = 0; b < n_blocks; ++b) h_deg[b] = mx;
      ctx->upload(m->ell_deg.p, h_deg.data(), n_blocks + 1);
      m->ell_uni_deg = mx;
    }
  }
  exclusive_sum<unsigned int, unsigned int>(ctx, m->ell_deg.p, m->ell_ptr.p, n_blocks + 1);
  unsigned int total = 0; ctx->download(&total, m->ell_ptr.p + n_blocks, 1);
  // compact 8-byte records when the dofs are the vertex ids and the offsets / slots fit their fields
  m->ell_compact = false;
  if (te->mesh->all_vertices_used && te->dof_map == te->mesh->cells.p && m->test.offset[0] == 0 && !std::getenv("TFELCU_NO_COMPACT")) {
    DevBuf<int> bad(1);
    ctx->zero(bad.p, 1);
    k_ell_check_compact<<<grid_for(re - rb, 256), 256, 0, ctx->stream>>>(te->adj_ptr.p, m->a_begin[0], rb, re,
                                                                       reinterpret_cast<const uint4*>(m->recs[0].p), bad.p);
    ctx->count(); TFELCU_LAUNCH_CHECK();
    int h_bad = 0; ctx->download(&h_bad, bad.p, 1);
    m->ell_compact = !h_bad;
  }
  if (m->ell_compact) {
    m->ell.alloc((size_t)total * kEllRows * 2);
    k_ell_fill_compact<<<(unsigned)n_blocks, kEllRows, 0, ctx->stream>>>(te->adj_ptr.p, m->a_begin[0], rb, re,
                                                                        reinterpret_cast<const uint4*>(m->recs[0].p), m->ell_ptr.p,
                                                                        m->ell_deg.p, reinterpret_cast<unsigned long long*>(m->ell.p));
  } else {
    m->ell.alloc((size_t)total * kEllRows * 4);
    k_ell_fill<<<(unsigned)n_blocks, kEllRows, 0, ctx->stream>>>(te->adj_ptr.p, m->a_begin[0], rb, re,
                                                                reinterpret_cast<const uint4*>(m->recs[0].p), m->ell_ptr.p,
                                                                m->ell_deg.p, reinterpret_cast<uint4*>(m->ell.p));
  }
  ctx->count(); TFELCU_LAUNCH_CHECK();
  m->ell_local = false;
  if (m->ell_compact && !std::getenv("TFELCU_NO_VTAB")) {
    DevBuf<unsigned int> cnt(n_blocks + 1), raw_cnt(n_blocks + 1);
    DevBuf<uint32_t> scratch(n_blocks * (size_t)kVtCap);      // the distinct ids of every block, until the tables are laid out
    DevBuf<int> bad(1);
    ctx->zero(cnt.p, n_blocks + 1); ctx->zero(bad.p, 1);
    unsigned long long* ell = reinterpret_cast<unsigned long long*>(m->ell.p);
    unsigned int deg_max = 0;
    {
      std::vector<unsigned int> h_deg(n_blocks);
      ctx->download(h_deg.data(), m->ell_deg.p, n_blocks);
      for (size_t b = 0; b < n_blocks; ++b) deg_max = std::max(deg_max, h_deg[b]);
    }
    if (2u * deg_max * kEllRows <= 2048u)
      k_ell_vertex_sets<2048><<<(unsigned)n_blocks, kEllRows, 0, ctx->stream>>>(rb, re, m->ell_ptr.p, m->ell_deg.p, ell, raw_cnt.p, cnt.p, scratch.p, bad.p);
    else
      k_ell_vertex_sets<4096><<<(unsigned)n_blocks, kEllRows, 0, ctx->stream>>>(rb, re, m->ell_ptr.p, m->ell_deg.p, ell, raw_cnt.p, cnt.p, scratch.p, bad.p);
    ctx->count(); TFELCU_LAUNCH_CHECK();
    int h_bad = 0; ctx->download(&h_bad, bad.p, 1);
    if (!h_bad) {
      m->vt_ptr.alloc(n_blocks + 1);
      m->vt_uni = 0;
      if (m->ell_uni_deg) {   // the same for the tables
        std::vector<unsigned int> h_cnt(n_blocks + 1, 0u);
        ctx->download(h_cnt.data(), cnt.p, n_blocks);
        unsigned long long sum = 0; unsigned int mx = 0;
        for (size_t b = 0; b < n_blocks; ++b) { sum += h_cnt[b]; mx = std::max(mx, h_cnt[b]); }
        if (mx > 0 && (double)mx * (double)n_blocks <= 1.08 * (double)sum) {
          for (size_t b = 0; b < n_blocks; ++b) h_cnt[b] = mx;
          ctx->upload(cnt.p, h_cnt.data(), n_blocks + 1);
          m->vt_uni = mx;
        }
      }
      exclusive_sum<unsigned int, unsigned int>(ctx, cnt.p, m->vt_ptr.p, n_blocks + 1);
      unsigned int n_list = 0; ctx->download(&n_list, m->vt_ptr.p + n_blocks, 1);
      m->vt_list.alloc((size_t)n_list + 1);
      k_ell_vertex_tables<<<(unsigned)n_blocks, kEllRows, 0, ctx->stream>>>(rb, re, m->ell_ptr.p, m->ell_deg.p, ell, raw_cnt.p, m->vt_ptr.p,
                                                                           scratch.p, m->vt_list.p);
      ctx->count(); TFELCU_LAUNCH_CHECK();
      m->ell_local = true;
      // extents and block descriptors of the streamed kernel
      std::vector<unsigned int> h_deg(n_blocks), h_cnt(n_blocks);
      ctx->download(h_deg.data(), m->ell_deg.p, n_blocks);
      ctx->download(h_cnt.data(), cnt.p, n_blocks);
      m->ell_deg_max = 0; m->vt_max = 0;
      for (size_t b = 0; b < n_blocks; ++b) { m->ell_deg_max = std::max(m->ell_deg_max, h_deg[b]); m->vt_max = std::max(m->vt_max, h_cnt[b]); }
      m->ell_desc.alloc(n_blocks * 8);
      k_ell_stream_desc<<<grid_for(n_blocks, 256), 256, 0, ctx->stream>>>(m->ell_ptr.p, m->ell_deg.p, m->vt_ptr.p, m->rowptr.p, rb, re,
                                                                          R.local_off[0], n_blocks, reinterpret_cast<int4*>(m->ell_desc.p));
      ctx->count(); TFELCU_LAUNCH_CHECK();
    }
  }
  ctx->sync();
  // the CSR-ordered records are only needed by the generic REC kernel: variable-coefficient forms on this pattern rebuild
  // them on demand
  m->recs.clear(); m->has_recs = false;
  return true;
}

bool matrix_ell_ready(tfelcu_matrix* m) { return p1rows_applicable(m) && build_ell(m); }

// a += (constant-coefficient form) [and b += (constant-coefficient linear form without derivatives)] in one launch
bool assemble_p1_rows(tfelcu_matrix* m, const FormBuild& fb, tfelcu_vector* v, const FormBuild* fbv) {
  if (!fb.const_coef || !matrix_ell_ready(m)) return false;
  Ctx* ctx = m->ctx;
  const RowMap& R = m->rows;
  const size_t rb = R.g_begin[0] - m->test.offset[0], re = R.g_end[0] - m->test.offset[0];
  if (re > rb) {
    const size_t stage_bytes = m->max_block_nnz * sizeof(double);
    if (m->rows_per_cta != kEllRows || stage_bytes > 200 * 1024) return false;
    const P1Form F = make_p1_form(fb, v ? fbv : nullptr);
    P1Rows B;
    B.vertices = m->mesh->vertices.p; B.dir_row = m->dir_row.p; B.rowptr = m->rowptr.p; B.colind = m->colind.p; B.val = m->val.p;
    B.init_mode = m->val_valid ? 0 : 1;
    B.ell = m->ell.p; B.ell_ptr = m->ell_ptr.p; B.ell_deg = m->ell_deg.p;
    B.row_begin = rb; B.row_end = re; B.local_off = R.local_off[0]; B.off_te = (uint32_t)m->test.offset[0];
    B.b = v ? v->data.p : nullptr; B.b_init = (v && v->zero_pending) ? 1 : 0;
    const bool local = m->ell_compact && m->ell_local;
    B.vt_ptr = m->vt_ptr.p; B.vt_list = m->vt_list.p;
    B.stage_doubles = (int)((m->max_block_nnz + 1) & ~(size_t)1);
    B.desc = reinterpret_cast<const int4*>(m->ell_desc.p);
    B.uni_deg = (m->ell_uni_deg && m->vt_uni) ? m->ell_uni_deg : 0u; B.uni_vt = B.uni_deg ? m->vt_uni : 0u;
    const size_t smem_bytes = local ? (size_t)B.stage_doubles * sizeof(double) + (size_t)kVtCap * sizeof(double2) : stage_bytes;
    if (v) v->zero_pending = false;
    const unsigned grid = grid_for(re - rb, kEllRows);
    const size_t n_blocks_total = grid;
    const bool stiff = F.hasrr && !F.has00 && !F.has0r && !F.hasr0;
    const int shape = (F.iso && m->ell_compact && !std::getenv("TFELCU_NO_ISO")) ? 2 : (stiff ? 1 : 0);
#define TFELCU_P1(FB, SH, CP, LV)                                                                  \
    do {                                                                                             \
      ensure_smem(k_assemble_p1_ell<FB, SH, CP, LV>, smem_bytes);                                    \
      k_assemble_p1_ell<FB, SH, CP, LV><<<grid, kEllRows, smem_bytes, ctx->stream>>>(F, B);          \
    } while (0)
#define TFELCU_P1C(FB, SH) do { if (local) TFELCU_P1(FB, SH, true, true); else TFELCU_P1(FB, SH, true, false); } while (0)
    const bool stream = local && m->ell_deg_max <= (unsigned)kStreamDeg && n_blocks_total < 0x7fffffffull && std::getenv("TFELCU_P1STREAM");
    if (stream) {
      P1Stream T;
      T.n_blocks = (int)n_blocks_total; T.deg_cap = (int)std::max(1u, m->ell_deg_max); T.vt_cap = (int)((m->vt_max + 3u) & ~3u);
      T.acc_doubles = B.stage_doubles; T.desc = reinterpret_cast<const int4*>(m->ell_desc.p);
      int stages = kStreamStages;
      if (const char* e = std::getenv("TFELCU_P1_STAGES")) stages = std::atoi(e) == 2 ? 2 : (std::atoi(e) == 4 ? 4 : 3);
      const size_t sm = (size_t)stages * p1_stream_stage_bytes(T.deg_cap, T.vt_cap) + 2 * (size_t)T.acc_doubles * sizeof(double);
      if (sm <= 200 * 1024) {
#define TFELCU_P1S_(FB, SH, ST)                                                                                    \
        do {                                                                                                        \
          ensure_smem(k_assemble_p1_stream<FB, SH, ST>, sm);                                                        \
          int per_sm = 0;                                                                                           \
          TFELCU_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_assemble_p1_stream<FB, SH, ST>, kEllRows + 32, sm)); \
          if (const char* e = std::getenv("TFELCU_P1_CTAS")) per_sm = std::min(per_sm, std::max(1, std::atoi(e)));     \
          const unsigned g = (unsigned)std::min<size_t>(n_blocks_total, (size_t)std::max(1, per_sm) * (size_t)ctx->sm_count); \
          k_assemble_p1_stream<FB, SH, ST><<<g, kEllRows + 32, sm, ctx->stream>>>(F, B, T);                          \
        } while (0)
#define TFELCU_P1S(FB, SH) do { if (stages == 2) TFELCU_P1S_(FB, SH, 2); else if (stages == 4) TFELCU_P1S_(FB, SH, 4); else TFELCU_P1S_(FB, SH, 3); } while (0)
        if (v) { if (shape == 2) TFELCU_P1S(true, 2); else if (shape == 1) TFELCU_P1S(true, 1); else TFELCU_P1S(true, 0); }
        else { if (shape == 2) TFELCU_P1S(false, 2); else if (shape == 1) TFELCU_P1S(false, 1); else TFELCU_P1S(false, 0); }
#undef TFELCU_P1S
#undef TFELCU_P1S_
        ctx->count(); TFELCU_LAUNCH_CHECK();
        m->val_valid = true;
        return true;
      }
    }
    const int sel = (v ? 8 : 0) | (shape << 1) | (m->ell_compact ? 1 : 0);
    switch (sel) {
      case 0: TFELCU_P1(false, 0, false, false); break; case 1: TFELCU_P1C(false, 0); break;
      case 2: TFELCU_P1(false, 1, false, false); break; case 3: TFELCU_P1C(false, 1); break;
      case 5: TFELCU_P1C(false, 2); break;
      case 8: TFELCU_P1(true, 0, false, false); break;  case 9: TFELCU_P1C(true, 0); break;
      case 10: TFELCU_P1(true, 1, false, false); break; case 11: TFELCU_P1C(true, 1); break;
      case 13: TFELCU_P1C(true, 2); break;
      default: fail("p1 rows: unexpected kernel selection");
    }
#undef TFELCU_P1C
#undef TFELCU_P1
    ctx->count(); TFELCU_LAUNCH_CHECK();
  }
  m->val_valid = true;
  return true;
}

}  // namespace tfelcu
