// Owner-computes rows for constant-coefficient forms on P1 triangles (configs[0] and [1]): the row kernel of assemble.cu
// (bilinear_form.hpp:64-109 restricted to the cells that touch a row, in ascending cell order) with
//   * the element matrix row in closed form (p1form.cuh) instead of the contraction of a pre-integrated tensor staged in
//     shared memory: 178 -> ~90 instructions per (row, cell);
//   * the packed 16-byte (cell vertices | slots | local dof) records in a blocked ELL layout -- record k of the 128 rows of
//     a CTA back to back -- so that a warp reads 512 contiguous bytes per step instead of 32 records 96 bytes apart
//     (ncu on the CSR-ordered records: 17 sectors per load request);
//   * a constant-coefficient linear form without derivatives on the same space riding along (linear_form.hpp:20-142:
//     b_i += |K| sum_q w_q f psi_i): the row thread has the cell volumes anyway, the second pass over the mesh (0.52 ms of
//     1.35 ms on configs[1]) disappears.
// Same stage / write-out as k_assemble_rows: every stored value is written once, coalesced; no atomics; the additions to
// an entry happen in ascending cell order (bitwise reproducible).
#include "p1form.cuh"
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
//   bits 0-20 offset a | 21-41 offset b | 42-59 three slots | 60-61 local dof i | 63 valid     (a < b: the local indices != i)
__device__ __host__ __forceinline__ bool compact_fits(long long da, long long db, uint32_t w) {
  const long long lim = 1ll << 20;
  return da > -lim && da < lim && db > -lim && db < lim && (w & 0xffu) < 64u && ((w >> 8) & 0xffu) < 64u && ((w >> 16) & 0xffu) < 64u;
}

__global__ void k_ell_check_compact(const uint32_t* __restrict__ adj_ptr, size_t a_begin, size_t row_begin, size_t row_end,
                                    const uint4* __restrict__ rec, int* __restrict__ bad) {
  const size_t r = row_begin + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r >= row_end) return;
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
      c = da | (db << 21) | ((unsigned long long)(q.w & 0x3fu) << 42) | ((unsigned long long)((q.w >> 8) & 0x3fu) << 48) |
          ((unsigned long long)((q.w >> 16) & 0x3fu) << 54) | ((unsigned long long)i << 60) | (1ull << 63);
    }
    out[(size_t)k * kEllRows] = c;
  }
}

struct P1Rows {
  const double* vertices;
  const uint8_t* dir_row; const int32_t* rowptr; const int32_t* colind; double* val; int init_mode;
  const void* ell; const unsigned int* ell_ptr; const unsigned int* ell_deg;
  size_t row_begin, row_end, local_off; uint32_t off_te;
  double* b;            // fused linear form (local rows), or nullptr
  int b_init;           // 1: its clear() is pending: write the entries
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
template <bool FUSE_B, int SHAPE, bool COMPACT>
__global__ void __launch_bounds__(kEllRows)
k_assemble_p1_ell(const __grid_constant__ P1Form F, const P1Rows B) {
  constexpr bool STIFF = SHAPE == 1;
  extern __shared__ double stage[];
  const size_t r0 = B.row_begin + (size_t)blockIdx.x * kEllRows;
  size_t r1 = r0 + kEllRows; if (r1 > B.row_end) r1 = B.row_end;
  const size_t to_local = B.local_off - B.row_begin;
  const int seg0 = B.rowptr[r0 + to_local];
  const int seg_len = B.rowptr[r1 + to_local] - seg0;
  for (int e = threadIdx.x; e < seg_len; e += kEllRows) stage[e] = 0.0;
  const unsigned D = B.ell_deg[blockIdx.x];
  const size_t r = r0 + threadIdx.x;
  int rs = 0;
  bool dir = false;
  if (r < r1) { rs = B.rowptr[r + to_local]; dir = B.dir_row[(size_t)B.off_te + r] != 0; }
  const double2* v2 = reinterpret_cast<const double2*>(B.vertices);
  const double bc0 = F.b_c[0], bc1 = F.b_c[1], bc2 = F.b_c[2];
  double bsum = 0.0;
  double* row = stage + (rs - seg0);
  if (COMPACT) {
    const unsigned long long* rec = reinterpret_cast<const unsigned long long*>(B.ell) + (size_t)B.ell_ptr[blockIdx.x] * kEllRows + threadIdx.x;
    // The record stream is the DRAM stream of this kernel: the first eight records of a row (all of them on every mesh
    // seen so far) are requested up front, 64 bytes in flight per thread (with one record ahead the loop waited on the
    // long scoreboard: 7.5 stall cycles per issue, 3.2 TB/s); the two corner gathers of cell k + 1 are in flight while
    // cell k is computed.
    const long long rc = (long long)(r < r1 ? r : r1 - 1);                   // rows past the end read a valid vertex
    constexpr int kAhead = 8;
    unsigned long long qa[kAhead];
#pragma unroll
    for (int k = 0; k < kAhead; ++k) qa[k] = (unsigned)k < D ? __ldg(rec + (size_t)k * kEllRows) : 0ull;
    const double2 pr = __ldg(v2 + rc);                                       // the row's own vertex: once per row
    auto off_a = [](unsigned long long c) { return ((long long)(c << 43)) >> 43; };   // sign-extended 21-bit fields
    auto off_b = [](unsigned long long c) { return ((long long)(c << 22)) >> 43; };
    double2 na = __ldg(v2 + rc + off_a(qa[0])), nb = __ldg(v2 + rc + off_b(qa[0]));
    __syncthreads();
    auto step = [&](unsigned long long cur, unsigned long long nxt) {
      const double2 pa = na, pb = nb;
      na = __ldg(v2 + rc + off_a(nxt)); nb = __ldg(v2 + rc + off_b(nxt));   // padding records: offsets 0
      if (!(cur >> 63)) return;
      const int i = (int)((cur >> 60) & 3u);
      if (SHAPE == 2) {
        // edges opposite to the row's vertex, to a and to b; K_ij = (c W / 4|K|) e_i . e_j + c00 |K| Mhat_ij
        const double ox = pb.x - pa.x, oy = pb.y - pa.y, ax = pr.x - pb.x, ay = pr.y - pb.y, bx = pa.x - pr.x, by = pa.y - pr.y;
        const double adet = fabs(bx * ay - by * ax);
        const double vol = 0.5 * adet;
        if (FUSE_B) bsum += (i == 0 ? bc0 : (i == 1 ? bc1 : bc2)) * vol;
        if (dir) return;
        const double sk = F.iso_k * fast_rcp(adet);
        double k_own = sk * (ox * ox + oy * oy), k_a = sk * (ox * ax + oy * ay), k_b = sk * (ox * bx + oy * by);
        if (F.has00) { k_own += F.iso_md * vol; k_a += F.iso_mo * vol; k_b += F.iso_mo * vol; }
        // slots are stored in local order j = 0, 1, 2: own = i, a = the smaller of the other two, b = the larger
        const unsigned sl = (unsigned)(cur >> 42) & 0x3ffffu;
        const int s_own = 6 * i, s_a = i == 0 ? 6 : 0, s_b = i == 2 ? 6 : 12;
        row[(sl >> s_own) & 0x3fu] += k_own;
        row[(sl >> s_a) & 0x3fu] += k_a;
        row[(sl >> s_b) & 0x3fu] += k_b;
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
      row[(cur >> 42) & 0x3fu] += K[0];
      row[(cur >> 48) & 0x3fu] += K[1];
      row[(cur >> 54) & 0x3fu] += K[2];
    };
#pragma unroll
    for (int k = 0; k < kAhead; ++k) {
      if ((unsigned)k < D) {
        const unsigned long long nxt = k + 1 < kAhead ? qa[k + 1 < kAhead ? k + 1 : 0]
                                                      : ((unsigned)(k + 1) < D ? __ldg(rec + (size_t)(k + 1) * kEllRows) : 0ull);
        step(qa[k], nxt);
      }
    }
    for (unsigned k = kAhead; k < D; ++k)
      step(__ldg(rec + (size_t)k * kEllRows), k + 1 < D ? __ldg(rec + (size_t)(k + 1) * kEllRows) : 0ull);
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
    if (v) v->zero_pending = false;
    const unsigned grid = grid_for(re - rb, kEllRows);
    const bool stiff = F.hasrr && !F.has00 && !F.has0r && !F.hasr0;
    const int shape = (F.iso && m->ell_compact && !std::getenv("TFELCU_NO_ISO")) ? 2 : (stiff ? 1 : 0);
#define TFELCU_P1(FB, SH, CP)                                                              \
    do {                                                                                     \
      ensure_smem(k_assemble_p1_ell<FB, SH, CP>, stage_bytes);                               \
      k_assemble_p1_ell<FB, SH, CP><<<grid, kEllRows, stage_bytes, ctx->stream>>>(F, B);     \
    } while (0)
    const int sel = (v ? 8 : 0) | (shape << 1) | (m->ell_compact ? 1 : 0);
    switch (sel) {
      case 0: TFELCU_P1(false, 0, false); break; case 1: TFELCU_P1(false, 0, true); break;
      case 2: TFELCU_P1(false, 1, false); break; case 3: TFELCU_P1(false, 1, true); break;
      case 5: TFELCU_P1(false, 2, true); break;
      case 8: TFELCU_P1(true, 0, false); break;  case 9: TFELCU_P1(true, 0, true); break;
      case 10: TFELCU_P1(true, 1, false); break; case 11: TFELCU_P1(true, 1, true); break;
      case 13: TFELCU_P1(true, 2, true); break;
      default: fail("p1 rows: unexpected kernel selection");
    }
#undef TFELCU_P1
    ctx->count(); TFELCU_LAUNCH_CHECK();
  }
  m->val_valid = true;
  return true;
}

}  // namespace tfelcu
