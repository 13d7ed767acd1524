// Sparse structure of a bilinear form on device: the CSR pattern the reference hands to PETSc,
// the transposed dof maps (row -> incident (cell, local dof)) that drive the owner-computes
// assembly, and the cell colouring of the colour-by-colour scatter.
// Reference: sparse_matrix as std::map<(i, j), double> (src/core/solver.hpp:259-265) turned into CRS
// in solver.cpp:71-96: rows ascending, columns ascending, one entry per (i, j) ever touched.
// bilinear_form.hpp:103-108,183-186: every (test dof, trial dof) pair of every cell is touched unless
// the row is a Dirichlet dof; clear() (:159-165) touches the diagonal of every Dirichlet row.
// Composite forms touch the pairs of ALL (m, n) blocks (composite_bilinear_form.hpp:173-182).
#include "structs.cuh"
#include "sortutil.cuh"

#include <cstdlib>

namespace tfelcu {

__global__ void k_copy_flags(const uint8_t* __restrict__ flag, const double* __restrict__ value, size_t n,
                             size_t off, uint8_t* __restrict__ out_flag, double* __restrict__ out_value) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  out_flag[off + i] = flag[i];
  out_value[off + i] = value[i];
}

void matrix_snapshot_dirichlet(tfelcu_matrix* m) {
  Ctx* ctx = m->ctx;
  if (!m->dir_row.n) { m->dir_row.alloc(m->n_rows); m->dir_value.alloc(m->n_rows); }
  for (int c = 0; c < m->test.n; ++c) {
    tfelcu_fes* f = m->test.fes[c];
    k_copy_flags<<<grid_for(f->n_dof, 256), 256, 0, ctx->stream>>>(f->dir_flag.p, f->dir_value.p, f->n_dof,
                                                                 m->test.offset[c], m->dir_row.p, m->dir_value.p);
    ctx->count(); TFELCU_LAUNCH_CHECK();
    m->dir_version[c] = f->dirichlet_version;
  }
}

// keys (local row << 32 | global col) of one (test component, trial component) block, generated from the OWNED rows:
// entry a of the transposed test dof map is (cell k, local dof i) of row r = dof_map_te[code]; it touches the nd_tr
// columns of cell k
__global__ void k_pattern_keys(const uint32_t* __restrict__ adj, size_t a_begin, size_t n_adj,
                               const uint32_t* __restrict__ map_te, int nd_te, uint32_t off_te,
                               unsigned long long row_begin, unsigned long long local_off,
                               const uint32_t* __restrict__ map_tr, int nd_tr, uint32_t off_tr,
                               const uint8_t* __restrict__ dir_row, uint64_t sentinel, uint64_t* __restrict__ keys) {
  const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= n_adj * (size_t)nd_tr) return;
  const size_t a = a_begin + p / nd_tr; const int j = (int)(p % nd_tr);
  const uint32_t code = adj[a];
  const size_t k = code / nd_te;
  const uint32_t r = map_te[code];
  const uint32_t col = off_tr + map_tr[k * nd_tr + j];
  const uint64_t lrow = local_off + (r - row_begin);
  keys[p] = dir_row[off_te + r] ? sentinel : ((lrow << 32) | col);
}

// identity entry of every owned Dirichlet row
__global__ void k_dirichlet_keys(const uint8_t* __restrict__ dir_row, RowMap R, uint64_t sentinel,
                                 uint64_t* __restrict__ keys) {
  const size_t lr = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (lr >= R.local_off[R.n]) return;
  const unsigned long long g = R.to_global(lr);
  keys[lr] = dir_row[g] ? (((uint64_t)lr << 32) | g) : sentinel;
}

__global__ void k_rowptr_from_keys(const uint64_t* __restrict__ keys, size_t nnz, size_t n_rows,
                                   int32_t* __restrict__ rowptr) {
  const size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  const uint64_t target = (uint64_t)r << 32;
  size_t lo = 0, hi = nnz;
  while (lo < hi) {
    const size_t mid = lo + (hi - lo) / 2;
    if (keys[mid] < target) lo = mid + 1; else hi = mid;
  }
  rowptr[r] = (int32_t)lo;
}

__global__ void k_colind_from_keys(const uint64_t* __restrict__ keys, size_t nnz, int32_t* __restrict__ colind) {
  const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (e < nnz) colind[e] = (int32_t)(keys[e] & 0xffffffffull);
}

__global__ void k_max_block_nnz(const int32_t* __restrict__ rowptr, size_t n_rows, int rows_per_block,
                                unsigned long long* __restrict__ out) {
  const size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t r0 = b * rows_per_block;
  if (r0 >= n_rows) return;
  size_t r1 = r0 + rows_per_block; if (r1 > n_rows) r1 = n_rows;
  atomicMax(out, (unsigned long long)(rowptr[r1] - rowptr[r0]));
}

void matrix_build_pattern(tfelcu_matrix* m) {
  Ctx* ctx = m->ctx;
  const RowMap& R = m->rows;
  const size_t n_local = R.n_local();
  TFELCU_REQUIRE(m->n_rows < 0xfffffffeull && m->n_cols < 0xfffffffeull, "matrix too large for 32-bit indices");
  // owned slices of the transposed test dof maps, one per row segment
  size_t* a_begin = m->a_begin; size_t* a_count = m->a_count;
  for (int sg = 0; sg < R.n; ++sg) {
    const int a = R.comp[sg];
    tfelcu_fes* te = m->test.fes[a];
    fes_transpose(te);
    uint32_t lo = 0, hi = 0;
    ctx->download(&lo, te->adj_ptr.p + (R.g_begin[sg] - m->test.offset[a]), 1);
    ctx->download(&hi, te->adj_ptr.p + (R.g_end[sg] - m->test.offset[a]), 1);
    a_begin[sg] = lo; a_count[sg] = hi - lo;
  }
  size_t n_keys = n_local;
  for (int sg = 0; sg < R.n; ++sg)
    for (int b = 0; b < m->trial.n; ++b) n_keys += a_count[sg] * (size_t)m->trial.fes[b]->nd;
  const uint64_t sentinel = (uint64_t)n_local << 32;   // sorts after every real key
  DevBuf<uint64_t> keys(n_keys), sorted(n_keys);
  size_t at = 0;
  for (int sg = 0; sg < R.n; ++sg)
    for (int b = 0; b < m->trial.n; ++b) {
      const int a = R.comp[sg];
      tfelcu_fes* te = m->test.fes[a]; tfelcu_fes* tr = m->trial.fes[b];
      const size_t n = a_count[sg] * (size_t)tr->nd;
      if (!n) continue;
      k_pattern_keys<<<grid_for(n, 256), 256, 0, ctx->stream>>>(te->adj.p, a_begin[sg], a_count[sg], te->dof_map, te->nd,
                                                              (uint32_t)m->test.offset[a], R.g_begin[sg] - m->test.offset[a],
                                                              R.local_off[sg], tr->dof_map, tr->nd,
                                                              (uint32_t)m->trial.offset[b], m->dir_row.p, sentinel, keys.p + at);
      ctx->count(); TFELCU_LAUNCH_CHECK();
      at += n;
    }
  if (n_local) {
    k_dirichlet_keys<<<grid_for(n_local, 256), 256, 0, ctx->stream>>>(m->dir_row.p, R, sentinel, keys.p + at);
    ctx->count(); TFELCU_LAUNCH_CHECK();
  }
  sort_keys<uint64_t>(ctx, keys.p, sorted.p, n_keys, 0, 32 + bits_for(n_local + 1));
  size_t nnz = unique_keys<uint64_t>(ctx, sorted.p, keys.p, n_keys);
  if (nnz) {
    uint64_t last = 0; ctx->download(&last, keys.p + (nnz - 1), 1);
    if (last == sentinel) --nnz;
  }
  TFELCU_REQUIRE(nnz < 0x7fffffffull, "more than 2^31 stored entries: int32 CSR indices (solver.cpp:71-74) overflow");
  m->nnz = nnz;
  m->rowptr.alloc(n_local + 1); m->colind.alloc(nnz); m->val.alloc(nnz);
  k_rowptr_from_keys<<<grid_for(n_local + 1, 256), 256, 0, ctx->stream>>>(keys.p, nnz, n_local, m->rowptr.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  if (nnz) {
    k_colind_from_keys<<<grid_for(nnz, 256), 256, 0, ctx->stream>>>(keys.p, nnz, m->colind.p);
    ctx->count(); TFELCU_LAUNCH_CHECK();
  }
  DevBuf<unsigned long long> mx(1);
  ctx->zero(mx.p, 1);
  // rows per CTA of the owner-computes assembly. Measured on tet P2 (n = 100): 128 rows 4.4 ms, 64 rows 5.9 ms,
  // 32 rows 9.3 ms -- the per-CTA staging of the reference tensor outweighs the occupancy gained by a smaller stage
  m->rows_per_cta = kRowsPerCta;
  if (const char* e = std::getenv("TFELCU_ROWS_PER_CTA")) { const int v = std::atoi(e); if (v == 32 || v == 64 || v == 128) m->rows_per_cta = v; }
  const size_t n_blocks = (n_local + m->rows_per_cta - 1) / m->rows_per_cta;
  if (n_blocks) {
    k_max_block_nnz<<<grid_for(n_blocks, 256), 256, 0, ctx->stream>>>(m->rowptr.p, n_local, m->rows_per_cta, mx.p);
    ctx->count(); TFELCU_LAUNCH_CHECK();
  }
  unsigned long long h = 0; ctx->download(&h, mx.p, 1);
  m->max_block_nnz = (size_t)h;
  m->has_colors = false;
  m->val_valid = false;
  m->has_slots = false; m->slots_failed = false;
  m->slots.clear();
  m->has_patches = false;
  m->patch_stage_off.clear();
  m->has_recs = false;
  m->recs.clear();
  m->tile_refused = false;
  m->has_ell = false; m->ell.release(); m->ell_ptr.release(); m->ell_deg.release();
  m->ell_local = false; m->vt_ptr.release(); m->vt_list.release(); m->ell_desc.release(); m->ell_uni_deg = 0; m->vt_uni = 0;
  static uint64_t next_pattern_id = 0;   // one host thread drives a context (API not re-entrant)
  m->pattern_id = ++next_pattern_id;
}

// one thread per entry a of the owned slice of the transposed test dof map: positions of its nd_tr columns in the row
__global__ void k_build_slots(const uint32_t* __restrict__ adj, size_t a_begin, size_t n_adj,
                              const uint32_t* __restrict__ map_te, int nd_te, uint32_t off_te,
                              unsigned long long row_begin, unsigned long long local_off,
                              const uint32_t* __restrict__ map_tr, int nd_tr, uint32_t off_tr,
                              const uint8_t* __restrict__ dir_row, const int32_t* __restrict__ rowptr,
                              const int32_t* __restrict__ colind, int words, uint32_t* __restrict__ slots, int* __restrict__ bad) {
  const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= n_adj) return;
  const uint32_t code = adj[a_begin + p];
  const size_t k = code / nd_te;
  const uint32_t r = map_te[code];
  uint32_t* out = slots + p * words;
  for (int t = 0; t < words; ++t) out[t] = 0u;
  if (dir_row[off_te + r]) return;
  const size_t lrow = local_off + (r - row_begin);
  const int rs = rowptr[lrow], re = rowptr[lrow + 1];
  for (int j = 0; j < nd_tr; ++j) {
    const int32_t col = (int32_t)(off_tr + map_tr[k * nd_tr + j]);
    int lo = rs, hi = re;
    while (lo < hi) {
      const int mid = lo + ((hi - lo) >> 1);   // lo + hi overflows int32 beyond 2^30 stored entries
      if (colind[mid] < col) lo = mid + 1; else hi = mid;
    }
    if (lo >= re || colind[lo] != col || lo - rs > 255) { atomicExch(bad, 1); return; }
    out[j >> 2] |= (uint32_t)(lo - rs) << (8 * (j & 3));
  }
}

void matrix_build_slots(tfelcu_matrix* m) {
  if (m->has_slots) return;
  Ctx* ctx = m->ctx;
  const RowMap& R = m->rows;
  m->slots.clear();
  m->slots.resize((size_t)R.n * m->trial.n);
  DevBuf<int> bad(1);
  ctx->zero(bad.p, 1);
  for (int sg = 0; sg < R.n; ++sg)
    for (int b = 0; b < m->trial.n; ++b) {
      const int a = R.comp[sg];
      tfelcu_fes* te = m->test.fes[a]; tfelcu_fes* tr = m->trial.fes[b];
      const int words = (tr->nd + 3) / 4;
      const size_t n = m->a_count[sg];
      DevBuf<uint32_t>& S = m->slots[(size_t)sg * m->trial.n + b];
      S.alloc(n * words);
      if (!n) continue;
      k_build_slots<<<grid_for(n, 256), 256, 0, ctx->stream>>>(te->adj.p, m->a_begin[sg], n, te->dof_map, te->nd,
                                                             (uint32_t)m->test.offset[a], R.g_begin[sg] - m->test.offset[a],
                                                             R.local_off[sg], tr->dof_map, tr->nd, (uint32_t)m->trial.offset[b],
                                                             m->dir_row.p, m->rowptr.p, m->colind.p, words, S.p, bad.p);
      ctx->count(); TFELCU_LAUNCH_CHECK();
    }
  int h = 0; ctx->download(&h, bad.p, 1);
  if (h) { m->slots.clear(); refuse("slot map: a row has more than 256 stored entries or the pattern is inconsistent"); }
  m->has_slots = true;
}

// packed row records of P1 triangles: (v0, v1, v2, slot bytes | local dof << 24) per entry of the transposed dof map
__global__ void k_build_rec(const uint32_t* __restrict__ adj, size_t a_begin, size_t n_adj, const uint32_t* __restrict__ cells,
                            const uint32_t* __restrict__ slots, uint4* __restrict__ rec) {
  const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= n_adj) return;
  const uint32_t code = adj[a_begin + p];
  const uint32_t k = code / 3u, i = code - 3u * k;
  const uint32_t* c = cells + (size_t)k * 3;
  rec[p] = make_uint4(c[0], c[1], c[2], (slots ? (slots[p] & 0x00ffffffu) : 0u) | (i << 24));
}

void matrix_build_recs(tfelcu_matrix* m) {
  if (m->has_recs) return;
  Ctx* ctx = m->ctx;
  const RowMap& R = m->rows;
  m->recs.clear();
  m->recs.resize((size_t)R.n * m->trial.n);
  m->has_recs = true;
  if (!m->has_slots || m->mesh->dim != 2) return;
  for (int sg = 0; sg < R.n; ++sg)
    for (int b = 0; b < m->trial.n; ++b) {
      tfelcu_fes* te = m->test.fes[R.comp[sg]]; tfelcu_fes* tr = m->trial.fes[b];
      if (te->nd != 3 || tr->nd != 3) continue;
      const size_t n = m->a_count[sg];
      DevBuf<uint32_t>& Rc = m->recs[(size_t)sg * m->trial.n + b];
      Rc.alloc(n * 4);
      if (!n) continue;
      k_build_rec<<<grid_for(n, 256), 256, 0, ctx->stream>>>(te->adj.p, m->a_begin[sg], n, m->mesh->cells.p,
                                                           m->slots[(size_t)sg * m->trial.n + b].p,
                                                           reinterpret_cast<uint4*>(Rc.p));
      ctx->count(); TFELCU_LAUNCH_CHECK();
    }
}

// clear(): zero everything, identity on the Dirichlet rows (bilinear_form.hpp:159-165)
__global__ void k_clear_values(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                               const uint8_t* __restrict__ dir_row, RowMap R, double* __restrict__ val) {
  const size_t lr = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (lr >= R.local_off[R.n]) return;
  const unsigned long long g = R.to_global(lr);
  const bool dir = dir_row[g];
  for (int32_t e = rowptr[lr]; e < rowptr[lr + 1]; ++e) val[e] = (dir && colind[e] == (int32_t)g) ? 1.0 : 0.0;
}

void matrix_ensure_cleared(tfelcu_matrix* m) {
  if (m->pending) matrix_flush_pending(m);
  if (m->val_valid) return;
  Ctx* ctx = m->ctx;
  if (m->rows.n_local()) {
    k_clear_values<<<grid_for(m->rows.n_local(), 256), 256, 0, ctx->stream>>>(m->rowptr.p, m->colind.p, m->dir_row.p, m->rows, m->val.p);
    ctx->count(); TFELCU_LAUNCH_CHECK();
  }
  m->val_valid = true;
}

// ------------------------------------------------------------------------------------------
// transposed dof map: for every dof the ascending list of k * nd + i with dof_map[k][i] == dof
// ------------------------------------------------------------------------------------------

__global__ void k_lower_bounds_u32(const uint32_t* __restrict__ sorted, size_t n, size_t n_targets,
                                   uint32_t* __restrict__ out) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t > n_targets) return;
  size_t lo = 0, hi = n;
  while (lo < hi) {
    const size_t mid = lo + (hi - lo) / 2;
    if (sorted[mid] < (uint32_t)t) lo = mid + 1; else hi = mid;
  }
  out[t] = (uint32_t)lo;
}

void build_transpose(Ctx* ctx, const uint32_t* dof_map, size_t K, int nd, size_t n_dof,
                     DevBuf<uint32_t>& ptr, DevBuf<uint32_t>& adj) {
  const size_t n = K * (size_t)nd;
  DevBuf<uint32_t> idx(n), keys_sorted(n);
  iota_u32(ctx, idx.p, n);
  adj.alloc(n);
  sort_pairs<uint32_t, uint32_t>(ctx, dof_map, keys_sorted.p, idx.p, adj.p, n, 0, bits_for(n_dof));
  ptr.alloc(n_dof + 1);
  k_lower_bounds_u32<<<grid_for(n_dof + 1, 256), 256, 0, ctx->stream>>>(keys_sorted.p, n, n_dof, ptr.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  ctx->sync();
}

void fes_transpose(tfelcu_fes* f) {
  if (f->has_adj) return;
  build_transpose(f->ctx, f->dof_map, f->mesh->K, f->nd, f->n_dof, f->adj_ptr, f->adj);
  f->has_adj = true;
}

// ------------------------------------------------------------------------------------------
// cell colouring: two cells conflict when they share a test dof (they would write the same rows).
// Jones-Plassmann with fixed hashed priorities: the result is a deterministic function of the mesh.
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t cell_priority(uint32_t k) { return k * 2654435761u; }

struct ColorSpace {
  int n;
  const uint32_t* dof_map[kMaxComp];
  const uint32_t* adj_ptr[kMaxComp];
  const uint32_t* adj[kMaxComp];
  int nd[kMaxComp];
};

// reads the colours of the previous round only (color), writes color_out: the outcome of a round does
// not depend on thread timing
__global__ void k_color_round(ColorSpace S, size_t K, const int* __restrict__ color, int* __restrict__ color_out,
                              unsigned long long* __restrict__ remaining, int* __restrict__ overflow) {
  const size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (k >= K || color[k] >= 0) return;
  const uint32_t mine = cell_priority((uint32_t)k);
  unsigned long long used[2] = {0ull, 0ull};
  bool is_max = true;
  for (int c = 0; c < S.n && is_max; ++c)
    for (int i = 0; i < S.nd[c] && is_max; ++i) {
      const uint32_t dof = S.dof_map[c][k * S.nd[c] + i];
      for (uint32_t a = S.adj_ptr[c][dof]; a < S.adj_ptr[c][dof + 1]; ++a) {
        const uint32_t other = S.adj[c][a] / S.nd[c];
        if (other == k) continue;
        const int oc = color[other];
        if (oc >= 0) used[oc >> 6] |= 1ull << (oc & 63);
        else if (cell_priority(other) > mine) { is_max = false; break; }
      }
    }
  if (!is_max) { atomicAdd(remaining, 1ull); return; }
  int pick = -1;
  for (int w = 0; w < 2 && pick < 0; ++w)
    if (~used[w]) pick = w * 64 + __ffsll((long long)~used[w]) - 1;
  if (pick < 0) { atomicExch(overflow, 1); return; }
  color_out[k] = pick;
}

void matrix_build_colors(tfelcu_matrix* m) {
  if (m->has_colors) return;
  Ctx* ctx = m->ctx;
  const size_t K = m->mesh->K;
  ColorSpace S; S.n = m->test.n;
  for (int c = 0; c < m->test.n; ++c) {
    fes_transpose(m->test.fes[c]);
    S.dof_map[c] = m->test.fes[c]->dof_map; S.adj_ptr[c] = m->test.fes[c]->adj_ptr.p; S.adj[c] = m->test.fes[c]->adj.p;
    S.nd[c] = m->test.fes[c]->nd;
  }
  DevBuf<int> color(K), color_next(K); DevBuf<unsigned long long> remaining(1); DevBuf<int> overflow(1);
  TFELCU_CK(cudaMemsetAsync(color.p, 0xff, K * sizeof(int), ctx->stream));
  ctx->zero(overflow.p, 1);
  for (int round = 0; round < 4096; ++round) {
    ctx->zero(remaining.p, 1);
    TFELCU_CK(cudaMemcpyAsync(color_next.p, color.p, K * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    k_color_round<<<grid_for(K, 256), 256, 0, ctx->stream>>>(S, K, color.p, color_next.p, remaining.p, overflow.p);
    ctx->count(); TFELCU_LAUNCH_CHECK();
    std::swap(color.p, color_next.p);
    unsigned long long left = 0; ctx->download(&left, remaining.p, 1);
    if (!left) break;
    TFELCU_REQUIRE(round < 4095, "cell colouring did not converge");
  }
  int ov = 0; ctx->download(&ov, overflow.p, 1);
  TFELCU_REQUIRE(!ov, "cell colouring needs more than 128 colours");
  // group cells by colour (stable: ascending cell id inside a colour)
  DevBuf<uint32_t> ids(K), key_sorted(K);
  iota_u32(ctx, ids.p, K);
  m->color_cells.alloc(K);
  sort_pairs<uint32_t, uint32_t>(ctx, reinterpret_cast<uint32_t*>(color.p), key_sorted.p, ids.p, m->color_cells.p, K, 0, 8);
  DevBuf<uint32_t> starts(130);
  k_lower_bounds_u32<<<1, 256, 0, ctx->stream>>>(key_sorted.p, K, 128, starts.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  uint32_t h[129]; ctx->download(h, starts.p, 129);
  m->color_start.clear();
  m->n_colors = 0;
  for (int c = 0; c < 128; ++c) if (h[c + 1] > h[c]) m->n_colors = c + 1;
  for (int c = 0; c <= m->n_colors; ++c) m->color_start.push_back(h[c]);
  m->has_colors = true;
}

}  // namespace tfelcu
