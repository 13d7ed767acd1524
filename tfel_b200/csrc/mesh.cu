// Mesh on device: upload / generators, per-cell geometry, sorted entity indices and the boundary
// submesh. Reference: src/core/mesh.hpp:142-158 (mesh ctor), :286-299 (admissibility),
// :485-521 (get_boundary_submesh), src/core/mesh.cpp:19-135 (generators),
// src/core/cell.hpp:383-435, 692-764 (entity lists as std::set), src/core/subdomain.hpp:36-43 (order).
#include <cub/cub.cuh>

#include "structs.cuh"
#include "geometry.cuh"
#include "sortutil.cuh"

namespace tfelcu {

// ------------------------------------------------------------------------------------------
// generators
// ------------------------------------------------------------------------------------------

__global__ void k_square_vertices(double* __restrict__ vertices, double x1, double x2, uint32_t n1, uint32_t n2) {
  const size_t id = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t V = (size_t)(n1 + 1) * (n2 + 1);
  if (id >= V) return;
  const uint32_t i = id % (n1 + 1), j = id / (n1 + 1);
  // mesh.cpp:26-27: x_1 / n_1 * i  (division first)
  vertices[2 * id] = __dmul_rn(__ddiv_rn(x1, (double)n1), (double)i);
  vertices[2 * id + 1] = __dmul_rn(__ddiv_rn(x2, (double)n2), (double)j);
}

__global__ void k_square_cells(uint32_t* __restrict__ cells, uint32_t n1, uint32_t n2) {
  const size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (q >= (size_t)n1 * n2) return;
  const uint32_t i = q % n1, j = q / n1;
  const uint32_t v00 = j * (n1 + 1) + i, v10 = v00 + 1, v01 = v00 + n1 + 1, v11 = v01 + 1;
  // mesh.cpp:41-47: "second diagonal": (v00, v10, v11), (v00, v01, v11)
  uint32_t* e = cells + 6 * q;
  e[0] = v00; e[1] = v10; e[2] = v11;
  e[3] = v00; e[4] = v01; e[5] = v11;
}

__global__ void k_cube_vertices(double* __restrict__ vertices, double hx, double hy, double hz,
                                uint32_t n1, uint32_t n2, uint32_t n3) {
  const size_t id = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t V = (size_t)(n1 + 1) * (n2 + 1) * (n3 + 1);
  if (id >= V) return;
  const uint32_t i = id % (n1 + 1), j = (id / (n1 + 1)) % (n2 + 1), k = id / ((size_t)(n1 + 1) * (n2 + 1));
  vertices[3 * id] = __dmul_rn((double)i, hx);     // mesh.cpp:74-76
  vertices[3 * id + 1] = __dmul_rn((double)j, hy);
  vertices[3 * id + 2] = __dmul_rn((double)k, hz);
}

// corner c = 4 dk + 2 dj + di of the unit cube; five tetrahedra, two orientations (mesh.cpp:82-131)
__constant__ int c_cube_even[5][4] = {{0, 1, 2, 4}, {1, 4, 5, 7}, {1, 2, 3, 7}, {2, 4, 6, 7}, {1, 2, 4, 7}};
__constant__ int c_cube_odd[5][4] = {{0, 1, 3, 5}, {0, 2, 3, 6}, {0, 4, 5, 6}, {3, 5, 6, 7}, {0, 3, 5, 6}};

__global__ void k_cube_cells(uint32_t* __restrict__ cells, uint32_t n1, uint32_t n2, uint32_t n3) {
  const size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (q >= (size_t)n1 * n2 * n3) return;
  const uint32_t i = q % n1, j = (q / n1) % n2, k = q / ((size_t)n1 * n2);
  const bool even = ((i + j + k) % 2) == 0;
  uint32_t* e = cells + 20 * q;
  for (int t = 0; t < 5; ++t)
    for (int v = 0; v < 4; ++v) {
      const int c = even ? c_cube_even[t][v] : c_cube_odd[t][v];
      const uint32_t kk = k + ((c >> 2) & 1), jj = j + ((c >> 1) & 1), ii = i + (c & 1);
      e[4 * t + v] = (uint32_t)(((size_t)kk * (n2 + 1) + jj) * (n1 + 1) + ii);
    }
}

// ------------------------------------------------------------------------------------------
// admissibility: every cell lists its vertices in ascending order (mesh.hpp:286-299)
// ------------------------------------------------------------------------------------------

__global__ void k_check_cells(const uint32_t* __restrict__ cells, size_t K, int nv, int* __restrict__ bad) {
  const size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (k >= K) return;
  bool ok = true;
  for (int i = 0; i + 1 < nv; ++i) ok = ok && (cells[k * nv + i] < cells[k * nv + i + 1]);
  if (!ok) atomicOr(bad, 1);
}

__global__ void k_sort_cells_like_reference(uint32_t* __restrict__ cells, size_t K, int nv) {
  // mesh.hpp:294-298 sorts [begin, begin + nv - 1): the first nv - 1 ids only
  const size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (k >= K) return;
  uint32_t* c = cells + k * nv;
  for (int i = 1; i < nv - 1; ++i) {
    const uint32_t x = c[i];
    int j = i - 1;
    while (j >= 0 && c[j] > x) { c[j + 1] = c[j]; --j; }
    c[j + 1] = x;
  }
}

void mesh_validate_cells(tfelcu_mesh* m) {
  Ctx* ctx = m->ctx;
  const int nv = m->dim + 1;
  DevBuf<int> bad(1);
  for (int attempt = 0; attempt < 2; ++attempt) {
    ctx->zero(bad.p, 1);
    k_check_cells<<<grid_for(m->K, 256), 256, 0, ctx->stream>>>(m->cells.p, m->K, nv, bad.p);
    ctx->count(); TFELCU_LAUNCH_CHECK();
    int h = 0; ctx->download(&h, bad.p, 1);
    if (!h) return;
    if (attempt == 1) fail("nodes must be ordered when inserted into a subdomain");  // subdomain.hpp:11-13
    k_sort_cells_like_reference<<<grid_for(m->K, 256), 256, 0, ctx->stream>>>(m->cells.p, m->K, nv);
    ctx->count(); TFELCU_LAUNCH_CHECK();
  }
}

// ------------------------------------------------------------------------------------------
// vertices used by cells and their ranks (get_vertex_list, cell.hpp:393-405)
// ------------------------------------------------------------------------------------------

__global__ void k_mark_used(const uint32_t* __restrict__ cells, size_t n, uint32_t* __restrict__ used) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) used[cells[i]] = 1u;
}

void mesh_build_vertex_rank(tfelcu_mesh* m) {
  if (m->vertex_rank.n) return;
  Ctx* ctx = m->ctx;
  DevBuf<uint32_t> used(m->V);
  ctx->zero(used.p, m->V);
  const size_t n = m->K * (m->dim + 1);
  k_mark_used<<<grid_for(n, 256), 256, 0, ctx->stream>>>(m->cells.p, n, used.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  m->vertex_rank.alloc(m->V);
  exclusive_sum(ctx, used.p, m->vertex_rank.p, m->V);
  uint32_t last_rank = 0, last_used = 0;
  if (m->V) {
    ctx->download(&last_rank, m->vertex_rank.p + (m->V - 1), 1);
    ctx->download(&last_used, used.p + (m->V - 1), 1);
  }
  m->n_used_vertices = (size_t)last_rank + last_used;
  m->all_vertices_used = (m->n_used_vertices == m->V);
}

// ------------------------------------------------------------------------------------------
// entity index of kind sd >= 1
// ------------------------------------------------------------------------------------------

__constant__ int c_tri_edge[3][2] = {{0, 1}, {0, 2}, {1, 2}};                                   // cell.hpp:362-373
__constant__ int c_tet_edge[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};          // cell.hpp:646-663
__constant__ int c_tet_face[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};              // cell.hpp:665-681

__device__ __forceinline__ int local_entity_vertex(int dim, int sd, int j, int t) {
  if (sd == dim) return t;
  if (sd == 1) return dim == 2 ? c_tri_edge[j][t] : c_tet_edge[j][t];
  return c_tet_face[j][t];
}

// hi / lo 64-bit words of the ascending vertex tuple of entity p = k * n_local + j
__global__ void k_entity_words(const uint32_t* __restrict__ cells, size_t K, int dim, int sd, int n_local,
                               uint64_t* __restrict__ hi, uint64_t* __restrict__ lo) {
  const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= K * (size_t)n_local) return;
  const size_t k = p / n_local; const int j = (int)(p % n_local);
  const int m = sd + 1;
  uint32_t v[4] = {0, 0, 0, 0};
  for (int t = 0; t < m; ++t) v[t] = cells[k * (dim + 1) + local_entity_vertex(dim, sd, j, t)];
  // tuple (v0, v1, v2, v3) padded with zeros compares like the reference's lexicographic order
  hi[p] = ((uint64_t)v[0] << 32) | v[1];
  if (lo) lo[p] = ((uint64_t)v[2] << 32) | v[3];
}

__global__ void k_gather_u64(const uint64_t* __restrict__ src, const uint32_t* __restrict__ idx, size_t n,
                             uint64_t* __restrict__ dst) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

__global__ void k_entity_heads(const uint64_t* __restrict__ hi, const uint64_t* __restrict__ lo,
                               const uint32_t* __restrict__ idx, size_t n, uint32_t* __restrict__ head) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool h = (i == 0);
  if (!h) {
    const uint32_t a = idx[i], b = idx[i - 1];
    h = hi[a] != hi[b] || (lo && lo[a] != lo[b]);
  }
  head[i] = h ? 1u : 0u;
}

__global__ void k_entity_finish(const uint64_t* __restrict__ hi, const uint64_t* __restrict__ lo,
                                const uint32_t* __restrict__ idx, const uint32_t* __restrict__ head,
                                const uint32_t* __restrict__ uid_incl, size_t n, int n_vert,
                                uint32_t* __restrict__ rank, uint32_t* __restrict__ unique_verts,
                                uint32_t* __restrict__ first, uint32_t* __restrict__ head_pos) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t u = uid_incl[i] - 1u, p = idx[i];
  rank[p] = u;
  if (head[i]) {
    const uint64_t h = hi[p], l = lo ? lo[p] : 0ull;
    const uint32_t v[4] = {(uint32_t)(h >> 32), (uint32_t)h, (uint32_t)(l >> 32), (uint32_t)l};
    for (int t = 0; t < n_vert; ++t) unique_verts[(size_t)u * n_vert + t] = v[t];
    first[u] = p;
    head_pos[u] = (uint32_t)i;
  }
}

__global__ void k_run_lengths(const uint32_t* __restrict__ head_pos, size_t n_unique, size_t n,
                              uint32_t* __restrict__ count) {
  const size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (u >= n_unique) return;
  const uint32_t end = (u + 1 < n_unique) ? head_pos[u + 1] : (uint32_t)n;
  count[u] = end - head_pos[u];
}

EntityIndex* mesh_entities(tfelcu_mesh* m, int sd) {
  TFELCU_REQUIRE(sd >= 1 && sd <= m->dim, "entity kind out of range");
  if (m->entities[sd]) return m->entities[sd].get();
  Ctx* ctx = m->ctx;
  std::unique_ptr<EntityIndex> E(new EntityIndex());
  const int n_local = n_sub_entities(m->dim, sd);
  const int n_vert = sd + 1;
  const size_t n = m->K * (size_t)n_local;
  TFELCU_REQUIRE(n < 0xffffffffull, "too many cell entities for 32-bit indices");
  E->n_vert = n_vert;
  E->rank.alloc(n);

  DevBuf<uint64_t> hi(n), lo;
  if (n_vert > 2) lo.alloc(n);
  k_entity_words<<<grid_for(n, 256), 256, 0, ctx->stream>>>(m->cells.p, m->K, m->dim, sd, n_local, hi.p,
                                                          n_vert > 2 ? lo.p : nullptr);
  ctx->count(); TFELCU_LAUNCH_CHECK();

  // LSD radix passes on (lo, hi); the payload starts as the identity so that ties keep ascending p
  DevBuf<uint32_t> idx(n), idx2(n);
  DevBuf<uint64_t> key(n), key2(n);
  iota_u32(ctx, idx.p, n);
  const int vbits = bits_for(m->V);
  if (n_vert > 2) {
    TFELCU_CK(cudaMemcpyAsync(key.p, lo.p, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    sort_pairs_u64_u32(ctx, key.p, key2.p, idx.p, idx2.p, n, n_vert == 3 ? 32 : 0, n_vert == 3 ? 32 + vbits : 32 + vbits);
    std::swap(idx.p, idx2.p);
    k_gather_u64<<<grid_for(n, 256), 256, 0, ctx->stream>>>(hi.p, idx.p, n, key.p);
    ctx->count(); TFELCU_LAUNCH_CHECK();
  } else {
    TFELCU_CK(cudaMemcpyAsync(key.p, hi.p, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  sort_pairs_u64_u32(ctx, key.p, key2.p, idx.p, idx2.p, n, 0, 32 + vbits);
  std::swap(idx.p, idx2.p);

  DevBuf<uint32_t> head(n), uid(n);
  k_entity_heads<<<grid_for(n, 256), 256, 0, ctx->stream>>>(hi.p, n_vert > 2 ? lo.p : nullptr, idx.p, n, head.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  inclusive_sum(ctx, head.p, uid.p, n);
  uint32_t n_unique = 0;
  if (n) ctx->download(&n_unique, uid.p + (n - 1), 1);
  E->n_unique = n_unique;
  E->unique_verts.alloc((size_t)n_unique * n_vert);
  E->first.alloc(n_unique);
  E->count.alloc(n_unique);
  DevBuf<uint32_t> head_pos(n_unique);
  k_entity_finish<<<grid_for(n, 256), 256, 0, ctx->stream>>>(hi.p, n_vert > 2 ? lo.p : nullptr, idx.p, head.p, uid.p,
                                                           n, n_vert, E->rank.p, E->unique_verts.p, E->first.p,
                                                           head_pos.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  k_run_lengths<<<grid_for(n_unique, 256), 256, 0, ctx->stream>>>(head_pos.p, n_unique, n, E->count.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  ctx->sync();
  m->entities[sd] = std::move(E);
  return m->entities[sd].get();
}

// ------------------------------------------------------------------------------------------
// boundary submesh: faces with exactly one incident cell, in ascending face order (mesh.hpp:485-521)
// ------------------------------------------------------------------------------------------

__global__ void k_is_boundary(const uint32_t* __restrict__ count, size_t n, uint32_t* __restrict__ flag) {
  const size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (u < n) flag[u] = count[u] == 1u ? 1u : 0u;
}

__global__ void k_boundary_compact(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos,
                                   const uint32_t* __restrict__ unique_verts, const uint32_t* __restrict__ first,
                                   size_t n, int n_vert, int n_local, uint32_t* __restrict__ el,
                                   uint32_t* __restrict__ parent_cell, uint32_t* __restrict__ parent_sd) {
  const size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (u >= n || !flag[u]) return;
  const uint32_t f = pos[u];
  for (int t = 0; t < n_vert; ++t) el[(size_t)f * n_vert + t] = unique_verts[u * n_vert + t];
  parent_cell[f] = first[u] / n_local;
  parent_sd[f] = first[u] % n_local;
}

void mesh_build_boundary(tfelcu_mesh* m) {
  if (m->has_boundary) return;
  Ctx* ctx = m->ctx;
  const int sd = m->dim - 1;
  EntityIndex* E = mesh_entities(m, sd);
  const size_t n = E->n_unique;
  DevBuf<uint32_t> flag(n), pos(n);
  k_is_boundary<<<grid_for(n, 256), 256, 0, ctx->stream>>>(E->count.p, n, flag.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  exclusive_sum(ctx, flag.p, pos.p, n);
  uint32_t last_pos = 0, last_flag = 0;
  if (n) { ctx->download(&last_pos, pos.p + (n - 1), 1); ctx->download(&last_flag, flag.p + (n - 1), 1); }
  m->F = (size_t)last_pos + last_flag;
  m->b_el.alloc(m->F * m->dim); m->b_cell.alloc(m->F); m->b_sd.alloc(m->F);
  k_boundary_compact<<<grid_for(n, 256), 256, 0, ctx->stream>>>(flag.p, pos.p, E->unique_verts.p, E->first.p, n,
                                                              E->n_vert, n_sub_entities(m->dim, sd),
                                                              m->b_el.p, m->b_cell.p, m->b_sd.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  ctx->sync();
  m->has_boundary = true;
}

// ------------------------------------------------------------------------------------------
// face-neighbour table: mesh<cell>::compute_cell_neighbours (mesh.hpp:301-332)
// ------------------------------------------------------------------------------------------

// largest k * n_local + j among the (cell, local face) pairs of every face
__global__ void k_face_last(const uint32_t* __restrict__ rank, size_t n, uint32_t* __restrict__ last) {
  const size_t code = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (code < n) atomicMax(last + rank[code], (uint32_t)code);
}

// The reference walks the cells in ascending order, remembers the first (cell, local face) that meets a face, links every
// later cell to that first one and the first one to the later cell (so the last one stays): first <-> last, others -> first.
__global__ void k_neighbours(const uint32_t* __restrict__ rank, const uint32_t* __restrict__ first,
                             const uint32_t* __restrict__ last, size_t n, int n_local, int32_t* __restrict__ nb) {
  const size_t code = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (code >= n) return;
  const uint32_t u = rank[code], f = first[u], l = last[u];
  nb[code] = (f == l) ? -1 : (int32_t)(((uint32_t)code == f ? l : f) / (uint32_t)n_local);
}

void mesh_neighbours(tfelcu_mesh* m, int32_t* d_nb) {
  Ctx* ctx = m->ctx;
  const int sd = m->dim - 1;
  EntityIndex* E = mesh_entities(m, sd);
  const int n_local = n_sub_entities(m->dim, sd);   // == dim + 1 faces per simplex
  const size_t n = m->K * (size_t)n_local;
  if (!n) return;
  DevBuf<uint32_t> last(E->n_unique);
  ctx->zero(last.p, E->n_unique);
  k_face_last<<<grid_for(n, 256), 256, 0, ctx->stream>>>(E->rank.p, n, last.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  k_neighbours<<<grid_for(n, 256), 256, 0, ctx->stream>>>(E->rank.p, E->first.p, last.p, n, n_local, d_nb);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  ctx->sync();
}

// ------------------------------------------------------------------------------------------
// geometry download (parity of a3)
// ------------------------------------------------------------------------------------------

template <int DIM>
__global__ void k_geometry(const double* __restrict__ vertices, const uint32_t* __restrict__ cells, size_t K,
                           double* __restrict__ vol, double* __restrict__ jmt) {
  const size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (k >= K) return;
  CellGeom<DIM> g;
  load_cell_geometry<DIM>(vertices, cells + k * (DIM + 1), g);
  vol[k] = g.vol;
  for (int s = 0; s < DIM; ++s)
    for (int t = 0; t < DIM; ++t) jmt[(k * DIM + s) * DIM + t] = g.jmt[s][t];
}

void mesh_geometry(const tfelcu_mesh* m, double* vol, double* jmt) {
  Ctx* ctx = m->ctx;
  DevBuf<double> d_vol(m->K), d_jmt(m->K * m->dim * m->dim);
  if (m->dim == 2)
    k_geometry<2><<<grid_for(m->K, 256), 256, 0, ctx->stream>>>(m->vertices.p, m->cells.p, m->K, d_vol.p, d_jmt.p);
  else
    k_geometry<3><<<grid_for(m->K, 256), 256, 0, ctx->stream>>>(m->vertices.p, m->cells.p, m->K, d_vol.p, d_jmt.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  ctx->download(vol, d_vol.p, m->K);
  ctx->download(jmt, d_jmt.p, m->K * m->dim * m->dim);
}

void mesh_generate_square(tfelcu_mesh* m, double x1, double x2, uint32_t n1, uint32_t n2) {
  Ctx* ctx = m->ctx;
  m->dim = 2; m->V = (size_t)(n1 + 1) * (n2 + 1); m->K = 2 * (size_t)n1 * n2;
  TFELCU_REQUIRE(m->V < 0xffffffffull, "too many vertices");
  m->vertices.alloc(m->V * 2); m->cells.alloc(m->K * 3);
  k_square_vertices<<<grid_for(m->V, 256), 256, 0, ctx->stream>>>(m->vertices.p, x1, x2, n1, n2);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  k_square_cells<<<grid_for((size_t)n1 * n2, 256), 256, 0, ctx->stream>>>(m->cells.p, n1, n2);
  ctx->count(); TFELCU_LAUNCH_CHECK();
}

void mesh_generate_cube(tfelcu_mesh* m, double x1, double x2, double x3, uint32_t n1, uint32_t n2, uint32_t n3) {
  Ctx* ctx = m->ctx;
  m->dim = 3; m->V = (size_t)(n1 + 1) * (n2 + 1) * (n3 + 1); m->K = 5 * (size_t)n1 * n2 * n3;
  TFELCU_REQUIRE(m->V < 0xffffffffull, "too many vertices");
  m->vertices.alloc(m->V * 3); m->cells.alloc(m->K * 4);
  const double hx = x1 / (double)n1, hy = x2 / (double)n2, hz = x3 / (double)n3;   // mesh.cpp:60-63
  k_cube_vertices<<<grid_for(m->V, 256), 256, 0, ctx->stream>>>(m->vertices.p, hx, hy, hz, n1, n2, n3);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  k_cube_cells<<<grid_for((size_t)n1 * n2 * n3, 256), 256, 0, ctx->stream>>>(m->cells.p, n1, n2, n3);
  ctx->count(); TFELCU_LAUNCH_CHECK();
}

}  // namespace tfelcu
