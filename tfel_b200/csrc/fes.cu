// finite_element_space<fe> on device: DOF numbering, global -> local map, Dirichlet sets.
// Reference: src/core/fes.hpp:18-82 (ctor), :100-149 (add_dirichlet_boundary), :190-202
// (get_dof_space_coordinate). The numbering rule: for every entity kind that carries dofs, in the
// order vertices, edges, faces, cell, dof = offset + rank(entity) * m + i where rank is the position
// of the entity in the std::set ordered by its ascending vertex tuple.
#include "structs.cuh"
#include "sortutil.cuh"

namespace tfelcu {

__global__ void k_dofs_vertex(const uint32_t* __restrict__ cells, const uint32_t* __restrict__ vertex_rank,
                              size_t K, int nv, int nd, uint32_t* __restrict__ dof_map) {
  const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= K * (size_t)nv) return;
  const size_t k = p / nv; const int j = (int)(p % nv);
  dof_map[k * nd + j] = vertex_rank[cells[p]];
}

__global__ void k_dofs_entity(const uint32_t* __restrict__ rank, size_t K, int n_local, int nd, int local_off,
                              uint32_t global_off, uint32_t* __restrict__ dof_map) {
  const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= K * (size_t)n_local) return;
  const size_t k = p / n_local; const int j = (int)(p % n_local);
  dof_map[k * nd + local_off + j] = global_off + rank[p];
}

// fes.hpp:76-81: the last (k, n) in cell-major order wins == the maximum of k * nd + n
__global__ void k_g2l_max(const uint32_t* __restrict__ dof_map, size_t n, unsigned long long* __restrict__ best) {
  const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p < n) atomicMax(best + dof_map[p], (unsigned long long)p);
}

__global__ void k_g2l_split(const unsigned long long* __restrict__ best, size_t n_dof, int nd,
                            uint32_t* __restrict__ g2l) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n_dof) return;
  g2l[2 * i] = (uint32_t)(best[i] / nd);
  g2l[2 * i + 1] = (uint32_t)(best[i] % nd);
}

void fes_build(tfelcu_fes* f) {
  Ctx* ctx = f->ctx;
  tfelcu_mesh* m = f->mesh;
  const int dim = m->dim;
  f->nd = fe_ndof(dim, f->fe);
  mesh_build_vertex_rank(m);

  size_t global_off = 0; int local_off = 0;
  const bool alias_cells = (f->fe == TFELCU_FE_P1 && m->all_vertices_used);
  if (!alias_cells) f->dof_map_store.alloc(m->K * f->nd);
  for (int sd = 0; sd <= dim; ++sd) {
    const int per = fe_dofs_on(dim, f->fe, sd);
    if (!per) continue;
    f->kind_offset[sd] = global_off;
    const int n_local = n_sub_entities(dim, sd);
    if (sd == 0) {
      if (!alias_cells) {
        k_dofs_vertex<<<grid_for(m->K * n_local, 256), 256, 0, ctx->stream>>>(
            m->cells.p, m->vertex_rank.p, m->K, n_local, f->nd, f->dof_map_store.p);
        ctx->count(); TFELCU_LAUNCH_CHECK();
      }
      global_off += m->n_used_vertices;
    } else {
      EntityIndex* E = mesh_entities(m, sd);
      TFELCU_REQUIRE(global_off + E->n_unique < 0xffffffffull, "too many dofs for 32-bit indices");
      k_dofs_entity<<<grid_for(m->K * n_local, 256), 256, 0, ctx->stream>>>(
          E->rank.p, m->K, n_local, f->nd, local_off, (uint32_t)global_off, f->dof_map_store.p);
      ctx->count(); TFELCU_LAUNCH_CHECK();
      global_off += E->n_unique;
    }
    local_off += per * n_local;
  }
  f->n_dof = global_off;
  // P1 on a mesh whose vertices are all used: the dof map IS the connectivity (rank(v) = v)
  f->dof_map = alias_cells ? m->cells.p : f->dof_map_store.p;

  DevBuf<unsigned long long> best(f->n_dof);
  ctx->zero(best.p, f->n_dof);
  const size_t n = m->K * (size_t)f->nd;
  k_g2l_max<<<grid_for(n, 256), 256, 0, ctx->stream>>>(f->dof_map, n, best.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  f->g2l.alloc(f->n_dof * 2);
  k_g2l_split<<<grid_for(f->n_dof, 256), 256, 0, ctx->stream>>>(best.p, f->n_dof, f->nd, f->g2l.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();

  f->dir_flag.alloc(f->n_dof); f->dir_value.alloc(f->n_dof);
  ctx->zero(f->dir_flag.p, f->n_dof);
  ctx->zero(f->dir_value.p, f->n_dof);
  ctx->sync();
}

// ------------------------------------------------------------------------------------------
// Dirichlet candidates of a submesh (fes.hpp:100-122)
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ int cmp_tuple(const uint32_t* a, const uint32_t* b, int n) {
  for (int t = 0; t < n; ++t)
    if (a[t] != b[t]) return a[t] < b[t] ? -1 : 1;
  return 0;
}

// sub-entity s (an m-subset of the nvb boundary-cell vertices, in lexicographic subset order)
__device__ __forceinline__ void subset_vertices(int nvb, int m, int s, int* idx) {
  // enumerate combinations in lexicographic order; nvb <= 3 so this is tiny
  int c[3] = {0, 1, 2};
  for (int it = 0; it < s; ++it) {
    int p = m - 1;
    while (p >= 0 && c[p] == nvb - m + p) --p;
    ++c[p];
    for (int q = p + 1; q < m; ++q) c[q] = c[q - 1] + 1;
  }
  for (int t = 0; t < m; ++t) idx[t] = c[t];
}

__global__ void k_boundary_vertex_dofs(const uint32_t* __restrict__ bcells, size_t n, const uint32_t* __restrict__ vertex_rank,
                                       uint32_t* __restrict__ out) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = vertex_rank[bcells[i]];
}

__global__ void k_boundary_entity_dofs(const uint32_t* __restrict__ bcells, size_t F, int nvb, int m, int n_sub,
                                       const uint32_t* __restrict__ unique_verts, size_t n_unique,
                                       uint32_t global_off, uint32_t* __restrict__ out) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= F * (size_t)n_sub) return;
  const size_t f = i / n_sub; const int s = (int)(i % n_sub);
  int idx[3]; subset_vertices(nvb, m, s, idx);
  uint32_t key[4];
  for (int t = 0; t < m; ++t) key[t] = bcells[f * nvb + idx[t]];
  // std::set::find (fes.hpp:112-113) as a binary search in the sorted unique list
  size_t lo = 0, hi = n_unique;
  while (lo < hi) {
    const size_t mid = lo + (hi - lo) / 2;
    if (cmp_tuple(unique_verts + mid * m, key, m) < 0) lo = mid + 1; else hi = mid;
  }
  const bool found = lo < n_unique && cmp_tuple(unique_verts + lo * m, key, m) == 0;
  out[i] = found ? global_off + (uint32_t)lo : 0xffffffffu;
}

static int n_choose(int n, int k) {
  if (k > n) return 0;
  int r = 1;
  for (int i = 0; i < k; ++i) r = r * (n - i) / (i + 1);
  return r;
}

void fes_boundary_dofs(tfelcu_fes* f, const uint32_t* d_bcells, size_t F, int nvb) {
  Ctx* ctx = f->ctx;
  tfelcu_mesh* m = f->mesh;
  const int dim = m->dim;
  TFELCU_REQUIRE(nvb >= 1 && nvb <= dim, "boundary cells must have between 1 and dim vertices");
  // candidate dof ids (with duplicates), one segment per entity kind < dim that carries dofs
  size_t total = 0;
  for (int sd = 0; sd < dim; ++sd)
    if (fe_dofs_on(dim, f->fe, sd)) total += F * (size_t)n_choose(nvb, sd + 1);
  DevBuf<uint32_t> raw(total), sorted(total);
  size_t at = 0;
  for (int sd = 0; sd < dim; ++sd) {
    if (!fe_dofs_on(dim, f->fe, sd)) continue;
    const int msub = sd + 1, n_sub = n_choose(nvb, msub);
    if (!n_sub || !F) continue;
    if (sd == 0) {
      k_boundary_vertex_dofs<<<grid_for(F * nvb, 256), 256, 0, ctx->stream>>>(d_bcells, F * nvb, m->vertex_rank.p, raw.p + at);
    } else {
      EntityIndex* E = mesh_entities(m, sd);
      k_boundary_entity_dofs<<<grid_for(F * n_sub, 256), 256, 0, ctx->stream>>>(
          d_bcells, F, nvb, msub, n_sub, E->unique_verts.p, E->n_unique, (uint32_t)f->kind_offset[sd], raw.p + at);
    }
    ctx->count(); TFELCU_LAUNCH_CHECK();
    at += F * (size_t)n_sub;
  }
  sort_keys<uint32_t>(ctx, raw.p, sorted.p, at, 0, 32);
  f->cand.alloc(at);
  size_t n = unique_keys<uint32_t>(ctx, sorted.p, f->cand.p, at);
  // entities that are not in the space (0xffffffff) sort last
  if (n) {
    uint32_t last = 0; ctx->download(&last, f->cand.p + (n - 1), 1);
    if (last == 0xffffffffu) --n;
  }
  f->n_cand = n;
}

__global__ void k_dof_coordinates(const uint32_t* __restrict__ dofs, size_t n, const uint32_t* __restrict__ g2l,
                                  const uint32_t* __restrict__ cells, const double* __restrict__ vertices,
                                  const double* __restrict__ nodes, int dim, double* __restrict__ x) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t dof = dofs[i];
  const uint32_t k = g2l[2 * (size_t)dof], ln = g2l[2 * (size_t)dof + 1];
  const uint32_t* cell = cells + (size_t)k * (dim + 1);
  // cell.hpp:456-468 / 783-797: v0 + sum_m xhat_m (v_{m+1} - v0), evaluated left to right without contraction
  for (int c = 0; c < dim; ++c) {
    const double v0 = vertices[(size_t)cell[0] * dim + c];
    double acc = v0;
    for (int mth = 0; mth < dim; ++mth)
      acc = __dadd_rn(acc, __dmul_rn(nodes[ln * dim + mth], __dsub_rn(vertices[(size_t)cell[mth + 1] * dim + c], v0)));
    x[i * dim + c] = acc;
  }
}

void fes_dof_coordinates(const tfelcu_fes* f, const uint32_t* d_dofs, size_t n, double* d_x) {
  Ctx* ctx = f->ctx;
  const int dim = f->mesh->dim;
  double h_nodes[kMaxNd * kMaxDim];
  fe_nodes(dim, f->fe, h_nodes);
  DevBuf<double> nodes((size_t)f->nd * dim);
  ctx->upload(nodes.p, h_nodes, (size_t)f->nd * dim);
  k_dof_coordinates<<<grid_for(n, 256), 256, 0, ctx->stream>>>(d_dofs, n, f->g2l.p, f->mesh->cells.p,
                                                             f->mesh->vertices.p, nodes.p, dim, d_x);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  ctx->sync();
}

__global__ void k_add_dirichlet(const uint32_t* __restrict__ dofs, const double* __restrict__ values, double constant,
                                size_t n, size_t n_dof, uint8_t* __restrict__ flag, double* __restrict__ value,
                                unsigned long long* __restrict__ n_new) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t d = dofs[i];
  if (d >= n_dof || flag[d]) return;     // std::map::insert keeps the first value (fes.hpp:116)
  flag[d] = 1;
  value[d] = values ? values[i] : constant;
  atomicAdd(n_new, 1ull);
}

void fes_add_dirichlet(tfelcu_fes* f, const uint32_t* d_dofs, const double* d_values, double constant, size_t n) {
  Ctx* ctx = f->ctx;
  if (!n) return;
  DevBuf<unsigned long long> n_new(1);
  ctx->zero(n_new.p, 1);
  k_add_dirichlet<<<grid_for(n, 256), 256, 0, ctx->stream>>>(d_dofs, d_values, constant, n, f->n_dof,
                                                           f->dir_flag.p, f->dir_value.p, n_new.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  unsigned long long h = 0; ctx->download(&h, n_new.p, 1);
  f->n_dirichlet += (size_t)h;
  if (h) ++f->dirichlet_version;
}

__global__ void k_flag_to_u32(const uint8_t* __restrict__ flag, size_t n, uint32_t* __restrict__ out) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = flag[i];
}

__global__ void k_compact_dirichlet(const uint8_t* __restrict__ flag, const uint32_t* __restrict__ pos,
                                    const double* __restrict__ value, size_t n, uint32_t* __restrict__ dofs,
                                    double* __restrict__ values) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n || !flag[i]) return;
  dofs[pos[i]] = (uint32_t)i;
  values[pos[i]] = value[i];
}

void fes_dirichlet_get(const tfelcu_fes* f, uint32_t* h_dofs, double* h_values) {
  Ctx* ctx = f->ctx;
  if (!f->n_dirichlet) return;
  DevBuf<uint32_t> f32(f->n_dof), pos(f->n_dof), dofs(f->n_dirichlet);
  DevBuf<double> values(f->n_dirichlet);
  k_flag_to_u32<<<grid_for(f->n_dof, 256), 256, 0, ctx->stream>>>(f->dir_flag.p, f->n_dof, f32.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  exclusive_sum(ctx, f32.p, pos.p, f->n_dof);
  k_compact_dirichlet<<<grid_for(f->n_dof, 256), 256, 0, ctx->stream>>>(f->dir_flag.p, pos.p, f->dir_value.p,
                                                                      f->n_dof, dofs.p, values.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  if (h_dofs) ctx->download(h_dofs, dofs.p, f->n_dirichlet);
  if (h_values) ctx->download(h_values, values.p, f->n_dirichlet);
}

// ------------------------------------------------------------------------------------------
// output side: to_mesh_vertex_data / to_mesh_cell_data (fes.hpp:464-504)
// ------------------------------------------------------------------------------------------

// values[cells[k][n]] = coefficients[dof(k, n)] for the vertex dofs n < dim + 1 of every cell (fes.hpp:498-501; all
// writers of a vertex store the same number for a continuous element)
__global__ void k_vertex_values(const uint32_t* __restrict__ cells, const uint32_t* __restrict__ dof_map, int nv, int nd,
                                size_t K, const double* __restrict__ coef, double* __restrict__ out) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= K * (size_t)nv) return;
  const size_t k = t / nv; const int n = (int)(t - k * nv);
  out[cells[k * nv + n]] = coef[dof_map[k * nd + n]];
}

struct BasisAt { double phi[kMaxNd]; };

// values[k] = sum_i coefficients[dof(k, i)] phi_i(barycentre) (fes.hpp:476-479: element::evaluate at cell_type::barycenter())
__global__ void k_cell_values(const uint32_t* __restrict__ dof_map, int nd, size_t K, BasisAt B,
                              const double* __restrict__ coef, double* __restrict__ out) {
  const size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (k >= K) return;
  double v = 0.0;
  for (int i = 0; i < nd; ++i) v = __dadd_rn(v, __dmul_rn(coef[dof_map[k * nd + i]], B.phi[i]));
  out[k] = v;
}

void fes_vertex_values(const tfelcu_fes* f, const double* d_coef, double* d_out) {
  Ctx* ctx = f->ctx;
  const tfelcu_mesh* m = f->mesh;
  ctx->zero(d_out, m->V);
  const int nv = m->dim + 1;
  const size_t n = m->K * (size_t)nv;
  if (!n) return;
  k_vertex_values<<<grid_for(n, 256), 256, 0, ctx->stream>>>(m->cells.p, f->dof_map, nv, f->nd, m->K, d_coef, d_out);
  ctx->count(); TFELCU_LAUNCH_CHECK();
}

void fes_cell_values(const tfelcu_fes* f, const double* d_coef, double* d_out) {
  Ctx* ctx = f->ctx;
  const tfelcu_mesh* m = f->mesh;
  if (!m->K) return;
  const int dim = m->dim, w = dim + 1;
  double bc[kMaxDim];
  for (int d = 0; d < dim; ++d) bc[d] = 1.0 / (dim + 1);
  double hat[kMaxNd * (kMaxDim + 1)];
  fe_eval_hat(dim, f->fe, bc, hat);
  BasisAt B;
  for (int i = 0; i < kMaxNd; ++i) B.phi[i] = i < f->nd ? hat[i * w] : 0.0;
  k_cell_values<<<grid_for(m->K, 256), 256, 0, ctx->stream>>>(f->dof_map, f->nd, m->K, B, d_coef, d_out);
  ctx->count(); TFELCU_LAUNCH_CHECK();
}

}  // namespace tfelcu
