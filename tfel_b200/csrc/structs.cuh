// Device-resident objects behind the opaque handles of include/tfel_cuda.h.
#pragma once

#include "common.cuh"
#include "tables.hpp"

namespace tfelcu {

// Sorted unique entities of one kind (vertices of an edge / face / cell as ascending id tuples),
// the device counterpart of cell::get_subdomain_list (cell.hpp:383-435, 692-764).
struct EntityIndex {
  int n_vert = 0;                 // vertices per entity
  size_t n_unique = 0;
  DevBuf<uint32_t> rank;          // [K * n_local]: rank of local entity j of cell k in the sorted unique list
  DevBuf<uint32_t> unique_verts;  // [n_unique][n_vert], ascending (size, lexicographic) order of subdomain.hpp:36-43
  DevBuf<uint32_t> count;         // [n_unique]: incidence count
  DevBuf<uint32_t> first;         // [n_unique]: smallest k * n_local + j among the incident (cell, local) pairs
};

// Patches of rows for the patch assembly (patch.cu): the dofs of one owned row range sorted along a Morton curve of
// their space coordinates and cut into groups of rows_per_patch; for each patch the ascending list of the cells
// incident to its rows, and the transposed dof map re-expressed in patch-local cell indices.
struct PatchPlan {
  size_t row_begin = 0, row_end = 0;   // dofs of the component covered by the plan
  int rows_per_patch = 0;              // capacity of a patch
  size_t n_patch = 0;
  size_t a_begin = 0, a_count = 0;     // slice of the transposed dof map of [row_begin, row_end)
  DevBuf<uint32_t> row_ptr;            // [n_patch + 1]: rows of patch p are rows[row_ptr[p] .. row_ptr[p + 1])
  DevBuf<uint32_t> rows;               // [row_end - row_begin]: dofs in patch order
  DevBuf<uint32_t> cell_ptr;           // [n_patch + 1]
  DevBuf<uint32_t> cells;              // cells of each patch, ascending
  DevBuf<uint32_t> adj_local;          // [a_count]: (index of the cell in its patch's list) * nd + local dof
  size_t max_cells = 0;                // largest patch, in cells
};

struct TilePlan;                         // tile.cu
struct FormBuild;                        // element.cuh
void tile_plan_free(TilePlan* p);
void form_build_free(FormBuild* p);

}  // namespace tfelcu

struct tfelcu_mesh {
  tfelcu::Ctx* ctx = nullptr;
  int dim = 0;
  size_t V = 0, K = 0;
  tfelcu::DevBuf<double> vertices;    // [V][dim]
  tfelcu::DevBuf<uint32_t> cells;     // [K][dim + 1]
  // boundary submesh (mesh.hpp:485-521), built on demand
  bool has_boundary = false;
  size_t F = 0;
  tfelcu::DevBuf<uint32_t> b_el, b_cell, b_sd;   // [F][dim], [F], [F]
  // cached entity indices per kind (1: edges, dim-1: faces, dim: cells); kind 0 uses vertex_rank
  std::unique_ptr<tfelcu::EntityIndex> entities[4];
  tfelcu::DevBuf<uint32_t> vertex_rank;  // [V]: rank of a vertex among the vertices used by cells
  size_t n_used_vertices = 0;
  bool all_vertices_used = false;
};

struct tfelcu_fes {
  tfelcu::Ctx* ctx = nullptr;
  tfelcu_mesh* mesh = nullptr;
  int fe = 0, nd = 0;                 // dofs per cell
  size_t n_dof = 0;
  size_t kind_offset[4] = {0, 0, 0, 0};  // global dof offset of each entity kind that carries dofs
  tfelcu::DevBuf<uint32_t> dof_map_store;  // [K][nd] unless aliased to mesh cells
  const uint32_t* dof_map = nullptr;       // device
  tfelcu::DevBuf<uint32_t> g2l;            // [n_dof][2], fes.hpp:76-81 (last (k, n) wins)
  tfelcu::DevBuf<uint8_t> dir_flag;        // [n_dof]
  tfelcu::DevBuf<double> dir_value;        // [n_dof]
  size_t n_dirichlet = 0;
  uint64_t dirichlet_version = 0;          // bumped whenever the Dirichlet set changes
  // result of the last tfelcu_fes_boundary_dofs call
  tfelcu::DevBuf<uint32_t> cand;
  size_t n_cand = 0;
  // transposed dof map, built on demand: adj_ptr[r] .. adj_ptr[r + 1] lists k * nd + i (ascending)
  // for every (cell k, local dof i) with dof_map[k][i] == r
  bool has_adj = false;
  tfelcu::DevBuf<uint32_t> adj_ptr, adj;
  std::vector<std::unique_ptr<tfelcu::PatchPlan>> patch_plans;   // built on demand, one per (row range, patch size)
};

namespace tfelcu {

constexpr int kMaxComp = 4;

struct SpaceList {
  int n = 0;
  tfelcu_fes* fes[kMaxComp] = {nullptr, nullptr, nullptr, nullptr};
  size_t offset[kMaxComp + 1] = {0, 0, 0, 0, 0};
  size_t total() const { return offset[n]; }
};

// Rows a process owns: up to kMaxSeg contiguous segments [g_begin, g_end) of GLOBAL rows (block layout
// [u0 | u1 | p]), ascending and disjoint, none straddling two components -- e.g. the vertex dofs and the edge dofs of one
// horizontal strip for every P2 component. Local rows are the segments concatenated. A single process owns one segment
// per component covering everything.
constexpr int kMaxSeg = 8;

struct RowMap {
  int n = 0;
  unsigned long long g_begin[kMaxSeg] = {0, 0, 0, 0, 0, 0, 0, 0};
  unsigned long long g_end[kMaxSeg] = {0, 0, 0, 0, 0, 0, 0, 0};
  unsigned long long local_off[kMaxSeg + 1] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // local row of the first row of a segment
  int comp[kMaxSeg] = {0, 0, 0, 0, 0, 0, 0, 0};                              // component the segment lies in
  unsigned long long n_global = 0;
  size_t n_local() const { return (size_t)local_off[n]; }
  bool whole() const { return local_off[n] == n_global; }
  // global row of a local row
  __host__ __device__ unsigned long long to_global(unsigned long long lr) const {
    int sg = 0;
    while (sg + 1 < n && lr >= local_off[sg + 1]) ++sg;
    return g_begin[sg] + (lr - local_off[sg]);
  }
  // local row of a global row, or ~0 when another process owns it
  __host__ __device__ unsigned long long to_local(unsigned long long g) const {
    for (int sg = 0; sg < n; ++sg)
      if (g >= g_begin[sg] && g < g_end[sg]) return local_off[sg] + (g - g_begin[sg]);
    return ~0ull;
  }
};

// seg_begin == nullptr: everything
inline RowMap make_row_map(const SpaceList& L, int n_seg, const size_t* seg_begin, const size_t* seg_end) {
  RowMap R;
  R.n_global = L.offset[L.n];
  if (!seg_begin) {
    R.n = L.n;
    for (int c = 0; c < L.n; ++c) {
      R.g_begin[c] = L.offset[c]; R.g_end[c] = L.offset[c + 1]; R.local_off[c] = L.offset[c]; R.comp[c] = c;
    }
    R.local_off[L.n] = L.offset[L.n];
    return R;
  }
  TFELCU_REQUIRE(n_seg >= 0 && n_seg <= kMaxSeg, "at most 8 row segments per process");
  R.n = n_seg;
  unsigned long long off = 0, prev_end = 0;
  for (int sg = 0; sg < n_seg; ++sg) {
    const unsigned long long b = seg_begin[sg], e = seg_end[sg];
    TFELCU_REQUIRE(b <= e && e <= R.n_global && b >= prev_end, "row segments must be ascending, disjoint and in range");
    int c = 0;
    while (c + 1 < L.n && b >= L.offset[c + 1]) ++c;
    TFELCU_REQUIRE(e <= L.offset[c + 1] || b == e, "a row segment must lie inside one component");
    R.g_begin[sg] = b; R.g_end[sg] = e; R.local_off[sg] = off; R.comp[sg] = c;
    off += e - b; prev_end = e;
  }
  R.local_off[n_seg] = off;
  return R;
}

}  // namespace tfelcu

struct tfelcu_vector {
  tfelcu::Ctx* ctx = nullptr;
  tfelcu::SpaceList space;
  tfelcu::RowMap rows;                  // owned rows; data holds rows.n_local() entries
  tfelcu::DevBuf<double> data;
  int scatter = TFELCU_SCATTER_AUTO;
  bool patch_refused = false;
  tfelcu::FormBuild* memo_form = nullptr;   // like tfelcu_matrix::memo_form, for the linear form that rides along
  std::vector<unsigned char> memo_key;
  ~tfelcu_vector() { if (memo_form) tfelcu::form_build_free(memo_form); }
  bool zero_pending = false;            // clear() is pending: the next writer starts from zero, every other user zeroes first
  void ensure_cleared() {
    if (!zero_pending) return;
    zero_pending = false;
    ctx->zero(data.p, data.n);
  }
};

struct tfelcu_matrix {
  tfelcu::Ctx* ctx = nullptr;
  tfelcu_mesh* mesh = nullptr;
  tfelcu::SpaceList test, trial;
  size_t n_rows = 0, n_cols = 0, nnz = 0;   // global sizes; nnz counts the stored entries of the OWNED rows
  tfelcu::RowMap rows;                  // owned rows: rowptr has rows.n_local() + 1 entries, colind holds global columns
  tfelcu::DevBuf<uint8_t> dir_row;      // [n_rows]: snapshot of the test spaces' Dirichlet flags
  tfelcu::DevBuf<double> dir_value;     // [n_rows]
  uint64_t dir_version[tfelcu::kMaxComp] = {0, 0, 0, 0};
  tfelcu::DevBuf<int32_t> rowptr, colind;
  tfelcu::DevBuf<double> val;
  uint64_t pattern_id = 0;              // unique per pattern build: solvers recognise an unchanged pattern by it
  bool pristine = true;                 // no += since clear()
  int scatter = TFELCU_SCATTER_AUTO;
  bool patch_refused = false;           // AUTO: the patch strategy did not apply to this form once; stay on rows
  bool val_valid = false;                // false: clear() is pending (performed lazily by the first user)
  int rows_per_cta = 128;               // rows (= threads) of one CTA of the owner-computes assembly
  size_t max_block_nnz = 0;             // max nnz over blocks of rows_per_cta consecutive rows
  // slot map: for entry a of the owned slice of the transposed test dof map (cell k, local dof i of row r) and local
  // trial dof j, the position of column dof_tr(k, j) inside row r -- so that the assembly never searches colind.
  // One byte per (a, j), padded to whole 32-bit words per entry.
  size_t a_begin[tfelcu::kMaxSeg] = {0}, a_count[tfelcu::kMaxSeg] = {0};
  bool has_slots = false, slots_failed = false;
  std::vector<tfelcu::DevBuf<uint32_t>> slots;   // [segment * n_trial + trial component][a_count * slot_words(nd_tr)]
  // P1 x P1 triangle blocks: 16-byte records (three vertex ids of the cell | three slot bytes + local dof) replacing the
  // separate reads of the transposed dof map, the slot map and the cell (same indexing as slots; empty when not built)
  bool has_recs = false;
  std::vector<tfelcu::DevBuf<uint32_t>> recs;
  // patch assembly (TFELCU_SCATTER_PATCH): per row segment the plan (owned by the test space) and, per row in patch
  // order, the offset of the row inside the shared-memory stage of its patch
  bool has_patches = false;
  tfelcu::PatchPlan* patch_plan[tfelcu::kMaxSeg] = {nullptr};
  std::vector<tfelcu::DevBuf<uint32_t>> patch_stage_off;
  size_t patch_max_stage[tfelcu::kMaxSeg] = {0};
  // tile assembly (TFELCU_SCATTER_TILE, tile.cu): one blob per tile of rows; built once per pattern
  tfelcu::TilePlan* tile_plan = nullptr;
  bool tile_refused = false;            // a tile exceeded a capacity of the plan once: stay on another strategy
  // a += (constant-coefficient form) that has been accepted but not launched yet: a linear form on the same space that
  // follows rides along in the same pass over the mesh (tile.cu); every other use of the matrix launches it first
  tfelcu::FormBuild* pending = nullptr;
  tfelcu::FormBuild* memo_form = nullptr;   // the last constant-coefficient form lowered for this matrix and its key
  std::vector<unsigned char> memo_key;      // (time / Newton loops repeat the same += : nothing is tabulated again)
  int pending_mode = 0;                 // 0: P1 rows (p1rows.cu), 1: tiles (tile.cu)
  // P1 x P1 triangle forms with constant coefficients (p1rows.cu): the packed records in a blocked ELL layout
  bool has_ell = false, ell_compact = false;   // compact: 8-byte records (vertex offsets from the row's own vertex)
  tfelcu::DevBuf<uint32_t> ell;         // uint4 records [ell_ptr[block] + k][128 rows of the block]
  tfelcu::DevBuf<unsigned int> ell_ptr, ell_deg;
  bool ell_local = false;               // compact records hold indices into the block's own vertex table (p1rows.cu)
  tfelcu::DevBuf<unsigned int> vt_ptr;  // [n_blocks + 1] start of a block's table in vt_list
  tfelcu::DevBuf<uint32_t> vt_list;     // ascending vertex ids every record of the block refers to (padded to 4 entries)
  tfelcu::DevBuf<int32_t> ell_desc;      // int4 per block for the streamed kernel
  unsigned int ell_deg_max = 0, vt_max = 0;
  unsigned int ell_uni_deg = 0, vt_uni = 0;   // > 0: every block holds this many record lines / table entries (padded): a
                                              // CTA addresses both from its block index alone   // largest degree / table over the blocks (shared-memory sizes of the streamed kernel)
  ~tfelcu_matrix() {
    if (tile_plan) tfelcu::tile_plan_free(tile_plan);
    if (pending) tfelcu::form_build_free(pending);
    if (memo_form) tfelcu::form_build_free(memo_form);
  }
  // cell colouring (TFELCU_SCATTER_COLORED): cells sorted by colour
  bool has_colors = false;
  int n_colors = 0;
  tfelcu::DevBuf<uint32_t> color_cells;     // [K] cell ids grouped by colour
  std::vector<size_t> color_start;          // [n_colors + 1]
};

namespace tfelcu {

constexpr int kRowsPerCta = 128;   // rows handled by one CTA of the owner-computes assembly

// ---- mesh.cu
void mesh_generate_square(tfelcu_mesh* m, double x1, double x2, uint32_t n1, uint32_t n2);
void mesh_generate_cube(tfelcu_mesh* m, double x1, double x2, double x3, uint32_t n1, uint32_t n2, uint32_t n3);
void mesh_geometry(const tfelcu_mesh* m, double* vol, double* jmt);
void mesh_validate_cells(tfelcu_mesh* m);
void mesh_build_vertex_rank(tfelcu_mesh* m);
EntityIndex* mesh_entities(tfelcu_mesh* m, int sd);
void mesh_build_boundary(tfelcu_mesh* m);
void mesh_neighbours(tfelcu_mesh* m, int32_t* d_nb);

// ---- fes.cu
void fes_build(tfelcu_fes* f);
void fes_boundary_dofs(tfelcu_fes* f, const uint32_t* d_bcells, size_t F, int nvb);
void fes_dof_coordinates(const tfelcu_fes* f, const uint32_t* d_dofs, size_t n, double* d_x);
void fes_add_dirichlet(tfelcu_fes* f, const uint32_t* d_dofs, const double* d_values, double constant, size_t n);
void fes_dirichlet_get(const tfelcu_fes* f, uint32_t* h_dofs, double* h_values);
void fes_vertex_values(const tfelcu_fes* f, const double* d_coef, double* d_out);
void fes_cell_values(const tfelcu_fes* f, const double* d_coef, double* d_out);

// ---- pattern.cu
void matrix_snapshot_dirichlet(tfelcu_matrix* m);
void matrix_build_pattern(tfelcu_matrix* m);
void matrix_build_colors(tfelcu_matrix* m);
void matrix_ensure_cleared(tfelcu_matrix* m);
// ---- tile.cu
bool tile_applicable(const tfelcu_matrix* m);
bool matrix_tile_plan_ready(tfelcu_matrix* m);
// ---- p1rows.cu
bool p1rows_applicable(const tfelcu_matrix* m);
bool matrix_ell_ready(tfelcu_matrix* m);
bool assemble_p1_rows(tfelcu_matrix* m, const FormBuild& fb, tfelcu_vector* v, const FormBuild* fbv);
bool assemble_tiles(tfelcu_matrix* m, const FormBuild& fb, tfelcu_vector* v, const FormBuild* fbv);
void matrix_flush_pending(tfelcu_matrix* m);     // assemble.cu: launch an accepted-but-deferred +=
void matrix_build_slots(tfelcu_matrix* m);
void matrix_build_recs(tfelcu_matrix* m);
void fes_transpose(tfelcu_fes* f);
void build_transpose(Ctx* ctx, const uint32_t* dof_map, size_t K, int nd, size_t n_dof,
                     DevBuf<uint32_t>& ptr, DevBuf<uint32_t>& adj);

// ---- patch.cu
PatchPlan* fes_patch_plan(tfelcu_fes* f, size_t row_begin, size_t row_end, int rows_per_patch);
void matrix_build_patches(tfelcu_matrix* m);
int resolve_scatter(int requested, bool single_block, int nd_te, int nd_tr);
bool assemble_bilinear_patch(tfelcu_matrix* m, int quad, const tfelcu_factor* factors, int n_factors,
                             const tfelcu_term* terms, int n_terms);
bool assemble_linear_patch(tfelcu_vector* v, int quad, const tfelcu_factor* factors, int n_factors,
                           const tfelcu_term* terms, int n_terms);

// ---- assemble.cu
void assemble_bilinear(tfelcu_matrix* m, int quad, const tfelcu_factor* factors, int n_factors,
                       const tfelcu_term* terms, int n_terms);
void assemble_linear(tfelcu_vector* v, int quad, const tfelcu_factor* factors, int n_factors,
                     const tfelcu_term* terms, int n_terms);
double integrate_scalar(Ctx* ctx, tfelcu_mesh* mesh, int quad, const tfelcu_factor* factors, int n_factors,
                        const tfelcu_term* terms, int n_terms);
int quadrature_size(int quad, int* dim);
void quadrature_points(Ctx* ctx, tfelcu_mesh* mesh, int quad, double* d_xq);

}  // namespace tfelcu
