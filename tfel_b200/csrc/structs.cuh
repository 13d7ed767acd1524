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
};

namespace tfelcu {

constexpr int kMaxComp = 4;

struct SpaceList {
  int n = 0;
  tfelcu_fes* fes[kMaxComp] = {nullptr, nullptr, nullptr, nullptr};
  size_t offset[kMaxComp + 1] = {0, 0, 0, 0, 0};
  size_t total() const { return offset[n]; }
};

}  // namespace tfelcu

struct tfelcu_vector {
  tfelcu::Ctx* ctx = nullptr;
  tfelcu::SpaceList space;
  tfelcu::DevBuf<double> data;
};

struct tfelcu_matrix {
  tfelcu::Ctx* ctx = nullptr;
  tfelcu_mesh* mesh = nullptr;
  tfelcu::SpaceList test, trial;
  size_t n_rows = 0, n_cols = 0, nnz = 0;
  tfelcu::DevBuf<uint8_t> dir_row;      // [n_rows]: snapshot of the test spaces' Dirichlet flags
  tfelcu::DevBuf<double> dir_value;     // [n_rows]
  uint64_t dir_version[tfelcu::kMaxComp] = {0, 0, 0, 0};
  tfelcu::DevBuf<int32_t> rowptr, colind;
  tfelcu::DevBuf<double> val;
  bool pristine = true;                 // no += since clear()
  int scatter = TFELCU_SCATTER_ROW_GATHER;
  bool val_valid = false;                // false: clear() is pending (performed lazily by the first user)
  size_t max_block_nnz = 0;             // max nnz over blocks of kRowsPerCta consecutive rows
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

// ---- fes.cu
void fes_build(tfelcu_fes* f);
void fes_boundary_dofs(tfelcu_fes* f, const uint32_t* d_bcells, size_t F, int nvb);
void fes_dof_coordinates(const tfelcu_fes* f, const uint32_t* d_dofs, size_t n, double* d_x);
void fes_add_dirichlet(tfelcu_fes* f, const uint32_t* d_dofs, const double* d_values, double constant, size_t n);
void fes_dirichlet_get(const tfelcu_fes* f, uint32_t* h_dofs, double* h_values);

// ---- pattern.cu
void matrix_snapshot_dirichlet(tfelcu_matrix* m);
void matrix_build_pattern(tfelcu_matrix* m);
void matrix_build_colors(tfelcu_matrix* m);
void matrix_ensure_cleared(tfelcu_matrix* m);
void fes_transpose(tfelcu_fes* f);
void build_transpose(Ctx* ctx, const uint32_t* dof_map, size_t K, int nd, size_t n_dof,
                     DevBuf<uint32_t>& ptr, DevBuf<uint32_t>& adj);

// ---- assemble.cu
void assemble_bilinear(tfelcu_matrix* m, int quad, const tfelcu_factor* factors, int n_factors,
                       const tfelcu_term* terms, int n_terms);
void assemble_linear(tfelcu_vector* v, int quad, const tfelcu_factor* factors, int n_factors,
                     const tfelcu_term* terms, int n_terms);
double integrate_scalar(Ctx* ctx, tfelcu_mesh* mesh, int quad, const tfelcu_factor* factors, int n_factors,
                        const tfelcu_term* terms, int n_terms);

}  // namespace tfelcu
