// Per-cell affine geometry on device: |K| and J^{-T} of the map x = v0 + J xhat.
// Reference: triangle::get_cell_volume / get_jmt (cell.hpp:474-502), tetrahedron:: (cell.hpp:799-837),
// where J^{-1} comes from LAPACK LU (linear_algebra.hpp:26-40); here it is the closed-form adjugate.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>

namespace tfelcu {

template <int DIM>
struct CellGeom {
  double vol;
  double jmt[DIM][DIM];   // jmt[s][t] = (J^{-1})[t][s]
  double v0[DIM];
  double J[DIM][DIM];     // J[r][c] = v_{c+1}[r] - v_0[r]
};

template <int DIM>
__device__ __forceinline__ void finish_cell_geometry(CellGeom<DIM>& g) {
  if constexpr (DIM == 2) {
    const double det = g.J[0][0] * g.J[1][1] - g.J[1][0] * g.J[0][1];
    g.vol = 0.5 * fabs(det);
    const double r = 1.0 / det;
    // J^{-1} = 1/det [[J11, -J01], [-J10, J00]]; jmt is its transpose
    g.jmt[0][0] = g.J[1][1] * r;
    g.jmt[0][1] = -g.J[1][0] * r;
    g.jmt[1][0] = -g.J[0][1] * r;
    g.jmt[1][1] = g.J[0][0] * r;
  } else {
    const double (*J)[DIM] = g.J;
    double adj[3][3];   // J^{-1} = adj / det
    adj[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    adj[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
    adj[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
    adj[1][0] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    adj[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
    adj[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
    adj[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    adj[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
    adj[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double det = J[0][0] * adj[0][0] + J[0][1] * adj[1][0] + J[0][2] * adj[2][0];
    g.vol = fabs(det) * (1.0 / 6.0);
    const double r = 1.0 / det;
#pragma unroll
    for (int s = 0; s < DIM; ++s)
#pragma unroll
      for (int t = 0; t < DIM; ++t) g.jmt[s][t] = adj[t][s] * r;
  }
}

// triangle whose three vertex ids are already in registers (packed row records of the owner-computes assembly)
__device__ __forceinline__ void triangle_geometry_from_ids(const double* __restrict__ vertices, uint32_t i0, uint32_t i1,
                                                           uint32_t i2, CellGeom<2>& g) {
  const double2* v2 = reinterpret_cast<const double2*>(vertices);
  const double2 p0 = __ldg(v2 + i0), p1 = __ldg(v2 + i1), p2 = __ldg(v2 + i2);
  g.v0[0] = p0.x; g.v0[1] = p0.y;
  g.J[0][0] = p1.x - p0.x; g.J[1][0] = p1.y - p0.y;
  g.J[0][1] = p2.x - p0.x; g.J[1][1] = p2.y - p0.y;
  finish_cell_geometry<2>(g);
}

template <int DIM>
__device__ __forceinline__ void load_cell_geometry(const double* __restrict__ vertices,
                                                   const uint32_t* __restrict__ cell,
                                                   CellGeom<DIM>& g) {
  uint32_t id[DIM + 1];
  if constexpr (DIM == 3) {   // four ids = one 16-byte load (cells are [K][4], 16-byte aligned)
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(cell));
    id[0] = q.x; id[1] = q.y; id[2] = q.z; id[3] = q.w;
  } else {
#pragma unroll
    for (int v = 0; v <= DIM; ++v) id[v] = __ldg(cell + v);
  }
  if constexpr (DIM == 2) {   // a vertex = one 16-byte load
    const double2* v2 = reinterpret_cast<const double2*>(vertices);
    const double2 p0 = __ldg(v2 + id[0]), p1 = __ldg(v2 + id[1]), p2 = __ldg(v2 + id[2]);
    g.v0[0] = p0.x; g.v0[1] = p0.y;
    g.J[0][0] = p1.x - p0.x; g.J[1][0] = p1.y - p0.y;
    g.J[0][1] = p2.x - p0.x; g.J[1][1] = p2.y - p0.y;
  } else {
#pragma unroll
    for (int r = 0; r < DIM; ++r) g.v0[r] = __ldg(vertices + (size_t)id[0] * DIM + r);
#pragma unroll
    for (int c = 0; c < DIM; ++c)
#pragma unroll
      for (int r = 0; r < DIM; ++r)
        g.J[r][c] = __ldg(vertices + (size_t)id[c + 1] * DIM + r) - g.v0[r];
  }
  finish_cell_geometry<DIM>(g);
}

}  // namespace tfelcu
