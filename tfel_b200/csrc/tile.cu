// Tile assembly: bilinear_form::operator+= (src/core/bilinear_form.hpp:23-110) and linear_form::operator+=
// (linear_form.hpp:20-142) for vertex-based (P1) spaces on triangles, cell-centric inside shared-memory tiles, so that
// every cell, every vertex coordinate and every stored value crosses HBM once and every element matrix is computed once.
//
// The owner-computes rows of assemble.cu recompute the geometry of every cell once per incident row (dim + 1 times) and
// stream a 16-byte record per (row, cell): 2.5 GB of traffic for 1.6 GB of algorithmic bytes on configs[1]. Here the rows
// are grouped into compact tiles of the mesh (the quadtree / octree leaves of patch.cu: 16 x 16 vertices on a uniform
// grid) and ONE contiguous blob per tile, built once per pattern, holds everything a CTA needs:
//     header | coordinates of the tile's vertices (owned + rim) | per-vertex row info | per-row CSR positions |
//     one 12-byte record per cell: three tile-local vertex indices (9 bits each), nine 4-bit slots (position of column
//     dof(k, j) in row dof(k, i)) and, for each of its three rows, the RANK of the cell among the cells of that row
// Phase 1: one thread per cell computes the element matrix (closed form for P1: the gradients are constant on a cell) and
// stores row i of it at [row of vertex i][rank] in shared memory -- a private location, no atomics, no colours.
// Phase 2: one thread per row adds its contributions in rank order = ascending cell order = the order of the reference's
// cell loop (the same sums as the owner-computes rows, bit for bit when the quadrature weights sum to exactly 1) into the
// shared-memory image of its CSR row. Rows are then written run by run, coalesced, each stored value exactly once.
// Cells on the rim of a tile are recomputed by the neighbouring tile (~13 %). A constant-coefficient linear form on the
// same space rides along as a fourth value per contribution (one pass over the mesh for A and b).
// (A first version accumulated colour by colour -- Jones-Plassmann colours inside the tile, one CTA barrier per colour --
// and was slower than the rows: 1700 instructions per warp and tile, a third of the stall samples on those barriers;
// profiles/r02_ncu_p1rows_tile.md.)
#include "p1form.cuh"
#include "sortutil.cuh"

#include <cstdlib>

namespace tfelcu {

namespace {

constexpr int kTileThreads = 256;
constexpr int kTileMaxRefs = 4096;     // (dim + 1) x cells of a tile, padded to a power of two
constexpr int kTileMaxVerts = 511;     // 9-bit local vertex indices
constexpr int kTileMaxDeg = 15;        // cells per row: 4-bit ranks
constexpr int kTileRows = 256;

struct alignas(16) TileHeader {        // 32 bytes
  uint16_t n_rows, n_verts, n_cells, max_deg;
  uint32_t stage_len, pad[5];
};
static_assert(sizeof(TileHeader) == 32, "tile header layout");

__host__ __device__ inline size_t pad16(size_t b) { return (b + 15) & ~(size_t)15; }

// byte offsets of the sections of a blob
struct TileLayout {
  size_t coords, vinfo, rows, rowmeta, rec0, rec1, rec2, total;
  __host__ __device__ TileLayout(int R, int V, int C) {
    coords = sizeof(TileHeader);
    vinfo = coords + (size_t)V * 16;
    rows = vinfo + pad16((size_t)V * 4);
    rowmeta = rows + pad16((size_t)R * 8);
    rec0 = rowmeta + pad16((size_t)R * 4);
    rec1 = rec0 + pad16((size_t)C * 4);
    rec2 = rec1 + pad16((size_t)C * 4);
    total = rec2 + pad16((size_t)C * 4);
  }
};

constexpr uint32_t kNoRow = 0xffffffffu;

// ---- small block-wide helpers (kTileThreads threads) ----------------------------------------------------------------

// exclusive prefix sum of a[0, n) in place; returns the total. scratch: kTileThreads + 1 ints
__device__ int block_exclusive_scan(int* a, int n, int* scratch) {
  const int per = (n + kTileThreads - 1) / kTileThreads;
  const int b = threadIdx.x * per, e = min(b + per, n);
  int s = 0;
  for (int i = b; i < e; ++i) s += a[i];
  scratch[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int t = 0; t < kTileThreads; ++t) { const int v = scratch[t]; scratch[t] = run; run += v; }
    scratch[kTileThreads] = run;
  }
  __syncthreads();
  int run = scratch[threadIdx.x];
  for (int i = b; i < e; ++i) { const int v = a[i]; a[i] = run; run += v; }
  const int total = scratch[kTileThreads];
  __syncthreads();
  return total;
}

__device__ void block_bitonic_sort(uint32_t* a, int P) {
  for (int k = 2; k <= P; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < P; t += kTileThreads) {
        const int u = t ^ j;
        if (u > t) {
          const uint32_t x = a[t], y = a[u];
          const bool up = (t & k) == 0;
          if ((x > y) == up) { a[t] = y; a[u] = x; }
        }
      }
      __syncthreads();
    }
}

__device__ __forceinline__ int lower_bound_u32(const uint32_t* a, int n, uint32_t v) {
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
  return lo;
}

struct TileBuildArgs {
  const uint32_t* patch_row_ptr; const uint32_t* patch_rows; const uint32_t* patch_cell_ptr; const uint32_t* patch_cells;
  const uint32_t* cells; const double* vertices;
  const int32_t* rowptr; const int32_t* colind; const uint8_t* dir_row;
  uint32_t off_te; size_t to_local;          // global row = off_te + dof; local row = dof + to_local
  unsigned long long* sizes;                 // PASS 0: bytes of every blob
  const unsigned long long* blob_ptr; unsigned char* blob;   // PASS 1
  int* status;                               // 1: a tile exceeds a capacity (vertices, cells per row, row length, stage)
  unsigned int* maxima;                      // [0] rows, [1] vertices, [2] stage entries, [3] cells per row
};

// One CTA per tile (triangles). PASS 0 sizes the blob, PASS 1 writes it.
template <int PASS>
__global__ void __launch_bounds__(kTileThreads) k_tile_build(const TileBuildArgs A) {
  constexpr int NV = 3;
  extern __shared__ unsigned char tb_sm[];
  uint32_t* sorted = reinterpret_cast<uint32_t*>(tb_sm);                 // [kTileMaxRefs]
  uint32_t* uvert = sorted + kTileMaxRefs;                               // [512]
  uint32_t* vinfo = uvert + 512;                                         // [512] owned ordinal or kNoRow
  int* vlen = reinterpret_cast<int*>(vinfo + 512);                       // [512] row length of owned vertices -> stage offsets
  int* vdeg = vlen + 512;                                                // [512] cells of the tile at the vertex
  int* scratch = vdeg + 512;                                             // [kTileThreads + 8]
  int* head = scratch + kTileThreads + 8;                                // [kTileMaxRefs] unique flags / offsets
  uint16_t* clv = reinterpret_cast<uint16_t*>(head + kTileMaxRefs);      // [kTileMaxRefs] local vertex of every cell ref
  uint8_t* rank = reinterpret_cast<uint8_t*>(clv + kTileMaxRefs);        // [kTileMaxRefs] rank of the cell at that vertex
  uint8_t* vdir = rank + kTileMaxRefs;                                   // [512] owned Dirichlet rows (identity rows)
  __shared__ int s_bad, s_maxdeg;

  const size_t tile = blockIdx.x;
  const uint32_t r_begin = A.patch_row_ptr[tile], r_end = A.patch_row_ptr[tile + 1];
  const uint32_t c_begin = A.patch_cell_ptr[tile], c_end = A.patch_cell_ptr[tile + 1];
  const int R = (int)(r_end - r_begin), C = (int)(c_end - c_begin);
  if (threadIdx.x == 0) { s_bad = 0; s_maxdeg = 0; }
  __syncthreads();
  if (C * NV > kTileMaxRefs || R > 511 || C > 1365) {
    if (threadIdx.x == 0) { atomicExch(A.status, 1); if (PASS == 0) A.sizes[tile] = 0; }
    return;
  }
  int P = 32;
  while (P < C * NV) P <<= 1;
  // 1. vertex references of the tile's cells
  for (int t = threadIdx.x; t < P; t += kTileThreads) {
    uint32_t v = 0xffffffffu;
    if (t < C * NV) v = A.cells[(size_t)A.patch_cells[c_begin + t / NV] * NV + t % NV];
    sorted[t] = v;
  }
  __syncthreads();
  block_bitonic_sort(sorted, P);
  // 2. unique vertices, ascending: local vertex order
  for (int t = threadIdx.x; t < P; t += kTileThreads)
    head[t] = (sorted[t] != 0xffffffffu && (t == 0 || sorted[t] != sorted[t - 1])) ? 1 : 0;
  __syncthreads();
  const int V = block_exclusive_scan(head, P, scratch);
  if (V > kTileMaxVerts) {
    if (threadIdx.x == 0) { atomicExch(A.status, 1); if (PASS == 0) A.sizes[tile] = 0; }
    return;
  }
  for (int t = threadIdx.x; t < P; t += kTileThreads)
    if (sorted[t] != 0xffffffffu && (t == 0 || sorted[t] != sorted[t - 1])) uvert[head[t]] = sorted[t];
  __syncthreads();
  // 3. owned rows (ascending in the patch plan): CSR positions, stage offsets
  for (int lv = threadIdx.x; lv < V; lv += kTileThreads) {
    const uint32_t v = uvert[lv];
    const int t = lower_bound_u32(A.patch_rows + r_begin, R, v);
    const bool owned = t < R && A.patch_rows[r_begin + t] == v;
    int len = 0;
    if (owned) { const size_t lrow = (size_t)v + A.to_local; len = A.rowptr[lrow + 1] - A.rowptr[lrow]; }
    vlen[lv] = len;
    vinfo[lv] = owned ? (uint32_t)t : kNoRow;
    vdir[lv] = (owned && A.dir_row[(size_t)A.off_te + v] != 0) ? 1 : 0;
    if (owned && len > 16) atomicExch(&s_bad, 1);  // 4-bit slots
  }
  __syncthreads();
  const int S = block_exclusive_scan(vlen, V, scratch);   // vlen[lv] = stage offset of the row of lv
  if (S > 0xfff0 || s_bad) {
    if (threadIdx.x == 0) { atomicExch(A.status, 1); if (PASS == 0) A.sizes[tile] = 0; }
    return;
  }
  // 4. local vertex of every cell reference (cells in the plan's order: ascending cell id) and the rank of the cell among
  //    the cells of the tile at that vertex: sort (vertex, cell, i) keys, rank = position inside the vertex's run
  for (int t = threadIdx.x; t < P; t += kTileThreads) {
    uint32_t key = 0xffffffffu;
    if (t < C * NV) {
      const uint32_t v = A.cells[(size_t)A.patch_cells[c_begin + t / NV] * NV + t % NV];
      const int lv = lower_bound_u32(uvert, V, v);
      clv[t] = (uint16_t)lv;
      key = ((uint32_t)lv << 13) | (uint32_t)t;      // t = cell * 3 + i < 4096 = 2^12
    }
    sorted[t] = key;
  }
  __syncthreads();
  block_bitonic_sort(sorted, P);
  for (int t = threadIdx.x; t < C * NV; t += kTileThreads) {
    const uint32_t key = sorted[t];
    const uint32_t lv = key >> 13;
    const int first = lower_bound_u32(sorted, C * NV, lv << 13);
    const int rk = t - first;
    if (rk > kTileMaxDeg) atomicExch(&s_bad, 1);
    rank[key & 0x1fffu] = (uint8_t)rk;
    const bool last = (t + 1 == C * NV) || (sorted[t + 1] >> 13) != lv;
    if (last) { vdeg[lv] = rk + 1; atomicMax(&s_maxdeg, rk + 1); }
  }
  __syncthreads();
  if (s_bad) {
    if (threadIdx.x == 0) { atomicExch(A.status, 1); if (PASS == 0) A.sizes[tile] = 0; }
    return;
  }
  if (PASS == 0) {
    if (threadIdx.x == 0) {
      A.sizes[tile] = TileLayout(R, V, C).total;
      atomicMax(A.maxima + 0, (unsigned)R); atomicMax(A.maxima + 1, (unsigned)V);
      atomicMax(A.maxima + 2, (unsigned)S); atomicMax(A.maxima + 3, (unsigned)s_maxdeg);
    }
    return;
  }
  // 5. write the blob
  unsigned char* out = A.blob + A.blob_ptr[tile];
  const TileLayout L(R, V, C);
  if (threadIdx.x == 0) {
    TileHeader H;
    H.n_rows = (uint16_t)R; H.n_verts = (uint16_t)V; H.n_cells = (uint16_t)C; H.max_deg = (uint16_t)s_maxdeg;
    H.stage_len = (uint32_t)S;
    for (int k = 0; k < 5; ++k) H.pad[k] = 0;
    *reinterpret_cast<TileHeader*>(out) = H;
  }
  double2* o_coords = reinterpret_cast<double2*>(out + L.coords);
  uint32_t* o_vinfo = reinterpret_cast<uint32_t*>(out + L.vinfo);
  int2* o_rows = reinterpret_cast<int2*>(out + L.rows);
  uint32_t* o_meta = reinterpret_cast<uint32_t*>(out + L.rowmeta);
  uint32_t* o_rec0 = reinterpret_cast<uint32_t*>(out + L.rec0);
  uint32_t* o_rec1 = reinterpret_cast<uint32_t*>(out + L.rec1);
  uint32_t* o_rec2 = reinterpret_cast<uint32_t*>(out + L.rec2);
  for (int lv = threadIdx.x; lv < V; lv += kTileThreads) {
    const uint32_t v = uvert[lv];
    o_coords[lv] = reinterpret_cast<const double2*>(A.vertices)[v];
    const uint32_t t = vinfo[lv];
    uint32_t info = 0u;
    if (t != kNoRow) {
      const size_t lrow = (size_t)v + A.to_local;
      const int rs = A.rowptr[lrow], len = A.rowptr[lrow + 1] - rs;
      const bool dir = vdir[lv] != 0;
      int diag = 0;
      if (dir) {   // clear(): identity row (bilinear_form.hpp:159-165)
        const int gcol = (int)(A.off_te + v);
        int lo = rs, hi = rs + len;
        while (lo < hi) { const int mid = lo + ((hi - lo) >> 1); if (A.colind[mid] < gcol) lo = mid + 1; else hi = mid; }
        diag = lo - rs;
      }
      o_rows[t] = make_int2(rs, (int)lrow);
      // stage offset (16) | row length (5) | cells at the row (4) | diagonal slot (4) | identity row
      o_meta[t] = (uint32_t)vlen[lv] | ((uint32_t)len << 16) | ((uint32_t)vdeg[lv] << 21) | ((uint32_t)diag << 25) | (dir ? 0x80000000u : 0u);
      info = t | 0x80000000u;     // owned ordinal | owned
    }
    o_vinfo[lv] = info;
  }
  for (int c = threadIdx.x; c < C; c += kTileThreads) {
    uint32_t ids[NV];
    unsigned long long slots = 0ull;
    uint32_t w0 = 0u, ranks = 0u;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      ids[v] = uvert[clv[c * NV + v]];
      w0 |= (uint32_t)clv[c * NV + v] << (9 * v);
      ranks |= (uint32_t)rank[c * NV + v] << (4 * v);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int lv = clv[c * NV + i];
      if (vinfo[lv] == kNoRow || vdir[lv]) continue;   // rows of other tiles; identity rows take no contributions
      const size_t lrow = (size_t)ids[i] + A.to_local;
      const int rs = A.rowptr[lrow], re = A.rowptr[lrow + 1];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int gcol = (int)(A.off_te + ids[j]);
        int lo = rs, hi = re;
        while (lo < hi) { const int mid = lo + ((hi - lo) >> 1); if (A.colind[mid] < gcol) lo = mid + 1; else hi = mid; }
        if (lo >= re || A.colind[lo] != gcol) atomicExch(A.status, 2);
        slots |= (unsigned long long)((lo - rs) & 0xf) << (4 * (i * NV + j));
      }
    }
    // w0: three local vertices (27 bits) | rank at vertex 0 (4 bits); w1: slots of rows 0 and 1 (24 bits) | ranks at
    // vertices 1 and 2 (8 bits); w2: slots of row 2 (12 bits)
    o_rec0[c] = w0 | ((ranks & 0xfu) << 27);
    o_rec1[c] = (uint32_t)(slots & 0xffffffull) | ((ranks >> 4) << 24);
    o_rec2[c] = (uint32_t)(slots >> 24);
  }
}

size_t tile_build_smem() {
  return (size_t)kTileMaxRefs * 4 + 512 * 4 * 4 + (kTileThreads + 8) * 4 + (size_t)kTileMaxRefs * 4 + (size_t)kTileMaxRefs * 2 +
         (size_t)kTileMaxRefs + 512 + 64;
}

// ------------------------------------------------------------------------------------------
// assembly kernel
// ------------------------------------------------------------------------------------------

struct TileAsm {
  const unsigned long long* blob_ptr; const unsigned char* blob;
  double* val; int init_mode;                  // 1: the values are uninitialised (pending clear()): write them fresh
  double* b;                                   // fused linear form (indexed by local row), or nullptr
  int b_init;                                  // 1: its clear() is pending: write the entries
  int max_deg;                                 // capacity of the contribution lists (cells per row), over all tiles
};

template <bool FUSE_B>
__global__ void __launch_bounds__(kTileThreads, 3)
k_assemble_tiles(const __grid_constant__ P1Form F, const TileAsm T) {
  extern __shared__ __align__(16) double tsm[];
  __shared__ TileHeader H;
  const unsigned char* blob = T.blob + T.blob_ptr[blockIdx.x];
  if (threadIdx.x < sizeof(TileHeader) / 4) reinterpret_cast<uint32_t*>(&H)[threadIdx.x] = reinterpret_cast<const uint32_t*>(blob)[threadIdx.x];
  __syncthreads();
  const int R = H.n_rows, V = H.n_verts, C = H.n_cells, S = (int)H.stage_len;
  const int D = T.max_deg;
  const TileLayout L(R, V, C);
  // shared memory: contributions [R][D] x 4 doubles (row i of an element matrix | its linear-form entry), coordinates,
  // stage (image of the CSR rows), per-vertex row info, slots of every contribution
  double4* s_con = reinterpret_cast<double4*>(tsm);
  double2* s_xy = reinterpret_cast<double2*>(s_con + (size_t)R * D);
  double* s_stage = reinterpret_cast<double*>(s_xy + V);
  uint32_t* s_vinfo = reinterpret_cast<uint32_t*>(s_stage + S);
  uint16_t* s_slot = reinterpret_cast<uint16_t*>(s_vinfo + V);
  const double2* g_xy = reinterpret_cast<const double2*>(blob + L.coords);
  const uint32_t* g_vinfo = reinterpret_cast<const uint32_t*>(blob + L.vinfo);
  const int2* g_rows = reinterpret_cast<const int2*>(blob + L.rows);
  const uint32_t* g_meta = reinterpret_cast<const uint32_t*>(blob + L.rowmeta);
  const uint32_t* g_rec0 = reinterpret_cast<const uint32_t*>(blob + L.rec0);
  const uint32_t* g_rec1 = reinterpret_cast<const uint32_t*>(blob + L.rec1);
  const uint32_t* g_rec2 = reinterpret_cast<const uint32_t*>(blob + L.rec2);
  for (int i = threadIdx.x; i < V; i += kTileThreads) { s_xy[i] = __ldg(g_xy + i); s_vinfo[i] = __ldg(g_vinfo + i); }
  // the records of this thread's cells are requested before the barrier
  uint32_t w0[2], w1[2], w2[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int c = threadIdx.x + u * kTileThreads;
    w0[u] = c < C ? __ldg(g_rec0 + c) : 0u; w1[u] = c < C ? __ldg(g_rec1 + c) : 0u; w2[u] = c < C ? __ldg(g_rec2 + c) : 0u;
  }
  __syncthreads();
  // ---- phase 1: element matrices, row i -> contribution list of the row of vertex i -----------------------------------
  auto cell = [&](uint32_t r0, uint32_t r1, uint32_t r2) {
    const int l0 = (int)(r0 & 0x1ff), l1 = (int)((r0 >> 9) & 0x1ff), l2 = (int)((r0 >> 18) & 0x1ff);
    const double2 p0 = s_xy[l0], p1 = s_xy[l1], p2 = s_xy[l2];
    const double J00 = p1.x - p0.x, J10 = p1.y - p0.y, J01 = p2.x - p0.x, J11 = p2.y - p0.y;
    const double det = J00 * J11 - J10 * J01;
    const double vol = 0.5 * fabs(det);
    const double rdet = 1.0 / det;
    double jmt[2][2];
    jmt[0][0] = J11 * rdet; jmt[0][1] = -J10 * rdet; jmt[1][0] = -J01 * rdet; jmt[1][1] = J00 * rdet;
    double K[9];
    p1_element(F, jmt, K);
    const uint32_t ranks = (r0 >> 27) | ((r1 >> 24) << 4);
    const unsigned long long slots = (unsigned long long)(r1 & 0xffffffu) | ((unsigned long long)r2 << 24);
    const uint32_t i0 = s_vinfo[l0], i1 = s_vinfo[l1], i2 = s_vinfo[l2];
    if (i0 & 0x80000000u) {                                  // else: the row belongs to another tile
      const int at = (int)(i0 & 0x1ffu) * D + (int)(ranks & 0xfu);
      s_con[at] = make_double4(K[0] * vol, K[1] * vol, K[2] * vol, FUSE_B ? F.b_c[0] * vol : 0.0);
      s_slot[at] = (uint16_t)(slots & 0xfffu);
    }
    if (i1 & 0x80000000u) {
      const int at = (int)(i1 & 0x1ffu) * D + (int)((ranks >> 4) & 0xfu);
      s_con[at] = make_double4(K[3] * vol, K[4] * vol, K[5] * vol, FUSE_B ? F.b_c[1] * vol : 0.0);
      s_slot[at] = (uint16_t)((slots >> 12) & 0xfffu);
    }
    if (i2 & 0x80000000u) {
      const int at = (int)(i2 & 0x1ffu) * D + (int)((ranks >> 8) & 0xfu);
      s_con[at] = make_double4(K[6] * vol, K[7] * vol, K[8] * vol, FUSE_B ? F.b_c[2] * vol : 0.0);
      s_slot[at] = (uint16_t)((slots >> 24) & 0xfffu);
    }
  };
  if ((int)threadIdx.x < C) cell(w0[0], w1[0], w2[0]);
  if ((int)threadIdx.x + kTileThreads < C) cell(w0[1], w1[1], w2[1]);
  for (int c = threadIdx.x + 2 * kTileThreads; c < C; c += kTileThreads) cell(__ldg(g_rec0 + c), __ldg(g_rec1 + c), __ldg(g_rec2 + c));
  __syncthreads();
  // ---- phase 2: every row adds its contributions in rank order (ascending cell id) ----------------------------------
  for (int t = threadIdx.x; t < R; t += kTileThreads) {
    const uint32_t m = __ldg(g_meta + t);
    const int off = (int)(m & 0xffffu), len = (int)((m >> 16) & 0x1fu), deg = (int)((m >> 21) & 0xfu);
    const bool dir = (m & 0x80000000u) != 0u;
    double* row = s_stage + off;
    for (int e = 0; e < len; ++e) row[e] = 0.0;
    if (dir && T.init_mode) row[(m >> 25) & 0xfu] = 1.0;   // identity rows of the Dirichlet dofs (clear())
    double bsum = 0.0;
    for (int k = 0; k < deg; ++k) {
      const double4 v = s_con[t * D + k];
      if (!dir) {
        const uint32_t sl = s_slot[t * D + k];
        row[sl & 0xfu] += v.x; row[(sl >> 4) & 0xfu] += v.y; row[(sl >> 8) & 0xfu] += v.z;
      }
      if (FUSE_B) bsum += v.w;
    }
    if (FUSE_B) { const int2 rr = __ldg(g_rows + t); T.b[rr.y] = T.b_init ? bsum : T.b[rr.y] + bsum; }
  }
  __syncthreads();
  // ---- write-out: a warp takes 32 consecutive owned rows, splits them into runs of rows that are consecutive in the CSR
  // arrays and copies every run with coalesced stores
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t0 = warp * 32; t0 < R; t0 += kTileThreads) {
    const int t = t0 + lane;
    int rs = 0, len = 0, off = 0;
    if (t < R) {
      const int2 rr = __ldg(g_rows + t); const uint32_t m = __ldg(g_meta + t);
      rs = rr.x; off = (int)(m & 0xffffu); len = (int)((m >> 16) & 0x1fu);
    }
    const int prev_end = __shfl_up_sync(0xffffffffu, rs + len, 1);
    const bool start = (t < R) && (lane == 0 || prev_end != rs);
    unsigned starts = __ballot_sync(0xffffffffu, start);
    const unsigned valid = __ballot_sync(0xffffffffu, t < R);
    while (starts) {
      const int a = __ffs(starts) - 1;
      starts &= starts - 1;
      const int z = starts ? __ffs(starts) - 1 : (32 - __clz(valid));   // first lane after the run
      const int run_rs = __shfl_sync(0xffffffffu, rs, a), run_off = __shfl_sync(0xffffffffu, off, a);
      const int end_rs = __shfl_sync(0xffffffffu, rs + len, z - 1);
      const int n = end_rs - run_rs;
      if (T.init_mode) for (int e = lane; e < n; e += 32) T.val[(size_t)run_rs + e] = s_stage[run_off + e];
      else for (int e = lane; e < n; e += 32) T.val[(size_t)run_rs + e] += s_stage[run_off + e];
    }
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------

struct TilePlan {
  size_t n_tile = 0;
  DevBuf<unsigned long long> blob_ptr;
  DevBuf<unsigned char> blob;
  unsigned max_rows = 0, max_verts = 0, max_stage = 0, max_deg = 0;
  uint64_t pattern_id = 0;
};

void tile_plan_free(TilePlan* p) { delete p; }

// the form can take the tile path: one scalar P1 space on triangles for test and trial, every vertex used by a cell
// (dof == vertex id), one owned row segment
bool tile_applicable(const tfelcu_matrix* m) {
  if (m->test.n != 1 || m->trial.n != 1 || m->test.fes[0] != m->trial.fes[0]) return false;
  const tfelcu_fes* f = m->test.fes[0];
  if (f->mesh->dim != 2 || f->fe != TFELCU_FE_P1 || f->nd != 3) return false;
  if (!f->mesh->all_vertices_used || f->dof_map != f->mesh->cells.p) return false;
  return m->rows.n == 1;
}

TilePlan* matrix_tile_plan(tfelcu_matrix* m) {
  if (m->tile_plan && m->tile_plan->pattern_id == m->pattern_id) return m->tile_plan;
  if (m->tile_refused) return nullptr;
  if (m->tile_plan) { delete m->tile_plan; m->tile_plan = nullptr; }
  Ctx* ctx = m->ctx;
  tfelcu_fes* te = m->test.fes[0];
  const RowMap& R = m->rows;
  const size_t rb = R.g_begin[0] - m->test.offset[0], re = R.g_end[0] - m->test.offset[0];
  PatchPlan* P = fes_patch_plan(te, rb, re, kTileRows);
  std::unique_ptr<TilePlan> T(new TilePlan());
  T->n_tile = P->n_patch;
  T->pattern_id = m->pattern_id;
  if (!T->n_tile) { m->tile_plan = T.release(); return m->tile_plan; }
  DevBuf<unsigned long long> sizes(T->n_tile + 1);
  DevBuf<int> status(1);
  DevBuf<unsigned int> maxima(4);
  ctx->zero(sizes.p, T->n_tile + 1); ctx->zero(status.p, 1); ctx->zero(maxima.p, 4);
  TileBuildArgs A;
  A.patch_row_ptr = P->row_ptr.p; A.patch_rows = P->rows.p; A.patch_cell_ptr = P->cell_ptr.p; A.patch_cells = P->cells.p;
  A.cells = m->mesh->cells.p; A.vertices = m->mesh->vertices.p;
  A.rowptr = m->rowptr.p; A.colind = m->colind.p; A.dir_row = m->dir_row.p;
  A.off_te = (uint32_t)m->test.offset[0]; A.to_local = (size_t)R.local_off[0] - rb;
  A.sizes = sizes.p; A.blob_ptr = nullptr; A.blob = nullptr; A.status = status.p; A.maxima = maxima.p;
  const size_t smem = tile_build_smem();
  TFELCU_CK(cudaFuncSetAttribute(k_tile_build<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TFELCU_CK(cudaFuncSetAttribute(k_tile_build<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_tile_build<0><<<(unsigned)T->n_tile, kTileThreads, smem, ctx->stream>>>(A);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  int h_status = 0; ctx->download(&h_status, status.p, 1);
  if (h_status) { m->tile_refused = true; return nullptr; }   // a tile exceeds a capacity: the caller stays on the rows
  unsigned h_max[4]; ctx->download(h_max, maxima.p, 4);
  T->max_rows = h_max[0]; T->max_verts = h_max[1]; T->max_stage = h_max[2]; T->max_deg = h_max[3];
  T->blob_ptr.alloc(T->n_tile + 1);
  exclusive_sum<unsigned long long, unsigned long long>(ctx, sizes.p, T->blob_ptr.p, T->n_tile + 1);
  unsigned long long total = 0; ctx->download(&total, T->blob_ptr.p + T->n_tile, 1);
  T->blob.alloc((size_t)total);
  A.blob_ptr = T->blob_ptr.p; A.blob = T->blob.p;
  k_tile_build<1><<<(unsigned)T->n_tile, kTileThreads, smem, ctx->stream>>>(A);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  ctx->download(&h_status, status.p, 1);
  TFELCU_REQUIRE(h_status == 0, "tile plan: a column of a cell is missing from the CSR pattern");
  if (std::getenv("TFELCU_TRACE"))
    std::fprintf(stderr, "[tfelcu trace] tile plan: %zu tiles, %.1f MB, max rows %u verts %u stage %u cells per row %u\n", T->n_tile,
                 (double)total / 1e6, T->max_rows, T->max_verts, T->max_stage, T->max_deg);
  m->tile_plan = T.release();
  return m->tile_plan;
}

bool matrix_tile_plan_ready(tfelcu_matrix* m) { return matrix_tile_plan(m) != nullptr; }

// a += (constant-coefficient form) [and b += (constant-coefficient, order-0 linear form)] in one pass over the tiles.
// Returns false when the tile path does not apply (the caller uses another strategy); nothing has been touched then.
bool assemble_tiles(tfelcu_matrix* m, const FormBuild& fb, tfelcu_vector* v, const FormBuild* fbv) {
  if (!tile_applicable(m) || !fb.const_coef) return false;
  TilePlan* T = matrix_tile_plan(m);
  if (!T) return false;
  Ctx* ctx = m->ctx;
  TileAsm K;
  K.blob_ptr = T->blob_ptr.p; K.blob = T->blob.p; K.val = m->val.p; K.init_mode = m->val_valid ? 0 : 1;
  K.b = nullptr; K.max_deg = (int)T->max_deg;
  const P1Form F = make_p1_form(fb, v ? fbv : nullptr);
  K.b_init = 0;
  if (v) { K.b = v->data.p; K.b_init = v->zero_pending ? 1 : 0; }
  if (!T->n_tile) { m->val_valid = true; return true; }
  const size_t con = (size_t)T->max_rows * T->max_deg;
  const size_t smem = con * 32 + (size_t)T->max_verts * 16 + (size_t)T->max_stage * 8 + (size_t)T->max_verts * 4 + con * 2 + 32;
  TFELCU_REQUIRE(smem <= 200 * 1024, "tile assembly: a tile needs more shared memory than an SM has");
  if (v) {
    ensure_smem(k_assemble_tiles<true>, smem);
    k_assemble_tiles<true><<<(unsigned)T->n_tile, kTileThreads, smem, ctx->stream>>>(F, K);
  } else {
    ensure_smem(k_assemble_tiles<false>, smem);
    k_assemble_tiles<false><<<(unsigned)T->n_tile, kTileThreads, smem, ctx->stream>>>(F, K);
  }
  ctx->count(); TFELCU_LAUNCH_CHECK();
  if (v) v->zero_pending = false;
  m->val_valid = true;
  return true;
}

}  // namespace tfelcu
