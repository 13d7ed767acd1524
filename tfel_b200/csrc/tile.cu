// Tile assembly: bilinear_form::operator+= (src/core/bilinear_form.hpp:23-110) and linear_form::operator+=
// (linear_form.hpp:20-142) for vertex-based (P1) spaces, cell-centric with graph colouring inside shared-memory tiles --
// the scatter north_star names ("graph coloring, not atomics, so the result is deterministic"), arranged so that every
// cell, every vertex coordinate and every stored value crosses HBM once.
//
// The owner-computes rows of assemble.cu recompute the geometry of every cell once per incident row (dim + 1 times) and
// stream a 16-byte record per (row, cell): 2.5 GB of traffic for 1.6 GB of algorithmic bytes on configs[1]. Here the rows
// are grouped into compact tiles of the mesh (the quadtree / octree leaves of patch.cu: 16 x 16 vertices on a uniform
// grid) and ONE contiguous blob per tile, built once per pattern, holds everything a CTA needs:
//     header | coordinates of the tile's vertices (owned + rim) | per-vertex stage offsets | per-row CSR positions |
//     one 8-byte record per cell: three tile-local vertex indices (9 bits each) + nine 4-bit slots
// The cells of a tile are sorted by COLOUR (no two cells of a colour share a vertex; Jones-Plassmann rounds with fixed
// hashed priorities in shared memory, deterministic). A CTA computes each element matrix once, in registers, and adds it
// colour by colour into the shared-memory image of its CSR rows: race-free without atomics, the order of the additions to
// an entry is the colour order -- bitwise reproducible. Rows are then written run by run, coalesced, each stored value
// exactly once. Cells on the rim of a tile are recomputed by the neighbouring tile (~13 %). A constant-coefficient linear
// form on the same space can ride along (one pass over the mesh for A and b).
#include "element.cuh"
#include "sortutil.cuh"

#include <cstdlib>

namespace tfelcu {

namespace {

constexpr int kTileThreads = 256;
constexpr int kTileMaxRefs = 4096;     // (dim + 1) x cells of a tile, padded to a power of two
constexpr int kTileMaxVerts = 511;     // 9-bit local vertex indices, 0x1ff never used
constexpr int kTileMaxColors = 24;
constexpr int kTileRows = 256;

struct alignas(16) TileHeader {        // 80 bytes
  uint16_t n_rows, n_verts, n_cells, n_colors;
  uint32_t stage_len, pad;
  uint16_t color_off[kTileMaxColors + 2];   // cells of colour c: [color_off[c], color_off[c + 1]) in record order
  uint16_t pad2[6];
};
static_assert(sizeof(TileHeader) == 80, "tile header layout");

__host__ __device__ inline size_t pad16(size_t b) { return (b + 15) & ~(size_t)15; }

// byte offsets of the sections of a blob
struct TileLayout {
  size_t coords, vinfo, rows, rowmeta, records, total;
  __host__ __device__ TileLayout(int R, int V, int C) {
    coords = sizeof(TileHeader);
    vinfo = coords + (size_t)V * 16;
    rows = vinfo + pad16((size_t)V * 4);
    rowmeta = rows + pad16((size_t)R * 8);
    records = rowmeta + pad16((size_t)R * 4);
    total = records + pad16((size_t)C * 8);
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

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

struct TileBuildArgs {
  const uint32_t* patch_row_ptr; const uint32_t* patch_rows; const uint32_t* patch_cell_ptr; const uint32_t* patch_cells;
  const uint32_t* cells; const double* vertices;
  const int32_t* rowptr; const int32_t* colind; const uint8_t* dir_row;
  uint32_t off_te; size_t to_local;          // global row = off_te + dof; local row = dof + to_local
  unsigned long long* sizes;                 // PASS 0: bytes of every blob
  const unsigned long long* blob_ptr; unsigned char* blob;   // PASS 1
  int* status;                               // 1: a tile exceeds a capacity (vertices, colours, row length, stage)
  unsigned int* maxima;                      // [0] rows, [1] vertices, [2] stage entries, [3] colours
};

// One CTA per tile. DIM = 2 (triangles). PASS 0 sizes the blob, PASS 1 writes it.
template <int PASS>
__global__ void __launch_bounds__(kTileThreads) k_tile_build(const TileBuildArgs A) {
  constexpr int NV = 3;
  extern __shared__ unsigned char tb_sm[];
  uint32_t* sorted = reinterpret_cast<uint32_t*>(tb_sm);                 // [kTileMaxRefs]
  uint32_t* uvert = sorted + kTileMaxRefs;                               // [512]
  uint32_t* vinfo = uvert + 512;                                         // [512]
  int* vlen = reinterpret_cast<int*>(vinfo + 512);                       // [512] row length of owned vertices -> stage offsets
  uint32_t* vmax = reinterpret_cast<uint32_t*>(vlen + 512);              // [512]
  uint32_t* used = vmax + 512;                                           // [512]
  int* scratch = reinterpret_cast<int*>(used + 512);                     // [kTileThreads + 1] + flags
  int* head = scratch + kTileThreads + 8;                                // [kTileMaxRefs] unique flags / offsets
  uint16_t* clv = reinterpret_cast<uint16_t*>(head + kTileMaxRefs);      // [kTileMaxRefs] local vertex of every cell ref
  uint8_t* color = reinterpret_cast<uint8_t*>(clv + kTileMaxRefs);       // [kTileMaxRefs / 3 + 1]
  uint16_t* order = reinterpret_cast<uint16_t*>(color + 1376);           // [1376] cells sorted by colour
  uint8_t* vdir = reinterpret_cast<uint8_t*>(order + 1376);              // [512] owned Dirichlet rows (identity rows)
  __shared__ int s_ncolors, s_left, s_bad;
  __shared__ int ccount[kTileMaxColors + 2];

  const size_t tile = blockIdx.x;
  const uint32_t r_begin = A.patch_row_ptr[tile], r_end = A.patch_row_ptr[tile + 1];
  const uint32_t c_begin = A.patch_cell_ptr[tile], c_end = A.patch_cell_ptr[tile + 1];
  const int R = (int)(r_end - r_begin), C = (int)(c_end - c_begin);
  if (threadIdx.x == 0) { s_bad = 0; s_ncolors = 0; }
  __syncthreads();
  if (C * NV > kTileMaxRefs || R > kTileRows * 2 || C > 1365) {
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
    vinfo[lv] = owned ? (uint32_t)t : kNoRow;      // owned ordinal for now
    vdir[lv] = (owned && A.dir_row[(size_t)A.off_te + v] != 0) ? 1 : 0;
    if (owned && len > 16) atomicExch(&s_bad, 1);  // 4-bit slots
  }
  __syncthreads();
  const int S = block_exclusive_scan(vlen, V, scratch);   // vlen[lv] = stage offset of the row of lv
  if (S > 0xfff0 || s_bad) {
    if (threadIdx.x == 0) { atomicExch(A.status, 1); if (PASS == 0) A.sizes[tile] = 0; }
    return;
  }
  // 4. local vertex of every cell reference (cells in the plan's order: ascending cell id)
  for (int t = threadIdx.x; t < C * NV; t += kTileThreads) {
    const uint32_t v = A.cells[(size_t)A.patch_cells[c_begin + t / NV] * NV + t % NV];
    clv[t] = (uint16_t)lower_bound_u32(uvert, V, v);
  }
  // 5. colouring: rounds of "local maxima of the hashed priority take the smallest colour free at their vertices"
  for (int lv = threadIdx.x; lv < V; lv += kTileThreads) { vmax[lv] = 0u; used[lv] = 0u; }
  for (int c = threadIdx.x; c < C; c += kTileThreads) color[c] = 0xff;
  if (threadIdx.x == 0) s_left = C;
  __syncthreads();
  for (int round = 0; round < 64; ++round) {
    if (s_left == 0) break;
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kTileThreads) {
      if (color[c] != 0xff) continue;
      const uint32_t pr = (mix32(A.patch_cells[c_begin + c]) & 0xfffff000u) | (uint32_t)(c + 1);   // distinct inside a tile
#pragma unroll
      for (int v = 0; v < NV; ++v) atomicMax(&vmax[clv[c * NV + v]], pr);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kTileThreads) {
      if (color[c] != 0xff) continue;
      const uint32_t pr = (mix32(A.patch_cells[c_begin + c]) & 0xfffff000u) | (uint32_t)(c + 1);
      bool top = true;
      uint32_t busy = 0u;
#pragma unroll
      for (int v = 0; v < NV; ++v) { const int lv = clv[c * NV + v]; top = top && vmax[lv] == pr; busy |= used[lv]; }
      if (!top) continue;
      const int col = __ffs(~busy) - 1;     // smallest colour not used at any of its vertices
      if (col < 0 || col >= kTileMaxColors) { atomicExch(&s_bad, 1); color[c] = 0; }
      else {
        color[c] = (uint8_t)col;
        // the winners of a round share no vertex (each vertex has one maximum): plain ORs would race only with other
        // winners of OTHER vertices of the same word -- use the atomic form
#pragma unroll
        for (int v = 0; v < NV; ++v) atomicOr(&used[clv[c * NV + v]], 1u << col);
        atomicMax(&s_ncolors, col + 1);
      }
      atomicSub(&s_left, 1);
    }
    __syncthreads();
    for (int lv = threadIdx.x; lv < V; lv += kTileThreads) vmax[lv] = 0u;
    __syncthreads();
  }
  if (s_left != 0 || s_bad) {
    if (threadIdx.x == 0) { atomicExch(A.status, 1); if (PASS == 0) A.sizes[tile] = 0; }
    return;
  }
  const int NC = s_ncolors;
  if (PASS == 0) {
    if (threadIdx.x == 0) {
      A.sizes[tile] = TileLayout(R, V, C).total;
      atomicMax(A.maxima + 0, (unsigned)R); atomicMax(A.maxima + 1, (unsigned)V);
      atomicMax(A.maxima + 2, (unsigned)S); atomicMax(A.maxima + 3, (unsigned)NC);
    }
    return;
  }
  // 6. cells sorted by colour (stable: ascending cell id inside a colour)
  if (threadIdx.x <= kTileMaxColors) ccount[threadIdx.x] = 0;
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kTileThreads) atomicAdd(&ccount[color[c]], 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int k = 0; k <= kTileMaxColors; ++k) { const int v = ccount[k]; ccount[k] = run; run += v; }
  }
  __syncthreads();
  // rank of a cell inside its colour = cells of the same colour before it (C <= 1365: a scan per thread is fine)
  for (int c = threadIdx.x; c < C; c += kTileThreads) {
    const int col = color[c];
    int rank = 0;
    for (int d = 0; d < c; ++d) rank += (color[d] == col) ? 1 : 0;
    order[ccount[col] + rank] = (uint16_t)c;
  }
  __syncthreads();
  // 7. write the blob
  unsigned char* out = A.blob + A.blob_ptr[tile];
  const TileLayout L(R, V, C);
  if (threadIdx.x == 0) {
    TileHeader H;
    H.n_rows = (uint16_t)R; H.n_verts = (uint16_t)V; H.n_cells = (uint16_t)C; H.n_colors = (uint16_t)NC;
    H.stage_len = (uint32_t)S; H.pad = 0;
    for (int k = 0; k <= kTileMaxColors + 1; ++k) H.color_off[k] = (uint16_t)(k <= kTileMaxColors ? ccount[k] : C);
    for (int k = NC; k <= kTileMaxColors + 1; ++k) H.color_off[k] = (uint16_t)C;
    for (int k = 0; k < 6; ++k) H.pad2[k] = 0;
    *reinterpret_cast<TileHeader*>(out) = H;
  }
  double2* o_coords = reinterpret_cast<double2*>(out + L.coords);
  uint32_t* o_vinfo = reinterpret_cast<uint32_t*>(out + L.vinfo);
  int2* o_rows = reinterpret_cast<int2*>(out + L.rows);
  uint32_t* o_meta = reinterpret_cast<uint32_t*>(out + L.rowmeta);
  unsigned long long* o_rec = reinterpret_cast<unsigned long long*>(out + L.records);
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
      o_meta[t] = (uint32_t)vlen[lv] | ((uint32_t)len << 16) | ((uint32_t)diag << 24) | (dir ? 0x80000000u : 0u);
      // stage offset | owned ordinal << 16 | owned | accumulates into A (identity rows take no contributions,
      // bilinear_form.hpp:183-186; the linear form is not filtered, linear_form.hpp:55-63)
      info = (uint32_t)vlen[lv] | (t << 16) | 0x40000000u | (dir ? 0u : 0x80000000u);
    }
    o_vinfo[lv] = info;
  }
  for (int e = threadIdx.x; e < C; e += kTileThreads) {
    const int c = order[e];
    unsigned long long rec = 0ull;
    uint32_t ids[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) { ids[v] = uvert[clv[c * NV + v]]; rec |= (unsigned long long)clv[c * NV + v] << (9 * v); }
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
        rec |= (unsigned long long)((lo - rs) & 0xf) << (27 + 4 * (i * NV + j));
      }
    }
    o_rec[e] = rec;
  }
}

size_t tile_build_smem() {
  return (size_t)kTileMaxRefs * 4 + 512 * 4 * 5 + (kTileThreads + 8) * 4 + (size_t)kTileMaxRefs * 4 + (size_t)kTileMaxRefs * 2 + 1376 + 1376 * 2 + 512 + 64;
}

// ------------------------------------------------------------------------------------------
// assembly kernel
// ------------------------------------------------------------------------------------------

struct TileAsm {
  const unsigned long long* blob_ptr; const unsigned char* blob;
  double* val; int init_mode;                  // 1: the values are uninitialised (pending clear()): write them fresh
  double* b; int b_mode;                       // fused linear form: 0 none, 1 b += ..., (b is indexed by local row)
  double b_c[3];                               // its reference weights: sum_t c_t mhat_i (constant coefficient, order 0)
};

// element matrix of a constant-coefficient form on a P1 triangle (without |K|): the W x W tensor
// G = [c00, c0r^T Jmt; Jmt^T cr0, Jmt^T C Jmt] contracted with the pre-integrated reference tensor in shared memory
__device__ __forceinline__ void tile_element(const FormDev& F, const double* sm_ref, const CellGeom<2>& g, double (&acc)[9]) {
  element_matrix_const<2, 3, 3>(F, sm_ref - F.off_ref, g, acc);
}

template <bool FUSE_B>
__global__ void __launch_bounds__(kTileThreads)
k_assemble_tiles(const __grid_constant__ FormDev F, const TileAsm T) {
  extern __shared__ __align__(16) double tsm[];
  __shared__ TileHeader H;
  const unsigned char* blob = T.blob + T.blob_ptr[blockIdx.x];
  if (threadIdx.x < sizeof(TileHeader) / 4) reinterpret_cast<uint32_t*>(&H)[threadIdx.x] = reinterpret_cast<const uint32_t*>(blob)[threadIdx.x];
  // reference tensor of the form (the last table)
  const int n_ref = F.n_stage - F.off_ref;
  double* s_ref = tsm;
  for (int i = threadIdx.x; i < n_ref; i += kTileThreads) s_ref[i] = __ldg(F.tables + F.off_ref + i);
  __syncthreads();
  const int R = H.n_rows, V = H.n_verts, C = H.n_cells, NC = H.n_colors, S = (int)H.stage_len;
  const TileLayout L(R, V, C);
  double2* s_xy = reinterpret_cast<double2*>(s_ref + ((n_ref + 1) & ~1));
  double* s_stage = reinterpret_cast<double*>(s_xy + V);
  double* s_b = s_stage + S;                                   // [R] when FUSE_B
  uint32_t* s_vinfo = reinterpret_cast<uint32_t*>(s_b + (FUSE_B ? R : 0));
  const double2* g_xy = reinterpret_cast<const double2*>(blob + L.coords);
  const uint32_t* g_vinfo = reinterpret_cast<const uint32_t*>(blob + L.vinfo);
  const int2* g_rows = reinterpret_cast<const int2*>(blob + L.rows);
  const uint32_t* g_meta = reinterpret_cast<const uint32_t*>(blob + L.rowmeta);
  const unsigned long long* g_rec = reinterpret_cast<const unsigned long long*>(blob + L.records);
  for (int i = threadIdx.x; i < V; i += kTileThreads) { s_xy[i] = __ldg(g_xy + i); s_vinfo[i] = __ldg(g_vinfo + i); }
  for (int i = threadIdx.x; i < S; i += kTileThreads) s_stage[i] = 0.0;
  if (FUSE_B) for (int i = threadIdx.x; i < R; i += kTileThreads) s_b[i] = 0.0;
  __syncthreads();
  if (T.init_mode)   // identity rows of the Dirichlet dofs
    for (int t = threadIdx.x; t < R; t += kTileThreads) {
      const uint32_t m = __ldg(g_meta + t);
      if (m & 0x80000000u) s_stage[(m & 0xffffu) + ((m >> 24) & 0x7fu)] = 1.0;
    }
  // cells in record order (sorted by colour), one chunk of kTileThreads cells at a time
  for (int c0 = 0; c0 < C; c0 += kTileThreads) {
    const int c = c0 + threadIdx.x;
    double acc[9];
    double vol = 0.0;
    uint32_t info[3] = {0u, 0u, 0u};
    unsigned long long rec = 0ull;
    int my_color = -1;
    if (c < C) {
      rec = __ldg(g_rec + c);
      const int l0 = (int)(rec & 0x1ff), l1 = (int)((rec >> 9) & 0x1ff), l2 = (int)((rec >> 18) & 0x1ff);
      const double2 p0 = s_xy[l0], p1 = s_xy[l1], p2 = s_xy[l2];
      CellGeom<2> g;
      g.v0[0] = p0.x; g.v0[1] = p0.y;
      g.J[0][0] = p1.x - p0.x; g.J[1][0] = p1.y - p0.y;
      g.J[0][1] = p2.x - p0.x; g.J[1][1] = p2.y - p0.y;
      finish_cell_geometry<2>(g);
      tile_element(F, s_ref, g, acc);
      vol = g.vol;
      info[0] = s_vinfo[l0]; info[1] = s_vinfo[l1]; info[2] = s_vinfo[l2];
      my_color = 0;
      while (c >= (int)H.color_off[my_color + 1]) ++my_color;
    }
    // colours present in this chunk
    int first = 0;
    while (c0 >= (int)H.color_off[first + 1]) ++first;
    const int last_cell = min(c0 + kTileThreads, C) - 1;
    int last = first;
    while (last_cell >= (int)H.color_off[last + 1]) ++last;
    for (int col = first; col <= last && col < NC; ++col) {
      if (my_color == col) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          if (info[i] & 0x80000000u) {
            double* row = s_stage + (info[i] & 0xffffu);
#pragma unroll
            for (int j = 0; j < 3; ++j) row[(rec >> (27 + 4 * (i * 3 + j))) & 0xf] += acc[i * 3 + j] * vol;
          }
          if (FUSE_B && (info[i] & 0x40000000u)) s_b[(info[i] >> 16) & 0x1ffu] += T.b_c[i] * vol;
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  // write-out: a warp takes 32 consecutive owned rows, splits them into runs of rows that are consecutive in the CSR
  // arrays and copies every run with coalesced stores
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t0 = warp * 32; t0 < R; t0 += (kTileThreads / 32) * 32) {
    const int t = t0 + lane;
    int rs = 0, len = 0, off = 0, lrow = 0;
    if (t < R) {
      const int2 rr = __ldg(g_rows + t); const uint32_t m = __ldg(g_meta + t);
      rs = rr.x; lrow = rr.y; off = (int)(m & 0xffffu); len = (int)((m >> 16) & 0xffu);
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
    if (FUSE_B && t < R) T.b[lrow] += s_b[t];
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
  unsigned max_rows = 0, max_verts = 0, max_stage = 0, max_colors = 0;
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
  T->max_rows = h_max[0]; T->max_verts = h_max[1]; T->max_stage = h_max[2]; T->max_colors = h_max[3];
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
    std::fprintf(stderr, "[tfelcu trace] tile plan: %zu tiles, %.1f MB, max rows %u verts %u stage %u colours %u\n", T->n_tile,
                 (double)total / 1e6, T->max_rows, T->max_verts, T->max_stage, T->max_colors);
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
  K.b = nullptr; K.b_mode = 0; K.b_c[0] = K.b_c[1] = K.b_c[2] = 0.0;
  if (v) {
    // sum_t c_t D^0 psi_i integrated on the reference cell: mhat[0][i] of the linear form's reference table
    const FormDev& G = fbv->F;
    double c = 0.0;
    for (int t = 0; t < G.n_terms; ++t) c += G.terms[t].c;
    for (int i = 0; i < 3; ++i) K.b_c[i] = c * fbv->tables[(size_t)G.off_ref + i];
    K.b = v->data.p; K.b_mode = 1;
  }
  if (!T->n_tile) { m->val_valid = true; return true; }
  const int n_ref = fb.F.n_stage - fb.F.off_ref;
  const size_t smem = ((size_t)((n_ref + 1) & ~1) + 2 * (size_t)T->max_verts + T->max_stage + (v ? T->max_rows : 0)) * sizeof(double) +
                      (size_t)T->max_verts * 4 + 16;
  if (v) {
    ensure_smem(k_assemble_tiles<true>, smem);
    k_assemble_tiles<true><<<(unsigned)T->n_tile, kTileThreads, smem, ctx->stream>>>(fb.F, K);
  } else {
    ensure_smem(k_assemble_tiles<false>, smem);
    k_assemble_tiles<false><<<(unsigned)T->n_tile, kTileThreads, smem, ctx->stream>>>(fb.F, K);
  }
  ctx->count(); TFELCU_LAUNCH_CHECK();
  m->val_valid = true;
  return true;
}

}  // namespace tfelcu
