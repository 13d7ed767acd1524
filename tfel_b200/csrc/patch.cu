// Patch assembly: bilinear_form::operator+= (src/core/bilinear_form.hpp:23-110) and linear_form::operator+=
// (linear_form.hpp:20-142) with the rows of a form grouped into compact patches of the mesh.
//
// The owner-computes kernel of assemble.cu gives every row to one thread, which recomputes the geometry and its row of
// the element matrix of each incident cell: (dim + 1)-fold redundant work for P1, and a block of consecutive rows in the
// reference numbering is a 1-D strip of the mesh, so sharing the element matrices inside a CTA still computes every
// cell twice. Here the rows of a CTA are a 2-D / 3-D patch instead: the dofs are sorted along a Morton curve of their
// space coordinates (fes.hpp:190-202) and the patches are the leaves of the quadtree / octree with at most
// rows_per_patch dofs (16 x 16 vertices on a uniform P1 grid). One CTA
//   1. computes the element matrix of every cell incident to its rows ONCE into shared memory (cells on the rim of the
//      patch are recomputed by the neighbouring patch: ~13 % for a 16 x 16 patch of P1 vertices),
//   2. lets every row sum its contributions in ascending cell order -- the order of the reference's cell loop -- into a
//      shared-memory image of its CSR row, through the same slot map as the owner-computes kernel,
//   3. writes the rows out, run of consecutive rows by run, coalesced; every stored value is written exactly once.
// No atomics, bitwise reproducible. The plan (patches, their cell lists, the transposed dof map in patch-local cell
// indices) is structure: built once per (space, row range) with radix sorts and reused by every += on it.
#include "element.cuh"
#include "sortutil.cuh"

#include <cstdlib>

namespace tfelcu {

// ------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ unsigned long long ordered_bits(double d) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(d);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

static double from_ordered_bits(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  double d; std::memcpy(&d, &b, sizeof(d));
  return d;
}

// bounding box of n points [n][dim]: box[d] = min, box[3 + d] = max (as order-preserving integers)
__global__ void k_bbox(const double* __restrict__ x, size_t n, int dim, unsigned long long* __restrict__ box) {
  unsigned long long lo[3] = {~0ull, ~0ull, ~0ull}, hi[3] = {0ull, 0ull, 0ull};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    for (int d = 0; d < dim; ++d) {
      const unsigned long long k = ordered_bits(x[i * dim + d]);
      lo[d] = k < lo[d] ? k : lo[d]; hi[d] = k > hi[d] ? k : hi[d];
    }
  for (int d = 0; d < dim; ++d) {
    for (int s = 16; s > 0; s >>= 1) {
      const unsigned long long ol = __shfl_xor_sync(0xffffffffu, lo[d], s), oh = __shfl_xor_sync(0xffffffffu, hi[d], s);
      lo[d] = ol < lo[d] ? ol : lo[d]; hi[d] = oh > hi[d] ? oh : hi[d];
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(box + d, lo[d]); atomicMax(box + 3 + d, hi[d]); }
  }
}

__device__ __forceinline__ uint32_t spread1(uint32_t v) {   // 16 bits -> every second bit
  v &= 0xffffu;
  v = (v | (v << 8)) & 0x00ff00ffu; v = (v | (v << 4)) & 0x0f0f0f0fu;
  v = (v | (v << 2)) & 0x33333333u; v = (v | (v << 1)) & 0x55555555u;
  return v;
}

__device__ __forceinline__ uint32_t spread2(uint32_t v) {   // 10 bits -> every third bit
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu; v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u; v = (v | (v << 2)) & 0x09249249u;
  return v;
}

struct Box { double lo[3], scale[3]; };

__global__ void k_iota_from(uint32_t* __restrict__ out, size_t n, uint32_t first) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = first + (uint32_t)i;
}

// key = Morton code of the quantised coordinates << 32 | index of the row in the range (ties: ascending dof).
// 2-D: 16 bits per axis, 3-D: 10 bits per axis; the extent maps onto 2^bits buckets, so the vertices of a uniform grid
// with a power-of-two number of intervals fall on bucket boundaries and quadtree nodes are exact squares of vertices.
__global__ void k_morton_keys(const double* __restrict__ x, size_t n, int dim, Box box, uint64_t* __restrict__ keys) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t qmax = dim == 2 ? 0xffffu : 0x3ffu;
  uint32_t q[3] = {0u, 0u, 0u};
  for (int d = 0; d < dim; ++d) {
    const double t = (x[i * dim + d] - box.lo[d]) * box.scale[d];
    q[d] = t <= 0.0 ? 0u : (t >= (double)qmax ? qmax : (uint32_t)t);
  }
  const uint32_t code = dim == 2 ? (spread1(q[0]) | (spread1(q[1]) << 1))
                                 : (spread2(q[0]) | (spread2(q[1]) << 1) | (spread2(q[2]) << 2));
  keys[i] = ((uint64_t)code << 32) | (uint64_t)i;
}

__device__ __forceinline__ size_t lower_bound_code(const uint64_t* __restrict__ sorted, size_t n, uint64_t code) {
  const uint64_t target = code << 32;   // code may be 2^32: target wraps to 0 only for code >= 2^32, handled by the caller
  size_t lo = 0, hi = n;
  while (lo < hi) {
    const size_t mid = lo + (hi - lo) / 2;
    if (sorted[mid] < target) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Leaves of the quadtree / octree over the sorted Morton codes with at most cap points: point t starts a patch when it
// is the first point of the coarsest node around it that holds <= cap points (more than cap points with one and the
// same code are cut every cap points).
__global__ void k_patch_starts(const uint64_t* __restrict__ sorted, size_t n, int dim, int cap, uint32_t* __restrict__ flag) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int levels = dim == 2 ? 16 : 10;
  const uint64_t code = sorted[t] >> 32;
  // node counts shrink with the level: binary search for the coarsest level whose node holds <= cap points
  int lo_l = 0, hi_l = levels;   // answer in [lo_l, hi_l]; level `levels` = one code
  size_t start = 0;
  auto node = [&](int level, size_t& first) -> size_t {
    const int shift = dim * (levels - level);
    const uint64_t prefix = code >> shift;
    first = lower_bound_code(sorted, n, prefix << shift);
    const uint64_t next = (prefix + 1) << shift;
    const size_t last = next >= (1ull << 32) ? n : lower_bound_code(sorted, n, next);
    return last - first;
  };
  while (lo_l < hi_l) {
    const int mid = (lo_l + hi_l) >> 1;
    size_t first;
    if (node(mid, first) <= (size_t)cap) hi_l = mid; else lo_l = mid + 1;
  }
  node(lo_l, start);
  flag[t] = ((t - start) % (size_t)cap == 0) ? 1u : 0u;
}

// the patch of every row of the range (patches numbered along the Morton curve)
__global__ void k_patch_of(const uint64_t* __restrict__ sorted, size_t n, const uint32_t* __restrict__ flag_sum,
                           uint32_t* __restrict__ patch_of, uint64_t* __restrict__ keys) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint32_t i = (uint32_t)(sorted[t] & 0xffffffffull);
  const uint32_t p = flag_sum[t] - 1u;   // inclusive sum of the start flags
  patch_of[i] = p;
  keys[i] = ((uint64_t)p << 32) | (uint64_t)i;
}

// rows of a patch in ascending dof order: consecutive dofs are consecutive CSR rows, which the write-out coalesces
__global__ void k_patch_rows(const uint64_t* __restrict__ sorted, size_t n, uint32_t row_begin, uint32_t* __restrict__ rows) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t < n) rows[t] = row_begin + (uint32_t)(sorted[t] & 0xffffffffull);
}

// (patch << 32 | cell) of every entry of the owned slice of the transposed dof map
__global__ void k_patch_pair_keys(const uint32_t* __restrict__ adj, size_t a_begin, size_t n_adj,
                                  const uint32_t* __restrict__ dof_map, int nd, uint32_t row_begin,
                                  const uint32_t* __restrict__ patch_of, uint64_t* __restrict__ keys) {
  const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= n_adj) return;
  const uint32_t code = adj[a_begin + p];
  const uint32_t r = dof_map[code];
  keys[p] = ((uint64_t)patch_of[r - row_begin] << 32) | (uint64_t)(code / (uint32_t)nd);
}

__global__ void k_patch_cell_ptr(const uint64_t* __restrict__ ukeys, size_t n_u, size_t n_patch,
                                 uint32_t* __restrict__ cell_ptr) {
  const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p > n_patch) return;
  cell_ptr[p] = (uint32_t)lower_bound_code(ukeys, n_u, (uint64_t)p);
}

__global__ void k_patch_cells(const uint64_t* __restrict__ ukeys, size_t n_u, uint32_t* __restrict__ cells) {
  const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (e < n_u) cells[e] = (uint32_t)(ukeys[e] & 0xffffffffull);
}

__global__ void k_patch_adj_local(const uint32_t* __restrict__ adj, size_t a_begin, size_t n_adj,
                                  const uint32_t* __restrict__ dof_map, int nd, uint32_t row_begin,
                                  const uint32_t* __restrict__ patch_of, const uint32_t* __restrict__ cell_ptr,
                                  const uint32_t* __restrict__ cells, uint32_t* __restrict__ adj_local) {
  const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= n_adj) return;
  const uint32_t code = adj[a_begin + p];
  const uint32_t k = code / (uint32_t)nd, i = code - k * (uint32_t)nd;
  const uint32_t patch = patch_of[dof_map[code] - row_begin];
  const uint32_t c0 = cell_ptr[patch];
  uint32_t lo = c0, hi = cell_ptr[patch + 1];
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (cells[mid] < k) lo = mid + 1; else hi = mid;
  }
  adj_local[p] = (lo - c0) * (uint32_t)nd + i;
}

__global__ void k_max_span(const uint32_t* __restrict__ ptr, size_t n, unsigned long long* __restrict__ out) {
  const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p < n) atomicMax(out, (unsigned long long)(ptr[p + 1] - ptr[p]));
}

PatchPlan* fes_patch_plan(tfelcu_fes* f, size_t row_begin, size_t row_end, int rows_per_patch) {
  for (auto& p : f->patch_plans)
    if (p->row_begin == row_begin && p->row_end == row_end && p->rows_per_patch == rows_per_patch) return p.get();
  Ctx* ctx = f->ctx;
  fes_transpose(f);
  std::unique_ptr<PatchPlan> P(new PatchPlan());
  P->row_begin = row_begin; P->row_end = row_end; P->rows_per_patch = rows_per_patch;
  const size_t n = row_end - row_begin;
  const int dim = f->mesh->dim;
  TFELCU_REQUIRE(rows_per_patch >= 1, "patch plan: bad patch size");
  if (n) {
    uint32_t a_lo = 0, a_hi = 0;
    ctx->download(&a_lo, f->adj_ptr.p + row_begin, 1);
    ctx->download(&a_hi, f->adj_ptr.p + row_end, 1);
    P->a_begin = a_lo; P->a_count = a_hi - a_lo;
    // 1. Morton order of the dof coordinates
    DevBuf<uint64_t> sorted(n);
    {
      DevBuf<uint32_t> ids(n);
      DevBuf<double> x(n * dim);
      k_iota_from<<<grid_for(n, 256), 256, 0, ctx->stream>>>(ids.p, n, (uint32_t)row_begin);
      ctx->count(); TFELCU_LAUNCH_CHECK();
      fes_dof_coordinates(f, ids.p, n, x.p);
      unsigned long long h_box[6] = {~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull};
      DevBuf<unsigned long long> box(6);
      ctx->upload(box.p, h_box, 6);
      unsigned g = grid_for(n, 256); if (g > 1184u) g = 1184u;
      k_bbox<<<g, 256, 0, ctx->stream>>>(x.p, n, dim, box.p);
      ctx->count(); TFELCU_LAUNCH_CHECK();
      ctx->download(h_box, box.p, 6);
      Box B;
      const double buckets = dim == 2 ? 65536.0 : 1024.0;
      for (int d = 0; d < 3; ++d) {
        B.lo[d] = 0.0; B.scale[d] = 0.0;
        if (d >= dim) continue;
        const double lo = from_ordered_bits(h_box[d]), hi = from_ordered_bits(h_box[3 + d]);
        B.lo[d] = lo;
        B.scale[d] = hi > lo ? buckets / (hi - lo) : 0.0;
      }
      DevBuf<uint64_t> keys(n);
      k_morton_keys<<<grid_for(n, 256), 256, 0, ctx->stream>>>(x.p, n, dim, B, keys.p);
      ctx->count(); TFELCU_LAUNCH_CHECK();
      sort_keys<uint64_t>(ctx, keys.p, sorted.p, n, 0, 64);
    }
    // 2. patches = tree leaves with at most rows_per_patch dofs
    DevBuf<uint32_t> patch_of(n);
    {
      DevBuf<uint32_t> flag(n), flag_sum(n);
      k_patch_starts<<<grid_for(n, 256), 256, 0, ctx->stream>>>(sorted.p, n, dim, rows_per_patch, flag.p);
      ctx->count(); TFELCU_LAUNCH_CHECK();
      inclusive_sum<uint32_t, uint32_t>(ctx, flag.p, flag_sum.p, n);
      uint32_t n_patch = 0;
      ctx->download(&n_patch, flag_sum.p + (n - 1), 1);
      P->n_patch = n_patch;
      P->rows.alloc(n); P->row_ptr.alloc((size_t)n_patch + 1);
      DevBuf<uint64_t> keys(n);
      k_patch_of<<<grid_for(n, 256), 256, 0, ctx->stream>>>(sorted.p, n, flag_sum.p, patch_of.p, keys.p);
      ctx->count(); TFELCU_LAUNCH_CHECK();
      sort_keys<uint64_t>(ctx, keys.p, sorted.p, n, 0, 32 + bits_for((size_t)n_patch + 1));
      k_patch_rows<<<grid_for(n, 256), 256, 0, ctx->stream>>>(sorted.p, n, (uint32_t)row_begin, P->rows.p);
      ctx->count(); TFELCU_LAUNCH_CHECK();
      k_patch_cell_ptr<<<grid_for((size_t)n_patch + 1, 256), 256, 0, ctx->stream>>>(sorted.p, n, n_patch, P->row_ptr.p);
      ctx->count(); TFELCU_LAUNCH_CHECK();
      ctx->sync();
    }
    sorted.release();
    // 3. cells of every patch and the transposed dof map in patch-local cell indices
    const size_t na = P->a_count;
    P->cell_ptr.alloc(P->n_patch + 1);
    P->adj_local.alloc(na);
    size_t n_u = 0;
    {
      DevBuf<uint64_t> keys(na), ukeys(na);
      if (na) {
        k_patch_pair_keys<<<grid_for(na, 256), 256, 0, ctx->stream>>>(f->adj.p, P->a_begin, na, f->dof_map, f->nd,
                                                                   (uint32_t)row_begin, patch_of.p, keys.p);
        ctx->count(); TFELCU_LAUNCH_CHECK();
        sort_keys<uint64_t>(ctx, keys.p, ukeys.p, na, 0, 32 + bits_for(P->n_patch + 1));
        n_u = unique_keys<uint64_t>(ctx, ukeys.p, keys.p, na);
      }
      TFELCU_REQUIRE(n_u < 0xffffffffull, "patch plan: too many (patch, cell) pairs");
      k_patch_cell_ptr<<<grid_for(P->n_patch + 1, 256), 256, 0, ctx->stream>>>(keys.p, n_u, P->n_patch, P->cell_ptr.p);
      ctx->count(); TFELCU_LAUNCH_CHECK();
      P->cells.alloc(n_u);
      if (n_u) {
        k_patch_cells<<<grid_for(n_u, 256), 256, 0, ctx->stream>>>(keys.p, n_u, P->cells.p);
        ctx->count(); TFELCU_LAUNCH_CHECK();
      }
      ctx->sync();
    }
    if (na) {
      k_patch_adj_local<<<grid_for(na, 256), 256, 0, ctx->stream>>>(f->adj.p, P->a_begin, na, f->dof_map, f->nd,
                                                                 (uint32_t)row_begin, patch_of.p, P->cell_ptr.p,
                                                                 P->cells.p, P->adj_local.p);
      ctx->count(); TFELCU_LAUNCH_CHECK();
    }
    DevBuf<unsigned long long> mx(1);
    ctx->zero(mx.p, 1);
    k_max_span<<<grid_for(P->n_patch, 256), 256, 0, ctx->stream>>>(P->cell_ptr.p, P->n_patch, mx.p);
    ctx->count(); TFELCU_LAUNCH_CHECK();
    unsigned long long h = 0; ctx->download(&h, mx.p, 1);
    P->max_cells = (size_t)h;
  }
  f->patch_plans.push_back(std::move(P));
  return f->patch_plans.back().get();
}

// offset of every row inside the shared-memory stage of its patch (rows of a patch back to back, in patch order)
__global__ void k_patch_stage_off(const uint32_t* __restrict__ row_ptr, size_t n_patch, const uint32_t* __restrict__ rows,
                                  size_t to_local, const int32_t* __restrict__ rowptr, uint32_t* __restrict__ stage_off,
                                  unsigned long long* __restrict__ max_stage) {
  const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= n_patch) return;
  uint32_t acc = 0;
  for (uint32_t t = row_ptr[p]; t < row_ptr[p + 1]; ++t) {
    const size_t lrow = rows[t] + to_local;
    stage_off[t] = acc;
    acc += (uint32_t)(rowptr[lrow + 1] - rowptr[lrow]);
  }
  atomicMax(max_stage, (unsigned long long)acc);
}

// rows per patch: enough cells per CTA to amortise the rim, small enough that the element matrices of a patch
// (max_cells * (nd_te * nd_tr | 1) doubles) and the stage fit a few CTAs per SM
static int patch_rows_for(int nn) {
  if (const char* e = std::getenv("TFELCU_PATCH_ROWS")) { const int v = std::atoi(e); if (v >= 8 && v <= 4096) return v; }
  if (nn <= 9) return 256;
  if (nn <= 36) return 128;
  return 64;
}

static int patch_threads() {
  if (const char* e = std::getenv("TFELCU_PATCH_THREADS")) { const int v = std::atoi(e); if (v >= 32 && v <= 640 && v % 32 == 0) return v; }
  return 256;
}

void matrix_build_patches(tfelcu_matrix* m) {
  if (m->has_patches) return;
  Ctx* ctx = m->ctx;
  const RowMap& R = m->rows;
  int nn = 1;
  for (int a = 0; a < m->test.n; ++a) for (int b = 0; b < m->trial.n; ++b)
    nn = std::max(nn, m->test.fes[a]->nd * m->trial.fes[b]->nd);
  const int cap = patch_rows_for(nn);
  m->patch_stage_off.clear();
  m->patch_stage_off.resize(R.n);
  DevBuf<unsigned long long> mx(1);
  for (int sg = 0; sg < R.n; ++sg) {
    const int a = R.comp[sg];
    tfelcu_fes* te = m->test.fes[a];
    const size_t rb = R.g_begin[sg] - m->test.offset[a], re = R.g_end[sg] - m->test.offset[a];
    PatchPlan* P = fes_patch_plan(te, rb, re, cap);
    m->patch_plan[sg] = P;
    m->patch_max_stage[sg] = 0;
    m->patch_stage_off[sg].alloc(re - rb);
    if (!P->n_patch) continue;
    ctx->zero(mx.p, 1);
    k_patch_stage_off<<<grid_for(P->n_patch, 128), 128, 0, ctx->stream>>>(P->row_ptr.p, P->n_patch, P->rows.p,
                                                                        (size_t)R.local_off[sg] - rb, m->rowptr.p,
                                                                        m->patch_stage_off[sg].p, mx.p);
    ctx->count(); TFELCU_LAUNCH_CHECK();
    unsigned long long h = 0; ctx->download(&h, mx.p, 1);
    m->patch_max_stage[sg] = (size_t)h;
  }
  m->has_patches = true;
}

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------

constexpr int kPatchMaxThreads = 640;

struct PatchBlock {
  const double* vertices; const uint32_t* cells;
  uint32_t off_te;
  size_t row_begin, local_off;
  const uint32_t* row_ptr; const uint32_t* prow; const uint32_t* stage_off;
  const uint32_t* cell_ptr; const uint32_t* pcells; const uint32_t* adj_local;
  const uint32_t* adj_ptr; size_t a_begin; const uint32_t* slots;
  const uint8_t* dir_row; const int32_t* rowptr; const int32_t* colind; double* val;
  int init_mode;     // 1: clear() is pending -- write the rows fresh (identity on Dirichlet rows); 0: add to the values
  int elem_doubles;  // shared-memory doubles reserved for the element matrices of a patch
  int max_stage;     // ... and for the image of its rows
};

template <int DIM, int ND_TE, int ND_TR, bool CONST_COEF>
__global__ void __launch_bounds__(kPatchMaxThreads) k_assemble_patch(const __grid_constant__ FormDev F, const PatchBlock B) {
  extern __shared__ double sm[];
  constexpr int NN = ND_TE * ND_TR, NNP = NN | 1;   // odd stride: conflict-free stores of one element matrix per thread
  for (int idx = (CONST_COEF ? F.off_ref : 0) + threadIdx.x; idx < F.n_stage; idx += blockDim.x) sm[idx] = __ldg(F.tables + idx);
  double* elem = sm + F.n_stage;
  double* stage = elem + B.elem_doubles;
  for (int e = threadIdx.x; e < B.max_stage; e += blockDim.x) stage[e] = 0.0;
  const uint32_t t0 = B.row_ptr[blockIdx.x];
  const int nrows = (int)(B.row_ptr[blockIdx.x + 1] - t0);
  const uint32_t c0 = B.cell_ptr[blockIdx.x];
  const int ncells = (int)(B.cell_ptr[blockIdx.x + 1] - c0);
  const size_t to_local = B.local_off - B.row_begin;
  __syncthreads();
  // 1. element matrices (times |K|) of the cells of the patch: bilinear_form.hpp:64-102
  for (int x = threadIdx.x; x < ncells; x += blockDim.x) {
    const size_t k = __ldg(B.pcells + c0 + x);
    CellGeom<DIM> g;
    load_cell_geometry<DIM>(B.vertices, B.cells + k * (DIM + 1), g);
    double* out = elem + (size_t)x * NNP;
    if constexpr (CONST_COEF && NN <= 36) {
      double acc[NN];
      element_matrix_const<DIM, ND_TE, ND_TR>(F, sm, g, acc);
#pragma unroll
      for (int e = 0; e < NN; ++e) out[e] = __dmul_rn(acc[e], g.vol);
    } else {
#pragma unroll 1
      for (int i = 0; i < ND_TE; ++i) {
        double acc[ND_TR];
        element_row<DIM, ND_TE, ND_TR, CONST_COEF>(F, sm, k, i, g, acc);
#pragma unroll
        for (int j = 0; j < ND_TR; ++j) out[i * ND_TR + j] = __dmul_rn(acc[j], g.vol);
      }
    }
  }
  __syncthreads();
  // 2. every row adds the contributions of its cells in ascending cell order (bilinear_form.hpp:103-108)
  for (int lr = threadIdx.x; lr < nrows; lr += blockDim.x) {
    const uint32_t r = __ldg(B.prow + t0 + lr);
    double* row = stage + __ldg(B.stage_off + t0 + lr);
    const size_t grow = (size_t)B.off_te + r;
    if (!B.dir_row[grow]) {
      const uint32_t a1 = __ldg(B.adj_ptr + r + 1);
      for (uint32_t a = __ldg(B.adj_ptr + r); a < a1; ++a) {
        const uint32_t code = __ldg(B.adj_local + (a - B.a_begin));
        const uint32_t lc = code / ND_TE; const int i = (int)(code - lc * ND_TE);
        const double* em = elem + (size_t)lc * NNP + i * ND_TR;
        constexpr int SW = slot_words(ND_TR);
        const uint32_t* sl = B.slots + (size_t)(a - B.a_begin) * SW;
        uint32_t w[SW];
#pragma unroll
        for (int t = 0; t < SW; ++t) w[t] = __ldg(sl + t);
#pragma unroll
        for (int j = 0; j < ND_TR; ++j) row[(w[j >> 2] >> (8 * (j & 3))) & 0xffu] += em[j];
      }
    } else if (B.init_mode) {
      const size_t lrow = r + to_local;
      const int rs = B.rowptr[lrow], re = B.rowptr[lrow + 1];
      row[find_in_row(B.colind, rs, re, (int32_t)grow) - rs] = 1.0;   // clear(): bilinear_form.hpp:159-165
    }
  }
  __syncthreads();
  // 3. write the rows: the rows of a warp's 32 slots split into runs of consecutive CSR rows, each copied coalesced
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int base = warp * 32; base < nrows; base += nwarps * 32) {
    const int lr = base + lane;
    const bool valid = lr < nrows;
    int rs = 0, len = 0; uint32_t off = 0;
    if (valid) {
      const size_t lrow = __ldg(B.prow + t0 + lr) + to_local;
      rs = B.rowptr[lrow]; len = B.rowptr[lrow + 1] - rs;
      off = __ldg(B.stage_off + t0 + lr);
    }
    const int prev_end = __shfl_up_sync(0xffffffffu, rs + len, 1);
    const bool start = valid && (lane == 0 || prev_end != rs);
    unsigned starts = __ballot_sync(0xffffffffu, start);
    const int n_valid = __popc(__ballot_sync(0xffffffffu, valid));
    while (starts) {
      const int s = __ffs(starts) - 1;
      starts &= starts - 1;
      const int e = starts ? (__ffs(starts) - 1) : n_valid;   // one past the last lane of the run
      const int run_rs = __shfl_sync(0xffffffffu, rs, s);
      const uint32_t run_off = __shfl_sync(0xffffffffu, off, s);
      const int total = __shfl_sync(0xffffffffu, rs + len, e - 1) - run_rs;
      if (B.init_mode) for (int x = lane; x < total; x += 32) B.val[run_rs + x] = stage[run_off + x];
      else for (int x = lane; x < total; x += 32) B.val[run_rs + x] += stage[run_off + x];
    }
  }
}

struct VecPatchBlock {
  const double* vertices; const uint32_t* cells;
  size_t row_begin, local_off;
  const uint32_t* row_ptr; const uint32_t* prow;
  const uint32_t* cell_ptr; const uint32_t* pcells; const uint32_t* adj_local;
  const uint32_t* adj_ptr; size_t a_begin;
  double* b;
};

template <int DIM, int ND_TE, bool CONST_COEF>
__global__ void __launch_bounds__(kPatchMaxThreads) k_assemble_vector_patch(const __grid_constant__ FormDev F, const VecPatchBlock B) {
  extern __shared__ double sm[];
  constexpr int NDP = ND_TE | 1;
  for (int idx = threadIdx.x; idx < F.n_stage; idx += blockDim.x) sm[idx] = __ldg(F.tables + idx);
  double* elem = sm + F.n_stage;
  const uint32_t t0 = B.row_ptr[blockIdx.x];
  const int nrows = (int)(B.row_ptr[blockIdx.x + 1] - t0);
  const uint32_t c0 = B.cell_ptr[blockIdx.x];
  const int ncells = (int)(B.cell_ptr[blockIdx.x + 1] - c0);
  __syncthreads();
  // element vectors (times |K|): linear_form.hpp:64-90
  for (int x = threadIdx.x; x < ncells; x += blockDim.x) {
    const size_t k = __ldg(B.pcells + c0 + x);
    CellGeom<DIM> g;
    load_cell_geometry<DIM>(B.vertices, B.cells + k * (DIM + 1), g);
    if constexpr (CONST_COEF) {
      const double* ref = sm + F.off_ref;
      double Ts[DIM + 1];
      const_test_weights<DIM>(F, g, Ts);
#pragma unroll
      for (int i = 0; i < ND_TE; ++i) {
        double v = 0.0;
#pragma unroll
        for (int a = 0; a <= DIM; ++a) v += Ts[a] * ref[a * ND_TE + i];
        elem[(size_t)x * NDP + i] = __dmul_rn(v, g.vol);
      }
    } else {
#pragma unroll 1
      for (int i = 0; i < ND_TE; ++i)
        elem[(size_t)x * NDP + i] = __dmul_rn(element_entry<DIM, ND_TE, CONST_COEF>(F, sm, k, i, g), g.vol);
    }
  }
  __syncthreads();
  // ordered scatter of linear_form.hpp:55-63 == ascending cell order per dof
  for (int lr = threadIdx.x; lr < nrows; lr += blockDim.x) {
    const uint32_t r = __ldg(B.prow + t0 + lr);
    double sum = 0.0;
    const uint32_t a1 = __ldg(B.adj_ptr + r + 1);
    for (uint32_t a = __ldg(B.adj_ptr + r); a < a1; ++a) {
      const uint32_t code = __ldg(B.adj_local + (a - B.a_begin));
      const uint32_t lc = code / ND_TE;
      sum += elem[(size_t)lc * NDP + (code - lc * ND_TE)];
    }
    B.b[B.local_off + (r - B.row_begin)] += sum;
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------

constexpr size_t kMaxPatchSmem = 200 * 1024;

template <int DIM, int ND_TE, int ND_TR>
static void launch_patch(Ctx* ctx, const FormBuild& fb, PatchBlock& B, const PatchPlan* P, size_t max_stage) {
  if (!P->n_patch) return;
  constexpr int NNP = (ND_TE * ND_TR) | 1;
  B.elem_doubles = (int)(P->max_cells * NNP);
  B.max_stage = (int)max_stage;
  const size_t smem = ((size_t)fb.F.n_stage + B.elem_doubles + max_stage) * sizeof(double);
  TFELCU_REQUIRE(smem <= kMaxPatchSmem, "patch assembly: the element matrices of a patch do not fit in shared memory");
  TFELCU_REQUIRE(P->n_patch <= 0x7fffffffull, "patch assembly: too many patches");
  const int threads = patch_threads();
  if (fb.const_coef) {
    auto kern = k_assemble_patch<DIM, ND_TE, ND_TR, true>;
    ensure_smem(kern, smem);
    kern<<<(unsigned)P->n_patch, threads, smem, ctx->stream>>>(fb.F, B);
  } else {
    auto kern = k_assemble_patch<DIM, ND_TE, ND_TR, false>;
    ensure_smem(kern, smem);
    kern<<<(unsigned)P->n_patch, threads, smem, ctx->stream>>>(fb.F, B);
  }
  ctx->count(); TFELCU_LAUNCH_CHECK();
}

template <int DIM>
static void assemble_bilinear_patch_dim(tfelcu_matrix* m, int quad, const tfelcu_factor* factors, int n_factors,
                                        const tfelcu_term* terms, int n_terms) {
  Ctx* ctx = m->ctx;
  tfelcu_mesh* mesh = m->mesh;
  if (!m->has_slots) {
    if (m->slots_failed) refuse("patch assembly needs the slot map (a row has more than 256 stored entries)");
    try { matrix_build_slots(m); } catch (const Refusal&) { m->has_slots = false; m->slots_failed = true; throw; }
  }
  matrix_build_patches(m);
  const bool single_block = (m->test.n == 1 && m->trial.n == 1);
  // pass 1: lower every block and check that it fits, so that a refusal happens before any value is touched
  std::vector<std::unique_ptr<FormBuild>> forms((size_t)m->test.n * m->trial.n);
  for (int a = 0; a < m->test.n; ++a)
    for (int b = 0; b < m->trial.n; ++b) {
      tfelcu_fes* te = m->test.fes[a]; tfelcu_fes* tr = m->trial.fes[b];
      std::unique_ptr<FormBuild> fb(new FormBuild());
      build_form(ctx, mesh, quad, factors, n_factors, terms, n_terms, a, te->fe, b, tr->fe, *fb);
      if (fb->F.n_terms == 0) continue;   // the block only contributes (already stored) zeros
      for (int sg = 0; sg < m->rows.n; ++sg) {
        if (m->rows.comp[sg] != a) continue;
        const size_t nnp = (size_t)(te->nd * tr->nd) | 1;
        const size_t smem = ((size_t)fb->F.n_stage + m->patch_plan[sg]->max_cells * nnp + m->patch_max_stage[sg]) * sizeof(double);
        if (smem > kMaxPatchSmem) refuse("patch assembly: the element matrices of a patch do not fit in shared memory");
      }
      forms[(size_t)a * m->trial.n + b] = std::move(fb);
    }
  // pass 2
  for (int a = 0; a < m->test.n; ++a)
    for (int b = 0; b < m->trial.n; ++b) {
      if (!forms[(size_t)a * m->trial.n + b]) continue;
      FormBuild& fb = *forms[(size_t)a * m->trial.n + b];
      tfelcu_fes* te = m->test.fes[a]; tfelcu_fes* tr = m->trial.fes[b];
      fb.F.tables = upload_form_tables(ctx, fb.tables);
      // a form with one block covers every owned row: the pending clear() is folded into the write-out
      const bool fresh = single_block && !m->val_valid;
      if (!fresh) matrix_ensure_cleared(m);
      PatchBlock B;
      B.vertices = mesh->vertices.p; B.cells = mesh->cells.p;
      B.off_te = (uint32_t)m->test.offset[a];
      B.adj_ptr = te->adj_ptr.p; B.dir_row = m->dir_row.p;
      B.rowptr = m->rowptr.p; B.colind = m->colind.p; B.val = m->val.p;
      B.init_mode = fresh ? 1 : 0;
      const int nd_te = te->nd, nd_tr = tr->nd;
      for (int sg = 0; sg < m->rows.n; ++sg) {
        if (m->rows.comp[sg] != a) continue;
        const PatchPlan* P = m->patch_plan[sg];
        B.row_begin = m->rows.g_begin[sg] - m->test.offset[a];
        B.local_off = m->rows.local_off[sg];
        B.row_ptr = P->row_ptr.p; B.prow = P->rows.p; B.stage_off = m->patch_stage_off[sg].p;
        B.cell_ptr = P->cell_ptr.p; B.pcells = P->cells.p; B.adj_local = P->adj_local.p;
        B.a_begin = m->a_begin[sg];
        B.slots = m->slots[(size_t)sg * m->trial.n + b].p;
        TFELCU_DISPATCH_ND(DIM, nd_te, NTE,
          TFELCU_DISPATCH_ND(DIM, nd_tr, NTR, launch_patch<DIM, NTE, NTR>(ctx, fb, B, P, m->patch_max_stage[sg]);))
      }
      if (fresh) m->val_valid = true;
    }
  ctx->sync();   // device copies of host coefficient data die with the forms
  matrix_ensure_cleared(m);
  m->pristine = false;
}

// Throws when the patch strategy does not apply to this form (the caller may fall back to owner-computes rows); a
// failure leaves the values untouched: every check happens before the first launch of a block.
bool assemble_bilinear_patch(tfelcu_matrix* m, int quad, const tfelcu_factor* factors, int n_factors,
                             const tfelcu_term* terms, int n_terms) {
  if (m->mesh->dim == 2) assemble_bilinear_patch_dim<2>(m, quad, factors, n_factors, terms, n_terms);
  else assemble_bilinear_patch_dim<3>(m, quad, factors, n_factors, terms, n_terms);
  return true;
}

template <int DIM, int ND_TE>
static void launch_vector_patch(Ctx* ctx, const FormBuild& fb, const VecPatchBlock& B, const PatchPlan* P) {
  if (!P->n_patch) return;
  constexpr int NDP = ND_TE | 1;
  const size_t smem = ((size_t)fb.F.n_stage + P->max_cells * NDP) * sizeof(double);
  if (smem > kMaxPatchSmem) refuse("patch assembly: the element vectors of a patch do not fit in shared memory");
  const int threads = patch_threads();
  if (fb.const_coef) {
    auto kern = k_assemble_vector_patch<DIM, ND_TE, true>;
    ensure_smem(kern, smem);
    kern<<<(unsigned)P->n_patch, threads, smem, ctx->stream>>>(fb.F, B);
  } else {
    auto kern = k_assemble_vector_patch<DIM, ND_TE, false>;
    ensure_smem(kern, smem);
    kern<<<(unsigned)P->n_patch, threads, smem, ctx->stream>>>(fb.F, B);
  }
  ctx->count(); TFELCU_LAUNCH_CHECK();
}

template <int DIM>
static void assemble_linear_patch_dim(tfelcu_vector* v, int quad, const tfelcu_factor* factors, int n_factors,
                                      const tfelcu_term* terms, int n_terms) {
  Ctx* ctx = v->ctx;
  for (int a = 0; a < v->space.n; ++a) {
    tfelcu_fes* te = v->space.fes[a];
    FormBuild fb;
    build_form(ctx, te->mesh, quad, factors, n_factors, terms, n_terms, a, te->fe, -1, 0, fb);
    if (fb.F.n_terms == 0) continue;
    fb.F.tables = upload_form_tables(ctx, fb.tables);
    fes_transpose(te);
    VecPatchBlock B;
    B.vertices = te->mesh->vertices.p; B.cells = te->mesh->cells.p;
    B.adj_ptr = te->adj_ptr.p; B.b = v->data.p;
    const int nd_te = te->nd;
    const int cap = patch_rows_for(nd_te * nd_te);   // the plan of a matrix on the same space is shared
    for (int sg = 0; sg < v->rows.n; ++sg) {
      if (v->rows.comp[sg] != a) continue;
      B.row_begin = v->rows.g_begin[sg] - v->space.offset[a];
      const size_t row_end = v->rows.g_end[sg] - v->space.offset[a];
      B.local_off = v->rows.local_off[sg];
      const PatchPlan* P = fes_patch_plan(te, B.row_begin, row_end, cap);
      B.row_ptr = P->row_ptr.p; B.prow = P->rows.p;
      B.cell_ptr = P->cell_ptr.p; B.pcells = P->cells.p; B.adj_local = P->adj_local.p;
      B.a_begin = P->a_begin;
      TFELCU_DISPATCH_ND(DIM, nd_te, NTE, launch_vector_patch<DIM, NTE>(ctx, fb, B, P);)
    }
    if (!fb.keep.empty()) ctx->sync();
  }
}

bool assemble_linear_patch(tfelcu_vector* v, int quad, const tfelcu_factor* factors, int n_factors,
                           const tfelcu_term* terms, int n_terms) {
  if (v->space.fes[0]->mesh->dim == 2) assemble_linear_patch_dim<2>(v, quad, factors, n_factors, terms, n_terms);
  else assemble_linear_patch_dim<3>(v, quad, factors, n_factors, terms, n_terms);
  return true;
}

}  // namespace tfelcu
