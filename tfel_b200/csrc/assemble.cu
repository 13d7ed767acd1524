// Element assembly on device: bilinear_form::operator+= (src/core/bilinear_form.hpp:23-110, composite:
// composite_bilinear_form.hpp:68-184), linear_form::operator+= (linear_form.hpp:20-142, composite:
// composite_linear_form.hpp:38-143) and the rank-0 integrate (form.hpp:87-125).
//
// The integrand arrives in sum-of-products form (include/tfel_cuda.h: tfelcu_term). Two scatter strategies,
// both free of atomics and bitwise reproducible:
//   * owner-computes rows (default): one thread per matrix row walks the cells incident to its dof in
//     ascending cell order -- the order in which the reference's cell loop adds into its std::map -- and
//     recomputes its row of each element matrix; every stored value is written exactly once per +=.
//   * colour-by-colour scatter: cells of one colour share no test dof; one launch per colour, one thread
//     per cell, read-modify-write into the CSR values.
// Reference-element tables (quadrature weights, basis values / gradients, and for constant-coefficient
// forms the pre-integrated reference tensor) are staged in shared memory; per-cell Jacobians are recomputed
// from the vertex coordinates (no J^{-T} array is materialised, cf. mesh.hpp:581-586).
#include "element.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace tfelcu {

struct RowBlock {
  const double* vertices; const uint32_t* cells;
  const uint32_t* map_tr; uint32_t off_te, off_tr;
  size_t row_begin, row_end, local_off;   // owned dofs of the test component and the local row of the first one
  const uint32_t* adj_ptr; const uint32_t* adj;
  const uint8_t* dir_row; const int32_t* rowptr; const int32_t* colind; double* val;
  int init_mode;   // STAGED only: 1 = values are uninitialised (pending clear()): write them fresh
  const uint32_t* slots; size_t a_begin;   // slot map of this (segment, trial component), or nullptr: search colind
  const uint4* rec;                        // P1 x P1 triangles: packed (cell vertices | slots + local dof) records, or nullptr
};


// REC (constant-coefficient P1 x P1 triangle blocks): everything a (row, cell) step needs besides the coordinates -- the
// three vertex ids, the three slot bytes and the local dof -- comes from ONE 16-byte record instead of five scattered
// 4-byte reads (transposed dof map, slot word, three ids of the cell): the per-thread gathers were what kept the L1 busy.
// Measured on configs[1]: A 0.974 -> 0.825 ms. The same record for the linear form was slower (0.529 -> 0.560 ms: it
// doubles the bytes that kernel streams from DRAM) and is not used.
template <int DIM, int ND_TE, int ND_TR, bool CONST_COEF, bool STAGED, bool REC = false>
__global__ void __launch_bounds__(kRowsPerCta) k_assemble_rows(const __grid_constant__ FormDev F, const RowBlock B) {
  extern __shared__ double sm[];
  // constant-coefficient forms only read the pre-integrated reference tensor (the last table)
  for (int idx = (CONST_COEF ? F.off_ref : 0) + threadIdx.x; idx < F.n_stage; idx += blockDim.x) sm[idx] = __ldg(F.tables + idx);
  double* stage = sm + F.n_stage;
  const size_t r0 = B.row_begin + (size_t)blockIdx.x * blockDim.x;   // one thread per row; blockDim.x = rows per CTA
  size_t r1 = r0 + blockDim.x; if (r1 > B.row_end) r1 = B.row_end;
  const size_t to_local = B.local_off - B.row_begin;   // local row = dof + to_local (mod 2^64)
  int seg0 = 0, seg_len = 0;
  if constexpr (STAGED) {
    seg0 = B.rowptr[r0 + to_local];
    seg_len = B.rowptr[r1 + to_local] - seg0;
    for (int e = threadIdx.x; e < seg_len; e += blockDim.x) stage[e] = 0.0;
  }
  __syncthreads();
  const size_t r = r0 + threadIdx.x;
  if (r < r1) {
    const size_t grow = B.off_te + r;
    const int rs = B.rowptr[r + to_local], re = B.rowptr[r + to_local + 1];
    if (!B.dir_row[grow]) {
      // bilinear_form.hpp:64-109 restricted to the cells that touch this row, in ascending cell order
      for (uint32_t a = B.adj_ptr[r]; a < B.adj_ptr[r + 1]; ++a) {
        if constexpr (REC) {
          const uint4 q = __ldg(B.rec + (a - B.a_begin));
          CellGeom<2> g;
          triangle_geometry_from_ids(B.vertices, q.x, q.y, q.z, g);
          double acc[ND_TR];
          element_row<DIM, ND_TE, ND_TR, CONST_COEF>(F, sm, 0, (int)(q.w >> 24), g, acc);
#pragma unroll
          for (int j = 0; j < ND_TR; ++j) {
            const int pos = rs + (int)((q.w >> (8 * j)) & 0xffu);
            if constexpr (STAGED) stage[pos - seg0] += acc[j] * g.vol;
            else B.val[pos] += acc[j] * g.vol;
          }
          continue;
        }
        const uint32_t code = B.adj[a];
        const size_t k = code / ND_TE; const int i = (int)(code - k * ND_TE);
        CellGeom<DIM> g;
        load_cell_geometry<DIM>(B.vertices, B.cells + k * (DIM + 1), g);
        double acc[ND_TR];
        element_row<DIM, ND_TE, ND_TR, CONST_COEF>(F, sm, k, i, g, acc);
        if (B.slots) {
          constexpr int SW = slot_words(ND_TR);
          const uint32_t* sl = B.slots + (size_t)(a - B.a_begin) * SW;
          uint32_t w[SW];
#pragma unroll
          for (int t = 0; t < SW; ++t) w[t] = __ldg(sl + t);
#pragma unroll
          for (int j = 0; j < ND_TR; ++j) {
            const int pos = rs + (int)((w[j >> 2] >> (8 * (j & 3))) & 0xffu);
            if constexpr (STAGED) stage[pos - seg0] += acc[j] * g.vol;
            else B.val[pos] += acc[j] * g.vol;
          }
        } else {
#pragma unroll
          for (int j = 0; j < ND_TR; ++j) {
            const int32_t col = (int32_t)(B.off_tr + B.map_tr[k * ND_TR + j]);
            const int pos = find_in_row(B.colind, rs, re, col);
            if constexpr (STAGED) stage[pos - seg0] += acc[j] * g.vol;
            else B.val[pos] += acc[j] * g.vol;
          }
        }
      }
    } else if (STAGED && B.init_mode) {
      stage[find_in_row(B.colind, rs, re, (int32_t)grow) - seg0] = 1.0;   // clear(): bilinear_form.hpp:159-165
    }
  }
  if constexpr (STAGED) {
    __syncthreads();
    if (B.init_mode) for (int e = threadIdx.x; e < seg_len; e += blockDim.x) B.val[seg0 + e] = stage[e];
    else for (int e = threadIdx.x; e < seg_len; e += blockDim.x) B.val[seg0 + e] += stage[e];
  }
}

struct ColorBlock {
  const double* vertices; const uint32_t* cells;
  const uint32_t* map_te; const uint32_t* map_tr; uint32_t off_te, off_tr;
  const uint32_t* color_cells; size_t n_cells;
  size_t row_begin, row_end, local_off;
  const uint8_t* dir_row; const int32_t* rowptr; const int32_t* colind; double* val;
};

template <int DIM, int ND_TE, int ND_TR, bool CONST_COEF>
__global__ void __launch_bounds__(128) k_assemble_colored(const __grid_constant__ FormDev F, const ColorBlock B) {
  extern __shared__ double sm[];
  for (int idx = threadIdx.x; idx < F.n_stage; idx += blockDim.x) sm[idx] = __ldg(F.tables + idx);
  __syncthreads();
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B.n_cells) return;
  const size_t k = B.color_cells[t];
  CellGeom<DIM> g;
  load_cell_geometry<DIM>(B.vertices, B.cells + k * (DIM + 1), g);
  for (int i = 0; i < ND_TE; ++i) {
    const size_t r = B.map_te[k * ND_TE + i];
    if (r < B.row_begin || r >= B.row_end) continue;   // another process owns the row
    const size_t grow = B.off_te + r;
    if (B.dir_row[grow]) continue;   // bilinear_form.hpp:183-186
    const size_t lrow = B.local_off + (r - B.row_begin);
    const int rs = B.rowptr[lrow], re = B.rowptr[lrow + 1];
    double acc[ND_TR];
    element_row<DIM, ND_TE, ND_TR, CONST_COEF>(F, sm, k, i, g, acc);
#pragma unroll
    for (int j = 0; j < ND_TR; ++j) {
      const int32_t col = (int32_t)(B.off_tr + B.map_tr[k * ND_TR + j]);
      B.val[find_in_row(B.colind, rs, re, col)] += acc[j] * g.vol;
    }
  }
}

struct VecBlock {
  const double* vertices; const uint32_t* cells;
  size_t row_begin, row_end, local_off;
  const uint32_t* adj_ptr; const uint32_t* adj;
  double* b;
};

template <int DIM, int ND_TE, bool CONST_COEF>
__global__ void __launch_bounds__(128) k_assemble_vector_rows(const __grid_constant__ FormDev F, const VecBlock B) {
  extern __shared__ double sm[];
  for (int idx = threadIdx.x; idx < F.n_stage; idx += blockDim.x) sm[idx] = __ldg(F.tables + idx);
  __syncthreads();
  const size_t r = B.row_begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B.row_end) return;
  double sum = 0.0;
  // linear_form.hpp:55-63: ordered scatter == ascending cell order per dof
  for (uint32_t a = B.adj_ptr[r]; a < B.adj_ptr[r + 1]; ++a) {
    const uint32_t code = B.adj[a];
    const size_t k = code / ND_TE; const int i = (int)(code - k * ND_TE);
    CellGeom<DIM> g;
    load_cell_geometry<DIM>(B.vertices, B.cells + k * (DIM + 1), g);
    sum += element_entry<DIM, ND_TE, CONST_COEF>(F, sm, k, i, g) * g.vol;
  }
  B.b[B.local_off + (r - B.row_begin)] += sum;
}

template <int DIM>
__global__ void __launch_bounds__(256) k_integrate(const __grid_constant__ FormDev F, const double* __restrict__ vertices,
                                                   const uint32_t* __restrict__ cells, size_t K,
                                                   double* __restrict__ partial) {
  extern __shared__ double sm[];
  for (int idx = threadIdx.x; idx < F.n_stage; idx += blockDim.x) sm[idx] = __ldg(F.tables + idx);
  __shared__ double red[256];
  __syncthreads();
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  double el = 0.0;
  if (k < K) {
    CellGeom<DIM> g;
    load_cell_geometry<DIM>(vertices, cells + k * (DIM + 1), g);
    const double* wq = sm + F.off_w;
    for (int q = 0; q < F.nq; ++q) {   // form.hpp:113-121
      double fv[kMaxFactors];
      eval_factors<DIM>(F, sm, k, q, g, fv);
      double f = 0.0;
      for (int t = 0; t < F.n_terms; ++t) f += term_coefficient(F.terms[t], fv);
      el += wq[q] * f;
    }
    el *= g.vol;
  }
  red[threadIdx.x] = el;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {   // fixed tree: reproducible
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256) k_sum_partials(const double* __restrict__ partial, size_t n, double* __restrict__ out) {
  __shared__ double red[256];
  double s = 0.0;
  for (size_t i = threadIdx.x; i < n; i += 256) s += partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = red[0];
}

// ------------------------------------------------------------------------------------------
// host side: lowering of the C-ABI description to FormDev
// ------------------------------------------------------------------------------------------


static void tabulate(int dim, int fe, const Quadrature& Q, std::vector<double>& out) {
  const int nd = fe_ndof(dim, fe), w = dim + 1;
  const size_t at = out.size();
  out.resize(at + (size_t)Q.nq * nd * w);
  for (int q = 0; q < Q.nq; ++q) fe_eval_hat(dim, fe, &Q.x[(size_t)q * dim], &out[at + (size_t)q * nd * w]);
}

// terms of block (test component a, trial component b); trial_comp < 0 selects linear-form terms
void build_form(Ctx* ctx, tfelcu_mesh* mesh, int quad, const tfelcu_factor* factors, int n_factors,
                       const tfelcu_term* terms, int n_terms, int comp_te, int fe_te, int comp_tr, int fe_tr,
                       FormBuild& out) {
  const int dim = mesh->dim, w = dim + 1;
  const Quadrature Q = make_quadrature(quad);
  TFELCU_REQUIRE(Q.dim == dim, "quadrature formula and mesh have different dimensions");
  TFELCU_REQUIRE(n_factors <= kMaxFactors, "too many coefficient factors in one integrand");
  FormDev& F = out.F;
  std::memset(&F, 0, sizeof(F));
  F.nq = Q.nq;
  // terms of this block
  bool used[kMaxFactors] = {false};
  for (int t = 0; t < n_terms; ++t) {
    const tfelcu_term& s = terms[t];
    if (comp_te >= 0 && s.test_comp != comp_te) continue;
    if (comp_tr >= 0 && s.trial_comp != comp_tr) continue;
    TFELCU_REQUIRE(F.n_terms < kMaxTerms, "too many terms in one block of the integrand");
    TFELCU_REQUIRE(s.test_d >= 0 && s.test_d <= dim && s.trial_d >= 0 && s.trial_d <= dim, "derivative index out of range");
    TFELCU_REQUIRE(s.n_factors >= 0 && s.n_factors <= 5, "too many factors in one term");
    DevTerm& d = F.terms[F.n_terms++];
    d.test_d = s.test_d; d.trial_d = s.trial_d; d.c = s.c; d.n_factors = s.n_factors;
    for (int f = 0; f < s.n_factors; ++f) {
      TFELCU_REQUIRE(s.factor[f] >= 0 && s.factor[f] < n_factors, "factor index out of range");
      d.factor[f] = s.factor[f]; used[s.factor[f]] = true;
      out.const_coef = false;
    }
  }
  if (out.const_coef) {
    for (int t = 0; t < F.n_terms; ++t) {
      const DevTerm& d = F.terms[t];
      if (d.test_d == 0 && d.trial_d == 0) { F.has00 = 1; F.c00 += d.c; }
      else if (d.test_d == 0) { F.has0r = 1; F.c0r[d.trial_d - 1] += d.c; }
      else if (d.trial_d == 0) { F.hasr0 = 1; F.cr0[d.test_d - 1] += d.c; }
      else { F.hasrr = 1; F.C[d.test_d - 1][d.trial_d - 1] += d.c; }
    }
  }
  // tables staged in shared memory
  std::vector<double>& T = out.tables;
  F.off_w = (int)T.size(); T.insert(T.end(), Q.w.begin(), Q.w.end());
  F.off_xhat = (int)T.size(); T.insert(T.end(), Q.x.begin(), Q.x.end());
  if (fe_te > 0) { F.off_hat_te = (int)T.size(); tabulate(dim, fe_te, Q, T); }
  if (fe_tr > 0) { F.off_hat_tr = (int)T.size(); tabulate(dim, fe_tr, Q, T); }
  if (out.const_coef && fe_te > 0) {
    const int nd_te = fe_ndof(dim, fe_te);
    const double* hte = &T[F.off_hat_te];
    F.off_ref = (int)T.size();
    if (fe_tr > 0) {
      const int nd_tr = fe_ndof(dim, fe_tr);
      const double* htr = &T[F.off_hat_tr];
      std::vector<double> ref((size_t)w * w * nd_te * nd_tr, 0.0);
      for (int a = 0; a < w; ++a) for (int b = 0; b < w; ++b)
        for (int i = 0; i < nd_te; ++i) for (int j = 0; j < nd_tr; ++j) {
          double s = 0.0;
          for (int q = 0; q < Q.nq; ++q) s += Q.w[q] * hte[((size_t)q * nd_te + i) * w + a] * htr[((size_t)q * nd_tr + j) * w + b];
          ref[(((size_t)a * w + b) * nd_te + i) * nd_tr + j] = s;
        }
      T.insert(T.end(), ref.begin(), ref.end());
    } else {
      std::vector<double> ref((size_t)w * nd_te, 0.0);
      for (int a = 0; a < w; ++a) for (int i = 0; i < nd_te; ++i) {
        double s = 0.0;
        for (int q = 0; q < Q.nq; ++q) s += Q.w[q] * hte[((size_t)q * nd_te + i) * w + a];
        ref[(size_t)a * nd_te + i] = s;
      }
      T.insert(T.end(), ref.begin(), ref.end());
    }
  }
  F.n_stage = (int)T.size();
  // coefficient factors
  F.n_factors = n_factors;
  int prog_at = 0;
  for (int f = 0; f < n_factors; ++f) {
    DevFactor& d = F.factors[f];
    d.kind = 0;
    if (!used[f]) continue;
    const tfelcu_factor& s = factors[f];
    d.kind = s.kind; d.d = s.d;
    if (s.kind == TFELCU_FACTOR_FIELD) {
      TFELCU_REQUIRE(s.fes && s.fes->mesh == mesh, "field factor lives on another mesh");
      TFELCU_REQUIRE(s.d >= 0 && s.d <= dim, "field derivative index out of range");
      d.nd = s.fes->nd; d.dof_map = s.fes->dof_map;
      d.tab_off = (int)T.size(); tabulate(dim, s.fes->fe, Q, T);
      if (is_device_pointer(s.data)) d.data = s.data;
      else {
        out.keep.emplace_back(s.fes->n_dof);
        ctx->upload(out.keep.back().p, s.data, s.fes->n_dof);
        d.data = out.keep.back().p;
      }
    } else if (s.kind == TFELCU_FACTOR_TABLE) {
      const size_t n = mesh->K * (size_t)Q.nq;
      if (is_device_pointer(s.data)) d.data = s.data;
      else {
        out.keep.emplace_back(n);
        ctx->upload(out.keep.back().p, s.data, n);
        d.data = out.keep.back().p;
      }
    } else if (s.kind == TFELCU_FACTOR_PROGRAM) {
      TFELCU_REQUIRE(prog_at + s.n_ops <= kMaxProgOps, "coefficient programs too long");
      d.prog_off = prog_at; d.prog_len = s.n_ops;
      int depth = 0;
      for (int p = 0; p < s.n_ops; ++p) {
        const int op = s.ops[p].op;
        if (op == TFELCU_OP_CONST || op == TFELCU_OP_COORD) ++depth;
        else if (op == TFELCU_OP_ADD || op == TFELCU_OP_SUB || op == TFELCU_OP_MUL || op == TFELCU_OP_DIV ||
                 op == TFELCU_OP_POW || op == TFELCU_OP_LT) { TFELCU_REQUIRE(depth >= 2, "malformed coefficient program"); --depth; }
        else TFELCU_REQUIRE(depth >= 1, "malformed coefficient program");
        TFELCU_REQUIRE(depth <= 8, "coefficient program needs a deeper stack than 8");
        F.prog[prog_at + p] = s.ops[p];
      }
      TFELCU_REQUIRE(depth == 1, "malformed coefficient program");
      prog_at += s.n_ops;
    } else if (s.kind == TFELCU_FACTOR_FUNCTION) {
      // device code of the caller: its module tabulates the function at the quadrature points of every cell, on our
      // stream, into a table the element kernels read (nothing is evaluated or copied on the host)
      TFELCU_REQUIRE(s.fn != nullptr, "function factor without device code");
      TFELCU_REQUIRE(s.closure == nullptr || is_device_pointer(s.closure), "the closure of a function factor must live in device memory");
      const size_t n = mesh->K * (size_t)Q.nq;
      out.keep.emplace_back(Q.x.size());
      double* xhat = out.keep.back().p;
      ctx->upload(xhat, Q.x.data(), Q.x.size());
      out.keep.emplace_back(n);
      tfelcu_points where;
      where.dim = dim; where.nq = Q.nq; where.n_cells = mesh->K;
      where.vertices = mesh->vertices.p; where.cells = mesh->cells.p; where.xhat = xhat;
      const int rc = s.fn(s.closure, &where, out.keep.back().p, (void*)ctx->stream);
      TFELCU_REQUIRE(rc == 0, "the tabulation kernel of a function factor could not be launched");
      d.kind = TFELCU_FACTOR_TABLE; d.data = out.keep.back().p;
    } else {
      fail("unknown coefficient factor kind");
    }
  }
}

// grow-only device scratch for the tables of the form being assembled (stream ordered reuse)
static const double* upload_tables(Ctx* ctx, const std::vector<double>& T, DevBuf<double>& scratch) {
  if (scratch.n < T.size()) { ctx->sync(); scratch.alloc(T.size() * 2); }
  TFELCU_CK(cudaMemcpyAsync(scratch.p, T.data(), T.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  return scratch.p;
}

static DevBuf<double>& table_scratch(Ctx* ctx) { return ctx->table_scratch; }

const double* upload_form_tables(Ctx* ctx, const std::vector<double>& T) { return upload_tables(ctx, T, table_scratch(ctx)); }

template <int DIM, int ND_TE, int ND_TR>
static void launch_rows(Ctx* ctx, const FormBuild& fb, const RowBlock& B, bool staged, size_t stage_doubles, int rows_per_cta) {
  if (B.row_end <= B.row_begin) return;
  const unsigned grid = grid_for(B.row_end - B.row_begin, (unsigned)rows_per_cta);
  const size_t smem = ((size_t)fb.F.n_stage + (staged ? stage_doubles : 0)) * sizeof(double);
#define TFELCU_ROWS(CC, ST)                                                                      \
  do {                                                                                           \
    auto kern = k_assemble_rows<DIM, ND_TE, ND_TR, CC, ST>;                                      \
    ensure_smem(kern, smem);                                                                     \
    kern<<<grid, rows_per_cta, smem, ctx->stream>>>(fb.F, B);                                    \
  } while (0)
  if constexpr (DIM == 2 && ND_TE == 3 && ND_TR == 3) {
    if (fb.const_coef && B.rec) {   // packed row records
      if (staged) {
        auto kern = k_assemble_rows<2, 3, 3, true, true, true>;
        ensure_smem(kern, smem);
        kern<<<grid, rows_per_cta, smem, ctx->stream>>>(fb.F, B);
      } else {
        auto kern = k_assemble_rows<2, 3, 3, true, false, true>;
        ensure_smem(kern, smem);
        kern<<<grid, rows_per_cta, smem, ctx->stream>>>(fb.F, B);
      }
      ctx->count(); TFELCU_LAUNCH_CHECK();
      return;
    }
  }
  if (fb.const_coef) { if (staged) TFELCU_ROWS(true, true); else TFELCU_ROWS(true, false); }
  else { if (staged) TFELCU_ROWS(false, true); else TFELCU_ROWS(false, false); }
#undef TFELCU_ROWS
  ctx->count(); TFELCU_LAUNCH_CHECK();
}

template <int DIM, int ND_TE, int ND_TR>
static void launch_colored(Ctx* ctx, const FormBuild& fb, const ColorBlock& B) {
  const unsigned grid = grid_for(B.n_cells, 128);
  const size_t smem = (size_t)fb.F.n_stage * sizeof(double);
  if (fb.const_coef) {
    auto kern = k_assemble_colored<DIM, ND_TE, ND_TR, true>;
    ensure_smem(kern, smem);
    kern<<<grid, 128, smem, ctx->stream>>>(fb.F, B);
  } else {
    auto kern = k_assemble_colored<DIM, ND_TE, ND_TR, false>;
    ensure_smem(kern, smem);
    kern<<<grid, 128, smem, ctx->stream>>>(fb.F, B);
  }
  ctx->count(); TFELCU_LAUNCH_CHECK();
}

template <int DIM, int ND_TE>
static void launch_vector(Ctx* ctx, const FormBuild& fb, const VecBlock& B) {
  if (B.row_end <= B.row_begin) return;
  const unsigned grid = grid_for(B.row_end - B.row_begin, 128);
  const size_t smem = (size_t)fb.F.n_stage * sizeof(double);
  if (fb.const_coef) {
    auto kern = k_assemble_vector_rows<DIM, ND_TE, true>;
    ensure_smem(kern, smem);
    kern<<<grid, 128, smem, ctx->stream>>>(fb.F, B);
  } else {
    auto kern = k_assemble_vector_rows<DIM, ND_TE, false>;
    ensure_smem(kern, smem);
    kern<<<grid, 128, smem, ctx->stream>>>(fb.F, B);
  }
  ctx->count(); TFELCU_LAUNCH_CHECK();
}

template <int DIM>
static void assemble_bilinear_dim(tfelcu_matrix* m, int quad, const tfelcu_factor* factors, int n_factors,
                                  const tfelcu_term* terms, int n_terms) {
  Ctx* ctx = m->ctx;
  tfelcu_mesh* mesh = m->mesh;
  const bool single_block = (m->test.n == 1 && m->trial.n == 1);
  const bool colored = (m->scatter == TFELCU_SCATTER_COLORED);
  if (colored) matrix_build_colors(m);
  else if (!m->has_slots && !m->slots_failed && !std::getenv("TFELCU_NO_SLOTS")) {
    try { matrix_build_slots(m); } catch (const Refusal&) { m->has_slots = false; m->slots_failed = true; }   // rows > 256 entries: search
  }
  const bool use_rec = !colored && m->has_slots && !std::getenv("TFELCU_NO_REC");
  if (use_rec) matrix_build_recs(m);
  for (int a = 0; a < m->test.n; ++a)
    for (int b = 0; b < m->trial.n; ++b) {
      tfelcu_fes* te = m->test.fes[a]; tfelcu_fes* tr = m->trial.fes[b];
      FormBuild fb;
      build_form(ctx, mesh, quad, factors, n_factors, terms, n_terms, a, te->fe, b, tr->fe, fb);
      if (fb.F.n_terms == 0) continue;   // the block only contributes (already stored) zeros
      fb.F.tables = upload_tables(ctx, fb.tables, table_scratch(ctx));
      const size_t stage_bytes = ((size_t)fb.F.n_stage + m->max_block_nnz) * sizeof(double);
      const bool staged = single_block && m->rows.n == 1 && !colored && te->nd == tr->nd && stage_bytes <= 160 * 1024;
      if (!staged) matrix_ensure_cleared(m);
      if (!colored) {
        fes_transpose(te);
        RowBlock B;
        B.vertices = mesh->vertices.p; B.cells = mesh->cells.p; B.map_tr = tr->dof_map;
        B.off_te = (uint32_t)m->test.offset[a]; B.off_tr = (uint32_t)m->trial.offset[b];
        B.adj_ptr = te->adj_ptr.p; B.adj = te->adj.p; B.dir_row = m->dir_row.p;
        B.rowptr = m->rowptr.p; B.colind = m->colind.p; B.val = m->val.p;
        B.init_mode = m->val_valid ? 0 : 1;
        const int nd_te = te->nd, nd_tr = tr->nd;
        for (int sg = 0; sg < m->rows.n; ++sg) {   // the owned row segments of this test component
          if (m->rows.comp[sg] != a) continue;
          B.row_begin = m->rows.g_begin[sg] - m->test.offset[a]; B.row_end = m->rows.g_end[sg] - m->test.offset[a];
          B.local_off = m->rows.local_off[sg];
          B.slots = m->has_slots ? m->slots[(size_t)sg * m->trial.n + b].p : nullptr;
          B.a_begin = m->a_begin[sg];
          B.rec = (use_rec && m->recs[(size_t)sg * m->trial.n + b].n)
                      ? reinterpret_cast<const uint4*>(m->recs[(size_t)sg * m->trial.n + b].p) : nullptr;
          TFELCU_DISPATCH_ND(DIM, nd_te, NTE,
            TFELCU_DISPATCH_ND(DIM, nd_tr, NTR,
              if (staged) { if constexpr (NTE == NTR) launch_rows<DIM, NTE, NTR>(ctx, fb, B, true, m->max_block_nnz, m->rows_per_cta); }
              else launch_rows<DIM, NTE, NTR>(ctx, fb, B, false, 0, m->rows_per_cta);))
        }
        if (staged) m->val_valid = true;
      } else {
        ColorBlock B;
        B.vertices = mesh->vertices.p; B.cells = mesh->cells.p; B.map_te = te->dof_map; B.map_tr = tr->dof_map;
        B.off_te = (uint32_t)m->test.offset[a]; B.off_tr = (uint32_t)m->trial.offset[b];
        B.dir_row = m->dir_row.p; B.rowptr = m->rowptr.p; B.colind = m->colind.p; B.val = m->val.p;
        const int nd_te = te->nd, nd_tr = tr->nd;
        for (int sg = 0; sg < m->rows.n; ++sg) {
          if (m->rows.comp[sg] != a) continue;
          B.row_begin = m->rows.g_begin[sg] - m->test.offset[a]; B.row_end = m->rows.g_end[sg] - m->test.offset[a];
          B.local_off = m->rows.local_off[sg];
          for (int c = 0; c < m->n_colors; ++c) {
            B.color_cells = m->color_cells.p + m->color_start[c];
            B.n_cells = m->color_start[c + 1] - m->color_start[c];
            if (!B.n_cells) continue;
            TFELCU_DISPATCH_ND(DIM, nd_te, NTE,
              TFELCU_DISPATCH_ND(DIM, nd_tr, NTR, launch_colored<DIM, NTE, NTR>(ctx, fb, B);))
          }
        }
      }
      if (!fb.keep.empty()) ctx->sync();   // device copies of host coefficient data die with fb
    }
  matrix_ensure_cleared(m);
  m->pristine = false;
}

// AUTO resolves to the measured winner. On configs[1] (Poisson P1, n = 4096, B200) the owner-computes rows take
// 0.97 ms (A) + 0.52 ms (b), the patch strategy 1.59-1.76 ms + 0.55 ms: the same ~5e8 warp instructions, but a patch CTA
// walks three dependent gathers per phase with only ~700 rows in flight per SM (57 KB of shared memory per 256 rows) and
// stalls on the long scoreboard (profiles/r01_ncu_patch_assembly.md). So AUTO = ROW_GATHER; TFELCU_SCATTER=patch|rows|colored
// in the environment, TFELCU_PATCH_AUTO_MAX_NN (largest nd_te * nd_tr that AUTO hands to PATCH) or
// tfelcu_matrix_set_scatter pick another one.
int resolve_scatter(int requested, bool single_block, int nd_te, int nd_tr) {
  if (requested != TFELCU_SCATTER_AUTO) return requested;
  if (const char* e = std::getenv("TFELCU_SCATTER")) {
    if (!std::strcmp(e, "tile")) return TFELCU_SCATTER_TILE;
    if (!std::strcmp(e, "rows")) return TFELCU_SCATTER_ROW_GATHER;
    if (!std::strcmp(e, "patch")) return TFELCU_SCATTER_PATCH;
    if (!std::strcmp(e, "colored")) return TFELCU_SCATTER_COLORED;
  }
  int max_nn = 0;
  if (const char* e = std::getenv("TFELCU_PATCH_AUTO_MAX_NN")) max_nn = std::atoi(e);
  return (single_block && nd_te * nd_tr <= max_nn) ? TFELCU_SCATTER_PATCH : TFELCU_SCATTER_ROW_GATHER;
}

void form_build_free(FormBuild* p) { delete p; }

// launch a += that was accepted but deferred (tile path): on its own, no linear form rode along
void matrix_flush_pending(tfelcu_matrix* m) {
  if (!m->pending) return;
  std::unique_ptr<FormBuild> fb(m->pending);
  m->pending = nullptr;
  if (m->ctx->pending_matrix == m) m->ctx->pending_matrix = nullptr;
  const bool ok = m->pending_mode == 1 ? assemble_tiles(m, *fb, nullptr, nullptr) : assemble_p1_rows(m, *fb, nullptr, nullptr);
  TFELCU_REQUIRE(ok, "assembly: the plan of a deferred += is gone");
  m->pristine = false;
}

// key of a constant-coefficient integrand: quadrature formula + the terms, byte for byte
static std::vector<unsigned char> form_key(int quad, const tfelcu_term* terms, int n_terms) {
  std::vector<unsigned char> k(sizeof(int) + (size_t)n_terms * sizeof(tfelcu_term));
  std::memcpy(k.data(), &quad, sizeof(int));
  if (n_terms) std::memcpy(k.data() + sizeof(int), terms, (size_t)n_terms * sizeof(tfelcu_term));
  return k;
}

// Constant-coefficient forms on P1 triangles: the closed-form row kernel of p1rows.cu (AUTO) or the tiles of tile.cu
// (TFELCU_SCATTER_TILE). The launch is deferred: a constant-coefficient linear form on the same space that follows is
// assembled in the same pass.
static bool try_p1_fast(tfelcu_matrix* m, int quad, const tfelcu_factor* factors, int n_factors, const tfelcu_term* terms,
                        int n_terms, bool tile, bool explicit_request) {
  if (tile ? !tile_applicable(m) : !p1rows_applicable(m)) {
    if (explicit_request) refuse("tile assembly: needs one scalar P1 space on triangles with every vertex used");
    return false;
  }
  for (int t = 0; t < n_terms; ++t) if (terms[t].n_factors != 0) { if (explicit_request) refuse("tile assembly: constant coefficients only"); return false; }
  std::unique_ptr<FormBuild> fb(new FormBuild());
  tfelcu_fes* te = m->test.fes[0];
  const std::vector<unsigned char> key = form_key(quad, terms, n_terms);
  if (m->memo_form && m->memo_key == key) {
    fb->F = m->memo_form->F; fb->tables = m->memo_form->tables; fb->const_coef = m->memo_form->const_coef;
  } else {
    build_form(m->ctx, m->mesh, quad, factors, n_factors, terms, n_terms, 0, te->fe, 0, te->fe, *fb);
    if (m->memo_form) form_build_free(m->memo_form);
    m->memo_form = new FormBuild();
    m->memo_form->F = fb->F; m->memo_form->tables = fb->tables; m->memo_form->const_coef = fb->const_coef;
    m->memo_key = key;
  }
  if (fb->F.n_terms == 0) return true;   // only (already stored) zeros
  if (!fb->const_coef) return false;
  if (tile) {
    if (!matrix_tile_plan_ready(m)) { if (explicit_request) refuse("tile assembly: a tile of this mesh exceeds the capacities of the plan"); return false; }
  } else if (m->rows_per_cta != 128 || m->max_block_nnz * sizeof(double) > 200 * 1024 || !matrix_ell_ready(m)) return false;
  // the tile kernel takes the form by value (tile.cu: TileForm): nothing of it lives in the shared table scratch
  Ctx* ctx = m->ctx;
  if (ctx->pending_matrix && ctx->pending_matrix != m) matrix_flush_pending(ctx->pending_matrix);
  m->pending = fb.release();
  m->pending_mode = tile ? 1 : 0;
  ctx->pending_matrix = m;
  if (std::getenv("TFELCU_NO_DEFER")) matrix_flush_pending(m);
  return true;
}

void assemble_bilinear(tfelcu_matrix* m, int quad, const tfelcu_factor* factors, int n_factors,
                       const tfelcu_term* terms, int n_terms) {
  for (int t = 0; t < n_terms; ++t) {
    TFELCU_REQUIRE(terms[t].test_comp >= 0 && terms[t].test_comp < m->test.n, "test component out of range");
    TFELCU_REQUIRE(terms[t].trial_comp >= 0 && terms[t].trial_comp < m->trial.n, "trial component out of range");
  }
  if (m->pending) matrix_flush_pending(m);   // a second += on the same form: the first one goes first
  const bool single_block = (m->test.n == 1 && m->trial.n == 1);
  const int requested = m->scatter;
  {
    const int want = resolve_scatter(requested, single_block, m->test.fes[0]->nd, m->trial.fes[0]->nd);
    // AUTO: the closed-form P1 rows where they apply (TFELCU_NO_P1ROWS=1: the generic kernels only)
    const bool auto_p1 = requested == TFELCU_SCATTER_AUTO && want == TFELCU_SCATTER_ROW_GATHER && !std::getenv("TFELCU_NO_P1ROWS");
    if (want == TFELCU_SCATTER_TILE || auto_p1)
      if (try_p1_fast(m, quad, factors, n_factors, terms, n_terms, want == TFELCU_SCATTER_TILE,
                      want == TFELCU_SCATTER_TILE && requested == TFELCU_SCATTER_TILE)) return;
  }
  int mode = (requested == TFELCU_SCATTER_AUTO && m->patch_refused)
                 ? TFELCU_SCATTER_ROW_GATHER
                 : resolve_scatter(requested, single_block, m->test.fes[0]->nd, m->trial.fes[0]->nd);
  if (mode == TFELCU_SCATTER_TILE) mode = TFELCU_SCATTER_ROW_GATHER;   // did not apply (AUTO / environment)
  if (mode == TFELCU_SCATTER_PATCH) {
    if (requested == TFELCU_SCATTER_PATCH) { assemble_bilinear_patch(m, quad, factors, n_factors, terms, n_terms); return; }
    try { assemble_bilinear_patch(m, quad, factors, n_factors, terms, n_terms); return; }
    catch (const Refusal&) { m->patch_refused = true; mode = TFELCU_SCATTER_ROW_GATHER; }   // refused before touching a value
  }
  const int saved = m->scatter;
  m->scatter = mode;
  try {
    if (m->mesh->dim == 2) assemble_bilinear_dim<2>(m, quad, factors, n_factors, terms, n_terms);
    else assemble_bilinear_dim<3>(m, quad, factors, n_factors, terms, n_terms);
  } catch (...) { m->scatter = saved; throw; }
  m->scatter = saved;
}

template <int DIM>
static void assemble_linear_dim(tfelcu_vector* v, int quad, const tfelcu_factor* factors, int n_factors,
                                const tfelcu_term* terms, int n_terms) {
  Ctx* ctx = v->ctx;
  for (int a = 0; a < v->space.n; ++a) {
    tfelcu_fes* te = v->space.fes[a];
    FormBuild fb;
    build_form(ctx, te->mesh, quad, factors, n_factors, terms, n_terms, a, te->fe, -1, 0, fb);
    if (fb.F.n_terms == 0) continue;
    fb.F.tables = upload_tables(ctx, fb.tables, table_scratch(ctx));
    fes_transpose(te);
    VecBlock B;
    B.vertices = te->mesh->vertices.p; B.cells = te->mesh->cells.p;
    B.adj_ptr = te->adj_ptr.p; B.adj = te->adj.p; B.b = v->data.p;
    const int nd_te = te->nd;
    for (int sg = 0; sg < v->rows.n; ++sg) {
      if (v->rows.comp[sg] != a) continue;
      B.row_begin = v->rows.g_begin[sg] - v->space.offset[a]; B.row_end = v->rows.g_end[sg] - v->space.offset[a];
      B.local_off = v->rows.local_off[sg];
      TFELCU_DISPATCH_ND(DIM, nd_te, NTE, launch_vector<DIM, NTE>(ctx, fb, B);)
    }
    if (!fb.keep.empty()) ctx->sync();
  }
}

// b += (constant-coefficient linear form without derivatives) on the space of a form whose += is still pending: both in
// one pass over the mesh
static bool try_ride_along(tfelcu_vector* v, int quad, const tfelcu_factor* factors, int n_factors,
                           const tfelcu_term* terms, int n_terms) {
  Ctx* ctx = v->ctx;
  tfelcu_matrix* m = ctx->pending_matrix;
  if (!m || !m->pending || v->space.n != 1 || v->space.fes[0] != m->test.fes[0] || v->rows.n != 1) return false;
  if (v->rows.g_begin[0] != m->rows.g_begin[0] || v->rows.g_end[0] != m->rows.g_end[0]) return false;
  if (v->scatter != TFELCU_SCATTER_AUTO) return false;
  for (int t = 0; t < n_terms; ++t) if (terms[t].n_factors != 0 || terms[t].test_d != 0) return false;
  FormBuild fbv;
  tfelcu_fes* te = v->space.fes[0];
  const std::vector<unsigned char> key = form_key(quad, terms, n_terms);
  if (v->memo_form && v->memo_key == key) {
    fbv.F = v->memo_form->F; fbv.tables = v->memo_form->tables; fbv.const_coef = v->memo_form->const_coef;
  } else {
    build_form(ctx, te->mesh, quad, factors, n_factors, terms, n_terms, 0, te->fe, -1, 0, fbv);
    if (v->memo_form) form_build_free(v->memo_form);
    v->memo_form = new FormBuild();
    v->memo_form->F = fbv.F; v->memo_form->tables = fbv.tables; v->memo_form->const_coef = fbv.const_coef;
    v->memo_key = key;
  }
  if (fbv.F.n_terms == 0) return true;
  if (!fbv.const_coef) return false;
  std::unique_ptr<FormBuild> fb(m->pending);
  m->pending = nullptr; ctx->pending_matrix = nullptr;
  const bool ok = m->pending_mode == 1 ? assemble_tiles(m, *fb, v, &fbv) : assemble_p1_rows(m, *fb, v, &fbv);
  TFELCU_REQUIRE(ok, "assembly: the plan of a deferred += is gone");
  m->pristine = false;
  return true;
}

void assemble_linear(tfelcu_vector* v, int quad, const tfelcu_factor* factors, int n_factors,
                     const tfelcu_term* terms, int n_terms) {
  for (int t = 0; t < n_terms; ++t)
    TFELCU_REQUIRE(terms[t].test_comp >= 0 && terms[t].test_comp < v->space.n, "test component out of range");
  if (try_ride_along(v, quad, factors, n_factors, terms, n_terms)) return;
  v->ensure_cleared();
  const int requested = v->scatter;
  const int nd = v->space.fes[0]->nd;
  int mode = (requested == TFELCU_SCATTER_AUTO && v->patch_refused) ? TFELCU_SCATTER_ROW_GATHER
                                                                    : resolve_scatter(requested, v->space.n == 1, nd, nd);
  if (mode == TFELCU_SCATTER_PATCH) {
    if (requested == TFELCU_SCATTER_PATCH) { assemble_linear_patch(v, quad, factors, n_factors, terms, n_terms); return; }
    // AUTO: a refusal (shared memory) before the first launch falls back to rows; later ones would double count
    try { assemble_linear_patch(v, quad, factors, n_factors, terms, n_terms); return; }
    catch (const Refusal&) { if (v->space.n != 1 || v->rows.n != 1) throw; v->patch_refused = true; }
  }
  if (v->space.fes[0]->mesh->dim == 2) assemble_linear_dim<2>(v, quad, factors, n_factors, terms, n_terms);
  else assemble_linear_dim<3>(v, quad, factors, n_factors, terms, n_terms);
}

// quadrature points in space coordinates [K][nq][dim] (cell.hpp:456-468: x = v0 + sum_m xhat_m (v_{m+1} - v0))
template <int DIM>
__global__ void k_quadrature_points(const double* __restrict__ vertices, const uint32_t* __restrict__ cells, size_t K,
                                    const double* __restrict__ xhat, int nq, double* __restrict__ xq) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= K * (size_t)nq) return;
  const size_t k = t / nq; const int q = (int)(t % nq);
  const uint32_t* cell = cells + k * (DIM + 1);
  for (int c = 0; c < DIM; ++c) {
    const double v0 = vertices[(size_t)cell[0] * DIM + c];
    double acc = v0;
    for (int m = 0; m < DIM; ++m)
      acc = __dadd_rn(acc, __dmul_rn(xhat[q * DIM + m], __dsub_rn(vertices[(size_t)cell[m + 1] * DIM + c], v0)));
    xq[t * DIM + c] = acc;
  }
}

int quadrature_size(int quad, int* dim) {
  const Quadrature Q = make_quadrature(quad);
  if (dim) *dim = Q.dim;
  return Q.nq;
}

void quadrature_points(Ctx* ctx, tfelcu_mesh* mesh, int quad, double* d_xq) {
  const Quadrature Q = make_quadrature(quad);
  TFELCU_REQUIRE(Q.dim == mesh->dim, "quadrature formula and mesh have different dimensions");
  DevBuf<double> xhat(Q.x.size());
  ctx->upload(xhat.p, Q.x.data(), Q.x.size());
  const size_t n = mesh->K * (size_t)Q.nq;
  if (mesh->dim == 2) k_quadrature_points<2><<<grid_for(n, 256), 256, 0, ctx->stream>>>(mesh->vertices.p, mesh->cells.p, mesh->K, xhat.p, Q.nq, d_xq);
  else k_quadrature_points<3><<<grid_for(n, 256), 256, 0, ctx->stream>>>(mesh->vertices.p, mesh->cells.p, mesh->K, xhat.p, Q.nq, d_xq);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  ctx->sync();
}

double integrate_scalar(Ctx* ctx, tfelcu_mesh* mesh, int quad, const tfelcu_factor* factors, int n_factors,
                        const tfelcu_term* terms, int n_terms) {
  FormBuild fb;
  build_form(ctx, mesh, quad, factors, n_factors, terms, n_terms, -1, 0, -1, 0, fb);
  fb.F.tables = upload_tables(ctx, fb.tables, table_scratch(ctx));
  const unsigned grid = grid_for(mesh->K, 256);
  DevBuf<double> partial(grid), out(1);
  const size_t smem = (size_t)fb.F.n_stage * sizeof(double);
  if (mesh->dim == 2) k_integrate<2><<<grid, 256, smem, ctx->stream>>>(fb.F, mesh->vertices.p, mesh->cells.p, mesh->K, partial.p);
  else k_integrate<3><<<grid, 256, smem, ctx->stream>>>(fb.F, mesh->vertices.p, mesh->cells.p, mesh->K, partial.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  k_sum_partials<<<1, 256, 0, ctx->stream>>>(partial.p, grid, out.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  double h = 0.0; ctx->download(&h, out.p, 1);
  return h;
}

}  // namespace tfelcu
