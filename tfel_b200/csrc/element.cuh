// Device pieces shared by the assembly kernels (assemble.cu: owner-computes rows / coloured scatter; patch.cu: patches
// of rows): the integrand in sum-of-products form and the element matrix / vector rows built from it.
// Reference: fe_value_manager::prepare (fe_value_manager.hpp:103-127), the expression leaves (expression.hpp:31-139)
// and the accumulation loops of bilinear_form.hpp:64-109 / linear_form.hpp:55-63.
#pragma once

#include "structs.cuh"
#include "geometry.cuh"

namespace tfelcu {

constexpr int kMaxTerms = 24;
constexpr int kMaxFactors = 12;
constexpr int kMaxProgOps = 48;

struct DevFactor {
  int kind, d;
  int nd;                    // FIELD: dofs per cell of the field's element
  int tab_off;               // FIELD: offset (doubles) of hat[nq][nd][DIM+1] in the tables buffer
  const uint32_t* dof_map;   // FIELD
  const double* data;        // FIELD: coefficients; TABLE: [K][nq]
  int prog_off, prog_len;    // PROGRAM
};

struct DevTerm {
  int test_d, trial_d, n_factors;
  int factor[5];
  double c;
};

struct FormDev {
  int nq, n_terms, n_factors, pad;
  int off_w, off_xhat, off_hat_te, off_hat_tr, off_ref, n_stage;
  // constant-coefficient forms: sum_t c_t D^{a_t} psi D^{b_t} phi collapses to the W x W tensor
  //   G = [ c00 , c0r^T Jmt ; Jmt^T cr0 , Jmt^T C Jmt ]   contracted with the pre-integrated reference tensor
  int has00, has0r, hasr0, hasrr;
  double c00, c0r[kMaxDim], cr0[kMaxDim], C[kMaxDim][kMaxDim];
  const double* tables;
  DevTerm terms[kMaxTerms];
  DevFactor factors[kMaxFactors];
  tfelcu_op prog[kMaxProgOps];
};

// ------------------------------------------------------------------------------------------
// device pieces
// ------------------------------------------------------------------------------------------

template <int DIM>
__device__ __forceinline__ double run_program(const FormDev& F, int off, int len, const double* x) {
  double st[8];
  int sp = 0;
  for (int p = 0; p < len; ++p) {
    const tfelcu_op o = F.prog[off + p];
    switch (o.op) {
      case TFELCU_OP_CONST: st[sp++] = o.imm; break;
      case TFELCU_OP_COORD: st[sp++] = x[(int)o.imm < DIM ? (int)o.imm : 0]; break;
      case TFELCU_OP_ADD: --sp; st[sp - 1] = st[sp - 1] + st[sp]; break;
      case TFELCU_OP_SUB: --sp; st[sp - 1] = st[sp - 1] - st[sp]; break;
      case TFELCU_OP_MUL: --sp; st[sp - 1] = st[sp - 1] * st[sp]; break;
      case TFELCU_OP_DIV: --sp; st[sp - 1] = st[sp - 1] / st[sp]; break;
      case TFELCU_OP_NEG: st[sp - 1] = -st[sp - 1]; break;
      case TFELCU_OP_SQRT: st[sp - 1] = sqrt(st[sp - 1]); break;
      case TFELCU_OP_EXP: st[sp - 1] = exp(st[sp - 1]); break;
      case TFELCU_OP_SIN: st[sp - 1] = sin(st[sp - 1]); break;
      case TFELCU_OP_COS: st[sp - 1] = cos(st[sp - 1]); break;
      case TFELCU_OP_POW: --sp; st[sp - 1] = pow(st[sp - 1], st[sp]); break;
      case TFELCU_OP_LT: --sp; st[sp - 1] = st[sp - 1] < st[sp] ? 1.0 : 0.0; break;
      case TFELCU_OP_ABS: st[sp - 1] = fabs(st[sp - 1]); break;
      default: break;
    }
  }
  return st[0];
}

// values of all coefficient factors at quadrature point q of cell k
template <int DIM>
__device__ __forceinline__ void eval_factors(const FormDev& F, const double* sm, size_t k, int q,
                                             const CellGeom<DIM>& g, double* fv) {
  constexpr int W = DIM + 1;
  for (int f = 0; f < F.n_factors; ++f) {
    const DevFactor& fa = F.factors[f];
    double v = 1.0;
    if (fa.kind == TFELCU_FACTOR_TABLE) {
      v = __ldg(fa.data + k * F.nq + q);
    } else if (fa.kind == TFELCU_FACTOR_FIELD) {
      // finite_element_function::prepare (expression.hpp:114-128): sum_i coef[dof(k, i)] * phi_i[d](x_q)
      const double* hat = F.tables + fa.tab_off + (size_t)q * fa.nd * W;
      v = 0.0;
      for (int i = 0; i < fa.nd; ++i) {
        const double c = __ldg(fa.data + fa.dof_map[k * fa.nd + i]);
        double b;
        if (fa.d == 0) b = __ldg(hat + i * W);
        else {
          b = 0.0;
#pragma unroll
          for (int t = 0; t < DIM; ++t) b += g.jmt[(fa.d - 1) < DIM ? (fa.d - 1) : 0][t] * __ldg(hat + i * W + 1 + t);
        }
        v += c * b;
      }
    } else if (fa.kind == TFELCU_FACTOR_PROGRAM) {
      // x_q = v0 + J xhat_q (cell.hpp:456-468)
      const double* xh = sm + F.off_xhat + q * DIM;
      double x[DIM];
#pragma unroll
      for (int r = 0; r < DIM; ++r) {
        double acc = g.v0[r];
#pragma unroll
        for (int c = 0; c < DIM; ++c) acc += xh[c] * g.J[r][c];
        x[r] = acc;
      }
      v = run_program<DIM>(F, fa.prog_off, fa.prog_len, x);
    }
    fv[f] = v;
  }
}

__device__ __forceinline__ double term_coefficient(const DevTerm& t, const double* fv) {
  double c = t.c;
  for (int f = 0; f < t.n_factors; ++f) c *= fv[t.factor[f]];
  return c;
}

// row i of the element matrix of cell k (without the |K| factor): acc[j], j = 0..ND_TR-1
template <int DIM, int ND_TE, int ND_TR, bool CONST_COEF>
__device__ __forceinline__ void element_row(const FormDev& F, const double* sm, size_t k, int i,
                                            const CellGeom<DIM>& g, double (&acc)[ND_TR]) {
  constexpr int W = DIM + 1;
#pragma unroll
  for (int j = 0; j < ND_TR; ++j) acc[j] = 0.0;
  if constexpr (CONST_COEF) {
    // G : Mhat with Mhat[a][b][i][j] = sum_q w_q hat_te[q][i][a] hat_tr[q][j][b] staged in shared memory
    const double* ref = sm + F.off_ref;
    if (F.has00) {
      const double* mrow = ref + (0 * ND_TE + i) * ND_TR;
#pragma unroll
      for (int j = 0; j < ND_TR; ++j) acc[j] += F.c00 * mrow[j];
    }
    if (F.has0r) {
#pragma unroll
      for (int u = 0; u < DIM; ++u) {
        double gv = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) gv += F.c0r[d] * g.jmt[d][u];
        const double* mrow = ref + ((0 * W + 1 + u) * ND_TE + i) * ND_TR;
#pragma unroll
        for (int j = 0; j < ND_TR; ++j) acc[j] += gv * mrow[j];
      }
    }
    if (F.hasr0) {
#pragma unroll
      for (int s2 = 0; s2 < DIM; ++s2) {
        double gv = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) gv += F.cr0[d] * g.jmt[d][s2];
        const double* mrow = ref + (((1 + s2) * W + 0) * ND_TE + i) * ND_TR;
#pragma unroll
        for (int j = 0; j < ND_TR; ++j) acc[j] += gv * mrow[j];
      }
    }
    if (F.hasrr) {
      // Jmt^T C Jmt
      double CJ[DIM][DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d)
#pragma unroll
        for (int u = 0; u < DIM; ++u) {
          double v = 0.0;
#pragma unroll
          for (int e = 0; e < DIM; ++e) v += F.C[d][e] * g.jmt[e][u];
          CJ[d][u] = v;
        }
#pragma unroll
      for (int s2 = 0; s2 < DIM; ++s2)
#pragma unroll
        for (int u = 0; u < DIM; ++u) {
          double gv = 0.0;
#pragma unroll
          for (int d = 0; d < DIM; ++d) gv += g.jmt[d][s2] * CJ[d][u];
          const double* mrow = ref + (((1 + s2) * W + 1 + u) * ND_TE + i) * ND_TR;
#pragma unroll
          for (int j = 0; j < ND_TR; ++j) acc[j] += gv * mrow[j];
        }
    }
  } else {
    const double* wq = sm + F.off_w;
    const double* hat_te = sm + F.off_hat_te;
    const double* hat_tr = sm + F.off_hat_tr;
    for (int q = 0; q < F.nq; ++q) {
      // fe_value_manager::prepare (fe_value_manager.hpp:103-127) for the one test function of this row
      const double* ht = hat_te + (q * ND_TE + i) * W;
      double psi[W];
      psi[0] = ht[0];
#pragma unroll
      for (int s = 0; s < DIM; ++s) {
        double v = 0.0;
#pragma unroll
        for (int t = 0; t < DIM; ++t) v += g.jmt[s][t] * ht[1 + t];
        psi[1 + s] = v;
      }
      double fv[kMaxFactors];
      eval_factors<DIM>(F, sm, k, q, g, fv);
      double Wv[W];
#pragma unroll
      for (int b = 0; b < W; ++b) Wv[b] = 0.0;
      for (int t = 0; t < F.n_terms; ++t) {
        const DevTerm& tm = F.terms[t];
        double p = psi[0];
#pragma unroll
        for (int a = 1; a < W; ++a) p = (tm.test_d == a) ? psi[a] : p;
        const double tw = wq[q] * term_coefficient(tm, fv) * p;
#pragma unroll
        for (int b = 0; b < W; ++b) Wv[b] += (tm.trial_d == b) ? tw : 0.0;
      }
      // D^b phi_j = sum_s jmt[b-1][s] hat_j[1+s]  =>  weights on the reference gradient
      double Wh[W];
      Wh[0] = Wv[0];
#pragma unroll
      for (int s = 0; s < DIM; ++s) {
        double v = 0.0;
#pragma unroll
        for (int b = 0; b < DIM; ++b) v += g.jmt[b][s] * Wv[1 + b];
        Wh[1 + s] = v;
      }
#pragma unroll
      for (int j = 0; j < ND_TR; ++j) {
        const double* hr = hat_tr + (q * ND_TR + j) * W;
        double v = acc[j];
#pragma unroll
        for (int a = 0; a < W; ++a) v += Wh[a] * hr[a];
        acc[j] = v;
      }
    }
  }
}

// whole element matrix of a constant-coefficient form (without |K|): the same sums as element_row, in the same order per
// entry, with the cell's W x W tensor G = [c00, c0r^T Jmt; Jmt^T cr0, Jmt^T C Jmt] formed once instead of once per row
template <int DIM, int ND_TE, int ND_TR>
__device__ __forceinline__ void element_matrix_const(const FormDev& F, const double* sm, const CellGeom<DIM>& g,
                                                     double (&acc)[ND_TE * ND_TR]) {
  constexpr int W = DIM + 1, NN = ND_TE * ND_TR;
  const double* ref = sm + F.off_ref;
#pragma unroll
  for (int e = 0; e < NN; ++e) acc[e] = 0.0;
  if (F.has00) {
    const double* mm = ref;
#pragma unroll
    for (int e = 0; e < NN; ++e) acc[e] += F.c00 * mm[e];
  }
  if (F.has0r) {
#pragma unroll
    for (int u = 0; u < DIM; ++u) {
      double gv = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) gv += F.c0r[d] * g.jmt[d][u];
      const double* mm = ref + (size_t)(0 * W + 1 + u) * NN;
#pragma unroll
      for (int e = 0; e < NN; ++e) acc[e] += gv * mm[e];
    }
  }
  if (F.hasr0) {
#pragma unroll
    for (int s2 = 0; s2 < DIM; ++s2) {
      double gv = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) gv += F.cr0[d] * g.jmt[d][s2];
      const double* mm = ref + (size_t)((1 + s2) * W + 0) * NN;
#pragma unroll
      for (int e = 0; e < NN; ++e) acc[e] += gv * mm[e];
    }
  }
  if (F.hasrr) {
    double CJ[DIM][DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
#pragma unroll
      for (int u = 0; u < DIM; ++u) {
        double v = 0.0;
#pragma unroll
        for (int e = 0; e < DIM; ++e) v += F.C[d][e] * g.jmt[e][u];
        CJ[d][u] = v;
      }
#pragma unroll
    for (int s2 = 0; s2 < DIM; ++s2)
#pragma unroll
      for (int u = 0; u < DIM; ++u) {
        double gv = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) gv += g.jmt[d][s2] * CJ[d][u];
        const double* mm = ref + (size_t)((1 + s2) * W + 1 + u) * NN;
#pragma unroll
        for (int e = 0; e < NN; ++e) acc[e] += gv * mm[e];
      }
  }
}

// constant-coefficient linear forms: sum_t c_t D^{d_t} psi_i = sum_a Ts[a] hat_i[a] with Ts formed once per cell
template <int DIM>
__device__ __forceinline__ void const_test_weights(const FormDev& F, const CellGeom<DIM>& g, double (&Ts)[DIM + 1]) {
#pragma unroll
  for (int a = 0; a <= DIM; ++a) Ts[a] = 0.0;
  for (int t = 0; t < F.n_terms; ++t) {
    const int d = F.terms[t].test_d;
    const double c = F.terms[t].c;
    Ts[0] += d == 0 ? c : 0.0;
#pragma unroll
    for (int s = 0; s < DIM; ++s) {
      double v = 0.0;
#pragma unroll
      for (int dd = 0; dd < DIM; ++dd) v = (d == dd + 1) ? g.jmt[dd][s] : v;
      Ts[1 + s] += c * v;
    }
  }
}

// entry i of the element vector of cell k (without |K|)
template <int DIM, int ND_TE, bool CONST_COEF>
__device__ __forceinline__ double element_entry(const FormDev& F, const double* sm, size_t k, int i,
                                                const CellGeom<DIM>& g) {
  constexpr int W = DIM + 1;
  double out = 0.0;
  if constexpr (CONST_COEF) {
    const double* ref = sm + F.off_ref;   // mhat[a][i] = sum_q w_q hat_te[q][i][a]
    double Ts[W];
    const_test_weights<DIM>(F, g, Ts);
#pragma unroll
    for (int a = 0; a < W; ++a) out += Ts[a] * ref[a * ND_TE + i];
  } else {
    const double* wq = sm + F.off_w;
    const double* hat_te = sm + F.off_hat_te;
    for (int q = 0; q < F.nq; ++q) {
      const double* ht = hat_te + (q * ND_TE + i) * W;
      double psi[W];
      psi[0] = ht[0];
#pragma unroll
      for (int s = 0; s < DIM; ++s) {
        double v = 0.0;
#pragma unroll
        for (int t = 0; t < DIM; ++t) v += g.jmt[s][t] * ht[1 + t];
        psi[1 + s] = v;
      }
      double fv[kMaxFactors];
      eval_factors<DIM>(F, sm, k, q, g, fv);
      double f = 0.0;
      for (int t = 0; t < F.n_terms; ++t) {
        const DevTerm& tm = F.terms[t];
        double p = psi[0];
#pragma unroll
        for (int a = 1; a < W; ++a) p = (tm.test_d == a) ? psi[a] : p;
        f += term_coefficient(tm, fv) * p;
      }
      out += wq[q] * f;
    }
  }
  return out;
}

__device__ __forceinline__ int find_in_row(const int32_t* __restrict__ colind, int rs, int re, int32_t col) {
  int lo = rs, hi = re;
  while (lo < hi) {
    const int mid = lo + ((hi - lo) >> 1);   // lo + hi overflows int32 beyond 2^30 stored entries
    if (__ldg(colind + mid) < col) lo = mid + 1; else hi = mid;
  }
  return lo;
}


// slots of one entry of the transposed dof map: nd_tr bytes padded to whole 32-bit words
__host__ __device__ constexpr int slot_words(int nd_tr) { return (nd_tr + 3) / 4; }

// host side: the C-ABI description of an integrand lowered to FormDev (build_form, assemble.cu)
struct FormBuild {
  FormDev F;
  std::vector<double> tables;
  std::vector<DevBuf<double>> keep;   // device copies of host-resident factor data
  bool const_coef = true;
};

template <typename Kern>
inline void ensure_smem(Kern kern, size_t bytes) {
  if (bytes > 48 * 1024) TFELCU_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

// dispatch on (dim, dofs per cell of the test element, of the trial element)
#define TFELCU_DISPATCH_ND(DIMV, nd, NDV, ...)                                                   \
  if constexpr (DIMV == 2) {                                                                             \
    if (nd == 3) { constexpr int NDV = 3; __VA_ARGS__ } else if (nd == 4) { constexpr int NDV = 4; __VA_ARGS__ }   \
    else if (nd == 6) { constexpr int NDV = 6; __VA_ARGS__ } else fail("unsupported element");   \
  } else {                                                                                       \
    if (nd == 4) { constexpr int NDV = 4; __VA_ARGS__ } else if (nd == 5) { constexpr int NDV = 5; __VA_ARGS__ }   \
    else if (nd == 10) { constexpr int NDV = 10; __VA_ARGS__ } else fail("unsupported element"); \
  }

void build_form(Ctx* ctx, tfelcu_mesh* mesh, int quad, const tfelcu_factor* factors, int n_factors,
                const tfelcu_term* terms, int n_terms, int comp_te, int fe_te, int comp_tr, int fe_tr, FormBuild& out);
const double* upload_form_tables(Ctx* ctx, const std::vector<double>& T);

}  // namespace tfelcu
