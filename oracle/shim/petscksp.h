// TEST INFRASTRUCTURE (oracle): recording stand-in for PETSc 3.7 <petscksp.h>.
// PETSc is a third-party dependency of the reference (README.md:11) that is absent
// from /root/reference and from this image. This header provides exactly the symbols
// src/core/solver.hpp:95-210 and src/core/solver.cpp:57-177 use, with these semantics:
//   * MatSetValues (called once per non-empty row with that row's sorted columns,
//     solver.cpp:108-115) records the CSR the reference hands to PETSc;
//   * VecSetValue records the right-hand side (solver.cpp:160-163);
//   * KSPSolve calls oracle_shim::capture().solve_hook (set by the driver to the CPU
//     Krylov restatement in oracle/tfel_oracle.c) or returns x = 0 when unset.
// Only oracle/_ref is built against it. Never included by the product.
#ifndef ORACLE_SHIM_PETSCKSP_H
#define ORACLE_SHIM_PETSCKSP_H

#include <vector>
#include <string>
#include <cstring>
#include <cctype>
#include <cmath>
#include <functional>

typedef int PetscErrorCode;
typedef int PetscInt;
typedef double PetscReal;
typedef double PetscScalar;
typedef enum { PETSC_FALSE, PETSC_TRUE } PetscBool;
typedef int MPI_Comm;
#define PETSC_COMM_WORLD 0
#define MATSEQAIJ "seqaij"
#define VECSEQ "seq"
#define KSPGMRES "gmres"
#define PCILU "ilu"
typedef enum { NOT_SET_VALUES, INSERT_VALUES, ADD_VALUES } InsertMode;
typedef enum { MAT_FLUSH_ASSEMBLY = 1, MAT_FINAL_ASSEMBLY = 0 } MatAssemblyType;
typedef enum { NORM_1 = 0, NORM_2 = 1, NORM_FROBENIUS = 2, NORM_INFINITY = 3 } NormType;
typedef enum { KSP_NORM_DEFAULT = -1, KSP_NORM_NONE = 0, KSP_NORM_PRECONDITIONED = 1,
               KSP_NORM_UNPRECONDITIONED = 2, KSP_NORM_NATURAL = 3 } KSPNormType;
static const char* const KSPNormTypes_storage[] = {"NONE", "PRECONDITIONED", "UNPRECONDITIONED", "NATURAL", 0};
static const char* const* const KSPNormTypes = KSPNormTypes_storage;

#define CHKERRV(ierr) do { (void)(ierr); } while (0)
#define CHKERRCONTINUE(ierr) do { (void)(ierr); } while (0)

struct _p_PetscObject { const char* prefix; };
typedef _p_PetscObject* PetscObject;

struct _p_Mat {
  PetscInt n_row, n_col;
  std::vector<int> rowptr, colind;
  std::vector<double> val;
  std::vector<int> row_nnz;   // preallocation as requested (solver.cpp:102)
  int last_row;
};
typedef _p_Mat* Mat;

struct _p_Vec { std::vector<double> v; };
typedef _p_Vec* Vec;

struct _p_PC { std::string type; PetscInt levels; };
typedef _p_PC* PC;

struct _p_KSP;
typedef _p_KSP* KSP;
typedef PetscErrorCode (*oracle_shim_monitor_t)(KSP, PetscInt, PetscReal, void*);

struct _p_KSP {
  _p_PetscObject hdr;      // must be first: the reference casts KSP to PetscObject (solver.hpp:192)
  Vec vec_rhs;             // solver.hpp:198
  KSPNormType normtype;    // solver.hpp:200
  std::string type;
  PetscBool guess_nonzero;
  PetscReal rtol, atol, dtol;
  PetscInt maxits, restart;
  bool mgs;
  oracle_shim_monitor_t monitor;
  _p_PC pc;
  Mat A;
  Vec last_x;
};

namespace oracle_shim {
  struct capture_t {
    // last operator / rhs / solver parameters seen at KSPSolve
    Mat A; Vec b; KSP ksp;
    // solve hook: (ksp, A, b, x) -> iterations (x presized & zeroed)
    std::function<int(KSP, Mat, Vec, Vec)> solve_hook;
    int last_its;
    capture_t() : A(0), b(0), ksp(0), last_its(0) {}
  };
  inline capture_t& capture() { static capture_t c; return c; }
}

inline PetscErrorCode PetscInitialize(int*, char***, const char*, const char*) { return 0; }
inline PetscErrorCode PetscFinalize() { return 0; }

inline PetscErrorCode MatCreate(MPI_Comm, Mat* a) { *a = new _p_Mat(); (*a)->n_row = (*a)->n_col = 0; (*a)->last_row = -1; return 0; }
inline PetscErrorCode MatSetType(Mat, const char*) { return 0; }
inline PetscErrorCode MatSetSizes(Mat a, PetscInt m, PetscInt n, PetscInt, PetscInt) {
  a->n_row = m; a->n_col = n;
  a->rowptr.assign((std::size_t)m + 1, 0); a->colind.clear(); a->val.clear(); a->last_row = -1;
  return 0;
}
inline PetscErrorCode MatSeqAIJSetPreallocation(Mat a, PetscInt, const PetscInt* nnz) {
  if (nnz) a->row_nnz.assign(nnz, nnz + a->n_row);
  return 0;
}
inline PetscErrorCode MatSetValues(Mat a, PetscInt m, const PetscInt* im, PetscInt n, const PetscInt* in,
                                   const PetscScalar* v, InsertMode) {
  // the reference inserts whole rows in ascending row order, one call per row
  for (PetscInt r(0); r < m; ++r) {
    const PetscInt row(im[r]);
    for (PetscInt fill(a->last_row + 1); fill <= row; ++fill)
      a->rowptr[fill] = (int)a->colind.size();
    a->last_row = row;
    for (PetscInt c(0); c < n; ++c) { a->colind.push_back(in[c]); a->val.push_back(v[(std::size_t)r * n + c]); }
    a->rowptr[row + 1] = (int)a->colind.size();
  }
  return 0;
}
inline PetscErrorCode MatAssemblyBegin(Mat, MatAssemblyType) { return 0; }
inline PetscErrorCode MatAssemblyEnd(Mat a, MatAssemblyType) {
  for (PetscInt fill(a->last_row + 1); fill < a->n_row; ++fill)
    a->rowptr[fill + 1] = (int)a->colind.size();
  if (a->n_row) a->rowptr[a->n_row] = (int)a->colind.size();
  return 0;
}
inline PetscErrorCode MatDestroy(Mat* a) { delete *a; *a = 0; return 0; }

inline PetscErrorCode VecCreate(MPI_Comm, Vec* v) { *v = new _p_Vec(); return 0; }
inline PetscErrorCode VecSetType(Vec, const char*) { return 0; }
inline PetscErrorCode VecSetSizes(Vec v, PetscInt n, PetscInt) { v->v.assign((std::size_t)n, 0.0); return 0; }
inline PetscErrorCode VecZeroEntries(Vec v) { std::fill(v->v.begin(), v->v.end(), 0.0); return 0; }
inline PetscErrorCode VecSetValue(Vec v, PetscInt i, PetscScalar x, InsertMode) { v->v[(std::size_t)i] = x; return 0; }
inline PetscErrorCode VecAssemblyBegin(Vec) { return 0; }
inline PetscErrorCode VecAssemblyEnd(Vec) { return 0; }
inline PetscErrorCode VecDuplicate(Vec v, Vec* w) { *w = new _p_Vec(); (*w)->v.assign(v->v.size(), 0.0); return 0; }
inline PetscErrorCode VecGetValues(Vec v, PetscInt n, const PetscInt* idx, PetscScalar* out) {
  for (PetscInt i(0); i < n; ++i) out[i] = v->v[(std::size_t)idx[i]];
  return 0;
}
inline PetscErrorCode VecNorm(Vec v, NormType, PetscReal* r) {
  double s(0.0); for (double x : v->v) s += x * x; *r = std::sqrt(s); return 0;
}
inline PetscErrorCode VecDestroy(Vec* v) { delete *v; *v = 0; return 0; }

inline PetscErrorCode KSPCreate(MPI_Comm, KSP* k) {
  *k = new _p_KSP(); (*k)->hdr.prefix = 0; (*k)->vec_rhs = 0; (*k)->normtype = KSP_NORM_PRECONDITIONED;
  (*k)->guess_nonzero = PETSC_FALSE; (*k)->rtol = 1e-5; (*k)->atol = 1e-50; (*k)->dtol = 1e5;
  (*k)->maxits = 10000; (*k)->restart = 30; (*k)->mgs = false; (*k)->monitor = 0;
  (*k)->pc.levels = 0; (*k)->A = 0; (*k)->last_x = 0;
  return 0;
}
inline PetscErrorCode KSPSetType(KSP k, const char* t) { k->type = t; return 0; }
inline PetscErrorCode KSPSetInitialGuessNonzero(KSP k, PetscBool f) { k->guess_nonzero = f; return 0; }
inline PetscErrorCode KSPSetTolerances(KSP k, PetscReal rtol, PetscReal atol, PetscReal dtol, PetscInt maxits) {
  k->rtol = rtol; k->atol = atol; k->dtol = dtol; k->maxits = maxits; return 0;
}
typedef int oracle_shim_orthog_tag;
static const oracle_shim_orthog_tag KSPGMRESModifiedGramSchmidtOrthogonalization = 1;
inline PetscErrorCode KSPGMRESSetOrthogonalization(KSP k, oracle_shim_orthog_tag t) { k->mgs = (t == 1); return 0; }
inline PetscErrorCode KSPGMRESSetRestart(KSP k, PetscInt r) { k->restart = r; return 0; }
inline PetscErrorCode KSPMonitorSet(KSP k, oracle_shim_monitor_t m, void*, void*) { k->monitor = m; return 0; }
inline PetscErrorCode KSPGetPC(KSP k, PC* pc) { *pc = &k->pc; return 0; }
inline PetscErrorCode KSPSetOperators(KSP k, Mat a, Mat) { k->A = a; return 0; }
inline PetscErrorCode PCSetType(PC pc, const char* t) { pc->type = t; return 0; }
inline PetscErrorCode PCFactorSetLevels(PC pc, PetscInt l) { pc->levels = l; return 0; }
inline PetscErrorCode KSPSolve(KSP k, Vec b, Vec x) {
  oracle_shim::capture_t& c(oracle_shim::capture());
  c.A = k->A; c.b = b; c.ksp = k; k->vec_rhs = b; k->last_x = x;
  std::fill(x->v.begin(), x->v.end(), 0.0);
  c.last_its = c.solve_hook ? c.solve_hook(k, k->A, b, x) : 0;
  return 0;
}
inline PetscErrorCode KSPBuildResidual(KSP k, Vec, Vec, Vec* r) {
  // r = b - A x for the current iterate
  *r = new _p_Vec(); (*r)->v = k->vec_rhs->v;
  if (k->A && k->last_x)
    for (PetscInt i(0); i < k->A->n_row; ++i)
      for (int e(k->A->rowptr[i]); e < k->A->rowptr[i + 1]; ++e)
        (*r)->v[i] -= k->A->val[e] * k->last_x->v[k->A->colind[e]];
  return 0;
}
inline PetscErrorCode KSPDestroy(KSP* k) { delete *k; *k = 0; return 0; }
inline PetscErrorCode PetscStrncpy(char* d, const char* s, std::size_t n) { std::strncpy(d, s, n); if (n) d[n - 1] = 0; return 0; }
inline PetscErrorCode PetscStrtolower(char* s) { for (; *s; ++s) *s = (char)std::tolower(*s); return 0; }

#endif
