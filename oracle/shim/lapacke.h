// TEST INFRASTRUCTURE (oracle): stand-in for LAPACKE 3.5 <lapacke.h>, absent in
// this image. Implements only what the reference calls on this path
// (src/core/linear_algebra.hpp:15-35, src/core/solver.hpp:70, src/core/solver.cpp:40):
// dgetrf / dgetri / dgetrs as plain partial-pivoting LU (Doolittle, first max pivot),
// row- or column-major. Results agree with LAPACK to rounding, not bit-for-bit.
#ifndef ORACLE_SHIM_LAPACKE_H
#define ORACLE_SHIM_LAPACKE_H

#include <cmath>
#include <vector>
#include <cstddef>

typedef int lapack_int;
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102

namespace oracle_shim_lapack {
  struct view {
    double* a; lapack_int lda; bool row_major;
    double& operator()(lapack_int i, lapack_int j) const {
      return row_major ? a[(std::size_t)i * lda + j] : a[(std::size_t)j * lda + i];
    }
  };
}

inline lapack_int LAPACKE_dgetrf(int layout, lapack_int m, lapack_int n,
                                 double* a, lapack_int lda, lapack_int* ipiv) {
  oracle_shim_lapack::view A{a, lda, layout == LAPACK_ROW_MAJOR};
  lapack_int info(0);
  const lapack_int mn(m < n ? m : n);
  for (lapack_int k(0); k < mn; ++k) {
    lapack_int p(k);
    double best(std::fabs(A(k, k)));
    for (lapack_int i(k + 1); i < m; ++i)
      if (std::fabs(A(i, k)) > best) { best = std::fabs(A(i, k)); p = i; }
    ipiv[k] = p + 1;
    if (A(p, k) != 0.0) {
      if (p != k)
        for (lapack_int j(0); j < n; ++j) { double t(A(k, j)); A(k, j) = A(p, j); A(p, j) = t; }
      const double r(1.0 / A(k, k));
      for (lapack_int i(k + 1); i < m; ++i) A(i, k) *= r;
    } else if (info == 0) {
      info = k + 1;
    }
    for (lapack_int j(k + 1); j < n; ++j)
      for (lapack_int i(k + 1); i < m; ++i)
        A(i, j) -= A(i, k) * A(k, j);
  }
  return info;
}

inline lapack_int LAPACKE_dgetrs(int layout, char trans, lapack_int n, lapack_int nrhs,
                                 const double* a, lapack_int lda, const lapack_int* ipiv,
                                 double* b, lapack_int ldb) {
  if (trans != 'N' && trans != 'n') return -2;
  oracle_shim_lapack::view A{const_cast<double*>(a), lda, layout == LAPACK_ROW_MAJOR};
  oracle_shim_lapack::view B{b, ldb, layout == LAPACK_ROW_MAJOR};
  for (lapack_int c(0); c < nrhs; ++c) {
    for (lapack_int k(0); k < n; ++k) {
      const lapack_int p(ipiv[k] - 1);
      if (p != k) { double t(B(k, c)); B(k, c) = B(p, c); B(p, c) = t; }
    }
    for (lapack_int i(0); i < n; ++i)
      for (lapack_int j(0); j < i; ++j)
        B(i, c) -= A(i, j) * B(j, c);
    for (lapack_int i(n - 1); i >= 0; --i) {
      for (lapack_int j(i + 1); j < n; ++j)
        B(i, c) -= A(i, j) * B(j, c);
      B(i, c) /= A(i, i);
    }
  }
  return 0;
}

inline lapack_int LAPACKE_dgetri(int layout, lapack_int n, double* a, lapack_int lda,
                                 const lapack_int* ipiv) {
  oracle_shim_lapack::view A{a, lda, layout == LAPACK_ROW_MAJOR};
  for (lapack_int i(0); i < n; ++i)
    if (A(i, i) == 0.0) return i + 1;
  std::vector<double> lu((std::size_t)n * n), inv((std::size_t)n * n, 0.0);
  for (lapack_int i(0); i < n; ++i)
    for (lapack_int j(0); j < n; ++j)
      lu[(std::size_t)i * n + j] = A(i, j);
  for (lapack_int i(0); i < n; ++i) inv[(std::size_t)i * n + i] = 1.0;
  LAPACKE_dgetrs(LAPACK_ROW_MAJOR, 'N', n, n, lu.data(), n, ipiv, inv.data(), n);
  for (lapack_int i(0); i < n; ++i)
    for (lapack_int j(0); j < n; ++j)
      A(i, j) = inv[(std::size_t)i * n + j];
  return 0;
}

#endif
