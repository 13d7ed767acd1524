/* TEST INFRASTRUCTURE (oracle): CPU restatement, in plain C, of the reference's
 * algorithm for the assembly + solve hot path (thomashilke/tfel). It exists to CHECK
 * the CUDA path: only tests/, __graft_entry__.smoke(), oracle/ref_driver.cpp and the
 * cpu_baseline / reference legs of bench.py may link, load or call it. The product
 * (tfel_b200/, include/) never does.
 *
 * Pinning: every function here is validated against the UNMODIFIED reference compiled
 * in oracle/_ref (see oracle/Makefile, tests/test_oracle_vs_reference.py) and against
 * the committed fixtures in tests/golden/ that the reference produced.
 * The Krylov part restates PETSc 3.7 semantics (KSPGMRES + PCILU, README.md:11 of the
 * reference; PETSc itself is not in /root/reference): iterate-for-iterate parity with
 * PETSc is UNPINNED (no reference test asserts anything at that boundary); solutions
 * are pinned through A x = b on the reference-captured system.
 */
#ifndef TFEL_ORACLE_H
#define TFEL_ORACLE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* finite elements (reference src/core/fe.hpp); ORC_FE_P2 on dim 3 is the NEW
 * tetrahedron P2 element the reference does not have (cell.hpp:948-952). */
enum { ORC_FE_P1 = 1, ORC_FE_P1_BUBBLE = 2, ORC_FE_P2 = 3 };

typedef struct {
  int test_comp, test_d;     /* component index, 0 = value / d = d-th space derivative */
  int trial_comp, trial_d;   /* trial_comp < 0: linear form */
  double c;                  /* constant factor */
  const double* table;       /* optional coefficient at (cell, q): [K][nq], or NULL */
} orc_term;

typedef struct {
  int dim;                    /* 2: triangles, 3: tetrahedra */
  int n_comp;
  const int* fe;              /* [n_comp] */
  const unsigned* const* dof_map; /* [n_comp] -> [K][nd_c] */
  const size_t* n_dof;        /* [n_comp] */
} orc_space;

void orc_free(void* p);

int orc_fe_ndof(int dim, int fe);
/* entity kinds carrying dofs: n_dof_per_subdomain(sd), sd = 0..dim */
void orc_fe_dofs_per_entity(int dim, int fe, int* out);
void orc_fe_nodes(int dim, int fe, double* nodes /* [nd][dim] */);
void orc_fe_tabulate(int dim, int fe, int npts, const double* pts, double* out /* [npts][nd][dim+1] */);
/* returns number of points, or -1 if unknown; x may be NULL to query size */
int orc_quadrature(const char* name, double* x, double* w, int* dim);

void orc_gen_square_mesh(double x1, double x2, unsigned n1, unsigned n2, double* vertices, unsigned* cells);
void orc_gen_cube_mesh(double x1, double x2, double x3, unsigned n1, unsigned n2, unsigned n3,
                       double* vertices, unsigned* cells);
void orc_geometry(int dim, const double* vertices, const unsigned* cells, size_t K, double* vol, double* jmt);

size_t orc_boundary_submesh(int dim, const unsigned* cells, size_t K,
                            unsigned** el, unsigned** el_id, unsigned** sd_id);
/* face-neighbour table of mesh<cell>::compute_cell_neighbours (mesh.hpp:301-332): nb[k][j] = the cell on the other
 * side of local face j of cell k, -1 on the boundary */
void orc_cell_neighbours(int dim, const unsigned* cells, size_t K, int* nb);

size_t orc_fes_build(int dim, int fe, const unsigned* cells, size_t K, unsigned* dof_map);
void orc_fes_g2l(const unsigned* dof_map, size_t K, int nd, size_t N, unsigned* g2l);
size_t orc_fes_dirichlet_dofs(int dim, int fe, const unsigned* cells, size_t K,
                              const unsigned* bcells, size_t F, int nvb, unsigned** dofs);
void orc_dof_coordinates(int dim, int fe, const double* vertices, const unsigned* cells,
                         const unsigned* g2l, const unsigned* dofs, size_t n, double* x);

/* CSR pattern the reference hands to PETSc (solver.cpp:71-96) for a bilinear form on
 * (test space, trial space): all (test dof, trial dof) pairs of every cell for rows
 * that are not Dirichlet, plus the identity entry of every Dirichlet row. */
size_t orc_pattern(size_t K, const orc_space* te, const orc_space* tr,
                   const unsigned char* dirichlet_row, int** rowptr, int** colind);

/* scalar form order (bilinear_form.hpp:64-109) when composite == 0, composite order
 * (composite_bilinear_form.hpp:68-120: one add per (q, i, j)) otherwise */
void orc_assemble_bilinear(int dim, const double* vertices, const unsigned* cells, size_t K,
                           const orc_space* te, const orc_space* tr,
                           const char* quadrature, const orc_term* terms, int n_terms,
                           const unsigned char* dirichlet_row, int composite,
                           const int* rowptr, const int* colind, double* val);
void orc_assemble_linear(int dim, const double* vertices, const unsigned* cells, size_t K,
                         const orc_space* te, const char* quadrature,
                         const orc_term* terms, int n_terms, int composite, double* b);
/* rank-0 integral: sum_k |K| sum_q w_q table[k][q] (form.hpp:87-125) */
double orc_integrate(int dim, const double* vertices, const unsigned* cells, size_t K,
                     const char* quadrature, const double* table);
/* quadrature points in space coordinates [K][nq][dim] (cell.hpp:456-468) */
void orc_quadrature_points(int dim, const double* vertices, const unsigned* cells, size_t K,
                           const char* quadrature, double* xq);
/* FE field (or its d-th derivative) at the quadrature points [K][nq] (expression.hpp:114-128) */
void orc_field_at_quadrature(int dim, int fe, const double* vertices, const unsigned* cells, size_t K,
                             const unsigned* dof_map, const double* coef, int d,
                             const char* quadrature, double* out);

void orc_spmv(int n, const int* rowptr, const int* colind, const double* val, const double* x, double* y);
int orc_cg_jacobi(int n, const int* rowptr, const int* colind, const double* val,
                  const double* b, double* x, double rtol, double atol, int maxits, double* final_rnorm);
int orc_gmres_ilu(int n, const int* rowptr, const int* colind, const double* val,
                  const double* b, double* x, int restart, int ilu_levels,
                  double rtol, double atol, double dtol, int maxits, double* final_rnorm);
/* ILU(0) in place on a copy of val: returns factors in lu[nnz] (unit L below diag, U on/above) */
int orc_ilu0(int n, const int* rowptr, const int* colind, const double* val, double* lu);
/* pattern of ILU(levels) as PCFactorSetLevels builds it (solver.hpp:162-163); *out_colind is malloc'ed (orc_free) */
size_t orc_iluk(int n, const int* rowptr, const int* colind, const double* val, int levels, int* out_rowptr,
                int** out_colind, double** out_lu);
size_t orc_iluk_pattern(int n, const int* rowptr, const int* colind, int levels, int* out_rowptr, int** out_colind);

#ifdef __cplusplus
}
#endif

#endif
