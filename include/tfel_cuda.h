/* tfel_cuda.h -- C ABI of libtfel_cuda.so: the B200 (sm_100a) assembly + solve path behind
 * TFEL's `bilinear_form::operator+=(integrate<quad>(expr, mesh))` / `linear_form` /
 * `a.solve(b, s)` surface.
 *
 * Plain C: opaque handles, pointers and sizes. Every entry point returns 0 on success and a
 * non-zero code on failure; tfelcu_last_error() returns the message (the C++ wrappers in
 * include/tfel/ rethrow it as std::string, the reference's error convention, e.g.
 * src/core/solver.hpp:132). Pointers documented "host or device" are classified with
 * cudaPointerGetAttributes. Citations are into the reference tree (thomashilke/tfel).
 *
 * There is no CPU fallback: every call needs a CUDA device.
 */
#ifndef TFEL_CUDA_H
#define TFEL_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tfelcu_ctx tfelcu_ctx;
typedef struct tfelcu_mesh tfelcu_mesh;
typedef struct tfelcu_fes tfelcu_fes;
typedef struct tfelcu_matrix tfelcu_matrix;
typedef struct tfelcu_vector tfelcu_vector;
typedef struct tfelcu_solver tfelcu_solver;

/* ---- enumerations ------------------------------------------------------------------- */

/* finite elements, src/core/fe.hpp:260-457 (triangle), :494-654 (tetrahedron).
 * TFELCU_FE_P2 on tetrahedra is new: the reference has none (cell.hpp:948-952). */
enum { TFELCU_FE_P1 = 1, TFELCU_FE_P1_BUBBLE = 2, TFELCU_FE_P2 = 3 };

/* quadrature formulas, src/core/quadrature.hpp:80-221, quadrature.cpp:52-367 */
enum {
  TFELCU_QUAD_TRI_QF1PT = 1, TFELCU_QUAD_TRI_QF2PT = 2, TFELCU_QUAD_TRI_QF5PT = 3,
  TFELCU_QUAD_TRI_QF1PTLUMP = 4,
  TFELCU_QUAD_TET_QF1PTET = 10, TFELCU_QUAD_TET_QF4PTET = 11, TFELCU_QUAD_TET_QF5PTET = 12,
  TFELCU_QUAD_TET_SYM1 = 13, TFELCU_QUAD_TET_SYM4 = 14, TFELCU_QUAD_TET_SYM10 = 15,
  TFELCU_QUAD_TET_SYM20 = 16, TFELCU_QUAD_TET_SYM35 = 17, TFELCU_QUAD_TET_SYM56 = 18
};

/* rank-0 coefficient factors of an integrand (leaves of src/core/expression.hpp:31-196) */
enum {
  TFELCU_FACTOR_FIELD = 1,    /* finite_element_function<fe, d>, expression.hpp:89-139 */
  TFELCU_FACTOR_TABLE = 2,    /* free_function / std_function tabulated at the quadrature points
                                 (expression.hpp:31-87): double[K][nq], host or device */
  TFELCU_FACTOR_PROGRAM = 3,  /* function of x evaluated on device (postfix program) */
  TFELCU_FACTOR_FUNCTION = 4  /* function of x as DEVICE CODE of the caller: a kernel of the caller's own module (nvcc, user's
                                 translation unit; include/tfel/tfel_device.cuh instantiates it from any functor type), which
                                 the library launches on its stream when the form is assembled -- see tfelcu_tabulate_fn.
                                 No host evaluation, no host <-> device copy. (Device function pointers do not cross
                                 CUDA modules on sm_100a without relocatable device code, and -rdc costs the element kernels
                                 their register allocation: DESIGN.md section 5.) */
};

/* postfix opcodes of TFELCU_FACTOR_PROGRAM */
enum {
  TFELCU_OP_CONST = 0, TFELCU_OP_COORD = 1, TFELCU_OP_ADD = 2, TFELCU_OP_SUB = 3, TFELCU_OP_MUL = 4,
  TFELCU_OP_DIV = 5, TFELCU_OP_NEG = 6, TFELCU_OP_SQRT = 7, TFELCU_OP_EXP = 8, TFELCU_OP_SIN = 9,
  TFELCU_OP_COS = 10, TFELCU_OP_POW = 11, TFELCU_OP_LT = 12, TFELCU_OP_ABS = 13
};

typedef struct { int32_t op; int32_t pad; double imm; } tfelcu_op;

/* Where a FUNCTION factor is evaluated: the quadrature points of every cell,
 * x_q = v_0 + sum_d xhat[q][d] (v_{d+1} - v_0) with v_j = vertices[cells[k][j]] (cell.hpp:456-468). All pointers are
 * device memory. */
typedef struct {
  int32_t dim, nq;
  size_t n_cells;
  const double* vertices;      /* [V][dim] */
  const uint32_t* cells;       /* [K][dim + 1] */
  const double* xhat;          /* [nq][dim] reference coordinates of the quadrature nodes */
} tfelcu_points;

/* The caller's launcher of a FUNCTION factor: enqueue on `stream` (a cudaStream_t) device work that writes
 * out[k * nq + q] = f(x_q of cell k) for every cell; return 0, or nonzero when the launch failed. Called on the host,
 * once per += that uses the factor; it must not synchronise. */
typedef int (*tfelcu_tabulate_fn)(const void* closure, const tfelcu_points* where, double* out, void* stream);

typedef struct {
  int32_t kind;               /* TFELCU_FACTOR_* */
  int32_t d;                  /* FIELD: 0 = value, k = derivative along x_k (d<k>(), expression.hpp:369-373) */
  const tfelcu_fes* fes;      /* FIELD: space of the field */
  const double* data;         /* FIELD: coefficients [n_dof] (host or device); TABLE: [K][nq] */
  const tfelcu_op* ops;       /* PROGRAM: host array */
  int32_t n_ops;
  int32_t pad;
  tfelcu_tabulate_fn fn;      /* FUNCTION: launcher of the caller's tabulation kernel */
  const void* closure;        /* FUNCTION: device memory handed to fn (the functor object), may be NULL */
} tfelcu_factor;

/* One monomial  c * prod(factors) * D^{test_d} psi^{test_comp}_i * D^{trial_d} phi^{trial_comp}_j  of an
 * integrand in sum-of-products form (what form<arg, rank, d> placeholders, expression.hpp:14-29,
 * combined with + - * reduce to). trial_comp < 0 for linear forms. */
typedef struct {
  int32_t test_comp, test_d;
  int32_t trial_comp, trial_d;
  double c;
  int32_t n_factors;
  int32_t factor[5];          /* indices into the factor array of the call */
} tfelcu_term;

enum { TFELCU_METHOD_CG = 1, TFELCU_METHOD_GMRES = 2 };
/* TFELCU_PC_ILU: incomplete LU in natural ordering with level of fill tfelcu_solver_params::ilu_levels -- PCILU +
 * PCFactorSetLevels(pc, ilufill) of the reference (solver.hpp:162-163). TFELCU_PC_ILU0 is the same value: ilu_levels = 0
 * (the default of a zero-initialised parameter block) is ILU(0). */
enum { TFELCU_PC_NONE = 0, TFELCU_PC_JACOBI = 1, TFELCU_PC_BLOCK_JACOBI = 2, TFELCU_PC_ILU0 = 3, TFELCU_PC_ILU = 3 };
/* Scatter strategies of the assembly, all free of atomics and bitwise reproducible. ROW_GATHER: one thread per row
 * recomputes its row of every incident element matrix. COLORED: one launch per colour of the cell conflict graph.
 * PATCH: rows grouped into compact patches of the mesh (quadtree / octree leaves of the dof coordinates); one CTA
 * computes every element matrix of its patch once in shared memory, then each row gathers. AUTO (default): the
 * measured winner; TFELCU_SCATTER=rows|patch|colored|tile in the environment overrides AUTO.
 * TILE: cell-centric inside shared-memory tiles of the mesh (vertex-based P1 spaces on triangles, constant
 * coefficients): every element matrix once, every value written once; explicit request only (measured slower than the
 * rows). For the same class of forms AUTO takes the closed-form P1 row kernel (p1rows.cu); with
 * either, a constant-coefficient linear form on the same space that follows the += is assembled in the same pass. */
enum { TFELCU_SCATTER_ROW_GATHER = 0, TFELCU_SCATTER_COLORED = 1, TFELCU_SCATTER_PATCH = 2, TFELCU_SCATTER_AUTO = 3,
       TFELCU_SCATTER_TILE = 4 };
enum { TFELCU_SPMV_AUTO = 0, TFELCU_SPMV_VECTOR = 1, TFELCU_SPMV_TMA_STREAM = 2 };

/* same keys as the dictionary solver::petsc::gmres_ilu requires (solver.hpp:125-133) */
typedef struct {
  int32_t method, precond;
  uint32_t maxits, restart;
  double rtol, atol, dtol;
  int32_t block_size;          /* TFELCU_PC_BLOCK_JACOBI */
  int32_t check_every;         /* CG: iterations between host convergence polls (0 = default) */
  int32_t spmv_kernel;         /* TFELCU_SPMV_* */
  int32_t profile_every;       /* > 0: bracket the SpMV launch of every profile_every-th iteration with CUDA events
                                  (at most 512 samples per solve) and report their average duration */
  int32_t ilu_levels;          /* TFELCU_PC_ILU: level of fill, the dictionary's "ilufill" (solver.hpp:163); 0 = ILU(0) */
  int32_t pad;
} tfelcu_solver_params;

typedef struct {
  uint32_t its;
  int32_t converged;           /* 1 converged, 0 maxits reached, -1 diverged (dtol), -2 breakdown,
                                  -3 a peer rank did not answer in time (distributed solves; the context falls back to NCCL) */
  double rnorm, bnorm;         /* norm the convergence test used */
  double t_setup_s, t_solve_s; /* device time, CUDA events */
  uint64_t n_launches;         /* kernels launched by the solve */
  uint64_t n_spmv;
  double spmv_avg_ms;          /* profile_every > 0: average duration of the sampled SpMV launches */
  uint32_t spmv_samples;
  uint32_t restart_used;       /* GMRES: restart length actually used (the dictionary's, bounded by maxits and by the device
                                  memory the Krylov basis may take); smaller than requested => a warning on stderr */
  uint64_t ilu_nnz;            /* TFELCU_PC_ILU: stored entries of the factors (0 otherwise) */
  int32_t ilu_levels;          /* level of fill the factorisation used */
  int32_t ilu_zero_pivot_row;  /* first row whose pivot is exactly 0, or -1 */
  double t_precond_setup_s;    /* part of t_setup_s spent building the preconditioner */
  double t_symbolic_s;         /* part of it spent in the symbolic phase (0 when the cached pattern was reused) */
} tfelcu_solver_report;

/* ---- context ------------------------------------------------------------------------ */

const char* tfelcu_last_error(void);
const char* tfelcu_version(void);
int tfelcu_ctx_create(int device, tfelcu_ctx** ctx);
/* one process per GPU: rank / n_ranks and the 128-byte NCCL unique id of rank 0 */
int tfelcu_nccl_unique_id(void* out128);
int tfelcu_ctx_create_distributed(int device, int rank, int n_ranks, const void* nccl_unique_id128,
                                  tfelcu_ctx** ctx);
int tfelcu_ctx_destroy(tfelcu_ctx* ctx);
int tfelcu_ctx_set_stream(tfelcu_ctx* ctx, void* cuda_stream);
void* tfelcu_ctx_stream(tfelcu_ctx* ctx);
int tfelcu_ctx_synchronize(tfelcu_ctx* ctx);
int tfelcu_ctx_rank(const tfelcu_ctx* ctx, int* rank, int* n_ranks);
/* kernels launched through this context since creation (all entry points) */
uint64_t tfelcu_ctx_launch_count(const tfelcu_ctx* ctx);
/* 1 when the distributed Krylov loop exchanges halos and reduction scalars through peer-mapped device memory
 * (CUDA IPC over NVLink), 0 when it uses NCCL send / recv / allreduce (TFELCU_NO_P2P=1, or IPC unavailable) */
int tfelcu_ctx_p2p_enabled(const tfelcu_ctx* ctx);
/* diagnostic: measured FP64 FMA throughput of the device in TFLOP/s (8 independent chains per thread, 8 CTAs per SM) */
int tfelcu_ctx_fp64_peak(tfelcu_ctx* ctx, double* tflops);
/* diagnostic (collective): average microseconds of one 2-value allreduce through the peer arenas */
int tfelcu_ctx_p2p_bench(tfelcu_ctx* ctx, int reps, double* allreduce_us);

/* ---- mesh: mesh<cell> / fe_mesh<cell>, src/core/mesh.hpp:142-158,355-364 ------------------ */

/* vertices double[V][dim], cells uint32[K][dim+1] (host or device; copied). Cells whose vertex
 * ids are not ascending are handled like mesh.hpp:286-299. */
int tfelcu_mesh_create(tfelcu_ctx* ctx, int dim, const double* vertices, size_t n_vertices,
                       const uint32_t* cells, size_t n_cells, tfelcu_mesh** mesh);
/* gen_square_mesh / gen_cube_mesh generated on device, src/core/mesh.cpp:19-51, 54-135 */
int tfelcu_mesh_gen_square(tfelcu_ctx* ctx, double x1, double x2, uint32_t n1, uint32_t n2, tfelcu_mesh** mesh);
int tfelcu_mesh_gen_cube(tfelcu_ctx* ctx, double x1, double x2, double x3,
                         uint32_t n1, uint32_t n2, uint32_t n3, tfelcu_mesh** mesh);
int tfelcu_mesh_destroy(tfelcu_mesh* mesh);
int tfelcu_mesh_info(const tfelcu_mesh* mesh, int* dim, size_t* n_vertices, size_t* n_cells);
int tfelcu_mesh_download(const tfelcu_mesh* mesh, double* vertices, uint32_t* cells);
/* get_cell_volume / get_jmt, cell.hpp:474-502, 799-837: vol[K], jmt[K][dim][dim] (host buffers) */
int tfelcu_mesh_geometry(const tfelcu_mesh* mesh, double* vol, double* jmt);
/* get_boundary_submesh, mesh.hpp:485-521: faces seen once, ascending face order */
int tfelcu_mesh_boundary(tfelcu_mesh* mesh, size_t* n_faces);
int tfelcu_mesh_boundary_download(const tfelcu_mesh* mesh, uint32_t* el, uint32_t* parent_cell,
                                  uint32_t* parent_subdomain);
/* compute_cell_neighbours, mesh.hpp:301-332: nb[K][dim + 1] (host or device buffer), nb[k][j] = the cell across local
 * face j of cell k, -1 on the boundary */
int tfelcu_mesh_neighbours(tfelcu_mesh* mesh, int32_t* nb);

/* ---- finite element space: finite_element_space<fe>, src/core/fes.hpp:18-82 ---------------- */

int tfelcu_fes_create(tfelcu_ctx* ctx, tfelcu_mesh* mesh, int fe, tfelcu_fes** fes);
int tfelcu_fes_destroy(tfelcu_fes* fes);
int tfelcu_fes_info(const tfelcu_fes* fes, int* fe, size_t* n_dof, int* n_dof_per_cell);
/* first dof of every entity kind (fes.hpp:33-72 numbers vertices, then edges, faces, cell): offsets[0..3] for kinds
 * 0..3, offsets[4] = n_dof; a kind without dofs has offsets[sd] == offsets[sd + 1] */
int tfelcu_fes_kind_offsets(const tfelcu_fes* fes, size_t* offsets /* [5] */);
int tfelcu_fes_download_dof_map(const tfelcu_fes* fes, uint32_t* dof_map /* [K][n_dof_per_cell] */);
int tfelcu_fes_download_g2l(const tfelcu_fes* fes, uint32_t* g2l /* [n_dof][2] */);
/* get_dof_space_coordinate, fes.hpp:190-202, for dofs[n] (host); x[n][dim] (host) */
int tfelcu_fes_dof_coordinates(const tfelcu_fes* fes, const uint32_t* dofs, size_t n, double* x);
/* Dofs add_dirichlet_boundary would insert for a submesh (fes.hpp:100-149): boundary cells
 * uint32[F][nvb] (host or device); bcells == NULL selects the whole boundary of the mesh.
 * The sorted list stays in the space until the next call; download it with _boundary_dofs_get. */
int tfelcu_fes_boundary_dofs(tfelcu_fes* fes, const uint32_t* bcells, size_t n_bcells, int nvb, size_t* n);
int tfelcu_fes_boundary_dofs_get(const tfelcu_fes* fes, uint32_t* dofs, double* x /* [n][dim] or NULL */);
/* std::map::insert semantics: a dof that already has a value keeps it (fes.hpp:116,143) */
int tfelcu_fes_add_dirichlet(tfelcu_fes* fes, const uint32_t* dofs, const double* values, size_t n);
/* add_dirichlet_boundary(dm, value) entirely on device for the list of the last _boundary_dofs call */
int tfelcu_fes_add_dirichlet_const(tfelcu_fes* fes, double value);
int tfelcu_fes_dirichlet_count(const tfelcu_fes* fes, size_t* n);
int tfelcu_fes_dirichlet_get(const tfelcu_fes* fes, uint32_t* dofs, double* values);
/* output side (fes.hpp:464-504). to_mesh_vertex_data: out[V], out[cells[k][n]] = coefficients[dof(k, n)] for the vertex
 * dofs of every cell (0 for vertices no cell uses); to_mesh_cell_data: out[K], the value of the discrete function at the
 * barycentre of every cell (element::evaluate, fes.hpp:247-255). coefficients / out: host or device. */
/* reference-element basis of a finite element (fe.hpp:260-654: phi / dphi), host only: out[i][0] = phi_i(x_hat),
 * out[i][1 + s] = d phi_i / d x_hat_s for the n_dof_per_element basis functions; returns their number in *n_dof */
int tfelcu_fe_tabulate(int dim, int fe, const double* x_hat, double* out, int* n_dof);
int tfelcu_fes_vertex_values(const tfelcu_fes* fes, const double* coefficients, double* out);
int tfelcu_fes_cell_values(const tfelcu_fes* fes, const double* coefficients, double* out);

/* ---- bilinear form / sparse matrix ----------------------------------------------------------
 * bilinear_form<te, tr> (bilinear_form.hpp:9-19; composite: composite_bilinear_form.hpp:33-62):
 * test / trial spaces are lists of scalar spaces ([u0 | u1 | p] block layout). Creation sizes the
 * matrix, builds the CSR pattern the reference would hand to PETSc (solver.cpp:71-96) and
 * performs clear() (identity rows of the Dirichlet dofs registered SO FAR, :159-165). */
int tfelcu_matrix_create(tfelcu_ctx* ctx, int n_test, tfelcu_fes* const* test,
                         int n_trial, tfelcu_fes* const* trial, tfelcu_matrix** m);
/* Row-block distribution (one process per GPU): the form holds only the rows this rank owns -- up to 8 contiguous
 * segments [row_begin[i], row_end[i]) of GLOBAL rows (block layout [u0 | u1 | p]), ascending, disjoint, each inside one
 * component: e.g. the vertex dofs and the edge dofs of a horizontal strip of gen_square_mesh for every P2 component.
 * The segments of all ranks must tile [0, n_rows). Local rows are the segments concatenated; columns stay global.
 * The reference has no distributed mode. */
int tfelcu_matrix_create_rows(tfelcu_ctx* ctx, int n_test, tfelcu_fes* const* test, int n_trial, tfelcu_fes* const* trial,
                              int n_segments, const size_t* row_begin, const size_t* row_end, tfelcu_matrix** m);
int tfelcu_matrix_local_rows(const tfelcu_matrix* m, size_t* n_local);
int tfelcu_matrix_destroy(tfelcu_matrix* m);
int tfelcu_matrix_clear(tfelcu_matrix* m);
int tfelcu_matrix_info(const tfelcu_matrix* m, size_t* n_rows, size_t* n_cols, size_t* nnz);
int tfelcu_matrix_set_scatter(tfelcu_matrix* m, int scatter /* TFELCU_SCATTER_* */);
/* operator+=(integrate<quad>(expr, mesh)), bilinear_form.hpp:23-110 */
int tfelcu_matrix_assemble(tfelcu_matrix* m, int quad, const tfelcu_factor* factors, int n_factors,
                           const tfelcu_term* terms, int n_terms);
/* rows ascending, columns ascending, int32 indices, explicit zeros kept (solver.cpp:71-88); host buffers;
 * any of the three may be NULL */
int tfelcu_matrix_download_csr(const tfelcu_matrix* m, int32_t* rowptr, int32_t* colind, double* val);
/* number of colours of the cell conflict graph (TFELCU_SCATTER_COLORED) */
int tfelcu_matrix_color_count(tfelcu_matrix* m, int* n_colors);

/* ---- linear form / vectors ------------------------------------------------------------------
 * linear_form<fes> (linear_form.hpp:9-20); also plain device vectors in the block layout. */
int tfelcu_vector_create(tfelcu_ctx* ctx, int n_test, tfelcu_fes* const* test, tfelcu_vector** v);
int tfelcu_vector_create_rows(tfelcu_ctx* ctx, int n_test, tfelcu_fes* const* test, int n_segments,
                              const size_t* row_begin, const size_t* row_end, tfelcu_vector** v);
int tfelcu_vector_destroy(tfelcu_vector* v);
int tfelcu_vector_clear(tfelcu_vector* v);
int tfelcu_vector_size(const tfelcu_vector* v, size_t* n);
int tfelcu_vector_assemble(tfelcu_vector* v, int quad, const tfelcu_factor* factors, int n_factors,
                           const tfelcu_term* terms, int n_terms);
/* TFELCU_SCATTER_ROW_GATHER, _PATCH or _AUTO (default) for the += of a linear form */
int tfelcu_vector_set_scatter(tfelcu_vector* v, int scatter);
int tfelcu_vector_download(const tfelcu_vector* v, double* host);
int tfelcu_vector_upload(tfelcu_vector* v, const double* host);
double* tfelcu_vector_device_ptr(tfelcu_vector* v);
/* rank-0 integrate<quad>(expr, mesh), form.hpp:87-139: sum_k |K| sum_q w_q prod(factors) */
int tfelcu_integrate(tfelcu_ctx* ctx, tfelcu_mesh* mesh, int quad, const tfelcu_factor* factors,
                     int n_factors, const tfelcu_term* terms, int n_terms, double* result);

/* number of points / dimension of a quadrature formula, and the points of every cell in space coordinates
 * xq[K][nq][dim] (cell.hpp:456-468; host or device buffer): what free_function leaves (expression.hpp:31-87) are
 * evaluated at when a host callback has to be tabulated for TFELCU_FACTOR_TABLE */
int tfelcu_quadrature_size(int quad, int* n_points, int* dim);
int tfelcu_mesh_quadrature_points(tfelcu_mesh* mesh, int quad, double* xq);

/* ---- plain device arrays ------------------------------------------------------------------------
 * Coefficient vectors of discrete functions (finite_element_space::element, fes.hpp:219-364) that stay on the device
 * between a solve, the algebra on elements (fes.hpp:314-405) and the next += that uses them as coefficient fields.
 * _copy takes host or device pointers on either side; _lincomb computes out = a x + b y (y may be NULL: out = a x) on
 * device arrays, out may alias x or y. */
int tfelcu_array_create(tfelcu_ctx* ctx, size_t n, double** out);
int tfelcu_array_destroy(tfelcu_ctx* ctx, double* p);
int tfelcu_array_copy(tfelcu_ctx* ctx, double* dst, const double* src, size_t n);
int tfelcu_array_lincomb(tfelcu_ctx* ctx, size_t n, double a, const double* x, double b, const double* y, double* out);

/* ---- solver plugin: solver::basic_solver, src/core/solver.hpp:29-40 ---------------------------- */

int tfelcu_solver_create(tfelcu_ctx* ctx, const tfelcu_solver_params* params, tfelcu_solver** s);
int tfelcu_solver_destroy(tfelcu_solver* s);
/* set_operator(const sparse_matrix&): like the reference (solver.cpp:71-96: the matrix is COPIED into the solver's own
 * storage), the solver keeps a device copy of the rows this process owns, so the form may be cleared, re-assembled or
 * destroyed afterwards; a later set_operator on a form with the same pattern only copies the values and reuses the
 * SpMV plan and the symbolic phase of ILU(k) (Newton / time loops, test/navier_stokes_2d_p2_p1.cpp:292-438).
 * tfelcu_matrix_solve below works on the form's arrays without a copy and detaches the solver on return. */
int tfelcu_solver_set_operator(tfelcu_solver* s, tfelcu_matrix* m);
/* ... or a host CSR (what solver.cpp:71-96 builds from the std::map) */
int tfelcu_solver_set_operator_csr(tfelcu_solver* s, size_t n, const int32_t* rowptr,
                                   const int32_t* colind, const double* val);
/* solve(rhs, x, report), solver.cpp:155-177; rhs / x host or device, length n */
int tfelcu_solver_solve(tfelcu_solver* s, const double* rhs, double* x, tfelcu_solver_report* report);
/* bilinear_form::solve (bilinear_form.hpp:122-157): rhs = b with the Dirichlet rows replaced by their
 * values, solve, return the coefficients of the solution (host buffer x[n_cols]) */
int tfelcu_matrix_solve(tfelcu_matrix* m, const tfelcu_vector* b, tfelcu_solver* s, double* x,
                        tfelcu_solver_report* report);
/* the right-hand side solve() builds (host buffer), for parity checks */
int tfelcu_matrix_make_rhs(const tfelcu_matrix* m, const tfelcu_vector* b, double* rhs);
/* y = A x with the solver's SpMV kernel; x, y host or device. Timed variant for the roofline:
 * runs `reps` launches and returns the average kernel time (CUDA events on the context stream). */
/* distributed operator: owned entries of every rank -> full vector in reference numbering on every rank */
int tfelcu_solver_gather(tfelcu_solver* s, const double* x_local, double* x_global);
int tfelcu_solver_spmv(tfelcu_solver* s, const double* x, double* y);
int tfelcu_solver_spmv_bench(tfelcu_solver* s, int reps, double* avg_ms, double* algorithmic_bytes);
/* TFELCU_PC_ILU, for parity checks: size of the factors, then the factors themselves -- rows ascending, columns
 * ascending, L (unit diagonal not stored) left of the diagonal, U from the diagonal on; host buffers rowptr[n + 1],
 * colind[nnz], lu[nnz], any of which may be NULL. Same layout as PETSc's / the oracle's row-merge result. */
int tfelcu_solver_ilu_info(const tfelcu_solver* s, size_t* n, size_t* nnz, int* levels);
int tfelcu_solver_ilu_download(const tfelcu_solver* s, int64_t* rowptr, int32_t* colind, double* lu);

#ifdef __cplusplus
}
#endif

#endif /* TFEL_CUDA_H */
