// solver::cuda::* behind solver::basic_solver (src/core/solver.hpp:29-40): device-resident CSR operator,
// FP64 Krylov methods and preconditioners.
#pragma once

#include "structs.cuh"

namespace tfelcu {

// scalar step of the single-reduction (Chronopoulos-Gear) CG from gamma = r.u, rr = r.r, delta = u.Au
__device__ __forceinline__ void cgs_scalar_step(struct KrylovScalars* S, double gamma, double rr, double delta);

// device-side scalars of a Krylov iteration (one cache line, read by every kernel)
struct KrylovScalars {
  double rz, pq, alpha, beta, rr, bnorm2, target2, dtol2;
  double rz_new, spare0, spare1, spare2;
  int done;        // 0 running, 1 converged, -1 diverged, -2 breakdown
  int its;
  int pad0, pad1;
};

__device__ __forceinline__ void cgs_scalar_step(KrylovScalars* S, double gamma, double rr, double delta) {
  if (S->done) return;
  S->rr = rr;
  if (!(rr == rr) || !(delta == delta)) { S->done = -2; return; }
  if (rr <= S->target2) { S->done = 1; return; }
  if (rr > S->dtol2) { S->done = -1; return; }
  if ((double)S->its >= S->spare1) { S->done = 2; return; }   // maxits updates done
  double beta = 0.0, alpha;
  if (S->its == 0) alpha = gamma / delta;
  else { beta = gamma / S->rz; alpha = gamma / (delta - beta * gamma / S->alpha); }
  if (!(alpha == alpha) || delta == 0.0) { S->done = -2; return; }
  S->alpha = alpha; S->beta = beta; S->rz = gamma; S->pq = delta;
  S->its += 1;
}

struct SpmvPlan {
  int kernel = TFELCU_SPMV_VECTOR;
  int lanes = 8;                         // lanes per row of the vector kernel
  // TMA-streamed kernel: slabs of bounded rows and entries
  DevBuf<int32_t> desc;                  // [n_blocks][4]: first row, end row, first entry, end entry
  int n_blocks = 0;
  int row_lanes = 1;                     // threads per row of the in-slab row reduction
  int cfg = 0;                           // tile shape (spmv.cu: kCfgs)
  int part_begin = 0, part_count = 0;    // slab range of the launch being issued
  int n_interior = -1;                   // distributed operators: slabs [0, n_interior) reference owned columns only
  DevBuf<int16_t> delta16;               // column - row of every stored entry, when all fit 16 bits
  bool use_delta16 = false;
  int tile_nnz = 0, tile_rows = 0;
};

// consecutive levels [l0, l1) of a level schedule launched together: one CTA walking several small levels
// (cta == true) or one grid for a single large level
struct LevelRun { int l0, l1; bool cta; };

struct Ilu0 {
  // the matrix that is factorised: the operator itself, or (row-distributed operators) the diagonal block of this rank
  const int32_t* rowptr = nullptr; const int32_t* colind = nullptr; const double* val = nullptr;
  size_t nnz = 0;
  DevBuf<int32_t> own_rowptr, own_colind;
  DevBuf<double> own_val;
  DevBuf<double> lu;                     // factors on the pattern of A
  DevBuf<int32_t> diag;                  // position of the diagonal in each row
  DevBuf<int32_t> order_l, order_u;      // rows grouped by level
  std::vector<int32_t> level_ptr_l, level_ptr_u;
  DevBuf<int32_t> d_level_ptr_l, d_level_ptr_u;
  DevBuf<int32_t> meta_l, meta_u;        // [n][4] in schedule order: row, first entry, end entry of its L / U part
  DevBuf<int32_t> pos_l, pos_u;          // [n]: schedule position of every row
  std::vector<LevelRun> runs_l, runs_u;
  // ---- ILU(k) without level scheduling (iluk.cu): factors on the level-of-fill pattern, 64-bit row pointers
  int levels = 0;                        // level of fill (the dictionary's "ilufill")
  bool syncfree = true;                  // false: level-scheduled launches (TFELCU_ILU=levels), ILU(0) only
  DevBuf<long long> fptr, fdiag;         // [n + 1] first entry of a row, [n] position of its diagonal
  const int32_t* fcol = nullptr;         // columns of the factors, ascending per row (levels == 0: the columns of A)
  DevBuf<int32_t> fcol_store;
  size_t fnnz = 0;
  DevBuf<unsigned int> ticket;           // rows are started in dependency order
  DevBuf<int> stuck, rowdone;
  DevBuf<double> ytmp;                   // result of the forward sweep
  int lanes = 8;                         // lanes per row of the triangular sweeps
  int gate_sleep_ns = 0;                 // pause between two reads at the gate of a row (< 0: no gate)
  size_t tri_smem = 0;                   // dynamic shared memory that limits the CTAs per SM of the sweeps (0: no limit)
  int zero_pivot_row = -1;
  int n_levels_l = 0, n_levels_u = 0;    // depth of the dependency graphs of the two sweeps
  double t_symbolic_s = 0.0;
};

// halo exchange plan of a row-block distributed operator (dist.cu)
struct DistPlan {
  RowMap rows;                           // rows of this rank
  std::vector<RowMap> all_rows;          // rows of every rank
  size_t n_halo = 0;                     // what a vector holds behind its owned entries: padding up to halo_base, then one
                                         // entry per column owned by another rank
  size_t halo_base = 0;                  // index of the first halo entry behind the owned ones (rounded up to a 128-byte line)
  size_t n_front = 0, front_pad = 0;     // halo entries owned by lower ranks sit IN FRONT of the owned entries, at
                                         // [-front_pad, -front_pad + n_front): their columns stay close to the rows that use
                                         // them, so every rank keeps 16-bit column deltas (spmv.cu)
  std::vector<long long> recv_pos;       // [n_ranks] where the entries received from rank q start in a vector
  DevBuf<int32_t> colind_local;          // column indices in [owned | halo] numbering
  DevBuf<uint32_t> send_idx;             // owned entries to send, grouped by destination rank
  DevBuf<double> send_buf;
  std::vector<size_t> send_off, recv_off;   // [n_ranks + 1]
  DevBuf<double> red;                    // reduction scalars (ncclAllReduce in place)
  bool p2p_halo = false;                 // halo values travel through the peer arenas
  std::vector<size_t> land_off;          // [n_ranks]: where this rank's values land in the arena of rank q
};

}  // namespace tfelcu

struct tfelcu_solver {
  tfelcu::Ctx* ctx = nullptr;
  tfelcu_solver_params params;
  size_t n = 0, nnz = 0;
  const int32_t* rowptr = nullptr;
  const int32_t* colind = nullptr;
  const double* val = nullptr;
  tfelcu::DevBuf<int32_t> own_rowptr, own_colind;
  tfelcu::DevBuf<double> own_val;
  tfelcu::SpmvPlan plan;
  // work space
  tfelcu::DevBuf<double> x, b, r, p, q, z, dinv, partial;
  tfelcu::DevBuf<double> sdir;             // s = A p of the single-reduction CG
  tfelcu::DevBuf<tfelcu::KrylovScalars> scal;
  tfelcu::DevBuf<double> block_inv;        // block Jacobi: [n_blocks][bs][bs]
  tfelcu::DevBuf<double> basis, hess;      // GMRES
  std::unique_ptr<tfelcu::Ilu0> ilu;
  double t_setup_s = 0.0;
  int n_partial = 0;
  int gmres_m = 0;                         // restart length the allocated basis holds
  std::unique_ptr<tfelcu::DistPlan> dist;  // set when the operator holds only the rows this rank owns
  size_t n_global = 0;
  // identity of the pattern the structural data (SpMV plan, halo plan, ILU symbolic phase) were built for: a later
  // set_operator with the same pattern only refreshes the values (0: none)
  uint64_t pattern_id = 0;
  bool borrowed = false;                   // the arrays belong to a bilinear form (tfelcu_matrix_solve): not usable afterwards
  bool borrowed_live = false;              // ... and the call that borrowed them is still running
  double t_precond_s = 0.0, t_symbolic_s = 0.0;
  uint32_t restart_used = 0;
};

namespace tfelcu {

// structure == false: the pattern is the one of the previous call, only the values changed
void solver_setup(tfelcu_solver* s, bool structure = true);
void solver_spmv(tfelcu_solver* s, const double* d_x, double* d_y);
void solver_solve(tfelcu_solver* s, const double* d_rhs, double* d_x, tfelcu_solver_report* rep);
void solve_cg(tfelcu_solver* s, tfelcu_solver_report* rep);
void solve_gmres(tfelcu_solver* s, tfelcu_solver_report* rep);
void spmv_plan(tfelcu_solver* s);
// y = A x; when dot_partial != nullptr also writes per-CTA partial sums of x . y and returns their number
int spmv_launch(tfelcu_solver* s, const double* d_x, double* d_y, double* dot_partial, const KrylovScalars* scal);
int spmv_partial_count(tfelcu_solver* s);
int spmv_launch_part(tfelcu_solver* s, int part, const double* d_x, double* d_y, double* dot_partial, const KrylovScalars* scal);
void spmv_split_boundary(tfelcu_solver* s, size_t n_local);
void ilu0_setup(tfelcu_solver* s, bool symbolic = true);
void ilu0_apply(tfelcu_solver* s, const double* d_in, double* d_out);
// ---- iluk.cu
void iluk_factorize(tfelcu_solver* s, bool symbolic);
void iluk_apply(tfelcu_solver* s, const double* d_in, double* d_out);
void iluk_check(tfelcu_solver* s);
void block_jacobi_setup(tfelcu_solver* s);
void block_jacobi_apply(tfelcu_solver* s, const double* d_in, double* d_out);
// ---- dist.cu
void dist_setup(tfelcu_solver* s);
void dist_halo_exchange(tfelcu_solver* s, double* v, const KrylovScalars* S, cudaStream_t stream = nullptr, int mode = 3);
// epilogue == 1: the reducing kernel also performs the scalar step of the single-reduction CG (k_cgs_scalars)
const double* dist_reduce(tfelcu_solver* s, const double* const* partial, const int* n, int k, int out_slot, int epilogue = 0);
void solve_cg_single_reduction(tfelcu_solver* s, tfelcu_solver_report* rep);
void dist_gather(tfelcu_solver* s, const double* x_local, double* x_global);
bool dist_peer_failed(tfelcu_solver* s);
void p2p_setup(Ctx* ctx);      // collective over the communicator of ctx
void p2p_teardown(Ctx* ctx);
double p2p_bench_allreduce(Ctx* ctx, int reps);

}  // namespace tfelcu
