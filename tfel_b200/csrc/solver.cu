// solver::cuda::* -- Krylov solvers behind solver::basic_solver (src/core/solver.hpp:29-40).
// This file: operator set-up, the preconditioned conjugate gradient method (Jacobi / block Jacobi / none) and
// the fused vector kernels. GMRES + ILU(0) live in krylov.cu.
//
// Dirichlet rows of the reference are identity rows whose columns are NOT eliminated
// (bilinear_form.hpp:159-165,183-186), so A is not symmetric when boundary data are non-zero. CG starts from the
// lift x0 = b on identity rows, 0 elsewhere: r0 then vanishes on those rows, every search direction does too,
// and the iteration runs on the symmetric interior block A_II with right-hand side b_I - A_ID g.
//
// Convergence: ||r||_2 <= max(rtol ||b||_2, atol); divergence: ||r||_2 > dtol ||b||_2 (the dictionary keys of
// solver.hpp:125-133). All reductions are per-CTA partial sums combined in a fixed order: bitwise reproducible.
#include "solver.cuh"

#include <cstdlib>

namespace tfelcu {

constexpr int kVecThreads = 256;

static int vec_grid(tfelcu_solver* s) {
  // grid-stride kernels sized to the machine: 8 CTAs of 256 threads per SM, two entries per thread and pass
  size_t want = (s->n + 2 * kVecThreads - 1) / (2 * kVecThreads);
  const size_t cap = (size_t)s->ctx->sm_count * 8;
  if (want > cap) want = cap;
  return want < 1 ? 1 : (int)want;
}

__device__ __forceinline__ double block_sum(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < kVecThreads / 32; ++w) t += red[w];
  __syncthreads();
  return t;   // valid on thread 0
}

// single-thread sum of per-CTA partials (tiny scalar kernels)
__device__ __forceinline__ double sum_partials(const double* __restrict__ partial, int n) {
  double t = 0.0;
  for (int i = 0; i < n; ++i) t += partial[i];
  return t;
}

// the same sum formed by every CTA: warp 0 adds lane-strided slices and combines them with a butterfly, which
// gives every lane (and every CTA) the same bits; broadcast through shared memory
__device__ __forceinline__ double cta_sum_partials(const double* __restrict__ partial, int n, double* sh) {
  if (threadIdx.x < 32) {
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) t += partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) sh[0] = t;
  }
  __syncthreads();
  const double v = sh[0];
  __syncthreads();
  return v;
}

__global__ void k_extract_diag(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                               const double* __restrict__ val, size_t n, int invert, double* __restrict__ d) {
  const size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  double v = 1.0;
  for (int e = rowptr[r]; e < rowptr[r + 1]; ++e)
    if (colind[e] == (int32_t)r) v = val[e];
  if (invert) d[r] = v != 0.0 ? 1.0 / v : 1.0;   // a zero diagonal (pressure block of Stokes) is left unscaled
  else d[r] = v;
}

// x0: the Dirichlet lift (identity rows take their right-hand side), zero elsewhere
__global__ void k_initial_guess(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                                const double* __restrict__ val, const double* __restrict__ b, size_t n,
                                double* __restrict__ x) {
  const size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int rs = rowptr[r];
  const bool identity = (rowptr[r + 1] - rs == 1) && colind[rs] == (int32_t)r && val[rs] == 1.0;
  x[r] = identity ? b[r] : 0.0;
}

// r = b - q, partial sums of b.b
__global__ void __launch_bounds__(kVecThreads)
k_residual(const double* __restrict__ b, const double* __restrict__ q, size_t n, double* __restrict__ r,
           double* __restrict__ partial_bb) {
  __shared__ double red[kVecThreads / 32];
  double bb = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double bi = b[i];
    r[i] = bi - q[i];
    bb += bi * bi;
  }
  const double t = block_sum(bb, red);
  if (threadIdx.x == 0) partial_bb[blockIdx.x] = t;
}

// z = M^-1 r (Jacobi: dinv * r, none: r), p = z, partial sums of r.z and r.r
__global__ void __launch_bounds__(kVecThreads)
k_cg_start(const double* __restrict__ r, const double* __restrict__ dinv, const double* __restrict__ z_in, size_t n,
           double* __restrict__ p, double* __restrict__ partial_rz, double* __restrict__ partial_rr) {
  __shared__ double red[kVecThreads / 32];
  double rz = 0.0, rr = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double ri = r[i];
    const double zi = z_in ? z_in[i] : (dinv ? dinv[i] * ri : ri);
    p[i] = zi;
    rz += ri * zi; rr += ri * ri;
  }
  const double t0 = block_sum(rz, red);
  const double t1 = block_sum(rr, red);
  if (threadIdx.x == 0) { partial_rz[blockIdx.x] = t0; partial_rr[blockIdx.x] = t1; }
}

// one warp: global sums of b.b, r.z, r.r from their per-CTA partials, tolerances and the initial stopping test
__global__ void k_cg_set_scalars(const double* __restrict__ partial_bb, const double* __restrict__ partial_rz,
                                 const double* __restrict__ partial_rr, int n_partial, double rtol, double atol,
                                 double dtol, KrylovScalars* __restrict__ S) {
  if (blockIdx.x || threadIdx.x >= 32) return;
  const double* p[3] = {partial_bb, partial_rz, partial_rr};
  double v[3];
  for (int k = 0; k < 3; ++k) {
    double t = 0.0;
    for (int i = threadIdx.x; i < n_partial; i += 32) t += p[k][i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    v[k] = t;
  }
  if (threadIdx.x) return;
  const double bb = v[0];
  const double target = fmax(rtol * sqrt(bb), atol);
  S->bnorm2 = bb; S->target2 = target * target; S->dtol2 = dtol * dtol * bb;
  S->rz = v[1];
  S->rr = v[2];
  S->its = 0;
  S->done = (S->rr <= S->target2) ? 1 : 0;
}

// alpha = rz / pq; r -= alpha q; [Jacobi / none: partial sums of r.z] and of r.r. The companion update
// x += alpha p is done by k_cg_direction, which reads p anyway (two vector passes less per iteration).
__global__ void __launch_bounds__(kVecThreads)
k_cg_update(const double* __restrict__ q, const double* __restrict__ dinv,
            int fused_precond, size_t n, const double* __restrict__ partial_pq, int n_pq,
            double* __restrict__ r, double* __restrict__ partial_rz,
            double* __restrict__ partial_rr, KrylovScalars* __restrict__ S) {
  __shared__ double red[kVecThreads / 32];
  __shared__ double bc[1];
  if (S->done) return;
  const double pq = cta_sum_partials(partial_pq, n_pq, bc);
  const double alpha = S->rz / pq;
  double rz = 0.0, rr = 0.0;
  const size_t n2 = n / 2;
  const double2* q2 = reinterpret_cast<const double2*>(q);
  const double2* d2 = reinterpret_cast<const double2*>(dinv);
  double2* r2 = reinterpret_cast<double2*>(r);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    const double2 qv = q2[i];
    double2 rv = r2[i];
    rv.x -= alpha * qv.x; rv.y -= alpha * qv.y;
    r2[i] = rv;
    rr += rv.x * rv.x; rr += rv.y * rv.y;
    if (fused_precond) {
      if (dinv) { const double2 dv = d2[i]; rz += rv.x * (dv.x * rv.x); rz += rv.y * (dv.y * rv.y); }
      else { rz += rv.x * rv.x; rz += rv.y * rv.y; }
    }
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const size_t i = n - 1;
    const double ri = r[i] - alpha * q[i];
    r[i] = ri;
    rr += ri * ri;
    if (fused_precond) rz += ri * (dinv ? dinv[i] * ri : ri);
  }
  const double t0 = block_sum(rz, red);
  const double t1 = block_sum(rr, red);
  if (threadIdx.x == 0) { partial_rz[blockIdx.x] = t0; partial_rr[blockIdx.x] = t1; }
  if (blockIdx.x == 0 && threadIdx.x == 0) { S->pq = pq; S->alpha = alpha; }
}

// r.z for a preconditioner that is applied by its own kernel
__global__ void __launch_bounds__(kVecThreads)
k_dot_partial(const double* __restrict__ a, const double* __restrict__ b, size_t n, double* __restrict__ partial,
              const KrylovScalars* __restrict__ S) {
  __shared__ double red[kVecThreads / 32];
  if (S && S->done) return;
  double v = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) v += a[i] * b[i];
  const double t = block_sum(v, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// x += alpha p (alpha of this iteration, published by k_cg_update); beta = rz_new / rz; p = z + beta p; the new
// scalars are published by a separate tiny kernel (k_cg_commit) after every CTA has read the old ones
__global__ void __launch_bounds__(kVecThreads)
k_cg_direction(const double* __restrict__ r, const double* __restrict__ dinv, const double* __restrict__ z_in, size_t n,
               const double* __restrict__ partial_rz, int n_partial,
               double* __restrict__ p, double* __restrict__ x, KrylovScalars* __restrict__ S) {
  __shared__ double bc[1];
  if (S->done) return;
  const double rz_new = cta_sum_partials(partial_rz, n_partial, bc);
  const double rz_old = S->rz;
  const double beta = rz_new / rz_old;
  const double alpha = S->alpha;
  double2* x2 = reinterpret_cast<double2*>(x);
  const size_t n2 = n / 2;
  const double2* r2 = reinterpret_cast<const double2*>(r);
  const double2* d2 = reinterpret_cast<const double2*>(dinv);
  const double2* z2 = reinterpret_cast<const double2*>(z_in);
  double2* p2 = reinterpret_cast<double2*>(p);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    double2 zv;
    if (z_in) zv = z2[i];
    else { zv = r2[i]; if (dinv) { const double2 dv = d2[i]; zv.x *= dv.x; zv.y *= dv.y; } }
    double2 pv = p2[i];
    double2 xv = x2[i];
    xv.x += alpha * pv.x; xv.y += alpha * pv.y;
    x2[i] = xv;
    pv.x = zv.x + beta * pv.x; pv.y = zv.y + beta * pv.y;
    p2[i] = pv;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const size_t i = n - 1;
    const double zi = z_in ? z_in[i] : (dinv ? dinv[i] * r[i] : r[i]);
    x[i] += alpha * p[i];
    p[i] = zi + beta * p[i];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    S->rz_new = rz_new; S->beta = beta;
  }
}

// one warp: r.r from its partial sums, then the scalars of the next iteration and the stopping test
__global__ void k_cg_commit(const double* __restrict__ partial_rr, int n_partial, KrylovScalars* __restrict__ S) {
  if (blockIdx.x || threadIdx.x >= 32 || S->done) return;
  double t = 0.0;
  for (int i = threadIdx.x; i < n_partial; i += 32) t += partial_rr[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (threadIdx.x) return;
  S->rz = S->rz_new; S->rr = t;
  S->its += 1;
  if (!(S->rr == S->rr) || !(S->pq == S->pq) || S->pq == 0.0) S->done = -2;   // NaN / breakdown
  else if (S->rr <= S->target2) S->done = 1;
  else if (S->rr > S->dtol2) S->done = -1;
}

// ------------------------------------------------------------------------------------------
// block Jacobi: contiguous blocks of bs rows, dense inverse of each diagonal block
// ------------------------------------------------------------------------------------------

__global__ void k_block_jacobi_setup(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                                     const double* __restrict__ val, size_t n, int bs, double* __restrict__ inv) {
  const size_t blk = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t r0 = blk * bs;
  if (r0 >= n) return;
  double a[8][16];   // [bs][2 bs]: Gauss-Jordan on [D | I]
  for (int i = 0; i < bs; ++i)
    for (int j = 0; j < 2 * bs; ++j) a[i][j] = (j == bs + i) ? 1.0 : 0.0;
  for (int i = 0; i < bs; ++i) {
    const size_t r = r0 + i;
    if (r >= n) { a[i][i] = 1.0; continue; }
    for (int e = rowptr[r]; e < rowptr[r + 1]; ++e) {
      const long long c = (long long)colind[e] - (long long)r0;
      if (c >= 0 && c < bs && (size_t)colind[e] < n) a[i][c] = val[e];   // halo columns (>= n) are not in a block
    }
  }
  for (int k = 0; k < bs; ++k) {
    int piv = k; double best = fabs(a[k][k]);
    for (int i = k + 1; i < bs; ++i) if (fabs(a[i][k]) > best) { best = fabs(a[i][k]); piv = i; }
    if (piv != k) for (int j = 0; j < 2 * bs; ++j) { const double t = a[k][j]; a[k][j] = a[piv][j]; a[piv][j] = t; }
    const double d = a[k][k] != 0.0 ? 1.0 / a[k][k] : 1.0;
    for (int j = 0; j < 2 * bs; ++j) a[k][j] *= d;
    for (int i = 0; i < bs; ++i) {
      if (i == k) continue;
      const double f = a[i][k];
      for (int j = 0; j < 2 * bs; ++j) a[i][j] -= f * a[k][j];
    }
  }
  for (int i = 0; i < bs; ++i)
    for (int j = 0; j < bs; ++j) inv[(blk * bs + i) * bs + j] = a[i][bs + j];
}

__global__ void k_block_jacobi_apply(const double* __restrict__ inv, const double* __restrict__ in, size_t n, int bs,
                                     double* __restrict__ out, const KrylovScalars* __restrict__ S) {
  if (S && S->done) return;
  const size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  const size_t blk = r / bs; const int i = (int)(r % bs);
  double acc = 0.0;
  for (int j = 0; j < bs; ++j) {
    const size_t c = blk * bs + j;
    if (c < n) acc += inv[(blk * bs + i) * bs + j] * in[c];
  }
  out[r] = acc;
}

void block_jacobi_setup(tfelcu_solver* s) {
  Ctx* ctx = s->ctx;
  int bs = s->params.block_size;
  if (bs <= 0) bs = 4;
  TFELCU_REQUIRE(bs <= 8, "block Jacobi supports blocks of up to 8 rows");
  s->params.block_size = bs;
  const size_t nb = (s->n + bs - 1) / bs;
  s->block_inv.alloc(nb * bs * bs);
  k_block_jacobi_setup<<<grid_for(nb, 128), 128, 0, ctx->stream>>>(s->rowptr, s->colind, s->val, s->n, bs, s->block_inv.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
}

void block_jacobi_apply(tfelcu_solver* s, const double* d_in, double* d_out) {
  Ctx* ctx = s->ctx;
  k_block_jacobi_apply<<<grid_for(s->n, 256), 256, 0, ctx->stream>>>(s->block_inv.p, d_in, s->n, s->params.block_size,
                                                                   d_out, s->scal.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------
// set-up and the CG driver
// ------------------------------------------------------------------------------------------

void solver_setup(tfelcu_solver* s, bool structure) {
  Ctx* ctx = s->ctx;
  cudaEvent_t e0, e1, e2;
  TFELCU_CK(cudaEventCreate(&e0)); TFELCU_CK(cudaEventCreate(&e1)); TFELCU_CK(cudaEventCreate(&e2));
  TFELCU_CK(cudaEventRecord(e0, ctx->stream));
  const size_t n = s->n;
  size_t n_halo = 0, front = 0;
  if (structure) {
    if (s->dist) {
      dist_setup(s);                       // needs the global column indices
      s->colind = s->dist->colind_local.p;
      n_halo = s->dist->n_halo; front = s->dist->front_pad;
    }
    // vectors the product reads carry the halo: [lower ranks' entries | owned (p points here) | pad, upper ranks' entries]
    s->x.alloc(n + n_halo, front); s->b.alloc(n); s->r.alloc(n); s->p.alloc(n + n_halo, front); s->q.alloc(n);
    ctx->zero(s->x.p - front, front + n + n_halo); ctx->zero(s->p.p - front, front + n + n_halo);
    s->scal.alloc(1);
    ctx->zero(s->scal.p, 1);
    spmv_plan(s);
    if (s->dist) spmv_split_boundary(s, n);   // interior slabs first: they overlap the halo exchange
    int np = spmv_partial_count(s);
    const int vg = vec_grid(s);
    if (vg > np) np = vg;
    s->n_partial = np;
    s->partial.alloc((size_t)np * 4);
    ctx->zero(s->partial.p, (size_t)np * 4);
  } else if (s->dist) {
    s->colind = s->dist->colind_local.p;   // the caller re-attached the global columns; the local ones are unchanged
  }
  TFELCU_CK(cudaEventRecord(e1, ctx->stream));
  const int pc = s->params.precond;
  if (pc == TFELCU_PC_JACOBI) {
    if (structure) s->dinv.alloc(n);
    k_extract_diag<<<grid_for(n, 256), 256, 0, ctx->stream>>>(s->rowptr, s->colind, s->val, n, 1, s->dinv.p);
    ctx->count(); TFELCU_LAUNCH_CHECK();
  } else if (pc == TFELCU_PC_BLOCK_JACOBI) {
    if (structure) s->z.alloc(n);
    block_jacobi_setup(s);
  } else if (pc == TFELCU_PC_ILU) {
    if (structure) s->z.alloc(n);
    ilu0_setup(s, structure);
  }
  TFELCU_CK(cudaEventRecord(e2, ctx->stream));
  TFELCU_CK(cudaEventSynchronize(e2));
  float ms = 0.f; TFELCU_CK(cudaEventElapsedTime(&ms, e0, e2));
  s->t_setup_s = ms * 1e-3;
  TFELCU_CK(cudaEventElapsedTime(&ms, e1, e2));
  s->t_precond_s = ms * 1e-3;
  s->t_symbolic_s = (pc == TFELCU_PC_ILU && s->ilu && structure) ? s->ilu->t_symbolic_s : 0.0;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
}

static void apply_precond(tfelcu_solver* s, const double* d_in, double* d_out) {
  if (s->params.precond == TFELCU_PC_BLOCK_JACOBI) block_jacobi_apply(s, d_in, d_out);
  else if (s->params.precond == TFELCU_PC_ILU) ilu0_apply(s, d_in, d_out);
  else fail("preconditioner has no stand-alone apply");
}

// ------------------------------------------------------------------------------------------
// Single-reduction CG (Chronopoulos & Gear): the same Krylov iterates as the standard method in exact arithmetic with
// ONE global reduction per iteration (gamma = r.u, r.r, delta = u.Au together) and one fused vector kernel, for the
// row-distributed solve where every synchronisation point costs more than the extra two vector passes:
//     w = A u ; (gamma, rr, delta) ; beta = gamma / gamma_old ; alpha = gamma / (delta - beta gamma / alpha_old)
//     p = u + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s ; u = M^-1 r          (M = diag(A) or I)
// Per iteration: halo exchange of u, SpMV (+ partials of u.Au), reduction (+ scalar step), fused update = 4 launches.
// ------------------------------------------------------------------------------------------

// u = M^-1 r, p = s = 0, partial sums of r.u and r.r
__global__ void __launch_bounds__(kVecThreads)
k_cgs_start(const double* __restrict__ r, const double* __restrict__ dinv, size_t n, double* __restrict__ u,
            double* __restrict__ p, double* __restrict__ sd, double* __restrict__ partial_g, double* __restrict__ partial_rr) {
  __shared__ double red[kVecThreads / 32];
  double g = 0.0, rr = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double ri = r[i];
    const double ui = dinv ? dinv[i] * ri : ri;
    u[i] = ui; p[i] = 0.0; sd[i] = 0.0;
    g += ri * ui; rr += ri * ri;
  }
  const double t0 = block_sum(g, red);
  const double t1 = block_sum(rr, red);
  if (threadIdx.x == 0) { partial_g[blockIdx.x] = t0; partial_rr[blockIdx.x] = t1; }
}

__global__ void k_cgs_init_scalars(const double* __restrict__ partial_bb, int n_partial, double rtol, double atol, double dtol,
                                   double maxits, KrylovScalars* __restrict__ S) {
  if (threadIdx.x || blockIdx.x) return;
  const double bb = sum_partials(partial_bb, n_partial);
  const double target = fmax(rtol * sqrt(bb), atol);
  S->bnorm2 = bb; S->target2 = target * target; S->dtol2 = dtol * dtol * bb;
  S->spare1 = maxits;
  S->its = 0; S->done = 0;
}

// single process: the three sums come as per-CTA partials
__global__ void k_cgs_scalars(const double* __restrict__ pg, int ng, const double* __restrict__ prr, int nrr,
                              const double* __restrict__ pd, int nd, KrylovScalars* __restrict__ S) {
  if (blockIdx.x || threadIdx.x >= 32) return;
  double v[3];
  const double* p[3] = {pg, prr, pd};
  const int n[3] = {ng, nrr, nd};
  for (int k = 0; k < 3; ++k) {
    double t = 0.0;
    for (int i = threadIdx.x; i < n[k]; i += 32) t += p[k][i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    v[k] = t;
  }
  if (threadIdx.x == 0) cgs_scalar_step(S, v[0], v[1], v[2]);
}

__global__ void __launch_bounds__(kVecThreads)
k_cgs_fused(const double* __restrict__ w, const double* __restrict__ dinv, size_t n, double* __restrict__ u,
            double* __restrict__ p, double* __restrict__ sd, double* __restrict__ x, double* __restrict__ r,
            double* __restrict__ partial_g, double* __restrict__ partial_rr, const KrylovScalars* __restrict__ S) {
  __shared__ double red[kVecThreads / 32];
  if (S->done) return;
  const double alpha = S->alpha, beta = S->beta;
  double g = 0.0, rr = 0.0;
  const size_t n2 = n / 2;
  const double2* w2 = reinterpret_cast<const double2*>(w);
  const double2* d2 = reinterpret_cast<const double2*>(dinv);
  double2* u2 = reinterpret_cast<double2*>(u);
  double2* p2 = reinterpret_cast<double2*>(p);
  double2* s2 = reinterpret_cast<double2*>(sd);
  double2* x2 = reinterpret_cast<double2*>(x);
  double2* r2 = reinterpret_cast<double2*>(r);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    double2 uv = u2[i], pv = p2[i], sv = s2[i], xv = x2[i], rv = r2[i];
    const double2 wv = w2[i];
    pv.x = uv.x + beta * pv.x; pv.y = uv.y + beta * pv.y;
    sv.x = wv.x + beta * sv.x; sv.y = wv.y + beta * sv.y;
    xv.x += alpha * pv.x; xv.y += alpha * pv.y;
    rv.x -= alpha * sv.x; rv.y -= alpha * sv.y;
    if (dinv) { const double2 dv = d2[i]; uv.x = dv.x * rv.x; uv.y = dv.y * rv.y; } else uv = rv;
    p2[i] = pv; s2[i] = sv; x2[i] = xv; r2[i] = rv; u2[i] = uv;
    g += rv.x * uv.x; g += rv.y * uv.y;
    rr += rv.x * rv.x; rr += rv.y * rv.y;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const size_t i = n - 1;
    const double pi = u[i] + beta * p[i], si = w[i] + beta * sd[i];
    p[i] = pi; sd[i] = si; x[i] += alpha * pi;
    const double ri = r[i] - alpha * si;
    r[i] = ri;
    const double ui = dinv ? dinv[i] * ri : ri;
    u[i] = ui;
    g += ri * ui; rr += ri * ri;
  }
  const double t0 = block_sum(g, red);
  const double t1 = block_sum(rr, red);
  if (threadIdx.x == 0) { partial_g[blockIdx.x] = t0; partial_rr[blockIdx.x] = t1; }
}

void solve_cg_single_reduction(tfelcu_solver* s, tfelcu_solver_report* rep) {
  Ctx* ctx = s->ctx;
  const size_t n = s->n;
  const int vg = vec_grid(s);
  const int pc = s->params.precond;
  const double* dinv = pc == TFELCU_PC_JACOBI ? s->dinv.p : nullptr;
  const bool dist = (bool)s->dist;
  const size_t n_halo = dist ? s->dist->n_halo : 0;
  const size_t front = dist ? s->dist->front_pad : 0;
  if (s->z.n < n + n_halo || s->z.front != front) s->z.alloc(n + n_halo, front);
  if (s->sdir.n < n) s->sdir.alloc(n);
  double* u = s->z.p; double* w = s->q.p; double* sd = s->sdir.p;
  ctx->zero(u - front, front + n + n_halo);
  double* P_d = s->partial.p;
  double* P_g = s->partial.p + s->n_partial;
  double* P_rr = s->partial.p + 2 * (size_t)s->n_partial;
  double* P_bb = s->partial.p + 3 * (size_t)s->n_partial;
  KrylovScalars* S = s->scal.p;
  uint64_t n_spmv = 0;
  const uint64_t launches0 = ctx->launches;

  ctx->zero(S, 1);
  k_initial_guess<<<grid_for(n, 256), 256, 0, ctx->stream>>>(s->rowptr, s->colind, s->val, s->b.p, n, s->x.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  if (dist) dist_halo_exchange(s, s->x.p, nullptr);
  spmv_launch(s, s->x.p, w, nullptr, nullptr); ++n_spmv;
  k_residual<<<vg, kVecThreads, 0, ctx->stream>>>(s->b.p, w, n, s->r.p, P_bb);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  {
    const double* bbp = P_bb; int nbb = vg;
    if (dist) { const double* ptrs[1] = {P_bb}; const int ns[1] = {vg}; bbp = dist_reduce(s, ptrs, ns, 1, 4); nbb = 1; }
    k_cgs_init_scalars<<<1, 32, 0, ctx->stream>>>(bbp, nbb, s->params.rtol, s->params.atol, s->params.dtol,
                                                  (double)s->params.maxits, S);
    ctx->count(); TFELCU_LAUNCH_CHECK();
  }
  k_cgs_start<<<vg, kVecThreads, 0, ctx->stream>>>(s->r.p, dinv, n, u, s->p.p, sd, P_g, P_rr);
  ctx->count(); TFELCU_LAUNCH_CHECK();

  const int prof_every = s->params.profile_every;
  std::vector<cudaEvent_t> prof;
  std::vector<uint32_t> prof_it;
  uint32_t last_sample = 0;
  // TFELCU_TRACE=1: time every launch of the eager iteration of each chunk (diagnostic, printed by rank 0)
  const bool trace = std::getenv("TFELCU_TRACE") != nullptr;
  std::vector<cudaEvent_t> tr;
  auto mark = [&](bool on) {
    if (!on) return;
    cudaEvent_t e; TFELCU_CK(cudaEventCreate(&e)); TFELCU_CK(cudaEventRecord(e, ctx->stream)); tr.push_back(e);
  };
  // distributed, optional (TFELCU_OVERLAP=1): the halo exchange runs on a side stream while the interior slabs (owned
  // columns only) are multiplied; the boundary slabs follow once the halo has landed. Measured at n = 4096: 8 GPUs
  // 0.122 vs 0.129 ms per iteration, 2 GPUs 0.362 vs 0.313 ms (the extra launch and the cross-stream joins cost more
  // than the hidden latency), so it is off by default.
  const bool overlap = dist && s->plan.n_interior > 0 && s->plan.n_interior < s->plan.n_blocks && std::getenv("TFELCU_OVERLAP");
  if (overlap && !ctx->side_stream) {
    TFELCU_CK(cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
    TFELCU_CK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    TFELCU_CK(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  }
  double* P_d2 = s->partial.p + 3 * (size_t)s->n_partial;   // partials of the boundary slabs (P_bb is free by then)
  auto iteration = [&](uint32_t it, bool may_sample) {
    const bool t_on = trace && may_sample && tr.size() < 5000;
    mark(t_on);
    if (overlap) {
      TFELCU_CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
      TFELCU_CK(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
      dist_halo_exchange(s, u, S, ctx->side_stream);
      TFELCU_CK(cudaEventRecord(ctx->ev_join, ctx->side_stream));
    } else if (dist) dist_halo_exchange(s, u, S);
    mark(t_on);
    const bool sampled = may_sample && prof_every > 0 && (it == 0 || it >= last_sample + (uint32_t)prof_every) && prof.size() < 1024;
    if (sampled) {
      cudaEvent_t a, b;
      TFELCU_CK(cudaEventCreate(&a)); TFELCU_CK(cudaEventCreate(&b));
      prof.push_back(a); prof.push_back(b); prof_it.push_back(it); last_sample = it;
      TFELCU_CK(cudaEventRecord(a, ctx->stream));
    }
    int n_d, n_d2 = 0;
    if (overlap) {
      n_d = spmv_launch_part(s, 1, u, w, P_d, S);
      TFELCU_CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
      n_d2 = spmv_launch_part(s, 2, u, w, P_d2, S);
    } else n_d = spmv_launch(s, u, w, P_d, S);
    ++n_spmv;
    if (sampled) TFELCU_CK(cudaEventRecord(prof.back(), ctx->stream));
    mark(t_on);
    if (dist && n_d2 > 0) {
      const double* ptrs[4] = {P_g, P_rr, P_d, P_d2}; const int ns[4] = {vg, vg, n_d, n_d2};
      dist_reduce(s, ptrs, ns, 4, 0, 1);
    } else if (dist) {
      const double* ptrs[3] = {P_g, P_rr, P_d}; const int ns[3] = {vg, vg, n_d};
      dist_reduce(s, ptrs, ns, 3, 0, 1);
    } else {
      k_cgs_scalars<<<1, 32, 0, ctx->stream>>>(P_g, vg, P_rr, vg, P_d, n_d, S);
      ctx->count(); TFELCU_LAUNCH_CHECK();
    }
    mark(t_on);
    k_cgs_fused<<<vg, kVecThreads, 0, ctx->stream>>>(w, dinv, n, u, s->p.p, sd, s->x.p, s->r.p, P_g, P_rr, S);
    ctx->count(); TFELCU_LAUNCH_CHECK();
    mark(t_on);
  };

  int check = s->params.check_every > 0 ? s->params.check_every : 50;
  KrylovScalars h;
  ctx->download(&h, S, 1);
  uint32_t issued = 0;
  const uint32_t maxits = s->params.maxits;
  const bool graph_ok = !std::getenv("TFELCU_NO_GRAPH") && (!dist || ctx->p2p) && check >= 8 && maxits >= (uint32_t)(2 * check);
  cudaGraphExec_t graph_exec = nullptr;
  uint64_t graph_launches = 0, graph_spmv = 0;
  if (graph_ok) {
    cudaGraph_t graph = nullptr;
    const uint64_t l0 = ctx->launches, s0 = n_spmv;
    TFELCU_CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    for (int it = 0; it < check - 1; ++it) iteration(0, false);
    TFELCU_CK(cudaStreamEndCapture(ctx->stream, &graph));
    graph_launches = ctx->launches - l0; graph_spmv = n_spmv - s0;
    ctx->launches = l0; n_spmv = s0;
    TFELCU_CK(cudaGraphInstantiate(&graph_exec, graph, 0));
    TFELCU_CK(cudaGraphDestroy(graph));
  }
  // one extra pass detects convergence of the last update (the test runs at the start of an iteration)
  while (!h.done && issued < maxits + 1) {
    uint32_t chunk = (uint32_t)check;
    if (issued + chunk > maxits + 1) chunk = maxits + 1 - issued;
    if (graph_exec && chunk == (uint32_t)check) {
      TFELCU_CK(cudaGraphLaunch(graph_exec, ctx->stream));
      ctx->count(graph_launches); n_spmv += graph_spmv;
      iteration(issued + chunk - 1, true);
    } else {
      for (uint32_t it = 0; it < chunk; ++it) iteration(issued + it, true);
    }
    issued += chunk;
    ctx->download(&h, S, 1);
  }
  if (graph_exec) TFELCU_CK(cudaGraphExecDestroy(graph_exec));
  if (!tr.empty()) {
    double acc[4] = {0, 0, 0, 0}; int cnt = 0;
    for (size_t e = 0; e + 4 < tr.size(); e += 5, ++cnt)
      for (int k = 0; k < 4; ++k) { float ms = 0.f; cudaEventElapsedTime(&ms, tr[e + k], tr[e + k + 1]); acc[k] += ms; }
    if (cnt)
      std::fprintf(stderr, "[tfelcu trace] rank %d eager iteration, avg of %d: halo %.1f us, spmv %.1f us, reduce %.1f us, fused %.1f us\n",
                   ctx->rank, cnt, 1e3 * acc[0] / cnt, 1e3 * acc[1] / cnt, 1e3 * acc[2] / cnt, 1e3 * acc[3] / cnt);
    for (cudaEvent_t e : tr) cudaEventDestroy(e);
  }
  if (!prof.empty()) {
    double total = 0.0; uint32_t cnt = 0;
    for (size_t e = 0; e + 1 < prof.size(); e += 2) {
      float ms = 0.f;
      TFELCU_CK(cudaEventElapsedTime(&ms, prof[e], prof[e + 1]));
      if (prof_it[e / 2] < (uint32_t)h.its) { total += ms; ++cnt; }
      cudaEventDestroy(prof[e]); cudaEventDestroy(prof[e + 1]);
    }
    rep->spmv_avg_ms = cnt ? total / cnt : 0.0;
    rep->spmv_samples = cnt;
  }
  rep->its = (uint32_t)h.its;
  rep->converged = h.done == 1 ? 1 : ((h.done == 0 || h.done == 2) ? 0 : h.done);
  rep->rnorm = sqrt(h.rr);
  rep->bnorm = sqrt(h.bnorm2);
  rep->n_spmv = n_spmv;
  rep->n_launches = ctx->launches - launches0;
}

void solve_cg(tfelcu_solver* s, tfelcu_solver_report* rep) {
  Ctx* ctx = s->ctx;
  const size_t n = s->n;
  const int vg = vec_grid(s);
  const int pc = s->params.precond;
  const bool fused = (pc == TFELCU_PC_JACOBI || pc == TFELCU_PC_NONE);
  const double* dinv = pc == TFELCU_PC_JACOBI ? s->dinv.p : nullptr;
  double* P_pq = s->partial.p;
  double* P_rz = s->partial.p + s->n_partial;
  double* P_rr = s->partial.p + 2 * (size_t)s->n_partial;
  double* P_bb = s->partial.p + 3 * (size_t)s->n_partial;
  KrylovScalars* S = s->scal.p;
  uint64_t n_spmv = 0;
  const uint64_t launches0 = ctx->launches;

  const bool dist = (bool)s->dist;
  struct Sum { const double* p; int n; };
  auto global_sums = [&](std::initializer_list<Sum> in, int slot, Sum* out) {
    // single process: consumers add the per-CTA partials themselves; distributed: one value per sum after an allreduce
    int k = 0;
    const double* ptrs[4]; int ns[4];
    for (const Sum& v : in) { ptrs[k] = v.p; ns[k] = v.n; out[k] = v; ++k; }
    if (!dist) return;
    const double* r = dist_reduce(s, ptrs, ns, k, slot);
    for (int i = 0; i < k; ++i) out[i] = Sum{r + i, 1};
  };

  ctx->zero(S, 1);
  k_initial_guess<<<grid_for(n, 256), 256, 0, ctx->stream>>>(s->rowptr, s->colind, s->val, s->b.p, n, s->x.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  if (dist) dist_halo_exchange(s, s->x.p, nullptr);
  spmv_launch(s, s->x.p, s->q.p, nullptr, nullptr); ++n_spmv;
  k_residual<<<vg, kVecThreads, 0, ctx->stream>>>(s->b.p, s->q.p, n, s->r.p, P_bb);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  if (!fused) apply_precond(s, s->r.p, s->z.p);
  k_cg_start<<<vg, kVecThreads, 0, ctx->stream>>>(s->r.p, dinv, fused ? nullptr : s->z.p, n, s->p.p, P_rz, P_rr);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  {
    Sum g[3];
    global_sums({Sum{P_bb, vg}, Sum{P_rz, vg}, Sum{P_rr, vg}}, 4, g);
    k_cg_set_scalars<<<1, 32, 0, ctx->stream>>>(g[0].p, g[1].p, g[2].p, g[0].n, s->params.rtol, s->params.atol, s->params.dtol, S);
    ctx->count(); TFELCU_LAUNCH_CHECK();
  }

  // optional SpMV sampling for the roofline report: events around every profile_every-th launch
  const int prof_every = s->params.profile_every;
  std::vector<cudaEvent_t> prof;
  std::vector<uint32_t> prof_it;
  uint32_t last_sample = 0;
  auto sample_begin = [&](uint32_t it) -> bool {
    if (prof_every <= 0 || (it != 0 && it < last_sample + (uint32_t)prof_every) || prof.size() >= 1024) return false;
    last_sample = it; prof_it.push_back(it);
    cudaEvent_t a, b;
    TFELCU_CK(cudaEventCreate(&a)); TFELCU_CK(cudaEventCreate(&b));
    prof.push_back(a); prof.push_back(b);
    TFELCU_CK(cudaEventRecord(a, ctx->stream));
    return true;
  };
  int check = s->params.check_every > 0 ? s->params.check_every : 50;
  KrylovScalars h;
  ctx->download(&h, S, 1);
  uint32_t issued = 0;
  const uint32_t maxits = s->params.maxits;

  // one CG iteration, enqueued on the context stream; every kernel returns at once when S->done is set
  auto iteration = [&](uint32_t it, bool may_sample) {
    if (dist) dist_halo_exchange(s, s->p.p, S);
    const bool sampled = may_sample && sample_begin(it);
    const int n_pq = spmv_launch(s, s->p.p, s->q.p, P_pq, S); ++n_spmv;
    if (sampled) TFELCU_CK(cudaEventRecord(prof.back(), ctx->stream));
    Sum pq[1];
    global_sums({Sum{P_pq, n_pq}}, 0, pq);
    k_cg_update<<<vg, kVecThreads, 0, ctx->stream>>>(s->q.p, dinv, fused ? 1 : 0, n, pq[0].p, pq[0].n, s->r.p, P_rz, P_rr, S);
    ctx->count(); TFELCU_LAUNCH_CHECK();
    if (!fused) {
      apply_precond(s, s->r.p, s->z.p);
      k_dot_partial<<<vg, kVecThreads, 0, ctx->stream>>>(s->r.p, s->z.p, n, P_rz, S);
      ctx->count(); TFELCU_LAUNCH_CHECK();
    }
    Sum g[2];
    global_sums({Sum{P_rz, vg}, Sum{P_rr, vg}}, 1, g);
    k_cg_direction<<<vg, kVecThreads, 0, ctx->stream>>>(s->r.p, dinv, fused ? nullptr : s->z.p, n, g[0].p, g[0].n, s->p.p,
                                                       s->x.p, S);
    ctx->count(); TFELCU_LAUNCH_CHECK();
    k_cg_commit<<<1, 32, 0, ctx->stream>>>(g[1].p, g[1].n, S);
    ctx->count(); TFELCU_LAUNCH_CHECK();
  };

  // The loop is launch-bound once a rank's share of the matrix is small (8 GPUs: ~60 us of HBM work per iteration):
  // `check - 1` iterations are captured once in a CUDA graph and replayed; one eager iteration per chunk carries the
  // optional SpMV timing events. NCCL fallbacks and ILU(0) level launches stay eager.
  const bool graph_ok = !std::getenv("TFELCU_NO_GRAPH") && pc != TFELCU_PC_ILU && (!dist || ctx->p2p) && check >= 8 &&
                        maxits >= (uint32_t)(2 * check);
  cudaGraphExec_t graph_exec = nullptr;
  uint64_t graph_launches = 0, graph_spmv = 0;
  if (graph_ok && !h.done) {
    cudaGraph_t graph = nullptr;
    const uint64_t l0 = ctx->launches, s0 = n_spmv;
    TFELCU_CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    for (int it = 0; it < check - 1; ++it) iteration(0, false);
    TFELCU_CK(cudaStreamEndCapture(ctx->stream, &graph));
    graph_launches = ctx->launches - l0; graph_spmv = n_spmv - s0;
    ctx->launches = l0; n_spmv = s0;   // captured, not launched yet
    TFELCU_CK(cudaGraphInstantiate(&graph_exec, graph, 0));
    TFELCU_CK(cudaGraphDestroy(graph));
  }
  while (!h.done && issued < maxits) {
    uint32_t chunk = (uint32_t)check;
    if (issued + chunk > maxits) chunk = maxits - issued;
    if (graph_exec && chunk == (uint32_t)check) {
      TFELCU_CK(cudaGraphLaunch(graph_exec, ctx->stream));
      ctx->count(graph_launches); n_spmv += graph_spmv;
      iteration(issued + chunk - 1, true);
    } else {
      for (uint32_t it = 0; it < chunk; ++it) iteration(issued + it, true);
    }
    issued += chunk;
    ctx->download(&h, S, 1);
  }
  if (graph_exec) TFELCU_CK(cudaGraphExecDestroy(graph_exec));
  if (!prof.empty()) {
    // samples taken after convergence (the kernels return at once) are discarded
    double total = 0.0; uint32_t cnt = 0;
    for (size_t e = 0; e + 1 < prof.size(); e += 2) {
      float ms = 0.f;
      TFELCU_CK(cudaEventElapsedTime(&ms, prof[e], prof[e + 1]));
      if (prof_it[e / 2] < (uint32_t)h.its) { total += ms; ++cnt; }
      cudaEventDestroy(prof[e]); cudaEventDestroy(prof[e + 1]);
    }
    rep->spmv_avg_ms = cnt ? total / cnt : 0.0;
    rep->spmv_samples = cnt;
  }
  if (pc == TFELCU_PC_ILU) iluk_check(s);
  rep->its = (uint32_t)h.its;
  rep->converged = h.done == 1 ? 1 : (h.done == 0 ? 0 : h.done);
  rep->rnorm = sqrt(h.rr);
  rep->bnorm = sqrt(h.bnorm2);
  rep->n_spmv = n_spmv;
  rep->n_launches = ctx->launches - launches0;
}

void solver_solve(tfelcu_solver* s, const double* d_rhs, double* d_x, tfelcu_solver_report* rep) {
  Ctx* ctx = s->ctx;
  TFELCU_REQUIRE(s->rowptr != nullptr, "solver has no operator: call set_operator first");
  TFELCU_REQUIRE(!(s->dist && ctx->p2p_failed), "a peer rank did not answer during an earlier distributed solve: recreate the contexts");
  TFELCU_REQUIRE(!s->borrowed || s->borrowed_live, "the solver only borrowed the arrays of a bilinear form during tfelcu_matrix_solve: call set_operator");
  std::memset(rep, 0, sizeof(*rep));
  cudaEvent_t e0, e1;
  TFELCU_CK(cudaEventCreate(&e0)); TFELCU_CK(cudaEventCreate(&e1));
  TFELCU_CK(cudaEventRecord(e0, ctx->stream));
  TFELCU_CK(cudaMemcpyAsync(s->b.p, d_rhs, s->n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  // row-distributed solves with a pointwise preconditioner use the single-reduction recurrences (fewer launches and
  // reductions per iteration); TFELCU_CG=standard / single forces either
  const char* cg_env = std::getenv("TFELCU_CG");
  const bool pointwise = s->params.precond == TFELCU_PC_JACOBI || s->params.precond == TFELCU_PC_NONE;
  const bool single = pointwise && (cg_env ? std::string(cg_env) == "single" : (bool)s->dist);
  if (s->params.method == TFELCU_METHOD_CG && single) solve_cg_single_reduction(s, rep);
  else if (s->params.method == TFELCU_METHOD_CG) solve_cg(s, rep);
  else if (s->params.method == TFELCU_METHOD_GMRES) solve_gmres(s, rep);
  else fail("unknown Krylov method");
  TFELCU_CK(cudaMemcpyAsync(d_x, s->x.p, s->n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  TFELCU_CK(cudaEventRecord(e1, ctx->stream));
  TFELCU_CK(cudaEventSynchronize(e1));
  float ms = 0.f; TFELCU_CK(cudaEventElapsedTime(&ms, e0, e1));
  if (s->dist && dist_peer_failed(s)) rep->converged = -3;   // also for GMRES, whose device state has no field for it
  rep->t_solve_s = ms * 1e-3;
  rep->t_setup_s = s->t_setup_s;
  rep->t_precond_setup_s = s->t_precond_s;
  rep->t_symbolic_s = s->t_symbolic_s;
  rep->restart_used = s->params.method == TFELCU_METHOD_GMRES ? s->restart_used : 0;
  rep->ilu_zero_pivot_row = -1;
  if (s->params.precond == TFELCU_PC_ILU && s->ilu) {
    rep->ilu_nnz = s->ilu->syncfree ? s->ilu->fnnz : s->ilu->nnz;
    rep->ilu_levels = s->ilu->levels;
    rep->ilu_zero_pivot_row = s->ilu->zero_pivot_row;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
}

}  // namespace tfelcu
