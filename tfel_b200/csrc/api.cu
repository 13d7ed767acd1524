// The C ABI of libtfel_cuda.so (include/tfel_cuda.h): argument checking, host <-> device staging and error
// translation around the device implementation in mesh.cu / fes.cu / pattern.cu / assemble.cu / solver.cu /
// spmv.cu / krylov.cu. No entry point computes anything on the host: without a CUDA device every call fails.
#include "structs.cuh"
#include "solver.cuh"

#include <algorithm>
#include <exception>
#include <new>

namespace {

thread_local std::string g_last_error;

template <typename F>
int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const tfelcu::Error& e) {
    g_last_error = e.what();
    return 1;
  } catch (const std::bad_alloc&) {
    g_last_error = "out of host memory";
    return 2;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return 3;
  }
}

using tfelcu::Ctx;
using tfelcu::DevBuf;
using tfelcu::fail;

// device view of a caller buffer that may live on the host: copies it when needed
template <typename T>
struct Staged {
  const T* p = nullptr;
  DevBuf<T> own;
  Staged(Ctx* ctx, const T* src, size_t n) {
    if (!n || !src) return;
    if (tfelcu::is_device_pointer(src)) { p = src; return; }
    own.alloc(n);
    ctx->upload(own.p, src, n);
    p = own.p;
  }
};

template <typename T>
void copy_out(Ctx* ctx, T* dst, const T* d_src, size_t n) {
  if (!n || !dst) return;
  TFELCU_CK(cudaMemcpyAsync(dst, d_src, n * sizeof(T), cudaMemcpyDefault, ctx->stream));
  ctx->sync();
}

void fill_space(tfelcu::SpaceList& L, int n, tfelcu_fes* const* list, const char* what) {
  TFELCU_REQUIRE(n >= 1 && n <= tfelcu::kMaxComp, std::string(what) + ": between 1 and 4 component spaces");
  L.n = n;
  size_t off = 0;
  for (int c = 0; c < n; ++c) {
    TFELCU_REQUIRE(list[c] != nullptr, std::string(what) + ": null space");
    TFELCU_REQUIRE(list[c]->mesh == list[0]->mesh, std::string(what) + ": component spaces live on different meshes");
    L.fes[c] = list[c];
    L.offset[c] = off;
    off += list[c]->n_dof;
  }
  L.offset[n] = off;
}

void use_device(const Ctx* ctx) { TFELCU_CK(cudaSetDevice(ctx->device)); }

}  // namespace

extern "C" {

const char* tfelcu_last_error(void) { return g_last_error.c_str(); }
const char* tfelcu_version(void) { return "tfel_b200 0.1 (sm_100a)"; }

// ---- context ---------------------------------------------------------------------------------

static void ctx_init(tfelcu_ctx* c, int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    fail("no CUDA device: libtfel_cuda has no CPU fallback");
  }
  TFELCU_REQUIRE(device >= 0 && device < count, "device index out of range");
  c->device = device;
  TFELCU_CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  TFELCU_CK(cudaGetDeviceProperties(&prop, device));
  TFELCU_REQUIRE(prop.major >= 10, "libtfel_cuda is built for sm_100a (Blackwell) only");
  c->sm_count = prop.multiProcessorCount;
  TFELCU_CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->own_stream = true;
}

int tfelcu_ctx_create(int device, tfelcu_ctx** ctx) {
  return guarded([&] {
    TFELCU_REQUIRE(ctx != nullptr, "null output pointer");
    std::unique_ptr<tfelcu_ctx> c(new tfelcu_ctx());
    ctx_init(c.get(), device);
    *ctx = c.release();
  });
}

int tfelcu_nccl_unique_id(void* out128) {
  return guarded([&] {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    TFELCU_NCCL(ncclGetUniqueId(&id));
    std::memcpy(out128, &id, sizeof(id));
  });
}

int tfelcu_ctx_create_distributed(int device, int rank, int n_ranks, const void* nccl_unique_id128, tfelcu_ctx** ctx) {
  return guarded([&] {
    TFELCU_REQUIRE(ctx != nullptr && nccl_unique_id128 != nullptr, "null pointer");
    TFELCU_REQUIRE(n_ranks >= 1 && rank >= 0 && rank < n_ranks, "rank out of range");
    std::unique_ptr<tfelcu_ctx> c(new tfelcu_ctx());
    ctx_init(c.get(), device);
    c->rank = rank; c->n_ranks = n_ranks;
    ncclUniqueId id;
    std::memcpy(&id, nccl_unique_id128, sizeof(id));
    TFELCU_NCCL(ncclCommInitRank(&c->comm, n_ranks, id, rank));
    tfelcu::p2p_setup(c.get());   // peer arenas over CUDA IPC; stays on NCCL when IPC is not available
    *ctx = c.release();
  });
}

int tfelcu_ctx_destroy(tfelcu_ctx* ctx) {
  return guarded([&] {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (tfelcu_matrix* m : ctx->matrix_pool) delete m;
    for (tfelcu_solver* sv : ctx->solver_pool) delete sv;
    ctx->matrix_pool.clear(); ctx->solver_pool.clear();
    tfelcu::p2p_teardown(ctx);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
    if (ctx->comm) ncclCommDestroy(ctx->comm);
    ctx->table_scratch.release();
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
  });
}

int tfelcu_ctx_set_stream(tfelcu_ctx* ctx, void* cuda_stream) {
  return guarded([&] {
    TFELCU_REQUIRE(ctx != nullptr, "null context");
    ctx->sync();
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    ctx->own_stream = false;
  });
}

void* tfelcu_ctx_stream(tfelcu_ctx* ctx) { return ctx ? static_cast<void*>(ctx->stream) : nullptr; }

int tfelcu_ctx_synchronize(tfelcu_ctx* ctx) {
  return guarded([&] { TFELCU_REQUIRE(ctx != nullptr, "null context"); use_device(ctx); ctx->sync(); });
}

int tfelcu_ctx_rank(const tfelcu_ctx* ctx, int* rank, int* n_ranks) {
  return guarded([&] {
    TFELCU_REQUIRE(ctx != nullptr, "null context");
    if (rank) *rank = ctx->rank;
    if (n_ranks) *n_ranks = ctx->n_ranks;
  });
}

uint64_t tfelcu_ctx_launch_count(const tfelcu_ctx* ctx) { return ctx ? ctx->launches : 0; }

int tfelcu_ctx_p2p_enabled(const tfelcu_ctx* ctx) { return ctx && ctx->p2p ? 1 : 0; }

int tfelcu_ctx_p2p_bench(tfelcu_ctx* ctx, int reps, double* allreduce_us) {
  return guarded([&] {
    TFELCU_REQUIRE(ctx && allreduce_us && reps > 0, "bad arguments");
    use_device(ctx);
    *allreduce_us = tfelcu::p2p_bench_allreduce(ctx, reps);
  });
}

// diagnostic: FP64 FMA throughput of the device (the denominator of the FP64-pipe figures in profiles/)
namespace {
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0, x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
}  // namespace

int tfelcu_ctx_fp64_peak(tfelcu_ctx* ctx, double* tflops) {
  return guarded([&] {
    TFELCU_REQUIRE(ctx && tflops, "bad arguments");
    use_device(ctx);
    const int grid = ctx->sm_count * 8, iters = 4096;
    DevBuf<double> out((size_t)grid * 256);
    cudaEvent_t e0, e1;
    TFELCU_CK(cudaEventCreate(&e0)); TFELCU_CK(cudaEventCreate(&e1));
    k_dfma_peak<<<grid, 256, 0, ctx->stream>>>(out.p, 64, 0.999999, 1e-9);   // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
      TFELCU_CK(cudaEventRecord(e0, ctx->stream));
      k_dfma_peak<<<grid, 256, 0, ctx->stream>>>(out.p, iters, 0.999999, 1e-9);
      TFELCU_CK(cudaEventRecord(e1, ctx->stream));
      TFELCU_CK(cudaEventSynchronize(e1));
      float ms = 0.f; TFELCU_CK(cudaEventElapsedTime(&ms, e0, e1));
      const double flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)grid * 256.0;
      best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    ctx->count(6); TFELCU_LAUNCH_CHECK();
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *tflops = best;
  });
}

// ---- mesh --------------------------------------------------------------------------------------

int tfelcu_mesh_create(tfelcu_ctx* ctx, int dim, const double* vertices, size_t n_vertices, const uint32_t* cells,
                       size_t n_cells, tfelcu_mesh** mesh) {
  return guarded([&] {
    TFELCU_REQUIRE(ctx && mesh, "null pointer");
    TFELCU_REQUIRE(dim == 2 || dim == 3, "meshes of triangles (dim 2) or tetrahedra (dim 3) only");
    TFELCU_REQUIRE(n_vertices < 0xffffffffull, "too many vertices");
    TFELCU_REQUIRE((vertices || !n_vertices) && (cells || !n_cells), "null mesh arrays");
    use_device(ctx);
    std::unique_ptr<tfelcu_mesh> m(new tfelcu_mesh());
    m->ctx = ctx; m->dim = dim; m->V = n_vertices; m->K = n_cells;
    m->vertices.alloc(n_vertices * dim); m->cells.alloc(n_cells * (dim + 1));
    TFELCU_CK(cudaMemcpyAsync(m->vertices.p, vertices, n_vertices * dim * sizeof(double), cudaMemcpyDefault, ctx->stream));
    TFELCU_CK(cudaMemcpyAsync(m->cells.p, cells, n_cells * (dim + 1) * sizeof(uint32_t), cudaMemcpyDefault, ctx->stream));
    ctx->sync();
    tfelcu::mesh_validate_cells(m.get());
    *mesh = m.release();
  });
}

int tfelcu_mesh_gen_square(tfelcu_ctx* ctx, double x1, double x2, uint32_t n1, uint32_t n2, tfelcu_mesh** mesh) {
  return guarded([&] {
    TFELCU_REQUIRE(ctx && mesh, "null pointer");
    TFELCU_REQUIRE(n1 >= 1 && n2 >= 1, "gen_square_mesh: at least one cell per direction");
    use_device(ctx);
    std::unique_ptr<tfelcu_mesh> m(new tfelcu_mesh());
    m->ctx = ctx;
    tfelcu::mesh_generate_square(m.get(), x1, x2, n1, n2);
    *mesh = m.release();
  });
}

int tfelcu_mesh_gen_cube(tfelcu_ctx* ctx, double x1, double x2, double x3, uint32_t n1, uint32_t n2, uint32_t n3,
                         tfelcu_mesh** mesh) {
  return guarded([&] {
    TFELCU_REQUIRE(ctx && mesh, "null pointer");
    TFELCU_REQUIRE(n1 >= 1 && n2 >= 1 && n3 >= 1, "gen_cube_mesh: at least one cell per direction");
    use_device(ctx);
    std::unique_ptr<tfelcu_mesh> m(new tfelcu_mesh());
    m->ctx = ctx;
    tfelcu::mesh_generate_cube(m.get(), x1, x2, x3, n1, n2, n3);
    *mesh = m.release();
  });
}

int tfelcu_mesh_destroy(tfelcu_mesh* mesh) {
  return guarded([&] { if (mesh) { use_device(mesh->ctx); mesh->ctx->sync(); delete mesh; } });
}

int tfelcu_mesh_info(const tfelcu_mesh* mesh, int* dim, size_t* n_vertices, size_t* n_cells) {
  return guarded([&] {
    TFELCU_REQUIRE(mesh != nullptr, "null mesh");
    if (dim) *dim = mesh->dim;
    if (n_vertices) *n_vertices = mesh->V;
    if (n_cells) *n_cells = mesh->K;
  });
}

int tfelcu_mesh_download(const tfelcu_mesh* mesh, double* vertices, uint32_t* cells) {
  return guarded([&] {
    TFELCU_REQUIRE(mesh != nullptr, "null mesh");
    use_device(mesh->ctx);
    copy_out(mesh->ctx, vertices, mesh->vertices.p, mesh->V * mesh->dim);
    copy_out(mesh->ctx, cells, mesh->cells.p, mesh->K * (mesh->dim + 1));
  });
}

int tfelcu_mesh_geometry(const tfelcu_mesh* mesh, double* vol, double* jmt) {
  return guarded([&] {
    TFELCU_REQUIRE(mesh && vol && jmt, "null pointer");
    use_device(mesh->ctx);
    tfelcu::mesh_geometry(mesh, vol, jmt);
  });
}

int tfelcu_mesh_boundary(tfelcu_mesh* mesh, size_t* n_faces) {
  return guarded([&] {
    TFELCU_REQUIRE(mesh != nullptr, "null mesh");
    use_device(mesh->ctx);
    tfelcu::mesh_build_boundary(mesh);
    if (n_faces) *n_faces = mesh->F;
  });
}

int tfelcu_mesh_boundary_download(const tfelcu_mesh* mesh, uint32_t* el, uint32_t* parent_cell, uint32_t* parent_subdomain) {
  return guarded([&] {
    TFELCU_REQUIRE(mesh && mesh->has_boundary, "call tfelcu_mesh_boundary first");
    use_device(mesh->ctx);
    copy_out(mesh->ctx, el, mesh->b_el.p, mesh->F * mesh->dim);
    copy_out(mesh->ctx, parent_cell, mesh->b_cell.p, mesh->F);
    copy_out(mesh->ctx, parent_subdomain, mesh->b_sd.p, mesh->F);
  });
}

int tfelcu_mesh_neighbours(tfelcu_mesh* mesh, int32_t* nb) {
  return guarded([&] {
    TFELCU_REQUIRE(mesh && nb, "null pointer");
    use_device(mesh->ctx);
    const size_t n = mesh->K * (size_t)(mesh->dim + 1);
    if (tfelcu::is_device_pointer(nb)) { tfelcu::mesh_neighbours(mesh, nb); return; }
    tfelcu::DevBuf<int32_t> d(n);
    tfelcu::mesh_neighbours(mesh, d.p);
    copy_out(mesh->ctx, nb, d.p, n);
  });
}

// ---- finite element space -------------------------------------------------------------------------

int tfelcu_fes_create(tfelcu_ctx* ctx, tfelcu_mesh* mesh, int fe, tfelcu_fes** fes) {
  return guarded([&] {
    TFELCU_REQUIRE(ctx && mesh && fes, "null pointer");
    TFELCU_REQUIRE(mesh->ctx == ctx, "mesh belongs to another context");
    TFELCU_REQUIRE(fe == TFELCU_FE_P1 || fe == TFELCU_FE_P1_BUBBLE || fe == TFELCU_FE_P2, "unknown finite element id");
    use_device(ctx);
    std::unique_ptr<tfelcu_fes> f(new tfelcu_fes());
    f->ctx = ctx; f->mesh = mesh; f->fe = fe;
    tfelcu::fes_build(f.get());
    ctx->live_fes.push_back(f.get());
    *fes = f.release();
  });
}

int tfelcu_fes_destroy(tfelcu_fes* fes) {
  return guarded([&] {
    if (!fes) return;
    use_device(fes->ctx); fes->ctx->sync();
    auto& live = fes->ctx->live_fes;
    for (size_t i = 0; i < live.size(); ++i) if (live[i] == fes) { live.erase(live.begin() + i); break; }
    auto& pool = fes->ctx->matrix_pool;   // parked forms on this space go with it
    for (size_t i = 0; i < pool.size();) {
      bool uses = false;
      for (int c = 0; c < pool[i]->test.n; ++c) uses = uses || pool[i]->test.fes[c] == fes;
      for (int c = 0; c < pool[i]->trial.n; ++c) uses = uses || pool[i]->trial.fes[c] == fes;
      if (uses) { delete pool[i]; pool.erase(pool.begin() + i); } else ++i;
    }
    delete fes;
  });
}

int tfelcu_fes_info(const tfelcu_fes* fes, int* fe, size_t* n_dof, int* n_dof_per_cell) {
  return guarded([&] {
    TFELCU_REQUIRE(fes != nullptr, "null space");
    if (fe) *fe = fes->fe;
    if (n_dof) *n_dof = fes->n_dof;
    if (n_dof_per_cell) *n_dof_per_cell = fes->nd;
  });
}

int tfelcu_fes_kind_offsets(const tfelcu_fes* fes, size_t* offsets) {
  return guarded([&] {
    TFELCU_REQUIRE(fes && offsets, "null pointer");
    const int dim = fes->mesh->dim;
    size_t next = fes->n_dof;
    for (int sd = 4; sd >= 0; --sd) {
      if (sd > dim || !tfelcu::fe_dofs_on(dim, fes->fe, sd)) { offsets[sd] = next; continue; }
      offsets[sd] = fes->kind_offset[sd];
      next = offsets[sd];
    }
    offsets[4] = fes->n_dof;
    // kinds without dofs collapse onto the start of the next kind that has some
    for (int sd = 3; sd >= 0; --sd)
      if (sd > dim || !tfelcu::fe_dofs_on(dim, fes->fe, sd)) offsets[sd] = offsets[sd + 1];
  });
}

int tfelcu_fes_download_dof_map(const tfelcu_fes* fes, uint32_t* dof_map) {
  return guarded([&] {
    TFELCU_REQUIRE(fes && dof_map, "null pointer");
    use_device(fes->ctx);
    copy_out(fes->ctx, dof_map, fes->dof_map, fes->mesh->K * (size_t)fes->nd);
  });
}

int tfelcu_fes_download_g2l(const tfelcu_fes* fes, uint32_t* g2l) {
  return guarded([&] {
    TFELCU_REQUIRE(fes && g2l, "null pointer");
    use_device(fes->ctx);
    copy_out(fes->ctx, g2l, fes->g2l.p, fes->n_dof * 2);
  });
}

int tfelcu_fes_dof_coordinates(const tfelcu_fes* fes, const uint32_t* dofs, size_t n, double* x) {
  return guarded([&] {
    TFELCU_REQUIRE(fes && (x || !n) && (dofs || !n), "null pointer");
    if (!n) return;
    use_device(fes->ctx);
    Ctx* ctx = fes->ctx;
    Staged<uint32_t> d(ctx, dofs, n);
    DevBuf<double> out(n * fes->mesh->dim);
    tfelcu::fes_dof_coordinates(fes, d.p, n, out.p);
    copy_out(ctx, x, out.p, n * fes->mesh->dim);
  });
}

int tfelcu_fes_boundary_dofs(tfelcu_fes* fes, const uint32_t* bcells, size_t n_bcells, int nvb, size_t* n) {
  return guarded([&] {
    TFELCU_REQUIRE(fes != nullptr, "null space");
    use_device(fes->ctx);
    Ctx* ctx = fes->ctx;
    if (!bcells) {
      tfelcu::mesh_build_boundary(fes->mesh);
      tfelcu::fes_boundary_dofs(fes, fes->mesh->b_el.p, fes->mesh->F, fes->mesh->dim);
    } else {
      Staged<uint32_t> b(ctx, bcells, n_bcells * (size_t)nvb);
      tfelcu::fes_boundary_dofs(fes, b.p, n_bcells, nvb);
      ctx->sync();
    }
    if (n) *n = fes->n_cand;
  });
}

int tfelcu_fes_boundary_dofs_get(const tfelcu_fes* fes, uint32_t* dofs, double* x) {
  return guarded([&] {
    TFELCU_REQUIRE(fes != nullptr, "null space");
    use_device(fes->ctx);
    Ctx* ctx = fes->ctx;
    copy_out(ctx, dofs, fes->cand.p, fes->n_cand);
    if (x && fes->n_cand) {
      DevBuf<double> out(fes->n_cand * fes->mesh->dim);
      tfelcu::fes_dof_coordinates(fes, fes->cand.p, fes->n_cand, out.p);
      copy_out(ctx, x, out.p, fes->n_cand * fes->mesh->dim);
    }
  });
}

int tfelcu_fes_add_dirichlet(tfelcu_fes* fes, const uint32_t* dofs, const double* values, size_t n) {
  return guarded([&] {
    TFELCU_REQUIRE(fes && (dofs || !n) && (values || !n), "null pointer");
    if (!n) return;
    use_device(fes->ctx);
    Ctx* ctx = fes->ctx;
    Staged<uint32_t> d(ctx, dofs, n);
    Staged<double> v(ctx, values, n);
    tfelcu::fes_add_dirichlet(fes, d.p, v.p, 0.0, n);
  });
}

int tfelcu_fes_add_dirichlet_const(tfelcu_fes* fes, double value) {
  return guarded([&] {
    TFELCU_REQUIRE(fes != nullptr, "null space");
    use_device(fes->ctx);
    tfelcu::fes_add_dirichlet(fes, fes->cand.p, nullptr, value, fes->n_cand);
  });
}

int tfelcu_fes_dirichlet_count(const tfelcu_fes* fes, size_t* n) {
  return guarded([&] { TFELCU_REQUIRE(fes && n, "null pointer"); *n = fes->n_dirichlet; });
}

int tfelcu_fes_dirichlet_get(const tfelcu_fes* fes, uint32_t* dofs, double* values) {
  return guarded([&] {
    TFELCU_REQUIRE(fes != nullptr, "null space");
    use_device(fes->ctx);
    tfelcu::fes_dirichlet_get(fes, dofs, values);
  });
}

static void fes_values(const tfelcu_fes* fes, const double* coefficients, double* out, bool vertex) {
  TFELCU_REQUIRE(fes && coefficients && out, "null pointer");
  use_device(fes->ctx);
  Ctx* ctx = fes->ctx;
  const size_t n_out = vertex ? fes->mesh->V : fes->mesh->K;
  Staged<double> c(ctx, coefficients, fes->n_dof);
  const bool dev_out = tfelcu::is_device_pointer(out);
  tfelcu::DevBuf<double> tmp;
  if (!dev_out) tmp.alloc(n_out);
  double* d_out = dev_out ? out : tmp.p;
  if (vertex) tfelcu::fes_vertex_values(fes, c.p, d_out); else tfelcu::fes_cell_values(fes, c.p, d_out);
  if (!dev_out) copy_out(ctx, out, d_out, n_out); else ctx->sync();
}

int tfelcu_fe_tabulate(int dim, int fe, const double* x_hat, double* out, int* n_dof) {
  return guarded([&] {
    TFELCU_REQUIRE(x_hat && out, "null pointer");
    TFELCU_REQUIRE(dim == 2 || dim == 3, "dimension must be 2 or 3");
    TFELCU_REQUIRE(fe == TFELCU_FE_P1 || fe == TFELCU_FE_P1_BUBBLE || fe == TFELCU_FE_P2, "unknown finite element id");
    tfelcu::fe_eval_hat(dim, fe, x_hat, out);
    if (n_dof) *n_dof = tfelcu::fe_ndof(dim, fe);
  });
}

int tfelcu_fes_vertex_values(const tfelcu_fes* fes, const double* coefficients, double* out) {
  return guarded([&] { fes_values(fes, coefficients, out, true); });
}

int tfelcu_fes_cell_values(const tfelcu_fes* fes, const double* coefficients, double* out) {
  return guarded([&] { fes_values(fes, coefficients, out, false); });
}

// ---- bilinear form / matrix -----------------------------------------------------------------------

static void matrix_create_impl(tfelcu_ctx* ctx, int n_test, tfelcu_fes* const* test, int n_trial, tfelcu_fes* const* trial,
                               int n_seg, const size_t* row_begin, const size_t* row_end, tfelcu_matrix** m) {
  TFELCU_REQUIRE(ctx && test && trial && m, "null pointer");
  use_device(ctx);
  std::unique_ptr<tfelcu_matrix> A(new tfelcu_matrix());
  A->ctx = ctx;
  fill_space(A->test, n_test, test, "test space");
  fill_space(A->trial, n_trial, trial, "trial space");
  TFELCU_REQUIRE(A->test.fes[0]->mesh == A->trial.fes[0]->mesh, "test and trial spaces live on different meshes");
  A->mesh = A->test.fes[0]->mesh;
  A->n_rows = A->test.total(); A->n_cols = A->trial.total();
  A->rows = tfelcu::make_row_map(A->test, n_seg, row_begin, row_end);
  // a parked form on the same spaces, the same Dirichlet sets and the same rows: reuse its structure (pattern, slot map,
  // colouring, patches); the values are cleared lazily like after clear()
  for (size_t i = 0; i < ctx->matrix_pool.size(); ++i) {
    tfelcu_matrix* P = ctx->matrix_pool[i];
    bool same = P->test.n == A->test.n && P->trial.n == A->trial.n && P->rows.n == A->rows.n && P->rows.n_global == A->rows.n_global;
    for (int c = 0; same && c < A->test.n; ++c) same = P->test.fes[c] == A->test.fes[c] && P->dir_version[c] == A->test.fes[c]->dirichlet_version;
    for (int c = 0; same && c < A->trial.n; ++c) same = P->trial.fes[c] == A->trial.fes[c];
    for (int g = 0; same && g < A->rows.n; ++g) same = P->rows.g_begin[g] == A->rows.g_begin[g] && P->rows.g_end[g] == A->rows.g_end[g];
    if (!same) continue;
    ctx->matrix_pool.erase(ctx->matrix_pool.begin() + i);
    P->val_valid = false; P->pristine = true;
    if (P->pending) { tfelcu::form_build_free(P->pending); P->pending = nullptr; }
    P->scatter = TFELCU_SCATTER_AUTO; P->patch_refused = false;
    *m = P;
    return;
  }
  tfelcu::matrix_snapshot_dirichlet(A.get());
  tfelcu::matrix_build_pattern(A.get());
  ctx->sync();
  *m = A.release();
}

int tfelcu_matrix_create(tfelcu_ctx* ctx, int n_test, tfelcu_fes* const* test, int n_trial, tfelcu_fes* const* trial,
                         tfelcu_matrix** m) {
  return guarded([&] { matrix_create_impl(ctx, n_test, test, n_trial, trial, 0, nullptr, nullptr, m); });
}

int tfelcu_matrix_create_rows(tfelcu_ctx* ctx, int n_test, tfelcu_fes* const* test, int n_trial, tfelcu_fes* const* trial,
                              int n_segments, const size_t* row_begin, const size_t* row_end, tfelcu_matrix** m) {
  return guarded([&] {
    TFELCU_REQUIRE(row_begin && row_end, "null row ranges");
    matrix_create_impl(ctx, n_test, test, n_trial, trial, n_segments, row_begin, row_end, m);
  });
}

int tfelcu_matrix_local_rows(const tfelcu_matrix* m, size_t* n_local) {
  return guarded([&] { TFELCU_REQUIRE(m && n_local, "null pointer"); *n_local = m->rows.n_local(); });
}

int tfelcu_matrix_destroy(tfelcu_matrix* m) {
  return guarded([&] {
    if (!m) return;
    use_device(m->ctx); m->ctx->sync();
    if (m->pending) { tfelcu::form_build_free(m->pending); m->pending = nullptr; }   // a += nobody looked at
    if (m->ctx->pending_matrix == m) m->ctx->pending_matrix = nullptr;
    auto& pool = m->ctx->matrix_pool;
    bool spaces_alive = true;
    auto alive = [&](const tfelcu_fes* f) {
      for (const void* p : m->ctx->live_fes) if (p == f) return true;
      return false;
    };
    for (int c = 0; c < m->test.n; ++c) spaces_alive = spaces_alive && alive(m->test.fes[c]);
    for (int c = 0; c < m->trial.n; ++c) spaces_alive = spaces_alive && alive(m->trial.fes[c]);
    if (!std::getenv("TFELCU_NO_POOL") && spaces_alive) {   // park it (see Ctx::matrix_pool); the oldest of more than three goes
      pool.push_back(m);
      if (pool.size() > 3) { delete pool.front(); pool.erase(pool.begin()); }
    } else delete m;
  });
}

// clear() re-reads the Dirichlet sets of the test spaces (bilinear_form.hpp:159-165). When they changed since the
// pattern was built the pattern changes with them (identity rows), exactly as a fresh std::map would.
int tfelcu_matrix_clear(tfelcu_matrix* m) {
  return guarded([&] {
    TFELCU_REQUIRE(m != nullptr, "null matrix");
    use_device(m->ctx);
    bool changed = false;
    for (int c = 0; c < m->test.n; ++c) changed = changed || (m->dir_version[c] != m->test.fes[c]->dirichlet_version);
    if (changed) {
      tfelcu::matrix_snapshot_dirichlet(m);
      tfelcu::matrix_build_pattern(m);
    }
    if (m->pending) { tfelcu::form_build_free(m->pending); m->pending = nullptr; }   // a += x; clear(): x is gone
    if (m->ctx->pending_matrix == m) m->ctx->pending_matrix = nullptr;
    m->val_valid = false;
    m->pristine = true;
  });
}

int tfelcu_matrix_info(const tfelcu_matrix* m, size_t* n_rows, size_t* n_cols, size_t* nnz) {
  return guarded([&] {
    TFELCU_REQUIRE(m != nullptr, "null matrix");
    if (n_rows) *n_rows = m->n_rows;
    if (n_cols) *n_cols = m->n_cols;
    if (nnz) *nnz = m->nnz;
  });
}

int tfelcu_matrix_set_scatter(tfelcu_matrix* m, int scatter) {
  return guarded([&] {
    TFELCU_REQUIRE(m != nullptr, "null matrix");
    TFELCU_REQUIRE(scatter >= TFELCU_SCATTER_ROW_GATHER && scatter <= TFELCU_SCATTER_TILE, "unknown scatter strategy");
    m->scatter = scatter;
    m->patch_refused = false;
  });
}

int tfelcu_matrix_assemble(tfelcu_matrix* m, int quad, const tfelcu_factor* factors, int n_factors,
                           const tfelcu_term* terms, int n_terms) {
  return guarded([&] {
    TFELCU_REQUIRE(m && (terms || !n_terms) && (factors || !n_factors), "null pointer");
    use_device(m->ctx);
    tfelcu::assemble_bilinear(m, quad, factors, n_factors, terms, n_terms);
  });
}

int tfelcu_matrix_download_csr(const tfelcu_matrix* m, int32_t* rowptr, int32_t* colind, double* val) {
  return guarded([&] {
    TFELCU_REQUIRE(m != nullptr, "null matrix");
    use_device(m->ctx);
    if (val) tfelcu::matrix_ensure_cleared(const_cast<tfelcu_matrix*>(m));
    copy_out(m->ctx, rowptr, m->rowptr.p, m->rows.n_local() + 1);
    copy_out(m->ctx, colind, m->colind.p, m->nnz);
    copy_out(m->ctx, val, m->val.p, m->nnz);
  });
}

int tfelcu_matrix_color_count(tfelcu_matrix* m, int* n_colors) {
  return guarded([&] {
    TFELCU_REQUIRE(m && n_colors, "null pointer");
    use_device(m->ctx);
    tfelcu::matrix_build_colors(m);
    *n_colors = m->n_colors;
  });
}

// ---- linear form / vectors --------------------------------------------------------------------------

static void vector_create_impl(tfelcu_ctx* ctx, int n_test, tfelcu_fes* const* test, int n_seg, const size_t* row_begin,
                               const size_t* row_end, tfelcu_vector** v) {
  TFELCU_REQUIRE(ctx && test && v, "null pointer");
  use_device(ctx);
  std::unique_ptr<tfelcu_vector> b(new tfelcu_vector());
  b->ctx = ctx;
  fill_space(b->space, n_test, test, "test space");
  b->rows = tfelcu::make_row_map(b->space, n_seg, row_begin, row_end);
  b->data.alloc(b->rows.n_local());
  ctx->zero(b->data.p, b->rows.n_local());
  *v = b.release();
}

int tfelcu_vector_create(tfelcu_ctx* ctx, int n_test, tfelcu_fes* const* test, tfelcu_vector** v) {
  return guarded([&] { vector_create_impl(ctx, n_test, test, 0, nullptr, nullptr, v); });
}

int tfelcu_vector_create_rows(tfelcu_ctx* ctx, int n_test, tfelcu_fes* const* test, int n_segments, const size_t* row_begin,
                              const size_t* row_end, tfelcu_vector** v) {
  return guarded([&] {
    TFELCU_REQUIRE(row_begin && row_end, "null row ranges");
    vector_create_impl(ctx, n_test, test, n_segments, row_begin, row_end, v);
  });
}

int tfelcu_vector_destroy(tfelcu_vector* v) {
  return guarded([&] { if (v) { use_device(v->ctx); v->ctx->sync(); delete v; } });
}

int tfelcu_vector_clear(tfelcu_vector* v) {
  // lazily: a fused assembly pass that follows writes the entries instead of adding to zeros (one memset and one read less)
  return guarded([&] { TFELCU_REQUIRE(v != nullptr, "null vector"); use_device(v->ctx); v->zero_pending = true; });
}

int tfelcu_vector_size(const tfelcu_vector* v, size_t* n) {
  return guarded([&] { TFELCU_REQUIRE(v && n, "null pointer"); *n = v->data.n; });
}

int tfelcu_vector_set_scatter(tfelcu_vector* v, int scatter) {
  return guarded([&] {
    TFELCU_REQUIRE(v != nullptr, "null vector");
    TFELCU_REQUIRE(scatter == TFELCU_SCATTER_ROW_GATHER || scatter == TFELCU_SCATTER_PATCH || scatter == TFELCU_SCATTER_AUTO ||
                   scatter == TFELCU_SCATTER_TILE, "unknown scatter strategy for a linear form");
    v->scatter = scatter == TFELCU_SCATTER_TILE ? TFELCU_SCATTER_AUTO : scatter;   // a linear form rides along with a tile += (AUTO)
    v->patch_refused = false;
  });
}

int tfelcu_vector_assemble(tfelcu_vector* v, int quad, const tfelcu_factor* factors, int n_factors,
                           const tfelcu_term* terms, int n_terms) {
  return guarded([&] {
    TFELCU_REQUIRE(v && (terms || !n_terms) && (factors || !n_factors), "null pointer");
    use_device(v->ctx);
    tfelcu::assemble_linear(v, quad, factors, n_factors, terms, n_terms);
  });
}

int tfelcu_vector_download(const tfelcu_vector* v, double* host) {
  return guarded([&] {
    TFELCU_REQUIRE(v && host, "null pointer");
    use_device(v->ctx);
    const_cast<tfelcu_vector*>(v)->ensure_cleared();
    copy_out(v->ctx, host, v->data.p, v->data.n);
  });
}

int tfelcu_vector_upload(tfelcu_vector* v, const double* host) {
  return guarded([&] {
    TFELCU_REQUIRE(v && host, "null pointer");
    use_device(v->ctx);
    v->zero_pending = false;
    TFELCU_CK(cudaMemcpyAsync(v->data.p, host, v->data.n * sizeof(double), cudaMemcpyDefault, v->ctx->stream));
    v->ctx->sync();
  });
}

double* tfelcu_vector_device_ptr(tfelcu_vector* v) {
  if (!v) return nullptr;
  cudaSetDevice(v->ctx->device);
  try { v->ensure_cleared(); } catch (...) { return nullptr; }
  return v->data.p;
}

int tfelcu_integrate(tfelcu_ctx* ctx, tfelcu_mesh* mesh, int quad, const tfelcu_factor* factors, int n_factors,
                     const tfelcu_term* terms, int n_terms, double* result) {
  return guarded([&] {
    TFELCU_REQUIRE(ctx && mesh && result && (terms || !n_terms), "null pointer");
    use_device(ctx);
    *result = tfelcu::integrate_scalar(ctx, mesh, quad, factors, n_factors, terms, n_terms);
  });
}

int tfelcu_quadrature_size(int quad, int* n_points, int* dim) {
  return guarded([&] {
    TFELCU_REQUIRE(n_points != nullptr, "null pointer");
    *n_points = tfelcu::quadrature_size(quad, dim);
  });
}

int tfelcu_mesh_quadrature_points(tfelcu_mesh* mesh, int quad, double* xq) {
  return guarded([&] {
    TFELCU_REQUIRE(mesh && xq, "null pointer");
    use_device(mesh->ctx);
    Ctx* ctx = mesh->ctx;
    const size_t n = mesh->K * (size_t)tfelcu::quadrature_size(quad, nullptr) * mesh->dim;
    if (tfelcu::is_device_pointer(xq)) { tfelcu::quadrature_points(ctx, mesh, quad, xq); return; }
    DevBuf<double> d(n);
    tfelcu::quadrature_points(ctx, mesh, quad, d.p);
    copy_out(ctx, xq, d.p, n);
  });
}

// ---- plain device arrays: coefficient vectors of discrete functions that stay on the device -------------------------
namespace {
__global__ void k_lincomb(size_t n, double a, const double* __restrict__ x, double b, const double* __restrict__ y, double* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = y ? a * x[i] + b * y[i] : a * x[i];
}
}  // namespace

int tfelcu_array_create(tfelcu_ctx* ctx, size_t n, double** out) {
  return guarded([&] {
    TFELCU_REQUIRE(ctx && out, "null pointer");
    use_device(ctx);
    double* p = nullptr;
    TFELCU_CK(cudaMalloc(&p, (n ? n : 1) * sizeof(double)));
    TFELCU_CK(cudaMemsetAsync(p, 0, (n ? n : 1) * sizeof(double), ctx->stream));
    *out = p;
  });
}

int tfelcu_array_destroy(tfelcu_ctx* ctx, double* p) {
  return guarded([&] {
    if (!ctx || !p) return;
    use_device(ctx); ctx->sync();
    cudaFree(p);
  });
}

int tfelcu_array_copy(tfelcu_ctx* ctx, double* dst, const double* src, size_t n) {
  return guarded([&] {
    TFELCU_REQUIRE(ctx && (dst || !n) && (src || !n), "null pointer");
    use_device(ctx);
    if (!n) return;
    TFELCU_CK(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDefault, ctx->stream));
    ctx->sync();
  });
}

int tfelcu_array_lincomb(tfelcu_ctx* ctx, size_t n, double a, const double* x, double b, const double* y, double* out) {
  return guarded([&] {
    TFELCU_REQUIRE(ctx && x && out, "null pointer");
    use_device(ctx);
    if (!n) return;
    TFELCU_REQUIRE(tfelcu::is_device_pointer(x) && tfelcu::is_device_pointer(out) && (!y || tfelcu::is_device_pointer(y)),
                   "array arithmetic works on device arrays");
    size_t g = (n + 255) / 256; if (g > (size_t)ctx->sm_count * 8) g = (size_t)ctx->sm_count * 8;
    k_lincomb<<<(unsigned)g, 256, 0, ctx->stream>>>(n, a, x, b, y, out);
    ctx->count(); TFELCU_LAUNCH_CHECK();
  });
}

// ---- solver --------------------------------------------------------------------------------------

int tfelcu_solver_create(tfelcu_ctx* ctx, const tfelcu_solver_params* params, tfelcu_solver** s) {
  return guarded([&] {
    TFELCU_REQUIRE(ctx && params && s, "null pointer");
    TFELCU_REQUIRE(params->method == TFELCU_METHOD_CG || params->method == TFELCU_METHOD_GMRES, "unknown Krylov method");
    TFELCU_REQUIRE(params->precond >= TFELCU_PC_NONE && params->precond <= TFELCU_PC_ILU, "unknown preconditioner");
    TFELCU_REQUIRE(params->rtol >= 0.0 && params->atol >= 0.0, "negative tolerance");
    TFELCU_REQUIRE(params->ilu_levels >= 0 && params->ilu_levels <= 16, "ilufill out of range (0 .. 16)");
    // a parked solver with the same structural parameters keeps its SpMV plan and ILU symbolic phase (Ctx::solver_pool)
    for (size_t i = 0; i < ctx->solver_pool.size(); ++i) {
      tfelcu_solver* P = ctx->solver_pool[i];
      if (P->params.method == params->method && P->params.precond == params->precond && P->params.ilu_levels == params->ilu_levels &&
          P->params.block_size == params->block_size && P->params.spmv_kernel == params->spmv_kernel) {
        ctx->solver_pool.erase(ctx->solver_pool.begin() + i);
        P->params = *params;
        *s = P;
        return;
      }
    }
    std::unique_ptr<tfelcu_solver> S(new tfelcu_solver());
    S->ctx = ctx; S->params = *params;
    *s = S.release();
  });
}

int tfelcu_solver_destroy(tfelcu_solver* s) {
  return guarded([&] {
    if (!s) return;
    use_device(s->ctx); s->ctx->sync();
    auto& pool = s->ctx->solver_pool;
    if (!std::getenv("TFELCU_NO_POOL") && s->pattern_id != 0) {
      s->basis.release(); s->hess.release(); s->gmres_m = 0;   // the Krylov basis is the large part: not kept while parked
      pool.push_back(s);
      if (pool.size() > 2) { delete pool.front(); pool.erase(pool.begin()); }
    } else delete s;
  });
}

// The solver either COPIES the rows of the bilinear form (set_operator, like solver.cpp:71-96) or borrows its device
// arrays for the duration of one tfelcu_matrix_solve. A form that holds only the rows this rank owns makes the solve
// distributed. Returns true when the pattern is the one the solver's structural data were built for.
static bool attach_matrix(tfelcu_solver* s, tfelcu_matrix* m, bool copy) {
  Ctx* ctx = s->ctx;
  const size_t n = m->rows.n_local();
  const bool same = s->pattern_id != 0 && s->pattern_id == m->pattern_id && s->n == n && s->nnz == m->nnz &&
                    (bool)s->dist == !m->rows.whole();
  s->n = n; s->n_global = m->n_rows; s->nnz = m->nnz;
  if (copy) {
    if (!same || s->own_val.n != m->nnz || s->own_rowptr.n != n + 1) {
      s->own_rowptr.alloc(n + 1); s->own_colind.alloc(m->nnz); s->own_val.alloc(m->nnz);
      TFELCU_CK(cudaMemcpyAsync(s->own_rowptr.p, m->rowptr.p, (n + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
      if (m->nnz) TFELCU_CK(cudaMemcpyAsync(s->own_colind.p, m->colind.p, m->nnz * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (m->nnz) TFELCU_CK(cudaMemcpyAsync(s->own_val.p, m->val.p, m->nnz * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    s->rowptr = s->own_rowptr.p; s->colind = s->own_colind.p; s->val = s->own_val.p;
    s->borrowed = false;
  } else {
    s->own_rowptr.release(); s->own_colind.release(); s->own_val.release();
    s->rowptr = m->rowptr.p; s->colind = m->colind.p; s->val = m->val.p;
    s->borrowed = true;
  }
  if (!same) {
    s->dist.reset();
    if (!m->rows.whole()) {
      s->dist.reset(new tfelcu::DistPlan());
      s->dist->rows = m->rows;
    }
  }
  s->pattern_id = 0;   // set by the caller once the set-up has succeeded
  return same;
}

int tfelcu_solver_set_operator(tfelcu_solver* s, tfelcu_matrix* m) {
  return guarded([&] {
    TFELCU_REQUIRE(s && m, "null pointer");
    TFELCU_REQUIRE(m->n_rows == m->n_cols, "the operator of a linear solve must be square");
    use_device(s->ctx);
    tfelcu::matrix_ensure_cleared(m);
    const bool same = attach_matrix(s, m, true);
    tfelcu::solver_setup(s, !same);
    s->pattern_id = m->pattern_id;
  });
}

int tfelcu_solver_set_operator_csr(tfelcu_solver* s, size_t n, const int32_t* rowptr, const int32_t* colind, const double* val) {
  return guarded([&] {
    TFELCU_REQUIRE(s && rowptr && (colind || !n) && (val || !n), "null pointer");
    use_device(s->ctx);
    Ctx* ctx = s->ctx;
    int32_t nnz = 0;
    TFELCU_CK(cudaMemcpyAsync(&nnz, rowptr + n, sizeof(int32_t), cudaMemcpyDefault, ctx->stream));
    ctx->sync();
    TFELCU_REQUIRE(nnz >= 0, "corrupt row pointer array");
    s->own_rowptr.alloc(n + 1); s->own_colind.alloc((size_t)nnz); s->own_val.alloc((size_t)nnz);
    TFELCU_CK(cudaMemcpyAsync(s->own_rowptr.p, rowptr, (n + 1) * sizeof(int32_t), cudaMemcpyDefault, ctx->stream));
    if (nnz) {
      TFELCU_CK(cudaMemcpyAsync(s->own_colind.p, colind, (size_t)nnz * sizeof(int32_t), cudaMemcpyDefault, ctx->stream));
      TFELCU_CK(cudaMemcpyAsync(s->own_val.p, val, (size_t)nnz * sizeof(double), cudaMemcpyDefault, ctx->stream));
    }
    ctx->sync();
    s->n = n; s->n_global = n; s->nnz = (size_t)nnz;
    s->rowptr = s->own_rowptr.p; s->colind = s->own_colind.p; s->val = s->own_val.p;
    s->dist.reset();
    s->pattern_id = 0; s->borrowed = false;
    tfelcu::solver_setup(s);
  });
}

int tfelcu_solver_solve(tfelcu_solver* s, const double* rhs, double* x, tfelcu_solver_report* report) {
  return guarded([&] {
    TFELCU_REQUIRE(s && rhs && x, "null pointer");
    use_device(s->ctx);
    Ctx* ctx = s->ctx;
    tfelcu_solver_report local;
    tfelcu_solver_report* rep = report ? report : &local;
    Staged<double> b(ctx, rhs, s->n);
    if (tfelcu::is_device_pointer(x)) {
      tfelcu::solver_solve(s, b.p, x, rep);
    } else {
      DevBuf<double> dx(s->n);
      tfelcu::solver_solve(s, b.p, dx.p, rep);
      copy_out(ctx, x, dx.p, s->n);
    }
  });
}

// rhs of bilinear_form::solve: b with the Dirichlet rows replaced by their values (bilinear_form.hpp:126-137)
namespace {
__global__ void k_make_rhs(const double* __restrict__ b, const uint8_t* __restrict__ dir_row,
                           const double* __restrict__ dir_value, tfelcu::RowMap R, double* __restrict__ rhs) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= R.local_off[R.n]) return;
  const unsigned long long g = R.to_global(i);
  rhs[i] = dir_row[g] ? dir_value[g] : b[i];
}

void make_rhs(const tfelcu_matrix* m, const tfelcu_vector* b, double* d_rhs) {
  bool same = b->space.total() == m->n_rows && b->rows.n == m->rows.n;
  for (int c = 0; same && c < m->rows.n; ++c) same = b->rows.g_begin[c] == m->rows.g_begin[c] && b->rows.g_end[c] == m->rows.g_end[c];
  TFELCU_REQUIRE(same, "linear form and bilinear form have different test spaces or own different rows");
  Ctx* ctx = m->ctx;
  if (!m->rows.n_local()) return;
  const_cast<tfelcu_vector*>(b)->ensure_cleared();
  k_make_rhs<<<tfelcu::grid_for(m->rows.n_local(), 256), 256, 0, ctx->stream>>>(b->data.p, m->dir_row.p, m->dir_value.p, m->rows, d_rhs);
  ctx->count(); TFELCU_LAUNCH_CHECK();
}
}  // namespace

int tfelcu_matrix_make_rhs(const tfelcu_matrix* m, const tfelcu_vector* b, double* rhs) {
  return guarded([&] {
    TFELCU_REQUIRE(m && b && rhs, "null pointer");
    use_device(m->ctx);
    DevBuf<double> d(m->rows.n_local());
    make_rhs(m, b, d.p);
    copy_out(m->ctx, rhs, d.p, m->rows.n_local());
  });
}

int tfelcu_matrix_solve(tfelcu_matrix* m, const tfelcu_vector* b, tfelcu_solver* s, double* x, tfelcu_solver_report* report) {
  return guarded([&] {
    TFELCU_REQUIRE(m && b && s && x, "null pointer");
    TFELCU_REQUIRE(m->n_rows == m->n_cols, "the operator of a linear solve must be square");
    use_device(m->ctx);
    Ctx* ctx = m->ctx;
    // s.set_operator(a) on every solve, like bilinear_form.hpp:139 -- without the copy: the form outlives this call
    tfelcu::matrix_ensure_cleared(m);
    const bool same = attach_matrix(s, m, false);
    struct Live { tfelcu_solver* s; ~Live() { s->borrowed_live = false; } } live{s};
    s->borrowed_live = true;
    tfelcu::solver_setup(s, !same);
    s->pattern_id = m->pattern_id;
    tfelcu_solver_report local;
    tfelcu_solver_report* rep = report ? report : &local;
    const size_t n_local = m->rows.n_local();
    DevBuf<double> rhs(n_local);
    make_rhs(m, b, rhs.p);
    // the solution comes back in reference numbering, complete on every rank (an element of the trial space)
    const bool x_on_device = tfelcu::is_device_pointer(x);
    DevBuf<double> dx_full;
    double* d_x = x;
    if (!x_on_device) { dx_full.alloc(m->n_cols); d_x = dx_full.p; }
    if (s->dist) {
      DevBuf<double> x_local(n_local);
      tfelcu::solver_solve(s, rhs.p, x_local.p, rep);
      tfelcu::dist_gather(s, x_local.p, d_x);
      ctx->sync();
    } else {
      tfelcu::solver_solve(s, rhs.p, d_x, rep);
    }
    if (!x_on_device) copy_out(ctx, x, d_x, m->n_cols);
  });
}

int tfelcu_solver_gather(tfelcu_solver* s, const double* x_local, double* x_global) {
  return guarded([&] {
    TFELCU_REQUIRE(s && x_local && x_global, "null pointer");
    TFELCU_REQUIRE(s->rowptr != nullptr, "solver has no operator: call set_operator first");
    use_device(s->ctx);
    Ctx* ctx = s->ctx;
    if (!s->dist) {   // nothing to gather
      TFELCU_CK(cudaMemcpyAsync(x_global, x_local, s->n * sizeof(double), cudaMemcpyDefault, ctx->stream));
      ctx->sync();
      return;
    }
    Staged<double> xl(ctx, x_local, s->n);
    if (tfelcu::is_device_pointer(x_global)) {
      tfelcu::dist_gather(s, xl.p, x_global);
      ctx->sync();
    } else {
      DevBuf<double> full(s->n_global);
      tfelcu::dist_gather(s, xl.p, full.p);
      copy_out(ctx, x_global, full.p, s->n_global);
    }
  });
}

int tfelcu_solver_spmv(tfelcu_solver* s, const double* x, double* y) {
  return guarded([&] {
    TFELCU_REQUIRE(s && x && y, "null pointer");
    TFELCU_REQUIRE(s->rowptr != nullptr && !s->borrowed, "solver has no operator: call set_operator first");
    use_device(s->ctx);
    Ctx* ctx = s->ctx;
    Staged<double> dx(ctx, x, s->n);
    const double* in = dx.p;
    if (s->dist) {   // owned entries + halo from the neighbours
      TFELCU_CK(cudaMemcpyAsync(s->p.p, dx.p, s->n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
      tfelcu::dist_halo_exchange(s, s->p.p, nullptr);
      in = s->p.p;
    }
    if (tfelcu::is_device_pointer(y)) {
      tfelcu::solver_spmv(s, in, y);
      ctx->sync();
    } else {
      DevBuf<double> dy(s->n);
      tfelcu::solver_spmv(s, in, dy.p);
      copy_out(ctx, y, dy.p, s->n);
    }
  });
}

int tfelcu_solver_spmv_bench(tfelcu_solver* s, int reps, double* avg_ms, double* algorithmic_bytes) {
  return guarded([&] {
    TFELCU_REQUIRE(s && avg_ms, "null pointer");
    TFELCU_REQUIRE(s->rowptr != nullptr && !s->borrowed, "solver has no operator: call set_operator first");
    TFELCU_REQUIRE(reps >= 1, "at least one repetition");
    use_device(s->ctx);
    Ctx* ctx = s->ctx;
    // x = the solver's p work vector filled with ones through the initial-guess path is not needed: any data
    ctx->zero(s->p.p, s->n);
    cudaEvent_t e0, e1;
    TFELCU_CK(cudaEventCreate(&e0)); TFELCU_CK(cudaEventCreate(&e1));
    tfelcu::solver_spmv(s, s->p.p, s->q.p);   // warm-up
    TFELCU_CK(cudaEventRecord(e0, ctx->stream));
    for (int r = 0; r < reps; ++r) tfelcu::solver_spmv(s, s->p.p, s->q.p);
    TFELCU_CK(cudaEventRecord(e1, ctx->stream));
    TFELCU_CK(cudaEventSynchronize(e1));
    float ms = 0.f; TFELCU_CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *avg_ms = (double)ms / reps;
    if (algorithmic_bytes) *algorithmic_bytes = 12.0 * (double)s->nnz + 4.0 * (double)(s->n + 1) + 16.0 * (double)s->n;
  });
}

int tfelcu_solver_ilu_info(const tfelcu_solver* s, size_t* n, size_t* nnz, int* levels) {
  return guarded([&] {
    TFELCU_REQUIRE(s && s->ilu && s->ilu->syncfree, "the solver holds no ILU factorisation");
    if (n) *n = s->n;
    if (nnz) *nnz = s->ilu->fnnz;
    if (levels) *levels = s->ilu->levels;
  });
}

int tfelcu_solver_ilu_download(const tfelcu_solver* s, int64_t* rowptr, int32_t* colind, double* lu) {
  return guarded([&] {
    TFELCU_REQUIRE(s && s->ilu && s->ilu->syncfree, "the solver holds no ILU factorisation");
    use_device(s->ctx);
    static_assert(sizeof(long long) == sizeof(int64_t), "64-bit row pointers");
    copy_out(s->ctx, reinterpret_cast<long long*>(rowptr), s->ilu->fptr.p, s->n + 1);
    copy_out(s->ctx, colind, s->ilu->fcol, s->ilu->fnnz);
    copy_out(s->ctx, lu, s->ilu->lu.p, s->ilu->fnnz);
  });
}

}  // extern "C"
