// <tfel/tfel_device.cuh> -- coefficient functions as __device__ functors (nvcc-compiled user code).
//
// The reference's leaves free_function / std_function hold HOST function pointers (expression.hpp:31-87); on the CUDA
// path a function of x can instead be any copyable type with   __device__ double operator()(const double* x) const:
//
//   struct gaussian { double sigma; __device__ double operator()(const double* x) const { ... } };
//   b += integrate<quad_type>(dev(gaussian{0.1}) * v, m);
//
// dev(f) instantiates the functor as DEVICE CODE in the user's translation unit: a kernel k_functor_at_quadrature<F, dim>
// that evaluates f at the quadrature points of every cell, and a host launcher for it whose address crosses the C ABI
// (TFELCU_FACTOR_FUNCTION, tfelcu_tabulate_fn) together with a device copy of the functor object. The library calls the
// launcher on its own stream when the form is assembled, exactly where the reference evaluates its free_function
// leaves (expression.hpp:46-48): no host evaluation, no host <-> device copy. TFEL_DEVICE_FUNCTION(name) wraps a plain
// __device__ double name(const double*) the same way.
// (Why a kernel of the user's module and not a __device__ function pointer called by the library's element kernels:
// without relocatable device code a function "address" on sm_100a is a module-local handle -- CALL.REL -- and the call
// faults with "invalid program counter"; with -rdc the element kernels go from 30-64 to 90-200 registers.)
// device_expr<quad>(f, mesh) is the older variant: it evaluates the functor at every quadrature point into a
// device-resident table first (TFELCU_FACTOR_TABLE).
// Include after <tfel/tfel.hpp>; compile with nvcc.
#ifndef TFEL_TFEL_DEVICE_CUH
#define TFEL_TFEL_DEVICE_CUH

#include <cuda_runtime.h>

#include "tfel.hpp"

namespace tfel_cuda {

template <typename F>
__global__ void k_tabulate_functor(F f, const double* __restrict__ xq, size_t n, int dim, double* __restrict__ out) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = f(xq + i * dim);
}

inline void cuda_ck(cudaError_t e) {
  if (e != cudaSuccess) throw std::string("tfel_device: ") + cudaGetErrorString(e);
}

// the device code the library launches: one instance per functor type and dimension, in the user's translation unit
template <typename F, int DIM>
__global__ void k_functor_at_quadrature(const F* __restrict__ f, tfelcu_points w, double* __restrict__ out) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= w.n_cells * (size_t)w.nq) return;
  const size_t k = i / w.nq;
  const int q = (int)(i - k * w.nq);
  const uint32_t* c = w.cells + k * (DIM + 1);
  double v0[DIM], x[DIM];
#pragma unroll
  for (int r = 0; r < DIM; ++r) x[r] = v0[r] = w.vertices[(size_t)c[0] * DIM + r];
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    const double t = w.xhat[q * DIM + d];
#pragma unroll
    for (int r = 0; r < DIM; ++r) x[r] += t * (w.vertices[(size_t)c[d + 1] * DIM + r] - v0[r]);
  }
  out[i] = (*f)(x);
}

template <typename F>
int launch_functor(const void* closure, const tfelcu_points* w, double* out, void* stream) {
  const size_t n = w->n_cells * (size_t)w->nq;
  if (n == 0) return 0;
  const unsigned grid = static_cast<unsigned>((n + 255) / 256);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const F* f = static_cast<const F*>(closure);
  if (w->dim == 2) k_functor_at_quadrature<F, 2><<<grid, 256, 0, s>>>(f, *w, out);
  else if (w->dim == 3) k_functor_at_quadrature<F, 3><<<grid, 256, 0, s>>>(f, *w, out);
  else if (w->dim == 1) k_functor_at_quadrature<F, 1><<<grid, 256, 0, s>>>(f, *w, out);
  else return 1;
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace tfel_cuda

// rank-0 expression whose value at a point x is f(x), evaluated by the assembly kernels themselves
template <typename F>
expression<false> dev(const F& f) {
  static_assert(std::is_trivially_copyable<F>::value, "a device functor is copied to the GPU byte for byte");
  void* closure(nullptr);
  tfel_cuda::cuda_ck(cudaMalloc(&closure, sizeof(F)));
  std::shared_ptr<void> owner(closure, [](void* p) { cudaFree(p); });
  cudaStream_t stream(static_cast<cudaStream_t>(tfelcu_ctx_stream(tfel_cuda::context())));
  tfel_cuda::cuda_ck(cudaMemcpyAsync(closure, &f, sizeof(F), cudaMemcpyHostToDevice, stream));
  tfel_cuda::cuda_ck(cudaStreamSynchronize(stream));
  expression<false> e;
  tfel_cuda::factor_desc d;
  d.kind = TFELCU_FACTOR_FUNCTION; d.device_fn = &tfel_cuda::launch_functor<F>; d.device_closure = owner;
  e.factors.push_back(d);
  tfel_cuda::monomial mono; mono.factors.push_back(0); e.terms.push_back(mono);
  return e;
}

// a plain  __device__ double name(const double* x)  as a functor type name##_functor (dev(name##_functor()) * v)
#define TFEL_DEVICE_FUNCTION(name) \
  struct name##_functor { __device__ double operator()(const double* x) const { return name(x); } }

// rank-0 expression whose value at a quadrature point is f(x_q), evaluated on the device
template <typename quad_t, typename F, typename mesh_t>
expression<false> device_expr(F f, const mesh_t& m) {
  int nq(0), dim(0);
  std::size_t K(0);
  tfel_cuda::ck(tfelcu_quadrature_size(quad_t::id, &nq, &dim));
  tfel_cuda::ck(tfelcu_mesh_info(m.handle(), nullptr, nullptr, &K));
  const std::size_t n(K * nq);
  double* xq(nullptr); double* table(nullptr);
  tfel_cuda::cuda_ck(cudaMalloc(&xq, n * dim * sizeof(double)));
  tfel_cuda::cuda_ck(cudaMalloc(&table, n * sizeof(double)));
  std::shared_ptr<void> owner(table, [](void* p) { cudaFree(p); });
  tfel_cuda::ck(tfelcu_mesh_quadrature_points(m.handle(), quad_t::id, xq));   // device buffer in, device buffer out
  cudaStream_t stream(static_cast<cudaStream_t>(tfelcu_ctx_stream(tfel_cuda::context())));
  tfel_cuda::k_tabulate_functor<F><<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(f, xq, n, dim, table);
  tfel_cuda::cuda_ck(cudaGetLastError());
  tfel_cuda::cuda_ck(cudaStreamSynchronize(stream));
  cudaFree(xq);
  expression<false> e;
  tfel_cuda::factor_desc d;
  d.kind = TFELCU_FACTOR_TABLE; d.device_table = owner; d.table_quad = quad_t::id;
  e.factors.push_back(d);
  tfel_cuda::monomial mono; mono.factors.push_back(0); e.terms.push_back(mono);
  return e;
}

#endif  // TFEL_TFEL_DEVICE_CUH
