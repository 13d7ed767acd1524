// <tfel/tfel_device.cuh> -- coefficient functions as __device__ functors (nvcc-compiled user code).
//
// The reference's leaves free_function / std_function hold HOST function pointers (expression.hpp:31-87); on the CUDA
// path a function of x can instead be any copyable type with   __device__ double operator()(const double* x) const:
//
//   struct gaussian { double sigma; __device__ double operator()(const double* x) const { ... } };
//   b += integrate<quad_type>(device_expr<quad_type>(gaussian{0.1}, m) * v, m);
//
// device_expr instantiates the functor in a kernel of the USER's translation unit, evaluates it at the quadrature
// points of every cell (cell.hpp:456-468) on the device, and hands the device-resident table to the assembly kernels
// (TFELCU_FACTOR_TABLE): no host evaluation, no host <-> device copy. Include after <tfel/tfel.hpp>; compile with nvcc.
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

}  // namespace tfel_cuda

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
