// Shared plumbing of libtfel_cuda: error handling, device buffers, the context.
#pragma once

#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include <memory>

#include "../../include/tfel_cuda.h"

struct tfelcu_matrix;
struct tfelcu_solver;

namespace tfelcu {

struct Error : std::runtime_error {
  explicit Error(const std::string& m) : std::runtime_error(m) {}
};

// a strategy that does not apply to this input (the caller may fall back to another one); never a CUDA failure
struct Refusal : Error {
  explicit Refusal(const std::string& m) : Error(m) {}
};

[[noreturn]] inline void fail(const std::string& m) { throw Error(m); }
[[noreturn]] inline void refuse(const std::string& m) { throw Refusal(m); }

#define TFELCU_CK(call)                                                                         \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      ::tfelcu::fail(std::string(#call) + " failed: " + cudaGetErrorString(e_) + " at " +       \
                     __FILE__ + ":" + std::to_string(__LINE__));                                \
  } while (0)

#define TFELCU_NCCL(call)                                                                       \
  do {                                                                                          \
    ncclResult_t r_ = (call);                                                                   \
    if (r_ != ncclSuccess)                                                                      \
      ::tfelcu::fail(std::string(#call) + " failed: " + ncclGetErrorString(r_) + " at " +       \
                     __FILE__ + ":" + std::to_string(__LINE__));                                \
  } while (0)

#define TFELCU_REQUIRE(cond, msg)                                                               \
  do { if (!(cond)) ::tfelcu::fail(std::string(msg)); } while (0)

// Device array with value semantics disabled (move only).
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  size_t front = 0;   // elements in front of p (p[-front .. -1]): the lower halo of a distributed vector
  DevBuf() = default;
  explicit DevBuf(size_t count) { alloc(count); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), front(o.front) { o.p = nullptr; o.n = 0; o.front = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; front = o.front; o.p = nullptr; o.n = 0; o.front = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t count, size_t in_front = 0) {
    release();
    n = count; front = in_front;
    // +8 elements of slack: bulk copies in the SpMV kernel read whole 16-byte chunks of 2-, 4- and 8-byte entries
    if (count + in_front) {
      T* raw = nullptr;
      TFELCU_CK(cudaMalloc(&raw, (in_front + count + 8) * sizeof(T)));
      p = raw + in_front;
    }
  }
  void release() {
    if (p) cudaFree(p - front);
    p = nullptr; n = 0; front = 0;
  }
  size_t bytes() const { return n * sizeof(T); }
};

inline bool is_device_pointer(const void* p) {
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  int rank = 0, n_ranks = 1;
  ncclComm_t comm = nullptr;
  int sm_count = 148;
  uint64_t launches = 0;
  // peer-to-peer arena (dist.cu): a small device buffer of every rank mapped into every other rank through CUDA IPC;
  // reductions and halo exchanges of the distributed Krylov loop are remote stores + flags over NVLink
  bool p2p = false;
  void* arena = nullptr;
  void* peer_arena[16] = {nullptr};
  cudaStream_t side_stream = nullptr;      // halo exchange overlapped with the interior SpMV
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int* p2p_err = nullptr;                  // device flag: a peer did not answer within the spin limit
  bool p2p_failed = false;                 // ... seen by the host: distributed solves on this context are refused from then on
  unsigned long long* seq_dev = nullptr;   // [0]: reductions done, [1 + q]: halo exchanges done with rank q (device counters,
                                           // advanced by the kernels themselves so that the loop can live in a CUDA graph)
  DevBuf<double> table_scratch;   // reference-element tables of the form being assembled (stream ordered reuse)
  // Newton / time loops construct a bilinear_form and a solver per step (test/navier_stokes_2d_p2_p1.cpp:292-438): destroyed
  // forms and solvers are parked here (a few of each) and handed out again when a form on the same spaces with the same
  // Dirichlet sets / a solver with the same parameters is created, so that the CSR pattern, slot map, colouring, SpMV plan
  // and the symbolic phase of ILU(k) are built once. TFELCU_NO_POOL=1 disables it.
  struct ::tfelcu_matrix* pending_matrix = nullptr;   // the form whose += is deferred (at most one per context)
  std::vector<const void*> live_fes;   // spaces alive in this context (a form is parked only while its spaces are)
  std::vector<struct ::tfelcu_matrix*> matrix_pool;
  std::vector<struct ::tfelcu_solver*> solver_pool;

  void count(uint64_t n = 1) { launches += n; }
  void sync() { TFELCU_CK(cudaStreamSynchronize(stream)); }

  // host <-> device copies on the context stream
  template <typename T>
  void upload(T* dst, const T* src, size_t n) {
    if (!n) return;
    TFELCU_CK(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDefault, stream));
  }
  template <typename T>
  void download(T* dst, const T* src, size_t n) {
    if (!n) return;
    TFELCU_CK(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDefault, stream));
    sync();
  }
  template <typename T>
  void zero(T* dst, size_t n) {
    if (n) TFELCU_CK(cudaMemsetAsync(dst, 0, n * sizeof(T), stream));
  }
};

inline unsigned grid_for(size_t n, unsigned block) {
  size_t g = (n + block - 1) / block;
  if (g == 0) g = 1;
  if (g > 0x7fffffffu) fail("grid too large");
  return static_cast<unsigned>(g);
}

#define TFELCU_LAUNCH_CHECK() TFELCU_CK(cudaGetLastError())

}  // namespace tfelcu

// Opaque handle types of the C ABI are the implementation structs below.
struct tfelcu_ctx : tfelcu::Ctx {};
