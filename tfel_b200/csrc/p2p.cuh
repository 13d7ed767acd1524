// Peer-to-peer arena of the distributed Krylov loop (one process per GPU, CUDA IPC over NVLink): device-side pieces
// shared by dist.cu (stand-alone halo / allreduce kernels), spmv.cu (halo receive before the boundary slabs, allreduce in
// the tail of the product) and solver.cu (halo send in the tail of the fused vector update).
#pragma once

#include "solver.cuh"

namespace tfelcu {

constexpr size_t kArenaHaloCap = (size_t)1 << 20;   // doubles per parity
constexpr long long kSpinLimit = 20000000000ll;     // clock64 ticks (~10 s): a lost peer must not hang the GPU

struct Arena {
  unsigned long long red_flag[2][16];   // [parity][source rank]: sequence number of the last complete contribution
  double red_val[2][16][4];
  unsigned long long halo_flag[16];     // [source rank]
  unsigned long long pad[16];
  double halo[2][kArenaHaloCap];        // landing zone of halo values, [parity]
};

struct Peers { Arena* a[16]; };

// block b: fixed-order sum of partial array b (one warp)
struct ReduceArgs { const double* p[4]; int n[4]; };

struct HaloArgs {
  unsigned long long send_off[17], recv_off[17], land_off[16];
};

// device counters (Ctx::seq_dev, 64 entries): advanced by the kernels themselves so that the loop can live in a CUDA graph
//   [0] reductions done | [1 + q] halo sends to rank q | [17 + q] halo receives from rank q | [40] fused products launched
constexpr int kSeqSend = 1, kSeqRecv = 17, kSeqLaunch = 40;

__device__ __forceinline__ bool spin_until(const volatile unsigned long long* flag, unsigned long long seq) {
  const long long t0 = clock64();
  while (*flag < seq) {
    if (clock64() - t0 > kSpinLimit) return false;
    __nanosleep(40);
  }
  return true;
}

// everything the distributed single-reduction CG folds into its two kernels per iteration
struct DistFuse {
  int enabled;
  int me, P;
  Peers peers;
  unsigned long long* seq_dev;
  unsigned int* ctr;            // [0] CTAs of the product that are done, [1] halo of this launch unpacked (launch sequence),
                                // [2] CTAs of the vector update that are done
  HaloArgs H;
  const uint32_t* send_idx;
  double* x_rw;                 // the input vector of the product (its halo entries are written by the product itself)
  size_t halo_base;             // first halo entry of a vector
  int n_interior;               // slabs [0, n_interior) reference owned columns only
  ReduceArgs A; int k;          // tail of the product: global sums of k partial arrays ...
  double* red_out;              // ... written here, then the scalar step of the CG
  KrylovScalars* S;
};

// One CTA: my boundary values of v -> landing zones of the neighbours (+ flags). All threads of the CTA (nt of them,
// synchronised by `sync`) take part.
template <typename Sync>
__device__ __forceinline__ void p2p_halo_send(const double* v, const DistFuse& F, int tid, int nt, Sync sync) {
  for (int q = 0; q < F.P; ++q) {
    if (q == F.me) continue;
    const size_t sc = F.H.send_off[q + 1] - F.H.send_off[q];
    if (!sc) continue;
    const unsigned long long seq = F.seq_dev[kSeqSend + q] + 1;
    const int par = (int)(seq & 1ull);
    double* dst = F.peers.a[q]->halo[par] + F.H.land_off[q];
    const uint32_t* idx = F.send_idx + F.H.send_off[q];
    // eight independent (index, value) loads in flight per thread before the first remote store: one element per trip
    // costs two dependent L2 latencies each (measured: 18 us for 4097 values with 256 threads)
    for (size_t i0 = tid; i0 < sc; i0 += (size_t)8 * nt) {
      uint32_t ix[8]; double val[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { const size_t i = i0 + (size_t)u * nt; ix[u] = i < sc ? __ldg(idx + i) : 0u; }
#pragma unroll
      for (int u = 0; u < 8; ++u) { const size_t i = i0 + (size_t)u * nt; val[u] = i < sc ? __ldcg(v + ix[u]) : 0.0; }
#pragma unroll
      for (int u = 0; u < 8; ++u) { const size_t i = i0 + (size_t)u * nt; if (i < sc) dst[i] = val[u]; }
    }
    __threadfence_system();
    sync();
    if (tid == 0) {
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long*>(&F.peers.a[q]->halo_flag[F.me]) = seq;
      F.seq_dev[kSeqSend + q] = seq;
    }
    sync();
  }
}

// One CTA: values of the neighbours -> halo entries of v. Returns false when a peer did not answer in time.
template <typename Sync>
__device__ __forceinline__ bool p2p_halo_recv(double* v, const DistFuse& F, int tid, int nt, Sync sync, int* ok_s) {
  Arena* mine = F.peers.a[F.me];
  bool all_ok = true;
  for (int q = 0; q < F.P; ++q) {
    if (q == F.me) continue;
    const size_t rc = F.H.recv_off[q + 1] - F.H.recv_off[q];
    if (!rc) continue;
    const unsigned long long seq = F.seq_dev[kSeqRecv + q] + 1;
    const int par = (int)(seq & 1ull);
    sync();
    if (tid == 0) { *ok_s = spin_until(&mine->halo_flag[q], seq) ? 1 : 0; F.seq_dev[kSeqRecv + q] = seq; }
    sync();
    __threadfence_system();
    if (!*ok_s) { all_ok = false; continue; }
    const double* src = mine->halo[par] + F.H.recv_off[q];
    double* dst = v + F.halo_base + F.H.recv_off[q];
    for (size_t i0 = tid; i0 < rc; i0 += (size_t)8 * nt) {   // eight loads in flight per thread (L2: written by the peer)
      double val[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { const size_t i = i0 + (size_t)u * nt; val[u] = i < rc ? __ldcg(src + i) : 0.0; }
#pragma unroll
      for (int u = 0; u < 8; ++u) { const size_t i = i0 + (size_t)u * nt; if (i < rc) dst[i] = val[u]; }
    }
  }
  return all_ok;
}

// Warps 0 .. k-1 of one CTA (lane = thread in warp): fixed-order local sum of partial array w, contribution written into
// every rank's arena, flag, wait for everybody's contribution, sum in rank order (every rank adds the same numbers in the
// same order: bitwise identical results, so every rank takes the same branch of the stopping test). epilogue == 1: the
// scalar step of the single-reduction CG from (gamma, rr, delta) in the same launch.
template <typename Sync>
__device__ __forceinline__ void p2p_allreduce(const ReduceArgs& A, int k, const Peers& peers, int me, int P,
                                              unsigned long long* seq_dev, double* out, KrylovScalars* S, int epilogue,
                                              int tid, Sync sync) {
  const int w = tid >> 5, lane = tid & 31;
  const unsigned long long seq = seq_dev[0] + 1;   // every thread reads the counter before thread 0 advances it
  sync();
  const int par = (int)(seq & 1ull);
  if (w < k) {
    const double* p = A.p[w];
    const int n = A.n[w];
    double t = 0.0;
    for (int i0 = lane; i0 < n; i0 += 8 * 32) {   // eight loads in flight; the order of the additions stays i ascending
      double val[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { const int i = i0 + u * 32; val[u] = i < n ? __ldcg(p + i) : 0.0; }
#pragma unroll
      for (int u = 0; u < 8; ++u) t += val[u];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane < P) {
      *reinterpret_cast<volatile double*>(&peers.a[lane]->red_val[par][me][w]) = t;
      __threadfence_system();
    }
  }
  sync();
  if (w == 0 && lane < P) {
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(&peers.a[lane]->red_flag[par][me]) = seq;
  }
  if (w < k) {
    Arena* mine = peers.a[me];
    double v = 0.0;
    bool ok = true;
    if (lane < P) {
      ok = spin_until(&mine->red_flag[par][lane], seq);
      __threadfence_system();
      v = *reinterpret_cast<volatile double*>(&mine->red_val[par][lane][w]);
    }
    double t = 0.0;
    for (int q = 0; q < P; ++q) t += __shfl_sync(0xffffffffu, v, q);
    if (lane == 0) out[w] = t;
    if (!__all_sync(0xffffffffu, ok) && lane == 0 && S) S->done = -3;
  }
  if (tid == 0) seq_dev[0] = seq;
  if (epilogue == 1) {   // single-reduction CG: (gamma, rr, delta) -> alpha, beta, stopping test
    sync();
    if (tid == 0) cgs_scalar_step(S, out[0], out[1], k == 4 ? out[2] + out[3] : out[2]);
  }
}

}  // namespace tfelcu
