// Peer-to-peer arena of the distributed Krylov loop (one process per GPU, CUDA IPC over NVLink): device-side pieces of
// the halo exchange and allreduce kernels of dist.cu.
// (Tried and removed: folding the halo receive and the allreduce into the product kernel and the halo send into the vector
// update -- two launches per iteration instead of four. At 8 GPUs it was slower, 137 vs 112 us per iteration: the exposed
// time is the latency of the peer-to-peer steps themselves, which the fused kernels serialise behind their last CTA, not
// the launches. What did help is in this file: words that carry their own sequence number instead of value / fence / flag.)
#pragma once

#include "solver.cuh"

namespace tfelcu {

constexpr size_t kArenaHaloCap = (size_t)1 << 20;   // doubles per parity
constexpr long long kSpinLimit = 20000000000ll;     // clock64 ticks (~10 s): a lost peer must not hang the GPU

// Every remote value travels as two 8-byte words (32 data bits | 32-bit sequence number): an aligned 8-byte store is
// atomic, so a word whose sequence number matches carries valid data and no separate flag, no system-scope fence and no
// second NVLink round trip are needed (the "LL" protocol of NCCL). Measured at 8 GPUs with value / fence / flag / fence:
// 22 us per halo exchange and 25 us per reduction, most of it fence round trips.
struct Arena {
  unsigned long long red[2][16][4][2];        // [parity][source rank][value][half]
  unsigned long long halo[2][2 * kArenaHaloCap];   // landing zone of halo values, [parity][2 words per value]
};

__device__ __forceinline__ void ll_store(unsigned long long* dst, double v, unsigned int seq) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  const unsigned long long lo = (b & 0xffffffffull) | ((unsigned long long)seq << 32);
  const unsigned long long hi = (b >> 32) | ((unsigned long long)seq << 32);
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(lo), "l"(hi) : "memory");
}

// n (<= 8) values at src[2 * stride * u], u = 0 .. n - 1: all loads are issued before the first check (one exposed latency
// per batch instead of one per value); words that are not there yet are re-read one by one
template <typename Store>
__device__ __forceinline__ bool ll_load_batch(const unsigned long long* src, size_t stride, int n, unsigned int seq, Store store) {
  unsigned long long lo[8], hi[8];
#pragma unroll
  for (int u = 0; u < 8; ++u)
    if (u < n) asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo[u]), "=l"(hi[u]) : "l"(src + 2 * stride * u) : "memory");
  bool ok = true;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    if (u >= n) continue;
    if ((unsigned int)(lo[u] >> 32) == seq && (unsigned int)(hi[u] >> 32) == seq) {
      store(u, __longlong_as_double((long long)((lo[u] & 0xffffffffull) | (hi[u] << 32))));
    } else {
      const long long t0 = clock64();
      int spins = 0;
      unsigned long long a, b;
      bool got = true;
      while (true) {
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(src + 2 * stride * u) : "memory");
        if ((unsigned int)(a >> 32) == seq && (unsigned int)(b >> 32) == seq) break;
        if ((++spins & 255) == 0 && clock64() - t0 > kSpinLimit) { got = false; break; }
      }
      ok = ok && got;
      store(u, got ? __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32))) : 0.0);
    }
  }
  return ok;
}

// both halves with sequence number seq -> value; false when they did not arrive in time
__device__ __forceinline__ bool ll_load(const unsigned long long* src, unsigned int seq, double* out) {
  unsigned long long lo, hi;
  const long long t0 = clock64();
  int spins = 0;
  while (true) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(src) : "memory");
    if ((unsigned int)(lo >> 32) == seq && (unsigned int)(hi >> 32) == seq) break;
    if ((++spins & 255) == 0 && clock64() - t0 > kSpinLimit) return false;
  }
  *out = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
  return true;
}

struct Peers { Arena* a[16]; };

// block b: fixed-order sum of partial array b (one warp)
struct ReduceArgs { const double* p[4]; int n[4]; };

struct HaloArgs {
  unsigned long long send_off[17], recv_off[17], land_off[16];
  long long recv_pos[16];   // where the entries received from rank q start in a vector (negative: in front of the owned entries)
};

// device counters (Ctx::seq_dev, 64 entries): advanced by the kernels themselves so that the loop can live in a CUDA graph
//   [0] reductions done | [1 + q] halo sends to rank q | [17 + q] halo receives from rank q
constexpr int kSeqSend = 1, kSeqRecv = 17;

// Warps 0 .. k-1 of one CTA (lane = thread in warp): fixed-order local sum of partial array w, contribution written into
// every rank's arena, every rank's contribution awaited, sum in rank order (every rank adds the same numbers in the same
// order: bitwise identical results, so every rank takes the same branch of the stopping test). epilogue == 1: the scalar
// step of the single-reduction CG from (gamma, rr, delta) in the same launch.
template <typename Sync>
__device__ __forceinline__ void p2p_allreduce(const ReduceArgs& A, int k, const Peers& peers, int me, int P,
                                              unsigned long long* seq_dev, double* out, KrylovScalars* S, int epilogue,
                                              int tid, Sync sync, int* err) {
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
    if (lane < P) ll_store(&peers.a[lane]->red[par][me][w][0], t, (unsigned int)seq);
    Arena* mine = peers.a[me];
    double v = 0.0;
    bool ok = true;
    if (lane < P) ok = ll_load(&mine->red[par][lane][w][0], (unsigned int)seq, &v);
    double sum = 0.0;
    for (int q = 0; q < P; ++q) sum += __shfl_sync(0xffffffffu, v, q);
    if (lane == 0) out[w] = sum;
    if (!__all_sync(0xffffffffu, ok) && lane == 0) { atomicExch(err, 1); if (S) S->done = -3; }
  }
  if (tid == 0) seq_dev[0] = seq;
  if (epilogue == 1) {   // single-reduction CG: (gamma, rr, delta) -> alpha, beta, stopping test
    sync();
    if (tid == 0) cgs_scalar_step(S, out[0], out[1], k == 4 ? out[2] + out[3] : out[2]);
  }
}

}  // namespace tfelcu
