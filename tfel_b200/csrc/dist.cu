// Row-block distribution of the Krylov solve over the GPUs of one box (one process per GPU, NCCL over NVLink).
// A process owns contiguous ranges of the reference DOF numbering (one per component: horizontal strips of
// gen_square_mesh, k-slabs of gen_cube_mesh) and the matching rows of the reference CSR with GLOBAL column indices
// (what the owner-computes assembly produces). Set-up turns the columns into local indices [owned | halo], builds the
// halo exchange plan with the neighbours and leaves three primitives to the Krylov methods:
//   dist_halo_exchange   owned entries of a vector -> halo entries of the neighbours (pack kernel + grouped
//                        ncclSend / ncclRecv)
//   dist_reduce          per-CTA partial sums -> one value per sum on every process (fixed-order local sum +
//                        ncclAllReduce), so that every rank takes the same branch of the stopping test
//   dist_gather          owned entries -> full vector in reference numbering on every process
// The reference has no distributed mode at all (sequential MATSEQAIJ, solver.hpp:139-143).
#include "p2p.cuh"
#include "sortutil.cuh"

#include <algorithm>
#include <cstdlib>
#include <utility>

namespace tfelcu {

struct RangeTable {   // owned row segments of every rank, by value in kernel arguments (P <= 16)
  int n_ranks;
  int n_seg[16];
  unsigned long long begin[16][kMaxSeg], end[16][kMaxSeg];
};

// key of every stored entry: (owner << 32 | global column) for columns another rank owns, all ones otherwise
__global__ void k_halo_keys(const int32_t* __restrict__ colind, size_t nnz, RangeTable T, int me, uint64_t* __restrict__ keys) {
  const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  const unsigned long long g = (unsigned long long)(uint32_t)colind[e];
  int owner = -1;
  for (int q = 0; q < T.n_ranks && owner < 0; ++q)
    for (int sg = 0; sg < T.n_seg[q]; ++sg)
      if (g >= T.begin[q][sg] && g < T.end[q][sg]) { owner = q; break; }
  keys[e] = (owner == me || owner < 0) ? ~0ull : (((uint64_t)owner << 32) | g);
}

__global__ void k_localize_columns(const int32_t* __restrict__ colind, size_t nnz, RowMap R, const uint64_t* __restrict__ keys,
                                   const uint64_t* __restrict__ halo, size_t n_halo, size_t halo_base, size_t n_front,
                                   size_t front_pad, int32_t* __restrict__ out) {
  const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  const uint64_t key = keys[e];
  if (key == ~0ull) { out[e] = (int32_t)R.to_local((unsigned long long)(uint32_t)colind[e]); return; }
  size_t lo = 0, hi = n_halo;
  while (lo < hi) {
    const size_t mid = lo + (hi - lo) / 2;
    if (halo[mid] < key) lo = mid + 1; else hi = mid;
  }
  out[e] = lo < n_front ? (int32_t)((long long)lo - (long long)front_pad) : (int32_t)(halo_base + (lo - n_front));
}

__global__ void k_lower_bounds_u64(const uint64_t* __restrict__ sorted, size_t n, int n_targets, unsigned long long* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > n_targets) return;
  const uint64_t target = (uint64_t)t << 32;
  size_t lo = 0, hi = n;
  while (lo < hi) {
    const size_t mid = lo + (hi - lo) / 2;
    if (sorted[mid] < target) lo = mid + 1; else hi = mid;
  }
  out[t] = lo;
}

__global__ void k_requests_to_local(const uint64_t* __restrict__ req, size_t n, RowMap R, uint32_t* __restrict__ idx, int* __restrict__ bad) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long l = R.to_local(req[i] & 0xffffffffull);
  if (l == ~0ull) { atomicExch(bad, 1); idx[i] = 0; }
  else idx[i] = (uint32_t)l;
}

__global__ void k_pack(const double* __restrict__ v, const uint32_t* __restrict__ idx, size_t n, double* __restrict__ out,
                       const KrylovScalars* __restrict__ S) {
  if (S && S->done) return;
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = v[idx[i]];
}

__global__ void k_reduce_many(ReduceArgs A, double* __restrict__ out) {
  const double* p = A.p[blockIdx.x];
  const int n = A.n[blockIdx.x];
  double t = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) t += p[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (threadIdx.x == 0) out[blockIdx.x] = t;
}

// ------------------------------------------------------------------------------------------
// peer-to-peer arena: remote stores + sequence flags over NVLink (one process per GPU, CUDA IPC)
// ------------------------------------------------------------------------------------------

// stand-alone allreduce of k partial arrays (warp w: array w); see p2p_allreduce
__global__ void k_p2p_allreduce(ReduceArgs A, int k, Peers peers, int me, int P, unsigned long long* __restrict__ seq_dev,
                                double* __restrict__ out, KrylovScalars* __restrict__ S, int epilogue, int* __restrict__ err) {
  p2p_allreduce(A, k, peers, me, P, seq_dev, out, S, epilogue, (int)threadIdx.x, [] { __syncthreads(); }, err);
}

// CTA q: my boundary values -> landing zone of rank q; values of rank q -> halo entries of v (words that carry their own
// sequence number: p2p.cuh)
__global__ void __launch_bounds__(256)
k_p2p_halo(double* __restrict__ v, const uint32_t* __restrict__ send_idx, HaloArgs H, Peers peers,
           int me, unsigned long long* __restrict__ seq_dev, KrylovScalars* __restrict__ S, int do_send, int do_recv,
           int* __restrict__ err) {
  if (S && S->done) return;
  const int q = blockIdx.x;
  if (q == me) return;
  const unsigned long long seq_s = seq_dev[kSeqSend + q] + 1, seq_r = seq_dev[kSeqRecv + q] + 1;
  __syncthreads();
  const size_t sc = H.send_off[q + 1] - H.send_off[q], rc = H.recv_off[q + 1] - H.recv_off[q];
  if (sc && do_send) {
    const int par = (int)(seq_s & 1ull);
    unsigned long long* dst = peers.a[q]->halo[par] + 2 * H.land_off[q];
    const uint32_t* idx = send_idx + H.send_off[q];
    for (size_t i0 = threadIdx.x; i0 < sc; i0 += (size_t)8 * blockDim.x) {   // see p2p_halo_send
      uint32_t ix[8]; double val[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { const size_t i = i0 + (size_t)u * blockDim.x; ix[u] = i < sc ? __ldg(idx + i) : 0u; }
#pragma unroll
      for (int u = 0; u < 8; ++u) { const size_t i = i0 + (size_t)u * blockDim.x; val[u] = i < sc ? v[ix[u]] : 0.0; }
#pragma unroll
      for (int u = 0; u < 8; ++u) { const size_t i = i0 + (size_t)u * blockDim.x; if (i < sc) ll_store(dst + 2 * i, val[u], (unsigned int)seq_s); }
    }
    if (threadIdx.x == 0) seq_dev[kSeqSend + q] = seq_s;
  }
  if (rc && do_recv) {
    const int par = (int)(seq_r & 1ull);
    const unsigned long long* src = peers.a[me]->halo[par] + 2 * H.recv_off[q];
    double* dst = v + H.recv_pos[q];
    bool ok = true;
    for (size_t i0 = threadIdx.x; i0 < rc; i0 += (size_t)8 * blockDim.x) {
      const size_t left = (rc - i0 + blockDim.x - 1) / blockDim.x;
      ok = ll_load_batch(src + 2 * i0, (size_t)blockDim.x, left < 8 ? (int)left : 8, (unsigned int)seq_r,
                         [&](int u, double val) { dst[i0 + (size_t)u * blockDim.x] = val; }) && ok;
    }
    if (!ok) { atomicExch(err, 1); if (S) S->done = -3; }   // the counters advance all the same: the ranks stay in step
    if (threadIdx.x == 0) seq_dev[kSeqRecv + q] = seq_r;
  }
}

void p2p_teardown(Ctx* ctx) {
  for (int q = 0; q < 16; ++q)
    if (ctx->peer_arena[q] && q != ctx->rank) cudaIpcCloseMemHandle(ctx->peer_arena[q]);
  if (ctx->arena) cudaFree(ctx->arena);
  if (ctx->seq_dev) cudaFree(ctx->seq_dev);
  if (ctx->p2p_err) cudaFree(ctx->p2p_err);
  ctx->arena = nullptr; ctx->seq_dev = nullptr; ctx->p2p_err = nullptr; ctx->p2p = false;
  for (int q = 0; q < 16; ++q) ctx->peer_arena[q] = nullptr;
  cudaGetLastError();
}

void p2p_setup(Ctx* ctx) {
  const int P = ctx->n_ranks, me = ctx->rank;
  if (P < 2 || P > 16 || std::getenv("TFELCU_NO_P2P")) return;
  int ok = 1;
  cudaIpcMemHandle_t mine_h;
  std::memset(&mine_h, 0, sizeof(mine_h));
  if (cudaMalloc(&ctx->arena, sizeof(Arena)) != cudaSuccess) { ok = 0; ctx->arena = nullptr; cudaGetLastError(); }
  if (ok && cudaMalloc(&ctx->seq_dev, 64 * sizeof(unsigned long long)) != cudaSuccess) { ok = 0; ctx->seq_dev = nullptr; cudaGetLastError(); }
  if (ok && cudaMalloc(&ctx->p2p_err, sizeof(int)) != cudaSuccess) { ok = 0; ctx->p2p_err = nullptr; cudaGetLastError(); }
  if (ok) {
    TFELCU_CK(cudaMemsetAsync(ctx->p2p_err, 0, sizeof(int), ctx->stream));
    TFELCU_CK(cudaMemsetAsync(ctx->arena, 0, sizeof(Arena), ctx->stream));
    TFELCU_CK(cudaMemsetAsync(ctx->seq_dev, 0, 64 * sizeof(unsigned long long), ctx->stream));
    ctx->sync();
    if (cudaIpcGetMemHandle(&mine_h, ctx->arena) != cudaSuccess) { ok = 0; cudaGetLastError(); }
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  DevBuf<char> d_mine(64), d_all((size_t)P * 64);
  ctx->upload(d_mine.p, reinterpret_cast<const char*>(&mine_h), 64);
  TFELCU_NCCL(ncclAllGather(d_mine.p, d_all.p, 64, ncclChar, ctx->comm, ctx->stream));
  std::vector<char> h_all((size_t)P * 64);
  ctx->download(h_all.data(), d_all.p, h_all.size());
  auto agree = [&](int flag) {   // 1 only when every rank says 1
    DevBuf<int> f(1);
    ctx->upload(f.p, &flag, 1);
    TFELCU_NCCL(ncclAllReduce(f.p, f.p, 1, ncclInt, ncclMin, ctx->comm, ctx->stream));
    int out = 0; ctx->download(&out, f.p, 1);
    return out;
  };
  ok = agree(ok);
  if (ok) {
    for (int q = 0; q < P; ++q) {
      if (q == me) { ctx->peer_arena[q] = ctx->arena; continue; }
      cudaIpcMemHandle_t hq;
      std::memcpy(&hq, &h_all[(size_t)q * 64], 64);
      if (cudaIpcOpenMemHandle(&ctx->peer_arena[q], hq, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        ok = 0; ctx->peer_arena[q] = nullptr; cudaGetLastError();
      }
    }
    ok = agree(ok);
  }
  if (!ok) { p2p_teardown(ctx); return; }
  ctx->p2p = true;
}

static Peers peers_of(Ctx* ctx);

// diagnostic: average time of one arena allreduce of 2 values (CUDA events, `reps` back-to-back launches)
double p2p_bench_allreduce(Ctx* ctx, int reps) {
  TFELCU_REQUIRE(ctx->p2p, "peer-to-peer arena is not enabled");
  DevBuf<double> part(64), out(4);
  ctx->zero(part.p, 64);
  ReduceArgs A;
  for (int i = 0; i < 4; ++i) { A.p[i] = part.p; A.n[i] = 32; }
  cudaEvent_t e0, e1;
  TFELCU_CK(cudaEventCreate(&e0)); TFELCU_CK(cudaEventCreate(&e1));
  for (int r = 0; r < 10; ++r)
    k_p2p_allreduce<<<1, 64, 0, ctx->stream>>>(A, 2, peers_of(ctx), ctx->rank, ctx->n_ranks, ctx->seq_dev, out.p, nullptr, 0, ctx->p2p_err);
  TFELCU_CK(cudaEventRecord(e0, ctx->stream));
  for (int r = 0; r < reps; ++r)
    k_p2p_allreduce<<<1, 64, 0, ctx->stream>>>(A, 2, peers_of(ctx), ctx->rank, ctx->n_ranks, ctx->seq_dev, out.p, nullptr, 0, ctx->p2p_err);
  TFELCU_CK(cudaEventRecord(e1, ctx->stream));
  TFELCU_CK(cudaEventSynchronize(e1));
  float ms = 0.f; TFELCU_CK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return (double)ms * 1e3 / reps;
}

static Peers peers_of(Ctx* ctx) {
  Peers p;
  for (int q = 0; q < 16; ++q) p.a[q] = static_cast<Arena*>(ctx->peer_arena[q]);
  return p;
}

void dist_setup(tfelcu_solver* s) {
  Ctx* ctx = s->ctx;
  DistPlan& D = *s->dist;
  const int P = ctx->n_ranks, me = ctx->rank;
  TFELCU_REQUIRE(ctx->comm != nullptr, "distributed solve needs a context created with tfelcu_ctx_create_distributed");
  TFELCU_REQUIRE(P <= 16, "at most 16 ranks");
  const RowMap& R = D.rows;
  const size_t n_local = R.n_local();
  // 1. everybody's row segments
  constexpr int W = 2 * kMaxSeg + 1;
  DevBuf<unsigned long long> mine(W), all((size_t)P * W);
  unsigned long long h_mine[W];
  for (int sg = 0; sg < kMaxSeg; ++sg) { h_mine[sg] = R.g_begin[sg]; h_mine[kMaxSeg + sg] = R.g_end[sg]; }
  h_mine[2 * kMaxSeg] = (unsigned long long)R.n;
  ctx->upload(mine.p, h_mine, W);
  TFELCU_NCCL(ncclAllGather(mine.p, all.p, W, ncclUint64, ctx->comm, ctx->stream));
  std::vector<unsigned long long> h_all((size_t)P * W);
  ctx->download(h_all.data(), all.p, h_all.size());
  RangeTable T;
  T.n_ranks = P;
  D.all_rows.assign(P, R);
  std::vector<std::pair<unsigned long long, unsigned long long>> cover;
  for (int q = 0; q < P; ++q) {
    RowMap& Q = D.all_rows[q];
    Q.n = (int)h_all[(size_t)q * W + 2 * kMaxSeg];
    T.n_seg[q] = Q.n;
    unsigned long long off = 0;
    for (int sg = 0; sg < Q.n; ++sg) {
      T.begin[q][sg] = Q.g_begin[sg] = h_all[(size_t)q * W + sg];
      T.end[q][sg] = Q.g_end[sg] = h_all[(size_t)q * W + kMaxSeg + sg];
      Q.local_off[sg] = off;
      off += Q.g_end[sg] - Q.g_begin[sg];
      if (Q.g_end[sg] > Q.g_begin[sg]) cover.push_back({Q.g_begin[sg], Q.g_end[sg]});
    }
    Q.local_off[Q.n] = off;
  }
  std::sort(cover.begin(), cover.end());
  unsigned long long at = 0;
  for (const auto& c : cover) {
    TFELCU_REQUIRE(c.first == at, "the row segments of the ranks must tile [0, n_rows) without gaps or overlaps");
    at = c.second;
  }
  TFELCU_REQUIRE(at == R.n_global, "the row segments of the ranks must cover every row");
  // 2. halo columns, sorted by (owner, global column)
  const size_t nnz = s->nnz;
  DevBuf<uint64_t> keys(nnz), sorted(nnz), halo(nnz);
  if (nnz) {
    k_halo_keys<<<grid_for(nnz, 256), 256, 0, ctx->stream>>>(s->colind, nnz, T, me, keys.p);
    ctx->count(); TFELCU_LAUNCH_CHECK();
  }
  sort_keys<uint64_t>(ctx, keys.p, sorted.p, nnz, 0, 64);
  size_t n_halo = unique_keys<uint64_t>(ctx, sorted.p, halo.p, nnz);
  if (n_halo) {
    uint64_t last = 0; ctx->download(&last, halo.p + (n_halo - 1), 1);
    if (last == ~0ull) --n_halo;
  }
  // the halo entries of a vector start on a 128-byte line of their own: the product reads them through the read-only
  // path after another CTA of the same launch has written them, which is only safe for lines no earlier read of the launch
  // (of owned entries) can have brought into an L1
  DevBuf<unsigned long long> bounds((size_t)P + 1);
  k_lower_bounds_u64<<<1, 32, 0, ctx->stream>>>(halo.p, n_halo, P, bounds.p);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  std::vector<unsigned long long> h_bounds((size_t)P + 1);
  ctx->download(h_bounds.data(), bounds.p, (size_t)P + 1);
  D.recv_off.assign(h_bounds.begin(), h_bounds.end());
  // what the lower ranks own goes in front of the owned entries, the rest behind them: with the rows of the ranks in
  // ascending order (the usual block partition) every halo column then lies within a few thousand entries of the rows
  // that use it, and the product keeps its 16-bit column deltas on every rank
  D.n_front = (size_t)D.recv_off[me];
  D.front_pad = (D.n_front + 15) & ~(size_t)15;
  D.halo_base = (n_local + 15) & ~(size_t)15;
  D.n_halo = (D.halo_base - n_local) + (n_halo - D.n_front);   // what a vector holds behind its owned entries: padding + halo
  TFELCU_REQUIRE(D.halo_base + n_halo < 0x7fffffffull, "too many local columns");
  D.recv_pos.assign((size_t)P, 0);
  for (int q = 0; q < P; ++q)
    D.recv_pos[q] = q < me ? (long long)D.recv_off[q] - (long long)D.front_pad
                           : (long long)D.halo_base + (long long)(D.recv_off[q] - D.n_front);
  D.colind_local.alloc(nnz);
  if (nnz) {
    k_localize_columns<<<grid_for(nnz, 256), 256, 0, ctx->stream>>>(s->colind, nnz, R, keys.p, halo.p, n_halo, D.halo_base, D.n_front,
                                                                    D.front_pad, D.colind_local.p);
    ctx->count(); TFELCU_LAUNCH_CHECK();
  }
  // 3. what every rank receives from every rank
  DevBuf<unsigned long long> cnt_mine(P), cnt_all((size_t)P * P);
  std::vector<unsigned long long> h_cnt(P);
  for (int q = 0; q < P; ++q) h_cnt[q] = D.recv_off[q + 1] - D.recv_off[q];
  ctx->upload(cnt_mine.p, h_cnt.data(), (size_t)P);
  TFELCU_NCCL(ncclAllGather(cnt_mine.p, cnt_all.p, (size_t)P, ncclUint64, ctx->comm, ctx->stream));
  std::vector<unsigned long long> h_cnt_all((size_t)P * P);
  ctx->download(h_cnt_all.data(), cnt_all.p, h_cnt_all.size());
  D.send_off.assign((size_t)P + 1, 0);
  for (int q = 0; q < P; ++q) D.send_off[q + 1] = D.send_off[q] + (size_t)h_cnt_all[(size_t)q * P + me];
  // where my values land at rank q (its receive offset for me) and whether every rank's halo fits the arena
  D.land_off.assign((size_t)P, 0);
  D.p2p_halo = ctx->p2p;
  for (int q = 0; q < P; ++q) {
    size_t off = 0, total = 0;
    for (int r = 0; r < P; ++r) {
      if (r == me) off = total;
      total += (size_t)h_cnt_all[(size_t)q * P + r];
    }
    D.land_off[q] = off;
    if (total > kArenaHaloCap) D.p2p_halo = false;
    // a parity buffer is reused every second exchange; what makes that safe is that every send to a rank is paired with
    // a receive from it (the sender cannot run two exchanges ahead). One-directional couplings take the NCCL path.
    const bool sends = h_cnt_all[(size_t)q * P + me] != 0, receives = h_cnt[q] != 0;
    if (q != me && sends != receives) D.p2p_halo = false;
  }
  {   // every rank must take the same path
    DevBuf<int> f(1);
    int flag = D.p2p_halo ? 1 : 0;
    ctx->upload(f.p, &flag, 1);
    TFELCU_NCCL(ncclAllReduce(f.p, f.p, 1, ncclInt, ncclMin, ctx->comm, ctx->stream));
    ctx->download(&flag, f.p, 1);
    D.p2p_halo = flag != 0;
  }
  const size_t n_send = D.send_off[P];
  // 4. tell the owners which of their rows we need
  DevBuf<uint64_t> req(n_send);
  TFELCU_NCCL(ncclGroupStart());
  for (int q = 0; q < P; ++q) {
    if (q == me) continue;
    const size_t rc = D.recv_off[q + 1] - D.recv_off[q], sc = D.send_off[q + 1] - D.send_off[q];
    if (rc) TFELCU_NCCL(ncclSend(halo.p + D.recv_off[q], rc, ncclUint64, q, ctx->comm, ctx->stream));
    if (sc) TFELCU_NCCL(ncclRecv(req.p + D.send_off[q], sc, ncclUint64, q, ctx->comm, ctx->stream));
  }
  TFELCU_NCCL(ncclGroupEnd());
  D.send_idx.alloc(n_send);
  D.send_buf.alloc(n_send);
  if (n_send) {
    DevBuf<int> bad(1);
    ctx->zero(bad.p, 1);
    k_requests_to_local<<<grid_for(n_send, 256), 256, 0, ctx->stream>>>(req.p, n_send, R, D.send_idx.p, bad.p);
    ctx->count(); TFELCU_LAUNCH_CHECK();
    int h_bad = 0; ctx->download(&h_bad, bad.p, 1);
    TFELCU_REQUIRE(!h_bad, "halo plan: a neighbour asked for a row this rank does not own");
  }
  D.red.alloc(8);
  ctx->sync();
}

// owned entries of v (length n_local + n_halo) -> halo entries of the neighbours' copies
// mode: 3 send + receive (default), 1 send only, 2 receive only (the fused loop of solver.cu opens with a send and, when it
// stops at maxits, closes with a receive; peer-to-peer arena only)
void dist_halo_exchange(tfelcu_solver* s, double* v, const KrylovScalars* S, cudaStream_t stream, int mode) {
  Ctx* ctx = s->ctx;
  DistPlan& D = *s->dist;
  const int P = ctx->n_ranks, me = ctx->rank;
  const size_t n_send = D.send_off[P];
  if (!stream) stream = ctx->stream;
  if (D.p2p_halo) {
    HaloArgs H;
    for (int q = 0; q <= P; ++q) { H.send_off[q] = D.send_off[q]; H.recv_off[q] = D.recv_off[q]; }
    for (int q = 0; q < P; ++q) H.recv_pos[q] = D.recv_pos[q];
    for (int q = 0; q < P; ++q) H.land_off[q] = D.land_off[q];
    k_p2p_halo<<<P, 256, 0, stream>>>(v, D.send_idx.p, H, peers_of(ctx), me, ctx->seq_dev,
                                     const_cast<KrylovScalars*>(S), mode & 1, (mode >> 1) & 1, ctx->p2p_err);
    ctx->count(); TFELCU_LAUNCH_CHECK();
    return;
  }
  TFELCU_REQUIRE(mode == 3, "one-sided halo steps need the peer-to-peer arena");
  if (n_send) {
    k_pack<<<grid_for(n_send, 256), 256, 0, stream>>>(v, D.send_idx.p, n_send, D.send_buf.p, S);
    ctx->count(); TFELCU_LAUNCH_CHECK();
  }
  TFELCU_NCCL(ncclGroupStart());
  for (int q = 0; q < P; ++q) {
    if (q == me) continue;
    const size_t rc = D.recv_off[q + 1] - D.recv_off[q], sc = D.send_off[q + 1] - D.send_off[q];
    if (sc) TFELCU_NCCL(ncclSend(D.send_buf.p + D.send_off[q], sc, ncclDouble, q, ctx->comm, stream));
    if (rc) TFELCU_NCCL(ncclRecv(v + D.recv_pos[q], rc, ncclDouble, q, ctx->comm, stream));
  }
  TFELCU_NCCL(ncclGroupEnd());
}

// out_slot .. out_slot + k - 1 of the reduction buffer <- global sums of k partial arrays; returns the device address
__global__ void k_cgs_scalars_from(const double* __restrict__ v, int k, KrylovScalars* __restrict__ S) {
  if (threadIdx.x == 0 && blockIdx.x == 0) cgs_scalar_step(S, v[0], v[1], k == 4 ? v[2] + v[3] : v[2]);
}

const double* dist_reduce(tfelcu_solver* s, const double* const* partial, const int* n, int k, int out_slot, int epilogue) {
  Ctx* ctx = s->ctx;
  DistPlan& D = *s->dist;
  ReduceArgs A;
  for (int i = 0; i < 4; ++i) { A.p[i] = i < k ? partial[i] : nullptr; A.n[i] = i < k ? n[i] : 0; }
  double* out = D.red.p + out_slot;
  if (ctx->p2p) {
    k_p2p_allreduce<<<1, 32 * k, 0, ctx->stream>>>(A, k, peers_of(ctx), ctx->rank, ctx->n_ranks, ctx->seq_dev, out,
                                                  s->scal.p, epilogue, ctx->p2p_err);
    ctx->count(); TFELCU_LAUNCH_CHECK();
    return out;
  }
  k_reduce_many<<<k, 32, 0, ctx->stream>>>(A, out);
  ctx->count(); TFELCU_LAUNCH_CHECK();
  TFELCU_NCCL(ncclAllReduce(out, out, (size_t)k, ncclDouble, ncclSum, ctx->comm, ctx->stream));
  if (epilogue == 1) {
    k_cgs_scalars_from<<<1, 32, 0, ctx->stream>>>(out, k, s->scal.p);
    ctx->count(); TFELCU_LAUNCH_CHECK();
  }
  return out;
}

// after a distributed solve: did a peer fail to answer? (the Krylov kernels stop at once when it happens: done = -3)
bool dist_peer_failed(tfelcu_solver* s) {
  Ctx* ctx = s->ctx;
  if (!ctx->p2p_err) return false;
  int h = 0;
  ctx->download(&h, ctx->p2p_err, 1);
  if (h) { ctx->p2p_failed = true; ctx->zero(ctx->p2p_err, 1); }
  return h != 0;
}

// x_global (device, n_rows of the global system) <- owned entries of every rank, reference numbering
void dist_gather(tfelcu_solver* s, const double* x_local, double* x_global) {
  Ctx* ctx = s->ctx;
  DistPlan& D = *s->dist;
  const int P = ctx->n_ranks;
  TFELCU_NCCL(ncclGroupStart());
  for (int q = 0; q < P; ++q) {
    const RowMap& Q = D.all_rows[q];
    for (int c = 0; c < Q.n; ++c) {
      const size_t cnt = (size_t)(Q.g_end[c] - Q.g_begin[c]);
      if (!cnt) continue;
      double* dst = x_global + Q.g_begin[c];
      const double* src = q == ctx->rank ? x_local + Q.local_off[c] : dst;
      TFELCU_NCCL(ncclBroadcast(src, dst, cnt, ncclDouble, q, ctx->comm, ctx->stream));
    }
  }
  TFELCU_NCCL(ncclGroupEnd());
}

}  // namespace tfelcu
