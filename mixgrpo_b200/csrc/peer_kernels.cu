// Fused collective + compute over NVLink peer memory for the path's only exchange step.
//
//   reward all-gather (TR:332-338, TR:417-425) + group-relative advantages (TR:439-501)   -> ONE launch
//   all-reduce(AVG) of the (loss, policy, kl, clip_frac) logging sums (TR:586-600)        -> ONE launch
//
// TR = /root/reference/fastvideo/train_grpo_flux.py.  The messages are tens to hundreds of bytes, i.e. pure latency:
// a NCCL all_gather (launch + protocol, ~10-20 us) followed by a separate advantage kernel is replaced by one
// single-CTA kernel that
//   1. PUSHES this rank's rewards straight into every peer's "region" over NVLink/NVSwitch (the regions are
//      cudaMalloc'ed once per rank and mapped into every peer with CUDA IPC).  Every float travels as ONE 64-bit word
//      { call number : 32 | float bits : 32 } written with a single st.relaxed.sys.u64 — a scalar 64-bit store is
//      atomic, so value and "this is call s" arrive together: no flag, no fence, no second round trip (the idea of
//      NCCL's LL protocol, applied to a one-shot exchange);
//   2. polls (ld.relaxed.sys.u64, bounded by a timeout) its OWN region until every word carries this call's number;
//   3. computes the advantages from shared memory with exactly the arithmetic of mg::group_adv_kernel.
// One-way NVLink latency is the whole cost.  No host involvement, no NCCL, no extra launch; the call counter lives in
// the region, so the kernel is CUDA-graph capturable (pointers and arguments never change between replays).  Two
// parity-alternating buffers make back-to-back calls safe without a handshake: a rank can only reach call k+2 after
// it has received every peer's words of call k+1, which a peer sends after its call-k kernel (the last reader of buffer
// k&1) has completed; a word of call k-2 still sitting in the buffer carries the wrong number and is never mistaken.
//
// Region layout (per rank, identical on all ranks):
//   [0,256)     header: u32 seq[2] (calls completed per channel), u32 status (1 = a wait timed out)
//   [256, ...)  channel 0: u64 [parity 2][world][cap]      (reward matrices, cap >= n_models*local_B)
//   then        channel 1: u64 [parity 2][world][kRedCap]  (logging sums)
#include "common.cuh"

namespace mg {

constexpr int kMaxWorld = MIXGRPO_PEER_MAX_WORLD;
constexpr int kRedCap = 256;
constexpr int kPeerThreads = 256;
constexpr int kPeerWarps = kPeerThreads / 32;
constexpr long long kHdrBytes = 256;

struct PeerArgs {
  char* region[kMaxWorld];      // region[p] = rank p's region as mapped in THIS process (own region at [rank])
  int rank, world;
  long long cap;                // words per rank in channel 0
  unsigned long long timeout_ns;
};

__host__ __device__ inline long long ch_data_offset(int ch, int world, long long cap) {
  return kHdrBytes + (ch == 0 ? 0 : 2ll * world * cap * (long long)sizeof(unsigned long long));
}
__host__ inline long long region_bytes(int world, long long cap) {
  return ch_data_offset(1, world, cap) + 2ll * world * kRedCap * (long long)sizeof(unsigned long long);
}

__device__ __forceinline__ unsigned long long* word_ptr(char* region, int ch, int par, int p, int world, long long cap) {
  const long long slot = ch == 0 ? cap : kRedCap;
  return reinterpret_cast<unsigned long long*>(region + ch_data_offset(ch, world, cap)) + ((long long)par * world + p) * slot;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct PeerCall {
  uint32_t seq;     // this call's number (calls completed + 1)
  int par;          // buffer parity
};

// Start a call: read the call counter and push `count` floats from src into slot [rank] of every rank's buffer.
// All kPeerThreads threads.  s_bcast: two shared words ([1] is the "no timeout so far" flag).
__device__ __forceinline__ PeerCall peer_push(const PeerArgs& a, int ch, const float* src, int count, uint32_t* s_bcast) {
  uint32_t* hdr = reinterpret_cast<uint32_t*>(a.region[a.rank]);
  if (threadIdx.x == 0) { s_bcast[0] = *reinterpret_cast<volatile uint32_t*>(hdr + ch) + 1u; s_bcast[1] = 1u; }
  __syncthreads();
  PeerCall c;
  c.seq = s_bcast[0];
  c.par = (int)(c.seq & 1u);
  for (int i = threadIdx.x; i < count; i += kPeerThreads) {
    const unsigned long long w = ((unsigned long long)c.seq << 32) | (unsigned long long)__float_as_uint(src[i]);
    for (int q = 1; q <= a.world; ++q) {
      const int p = (a.rank + q) % a.world;                       // peers first (their latency is the longer one), ourselves last
      st_relaxed_sys(word_ptr(a.region[p], ch, c.par, a.rank, a.world, a.cap) + i, w);
    }
  }
  return c;
}

// Wait for word i of rank q's slot in OUR region to carry this call's number and return its float; NaN after a timeout.
__device__ __forceinline__ float peer_poll(const PeerArgs& a, int ch, const PeerCall& c, int q, int i, uint32_t* s_bcast) {
  const unsigned long long* wp = word_ptr(a.region[a.rank], ch, c.par, q, a.world, a.cap) + i;
  unsigned long long w = ld_relaxed_sys(wp);
  if ((uint32_t)(w >> 32) != c.seq) {
    const unsigned long long t0 = global_ns();
    unsigned spins = 0;
    while ((uint32_t)((w = ld_relaxed_sys(wp)) >> 32) != c.seq) {
      if ((++spins & 0xffu) == 0 && (*reinterpret_cast<volatile uint32_t*>(s_bcast + 1) == 0u ||
                                    (a.timeout_ns && global_ns() - t0 > a.timeout_ns))) {
        *reinterpret_cast<volatile uint32_t*>(s_bcast + 1) = 0u;  // tell the other pollers of this CTA to give up too
        reinterpret_cast<volatile uint32_t*>(a.region[a.rank])[2] = 1u;   // status: a wait timed out
        return __int_as_float(0x7fc00000);
      }
    }
  }
  return __uint_as_float((uint32_t)w);
}

__device__ __forceinline__ void peer_commit(const PeerArgs& a, int ch, const PeerCall& c) {   // after the last poll
  __syncthreads();
  if (threadIdx.x == 0) *reinterpret_cast<volatile uint32_t*>(reinterpret_cast<uint32_t*>(a.region[a.rank]) + ch) = c.seq;
}

// ------------------------------------------------------------------ gather + advantages
struct GatherAdvParams {
  PeerArgs peer;
  const float* rewards;      // [n_models, local_B] this rank's rewards
  const float* weights;      // [n_models] or nullptr (single model, stored unweighted, TR:489)
  float* gathered_out;       // [n_models, world*local_B] rank-major columns (torch.cat order of TR:338) or nullptr
  float* adv_out;            // [local_B]
  int n_models, local_B, G, trim, mode;
};

// mean and (Bessel std + 1e-8) of r[0..G) without its `trim` smallest members (stable rank), by one warp: the
// arithmetic of mg::group_adv_kernel (csrc/grpo_kernels.cu) — fp64 sums of <= a few thousand numbers rounded once.
__device__ __forceinline__ void warp_group_stats(const float* r, int G, int trim, int lane, float* out2) {
  const int kept_n = G - trim;
  double sum = 0.0;
  for (int i = lane; i < G; i += 32) {
    const float ri = r[i];
    bool keep = true;
    if (trim > 0) {
      int rank = 0;
      for (int j = 0; j < G; ++j) rank += (r[j] < ri) || (r[j] == ri && j < i);
      keep = rank >= trim;
    }
    if (keep) sum += (double)ri;
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const double mean_d = sum / (double)kept_n;
  double sq = 0.0;
  for (int i = lane; i < G; i += 32) {
    const float ri = r[i];
    bool keep = true;
    if (trim > 0) {
      int rank = 0;
      for (int j = 0; j < G; ++j) rank += (r[j] < ri) || (r[j] == ri && j < i);
      keep = rank >= trim;
    }
    if (keep) { const double d = (double)ri - mean_d; sq += d * d; }
  }
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if (lane == 0) {
    out2[0] = (float)mean_d;
    out2[1] = __fadd_rn((float)sqrt(sq / (double)(kept_n - 1)), 1e-8f);
  }
}

__global__ void __launch_bounds__(kPeerThreads) peer_gather_adv_kernel(const __grid_constant__ GatherAdvParams p) {
  pdl_prologue();
  extern __shared__ float s_dyn[];
  __shared__ uint32_t s_bcast[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = p.peer.world, LB = p.local_B, M = p.n_models;
  const int NB = W * LB, count = M * LB;
  float* s_g = s_dyn;                         // [M][NB] gathered rewards, columns rank-major
  float* s_stat = s_dyn + (size_t)M * NB;     // [groups][M][2]

  const PeerCall call = peer_push(p.peer, 0, p.rewards, count, s_bcast);
  for (int idx = tid; idx < W * count; idx += kPeerThreads) {
    const int q = idx / count, r = idx - q * count, m = r / LB, b = r - m * LB;
    const float val = peer_poll(p.peer, 0, call, q, r, s_bcast);
    s_g[(size_t)m * NB + q * LB + b] = val;
    if (p.gathered_out) p.gathered_out[(size_t)m * NB + q * LB + b] = val;
  }
  for (int b = tid; b < LB; b += kPeerThreads) p.adv_out[b] = 0.f;   // ragged tails stay zero like the reference (TR:445)
  __syncthreads();
  if (s_bcast[1] == 0u) {                                            // a peer never showed up: poison every output
    const float nan = __int_as_float(0x7fc00000);
    for (int b = tid; b < LB; b += kPeerThreads) p.adv_out[b] = nan;
    if (p.gathered_out) for (int i = tid; i < M * NB; i += kPeerThreads) p.gathered_out[i] = nan;
    peer_commit(p.peer, 0, call);
    return;
  }

  const int lo = p.peer.rank * LB, hi = lo + LB;                     // this rank's columns
  if (p.mode == MIXGRPO_ADV_GLOBAL) {                                // TR:498: statistics of the gathered vector
    if (warp == 0) warp_group_stats(s_g, NB, 0, lane, s_stat);
    __syncthreads();
    for (int b = tid; b < LB; b += kPeerThreads) p.adv_out[b] = __fdiv_rn(__fsub_rn(s_g[lo + b], s_stat[0]), s_stat[1]);
  } else {
    // groups are consecutive runs of G columns starting at `origin`: this rank's own columns (whole groups per rank,
    // TR:443-461) or the gathered order (a group split across ranks, SURVEY §8e extended mode)
    const int G = p.G;
    const int origin = p.mode == MIXGRPO_ADV_GROUP_LOCAL ? lo : 0;
    const int span = p.mode == MIXGRPO_ADV_GROUP_LOCAL ? LB : NB;
    const int n_groups = span / G;
    int g_lo = 0, g_hi = n_groups - 1;
    if (p.mode == MIXGRPO_ADV_GROUP_SPLIT) { g_lo = lo / G; g_hi = min((hi - 1) / G, n_groups - 1); }
    const int n_rel = g_hi - g_lo + 1;
    for (int pr = warp; pr < n_rel * M; pr += kPeerWarps) {
      const int g = g_lo + pr / M, m = pr % M;
      warp_group_stats(s_g + (size_t)m * NB + origin + g * G, G, p.trim, lane, s_stat + 2 * pr);
    }
    __syncthreads();
    for (int e = tid; e < n_rel * G; e += kPeerThreads) {
      const int gl = e / G, i = e - gl * G;
      const int j = origin + (g_lo + gl) * G + i;
      if (j < lo || j >= hi) continue;
      float out = 0.f;
      for (int m = 0; m < M; ++m) {
        const float* st = s_stat + 2 * (gl * M + m);
        const float a = __fdiv_rn(__fsub_rn(s_g[(size_t)m * NB + j], st[0]), st[1]);
        out = p.weights ? __fadd_rn(out, __fmul_rn(a, p.weights[m])) : a;      // TR:467 | TR:489
      }
      p.adv_out[j - lo] = out;
    }
  }
  peer_commit(p.peer, 0, call);
}

// ------------------------------------------------------------------ small all-reduce (sum in rank order / average)
struct ReduceParams {
  PeerArgs peer;
  float* vals;    // [count] in place
  int count, average;
};

__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_kernel(const __grid_constant__ ReduceParams p) {
  pdl_prologue();
  __shared__ uint32_t s_bcast[2];
  const PeerCall call = peer_push(p.peer, 1, p.vals, p.count, s_bcast);
  const int W = p.peer.world;
  if ((int)threadIdx.x < p.count) {
    // every rank adds the W contributions in rank order: the result is bit-identical on all ranks and run to run
    float acc = peer_poll(p.peer, 1, call, 0, threadIdx.x, s_bcast);
    for (int q = 1; q < W; ++q) acc = __fadd_rn(acc, peer_poll(p.peer, 1, call, q, threadIdx.x, s_bcast));
    if (p.average) acc = __fdiv_rn(acc, (float)W);
    p.vals[threadIdx.x] = acc;
  }
  __syncthreads();
  if (s_bcast[1] == 0u && (int)threadIdx.x < p.count) p.vals[threadIdx.x] = __int_as_float(0x7fc00000);   // timed out
  peer_commit(p.peer, 1, call);
}

static unsigned long long g_peer_timeout_ms = 600000ull;   // mixgrpo_set_tuning key 2 (0 = wait forever); 10 min like NCCL's watchdog

static int fill_peer(PeerArgs& a, void* const* regions_host, int rank, int world, int64_t cap) {
  if (!regions_host || world < 1 || world > kMaxWorld || rank < 0 || rank >= world || cap < 1) return MIXGRPO_EINVAL;
  for (int q = 0; q < kMaxWorld; ++q) a.region[q] = nullptr;
  for (int q = 0; q < world; ++q) {
    if (!regions_host[q] || (reinterpret_cast<uintptr_t>(regions_host[q]) % 256) != 0) return MIXGRPO_EINVAL;
    a.region[q] = static_cast<char*>(regions_host[q]);
  }
  a.rank = rank; a.world = world; a.cap = cap; a.timeout_ns = g_peer_timeout_ms * 1000000ull;
  return 0;
}

}  // namespace mg

using namespace mg;

int mixgrpo_peer_set_timeout_ms(int ms) {   // reached through mixgrpo_set_tuning(2, ms); returns the previous value
  const int old = (int)g_peer_timeout_ms;
  g_peer_timeout_ms = (unsigned long long)ms;
  return old;
}

extern "C" __attribute__((visibility("default"))) int64_t mixgrpo_peer_region_bytes(int world, int64_t cap_floats) {
  if (world < 1 || world > kMaxWorld || cap_floats < 1) return 0;
  return region_bytes(world, cap_floats);
}

extern "C" __attribute__((visibility("default"))) int mixgrpo_peer_region_alloc(int world, int64_t cap_floats, void** region_out,
                                                                                void* ipc_handle_out) {
  if (!region_out || world < 1 || world > kMaxWorld || cap_floats < 1) return MIXGRPO_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == MIXGRPO_PEER_HANDLE_BYTES, "IPC handle size");
  void* ptr = nullptr;
  const long long bytes = region_bytes(world, cap_floats);
  cudaError_t e = cudaMalloc(&ptr, (size_t)bytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemset(ptr, 0, (size_t)bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();      // peers must never observe a non-zero flag of a previous owner
  if (e == cudaSuccess && ipc_handle_out) e = cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(ipc_handle_out), ptr);
  if (e != cudaSuccess) { cudaFree(ptr); return (int)e; }
  *region_out = ptr;
  return 0;
}

extern "C" __attribute__((visibility("default"))) int mixgrpo_peer_region_open(const void* ipc_handle, void** region_out) {
  if (!ipc_handle || !region_out) return MIXGRPO_EINVAL;
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  return (int)cudaIpcOpenMemHandle(region_out, h, cudaIpcMemLazyEnablePeerAccess);
}

extern "C" __attribute__((visibility("default"))) int mixgrpo_peer_region_close(void* region) {
  return region ? (int)cudaIpcCloseMemHandle(region) : MIXGRPO_EINVAL;
}

extern "C" __attribute__((visibility("default"))) int mixgrpo_peer_region_free(void* region) {
  return region ? (int)cudaFree(region) : MIXGRPO_EINVAL;
}

extern "C" __attribute__((visibility("default"))) int mixgrpo_peer_region_status(const void* region, int* seq_gather_host, int* seq_reduce_host,
                                                                                 int* status_host) {
  if (!region) return MIXGRPO_EINVAL;
  uint32_t h[3];
  cudaError_t e = cudaMemcpy(h, region, sizeof(h), cudaMemcpyDeviceToHost);   // synchronising: tests / teardown only
  if (e != cudaSuccess) return (int)e;
  if (seq_gather_host) *seq_gather_host = (int)h[0];
  if (seq_reduce_host) *seq_reduce_host = (int)h[1];
  if (status_host) *status_host = (int)h[2];
  return 0;
}

extern "C" __attribute__((visibility("default"))) int mixgrpo_peer_gather_advantages(void* const* regions_host, int rank, int world, int64_t cap_floats,
                                                                                     const float* rewards, const float* weights, int n_models,
                                                                                     int64_t local_B, int num_generations, int trim_size, int mode,
                                                                                     float* gathered_out, float* advantages, void* stream) {
  GatherAdvParams p;
  const int rc = fill_peer(p.peer, regions_host, rank, world, cap_floats);
  if (rc) return rc;
  if (!rewards || !advantages || n_models <= 0 || local_B <= 0 || (n_models > 1 && !weights)) return MIXGRPO_EINVAL;
  if ((int64_t)n_models * local_B > cap_floats) return MIXGRPO_ENOSPACE;
  if (mode != MIXGRPO_ADV_GROUP_LOCAL && mode != MIXGRPO_ADV_GROUP_SPLIT && mode != MIXGRPO_ADV_GLOBAL) return MIXGRPO_EINVAL;
  const int64_t NB = (int64_t)world * local_B;
  int64_t stat_floats = 2;
  if (mode == MIXGRPO_ADV_GLOBAL) {
    if (n_models != 1 || NB < 1) return MIXGRPO_EINVAL;                        // TR:495-496
  } else {
    const int G = num_generations;
    if (G <= 0 || trim_size < 0 || trim_size > G - 1) return MIXGRPO_EINVAL;
    const int64_t span = mode == MIXGRPO_ADV_GROUP_LOCAL ? local_B : NB;
    stat_floats = 2 * (span / G + 2) * n_models;
  }
  const size_t smem = ((size_t)n_models * NB + (size_t)stat_floats) * sizeof(float);
  if (smem > 48 * 1024) return MIXGRPO_ENOSPACE;
  p.rewards = rewards; p.weights = weights; p.gathered_out = gathered_out; p.adv_out = advantages;
  p.n_models = n_models; p.local_B = (int)local_B; p.G = num_generations; p.trim = trim_size; p.mode = mode;
  launch_pdl(peer_gather_adv_kernel, 1, kPeerThreads, smem, reinterpret_cast<cudaStream_t>(stream), p);
  return (int)cudaGetLastError();
}

extern "C" __attribute__((visibility("default"))) int mixgrpo_peer_allreduce(void* const* regions_host, int rank, int world, int64_t cap_floats,
                                                                             float* values, int count, int average, void* stream) {
  ReduceParams p;
  const int rc = fill_peer(p.peer, regions_host, rank, world, cap_floats);
  if (rc) return rc;
  if (!values || count < 1) return MIXGRPO_EINVAL;
  if (count > kRedCap) return MIXGRPO_ENOSPACE;
  p.vals = values; p.count = count; p.average = average ? 1 : 0;
  launch_pdl(peer_allreduce_kernel, 1, kPeerThreads, 0, reinterpret_cast<cudaStream_t>(stream), p);
  return (int)cudaGetLastError();
}
