// flow_grpo_step family (SU:157-210) of the fused step kernel + the entry points that do not belong to one family.
#include "step_kernel.cuh"

namespace mg {
int g_max_ctas_per_sample = kMaxCtasPerSample;          // bench knob (mixgrpo_set_tuning key 0)
int g_use_pdl = 1;                                      // bench knob (key 1): programmatic dependent launch on/off
int g_bwd_threads = kThreads;                           // knob (key 7): CTA size of the log-prob backward kernels
int g_half_ctas = 1;                                    // knob (key 6): deferred launches in the 128-thread shape (0 never, 1 auto, 2 always)
std::atomic<long long> g_half_launches{0};             // host threads of one process may launch concurrently
int sm_count() {
  static std::atomic<int> cached[64];               // zero-initialised; two threads may both fill an entry with the same value
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int sms = cached[dev].load(std::memory_order_relaxed);
  if (sms == 0) {
    int n = 0;
    sms = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
    cached[dev].store(sms, std::memory_order_relaxed);
  }
  return sms;
}
int policy_fwd_dance(StepParams& p, int v_dtype, int64_t B, bool vec, bool rnd, cudaStream_t st);   // step_dance.cu
}  // namespace mg

using namespace mg;

extern "C" __attribute__((visibility("default"))) int64_t mixgrpo_step_workspace_bytes(int64_t B, int64_t n) {
  if (B <= 0 || n <= 0) return 0;
  // one 32-byte record per sample: { packed 64-bit accumulator | 32-bit epoch | 32-bit status (record 0) | 64-bit side
  // accumulator (Q39.24) for oversized shares | 64-bit integer accumulator for shares above 2^27 } — the layout does not depend on B, so calls with different batch sizes can
  // share one zero-initialised allocation
  return ((B * (int64_t)(kWsStride * sizeof(unsigned long long)) + 255) / 256) * 256;
}

extern "C" __attribute__((visibility("default"))) int64_t mixgrpo_deferred_workspace_bytes(int64_t B, int64_t n) {
  if (B <= 0 || n <= 0) return 0;
  // kDeferSubs records per sample: a deferred launch spreads the sample's arrivals over them (step_math.cuh)
  return ((B * (int64_t)(kDeferSubs * kWsStride * sizeof(unsigned long long)) + 255) / 256) * 256;
}

extern "C" __attribute__((visibility("default"))) int mixgrpo_set_tuning(int key, int value) {
  if (key == 2) return value < 0 ? MIXGRPO_EINVAL : mixgrpo_peer_set_timeout_ms(value);
  if (key >= 3 && key <= 5) return mixgrpo_policy_set_tuning(key, value);
  if (key == 8) return (int)(g_half_launches.load(std::memory_order_relaxed) & 0x7fffffff);       // read-only: launches issued in the 128-thread shape
  if (key == 7) {
    if (value != kThreads && value != kHalfThreads) return MIXGRPO_EINVAL;
    const int old = g_bwd_threads;
    g_bwd_threads = value;
    return old;
  }
  if (key == 1 || key == 6) {
    if (value < 0 || value > (key == 6 ? 2 : 1)) return MIXGRPO_EINVAL;
    int& knob = key == 1 ? g_use_pdl : g_half_ctas;
    const int old = knob;
    knob = value;
    return old;
  }
  if (key != 0 || value < 1 || value > kMaxCtasPerSample) return MIXGRPO_EINVAL;
  const int old = g_max_ctas_per_sample;
  g_max_ctas_per_sample = value;
  return old;
}

extern "C" __attribute__((visibility("default"))) int mixgrpo_flow_step(const void* v, int v_dtype, const float* x, int64_t x_bs, const void* noise,
                                 const float* x_next_in, int64_t in_bs, float* x_next_out, int64_t out_bs,
                                 float* x0_out, float* mean_out, float* logp_out, void* workspace,
                                 int64_t workspace_bytes, int64_t B, int64_t n, const mixgrpo_step_coefs* coefs_host,
                                 int src, unsigned flags, void* stream, const mixgrpo_step_ext* ext) {
  int err = 0;
  if (!coefs_host || !check_common(v, x, B, n, v_dtype, workspace, workspace_bytes, logp_out, &err, (flags & MIXGRPO_FLAG_DEFER_LOGP) != 0)) return err ? err : MIXGRPO_EINVAL;
  if (((src == MIXGRPO_SRC_NOISE || src == MIXGRPO_SRC_PHILOX) && !noise) || (src == MIXGRPO_SRC_GIVEN && !x_next_in)) return MIXGRPO_EINVAL;
  StepParams p;
  fill(p, v, x, x_bs, src == MIXGRPO_SRC_PHILOX ? nullptr : noise, x_next_in, in_bs, nullptr, nullptr, x_next_out, out_bs, x0_out, mean_out, logp_out, workspace, B, n, coefs_host);
  if (src == MIXGRPO_SRC_PHILOX) set_philox(p, noise);
  set_early(p, flags);
  const bool vec = vector_ok(p, v_dtype, v_dtype, n);
  if ((err = set_ext(p, ext, n, vec, src)) != 0) return err;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (v_dtype == MIXGRPO_F32) return pick_src<kFlow, float, float, 1, false, false>(p, B, src, vec, st);
  if (flags & MIXGRPO_FLAG_ROUND_LIKE_TORCH) return pick_src<kFlow, __nv_bfloat16, __nv_bfloat16, 1, true, false>(p, B, src, vec, st);
  return pick_src<kFlow, __nv_bfloat16, __nv_bfloat16, 1, false, false>(p, B, src, vec, st);
}

// Fused policy-update forward: log p(x_next | x, v) for the stored transition (TR:149-168 via grpo_one_step) AND the
// per-sample clipped-ratio loss terms (TR:560-583) accumulated into stats_rows — one launch, nothing else.
extern "C" __attribute__((visibility("default"))) int mixgrpo_policy_fwd(int family, const void* v, int v_dtype, const float* x, int64_t x_bs,
                                  const float* x_next, int64_t in_bs, float* logp_out, void* workspace, int64_t workspace_bytes,
                                  int64_t B, int64_t n, const mixgrpo_step_coefs* coefs_host, const mixgrpo_loss_args* loss,
                                  unsigned flags, void* stream) {
  int err = 0;
  if (!coefs_host || !logp_out || !x_next || !check_common(v, x, B, n, v_dtype, workspace, workspace_bytes, logp_out, &err))
    return err ? err : MIXGRPO_EINVAL;
  if (family != kFlow && family != kDance) return MIXGRPO_EINVAL;
  if (flags & MIXGRPO_FLAG_DEFER_LOGP) return MIXGRPO_EINVAL;             // the fused loss needs the log-prob inside this launch
  if (loss && (!loss->old_logp || !loss->advantages)) return MIXGRPO_EINVAL;
  StepParams p;
  fill(p, v, x, x_bs, nullptr, x_next, in_bs, nullptr, nullptr, nullptr, n, nullptr, nullptr, logp_out, workspace, B, n, coefs_host);
  if (loss) {
    if (loss->stats_rows && (reinterpret_cast<uintptr_t>(loss->stats_rows) % 16) != 0) return MIXGRPO_EINVAL;   // rows are float4
    p.loss = make_loss_params(loss->old_logp, loss->advantages, loss->stats_rows, loss->clip_range, loss->adv_clip_max, loss->kl_coeff, loss->denom);
    p.loss.accumulate = loss->accumulate ? 1 : 0;
  }
  set_early(p, flags);
  const bool vec = vector_ok(p, v_dtype, MIXGRPO_F32, n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool rnd = (flags & MIXGRPO_FLAG_ROUND_LIKE_TORCH) != 0;
  const int src = MIXGRPO_SRC_GIVEN;
  if (family == kFlow) {
    if (v_dtype == MIXGRPO_F32) return pick_src<kFlow, float, float, 1, false, false>(p, B, src, vec, st);
    if (rnd) return pick_src<kFlow, __nv_bfloat16, __nv_bfloat16, 1, true, false>(p, B, src, vec, st);
    return pick_src<kFlow, __nv_bfloat16, __nv_bfloat16, 1, false, false>(p, B, src, vec, st);
  }
  return policy_fwd_dance(p, v_dtype, B, vec, rnd, st);
}

namespace mg {
constexpr int kFinalizeChunk = 64;
struct FinalizeParams {
  unsigned long long* ws;
  float* out;
  long long stride_words, out_stride;
  int B, n_launches;
  float log_scale[kFinalizeChunk], log_norm[kFinalizeChunk];
  unsigned long long active;                      // bit i: launch i accumulated
};

// kDeferSubs lanes per (launch, sample): each takes one sub-record (accumulator + side words), the group adds them up as integers,
// its first lane writes the log-prob; every word is left zeroed
__global__ void __launch_bounds__(128) logp_finalize_kernel(const __grid_constant__ FinalizeParams p) {
  pdl_prologue();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = idx / kDeferSubs, sub = idx - r * kDeferSubs;
  const bool live = r < p.n_launches * p.B;
  const int i = live ? r / p.B : 0, b = live ? r - i * p.B : 0;
  const bool active = live && ((p.active >> i) & 1ull);
  unsigned long long tot = 0ull, wide = 0ull, huge = 0ull;
  if (active) {
    unsigned long long* rec = p.ws + (long long)i * p.stride_words + ((long long)kDeferSubs * b + sub) * kWsStride;
    tot = *rec;
    *rec = 0ull;
    if ((tot >> kCountBits) & ((1ull << kPoisonBits) - 1)) {       // some part of this sub-record went to the side words
      wide = atomicExch(rec + kWsWide, 0ull);
      huge = atomicExch(rec + kWsHuge, 0ull);
    }
  }
  unsigned long long bad = wide >> 63;
  wide &= ~(1ull << 63);
#pragma unroll
  for (int o = kDeferSubs / 2; o > 0; o >>= 1) {                   // groups of kDeferSubs lanes are aligned inside the warp
    tot += __shfl_xor_sync(0xffffffffu, tot, o);
    wide += __shfl_xor_sync(0xffffffffu, wide, o);
    huge += __shfl_xor_sync(0xffffffffu, huge, o);
    bad |= __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if (!live || sub != 0) return;
  float lp = __int_as_float(0x7fc00000);
  if (active && !bad) {
    double q = (double)(tot >> (kCountBits + kPoisonBits)) * (1.0 / 4294967296.0);
    q += (double)wide * (1.0 / 16777216.0) + (double)huge;                        // packed_total's association (adds +0.0 when nothing went wide)
    lp = __fsub_rn(__fsub_rn(-(float)q, p.log_scale[i]), p.log_norm[i]);          // SU:201-208
  }
  p.out[(long long)i * p.out_stride + b] = lp;
}

__global__ void philox_advance_kernel(unsigned long long* state, unsigned long long inc) {
  pdl_prologue();
  state[1] += inc;
}
}  // namespace mg

extern "C" __attribute__((visibility("default"))) int mixgrpo_logp_finalize(void* workspace, int64_t launch_stride_bytes, int64_t n_launches, int64_t B,
                                                                            const float* log_scale_host, const float* log_norm_host,
                                                                            const int* active_host, float* logp_out, int64_t out_stride, void* stream) {
  if (!workspace || !logp_out || !log_scale_host || !log_norm_host || n_launches <= 0 || n_launches > 4096 || B <= 0 || B > 65535 ||
      launch_stride_bytes < B * (int64_t)(kDeferSubs * kWsStride * sizeof(unsigned long long)) || (launch_stride_bytes % 8) != 0 || out_stride < B ||
      (reinterpret_cast<uintptr_t>(workspace) % 8) != 0)
    return MIXGRPO_EINVAL;
  for (int64_t lo = 0; lo < n_launches; lo += kFinalizeChunk) {
    FinalizeParams p;
    p.n_launches = (int)((n_launches - lo) < kFinalizeChunk ? (n_launches - lo) : kFinalizeChunk);
    p.B = (int)B;
    p.stride_words = launch_stride_bytes / 8;
    p.ws = reinterpret_cast<unsigned long long*>(workspace) + lo * p.stride_words;
    p.out = logp_out + lo * out_stride;
    p.out_stride = out_stride;
    p.active = 0ull;
    for (int i = 0; i < p.n_launches; ++i) {
      p.log_scale[i] = log_scale_host[lo + i];
      p.log_norm[i] = log_norm_host[lo + i];
      if (!active_host || active_host[lo + i]) p.active |= 1ull << i;
    }
    const int total = p.n_launches * p.B * kDeferSubs;
    launch_pdl(logp_finalize_kernel, (total + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream), p);
    const int rc = (int)cudaGetLastError();
    if (rc != 0) return rc;
  }
  return 0;
}

extern "C" __attribute__((visibility("default"))) int mixgrpo_philox_advance(uint64_t* device_state, uint64_t increment, void* stream) {
  if (!device_state || (reinterpret_cast<uintptr_t>(device_state) % 8) != 0) return MIXGRPO_EINVAL;
  launch_pdl(philox_advance_kernel, 1, 1, 0, reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<unsigned long long*>(device_state),
             (unsigned long long)increment);
  return (int)cudaGetLastError();
}
