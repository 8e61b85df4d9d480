// Reward -> group-relative advantage (TR:439-501) and the clipped-ratio GRPO loss with its
// closed-form backward (TR:560-585).  TR = /root/reference/fastvideo/train_grpo_flux.py.
//
// These are tiny (tens of scalars) — the reference spends ~10 eager kernels per group per model and
// ~15 scalar kernels + 4 all-reduces + 5 .item() syncs per (sample, step) on them.  Here each is ONE
// launch with no host sync; the loss kernel also emits dL/dlogp on the device so the log-prob
// backward kernel can consume it directly, and accumulates the logging scalars on the device.  (The
// fused policy path, mixgrpo_policy_fwd/_bwd, evaluates the same per-sample loss inside the log-prob
// kernels and needs no loss launch at all — see loss_terms() in common.cuh.)
#include "common.cuh"

namespace mg {

constexpr int kAdvThreads = 128;
constexpr int kAdvWarps = kAdvThreads / 32;
constexpr int kMaxGroup = 8192;

// One CTA per prompt group.  All models' rewards of the group are staged in shared memory with one
// round of loads; warp w then owns models w, w+4, ...: group statistics in fp64 (12..24 numbers) rounded
// once to fp32, then the reference's fp32 expression (r - mean) / (std + 1e-8) (TR:459-461) with separately
// rounded ops.  The weighted merge (TR:465-468) runs in model order so the fp32 sum matches the reference's.
__global__ void __launch_bounds__(kAdvThreads) group_adv_kernel(const float* __restrict__ rewards,
                                                               const float* __restrict__ weights, int n_models,
                                                               long long local_B, int G, int trim,
                                                               float* __restrict__ adv) {
  pdl_prologue();
  extern __shared__ float s_r[];            // [n_models][G] rewards, then [n_models][2] (mean, std+1e-8)
  float* s_stat = s_r + (size_t)n_models * G;
  const int g = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long base = (long long)g * G;
  for (int i = tid; i < n_models * G; i += kAdvThreads) {
    const int m = i / G, j = i - m * G;
    s_r[i] = rewards[(long long)m * local_B + base + j];
  }
  __syncthreads();
  const int kept_n = G - trim;
  for (int m = warp; m < n_models; m += kAdvWarps) {
    const float* r = s_r + (size_t)m * G;
    // which members enter the statistics: all, or all but the `trim` smallest (stable rank), TR:451-457
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
      // Bessel-corrected std; a single kept element gives 0/0 = NaN exactly like torch.std
      s_stat[2 * m] = (float)mean_d;
      s_stat[2 * m + 1] = __fadd_rn((float)sqrt(sq / (double)(kept_n - 1)), 1e-8f);
    }
  }
  __syncthreads();
  for (int i = tid; i < G; i += kAdvThreads) {
    float out = 0.f;
    for (int m = 0; m < n_models; ++m) {
      const float a = __fdiv_rn(__fsub_rn(s_r[(size_t)m * G + i], s_stat[2 * m]), s_stat[2 * m + 1]);
      out = weights ? __fadd_rn(out, __fmul_rn(a, weights[m])) : a;       // merged += adv * w (TR:467) | plain (TR:489)
    }
    adv[base + i] = out;
  }
}

// No-group path (TR:498): one model, statistics of the all-gathered vector.
__global__ void __launch_bounds__(256) global_adv_kernel(const float* __restrict__ rewards, long long local_B,
                                                        const float* __restrict__ stat, long long n_stat,
                                                        float* __restrict__ adv) {
  pdl_prologue();
  __shared__ double s_red[8];
  __shared__ float s_stat[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double sum = 0.0;
  for (long long i = tid; i < n_stat; i += 256) sum += (double)stat[i];
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  double tot = 0.0;
  for (int w = 0; w < 8; ++w) tot += s_red[w];
  const double mean_d = tot / (double)n_stat;
  __syncthreads();
  double sq = 0.0;
  for (long long i = tid; i < n_stat; i += 256) { const double d = (double)stat[i] - mean_d; sq += d * d; }
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if (lane == 0) s_red[warp] = sq;
  __syncthreads();
  if (tid == 0) {
    double tq = 0.0;
    for (int w = 0; w < 8; ++w) tq += s_red[w];
    s_stat[0] = (float)mean_d;
    s_stat[1] = __fadd_rn((float)sqrt(tq / (double)(n_stat - 1)), 1e-8f);
  }
  __syncthreads();
  const float mean = s_stat[0], sd = s_stat[1];
  for (long long i = tid; i < local_B; i += 256) adv[i] = __fdiv_rn(__fsub_rn(rewards[i], mean), sd);
}

// Clipped surrogate + KL over a batch of B log-probs, forward and dL/dnew_logp in one launch (TR:560-583).
// One warp when B <= 32 (the usual case: no barrier at all), else 256 threads.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) grpo_loss_kernel(const float* new_lp, const float* old_lp,     // no __restrict__: the launch before may
                                                           const float* adv, long long B, LossParams q,                  // have written them (no .nc loads, see ld_dep)
                                                           float* __restrict__ stats, float* __restrict__ grad,
                                                           float* __restrict__ accum) {
  pdl_prologue();
  __shared__ float s_warp[3][THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float s_pol = 0.f, s_kl = 0.f, s_cf = 0.f;
  const float invB = 1.f / (float)B;
  for (long long i = tid; i < B; i += THREADS) {
    const LossTerms t = loss_terms(new_lp[i], old_lp[i], adv[i], q, invB);
    s_pol += t.policy_num;
    s_kl += t.kl_num;
    s_cf += t.clip;
    if (grad) grad[i] = t.grad;
  }
  s_pol = warp_sum(s_pol); s_kl = warp_sum(s_kl); s_cf = warp_sum(s_cf);
  if constexpr (THREADS > 32) {
    if (lane == 0) { s_warp[0][warp] = s_pol; s_warp[1][warp] = s_kl; s_warp[2][warp] = s_cf; }
    __syncthreads();
    if (warp != 0) return;
    s_pol = warp_sum(lane < THREADS / 32 ? s_warp[0][lane] : 0.f);
    s_kl = warp_sum(lane < THREADS / 32 ? s_warp[1][lane] : 0.f);
    s_cf = warp_sum(lane < THREADS / 32 ? s_warp[2][lane] : 0.f);
  }
  if (tid == 0) {
    const float policy = __fdiv_rn(__fdiv_rn(s_pol, (float)B), q.denom);              // TR:575-577
    const float kl = __fdiv_rn(__fmul_rn(0.5f, __fdiv_rn(s_kl, (float)B)), q.denom);  // TR:578-582
    const float loss = __fadd_rn(policy, __fmul_rn(q.klc, kl));                       // TR:583
    const float cf = __fdiv_rn(s_cf, (float)B);
    stats[0] = loss; stats[1] = policy; stats[2] = kl; stats[3] = cf;
    if (accum) {   // fire-and-forget REDs: this launch is the only writer on the stream, so the running sum stays ordered
      atomicAdd(accum + 0, loss); atomicAdd(accum + 1, policy); atomicAdd(accum + 2, kl); atomicAdd(accum + 3, cf);
    }
  }
}

}  // namespace mg

using namespace mg;

extern "C" __attribute__((visibility("default"))) int mixgrpo_group_advantages(const float* rewards, const float* weights, int n_models, int64_t local_B,
                                        int num_generations, int trim_size, int use_group, const float* stat_rewards,
                                        int64_t n_stat, float* advantages, void* stream) {
  if (!rewards || !advantages || n_models <= 0 || local_B <= 0 || (n_models > 1 && !weights)) return MIXGRPO_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!use_group) {
    if (n_models != 1 || !stat_rewards || n_stat <= 0) return MIXGRPO_EINVAL;   // TR:495-496
    launch_pdl(global_adv_kernel, 1, 256, 0, st, rewards, local_B, stat_rewards, n_stat, advantages);
    return (int)cudaGetLastError();
  }
  const int G = num_generations;
  if (G <= 0 || G > kMaxGroup || trim_size < 0 || trim_size > G - 1) return MIXGRPO_EINVAL;
  const size_t smem = ((size_t)n_models * G + 2 * (size_t)n_models) * sizeof(float);
  if (smem > 48 * 1024) return MIXGRPO_EINVAL;                                 // n_models * G <= ~12k rewards per group
  const int64_t n_groups = local_B / G;                                        // TR:444 (floor; tail untouched)
  if (n_groups <= 0) return 0;
  launch_pdl(group_adv_kernel, (unsigned)n_groups, kAdvThreads, smem, st, rewards, weights, n_models, local_B, G, trim_size, advantages);
  return (int)cudaGetLastError();
}

extern "C" __attribute__((visibility("default"))) int mixgrpo_grpo_loss(const float* new_logp, const float* old_logp, const float* advantages, int64_t B,
                                 double clip_range, double adv_clip_max, double kl_coeff, double denom,
                                 float* stats_out, float* grad_new_logp, float* stats_accum, void* stream) {
  if (!new_logp || !old_logp || !advantages || !stats_out || B <= 0) return MIXGRPO_EINVAL;
  const LossParams q = make_loss_params(nullptr, nullptr, nullptr, clip_range, adv_clip_max, kl_coeff, denom);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (B <= 32) launch_pdl(grpo_loss_kernel<32>, 1, 32, 0, st, new_logp, old_logp, advantages, B, q, stats_out, grad_new_logp, stats_accum);
  else launch_pdl(grpo_loss_kernel<256>, 1, 256, 0, st, new_logp, old_logp, advantages, B, q, stats_out, grad_new_logp, stats_accum);
  return (int)cudaGetLastError();
}
