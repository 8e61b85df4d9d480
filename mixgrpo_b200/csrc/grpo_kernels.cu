// Reward -> group-relative advantage (TR:439-501) and the clipped-ratio GRPO loss with its
// closed-form backward (TR:560-585).  TR = /root/reference/fastvideo/train_grpo_flux.py.
//
// These are tiny (tens of scalars) — the reference spends ~10 eager kernels per group per model and
// ~15 scalar kernels + 4 all-reduces + 5 .item() syncs per (sample, step) on them.  Here each is ONE
// launch with no host sync; the loss kernel also emits dL/dlogp on the device so the log-prob
// backward kernel can consume it directly, and accumulates the logging scalars on the device.
#include "common.cuh"

namespace mg {

constexpr int kAdvThreads = 128;
constexpr int kMaxGroup = 8192;

// One CTA per prompt group; loops over reward models.  Group statistics are accumulated in fp64
// (12..24 numbers) and rounded once to fp32, then the reference's fp32 expression
//   (r - mean) / (std + 1e-8)  (TR:459-461),  merged += adv * w  (TR:465-468)
// is evaluated with separately rounded fp32 ops.
__global__ void __launch_bounds__(kAdvThreads) group_adv_kernel(const float* __restrict__ rewards,
                                                               const float* __restrict__ weights, int n_models,
                                                               long long local_B, int G, int trim,
                                                               float* __restrict__ adv) {
  extern __shared__ float s_r[];            // [G]
  __shared__ double s_red[2][kAdvThreads / 32];
  __shared__ float s_stat[2];
  const int g = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long base = (long long)g * G;
  for (int m = 0; m < n_models; ++m) {
    const float* r = rewards + (long long)m * local_B + base;
    for (int i = tid; i < G; i += kAdvThreads) s_r[i] = r[i];
    __syncthreads();
    // which members enter the statistics: all, or all but the `trim` smallest (stable rank), TR:451-457
    double sum = 0.0, sq = 0.0;
    int kept_n = G - trim;
    for (int i = tid; i < G; i += kAdvThreads) {
      bool keep = true;
      const float ri = s_r[i];
      if (trim > 0) {
        int rank = 0;
        for (int j = 0; j < G; ++j) rank += (s_r[j] < ri) || (s_r[j] == ri && j < i);
        keep = rank >= trim;
      }
      if (keep) sum += (double)ri;
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) s_red[0][warp] = sum;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < kAdvThreads / 32; ++w) tot += s_red[0][w];
    const double mean_d = tot / (double)kept_n;
    for (int i = tid; i < G; i += kAdvThreads) {
      bool keep = true;
      const float ri = s_r[i];
      if (trim > 0) {
        int rank = 0;
        for (int j = 0; j < G; ++j) rank += (s_r[j] < ri) || (s_r[j] == ri && j < i);
        keep = rank >= trim;
      }
      if (keep) { const double dlt = (double)ri - mean_d; sq += dlt * dlt; }
    }
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (lane == 0) s_red[1][warp] = sq;
    __syncthreads();
    if (tid == 0) {
      double tq = 0.0;
      for (int w = 0; w < kAdvThreads / 32; ++w) tq += s_red[1][w];
      // Bessel-corrected std; a single kept element gives 0/0 = NaN exactly like torch.std
      const double var = tq / (double)(kept_n - 1);
      s_stat[0] = (float)mean_d;
      s_stat[1] = __fadd_rn((float)sqrt(var), 1e-8f);
    }
    __syncthreads();
    const float mean = s_stat[0], sd = s_stat[1];
    const float w = weights ? weights[m] : 1.f;
    for (int i = tid; i < G; i += kAdvThreads) {
      const float a = __fdiv_rn(__fsub_rn(s_r[i], mean), sd);
      float* out = adv + base + i;
      if (weights) *out = (m == 0) ? __fadd_rn(0.f, __fmul_rn(a, w)) : __fadd_rn(*out, __fmul_rn(a, w));
      else *out = a;
    }
    __syncthreads();
  }
}

// No-group path (TR:498): one model, statistics of the all-gathered vector.
__global__ void __launch_bounds__(256) global_adv_kernel(const float* __restrict__ rewards, long long local_B,
                                                        const float* __restrict__ stat, long long n_stat,
                                                        float* __restrict__ adv) {
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

// Clipped surrogate + KL, forward and dL/dnew_logp in one launch (TR:560-583).
//   torch.maximum ties split the gradient 1/2 + 1/2; clamp passes gradient on [lo, hi] inclusive.
__global__ void __launch_bounds__(256) grpo_loss_kernel(const float* __restrict__ new_lp, const float* __restrict__ old_lp,
                                                       const float* __restrict__ adv, long long B, float clip,
                                                       float lo, float hi, float amax, float kl_coeff, float denom,
                                                       float* __restrict__ stats, float* __restrict__ grad,
                                                       float* __restrict__ accum) {
  __shared__ float s_warp[8];
  const int tid = threadIdx.x;
  float s_pol = 0.f, s_kl = 0.f, s_cf = 0.f;
  const float invB = 1.f / (float)B;
  for (long long i = tid; i < B; i += 256) {
    const float a = fminf(fmaxf(adv[i], -amax), amax);            // TR:560-564
    const float lr = __fsub_rn(new_lp[i], old_lp[i]);
    const float r = expf(lr);                                     // TR:566
    const float rc = fminf(fmaxf(r, lo), hi);
    const float un = __fmul_rn(-a, r), cl = __fmul_rn(-a, rc);     // TR:568-573
    s_pol += fmaxf(un, cl);
    s_cf += (fabsf(__fsub_rn(r, 1.f)) > clip) ? 1.f : 0.f;        // TR:574
    s_kl += __fmul_rn(lr, lr);                                    // TR:580
    if (grad) {
      const float inside = (r >= lo && r <= hi) ? 1.f : 0.f;
      float dpl_dr;
      if (un > cl) dpl_dr = -a;
      else if (un < cl) dpl_dr = -a * inside;
      else dpl_dr = 0.5f * (-a) + 0.5f * (-a) * inside;
      const float g_pol = dpl_dr * r * invB / denom;
      const float g_kl = kl_coeff * lr * invB / denom;            // d/dlr [0.5*mean(lr^2)/denom]
      grad[i] = g_pol + g_kl;
    }
  }
  const float t_pol = block_sum(s_pol, s_warp);
  const float t_kl = block_sum(s_kl, s_warp);
  const float t_cf = block_sum(s_cf, s_warp);
  if (tid == 0) {
    const float policy = __fdiv_rn(__fdiv_rn(t_pol, (float)B), denom);              // TR:575-577
    const float kl = __fdiv_rn(__fmul_rn(0.5f, __fdiv_rn(t_kl, (float)B)), denom);  // TR:578-582
    const float loss = __fadd_rn(policy, __fmul_rn(kl_coeff, kl));                  // TR:583
    const float cf = __fdiv_rn(t_cf, (float)B);
    stats[0] = loss; stats[1] = policy; stats[2] = kl; stats[3] = cf;
    if (accum) { accum[0] += loss; accum[1] += policy; accum[2] += kl; accum[3] += cf; }
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
    global_adv_kernel<<<1, 256, 0, st>>>(rewards, local_B, stat_rewards, n_stat, advantages);
    return (int)cudaGetLastError();
  }
  const int G = num_generations;
  if (G <= 0 || G > kMaxGroup || trim_size < 0 || trim_size > G - 1) return MIXGRPO_EINVAL;
  const int64_t n_groups = local_B / G;                                        // TR:444 (floor; tail untouched)
  if (n_groups <= 0) return 0;
  group_adv_kernel<<<(unsigned)n_groups, kAdvThreads, (size_t)G * sizeof(float), st>>>(
      rewards, weights, n_models, local_B, G, trim_size, advantages);
  return (int)cudaGetLastError();
}

extern "C" __attribute__((visibility("default"))) int mixgrpo_grpo_loss(const float* new_logp, const float* old_logp, const float* advantages, int64_t B,
                                 double clip_range, double adv_clip_max, double kl_coeff, double denom,
                                 float* stats_out, float* grad_new_logp, float* stats_accum, void* stream) {
  if (!new_logp || !old_logp || !advantages || !stats_out || B <= 0) return MIXGRPO_EINVAL;
  // python scalars are cast to the tensor dtype (fp32) where torch compares / clamps with them
  const float clip = (float)clip_range, lo = (float)(1.0 - clip_range), hi = (float)(1.0 + clip_range);
  grpo_loss_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      new_logp, old_logp, advantages, B, clip, lo, hi, (float)adv_clip_max, (float)kl_coeff, (float)denom, stats_out,
      grad_new_logp, stats_accum);
  return (int)cudaGetLastError();
}
