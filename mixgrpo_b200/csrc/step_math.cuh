// Per-tile arithmetic of the three operator families (flow_grpo_step SU:157-210, dance_grpo_step SU:212-253, dpm_step
// SU:273-639; SU = /root/reference/fastvideo/utils/sampling_utils.py), shared by the streaming step kernel
// (step_kernel.cuh) and the single-pass policy kernel (policy_kernels.cu).  Every product/sum the reference performs as
// a separate torch kernel is one __fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn here (no FMA contraction), in the same order,
// with torch's bf16 promotion roundings when RND is set.
#pragma once
#include "common.cuh"

namespace mg {

enum Family { kFlow = 0, kDance = 1, kDpm = 2 };

// packed log-prob accumulator (one 64-bit word per sample): [ sum Q8.32 : 40 | wide-part count : 12 | arrivals : 12 ]
constexpr int kTile = kThreads * kVec;          // scalars per CTA-tile of the 256-thread kernels
constexpr int kHalfThreads = kThreads / 2;      // the 128-thread (4-warp) CTA of the deferred rollout kernels ...
constexpr int kHalfTile = kHalfThreads * kVec;  // ... owns one HALF-tile: 1024 consecutive scalars
constexpr int kCountBits = 12, kPoisonBits = 12;
constexpr unsigned long long kArrivalMask = (1ull << kCountBits) - 1;
// a CTA contributes two parts (its half-tiles), each of which may be counted in the 12-bit "wide" field: <= 2047 CTAs per sample
constexpr int kMaxCtasPerSample = (1 << (kCountBits - 1)) - 1;
// workspace record per sample (in 64-bit words): { accumulator, { epoch : 32 | status : 32 }, wide side accumulator, integer side
// accumulator } — see mixgrpo_step_workspace_bytes.  A DEFERRED launch (MIXGRPO_FLAG_DEFER_LOGP) spreads a sample's arrivals over
// kDeferSubs such records (mixgrpo_deferred_workspace_bytes): 3072 fire-and-forget reductions into 12 addresses serialise in the
// L2 (measured: +0.5 us per launch), into 12 x 8 they do not; mixgrpo_logp_finalize adds the sub-records up (integers: exact).
constexpr int kWsStride = 4;
constexpr int kWsWide = 2;                      // word index of the side accumulator inside a record
constexpr int kWsHuge = 3;                      // word index of the second side accumulator (integer units)
constexpr int kDeferSubs = 8;
constexpr float kWideCap = 134217728.f;         // 2^27: largest share the Q39.24 side accumulator takes (4094 parts x 2^27 x 2^24 < 2^63)
constexpr float kHugeCap = 1125899906842624.f;  // 2^50: largest share the integer side accumulator takes (4094 x 2^50 < 2^63)

// The unit of the reduction is the HALF-TILE: the 1024 consecutive scalars four warps own.  Its d^2 sum is taken in fp32 in a
// fixed order (per thread pair-wise, five xor-shuffle levels per warp, two over the four warp sums), scaled to
// r = sum / (n * 2 s^2) and converted to Q8.32 fixed point ON ITS OWN; everything above that level is integer addition, which
// commutes.  So a 256-thread CTA (two half-tiles, added as integers before its one atomic), a 128-thread CTA of the deferred
// rollout kernels (one half-tile, one fire-and-forget reduction), the batched window kernel and the single-pass policy kernel
// all produce bit-identical log-probs, in any arrival order.
//
// The packed word holds parts of mean(d^2 / 2 s^2) up to 255/parts each — ample for any transition a sane policy
// produces (the value is ~0.5 on the rollout's own samples).  The reference, though, returns a FINITE log-prob however far
// x_next is from the mean (SU:201-208), so a part that does not fit is not dropped: it goes, as Q39.24 fixed point, into
// the record's 64-bit side accumulator (integer adds: still order-independent, still bitwise reproducible) and only flags the
// fact in the packed word's 12-bit "wide" count.  The finalizer folds the side word in and re-zeroes it.  A part above 2^27
// (|d|/s > 16000) goes, rounded to an integer (relative resolution 2^-27), into a second side word; only a non-finite /
// negative part, or one above 2^50 (|d|/s > 4e7, where the reference's own fp32 mean has long lost its digits), yields NaN.
// The common path is unchanged: one atomic per CTA, no fence.
// Returns the part's bits WITHOUT an arrival; `parts` = half-tiles per sample of the 256-thread tiling (2 x CTAs per sample).
__device__ __forceinline__ unsigned long long packed_part(float r, int parts, unsigned long long* rec) {
  const float cap = 255.0f / (float)parts;
  unsigned long long add = 0ull;
  if (!(r >= 0.f && r <= cap)) {
    if (r > cap && r <= kWideCap) atomicAdd(rec + kWsWide, __float2ull_rn(r * 16777216.0f));
    else if (r > kWideCap && r <= kHugeCap) atomicAdd(rec + kWsHuge, __float2ull_rn(r));
    else atomicOr(rec + kWsWide, 1ull << 63);
    __threadfence();                             // the side word is visible before this CTA's arrival is counted
    add = 1ull << kCountBits;
    r = 0.f;
  }
  return add + (__float2ull_rn(r * 4294967296.0f) << (kCountBits + kPoisonBits));
}

// Warp 0 of a 256-thread CTA, all 32 lanes: s_warp = the 8 warp sums of d^2 (warps 0-3 = first half-tile, 4-7 = second).
// Returns, in lane 0, the word the CTA adds to its sample's record: one arrival + both half-tile parts.
static_assert(kThreads / 32 == 8, "cta_share assumes 8 warps per CTA");
__device__ __forceinline__ unsigned long long cta_share(const float* s_warp, int lane, float denom, int parts, unsigned long long* rec) {
  float t = lane < 8 ? s_warp[lane] : 0.f;
  t += __shfl_xor_sync(0xffffffffu, t, 2);
  t += __shfl_xor_sync(0xffffffffu, t, 1);        // lanes 0 and 4: (w0 + w2) + (w1 + w3) of their half
  unsigned long long add = 0ull;
  if (lane == 0 || lane == 4) add = packed_part(__fdiv_rn(t, denom), parts, rec);
  add += __shfl_down_sync(0xffffffffu, add, 4);
  return add + 1ull;
}
// the same two values from the 8 warp sums held by ONE thread (single-pass policy kernel)
__device__ __forceinline__ unsigned long long cta_share_serial(const float* w, float denom, int parts, unsigned long long* rec) {
  const float h0 = __fadd_rn(__fadd_rn(w[0], w[2]), __fadd_rn(w[1], w[3]));
  const float h1 = __fadd_rn(__fadd_rn(w[4], w[6]), __fadd_rn(w[5], w[7]));
  return packed_part(__fdiv_rn(h0, denom), parts, rec) + packed_part(__fdiv_rn(h1, denom), parts, rec) + 1ull;
}
// warp 0 of a 128-thread CTA: s_warp = its 4 warp sums; lane 0 gets the half-tile's sum in the same order
__device__ __forceinline__ float half_sum(const float* s_warp, int lane) {
  float t = lane < 4 ? s_warp[lane] : 0.f;
  t += __shfl_xor_sync(0xffffffffu, t, 2);
  t += __shfl_xor_sync(0xffffffffu, t, 1);
  return t;
}

// tot = the packed word after the LAST arrival; returns mean(d^2 / 2 s^2) of the sample
__device__ __forceinline__ float packed_total(unsigned long long tot, unsigned long long* rec) {
  double q = (double)(tot >> (kCountBits + kPoisonBits)) * (1.0 / 4294967296.0);
  if ((tot >> kCountBits) & ((1ull << kPoisonBits) - 1)) {
    __threadfence();
    const unsigned long long wide = atomicExch(rec + kWsWide, 0ull);      // read and leave zeroed for the next launch
    const unsigned long long huge = atomicExch(rec + kWsHuge, 0ull);
    if (wide >> 63) return __int_as_float(0x7fc00000);
    q += (double)wide * (1.0 / 16777216.0) + (double)huge;
  }
  return (float)q;
}

// ------------------------------------------------------------------ per-tile arithmetic
// FAM/SRC/ORDER/RND/SDE are compile-time so each instantiation is straight-line code.
template <int FAM, int SRC, int ORDER, bool RND, bool SDE, int N>
__device__ __forceinline__ void tile_math(const mixgrpo_step_coefs& k, const float (&v)[N], const float (&x)[N],
                                           const float (&a)[N], const float (&m1)[N], const float (&m2)[N],
                                           float (&xn)[N], float (&x0)[N], float (&mu)[N], float (&dd)[N]) {
  const float* c = k.c;
  float t[N];
  // x0 = x - sigma*v          (SU:175, SU:226, SU:394)
#pragma unroll
  for (int i = 0; i < N; ++i) t[i] = __fmul_rn(c[0], v[i]);
  round_like_torch<RND>(t);
#pragma unroll
  for (int i = 0; i < N; ++i) x0[i] = __fsub_rn(x[i], t[i]);

  if constexpr (FAM == kFlow) {
    // mean = x*c_x + (v*c_v)*dt   (SU:186)
#pragma unroll
    for (int i = 0; i < N; ++i) t[i] = __fmul_rn(v[i], c[2]);
    round_like_torch<RND>(t);
#pragma unroll
    for (int i = 0; i < N; ++i) t[i] = __fmul_rn(t[i], c[3]);
    round_like_torch<RND>(t);
#pragma unroll
    for (int i = 0; i < N; ++i) mu[i] = __fadd_rn(__fmul_rn(x[i], c[1]), t[i]);
    if constexpr (SRC == MIXGRPO_SRC_NOISE) {            // SU:195
#pragma unroll
      for (int i = 0; i < N; ++i) t[i] = __fmul_rn(c[4], a[i]);
      round_like_torch<RND>(t);
#pragma unroll
      for (int i = 0; i < N; ++i) xn[i] = __fadd_rn(mu[i], t[i]);
    } else if constexpr (SRC == MIXGRPO_SRC_DETERMINISTIC) {   // SU:198-199
#pragma unroll
      for (int i = 0; i < N; ++i) t[i] = __fmul_rn(c[5], v[i]);
      round_like_torch<RND>(t);
#pragma unroll
      for (int i = 0; i < N; ++i) xn[i] = __fadd_rn(x[i], t[i]);
    }
  } else if constexpr (FAM == kDance) {
    // mean = x + dsigma*v       (SU:224)
#pragma unroll
    for (int i = 0; i < N; ++i) t[i] = __fmul_rn(c[1], v[i]);
    round_like_torch<RND>(t);
#pragma unroll
    for (int i = 0; i < N; ++i) mu[i] = __fadd_rn(x[i], t[i]);
    if constexpr (SDE) {         // score / drift correction, SU:231-234
#pragma unroll
      for (int i = 0; i < N; ++i) {
        float s = __fdiv_rn(-__fsub_rn(x[i], __fmul_rn(x0[i], c[2])), c[3]);
        mu[i] = __fadd_rn(mu[i], __fmul_rn(__fmul_rn(s, c[4]), c[5]));
      }
    }
    if constexpr (SRC == MIXGRPO_SRC_NOISE) {            // SU:238
#pragma unroll
      for (int i = 0; i < N; ++i) xn[i] = __fadd_rn(mu[i], __fmul_rn(a[i], c[6]));
    } else if constexpr (SRC == MIXGRPO_SRC_DETERMINISTIC) {   // SU:240
#pragma unroll
      for (int i = 0; i < N; ++i) xn[i] = mu[i];
    }
  } else {  // kDpm: data-prediction multistep, signs folded into the coefficients
    float d1[N], d2[N];
    if constexpr (ORDER == 2) {                          // SU:490
#pragma unroll
      for (int i = 0; i < N; ++i) d1[i] = __fmul_rn(c[1], __fsub_rn(x0[i], m1[i]));
    } else if constexpr (ORDER == 3) {                   // SU:607-610
#pragma unroll
      for (int i = 0; i < N; ++i) {
        float d10 = __fmul_rn(c[1], __fsub_rn(x0[i], m1[i]));
        float d11 = __fmul_rn(c[2], __fsub_rn(m1[i], m2[i]));
        float dd = __fsub_rn(d10, d11);
        d1[i] = __fadd_rn(d10, __fmul_rn(c[3], dd));
        d2[i] = __fmul_rn(c[4], dd);
      }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float m = __fadd_rn(__fmul_rn(c[5], x[i]), __fmul_rn(c[6], x0[i]));
      if constexpr (ORDER >= 2) m = __fadd_rn(m, __fmul_rn(c[7], d1[i]));
      if constexpr (ORDER == 3) m = __fadd_rn(m, __fmul_rn(c[8], d2[i]));
      mu[i] = m;
    }
    if constexpr (SRC == MIXGRPO_SRC_NOISE) {            // SU:434, SU:510, SU:620
#pragma unroll
      for (int i = 0; i < N; ++i) xn[i] = __fadd_rn(mu[i], __fmul_rn(c[13], a[i]));
    } else if constexpr (SRC == MIXGRPO_SRC_DETERMINISTIC) {   // SU:436, SU:516-526, SU:623-628
#pragma unroll
      for (int i = 0; i < N; ++i) {
        float o = __fadd_rn(__fmul_rn(c[9], x[i]), __fmul_rn(c[10], x0[i]));
        if constexpr (ORDER >= 2) o = __fadd_rn(o, __fmul_rn(c[11], d1[i]));
        if constexpr (ORDER == 3) o = __fadd_rn(o, __fmul_rn(c[12], d2[i]));
        xn[i] = o;
      }
    }
  }
  if constexpr (SRC == MIXGRPO_SRC_GIVEN) {
#pragma unroll
    for (int i = 0; i < N; ++i) xn[i] = a[i];
  }
  // squared residual of the transition (SU:202, SU:245, SU:377)
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float d = __fsub_rn(xn[i], mu[i]);
    dd[i] = d * d;
  }
}

}  // namespace mg
