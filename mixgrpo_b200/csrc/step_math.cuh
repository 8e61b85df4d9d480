// Per-tile arithmetic of the three operator families (flow_grpo_step SU:157-210, dance_grpo_step SU:212-253, dpm_step
// SU:273-639; SU = /root/reference/fastvideo/utils/sampling_utils.py), shared by the streaming step kernel
// (step_kernel.cuh) and the single-pass policy kernel (policy_kernels.cu).  Every product/sum the reference performs as
// a separate torch kernel is one __fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn here (no FMA contraction), in the same order,
// with torch's bf16 promotion roundings when RND is set.
#pragma once
#include "common.cuh"

namespace mg {

enum Family { kFlow = 0, kDance = 1, kDpm = 2 };

// packed log-prob accumulator (one 64-bit word per sample): [ sum Q8.32 : 40 | wide-share count : 12 | arrivals : 12 ]
constexpr int kTile = kThreads * kVec;          // scalars per CTA-tile
constexpr int kCountBits = 12, kPoisonBits = 12;
constexpr int kMaxCtasPerSample = (1 << kCountBits) - 1;
// workspace record per sample (in 64-bit words): { accumulator, { epoch : 32 | status : 32 }, wide side accumulator, pad }
// — see mixgrpo_step_workspace_bytes
constexpr int kWsStride = 4;
constexpr int kWsWide = 2;                      // word index of the side accumulator inside a record
constexpr int kWsHuge = 3;                      // word index of the second side accumulator (integer units)
constexpr float kWideCap = 134217728.f;         // 2^27: largest per-CTA share the Q39.24 side accumulator takes (4095 CTAs x 2^27 x 2^24 < 2^63)
constexpr float kHugeCap = 1125899906842624.f;  // 2^50: largest share the integer side accumulator takes (4095 x 2^50 < 2^63)

// The packed word holds shares of mean(d^2 / 2 s^2) up to 255/ctas each — ample for any transition a sane policy
// produces (the value is ~0.5 on the rollout's own samples).  The reference, though, returns a FINITE log-prob however far
// x_next is from the mean (SU:201-208), so a share that does not fit is not dropped: it goes, as Q39.24 fixed point, into
// the record's 64-bit side accumulator (integer adds: still order-independent, still bitwise reproducible) and the CTA
// only flags the fact in the packed word's 12-bit "wide" count.  The finalizer — the CTA that sees the last arrival —
// folds the side word in and re-zeroes it.  A share above 2^27 (|d|/s > 16000) goes, rounded to an integer (relative
// resolution 2^-27), into a second side word; only a non-finite / negative share, or one above 2^50 (|d|/s > 4e7, where
// the reference's own fp32 mean has long lost its digits), yields NaN.  The common path is unchanged: one atomicAdd per
// CTA, no fence.
__device__ __forceinline__ unsigned long long packed_share(float r, int ctas, unsigned long long* rec) {
  const float cap = 255.0f / (float)ctas;
  unsigned long long add = 1ull;
  if (!(r >= 0.f && r <= cap)) {
    if (r > cap && r <= kWideCap) atomicAdd(rec + kWsWide, __float2ull_rn(r * 16777216.0f));
    else if (r > kWideCap && r <= kHugeCap) atomicAdd(rec + kWsHuge, __float2ull_rn(r));
    else atomicOr(rec + kWsWide, 1ull << 63);
    __threadfence();                             // the side word is visible before this CTA's arrival is counted
    add += 1ull << kCountBits;
    r = 0.f;
  }
  return add + (__float2ull_rn(r * 4294967296.0f) << (kCountBits + kPoisonBits));
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
