// Per-tile arithmetic of the three operator families (flow_grpo_step SU:157-210, dance_grpo_step SU:212-253, dpm_step
// SU:273-639; SU = /root/reference/fastvideo/utils/sampling_utils.py), shared by the streaming step kernel
// (step_kernels.cu) and the single-pass policy kernel (policy_kernels.cu).  Every product/sum the reference performs as
// a separate torch kernel is one __fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn here (no FMA contraction), in the same order,
// with torch's bf16 promotion roundings when RND is set.
#pragma once
#include "common.cuh"

namespace mg {

enum Family { kFlow = 0, kDance = 1, kDpm = 2 };

// packed log-prob accumulator (one 64-bit word per sample): [ sum Q8.32 : 40 | poison : 12 | arrivals : 12 ]
constexpr int kTile = kThreads * kVec;          // scalars per CTA-tile
constexpr int kCountBits = 12, kPoisonBits = 12;
constexpr int kMaxCtasPerSample = (1 << kCountBits) - 1;
// workspace record per sample (in 64-bit words): { accumulator, { epoch : 32 | status : 32 } } — see mixgrpo_step_workspace_bytes
constexpr int kWsStride = 2;

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
