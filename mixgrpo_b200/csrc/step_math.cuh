// Per-tile arithmetic of the three operator families (flow_grpo_step SU:157-210, dance_grpo_step SU:212-253, dpm_step
// SU:273-639; SU = /root/reference/fastvideo/utils/sampling_utils.py), shared by the streaming step kernel
// (step_kernel.cuh) and the single-pass policy kernel (policy_kernels.cu).  Every product/sum the reference performs as
// a separate torch kernel is one __fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn here (no FMA contraction), in the same order,
// with torch's bf16 promotion roundings when RND is set.
#pragma once
#include "common.cuh"

namespace mg {

enum Family { kFlow = 0, kDance = 1, kDpm = 2 };

// packed log-prob accumulator (one 64-bit word per sample): [ sum Q8.32 : 40 | wide flag count : 12 | arrivals : 12 ]
constexpr int kTile = kThreads * kVec;          // scalars per CTA-tile of the 256-thread kernels
constexpr int kHalfThreads = kThreads / 2;      // the 128-thread (4-warp) CTA of the deferred rollout kernels ...
constexpr int kHalfTile = kHalfThreads * kVec;  // ... owns one half-tile: 1024 consecutive scalars
constexpr int kCountBits = 12, kPoisonBits = 12;
constexpr unsigned long long kArrivalMask = (1ull << kCountBits) - 1;
constexpr unsigned long long kWideMask = ((1ull << kPoisonBits) - 1) << kCountBits;
// a sample's CTAs each add one arrival and at most one wide flag; 128-thread CTAs are twice as many: <= 2047 256-thread CTAs per sample
constexpr int kMaxCtasPerSample = (1 << (kCountBits - 1)) - 1;
// workspace record per sample (in 64-bit words): { accumulator, { epoch : 32 | status : 32 }, wide side accumulator, integer side
// accumulator } — see mixgrpo_step_workspace_bytes.  A DEFERRED launch (MIXGRPO_FLAG_DEFER_LOGP) spreads a sample's arrivals over
// kDeferSubs such records (mixgrpo_deferred_workspace_bytes): 3072 fire-and-forget reductions into 12 addresses serialise in the
// L2 (measured: +0.5 us per launch), into 12 x 8 they do not; mixgrpo_logp_finalize adds the sub-records up (integers: exact).
constexpr int kWsStride = 4;
constexpr int kWsWide = 2;                      // word index of the side accumulator inside a record
constexpr int kWsHuge = 3;                      // word index of the second side accumulator (integer units)
constexpr int kDeferSubs = 8;
constexpr float kWideCap = 16777216.f;          // 2^24: largest warp part the Q39.24 side accumulator takes (16376 warps x 2^24 x 2^24 < 2^63)
constexpr float kHugeCap = 281474976710656.f;   // 2^48: largest warp part the integer side accumulator takes (16376 x 2^48 < 2^63)

// The log-prob reduction is INTEGER from the thread up.  A thread's sum of d^2 over its 8 scalars (fp32, pair-wise, fixed
// order) is scaled by 2^32 / (n * 2 s^2) and rounded to an unsigned integer — its share of mean(d^2 / 2 s^2) in units of 2^-32
// — and everything above is integer addition, which commutes: ONE redux.sync.add.u32 per warp instead of five dependent
// shuffles, a 64-bit word per warp through shared memory, one atomic (or fire-and-forget reduction) per CTA.  So any CTA
// shape, any arrival order and any of the kernels that evaluate the same transition — the 256-thread step / policy kernels, the
// 128-thread deferred rollout shape, the batched window kernel, the single-pass policy kernel — give bit-identical log-probs.
// Rounding per thread costs <= 0.5 unit each: 0.29 * sqrt(n/8) units in total = 1.2e-8 absolute on a log-prob at 1024^2.
//
// A thread's share must stay below `cap` = min(2^26, 255 * 2^32 / threads-per-sample): 32 lanes fit the 32-bit redux and a
// whole sample fits the 40-bit field (total < 255).  That is |x' - mean| / s up to ~8..23 on EVERY scalar of a thread — ample for
// any transition a sane policy produces (the value is ~1 on the rollout's own samples).  The reference, though, returns a FINITE
// log-prob however far x_next is from the mean (SU:201-208), so a warp with a thread beyond the cap is not dropped: its fp32
// sum goes, as Q39.24 fixed point, into the record's 64-bit side accumulator (integer adds: still order-independent, still
// bitwise reproducible), above 2^24 rounded to an integer into a second side word, and the warp raises the word's "wide"
// flag; the finalizer folds the side words in and re-zeroes them.  Only a non-finite sum, or one above 2^48 (|d|/s > 1e8,
// where the reference's own fp32 mean has long lost its digits), yields NaN.  The common path has no branch but the vote.
struct LpQuant { float scale, cap, denom; };
// evaluated ONCE per launch on the host (IEEE single operations, the same three for every kernel) and passed by value: two
// divisions per thread are 15 % of a streaming thread's instructions
inline LpQuant lp_quant(long long n, float two_var) {
  LpQuant q;
  q.denom = (float)n * two_var;
  q.scale = 4294967296.f / q.denom;
  const float fit = 1095216660480.f / (float)((n + kVec - 1) / kVec);                      // 255 * 2^32 / threads per sample
  q.cap = fit < 67108864.f ? fit : 67108864.f;
  return q;
}

// All 32 lanes of a warp: acc = the thread's sum of d^2.  Returns (in every lane) the warp's part of the packed word — the
// fixed-point sum in the upper 40 bits, or zero there and one wide flag when the warp went to the side accumulators of `rec`.
__device__ __forceinline__ unsigned long long warp_part(float acc, const LpQuant& q, unsigned long long* rec) {
  const float u = __fmul_rn(acc, q.scale);
  if (__all_sync(0xffffffffu, u < q.cap))                                         // NaN fails the test
    return (unsigned long long)__reduce_add_sync(0xffffffffu, __float2uint_rn(u)) << (kCountBits + kPoisonBits);
  const float r = __fdiv_rn(warp_sum(acc), q.denom);                              // far transition: fp32 warp sum, side words
  if ((threadIdx.x & 31) == 0) {
    if (r >= 0.f && r <= kWideCap) atomicAdd(rec + kWsWide, __float2ull_rn(r * 16777216.0f));
    else if (r > kWideCap && r <= kHugeCap) atomicAdd(rec + kWsHuge, __float2ull_rn(r));
    else atomicOr(rec + kWsWide, 1ull << 63);
    __threadfence();                             // the side word is visible before this CTA's arrival is counted
  }
  return 1ull << kCountBits;
}

// One thread: the word a CTA adds to its sample's record from its WARPS warp parts — their integer sum, at most ONE wide flag
// (flags only say "read the side words"; counting CTAs keeps the 12-bit field from overflowing), one arrival.
template <int WARPS>
__device__ __forceinline__ unsigned long long cta_word(const unsigned long long* parts) {
  unsigned long long t = 0ull;
#pragma unroll
  for (int w = 0; w < WARPS; ++w) t += parts[w];
  return (t & ~kWideMask) + ((t & kWideMask) ? (1ull << kCountBits) : 0ull) + 1ull;
}

// tot = the packed word after the LAST arrival; returns mean(d^2 / 2 s^2) of the sample
__device__ __forceinline__ float packed_total(unsigned long long tot, unsigned long long* rec) {
  double q = (double)(tot >> (kCountBits + kPoisonBits)) * (1.0 / 4294967296.0);
  if (tot & kWideMask) {
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
