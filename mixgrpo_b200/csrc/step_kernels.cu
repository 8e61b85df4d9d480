// Fused sampler step + Gaussian transition log-prob (one HBM pass) for the three operator families
// of the reference: flow_grpo_step (SU:157-210), dance_grpo_step (SU:212-253) and dpm_step with its
// order-1/2/3 updates (SU:273-639).  SU = /root/reference/fastvideo/utils/sampling_utils.py.
//
// Data layout: every tensor is (B, n) with n = S*64 packed-latent scalars; thread t of CTA (bx, b)
// owns scalars [8*(tile*256+t), +8) of sample b in EVERY stream, so v (bf16, 16 B), x / noise /
// history / outputs (fp32, 32 B) are each a single LDG.128 / LDG.256 / STG.256 per thread and a
// warp request covers whole 128-B lines.  All loads of the UNROLL tiles a CTA owns are issued before
// any math (memory-level parallelism), results are stored straight from registers; nothing is
// staged in shared memory because no byte is touched twice.
//
// Arithmetic: every product/sum the reference performs as a separate torch kernel is performed here
// with __fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn (no FMA contraction) in the same order, and with the
// same bf16 rounding points when MIXGRPO_FLAG_ROUND_LIKE_TORCH is set, so x_next / x0 / mean are
// bit-identical to the reference on identical inputs.  Only the log-prob reduction order differs.
//
// log-prob: per-thread sum of (x_next-mean)^2 -> warp shuffle -> CTA -> partials[b][bx]; the last CTA
// of a sample (arrival counter) adds the partials in index order, so the result does not depend on
// CTA scheduling, and resets the counter (graph-replay safe).
#include "common.cuh"

namespace mg {

enum Family { kFlow = 0, kDance = 1, kDpm = 2 };

struct StepParams {
  const void* v;
  const float* x;
  const void* noise;
  const float* x_in;
  const float* m1;
  const float* m2;
  float* x_out;
  float* x0_out;
  float* mean_out;
  float* logp_out;
  float* partials;
  unsigned* counters;
  long long n, x_bs, in_bs, out_bs;
  int nblk;
  mixgrpo_step_coefs k;
};

// ------------------------------------------------------------------ per-tile arithmetic
// FAM/SRC/ORDER/RND/SDE are compile-time so each instantiation is straight-line code.
template <int FAM, int SRC, int ORDER, bool RND, bool SDE, int N>
__device__ __forceinline__ float tile_math(const mixgrpo_step_coefs& k, const float (&v)[N], const float (&x)[N],
                                           const float (&a)[N], const float (&m1)[N], const float (&m2)[N],
                                           float (&xn)[N], float (&x0)[N], float (&mu)[N]) {
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
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float d = __fsub_rn(xn[i], mu[i]);
    acc = fmaf(d, d, acc);
  }
  return acc;
}

// ------------------------------------------------------------------ the streaming kernel
template <int FAM, class VT, class NT, int SRC, int ORDER, bool RND, bool SDE, int VEC, int UNROLL>
__global__ void __launch_bounds__(kThreads) step_kernel(const __grid_constant__ StepParams p) {
  __shared__ float s_warp[kThreads / 32];
  __shared__ int s_last;
  const int b = blockIdx.y;
  const long long n = p.n;
  const VT* vp = reinterpret_cast<const VT*>(p.v) + (long long)b * n;
  const float* xp = p.x + (long long)b * p.x_bs;
  const NT* np = reinterpret_cast<const NT*>(p.noise) + (long long)b * n;
  const float* ip = p.x_in + (long long)b * p.in_bs;
  const float* m1p = p.m1 + (long long)b * n;
  const float* m2p = p.m2 + (long long)b * n;

  float v[UNROLL][VEC], x[UNROLL][VEC], a[UNROLL][VEC], m1[UNROLL][VEC], m2[UNROLL][VEC];
  long long idx[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    idx[u] = (((long long)blockIdx.x * UNROLL + u) * kThreads + threadIdx.x) * VEC;
    if (idx[u] < n) {
      ld_stream(vp + idx[u], v[u]);
      ld_stream(xp + idx[u], x[u]);
      if constexpr (SRC == MIXGRPO_SRC_NOISE) ld_stream(np + idx[u], a[u]);
      if constexpr (SRC == MIXGRPO_SRC_GIVEN) ld_stream(ip + idx[u], a[u]);
      if constexpr (FAM == kDpm && ORDER >= 2) ld_stream(m1p + idx[u], m1[u]);
      if constexpr (FAM == kDpm && ORDER == 3) ld_stream(m2p + idx[u], m2[u]);
    }
  }
  float acc = 0.f;
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    if (idx[u] < n) {
      float xn[VEC], x0[VEC], mu[VEC];
      acc += tile_math<FAM, SRC, ORDER, RND, SDE>(p.k, v[u], x[u], a[u], m1[u], m2[u], xn, x0, mu);
      if constexpr (SRC != MIXGRPO_SRC_GIVEN) {
        if (p.x_out) st_stream(p.x_out + (long long)b * p.out_bs + idx[u], xn);
      }
      if (p.x0_out) st_stream(p.x0_out + (long long)b * n + idx[u], x0);
      if (p.mean_out) st_stream(p.mean_out + (long long)b * n + idx[u], mu);
    }
  }
  if (p.logp_out == nullptr) return;

  // deterministic cross-CTA finish
  const float bsum = block_sum(acc, s_warp);
  const int nblk = p.nblk;
  if (threadIdx.x == 0) {
    p.partials[(long long)b * nblk + blockIdx.x] = bsum;
    __threadfence();
    const unsigned ticket = atomicAdd(&p.counters[b], 1u);
    s_last = (ticket == (unsigned)(nblk - 1));
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    float s = 0.f;
    for (int i = threadIdx.x; i < nblk; i += kThreads) s += __ldcg(&p.partials[(long long)b * nblk + i]);
    const float tot = block_sum(s, s_warp);
    if (threadIdx.x == 0) {
      const float msq = __fdiv_rn(tot, (float)n);
      // mean_i[ -(d_i^2)/(2 s^2) - log s - log sqrt(2 pi) ]   (SU:201-208)
      p.logp_out[b] = __fsub_rn(__fsub_rn(__fdiv_rn(-msq, p.k.two_var), p.k.log_scale), p.k.log_norm);
      p.counters[b] = 0u;
    }
  }
}

// ------------------------------------------------------------------ host-side dispatch
static int g_unroll = 2;   // tiles per CTA on the vector path (bench knob, mixgrpo_set_tuning key 0)

static inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

template <int FAM, class VT, class NT, int SRC, int ORDER, bool RND, bool SDE, int VEC, int UNROLL>
static int launch(StepParams& p, int64_t B, cudaStream_t st) {
  const long long per_cta = (long long)kThreads * VEC * UNROLL;
  p.nblk = (int)((p.n + per_cta - 1) / per_cta);
  dim3 grid((unsigned)p.nblk, (unsigned)B);
  step_kernel<FAM, VT, NT, SRC, ORDER, RND, SDE, VEC, UNROLL><<<grid, kThreads, 0, st>>>(p);
  return (int)cudaGetLastError();
}

template <int FAM, class VT, class NT, int SRC, int ORDER, bool RND, bool SDE>
static int pick_width(StepParams& p, int64_t B, bool vec_ok, cudaStream_t st) {
  if (!vec_ok) return launch<FAM, VT, NT, SRC, ORDER, RND, SDE, 1, 4>(p, B, st);
  if constexpr (FAM == kFlow) {
    switch (g_unroll) {
      case 1: return launch<FAM, VT, NT, SRC, ORDER, RND, SDE, kVec, 1>(p, B, st);
      case 4: return launch<FAM, VT, NT, SRC, ORDER, RND, SDE, kVec, 4>(p, B, st);
      default: break;
    }
  }
  return launch<FAM, VT, NT, SRC, ORDER, RND, SDE, kVec, 2>(p, B, st);
}

template <int FAM, class VT, class NT, int ORDER, bool RND, bool SDE>
static int pick_src(StepParams& p, int64_t B, int src, bool vec_ok, cudaStream_t st) {
  switch (src) {
    case MIXGRPO_SRC_NOISE: return pick_width<FAM, VT, NT, MIXGRPO_SRC_NOISE, ORDER, RND, SDE>(p, B, vec_ok, st);
    case MIXGRPO_SRC_GIVEN:
      if constexpr (FAM == kDpm) return MIXGRPO_EINVAL;   // dpm_step has no prev_sample argument (SU:273-284)
      else return pick_width<FAM, VT, NT, MIXGRPO_SRC_GIVEN, ORDER, RND, SDE>(p, B, vec_ok, st);
    case MIXGRPO_SRC_DETERMINISTIC: return pick_width<FAM, VT, NT, MIXGRPO_SRC_DETERMINISTIC, ORDER, RND, SDE>(p, B, vec_ok, st);
  }
  return MIXGRPO_EINVAL;
}

static bool check_common(const void* v, const float* x, int64_t B, int64_t n, int v_dtype, void* ws, int64_t ws_bytes,
                         float* logp, int* err) {
  if (!v || !x || B <= 0 || n <= 0 || B > 65535 || (v_dtype != MIXGRPO_F32 && v_dtype != MIXGRPO_BF16)) {
    *err = MIXGRPO_EINVAL;
    return false;
  }
  if (logp && (!ws || ws_bytes < mixgrpo_step_workspace_bytes(B, n))) {
    *err = ws ? MIXGRPO_ENOSPACE : MIXGRPO_EINVAL;
    return false;
  }
  return true;
}

static void fill(StepParams& p, const void* v, const float* x, int64_t x_bs, const void* noise, const float* x_in,
                 int64_t in_bs, const float* m1, const float* m2, float* x_out, int64_t out_bs, float* x0_out,
                 float* mean_out, float* logp_out, void* ws, int64_t B, int64_t n, const mixgrpo_step_coefs* k) {
  p.v = v; p.x = x; p.noise = noise; p.x_in = x_in; p.m1 = m1; p.m2 = m2;
  p.x_out = x_out; p.x0_out = x0_out; p.mean_out = mean_out; p.logp_out = logp_out;
  p.counters = reinterpret_cast<unsigned*>(ws);
  p.partials = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + ws_counter_bytes(B));
  p.n = n; p.x_bs = x_bs; p.in_bs = in_bs; p.out_bs = out_bs; p.nblk = 0; p.k = *k;
}

// 256-bit path needs 32-B aligned fp32 streams, 16-B aligned bf16 streams and n, strides % 8 == 0.
static bool vector_ok(const StepParams& p, int v_dtype, int noise_dtype, int64_t n) {
  bool ok = (n % kVec == 0) && (p.x_bs % kVec == 0) && (p.in_bs % kVec == 0) && (p.out_bs % kVec == 0);
  ok = ok && aligned(p.v, v_dtype == MIXGRPO_BF16 ? 16 : 32) && aligned(p.x, 32);
  ok = ok && aligned(p.noise, noise_dtype == MIXGRPO_BF16 ? 16 : 32) && aligned(p.x_in, 32);
  ok = ok && aligned(p.m1, 32) && aligned(p.m2, 32) && aligned(p.x_out, 32) && aligned(p.x0_out, 32) && aligned(p.mean_out, 32);
  return ok;
}

}  // namespace mg

using namespace mg;

extern "C" __attribute__((visibility("default"))) int64_t mixgrpo_step_workspace_bytes(int64_t B, int64_t n) {
  if (B <= 0 || n <= 0) return 0;
  const int64_t nblk_max = (n + 1023) / 1024;   // smallest CTA footprint (scalar path: 256 thr x 1 x 4)
  return ws_counter_bytes(B) + B * nblk_max * (int64_t)sizeof(float);
}

extern "C" __attribute__((visibility("default"))) int mixgrpo_set_tuning(int key, int value) {
  if (key == 0) {
    if (value != 1 && value != 2 && value != 4) return MIXGRPO_EINVAL;
    int old = g_unroll;
    g_unroll = value;
    return old;
  }
  return MIXGRPO_EINVAL;
}

extern "C" __attribute__((visibility("default"))) int mixgrpo_flow_step(const void* v, int v_dtype, const float* x, int64_t x_bs, const void* noise,
                                 const float* x_next_in, int64_t in_bs, float* x_next_out, int64_t out_bs,
                                 float* x0_out, float* mean_out, float* logp_out, void* workspace,
                                 int64_t workspace_bytes, int64_t B, int64_t n, const mixgrpo_step_coefs* coefs_host,
                                 int src, unsigned flags, void* stream) {
  int err = 0;
  if (!coefs_host || !check_common(v, x, B, n, v_dtype, workspace, workspace_bytes, logp_out, &err)) return err ? err : MIXGRPO_EINVAL;
  if ((src == MIXGRPO_SRC_NOISE && !noise) || (src == MIXGRPO_SRC_GIVEN && !x_next_in)) return MIXGRPO_EINVAL;
  StepParams p;
  fill(p, v, x, x_bs, noise, x_next_in, in_bs, nullptr, nullptr, x_next_out, out_bs, x0_out, mean_out, logp_out, workspace, B, n, coefs_host);
  const bool vec = vector_ok(p, v_dtype, v_dtype, n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (v_dtype == MIXGRPO_F32) return pick_src<kFlow, float, float, 1, false, false>(p, B, src, vec, st);
  if (flags & MIXGRPO_FLAG_ROUND_LIKE_TORCH) return pick_src<kFlow, __nv_bfloat16, __nv_bfloat16, 1, true, false>(p, B, src, vec, st);
  return pick_src<kFlow, __nv_bfloat16, __nv_bfloat16, 1, false, false>(p, B, src, vec, st);
}

extern "C" __attribute__((visibility("default"))) int mixgrpo_dance_step(const void* v, int v_dtype, const float* x, int64_t x_bs, const float* noise,
                                  const float* x_next_in, int64_t in_bs, float* x_next_out, int64_t out_bs,
                                  float* x0_out, float* mean_out, float* logp_out, void* workspace,
                                  int64_t workspace_bytes, int64_t B, int64_t n, const mixgrpo_step_coefs* coefs_host,
                                  int src, int sde_solver, unsigned flags, void* stream) {
  int err = 0;
  if (!coefs_host || !check_common(v, x, B, n, v_dtype, workspace, workspace_bytes, logp_out, &err)) return err ? err : MIXGRPO_EINVAL;
  if ((src == MIXGRPO_SRC_NOISE && !noise) || (src == MIXGRPO_SRC_GIVEN && !x_next_in)) return MIXGRPO_EINVAL;
  StepParams p;
  fill(p, v, x, x_bs, noise, x_next_in, in_bs, nullptr, nullptr, x_next_out, out_bs, x0_out, mean_out, logp_out, workspace, B, n, coefs_host);
  const bool vec = vector_ok(p, v_dtype, MIXGRPO_F32, n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool rnd = (flags & MIXGRPO_FLAG_ROUND_LIKE_TORCH) != 0;
  if (v_dtype == MIXGRPO_F32) {
    return sde_solver ? pick_src<kDance, float, float, 1, false, true>(p, B, src, vec, st)
                      : pick_src<kDance, float, float, 1, false, false>(p, B, src, vec, st);
  }
  if (rnd) {
    return sde_solver ? pick_src<kDance, __nv_bfloat16, float, 1, true, true>(p, B, src, vec, st)
                      : pick_src<kDance, __nv_bfloat16, float, 1, true, false>(p, B, src, vec, st);
  }
  return sde_solver ? pick_src<kDance, __nv_bfloat16, float, 1, false, true>(p, B, src, vec, st)
                    : pick_src<kDance, __nv_bfloat16, float, 1, false, false>(p, B, src, vec, st);
}

template <int ORDER>
static int dpm_dispatch(StepParams& p, int v_dtype, int64_t B, int src, bool vec, bool rnd, cudaStream_t st) {
  if (v_dtype == MIXGRPO_F32) return pick_src<kDpm, float, float, ORDER, false, false>(p, B, src, vec, st);
  if (rnd) return pick_src<kDpm, __nv_bfloat16, float, ORDER, true, false>(p, B, src, vec, st);
  return pick_src<kDpm, __nv_bfloat16, float, ORDER, false, false>(p, B, src, vec, st);
}

extern "C" __attribute__((visibility("default"))) int mixgrpo_dpm_step(const void* v, int v_dtype, const float* x, int64_t x_bs, const float* noise,
                                const float* m1, const float* m2, int order, float* x_next_out, int64_t out_bs,
                                float* x0_out, float* mean_out, float* logp_out, void* workspace,
                                int64_t workspace_bytes, int64_t B, int64_t n, const mixgrpo_step_coefs* coefs_host,
                                int src, unsigned flags, void* stream) {
  int err = 0;
  if (!coefs_host || !check_common(v, x, B, n, v_dtype, workspace, workspace_bytes, logp_out, &err)) return err ? err : MIXGRPO_EINVAL;
  if (order < 1 || order > 3 || (order >= 2 && !m1) || (order == 3 && !m2)) return MIXGRPO_EINVAL;
  if (src == MIXGRPO_SRC_NOISE && !noise) return MIXGRPO_EINVAL;
  StepParams p;
  fill(p, v, x, x_bs, noise, nullptr, n, m1, m2, x_next_out, out_bs, x0_out, mean_out, logp_out, workspace, B, n, coefs_host);
  const bool vec = vector_ok(p, v_dtype, MIXGRPO_F32, n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool rnd = (flags & MIXGRPO_FLAG_ROUND_LIKE_TORCH) != 0;
  switch (order) {
    case 1: return dpm_dispatch<1>(p, v_dtype, B, src, vec, rnd, st);
    case 2: return dpm_dispatch<2>(p, v_dtype, B, src, vec, rnd, st);
    default: return dpm_dispatch<3>(p, v_dtype, B, src, vec, rnd, st);
  }
}
