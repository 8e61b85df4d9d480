// Fused sampler step + Gaussian transition log-prob (one HBM pass) for the three operator families
// of the reference: flow_grpo_step (SU:157-210), dance_grpo_step (SU:212-253) and dpm_step with its
// order-1/2/3 updates (SU:273-639).  SU = /root/reference/fastvideo/utils/sampling_utils.py.
//
// Data layout: every tensor is (B, n) with n = S*64 packed-latent scalars; thread t of a CTA owns
// scalars [8t, 8t+8) of the CTA's 2048-scalar tile in EVERY stream, so v (bf16, 16 B), x / noise /
// history / outputs (fp32, 32 B) are each a single LDG.128 / LDG.256 / STG.256 per thread and a warp
// request covers whole 128-B lines with every sector used.  Results are stored straight from registers;
// nothing is staged in shared memory because no byte is touched twice.
//
// Arithmetic: every product/sum the reference performs as a separate torch kernel is performed here
// with __fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn (no FMA contraction) in the same order, and with the
// same bf16 rounding points when MIXGRPO_FLAG_ROUND_LIKE_TORCH is set, so x_next / x0 / mean are
// bit-identical to the reference on identical inputs.  Only the log-prob reduction order differs.
//
// log-prob: per-thread sum of (x_next-mean)^2 -> warp shuffle -> CTA -> ONE packed fixed-point atomicAdd per
// CTA (count + sum in a 64-bit word): order-independent, hence bitwise reproducible; the last arriver
// writes logp[b] and re-zeroes the word (graph-replay safe).  Details at step_kernel below.
#include "step_math.cuh"

namespace mg {

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
  unsigned long long* acc;
  long long n, x_bs, in_bs, out_bs;
  int B, tiles;
  mixgrpo_step_coefs k;
  unsigned long long philox_seed, philox_offset;   // SRC_PHILOX
  LossParams loss;          // fused policy path (SRC_GIVEN only): old log-probs / advantages / stats rows, or nullptrs
};

// ------------------------------------------------------------------ the streaming kernel
// Grid (ctas_per_sample, B); CTA = 256 threads; a CTA-tile is 2048 consecutive scalars of one sample and
// thread t owns scalars [8t, 8t+8) of it in every stream (one LDG.128 for bf16, one LDG.256 for fp32, one
// STG.256 per output).  Measured on B200 (tools/ubench.cu, profiles/r01_design_space.md): what matters for
// this 50-100 MB pass is bytes in flight — all of a thread's loads are issued before any math, the math
// is done pair-wise so the kernel stays at <= 40 registers (6 CTAs = 1536 threads per SM), and CTAs retire
// without waiting on anything (a persistent / software-pipelined variant and every fence+ticket
// reduction were 20-30 % slower).
//
// log-prob reduction — one atomic per CTA, deterministic, no fence:
//   acc[b] is a 64-bit word  [ sum : 40 bit fixed point Q8.32 | poison : 12 | arrivals : 12 ].
//   A CTA adds  (1, poison?, round(r * 2^32))  with r = sum_cta(d^2) / (n * 2 s^2)  in ONE atomicAdd.
//   Integer addition commutes, so the total is bit-identical whatever order CTAs arrive in; the CTA
//   whose returned count is the last one owns the complete sum in (old + mine), writes
//   logp[b] = -sum - log s - log sqrt(2 pi) and zeroes the word for the next launch.
//   Resolution 2^-32 per CTA (<= 1.5e-8 absolute on logp at 1024^2); a contribution that is not finite
//   or would overflow the field (mean squared normalised residual > 510) poisons the sample -> NaN.

template <class T, bool VECTOR>
__device__ __forceinline__ void load_tile(const T* base, long long off, long long n, float (&r)[kVec]) {
  if constexpr (VECTOR) {
    ld_stream(base + off + threadIdx.x * kVec, r);          // caller guarantees off + 8t < n
  } else {
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
      const long long i = off + j * kThreads + threadIdx.x;
      float one[1] = {0.f};
      if (i < n) ld_stream(base + i, one);
      r[j] = one[0];
    }
  }
}

template <bool VECTOR>
__device__ __forceinline__ void store_tile(float* base, long long off, long long n, const float (&r)[kVec]) {
  if constexpr (VECTOR) {
    st_stream(base + off + threadIdx.x * kVec, r);
  } else {
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
      const long long i = off + j * kThreads + threadIdx.x;
      if (i < n) base[i] = r[j];
    }
  }
}

// OUT (compile time): 0 = neither x0 nor mean is stored (the rollout driver's steps: x0 is dead code), 1 = x0, 2 = x0 + mean
template <int FAM, class VT, class NT, int SRC, int ORDER, bool RND, bool SDE, bool VECTOR, int OUT>
__global__ void __launch_bounds__(kThreads, ((FAM == kDpm && ORDER >= 2) || OUT == 2 || !VECTOR || (FAM == kDance && SDE)) ? 4 : (SRC == MIXGRPO_SRC_PHILOX ? 5 : 6))
step_kernel(const __grid_constant__ StepParams p) {
  pdl_prologue();
  const int b = blockIdx.y;
  const long long n = p.n;
  const VT* vp = reinterpret_cast<const VT*>(p.v) + (long long)b * n;
  const float* xp = p.x + (long long)b * p.x_bs;
  float acc = 0.f;

  for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
    const long long off = (long long)tile * kTile;
    // vector path: n % 8 == 0, so a thread's 8 scalars are all inside or all outside the sample
    if (VECTOR && off + threadIdx.x * kVec >= n) continue;
    float v[kVec], x[kVec], a[kVec], m1[kVec], m2[kVec];
    load_tile<VT, VECTOR>(vp, off, n, v);
    load_tile<float, VECTOR>(xp, off, n, x);
    if constexpr (SRC == MIXGRPO_SRC_NOISE) load_tile<NT, VECTOR>(reinterpret_cast<const NT*>(p.noise) + (long long)b * n, off, n, a);
    if constexpr (SRC == MIXGRPO_SRC_PHILOX) {     // draw the noise here: element e -> component e%4 of Philox(e/4)
      if constexpr (VECTOR) {
        const unsigned long long e0 = (unsigned long long)b * n + off + threadIdx.x * kVec;
        float z0[4], z1[4];
        philox_normal4(e0 >> 2, p.philox_seed, p.philox_offset, z0);
        philox_normal4((e0 >> 2) + 1, p.philox_seed, p.philox_offset, z1);
#pragma unroll
        for (int j = 0; j < 4; ++j) { a[j] = z0[j]; a[4 + j] = z1[j]; }
      } else {
#pragma unroll
        for (int j = 0; j < kVec; ++j) {
          const unsigned long long e = (unsigned long long)b * n + off + j * kThreads + threadIdx.x;
          float z[4];
          philox_normal4(e >> 2, p.philox_seed, p.philox_offset, z);
          a[j] = z[e & 3];
        }
      }
      if constexpr (sizeof(NT) == 2) round_like_torch<true>(a);        // the noise tensor itself is bf16 (SU:193)
    }
    if constexpr (SRC == MIXGRPO_SRC_GIVEN) load_tile<float, VECTOR>(p.x_in + (long long)b * p.in_bs, off, n, a);
    if constexpr (FAM == kDpm && ORDER >= 2) load_tile<float, VECTOR>(p.m1 + (long long)b * n, off, n, m1);
    if constexpr (FAM == kDpm && ORDER == 3) load_tile<float, VECTOR>(p.m2 + (long long)b * n, off, n, m2);

    float xn[kVec], x0[OUT >= 1 ? kVec : 2], mu[OUT == 2 ? kVec : 2];
#pragma unroll
    for (int j = 0; j < kVec; j += 2) {           // pair-wise: short live ranges, packed bf16 rounding
      const float v2[2] = {v[j], v[j + 1]}, x2[2] = {x[j], x[j + 1]}, a2[2] = {a[j], a[j + 1]};
      const float m12[2] = {m1[j], m1[j + 1]}, m22[2] = {m2[j], m2[j + 1]};
      float xn2[2], x02[2], mu2[2], dd2[2];
      tile_math<FAM, (SRC == MIXGRPO_SRC_PHILOX ? MIXGRPO_SRC_NOISE : SRC), ORDER, RND, SDE>(p.k, v2, x2, a2, m12, m22, xn2, x02, mu2, dd2);
      xn[j] = xn2[0]; xn[j + 1] = xn2[1];
      if constexpr (OUT >= 1) { x0[j] = x02[0]; x0[j + 1] = x02[1]; }
      if constexpr (OUT == 2) { mu[j] = mu2[0]; mu[j + 1] = mu2[1]; }
      if constexpr (VECTOR) {
        acc += dd2[0] + dd2[1];
      } else {                                    // ragged tail: mask scalars beyond the sample
        if (off + j * kThreads + threadIdx.x < n) acc += dd2[0];
        if (off + (j + 1) * kThreads + threadIdx.x < n) acc += dd2[1];
      }
    }
    if constexpr (SRC != MIXGRPO_SRC_GIVEN) {
      if (p.x_out) store_tile<VECTOR>(p.x_out + (long long)b * p.out_bs, off, n, xn);
    }
    if constexpr (OUT >= 1) store_tile<VECTOR>(p.x0_out + (long long)b * n, off, n, x0);
    if constexpr (OUT == 2) store_tile<VECTOR>(p.mean_out + (long long)b * n, off, n, mu);
  }
  if (p.logp_out == nullptr) return;

  __shared__ float s_warp[kThreads / 32];
  acc = warp_sum(acc);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_warp[warp] = acc;
  __syncthreads();
  if (warp != 0) return;                           // warps 1..7 are done: nothing waits on the atomic
  float t = lane < kThreads / 32 ? s_warp[lane] : 0.f;
  t = warp_sum(t);
  if (lane == 0) {
    const int ctas = gridDim.x;
    float r = __fdiv_rn(t, __fmul_rn((float)n, p.k.two_var));
    const float cap = 255.0f / (float)ctas;
    unsigned long long add = 1ull;
    if (!(r >= 0.f && r <= cap)) {                 // NaN, inf, negative (two_var < 0) or would overflow
      add += 1ull << kCountBits;
      r = 0.f;
    }
    add += __float2ull_rn(r * 4294967296.0f) << (kCountBits + kPoisonBits);
    const unsigned long long old = atomicAdd(&p.acc[kWsStride * b], add);
    if ((old & (unsigned long long)kMaxCtasPerSample) == (unsigned long long)(ctas - 1)) {
      const unsigned long long tot = old + add;
      float q = (float)((double)(tot >> (kCountBits + kPoisonBits)) * (1.0 / 4294967296.0));
      if ((tot >> kCountBits) & ((1ull << kPoisonBits) - 1)) q = __int_as_float(0x7fc00000);
      // mean_i[ -(d_i^2)/(2 s^2) - log s - log sqrt(2 pi) ]   (SU:201-208)
      const float lp = __fsub_rn(__fsub_rn(-q, p.k.log_scale), p.k.log_norm);
      p.logp_out[b] = lp;
      p.acc[kWsStride * b] = 0ull;
      if constexpr (SRC == MIXGRPO_SRC_GIVEN) {
        // fused policy path: the sample's clipped-ratio loss terms (TR:560-583, one sample = the reference's B == 1)
        // are added to its own stats row by this single thread — ordered across launches, no extra kernel
        if (p.loss.rows) {
          const LossTerms t = loss_terms(lp, p.loss.old_lp[b], p.loss.adv[b], p.loss, 1.f);
          const float policy = __fdiv_rn(t.policy_num, p.loss.denom);
          const float kl = __fdiv_rn(__fmul_rn(0.5f, t.kl_num), p.loss.denom);
          float* row = p.loss.rows + 4 * (long long)b;
          const float4 prev = p.loss.accumulate ? *reinterpret_cast<const float4*>(row) : make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(row) = make_float4(prev.x + __fadd_rn(policy, __fmul_rn(p.loss.klc, kl)), prev.y + policy,
                                                        prev.z + kl, prev.w + t.clip);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ host-side dispatch
static int g_max_ctas_per_sample = kMaxCtasPerSample;   // bench knob (mixgrpo_set_tuning key 0)
int g_use_pdl = 1;                                      // bench knob (key 1): programmatic dependent launch on/off

static inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

template <int FAM, class VT, class NT, int SRC, int ORDER, bool RND, bool SDE, bool VECTOR, int OUT>
static int launch(StepParams& p, cudaStream_t st) {
  p.tiles = (int)((p.n + kTile - 1) / kTile);
  int ctas = p.tiles < g_max_ctas_per_sample ? p.tiles : g_max_ctas_per_sample;
  dim3 grid((unsigned)ctas, (unsigned)p.B);
  launch_pdl(step_kernel<FAM, VT, NT, SRC, ORDER, RND, SDE, VECTOR, OUT>, grid, kThreads, 0, st, p);
  return (int)cudaGetLastError();
}

template <int FAM, class VT, class NT, int SRC, int ORDER, bool RND, bool SDE, int OUT>
static int pick_vec(StepParams& p, bool vec_ok, cudaStream_t st) {
  if (!vec_ok) return launch<FAM, VT, NT, SRC, ORDER, RND, SDE, false, OUT>(p, st);
  return launch<FAM, VT, NT, SRC, ORDER, RND, SDE, true, OUT>(p, st);
}

template <int FAM, class VT, class NT, int SRC, int ORDER, bool RND, bool SDE>
static int pick_width(StepParams& p, int64_t, bool vec_ok, cudaStream_t st) {
  if (p.mean_out && !p.x0_out) return MIXGRPO_EINVAL;     // the mean is only offered together with x0 (SU:210)
  if constexpr (FAM == kDpm) {                             // dpm_step never returns the mean (SU:385)
    if (p.mean_out) return MIXGRPO_EINVAL;
    if (p.x0_out) return pick_vec<FAM, VT, NT, SRC, ORDER, RND, SDE, 1>(p, vec_ok, st);
    return pick_vec<FAM, VT, NT, SRC, ORDER, RND, SDE, 0>(p, vec_ok, st);
  } else {
    if (p.mean_out) return pick_vec<FAM, VT, NT, SRC, ORDER, RND, SDE, 2>(p, vec_ok, st);
    if (p.x0_out) return pick_vec<FAM, VT, NT, SRC, ORDER, RND, SDE, 1>(p, vec_ok, st);
    return pick_vec<FAM, VT, NT, SRC, ORDER, RND, SDE, 0>(p, vec_ok, st);
  }
}

template <int FAM, class VT, class NT, int ORDER, bool RND, bool SDE>
static int pick_src(StepParams& p, int64_t B, int src, bool vec_ok, cudaStream_t st) {
  switch (src) {
    case MIXGRPO_SRC_NOISE: return pick_width<FAM, VT, NT, MIXGRPO_SRC_NOISE, ORDER, RND, SDE>(p, B, vec_ok, st);
    case MIXGRPO_SRC_GIVEN:
      if constexpr (FAM == kDpm) return MIXGRPO_EINVAL;   // dpm_step has no prev_sample argument (SU:273-284)
      else return pick_width<FAM, VT, NT, MIXGRPO_SRC_GIVEN, ORDER, RND, SDE>(p, B, vec_ok, st);
    case MIXGRPO_SRC_DETERMINISTIC: return pick_width<FAM, VT, NT, MIXGRPO_SRC_DETERMINISTIC, ORDER, RND, SDE>(p, B, vec_ok, st);
    case MIXGRPO_SRC_PHILOX: return pick_width<FAM, VT, NT, MIXGRPO_SRC_PHILOX, ORDER, RND, SDE>(p, B, vec_ok, st);
  }
  return MIXGRPO_EINVAL;
}

static bool check_common(const void* v, const float* x, int64_t B, int64_t n, int v_dtype, void* ws, int64_t ws_bytes,
                         float* logp, int* err) {
  // grid.y carries the sample index; tile indices are 32-bit inside the kernel
  if (!v || !x || B <= 0 || B > 65535 || n <= 0 || (v_dtype != MIXGRPO_F32 && v_dtype != MIXGRPO_BF16) ||
      (n + kTile - 1) / kTile >= 2147483647LL) {
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
  p.acc = reinterpret_cast<unsigned long long*>(ws);
  p.n = n; p.x_bs = x_bs; p.in_bs = in_bs; p.out_bs = out_bs;
  p.B = (int)B; p.tiles = 0; p.k = *k;
  p.philox_seed = p.philox_offset = 0ull;
  p.loss = LossParams{nullptr, nullptr, nullptr, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 1};
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
  // one 16-byte record per sample: { packed 64-bit accumulator | 32-bit epoch | 32-bit status (record 0) } — the layout
  // does not depend on B, so calls with different batch sizes can share one zero-initialised allocation
  return ((B * (int64_t)(kWsStride * sizeof(unsigned long long)) + 255) / 256) * 256;
}

extern "C" __attribute__((visibility("default"))) int mixgrpo_set_tuning(int key, int value) {
  if (key == 2) return value < 0 ? MIXGRPO_EINVAL : mixgrpo_peer_set_timeout_ms(value);
  if (key >= 3 && key <= 5) return mixgrpo_policy_set_tuning(key, value);
  if (key == 1) {
    if (value != 0 && value != 1) return MIXGRPO_EINVAL;
    const int old = g_use_pdl;
    g_use_pdl = value;
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
                                 int src, unsigned flags, void* stream) {
  int err = 0;
  if (!coefs_host || !check_common(v, x, B, n, v_dtype, workspace, workspace_bytes, logp_out, &err)) return err ? err : MIXGRPO_EINVAL;
  if (((src == MIXGRPO_SRC_NOISE || src == MIXGRPO_SRC_PHILOX) && !noise) || (src == MIXGRPO_SRC_GIVEN && !x_next_in)) return MIXGRPO_EINVAL;
  StepParams p;
  fill(p, v, x, x_bs, src == MIXGRPO_SRC_PHILOX ? nullptr : noise, x_next_in, in_bs, nullptr, nullptr, x_next_out, out_bs, x0_out, mean_out, logp_out, workspace, B, n, coefs_host);
  if (src == MIXGRPO_SRC_PHILOX) { const mixgrpo_philox_args* ph = static_cast<const mixgrpo_philox_args*>(noise); p.philox_seed = ph->seed; p.philox_offset = ph->offset; }
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
  if (((src == MIXGRPO_SRC_NOISE || src == MIXGRPO_SRC_PHILOX) && !noise) || (src == MIXGRPO_SRC_GIVEN && !x_next_in)) return MIXGRPO_EINVAL;
  StepParams p;
  fill(p, v, x, x_bs, src == MIXGRPO_SRC_PHILOX ? nullptr : noise, x_next_in, in_bs, nullptr, nullptr, x_next_out, out_bs, x0_out, mean_out, logp_out, workspace, B, n, coefs_host);
  if (src == MIXGRPO_SRC_PHILOX) { const mixgrpo_philox_args* ph = reinterpret_cast<const mixgrpo_philox_args*>(noise); p.philox_seed = ph->seed; p.philox_offset = ph->offset; }
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
  if ((src == MIXGRPO_SRC_NOISE || src == MIXGRPO_SRC_PHILOX) && !noise) return MIXGRPO_EINVAL;
  StepParams p;
  fill(p, v, x, x_bs, src == MIXGRPO_SRC_PHILOX ? nullptr : noise, nullptr, n, m1, m2, x_next_out, out_bs, x0_out, mean_out, logp_out, workspace, B, n, coefs_host);
  if (src == MIXGRPO_SRC_PHILOX) { const mixgrpo_philox_args* ph = reinterpret_cast<const mixgrpo_philox_args*>(noise); p.philox_seed = ph->seed; p.philox_offset = ph->offset; }
  const bool vec = vector_ok(p, v_dtype, MIXGRPO_F32, n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool rnd = (flags & MIXGRPO_FLAG_ROUND_LIKE_TORCH) != 0;
  switch (order) {
    case 1: return dpm_dispatch<1>(p, v_dtype, B, src, vec, rnd, st);
    case 2: return dpm_dispatch<2>(p, v_dtype, B, src, vec, rnd, st);
    default: return dpm_dispatch<3>(p, v_dtype, B, src, vec, rnd, st);
  }
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
  if (loss && (!loss->old_logp || !loss->advantages)) return MIXGRPO_EINVAL;
  StepParams p;
  fill(p, v, x, x_bs, nullptr, x_next, in_bs, nullptr, nullptr, nullptr, n, nullptr, nullptr, logp_out, workspace, B, n, coefs_host);
  if (loss) {
    if (loss->stats_rows && (reinterpret_cast<uintptr_t>(loss->stats_rows) % 16) != 0) return MIXGRPO_EINVAL;   // rows are float4
    p.loss = make_loss_params(loss->old_logp, loss->advantages, loss->stats_rows, loss->clip_range, loss->adv_clip_max, loss->kl_coeff, loss->denom);
    p.loss.accumulate = loss->accumulate ? 1 : 0;
  }
  const bool vec = vector_ok(p, v_dtype, MIXGRPO_F32, n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool rnd = (flags & MIXGRPO_FLAG_ROUND_LIKE_TORCH) != 0;
  const int src = MIXGRPO_SRC_GIVEN;
  if (family == kFlow) {
    if (v_dtype == MIXGRPO_F32) return pick_src<kFlow, float, float, 1, false, false>(p, B, src, vec, st);
    if (rnd) return pick_src<kFlow, __nv_bfloat16, __nv_bfloat16, 1, true, false>(p, B, src, vec, st);
    return pick_src<kFlow, __nv_bfloat16, __nv_bfloat16, 1, false, false>(p, B, src, vec, st);
  }
  if (v_dtype == MIXGRPO_F32) return pick_src<kDance, float, float, 1, false, true>(p, B, src, vec, st);
  if (rnd) return pick_src<kDance, __nv_bfloat16, float, 1, true, true>(p, B, src, vec, st);
  return pick_src<kDance, __nv_bfloat16, float, 1, false, true>(p, B, src, vec, st);
}
