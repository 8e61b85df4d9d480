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
// log-prob: per-thread sum of (x_next-mean)^2 -> fixed point -> ONE redux.sync.add per warp -> 64-bit warp parts through shared
// memory -> ONE packed atomicAdd per CTA (count + sum in a 64-bit word): integer from the thread up, so order- and
// shape-independent, hence bitwise reproducible; the last arriver writes logp[b] and re-zeroes the word (graph-replay safe),
// or — a rollout's deferred launches — nobody does and one finalize launch writes all log-probs.  Details at step_kernel below
// and in step_math.cuh.
#pragma once
#include <atomic>

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
  LpQuant lpq;              // fixed-point scale / per-thread cap / n * 2 s^2 of the log-prob reduction (host-evaluated, step_math.cuh)
  mixgrpo_step_coefs k;
  unsigned long long philox_seed, philox_offset;   // SRC_PHILOX
  const unsigned long long* philox_state;          // SRC_PHILOX, graph-safe: device {seed, base offset} (or nullptr)
  LossParams loss;          // fused policy path (SRC_GIVEN only): old log-probs / advantages / stats rows, or nullptrs
  int defer;                // MIXGRPO_FLAG_DEFER_LOGP: accumulate only (fire-and-forget), mixgrpo_logp_finalize writes the log-probs
  int early;                // programmatic dependent launch: 0 = wait before any load, 1 = v / noise first, 2 = every input first
  // optional second output: x_next (or x0) unpacked to (B,C,H,W) and de-normalised for the VAE (TR:102-115, TR:286-287)
  float* decode_out;
  int dC, dH, dW, d_from_x0, d_recip;
  float d_div, d_shift;
  // optional trajectory seed: x is the bf16 initial latent; its fp32 widening is also written here (all_latents[:, 0])
  float* seed_out;
  long long seed_bs;
};

// (B, S, 4C) packed scalars [i, i+8) of sample b -> (B, C, H, W): two channels x (2 x 2) patch = four 8-byte stores; a warp's
// 32 threads cover 4 tokens x 64 channels, i.e. for each (channel, row) one fully written 32-byte sector.
__device__ __forceinline__ void store_decoded(const StepParams& p, int b, long long i, const float (&r)[kVec]) {
  const int c4 = 4 * p.dC, Wp = p.dW >> 1;
  const long long tok = i / c4;
  const int c0 = (int)(i - tok * c4) >> 2;
  const int hp = (int)(tok / Wp), wp = (int)(tok - (long long)hp * Wp);
  const float inv = __fdiv_rn(1.f, p.d_div);
#pragma unroll
  for (int q = 0; q < 4; ++q) {                    // q = cc*2 + dh
    float a0 = r[2 * q], a1 = r[2 * q + 1];
    if (p.d_recip) { a0 = __fmul_rn(a0, inv); a1 = __fmul_rn(a1, inv); }       // torch's CUDA div-by-scalar: a * (1/b)
    else { a0 = __fdiv_rn(a0, p.d_div); a1 = __fdiv_rn(a1, p.d_div); }          // true division (torch CPU)
    a0 = __fadd_rn(a0, p.d_shift); a1 = __fadd_rn(a1, p.d_shift);
    float* dst = p.decode_out + (((long long)b * p.dC + c0 + (q >> 1)) * p.dH + 2 * hp + (q & 1)) * (long long)p.dW + 2 * wp;
    *reinterpret_cast<float2*>(dst) = make_float2(a0, a1);
  }
}

// ------------------------------------------------------------------ the streaming kernel
// Grid (ctas_per_sample, B); CTA = 256 threads; a CTA-tile is 2048 consecutive scalars of one sample and
// thread t owns scalars [8t, 8t+8) of it in every stream (one LDG.128 for bf16, one LDG.256 for fp32, one
// STG.256 per output).  Measured on B200 (tools/ubench.cu, profiles/r01_design_space.md): what matters for
// this 50-100 MB pass is bytes in flight — all of a thread's loads are issued before any math, the math
// is done pair-wise so the kernel stays at <= 40 registers (6 CTAs = 1536 threads per SM), and CTAs retire
// without waiting on anything (a persistent / software-pipelined variant and every fence+ticket
// reduction were 20-30 % slower).
//
// log-prob reduction — one atomic per CTA, deterministic, no fence (step_math.cuh):
//   acc[b] is a 64-bit word  [ sum : 40 bit fixed point Q8.32 | wide flags : 12 | arrivals : 12 ].
//   A thread converts its sum of d^2 to units of 2^-32 of mean(d^2 / 2 s^2); a warp adds its 32 integers with ONE
//   redux.sync; thread 0 adds the CTA's warp parts and issues ONE atomicAdd of (sum, wide?, 1 arrival).
//   Integer addition commutes, so the total is bit-identical whatever order CTAs arrive in and whatever the CTA shape; the
//   CTA whose returned count is the last one owns the complete sum in (old + mine), writes
//   logp[b] = -sum - log s - log sqrt(2 pi) and zeroes the word for the next launch.
//   Resolution 2^-32 per thread (1.2e-8 absolute on logp at 1024^2).  A warp with a thread beyond the fixed-point range sends
//   its fp32 sum to the record's 64-bit side accumulators instead, so the log-prob stays finite like the reference's up to
//   |d|/s ~ 1e8; only a non-finite sum (or one beyond that) gives NaN.

// DEP: the stream may have been written by the launch immediately before this one (latents, DPM history, stored x_next) ->
// coherent loads that stay behind griddepcontrol.wait (ld_dep, common.cuh); model output and noise come from several launches back
template <class T, bool VECTOR, bool DEP = false>
__device__ __forceinline__ void load_tile(const T* base, long long off, long long n, float (&r)[kVec]) {
  if constexpr (VECTOR) {
    if constexpr (DEP) ld_dep(base + off + threadIdx.x * kVec, r);          // caller guarantees off + 8t < n
    else ld_stream(base + off + threadIdx.x * kVec, r);
  } else {
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
      const long long i = off + j * kThreads + threadIdx.x;
      float one[1] = {0.f};
      if (i < n) { if constexpr (DEP) ld_dep(base + i, one); else ld_stream(base + i, one); }
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
// EXT (compile time): 0 none; 1 also emit the decode-ready tensor; 2 trajectory seed (x read as bf16, its fp32 widening stored
// too) — mixgrpo_step_ext; their own instantiations, so the other steps of a
// rollout that do not need it keep their register budget
// HALF (compile time): the deferred rollout launches' shape — 128-thread CTAs, one half-tile (1024 scalars) each, twice as many of
// them per SM, no finalization code: finer-grained CTA turnover and one fire-and-forget reduction into one of the sample's
// kDeferSubs sub-records.  Measured on the Euler-ODE step at (12,4096,64): -0.45 us of 6.15 (profiles/r02_design_space.md §4).
template <int FAM, class VT, class NT, int SRC, int ORDER, bool RND, bool SDE, bool VECTOR, int OUT, int EXT, bool HALF>
__global__ void __launch_bounds__(HALF ? kHalfThreads : kThreads,
                                  (((FAM == kDpm && ORDER >= 2) || OUT == 2 || !VECTOR || (FAM == kDance && SDE) || EXT == 1 || SRC == MIXGRPO_SRC_PHILOX) ? 4 : 6) * (HALF ? 2 : 1))
step_kernel(const __grid_constant__ StepParams p) {
  static_assert(!HALF || VECTOR, "the 128-thread shape is vector-only");
  constexpr int TILE = HALF ? kHalfTile : kTile;
  // Programmatic dependent launch: inputs the immediately preceding launch cannot have written are requested BEFORE
  // griddepcontrol.wait, so their DRAM latency overlaps that launch's drain (MIXGRPO_FLAG_PDL_EARLY_V / _EARLY_LOADS)
  // (the host only sets p.early when every CTA owns exactly one tile, so "first iteration" is the only iteration)
  if (p.early == 0) pdl_prologue();
  const int b = blockIdx.y;
  const long long n = p.n;
  const VT* vp = reinterpret_cast<const VT*>(p.v) + (long long)b * n;
  const float* xp = p.x + (long long)b * p.x_bs;
  float acc = 0.f;
  const LpQuant lpq = p.lpq;

  for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
    const long long off = (long long)tile * TILE;
    // vector path: n % 8 == 0, so a thread's 8 scalars are all inside or all outside the sample
    if (VECTOR && off + threadIdx.x * kVec >= n) continue;
    float v[kVec], x[kVec], a[kVec], m1[kVec], m2[kVec];
    load_tile<VT, VECTOR>(vp, off, n, v);
    if constexpr (SRC == MIXGRPO_SRC_NOISE) load_tile<NT, VECTOR>(reinterpret_cast<const NT*>(p.noise) + (long long)b * n, off, n, a);
    if (p.early == 1) pdl_prologue();
    if constexpr (EXT == 2) load_tile<__nv_bfloat16, VECTOR>(reinterpret_cast<const __nv_bfloat16*>(p.x) + (long long)b * p.x_bs, off, n, x);
    else load_tile<float, VECTOR, true>(xp, off, n, x);
    if constexpr (SRC == MIXGRPO_SRC_PHILOX) {     // draw the noise here: element e -> component e%4 of Philox(e/4)
      unsigned long long ph_seed = p.philox_seed, ph_off = p.philox_offset;
      if (p.philox_state) { ph_seed = ld_dep(p.philox_state); ph_off += ld_dep(p.philox_state + 1); }   // graph-safe state (moved by mixgrpo_philox_advance)
      if constexpr (VECTOR) {
        const unsigned long long e0 = (unsigned long long)b * n + off + threadIdx.x * kVec;
        float z0[4], z1[4];
        philox_normal4(e0 >> 2, ph_seed, ph_off, z0);
        philox_normal4((e0 >> 2) + 1, ph_seed, ph_off, z1);
#pragma unroll
        for (int j = 0; j < 4; ++j) { a[j] = z0[j]; a[4 + j] = z1[j]; }
      } else {
#pragma unroll
        for (int j = 0; j < kVec; ++j) {
          const unsigned long long e = (unsigned long long)b * n + off + j * kThreads + threadIdx.x;
          float z[4];
          philox_normal4(e >> 2, ph_seed, ph_off, z);
          a[j] = z[e & 3];
        }
      }
      if constexpr (sizeof(NT) == 2) round_like_torch<true>(a);        // the noise tensor itself is bf16 (SU:193)
    }
    if constexpr (SRC == MIXGRPO_SRC_GIVEN) load_tile<float, VECTOR, true>(p.x_in + (long long)b * p.in_bs, off, n, a);
    if constexpr (FAM == kDpm && ORDER >= 2) load_tile<float, VECTOR, true>(p.m1 + (long long)b * n, off, n, m1);
    if constexpr (FAM == kDpm && ORDER == 3) load_tile<float, VECTOR, true>(p.m2 + (long long)b * n, off, n, m2);
    if (p.early == 2) pdl_prologue();

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
    if constexpr (EXT == 2) store_tile<VECTOR>(p.seed_out + (long long)b * p.seed_bs, off, n, x);   // all_latents[:, 0] = float(z)
    if constexpr (EXT == 1) {                      // the rollout's last step: the VAE's input leaves from the same registers
      if constexpr (OUT >= 1) { if (p.d_from_x0) store_decoded(p, b, off + threadIdx.x * kVec, x0); else store_decoded(p, b, off + threadIdx.x * kVec, xn); }
      else store_decoded(p, b, off + threadIdx.x * kVec, xn);
    }
  }
  if (p.logp_out == nullptr && !p.defer) return;

  constexpr int WARPS = (HALF ? kHalfThreads : kThreads) / 32;
  __shared__ unsigned long long s_part[WARPS];
  // a deferred launch spreads a sample's arrivals over kDeferSubs records (side words of the same sub-record)
  unsigned long long* rec = p.defer ? p.acc + ((long long)kDeferSubs * b + (blockIdx.x & (kDeferSubs - 1))) * kWsStride : p.acc + kWsStride * b;
  const unsigned long long part = warp_part(acc, lpq, rec);                        // step_math.cuh: integer from the thread up
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x != 0) return;                    // nothing else waits on the atomic
  {
    const unsigned long long add = cta_word<WARPS>(s_part);
    if (HALF || p.defer) {
      // a true reduction (REDG.E.ADD.64): nothing comes back, so the CTA retires without an L2 round trip; the sums are
      // turned into log-probs by mixgrpo_logp_finalize after the rollout.  Spelled in PTX: nvcc keeps an ATOMG otherwise.
      asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" :: "l"(rec), "l"(add) : "memory");
      return;
    }
    const int ctas = gridDim.x;
    const unsigned long long old = atomicAdd(rec, add);
    if ((old & kArrivalMask) == (unsigned long long)(ctas - 1)) {
      const float q = packed_total(old + add, rec);
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
extern int g_max_ctas_per_sample;                       // bench knob (mixgrpo_set_tuning key 0), defined in step_flow.cu
extern int g_half_ctas;                                 // knob (key 6): deferred launches use the 128-thread shape — 0 never, 1 when the
                                                        // 256-thread grid exceeds one wave (default), 2 always
int sm_count();                                         // SMs of the current device (cached), step_flow.cu
extern std::atomic<long long> g_half_launches;          // launches issued in the 128-thread shape so far (key 8 reads it)

static inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

template <int FAM, class VT, class NT, int SRC, int ORDER, bool RND, bool SDE, bool VECTOR, int OUT, int EXT>
static int launch(StepParams& p, cudaStream_t st) {
  p.tiles = (int)((p.n + kTile - 1) / kTile);
  int ctas = p.tiles < g_max_ctas_per_sample ? p.tiles : g_max_ctas_per_sample;
  if constexpr (VECTOR) {
    // a rollout's deferred launches: one 128-thread CTA per half-tile (never looping, so the parts are the same half-tiles) —
    // when the 256-thread grid is more than one wave of 6 CTAs per SM.  Measured (B200): (12,4096,64) 1536 CTAs -0.2 us of 6.15,
    // (24,4096,64) -4 % on the whole step; a sub-wave grid, (24,1024,64) = 768 CTAs, is 0.25 us SLOWER with twice the CTAs to dispatch
    if (p.defer && ctas == p.tiles && (g_half_ctas == 2 || (g_half_ctas == 1 && (long long)ctas * p.B > 6LL * sm_count()))) {
      p.tiles = (int)((p.n + kHalfTile - 1) / kHalfTile);
      g_half_launches.fetch_add(1, std::memory_order_relaxed);
      dim3 grid((unsigned)p.tiles, (unsigned)p.B);
      launch_pdl(step_kernel<FAM, VT, NT, SRC, ORDER, RND, SDE, true, OUT, EXT, true>, grid, kHalfThreads, 0, st, p);
      return (int)cudaGetLastError();
    }
  }
  if (ctas != p.tiles || !VECTOR) p.early = 0;           // early loads assume one tile per CTA (no loop-carried state in the kernel)
  dim3 grid((unsigned)ctas, (unsigned)p.B);
  launch_pdl(step_kernel<FAM, VT, NT, SRC, ORDER, RND, SDE, VECTOR, OUT, EXT, false>, grid, kThreads, 0, st, p);
  return (int)cudaGetLastError();
}

template <int FAM, class VT, class NT, int SRC, int ORDER, bool RND, bool SDE, int OUT>
static int pick_vec(StepParams& p, bool vec_ok, cudaStream_t st) {
  if (p.seed_out) {                                        // set_ext() admitted it: vector path, bf16 x
    if constexpr (FAM == kFlow && SRC != MIXGRPO_SRC_GIVEN && OUT <= 1) return launch<FAM, VT, NT, SRC, ORDER, RND, SDE, true, OUT, 2>(p, st);
    else return MIXGRPO_EUNSUPPORTED;                      // the flow family's rollout steps only
  }
  if (p.decode_out) {                                      // set_ext() admitted it: vector path, a computed x_next (or x0)
    if constexpr (SRC != MIXGRPO_SRC_GIVEN && OUT <= 1) return launch<FAM, VT, NT, SRC, ORDER, RND, SDE, true, OUT, 1>(p, st);
    else return MIXGRPO_EUNSUPPORTED;                      // not together with the mean output
  }
  if (!vec_ok) return launch<FAM, VT, NT, SRC, ORDER, RND, SDE, false, OUT, 0>(p, st);
  return launch<FAM, VT, NT, SRC, ORDER, RND, SDE, true, OUT, 0>(p, st);
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

static inline bool check_common(const void* v, const float* x, int64_t B, int64_t n, int v_dtype, void* ws, int64_t ws_bytes,
                         float* logp, int* err, bool defer = false) {
  // grid.y carries the sample index; tile indices are 32-bit inside the kernel
  if (!v || !x || B <= 0 || B > 65535 || n <= 0 || (v_dtype != MIXGRPO_F32 && v_dtype != MIXGRPO_BF16) ||
      (n + kTile - 1) / kTile >= 2147483647LL) {
    *err = MIXGRPO_EINVAL;
    return false;
  }
  if ((logp || defer) && (!ws || ws_bytes < (defer ? mixgrpo_deferred_workspace_bytes(B, n) : mixgrpo_step_workspace_bytes(B, n)))) {
    *err = ws ? MIXGRPO_ENOSPACE : MIXGRPO_EINVAL;
    return false;
  }
  return true;
}

static inline void fill(StepParams& p, const void* v, const float* x, int64_t x_bs, const void* noise, const float* x_in,
                 int64_t in_bs, const float* m1, const float* m2, float* x_out, int64_t out_bs, float* x0_out,
                 float* mean_out, float* logp_out, void* ws, int64_t B, int64_t n, const mixgrpo_step_coefs* k) {
  p.v = v; p.x = x; p.noise = noise; p.x_in = x_in; p.m1 = m1; p.m2 = m2;
  p.x_out = x_out; p.x0_out = x0_out; p.mean_out = mean_out; p.logp_out = logp_out;
  p.acc = reinterpret_cast<unsigned long long*>(ws);
  p.n = n; p.x_bs = x_bs; p.in_bs = in_bs; p.out_bs = out_bs;
  p.B = (int)B; p.tiles = 0; p.k = *k;
  p.lpq = lp_quant(n, k->two_var);
  p.philox_seed = p.philox_offset = 0ull;
  p.philox_state = nullptr;
  p.loss = LossParams{nullptr, nullptr, nullptr, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 1};
  p.early = 0; p.defer = 0;
  p.decode_out = nullptr; p.dC = p.dH = p.dW = p.d_from_x0 = p.d_recip = 0; p.d_div = 1.f; p.d_shift = 0.f;
  p.seed_out = nullptr; p.seed_bs = 0;
}

static inline void set_philox(StepParams& p, const void* noise_host) {
  const mixgrpo_philox_args* ph = static_cast<const mixgrpo_philox_args*>(noise_host);
  p.philox_seed = ph->seed; p.philox_offset = ph->offset;
  p.philox_state = reinterpret_cast<const unsigned long long*>(ph->device_state);
}

static inline void set_early(StepParams& p, unsigned flags) {
  p.defer = (flags & MIXGRPO_FLAG_DEFER_LOGP) ? 1 : 0;
  p.early = (flags & MIXGRPO_FLAG_PDL_EARLY_LOADS) ? 2 : ((flags & MIXGRPO_FLAG_PDL_EARLY_V) ? 1 : 0);
}

// validates and installs the optional decode output; returns 0 or a MIXGRPO_E* code
static inline int set_ext(StepParams& p, const mixgrpo_step_ext* ext, int64_t n, bool vec, int src) {
  if (!ext) return 0;
  if (ext->x_f32_out) {                                    // trajectory seed: x is bf16 (16-B aligned is enough: vector_ok asked for 32)
    if (!ext->x_is_bf16 || ext->decode_out || src == MIXGRPO_SRC_GIVEN) return MIXGRPO_EINVAL;
    if (!vec || !aligned(ext->x_f32_out, 32) || (ext->x_f32_out_bs % kVec) != 0 || ext->x_f32_out_bs < n) return MIXGRPO_EUNSUPPORTED;
    p.seed_out = ext->x_f32_out; p.seed_bs = ext->x_f32_out_bs;
    return 0;
  }
  if (ext->x_is_bf16) return MIXGRPO_EINVAL;               // a bf16 x is only understood together with its fp32 copy-out
  if (!ext->decode_out) return 0;
  if (ext->C <= 0 || ext->H <= 0 || ext->W <= 0 || (ext->C % 2) || (ext->H % 2) || (ext->W % 2) || (int64_t)ext->C * ext->H * ext->W != n ||
      !(ext->divisor != 0.f)) return MIXGRPO_EINVAL;
  if (ext->from_x0 ? !p.x0_out : (src == MIXGRPO_SRC_GIVEN)) return MIXGRPO_EINVAL;   // the decoded tensor must be one this launch computes
  if (!vec || !aligned(ext->decode_out, 8)) return MIXGRPO_EUNSUPPORTED;
  p.decode_out = ext->decode_out; p.dC = ext->C; p.dH = ext->H; p.dW = ext->W;
  p.d_from_x0 = ext->from_x0 ? 1 : 0; p.d_div = ext->divisor; p.d_shift = ext->shift;
  p.d_recip = ext->reciprocal ? 1 : 0;
  return 0;
}

// 256-bit path needs 32-B aligned fp32 streams, 16-B aligned bf16 streams and n, strides % 8 == 0.
static inline bool vector_ok(const StepParams& p, int v_dtype, int noise_dtype, int64_t n) {
  bool ok = (n % kVec == 0) && (p.x_bs % kVec == 0) && (p.in_bs % kVec == 0) && (p.out_bs % kVec == 0);
  ok = ok && aligned(p.v, v_dtype == MIXGRPO_BF16 ? 16 : 32) && aligned(p.x, 32);
  ok = ok && aligned(p.noise, noise_dtype == MIXGRPO_BF16 ? 16 : 32) && aligned(p.x_in, 32);
  ok = ok && aligned(p.m1, 32) && aligned(p.m2, 32) && aligned(p.x_out, 32) && aligned(p.x0_out, 32) && aligned(p.mean_out, 32);
  return ok;
}

}  // namespace mg
