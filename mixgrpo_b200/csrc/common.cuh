// Shared device helpers for the mixgrpo_b200 streaming kernels (sm_100a only).
//
// The hot path is HBM-bound elementwise + reduction work (≈1 flop/byte), so everything here is
// about moving bytes: 256-bit LDG/STG (sm_100 adds .v8.f32 / LDG.E.256) so that one thread owns 8
// consecutive latent scalars in every stream — 32 B of fp32, 16 B of bf16 — and every warp-level
// request covers whole 128-B lines with all sectors used; streaming (no-L1-allocate) hints; a
// deterministic warp-shuffle → shared → packed fixed-point atomic reduction for the per-sample log-prob;
// programmatic dependent launch so chained kernels overlap their tails; counter-based in-kernel noise.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/mixgrpo_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "mixgrpo_b200 kernels target sm_100a (B200) only"
#endif

namespace mg {

constexpr int kThreads = 256;   // threads per CTA for the streaming kernels
constexpr int kVec = 8;         // latent scalars owned by one thread per tile

// ---------------------------------------------------------------- 256/128-bit streaming access
__device__ __forceinline__ void ld_stream(const float* p, float (&r)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void ld_stream(const float* p, float (&r)[1]) { r[0] = __ldg(p); }

// 8 bf16 (16 B) -> 8 floats.  Widening is exact: bf16 is the top half of an fp32.
__device__ __forceinline__ void ld_stream(const __nv_bfloat16* p, float (&r)[8]) {
  uint32_t a, b, c, d;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
               : "l"(p));
  r[0] = __uint_as_float(a << 16); r[1] = __uint_as_float(a & 0xffff0000u);
  r[2] = __uint_as_float(b << 16); r[3] = __uint_as_float(b & 0xffff0000u);
  r[4] = __uint_as_float(c << 16); r[5] = __uint_as_float(c & 0xffff0000u);
  r[6] = __uint_as_float(d << 16); r[7] = __uint_as_float(d & 0xffff0000u);
}
__device__ __forceinline__ void ld_stream(const __nv_bfloat16* p, float (&r)[1]) {
  r[0] = __bfloat162float(*p);
}

// Data the launch immediately BEFORE this one may have written — the latents a rollout step reads, a DPM history, the stored
// x_next, the new log-probs a backward reads: COHERENT loads that are also compiler barriers.  ld.global.nc promises that the
// data is read-only for the kernel's whole lifetime, which under programmatic dependent launch begins before the previous
// kernel has ended; ptxas does move such loads across griddepcontrol.wait (it had put the wait of an early-loads launch right
// in front of the kernel's first STORE), so .nc is only for streams produced several launches back (model output, noise).
__device__ __forceinline__ void ld_dep(const float* p, float (&r)[8]) {
  asm volatile("ld.global.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
               : "l"(p) : "memory");
}
__device__ __forceinline__ void ld_dep(const float* p, float (&r)[1]) {
  asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(r[0]) : "l"(p) : "memory");
}
__device__ __forceinline__ float ld_dep(const float* p) {
  float r;
  asm volatile("ld.global.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ unsigned long long ld_dep(const unsigned long long* p) {
  unsigned long long r;
  asm volatile("ld.global.u64 %0, [%1];" : "=l"(r) : "l"(p) : "memory");
  return r;
}

__device__ __forceinline__ void st_stream(float* p, const float (&r)[8]) {
  asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(p), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7])
               : "memory");
}
__device__ __forceinline__ void st_stream(float* p, const float (&r)[1]) { *p = r[0]; }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);   // .x = lo (low half-word), .y = hi
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void st_stream(__nv_bfloat16* p, const float (&r)[8]) {
  uint32_t a = pack_bf16x2(r[0], r[1]), b = pack_bf16x2(r[2], r[3]);
  uint32_t c = pack_bf16x2(r[4], r[5]), d = pack_bf16x2(r[6], r[7]);
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_stream(__nv_bfloat16* p, const float (&r)[1]) {
  *p = __float2bfloat16_rn(r[0]);
}

// ---------------------------------------------------------------- torch-promotion rounding
// R<RND>(x): when torch's type promotion would have produced a bf16 intermediate, round x to bf16
// (RNE, like c10::BFloat16) and widen back.  Pairs share one F2FP.BF16.F32.PACK_AB.
template <bool RND, int N>
__device__ __forceinline__ void round_like_torch(float (&r)[N]) {
  if constexpr (RND) {
    if constexpr (N % 2 == 0) {
#pragma unroll
      for (int k = 0; k < N; k += 2) {
        uint32_t u = pack_bf16x2(r[k], r[k + 1]);
        r[k] = __uint_as_float(u << 16);
        r[k + 1] = __uint_as_float(u & 0xffff0000u);
      }
    } else {
#pragma unroll
      for (int k = 0; k < N; ++k) r[k] = __bfloat162float(__float2bfloat16_rn(r[k]));
    }
  }
}

// ---------------------------------------------------------------- in-kernel normal noise (MIXGRPO_SRC_PHILOX)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// four N(0,1) samples for elements 4q .. 4q+3 (q = element index / 4)
__device__ __forceinline__ void philox_normal4(unsigned long long q, unsigned long long seed, unsigned long long offset, float (&z)[4]) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x6d697867u));
  // 24-bit uniforms in (0,1): never 0, so the log is finite
  const float u0 = ((float)(r.x >> 8) + 0.5f) * 5.9604644775390625e-08f, u1 = ((float)(r.y >> 8) + 0.5f) * 5.9604644775390625e-08f;
  const float u2 = ((float)(r.z >> 8) + 0.5f) * 5.9604644775390625e-08f, u3 = ((float)(r.w >> 8) + 0.5f) * 5.9604644775390625e-08f;
  const float ra = sqrtf(-2.f * __logf(u0)), rb = sqrtf(-2.f * __logf(u2));
  float sa, ca, sb, cb;
  __sincosf(6.283185307179586f * u1, &sa, &ca);
  __sincosf(6.283185307179586f * u3, &sb, &cb);
  z[0] = ra * ca; z[1] = ra * sa; z[2] = rb * cb; z[3] = rb * sb;
}

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// Every kernel of this library is launched with cudaLaunchAttributeProgrammaticStreamSerialization and starts with
//   griddepcontrol.wait              -> the previous grid on the stream has completed and its writes are visible
//   griddepcontrol.launch_dependents -> the NEXT grid may start filling SMs as this grid's CTAs retire
// so in a chain of our launches (25 sampler steps, policy forward -> backward) the next kernel's launch latency and
// ramp-up overlap this kernel's tail instead of following it.  The trigger comes AFTER the wait on purpose: a dependent
// can then only start once everything two launches back has completed, which is what makes "early loads" (inputs that
// the immediately preceding kernel does not write, MIXGRPO_FLAG_PDL_EARLY_LOADS) safe.  After a non-PDL kernel (torch,
// cuBLAS) the attribute is inert and ordering is the ordinary stream order.  Measured: -0.4 us per chained launch,
// -1.1 us with early loads (profiles/r01_design_space.md).
extern int g_use_pdl;   // mixgrpo_set_tuning key 1 (bench A/B knob), defined in step_flow.cu
extern int g_bwd_threads;   // key 7: threads per CTA of the log-prob backward kernels (128 | 256), defined in step_flow.cu
}  // namespace mg
int mixgrpo_peer_set_timeout_ms(int ms);   // mixgrpo_set_tuning key 2, defined in peer_kernels.cu
int mixgrpo_policy_set_tuning(int key, int value);   // keys 3-5, defined in policy_kernels.cu
namespace mg {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() { pdl_wait(); pdl_launch_dependents(); }

template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = g_use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------- deterministic reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// ---------------------------------------------------------------- clipped-ratio GRPO loss, one sample
// TR:560-583 for ONE log-prob (TR = fastvideo/train_grpo_flux.py).  Shared by the batch loss kernel and by the
// fused policy path, where the forward log-prob kernel's finalizer and every CTA of the backward kernel
// evaluate it in place (no loss launch, no host sync).
struct LossParams {
  const float* old_lp;     // [B]   (fused path only)
  const float* adv;        // [B]
  float* rows;             // [B,4] running (loss, policy, kl, clip_frac) per sample, or nullptr
  float clip, lo, hi, amax, klc, denom;
  int accumulate;          // rows[b] += terms (1) or rows[b] = terms (0)
};

struct LossTerms {
  float policy_num;        // max(-A r, -A clamp(r))           -> mean / denom = policy loss
  float kl_num;            // (new - old)^2                    -> 0.5 * mean / denom = kl loss
  float clip;              // |r - 1| > clip_range ? 1 : 0
  float grad;              // d loss / d new_logp  (includes the 1/B of the batch mean via inv_b)
};

__host__ inline LossParams make_loss_params(const float* old_lp, const float* adv, float* rows, double clip_range,
                                            double adv_clip_max, double kl_coeff, double denom) {
  LossParams q;
  q.old_lp = old_lp; q.adv = adv; q.rows = rows;
  // python scalars are narrowed to fp32 exactly where torch narrows them; 1 -/+ clip is formed in double (TR:571-572)
  q.clip = (float)clip_range; q.lo = (float)(1.0 - clip_range); q.hi = (float)(1.0 + clip_range);
  q.amax = (float)adv_clip_max; q.klc = (float)kl_coeff; q.denom = (float)denom;
  q.accumulate = 1;
  return q;
}

__device__ __forceinline__ LossTerms loss_terms(float new_lp, float old_lp, float adv, const LossParams& q, float inv_b) {
  LossTerms t;
  const float a = fminf(fmaxf(adv, -q.amax), q.amax);              // TR:560-564
  const float lr = __fsub_rn(new_lp, old_lp);
  const float r = expf(lr);                                        // TR:566
  const float rc = fminf(fmaxf(r, q.lo), q.hi);
  const float un = __fmul_rn(-a, r), cl = __fmul_rn(-a, rc);        // TR:568-573
  t.policy_num = fmaxf(un, cl);
  t.clip = (fabsf(__fsub_rn(r, 1.f)) > q.clip) ? 1.f : 0.f;        // TR:574
  t.kl_num = __fmul_rn(lr, lr);                                    // TR:580
  // torch.maximum splits the gradient 1/2 + 1/2 on ties; clamp passes gradient on [lo, hi] inclusive
  const float inside = (r >= q.lo && r <= q.hi) ? 1.f : 0.f;
  float dpl_dr;
  if (un > cl) dpl_dr = -a;
  else if (un < cl) dpl_dr = -a * inside;
  else dpl_dr = 0.5f * (-a) + 0.5f * (-a) * inside;
  t.grad = dpl_dr * r * inv_b / q.denom + q.klc * lr * inv_b / q.denom;
  return t;
}

}  // namespace mg
