// Backward of the transition log-prob w.r.t. model_output (policy update, TR:585 loss.backward()
// flowing through SU:175-208 / SU:224-250).  One streaming pass: read v, x, x_next (10 B/elem with
// bf16 v), recompute the mean exactly as the forward did, write grad_v (2 B/elem).
//
// Autograd chain being reproduced (g = dL/dlogp[b], n = elements per sample, d = x_next - mean):
//   mean(dim)        -> g/n
//   (.)/(2 s^2)      -> (g/n)/(2 s^2)
//   -(d^2)           -> that * (2 d)          and d = x_next.detach() - mean gives +1 overall
//   flow : mean = x*c_x + (v*c_v)*dt          -> grad_v = ((gm)*dt)*c_v        (bf16 at every step
//   dance: mean = x + ds*v + drift(x0(v))*ds  -> two bf16 contributions summed  when v is bf16)
#include "common.cuh"

namespace mg {

struct BwdParams {
  const void* v;
  const float* x;
  const float* x_next;
  const float* grad_logp;
  void* grad_v;
  long long n, x_bs, in_bs;
  mixgrpo_step_coefs k;
  LossParams loss;          // fused policy path: grad_logp then holds the NEW log-probs and dL/dlogp is evaluated here
};

// EARLY (MIXGRPO_FLAG_PDL_EARLY_LOADS): the caller guarantees that v / x / x_next were not written by the kernel
// launched immediately before this one (true right after mixgrpo_policy_fwd, which only writes log-probs), so their
// loads are issued BEFORE griddepcontrol.wait and overlap the predecessor's tail; only dL/dlogp is read after it.
// FUSED (compile time): dL/dlogp comes from the clipped-ratio loss evaluated in place (policy path) instead of a given gradient
template <int FAM, class VT, bool RND, int VEC, bool EARLY, bool FUSED>
__global__ void __launch_bounds__(kThreads) logprob_bwd_kernel(const __grid_constant__ BwdParams p) {
  if constexpr (!EARLY) pdl_prologue();
  const int b = blockIdx.y;
  const long long idx = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;        // blockDim.x = g_bwd_threads (128 | 256)
  const bool active = idx < p.n;
  const float* c = p.k.c;
  float v[VEC], x[VEC], xn[VEC], t[VEC], mu[VEC], x0[VEC], g[VEC];
  if (active) {
    ld_stream(reinterpret_cast<const VT*>(p.v) + (long long)b * p.n + idx, v);
    ld_dep(p.x + (long long)b * p.x_bs + idx, x);
    ld_dep(p.x_next + (long long)b * p.in_bs + idx, xn);
  }
  if constexpr (EARLY) pdl_prologue();
  // (g/n)/(2 s^2): per-SAMPLE scalar, same two divisions autograd performs.  With the fused loss (TR:560-585 evaluated in place:
  // exp, four divisions, the clip logic — ~70 of a thread's ~250 instructions, and the window backward issues on 69 % of its
  // cycles) ONE thread per CTA evaluates it while the loads are in flight; a plain dL/dlogp is cheaper to read per thread than to
  // pass through a barrier (measured: 7.1 vs 7.6 us).
  float gs;
  if constexpr (FUSED) {
    __shared__ float s_gs;
    if (threadIdx.x == 0) {
      const float g_lp = loss_terms(ld_dep(p.grad_logp + b), ld_dep(p.loss.old_lp + b), ld_dep(p.loss.adv + b), p.loss, 1.f).grad;
      s_gs = __fdiv_rn(__fdiv_rn(g_lp, (float)p.n), p.k.two_var);
    }
    __syncthreads();
    gs = s_gs;
  } else {
    gs = __fdiv_rn(__fdiv_rn(ld_dep(p.grad_logp + b), (float)p.n), p.k.two_var);   // the forward launched just before may have written it: coherent
  }
  if (!active) return;

  if constexpr (FAM == 0) {   // flow, SU:186
#pragma unroll
    for (int i = 0; i < VEC; ++i) t[i] = __fmul_rn(v[i], c[2]);
    round_like_torch<RND>(t);
#pragma unroll
    for (int i = 0; i < VEC; ++i) t[i] = __fmul_rn(t[i], c[3]);
    round_like_torch<RND>(t);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      mu[i] = __fadd_rn(__fmul_rn(x[i], c[1]), t[i]);
      g[i] = __fmul_rn(gs, __fmul_rn(2.f, __fsub_rn(xn[i], mu[i])));
    }
    round_like_torch<RND>(g);                      // grad of the bf16 term
#pragma unroll
    for (int i = 0; i < VEC; ++i) g[i] = __fmul_rn(g[i], c[3]);
    round_like_torch<RND>(g);
#pragma unroll
    for (int i = 0; i < VEC; ++i) g[i] = __fmul_rn(g[i], c[2]);
  } else if constexpr (FAM == 2) {   // dpm order 1 (TR:169-180: dpm_state=None): mean = c5*x + c6*x0, x0 = x - sigma_s*v (SU:394, SU:426-444)
#pragma unroll
    for (int i = 0; i < VEC; ++i) t[i] = __fmul_rn(c[0], v[i]);
    round_like_torch<RND>(t);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      x0[i] = __fsub_rn(x[i], t[i]);
      mu[i] = __fadd_rn(__fmul_rn(c[5], x[i]), __fmul_rn(c[6], x0[i]));
      const float gm = __fmul_rn(gs, __fmul_rn(2.f, __fsub_rn(xn[i], mu[i])));
      g[i] = -__fmul_rn(gm, c[6]);                  // d/dx0 through the (sign-folded) coefficient, then x0 = x - t
    }
    round_like_torch<RND>(g);                       // grad of the bf16 product t = sigma_s*v
#pragma unroll
    for (int i = 0; i < VEC; ++i) g[i] = __fmul_rn(g[i], c[14]);   // c[14]: sigma_s as autograd's `grad * other` sees it
  } else {                    // dance with sde_solver=True, SU:224-234
#pragma unroll
    for (int i = 0; i < VEC; ++i) t[i] = __fmul_rn(c[0], v[i]);
    round_like_torch<RND>(t);
#pragma unroll
    for (int i = 0; i < VEC; ++i) x0[i] = __fsub_rn(x[i], t[i]);
#pragma unroll
    for (int i = 0; i < VEC; ++i) t[i] = __fmul_rn(c[1], v[i]);
    round_like_torch<RND>(t);
    float g1[VEC], g2[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float m = __fadd_rn(x[i], t[i]);
      float s = __fdiv_rn(-__fsub_rn(x[i], __fmul_rn(x0[i], c[2])), c[3]);
      m = __fadd_rn(m, __fmul_rn(__fmul_rn(s, c[4]), c[5]));
      const float gm = __fmul_rn(gs, __fmul_rn(2.f, __fsub_rn(xn[i], m)));
      g1[i] = gm;                                                     // via mean = x + ds*v
      // via drift: gm*ds -> *k -> /sigma^2 -> (two negations cancel) -> *(1-sigma) -> x0 = x - sigma*v
      g2[i] = -__fmul_rn(__fdiv_rn(__fmul_rn(__fmul_rn(gm, c[5]), c[4]), c[3]), c[2]);
    }
    round_like_torch<RND>(g1);
    round_like_torch<RND>(g2);
#pragma unroll
    for (int i = 0; i < VEC; ++i) { g1[i] = __fmul_rn(g1[i], c[1]); g2[i] = __fmul_rn(g2[i], c[0]); }
    round_like_torch<RND>(g1);
    round_like_torch<RND>(g2);
#pragma unroll
    for (int i = 0; i < VEC; ++i) g[i] = __fadd_rn(g1[i], g2[i]);
  }
  st_stream(reinterpret_cast<VT*>(p.grad_v) + (long long)b * p.n + idx, g);   // bf16 store rounds (RNE)
}

template <int FAM, class VT, bool RND, bool FUSED>
static int launch_bwd_f(const BwdParams& p, int64_t B, bool vec, bool early, cudaStream_t st) {
  const int thr = g_bwd_threads;
  if (vec) {
    dim3 grid((unsigned)((p.n + (long long)thr * kVec - 1) / ((long long)thr * kVec)), (unsigned)B);
    if (early) launch_pdl(logprob_bwd_kernel<FAM, VT, RND, kVec, true, FUSED>, grid, thr, 0, st, p);
    else launch_pdl(logprob_bwd_kernel<FAM, VT, RND, kVec, false, FUSED>, grid, thr, 0, st, p);
  } else {
    dim3 grid((unsigned)((p.n + thr - 1) / thr), (unsigned)B);
    launch_pdl(logprob_bwd_kernel<FAM, VT, RND, 1, false, FUSED>, grid, thr, 0, st, p);
  }
  return (int)cudaGetLastError();
}

template <int FAM, class VT, bool RND>
static int launch_bwd(const BwdParams& p, int64_t B, bool vec, bool early, cudaStream_t st) {
  return p.loss.old_lp ? launch_bwd_f<FAM, VT, RND, true>(p, B, vec, early, st) : launch_bwd_f<FAM, VT, RND, false>(p, B, vec, early, st);
}

template <int FAM>
static int bwd_family(const BwdParams& p, int v_dtype, int64_t B, bool vec, bool rnd, bool early, cudaStream_t st) {
  if (v_dtype == MIXGRPO_F32) return launch_bwd<FAM, float, false>(p, B, vec, early, st);
  if (rnd) return launch_bwd<FAM, __nv_bfloat16, true>(p, B, vec, early, st);
  return launch_bwd<FAM, __nv_bfloat16, false>(p, B, vec, early, st);
}

}  // namespace mg

using namespace mg;

static int logprob_bwd_impl(int family, const void* v, int v_dtype, const float* x, int64_t x_bs, const float* x_next,
                            int64_t in_bs, const float* grad_logp, void* grad_v, int64_t B, int64_t n,
                            const mixgrpo_step_coefs* coefs_host, const mixgrpo_loss_args* loss, unsigned flags, void* stream) {
  if (!v || !x || !x_next || !grad_logp || !grad_v || !coefs_host || B <= 0 || B > 65535 || n <= 0) return MIXGRPO_EINVAL;
  if (v_dtype != MIXGRPO_F32 && v_dtype != MIXGRPO_BF16) return MIXGRPO_EINVAL;
  if (family < 0 || family > 2 || (family == 2 && loss)) return MIXGRPO_EINVAL;
  BwdParams p;
  p.v = v; p.x = x; p.x_next = x_next; p.grad_logp = grad_logp; p.grad_v = grad_v;
  p.n = n; p.x_bs = x_bs; p.in_bs = in_bs; p.k = *coefs_host;
  p.loss = LossParams{nullptr, nullptr, nullptr, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 1};
  if (loss) p.loss = make_loss_params(loss->old_logp, loss->advantages, nullptr, loss->clip_range, loss->adv_clip_max, loss->kl_coeff, loss->denom);
  auto al = [](const void* q, size_t a) { return (reinterpret_cast<uintptr_t>(q) % a) == 0; };
  const size_t va = v_dtype == MIXGRPO_BF16 ? 16 : 32;
  const bool vec = (n % kVec == 0) && (x_bs % kVec == 0) && (in_bs % kVec == 0) && al(v, va) && al(grad_v, va) &&
                   al(x, 32) && al(x_next, 32);
  const bool rnd = (flags & MIXGRPO_FLAG_ROUND_LIKE_TORCH) != 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool early = (flags & MIXGRPO_FLAG_PDL_EARLY_LOADS) != 0;
  if (family == 2) return bwd_family<2>(p, v_dtype, B, vec, rnd, early, st);
  return family == 0 ? bwd_family<0>(p, v_dtype, B, vec, rnd, early, st) : bwd_family<1>(p, v_dtype, B, vec, rnd, early, st);
}

extern "C" __attribute__((visibility("default"))) int mixgrpo_logprob_bwd(int family, const void* v, int v_dtype, const float* x, int64_t x_bs,
                                   const float* x_next, int64_t in_bs, const float* grad_logp, void* grad_v,
                                   int64_t B, int64_t n, const mixgrpo_step_coefs* coefs_host, unsigned flags,
                                   void* stream) {
  return logprob_bwd_impl(family, v, v_dtype, x, x_bs, x_next, in_bs, grad_logp, grad_v, B, n, coefs_host, nullptr, flags, stream);
}

// Fused policy-update backward: dL/dlogp[b] is evaluated in place from (new_logp, old_logp, advantage)[b] — the
// loss never exists as a separate launch — then the same closed-form chain as mixgrpo_logprob_bwd.
extern "C" __attribute__((visibility("default"))) int mixgrpo_policy_bwd(int family, const void* v, int v_dtype, const float* x, int64_t x_bs,
                                  const float* x_next, int64_t in_bs, const float* new_logp, const mixgrpo_loss_args* loss,
                                  void* grad_v, int64_t B, int64_t n, const mixgrpo_step_coefs* coefs_host, unsigned flags,
                                  void* stream) {
  if (!loss || !loss->old_logp || !loss->advantages) return MIXGRPO_EINVAL;
  return logprob_bwd_impl(family, v, v_dtype, x, x_bs, x_next, in_bs, new_logp, grad_v, B, n, coefs_host, loss, flags, stream);
}
