// The SDE window's policy updates as TWO launches: one log-prob + loss forward and one backward for up to 8 independent
// (batch, window step) items (TR:536-585 loops over them one sample and one step at a time; TR =
// /root/reference/fastvideo/train_grpo_flux.py).  Given their model outputs the items do not depend on one another, so
// grid.y = n_items * B turns four 31 MB launches — each ~2 us of ramp + drain around ~5 us of streaming at group 12
// (profiles/r01_step_kernel_ncu.md: 40 % of a launch has no resident warp) — into one 126 MB launch, the regime where
// the same code reaches 0.9 of the HBM roofline.  Arithmetic, reduction order and the packed accumulator are those of
// mg::step_kernel<SRC_GIVEN> / mg::logprob_bwd_kernel, so every output is bit-identical to the per-item entry points.
#include "step_math.cuh"

namespace mg {

struct MultiItem {
  const void* v;
  const float* x;
  const float* x_in;
  float* logp;              // forward: out; backward: the new log-probs
  const float* old_lp;
  float* rows;              // [B,4] stats rows or nullptr
  void* grad_v;
  long long x_bs, in_bs;
  mixgrpo_step_coefs k;
  LpQuant lpq;              // host-evaluated (step_math.cuh)
};

struct MultiParams {
  MultiItem it[MIXGRPO_POLICY_MAX_ITEMS];
  unsigned long long* acc;  // n_items * B workspace records
  long long n;
  int B, tiles, n_items, early;
  LossParams loss;          // shared scalars + advantages; old_lp / rows come from the item
};

// dance trains only its sde_solver=True variant (TR:159-168)
template <int FAM, class VT, bool RND>
__global__ void __launch_bounds__(kThreads, FAM == kDance ? 4 : 6) policy_fwd_multi_kernel(const __grid_constant__ MultiParams p) {
  if (p.early == 0) pdl_prologue();
  const int item = blockIdx.y / p.B, b = blockIdx.y - item * p.B;
  const MultiItem& q = p.it[item];
  const long long n = p.n;
  const VT* vp = reinterpret_cast<const VT*>(q.v) + (long long)b * n;
  const float* xp = q.x + (long long)b * q.x_bs;
  const float* ap = q.x_in + (long long)b * q.in_bs;
  float acc = 0.f;
  const LpQuant lpq = q.lpq;
  for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
    const long long off = (long long)tile * kTile + threadIdx.x * kVec;
    if (off >= n) continue;
    float v[kVec], x[kVec], a[kVec];
    ld_stream(vp + off, v);
    ld_dep(xp + off, x);
    ld_dep(ap + off, a);
    if (p.early != 0) pdl_prologue();               // grid.x == tiles: the first iteration is the only one
#pragma unroll
    for (int j = 0; j < kVec; j += 2) {
      const float v2[2] = {v[j], v[j + 1]}, x2[2] = {x[j], x[j + 1]}, a2[2] = {a[j], a[j + 1]}, z2[2] = {0.f, 0.f};
      float xn2[2], x02[2], mu2[2], dd2[2];
      tile_math<FAM, MIXGRPO_SRC_GIVEN, 1, RND, true>(q.k, v2, x2, a2, z2, z2, xn2, x02, mu2, dd2);
      acc += dd2[0] + dd2[1];
    }
  }

  __shared__ unsigned long long s_part[kThreads / 32];
  unsigned long long* rec = p.acc + kWsStride * (long long)blockIdx.y;
  const unsigned long long part = warp_part(acc, lpq, rec);                          // step_math.cuh: integer from the thread up
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x != 0) return;
  {
    const unsigned long long add = cta_word<kThreads / 32>(s_part);
    const int ctas = gridDim.x;
    const unsigned long long old = atomicAdd(rec, add);
    if ((old & kArrivalMask) == (unsigned long long)(ctas - 1)) {
      const float s = packed_total(old + add, rec);
      const float lp = __fsub_rn(__fsub_rn(-s, q.k.log_scale), q.k.log_norm);        // SU:201-208
      q.logp[b] = lp;
      *rec = 0ull;
      if (q.rows) {                                                                    // TR:560-583, the reference's B == 1 evaluation
        const LossTerms lt = loss_terms(lp, q.old_lp[b], p.loss.adv[b], p.loss, 1.f);
        const float policy = __fdiv_rn(lt.policy_num, p.loss.denom);
        const float kl = __fdiv_rn(__fmul_rn(0.5f, lt.kl_num), p.loss.denom);
        float* row = q.rows + 4 * (long long)b;
        const float4 prev = p.loss.accumulate ? *reinterpret_cast<const float4*>(row) : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(row) = make_float4(prev.x + __fadd_rn(policy, __fmul_rn(p.loss.klc, kl)), prev.y + policy,
                                                      prev.z + kl, prev.w + lt.clip);
      }
    }
  }
}

// the chain of mg::logprob_bwd_kernel (csrc/bwd_kernels.cu), item-indexed; dL/dlogp is evaluated in place (TR:560-585)
template <int FAM, class VT, bool RND>
__global__ void __launch_bounds__(kThreads, 6) policy_bwd_multi_kernel(const __grid_constant__ MultiParams p) {
  if (p.early == 0) pdl_prologue();
  const int item = blockIdx.y / p.B, b = blockIdx.y - item * p.B;
  const MultiItem& q = p.it[item];
  const long long idx = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * kVec;       // blockDim.x = g_bwd_threads (128 | 256)
  const bool active = idx < p.n;
  const float* c = q.k.c;
  float v[kVec], x[kVec], xn[kVec], t[kVec], g[kVec];
  if (active) {
    ld_stream(reinterpret_cast<const VT*>(q.v) + (long long)b * p.n + idx, v);
    ld_dep(q.x + (long long)b * q.x_bs + idx, x);
    ld_dep(q.x_in + (long long)b * q.in_bs + idx, xn);
  }
  if (p.early != 0) pdl_prologue();
  // dL/dlogp and (g/n)/(2 s^2) are per-SAMPLE scalars: ONE thread per CTA evaluates them while the loads are in flight
  __shared__ float s_gs;
  if (threadIdx.x == 0) {
    const float g_lp = loss_terms(ld_dep(q.logp + b), ld_dep(q.old_lp + b), ld_dep(p.loss.adv + b), p.loss, 1.f).grad;   // the forward just wrote logp
    s_gs = __fdiv_rn(__fdiv_rn(g_lp, (float)p.n), q.k.two_var);
  }
  __syncthreads();
  if (!active) return;
  const float gs = s_gs;
  if constexpr (FAM == kFlow) {
#pragma unroll
    for (int i = 0; i < kVec; ++i) t[i] = __fmul_rn(v[i], c[2]);
    round_like_torch<RND>(t);
#pragma unroll
    for (int i = 0; i < kVec; ++i) t[i] = __fmul_rn(t[i], c[3]);
    round_like_torch<RND>(t);
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const float mu = __fadd_rn(__fmul_rn(x[i], c[1]), t[i]);
      g[i] = __fmul_rn(gs, __fmul_rn(2.f, __fsub_rn(xn[i], mu)));
    }
    round_like_torch<RND>(g);
#pragma unroll
    for (int i = 0; i < kVec; ++i) g[i] = __fmul_rn(g[i], c[3]);
    round_like_torch<RND>(g);
#pragma unroll
    for (int i = 0; i < kVec; ++i) g[i] = __fmul_rn(g[i], c[2]);
  } else {
    float x0[kVec], g1[kVec], g2[kVec];
#pragma unroll
    for (int i = 0; i < kVec; ++i) t[i] = __fmul_rn(c[0], v[i]);
    round_like_torch<RND>(t);
#pragma unroll
    for (int i = 0; i < kVec; ++i) x0[i] = __fsub_rn(x[i], t[i]);
#pragma unroll
    for (int i = 0; i < kVec; ++i) t[i] = __fmul_rn(c[1], v[i]);
    round_like_torch<RND>(t);
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      float m = __fadd_rn(x[i], t[i]);
      const float s = __fdiv_rn(-__fsub_rn(x[i], __fmul_rn(x0[i], c[2])), c[3]);
      m = __fadd_rn(m, __fmul_rn(__fmul_rn(s, c[4]), c[5]));
      const float gm = __fmul_rn(gs, __fmul_rn(2.f, __fsub_rn(xn[i], m)));
      g1[i] = gm;
      g2[i] = -__fmul_rn(__fdiv_rn(__fmul_rn(__fmul_rn(gm, c[5]), c[4]), c[3]), c[2]);
    }
    round_like_torch<RND>(g1);
    round_like_torch<RND>(g2);
#pragma unroll
    for (int i = 0; i < kVec; ++i) { g1[i] = __fmul_rn(g1[i], c[1]); g2[i] = __fmul_rn(g2[i], c[0]); }
    round_like_torch<RND>(g1);
    round_like_torch<RND>(g2);
#pragma unroll
    for (int i = 0; i < kVec; ++i) g[i] = __fadd_rn(g1[i], g2[i]);
  }
  st_stream(reinterpret_cast<VT*>(q.grad_v) + (long long)b * p.n + idx, g);
}

static inline bool al(const void* q, size_t a) { return (reinterpret_cast<uintptr_t>(q) % a) == 0; }

// returns 0, MIXGRPO_EINVAL or MIXGRPO_EUNSUPPORTED
static int fill_multi(MultiParams& p, int family, int v_dtype, const mixgrpo_policy_item* items, int n_items, const float* adv,
                      double clip_range, double adv_clip_max, double kl_coeff, double denom, int64_t B, int64_t n, bool backward) {
  if (!items || !adv || n_items <= 0 || n_items > MIXGRPO_POLICY_MAX_ITEMS || B <= 0 || n <= 0 || B * (int64_t)n_items > 65535) return MIXGRPO_EINVAL;
  if ((v_dtype != MIXGRPO_F32 && v_dtype != MIXGRPO_BF16) || (family != kFlow && family != kDance)) return MIXGRPO_EINVAL;
  const size_t va = v_dtype == MIXGRPO_BF16 ? 16 : 32;
  bool vec = (n % kVec == 0);
  for (int j = 0; j < n_items; ++j) {
    const mixgrpo_policy_item& s = items[j];
    if (!s.v || !s.x || !s.x_next || !s.logp || !s.old_logp || (backward && !s.grad_v)) return MIXGRPO_EINVAL;
    if (s.stats_rows && !al(s.stats_rows, 16)) return MIXGRPO_EINVAL;
    vec = vec && (s.x_bs % kVec == 0) && (s.in_bs % kVec == 0) && al(s.v, va) && al(s.x, 32) && al(s.x_next, 32) && (!backward || al(s.grad_v, va));
    MultiItem& d = p.it[j];
    d.v = s.v; d.x = s.x; d.x_in = s.x_next; d.logp = s.logp; d.old_lp = s.old_logp; d.rows = s.stats_rows; d.grad_v = s.grad_v;
    d.x_bs = s.x_bs; d.in_bs = s.in_bs; d.k = s.coefs; d.lpq = lp_quant(n, s.coefs.two_var);
  }
  if (!vec) return MIXGRPO_EUNSUPPORTED;
  p.n = n; p.B = (int)B; p.n_items = n_items; p.early = 0; p.acc = nullptr;
  p.tiles = (int)((n + kTile - 1) / kTile);
  p.loss = make_loss_params(nullptr, adv, nullptr, clip_range, adv_clip_max, kl_coeff, denom);
  return 0;
}

}  // namespace mg

using namespace mg;

extern "C" __attribute__((visibility("default"))) int mixgrpo_policy_fwd_multi(int family, int v_dtype, const mixgrpo_policy_item* items_host, int n_items,
                                                                               const float* advantages, double clip_range, double adv_clip_max,
                                                                               double kl_coeff, double denom, int accumulate, void* workspace,
                                                                               int64_t workspace_bytes, int64_t B, int64_t n, unsigned flags, void* stream) {
  MultiParams p;
  const int rc = fill_multi(p, family, v_dtype, items_host, n_items, advantages, clip_range, adv_clip_max, kl_coeff, denom, B, n, false);
  if (rc != 0) return rc;
  if (!workspace) return MIXGRPO_EINVAL;
  if (workspace_bytes < mixgrpo_step_workspace_bytes(B * n_items, n)) return MIXGRPO_ENOSPACE;
  if (p.tiles > kMaxCtasPerSample) return MIXGRPO_EUNSUPPORTED;
  p.acc = reinterpret_cast<unsigned long long*>(workspace);
  p.loss.accumulate = accumulate ? 1 : 0;
  p.early = (flags & MIXGRPO_FLAG_PDL_EARLY_LOADS) ? 2 : 0;
  const bool rnd = (flags & MIXGRPO_FLAG_ROUND_LIKE_TORCH) != 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)p.tiles, (unsigned)(B * n_items));
  if (family == kFlow) {
    if (v_dtype == MIXGRPO_F32) launch_pdl(policy_fwd_multi_kernel<kFlow, float, false>, grid, kThreads, 0, st, p);
    else if (rnd) launch_pdl(policy_fwd_multi_kernel<kFlow, __nv_bfloat16, true>, grid, kThreads, 0, st, p);
    else launch_pdl(policy_fwd_multi_kernel<kFlow, __nv_bfloat16, false>, grid, kThreads, 0, st, p);
  } else {
    if (v_dtype == MIXGRPO_F32) launch_pdl(policy_fwd_multi_kernel<kDance, float, false>, grid, kThreads, 0, st, p);
    else if (rnd) launch_pdl(policy_fwd_multi_kernel<kDance, __nv_bfloat16, true>, grid, kThreads, 0, st, p);
    else launch_pdl(policy_fwd_multi_kernel<kDance, __nv_bfloat16, false>, grid, kThreads, 0, st, p);
  }
  return (int)cudaGetLastError();
}

extern "C" __attribute__((visibility("default"))) int mixgrpo_policy_bwd_multi(int family, int v_dtype, const mixgrpo_policy_item* items_host, int n_items,
                                                                               const float* advantages, double clip_range, double adv_clip_max,
                                                                               double kl_coeff, double denom, int64_t B, int64_t n, unsigned flags,
                                                                               void* stream) {
  MultiParams p;
  const int rc = fill_multi(p, family, v_dtype, items_host, n_items, advantages, clip_range, adv_clip_max, kl_coeff, denom, B, n, true);
  if (rc != 0) return rc;
  p.early = (flags & MIXGRPO_FLAG_PDL_EARLY_LOADS) ? 2 : 0;
  const bool rnd = (flags & MIXGRPO_FLAG_ROUND_LIKE_TORCH) != 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int thr = g_bwd_threads;
  dim3 grid((unsigned)((n + (long long)thr * kVec - 1) / ((long long)thr * kVec)), (unsigned)(B * n_items));
  if (family == kFlow) {
    if (v_dtype == MIXGRPO_F32) launch_pdl(policy_bwd_multi_kernel<kFlow, float, false>, grid, thr, 0, st, p);
    else if (rnd) launch_pdl(policy_bwd_multi_kernel<kFlow, __nv_bfloat16, true>, grid, thr, 0, st, p);
    else launch_pdl(policy_bwd_multi_kernel<kFlow, __nv_bfloat16, false>, grid, thr, 0, st, p);
  } else {
    if (v_dtype == MIXGRPO_F32) launch_pdl(policy_bwd_multi_kernel<kDance, float, false>, grid, thr, 0, st, p);
    else if (rnd) launch_pdl(policy_bwd_multi_kernel<kDance, __nv_bfloat16, true>, grid, thr, 0, st, p);
    else launch_pdl(policy_bwd_multi_kernel<kDance, __nv_bfloat16, false>, grid, thr, 0, st, p);
  }
  return (int)cudaGetLastError();
}
