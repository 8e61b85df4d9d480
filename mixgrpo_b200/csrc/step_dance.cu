// dance_grpo_step family (SU:212-253) of the fused step kernel.
#include "step_kernel.cuh"

using namespace mg;

namespace mg {
// the dance half of mixgrpo_policy_fwd (step_flow.cu): the stored transition's log-prob with sde_solver=True (TR:159-168)
int policy_fwd_dance(StepParams& p, int v_dtype, int64_t B, bool vec, bool rnd, cudaStream_t st) {
  const int src = MIXGRPO_SRC_GIVEN;
  if (v_dtype == MIXGRPO_F32) return pick_src<kDance, float, float, 1, false, true>(p, B, src, vec, st);
  if (rnd) return pick_src<kDance, __nv_bfloat16, float, 1, true, true>(p, B, src, vec, st);
  return pick_src<kDance, __nv_bfloat16, float, 1, false, true>(p, B, src, vec, st);
}
}  // namespace mg

extern "C" __attribute__((visibility("default"))) int mixgrpo_dance_step(const void* v, int v_dtype, const float* x, int64_t x_bs, const float* noise,
                                  const float* x_next_in, int64_t in_bs, float* x_next_out, int64_t out_bs,
                                  float* x0_out, float* mean_out, float* logp_out, void* workspace,
                                  int64_t workspace_bytes, int64_t B, int64_t n, const mixgrpo_step_coefs* coefs_host,
                                  int src, int sde_solver, unsigned flags, void* stream, const mixgrpo_step_ext* ext) {
  int err = 0;
  if (!coefs_host || !check_common(v, x, B, n, v_dtype, workspace, workspace_bytes, logp_out, &err, (flags & MIXGRPO_FLAG_DEFER_LOGP) != 0)) return err ? err : MIXGRPO_EINVAL;
  if (((src == MIXGRPO_SRC_NOISE || src == MIXGRPO_SRC_PHILOX) && !noise) || (src == MIXGRPO_SRC_GIVEN && !x_next_in)) return MIXGRPO_EINVAL;
  StepParams p;
  fill(p, v, x, x_bs, src == MIXGRPO_SRC_PHILOX ? nullptr : noise, x_next_in, in_bs, nullptr, nullptr, x_next_out, out_bs, x0_out, mean_out, logp_out, workspace, B, n, coefs_host);
  if (src == MIXGRPO_SRC_PHILOX) set_philox(p, noise);
  set_early(p, flags);
  const bool vec = vector_ok(p, v_dtype, MIXGRPO_F32, n);
  if ((err = set_ext(p, ext, n, vec, src)) != 0) return err;
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
