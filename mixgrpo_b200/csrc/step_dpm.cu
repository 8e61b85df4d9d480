// dpm_step family (SU:273-639: DPM-Solver / DPM-Solver++ order 1-3 in x0-prediction form) of the fused step kernel.
#include "step_kernel.cuh"

using namespace mg;

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
                                int src, unsigned flags, void* stream, const mixgrpo_step_ext* ext) {
  int err = 0;
  if (!coefs_host || !check_common(v, x, B, n, v_dtype, workspace, workspace_bytes, logp_out, &err, (flags & MIXGRPO_FLAG_DEFER_LOGP) != 0)) return err ? err : MIXGRPO_EINVAL;
  if (order < 1 || order > 3 || (order >= 2 && !m1) || (order == 3 && !m2)) return MIXGRPO_EINVAL;
  if ((src == MIXGRPO_SRC_NOISE || src == MIXGRPO_SRC_PHILOX) && !noise) return MIXGRPO_EINVAL;
  StepParams p;
  fill(p, v, x, x_bs, src == MIXGRPO_SRC_PHILOX ? nullptr : noise, nullptr, n, m1, m2, x_next_out, out_bs, x0_out, mean_out, logp_out, workspace, B, n, coefs_host);
  if (src == MIXGRPO_SRC_PHILOX) set_philox(p, noise);
  set_early(p, flags);
  const bool vec = vector_ok(p, v_dtype, MIXGRPO_F32, n);
  if ((err = set_ext(p, ext, n, vec, src)) != 0) return err;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool rnd = (flags & MIXGRPO_FLAG_ROUND_LIKE_TORCH) != 0;
  switch (order) {
    case 1: return dpm_dispatch<1>(p, v_dtype, B, src, vec, rnd, st);
    case 2: return dpm_dispatch<2>(p, v_dtype, B, src, vec, rnd, st);
    default: return dpm_dispatch<3>(p, v_dtype, B, src, vec, rnd, st);
  }
}
