/*
 * mixgrpo_b200 — C ABI of the B200-native MixGRPO rollout / policy-update hot path.
 *
 * The reference (zqqqqz2000/MixGRPO) has no FFI: its boundary for this path is a set of Python
 * functions operating on torch tensors (fastvideo/utils/sampling_utils.py, "SU") and inline
 * arithmetic in fastvideo/train_grpo_flux.py ("TR").  This header is the plain-C boundary a
 * binding (ctypes / cffi / pybind / TORCH_LIBRARY shim) attaches to; every entry point names
 * the reference code it replaces.  No torch types appear here: device pointers, sizes, a
 * coefficient block and a cudaStream_t (as void*) only.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host;
 *   - tensors are (B, n) row-major: B samples of n = S*64 packed-latent scalars; a "batch
 *     stride" (in elements) lets a view such as all_latents[:, i] (SU:153) be read/written in
 *     place; pass n for a contiguous tensor;
 *   - every function returns 0 on success, a cudaError_t (>0) when the launch failed, or a
 *     negative MIXGRPO_E* code for argument errors.  Nothing throws, nothing synchronises;
 *   - kernels are launched on `stream`; all are CUDA-graph capturable;
 *   - there is NO CPU fallback: without a CUDA device the calls fail.
 */
#ifndef MIXGRPO_B200_H
#define MIXGRPO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIXGRPO_ABI_VERSION 6

/* element type of model_output / noise / grad_model_output */
enum { MIXGRPO_F32 = 0, MIXGRPO_BF16 = 1 };

/* where x_next comes from */
enum {
  MIXGRPO_SRC_NOISE = 0,        /* rollout, SDE:  x_next = mean + scale*noise   (SU:188-195, SU:238, SU:434) */
  MIXGRPO_SRC_GIVEN = 1,        /* policy update: x_next supplied (prev_sample=, TR:149-157)               */
  MIXGRPO_SRC_DETERMINISTIC = 2,/* rollout, ODE:  Euler / DPM closed form       (SU:198-199, SU:240, SU:436) */
  MIXGRPO_SRC_PHILOX = 3        /* rollout, SDE with the noise drawn IN the kernel (replaces the randn_tensor /
                                   randn_like launch and its 2-4 B/elem read, SU:189-194, SU:238, SU:319-321):
                                   `noise` is then a HOST pointer to a mixgrpo_philox_args                     */
};

/* Counter-based noise for MIXGRPO_SRC_PHILOX: element e = b*n + i of the (B, n) tensor gets component (e mod 4) of
 * Philox4x32-10(counter = {e/4 (64 bit), offset (64 bit)}, key = seed ^ tag) through Box-Muller, i.e. a pure function
 * of (seed, offset, e) — independent of the launch configuration, reproducible, and reproducible on the host
 * (oracle/philox_oracle.py).  Rounded to bf16 when the family's noise dtype is bf16 (flow with bf16 model output). */
typedef struct mixgrpo_philox_args {
  uint64_t seed;
  uint64_t offset;
  /* CUDA-graph-safe state (nullable): DEVICE pointer to {seed, base offset}.  When set, the kernel reads both from
   * device memory at run time — seed = device_state[0], offset = device_state[1] + `offset` (the host-known position
   * inside the graph) — so replays of a captured launch draw fresh noise once mixgrpo_philox_advance has moved the base
   * (torch's graph-safe generator does the same with its seed / offset-extragraph pointers).  `seed` is then ignored. */
  const uint64_t* device_state;
} mixgrpo_philox_args;

/* flags */
#define MIXGRPO_FLAG_ROUND_LIKE_TORCH 1u /* reproduce torch's bf16 type-promotion roundings */
#define MIXGRPO_FLAG_PDL_EARLY_LOADS  2u /* NONE of the streamed inputs (v / x / x_next / noise / history) was written by
                                            the kernel launched immediately before on this stream (e.g. a backward right
                                            after mixgrpo_policy_fwd, which only writes [B] log-probs): their loads are
                                            issued before griddepcontrol.wait and overlap that kernel's tail */
#define MIXGRPO_FLAG_PDL_EARLY_V      4u /* step kernels: model_output and noise were not written by the immediately
                                            preceding launch (they come from the DiT / randn, several launches back) but
                                            the latents may have been (step i reads what step i-1 wrote): only v / noise
                                            are loaded before the wait */

#define MIXGRPO_FLAG_DEFER_LOGP       8u /* step kernels: accumulate the launch's per-sample log-prob sums into `workspace` — then the
                                            CALLER's own records for THIS launch, mixgrpo_deferred_workspace_bytes(B, n) bytes, zeroed —
                                            with a fire-and-forget reduction and do not finalize; logp_out may be NULL.  A rollout's
                                            log-probs are only read after the rollout (SU:153-155), so its 25 step launches skip the
                                            returning atomic that otherwise keeps every CTA resident for an L2 round trip (0.7 us of a
                                            7 us launch at (12,4096,64)), and — nobody having to see the last arrival — run as
                                            128-thread CTAs of one half-tile each whose reductions are spread over 8 sub-records per
                                            sample (another 0.45 us); ONE mixgrpo_logp_finalize launch then turns all records into
                                            log-probs.  Same per-half-tile integer sums, hence the same bits as the immediate path. */

/* error codes */
#define MIXGRPO_EINVAL   (-1)  /* bad argument (null pointer, bad enum, B<=0, n<=0) */
#define MIXGRPO_EALIGN   (-2)  /* reserved */
#define MIXGRPO_ENOSPACE (-3)  /* workspace too small */
#define MIXGRPO_EUNSUPPORTED (-4) /* this entry point does not cover the shape (see mixgrpo_policy_step); nothing was launched */

/* Per-step scalar block.  All values are fp32 numbers computed on the HOST with the reference's
 * operator order (mixgrpo_b200/coefs.py); where torch would cast a 0-dim scalar to bf16 before a
 * multiply, the host stores the already-rounded value.  Layout per family:
 *
 *  common (all families)
 *    two_var   2*scale^2            divisor of the squared residual          (SU:202, SU:245, SU:377)
 *    log_scale log(scale)           0 for the dance family (SU:247 quirk)
 *    log_norm  log(sqrt(2*pi))      0 for the dance family
 *
 *  flow  (flow_grpo_step, SU:157-210)         c[0..5]
 *    0 sigma   (x0 = x - sigma*v)               1 c_x  = 1 + std^2/(2 sigma)*dt
 *    2 c_v   = 1 + std^2 (1-sigma)/(2 sigma)    3 dt   (factor of the mean's v term)
 *    4 scale = std*sqrt(-dt) (noise factor)     5 dt   (Euler ODE factor, SU:199)
 *
 *  dance (dance_grpo_step, SU:212-253)        c[0..6]
 *    0 sigma (x0)   1 dsigma (mean = x + dsigma*v)   2 1-sigma   3 sigma^2
 *    4 -0.5*eta^2   5 dsigma (drift factor, fp32)    6 std = eta*sqrt(-dsigma)
 *
 *  dpm   (dpm_step and its order-1/2/3 updates, SU:273-639)   c[0..13]
 *    0 sigma_s (x0 = x - sigma_s*v, SU:394)
 *    1 1/r0     2 1/r1     3 r0/(r0+r1)    4 1/(r0+r1)          (finite differences, SU:490, SU:608-610)
 *    5..8   mean   = c5*x + c6*D0 + c7*D1 + c8*D2                (signs folded into the coefficients)
 *    9..12  x_ode  = c9*x + c10*D0 + c11*D1 + c12*D2
 *    13 scale = sigma_t*sqrt(1-exp(-2h))  (noise factor)
 *    14 sigma_s as the log-prob backward multiplies by it (mixgrpo_logprob_bwd family 2)
 */
typedef struct mixgrpo_step_coefs {
  float two_var;
  float log_scale;
  float log_norm;
  float c[16];
} mixgrpo_step_coefs;

/* Bytes of zero-initialised device workspace the step / policy kernels need for (B, n): one 32-byte record per
 * sample — a 64-bit packed accumulator ([fixed-point sum | wide-share count | arrival count], csrc/step_kernel.cuh) for
 * the deterministic log-prob reduction, a 32-bit epoch and a status word (csrc/policy_kernels.cu), and a 64-bit side
 * accumulator for shares too large for the packed field (csrc/step_math.cuh).  The layout does
 * not depend on B and kernels leave the accumulators zeroed again, so one allocation can be reused by successive
 * launches of any batch size on the same stream (never by launches that may run concurrently). */
int64_t mixgrpo_step_workspace_bytes(int64_t B, int64_t n);

/* Bytes of zero-initialised device workspace ONE launch with MIXGRPO_FLAG_DEFER_LOGP needs for (B, n): eight such 32-byte
 * records per sample (256 B), over which the launch spreads the sample's arrivals so that its thousands of fire-and-forget
 * reductions do not serialise on 12 addresses in the L2.  mixgrpo_logp_finalize adds them up (integers: exact) and leaves them
 * zeroed.  A rollout of N step launches holds N such blocks, launch_stride_bytes apart. */
int64_t mixgrpo_deferred_workspace_bytes(int64_t B, int64_t n);

/* ABI / build introspection. */
int mixgrpo_abi_version(void);
const char* mixgrpo_build_info(void);          /* e.g. "sm_100a nvcc 12.9 ..." (static string) */
int mixgrpo_set_tuning(int key, int value);    /* knobs (0: max CTAs/sample, 1: PDL on/off, 2: peer-wait timeout in ms, 0 = forever,
                                                  3: single-pass policy kernel CTAs/SM, 4: its cooperative launch on/off, 5: its wait
                                                  timeout in ms, 6: deferred step launches as 128-thread CTAs — 0 never, 1 when the grid exceeds one
                                                  wave (default), 2 always, 7: CTA size of the log-prob backward kernels (128 | 256), 8: read-only count of
                                                  step launches issued in the 128-thread shape); returns the
                                                  previous value or <0 */
const char* mixgrpo_error_string(int code);    /* text for a return code (cudaGetErrorString for >0) */

/* ---- fused sampler step + Gaussian transition log-prob ---------------------------------------
 * One pass over the latents: reads v (and noise or the stored x_next), writes x_next / x0
 * (/ mean) and the per-sample log-prob.  Output pointers may be NULL to skip that stream.
 *
 * mixgrpo_flow_step   replaces flow_grpo_step            SU:157-210 (rollout SU:85-93; train TR:149-157)
 * mixgrpo_dance_step  replaces dance_grpo_step           SU:212-253 (rollout SU:95-98; train TR:159-168)
 * mixgrpo_dpm_step    replaces dpm_step + convert_model_output + order-1/2/3 updates
 *                              SU:273-639 (rollout SU:101-111, SU:135-144; train TR:170-180)
 *
 *   v, v_dtype      model_output (B,n), contiguous, MIXGRPO_F32 | MIXGRPO_BF16
 *   x, x_bs         latents fp32 (B,n) with batch stride x_bs
 *   noise           src==NOISE: (B,n) contiguous; dtype v_dtype for flow, fp32 for dance/dpm
 *   x_next_in,in_bs src==GIVEN: stored next latents fp32
 *   m1, m2          dpm order>=2 / ==3: previous x0 predictions fp32 (B,n) contiguous (DPMState, SU:255-271)
 *   x_next_out,out_bs   fp32 (B,n) or NULL      x0_out  fp32 (B,n) contiguous or NULL
 *   mean_out        fp32 (B,n) contiguous or NULL (the 5-tuple's prev_sample_mean, SU:210)
 *   logp_out        fp32 [B] or NULL
 *   workspace       >= mixgrpo_step_workspace_bytes(B,n), zeroed once at allocation; with MIXGRPO_FLAG_DEFER_LOGP the launch's own
 *                   >= mixgrpo_deferred_workspace_bytes(B,n) block (MIXGRPO_ENOSPACE otherwise)
 *   ordering        x, x_next_in, m1, m2 may have been written by the launch immediately before on the stream (they are read
 *                   with coherent loads behind the dependency wait); v and noise must not have been when a
 *                   MIXGRPO_FLAG_PDL_EARLY_* flag is passed (they are requested before it, through the read-only path)
 */
/* Optional second output of a step launch (nullable `ext`): the latents handed to the VAE, i.e. what the reference
 * computes after the rollout with two more passes over the final latent — unpack_latents (TR:102-115) then
 * `latents / 0.3611 + 0.1159` (TR:286-287).  The rollout's LAST step writes it straight from its registers:
 *   decode_out[b][c][2hp+dh][2wp+dw] = src[b][hp*(W/2)+wp][c*4+dh*2+dw] / divisor + shift,   n == C*H*W
 * with src = x_next (from_x0 == 0; TR:150-152 `latents = z`) or x0 (from_x0 != 0; --drop_last_sample, SU:149-150).
 * fp32 out, (B,C,H,W) contiguous, H and W even; needs the vector path (n % 8 == 0, aligned pointers), C % 2 == 0. */
typedef struct mixgrpo_step_ext {
  float* decode_out;
  int C, H, W;
  float divisor, shift;
  int from_x0;
  int reciprocal;   /* 0: true division (what torch computes on CPU tensors); 1: src * (1.0f / divisor), which is what torch's CUDA
                       div-by-python-scalar kernel computes (one rounding more; differs from true division by <= 1 ulp) */
  /* Trajectory seed (nullable x_f32_out; flow family, vector path, not together with decode_out): the rollout's FIRST step
   * takes `x` as the bf16 initial latent z itself (the `x` argument is then a bf16 pointer, x_bs in elements) and also writes
   * its fp32 widening — all_latents[:, 0], what SU:26 / SU:153's torch.stack promotion produces — to x_f32_out (batch stride
   * x_f32_out_bs): no separate cast launch, and z is read once as 2 B/elem instead of once as 2 and once more as 4. */
  int x_is_bf16;
  float* x_f32_out;
  int64_t x_f32_out_bs;
} mixgrpo_step_ext;

int mixgrpo_flow_step(const void* v, int v_dtype, const float* x, int64_t x_bs,
                      const void* noise, const float* x_next_in, int64_t in_bs,
                      float* x_next_out, int64_t out_bs, float* x0_out, float* mean_out,
                      float* logp_out, void* workspace, int64_t workspace_bytes,
                      int64_t B, int64_t n, const mixgrpo_step_coefs* coefs_host,
                      int src, unsigned flags, void* stream, const mixgrpo_step_ext* ext);

int mixgrpo_dance_step(const void* v, int v_dtype, const float* x, int64_t x_bs,
                       const float* noise, const float* x_next_in, int64_t in_bs,
                       float* x_next_out, int64_t out_bs, float* x0_out, float* mean_out,
                       float* logp_out, void* workspace, int64_t workspace_bytes,
                       int64_t B, int64_t n, const mixgrpo_step_coefs* coefs_host,
                       int src, int sde_solver, unsigned flags, void* stream, const mixgrpo_step_ext* ext);

int mixgrpo_dpm_step(const void* v, int v_dtype, const float* x, int64_t x_bs,
                     const float* noise, const float* m1, const float* m2, int order,
                     float* x_next_out, int64_t out_bs, float* x0_out, float* mean_out,
                     float* logp_out, void* workspace, int64_t workspace_bytes,
                     int64_t B, int64_t n, const mixgrpo_step_coefs* coefs_host,
                     int src, unsigned flags, void* stream, const mixgrpo_step_ext* ext);

/* Finalizes the log-probs of `n_launches` step launches issued with MIXGRPO_FLAG_DEFER_LOGP (stream-ordered after them; ONE
 * launch): launch i accumulated into the records at workspace + i * launch_stride_bytes (launch_stride_bytes >=
 * mixgrpo_deferred_workspace_bytes(B, n), a multiple of 8);
 *   logp_out[i * out_stride + b] = -mean(d^2 / 2 s^2) - log_scale_host[i] - log_norm_host[i]      (SU:201-208)
 * with the per-step scalars of that launch's mixgrpo_step_coefs.  active_host[i] == 0 (nullable = all active): launch i did not
 * accumulate (e.g. a deterministic step whose log-prob was skipped) — its row is filled with NaN.  Records are left zeroed.
 * n_launches <= 4096 (chunks of 64 per kernel launch). */
int mixgrpo_logp_finalize(void* workspace, int64_t launch_stride_bytes, int64_t n_launches, int64_t B,
                          const float* log_scale_host, const float* log_norm_host, const int* active_host,
                          float* logp_out, int64_t out_stride, void* stream);

/* Moves a graph-safe Philox state's base offset: device_state[1] += increment (one tiny launch, stream-ordered after the
 * step launches that consumed the numbers; issue it once per rollout, captured in the same graph). */
int mixgrpo_philox_advance(uint64_t* device_state, uint64_t increment, void* stream);

/* ---- backward of the transition log-prob w.r.t. model_output ----------------------------------
 * Replaces autograd through SU:175-208 / SU:224-250 when TR:585 calls loss.backward():
 *   grad_v = dL/dlogp[b] * (x_next - mean)/scale^2 * dmean/dv / n      (closed form)
 * mean is recomputed from (v, x) in the same pass; grad_logp is a DEVICE vector [B], so no host
 * sync sits between the loss kernel and this one.  grad_v has dtype v_dtype.
 * family: 0 flow, 1 dance (sde_solver=True, the only trained variant TR:159-168), 2 dpm order 1 (the
 * dpm_apply_strategy == "all" training call, TR:169-180: dpm_state=None, so always the first-order update;
 * x_next is the sample that forward call drew; coefs from the same dpm table, c[0] sigma_s, c[6] the x0 coefficient). */
int mixgrpo_logprob_bwd(int family, const void* v, int v_dtype, const float* x, int64_t x_bs,
                        const float* x_next, int64_t in_bs, const float* grad_logp,
                        void* grad_v, int64_t B, int64_t n,
                        const mixgrpo_step_coefs* coefs_host, unsigned flags, void* stream);

/* ---- fused policy update: log-prob + clipped-ratio loss, forward and backward, two launches -------
 * Replaces, for a batch of B stored transitions, grpo_one_step's operator call (TR:149-168), the loss
 * (TR:560-583) and loss.backward() down to d/d model_output (TR:585) with no separate loss launch:
 *   mixgrpo_policy_fwd  writes new log-probs [B]; its per-sample finalizer also adds that sample's
 *                       (loss, policy_loss, kl_loss, clip_frac) — the reference's B == 1 evaluation — to
 *                       stats_rows[b][0..3] (nullable).  Rows are summed by the caller when it logs.
 *   mixgrpo_policy_bwd  evaluates dL/dlogp[b] in place from (new_logp, old_logp, advantages)[b] and
 *                       writes grad_v (dtype v_dtype).
 * denom = gradient_accumulation_steps * len(train_timesteps) (TR:576).  family: 0 flow, 1 dance (sde). */
typedef struct mixgrpo_loss_args {
  const float* old_logp;     /* [B] device */
  const float* advantages;   /* [B] device (unclamped; adv_clip_max is applied inside, TR:560-564) */
  float* stats_rows;         /* [B,4] device; may be NULL */
  double clip_range, adv_clip_max, kl_coeff, denom;
  int accumulate;            /* 1: stats_rows[b] += terms (running sum over window steps); 0: stats_rows[b] = terms */
} mixgrpo_loss_args;

int mixgrpo_policy_fwd(int family, const void* v, int v_dtype, const float* x, int64_t x_bs,
                       const float* x_next, int64_t in_bs, float* logp_out, void* workspace,
                       int64_t workspace_bytes, int64_t B, int64_t n, const mixgrpo_step_coefs* coefs_host,
                       const mixgrpo_loss_args* loss, unsigned flags, void* stream);

int mixgrpo_policy_bwd(int family, const void* v, int v_dtype, const float* x, int64_t x_bs,
                       const float* x_next, int64_t in_bs, const float* new_logp,
                       const mixgrpo_loss_args* loss, void* grad_v, int64_t B, int64_t n,
                       const mixgrpo_step_coefs* coefs_host, unsigned flags, void* stream);

/* ---- the window's policy updates batched: ONE forward launch and ONE backward launch for n_items (<= 8) independent
 * (batch, window step) updates — the 4 SDE-window steps of TR:536-585 are independent of one another given their model
 * outputs, so their log-prob + loss forwards share a launch (grid.y = n_items * B: 126 MB instead of 4 x 31 MB at group
 * 12, where a 31 MB launch is ~2 us of ramp/drain on 5 us of streaming) and so do their backwards.  Item j has its own
 * tensors, per-step scalar block, old log-probs and stats rows; advantages [B] and the loss scalars are shared.  Results
 * are bit-identical to n_items calls of mixgrpo_policy_fwd / mixgrpo_policy_bwd.  workspace >=
 * mixgrpo_step_workspace_bytes(n_items * B, n).  Ragged / unaligned tensors return MIXGRPO_EUNSUPPORTED (nothing
 * launched): call the single-item entry points then. */
#define MIXGRPO_POLICY_MAX_ITEMS 8
typedef struct mixgrpo_policy_item {
  const void* v;             /* model output (B,n), dtype v_dtype, contiguous */
  const float* x;            /* latents before the step, batch stride x_bs */
  const float* x_next;       /* stored latents after the step, batch stride in_bs */
  int64_t x_bs, in_bs;
  float* logp;               /* [B]: written by _fwd_multi, read by _bwd_multi (the new log-probs) */
  const float* old_logp;     /* [B] */
  float* stats_rows;         /* [B,4] or NULL (forward) */
  void* grad_v;              /* (B,n) dtype v_dtype (backward) */
  mixgrpo_step_coefs coefs;  /* this step's scalar block */
} mixgrpo_policy_item;

int mixgrpo_policy_fwd_multi(int family, int v_dtype, const mixgrpo_policy_item* items_host, int n_items,
                             const float* advantages, double clip_range, double adv_clip_max, double kl_coeff,
                             double denom, int accumulate, void* workspace, int64_t workspace_bytes, int64_t B,
                             int64_t n, unsigned flags, void* stream);
int mixgrpo_policy_bwd_multi(int family, int v_dtype, const mixgrpo_policy_item* items_host, int n_items,
                             const float* advantages, double clip_range, double adv_clip_max, double kl_coeff,
                             double denom, int64_t B, int64_t n, unsigned flags, void* stream);

/* ---- single-pass policy update: log-prob forward + loss + log-prob backward in ONE launch, 12 B/elem ----------
 * Same inputs, outputs and bits as mixgrpo_policy_fwd followed by mixgrpo_policy_bwd (new log-probs [B], stats rows,
 * grad_v), but v / x / x_next are read ONCE: a persistent, co-resident (cooperatively launched) grid keeps the
 * residuals x_next - mean in shared memory while the per-sample sums complete, then turns them into grad_v
 * (csrc/policy_kernels.cu).  Returns MIXGRPO_EUNSUPPORTED — nothing launched — when the residuals do not fit on
 * chip (more than ~27 tiles of 2048 scalars per SM, i.e. B*n > ~8 M), for ragged / unaligned tensors (n % 8, strides,
 * 32-byte alignment); call the two-launch pair then.  workspace as for the step kernels; it must not be shared with
 * a launch that can run concurrently.  A wait longer than the timeout (mixgrpo_set_tuning key 5, default 10 s) sets
 * the workspace's status word and yields NaN gradients instead of hanging the GPU. */
int mixgrpo_policy_step(int family, const void* v, int v_dtype, const float* x, int64_t x_bs,
                        const float* x_next, int64_t in_bs, float* logp_out, void* grad_v, void* workspace,
                        int64_t workspace_bytes, int64_t B, int64_t n, const mixgrpo_step_coefs* coefs_host,
                        const mixgrpo_loss_args* loss, unsigned flags, void* stream);

/* ---- reward -> group-relative advantage -------------------------------------------------------
 * Replaces TR:439-501.  rewards is [n_models, local_B] fp32 (one all-gathered or rank-local
 * matrix), groups are consecutive runs of num_generations samples (TR:444-450).
 *   A[b] = sum_m w[m] * (r[m,b] - mean_g) / (std_g + 1e-8),  std Bessel-corrected (TR:459-461)
 * trim_size > 0: mean/std over the group's rewards with the trim_size smallest removed (TR:451-457).
 * use_group == 0: single-model global normalisation with statistics of `stat_rewards`
 * [n_stat] (the all-gathered vector, TR:498).  weights may be NULL only when n_models == 1
 * (reward_aggr, TR:470-491: the advantage is stored unweighted). */
int mixgrpo_group_advantages(const float* rewards, const float* weights, int n_models,
                             int64_t local_B, int num_generations, int trim_size,
                             int use_group, const float* stat_rewards, int64_t n_stat,
                             float* advantages, void* stream);

/* ---- clipped-ratio GRPO loss, forward + closed-form backward ---------------------------------
 * Replaces TR:560-583 and the scalar part of TR:585.  All vectors are [B] fp32 on the device.
 *   stats_out[0..3] = loss, policy_loss, kl_loss, clip_frac          (TR:575-583)
 *   grad_new_logp[b] = dloss/dnew_logp[b]  (may be NULL)
 * denom = gradient_accumulation_steps * len(train_timesteps) (TR:576).  The four scalars are the
 * python floats of the reference call site (doubles); they are narrowed to fp32 exactly where torch
 * narrows them (clamp bounds 1-clip / 1+clip are formed in double first, TR:571-572).
 * stats_accum (nullable, [4]) += stats_out: device-side running sums replacing the four
 * all_reduce+.item() pairs per (sample, step) at TR:586-600. */
int mixgrpo_grpo_loss(const float* new_logp, const float* old_logp, const float* advantages,
                      int64_t B, double clip_range, double adv_clip_max, double kl_coeff, double denom,
                      float* stats_out, float* grad_new_logp, float* stats_accum, void* stream);

/* ---- fused peer-memory exchange: reward gather + advantages, logging all-reduce (one node, NVLink) ----
 * The path's only collective as ONE kernel per exchange instead of NCCL + a compute launch:
 *   mixgrpo_peer_gather_advantages  replaces gather_tensor per reward model (TR:332-338, TR:417-425) AND the
 *                                   advantage computation (TR:439-501): every rank pushes its [n_models, local_B]
 *                                   rewards into every peer's region with st.global over NVLink, publishes a
 *                                   sequence flag (st.release.sys), waits for the peers' flags (ld.acquire.sys) and
 *                                   computes its advantages from the gathered matrix in the same CTA.
 *   mixgrpo_peer_allreduce          replaces the all_reduce(AVG)+.item() pairs of TR:586-600 for <= 256 floats:
 *                                   contributions are summed in rank order, so every rank gets the same bits.
 * A "region" is mixgrpo_peer_region_bytes(world, cap_floats) bytes of cudaMalloc'ed memory per rank
 * (cap_floats >= n_models*local_B), created with _alloc (which also returns a 64-byte CUDA IPC handle to ship to
 * the peers by any host channel, e.g. torch.distributed.all_gather_object) and mapped by peers with _open.
 * regions_host[q] is rank q's region as mapped in the CALLING process (its own allocation at [rank]).  All ranks
 * must issue the same sequence of exchange calls (collective semantics); the call counter lives in the region, so
 * the launches are CUDA-graph capturable.  A wait that exceeds the timeout (mixgrpo_set_tuning key 2, default
 * 10 min, the order of NCCL's watchdog: ranks may be skewed by a slow reward model) sets the region's status word and yields NaN outputs instead of hanging the GPU.
 * Several "ranks" may live on ONE device (regions_host = plain device pointers, one stream per rank): that is how
 * the single-GPU parity tests drive the protocol. */
#define MIXGRPO_PEER_MAX_WORLD 16
#define MIXGRPO_PEER_HANDLE_BYTES 64

/* which columns form a group */
enum {
  MIXGRPO_ADV_GROUP_LOCAL = 0,  /* consecutive runs of num_generations of THIS rank's samples (TR:443-461)            */
  MIXGRPO_ADV_GROUP_SPLIT = 1,  /* consecutive runs in the rank-major gathered order: a group may span ranks (§8e)    */
  MIXGRPO_ADV_GLOBAL = 2        /* no groups: statistics of the whole gathered vector, one model (TR:495-499)         */
};

int64_t mixgrpo_peer_region_bytes(int world, int64_t cap_floats);
int mixgrpo_peer_region_alloc(int world, int64_t cap_floats, void** region_out, void* ipc_handle_out /* 64 B or NULL */);
int mixgrpo_peer_region_open(const void* ipc_handle, void** region_out);
int mixgrpo_peer_region_close(void* region);   /* a region obtained from _open  */
int mixgrpo_peer_region_free(void* region);    /* a region obtained from _alloc */
/* synchronising debug read of (gather calls completed, all-reduce calls completed, status: 1 = a wait timed out) */
int mixgrpo_peer_region_status(const void* region, int* seq_gather_host, int* seq_reduce_host, int* status_host);

/*   rewards [n_models, local_B] fp32 (this rank), weights [n_models] or NULL (single model)
 *   gathered_out [n_models, world*local_B] fp32 or NULL: column q*local_B + b = rank q's sample b (torch.cat order)
 *   advantages [local_B] fp32: this rank's samples; entries outside a whole group are 0 (TR:445) */
int mixgrpo_peer_gather_advantages(void* const* regions_host, int rank, int world, int64_t cap_floats,
                                   const float* rewards, const float* weights, int n_models, int64_t local_B,
                                   int num_generations, int trim_size, int mode, float* gathered_out,
                                   float* advantages, void* stream);
/*   values [count <= 256] fp32, reduced in place: sum over ranks (in rank order), divided by world when average != 0 */
int mixgrpo_peer_allreduce(void* const* regions_host, int rank, int world, int64_t cap_floats, float* values,
                           int count, int average, void* stream);

/* ---- layout helpers either side of the path ----------------------------------------------------
 * mixgrpo_pack_latents    (B,C,H,W) -> (B,(H/2)(W/2),4C)     TR:94-99   (src dtype bf16|f32 -> same)
 * mixgrpo_unpack_latents  inverse, fused with the VAE de-normalisation  x/divisor + shift
 *                         (TR:102-115 then TR:287 `latents / 0.3611 + 0.1159`; pass divisor=1, shift=0
 *                         for the plain permute).  H, W are the UNPACKED latent height/width. */
int mixgrpo_pack_latents(const void* src, void* dst, int dtype, int64_t B, int C, int H, int W,
                         void* stream);
/* mixgrpo_cast_rows       (B,n) bf16|f32 contiguous -> fp32 rows with batch stride dst_bs: seeds slot 0 of the
 *                         fp32 trajectory with the bf16 initial latents (SU:26, the list head `z` that torch.stack
 *                         promotes at SU:153). */
int mixgrpo_cast_rows(const void* src, int src_dtype, float* dst, int64_t dst_bs, int64_t B, int64_t n, void* stream);
int mixgrpo_unpack_latents(const void* src, void* dst, int dtype, int64_t B, int C, int H, int W,
                           float divisor, float shift, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MIXGRPO_B200_H */
