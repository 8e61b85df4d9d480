// Single-pass policy update: transition log-prob forward + clipped-ratio loss + log-prob backward in ONE launch that
// touches every latent byte once (v 2 B + x 4 B + x_next 4 B read, grad_v 2 B written = 12 B/elem instead of the
// 10 + 12 = 22 B/elem of mixgrpo_policy_fwd followed by mixgrpo_policy_bwd).
//
// Replaces, for B stored transitions: grpo_one_step's operator call (TR:149-168), the loss (TR:560-583) and
// loss.backward() down to d/d model_output (TR:585).  TR = /root/reference/fastvideo/train_grpo_flux.py.
//
// Why it needs a different shape than the streaming kernels: dL/dlogp[b] depends on the COMPLETE per-sample sum of
// (x_next - mean)^2, i.e. on every tile of sample b, while grad_v needs the per-element residual d = x_next - mean
// again.  So the residuals stay ON CHIP between the two halves: a persistent grid of (SMs x R) co-resident CTAs
// (cooperative launch) owns contiguous runs of 2048-scalar tiles, keeps each tile's d in shared memory (8 KB per tile;
// 12 x 4096 x 64 fp32 residuals = 12.6 MB of the chip's 33 MB), and
//   1. streams v, x, x_next once (next tile's loads in flight during this tile's math), d -> shared memory, per-tile
//      sums reduced exactly like mg::step_kernel does (same order -> the log-probs are BIT-IDENTICAL to mixgrpo_policy_fwd);
//   2. one packed fixed-point atomicAdd per tile into the sample's 64-bit accumulator; the last arriver finalizes
//      logp[b], the sample's loss terms / stats row, re-zeroes the word and bumps the sample's epoch (st.release.gpu);
//   3. one thread per owned tile waits for its sample's epoch to move (ld.acquire.gpu, bounded by a timeout), evaluates
//      dL/dlogp[b] in place from (new_logp, old_logp, advantage)[b] — the loss never exists as a launch;
//   4. grad_v = chain(d) straight from shared memory: the arithmetic of mg::logprob_bwd_kernel (bit-identical grads).
// Shapes whose residuals do not fit on chip (> SMs x 27 tiles), ragged / unaligned tensors and the dpm family return
// MIXGRPO_EUNSUPPORTED: the caller then issues the two-launch path (both are CUDA; there is no CPU path).
#include "step_math.cuh"

namespace mg {

struct PolicyStepParams {
  const void* v;
  const float* x;
  const float* x_next;
  void* grad_v;
  float* logp_out;
  unsigned long long* acc;     // workspace records, 32 B per sample: acc[kWsStride*b] = packed accumulator (as the step kernels use it)
  uint32_t* epoch;             // epoch[2*kWsStride*b]: bumped once per launch per sample by the sample's finalizer
  uint32_t* status;            // set to 1 when a wait timed out
  long long n, x_bs, in_bs, T; // T = B * tps tiles in total
  int B, tps;                  // tiles per sample
  mixgrpo_step_coefs k;
  LpQuant lpq;                 // host-evaluated (step_math.cuh)
  LossParams loss;
  unsigned long long timeout_ns;
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_gpu(const float* p) {
  float v;
  asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns_p() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// shared memory per CTA for K owned tiles: residuals as two float4 planes (conflict-free 128-bit accesses), then
// per-tile warp partials, start epochs and per-tile backward scalars
__host__ __device__ inline size_t policy_smem_bytes(int K) {
  return (size_t)K * (kTile * sizeof(float) + (kThreads / 32) * sizeof(unsigned long long) + sizeof(uint32_t) + sizeof(float));
}

// PF: software-prefetch the next tile's inputs into registers (<= 3 CTAs/SM); without it the kernel fits 42 registers
// and 6 CTAs/SM overlap each other's loads, which is what the streaming kernels found fastest on B200.
template <int FAM, class VT, bool RND, bool PF>
__global__ void __launch_bounds__(kThreads, PF ? 3 : 6) policy_step_kernel(const __grid_constant__ PolicyStepParams p, const int Kmax) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char s_raw[];
  float4* s_lo = reinterpret_cast<float4*>(s_raw);                       // [Kmax][256] residuals 0..3 of each thread
  float4* s_hi = s_lo + (size_t)Kmax * kThreads;                         // [Kmax][256] residuals 4..7
  unsigned long long* s_part = reinterpret_cast<unsigned long long*>(s_hi + (size_t)Kmax * kThreads);   // [Kmax][8] per-warp parts of the packed word
  uint32_t* s_e0 = reinterpret_cast<uint32_t*>(s_part + (size_t)Kmax * (kThreads / 32));   // [Kmax] epoch at kernel start
  float* s_gs = reinterpret_cast<float*>(s_e0 + Kmax);                   // [Kmax] (dL/dlogp / n) / (2 s^2) of the tile's sample

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // tile bookkeeping in 32 bits: T = B * tps < 65536 * 2048, in-sample offsets < 2047 * 2048 + 2048
  const int t_lo = (int)(((long long)blockIdx.x * p.T) / gridDim.x), t_hi = (int)(((long long)(blockIdx.x + 1) * p.T) / gridDim.x);
  const int K = t_hi - t_lo;                                             // 1 <= K <= Kmax (host guarantees G <= T)
  const int n = (int)p.n;
  const LpQuant lpq = p.lpq;
  const VT* vbase = reinterpret_cast<const VT*>(p.v);

  if (tid < K) s_e0[tid] = ld_relaxed_gpu(p.epoch + 2 * kWsStride * ((t_lo + tid) / p.tps));

  // ---- 1. stream the inputs once; residuals stay in shared memory
  auto tile_pos = [&](int k, int& b, int& off) {
    const int tt = t_lo + k;
    b = tt / p.tps;
    off = (tt - b * p.tps) * kTile + tid * kVec;
  };
  float v[kVec], x[kVec], a[kVec];
  bool act = false;
  auto load_tile = [&](int k, float (&vv)[kVec], float (&xx)[kVec], float (&aa)[kVec]) -> bool {
    int b, off;
    tile_pos(k, b, off);
    if (off >= n) return false;
    ld_stream(vbase + ((long long)b * n + off), vv);
    ld_dep(p.x + ((long long)b * p.x_bs + off), xx);
    ld_dep(p.x_next + ((long long)b * p.in_bs + off), aa);
    return true;
  };
  if constexpr (PF) act = load_tile(0, v, x, a);
  for (int k = 0; k < K; ++k) {
    float vn[kVec], xn[kVec], an[kVec];
    bool actn = false;
    if constexpr (PF) {
      if (k + 1 < K) actn = load_tile(k + 1, vn, xn, an);               // next tile's loads fly during this tile's math
    } else {
      act = load_tile(k, v, x, a);
    }
    float d[kVec];
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < kVec; j += 2) {                                  // pair-wise, like mg::step_kernel
      float v2[2] = {0.f, 0.f}, x2[2] = {0.f, 0.f}, a2[2] = {0.f, 0.f};
      if (act) { v2[0] = v[j]; v2[1] = v[j + 1]; x2[0] = x[j]; x2[1] = x[j + 1]; a2[0] = a[j]; a2[1] = a[j + 1]; }
      const float z2[2] = {0.f, 0.f};
      float o2[2], x02[2], mu2[2], dd2[2];
      tile_math<FAM, MIXGRPO_SRC_GIVEN, 1, RND, true>(p.k, v2, x2, a2, z2, z2, o2, x02, mu2, dd2);
      d[j] = __fsub_rn(o2[0], mu2[0]);
      d[j + 1] = __fsub_rn(o2[1], mu2[1]);
      acc += dd2[0] + dd2[1];
    }
    if (!act) {
      acc = 0.f;
#pragma unroll
      for (int j = 0; j < kVec; ++j) d[j] = 0.f;
    }
    s_lo[(size_t)k * kThreads + tid] = make_float4(d[0], d[1], d[2], d[3]);
    s_hi[(size_t)k * kThreads + tid] = make_float4(d[4], d[5], d[6], d[7]);
    {
      int bk, offk;
      tile_pos(k, bk, offk);
      const unsigned long long part = warp_part(acc, lpq, p.acc + kWsStride * bk);      // step_math.cuh: integer from the thread up
      if (lane == 0) s_part[k * (kThreads / 32) + warp] = part;
    }
    if constexpr (PF) {
      if (k + 1 < K) {
#pragma unroll
        for (int j = 0; j < kVec; ++j) { v[j] = vn[j]; x[j] = xn[j]; a[j] = an[j]; }
        act = actn;
      }
    }
  }
  __syncthreads();

  // ---- 2. one packed atomic per owned tile; the last arriver of a sample finalizes it and bumps its epoch
  int my_b = 0;
  if (tid < K) {
    my_b = (t_lo + tid) / p.tps;
    unsigned long long* rec = p.acc + kWsStride * my_b;
    const unsigned long long add = cta_word<kThreads / 32>(s_part + tid * (kThreads / 32));
    const unsigned long long old = atomicAdd(rec, add);
    if ((old & kArrivalMask) == (unsigned long long)(p.tps - 1)) {
      const float q = packed_total(old + add, rec);
      const float lp = __fsub_rn(__fsub_rn(-q, p.k.log_scale), p.k.log_norm);      // SU:201-208
      p.logp_out[my_b] = lp;
      p.acc[kWsStride * my_b] = 0ull;
      if (p.loss.rows) {                                                             // TR:560-583, the reference's B == 1 evaluation
        const LossTerms lt = loss_terms(lp, p.loss.old_lp[my_b], p.loss.adv[my_b], p.loss, 1.f);
        const float policy = __fdiv_rn(lt.policy_num, p.loss.denom);
        const float kl = __fdiv_rn(__fmul_rn(0.5f, lt.kl_num), p.loss.denom);
        float* row = p.loss.rows + 4 * (long long)my_b;
        const float4 prev = p.loss.accumulate ? *reinterpret_cast<const float4*>(row) : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(row) = make_float4(prev.x + __fadd_rn(policy, __fmul_rn(p.loss.klc, kl)), prev.y + policy,
                                                      prev.z + kl, prev.w + lt.clip);
      }
      __threadfence();
      st_release_gpu(p.epoch + 2 * kWsStride * my_b, s_e0[tid] + 1u);
    }
  }

  // ---- 3. wait for the owned tiles' samples, evaluate dL/dlogp in place
  if (tid < K) {
    const uint32_t e0 = s_e0[tid];
    float gs = __int_as_float(0x7fc00000);
    bool ready = ld_acquire_gpu(p.epoch + 2 * kWsStride * my_b) != e0;
    if (!ready) {
      const unsigned long long t0 = global_ns_p();
      unsigned spins = 0;
      while (!(ready = ld_acquire_gpu(p.epoch + 2 * kWsStride * my_b) != e0)) {
        if ((++spins & 0xffu) == 0 && p.timeout_ns && global_ns_p() - t0 > p.timeout_ns) {
          *reinterpret_cast<volatile uint32_t*>(p.status) = 1u;
          break;
        }
      }
    }
    if (ready) {
      const float lp = ld_relaxed_gpu(p.logp_out + my_b);
      const float g_lp = loss_terms(lp, ld_dep(p.loss.old_lp + my_b), ld_dep(p.loss.adv + my_b), p.loss, 1.f).grad;   // TR:560-585
      gs = __fdiv_rn(__fdiv_rn(g_lp, (float)n), p.k.two_var);            // (g/n)/(2 s^2): the two divisions autograd performs
    }
    s_gs[tid] = gs;
  }
  __syncthreads();

  // ---- 4. grad_v from the on-chip residuals: the chain of mg::logprob_bwd_kernel
  const float* cf = p.k.c;
  for (int k = 0; k < K; ++k) {
    int b, off;
    tile_pos(k, b, off);
    if (off >= n) continue;
    const float gs = s_gs[k];
    const float4 lo = s_lo[(size_t)k * kThreads + tid], hi = s_hi[(size_t)k * kThreads + tid];
    const float d[kVec] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    float g[kVec];
    if constexpr (FAM == kFlow) {
#pragma unroll
      for (int i = 0; i < kVec; ++i) g[i] = __fmul_rn(gs, __fmul_rn(2.f, d[i]));
      round_like_torch<RND>(g);
#pragma unroll
      for (int i = 0; i < kVec; ++i) g[i] = __fmul_rn(g[i], cf[3]);
      round_like_torch<RND>(g);
#pragma unroll
      for (int i = 0; i < kVec; ++i) g[i] = __fmul_rn(g[i], cf[2]);
    } else {
      float g1[kVec], g2[kVec];
#pragma unroll
      for (int i = 0; i < kVec; ++i) {
        const float gm = __fmul_rn(gs, __fmul_rn(2.f, d[i]));
        g1[i] = gm;
        g2[i] = -__fmul_rn(__fdiv_rn(__fmul_rn(__fmul_rn(gm, cf[5]), cf[4]), cf[3]), cf[2]);
      }
      round_like_torch<RND>(g1);
      round_like_torch<RND>(g2);
#pragma unroll
      for (int i = 0; i < kVec; ++i) { g1[i] = __fmul_rn(g1[i], cf[1]); g2[i] = __fmul_rn(g2[i], cf[0]); }
      round_like_torch<RND>(g1);
      round_like_torch<RND>(g2);
#pragma unroll
      for (int i = 0; i < kVec; ++i) g[i] = __fadd_rn(g1[i], g2[i]);
    }
    st_stream(reinterpret_cast<VT*>(p.grad_v) + ((long long)b * n + off), g);
  }
}

static int g_policy_ctas_per_sm = 6;        // mixgrpo_set_tuning key 3 (<= 3 selects the register-prefetch variant)
static int g_policy_cooperative = 1;        // mixgrpo_set_tuning key 4
static unsigned long long g_policy_timeout_ms = 10000ull;   // key 5 (0 = wait forever)

// launch configuration of one (device, T) pair: resolved once (occupancy query), reused by every later launch
struct PolicyCfg { int dev = -1, r_req = 0, K = 0; long long T = 0, G = 0; size_t smem = 0; };

template <int FAM, class VT, bool RND, bool PF>
static int launch_policy_pf(PolicyStepParams& p, cudaStream_t st) {
  static PolicyCfg cache;
  auto kern = policy_step_kernel<FAM, VT, RND, PF>;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (cache.dev != dev || cache.T != p.T || cache.r_req != g_policy_ctas_per_sm) {
    int sms = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    PolicyCfg c;
    // largest co-resident grid: R CTAs per SM such that R x (K tiles of residuals) fits in shared memory
    for (int R = g_policy_ctas_per_sm; R >= 1 && c.G == 0; --R) {
      long long G = (long long)sms * R;
      if (G > p.T) G = p.T;
      const int K = (int)((p.T + G - 1) / G);
      const size_t smem = policy_smem_bytes(K);
      if (smem > 200 * 1024) continue;
      int occ = 0;
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem);
      if (e != cudaSuccess) return (int)e;
      if ((long long)occ * sms < G) continue;
      c.G = G; c.K = K; c.smem = smem;
    }
    if (c.G == 0) return MIXGRPO_EUNSUPPORTED;          // the residuals do not fit on chip: use policy_fwd + policy_bwd
    c.dev = dev; c.T = p.T; c.r_req = g_policy_ctas_per_sm;
    cache = c;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)cache.G); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = cache.smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  int na = 0;
  // cooperative launch = the driver guarantees that all G CTAs are co-resident, which the epoch wait relies on
  if (g_policy_cooperative) { at[na].id = cudaLaunchAttributeCooperative; at[na].val.cooperative = 1; ++na; }
  else if (g_use_pdl) { at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[na].val.programmaticStreamSerializationAllowed = 1; ++na; }
  cfg.attrs = at; cfg.numAttrs = na;
  e = cudaLaunchKernelEx(&cfg, kern, p, cache.K);
  return (int)(e != cudaSuccess ? e : cudaGetLastError());
}

template <int FAM, class VT, bool RND>
static int launch_policy(PolicyStepParams& p, cudaStream_t st) {
  return g_policy_ctas_per_sm <= 3 ? launch_policy_pf<FAM, VT, RND, true>(p, st) : launch_policy_pf<FAM, VT, RND, false>(p, st);
}

}  // namespace mg

using namespace mg;

int mixgrpo_policy_set_tuning(int key, int value) {   // reached through mixgrpo_set_tuning(3|4|5, value)
  int old = -1;
  if (key == 3) { if (value < 1 || value > 8) return MIXGRPO_EINVAL; old = g_policy_ctas_per_sm; g_policy_ctas_per_sm = value; }
  else if (key == 4) { if (value != 0 && value != 1) return MIXGRPO_EINVAL; old = g_policy_cooperative; g_policy_cooperative = value; }
  else if (key == 5) { if (value < 0) return MIXGRPO_EINVAL; old = (int)g_policy_timeout_ms; g_policy_timeout_ms = (unsigned long long)value; }
  return old;
}

extern "C" __attribute__((visibility("default"))) int mixgrpo_policy_step(int family, const void* v, int v_dtype, const float* x, int64_t x_bs,
                                                                          const float* x_next, int64_t in_bs, float* logp_out, void* grad_v,
                                                                          void* workspace, int64_t workspace_bytes, int64_t B, int64_t n,
                                                                          const mixgrpo_step_coefs* coefs_host, const mixgrpo_loss_args* loss,
                                                                          unsigned flags, void* stream) {
  if (!v || !x || !x_next || !logp_out || !grad_v || !coefs_host || !loss || !loss->old_logp || !loss->advantages || B <= 0 || B > 65535 || n <= 0)
    return MIXGRPO_EINVAL;
  if (v_dtype != MIXGRPO_F32 && v_dtype != MIXGRPO_BF16) return MIXGRPO_EINVAL;
  if (family != kFlow && family != kDance) return MIXGRPO_EINVAL;
  if (!workspace) return MIXGRPO_EINVAL;
  if (workspace_bytes < mixgrpo_step_workspace_bytes(B, n)) return MIXGRPO_ENOSPACE;
  if (loss->stats_rows && (reinterpret_cast<uintptr_t>(loss->stats_rows) % 16) != 0) return MIXGRPO_EINVAL;
  auto al = [](const void* q, size_t a) { return (reinterpret_cast<uintptr_t>(q) % a) == 0; };
  const size_t va = v_dtype == MIXGRPO_BF16 ? 16 : 32;
  const bool vec = (n % kVec == 0) && (x_bs % kVec == 0) && (in_bs % kVec == 0) && al(v, va) && al(grad_v, va) && al(x, 32) && al(x_next, 32);
  const long long tps = (n + kTile - 1) / kTile;
  if (!vec || tps > kMaxCtasPerSample) return MIXGRPO_EUNSUPPORTED;
  PolicyStepParams p;
  p.v = v; p.x = x; p.x_next = x_next; p.grad_v = grad_v; p.logp_out = logp_out;
  p.acc = reinterpret_cast<unsigned long long*>(workspace);
  p.epoch = reinterpret_cast<uint32_t*>(p.acc) + 2;      // record b: u64 acc | u32 epoch | u32 (status in record 0)
  p.status = reinterpret_cast<uint32_t*>(p.acc) + 3;
  p.n = n; p.x_bs = x_bs; p.in_bs = in_bs; p.B = (int)B; p.tps = (int)tps; p.T = (long long)B * tps;
  p.k = *coefs_host;
  p.lpq = lp_quant(n, coefs_host->two_var);
  p.loss = make_loss_params(loss->old_logp, loss->advantages, loss->stats_rows, loss->clip_range, loss->adv_clip_max, loss->kl_coeff, loss->denom);
  p.loss.accumulate = loss->accumulate ? 1 : 0;
  p.timeout_ns = g_policy_timeout_ms * 1000000ull;
  const bool rnd = (flags & MIXGRPO_FLAG_ROUND_LIKE_TORCH) != 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (family == kFlow) {
    if (v_dtype == MIXGRPO_F32) return launch_policy<kFlow, float, false>(p, st);
    return rnd ? launch_policy<kFlow, __nv_bfloat16, true>(p, st) : launch_policy<kFlow, __nv_bfloat16, false>(p, st);
  }
  if (v_dtype == MIXGRPO_F32) return launch_policy<kDance, float, false>(p, st);
  return rnd ? launch_policy<kDance, __nv_bfloat16, true>(p, st) : launch_policy<kDance, __nv_bfloat16, false>(p, st);
}
