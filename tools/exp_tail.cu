// Measurement-only variants of the Euler-ODE sampler step's log-prob tail (round 2, profiles/r02_design_space.md §4).
// Not product code: built by tools/exp_tail.py into build/libexp_tail.so and timed exactly like bench.py's roofline leg.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -shared -Xcompiler -fPIC tools/exp_tail.cu -o build/libexp_tail.so
// Every variant does the product kernel's loads, arithmetic and stores for <flow, bf16, SRC_DETERMINISTIC, OUT=0> with the model
// output requested ahead of the dependency wait; only what happens after the last store differs:
//   TAIL 0  what ships: 5 shuffles -> shared -> CTA barrier -> 3 shuffles -> ONE fire-and-forget 64-bit reduction per CTA
//   TAIL 1  5 shuffles -> one reduction PER WARP into one of 64 sub-words of the sample (no shared memory, no barrier)
//   TAIL 2  per-thread fixed point -> ONE redux.sync.add.u32 -> one reduction per warp into a sub-word
//   TAIL 3  per-thread fixed point -> redux -> shared -> barrier -> redux -> one reduction per CTA
//   TAIL 4  no log-prob at all (the floor)
#include "../mixgrpo_b200/csrc/step_kernel.cuh"

namespace mg {
int g_use_pdl = 1;
int g_max_ctas_per_sample = kMaxCtasPerSample;
}  // namespace mg
int mixgrpo_peer_set_timeout_ms(int) { return 0; }
int mixgrpo_policy_set_tuning(int, int) { return 0; }
extern "C" int64_t mixgrpo_step_workspace_bytes(int64_t B, int64_t) { return B * 32; }
extern "C" int64_t mixgrpo_deferred_workspace_bytes(int64_t B, int64_t) { return B * 256; }
namespace mg { int g_half_ctas = 0; std::atomic<long long> g_half_launches{0}; int g_bwd_threads = 256; int sm_count() { return 148; } }

using namespace mg;



// the CTA-level conversion the float variants use (what shipped before the reduction became integer from the thread up)
__device__ __forceinline__ unsigned long long exp_packed_part(float r, int parts, unsigned long long* rec) {
  const float cap = 255.0f / (float)parts;
  unsigned long long add = 0ull;
  if (!(r >= 0.f && r <= cap)) {
    atomicAdd(rec + kWsWide, __float2ull_rn(r * 16777216.0f));
    add = 1ull << kCountBits;
    r = 0.f;
  }
  return add + (__float2ull_rn(r * 4294967296.0f) << (kCountBits + kPoisonBits));
}

__device__ __forceinline__ void red_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

template <int TAIL, int THREADS, int MINB, int UNROLL>
__global__ void __launch_bounds__(THREADS, MINB) ode_exp_kernel(const __grid_constant__ StepParams p, unsigned long long* sub, float scale, int nsub, int stride) {
  constexpr int TILE = THREADS * kVec;
  constexpr int WARPS = THREADS / 32;
  const int b = blockIdx.y;
  const long long n = p.n;
  const __nv_bfloat16* vp = reinterpret_cast<const __nv_bfloat16*>(p.v) + (long long)b * n;
  const float* xp = p.x + (long long)b * p.x_bs;
  float* op = p.x_out + (long long)b * p.out_bs;
  float v[UNROLL][kVec], x[UNROLL][kVec];
  long long off[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) off[u] = ((long long)blockIdx.x * UNROLL + u) * TILE + threadIdx.x * kVec;
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) if (off[u] < n) ld_stream(vp + off[u], v[u]);
  pdl_prologue();
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) if (off[u] < n) ld_dep(xp + off[u], x[u]);        // coherent, behind the wait (common.cuh)
  float acc = 0.f;
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    if (off[u] >= n) continue;
    float xn[kVec];
#pragma unroll
    for (int j = 0; j < kVec; j += 2) {
      const float v2[2] = {v[u][j], v[u][j + 1]}, x2[2] = {x[u][j], x[u][j + 1]}, a2[2] = {0.f, 0.f};
      float xn2[2], x02[2], mu2[2], dd2[2];
      tile_math<kFlow, MIXGRPO_SRC_DETERMINISTIC, 1, true, false>(p.k, v2, x2, a2, a2, a2, xn2, x02, mu2, dd2);
      xn[j] = xn2[0]; xn[j + 1] = xn2[1];
      acc += dd2[0] + dd2[1];
    }
    st_stream(op + off[u], xn);
  }
  if constexpr (TAIL == 4) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long* rec = p.acc + kWsStride * b;
  // TAIL 1, 2: one sub-word per warp; TAIL 5, 6: one per CTA.  `stride` words apart (16 = a 128-byte line of its own)
  const int unit = (TAIL == 5 || TAIL == 6) ? blockIdx.x : blockIdx.x * WARPS + warp;
  unsigned long long* mine = sub + (long long)b * (256 * 16) + (long long)(unit & (nsub - 1)) * stride;   // the sample's 256-line area (exp_collect_kernel)
  const int per_sub = (gridDim.x * WARPS + nsub - 1) / nsub;

  if constexpr (TAIL == 0 || TAIL == 6) {
    __shared__ float s_warp[WARPS];
    acc = warp_sum(acc);
    if (lane == 0) s_warp[warp] = acc;
    __syncthreads();
    if (warp != 0) return;
    float t = lane < WARPS ? s_warp[lane] : 0.f;
#pragma unroll
    for (int o = WARPS / 2; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) {
      const float r = __fdiv_rn(t, __fmul_rn((float)n, p.k.two_var));
      red_u64(TAIL == 6 ? mine : rec, exp_packed_part(r, 2 * (int)gridDim.x, rec) + 1ull);
    }
  } else if constexpr (TAIL == 1) {
    acc = warp_sum(acc);
    if (lane == 0) {
      const float r = __fdiv_rn(acc, __fmul_rn((float)n, p.k.two_var));
      red_u64(mine, exp_packed_part(r, 2 * per_sub, rec) + 1ull);
    }
  } else if constexpr (TAIL == 2) {
    const float q = acc * scale;                               // units of 2^-32 of mean(d^2 / 2 s^2)
    if (__all_sync(0xffffffffu, q < 67108864.f)) {            // 32 lanes x 2^26 < 2^32 (NaN fails the test)
      const unsigned tot = __reduce_add_sync(0xffffffffu, __float2uint_rn(q));
      if (lane == 0) red_u64(mine, ((unsigned long long)tot << (kCountBits + kPoisonBits)) + 1ull);
    } else {
      acc = warp_sum(acc);
      if (lane == 0) {
        const float r = __fdiv_rn(acc, __fmul_rn((float)n, p.k.two_var));
        red_u64(mine, exp_packed_part(r, 2 * per_sub, rec) + 1ull);
      }
    }
  } else if constexpr (TAIL == 3 || TAIL == 5) {
    __shared__ unsigned s_tot[WARPS];
    __shared__ int s_bad;
    const float q = acc * scale;
    const bool ok = __all_sync(0xffffffffu, q < 67108864.f / WARPS);
    const unsigned tot = __reduce_add_sync(0xffffffffu, ok ? __float2uint_rn(q) : 0u);
    if (lane == 0) s_tot[warp] = tot;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    if (!ok && lane == 0) s_bad = 1;                           // (measurement build: the wide path is not completed)
    if (warp != 0) return;
    const unsigned t = __reduce_add_sync(0xffffffffu, lane < WARPS ? s_tot[lane] : 0u);
    if (lane == 0) red_u64(TAIL == 5 ? mine : rec, ((unsigned long long)t << (kCountBits + kPoisonBits)) + 1ull);
  }
}

template <int TAIL, int THREADS, int MINB, int UNROLL>
static int go(StepParams& p, unsigned long long* sub, float scale, cudaStream_t st, int nsub = 64, int stride = 16) {
  constexpr long long per_cta = (long long)THREADS * kVec * UNROLL;
  dim3 grid((unsigned)((p.n + per_cta - 1) / per_cta), (unsigned)p.B);
  launch_pdl(ode_exp_kernel<TAIL, THREADS, MINB, UNROLL>, grid, THREADS, 0, st, p, sub, scale, nsub, stride);
  return (int)cudaGetLastError();
}

extern "C" __attribute__((visibility("default"))) const char* exp_variant_name(int variant) {
  switch (variant) {
    case 0: return "T0 ships: shuffles+smem+barrier, 1 RED/CTA -> record      (256 thr, 6/SM)";
    case 1: return "T4 no log-prob (floor)                                    (128 thr, 12/SM)";
    case 2: return "T5 redux+smem+barrier+redux, 1 RED/CTA -> 16 lines        (128 thr, 12/SM)";
    case 3: return "T6 shuffles+smem+barrier, 1 RED/CTA -> 16 lines           (128 thr, 12/SM)";
    case 4: return "T6 shuffles+smem+barrier, 1 RED/CTA -> 16 x 32 B          (128 thr, 12/SM)";
    case 5: return "T6 shuffles+smem+barrier, 1 RED/CTA -> 8 lines            (128 thr, 12/SM)";
    case 6: return "T6 shuffles+smem+barrier, 1 RED/CTA -> 32 lines           (128 thr, 12/SM)";
    case 7: return "T6 shuffles+smem+barrier, 1 RED/CTA -> 16 lines           (128 thr, 16/SM)";
    case 8: return "T5 redux+smem+barrier+redux, 1 RED/CTA -> 16 lines        (128 thr, 16/SM)";
    case 9: return "T6 shuffles+smem+barrier, 1 RED/CTA -> 32 lines           (64 thr, 24/SM)";
    case 10: return "T4 no log-prob (floor)                                    (64 thr, 24/SM)";
    case 11: return "T6 shuffles+smem+barrier, 1 RED/CTA -> 64 lines           (128 thr, 12/SM)";
    case 12: return "T6 shuffles+smem+barrier, 1 RED/CTA -> 16 lines           (256 thr, 6/SM)";
    case 13: return "T6 shuffles+smem+barrier, 1 RED/CTA -> 16 lines           (128 thr, 8/SM)";
    case 14: return "T6 shuffles+smem+barrier, 1 RED/CTA -> 4 lines            (128 thr, 12/SM)";
    case 15: return "PRODUCT mg::step_kernel, 128-thread half-tile shape       (128 thr, 12/SM)";
    case 16: return "PRODUCT mg::step_kernel, 256-thread shape, 8 sub-records  (256 thr, 6/SM)";
  }
  return nullptr;
}

extern "C" __attribute__((visibility("default"))) int exp_ode(int variant, const void* v, const float* x, float* out, void* rec, void* sub,
                                                              int64_t B, int64_t n, const mixgrpo_step_coefs* k, void* stream) {
  if (!v || !x || !out || !rec || !sub || !k || B <= 0 || n <= 0 || (n % kVec) != 0) return MIXGRPO_EINVAL;
  StepParams p;
  fill(p, v, x, n, nullptr, nullptr, n, nullptr, nullptr, out, n, nullptr, nullptr, nullptr, rec, B, n, k);
  p.defer = 1; p.early = 1;
  const float scale = 4294967296.f / ((float)n * k->two_var);
  unsigned long long* s = reinterpret_cast<unsigned long long*>(sub);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (variant) {
    case 0: return go<0, 256, 6, 1>(p, s, scale, st);
    case 1: return go<4, 128, 12, 1>(p, s, scale, st);
    case 2: return go<5, 128, 12, 1>(p, s, scale, st, 16, 16);
    case 3: return go<6, 128, 12, 1>(p, s, scale, st, 16, 16);
    case 4: return go<6, 128, 12, 1>(p, s, scale, st, 16, 4);
    case 5: return go<6, 128, 12, 1>(p, s, scale, st, 8, 16);
    case 6: return go<6, 128, 12, 1>(p, s, scale, st, 32, 16);
    case 7: return go<6, 128, 16, 1>(p, s, scale, st, 16, 16);
    case 8: return go<5, 128, 16, 1>(p, s, scale, st, 16, 16);
    case 9: return go<6, 64, 24, 1>(p, s, scale, st, 32, 16);
    case 10: return go<4, 64, 24, 1>(p, s, scale, st);
    case 11: return go<6, 128, 12, 1>(p, s, scale, st, 64, 16);
    case 12: return go<6, 256, 6, 1>(p, s, scale, st, 16, 16);
    case 13: return go<6, 128, 8, 1>(p, s, scale, st, 16, 16);
    case 14: return go<6, 128, 12, 1>(p, s, scale, st, 4, 16);
    case 15: g_half_ctas = 2; p.acc = s; return launch<kFlow, __nv_bfloat16, __nv_bfloat16, MIXGRPO_SRC_DETERMINISTIC, 1, true, false, true, 0, 0>(p, st);
    case 16: g_half_ctas = 0; p.acc = s; return launch<kFlow, __nv_bfloat16, __nv_bfloat16, MIXGRPO_SRC_DETERMINISTIC, 1, true, false, true, 0, 0>(p, st);
  }
  return MIXGRPO_EINVAL;
}

constexpr int kMaxSub = 256, kLine = 16;
// one CTA per (launch, sample): sums the record word and every sub-word slot (checks the variants against each other), re-zeroes them
__global__ void __launch_bounds__(kMaxSub) exp_collect_kernel(unsigned long long* sub, unsigned long long* rec, double* out) {
  pdl_prologue();
  __shared__ unsigned long long s_part[kMaxSub / 32];
  const long long r = blockIdx.x;
  unsigned long long* w = sub + (r * kMaxSub + threadIdx.x) * kLine;
  unsigned long long t = 0ull;
#pragma unroll
  for (int q = 0; q < kLine; q += 4) { t += w[q]; w[q] = 0ull; }     // 32-byte-apart slots of the line (stride 4 or 16 variants)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long tot = rec[kWsStride * r];
    rec[kWsStride * r] = 0ull;
    for (int i = 0; i < kMaxSub / 32; ++i) tot += s_part[i];
    out[r] = (double)(tot >> (kCountBits + kPoisonBits)) * (1.0 / 4294967296.0);
  }
}

extern "C" __attribute__((visibility("default"))) int64_t exp_sub_words(int64_t records) { return records * kMaxSub * kLine; }

extern "C" __attribute__((visibility("default"))) int exp_collect(void* sub, void* rec, double* out, int64_t records, void* stream) {
  launch_pdl(exp_collect_kernel, dim3((unsigned)records), dim3(kMaxSub), 0, reinterpret_cast<cudaStream_t>(stream),
             reinterpret_cast<unsigned long long*>(sub), reinterpret_cast<unsigned long long*>(rec), out);
  return (int)cudaGetLastError();
}
