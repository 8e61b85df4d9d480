// Design-space micro-benchmark for the fused SDE step + log-prob kernel (standalone; not product code).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo tools/ubench.cu -o build/ubench
// Each variant runs the flow SDE rollout arithmetic (bf16 v/eps, fp32 x -> fp32 x', x0, logp) over
// rotating buffer sets (> L2) from a CUDA graph and prints us/launch and GB/s at 16 B/elem.
#include <cuda_bf16.h>
#include <string.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

struct Coef { float sig, cx, cv, dt, s, two_var, log_s, log_c; };

enum Hint { H_NC_NA = 0, H_DEFAULT = 1, H_EVICT_FIRST = 2, H_CS = 3 };

__device__ __forceinline__ uint64_t ef_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
template <int HINT> __device__ __forceinline__ uint4 ld16(const void* p) {
  uint4 r;
  if constexpr (HINT == H_NC_NA) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  else if constexpr (HINT == H_EVICT_FIRST) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(ef_policy()));
  else if constexpr (HINT == H_CS) asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  else r = *reinterpret_cast<const uint4*>(p);
  return r;
}
template <int HINT> __device__ __forceinline__ uint2 ld8(const void* p) {
  uint2 r;
  if constexpr (HINT == H_NC_NA) asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  else if constexpr (HINT == H_EVICT_FIRST) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(ef_policy()));
  else if constexpr (HINT == H_CS) asm volatile("ld.global.cs.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  else r = *reinterpret_cast<const uint2*>(p);
  return r;
}
template <int HINT> __device__ __forceinline__ void ld32(const float* p, float (&r)[8]) {
  if constexpr (HINT == H_NC_NA) asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "l"(p));
  else if constexpr (HINT == H_EVICT_FIRST) {
    uint32_t u[8];
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]) : "l"(p));
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(u[i]);
  }
  else if constexpr (HINT == H_CS) asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "l"(p));
  else asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "l"(p));
}
template <int HINT> __device__ __forceinline__ void st32(float* p, const float (&r)[8]) {
  if constexpr (HINT == H_NC_NA) asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]) : "memory");
  else if constexpr (HINT == H_EVICT_FIRST) asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "r"(__float_as_uint(r[0])), "r"(__float_as_uint(r[1])), "r"(__float_as_uint(r[2])), "r"(__float_as_uint(r[3])), "r"(__float_as_uint(r[4])), "r"(__float_as_uint(r[5])), "r"(__float_as_uint(r[6])), "r"(__float_as_uint(r[7])) : "memory");
  else if constexpr (HINT == H_CS) asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]) : "memory");
  else asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]) : "memory");
}
template <int HINT> __device__ __forceinline__ void st16(float* p, float a, float b, float c, float d) {
  if constexpr (HINT == H_NC_NA) asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
  else if constexpr (HINT == H_EVICT_FIRST) asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "l"(ef_policy()) : "memory");
  else if constexpr (HINT == H_CS) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
  else *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}

__device__ __forceinline__ float lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ void rnd2(float& a, float& b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  uint32_t u = *reinterpret_cast<uint32_t*>(&t);
  a = lo(u); b = hi(u);
}

// flow SDE math on a pair of elements; returns d^2 sum
__device__ __forceinline__ float pair_math(const Coef& c, float v0, float v1, float x0i, float x1i, float e0, float e1,
                                           float& xn0, float& xn1, float& p0, float& p1) {
  float t0 = __fmul_rn(c.sig, v0), t1 = __fmul_rn(c.sig, v1); rnd2(t0, t1);
  p0 = __fsub_rn(x0i, t0); p1 = __fsub_rn(x1i, t1);
  float a0 = __fmul_rn(v0, c.cv), a1 = __fmul_rn(v1, c.cv); rnd2(a0, a1);
  a0 = __fmul_rn(a0, c.dt); a1 = __fmul_rn(a1, c.dt); rnd2(a0, a1);
  float m0 = __fadd_rn(__fmul_rn(x0i, c.cx), a0), m1 = __fadd_rn(__fmul_rn(x1i, c.cx), a1);
  float n0 = __fmul_rn(c.s, e0), n1 = __fmul_rn(c.s, e1); rnd2(n0, n1);
  xn0 = __fadd_rn(m0, n0); xn1 = __fadd_rn(m1, n1);
  float d0 = __fsub_rn(xn0, m0), d1 = __fsub_rn(xn1, m1);
  return fmaf(d0, d0, d1 * d1);
}

// ref_cuda rounding with packed bf16x2 multiplies: bf16*bf16 is exact in fp32, so one HMUL2.BF16 (RN) equals
// "widen, fp32 multiply, round to bf16".  cb = {sig, cv, dt, s} as bf16x2 broadcast pairs.
struct CoefB { uint32_t sig, cv, dt, s; };
__device__ __forceinline__ uint32_t bmul(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmul2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ float pair_math_b(const Coef& c, const CoefB& cb, uint32_t v, float x0i, float x1i, uint32_t e,
                                             float& xn0, float& xn1, float& p0, float& p1) {
  const uint32_t t = bmul(cb.sig, v);
  p0 = __fsub_rn(x0i, lo(t)); p1 = __fsub_rn(x1i, hi(t));
  const uint32_t a = bmul(bmul(v, cb.cv), cb.dt);
  const float m0 = __fadd_rn(__fmul_rn(x0i, c.cx), lo(a)), m1 = __fadd_rn(__fmul_rn(x1i, c.cx), hi(a));
  const uint32_t nz = bmul(cb.s, e);
  xn0 = __fadd_rn(m0, lo(nz)); xn1 = __fadd_rn(m1, hi(nz));
  const float d0 = __fsub_rn(xn0, m0), d1 = __fsub_rn(xn1, m1);
  return fmaf(d0, d0, d1 * d1);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct P {
  const __nv_bfloat16* v; const float* x; const __nv_bfloat16* e; float* xo; float* x0; float* logp; float* partials; unsigned* counters; unsigned long long* packed;
  long long n; int nblk; Coef c; CoefB cb;
  int pf_dist, nsamp;      // MATH == 4: CTA (linear id L) bulk-prefetches the tile of CTA L + pf_dist into L2
};

__device__ __forceinline__ void l2_prefetch_bulk(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// REDUCE: 0 none, 1 CTA partial + fence + ticket + last-CTA finalize, 2 CTA partial store only (finalize kernel separate),
//         3 float RED into logp accumulator (no ticket)
template <int VEC, int UNROLL, int HINT, int REDUCE, int BLOCK, int MINB, int MATH = 0>
__global__ void __launch_bounds__(BLOCK, MINB) k_sde(const __grid_constant__ P p) {
  constexpr int WORK = (REDUCE == 5) ? BLOCK - 32 : BLOCK;   // REDUCE 5: last warp is a dedicated reducer
  const bool worker = threadIdx.x < WORK;
  const int b = blockIdx.y;
  const long long n = p.n;
  const __nv_bfloat16* vp = p.v + (long long)b * n;
  const float* xp = p.x + (long long)b * n;
  const __nv_bfloat16* ep = p.e + (long long)b * n;
  float* xo = p.xo + (long long)b * n;
  float* x0 = p.x0 + (long long)b * n;
  float acc = 0.f;
  __shared__ unsigned long long s_fx;
  __shared__ unsigned s_cnt;
  if constexpr (REDUCE == 8) {                    // init before the loads are even issued: the barrier costs nothing here
    if (threadIdx.x == 0) { s_fx = 0ull; s_cnt = 0u; }
    __syncthreads();
  }
  if constexpr (MATH >= 2) asm volatile("griddepcontrol.launch_dependents;");   // PDL: let the next grid start filling freed SMs
  if constexpr (MATH == 2 || MATH == 4) asm volatile("griddepcontrol.wait;" ::: "memory");      // wait for the previous grid before ANY load
  if constexpr (MATH == 4) {
    // one thread pulls the inputs of the CTA that will run in this slot one wave later into L2 (3 bulk prefetches)
    if (threadIdx.x == 0 && UNROLL == 1) {
      const long long L = (long long)blockIdx.y * gridDim.x + blockIdx.x + p.pf_dist;
      const int by = (int)(L / gridDim.x), bx = (int)(L - (long long)by * gridDim.x);
      if (by < p.nsamp) {
        const long long off = (long long)bx * WORK * 8;
        const long long rem = n - off;
        if (rem > 0) {
          const unsigned cnt = (unsigned)(rem < (long long)WORK * 8 ? rem : (long long)WORK * 8);
          l2_prefetch_bulk(p.v + (long long)by * n + off, cnt * 2u);
          l2_prefetch_bulk(p.x + (long long)by * n + off, cnt * 4u);
          l2_prefetch_bulk(p.e + (long long)by * n + off, cnt * 2u);
        }
      }
    }
  }
  if constexpr (VEC == 8) {
    uint4 v[UNROLL], e[UNROLL];
    float x[UNROLL][8];
    long long idx[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      idx[u] = (((long long)blockIdx.x * UNROLL + u) * WORK + threadIdx.x) * 8;
      if (!worker) idx[u] = n;
      if (idx[u] < n) { v[u] = ld16<HINT>(vp + idx[u]); e[u] = ld16<HINT>(ep + idx[u]); }
    }
    if constexpr (MATH == 3) asm volatile("griddepcontrol.wait;" ::: "memory");    // only x depends on the previous step
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (idx[u] < n) ld32<HINT>(xp + idx[u], x[u]);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (idx[u] < n) {
        float xn[8], p0[8];
        const uint32_t vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w}, ee[4] = {e[u].x, e[u].y, e[u].z, e[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if constexpr (MATH == 1) acc += pair_math_b(p.c, p.cb, vv[j], x[u][2 * j], x[u][2 * j + 1], ee[j], xn[2 * j], xn[2 * j + 1], p0[2 * j], p0[2 * j + 1]);
          else acc += pair_math(p.c, lo(vv[j]), hi(vv[j]), x[u][2 * j], x[u][2 * j + 1], lo(ee[j]), hi(ee[j]), xn[2 * j], xn[2 * j + 1], p0[2 * j], p0[2 * j + 1]);
        }
        st32<HINT>(xo + idx[u], xn);
        st32<HINT>(x0 + idx[u], p0);
      }
    }
  } else {  // VEC == 4
    uint2 v[UNROLL], e[UNROLL];
    uint4 x[UNROLL];
    long long idx[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      idx[u] = (((long long)blockIdx.x * UNROLL + u) * WORK + threadIdx.x) * 4;
      if (!worker) idx[u] = n;
      if (idx[u] < n) { v[u] = ld8<HINT>(vp + idx[u]); x[u] = ld16<HINT>(xp + idx[u]); e[u] = ld8<HINT>(ep + idx[u]); }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (idx[u] < n) {
        float xn[4], p0[4];
        const float xf[4] = {__uint_as_float(x[u].x), __uint_as_float(x[u].y), __uint_as_float(x[u].z), __uint_as_float(x[u].w)};
        acc += pair_math(p.c, lo(v[u].x), hi(v[u].x), xf[0], xf[1], lo(e[u].x), hi(e[u].x), xn[0], xn[1], p0[0], p0[1]);
        acc += pair_math(p.c, lo(v[u].y), hi(v[u].y), xf[2], xf[3], lo(e[u].y), hi(e[u].y), xn[2], xn[3], p0[2], p0[3]);
        st16<HINT>(xo + idx[u], xn[0], xn[1], xn[2], xn[3]);
        st16<HINT>(x0 + idx[u], p0[0], p0[1], p0[2], p0[3]);
      }
    }
  }
  if constexpr (REDUCE == 0) {
    if (acc == 123.456f) p.logp[b] = acc;
    return;
  } else if constexpr (REDUCE == 7) {
    // ONE atomic carries both the arrival count and the CTA's contribution as 40-bit fixed point
    // ([sum Q8.32 | poison 12 | count 12]): integer adds commute -> deterministic, no fence, no partial reads.
    __shared__ float s_w[BLOCK / 32];
    acc = warp_sum(acc);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_w[warp] = acc;
    __syncthreads();
    if (warp != 0) return;
    float t = lane < BLOCK / 32 ? s_w[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) {
      float r = t / ((float)n * p.c.two_var);
      const float cap = 255.0f / (float)p.nblk;
      unsigned long long add = 1ull;
      if (!(r <= cap)) { add += 1ull << 12; r = (r != r) ? 0.f : cap; }
      add += __float2ull_rn(r * 4294967296.0f) << 24;
      const unsigned long long old = atomicAdd(&p.packed[b], add);
      if ((old & 0xfffull) == (unsigned long long)(p.nblk - 1)) {
        const unsigned long long tot = old + add;
        float q = (float)((double)(tot >> 24) * (1.0 / 4294967296.0));
        if ((tot >> 12) & 0xfffull) q = __int_as_float(0x7fc00000);
        p.logp[b] = -q - p.c.log_s - p.c.log_c;
        p.packed[b] = 0ull;
      }
    }
  } else if constexpr (REDUCE == 8) {
    // barrier-free: each warp adds its fixed-point share to a shared-memory word and bumps a shared counter; the last
    // warp to arrive publishes the CTA with the single packed global atomic.  No warp waits for another.
    const int lane = threadIdx.x & 31;
    acc = warp_sum(acc);
    if (lane == 0) {
      float r = acc / ((float)n * p.c.two_var);
      unsigned long long fx = (r >= 0.f && r <= 255.f) ? __float2ull_rn(r * 4294967296.0f) : (1ull << 62);
      atomicAdd(&s_fx, fx);
      __threadfence_block();
      const unsigned old = atomicAdd(&s_cnt, 1u);
      if (old == BLOCK / 32 - 1) {
        __threadfence_block();
        unsigned long long tot = s_fx;
        unsigned long long add = 1ull;
        const unsigned long long capfx = (unsigned long long)(255.0 * 4294967296.0 / p.nblk);
        if (tot > capfx) { add += 1ull << 12; tot = 0; }
        add += tot << 24;
        const unsigned long long o = atomicAdd(&p.packed[b], add);
        if ((o & 0xfffull) == (unsigned long long)(p.nblk - 1)) {
          const unsigned long long tt = o + add;
          float q = (float)((double)(tt >> 24) * (1.0 / 4294967296.0));
          if ((tt >> 12) & 0xfffull) q = __int_as_float(0x7fc00000);
          p.logp[b] = -q - p.c.log_s - p.c.log_c;
          p.packed[b] = 0ull;
        }
      }
    }
  } else if constexpr (REDUCE == 4 || REDUCE == 5 || REDUCE == 6) {
    // 4: warp 0 lingers for the ticket, warps 1.. exit right after the barrier
    // 5: a dedicated extra warp (no streaming stores of its own -> cheap fence) does the ticket
    // 6: partial store only (no ticket, no finalize): lower bound for a deferred-finalize design
    __shared__ float s_w[BLOCK / 32];
    acc = warp_sum(acc);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int RW = (REDUCE == 5) ? BLOCK / 32 - 1 : 0;     // reducer warp
    if (lane == 0) s_w[warp] = acc;
    __syncthreads();
    if (warp != RW) return;
    float t = lane < WORK / 32 ? s_w[lane] : 0.f;
    t = warp_sum(t);
    int last = 0;
    if (lane == 0) {
      p.partials[(long long)b * p.nblk + blockIdx.x] = t;
      if constexpr (REDUCE != 6) {
        __threadfence();
        const unsigned ticket = atomicAdd(&p.counters[b], 1u);
        last = (ticket == (unsigned)(p.nblk - 1));
      }
    }
    if constexpr (REDUCE == 6) return;
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
      __threadfence();
      float s2 = 0.f;
      for (int i = lane; i < p.nblk; i += 32) s2 += __ldcg(&p.partials[(long long)b * p.nblk + i]);
      s2 = warp_sum(s2);
      if (lane == 0) { p.logp[b] = -(s2 / (float)n) / p.c.two_var - p.c.log_s - p.c.log_c; p.counters[b] = 0; }
    }
  } else {
    __shared__ float s_w[BLOCK / 32];
    __shared__ int s_last;
    acc = warp_sum(acc);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_w[warp] = acc;
    __syncthreads();
    if (warp == 0) {
      float t = lane < BLOCK / 32 ? s_w[lane] : 0.f;
      t = warp_sum(t);
      if (lane == 0) {
        if constexpr (REDUCE == 3) {
          atomicAdd(&p.logp[b], -t / ((float)n * p.c.two_var));
        } else {
          p.partials[(long long)b * p.nblk + blockIdx.x] = t;
          if constexpr (REDUCE == 1) {
            __threadfence();
            const unsigned ticket = atomicAdd(&p.counters[b], 1u);
            s_last = (ticket == (unsigned)(p.nblk - 1));
          }
        }
      }
    }
    if constexpr (REDUCE == 1) {
      __syncthreads();
      if (s_last) {
        __threadfence();
        float s = 0.f;
        for (int i = threadIdx.x; i < p.nblk; i += BLOCK) s += __ldcg(&p.partials[(long long)b * p.nblk + i]);
        s = warp_sum(s);
        if (lane == 0) s_w[warp] = s;
        __syncthreads();
        if (warp == 0) {
          float t = lane < BLOCK / 32 ? s_w[lane] : 0.f;
          t = warp_sum(t);
          if (lane == 0) { p.logp[b] = -(t / (float)n) / p.c.two_var - p.c.log_s - p.c.log_c; p.counters[b] = 0; }
        }
      }
    }
  }
}

__global__ void k_finalize(const float* partials, int nblk, long long n, Coef c, float* logp) {
  const int b = blockIdx.x;
  float s = 0.f;
  for (int i = threadIdx.x; i < nblk; i += 32) s += partials[(long long)b * nblk + i];
  s = warp_sum(s);
  if (threadIdx.x == 0) logp[b] = -(s / (float)n) / c.two_var - c.log_s - c.log_c;
}

// plain copy with the same traffic (read 8 B, write 8 B per element): ceiling for this size
template <int HINT> __global__ void __launch_bounds__(256) k_copy(const float* a, const float* b2, float* c, float* d, long long n) {
  const long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 8;
  if (i < n) { float r[8], s[8]; ld32<HINT>(a + i, r); ld32<HINT>(b2 + i, s); st32<HINT>(c + i, r); st32<HINT>(d + i, s); }
}


// ---------------------------------------------------------------------------------------------------------
// Prototype: persistent CTA per SM, inputs staged by 1-D bulk async copies (cp.async.bulk -> UBLKCP) into a deep
// shared-memory ring guarded by mbarriers, so (nearly) all of an SM's share of the input is in flight from t=0.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" :: "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int STAGES, int CT /*consumer threads*/, int MATH>
__global__ void __launch_bounds__(CT + 32, 1) k_sde_tma(const __grid_constant__ P p, int tiles_per_sample, int total_tiles) {
  constexpr int TILE = 2048;
  constexpr int EPT = TILE / CT;                 // elements per consumer thread per tile (4 for 512 threads)
  static_assert(EPT == 4 || EPT == 8, "");
  constexpr int STAGE_BYTES = TILE * 2 + TILE * 4 + TILE * 2;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + STAGES;
  unsigned long long* s_acc = reinterpret_cast<unsigned long long*>(empty + STAGES);
  unsigned* s_arr = reinterpret_cast<unsigned*>(s_acc + STAGES + 1);
  unsigned char* ring = smem + 1024;
  const int tid = threadIdx.x;
  const long long n = p.n;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CT / 32); }
    for (int s = 0; s <= STAGES; ++s) { s_acc[s] = 0ull; s_arr[s] = 0u; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // contiguous balanced share of the tiles
  const int G = gridDim.x, c = blockIdx.x;
  const int q = total_tiles / G, r = total_tiles % G;
  const int t0 = c * q + min(c, r), t1 = t0 + q + (c < r ? 1 : 0);

  if (tid >= CT) {                               // ---------------- producer warp
    if (tid == CT) {
      for (int t = t0, i = 0; t < t1; ++t, ++i) {
        const int s = i % STAGES;
        if (i >= STAGES) mbar_wait(&empty[s], ((i / STAGES) - 1) & 1);
        const int b = t / tiles_per_sample;
        const long long off = (long long)(t - b * tiles_per_sample) * TILE;
        const long long rem = n - off;
        const uint32_t cnt = (uint32_t)(rem < TILE ? rem : TILE);
        unsigned char* st = ring + (size_t)s * STAGE_BYTES;
        mbar_expect_tx(&full[s], cnt * 8);
        bulk_g2s(st, p.v + (long long)b * n + off, cnt * 2, &full[s]);
        bulk_g2s(st + TILE * 2, p.x + (long long)b * n + off, cnt * 4, &full[s]);
        bulk_g2s(st + TILE * 6, p.e + (long long)b * n + off, cnt * 2, &full[s]);
      }
    }
    return;
  }
  // ---------------- consumers
  const int lane = tid & 31;
  float run_acc = 0.f;
  int run_len = 0, run_id = 0;
  for (int t = t0, i = 0; t < t1; ++t, ++i) {
    const int s = i % STAGES;
    const int b = t / tiles_per_sample;
    const long long off = (long long)(t - b * tiles_per_sample) * TILE;
    mbar_wait(&full[s], (i / STAGES) & 1);
    const unsigned char* st = ring + (size_t)s * STAGE_BYTES;
    float acc = 0.f;
    const bool valid = off + (long long)tid * EPT < n;
    if constexpr (EPT == 4) {
      const uint2 v = *reinterpret_cast<const uint2*>(st + tid * 8);
      const float4 x = *reinterpret_cast<const float4*>(st + TILE * 2 + tid * 16);
      const uint2 e = *reinterpret_cast<const uint2*>(st + TILE * 6 + tid * 8);
      float xn[4], p0[4];
      if constexpr (MATH == 1) {
        acc += pair_math_b(p.c, p.cb, v.x, x.x, x.y, e.x, xn[0], xn[1], p0[0], p0[1]);
        acc += pair_math_b(p.c, p.cb, v.y, x.z, x.w, e.y, xn[2], xn[3], p0[2], p0[3]);
      } else {
        acc += pair_math(p.c, lo(v.x), hi(v.x), x.x, x.y, lo(e.x), hi(e.x), xn[0], xn[1], p0[0], p0[1]);
        acc += pair_math(p.c, lo(v.y), hi(v.y), x.z, x.w, lo(e.y), hi(e.y), xn[2], xn[3], p0[2], p0[3]);
      }
      if (valid) {
        st16<H_NC_NA>(p.xo + (long long)b * n + off + tid * 4, xn[0], xn[1], xn[2], xn[3]);
        st16<H_NC_NA>(p.x0 + (long long)b * n + off + tid * 4, p0[0], p0[1], p0[2], p0[3]);
      }
    } else {
      const uint4 v = *reinterpret_cast<const uint4*>(st + tid * 16);
      const float4 xa = *reinterpret_cast<const float4*>(st + TILE * 2 + tid * 32);
      const float4 xb = *reinterpret_cast<const float4*>(st + TILE * 2 + tid * 32 + 16);
      const uint4 e = *reinterpret_cast<const uint4*>(st + TILE * 6 + tid * 16);
      float xn[8], p0[8];
      acc += pair_math(p.c, lo(v.x), hi(v.x), xa.x, xa.y, lo(e.x), hi(e.x), xn[0], xn[1], p0[0], p0[1]);
      acc += pair_math(p.c, lo(v.y), hi(v.y), xa.z, xa.w, lo(e.y), hi(e.y), xn[2], xn[3], p0[2], p0[3]);
      acc += pair_math(p.c, lo(v.z), hi(v.z), xb.x, xb.y, lo(e.z), hi(e.z), xn[4], xn[5], p0[4], p0[5]);
      acc += pair_math(p.c, lo(v.w), hi(v.w), xb.z, xb.w, lo(e.w), hi(e.w), xn[6], xn[7], p0[6], p0[7]);
      if (valid) { st32<H_NC_NA>(p.xo + (long long)b * n + off + tid * 8, xn); st32<H_NC_NA>(p.x0 + (long long)b * n + off + tid * 8, p0); }
    }
    if (valid) run_acc += acc;
    run_len += 1;
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);                         // stage can be refilled
    const bool run_ends = (t + 1 == t1) || ((t + 1) / tiles_per_sample != b);
    if (run_ends) {
      // deterministic CTA reduction for this run of tiles without a barrier: every warp adds its fixed-point
      // share to a shared-memory slot; the last warp to arrive publishes the run with ONE global atomic.
      const float ws = warp_sum(run_acc);
      if (lane == 0) {
        const int slot = run_id % (STAGES + 1);
        float rr = ws / ((float)n * p.c.two_var);
        unsigned long long fx = (rr >= 0.f && rr <= 255.f) ? __float2ull_rn(rr * 4294967296.0f) : (1ull << 62);
        atomicAdd(&s_acc[slot], fx);
        __threadfence_block();
        const unsigned old = atomicAdd(&s_arr[slot], 1u);
        if (old == CT / 32 - 1) {
          __threadfence_block();
          unsigned long long tot = atomicExch(&s_acc[slot], 0ull);
          s_arr[slot] = 0u;
          unsigned long long add = (unsigned long long)run_len;
          const unsigned long long capfx = (unsigned long long)(255.0 * 4294967296.0 * run_len / tiles_per_sample);
          if (tot > capfx) { add += 1ull << 12; tot = 0; }
          add += tot << 24;
          const unsigned long long o = atomicAdd(&p.packed[b], add);
          if ((o & 0xfffull) + run_len == (unsigned long long)tiles_per_sample) {
            const unsigned long long tt = o + add;
            float qv = (float)((double)(tt >> 24) * (1.0 / 4294967296.0));
            if ((tt >> 12) & 0xfffull) qv = __int_as_float(0x7fc00000);
            p.logp[b] = -qv - p.c.log_s - p.c.log_c;
            p.packed[b] = 0ull;
          }
        }
      }
      run_acc = 0.f; run_len = 0; run_id += 1;
    }
  }
}

struct Bufs { __nv_bfloat16 *v, *e; float *x, *xo, *x0; };

template <class F> static float time_graph(F launch, int nsets, int reps) {
  cudaStream_t st; CK(cudaStreamCreate(&st));
  launch(0, st); CK(cudaStreamSynchronize(st));
  cudaGraph_t g; cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
  for (int i = 0; i < nsets; ++i) launch(i, st);
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  for (int i = 0; i < 3; ++i) CK(cudaGraphLaunch(ge, st));
  CK(cudaStreamSynchronize(st));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  CK(cudaEventRecord(a, st));
  for (int i = 0; i < reps; ++i) CK(cudaGraphLaunch(ge, st));
  CK(cudaEventRecord(b, st));
  CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g)); CK(cudaStreamDestroy(st));
  return ms * 1e3f / (reps * nsets);
}

static int B = 12, S = 4096, NS = 10, REPS = 30;
static long long n;
static std::vector<Bufs> bufs;
static float *d_logp, *d_partials; static unsigned* d_counters; static unsigned long long* d_packed;
static int g_pf_dist = 888;
static Coef coef = {0.8125f, 0.9937f, 1.0234f, -0.0262f, 0.1367f, 0.0372f, -1.99f, 0.9189f};
static CoefB coefb = {0x3f503f50u, 0x3f833f83u, 0xbcd7bcd7u, 0x3e0c3e0cu};

template <int STAGES, int CT, int MATH>
static void run_tma(const char* name, int grid_cap) {
  constexpr int STAGE_BYTES = 2048 * 8;
  const int tps = (int)((n + 2047) / 2048);
  const int total = tps * B;
  const size_t smem = 1024 + (size_t)STAGES * STAGE_BYTES;
  auto kern = k_sde_tma<STAGES, CT, MATH>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
  int grid = total < grid_cap ? total : grid_cap;
  auto launch = [&](int i, cudaStream_t st) {
    P p{bufs[i].v, bufs[i].x, bufs[i].e, bufs[i].xo, bufs[i].x0, d_logp, d_partials, d_counters, d_packed, n, tps, coef, coefb};
    kern<<<grid, CT + 32, smem, st>>>(p, tps, total);
  };
  const float us = time_graph(launch, NS, REPS);
  CK(cudaDeviceSynchronize());
  const double gbs = (double)B * n * 16 / us / 1e3;
  printf("%-34s STAGES=%d CT=%d MATH=%d grid=%d regs=%3d smem=%zuK  %7.2f us  %7.1f GB/s  %.3f\n", name, STAGES, CT, MATH, grid, fa.numRegs, smem / 1024, us, gbs, gbs / 6533.5);
  fflush(stdout);
}


template <int VEC, int UNROLL, int HINT, int REDUCE, int BLOCK, int MINB, int MATH = 0>
static void run(const char* name) {
  const long long per = (long long)(REDUCE == 5 ? BLOCK - 32 : BLOCK) * VEC * UNROLL;
  const int nblk = (int)((n + per - 1) / per);
  auto kern = k_sde<VEC, UNROLL, HINT, REDUCE, BLOCK, MINB, MATH>;
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
  int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BLOCK, 0));
  auto launch = [&](int i, cudaStream_t st) {
    P p{bufs[i].v, bufs[i].x, bufs[i].e, bufs[i].xo, bufs[i].x0, d_logp, d_partials, d_counters, d_packed, n, nblk, coef, coefb, g_pf_dist, B};
    if (REDUCE == 3) cudaMemsetAsync(d_logp, 0, B * sizeof(float), st);
    if (MATH >= 2) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(nblk, B); cfg.blockDim = dim3(BLOCK); cfg.dynamicSmemBytes = 0; cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      CK(cudaLaunchKernelEx(&cfg, kern, p));
    } else {
      kern<<<dim3(nblk, B), BLOCK, 0, st>>>(p);
    }
    if (REDUCE == 2) k_finalize<<<B, 32, 0, st>>>(d_partials, nblk, n, coef, d_logp);
  };
  const float us = time_graph(launch, NS, REPS);
  const double gbs = (double)B * n * 16 / us / 1e3;
  printf("%-34s VEC=%d UNR=%d HINT=%d RED=%d BLK=%d regs=%3d occ=%d thr/SM=%4d  %7.2f us  %7.1f GB/s  %.3f\n", name, VEC, UNROLL, HINT,
         REDUCE, BLOCK, fa.numRegs, occ, occ * BLOCK, us, gbs, gbs / 6533.5);
  fflush(stdout);
}

int main(int argc, char** argv) {
  if (argc > 1) B = atoi(argv[1]);
  if (argc > 2) S = atoi(argv[2]);
  n = (long long)S * 64;
  const long long E = (long long)B * n;
  bufs.resize(NS);
  std::vector<float> h(E);
  for (long long i = 0; i < E; ++i) h[i] = (float)((i * 2654435761u) % 2001) / 1000.f - 1.f;
  std::vector<__nv_bfloat16> hb(E);
  for (long long i = 0; i < E; ++i) hb[i] = __float2bfloat16(h[i]);
  for (auto& bf : bufs) {
    CK(cudaMalloc(&bf.v, E * 2)); CK(cudaMalloc(&bf.e, E * 2)); CK(cudaMalloc(&bf.x, E * 4)); CK(cudaMalloc(&bf.xo, E * 4)); CK(cudaMalloc(&bf.x0, E * 4));
    CK(cudaMemcpy(bf.v, hb.data(), E * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(bf.e, hb.data(), E * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(bf.x, h.data(), E * 4, cudaMemcpyHostToDevice));
  }
  CK(cudaMalloc(&d_logp, B * 4)); CK(cudaMalloc(&d_partials, (size_t)B * 65536 * 4)); CK(cudaMalloc(&d_counters, B * 4));
  CK(cudaMemset(d_counters, 0, B * 4));
  CK(cudaMalloc(&d_packed, B * 8)); CK(cudaMemset(d_packed, 0, B * 8));
  printf("B=%d S=%d E=%lld  bytes/launch=%.1f MB  ideal@6533GB/s=%.2f us\n", B, S, E, E * 16 / 1e6, E * 16 / 6533.5e3);
  {
    auto launch = [&](int i, cudaStream_t st) { k_copy<H_NC_NA><<<(unsigned)((E / 8 + 255) / 256), 256, 0, st>>>(bufs[i].x, bufs[(i + 1) % NS].x, bufs[i].xo, bufs[i].x0, E); };
    float us = time_graph(launch, NS, REPS);
    printf("%-34s %7.2f us  %7.1f GB/s  %.3f\n", "copy2x(8B rd + 8B wr)/elem", us, E * 16 / us / 1e3, E * 16 / us / 1e3 / 6533.5);
    auto launch2 = [&](int i, cudaStream_t st) { k_copy<H_DEFAULT><<<(unsigned)((E / 8 + 255) / 256), 256, 0, st>>>(bufs[i].x, bufs[(i + 1) % NS].x, bufs[i].xo, bufs[i].x0, E); };
    us = time_graph(launch2, NS, REPS);
    printf("%-34s %7.2f us  %7.1f GB/s  %.3f\n", "copy2x default hints", us, E * 16 / us / 1e3, E * 16 / us / 1e3 / 6533.5);
  }
  if (argc > 3 && !strcmp(argv[3], "pf")) {
    // L2 prefetch of the NEXT wave's inputs (cp.async.bulk.prefetch.L2), distance in CTAs
    run<8, 1, H_NC_NA, 7, 256, 6, 2>("packed + PDL (shipped)");
    for (int d : {888, 740, 592, 444, 296, 148, 1036, 1184}) {
      g_pf_dist = d;
      char nm[64]; snprintf(nm, sizeof nm, "+ L2 bulk prefetch, dist %d", d);
      run<8, 1, H_NC_NA, 7, 256, 6, 4>(nm);
    }
    g_pf_dist = 888;
    run<8, 1, H_NC_NA, 0, 256, 6, 4>("prefetch 888, no reduction");
    run<8, 1, H_NC_NA, 7, 192, 8, 4>("prefetch 888, b192 c8");
    return 0;
  }
  if (argc > 3 && !strcmp(argv[3], "wave")) {
    // one-wave shapes at (12,4096,64): every load of the launch is issued before the first CTA retires
    run<8, 1, H_NC_NA, 7, 256, 6, 2>("packed + PDL b256 c6 u1 (1.73 waves)");
    run<8, 2, H_NC_NA, 7, 224, 6, 2>("u2 b224 c6 (888 CTAs, 1.00 waves)");
    run<8, 2, H_NC_NA, 7, 256, 6, 2>("u2 b256 c6 (768 CTAs)");
    run<8, 2, H_NC_NA, 7, 256, 5, 2>("u2 b256 c5 (768 CTAs, 1.04 waves)");
    run<8, 2, H_NC_NA, 7, 288, 5, 2>("u2 b288 c5 (684 CTAs)");
    run<8, 2, H_NC_NA, 7, 352, 4, 2>("u2 b352 c4 (564 CTAs)");
    run<8, 2, H_NC_NA, 7, 512, 3, 2>("u2 b512 c3 (384 CTAs)");
    run<8, 3, H_NC_NA, 7, 256, 4, 2>("u3 b256 c4 (516 CTAs)");
    run<8, 3, H_NC_NA, 7, 384, 3, 2>("u3 b384 c3 (348 CTAs)");
    run<8, 4, H_NC_NA, 7, 256, 3, 2>("u4 b256 c3 (384 CTAs)");
    run<8, 4, H_NC_NA, 7, 128, 6, 2>("u4 b128 c6 (768 CTAs)");
    run<8, 4, H_NC_NA, 7, 128, 7, 2>("u4 b128 c7 (768 CTAs)");
    run<8, 2, H_NC_NA, 0, 224, 6, 2>("u2 b224 c6 no reduction");
    return 0;
  }
  run<8, 1, H_NC_NA, 7, 256, 6>("v8 u1 packed atomic (current)");
  run<8, 1, H_NC_NA, 7, 256, 6, 2>("packed + PDL b256 c6 (1.73 waves)");
  run<8, 1, H_NC_NA, 7, 448, 3, 2>("packed + PDL b448 c3 (2.00 waves)");
  run<8, 1, H_NC_NA, 7, 448, 4, 2>("packed + PDL b448 c4");
  run<8, 1, H_NC_NA, 7, 192, 8, 2>("packed + PDL b192 c8");
  run<8, 1, H_NC_NA, 7, 320, 5, 2>("packed + PDL b320 c5");
  run<8, 1, H_NC_NA, 7, 384, 4, 2>("packed + PDL b384 c4");
  run<8, 1, H_NC_NA, 7, 224, 7, 2>("packed + PDL b224 c7");
  run<8, 1, H_NC_NA, 7, 160, 9, 2>("packed + PDL b160 c9");
  run<8, 1, H_NC_NA, 7, 96, 16, 2>("packed + PDL b96 c16");
  run<8, 1, H_NC_NA, 7, 64, 24, 2>("packed + PDL b64 c24");
  return 0;
}
