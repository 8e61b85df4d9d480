// Layout helpers either side of the sampler path: FLUX 2x2 patchify of VAE latents
//   pack   (B,C,H,W) -> (B,(H/2)(W/2),4C)   TR:94-99
//   unpack inverse, optionally fused with the VAE de-normalisation x/0.3611 + 0.1159   TR:102-115, TR:287
// TR = /root/reference/fastvideo/train_grpo_flux.py.  Pure permutes (HBM-bound, 2x element size per
// element); one thread moves one 2-element row fragment so both sides see >= 8-byte accesses for fp32.
#include "common.cuh"

namespace mg {

// packed index: [b][hp][wp][c][dh][dw]   <->   unpacked: [b][c][2hp+dh][2wp+dw]
template <class T, bool PACK, bool AFFINE>
__global__ void __launch_bounds__(256) pack_kernel(const T* src, T* __restrict__ dst,   // src: maybe the last step's output (no .nc loads, see ld_dep)
                                                    long long total_pairs,
                                                  int C, int H, int W, float divisor, float shift) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;   // one (dw=0,1) pair
  if (i >= total_pairs) return;
  const int Wp = W / 2, Hp = H / 2;
  // enumerate in PACKED order so the packed side is fully coalesced
  long long r = i;
  const int dh = (int)(r % 2); r /= 2;
  const int c = (int)(r % C); r /= C;
  const int wp = (int)(r % Wp); r /= Wp;
  const int hp = (int)(r % Hp); r /= Hp;
  const long long b = r;
  const long long packed = i * 2;
  const long long unpacked = ((b * C + c) * H + (2 * hp + dh)) * (long long)W + 2 * wp;
  if constexpr (PACK) {
    dst[packed] = src[unpacked];
    dst[packed + 1] = src[unpacked + 1];
  } else {
    float a0 = (float)src[packed], a1 = (float)src[packed + 1];
    if constexpr (AFFINE) {
      a0 = __fadd_rn(__fdiv_rn(a0, divisor), shift);
      a1 = __fadd_rn(__fdiv_rn(a1, divisor), shift);
    }
    dst[unpacked] = (T)a0;
    dst[unpacked + 1] = (T)a1;
  }
}

template <bool PACK>
static int launch_pack(const void* src, void* dst, int dtype, int64_t B, int C, int H, int W, float divisor, float shift,
                       cudaStream_t st) {
  if (!src || !dst || B <= 0 || C <= 0 || H <= 0 || W <= 0 || (H % 2) || (W % 2)) return MIXGRPO_EINVAL;
  const long long pairs = (long long)B * C * H * W / 2;
  const unsigned grid = (unsigned)((pairs + 255) / 256);
  const bool affine = !PACK && !(divisor == 1.f && shift == 0.f);
  if (dtype == MIXGRPO_F32) {
    if (affine) launch_pdl(pack_kernel<float, PACK, true>, grid, 256, 0, st, (const float*)src, (float*)dst, pairs, C, H, W, divisor, shift);
    else launch_pdl(pack_kernel<float, PACK, false>, grid, 256, 0, st, (const float*)src, (float*)dst, pairs, C, H, W, divisor, shift);
  } else if (dtype == MIXGRPO_BF16) {
    if (affine) launch_pdl(pack_kernel<__nv_bfloat16, PACK, true>, grid, 256, 0, st, (const __nv_bfloat16*)src, (__nv_bfloat16*)dst, pairs, C, H, W, divisor, shift);
    else launch_pdl(pack_kernel<__nv_bfloat16, PACK, false>, grid, 256, 0, st, (const __nv_bfloat16*)src, (__nv_bfloat16*)dst, pairs, C, H, W, divisor, shift);
  } else {
    return MIXGRPO_EINVAL;
  }
  return (int)cudaGetLastError();
}

}  // namespace mg

extern "C" __attribute__((visibility("default"))) int mixgrpo_pack_latents(const void* src, void* dst, int dtype, int64_t B, int C, int H, int W, void* stream) {
  return mg::launch_pack<true>(src, dst, dtype, B, C, H, W, 1.f, 0.f, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" __attribute__((visibility("default"))) int mixgrpo_unpack_latents(const void* src, void* dst, int dtype, int64_t B, int C, int H, int W, float divisor,
                                      float shift, void* stream) {
  return mg::launch_pack<false>(src, dst, dtype, B, C, H, W, divisor, shift, reinterpret_cast<cudaStream_t>(stream));
}

namespace mg {
template <class T, bool VECTOR>
__global__ void __launch_bounds__(256) cast_rows_kernel(const T* __restrict__ src, float* __restrict__ dst, long long dst_bs, long long n) {
  pdl_prologue();
  const int b = blockIdx.y;
  const long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * (VECTOR ? kVec : 1);
  if (i >= n) return;
  if constexpr (VECTOR) {
    float r[kVec];
    ld_stream(src + (long long)b * n + i, r);
    st_stream(dst + (long long)b * dst_bs + i, r);
  } else {
    dst[(long long)b * dst_bs + i] = (float)src[(long long)b * n + i];
  }
}
}  // namespace mg

extern "C" __attribute__((visibility("default"))) int mixgrpo_cast_rows(const void* src, int src_dtype, float* dst, int64_t dst_bs, int64_t B, int64_t n,
                                                                       void* stream) {
  if (!src || !dst || B <= 0 || B > 65535 || n <= 0 || (src_dtype != MIXGRPO_F32 && src_dtype != MIXGRPO_BF16)) return MIXGRPO_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const uintptr_t s = reinterpret_cast<uintptr_t>(src), d = reinterpret_cast<uintptr_t>(dst);
  const bool vec = (n % mg::kVec == 0) && (dst_bs % mg::kVec == 0) && (d % 32 == 0) && (s % (src_dtype == MIXGRPO_BF16 ? 16 : 32) == 0);
  const long long per = 256LL * (vec ? mg::kVec : 1);
  dim3 grid((unsigned)((n + per - 1) / per), (unsigned)B);
  if (src_dtype == MIXGRPO_BF16) {
    if (vec) mg::launch_pdl(mg::cast_rows_kernel<__nv_bfloat16, true>, grid, 256, 0, st, (const __nv_bfloat16*)src, dst, dst_bs, n);
    else mg::launch_pdl(mg::cast_rows_kernel<__nv_bfloat16, false>, grid, 256, 0, st, (const __nv_bfloat16*)src, dst, dst_bs, n);
  } else {
    if (vec) mg::launch_pdl(mg::cast_rows_kernel<float, true>, grid, 256, 0, st, (const float*)src, dst, dst_bs, n);
    else mg::launch_pdl(mg::cast_rows_kernel<float, false>, grid, 256, 0, st, (const float*)src, dst, dst_bs, n);
  }
  return (int)cudaGetLastError();
}

extern "C" __attribute__((visibility("default"))) int mixgrpo_abi_version(void) { return MIXGRPO_ABI_VERSION; }

#define MG_STR2(x) #x
#define MG_STR(x) MG_STR2(x)
extern "C" __attribute__((visibility("default"))) const char* mixgrpo_build_info(void) {
  return "mixgrpo_b200 abi " MG_STR(MIXGRPO_ABI_VERSION) " sm_100a nvcc " MG_STR(__CUDACC_VER_MAJOR__) "." MG_STR(__CUDACC_VER_MINOR__) " built " __DATE__;
}

extern "C" __attribute__((visibility("default"))) const char* mixgrpo_error_string(int code) {
  switch (code) {
    case 0: return "success";
    case MIXGRPO_EINVAL: return "mixgrpo: invalid argument";
    case MIXGRPO_EALIGN: return "mixgrpo: misaligned pointer";
    case MIXGRPO_ENOSPACE: return "mixgrpo: workspace too small";
    case MIXGRPO_EUNSUPPORTED: return "mixgrpo: shape not covered by this entry point (nothing launched)";
    default: return code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "mixgrpo: unknown error";
  }
}
