// Stand-alone C++ client of the C ABI (include/mixgrpo_b200.h): no torch, no Python — cudaMalloc'ed buffers, one fused
// SDE sampler step + log-prob launch (the replacement of flow_grpo_step's rollout mode, SU:157-210), results printed as
// hex floats.  tests/test_gpu_cabi_client.py builds it with nvcc, feeds it the coefficient block the Python layer computes
// and checks the output bit-for-bit against the Python path on the same synthetic inputs.
//
//   nvcc -std=c++17 -I include examples/c_abi_demo.cpp -o build/c_abi_demo -L mixgrpo_b200/_lib -lmixgrpo_b200 -Xlinker -rpath=$PWD/mixgrpo_b200/_lib
//   build/c_abi_demo B n  two_var log_scale log_norm c0 ... c15      (19 floats as %a hex)
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mixgrpo_b200.h"

static uint16_t f2bf(float f) {   // round-to-nearest-even, like c10::BFloat16
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

// the same integer pattern tests/test_gpu_cabi_client.py builds with numpy: value in [-1, 1) with 11 bits
static float pattern(uint64_t i, uint32_t salt) { return (float)(((i * 2654435761ull + salt) % 2048ull)) / 1024.0f - 1.0f; }

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 10; } } while (0)

int main(int argc, char** argv) {
  if (argc != 3 + 19) { fprintf(stderr, "usage: %s B n <19 coefficient floats>\n", argv[0]); return 2; }
  const int64_t B = atoll(argv[1]), n = atoll(argv[2]);
  mixgrpo_step_coefs k;
  k.two_var = strtof(argv[3], nullptr); k.log_scale = strtof(argv[4], nullptr); k.log_norm = strtof(argv[5], nullptr);
  for (int i = 0; i < 16; ++i) k.c[i] = strtof(argv[6 + i], nullptr);
  const int64_t E = B * n;
  std::vector<float> hx(E);
  std::vector<uint16_t> hv(E), he(E);
  for (int64_t i = 0; i < E; ++i) { hx[i] = pattern(i, 1); hv[i] = f2bf(pattern(i, 7)); he[i] = f2bf(pattern(i, 13)); }
  float *x, *xn, *x0, *lp; uint16_t *v, *e; void* ws;
  const int64_t wsb = mixgrpo_step_workspace_bytes(B, n);
  CK(cudaMalloc(&x, E * 4)); CK(cudaMalloc(&xn, E * 4)); CK(cudaMalloc(&x0, E * 4)); CK(cudaMalloc(&lp, B * 4));
  CK(cudaMalloc(&v, E * 2)); CK(cudaMalloc(&e, E * 2)); CK(cudaMalloc(&ws, wsb)); CK(cudaMemset(ws, 0, wsb));
  CK(cudaMemcpy(x, hx.data(), E * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(v, hv.data(), E * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e, he.data(), E * 2, cudaMemcpyHostToDevice));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  const int rc = mixgrpo_flow_step(v, MIXGRPO_BF16, x, n, e, nullptr, n, xn, n, x0, nullptr, lp, ws, wsb, B, n, &k, MIXGRPO_SRC_NOISE,
                                   MIXGRPO_FLAG_ROUND_LIKE_TORCH, st, nullptr);
  if (rc != 0) { fprintf(stderr, "mixgrpo_flow_step: %d (%s)\n", rc, mixgrpo_error_string(rc)); return 3; }
  CK(cudaStreamSynchronize(st));
  std::vector<float> hlp(B), hxn(E), hx0(E);
  CK(cudaMemcpy(hlp.data(), lp, B * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hxn.data(), xn, E * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hx0.data(), x0, E * 4, cudaMemcpyDeviceToHost));
  uint64_t h1 = 1469598103934665603ull, h2 = h1;       // FNV-1a over the raw bits of x_next and x0
  for (int64_t i = 0; i < E; ++i) { uint32_t a, b; memcpy(&a, &hxn[i], 4); memcpy(&b, &hx0[i], 4); h1 = (h1 ^ a) * 1099511628211ull; h2 = (h2 ^ b) * 1099511628211ull; }
  printf("abi %d\n", mixgrpo_abi_version());
  for (int64_t b = 0; b < B; ++b) printf("logp %lld %a\n", (long long)b, hlp[b]);
  printf("x_next_fnv %016llx\nx0_fnv %016llx\n", (unsigned long long)h1, (unsigned long long)h2);
  return 0;
}
