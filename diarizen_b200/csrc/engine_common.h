// Pieces shared by the segmentation and embedding engines: device buffers, GEMM weight packing
// (fp32 -> 16-bit hi/lo planes, rows padded to 16 bytes), host-side rounding helpers.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <functional>
#include <string>
#include <vector>

namespace dz {

typedef __nv_bfloat16 bf16;
static inline int rup(int x, int m) { return (x + m - 1) / m * m; }

struct DevMem {
  void* p = nullptr;
  size_t bytes = 0;
  DevMem() = default;
  DevMem(const DevMem&) = delete;
  DevMem& operator=(const DevMem&) = delete;
  DevMem(DevMem&& o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; o.bytes = 0; }
  DevMem& operator=(DevMem&& o) noexcept {
    if (this != &o) { release(); p = o.p; bytes = o.bytes; o.p = nullptr; o.bytes = 0; }
    return *this;
  }
  cudaError_t alloc(size_t n, bool zero) {
    release();
    cudaError_t e = cudaMalloc(&p, n ? n : 16);
    if (e != cudaSuccess) { p = nullptr; return e; }
    bytes = n;
    return zero ? cudaMemset(p, 0, n ? n : 16) : cudaSuccess;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  ~DevMem() { release(); }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Weight {  // GEMM B operand, bf16 planes [2][groups][N][ldb]
  DevMem w, bias;
  int N = 0, K = 0, ldb = 0, groups = 1;
  long long plane = 0, gstride = 0;
};

struct Planes {  // activation planes view
  bf16* p = nullptr;
  long long plane = 0;
};

typedef std::function<cudaError_t(cudaStream_t)> StepFn;
struct Step { std::string name; StepFn fn; double flops = 0.0; double bytes = 0.0; };

inline cudaError_t upload(DevMem& m, const float* h, size_t n) {
  cudaError_t e = m.alloc(n * sizeof(float), false);
  if (e != cudaSuccess) return e;
  return cudaMemcpy(m.p, h, n * sizeof(float), cudaMemcpyHostToDevice);
}
inline cudaError_t upload_vec(DevMem& m, const std::vector<float>& v) { return upload(m, v.data(), v.size()); }

inline uint16_t f2bf(float x) {  // round-to-nearest-even, matches __float2bfloat16_rn for finite x
  uint32_t u;
  memcpy(&u, &x, 4);
  const uint32_t r = u + 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(r >> 16);
}
inline float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float x;
  memcpy(&x, &u, 4);
  return x;
}

inline uint16_t f2h(float x) {  // fp32 -> IEEE half, round-to-nearest-even, saturating
  if (x > 65504.f) x = 65504.f;
  if (x < -65504.f) x = -65504.f;
  __half h = __float2half_rn(x);  // host-callable conversion (cuda_fp16.h)
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}
inline float h2f(uint16_t u) {
  __half h;
  memcpy(&h, &u, 2);
  return __half2float(h);
}
inline int& g_weight_fp16() { static int v = 0; return v; }  // set by the engines while they build weights

// w: [groups][N][K] fp32 row-major -> device planes [2][groups][N][ldb]; bias (may be null) -> fp32 padded.
inline cudaError_t make_weight(Weight& W, const float* w, int groups, int N, int K, const float* bias, int nbias) {
  W.N = N; W.K = K; W.groups = groups;
  W.ldb = rup(K, 8);
  W.gstride = (long long)N * W.ldb;
  W.plane = W.gstride * groups;
  std::vector<uint16_t> h((size_t)W.plane * 2, 0);
  for (int g = 0; g < groups; ++g)
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) {
        const float x = w[((size_t)g * N + n) * K + k];
        const uint16_t hi = g_weight_fp16() ? f2h(x) : f2bf(x);
        const uint16_t lo = g_weight_fp16() ? f2h(x - h2f(hi)) : f2bf(x - bf2f(hi));
        const size_t o = (size_t)g * W.gstride + (size_t)n * W.ldb + k;
        h[o] = hi;
        h[(size_t)W.plane + o] = lo;
      }
  cudaError_t e = W.w.alloc(h.size() * 2, false);
  if (e != cudaSuccess) return e;
  e = cudaMemcpy(W.w.p, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return e;
  if (bias != nullptr) {
    std::vector<float> bb((size_t)rup(nbias, 64) + 64, 0.f);
    for (int i = 0; i < nbias; ++i) bb[i] = bias[i];
    e = upload_vec(W.bias, bb);
  }
  return e;
}


}  // namespace dz
