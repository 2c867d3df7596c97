// CUDA-core implementation of the GemmDesc contract (gemm.h).  Used (a) as the on-device checker for the
// tcgen05 kernel in tests, (b) for shapes too small to be worth a tensor-core tile (classifier 256->11).
// fp32 FMA over (hi [+ lo]) bf16 planes; same fused epilogue as the tcgen05 kernel.
#include "common.cuh"
#include "gemm.h"
#include "gemm_epilogue.cuh"

namespace dz {

static constexpr int SM_ROWS = 64;
static constexpr int SM_COLS = 128;
static constexpr int SM_KT = 16;

__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmDesc d) {
  __shared__ float As[SM_KT][SM_ROWS + 1];
  __shared__ float Bs[SM_KT][SM_COLS + 1];
  const int tid = threadIdx.x;
  const int r = tid & 63;
  const int cc = tid >> 6;  // 0..3 -> 32-column chunk
  const int m0 = blockIdx.x * SM_ROWS;
  const int g = (d.groups > 1) ? (int)blockIdx.y : 0;
  const int n0 = (d.groups > 1) ? 0 : (int)blockIdx.y * SM_COLS;
  const int b = blockIdx.z;
  const bool two = d.npass > 1;

  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;

  const __nv_bfloat16* abase = (const __nv_bfloat16*)d.a + (long long)b * d.a_bstride + (long long)g * d.a_gstride;
  const __nv_bfloat16* bbase = (const __nv_bfloat16*)d.b + (long long)g * d.b_gstride;

  const bool conv2d = d.conv_runs > 0;
  const int krun = conv2d ? ((d.conv_run_len + 63) / 64) * 64 : 0;   // padded run length in the B layout
  const int Ktot = conv2d ? d.conv_runs * krun : d.K;
  const int img = conv2d ? b / d.conv_Ho : 0, ho = conv2d ? b - img * d.conv_Ho : 0;
  for (int k0 = 0; k0 < Ktot; k0 += SM_KT) {
    // A tile: 64 rows x 16 k -> 1024 elements, 4 per thread
    for (int e = tid; e < SM_ROWS * SM_KT; e += 256) {
      const int kk = e % SM_KT, rr = e / SM_KT;
      const int k = k0 + kk, m = m0 + rr;
      float v = 0.f;
      if (conv2d) {
        const int run = k / krun, kk = k - run * krun;
        const int hin = d.conv_hs * ho + d.conv_h0 + run;
        if (k < Ktot && m < d.M && kk < d.conv_run_len && hin >= 0 && hin < d.conv_H) {
          const __nv_bfloat16* ab = (const __nv_bfloat16*)d.a + (long long)img * d.a_bstride + (long long)hin * d.a_hstride +
                                    (long long)m * d.a_rstride + d.conv_x0 + kk;
          v = from16(*ab, d.fp16);
          if (two) v += from16(ab[d.a_plane], d.fp16);
        }
      } else if (k < d.K && m < d.M) {
        const long long off = (long long)m * d.a_rstride + (long long)(k / d.a_kinner) * d.a_kouter + (k % d.a_kinner);
        v = from16(abase[off], d.fp16);
        if (two) v += from16(abase[d.a_plane + off], d.fp16);
      }
      As[kk][rr] = v;
    }
    for (int e = tid; e < SM_COLS * SM_KT; e += 256) {
      const int kk = e % SM_KT, nn = e / SM_KT;
      const int k = k0 + kk, n = n0 + nn;
      float v = 0.f;
      if (k < Ktot && n < d.N) {
        const long long off = (long long)n * d.ldb + k;
        v = from16(bbase[off], d.fp16);
        if (two) v += from16(bbase[d.b_plane + off], d.fp16);
      }
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SM_KT; ++kk) {
      const float a = As[kk][r];
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = fmaf(a, Bs[kk][cc * 32 + j], acc[j]);
    }
    __syncthreads();
  }
  const int m = m0 + r;
  const int nc = n0 + cc * 32;
  if (m < d.M && nc < max(d.N, d.zero_pad_to)) gemm_epilogue_chunk(d, b, g, m, nc, acc);
}

cudaError_t gemm_simt_launch(const GemmDesc& d, cudaStream_t st) {
  if (d.ln_gamma != nullptr) return cudaErrorNotSupported;   // the fused row LayerNorm exists in the tcgen05 epilogue only
  dim3 grid((d.M + SM_ROWS - 1) / SM_ROWS, d.groups > 1 ? d.groups : (max(d.N, d.zero_pad_to) + SM_COLS - 1) / SM_COLS,
            d.batches);
  gemm_simt_kernel<<<grid, 256, 0, st>>>(d);
  return cudaGetLastError();
}

}  // namespace dz
