// Argument blocks + launch wrappers of the embedding-path kernels (emb_kernels.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dz {

struct FbankArgs {
  const float* wav; int N; int F;
  const float2* twiddle;    // [256] exp(-2 pi i k / 512)
  const float* window;      // [400] hamming
  const float* mel_w;       // [80][257]
  const int* mel_range;     // [80][2] first / one-past-last non-zero FFT bin
  float* out;               // [B][80][F] log-mel, mel-major (before mean subtraction)
};
struct Conv1Args {
  const float* fb; const float* mean; int B; int F;
  const float* w;           // [32][9]
  const float* scale; const float* shift;   // folded BatchNorm
  __nv_bfloat16* out; long long out_plane; int planes; int fp16;   // [B][80][F+2][32]
};
struct PoolArgs {
  const __nv_bfloat16* x; long long x_plane; int planes; int fp16;
  int H; int W; int C;
  const float* masks; int S; int T;
  const int* widx;          // [W] nearest-interpolation source frame
  __nv_bfloat16* out; long long out_plane; int ldo;
};

cudaError_t launch_fbank(const FbankArgs& a, int B, cudaStream_t st);
cudaError_t launch_fbank_mean(const float* fb, int B, int F, float* mean, cudaStream_t st);
cudaError_t launch_emb_conv1(const Conv1Args& a, cudaStream_t st);
cudaError_t launch_stats_pool(const PoolArgs& a, int B, cudaStream_t st);
// 3x3 / stride 1 / pad 1, C -> C channels with C = 32 or 64 (conv3x3_c32.cu).  in/out/res: zero-bordered NHWC planes
// [B][H][W+2][C]; w: the GEMM B matrix of the same layer, [C][ldw] with element (kh, kw, ci) at kh*rup(3C,64) + kw*C + ci.
struct Conv3Args {
  const __nv_bfloat16* in; __nv_bfloat16* out; const __nv_bfloat16* res;
  const __nv_bfloat16* w; int ldw; const float* bias;
  int B; int H; int W; int relu; int fp16; int C;
};
struct Conv3Plan;
Conv3Plan* conv3x3_c32_plan_create(const Conv3Args& a);
void conv3x3_c32_plan_destroy(Conv3Plan* p);
cudaError_t conv3x3_c32_plan_launch(const Conv3Plan* p, cudaStream_t st);
// same contract for C = 128 (layer3) with streamed weights and two output rows per tile (conv3x3_c128.cu)
struct ConvSPlan;
ConvSPlan* conv3x3_stream_plan_create(const Conv3Args& a);
void conv3x3_stream_plan_destroy(ConvSPlan* p);
cudaError_t conv3x3_stream_plan_launch(const ConvSPlan* p, cudaStream_t st);
cudaError_t launch_zero_borders(__nv_bfloat16* p, long long plane, int planes, long long rows, int W, int C, cudaStream_t st);

}  // namespace dz
