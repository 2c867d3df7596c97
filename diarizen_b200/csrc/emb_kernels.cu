// CUDA-core kernels of the speaker-embedding path: Kaldi-compatible log-mel filterbank (framing, DC removal,
// pre-emphasis, Hamming window, 512-point FFT, mel bins, log), cepstral mean subtraction fused into the first
// 3x3 convolution (C_in = 1 stencil) and the masked statistics pooling (4 speaker masks per trunk output).
// reference: pyannote-audio/pyannote/audio/models/embedding/wespeaker/__init__.py:80-103 (compute_fbank),
//            wespeaker/resnet.py:358 (conv1 + bn1 + relu), resnet.py:49-66 + models/blocks/pooling.py:44-131 (TSTP).
#include "common.cuh"
#include "emb_kernels.h"

namespace dz {

// ------------------------------------------------------------------------------------------------
// K13 fbank: one warp per frame.  torchaudio.compliance.kaldi.fbank(num_mel_bins=80, frame 25 ms / 10 ms,
// dither 0, hamming, remove_dc_offset, preemphasis 0.97, round_to_power_of_two, use_power, use_log_fbank),
// applied to waveform * 2^15.
// ------------------------------------------------------------------------------------------------
static constexpr int FB_WIN = 400, FB_HOP = 160, FB_NFFT = 512, FB_BINS = 257;

__global__ void __launch_bounds__(256) fbank_kernel(FbankArgs a) {
  extern __shared__ float2 smf[];
  float2* tw = smf;                       // 256 twiddles
  float2* bufs = smf + 256;               // 8 x 512 complex
  __shared__ float win[FB_WIN];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) tw[i] = a.twiddle[i];
  for (int i = threadIdx.x; i < FB_WIN; i += blockDim.x) win[i] = a.window[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.x * 8 + warp;
  const int b = blockIdx.y;
  if (f >= a.F) return;
  float2* X = bufs + warp * FB_NFFT;
  float* xr = reinterpret_cast<float*>(X);   // scratch for the raw frame (first 400 floats)
  const float* src = a.wav + (long long)b * a.N + (long long)f * FB_HOP;
  float s = 0.f;
  for (int i = lane; i < FB_WIN; i += 32) {
    const float v = src[i] * 32768.0f;
    xr[i] = v;
    s += v;
  }
  const float mean = warp_sum(s) / FB_WIN;
  __syncwarp();
  float y[13];
#pragma unroll
  for (int j = 0; j < 13; ++j) {
    const int i = lane + 32 * j;
    float v = 0.f;
    if (i < FB_WIN) {
      const float cur = xr[i] - mean;
      const float prev = xr[i > 0 ? i - 1 : 0] - mean;
      v = (cur - 0.97f * prev) * win[i];
    }
    y[j] = v;
  }
  __syncwarp();
  // bit-reversed load of the (zero padded) real frame
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int i = lane + 32 * j;
    const float v = (j < 13) ? y[j < 13 ? j : 0] : 0.f;
    const int r = __brev((unsigned)i) >> 23;  // 9-bit reversal
    X[r] = make_float2((i < FB_WIN) ? v : 0.f, 0.f);
  }
  __syncwarp();
  // radix-2 decimation-in-time, 9 stages, 256 butterflies per stage, 8 per lane
#pragma unroll 1
  for (int st = 0; st < 9; ++st) {
    const int half = 1 << st;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int t = lane + 32 * j;           // butterfly index 0..255
      const int k = t & (half - 1);
      const int i0 = ((t >> st) << (st + 1)) + k;
      const int i1 = i0 + half;
      const float2 w = tw[k << (8 - st)];
      const float2 u = X[i0], v = X[i1];
      const float2 wv = make_float2(w.x * v.x - w.y * v.y, w.x * v.y + w.y * v.x);
      X[i0] = make_float2(u.x + wv.x, u.y + wv.y);
      X[i1] = make_float2(u.x - wv.x, u.y - wv.y);
    }
    __syncwarp();
  }
  // power spectrum into the (now free) upper half of the buffer is not possible in place: use registers
  float pw[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    const int k = lane + 32 * j;
    float p = 0.f;
    if (k < FB_BINS) { const float2 c = X[k]; p = c.x * c.x + c.y * c.y; }
    pw[j] = p;
  }
  __syncwarp();
  float* P = reinterpret_cast<float*>(X);
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    const int k = lane + 32 * j;
    if (k < FB_BINS) P[k] = pw[j];
  }
  __syncwarp();
  // mel-major output [b][mel][frame]: the consumers (mean over frames, 3x3 conv over (mel, frame)) walk along frames
  float* dst = a.out + (long long)b * 80 * a.F + f;
  for (int m = lane; m < 80; m += 32) {
    const int ks = a.mel_range[2 * m], ke = a.mel_range[2 * m + 1];
    const float* wrow = a.mel_w + m * FB_BINS;
    float acc = 0.f;
    for (int k = ks; k < ke; ++k) acc = fmaf(P[k], __ldg(wrow + k), acc);
    dst[(long long)m * a.F] = logf(fmaxf(acc, 1.1920928955078125e-07f));
  }
}

cudaError_t launch_fbank(const FbankArgs& a, int B, cudaStream_t st) {
  dim3 grid((a.F + 7) / 8, B);
  const size_t smem = sizeof(float2) * (256 + 8 * FB_NFFT);
  fbank_kernel<<<grid, 256, smem, st>>>(a);
  return cudaGetLastError();
}

// per (window, mel) mean over frames: one warp per contiguous row of the mel-major features.  Lane l sums frames l, l + 32, ...
// in order and the 32 partial sums are combined by the shuffle tree (fp32; the reference's torch.mean uses another order).
__global__ void __launch_bounds__(256) fbank_mean_kernel(const float* __restrict__ fb, int rows, int F, float* __restrict__ mean) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* p = fb + (long long)r * F;
  float s = 0.f;
  for (int f = lane; f < F; f += 32) s += p[f];
  s = warp_sum(s);
  if (lane == 0) mean[r] = s / F;
}
cudaError_t launch_fbank_mean(const float* fb, int B, int F, float* mean, cudaStream_t st) {
  fbank_mean_kernel<<<(B * 80 + 7) / 8, 256, 0, st>>>(fb, B * 80, F, mean);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// conv1 (1 -> 32, 3x3, pad 1) on the mean-subtracted fbank + folded BN + ReLU -> zero-bordered NHWC planes
// [b][h = mel][1 + w = frame][32].  One thread per output pixel, 32 channels.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) emb_conv1_kernel(Conv1Args a) {
  __shared__ float w[9 * 32], sc[32], sh[32];
  for (int i = threadIdx.x; i < 288; i += blockDim.x) {
    const int c = i / 9, t = i - c * 9;
    w[t * 32 + c] = a.w[i];   // [tap][c]
  }
  for (int i = threadIdx.x; i < 32; i += blockDim.x) { sc[i] = a.scale[i]; sh[i] = a.shift[i]; }
  __syncthreads();
  const long long total = (long long)a.B * 80 * a.F;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int wf = (int)(i % a.F);
    const long long r = i / a.F;
    const int h = (int)(r % 80), b = (int)(r / 80);
    float in[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int hh = h + kh - 1, ww = wf + kw - 1;
        float v = 0.f;
        if (hh >= 0 && hh < 80 && ww >= 0 && ww < a.F) v = a.fb[((long long)b * 80 + hh) * a.F + ww] - a.mean[b * 80 + hh];
        in[kh * 3 + kw] = v;
      }
    bf16* o = a.out + (((long long)b * 80 + h) * (a.F + 2) + wf + 1) * 32;
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
      uint32_t hw[4], lw[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v2[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c = c8 * 8 + e * 2 + q;
          float acc = 0.f;
#pragma unroll
          for (int t = 0; t < 9; ++t) acc = fmaf(in[t], w[t * 32 + c], acc);
          v2[q] = fmaxf(acc * sc[c] + sh[c], 0.f);
        }
        bf16 h0, l0, h1, l1;
        split_bf16(v2[0], h0, l0, a.fp16);
        split_bf16(v2[1], h1, l1, a.fp16);
        hw[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        lw[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
      }
      *reinterpret_cast<uint4*>(o + c8 * 8) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      if (a.planes > 1) *reinterpret_cast<uint4*>(o + a.out_plane + c8 * 8) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
  }
}
cudaError_t launch_emb_conv1(const Conv1Args& a, cudaStream_t st) {
  const long long total = (long long)a.B * 80 * a.F;
  const int grid = (int)min((long long)148 * 16, (total + 255) / 256);
  emb_conv1_kernel<<<grid, 256, 0, st>>>(a);
  return cudaGetLastError();
}

// zero the two border columns (w = 0 and w = W + 1) of a zero-bordered NHWC plane set [B*H][W + 2][C] (+ slack)
__global__ void zero_borders_kernel(bf16* __restrict__ p, long long plane, int planes, long long rows, int W, int C) {
  const long long total = rows * 2 * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long r = i / C;
    const long long row = r >> 1;
    const int side = (int)(r & 1);
    const long long o = (row * (W + 2) + (side ? W + 1 : 0)) * C + c;
    p[o] = __float2bfloat16_rn(0.f);
    if (planes > 1) p[plane + o] = __float2bfloat16_rn(0.f);
  }
}
cudaError_t launch_zero_borders(bf16* p, long long plane, int planes, long long rows, int W, int C, cudaStream_t st) {
  const long long total = rows * 2 * C;
  const int grid = (int)min((long long)148 * 8, (total + 255) / 256);
  zero_borders_kernel<<<grid, 256, 0, st>>>(p, plane, planes, rows, W, C);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K15 masked statistics pooling.  x: zero-bordered NHWC planes [b][h][1 + w][C] (C = 256, h < H = 10);
// masks [b][S][T] -> weights nearest-interpolated to W frames through widx[w];
//   v1 = sum w + 1e-8; mean = sum x w / v1; var = sum (x - mean)^2 w / (v1 - sum w^2 / v1 + 1e-8); std = sqrt(var)
// out planes [b*S + s][2*C*H]: mean at c*H + h, std at C*H + c*H + h ("(dimension channel)" flattening).
// One CTA per (b, h), one thread per channel.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stats_pool_kernel(PoolArgs a) {
  extern __shared__ float wsm[];  // [S][W]
  const int b = blockIdx.y, h = blockIdx.x, c = threadIdx.x;
  for (int i = threadIdx.x; i < a.S * a.W; i += blockDim.x) {
    const int s = i / a.W, w = i - s * a.W;
    wsm[i] = a.masks[((long long)b * a.S + s) * a.T + a.widx[w]];
  }
  __syncthreads();
  if (c >= a.C) return;
  const bf16* xp = a.x + (((long long)b * a.H + h) * (a.W + 2) + 1) * a.C + c;
  float v1[4] = {0.f, 0.f, 0.f, 0.f}, v2[4] = {0.f, 0.f, 0.f, 0.f}, sx[4] = {0.f, 0.f, 0.f, 0.f};
  for (int w = 0; w < a.W; ++w) {
    float x = from16(xp[(long long)w * a.C], a.fp16);
    if (a.planes > 1) x += from16(xp[a.x_plane + (long long)w * a.C], a.fp16);
#pragma unroll
    for (int s = 0; s < 4; ++s)
      if (s < a.S) { const float wt = wsm[s * a.W + w]; v1[s] += wt; v2[s] += wt * wt; sx[s] += x * wt; }
  }
  float mean[4], dx[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int s = 0; s < 4; ++s) { v1[s] += 1e-8f; mean[s] = sx[s] / v1[s]; }
  for (int w = 0; w < a.W; ++w) {
    float x = from16(xp[(long long)w * a.C], a.fp16);
    if (a.planes > 1) x += from16(xp[a.x_plane + (long long)w * a.C], a.fp16);
#pragma unroll
    for (int s = 0; s < 4; ++s)
      if (s < a.S) { const float dl = x - mean[s]; dx[s] += dl * dl * wsm[s * a.W + w]; }
  }
  const int feat = a.C * a.H;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    if (s >= a.S) break;
    const float var = dx[s] / (v1[s] - v2[s] / v1[s] + 1e-8f);
    const float sd = sqrtf(var);
    bf16* o = a.out + ((long long)b * a.S + s) * a.ldo;
    bf16 hh, ll;
    split_bf16(mean[s], hh, ll, a.fp16);
    o[c * a.H + h] = hh;
    if (a.planes > 1) o[a.out_plane + c * a.H + h] = ll;
    split_bf16(sd, hh, ll, a.fp16);
    o[feat + c * a.H + h] = hh;
    if (a.planes > 1) o[a.out_plane + feat + c * a.H + h] = ll;
  }
}
cudaError_t launch_stats_pool(const PoolArgs& a, int B, cudaStream_t st) {
  if (a.S > 4 || a.C > 256) return cudaErrorInvalidValue;
  dim3 grid(a.H, B);
  const size_t smem = sizeof(float) * (size_t)a.S * a.W;
  stats_pool_kernel<<<grid, 256, smem, st>>>(a);
  return cudaGetLastError();
}

}  // namespace dz
