// Layer 0 of the WavLM feature extractor on tcgen05: conv1d(1 -> C0, k = 10, stride 5, no bias) + normalisation + GELU,
// channels-last 16-bit output.  reference: diarizen/models/module/wav2vec2/components.py:119-122 (conv -> norm -> gelu),
// :63-70 (LayerNorm over channels, "large") / GroupNorm(C0, C0) over time ("base", as a per-(window, channel) affine from
// conv0_gn_coef_kernel), model.py waveform normalisation for "large".
//
// The CUDA-core kernel (seg_kernels.cu conv0_kernel) spends a third of its issue slots on the 10-tap FMAs and their
// shared-memory weight loads and, for LayerNorm, keeps a whole 512-channel row in registers.  Here
//   * a tile is 128 consecutive frames x all channels: A = [128 frames][16] (10 taps + zero pad) as fp16 hi and lo planes
//     (two accumulating UMMAs: the waveform keeps ~22 bits), built in shared memory by one warp in the canonical
//     no-swizzle K-major core-matrix layout (8 rows x 16 B; LBO = 128 B between the two K halves, SBO = 256 B between
//     8-row groups); B = W[C0][16] fp16, resident;
//   * D = [128][C0 <= 512] fp32 is the whole TMEM of the SM.  LayerNorm needs the row statistics before anything can be
//     written, so the accumulator is simply read three times (sum, centred sum of squares, normalise + GELU + store) -
//     TMEM reads are cheap, nothing is recomputed and no row ever lives in registers;
//   * 16 epilogue warps: warp = (TMEM lane quarter, column group); the four warps of a quarter exchange their partial row
//     sums through shared memory with a 128-thread named barrier.
// fp16 operand mode only (weights are rounded to fp16: relative 2^-11, below the fp16 rounding of the outputs).
#include <cstdint>

#include "common.cuh"
#include "seg_kernels.h"

namespace dz {

static constexpr int C0T_THREADS = 576;   // 16 epilogue warps + UMMA issuer + A builder

DZ_DEVINL uint64_t umma_desc_nosw(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(128 >> 4) << 16;   // leading byte offset: next core matrix along K
  d |= (uint64_t)(256 >> 4) << 32;   // stride byte offset: next 8-row group
  d |= (uint64_t)1 << 46;            // descriptor version; layout type 0 = no swizzle
  return d;
}
DZ_DEVINL uint32_t core_off(int r, int k) { return (uint32_t)((r >> 3) * 256 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2); }

template <bool LARGE>
__global__ void __launch_bounds__(C0T_THREADS, 1) conv0_tc_kernel(const Conv0Args a, const int tpw, const int ntiles, const int per_cta) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* Bw = smem;                         // 512 rows x 32 B
  uint8_t* As = Bw + 16384;                   // 2 buffers x (hi plane 4 KB | lo plane 4 KB)
  float* xs = reinterpret_cast<float*>(As + 16384);   // 648 samples
  float* cf = xs + 656;                       // [512][2]: gamma/beta (large) or scale/shift of the current window (base)
  float* red1 = cf + 1024;                    // [128][4]
  float* red2 = red1 + 512;                   // [128][4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(red2 + 512);
  uint64_t* a_full = bars;        // [2]
  uint64_t* a_empty = bars + 2;   // [2]
  uint64_t* d_full = bars + 4;
  uint64_t* d_empty = bars + 5;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C0 = a.C0, NP = (C0 + 15) & ~15;
  const int tile_lo = blockIdx.x * per_cta, tile_hi = min(ntiles, tile_lo + per_cta);

  if (threadIdx.x == 0) {
    mbar_init(&a_full[0], 1); mbar_init(&a_full[1], 1);
    mbar_init(&a_empty[0], 1); mbar_init(&a_empty[1], 1);
    mbar_init(d_full, 1);
    mbar_init(d_empty, 16);
    mbar_fence_init();
  }
  if (warp == 16) tmem_alloc(tmem_ptr, 512);
  // resident weights (fp16) and the zero K padding of both A buffers
  for (int i = threadIdx.x; i < 512 * 16; i += C0T_THREADS) {
    const int n = i >> 4, k = i & 15;
    const float w = (n < C0 && k < 10) ? a.w[n * 10 + k] : 0.f;
    *reinterpret_cast<__half*>(Bw + core_off(n, k)) = __float2half_rn(w);
  }
  for (int i = threadIdx.x; i < 16384 / 4; i += C0T_THREADS) reinterpret_cast<uint32_t*>(As)[i] = 0u;
  if (LARGE)
    for (int i = threadIdx.x; i < 512; i += C0T_THREADS) { cf[2 * i] = i < C0 ? a.gamma[i] : 0.f; cf[2 * i + 1] = i < C0 ? a.beta[i] : 0.f; }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 17) {
    // ---------------- A builder ----------------
    uint32_t n = 0;
    for (int tile = tile_lo; tile < tile_hi; ++tile, ++n) {
      const int b = tile / tpw, t0 = (tile - b * tpw) * 128;
      const uint32_t buf = n & 1;
      if (n >= 2) mbar_wait(&a_empty[buf], ((n >> 1) - 1) & 1);
      const float* x = a.wav + (long long)b * a.N;
      float mu = 0.f, rs = 1.f;
      if (LARGE) { mu = a.wstats[2 * b]; rs = a.wstats[2 * b + 1]; }
      __syncwarp();
      for (int i = lane; i < 5 * 128 + 5; i += 32) {
        const long long s = 5LL * t0 + i;
        xs[i] = (s < a.N) ? (x[s] - mu) * rs : 0.f;
      }
      __syncwarp();
      uint8_t* Ah = As + buf * 8192;
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int r = lane + 32 * rr;
        uint32_t hw[5], lw[5];
#pragma unroll
        for (int k2 = 0; k2 < 5; ++k2) {
          const float v0 = xs[5 * r + 2 * k2], v1 = xs[5 * r + 2 * k2 + 1];
          const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
          const __half l0 = __float2half_rn(v0 - __half2float(h0)), l1 = __float2half_rn(v1 - __half2float(h1));
          hw[k2] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
          lw[k2] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
        }
        const uint32_t o = core_off(r, 0);
        *reinterpret_cast<uint4*>(Ah + o) = make_uint4(hw[0], hw[1], hw[2], hw[3]);             // taps 0..7
        *reinterpret_cast<uint32_t*>(Ah + o + 128) = hw[4];                                      // taps 8, 9 (10..15 stay 0)
        *reinterpret_cast<uint4*>(Ah + 4096 + o) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        *reinterpret_cast<uint32_t*>(Ah + 4096 + o + 128) = lw[4];
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[buf]);
    }
  } else if (warp == 16) {
    // ---------------- UMMA issuer ----------------
    if (lane == 0) {
      const int N0 = NP < 256 ? NP : 256, N1 = NP - N0;
      const uint32_t idesc0 = umma_idesc_bf16(128, (uint32_t)N0, 1);
      const uint32_t idesc1 = N1 > 0 ? umma_idesc_bf16(128, (uint32_t)N1, 1) : 0u;
      const uint32_t bw = smem_u32(Bw);
      uint32_t n = 0;
      for (int tile = tile_lo; tile < tile_hi; ++tile, ++n) {
        const uint32_t buf = n & 1;
        mbar_wait(&a_full[buf], (n >> 1) & 1);
        if (n >= 1) mbar_wait(d_empty, (n - 1) & 1);   // the previous tile has been read out of TMEM
        tc_fence_after();
        const uint32_t ah = smem_u32(As + buf * 8192);
        umma_bf16(tmem_base, umma_desc_nosw(ah), umma_desc_nosw(bw), idesc0, 0u);
        umma_bf16(tmem_base, umma_desc_nosw(ah + 4096), umma_desc_nosw(bw), idesc0, 1u);
        if (N1 > 0) {
          umma_bf16(tmem_base + 256, umma_desc_nosw(ah), umma_desc_nosw(bw + 8192), idesc1, 0u);
          umma_bf16(tmem_base + 256, umma_desc_nosw(ah + 4096), umma_desc_nosw(bw + 8192), idesc1, 1u);
        }
        umma_commit(&a_empty[buf]);
        umma_commit(d_full);
      }
    }
  } else {
    // ---------------- epilogue: warp = (lane quarter, column group) ----------------
    const int quad = warp & 3, cg = warp >> 2;
    const int row = quad * 32 + lane;
    const int ldo = a.ldo;
    const int CGW = 32 * ((ldo + 127) / 128);          // columns per group: 32 / 64 / 96 / 128
    const int col_lo = cg * CGW;
    const int nj = (col_lo < ldo) ? min(CGW, ldo - col_lo + 31) / 32 : 0;   // 32-column chunks this warp owns
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const float invC = 1.0f / (float)C0;
    int cur_b = -1;
    uint32_t n = 0;
    for (int tile = tile_lo; tile < tile_hi; ++tile, ++n) {
      const int b = tile / tpw, t0 = (tile - b * tpw) * 128;
      const int t = t0 + row;
      if (!LARGE && b != cur_b) {
        asm volatile("bar.sync 5, 512;" ::: "memory");   // everyone is done with the previous window's coefficients
        const int i = threadIdx.x;                        // 512 epilogue threads <-> 512 channels
        float2 v = make_float2(0.f, 0.f);
        if (i < C0) v = *reinterpret_cast<const float2*>(a.coef + ((long long)b * C0 + i) * 2);
        cf[2 * i] = v.x; cf[2 * i + 1] = v.y;
        asm volatile("bar.sync 5, 512;" ::: "memory");
        cur_b = b;
      }
      mbar_wait(d_full, n & 1);
      tc_fence_after();
      float sc = 1.f, of = 0.f;
      if (LARGE) {
        float s = 0.f;
        for (int j = 0; j < nj; ++j) {
          const int c0 = col_lo + 32 * j;
          if (c0 >= NP) break;
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + lane_off + (uint32_t)c0, r);
          tmem_ld_wait();
          const int nv = NP - c0;   // columns >= NP were never written
          if (nv >= 32) {           // whole chunk valid (every chunk when C0 is a multiple of 32): no per-element predicates
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              s0 += __uint_as_float(r[c]); s1 += __uint_as_float(r[c + 1]); s2 += __uint_as_float(r[c + 2]); s3 += __uint_as_float(r[c + 3]);
            }
            s += (s0 + s1) + (s2 + s3);
            continue;
          }
#pragma unroll
          for (int c = 0; c < 32; ++c) s += (c < nv) ? __uint_as_float(r[c]) : 0.f;   // channels C0..NP-1 are exact zeros
        }
        red1[row * 4 + cg] = s;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + quad) : "memory");
        const float4 p4 = *reinterpret_cast<const float4*>(red1 + row * 4);
        const float mean = ((p4.x + p4.y) + (p4.z + p4.w)) * invC;
        float q = 0.f;
        for (int j = 0; j < nj; ++j) {
          const int c0 = col_lo + 32 * j;
          if (c0 >= C0) break;
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + lane_off + (uint32_t)c0, r);
          tmem_ld_wait();
          const int nv = C0 - c0;
          if (nv >= 32) {
            float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              const float d0 = __uint_as_float(r[c]) - mean, d1 = __uint_as_float(r[c + 1]) - mean;
              const float d2 = __uint_as_float(r[c + 2]) - mean, d3 = __uint_as_float(r[c + 3]) - mean;
              q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); q2 = fmaf(d2, d2, q2); q3 = fmaf(d3, d3, q3);
            }
            q += (q0 + q1) + (q2 + q3);
            continue;
          }
#pragma unroll
          for (int c = 0; c < 32; ++c) { const float dl = __uint_as_float(r[c]) - mean; q = (c < nv) ? fmaf(dl, dl, q) : q; }
        }
        red2[row * 4 + cg] = q;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + quad) : "memory");
        const float4 q4 = *reinterpret_cast<const float4*>(red2 + row * 4);
        sc = rsqrtf(((q4.x + q4.y) + (q4.z + q4.w)) * invC + 1e-5f);
        of = -mean * sc;
      }
      {
        // tcgen05.ld is warp-collective (.sync.aligned): every lane takes part in the loads, also the lanes whose frame lies
        // past the end of the window (the last tile of a window has 127 valid rows); only the stores are predicated
        const bool rvalid = t < a.T0;
        bf16* orow = a.out + (long long)b * a.out_bstride + (long long)(rvalid ? t : 0) * ldo;
        for (int j = 0; j < nj; ++j) {
          const int c0 = col_lo + 32 * j;
          uint32_t r[32];
          if (c0 < NP) {
            tmem_ld_32x32(tmem_base + lane_off + (uint32_t)c0, r);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) r[c] = 0u;
          }
          if (c0 + 32 <= C0) {
            // whole chunk inside the channel range: straight-line code, no per-element predicates or branches around the GELU
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
              uint32_t w4[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float4 g = *reinterpret_cast<const float4*>(cf + 2 * (c0 + 8 * c8 + 2 * e));
                float y0 = __uint_as_float(r[8 * c8 + 2 * e]), y1 = __uint_as_float(r[8 * c8 + 2 * e + 1]);
                if (LARGE) { y0 = fmaf(y0, sc, of); y1 = fmaf(y1, sc, of); }
                w4[e] = pack2_16<1>(gelu_erf(fmaf(y0, g.x, g.y)), gelu_erf(fmaf(y1, g.z, g.w)));
              }
              if (rvalid) *reinterpret_cast<uint4*>(orow + c0 + 8 * c8) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
            }
            continue;
          }
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            const int cb = c0 + 8 * c8;
            if (cb >= ldo) break;
            uint32_t w4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = cb + 2 * e;
              const float4 g = *reinterpret_cast<const float4*>(cf + 2 * min(c, 510));   // (p0, q0, p1, q1) of channels c, c + 1
              float y0 = __uint_as_float(r[8 * c8 + 2 * e]), y1 = __uint_as_float(r[8 * c8 + 2 * e + 1]);
              if (LARGE) { y0 = fmaf(y0, sc, of); y1 = fmaf(y1, sc, of); }
              y0 = (c < C0) ? gelu_erf(fmaf(y0, g.x, g.y)) : 0.f;
              y1 = (c + 1 < C0) ? gelu_erf(fmaf(y1, g.z, g.w)) : 0.f;
              w4[e] = pack2_16<1>(y0, y1);
            }
            if (rvalid) *reinterpret_cast<uint4*>(orow + cb) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

bool conv0_tc_eligible(const Conv0Args& a) {
  return a.fp16 == 1 && a.planes == 1 && a.C0 <= 512 && (a.ldo % 8) == 0 && a.ldo <= 512 && a.ldo >= a.C0;
}

cudaError_t launch_conv0_tc(const Conv0Args& a, int B, bool large, cudaStream_t st) {
  const size_t smem = 1024 + 16384 + 16384 + sizeof(float) * (656 + 1024 + 512 + 512) + 64 + 16;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(conv0_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv0_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tpw = (a.T0 + 127) / 128;
  const long long ntiles = (long long)B * tpw;
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  const int per_cta = (int)((ntiles + grid - 1) / grid);
  if (large) conv0_tc_kernel<true><<<grid, C0T_THREADS, smem, st>>>(a, tpw, (int)ntiles, per_cta);
  else conv0_tc_kernel<false><<<grid, C0T_THREADS, smem, st>>>(a, tpw, (int)ntiles, per_cta);
  return cudaGetLastError();
}

}  // namespace dz
