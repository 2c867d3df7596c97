// WavLM positional convolution (grouped conv1d, 16 groups, k = 128, pad 64, weight-norm folded) + bias + GELU + residual
// add on tcgen05.  reference: diarizen/models/module/wav2vec2/components.py:352-386 (ConvolutionalPositionalEmbedding),
// :903-905 (x = x + pos_conv(x)).
//
// As an implicit GEMM (gemm_tc.cu, rank-5 map over the zero-padded staging copy) every tap re-fetches a 128 x 64 slice of
// the input that overlaps the previous tap's slice in 127 of its 128 rows, and the loop is bound by that L2 -> shared
// traffic (24 KB per 1 MFLOP k-block).  Here the input window of a tile is loaded ONCE:
//   tile = (window b, group g, up to 256 consecutive frames)  ->  A window = (256 + 127) rows x 64 channels (128-byte rows,
//   128-byte swizzle, 48 KB); tap k of accumulator block m reads it through a descriptor whose start address is shifted by
//   (128 m + k) rows - valid because the swizzle is a function of the shared-memory address (same trick as conv3x3_c32.cu);
//   only the weights stream: one 64 x 64 tap slice (8 KB) per 2 x 4 UMMAs, 32 B/clk/SM.
// Roles: warp 0 TMA producer, warp 1 UMMA issuer, warps 2..5 epilogue (one accumulator row per thread: 64 channels of one
// frame = 256 contiguous bytes of the fp32 residual stream, read-modify-written in place).  TMEM: two tiles x two 64-column
// accumulators, so the epilogue of a tile overlaps the taps of the next one.
#include <cstdint>
#include <string>

#include "common.cuh"
#include "gemm.h"
#include "seg_kernels.h"

namespace dz {

bool make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                    const uint32_t* box);

static constexpr int PC_THREADS = 192;
static constexpr int PC_AWIN = 48 * 1024;     // 3 boxes of 128 rows x 128 B (256 + 127 rows are read)
static constexpr int PC_NSB = 8;              // weight tap stages
static constexpr int PC_TAPB = 8192;          // 64 rows x 128 B
static constexpr int PC_TAPS = 128;

struct PosConvMaps { CUtensorMap x, w; };

template <int FP16>
__global__ void __launch_bounds__(PC_THREADS, 1) posconv_tc_kernel(const __grid_constant__ PosConvMaps maps, const PosConvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* Aw = smem;                                  // 2 x PC_AWIN
  uint8_t* Bs = Aw + 2 * PC_AWIN;                      // PC_NSB x PC_TAPB
  uint64_t* bars = reinterpret_cast<uint64_t*>(Bs + PC_NSB * PC_TAPB);
  uint64_t* a_full = bars;                // [2]
  uint64_t* a_empty = bars + 2;           // [2]
  uint64_t* b_full = bars + 4;            // [NSB]
  uint64_t* b_empty = b_full + PC_NSB;    // [NSB]
  uint64_t* t_full = b_empty + PC_NSB;    // [2]
  uint64_t* t_empty = t_full + 2;         // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(t_empty + 2);
  float* sbias = reinterpret_cast<float*>(tmem_ptr + 2);   // [16 groups x 64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk_seq = (a.T + 127) / 128;              // 128-frame blocks per window
  const int tps = (nblk_seq + 1) / 2;                  // tiles per (window, group)
  const int ntiles = a.B * 16 * tps;
  const int NN = (a.Dg + 15) & ~15;                    // UMMA N

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
    for (int s = 0; s < PC_NSB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_fence_init();
    tma_prefetch_desc(&maps.x); tma_prefetch_desc(&maps.w);
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 256);
  for (int i = threadIdx.x; i < 16 * 64; i += PC_THREADS) {
    const int g = i >> 6, n = i & 63;
    sbias[i] = (a.bias != nullptr && n < a.Dg) ? a.bias[g * a.Dg + n] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // tile -> (window, group, first frame, number of 128-frame blocks); the group is the fastest index so that concurrently
  // running CTAs share the same input rows in L2
  auto decode = [&](int tile, int& b, int& g, int& t0, int& nb) {
    g = tile & 15;
    const int r = tile >> 4;
    const int tb = r % tps;
    b = r / tps;
    t0 = tb * 256;
    nb = (nblk_seq - 2 * tb) >= 2 ? 2 : 1;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t gt = 0, n = 0;   // running tap counter (weight ring), tile counter
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++n) {
        int b, g, t0, nb;
        decode(tile, b, g, t0, nb);
        const uint32_t ab = n & 1;
        if (n >= 2) mbar_wait(&a_empty[ab], ((n >> 1) - 1) & 1);
        // staging rows t0 .. t0 + 128 nb + 126 (frame t, tap k reads staging row t + k), fetched as nb + 1 boxes of 128
        // rows; rows past the end of the staging buffer are zero-filled
        mbar_expect_tx(&a_full[ab], (uint32_t)(nb + 1) * 16384u);
        for (int i = 0; i <= nb; ++i) tma_load_3d(Aw + ab * PC_AWIN + i * 16384, &maps.x, &a_full[ab], g * 64, t0 + i * 128, b);
        for (int k = 0; k < PC_TAPS; ++k, ++gt) {
          const uint32_t s = gt % PC_NSB, use = gt / PC_NSB;
          if (use > 0) mbar_wait(&b_empty[s], (use - 1) & 1);
          mbar_expect_tx(&b_full[s], PC_TAPB);
          tma_load_3d(Bs + s * PC_TAPB, &maps.w, &b_full[s], k * 64, 0, g);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)NN, FP16);
      uint32_t gt = 0, n = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++n) {
        int b, g, t0, nb;
        decode(tile, b, g, t0, nb);
        const uint32_t ab = n & 1;
        mbar_wait(&a_full[ab], (n >> 1) & 1);
        if (n >= 2) mbar_wait(&t_empty[ab], ((n >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t aw = smem_u32(Aw + ab * PC_AWIN);
        const uint32_t tacc = tmem_base + ab * 128;
        for (int k = 0; k < PC_TAPS; ++k, ++gt) {
          const uint32_t s = gt % PC_NSB;
          mbar_wait(&b_full[s], (gt / PC_NSB) & 1);
          tc_fence_after();
          const uint32_t bs = smem_u32(Bs + s * PC_TAPB);
          for (int m = 0; m < nb; ++m) {
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2)
              umma_bf16(tacc + m * 64, umma_desc_sw128(aw + (m * 128 + k) * 128 + k2 * 32), umma_desc_sw128(bs + k2 * 32), idesc,
                        (k | k2) ? 1u : 0u);
          }
          umma_commit(&b_empty[s]);
        }
        umma_commit(&a_empty[ab]);
        umma_commit(&t_full[ab]);
      }
    }
  } else {
    const int quad = warp & 3;              // TMEM lane quarter (warps 2..5 -> 2, 3, 0, 1)
    const int row = quad * 32 + lane;
    uint32_t n = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++n) {
      int b, g, t0, nb;
      decode(tile, b, g, t0, nb);
      const uint32_t ab = n & 1;
      mbar_wait(&t_full[ab], (n >> 1) & 1);
      tc_fence_after();
      const float* bias = sbias + g * 64;
      for (int m = 0; m < nb; ++m) {
        const int t = t0 + m * 128 + row;
        float* xp = a.x + ((long long)b * a.T + t) * a.ldx + g * a.Dg;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + ab * 128 + m * 64 + h * 32 + ((uint32_t)(quad * 32) << 16), r);
          tmem_ld_wait();
          if (t < a.T) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int c = h * 32 + 4 * q;
              if (c < a.Dg) {   // Dg is a multiple of 4 (checked on the host)
                float4 o = *reinterpret_cast<const float4*>(xp + c);
                o.x += gelu_erf(__uint_as_float(r[4 * q]) + bias[c]);
                o.y += gelu_erf(__uint_as_float(r[4 * q + 1]) + bias[c + 1]);
                o.z += gelu_erf(__uint_as_float(r[4 * q + 2]) + bias[c + 2]);
                o.w += gelu_erf(__uint_as_float(r[4 * q + 3]) + bias[c + 3]);
                *reinterpret_cast<float4*>(xp + c) = o;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[ab]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

struct PosConvPlan {
  PosConvArgs a;
  PosConvMaps maps;
  size_t smem;
  int grid;
};

PosConvPlan* posconv_plan_create(const PosConvArgs& a) {
  if (a.Dg > 64 || (a.Dg % 4) != 0 || (a.ldx % 4) != 0) return nullptr;
  PosConvPlan* p = new PosConvPlan();
  p->a = a;
  {
    uint64_t dims[3] = {(uint64_t)a.stage_ld, (uint64_t)a.stage_rows, (uint64_t)a.B};
    uint64_t str[3] = {1, (uint64_t)a.stage_ld, (uint64_t)a.stage_rows * a.stage_ld};
    uint32_t box[3] = {64, 128, 1};
    if (!make_tmap_bf16(&p->maps.x, a.stage, 3, dims, str, box)) { delete p; return nullptr; }
  }
  {
    uint64_t dims[3] = {(uint64_t)a.ldw, (uint64_t)a.Dg, 16};
    uint64_t str[3] = {1, (uint64_t)a.ldw, (uint64_t)a.w_gstride};
    uint32_t box[3] = {64, 64, 1};
    if (!make_tmap_bf16(&p->maps.w, a.w, 3, dims, str, box)) { delete p; return nullptr; }
  }
  p->smem = 1024 + 2 * PC_AWIN + PC_NSB * PC_TAPB + 8 * (8 + 2 * PC_NSB) + 8 + 16 * 64 * 4 + 16;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int nblk_seq = (a.T + 127) / 128;
  const long long tiles = (long long)a.B * 16 * ((nblk_seq + 1) / 2);
  p->grid = (int)(tiles < sms ? tiles : sms);
  return p;
}
void posconv_plan_destroy(PosConvPlan* p) { delete p; }

cudaError_t posconv_plan_launch(const PosConvPlan* p, cudaStream_t st) {
  static size_t attr = 0;
  if (p->smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(posconv_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(posconv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem);
    if (e != cudaSuccess) return e;
    attr = p->smem;
  }
  if (p->a.fp16) posconv_tc_kernel<1><<<p->grid, PC_THREADS, p->smem, st>>>(p->maps, p->a);
  else posconv_tc_kernel<0><<<p->grid, PC_THREADS, p->smem, st>>>(p->maps, p->a);
  return cudaGetLastError();
}

}  // namespace dz
