// 3x3 / stride 1 / pad 1 convolution, 128 -> 128 channels, on tcgen05 with STREAMED weights: the stride-1 convolutions of
// layer3 of the WeSpeaker ResNet34 (pyannote-audio/pyannote/audio/models/embedding/wespeaker/resnet.py:56-79,:141-147; 11
// of the 12 layer3 convolutions, 37 % of the embedding forward), folded BatchNorm bias, optional 16-bit residual, ReLU.
//
// The generic implicit GEMM (gemm_tc.cu, conv2d mode) runs these at ~440 TFLOP/s: a 128-pixel tile pulls 576 KB from L2
// (every input pixel nine times, the 295 KB weight once per tile) through a 4-deep 32 KB ring, and with ~1 us of L2 latency
// that ring depth - not the tensor pipe (24 % active) - sets the pace.  This kernel needs 2.7x fewer bytes per FLOP:
//   * a tile is 128 pixels x TWO output rows (h, h + 1): the four input rows h - 1 .. h + 2 are loaded once and serve both
//     rows and all nine taps - the three kw taps of a row are UMMA A-descriptors shifted by kw pixel rows (kw x 128 B) into
//     the same 130-pixel box (128-byte swizzle is a function of the address, so a row-shifted view is still K-major);
//   * the weight is streamed one (tap, 64-input-channel half) slab = [128 out][64 in] = 16 KB at a time through a 4-slot
//     ring and multiplied against both output rows before the slot is released: 288 KB of weights per 2 x 128 pixels;
//   * the input is split by channel half too: while the MMAs of half 0 run, half 1 (and then half 0 of the next tile) is
//     in flight - a two-slot pipeline of 68 KB loads without a second full tile buffer.
// Roles (352 threads, 1 CTA / SM, persistent): warp 0 input producer, warp 1 weight producer, warp 2 UMMA issuer, warps
// 3..10 epilogue (TMEM lane quarter x output row).  Accumulators: 2 rows x 128 fp32 columns, double buffered (512 TMEM
// columns), so the epilogue of tile t overlaps the MMAs of tile t + 1.
#include <cstdint>
#include <string>

#include "common.cuh"
#include "emb_kernels.h"
#include "gemm.h"

namespace dz {

bool make_tmap_sw(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                  const uint32_t* box, int swizzle_bytes);

namespace {
constexpr int CS_THREADS = 352;
constexpr int CS_C = 128;
constexpr int CS_BOX = 17 * 1024;               // one input row box: 130 pixels x 64 channels x 2 B = 16640 B, 1 KB aligned
constexpr int CS_HALF = 4 * CS_BOX;             // rows h-1 .. h+2 of one channel half
constexpr int CS_WSLAB = 128 * 128;             // [128 out][64 in] x 2 B
constexpr int CS_WSLOTS = 4;
constexpr int CS_SMEM = 1024 + 2 * CS_HALF + CS_WSLOTS * CS_WSLAB + 512 + 128 * 4;

struct ConvSMaps { CUtensorMap in, w; };

DZ_DEVINL uint64_t desc_rows128(uint32_t smem_addr) {   // K-major rows of 128 B, SWIZZLE_128B, 8-row groups 1024 B apart
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
}  // namespace

template <int FP16>
__global__ void __launch_bounds__(CS_THREADS, 1) conv3x3_stream_kernel(const __grid_constant__ ConvSMaps maps, const Conv3Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* In = smem;                                   // [2 halves][4 rows][CS_BOX]
  uint8_t* Ws = In + 2 * CS_HALF;                       // [CS_WSLOTS][CS_WSLAB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Ws + CS_WSLOTS * CS_WSLAB);
  uint64_t* in_full = bars;            // [2]
  uint64_t* in_empty = bars + 2;       // [2]
  uint64_t* w_full = bars + 4;         // [4]
  uint64_t* w_empty = bars + 8;        // [4]
  uint64_t* t_full = bars + 12;        // [2]
  uint64_t* t_empty = bars + 14;       // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 16);
  float* sbias = reinterpret_cast<float*>(bars + 64);   // [128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tpr = (a.W + 127) / 128;                    // tiles per image row
  const int hp = (a.H + 1) / 2;                         // row pairs
  const int ntiles = a.B * hp * tpr;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], 1); mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 8); }
    for (int i = 0; i < CS_WSLOTS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    mbar_fence_init();
    tma_prefetch_desc(&maps.in); tma_prefetch_desc(&maps.w);
  }
  if (warp == 2) tmem_alloc(tmem_ptr, 512);
  if (threadIdx.x < CS_C) sbias[threadIdx.x] = a.bias ? a.bias[threadIdx.x] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ---------------- input producer: per tile, channel half 0 then half 1, four row boxes each ----------------
    if (lane == 0) {
      uint32_t n = 0;                                   // (tile, half) counter
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int wt = tile % tpr, r = tile / tpr, p = r % hp, b = r / hp;
        for (int half = 0; half < 2; ++half, ++n) {
          const uint32_t use = n >> 1;                  // how many times this half's buffer has been filled before
          if (use > 0) mbar_wait(&in_empty[half], (use - 1) & 1);
          mbar_expect_tx(&in_full[half], 4 * 130 * 128);
          // rows outside [0, H) and pixels past the right border are zero-filled by TMA (the padded buffer has W + 2 pixels)
#pragma unroll
          for (int rr = 0; rr < 4; ++rr)
            tma_load_4d(In + half * CS_HALF + rr * CS_BOX, &maps.in, &in_full[half], half * 64, wt * 128, 2 * p - 1 + rr, b);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- weight producer: 18 slabs per tile (half-major, then tap) through the slot ring ----------------
    if (lane == 0) {
      constexpr int KRUN = (3 * CS_C + 63) / 64 * 64;   // K stride between kh blocks in the GEMM weight layout
      uint32_t n = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int half = 0; half < 2; ++half)
          for (int tap = 0; tap < 9; ++tap, ++n) {
            const uint32_t s = n % CS_WSLOTS, use = n / CS_WSLOTS;
            if (use > 0) mbar_wait(&w_empty[s], (use - 1) & 1);
            mbar_expect_tx(&w_full[s], CS_WSLAB);
            tma_load_2d(Ws + s * CS_WSLAB, &maps.w, &w_full[s], (tap / 3) * KRUN + (tap % 3) * CS_C + half * 64, 0);
          }
      }
    }
  } else if (warp == 2) {
    // ---------------- UMMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, CS_C, FP16);
      const uint32_t in0 = smem_u32(In), ws0 = smem_u32(Ws);
      uint32_t t = 0, nh = 0, nw = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
        const uint32_t acc = t & 1;
        if (t >= 2) mbar_wait(&t_empty[acc], ((t >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + acc * 256;
        for (int half = 0; half < 2; ++half, ++nh) {
          mbar_wait(&in_full[half], (nh >> 1) & 1);
          tc_fence_after();
          const uint32_t ia = in0 + half * CS_HALF;
          for (int tap = 0; tap < 9; ++tap, ++nw) {
            const uint32_t s = nw % CS_WSLOTS;
            mbar_wait(&w_full[s], (nw / CS_WSLOTS) & 1);
            tc_fence_after();
            const int kh = tap / 3, kw = tap - 3 * kh;
            const uint32_t wa = ws0 + s * CS_WSLAB;
#pragma unroll
            for (int row = 0; row < 2; ++row)
#pragma unroll
              for (int k2 = 0; k2 < 4; ++k2)            // 64 input channels = 4 steps of K = 16 (32 bytes)
                umma_bf16(tacc + row * 128, desc_rows128(ia + (row + kh) * CS_BOX + kw * 128 + k2 * 32), desc_rows128(wa + k2 * 32), idesc,
                          (half | tap | k2) ? 1u : 0u);
            umma_commit(&w_empty[s]);
          }
          umma_commit(&in_empty[half]);
        }
        umma_commit(&t_full[acc]);
      }
    }
  } else {
    // ---------------- epilogue: warp = (TMEM lane quarter, output row of the pair) ----------------
    const int quad = warp & 3;
    const int row = (warp - 3) >> 2;
    const long long Wp = a.W + 2;
    uint32_t t = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
      const int wt = tile % tpr, r = tile / tpr, p = r % hp, b = r / hp;
      const int h = 2 * p + row, wo = wt * 128 + quad * 32 + lane;
      const bool valid = h < a.H && wo < a.W;
      const long long pix = (((long long)b * a.H + (valid ? h : 0)) * Wp + (valid ? wo : 0) + 1) * CS_C;
      const uint32_t acc = t & 1;
      mbar_wait(&t_full[acc], (t >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < 4; ++g) {                     // 32 output channels at a time
        uint4 rq[4];
        if (a.res != nullptr && valid) {
          const uint4* rp = reinterpret_cast<const uint4*>(a.res + pix + 32 * g);
#pragma unroll
          for (int i = 0; i < 4; ++i) rq[i] = rp[i];
        }
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + acc * 256 + row * 128 + g * 32 + ((uint32_t)(quad * 32) << 16), v);
        tmem_ld_wait();
        if (valid) {
          uint4* op = reinterpret_cast<uint4*>(a.out + pix + 32 * g);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint32_t w4[4];
            const uint32_t rr[4] = {rq[i].x, rq[i].y, rq[i].z, rq[i].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = 8 * i + 2 * e;
              float v0 = __uint_as_float(v[c]) + sbias[32 * g + c], v1 = __uint_as_float(v[c + 1]) + sbias[32 * g + c + 1];
              if (a.res != nullptr) {
                v0 += from16(__ushort_as_bfloat16((unsigned short)(rr[e] & 0xffff)), FP16);
                v1 += from16(__ushort_as_bfloat16((unsigned short)(rr[e] >> 16)), FP16);
              }
              if (a.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
              w4[e] = pack2_16<FP16>(v0, v1);
            }
            op[i] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

struct ConvSPlan {
  Conv3Args a;
  ConvSMaps maps;
  int grid;
};

ConvSPlan* conv3x3_stream_plan_create(const Conv3Args& a) {
  if (a.C != CS_C) return nullptr;
  ConvSPlan* p = new ConvSPlan();
  p->a = a;
  {
    uint64_t dims[4] = {(uint64_t)CS_C, (uint64_t)a.W + 2, (uint64_t)a.H, (uint64_t)a.B};
    uint64_t str[4] = {1, (uint64_t)CS_C, (uint64_t)(a.W + 2) * CS_C, (uint64_t)a.H * (a.W + 2) * CS_C};
    uint32_t box[4] = {64, 130, 1, 1};
    if (!make_tmap_sw(&p->maps.in, a.in, 4, dims, str, box, 128)) { delete p; return nullptr; }
  }
  {
    uint64_t dims[2] = {(uint64_t)a.ldw, (uint64_t)CS_C};
    uint64_t str[2] = {1, (uint64_t)a.ldw};
    uint32_t box[2] = {64, (uint32_t)CS_C};
    if (!make_tmap_sw(&p->maps.w, a.w, 2, dims, str, box, 128)) { delete p; return nullptr; }
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long tiles = (long long)a.B * ((a.H + 1) / 2) * ((a.W + 127) / 128);
  p->grid = (int)(tiles < sms ? tiles : sms);
  return p;
}
void conv3x3_stream_plan_destroy(ConvSPlan* p) { delete p; }

cudaError_t conv3x3_stream_plan_launch(const ConvSPlan* p, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_stream_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv3x3_stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  if (p->a.fp16) conv3x3_stream_kernel<1><<<p->grid, CS_THREADS, CS_SMEM, st>>>(p->maps, p->a);
  else conv3x3_stream_kernel<0><<<p->grid, CS_THREADS, CS_SMEM, st>>>(p->maps, p->a);
  return cudaGetLastError();
}

}  // namespace dz
