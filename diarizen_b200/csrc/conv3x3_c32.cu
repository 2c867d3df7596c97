// 3x3 / stride 1 / pad 1 convolution for C -> C channels, C = 32 or 64, on tcgen05 (the stride-1 convolutions of layer1
// and layer2 of the WeSpeaker ResNet34: pyannote-audio/pyannote/audio/models/embedding/wespeaker/resnet.py:56-79,
// :141-147), folded BatchNorm bias, optional 16-bit residual and ReLU.  The numbers below are for C = 32.
//
// The generic implicit GEMM (gemm_tc.cu, conv2d mode) moves ~165 KB from L2 into shared memory per 128-pixel tile of this
// layer (every input pixel three times for the kw taps, K padded 96 -> 128, the 32 output channels padded to a 64-wide
// weight tile that is re-fetched by every tile) and is bound by exactly that traffic.  This kernel moves 25 KB:
//   * the whole folded weight (9 taps x [32 out][32 in], 18 KB) is loaded once per persistent CTA and stays resident;
//   * per tile and per kh ONE box of 130 input pixels x 32 channels (64-byte rows, 64-byte swizzle) is loaded; the three
//     kw taps are three UMMA A-descriptors whose start address is shifted by kw rows (kw x 64 B) into the same box -
//     the swizzle is a function of the shared-memory address, so a row-shifted view of a swizzled box is still a valid
//     K-major operand;
//   * the accumulator (128 pixels x 32 channels fp32) is read one row per thread; a row is 64 contiguous bytes of the NHWC
//     output, so residual loads and stores go straight to global memory, coalesced, without a staging patch.
// Roles per CTA (2 CTAs / SM): warp 0 TMA producer, warp 1 UMMA issuer, warps 2..9 two epilogue sets that alternate tiles;
// four 32-column accumulators in TMEM.
#include <cstdint>
#include <string>

#include "common.cuh"
#include "emb_kernels.h"
#include "gemm.h"

namespace dz {

bool make_tmap_sw(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                  const uint32_t* box, int swizzle_bytes);

static constexpr int C3_THREADS = 320;
template <int C>
struct C3Cfg {
  static constexpr int ROWB = 2 * C;                               // bytes per pixel row = swizzle span (64 or 128)
  static constexpr int NST = C == 32 ? 3 : 2;                      // input stages
  static constexpr int BOX = (130 * ROWB + 1023) / 1024 * 1024;    // one kh box: 130 pixels
  static constexpr int STAGE = 3 * BOX;
  static constexpr int TAPB = C * ROWB;                            // one tap of the weight: [C out][C in]
  static constexpr int WBYTES = 9 * TAPB;
  static constexpr int KRUN = (3 * C + 63) / 64 * 64;              // K stride between kh blocks in the GEMM weight layout
  static constexpr int CTAS = C == 32 ? 2 : 1;                     // resident CTAs per SM
  static constexpr int TMEM = 4 * C;                               // four accumulators
};

struct Conv3Maps { CUtensorMap in, w; };

// K-major operand whose rows are exactly one swizzle span (64 B -> SWIZZLE_64B, 128 B -> SWIZZLE_128B); 8-row groups
template <int ROWB>
DZ_DEVINL uint64_t umma_desc_rows(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * ROWB) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(ROWB == 64 ? 4 : 2) << 61;
  return d;
}

template <int C, int FP16>
__global__ void __launch_bounds__(C3_THREADS, C3Cfg<C>::CTAS) conv3x3_kernel(const __grid_constant__ Conv3Maps maps, const Conv3Args a) {
  using K = C3Cfg<C>;
  constexpr int C3_NST = K::NST, C3_BOX = K::BOX, C3_STAGE = K::STAGE, C3_WBYTES = K::WBYTES, ROWB = K::ROWB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* Wt = smem;                               // resident weights
  uint8_t* St = Wt + C3_WBYTES;                     // C3_NST stages x 3 boxes (kh = 0, 1, 2)
  uint64_t* bars = reinterpret_cast<uint64_t*>(St + C3_NST * C3_STAGE);
  uint64_t* w_full = bars;
  uint64_t* full = bars + 1;          // [NST]
  uint64_t* empty = full + C3_NST;    // [NST]
  uint64_t* t_full = empty + C3_NST;  // [4]
  uint64_t* t_empty = t_full + 4;     // [4]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(t_empty + 4);
  float* sbias = reinterpret_cast<float*>(tmem_ptr + 2);   // [C]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tpr = (a.W + 127) / 128;                  // tiles per image row
  const int ntiles = a.B * a.H * tpr;

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int s = 0; s < C3_NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
    mbar_fence_init();
    tma_prefetch_desc(&maps.in); tma_prefetch_desc(&maps.w);
  }
  if (warp == 1) tmem_alloc(tmem_ptr, K::TMEM);
  if (threadIdx.x < C) sbias[threadIdx.x] = a.bias ? a.bias[threadIdx.x] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(w_full, C3_WBYTES);
      for (int tap = 0; tap < 9; ++tap)   // B[n][kh*KRUN + kw*C + ci]: one C x C box per tap
        tma_load_2d(Wt + tap * K::TAPB, &maps.w, w_full, (tap / 3) * K::KRUN + (tap % 3) * C, 0);
      uint32_t t = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
        const int wt = tile % tpr, bh = tile / tpr, h = bh % a.H, b = bh / a.H;
        const uint32_t s = t % C3_NST, use = t / C3_NST;
        if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
        mbar_expect_tx(&full[s], 3 * 130 * ROWB);
        // padded pixel index of tap kw for output pixel wo is wo + kw: the box starts at wo0 and spans 130 pixels;
        // rows h - 1 / h + 1 outside the image and pixels past the right border are zero-filled by TMA
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) tma_load_4d(St + s * C3_STAGE + kh * C3_BOX, &maps.in, &full[s], 0, wt * 128, h + kh - 1, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, C, FP16);
      mbar_wait(w_full, 0);
      const uint32_t wa = smem_u32(Wt);
      uint32_t t = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
        const uint32_t s = t % C3_NST, acc = t & 3;
        mbar_wait(&full[s], (t / C3_NST) & 1);
        if (t >= 4) mbar_wait(&t_empty[acc], ((t >> 2) - 1) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(St + s * C3_STAGE);
        const uint32_t tacc = tmem_base + acc * C;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
#pragma unroll
            for (int k2 = 0; k2 < C / 16; ++k2)   // C input channels = C / 16 steps of K = 16 (32 bytes)
              umma_bf16(tacc, umma_desc_rows<ROWB>(sa + kh * C3_BOX + kw * ROWB + k2 * 32),
                        umma_desc_rows<ROWB>(wa + (kh * 3 + kw) * K::TAPB + k2 * 32), idesc, (kh | kw | k2) ? 1u : 0u);
        umma_commit(&empty[s]);
        umma_commit(&t_full[acc]);
      }
    }
  } else {
    const int set = (warp - 2) >> 2;        // tiles with (t & 1) == set
    const int quad = warp & 3;              // TMEM lane quarter this warp may read
    const int row = quad * 32 + lane;       // accumulator row = pixel within the tile
    const long long Wp = a.W + 2;
    uint32_t t = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
      if ((int)(t & 1) != set) continue;
      const int wt = tile % tpr, bh = tile / tpr;
      const int wo = wt * 128 + row;
      const bool valid = wo < a.W;
      const long long pix = ((long long)bh * Wp + wo + 1) * C;   // element offset of this pixel's C channels
      constexpr int NQ = C / 8;                                   // 16-byte pieces per pixel row
      uint4 rq[NQ];
      if (a.res != nullptr && valid) {
        const uint4* rp = reinterpret_cast<const uint4*>(a.res + pix);
#pragma unroll
        for (int i = 0; i < NQ; ++i) rq[i] = rp[i];
      }
      const uint32_t acc = t & 3;
      mbar_wait(&t_full[acc], (t >> 2) & 1);
      tc_fence_after();
      uint32_t r[C];
#pragma unroll
      for (int g = 0; g < C / 32; ++g) tmem_ld_32x32(tmem_base + acc * C + g * 32 + ((uint32_t)(quad * 32) << 16), r + 32 * g);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acc]);
      if (valid) {
        uint4* op = reinterpret_cast<uint4*>(a.out + pix);
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
          uint32_t w4[4];
          const uint32_t rr[4] = {rq[i].x, rq[i].y, rq[i].z, rq[i].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = 8 * i + 2 * e;
            float v0 = __uint_as_float(r[c]) + sbias[c], v1 = __uint_as_float(r[c + 1]) + sbias[c + 1];
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, K::TMEM);
  }
}

struct Conv3Plan {
  Conv3Args a;
  Conv3Maps maps;
  size_t smem;
  int grid;
};

Conv3Plan* conv3x3_c32_plan_create(const Conv3Args& a) {
  if (a.C != 32 && a.C != 64) return nullptr;
  Conv3Plan* p = new Conv3Plan();
  p->a = a;
  const int C = a.C, sw = 2 * C;
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)a.W + 2, (uint64_t)a.H, (uint64_t)a.B};
    uint64_t str[4] = {1, (uint64_t)C, (uint64_t)(a.W + 2) * C, (uint64_t)a.H * (a.W + 2) * C};
    uint32_t box[4] = {(uint32_t)C, 130, 1, 1};
    if (!make_tmap_sw(&p->maps.in, a.in, 4, dims, str, box, sw)) { delete p; return nullptr; }
  }
  {
    uint64_t dims[2] = {(uint64_t)a.ldw, (uint64_t)C};
    uint64_t str[2] = {1, (uint64_t)a.ldw};
    uint32_t box[2] = {(uint32_t)C, (uint32_t)C};
    if (!make_tmap_sw(&p->maps.w, a.w, 2, dims, str, box, sw)) { delete p; return nullptr; }
  }
  const size_t wbytes = C == 32 ? C3Cfg<32>::WBYTES : C3Cfg<64>::WBYTES;
  const size_t stages = C == 32 ? (size_t)C3Cfg<32>::NST * C3Cfg<32>::STAGE : (size_t)C3Cfg<64>::NST * C3Cfg<64>::STAGE;
  p->smem = 1024 + wbytes + stages + 8 * (1 + 2 * 3 + 8) + 8 + 64 * 4 + 16;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long tiles = (long long)a.B * a.H * ((a.W + 127) / 128);
  const long long slots = (long long)sms * (C == 32 ? 2 : 1);
  p->grid = (int)(tiles < slots ? tiles : slots);
  return p;
}
void conv3x3_c32_plan_destroy(Conv3Plan* p) { delete p; }

cudaError_t conv3x3_c32_plan_launch(const Conv3Plan* p, cudaStream_t st) {
  static size_t attr32 = 0, attr64 = 0;
  size_t& attr = p->a.C == 32 ? attr32 : attr64;
  if (p->smem > attr) {
    cudaError_t e;
    if (p->a.C == 32) {
      e = cudaFuncSetAttribute(conv3x3_kernel<32, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(conv3x3_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem);
    } else {
      e = cudaFuncSetAttribute(conv3x3_kernel<64, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(conv3x3_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem);
    }
    if (e != cudaSuccess) return e;
    attr = p->smem;
  }
  const int variant = (p->a.C == 64 ? 2 : 0) | (p->a.fp16 ? 1 : 0);
  switch (variant) {
    case 0: conv3x3_kernel<32, 0><<<p->grid, C3_THREADS, p->smem, st>>>(p->maps, p->a); break;
    case 1: conv3x3_kernel<32, 1><<<p->grid, C3_THREADS, p->smem, st>>>(p->maps, p->a); break;
    case 2: conv3x3_kernel<64, 0><<<p->grid, C3_THREADS, p->smem, st>>>(p->maps, p->a); break;
    default: conv3x3_kernel<64, 1><<<p->grid, C3_THREADS, p->smem, st>>>(p->maps, p->a); break;
  }
  return cudaGetLastError();
}

}  // namespace dz
