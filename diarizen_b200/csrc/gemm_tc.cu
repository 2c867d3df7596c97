// tcgen05 GEMM for sm_100a:  C = epilogue(A * B^T), bf16 operands (hi/lo planes), fp32 accumulation in TMEM.
//
// One CTA computes a 128 x BN output tile.  Warp roles:
//   warp 0      : TMA producer (one elected lane): A box 128 rows x 64 k, B box BN rows x 64 k, SWIZZLE_128B
//   warp 1      : TMEM allocator + UMMA issuer (one elected lane): 4 x tcgen05.mma (K = 16) per 64-wide k block
//   warps 2..5  : epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> fused bias/act/residual -> HBM
// Pipeline: NSTAGE-deep smem ring with full/empty mbarriers; tcgen05.commit releases a stage back to the
// producer and finally signals the epilogue.  With npass = 3 the k loop runs three times over
// (A_hi,B_hi), (A_lo,B_hi), (A_hi,B_lo) into the same accumulator.
//
// This kernel replaces the cuBLAS/cuDNN calls behind nn.Linear / nn.Conv1d on the reference hot path
// (reference: diarizen/models/module/wav2vec2/components.py:119 conv1d, :305-306 projection, :374 pos-conv,
//  :455-480 q/k/v/out projections, :805-814 FFN; diarizen/models/module/conformer.py:116-214).
#include <mutex>
#include <string>

#include "common.cuh"
#include "gemm.h"
#include "gemm_epilogue.cuh"

namespace dz {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int NUM_THREADS = 192;

template <int BN>
struct TcCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NSTAGE = (BN == 256) ? 4 : (BN == 128 ? 3 : 4);
  static constexpr int SMEM = NSTAGE * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct TcMaps {
  CUtensorMap a[2];  // hi, lo
  CUtensorMap b[2];
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ TcMaps maps, const GemmDesc d, const int a_rank5) {
  using C = TcCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::NSTAGE * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::NSTAGE;
  uint64_t* tmem_full = empty_bar + C::NSTAGE;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;
  const int g = (d.groups > 1) ? (int)blockIdx.y : 0;
  const int n0 = (d.groups > 1) ? 0 : (int)blockIdx.y * BN;
  const int b = blockIdx.z;
  const int kblocks = (d.K + BK - 1) / BK;
  const int iters = kblocks * d.npass;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::NSTAGE; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.b[0]);
  }
  if (warp == 1) tmem_alloc(tmem_ptr, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const int s = it % C::NSTAGE;
        const uint32_t ph = (it / C::NSTAGE) & 1;
        const int pass = it / kblocks;
        const int kb = it - pass * kblocks;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * C::STAGE_BYTES;
        uint8_t* sb = sa + C::A_BYTES;
        mbar_expect_tx(&full_bar[s], C::STAGE_BYTES);
        const CUtensorMap* ma = &maps.a[pass == 1 ? 1 : 0];
        const CUtensorMap* mb = &maps.b[pass == 2 ? 1 : 0];
        if (a_rank5) {
          // (k_inner, k_outer, row, group, batch): one k block = one run of a_kinner (=64) elements
          tma_load_5d(sa, ma, &full_bar[s], 0, kb, m0, g, b);
        } else {
          tma_load_3d(sa, ma, &full_bar[s], kb * BK, m0, b);
        }
        tma_load_3d(sb, mb, &full_bar[s], kb * BK, n0, g);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const int n_valid = min(BN, d.N - n0);
      const uint32_t umma_n = (uint32_t)((n_valid + 15) & ~15);
      const uint32_t idesc = umma_idesc_bf16(BM, umma_n, d.fp16);
      for (int it = 0; it < iters; ++it) {
        const int s = it % C::NSTAGE;
        const uint32_t ph = (it / C::NSTAGE) & 1;
        const int kb = it % kblocks;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * C::STAGE_BYTES);
        const uint32_t sb = sa + C::A_BYTES;
        const int krem = d.K - kb * BK;
        const int ksteps = krem >= BK ? (BK / 16) : ((krem + 15) / 16);
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t adesc = umma_desc_sw128(sa + k * 32);
          const uint64_t bdesc = umma_desc_sw128(sb + k * 32);
          umma_bf16(tmem_base, adesc, bdesc, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tmem_full);
    }
  } else {
    // epilogue: warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32)
    const int quad = warp & 3;
    const int m = m0 + quad * 32 + lane;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int n_valid = min(BN, d.N - n0);
    const int n_store = max(n_valid, min(BN, d.zero_pad_to - n0));
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      if (c >= n_store) break;
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c, r);
      tmem_ld_wait();
      if (m < d.M) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        gemm_epilogue_chunk(d, b, g, m, n0 + c, v);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static thread_local std::string g_err;
const char* gemm_last_error() { return g_err.c_str(); }

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

bool make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                    const uint32_t* box) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { g_err = "cuTensorMapEncodeTiled entry point unavailable"; return false; }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_elems[i] * 2;
      if (gstr[i - 1] % 16 != 0) { g_err = "tensor map stride not a multiple of 16 bytes"; return false; }
    }
  }
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0) { g_err = "tensor map base not 16-byte aligned"; return false; }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_err = "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r);
    return false;
  }
  return true;
}

struct GemmPlan {
  GemmDesc d;
  TcMaps maps;
  int bn = 128;
  int rank5 = 0;
  dim3 grid;
};

template <int BN>
static cudaError_t launch_bn(const GemmPlan* p, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN>::SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  gemm_tc_kernel<BN><<<p->grid, NUM_THREADS, TcCfg<BN>::SMEM, st>>>(p->maps, p->d, p->rank5);
  return cudaGetLastError();
}

static int choose_bn(const GemmDesc& d) {
  if (d.groups > 1 || d.N <= 64) return 64;
  if (d.N <= 128) return 128;
  // prefer the wider tile (less smem traffic per MAC) unless it leaves a mostly empty tail tile
  const int t256 = (d.N + 255) / 256, t128 = (d.N + 127) / 128;
  const long long pad256 = (long long)t256 * 256 - d.N, pad128 = (long long)t128 * 128 - d.N;
  (void)pad128;
  return (pad256 >= 128) ? 128 : 256;
}

GemmPlan* gemm_plan_create(const GemmDesc& d, int force_bn) {
  GemmPlan* p = new GemmPlan();
  p->d = d;
  p->bn = force_bn ? force_bn : choose_bn(d);
  if (d.groups > 1 && d.N > p->bn) { g_err = "grouped GEMM needs N <= BN"; delete p; return nullptr; }
  if (d.npass != 1 && d.npass != 3) { g_err = "npass must be 1 or 3"; delete p; return nullptr; }
  p->rank5 = (d.a_kinner != d.K) ? 1 : 0;
  const long long rows_alloc = d.a_rows_alloc > 0 ? d.a_rows_alloc : d.M;
  for (int pl = 0; pl < 2; ++pl) {
    const __nv_bfloat16* abase = (const __nv_bfloat16*)d.a + (pl ? d.a_plane : 0);
    const __nv_bfloat16* bbase = (const __nv_bfloat16*)d.b + (pl ? d.b_plane : 0);
    if (pl == 1 && d.npass == 1) { abase = (const __nv_bfloat16*)d.a; bbase = (const __nv_bfloat16*)d.b; }
    bool ok;
    if (p->rank5) {
      if (d.a_kinner != BK) { g_err = "rank-5 A operand needs a_kinner == 64"; delete p; return nullptr; }
      uint64_t dims[5] = {(uint64_t)d.a_kinner, (uint64_t)(d.K / d.a_kinner), (uint64_t)rows_alloc,
                          (uint64_t)d.groups, (uint64_t)d.batches};
      uint64_t str[5] = {1, (uint64_t)d.a_kouter, (uint64_t)d.a_rstride, (uint64_t)d.a_gstride, (uint64_t)d.a_bstride};
      uint32_t box[5] = {BK, 1, BM, 1, 1};
      // rows_alloc bounds row + tap: the view is (tap, row) -> row+tap; expose rows so that row+tap stays in the batch
      ok = make_tmap_bf16(&p->maps.a[pl], abase, 5, dims, str, box);
    } else {
      uint64_t dims[3] = {(uint64_t)d.K, (uint64_t)rows_alloc, (uint64_t)d.batches};
      uint64_t str[3] = {1, (uint64_t)d.a_rstride, (uint64_t)(d.batches > 1 ? d.a_bstride : d.a_rstride * rows_alloc)};
      uint32_t box[3] = {BK, BM, 1};
      ok = make_tmap_bf16(&p->maps.a[pl], abase, 3, dims, str, box);
    }
    if (!ok) { delete p; return nullptr; }
    uint64_t bdims[3] = {(uint64_t)d.K, (uint64_t)d.N, (uint64_t)d.groups};
    uint64_t bstr[3] = {1, (uint64_t)d.ldb, (uint64_t)(d.groups > 1 ? d.b_gstride : (long long)d.ldb * d.N)};
    uint32_t bbox[3] = {BK, (uint32_t)p->bn, 1};
    if (!make_tmap_bf16(&p->maps.b[pl], bbase, 3, bdims, bstr, bbox)) { delete p; return nullptr; }
  }
  const int mt = (d.M + BM - 1) / BM;
  const int nt = d.groups > 1 ? d.groups : (d.N + p->bn - 1) / p->bn;
  p->grid = dim3(mt, nt, d.batches);
  return p;
}

void gemm_plan_destroy(GemmPlan* p) { delete p; }
const GemmDesc& gemm_plan_desc(const GemmPlan* p) { return p->d; }

cudaError_t gemm_plan_launch(const GemmPlan* p, cudaStream_t st) {
  switch (p->bn) {
    case 64: return launch_bn<64>(p, st);
    case 128: return launch_bn<128>(p, st);
    case 256: return launch_bn<256>(p, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace dz
