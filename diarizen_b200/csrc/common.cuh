// Shared device helpers for the sm_100a kernels: mbarrier / TMA / tcgen05 PTX wrappers,
// bf16 hi/lo plane packing, warp reductions.  Everything here is inline PTX for sm_100a;
// there is no fallback path for other architectures.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace dz {

typedef __nv_bfloat16 bf16;

#define DZ_DEVINL __device__ __forceinline__

DZ_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
DZ_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
DZ_DEVINL void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
DZ_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DZ_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DZ_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  // the last operand is the suspend-time hint (ns): the warp may sleep in hardware until the phase flips instead of
  // re-issuing the poll, which keeps the spin loops of the producer / issuer warps out of the math warps' issue slots
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must become a trap (an error the host sees), never a hung GPU.  The clock is sampled only
// every 4096 polls so that the common path is poll + branch.
DZ_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t n = 0;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if ((++n & 0x3Fu) == 0) {
      if (clock64() - t0 > 4000000000LL) {   // ~2 s: three orders of magnitude above any legitimate wait on this path
        printf("dz: mbarrier timeout block(%d,%d,%d) thread %d barrier@%u parity %u\n", blockIdx.x, blockIdx.y, blockIdx.z,
               threadIdx.x, smem_u32(bar), parity);
        __trap();
      }
    }
  }
}
// dynamic shared memory rounded up to the 1024-byte alignment the 128-byte swizzle atoms need.  Pointer arithmetic (not an
// integer round trip) so that the compiler keeps the shared address space and emits LDS/STS instead of generic LD/ST.
DZ_DEVINL uint8_t* smem_align1024(uint8_t* p) { return p + ((1024u - (smem_u32(p) & 1023u)) & 1023u); }
DZ_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), tile mode, completion on an mbarrier
// ----------------------------------------------------------------------------------------------
DZ_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
DZ_DEVINL void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
DZ_DEVINL void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
DZ_DEVINL void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
DZ_DEVINL void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// TMA store (shared -> global), bulk async-group completion
DZ_DEVINL void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(m),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
DZ_DEVINL void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
DZ_DEVINL void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
DZ_DEVINL void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// ----------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, UMMA issue, commit, TMEM loads
// ----------------------------------------------------------------------------------------------
DZ_DEVINL void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // whole warp, ncols = 2^k >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
DZ_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
DZ_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
DZ_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows of 128 B, 8-row groups
// 1024 B apart (SBO), descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
DZ_DEVINL uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address, 16-byte units
  d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                       // descriptor version
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}
// MN-major operand (the N or M index is the contiguous one), 128-byte swizzle: one row per K index holding 64 MN elements
// (128 B); 8-row K groups `SBO` = 1024 B apart; the next block of 64 MN elements `lbo_bytes` after the first.
DZ_DEVINL uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor for kind::f16 with bf16 A/B (K-major both), fp32 accumulate.
DZ_DEVINL uint32_t umma_idesc_bf16(uint32_t m, uint32_t n, int fp16 = 0) {
  uint32_t d = 0;
  d |= 1u << 4;                    // D format = F32
  d |= (fp16 ? 0u : 1u) << 7;      // A format: 0 = F16, 1 = BF16
  d |= (fp16 ? 0u : 1u) << 10;     // B format
  d |= (n >> 3) << 17;   // N
  d |= (m >> 4) << 24;   // M
  return d;
}
// D[tmem] (+)= A[smem] * B[smem]^T ; single thread issues.
DZ_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued UMMAs of this thread have completed.
DZ_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane_base + i).
DZ_DEVINL void tmem_ld_32x32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
DZ_DEVINL void tmem_ld_32x32_x16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
DZ_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
DZ_DEVINL void tmem_st_32x32_x1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
DZ_DEVINL void tmem_st_32x32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
DZ_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// numerics
// ----------------------------------------------------------------------------------------------
// GELU (exact-erf form).  erf through Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, i.e. below fp32 resolution of the
// result for |x| < ~4 and far below every tolerance in this path): one MUFU.RCP + one MUFU.EX2 instead of the ~25
// instruction libdevice erff - the activation sits on the critical path of the conv0 kernel and the FFN epilogues.
DZ_DEVINL float erf_as(float z) {
  const float a = fabsf(z);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, a, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.0f - p * t * __expf(-a * a);
  return copysignf(e, z);
}
// gelu(x) = x * Phi(x) with the same A-S polynomial folded through: Phi(x) = h for x < 0 and 1 - h for x >= 0, where
// h = 0.5 * P(t) * exp(-x^2 / 2), t = 1 / (1 + 0.3275911 |x| / sqrt(2)).  12 FP32 ops + RCP + EX2 per element.
DZ_DEVINL float gelu_erf(float x) {
  float t, e;
  const float d = fmaf(0.23164189f, fabsf(x), 1.0f);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(d));
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  const float xx = x * x;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(xx * -0.72134752044448170368f));   // exp(-x^2 / 2)
  const float xh = x * (p * t * e);
  return x > 0.0f ? x - xh : xh;
}
DZ_DEVINL float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
DZ_DEVINL float apply_act(float x, int act) {
  switch (act) {
    case 1: return gelu_erf(x);
    case 2: return x / (1.0f + expf(-x));  // swish
    case 3: return fmaxf(x, 0.0f);
    default: return x;
  }
}
// 16-bit operand planes.  The storage type is `bf16` (a 16-bit container); `fp16` selects how the bits are
// interpreted: 0 = bfloat16 (8-bit mantissa, fp32 range), 1 = IEEE half (11-bit mantissa, saturating at 65504).
DZ_DEVINL bf16 to16(float x, int fp16) {
  if (fp16) {
    x = fminf(fmaxf(x, -65504.f), 65504.f);
    return __ushort_as_bfloat16(__half_as_ushort(__float2half_rn(x)));
  }
  return __float2bfloat16_rn(x);
}
DZ_DEVINL float from16(bf16 h, int fp16) {
  return fp16 ? __half2float(__ushort_as_half(__bfloat16_as_ushort(h))) : __bfloat162float(h);
}
// fp32 -> (hi, lo) planes: hi = rn(x), lo = rn(x - hi); with bf16, hi + lo carries ~16 mantissa bits.
DZ_DEVINL void split_bf16(float x, bf16& hi, bf16& lo, int fp16) {
  hi = to16(x, fp16);
  lo = to16(x - from16(hi, fp16), fp16);
}
// two fp32 -> one packed pair of 16-bit operands (lo in the low half); the fp16 flavour saturates like to16()
template <int FP16>
DZ_DEVINL uint32_t pack2_16(float lo, float hi) {
  uint32_t r;
  if (FP16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
DZ_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
DZ_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace dz
