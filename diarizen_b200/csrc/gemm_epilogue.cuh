// Fused GEMM epilogue shared by gemm_tc.cu and gemm_simt.cu.  One thread owns one output row and a
// chunk of 32 consecutive accumulator columns (exactly what tcgen05.ld 32x32b.x32 hands a thread).
#pragma once
#include "common.cuh"
#include "gemm.h"

namespace dz {

// v[j] = accumulator of column n0 + j (n0 % 32 == 0) of row m in (batch b, group g).
DZ_DEVINL void gemm_epilogue_chunk(const GemmDesc& d, int b, int g, int m, int n0, float (&v)[32]) {
  const int gcol0 = g * d.group_cols + n0;
  const int nvalid = d.N - n0;  // may be <= 0 or >= 32
  // ---- bias / activation / scale ----
  if (d.bias != nullptr) {
    if (nvalid >= 32) {
      const float4* bp = reinterpret_cast<const float4*>(d.bias + gcol0);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 t = __ldg(bp + q);
        v[4 * q + 0] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) v[j] += __ldg(d.bias + gcol0 + j);
    }
  }
  if (d.act_after_res) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = d.alpha * v[j];
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = d.alpha * apply_act(v[j], d.act);
  }
  // ---- residual ----
  if (d.res16 != nullptr) {
    const bf16* rp = (const bf16*)d.res16 + (long long)b * d.res16_bstride + (long long)(m + d.res16_row_off) * d.ldr16 + gcol0;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < nvalid) {
        v[j] += from16(rp[j], d.fp16);
        if (d.out_planes > 1) v[j] += from16(rp[d.res16_plane + j], d.fp16);
      }
  }
  if (d.residual != nullptr) {
    const float* rp = d.residual + (long long)b * d.res_bstride + (long long)m * d.ldr + gcol0;
    if (nvalid >= 32) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 t = *reinterpret_cast<const float4*>(rp + 4 * q);
        v[4 * q + 0] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) v[j] += rp[j];
    }
  }
  if (d.act_after_res) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], d.act);
  }
  // ---- fp32 row-major output ----
  if (d.out_f32 != nullptr) {
    float* op = d.out_f32 + (long long)b * d.of_bstride + (long long)m * d.ldo + gcol0;
    if (nvalid >= 32) {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(op + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) op[j] = v[j];
    }
  }
  // ---- bf16 plane outputs ----
  const int rm_cols = (d.out_t != nullptr) ? d.tr_col0 : d.N;  // columns that go to the row-major planes
  if (d.out_bf != nullptr && n0 < max(rm_cols, d.zero_pad_to)) {
    bf16* hp = (bf16*)d.out_bf + (long long)b * d.ob_bstride + (long long)(m + d.out_row_off) * d.ldob + gcol0;
    if (n0 + 32 <= rm_cols) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          bf16 h0, l0, h1, l1;
          split_bf16(v[8 * q + 2 * e], h0, l0, d.fp16);
          split_bf16(v[8 * q + 2 * e + 1], h1, l1, d.fp16);
          hw[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          lw[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
        *reinterpret_cast<uint4*>(hp + 8 * q) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        if (d.out_planes > 1) *reinterpret_cast<uint4*>(hp + d.ob_plane + 8 * q) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int c = n0 + j;
        if (c < rm_cols) {
          bf16 h, l;
          split_bf16(v[j], h, l, d.fp16);
          hp[j] = h;
          if (d.out_planes > 1) hp[d.ob_plane + j] = l;
        } else if (c < d.zero_pad_to && d.out_t == nullptr) {
          hp[j] = __float2bfloat16_rn(0.0f);
          if (d.out_planes > 1) hp[d.ob_plane + j] = __float2bfloat16_rn(0.0f);
        }
      }
    }
  }
  // ---- transposed bf16 output (e.g. V^T for the P*V product) ----
  if (d.out_t != nullptr && n0 + 32 > d.tr_col0 && nvalid > 0) {
    const int sb = m / d.seq_len, st = m - sb * d.seq_len;
    bf16* tp = (bf16*)d.out_t + (long long)sb * d.ot_bstride + st;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int c = gcol0 + j;
      if (c >= d.tr_col0 && j < nvalid) {
        bf16 h, l;
        split_bf16(v[j], h, l, d.fp16);
        bf16* q = tp + (long long)(c - d.tr_col0) * d.ldt;
        *q = h;
        if (d.out_planes > 1) q[d.ot_plane] = l;
      }
    }
  }
}

}  // namespace dz
