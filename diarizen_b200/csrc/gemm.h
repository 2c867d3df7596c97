// Descriptor of one "C = epilogue(A * B^T)" launch shared by the tcgen05 kernel (gemm_tc.cu) and the
// CUDA-core checker kernel (gemm_simt.cu).  Operands are bf16 *planes*: plane 0 = rn(x) ("hi"),
// plane 1 = rn(x - hi) ("lo").  npass = 1 multiplies the hi planes only (bf16 mode); npass = 3
// accumulates hi*hi + lo*hi + hi*lo into the same fp32 accumulator (fp32-class mode, error ~2^-17).
//
// The A operand is described by strides so that convolutions become GEMMs without an im2col copy:
//   element (batch b, group g, row m, k) lives at
//     a + b*a_bstride + g*a_gstride + m*a_rstride + (k / a_kinner)*a_kouter + (k % a_kinner)
//   * linear layer      : a_rstride = ld, a_kinner = K
//   * conv1d k, stride 2: channels-last input (T_in, C); a_rstride = 2*C, K = k*C  (rows overlap)
//   * grouped pos-conv  : a_kinner = 64 (channels of one group, padded), a_kouter = row pitch (one tap),
//                         a_rstride = row pitch, a_gstride = 64
// The same description is turned into a TMA tensor map (rank 3 or 5) for the tcgen05 path.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/diarizen_b200.h"

namespace dz {

typedef dz_gemm_desc GemmDesc;  // field meanings below; layout is the public C struct

// M: valid rows per batch; N: valid output columns per group; K: reduction length per pass;
// npass: 1 (bf16) or 3 (bf16x3); a_plane / b_plane: elements between hi and lo planes;
// a_rows_alloc: rows addressable per batch (TMA bound), >= M (0 -> M);
// epilogue: v = alpha * act(acc + bias[g*group_cols + n]) (+ residual[b][m][col]);
// outputs: out_f32[b][m][col]; out_bf planes [b][m + out_row_off][col] (zero_pad_to: also zero columns
// [N, zero_pad_to)); out_t: transposed planes for columns >= tr_col0: [m / seq_len][col - tr_col0][m % seq_len].
inline GemmDesc gemm_desc_default() {
  GemmDesc d;
  memset(&d, 0, sizeof(d));
  d.npass = 1; d.batches = 1; d.groups = 1; d.alpha = 1.0f; d.out_planes = 1; d.seq_len = 1;
  return d;
}

// Host API (C++): plan = tensor maps + launch geometry, built once per shape and replayed.
struct GemmPlan;
GemmPlan* gemm_plan_create(const GemmDesc& d, int force_bn /*0 = auto*/);
void gemm_plan_destroy(GemmPlan* p);
cudaError_t gemm_plan_launch(const GemmPlan* p, cudaStream_t stream);
const GemmDesc& gemm_plan_desc(const GemmPlan* p);
// CUDA-core implementation of the same contract (checker / tiny shapes).
cudaError_t gemm_simt_launch(const GemmDesc& d, cudaStream_t stream);
const char* gemm_last_error();

}  // namespace dz
