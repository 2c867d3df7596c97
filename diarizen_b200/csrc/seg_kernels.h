// Launch wrappers + argument blocks of the non-GEMM kernels of the segmentation path (seg_kernels.cu,
// attention_tc.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dz {

struct Conv0Args {
  const float* wav; int N; int T0; int C0; int C0p64;  // C0 rounded up to a multiple of 64
  const float* w;        // [C0][10]
  const float* wstats;   // large: [B][2] (mean, rstd) of the waveform
  const float* coef;     // base: [B][C0][2] GroupNorm scale/shift
  const float* gamma; const float* beta;  // large: LayerNorm(C0) affine
  __nv_bfloat16* out; long long out_plane; long long out_bstride; int ldo; int planes; int fp16;
};

struct LnArgs {
  const float* x; long long rows; int C; int ldx;
  const float* prescale; const float* gamma; const float* beta; int act;
  float* y_f32; int ldy;
  __nv_bfloat16* y_bf; long long bf_plane; int ldb; int planes;
  float* mix; float mix_w; int mix_src; /*0 none, 1 input x, 2 output y*/ int mix_init;
  int fp16;
};

struct GateArgs {
  const __nv_bfloat16* x; long long x_plane; int planes; long long rows; int ldx; int seq_len;
  const float* wab;   // [128]: summed rows 0..3 then rows 4..7 of gru_rel_pos_linear.weight
  float ba, bb;       // summed biases
  const float* gconst;      // [total_heads]
  const int* head_index;    // [nheads] device
  int nheads;
  float* gate;        // [B][nheads][T]
  int fp16;
};

struct AttnArgs {   // layout mirrors dz_attn_args (include/diarizen_b200.h)
  int T; int nheads;
  const __nv_bfloat16* q; const __nv_bfloat16* k; long long qk_plane; int ldqk; int q_col; int k_col; int fp16;
  const __nv_bfloat16* vt; long long vt_plane; int ldvt;   // [B][nheads*64][ldvt]
  int planes;
  const float* bias_tab;   // [nheads][2T-1] or null
  const float* gate;       // [B][nheads][T] or null
  __nv_bfloat16* out; long long out_plane; int ldo; int out_planes;  // [B*T][ldo]
  const __nv_bfloat16* v; int v_col; int _pad;   // optional row-major V [B*T][ldqk] (planes qk_plane apart); replaces vt when set
};

struct DwArgs {
  const float* x; int ldx; int T; int A; int ksize;
  const float* w;      // [A][ksize]
  const float* scale; const float* shift;  // folded BatchNorm (+ conv bias)
  __nv_bfloat16* out; long long out_plane; int ldo; int planes; int fp16;
};

struct HeadArgs {
  const float* x; long long rows; int ldx; int A; int NC;
  const float* w; const float* bias;
  float* logp; uint8_t* multilabel;
};

size_t wave_stats_scratch_bytes(int B);
cudaError_t launch_wave_stats(const float* wav, int B, int N, float* stats, void* scratch, cudaStream_t st);   // scratch: zeroed once
cudaError_t launch_conv0_moments(const float* wav, int B, int N, int T0, double* mom, cudaStream_t st);
cudaError_t launch_conv0_gn_coef(const double* mom, const float* w, const float* gamma, const float* beta, int B, int C0,
                                 int T0, float* coef, cudaStream_t st);
cudaError_t launch_conv0(const Conv0Args& a, int B, bool large, cudaStream_t st);
bool conv0_tc_eligible(const Conv0Args& a);                                         // conv0_tc.cu (tcgen05 variant, fp16 mode)
cudaError_t launch_conv0_tc(const Conv0Args& a, int B, bool large, cudaStream_t st);
cudaError_t launch_layernorm(const LnArgs& a, cudaStream_t st);
cudaError_t launch_axpy_mix(const float* x, float* mix, float w, int init, long long n, cudaStream_t st);
cudaError_t launch_regroup(const float* x, long long rows, int C, int ldx, int seq_len, int seq_rows_out, int row_off,
                           int gin, int gout, __nv_bfloat16* out, long long out_plane, int ldo, int planes,
                           int fp16, cudaStream_t st);
cudaError_t launch_gate(const GateArgs& a, cudaStream_t st);
cudaError_t launch_attention_simt(const AttnArgs& a, int B, cudaStream_t st);
cudaError_t launch_attention_tc(const AttnArgs& a, int B, cudaStream_t st);
// positional convolution + bias + GELU + residual, in place on the fp32 residual stream (posconv_tc.cu)
struct PosConvArgs {
  const __nv_bfloat16* stage; int stage_rows; int stage_ld;   // [B][stage_rows = T + 128][16 x 64] zero-padded 16-bit copy
  const __nv_bfloat16* w; int ldw; long long w_gstride;        // [16][Dg][128 taps x 64]
  const float* bias;                                           // [16 x Dg]
  float* x; int ldx;                                           // [B*T][ldx] residual stream, updated in place
  int B; int T; int Dg; int fp16;
};
struct PosConvPlan;
PosConvPlan* posconv_plan_create(const PosConvArgs& a);
void posconv_plan_destroy(PosConvPlan* p);
cudaError_t posconv_plan_launch(const PosConvPlan* p, cudaStream_t st);
struct AttnPlan;  // tensor maps + launch geometry of the tcgen05 attention kernel (attention_tc.cu)
AttnPlan* attention_tc_plan_create(const AttnArgs& a, int B);
void attention_tc_plan_destroy(AttnPlan* p);
cudaError_t attention_tc_plan_launch(const AttnPlan* p, cudaStream_t st);
cudaError_t launch_glu_dwconv(const DwArgs& a, int B, cudaStream_t st);
cudaError_t launch_classifier_head(const HeadArgs& a, cudaStream_t st);

}  // namespace dz
