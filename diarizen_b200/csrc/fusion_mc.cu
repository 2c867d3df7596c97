// Cross-channel fusion of the multi-channel segmentation model.
// reference: diarizen/models/module/utils_mc.py:13-64 (CrossChannelAttention), applied between the first WavLM layers by
// Transformer.get_intermediate_outputs_mc (diarizen/models/module/wav2vec2/components.py:1026-1070):
//   x (B, C, T, D);  q/k/v = Linear(D -> hu)(x);  per (b, t) and head: softmax over channels of q k^T / sqrt(dk), times v;
//   x <- LayerNorm(Linear(hu -> D)(ctx)) + x;   att = the (B*T, heads, C, C) attention weights (the recipe averages them).
// Rows of x are ordered (b, c, t) - the engine runs the B*C channel signals as B*C windows.
// The two projections are tcgen05 GEMMs (gemm_tc.cu) on 16-bit planes of x; the C x C attention over channels (C <= 8) and
// the LayerNorm + residual are warp-per-row kernels.  dz_fusion_* is the C ABI the Python MC model drives.
#include <map>
#include <string>
#include <vector>

#include "../../include/diarizen_b200.h"
#include "common.cuh"
#include "engine_common.h"
#include "gemm.h"
#include "seg_kernels.h"

namespace dz {
std::string& tls_error();
int fail(int code, const std::string& msg);

// one warp per (b, t): q, k, v planes [R][3*hu] (q | k | v), rows (b*C + c)*T + t.  ctx planes [R][hu]; att [B*T][C][C] = mean over heads.
__global__ void __launch_bounds__(256) cross_channel_attention_kernel(const bf16* __restrict__ qkv, long long plane, int planes, int fp16,
                                                                      int B, int C, int T, int hu, int heads, bf16* __restrict__ ctx,
                                                                      long long ctx_plane, float* __restrict__ att) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long bt = (long long)blockIdx.x * 8 + warp;
  if (bt >= (long long)B * T) return;
  const int b = (int)(bt / T), t = (int)(bt - (long long)b * T);
  const int dk = hu / heads;
  const float scale = rsqrtf((float)dk);
  const int ld = 3 * hu;
  auto val = [&](long long row, int col) -> float {
    float v = from16(qkv[row * ld + col], fp16);
    if (planes > 1) v += from16(qkv[plane + row * ld + col], fp16);
    return v;
  };
  float amean[64];   // C * C <= 64, lane 0 accumulates the head mean
#pragma unroll
  for (int i = 0; i < 64; ++i) amean[i] = 0.f;
  for (int h = 0; h < heads; ++h) {
    // scores s[i][j] = q_i . k_j * scale; lanes split the dk dimension
    float s[64];
    for (int i = 0; i < C; ++i)
      for (int j = 0; j < C; ++j) {
        const long long ri = ((long long)b * C + i) * T + t, rj = ((long long)b * C + j) * T + t;
        float acc = 0.f;
        for (int d = lane; d < dk; d += 32) acc += val(ri, h * dk + d) * val(rj, hu + h * dk + d);
        s[i * C + j] = warp_sum(acc) * scale;
      }
    for (int i = 0; i < C; ++i) {
      float mx = -INFINITY;
      for (int j = 0; j < C; ++j) mx = fmaxf(mx, s[i * C + j]);
      float se = 0.f;
      for (int j = 0; j < C; ++j) { s[i * C + j] = expf(s[i * C + j] - mx); se += s[i * C + j]; }
      const float inv = 1.0f / se;
      for (int j = 0; j < C; ++j) { s[i * C + j] *= inv; amean[i * C + j] += s[i * C + j]; }
      // ctx_i = sum_j p_ij v_j
      const long long ri = ((long long)b * C + i) * T + t;
      for (int d = lane; d < dk; d += 32) {
        float acc = 0.f;
        for (int j = 0; j < C; ++j) acc += s[i * C + j] * val(((long long)b * C + j) * T + t, 2 * hu + h * dk + d);
        bf16 hi, lo;
        split_bf16(acc, hi, lo, fp16);
        ctx[ri * hu + h * dk + d] = hi;
        if (planes > 1) ctx[ctx_plane + ri * hu + h * dk + d] = lo;
      }
    }
  }
  if (lane == 0 && att != nullptr) {
    const float ih = 1.0f / (float)heads;
    for (int i = 0; i < C * C; ++i) att[bt * C * C + i] = amean[i] * ih;
  }
}

// x[r] += LayerNorm(o[r]) (gamma, beta); optionally refresh the 16-bit planes of x and add w * LayerNorm(o[r]) to mix.  One warp per row.
__global__ void __launch_bounds__(256) ln_residual_kernel(const float* __restrict__ o, int ldo, float* __restrict__ x, int ldx, long long rows, int D,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta, bf16* __restrict__ xbf,
                                                          long long xbf_plane, int ldb, int planes, int fp16, float* __restrict__ mix, float mix_w) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + warp;
  if (r >= rows) return;
  const float* orow = o + r * ldo;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) s += orow[c];
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
  for (int c = lane; c < D; c += 32) { const float d = orow[c] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + 1e-5f);
  for (int c = lane; c < D; c += 32) {
    const float y = (orow[c] - mean) * rstd * gamma[c] + beta[c];
    const float v = x[r * ldx + c] + y;
    x[r * ldx + c] = v;
    if (xbf != nullptr) {
      bf16 hi, lo;
      split_bf16(v, hi, lo, fp16);
      xbf[r * ldb + c] = hi;
      if (planes > 1) xbf[xbf_plane + r * ldb + c] = lo;
    }
    if (mix != nullptr) mix[r * ldx + c] += mix_w * y;
  }
}

// out[(b, t)] = mean over c of in[(b, c, t)]  (fp32 rows of width D, leading dimension ld)
__global__ void channel_mean_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int C, int T, int D, int ld) {
  const long long total = (long long)B * T * D;
  const float inv = 1.0f / (float)C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const long long bt = i / D;
    const int b = (int)(bt / T), t = (int)(bt - (long long)b * T);
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += in[(((long long)b * C + c) * T + t) * ld + d];
    out[bt * ld + d] = s * inv;
  }
}

}  // namespace dz

using namespace dz;

struct dz_fusion {
  int D = 0, hu = 0, heads = 0, precision = 2, planes = 1, npass = 1, fp16 = 1;
  std::map<std::string, std::vector<float>> params;
  bool finalized = false;
  Weight wqkv, wo;
  DevMem gamma, beta;
  // workspace for the current (rows) shape
  long long rows = 0;
  DevMem xpl, qkv, ctx, o;
  GemmPlan* g1 = nullptr;
  GemmPlan* g2 = nullptr;
  ~dz_fusion() { if (g1) gemm_plan_destroy(g1); if (g2) gemm_plan_destroy(g2); }
};

extern "C" {

dz_fusion* dz_fusion_create(int D, int hidden, int heads, int precision) {
  if (D < 8 || hidden < heads || heads < 1 || hidden % heads != 0 || (precision != 1 && precision != 2 && precision != 3)) {
    fail(DZ_ERR_INVALID, "bad fusion configuration");
    return nullptr;
  }
  dz_fusion* f = new dz_fusion();
  f->D = D; f->hu = hidden; f->heads = heads; f->precision = precision;
  f->planes = precision == 3 ? 2 : 1; f->npass = precision == 3 ? 3 : 1; f->fp16 = precision == 2 ? 1 : 0;
  return f;
}
void dz_fusion_destroy(dz_fusion* f) { delete f; }

int dz_fusion_set_param(dz_fusion* f, const char* name, const float* host, int64_t n) {
  if (!f || !name || !host || n <= 0) return fail(DZ_ERR_INVALID, "bad argument");
  f->params[name].assign(host, host + n);
  return DZ_OK;
}

// expects linearQ/K/V/O .weight/.bias and ln_norm.weight/.bias (state-dict names of utils_mc.CrossChannelAttention)
int dz_fusion_finalize(dz_fusion* f) {
  if (!f) return fail(DZ_ERR_INVALID, "null handle");
  const int D = f->D, hu = f->hu;
  auto need = [&](const std::string& k, size_t n) -> const std::vector<float>* {
    auto it = f->params.find(k);
    if (it == f->params.end() || it->second.size() != n) { fail(DZ_ERR_INVALID, "fusion parameter '" + k + "' missing or of the wrong size"); return nullptr; }
    return &it->second;
  };
  const std::vector<float>*wq = need("linearQ.weight", (size_t)hu * D), *wk = need("linearK.weight", (size_t)hu * D), *wv = need("linearV.weight", (size_t)hu * D);
  const std::vector<float>*bq = need("linearQ.bias", hu), *bk = need("linearK.bias", hu), *bv = need("linearV.bias", hu);
  const std::vector<float>*wo = need("linearO.weight", (size_t)D * hu), *bo = need("linearO.bias", D);
  const std::vector<float>*g = need("ln_norm.weight", D), *be = need("ln_norm.bias", D);
  if (!wq || !wk || !wv || !bq || !bk || !bv || !wo || !bo || !g || !be) return DZ_ERR_INVALID;
  std::vector<float> w3((size_t)3 * hu * D), b3((size_t)3 * hu);
  std::copy(wq->begin(), wq->end(), w3.begin()); std::copy(wk->begin(), wk->end(), w3.begin() + (size_t)hu * D);
  std::copy(wv->begin(), wv->end(), w3.begin() + (size_t)2 * hu * D);
  std::copy(bq->begin(), bq->end(), b3.begin()); std::copy(bk->begin(), bk->end(), b3.begin() + hu); std::copy(bv->begin(), bv->end(), b3.begin() + 2 * hu);
  g_weight_fp16() = f->fp16;
  cudaError_t e = make_weight(f->wqkv, w3.data(), 1, 3 * hu, D, b3.data(), 3 * hu);
  if (e == cudaSuccess) e = make_weight(f->wo, wo->data(), 1, D, hu, bo->data(), D);
  if (e == cudaSuccess) e = upload_vec(f->gamma, *g);
  if (e == cudaSuccess) e = upload_vec(f->beta, *be);
  if (e != cudaSuccess) return fail(DZ_ERR_CUDA, std::string("fusion weights: ") + cudaGetErrorString(e));
  f->finalized = true;
  return DZ_OK;
}

/* x_dev: fp32 [B*C*T][ldx], rows (b, c, t), updated in place.  xbf_dev (optional): 16-bit planes [B*C*T][ldb] of x to refresh;
 * mix_dev (optional): fp32 [B*C*T][ldx], mix += mix_w * LayerNorm(o).  att_dev (optional): fp32 [B*T][C][C], mean over heads. */
int dz_fusion_forward(dz_fusion* f, float* x_dev, int B, int C, int T, int ldx, void* xbf_dev, int64_t xbf_plane, int ldb, float* mix_dev,
                      float mix_w, float* att_dev, void* stream) {
  if (!f || !f->finalized || !x_dev || B < 1 || C < 1 || C > 8 || T < 1) return fail(DZ_ERR_INVALID, "bad argument (1 <= channels <= 8)");
  cudaStream_t st = (cudaStream_t)stream;
  const int D = f->D, hu = f->hu, P = f->planes, Dp = rup(D, 8), hup = rup(hu, 8);
  const long long R = (long long)B * C * T;
  if (R != f->rows) {
    cudaDeviceSynchronize();
    if (f->g1) { gemm_plan_destroy(f->g1); f->g1 = nullptr; }
    if (f->g2) { gemm_plan_destroy(f->g2); f->g2 = nullptr; }
    cudaError_t e = f->xpl.alloc(((size_t)R * Dp + 64) * 2 * P, true);
    if (e == cudaSuccess) e = f->qkv.alloc(((size_t)R * 3 * hup + 64) * 2 * P, true);
    if (e == cudaSuccess) e = f->ctx.alloc(((size_t)R * hup + 64) * 2 * P, true);
    if (e == cudaSuccess) e = f->o.alloc(((size_t)R * Dp + 64) * 4, true);
    if (e != cudaSuccess) return fail(DZ_ERR_CUDA, std::string("fusion workspace: ") + cudaGetErrorString(e));
    if (hu % 8 != 0) return fail(DZ_ERR_INVALID, "fusion hidden size must be a multiple of 8");
    GemmDesc d = gemm_desc_default();
    d.M = (int)R; d.N = 3 * hu; d.K = f->wqkv.K; d.npass = f->npass; d.out_planes = P; d.fp16 = f->fp16;
    d.a = f->xpl.p; d.a_plane = (long long)R * Dp + 64; d.a_rstride = Dp; d.a_kinner = f->wqkv.K; d.a_rows_alloc = R;
    d.b = f->wqkv.w.p; d.b_plane = f->wqkv.plane; d.ldb = f->wqkv.ldb; d.b_gstride = f->wqkv.gstride; d.bias = f->wqkv.bias.as<float>();
    d.out_bf = f->qkv.p; d.ob_plane = (long long)R * 3 * hu + 64; d.ldob = 3 * hu;
    f->g1 = gemm_plan_create(d, 0);
    GemmDesc o = gemm_desc_default();
    o.M = (int)R; o.N = D; o.K = f->wo.K; o.npass = f->npass; o.out_planes = P; o.fp16 = f->fp16;
    o.a = f->ctx.p; o.a_plane = (long long)R * hu + 64; o.a_rstride = hu; o.a_kinner = f->wo.K; o.a_rows_alloc = R;
    o.b = f->wo.w.p; o.b_plane = f->wo.plane; o.ldb = f->wo.ldb; o.b_gstride = f->wo.gstride; o.bias = f->wo.bias.as<float>();
    o.out_f32 = f->o.as<float>(); o.ldo = Dp;
    f->g2 = gemm_plan_create(o, 0);
    if (!f->g1 || !f->g2) return fail(DZ_ERR_CUDA, std::string("fusion GEMM plan: ") + gemm_last_error());
    f->rows = R;
    cudaDeviceSynchronize();
  }
  cudaError_t e = launch_regroup(x_dev, R, D, ldx, 1, 1, 0, D, D, f->xpl.as<bf16>(), (long long)R * Dp + 64, Dp, P, f->fp16, st);
  if (e == cudaSuccess) e = gemm_plan_launch(f->g1, st);
  if (e == cudaSuccess) {
    const long long bt = (long long)B * T;
    cross_channel_attention_kernel<<<(unsigned)((bt + 7) / 8), 256, 0, st>>>(f->qkv.as<bf16>(), (long long)R * 3 * hu + 64, P, f->fp16, B, C, T, hu, f->heads,
                                                                            f->ctx.as<bf16>(), (long long)R * hu + 64, att_dev);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = gemm_plan_launch(f->g2, st);
  if (e == cudaSuccess) {
    ln_residual_kernel<<<(unsigned)((R + 7) / 8), 256, 0, st>>>(f->o.as<float>(), Dp, x_dev, ldx, R, D, f->gamma.as<float>(), f->beta.as<float>(), (bf16*)xbf_dev,
                                                                xbf_plane, ldb, P, f->fp16, mix_dev, mix_w);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) return fail(DZ_ERR_CUDA, std::string("fusion launch: ") + cudaGetErrorString(e) + " " + gemm_last_error());
  return DZ_OK;
}

/* fp32 rows [rows][ldx] (first C columns) -> 16-bit operand planes [rows][ldo] (hi [+ lo], `plane_elems` apart) */
int dz_rows_to_planes(const float* x_dev, int64_t rows, int C, int ldx, void* out_dev, int64_t plane_elems, int ldo, int planes, int fp16, void* stream) {
  if (!x_dev || !out_dev || rows < 1 || C < 1) return fail(DZ_ERR_INVALID, "bad argument");
  cudaError_t e = launch_regroup(x_dev, rows, C, ldx, 1, 1, 0, C, C, (bf16*)out_dev, plane_elems, ldo, planes, fp16, (cudaStream_t)stream);
  return e == cudaSuccess ? DZ_OK : fail(DZ_ERR_CUDA, cudaGetErrorString(e));
}

int dz_channel_mean(const float* in_dev, float* out_dev, int B, int C, int T, int D, int ld, void* stream) {
  if (!in_dev || !out_dev || B < 1 || C < 1 || T < 1 || D < 1) return fail(DZ_ERR_INVALID, "bad argument");
  channel_mean_kernel<<<592, 256, 0, (cudaStream_t)stream>>>(in_dev, out_dev, B, C, T, D, ld);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DZ_OK : fail(DZ_ERR_CUDA, cudaGetErrorString(e));
}

}  // extern "C"
