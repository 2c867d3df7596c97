// dz_dendrogram_cut: the flat-cluster selection of the reference's AgglomerativeClustering.cluster
// (pyannote-audio/pyannote/audio/pipelines/clustering.py:418-492) and scipy's fcluster numbering, as one single-CTA
// kernel over the merge list the linkage kernel left on the device (algorithm: dendro_cut.cuh).
// Latency-bound by construction (a few dependent passes over N-1 merges); the child / parent tables live in shared
// memory when N <= ~13 k so that the two sequential passes (cut closure, numbering walk) run at shared-memory latency.
#include <string>

#include "../../include/diarizen_b200.h"
#include "dendro_cut.cuh"

namespace dz {
std::string& tls_error();
int fail(int code, const std::string& msg);

struct DevCtx {
  double* red_d;   // [2 * 32]
  int* red_i;      // [32 + 1]
  __device__ int tid() const { return threadIdx.x; }
  __device__ int nt() const { return blockDim.x; }
  __device__ void sync() const { __syncthreads(); }
  __device__ int sum(int v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red_i[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red_i[w];
    __syncthreads();
    return t;
  }
  __device__ CutKey argmin(CutKey k) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      CutKey t;
      t.a = __shfl_xor_sync(0xffffffffu, k.a, o);
      t.b = __shfl_xor_sync(0xffffffffu, k.b, o);
      t.i = __shfl_xor_sync(0xffffffffu, k.i, o);
      if (cut_less(t, k)) k = t;
    }
    __syncthreads();
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { red_d[2 * w] = k.a; red_d[2 * w + 1] = k.b; red_i[w] = k.i; }
    __syncthreads();
    CutKey best{red_d[0], red_d[1], red_i[0]};
    for (int q = 1; q < (int)(blockDim.x >> 5); ++q) {
      const CutKey t{red_d[2 * q], red_d[2 * q + 1], red_i[q]};
      if (cut_less(t, best)) best = t;
    }
    __syncthreads();
    return best;
  }
  // in-place inclusive prefix sum of a[0..m) plus `base`; contiguous chunk per thread, chunk sums scanned by warp shuffles
  __device__ void inclusive_scan(int* a, int m, int base) const {
    const int per = (m + blockDim.x - 1) / blockDim.x;
    const int lo = min(m, (int)threadIdx.x * per), hi = min(m, lo + per);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += a[i];
    int incl = s;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    __syncthreads();
    if (lane == 31) red_i[w] = incl;
    __syncthreads();
    int woff = 0;
    for (int q = 0; q < w; ++q) woff += red_i[q];
    int run = base + woff + incl - s;
    for (int i = lo; i < hi; ++i) { run += a[i]; a[i] = run; }
    __syncthreads();
  }
};

__global__ void __launch_bounds__(1024) dendro_cut_kernel(const double* __restrict__ Z, CutParams p, int* ws, int use_smem,
                                                          int* __restrict__ labels, int* __restrict__ info) {
  extern __shared__ int dsm[];
  __shared__ double red_d[64];
  __shared__ int red_i[33];
  const int n = p.n, m = n - 1;
  // workspace layout (ints): nlarge[m] stack[2m] node_label[m] | left[m] right[m] parent[2n-1] in[(m+3)/4]
  int* nlarge = ws;
  int* stack = nlarge + m;
  int* node_label = stack + 2 * m;
  int* fast = use_smem ? dsm : node_label + m;
  int* left = fast;
  int* right = left + m;
  int* parent = right + m;
  unsigned char* in = reinterpret_cast<unsigned char*>(parent + (2 * n - 1));
  DevCtx cx{red_d, red_i};
  dendro_cut_body(cx, Z, p, left, right, parent, in, nlarge, stack, node_label, labels, info);
}

}  // namespace dz

using namespace dz;

extern "C" {

int64_t dz_dendrogram_cut_workspace_bytes(int N) { return (int64_t)4 * (4 * (int64_t)N + 2 * (int64_t)N + 2 * (int64_t)N + N / 4 + 16); }

int dz_dendrogram_cut(const double* z_dev, int N, double threshold, int min_cluster_size, int min_clusters, int max_clusters,
                      int num_clusters, int force_iteration, int32_t* labels_dev, int32_t* info_dev, void* workspace_dev, void* stream) {
  if (!z_dev || !labels_dev || !info_dev || !workspace_dev || N < 2 || min_cluster_size < 1 || min_clusters < 1 || max_clusters < min_clusters)
    return fail(DZ_ERR_INVALID, "bad argument");
  CutParams p{N, threshold, min_cluster_size, min_clusters, max_clusters, num_clusters > 0 ? num_clusters : 0, force_iteration};
  const size_t smem = sizeof(int) * ((size_t)2 * (N - 1) + (2 * N - 1)) + (size_t)N + 16;
  const int use_smem = smem <= 220 * 1024;
  if (use_smem) {
    static size_t attr = 0;
    if (smem > attr) {
      cudaError_t e = cudaFuncSetAttribute(dendro_cut_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return fail(DZ_ERR_CUDA, cudaGetErrorString(e));
      attr = smem;
    }
  }
  dendro_cut_kernel<<<1, 1024, use_smem ? smem : 0, (cudaStream_t)stream>>>(z_dev, p, (int*)workspace_dev, use_smem, labels_dev, info_dev);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(DZ_ERR_CUDA, cudaGetErrorString(e));
  return DZ_OK;
}

}  // extern "C"
