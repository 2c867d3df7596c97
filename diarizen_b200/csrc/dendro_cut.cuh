// Flat clusters from a centroid-linkage dendrogram, chosen the way the reference's AgglomerativeClustering.cluster does
// (pyannote-audio/pyannote/audio/pipelines/clustering.py:418-477) but as ONE pass over the merge list instead of one
// scipy.fcluster call per candidate cut:
//
//   * a cut is a predicate in(i) on merges that is closed downwards (in(parent) => in(children)):
//       threshold cut : in(i) = height(i) <= t and in(left) and in(right)      [= scipy's "max height in the subtree <= t",
//                                                                                which also handles centroid inversions]
//       iteration cut : in(i) = i <= it                                        [the state after the first it+1 merges]
//     flat clusters = maximal in() subtrees + the leaves hanging directly above the cut;
//   * the number of clusters with at least `mcs` members after EVERY iteration cut is a prefix sum over merges of
//     [size_i >= mcs] - [size_left >= mcs] - [size_right >= mcs] - no tree walk, no per-candidate relabelling;
//   * the candidate the reference's loop would stop at (closest |height - t| first, first exact hit, else the first
//     closest count, ties in |height - t| towards the lower merge index) is three block-wide arg-min reductions;
//   * labels are numbered exactly as scipy.cluster.hierarchy.fcluster numbers them (the order in which its depth-first
//     walk meets the clusters: at every node above the cut first the internal left child, then the internal right child,
//     then the leaf children), because those numbers leak into the RTTM speaker labels.
//
// The body is written against a small execution context (thread id, thread count, barrier, block arg-min, block scan) so
// that the identical source also compiles for the host with one "thread" - tests/test_dendro_cut_host.py checks it there
// against scipy and the reference-produced goldens; the CUDA instantiation is dendro_cut.cu.
#pragma once
#include <stdint.h>

#ifndef DZ_HD
#ifdef __CUDACC__
#define DZ_HD __host__ __device__ __forceinline__
#else
#define DZ_HD inline
#endif
#endif

namespace dz {

struct CutParams {
  int n;                // observations (merges = n - 1)
  double threshold;     // distance threshold of the first cut
  int mcs;              // effective min_cluster_size (already min(mcs, max(1, round(0.1 n))))
  int min_clusters, max_clusters;
  int num_clusters;     // 0 = not imposed
  int force_iteration;  // >= 0: skip the selection and cut after this merge (used by tests); -1 otherwise
};

struct CutKey {   // lexicographic (a, b, i), smaller wins
  double a, b;
  int i;
};
DZ_HD bool cut_less(const CutKey& x, const CutKey& y) {
  if (x.a != y.a) return x.a < y.a;
  if (x.b != y.b) return x.b < y.b;
  return x.i < y.i;
}

// info[] layout
enum { CUT_NUM_LARGE = 0, CUT_ITERATION = 1, CUT_FOUND_ONLY = 2, CUT_NUM_FLAT = 3, CUT_NUM_LARGE_AT_THRESHOLD = 4, CUT_TARGET = 5, CUT_INFO_LEN = 8 };

// Z [n-1][4] float64 (scipy layout).  left/right/parent: the caller provides them in fast memory when it can.
//   left, right : [n-1] child ids (0..2n-2)            parent : [2n-1] merge index of the parent, -1 for the root
//   in          : [n-1] bytes                          nlarge : [n-1] ints (scan)      stack : [2(n-1)] ints
//   node_label  : [n-1] ints                           labels : [n] ints (output, 0-based)
template <class Ctx>
DZ_HD void dendro_cut_body(Ctx& cx, const double* Z, const CutParams p, int* left, int* right, int* parent, unsigned char* in,
                           int* nlarge, int* stack, int* node_label, int* labels, int* info) {
  const int n = p.n, m = n - 1;
  const int tid = cx.tid(), nt = cx.nt();
  for (int i = tid; i < m; i += nt) {
    left[i] = (int)Z[4 * i + 0];
    right[i] = (int)Z[4 * i + 1];
  }
  for (int i = tid; i < 2 * n - 1; i += nt) parent[i] = -1;
  cx.sync();
  for (int i = tid; i < m; i += nt) { parent[left[i]] = i; parent[right[i]] = i; }
  // ---- prefix count of large clusters after every merge ----
  for (int i = tid; i < m; i += nt) {
    const int l = left[i], r = right[i];
    const int sl = l < n ? 1 : (int)Z[4 * (l - n) + 3], sr = r < n ? 1 : (int)Z[4 * (r - n) + 3];
    nlarge[i] = ((int)Z[4 * i + 3] >= p.mcs) - (sl >= p.mcs) - (sr >= p.mcs);
  }
  cx.sync();
  cx.inclusive_scan(nlarge, m, p.mcs <= 1 ? n : 0);   // with mcs <= 1 every leaf is already a large cluster
  // ---- the threshold cut ----
  for (int i = tid; i < m; i += nt) in[i] = Z[4 * i + 2] <= p.threshold ? 1 : 0;
  cx.sync();
  if (tid == 0)
    for (int i = 0; i < m; ++i)
      if (in[i]) {
        const int l = left[i], r = right[i];
        if ((l >= n && !in[l - n]) || (r >= n && !in[r - n])) in[i] = 0;
      }
  cx.sync();
  int cnt = 0;
  for (int i = tid; i < m; i += nt) {
    if (in[i] && (parent[n + i] < 0 || !in[parent[n + i]]) && (int)Z[4 * i + 3] >= p.mcs) ++cnt;
  }
  if (p.mcs <= 1)
    for (int l = tid; l < n; l += nt)
      if (!in[parent[l]]) ++cnt;
  const int large_t = cx.sum(cnt);
  // ---- which cut ----
  int target = p.num_clusters;
  if (large_t < p.min_clusters) target = p.min_clusters;
  else if (large_t > p.max_clusters) target = p.max_clusters;
  int it = -1, found_only = 0, num_large = large_t;
  if (p.force_iteration >= 0) {
    it = p.force_iteration < m ? p.force_iteration : m - 1;
  } else if (target > 0 && large_t != target) {
    CutKey exact{1e300, 1e300, 0x7fffffff}, close{1e300, 1e300, 0x7fffffff};
    for (int i = tid; i < m; i += nt) {
      if ((int)Z[4 * i + 3] < p.mcs) continue;
      double a = Z[4 * i + 2] - p.threshold;
      a = a < 0 ? -a : a;
      int d = nlarge[i] - target;
      d = d < 0 ? -d : d;
      if (d == 0) { const CutKey k{a, 0.0, i}; if (cut_less(k, exact)) exact = k; }
      const CutKey k2{(double)d, a, i};
      if (cut_less(k2, close)) close = k2;
    }
    exact = cx.argmin(exact);
    close = cx.argmin(close);
    if (exact.i != 0x7fffffff) {
      it = exact.i;
    } else {
      // the reference starts from "everything merged" (one large cluster) and only moves to a strictly closer count
      const int d0 = target > 1 ? target - 1 : 1 - target;
      it = (close.i != 0x7fffffff && close.a < (double)d0) ? close.i : m - 1;
      found_only = 1;
    }
  }
  if (it >= 0) {
    cx.sync();
    for (int i = tid; i < m; i += nt) in[i] = i <= it ? 1 : 0;
    num_large = nlarge[it];
  }
  cx.sync();
  // ---- numbering (scipy's walk restricted to the nodes above the cut) ----
  for (int i = tid; i < m; i += nt) node_label[i] = -1;
  for (int l = tid; l < n; l += nt) labels[l] = -1;
  cx.sync();
  if (tid == 0) {
    int ncl = 0;
    if (in[m - 1]) {
      node_label[m - 1] = ncl++;
    } else {
      int sp = 0;
      stack[0] = m - 1; stack[1] = 0; sp = 1;
      while (sp > 0) {
        const int node = stack[2 * (sp - 1)], st = stack[2 * (sp - 1) + 1];
        if (st < 2) {
          stack[2 * (sp - 1) + 1] = st + 1;
          const int c = st == 0 ? left[node] : right[node];
          if (c >= n) {
            if (in[c - n]) node_label[c - n] = ncl++;
            else { stack[2 * sp] = c - n; stack[2 * sp + 1] = 0; ++sp; }
          }
        } else {
          const int l = left[node], r = right[node];
          if (l < n) labels[l] = ncl++;
          if (r < n) labels[r] = ncl++;
          --sp;
        }
      }
    }
    info[CUT_NUM_LARGE] = num_large;
    info[CUT_ITERATION] = it;
    info[CUT_FOUND_ONLY] = found_only;
    info[CUT_NUM_FLAT] = ncl;
    info[CUT_NUM_LARGE_AT_THRESHOLD] = large_t;
    info[CUT_TARGET] = target;
  }
  cx.sync();
  for (int l = tid; l < n; l += nt) {
    if (labels[l] >= 0) continue;
    int q = parent[l];                       // in[q] holds: a leaf whose parent is above the cut was numbered by the walk
    while (parent[n + q] >= 0 && in[parent[n + q]]) q = parent[n + q];
    labels[l] = node_label[q];
  }
}

}  // namespace dz
