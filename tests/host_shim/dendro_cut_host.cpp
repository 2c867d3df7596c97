// TEST HARNESS ONLY (never part of libdiarizen_b200.so): compiles the body of the dz_dendrogram_cut kernel
// (diarizen_b200/csrc/dendro_cut.cuh) for the host with ONE "thread", so that the selection / numbering logic can be
// checked against scipy and the reference-produced goldens in the CPU test suite.  The CUDA instantiation (1024 threads,
// shared-memory reductions) is checked on the GPU by tests/test_glue_golden_gpu.py.
#include <vector>

#include "../../diarizen_b200/csrc/dendro_cut.cuh"

namespace {
struct HostCtx {
  int tid() const { return 0; }
  int nt() const { return 1; }
  void sync() const {}
  int sum(int v) const { return v; }
  dz::CutKey argmin(dz::CutKey k) const { return k; }
  void inclusive_scan(int* a, int m, int base) const {
    int run = base;
    for (int i = 0; i < m; ++i) { run += a[i]; a[i] = run; }
  }
};
}  // namespace

extern "C" int dendro_cut_host(const double* Z, int n, double threshold, int mcs, int min_clusters, int max_clusters, int num_clusters,
                               int force_iteration, int* labels, int* info) {
  if (n < 2) return -1;
  const int m = n - 1;
  std::vector<int> left(m), right(m), parent(2 * n - 1), nlarge(m), stack(2 * m), node_label(m);
  std::vector<unsigned char> in(m);
  HostCtx cx;
  dz::CutParams p{n, threshold, mcs, min_clusters, max_clusters, num_clusters > 0 ? num_clusters : 0, force_iteration};
  dz::dendro_cut_body(cx, Z, p, left.data(), right.data(), parent.data(), in.data(), nlarge.data(), stack.data(), node_label.data(), labels, info);
  return 0;
}
