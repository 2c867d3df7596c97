// Per-chunk constrained assignment (pyannote-audio/pyannote/audio/pipelines/clustering.py:159-173):
// scipy.optimize.linear_sum_assignment(cost, maximize=True) on a (speakers <= 4) x (clusters <= 127) score matrix.
// scipy's solver (third party; not in /root/reference) is the shortest-augmenting-path algorithm of Crouse, "On implementing 2D
// rectangular assignment algorithms" (2016), as shipped in scipy/optimize/rectangular_lsap: rows are added one at a time, the
// column scan runs over a `remaining` list initialised in REVERSE order, ties in the shortest path cost prefer a column
// that is still unassigned, a tall matrix is transposed first and a maximisation negates the costs.  The restatement
// below follows those steps one for one, because identical rows - local speakers that are inactive in a chunk all carry
// the same embedding - make ties the common case, and which of them gets which cluster is decided by exactly these
// details.  Compiles for the host too (tests/host_shim/lsap_host.cpp checks it against scipy on tie-heavy inputs).
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef DZ_HD
#ifdef __CUDACC__
#define DZ_HD __host__ __device__ __forceinline__
#else
#define DZ_HD inline
#endif
#endif

namespace dz {

static constexpr int LSAP_MAX_COLS = 128;

// soft: [S][K] row-major scores (to be maximised); hard: [S] assigned cluster or -2.  S <= 4, K <= 127.
DZ_HD void lsap_assign_max(const double* soft, int S, int K, int8_t* hard) {
  const bool tr = K < S;                       // tall matrix: solve the transposed problem (rows' = clusters)
  const int nr = tr ? K : S, nc = tr ? S : K;
  double u[4], v[LSAP_MAX_COLS], spc[LSAP_MAX_COLS];
  int path[LSAP_MAX_COLS], row4col[LSAP_MAX_COLS], remaining[LSAP_MAX_COLS], col4row[4];
  bool SR[4], SC[LSAP_MAX_COLS];
  for (int i = 0; i < nr; ++i) { u[i] = 0.0; col4row[i] = -1; }
  for (int j = 0; j < nc; ++j) { v[j] = 0.0; path[j] = -1; row4col[j] = -1; }
  auto cost = [&](int i, int j) -> double { return tr ? -soft[j * K + i] : -soft[i * K + j]; };
  for (int cur = 0; cur < nr; ++cur) {
    // ---- shortest augmenting path from row `cur` ----
    double minVal = 0.0;
    int num_remaining = nc;
    for (int it = 0; it < nc; ++it) remaining[it] = nc - it - 1;
    for (int i = 0; i < nr; ++i) SR[i] = false;
    for (int j = 0; j < nc; ++j) { SC[j] = false; spc[j] = INFINITY; }
    int sink = -1, i = cur;
    while (sink == -1) {
      int index = -1;
      double lowest = INFINITY;
      SR[i] = true;
      for (int it = 0; it < num_remaining; ++it) {
        const int j = remaining[it];
        const double r = minVal + cost(i, j) - u[i] - v[j];
        if (r < spc[j]) { path[j] = i; spc[j] = r; }
        if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) { lowest = spc[j]; index = it; }
      }
      minVal = lowest;
      if (index < 0) break;                    // infeasible (cannot happen with finite scores)
      const int j = remaining[index];
      if (row4col[j] == -1) sink = j; else i = row4col[j];
      SC[j] = true;
      remaining[index] = remaining[--num_remaining];
    }
    if (sink < 0) break;
    // ---- dual update ----
    u[cur] += minVal;
    for (int r = 0; r < nr; ++r)
      if (SR[r] && r != cur) u[r] += minVal - spc[col4row[r]];
    for (int j = 0; j < nc; ++j)
      if (SC[j]) v[j] -= minVal - spc[j];
    // ---- augment ----
    int j = sink;
    while (true) {
      const int r = path[j];
      row4col[j] = r;
      const int t = col4row[r]; col4row[r] = j; j = t;
      if (r == cur) break;
    }
  }
  for (int s = 0; s < S; ++s) hard[s] = -2;
  if (tr) { for (int k = 0; k < nr; ++k) if (col4row[k] >= 0) hard[col4row[k]] = (int8_t)k; }
  else { for (int s = 0; s < nr; ++s) if (col4row[s] >= 0) hard[s] = (int8_t)col4row[s]; }
}

}  // namespace dz
