// TEST HARNESS ONLY: the per-chunk assignment routine of dz_assign (diarizen_b200/csrc/lsap_small.cuh) compiled for the host, so
// that the CPU suite can check it against scipy.optimize.linear_sum_assignment on tie-heavy inputs.
#include "../../diarizen_b200/csrc/lsap_small.cuh"

extern "C" void lsap_host(const double* soft, int C, int S, int K, signed char* hard) {
  for (int c = 0; c < C; ++c) dz::lsap_assign_max(soft + (long long)c * S * K, S, K, reinterpret_cast<int8_t*>(hard) + c * S);
}
