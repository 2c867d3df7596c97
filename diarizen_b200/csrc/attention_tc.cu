// tcgen05 attention (placeholder until the tensor-core kernel lands: fails loudly, never falls back).
#include "common.cuh"
#include "seg_kernels.h"
namespace dz {
cudaError_t launch_attention_tc(const AttnArgs& a, int B, cudaStream_t st) {
  (void)a; (void)B; (void)st;
  return cudaErrorNotSupported;
}
}  // namespace dz
