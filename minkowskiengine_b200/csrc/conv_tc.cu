// Placeholder until the tcgen05 kernels land: nothing is "supported", so every call
// takes the SIMT path.  (Replaced by the real implementation in the next milestone.)
#include "conv_tc.cuh"

namespace meb200 {

bool conv_tc_supported(int, uint32_t, uint32_t) { return false; }
bool conv_wgrad_tc_supported(int, uint32_t, uint32_t) { return false; }

int conv_forward_tc(const void *, int, uint32_t, uint32_t, const void *, uint32_t, uint32_t, bool,
                    const int32_t *, uint32_t, void *, int, void *, cudaStream_t) {
  set_error("tcgen05 convolution path not built");
  return MEB200_ERR_UNSUPPORTED;
}

int conv_wgrad_tc(const void *, const void *, int, uint32_t, uint32_t, uint32_t, const int32_t *,
                  uint32_t, float *, cudaStream_t) {
  set_error("tcgen05 wgrad path not built");
  return MEB200_ERR_UNSUPPORTED;
}

}  // namespace meb200
