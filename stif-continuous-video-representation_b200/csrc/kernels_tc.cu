// kernels_tc.cu -- (temporary stub; replaced by the fused tcgen05 kernels)
#include "stif_internal.h"
namespace stif {
struct TcWeights { int dummy; };
TcWeights* tc_weights_create(const FoldedWeights&, std::string&) { return new TcWeights{0}; }
void tc_weights_destroy(TcWeights* t) { delete t; }
cudaError_t decode_slab_tc(const LaunchCtx&, const TcWeights*, const Geometry&, const Workspace&, float, int, int, int, int, float*, int) {
  return cudaErrorNotSupported;
}
}  // namespace stif
