// Tensor-core (tcgen05) E-step of the spherical k-means -- placeholder build.
// The real kernel lands in the next commit; until then the shape test says
// "unsupported" so every call takes the fp32 CUDA-core E-step.
#include "kmeans.cuh"

namespace hsg {

bool tc_shape_supported(int, int, int) { return false; }
size_t tc_workspace_bytes(int, int, int) { return 0; }
void tc_carve(Carver&, TcState& t, int, int, int) { t.enabled = false; }
int tc_prepare(TcState&, int64_t, int) { return HSG_OK; }
int tc_convert_centroids(const EStepArgs&, const TcState&, cudaStream_t) { return HSG_OK; }
int estep_tc(const EStepArgs&, const TcState&, cudaStream_t) {
  set_error("tensor-core E-step not built");
  return HSG_E_UNSUPPORTED;
}

}  // namespace hsg
