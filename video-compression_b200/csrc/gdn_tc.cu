// K-GDN on tcgen05 (3xTF32): placeholder until the tensor-core kernel lands; the dispatcher in gdn.cu falls
// through to the CUDA-core fp32 kernel when this returns B200VC_EUNSUPPORTED.
#include "common.cuh"

namespace b200vc {
int launch_gdn_tc(const float*, const float*, const float*, float*, int, int, int64_t, int, cudaStream_t) {
  set_error("gdn_f32: tcgen05 kernel not built");
  return B200VC_EUNSUPPORTED;
}
}  // namespace b200vc
