// Backward of the camera->BEV lift (gradient w.r.t. the head tensor).  Placeholder until the kernel lands.
#include "lift_tile.cuh"

namespace fiery {

int launch_lift_backward(const LiftParams& P, const void* head, int head_dtype, cudaStream_t stream) {
    (void)P; (void)head; (void)head_dtype; (void)stream;
    return set_error(FIERY_E_UNSUPPORTED, "fiery_lift_backward is not implemented in this build");
}

}  // namespace fiery
