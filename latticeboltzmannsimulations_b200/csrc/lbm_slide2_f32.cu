// Sliding-window two-step kernel, fp32 instantiations (see lbm_slide2.cuh).
#include "lbm_slide2_inst.cuh"

namespace lbm {
cudaError_t launch_slide2_f32(const StepArgs& a, const Slide2Launch& L) { return slide_launch<float, 3>(a, L); }
}  // namespace lbm
