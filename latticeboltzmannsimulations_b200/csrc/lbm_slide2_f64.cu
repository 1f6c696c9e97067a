// Sliding-window two-step kernel, fp64 instantiations (see lbm_slide2.cuh).
#include "lbm_slide2_inst.cuh"

namespace lbm {
cudaError_t launch_slide2_f64(const StepArgs& a, const Slide2Launch& L) { return slide_launch<double, 3>(a, L); }
}  // namespace lbm
