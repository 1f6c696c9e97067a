// Internal interface between the translation units of liblbm_b200.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "lbm_device.cuh"

namespace lbm {

// Sliding-window two-step kernel (lbm_slide2.cuh), compiled per dtype in lbm_slide2_f64.cu / lbm_slide2_f32.cu.
struct Slide2Launch {
    int coll;            // COLL_*
    bool turb, macros;
    int batch;
    bool pdl;
    cudaStream_t st;
    const CUtensorMap* tmap;   // tensor map of the source buffer, box = (staged row width) x (rows per iteration); NULL: none
};
cudaError_t launch_slide2_f64(const StepArgs& a, const Slide2Launch& L);
cudaError_t launch_slide2_f32(const StepArgs& a, const Slide2Launch& L);

}  // namespace lbm
