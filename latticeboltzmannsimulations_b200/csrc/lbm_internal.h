// Internal interface between the translation units of liblbm_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include "lbm_device.cuh"

namespace lbm {

// Marching two-step kernel (lbm_march2.cuh), compiled per dtype in lbm_march2_f64.cu / lbm_march2_f32.cu.
// `variant` selects (nodes per lane V, register budget); variant 0 is the shipped default of the dtype.
struct March2Launch {
    int coll;            // COLL_*
    bool turb, macros;
    int variant;
    int batch;
    bool pdl;            // programmatic dependent launch attribute
    cudaStream_t st;
};
int march2_variants(int esz);                      // number of compiled variants for this element size
int march2_cols(int esz, int variant);             // columns per warp (32 * V), 0 if the variant does not exist
cudaError_t launch_march2_f64(const StepArgs& a, const March2Launch& L);
cudaError_t launch_march2_f32(const StepArgs& a, const March2Launch& L);

// Sliding-window two-step kernel (lbm_slide2.cuh), compiled per dtype in lbm_slide2_f64.cu / lbm_slide2_f32.cu.
struct Slide2Launch {
    int coll;            // COLL_*
    bool turb, macros;
    int batch;
    bool pdl;
    cudaStream_t st;
};
cudaError_t launch_slide2_f64(const StepArgs& a, const Slide2Launch& L);
cudaError_t launch_slide2_f32(const StepArgs& a, const Slide2Launch& L);

}  // namespace lbm
