// Marching two-step kernel, fp64 instantiations (see lbm_march2.cuh).
#include "lbm_march2_inst.cuh"

namespace lbm {

// variant -> columns per warp (32 * V); template arguments below: V nodes per lane, warps per CTA, min CTAs per SM
// (= register budget), D stages of the cp.async ring
static const int kColsF64[] = {32, 32, 32, 32, 64, 64, 64, 64};

cudaError_t launch_march2_f64(const StepArgs& a, const March2Launch& L) {
    const bool plain = L.coll == COLL_MRT && !L.turb && !L.macros;
    switch (plain ? L.variant : 0) {
        case 1: return launch_tuning<double, 1, 4, 5, 3>(a, L);
        case 2: return launch_tuning<double, 1, 4, 4, 5>(a, L);
        case 3: return launch_tuning<double, 1, 4, 3, 6>(a, L);
        case 4: return launch_tuning<double, 2, 4, 3, 3>(a, L);
        case 5: return launch_tuning<double, 2, 4, 3, 2>(a, L);
        case 6: return launch_tuning<double, 2, 4, 2, 4>(a, L);
        case 7: return launch_tuning<double, 2, 2, 6, 3>(a, L);
        default: return launch_default<double, 1, 4, 5, 4>(a, L);
    }
}

int march2_variants_f64() { return (int)(sizeof(kColsF64) / sizeof(int)); }
int march2_cols_f64(int variant) { return variant >= 0 && variant < march2_variants_f64() ? kColsF64[variant] : 0; }

}  // namespace lbm
