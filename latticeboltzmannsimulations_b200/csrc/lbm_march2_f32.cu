// Marching two-step kernel, fp32 instantiations (see lbm_march2.cuh).
#include "lbm_march2_inst.cuh"

namespace lbm {

static const int kColsF32[] = {128, 128, 128, 64, 64, 32, 128, 64};

cudaError_t launch_march2_f32(const StepArgs& a, const March2Launch& L) {
    const bool plain = L.coll == COLL_MRT && !L.turb && !L.macros;
    switch (plain ? L.variant : 0) {
        case 1: return launch_tuning<float, 4, 4, 3, 2>(a, L);
        case 2: return launch_tuning<float, 4, 4, 2, 4>(a, L);
        case 3: return launch_tuning<float, 2, 4, 4, 4>(a, L);
        case 4: return launch_tuning<float, 2, 4, 5, 3>(a, L);
        case 5: return launch_tuning<float, 1, 4, 6, 4>(a, L);
        case 6: return launch_tuning<float, 4, 2, 6, 3>(a, L);
        case 7: return launch_tuning<float, 2, 4, 3, 6>(a, L);
        default: return launch_default<float, 4, 4, 3, 3>(a, L);
    }
}

int march2_variants_f32() { return (int)(sizeof(kColsF32) / sizeof(int)); }
int march2_cols_f32(int variant) { return variant >= 0 && variant < march2_variants_f32() ? kColsF32[variant] : 0; }

}  // namespace lbm
