// Instantiations + launcher of the marching two-step kernel for ONE element type (included by lbm_march2_f64.cu and
// lbm_march2_f32.cu with LBM_MARCH_T / LBM_MARCH_SUFFIX defined; two translation units so that they compile in parallel).
#include "lbm_internal.h"
#include "lbm_march2.cuh"

namespace lbm {
namespace {

template <typename T, int COLL, bool TURB, bool MACROS, int V, int NW, int MINB, int D>
cudaError_t launch_cfg(const StepArgs& a0, const March2Launch& L) {
    using Cfg = MarchCfg<T, V, TURB, D>;
    auto kern = lbm_step_march2<T, COLL, TURB, MACROS, V, NW, MINB, D>;
    static bool attr_done[64] = {};                // cudaFuncSetAttribute is per device
    int dev = 0;
    if (cudaError_t e = cudaGetDevice(&dev)) return e;
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        if (cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM(NW))) return e;
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    StepArgs a = a0;
    a.nsx = (a.nx + 32 * V - 1) / (32 * V);
    const int nseg = (a.row_count + a.seg_h - 1) / a.seg_h;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((a.nsx + NW - 1) / NW, nseg, L.batch);
    cfg.blockDim = dim3(NW * 32, 1, 1);
    cfg.dynamicSmemBytes = Cfg::SMEM(NW);
    cfg.stream = L.st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = L.pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, a);
}

template <typename T, int COLL, int V, int NW, int MINB, int D>
cudaError_t launch_flags(const StepArgs& a, const March2Launch& L) {
    if (L.turb) {
        if (L.macros) return launch_cfg<T, COLL, true, true, V, NW, MINB, D>(a, L);
        return launch_cfg<T, COLL, true, false, V, NW, MINB, D>(a, L);
    }
    if (L.macros) return launch_cfg<T, COLL, false, true, V, NW, MINB, D>(a, L);
    return launch_cfg<T, COLL, false, false, V, NW, MINB, D>(a, L);
}

// the shipped variant: every collision / closure / output combination
template <typename T, int V, int NW, int MINB, int D>
cudaError_t launch_default(const StepArgs& a, const March2Launch& L) {
    switch (L.coll) {
        case COLL_SRT: return launch_flags<T, COLL_SRT, V, NW, MINB, D>(a, L);
        case COLL_TRT: return launch_flags<T, COLL_TRT, V, NW, MINB, D>(a, L);
        default: return launch_flags<T, COLL_MRT, V, NW, MINB, D>(a, L);
    }
}

// tuning variants: MRT without closure and without macro output only (everything else falls back to variant 0)
template <typename T, int V, int NW, int MINB, int D>
cudaError_t launch_tuning(const StepArgs& a, const March2Launch& L) {
    return launch_cfg<T, COLL_MRT, false, false, V, NW, MINB, D>(a, L);
}

}  // namespace
}  // namespace lbm
