// Instantiations + launcher of the sliding-window two-step kernel for ONE element type (included by
// lbm_slide2_f64.cu and lbm_slide2_f32.cu; two translation units so that they compile in parallel).
#include "lbm_internal.h"
#include "lbm_slide2.cuh"

namespace lbm {
namespace {

template <typename T, int COLL, bool MACROS, int MINB, bool TURB>
cudaError_t slide_launch_cfg(const StepArgs& a, const Slide2Launch& L) {
    using Cfg = SlideCfg<T, TURB>;
    auto kern = lbm_step_slide2<T, COLL, MACROS, MINB, TURB>;
    static bool attr_done[64] = {};                // cudaFuncSetAttribute is per device
    int dev = 0;
    if (cudaError_t e = cudaGetDevice(&dev)) return e;
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        if (cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM)) return e;
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    const int nseg = a.seg_stride ? 2 : (a.row_count + a.seg_h - 1) / a.seg_h;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((a.nx + Cfg::TX - 1) / Cfg::TX, nseg, L.batch);
    cfg.blockDim = dim3(Cfg::NT, 1, 1);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = L.st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = L.pdl ? 1 : 0;
    static const CUtensorMap no_map = {};
    return cudaLaunchKernelEx(&cfg, kern, a, L.tmap ? *L.tmap : no_map);
}

template <typename T, int COLL, int MINB>
cudaError_t slide_launch_flags(const StepArgs& a, const Slide2Launch& L) {
    if (L.turb) {
        if (L.macros) return slide_launch_cfg<T, COLL, true, MINB, true>(a, L);
        return slide_launch_cfg<T, COLL, false, MINB, true>(a, L);
    }
    if (L.macros) return slide_launch_cfg<T, COLL, true, MINB, false>(a, L);
    return slide_launch_cfg<T, COLL, false, MINB, false>(a, L);
}

template <typename T, int MINB>
cudaError_t slide_launch(const StepArgs& a, const Slide2Launch& L) {
    switch (L.coll) {
        case COLL_SRT: return slide_launch_flags<T, COLL_SRT, MINB>(a, L);
        case COLL_TRT: return slide_launch_flags<T, COLL_TRT, MINB>(a, L);
        default: return slide_launch_flags<T, COLL_MRT, MINB>(a, L);
    }
}

}  // namespace
}  // namespace lbm
