// C-ABI implementation (include/lbm_b200.h): solver object, kernel dispatch, host<->device plumbing.
// Kernels: lbm_kernels.cuh ("ldg" family, auxiliary), lbm_fused2.cuh (temporal blocking), lbm_tma.cuh (TMA engine).
//
// Device layout (private; lbm_get_layout): population buffers A and B, each
//     [cavity b][population k][stored row r = 0 .. ny_local+1][pitch]      x fastest,
// stored row r holds local row r-1; rows 0 and ny_local+1 are ghost rows (neighbour strip's edge rows, filled by the
// halo exchange; never read where they fall outside the physical cavity).  Between steps a buffer holds the
// POST-collision populations; the next launch pulls them (streaming), applies the wall rule, collides and stores
// into the other buffer.  Side buffers carry what the wall rule needs from the previous step: rho_lid[b][x] and the
// four doubly-orphaned corner populations carry[b][4] (see oracle/lbm_oracle.py PullState).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/lbm_b200.h"
#include "lbm_device.cuh"
#include "lbm_kernels.cuh"
#include "lbm_fused2.cuh"
#include "lbm_aa.cuh"
#include "lbm_internal.h"
#include "lbm_tma.cuh"

using namespace lbm;


// ------------------------------------------------------------------------------------------------------------
// error handling
// ------------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(e__ == cudaErrorMemoryAllocation ? LBM_ENOMEM : LBM_ECUDA,                       \
                        std::string(#call) + ": " + cudaGetErrorString(e__));                            \
    } while (0)

// Where the shared-memory tiles pay (round-1 sweeps, fp64; tools/size_sweep2.py): large cavities with 64x8 tiles
// (1024^2: 45 807 -> 63 586 MLUPS, 4096^2: 47 558 -> 74 393); launch-latency-bound small cavities with 32x8 tiles, so
// that one wave still covers all SMs (384^2: 3.46 -> 2.98 us/step, 128^2: 2.11 -> 1.97).  In between (640^2:
// 6.23 vs 6.65 us/step) neither tile shape fills the machine well and the one-step kernel stays.
#define LBM_FUSED2_MIN_NODES 10000
#define LBM_FUSED2_SMALL_NODES 250000     // below: 32x8 tiles
#define LBM_FUSED2_LARGE_NODES 600000     // from here: 64x8 (fp64) / 32x16 (fp32) tiles
// The sliding-window two-step kernel (lbm_slide2.cuh: one CTA per column strip, bulk-copy double-buffered source rows)
// takes over where its segments fill the machine and the state no longer fits L2 (tools/size_sweep2.py mid,
// tools/h_sweep_mid.py, best segment height each: 1024^2 fp64 tiles 61 563 / sliding 60 102 MLUPS, fp32 one-step
// 105 014 / sliding 98 324; 1280^2 fp64 64 311 / 67 018, fp32 85 838 / 107 881; 8 x 384^2 fp64 61 889 / 56 007;
// 2048^2 fp64 69 576 / 80 187, fp32 88 467 / 135 781).  It also covers the Smagorinsky closure (whole cavities) and
// batches with frozen cavities, which the tiles do not.
#define LBM_SLIDE_MIN_NODES 1500000

// ------------------------------------------------------------------------------------------------------------
// solver object
// ------------------------------------------------------------------------------------------------------------
struct lbm_solver {
    lbm_config_t cfg{};
    int device = 0;
    int esz = 8;
    int pitch = 0, nyl = 0;
    long long plane = 0, cavity = 0, mplane = 0;
    size_t state_bytes = 0;
    void* f[2] = {nullptr, nullptr};
    bool own_f = true;
    int cur = 0;               // buffer read by the next step
    bool pre = true;           // cur holds pre-collision `fin` (just uploaded / initialised)
    void* rho = nullptr;       // [batch][nyl][pitch]
    void* ux = nullptr;
    void* uy = nullptr;
    void* rho_lid = nullptr;
    void* carry = nullptr;
    void* pi_eq = nullptr;     // Smagorinsky state (turb = 1 only)
    void* rho_prev = nullptr;
    int* active = nullptr;     // device copy of the per-cavity active flags (NULL until lbm_set_active is used)
    std::vector<int> active_host;
    double* usum = nullptr;    // [batch] accumulator of lbm_mean_u
    double* conv_past = nullptr;   // [batch] mean(u) at the previous convergence check (device)
    int* conv_count = nullptr;     // [batch] hits so far; [batch .. 2 batch) cavities retired by the last check
    CavityParams* cav = nullptr;
    std::vector<CavityParams> cav_host;
    bool cav_dirty = true;
    void* staging = nullptr;   // host-layout staging for uploads/downloads (two plane-sized slots)
    cudaStream_t copy_stream = nullptr;            // host link copies, overlapped with the transposes
    cudaEvent_t ev_slot_full[2] = {nullptr, nullptr}, ev_slot_free[2] = {nullptr, nullptr};
    size_t staging_bytes = 0;
    void* scratch = nullptr;   // finalize target (populations) when B must stay intact
    int64_t steps = 0, launches = 0;
    int engine = LBM_ENGINE_LDG;
    // tma family
    CUtensorMap tmap[2][2];    // [buffer][0 = narrow box, 1 = wide box]
    bool tmap_ok = false;
    int num_sms = 148;
    int tma_variant = 0;       // index into the compiled (TY, STAGES) configurations
    int tma_ctas_per_sm = 1;
    // nodes per thread of the ldg family (1 = scalar kernel).  Measured at 4096^2 on B200 (round-1 sweep):
    // fp64 scalar 47 067 vs vec2 45 905 MLUPS; fp32 scalar 87 132, vec2 88 321, vec4 91 055 MLUPS.
    // fp32: 0 = by size (tools/vec_sweep.py, us/step V = 1 / 2 / 4: 384^2 3.08 / 3.72 / 4.05, 640^2 5.83 / 5.20 / 5.82,
    // 1024^2 12.7 / 10.0 / 10.1, 1400^2 27.5 / 25.8 / 25.4; with the closure 640^2 6.62 / 6.52 / 8.21, 1400^2 32.9 / 35.3 /
    // 35.2): small launches want threads, large ones fewer instructions per node
    int vec_f64 = 1, vec_f32 = 0;
    // CUDA graphs of the steady step loop: graph[p] = 2*GRAPH_PAIRS launches starting with buffer p as source
    cudaGraphExec_t graph[4] = {nullptr, nullptr, nullptr, nullptr};   // index = cur * 2 + side
    cudaStream_t capture_stream = nullptr;
    int use_graph = 1;
    int use_pdl = 1;
    int use_fused2 = 1;        // temporal blocking (two steps per launch) for whole cavities
    int side = 0;              // which half of the double-buffered rho_lid / carry arrays is current
    int fused2_tile = -1;      // tile-shape variant of the fused kernel (-1 = per-dtype default)
    long long fused2_min_nodes = LBM_FUSED2_MIN_NODES;
    int slide_tma = 1;         // sliding-window kernel: stage interior blocks by tensor copies (0: one bulk copy per row)
    CUtensorMap tmap_slide[2]; // [buffer], box = staged row width x 4 rows
    int tmap_slide_state = 0;  // 0 not made yet, 1 ready
    int use_slide = 1;         // sliding-window two-step kernel for large cavities / batches (no closure)
    int slide_h = 0;           // rows per segment, 0 = chosen from the number of work items
    long long slide_min_nodes = LBM_SLIDE_MIN_NODES;
    // AA pattern (LBM_ENGINE_AA, lbm_aa.cuh): one population buffer f[0]; f[1] stays NULL
    bool aa = false;
    bool aa_swapped = false;   // the buffer is in the SWAPPED layout (an odd number of AA steps since it was last NATURAL)
};

// A fresh state (init / upload) un-freezes every cavity; graphs captured with the old flag pointer are dropped.
static void reset_active(lbm_solver* s) {
    if (s->conv_past) { cudaFree(s->conv_past); s->conv_past = nullptr; cudaFree(s->conv_count); s->conv_count = nullptr; }
    if (!s->active) return;
    cudaFree(s->active);
    s->active = nullptr;
    for (int i = 0; i < 4; ++i)
        if (s->graph[i]) { cudaGraphExecDestroy(s->graph[i]); s->graph[i] = nullptr; }
}

static int set_device(lbm_solver* s) {
    CK(cudaSetDevice(s->device));
    return LBM_OK;
}

static StepArgs make_args(lbm_solver* s, const void* src, void* dst) {
    StepArgs a{};
    a.src = src; a.dst = dst;
    a.rho = s->rho; a.ux = s->ux; a.uy = s->uy;
    const size_t rl_half = (size_t)s->cfg.batch * s->pitch * s->esz, ca_half = (size_t)s->cfg.batch * 4 * s->esz;
    a.rho_lid = (char*)s->rho_lid + s->side * rl_half;
    a.carry = (char*)s->carry + s->side * ca_half;
    a.rho_lid_out = (char*)s->rho_lid + (s->side ^ 1) * rl_half;
    a.carry_out = (char*)s->carry + (s->side ^ 1) * ca_half;
    a.cav = s->cav;
    if (s->pi_eq) {   // Smagorinsky state, two halves like rho_lid: one-step kernels update the current half in place
        const size_t m_half = (size_t)s->cfg.batch * s->mplane * s->esz;
        a.pi_eq = (char*)s->pi_eq + s->side * m_half;
        a.rho_prev = (char*)s->rho_prev + s->side * m_half;
        a.pi_eq_out = (char*)s->pi_eq + (s->side ^ 1) * m_half;
        a.rho_prev_out = (char*)s->rho_prev + (s->side ^ 1) * m_half;
    }
    a.active = s->active;
    a.nx = s->cfg.nx; a.ny = s->cfg.ny; a.y0 = s->cfg.y0; a.nyl = s->nyl; a.pitch = s->pitch;
    a.plane = s->plane; a.cavity = s->cavity; a.mplane = s->mplane;
    a.row_begin = 0; a.row_stride = 1;
    return a;
}

static int sync_params(lbm_solver* s, cudaStream_t st) {
    if (!s->cav_dirty) return LBM_OK;
    CK(cudaMemcpyAsync(s->cav, s->cav_host.data(), sizeof(CavityParams) * s->cav_host.size(),
                       cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));   // cav_host may be modified again by the caller right away
    s->cav_dirty = false;
    return LBM_OK;
}


// cudaFuncSetAttribute is per device: remember for which devices a kernel has been given its dynamic-smem limit.
struct AttrOnce {
    bool done[64] = {};
    template <typename K>
    cudaError_t ensure(K kern, int bytes) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
        return e;
    }
};

// ---- tma family: tensor maps and launch ------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename T> struct TmaVec { static constexpr int V = 16 / sizeof(T); };

// compiled tile configurations: {TY, STAGES}
#define LBM_TMA_VARIANTS 3
static const int kTmaTY[LBM_TMA_VARIANTS] = {4, 4, 8};      // rows per tile (fp32 tiles use twice as many rows)
// stages per variant: {4, 3, 2}

static int make_tensor_maps(lbm_solver* s) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn)
        return fail(LBM_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    EncodeTiledFn encode = (EncodeTiledFn)fn;
    const int V = 16 / s->esz;
    const int ty = kTmaTY[s->tma_variant] * (s->esz == 4 ? 2 : 1);
    const int txn = (s->esz == 8 ? 64 : 32) * V;       // TmaCfg::TX
    cuuint64_t dims[2] = {(cuuint64_t)s->pitch, (cuuint64_t)s->cfg.batch * 9 * (cuuint64_t)(s->nyl + 2)};
    cuuint64_t strides[1] = {(cuuint64_t)s->pitch * s->esz};
    cuuint32_t estr[2] = {1, 1};
    for (int i = 0; i < 2; ++i) {
        for (int w = 0; w < 2; ++w) {
            cuuint32_t box[2] = {(cuuint32_t)(txn + (w ? 128 / s->esz : 0)), (cuuint32_t)ty};
            CUresult r = encode(&s->tmap[i][w],
                                s->esz == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, s->f[i],
                                dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS)
                return fail(LBM_ECUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        }
    }
    s->tmap_ok = true;
    return LBM_OK;
}

// Tensor maps of both population buffers for the sliding-window kernel: all planes of all cavities stacked into one
// 2-D tensor [batch * 9 * (nyl + 2)][pitch]; a box is the staged row width (512 B + 2 x 16 B halo) x 4 rows.
static int make_slide_maps(lbm_solver* s) {
    if (s->tmap_slide_state == 1) return LBM_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn)
        return fail(LBM_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    EncodeTiledFn encode = (EncodeTiledFn)fn;
    cuuint64_t dims[2] = {(cuuint64_t)s->pitch, (cuuint64_t)s->cfg.batch * 9 * (cuuint64_t)(s->nyl + 2)};
    cuuint64_t strides[1] = {(cuuint64_t)s->pitch * s->esz};
    cuuint32_t estr[2] = {1, 1};
    cuuint32_t box[2] = {(cuuint32_t)(512 / s->esz + 2 * (16 / s->esz)), 4};       // SlideCfg::SW x SlideCfg::R
    for (int i = 0; i < 2; ++i) {
        CUresult r = encode(&s->tmap_slide[i], s->esz == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                            2, s->f[i], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS)
            return fail(LBM_ECUDA, "cuTensorMapEncodeTiled (sliding-window kernel) failed with CUresult " + std::to_string((int)r));
    }
    s->tmap_slide_state = 1;
    return LBM_OK;
}

template <typename T, int COLL, int TY, int STAGES, int MINB>
static cudaError_t launch_tma_cfg(lbm_solver* s, const CUtensorMap* tm, const StepArgs& a, const TileSched& ts, cudaStream_t st) {
    constexpr int V = TmaVec<T>::V;
    using Cfg = TmaCfg<T, V, TY, STAGES>;
    auto kern = lbm_step_tma<T, COLL, false, V, TY, STAGES, MINB>;
    static AttrOnce once;
    if (cudaError_t e = once.ensure(kern, (int)Cfg::SMEM_BYTES)) return e;
    long long want = (long long)s->num_sms * s->tma_ctas_per_sm;
    int grid = (int)(ts.tiles_total < want ? ts.tiles_total : want);
    TileSched t2 = ts;
    const int per = ts.tiles_x * ts.tiles_y;
    t2.adv_b = grid / per;
    t2.adv_y = (grid - t2.adv_b * per) / ts.tiles_x;
    t2.adv_x = grid - t2.adv_b * per - t2.adv_y * ts.tiles_x;
    kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tm[0], tm[1], a, t2);
    return cudaGetLastError();
}

template <typename T, int COLL>
static cudaError_t launch_tma_variant(lbm_solver* s, const CUtensorMap* tm, const StepArgs& a, TileSched ts, cudaStream_t st) {
    constexpr int V = TmaVec<T>::V;
    constexpr int M = sizeof(T) == 4 ? 2 : 1;          // fp32 tiles: half as many threads per row, twice the rows
    constexpr int TX = (sizeof(T) == 8 ? 64 : 32) * V;
    const int ty = kTmaTY[s->tma_variant] * M;
    ts.tiles_x = (s->cfg.nx + TX - 1) / TX;
    ts.tiles_y = (ts.row_count + ty - 1) / ty;
    ts.tiles_total = (long long)ts.tiles_x * ts.tiles_y * s->cfg.batch;
    switch (s->tma_variant) {
        case 1: return launch_tma_cfg<T, COLL, 4 * M, 3, 2>(s, tm, a, ts, st);
        case 2: return launch_tma_cfg<T, COLL, 8 * M, 2, 1>(s, tm, a, ts, st);
        default: return launch_tma_cfg<T, COLL, 4 * M, 4, 1>(s, tm, a, ts, st);
    }
}

template <typename T>
static cudaError_t launch_tma_coll(lbm_solver* s, const CUtensorMap* tm, const StepArgs& a, const TileSched& ts, cudaStream_t st) {
    switch (s->cfg.collision) {
        case LBM_SRT: return launch_tma_variant<T, COLL_SRT>(s, tm, a, ts, st);
        case LBM_TRT: return launch_tma_variant<T, COLL_TRT>(s, tm, a, ts, st);
        default: return launch_tma_variant<T, COLL_MRT>(s, tm, a, ts, st);
    }
}

// ---- kernel dispatch ----------------------------------------------------------------------------------------
struct Launch {
    dim3 grid, block;
    cudaStream_t st;
    bool pdl;          // launch with programmatic stream serialization (step kernels only)
};

// Launch a step kernel.  With L.pdl the launch carries cudaLaunchAttributeProgrammaticStreamSerialization: the next
// step's CTAs may be scheduled while this one drains; they block in griddepcontrol.wait (first instruction of the
// kernels) until the previous grid has completed and flushed, so the A/B read/write ordering is unchanged.
template <typename K>
static void launch_step(K kern, const Launch& L, const StepArgs& a) {
    if (!L.pdl) {
        kern<<<L.grid, L.block, 0, L.st>>>(a);
        return;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = L.grid; cfg.blockDim = L.block; cfg.dynamicSmemBytes = 0; cfg.stream = L.st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, a);
}

template <typename T, int COLL, bool TURB>
static void dispatch_step(const StepArgs& a, const Launch& L, bool gather, bool macros) {
    if (gather) {
        if (macros) launch_step(lbm_step_ldg<T, COLL, true, true, MODE_STEP, TURB>, L, a);
        else launch_step(lbm_step_ldg<T, COLL, true, false, MODE_STEP, TURB>, L, a);
    } else {
        if (macros) launch_step(lbm_step_ldg<T, COLL, false, true, MODE_STEP, TURB>, L, a);
        else launch_step(lbm_step_ldg<T, COLL, false, false, MODE_STEP, TURB>, L, a);
    }
}

template <typename T, int COLL>
static void dispatch_flags(const StepArgs& a, const Launch& L, bool gather, bool macros, int mode) {
    if (mode == MODE_FINALIZE) { launch_step(lbm_step_ldg<T, COLL, true, false, MODE_FINALIZE>, L, a); return; }
    if (mode == MODE_MACROS) {
        if (gather) launch_step(lbm_step_ldg<T, COLL, true, true, MODE_MACROS>, L, a);
        else launch_step(lbm_step_ldg<T, COLL, false, true, MODE_MACROS>, L, a);
        return;
    }
    if (a.pi_eq) dispatch_step<T, COLL, true>(a, L, gather, macros);
    else dispatch_step<T, COLL, false>(a, L, gather, macros);
}

template <typename T>
static void dispatch_coll(int coll, const StepArgs& a, const Launch& L, bool gather, bool macros, int mode) {
    if (mode != MODE_STEP) { dispatch_flags<T, COLL_MRT>(a, L, gather, macros, mode); return; }
    switch (coll) {
        case LBM_SRT: dispatch_flags<T, COLL_SRT>(a, L, gather, macros, mode); break;
        case LBM_TRT: dispatch_flags<T, COLL_TRT>(a, L, gather, macros, mode); break;
        default: dispatch_flags<T, COLL_MRT>(a, L, gather, macros, mode); break;
    }
}

// Block shape for `threads_x` threads along a row: blockDim.x from {256,...,32} with the least idle lanes (ties go
// to the wider block), blockDim.y rows so that a block has up to 256 threads.
static void block_shape(int threads_x, int rows, int batch, Launch* L) {
    int best = 256, waste = 1 << 30;
    for (int bx = 256; bx >= 32; bx -= 32) {
        const int w = (threads_x + bx - 1) / bx * bx - threads_x;
        if (w < waste) { waste = w; best = bx; }
    }
    int by = 256 / best;
    if (by > rows) by = rows;
    if (by < 1) by = 1;
    L->block = dim3(best, by, 1);
    L->grid = dim3((threads_x + best - 1) / best, (rows + by - 1) / by, batch);
}

// ---- vector ldg dispatch (hot path) ---------------------------------------------------------------------------
template <typename T, int COLL, int V>
static void launch_vec_flags(const StepArgs& a, const Launch& L, bool macros) {
    if (a.pi_eq) {
        if (macros) launch_step(lbm_step_vec<T, COLL, true, V, true>, L, a);
        else launch_step(lbm_step_vec<T, COLL, false, V, true>, L, a);
    } else {
        if (macros) launch_step(lbm_step_vec<T, COLL, true, V>, L, a);
        else launch_step(lbm_step_vec<T, COLL, false, V>, L, a);
    }
}
template <typename T, int V>
static void launch_vec_coll(int coll, const StepArgs& a, const Launch& L, bool macros) {
    switch (coll) {
        case LBM_SRT: launch_vec_flags<T, COLL_SRT, V>(a, L, macros); break;
        case LBM_TRT: launch_vec_flags<T, COLL_TRT, V>(a, L, macros); break;
        default: launch_vec_flags<T, COLL_MRT, V>(a, L, macros); break;
    }
}

// Which two-step (temporal blocking) kernel advances this handle, if any (whole cavity or y-strip of >= 2 rows).
enum { TWO_NONE = 0, TWO_TILE = 1, TWO_SLIDE = 3 };
static int two_step_kind(const lbm_solver* s) {
    if (s->aa) return TWO_NONE;
    if (!s->use_fused2 || s->engine != LBM_ENGINE_LDG || s->cfg.semantics != LBM_SEMANTICS_C || s->nyl < 2) return TWO_NONE;
    const long long nodes = (long long)s->cfg.nx * s->cfg.ny * s->cfg.batch;
    if (nodes < s->fused2_min_nodes) return TWO_NONE;
    const bool whole = s->nyl == s->cfg.ny;
    // the closure's per-node state has no halo exchange: two-step with turb = 1 on whole cavities only
    if (s->use_slide && nodes >= s->slide_min_nodes && (!s->cfg.turb || whole)) return TWO_SLIDE;
    // shared-memory tiles: no closure, no frozen cavities
    if (s->cfg.turb || s->active) return TWO_NONE;
    // fp32 never gains from the tiles (ALU-bound): 1024^2 one-step 107 088 / tiles 98 466 MLUPS, narrower cavities alike
    if (s->esz == 4 && s->fused2_tile < 0) return TWO_NONE;
    if (s->fused2_tile < 0 && nodes >= LBM_FUSED2_SMALL_NODES && nodes < LBM_FUSED2_LARGE_NODES) return TWO_NONE;
    return TWO_TILE;
}
static bool fused2_capable(const lbm_solver* s) { return two_step_kind(s) != TWO_NONE; }
// ... and to lbm_step, which owns whole cavities only
static bool fused2_usable(const lbm_solver* s) { return fused2_capable(s) && s->nyl == s->cfg.ny; }

// Rows per segment of the sliding-window kernel: 4m + 2 (m + 1 iterations of four rows cover the segment and its two
// halo rows exactly).  Cost model fitted to tools/h_sweep_mid.py and tools/h_sweep_fine.py (20 shapes from 1024^2 to
// 32768 x 4096, both dtypes; worst regret of the pick against the best measured height 2 %):
//   time ~ [(h + 2.5) / h + m(h) 2 / h] x (W + 0.5) / W
//   (h + 2.5) / h   start-up of a segment: two halo rows and an empty pipeline
//   W               CTAs / (3 per SM x SMs), NOT rounded up: CTAs are handed out as others finish, so what a launch
//                   with few CTAs loses is the under-occupied tail, about half a CTA lifetime -> short segments
//                   (1024^2 fp64: 60 102 MLUPS at 14 rows, 58 027 at 22, 49 670 at 34; fp32 98 324 at 10, 58 131 at 34)
//   m(h) 2 / h      the halo rows are re-read by the segment below one CTA lifetime later; they come from L2 only
//                   while what the resident CTAs stream in between (h x 512 B x 18 each) stays well inside it --
//                   above ~80 MB the re-reads start to cost DRAM time (3072^2 fp64: 85 794 at 22, 82 889 at 34)
static int slide_seg_h(const lbm_solver* s) {
    if (s->slide_h > 0) return s->slide_h;
    const int tx = 512 / s->esz;
    const long long nsx = (s->cfg.nx + tx - 1) / tx;
    const double slots = 3.0 * s->num_sms;
    int best = 26;
    double best_cost = 1e300;
    for (int h = s->esz == 4 ? 10 : 14; h <= 30; h += 4) {
        const double items = (double)(nsx * ((s->nyl + h - 1) / h) * s->cfg.batch);
        const double waves = items / slots;
        const double streamed_mb = (items < slots ? items : slots) * h * 512.0 * 18.0 / 1e6;
        double miss = (streamed_mb - 80.0) / 100.0;
        miss = miss < 0.0 ? 0.0 : (miss > 1.0 ? 1.0 : miss);
        const double cost = ((h + 2.5) / h + miss * 2.0 / h) * (waves + 0.5) / waves;
        if (cost < best_cost - 1e-12) { best_cost = cost; best = h; }
    }
    return best;
}

// Tile-shape variant of the two-step kernel.  Defaults from the round-1 tile sweep at 4096^2: fp64 64x8 tiles at
// 4 CTAs/SM, fp32 32x16; cavities small enough to be launch-latency-bound take 32x8 tiles so that one wave still
// covers all SMs (384^2 -> 576 CTAs).
static int fused2_variant(const lbm_solver* s) {
    if (s->fused2_tile >= 0) return s->fused2_tile;
    if ((long long)s->cfg.nx * s->cfg.ny * s->cfg.batch < LBM_FUSED2_SMALL_NODES) return 9;
    return s->esz == 8 ? 3 : 4;
}

// ---- fused two-step launch ------------------------------------------------------------------------------------
template <typename T, int COLL, int TX, int TY, int MINB, bool GHOST2>
static cudaError_t launch_fused2_cfg(lbm_solver* s, const StepArgs& a, bool macros, cudaStream_t st) {
    using Cfg = Fused2Cfg<T, TX, TY>;
    static AttrOnce once_plain, once_macros;
    if (cudaError_t e = once_plain.ensure(lbm_step_fused2<T, COLL, false, TX, TY, MINB, GHOST2>, (int)Cfg::SMEM)) return e;
    if (cudaError_t e = once_macros.ensure(lbm_step_fused2<T, COLL, true, TX, TY, MINB, GHOST2>, (int)Cfg::SMEM)) return e;
    // a.row_begin / a.row_count arrive in LOCAL ROWS (multiples of the tile height, see lbm_step2_region): convert
    StepArgs t = a;
    t.row_begin = a.row_begin / Cfg::TY;
    const int tile_rows = (a.row_count + Cfg::TY - 1) / Cfg::TY;
    dim3 grid((s->cfg.nx + Cfg::TX - 1) / Cfg::TX, tile_rows, s->cfg.batch);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(256, 1, 1); cfg.dynamicSmemBytes = Cfg::SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = s->use_pdl ? 1 : 0;
    if (macros) return cudaLaunchKernelEx(&cfg, lbm_step_fused2<T, COLL, true, TX, TY, MINB, GHOST2>, t);
    return cudaLaunchKernelEx(&cfg, lbm_step_fused2<T, COLL, false, TX, TY, MINB, GHOST2>, t);
}

template <typename T, int COLL, int TX, int TY, int MINB>
static cudaError_t launch_fused2_tile(lbm_solver* s, const StepArgs& a, bool macros, cudaStream_t st) {
    // only a band that contains the first or last row of a strip that is not the whole cavity reads ghost2
    const bool strip = s->nyl != s->cfg.ny;
    const bool touches_end = a.row_begin == 0 || a.row_begin + a.row_count >= s->nyl;
    if (strip && touches_end) return launch_fused2_cfg<T, COLL, TX, TY, MINB, true>(s, a, macros, st);
    return launch_fused2_cfg<T, COLL, TX, TY, MINB, false>(s, a, macros, st);
}

template <typename T, int COLL>
static cudaError_t launch_fused2_t(lbm_solver* s, const StepArgs& a, bool macros, cudaStream_t st) {
    const int variant = fused2_variant(s);
    switch (variant) {
        case 9: return launch_fused2_tile<T, COLL, 32, 8, 4>(s, a, macros, st);
        case 1: return launch_fused2_tile<T, COLL, 64, 8, 3>(s, a, macros, st);
        case 2: return launch_fused2_tile<T, COLL, 32, 16, 3>(s, a, macros, st);
        case 3: return launch_fused2_tile<T, COLL, 64, 8, 4>(s, a, macros, st);
        case 4: return launch_fused2_tile<T, COLL, 32, 16, 4>(s, a, macros, st);
        default: return launch_fused2_tile<T, COLL, 64, 16, 2>(s, a, macros, st);
    }
}

// Tile height of the fused kernel in use (needed to cut a strip into edge / interior bands of whole tile rows).
static int fused2_tile_height(const lbm_solver* s) {
    if (two_step_kind(s) == TWO_SLIDE) return slide_seg_h(s);
    switch (fused2_variant(s)) {
        case 1: case 3: case 9: return 8;
        default: return 16;
    }
}

// Fused launch over local rows [row_begin, row_begin + row_count) (row_begin a multiple of the tile height); no
// state change -- the caller flips cur / side once every band of the double step has been launched.
// Sliding-window kernel only: band_h > 0 launches the two edge bands of the strip -- rows [0, band_h) and
// [nyl - band_h, nyl) -- as one grid of two segments.
static int launch_fused2_rows(lbm_solver* s, int row_begin, int row_count, bool macros, cudaStream_t st, int band_h = 0) {
    if (row_count <= 0) return LBM_OK;
    StepArgs a = make_args(s, s->f[s->cur], s->f[s->cur ^ 1]);
    a.row_begin = row_begin; a.row_count = row_count;
    a.ghost2 = (char*)s->f[s->cur] + (size_t)s->cfg.batch * s->cavity * s->esz;
    cudaError_t e;
    if (two_step_kind(s) == TWO_SLIDE) {
        a.seg_h = slide_seg_h(s);
        if (band_h > 0) { a.seg_h = band_h; a.seg_stride = row_count - band_h; }
        a.slide_tma = s->slide_tma;
        if (s->slide_tma) { int rc = make_slide_maps(s); if (rc) return rc; }
        Slide2Launch L{};
        L.tmap = s->slide_tma ? &s->tmap_slide[s->cur] : nullptr;
        L.coll = s->cfg.collision == LBM_SRT ? COLL_SRT : s->cfg.collision == LBM_TRT ? COLL_TRT : COLL_MRT;
        L.turb = s->cfg.turb != 0; L.macros = macros; L.batch = s->cfg.batch; L.pdl = s->use_pdl != 0; L.st = st;
        e = s->esz == 8 ? launch_slide2_f64(a, L) : launch_slide2_f32(a, L);
        if (e != cudaSuccess) return fail(LBM_ECUDA, std::string("sliding two-step launch: ") + cudaGetErrorString(e));
        s->launches++;
        return LBM_OK;
    }
    if (s->cfg.dtype == LBM_F64) {
        e = s->cfg.collision == LBM_SRT ? launch_fused2_t<double, COLL_SRT>(s, a, macros, st)
          : s->cfg.collision == LBM_TRT ? launch_fused2_t<double, COLL_TRT>(s, a, macros, st)
                                        : launch_fused2_t<double, COLL_MRT>(s, a, macros, st);
    } else {
        e = s->cfg.collision == LBM_SRT ? launch_fused2_t<float, COLL_SRT>(s, a, macros, st)
          : s->cfg.collision == LBM_TRT ? launch_fused2_t<float, COLL_TRT>(s, a, macros, st)
                                        : launch_fused2_t<float, COLL_MRT>(s, a, macros, st);
    }
    if (e != cudaSuccess) return fail(LBM_ECUDA, std::string("fused two-step launch: ") + cudaGetErrorString(e));
    s->launches++;
    return LBM_OK;
}

// Two steps in one launch over the whole strip: cur -> cur^1, side -> side^1.
static int launch_fused2(lbm_solver* s, bool macros, cudaStream_t st) {
    int rc = launch_fused2_rows(s, 0, s->nyl, macros, st);
    if (rc) return rc;
    s->cur ^= 1; s->side ^= 1; s->steps += 2;
    return LBM_OK;
}

// Launch one pass over a row region. rows: begin, count, stride.
static int launch_pass(lbm_solver* s, const void* src, void* dst, int row_begin, int row_count, int row_stride,
                       bool gather, bool macros, int mode, cudaStream_t st) {
    if (row_count <= 0) return LBM_OK;
    StepArgs a = make_args(s, src, dst);
    a.row_begin = row_begin; a.row_stride = row_stride;
    if (s->engine == LBM_ENGINE_TMA && s->tmap_ok && !s->cfg.turb && !s->active && mode == MODE_STEP && gather && !macros && row_stride == 1 &&
        (src == s->f[0] || src == s->f[1])) {
        TileSched ts{};
        ts.row_begin = row_begin; ts.row_count = row_count; ts.rows_per_plane = s->nyl + 2;
        const CUtensorMap* tm = s->tmap[src == s->f[0] ? 0 : 1];
        cudaError_t e = s->cfg.dtype == LBM_F64 ? launch_tma_coll<double>(s, tm, a, ts, st)
                                                : launch_tma_coll<float>(s, tm, a, ts, st);
        s->launches++;
        if (e != cudaSuccess) return fail(LBM_ECUDA, std::string("tma launch: ") + cudaGetErrorString(e));
        return LBM_OK;
    }
    a.row_count = row_count;
    int vw = s->cfg.dtype == LBM_F64 ? s->vec_f64 : s->vec_f32;
    if (vw == 0) {                                   // fp32, automatic
        const long long nodes = (long long)s->cfg.nx * s->nyl * s->cfg.batch;
        if (s->cfg.turb) vw = (nodes <= 200000 || nodes > 1200000) ? 1 : 2;
        else vw = nodes <= 200000 ? 1 : (nodes <= 1200000 ? 2 : 4);
    }
    const bool vec = mode == MODE_STEP && gather && vw > 1;
    Launch L{};
    L.st = st;
    L.pdl = s->use_pdl && mode == MODE_STEP;
    block_shape(vec ? (s->cfg.nx + vw - 1) / vw : s->cfg.nx, row_count, s->cfg.batch, &L);
    if (L.grid.y > 65535u) {
        // split over rows in chunks the grid can express
        const int chunk = 65535;
        for (int off = 0; off < row_count; off += chunk) {
            const int n = row_count - off < chunk ? row_count - off : chunk;
            int rc = launch_pass(s, src, dst, row_begin + off * row_stride, n, row_stride, gather, macros, mode, st);
            if (rc) return rc;
        }
        return LBM_OK;
    }
    if (vec) {
        if (s->cfg.dtype == LBM_F64) launch_vec_coll<double, 2>(s->cfg.collision, a, L, macros);
        else if (vw == 4) launch_vec_coll<float, 4>(s->cfg.collision, a, L, macros);
        else launch_vec_coll<float, 2>(s->cfg.collision, a, L, macros);
    } else if (s->cfg.dtype == LBM_F64) dispatch_coll<double>(s->cfg.collision, a, L, gather, macros, mode);
    else dispatch_coll<float>(s->cfg.collision, a, L, gather, macros, mode);
    s->launches++;
    CK(cudaGetLastError());
    return LBM_OK;
}

// ---- AA pattern (one population buffer) ---------------------------------------------------------------------------
template <typename T, int COLL, bool TURB>
static void dispatch_aa_step(const StepArgs& a, const Launch& L, bool odd, bool walls, bool macros) {
    if (odd) {
        if (macros) launch_step(lbm_step_aa<T, COLL, true, true, true, MODE_STEP, TURB>, L, a);
        else launch_step(lbm_step_aa<T, COLL, true, true, false, MODE_STEP, TURB>, L, a);
    } else if (walls) {
        if (macros) launch_step(lbm_step_aa<T, COLL, false, true, true, MODE_STEP, TURB>, L, a);
        else launch_step(lbm_step_aa<T, COLL, false, true, false, MODE_STEP, TURB>, L, a);
    } else {
        if (macros) launch_step(lbm_step_aa<T, COLL, false, false, true, MODE_STEP, TURB>, L, a);
        else launch_step(lbm_step_aa<T, COLL, false, false, false, MODE_STEP, TURB>, L, a);
    }
}

template <typename T, int COLL>
static void dispatch_aa_turb(const StepArgs& a, const Launch& L, bool odd, bool walls, bool macros) {
    if (a.pi_eq) dispatch_aa_step<T, COLL, true>(a, L, odd, walls, macros);
    else dispatch_aa_step<T, COLL, false>(a, L, odd, walls, macros);
}

template <typename T>
static void dispatch_aa(int coll, const StepArgs& a, const Launch& L, bool odd, bool walls, bool macros, int mode) {
    if (mode == MODE_FINALIZE) {
        if (odd) launch_step(lbm_step_aa<T, COLL_MRT, true, true, false, MODE_FINALIZE>, L, a);
        else launch_step(lbm_step_aa<T, COLL_MRT, false, true, false, MODE_FINALIZE>, L, a);
        return;
    }
    if (mode == MODE_MACROS) {
        if (odd) launch_step(lbm_step_aa<T, COLL_MRT, true, true, true, MODE_MACROS>, L, a);
        else if (walls) launch_step(lbm_step_aa<T, COLL_MRT, false, true, true, MODE_MACROS>, L, a);
        else launch_step(lbm_step_aa<T, COLL_MRT, false, false, true, MODE_MACROS>, L, a);
        return;
    }
    switch (coll) {
        case LBM_SRT: dispatch_aa_turb<T, COLL_SRT>(a, L, odd, walls, macros); break;
        case LBM_TRT: dispatch_aa_turb<T, COLL_TRT>(a, L, odd, walls, macros); break;
        default: dispatch_aa_turb<T, COLL_MRT>(a, L, odd, walls, macros); break;
    }
}

// One AA pass over the whole cavity.  MODE_STEP works in place on f[0]; MODE_FINALIZE writes the reference's `fin` into
// `out` (a second buffer borrowed for a download); MODE_MACROS only stores rho, u.  odd = the buffer is in the SWAPPED
// layout; walls = rebuild what a wall node cannot receive (false only for a freshly uploaded / initialised state).
static int launch_aa_pass(lbm_solver* s, bool odd, bool walls, bool macros, int mode, void* out, cudaStream_t st) {
    StepArgs a = make_args(s, s->f[0], mode == MODE_STEP ? s->f[0] : out);
    Launch L{};
    L.st = st;
    L.pdl = s->use_pdl && mode == MODE_STEP;
    for (int off = 0; off < s->nyl; off += 65535) {
        const int n = s->nyl - off < 65535 ? s->nyl - off : 65535;
        a.row_begin = off; a.row_stride = 1; a.row_count = n;
        block_shape(s->cfg.nx, n, s->cfg.batch, &L);
        if (s->cfg.dtype == LBM_F64) dispatch_aa<double>(s->cfg.collision, a, L, odd, walls, macros, mode);
        else dispatch_aa<float>(s->cfg.collision, a, L, odd, walls, macros, mode);
        s->launches++;
    }
    CK(cudaGetLastError());
    return LBM_OK;
}

// Row bands of a one-step region launch.  EDGE = the first TWO and the last TWO rows of the strip: the halo exchange
// that follows an EDGE launch (while INTERIOR is still running) ships rows 0, 1, nyl-2 and nyl-1 -- the second ones
// feed the neighbour's second ghost rows for the two-step kernel -- so all four must be complete by then.
static int region_rows(lbm_solver* s, int region, int rows[2][3], int* n) {
    // rows[i] = {begin, count, stride}
    const int nyl = s->nyl;
    *n = 0;
    const bool split = nyl >= 5;            // otherwise the edge bands are the whole strip
    if (region == LBM_REGION_ALL || (region == LBM_REGION_EDGE && !split)) {
        rows[0][0] = 0; rows[0][1] = nyl; rows[0][2] = 1; *n = 1;
    } else if (region == LBM_REGION_EDGE) {
        rows[0][0] = 0; rows[0][1] = 2; rows[0][2] = 1;
        rows[1][0] = nyl - 2; rows[1][1] = 2; rows[1][2] = 1;
        *n = 2;
    } else if (region == LBM_REGION_INTERIOR) {
        if (split) { rows[0][0] = 2; rows[0][1] = nyl - 4; rows[0][2] = 1; *n = 1; }
    } else {
        return fail(LBM_EINVAL, "bad region");
    }
    return LBM_OK;
}

// ------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------
extern "C" {

const char* lbm_last_error(void) { return g_err.c_str(); }
int lbm_abi_version(void) { return LBM_B200_ABI_VERSION; }

int lbm_device_count(int* count) {
    if (!count) return fail(LBM_EINVAL, "count == NULL");
    CK(cudaGetDeviceCount(count));
    return LBM_OK;
}

static int check_cfg(const lbm_config_t* c, int* nyl_out) {
    if (!c) return fail(LBM_EINVAL, "cfg == NULL");
    if (c->nx < 3 || c->ny < 3) return fail(LBM_EINVAL, "nx and ny must be >= 3");
    if (c->batch < 1 || c->batch > 65535) return fail(LBM_EINVAL, "batch must lie in [1, 65535] (grid.z = cavity)");
    if (c->dtype != LBM_F32 && c->dtype != LBM_F64) return fail(LBM_EINVAL, "dtype must be LBM_F32 or LBM_F64");
    if (c->collision < LBM_SRT || c->collision > LBM_MRT) return fail(LBM_EINVAL, "bad collision");
    if (c->turb != 0 && c->turb != 1) return fail(LBM_EINVAL, "turb must be 0 or 1");
    int nyl = c->ny_local == 0 ? c->ny : c->ny_local;
    if (c->ny_local == 0 && c->y0 != 0) return fail(LBM_EINVAL, "y0 must be 0 when ny_local == 0");
    if (c->y0 < 0 || nyl < 1 || c->y0 + nyl > c->ny) return fail(LBM_EINVAL, "y-strip [y0, y0+ny_local) outside [0, ny)");
    if (c->engine < LBM_ENGINE_AUTO || c->engine > LBM_ENGINE_AA) return fail(LBM_EINVAL, "bad engine");
    if (c->semantics != LBM_SEMANTICS_C && c->semantics != LBM_SEMANTICS_A) return fail(LBM_EINVAL, "bad semantics");
    if (c->reserved != 0) return fail(LBM_EINVAL, "reserved must be 0");
    if ((c->ext_f[0] == nullptr) != (c->ext_f[1] == nullptr))
        return fail(LBM_EINVAL, "ext_f: give both population buffers or neither");
    if (c->ext_f[0] && c->ext_f[0] == c->ext_f[1]) return fail(LBM_EINVAL, "ext_f: the two population buffers must differ");
    if (c->semantics == LBM_SEMANTICS_A) {
        if (c->collision != LBM_SRT || c->turb) return fail(LBM_EINVAL, "semantics A (MRT.py) is SRT without turbulence model");
        if (nyl != c->ny) return fail(LBM_EINVAL, "semantics A does not support y-strips");
        if (c->ny > 65535) return fail(LBM_EINVAL, "semantics A supports ny <= 65535");
    }
    if (c->engine == LBM_ENGINE_AA) {
        if (c->semantics != LBM_SEMANTICS_C) return fail(LBM_EINVAL, "the AA pattern implements semantics C");
        if (nyl != c->ny) return fail(LBM_EINVAL, "the AA pattern holds whole cavities (no y-strips)");
        if (c->ext_f[0]) return fail(LBM_EINVAL, "the AA pattern owns its single population buffer (no ext_f)");
    }
    *nyl_out = nyl;
    return LBM_OK;
}

static void layout_of(const lbm_config_t* c, int nyl, lbm_layout_t* L) {
    L->elem_size = c->dtype == LBM_F64 ? 8 : 4;
    L->pitch = ((int64_t)c->nx + 31) / 32 * 32;
    L->rows = nyl + 2;
    L->plane = L->rows * L->pitch;
    L->cavity = 9 * L->plane;
    L->ghost2_offset = (int64_t)c->batch * L->cavity;            // tail: [batch][top|bottom][3][pitch]
    L->state_bytes = ((int64_t)c->batch * L->cavity + (int64_t)c->batch * 6 * L->pitch) * L->elem_size;
}

int lbm_state_bytes(const lbm_config_t* cfg, size_t* bytes) {
    int nyl;
    int rc = check_cfg(cfg, &nyl);
    if (rc) return rc;
    if (!bytes) return fail(LBM_EINVAL, "bytes == NULL");
    lbm_layout_t L;
    layout_of(cfg, nyl, &L);
    *bytes = (size_t)L.state_bytes;
    return LBM_OK;
}

int lbm_destroy(lbm_handle_t s) {
    if (!s) return LBM_OK;
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    if (s->own_f) { cudaFree(s->f[0]); cudaFree(s->f[1]); }
    cudaFree(s->rho); cudaFree(s->ux); cudaFree(s->uy);
    cudaFree(s->rho_lid); cudaFree(s->carry); cudaFree(s->cav);
    cudaFree(s->pi_eq); cudaFree(s->rho_prev);
    cudaFree(s->active); cudaFree(s->usum); cudaFree(s->conv_past); cudaFree(s->conv_count);
    cudaFree(s->staging); cudaFree(s->scratch);
    for (int i = 0; i < 4; ++i) if (s->graph[i]) cudaGraphExecDestroy(s->graph[i]);
    if (s->capture_stream) cudaStreamDestroy(s->capture_stream);
    if (s->copy_stream) {
        cudaStreamDestroy(s->copy_stream);
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(s->ev_slot_full[i]); cudaEventDestroy(s->ev_slot_free[i]); }
    }
    delete s;
    return LBM_OK;
}

int lbm_create(const lbm_config_t* cfg, lbm_handle_t* out) {
    int nyl;
    int rc = check_cfg(cfg, &nyl);
    if (rc) return rc;
    if (!out) return fail(LBM_EINVAL, "out == NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(LBM_ECUDA, std::string("no usable CUDA device (there is no CPU fallback): ") +
                                   (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    lbm_solver* s = new (std::nothrow) lbm_solver();
    if (!s) return fail(LBM_ENOMEM, "host allocation failed");
    s->cfg = *cfg;
    if (cfg->device < 0) {
        if (cudaGetDevice(&s->device) != cudaSuccess) { delete s; return fail(LBM_ECUDA, "cudaGetDevice failed"); }
    } else {
        s->device = cfg->device;
    }
    if (s->device >= ndev) { delete s; return fail(LBM_EINVAL, "device ordinal out of range"); }
    lbm_layout_t L;
    layout_of(cfg, nyl, &L);
    s->esz = (int)L.elem_size; s->pitch = (int)L.pitch; s->nyl = nyl;
    s->plane = L.plane; s->cavity = L.cavity; s->mplane = (long long)nyl * L.pitch;
    s->state_bytes = (size_t)L.state_bytes;
    s->engine = cfg->engine == LBM_ENGINE_TMA ? LBM_ENGINE_TMA : LBM_ENGINE_LDG;   // AUTO -> ldg
    s->aa = cfg->engine == LBM_ENGINE_AA;
#define CKD(call)                                                                       \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            std::string m__ = std::string(#call) + ": " + cudaGetErrorString(e__);      \
            lbm_destroy(s);                                                             \
            return fail(e__ == cudaErrorMemoryAllocation ? LBM_ENOMEM : LBM_ECUDA, m__); \
        }                                                                               \
    } while (0)
    CKD(cudaSetDevice(s->device));
    if (cfg->ext_f[0] && cfg->ext_f[1]) {
        s->f[0] = cfg->ext_f[0]; s->f[1] = cfg->ext_f[1]; s->own_f = false;
    } else {
        s->own_f = true;
        CKD(cudaMalloc(&s->f[0], s->state_bytes));
        if (!s->aa) CKD(cudaMalloc(&s->f[1], s->state_bytes));
    }
    // ghost rows and pitch padding must hold finite values: they are read (and discarded) by masked lanes only in
    // the TMA family, but zero them once for determinism.
    CKD(cudaMemset(s->f[0], 0, s->state_bytes));
    if (s->f[1]) CKD(cudaMemset(s->f[1], 0, s->state_bytes));
    const size_t mbytes = (size_t)cfg->batch * s->mplane * s->esz;
    CKD(cudaMalloc(&s->rho, mbytes));
    CKD(cudaMalloc(&s->ux, mbytes));
    CKD(cudaMalloc(&s->uy, mbytes));
    CKD(cudaMemset(s->rho, 0, mbytes)); CKD(cudaMemset(s->ux, 0, mbytes)); CKD(cudaMemset(s->uy, 0, mbytes));
    CKD(cudaMalloc(&s->rho_lid, 2 * (size_t)cfg->batch * s->pitch * s->esz));      // two halves (see `side`)
    CKD(cudaMemset(s->rho_lid, 0, 2 * (size_t)cfg->batch * s->pitch * s->esz));
    CKD(cudaMalloc(&s->carry, 2 * (size_t)cfg->batch * 4 * s->esz));
    CKD(cudaMemset(s->carry, 0, 2 * (size_t)cfg->batch * 4 * s->esz));
    CKD(cudaMalloc(&s->cav, sizeof(CavityParams) * cfg->batch));
    if (cfg->turb) {       // two halves each: the two-step kernel reads one and writes the other (see `side`)
        CKD(cudaMalloc(&s->pi_eq, 2 * mbytes));
        CKD(cudaMalloc(&s->rho_prev, 2 * mbytes));
        CKD(cudaMemset(s->pi_eq, 0, 2 * mbytes));
        CKD(cudaMemset(s->rho_prev, 0, 2 * mbytes));
    }
#undef CKD
    s->cav_host.resize(cfg->batch);
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, s->device) == cudaSuccess) s->num_sms = prop.multiProcessorCount;
    }
    if (s->engine == LBM_ENGINE_TMA) {
        int trc = make_tensor_maps(s);
        if (trc) { std::string m = g_err; lbm_destroy(s); return fail(trc, m); }
    }
    *out = s;
    // defaults of the reference GPU script: Re = 100 placeholder, uLB = 0.08 (MRT_GPU.py:58)
    rc = lbm_set_reynolds(s, -1, 0.08, 100.0);
    if (rc) { lbm_destroy(s); *out = nullptr; return rc; }
    return LBM_OK;
}

int lbm_get_layout(lbm_handle_t s, lbm_layout_t* out) {
    if (!s || !out) return fail(LBM_EINVAL, "NULL argument");
    layout_of(&s->cfg, s->nyl, out);
    return LBM_OK;
}

static void drop_graphs(lbm_solver* s) {
    for (int i = 0; i < 4; ++i)
        if (s->graph[i]) { cudaGraphExecDestroy(s->graph[i]); s->graph[i] = nullptr; }
}

int lbm_set_tuning(lbm_handle_t s, const char* key, int64_t value) {
    if (!s || !key) return fail(LBM_EINVAL, "NULL argument");
    const std::string k(key);
    const int v = (int)value;
    if (k == "two_step") s->use_fused2 = v != 0;
    else if (k == "slide") s->use_slide = v != 0;
    else if (k == "slide_tma") s->slide_tma = v != 0;
    else if (k == "slide_h") { if (v < 0 || v > 4096) return fail(LBM_EINVAL, "slide_h must lie in [0, 4096]"); s->slide_h = v; }
    else if (k == "slide_min_nodes") s->slide_min_nodes = value;
    else if (k == "tile") { if (v < -1 || v > 9) return fail(LBM_EINVAL, "tile must lie in [-1, 9]"); s->fused2_tile = v; }
    else if (k == "two_step_min_nodes") s->fused2_min_nodes = value;
    else if (k == "vec_f64") { if (v != 1 && v != 2) return fail(LBM_EINVAL, "vec_f64 must be 1 or 2"); s->vec_f64 = v; }
    else if (k == "vec_f32") { if (v != 0 && v != 1 && v != 2 && v != 4) return fail(LBM_EINVAL, "vec_f32 must be 0 (by size), 1, 2 or 4"); s->vec_f32 = v; }
    else if (k == "graph") s->use_graph = v != 0;
    else if (k == "pdl") s->use_pdl = v != 0;
    else if (k == "tma_ctas") { if (v < 1) return fail(LBM_EINVAL, "tma_ctas must be >= 1"); s->tma_ctas_per_sm = v; }
    else if (k == "tma_variant") {
        if (v < 0 || v >= LBM_TMA_VARIANTS) return fail(LBM_EINVAL, "tma_variant out of range");
        s->tma_variant = v;
        if (s->engine == LBM_ENGINE_TMA) { int rc = make_tensor_maps(s); if (rc) return rc; }
    }
    else return fail(LBM_EINVAL, "unknown tuning key '" + k + "'");
    drop_graphs(s);          // captured launch sequences depend on every one of these
    return LBM_OK;
}

int lbm_set_rates(lbm_handle_t s, int cavity, double uLB, double omega_nu, double omega_e, double omega_eps,
                  double omega_q, double omega_minus) {
    if (!s) return fail(LBM_EINVAL, "NULL handle");
    if (cavity < -1 || cavity >= s->cfg.batch) return fail(LBM_EINVAL, "cavity index out of range");
    if (!(omega_nu > 0.0 && omega_nu < 2.0)) return fail(LBM_EINVAL, "omega_nu must lie in (0, 2)");
    const int b0 = cavity < 0 ? 0 : cavity, b1 = cavity < 0 ? s->cfg.batch : cavity + 1;
    for (int b = b0; b < b1; ++b) {
        CavityParams& p = s->cav_host[b];
        p.uLB = uLB; p.omega = omega_nu; p.omegam = omega_minus;
        p.s_e = omega_e; p.s_eps = omega_eps; p.s_q = omega_q;
        p.tau0 = 1.0 / omega_nu;
    }
    s->cav_dirty = true;
    return LBM_OK;
}

int lbm_set_reynolds(lbm_handle_t s, int cavity, double uLB, double Re) {
    if (!s) return fail(LBM_EINVAL, "NULL handle");
    if (!(Re > 0.0)) return fail(LBM_EINVAL, "Re must be positive");
    const double nuLB = uLB * s->cfg.ny / Re;                 // functions.pyx:41, MRT_GPU.py:63
    const double omega = 2.0 / (6. * nuLB + 1);               // functions.pyx:43, MRT_GPU.py:65
    const double delTRT = 1.0 / 3.5;                          // MRT_GPU.py:83
    const double omegam = 1.0 / (0.5 + (delTRT / ((1 / omega) - 0.5)));   // MRT_GPU.py:84
    return lbm_set_rates(s, cavity, uLB, omega, 1.0, 1.2, 1.2, omegam);   // MRT_GPU.py:88-91
}

int lbm_init_equilibrium(lbm_handle_t s) {
    if (!s) return fail(LBM_EINVAL, "NULL handle");
    int rc = set_device(s);
    if (rc) return rc;
    rc = sync_params(s, 0);
    if (rc) return rc;
    s->side = 0;
    StepArgs a = make_args(s, nullptr, s->f[0]);
    dim3 grid((s->cfg.nx + 255) / 256, 1, s->cfg.batch);
    for (int off = 0; off < s->nyl; off += 65535) {
        // blockIdx.y covers rows [off, off+n): shift through y0/dst offsets is avoided by a small loop
        const int n = s->nyl - off < 65535 ? s->nyl - off : 65535;
        StepArgs b = a;
        b.y0 = a.y0 + off;
        b.dst = (char*)a.dst + (size_t)off * s->pitch * s->esz;
        b.rho = (char*)a.rho + (size_t)off * s->pitch * s->esz;
        b.ux = (char*)a.ux + (size_t)off * s->pitch * s->esz;
        b.uy = (char*)a.uy + (size_t)off * s->pitch * s->esz;
        if (a.pi_eq) {
            b.pi_eq = (char*)a.pi_eq + (size_t)off * s->pitch * s->esz;
            b.rho_prev = (char*)a.rho_prev + (size_t)off * s->pitch * s->esz;
        }
        grid.y = n;
        if (s->cfg.dtype == LBM_F64) lbm_init_eq<double><<<grid, 256>>>(b);
        else lbm_init_eq<float><<<grid, 256>>>(b);
        s->launches++;
    }
    CK(cudaGetLastError());
    s->cur = 0; s->pre = true; s->steps = 0; s->aa_swapped = false;
    reset_active(s);
    return LBM_OK;
}

static int ensure_staging(lbm_solver* s, size_t bytes) {
    if (s->staging_bytes >= bytes) return LBM_OK;
    if (s->staging) { CK(cudaFree(s->staging)); s->staging = nullptr; s->staging_bytes = 0; }
    CK(cudaMalloc(&s->staging, bytes));
    s->staging_bytes = bytes;
    return LBM_OK;
}

static void launch_transpose(lbm_solver* s, bool to_device, void* dev, void* lin, int planes, long long dev_plane_stride,
                             long long dev_row0, cudaStream_t st) {
    const int nx = s->cfg.nx, nyl = s->nyl;
    dim3 blk(32, 8), grid((nx + 31) / 32, (nyl + 31) / 32, planes);
    if (s->esz == 8) {
        if (to_device) lbm_transpose<double, true><<<grid, blk, 0, st>>>((double*)dev, (double*)lin, nx, nyl, s->pitch, dev_plane_stride, dev_row0);
        else lbm_transpose<double, false><<<grid, blk, 0, st>>>((double*)dev, (double*)lin, nx, nyl, s->pitch, dev_plane_stride, dev_row0);
    } else {
        if (to_device) lbm_transpose<float, true><<<grid, blk, 0, st>>>((float*)dev, (float*)lin, nx, nyl, s->pitch, dev_plane_stride, dev_row0);
        else lbm_transpose<float, false><<<grid, blk, 0, st>>>((float*)dev, (float*)lin, nx, nyl, s->pitch, dev_plane_stride, dev_row0);
    }
    s->launches++;
}

// Move `ncav` cavities of `planes` [nx][nyl] planes each between a reference-layout array (`lin`: host, or device
// when on_device) and device planes.  lin_cavity_bytes: distance between cavities in `lin`; dev_*_stride in elements;
// dev_row0: element offset of local row 0 inside a device plane.
// Host arrays go plane by plane through two device staging slots: the copy engine (internal copy stream) moves plane
// i+1 over the host link while the transpose kernel of plane i runs on the caller's stream -- the layout change costs
// no time on top of the copies.  Device-resident arrays are transposed in place by one launch per cavity.
static int move_planes(lbm_solver* s, void* lin_base, size_t lin_cavity_bytes, bool on_device, bool to_device,
                       int ncav, int planes, void* dev_base, long long dev_plane_stride, long long dev_cavity_stride,
                       long long dev_row0, cudaStream_t st) {
    const int nx = s->cfg.nx, nyl = s->nyl;
    if (on_device) {
        for (int b = 0; b < ncav; ++b) {
            launch_transpose(s, to_device, (char*)dev_base + (size_t)b * dev_cavity_stride * s->esz,
                             (char*)lin_base + (size_t)b * lin_cavity_bytes, planes, dev_plane_stride, dev_row0, st);
            CK(cudaGetLastError());
        }
        return LBM_OK;
    }
    const size_t pbytes = (size_t)nx * nyl * s->esz;                     // one plane in the reference layout
    int rc = ensure_staging(s, 2 * pbytes);
    if (rc) return rc;
    if (!s->copy_stream) {
        CK(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CK(cudaEventCreateWithFlags(&s->ev_slot_full[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&s->ev_slot_free[i], cudaEventDisableTiming));
        }
    }
    cudaStream_t cs = s->copy_stream;
    // the copy stream starts after everything already queued on the caller's stream (e.g. the finalize pass)
    CK(cudaEventRecord(s->ev_slot_free[0], st));
    CK(cudaStreamWaitEvent(cs, s->ev_slot_free[0], 0));
    int i = 0;
    for (int b = 0; b < ncav; ++b) {
        for (int k = 0; k < planes; ++k, ++i) {
            const int slot = i & 1;
            char* h = (char*)lin_base + (size_t)b * lin_cavity_bytes + (size_t)k * pbytes;
            char* lin = (char*)s->staging + (size_t)slot * pbytes;
            char* dev = (char*)dev_base + ((size_t)b * dev_cavity_stride + (size_t)k * dev_plane_stride) * s->esz;
            if (to_device) {
                if (i >= 2) CK(cudaStreamWaitEvent(cs, s->ev_slot_free[slot], 0));      // its last transpose has read it
                CK(cudaMemcpyAsync(lin, h, pbytes, cudaMemcpyHostToDevice, cs));
                CK(cudaEventRecord(s->ev_slot_full[slot], cs));
                CK(cudaStreamWaitEvent(st, s->ev_slot_full[slot], 0));
                launch_transpose(s, true, dev, lin, 1, dev_plane_stride, dev_row0, st);
                CK(cudaEventRecord(s->ev_slot_free[slot], st));
            } else {
                if (i >= 2) CK(cudaStreamWaitEvent(st, s->ev_slot_free[slot], 0));      // its last copy out has left
                launch_transpose(s, false, dev, lin, 1, dev_plane_stride, dev_row0, st);
                CK(cudaEventRecord(s->ev_slot_full[slot], st));
                CK(cudaStreamWaitEvent(cs, s->ev_slot_full[slot], 0));
                CK(cudaMemcpyAsync(h, lin, pbytes, cudaMemcpyDeviceToHost, cs));
                CK(cudaEventRecord(s->ev_slot_free[slot], cs));
            }
            CK(cudaGetLastError());
        }
    }
    CK(cudaStreamSynchronize(cs));
    CK(cudaStreamSynchronize(st));
    return LBM_OK;
}

int lbm_upload_f(lbm_handle_t s, const void* f, int on_device, void* stream) {
    if (!s || !f) return fail(LBM_EINVAL, "NULL argument");
    int rc = set_device(s);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    rc = sync_params(s, st);
    if (rc) return rc;
    const size_t fcav = (size_t)9 * s->cfg.nx * s->nyl * s->esz;
    rc = move_planes(s, const_cast<void*>(f), fcav, on_device != 0, true, s->cfg.batch, 9, s->f[0], s->plane, s->cavity,
                     s->pitch, st);
    if (rc) return rc;
    s->side = 0;
    StepArgs a = make_args(s, s->f[0], nullptr);
    if (s->esz == 8) lbm_seed_carry<double><<<s->cfg.batch, 4, 0, st>>>(a);
    else lbm_seed_carry<float><<<s->cfg.batch, 4, 0, st>>>(a);
    const long long n = (long long)s->cfg.batch * s->mplane;
    if (s->esz == 8) {
        lbm_fill<double><<<1024, 256, 0, st>>>((double*)s->rho, n, 1.0);
        lbm_fill<double><<<1024, 256, 0, st>>>((double*)s->ux, n, 0.0);
        lbm_fill<double><<<1024, 256, 0, st>>>((double*)s->uy, n, 0.0);
    } else {
        lbm_fill<float><<<1024, 256, 0, st>>>((float*)s->rho, n, 1.0f);
        lbm_fill<float><<<1024, 256, 0, st>>>((float*)s->ux, n, 0.0f);
        lbm_fill<float><<<1024, 256, 0, st>>>((float*)s->uy, n, 0.0f);
    }
    s->launches += 4;
    if (s->cfg.turb) {
        dim3 g((s->cfg.nx + 255) / 256, s->nyl, s->cfg.batch);
        if (s->esz == 8) lbm_seed_turb<double><<<g, 256, 0, st>>>(a);
        else lbm_seed_turb<float><<<g, 256, 0, st>>>(a);
        s->launches++;
    }
    CK(cudaGetLastError());
    s->cur = 0; s->pre = true; s->steps = 0; s->aa_swapped = false;
    reset_active(s);
    return LBM_OK;
}

int lbm_download_f(lbm_handle_t s, void* f, int on_device, void* stream) {
    if (!s || !f) return fail(LBM_EINVAL, "NULL argument");
    int rc = set_device(s);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    rc = sync_params(s, st);
    if (rc) return rc;
    if (s->aa) {
        const size_t fc = (size_t)9 * s->cfg.nx * s->nyl * s->esz;
        if (s->pre) return move_planes(s, f, fc, on_device != 0, false, s->cfg.batch, 9, s->f[0], s->plane, s->cavity, s->pitch, st);
        // the reference's `fin` (gather + wall rule, no collision) needs a second buffer: borrowed for this call only
        if (!s->scratch) CK(cudaMalloc(&s->scratch, s->state_bytes));
        rc = launch_aa_pass(s, s->aa_swapped, true, false, MODE_FINALIZE, s->scratch, st);
        if (rc == LBM_OK)
            rc = move_planes(s, f, fc, on_device != 0, false, s->cfg.batch, 9, s->scratch, s->plane, s->cavity, s->pitch, st);
        cudaFree(s->scratch);              // synchronises the device: everything queued above has run
        s->scratch = nullptr;
        return rc;
    }
    void* fin = s->f[s->cur];
    if (!s->pre) {
        // gather + wall rule (no collision) into the buffer the next step will overwrite anyway -- except with frozen
        // cavities, whose two buffers must both keep the post-collision state (no step rewrites them): scratch then
        void* dst = s->f[s->cur ^ 1];
        if (s->active) {
            if (!s->scratch) CK(cudaMalloc(&s->scratch, s->state_bytes));
            dst = s->scratch;
        }
        rc = launch_pass(s, s->f[s->cur], dst, 0, s->nyl, 1, true, false, MODE_FINALIZE, st);
        if (rc) return rc;
        fin = dst;
    }
    const size_t fcav = (size_t)9 * s->cfg.nx * s->nyl * s->esz;
    return move_planes(s, f, fcav, on_device != 0, false, s->cfg.batch, 9, fin, s->plane, s->cavity, s->pitch, st);
}

// One step of semantics A: collide (f[0] -> f[1], rho, u), then stream + walls (f[1], f[0] -> f[0] in place).
static int step_A(lbm_solver* s, cudaStream_t st) {
    dim3 grid((s->cfg.nx + 255) / 256, s->cfg.ny, s->cfg.batch);
    StepArgs a = make_args(s, s->f[0], s->f[1]);
    StepArgs b = make_args(s, s->f[1], s->f[0]);
    if (s->esz == 8) { lbm_A_collide<double><<<grid, 256, 0, st>>>(a); lbm_A_stream_bc<double><<<grid, 256, 0, st>>>(b); }
    else { lbm_A_collide<float><<<grid, 256, 0, st>>>(a); lbm_A_stream_bc<float><<<grid, 256, 0, st>>>(b); }
    s->launches += 2;
    CK(cudaGetLastError());
    s->steps++;
    return LBM_OK;
}

int lbm_step_region(lbm_handle_t s, int region, int write_macros, void* stream) {
    if (!s) return fail(LBM_EINVAL, "NULL handle");
    if (s->cfg.semantics == LBM_SEMANTICS_A) return fail(LBM_ESTATE, "semantics A has no region stepping: use lbm_step");
    if (s->aa) return fail(LBM_ESTATE, "the AA pattern has no region stepping: use lbm_step");
    int rc = set_device(s);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    rc = sync_params(s, st);
    if (rc) return rc;
    int rows[2][3], n;
    rc = region_rows(s, region, rows, &n);
    if (rc) return rc;
    for (int i = 0; i < n; ++i) {
        rc = launch_pass(s, s->f[s->cur], s->f[s->cur ^ 1], rows[i][0], rows[i][1], rows[i][2], !s->pre,
                         write_macros != 0, MODE_STEP, st);
        if (rc) return rc;
    }
    return LBM_OK;
}

int lbm_swap(lbm_handle_t s) {
    if (!s) return fail(LBM_EINVAL, "NULL handle");
    if (s->cfg.semantics == LBM_SEMANTICS_A || s->aa) return fail(LBM_ESTATE, "semantics A and the AA pattern have no region stepping: use lbm_step");
    s->cur ^= 1; s->pre = false; s->steps++;
    return LBM_OK;
}

int lbm_step2_available(lbm_handle_t s) { return s && fused2_capable(s) && !s->pre ? 1 : 0; }

int lbm_step2_region(lbm_handle_t s, int region, int write_macros, void* stream) {
    if (!s) return fail(LBM_EINVAL, "NULL handle");
    if (!fused2_capable(s) || s->pre)
        return fail(LBM_ESTATE, "two-step kernel not available for this handle/state (see lbm_step2_available)");
    int rc = set_device(s);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    rc = sync_params(s, st);
    if (rc) return rc;
    const int nyl = s->nyl;
    const bool wm = write_macros != 0;
    if (two_step_kind(s) == TWO_SLIDE) {
        // sliding-window kernel: segments of any height give the same bits, so the edge bands are as thin as the kernel
        // allows (4m + 2 rows, m = 1: two iterations) and go out as ONE launch of two segments per column strip; the
        // interior takes the rows in between with the height its own size asks for.  (With bands one interior segment
        // tall the edge pass of a 32768 x 4096 strip took 134 us on the priority stream, 4 % of the pass.)
        const int eb = 6;
        const bool split = nyl >= 4 * eb;
        if (region == LBM_REGION_ALL || (region == LBM_REGION_EDGE && !split)) return launch_fused2_rows(s, 0, nyl, wm, st);
        if (region == LBM_REGION_EDGE) return launch_fused2_rows(s, 0, nyl, wm, st, eb);
        if (region == LBM_REGION_INTERIOR) return split ? launch_fused2_rows(s, eb, nyl - 2 * eb, wm, st) : LBM_OK;
        return fail(LBM_EINVAL, "bad region");
    }
    const int ty = fused2_tile_height(s);
    const int ntr = (nyl + ty - 1) / ty;                              // tile rows of the strip
    const int nb = (nyl % ty == 1 && ntr > 1) ? 2 : 1;                // bottom band must contain rows nyl-2 and nyl-1
    const bool split = ntr > 1 + nb;                                  // otherwise the edge bands are the whole strip
    const int bot0 = (ntr - nb) * ty;                                 // first row of the bottom band
    if (region == LBM_REGION_ALL || (region == LBM_REGION_EDGE && !split)) return launch_fused2_rows(s, 0, nyl, wm, st);
    if (region == LBM_REGION_EDGE) {
        rc = launch_fused2_rows(s, 0, ty, wm, st);
        if (rc) return rc;
        return launch_fused2_rows(s, bot0, nyl - bot0, wm, st);
    }
    if (region == LBM_REGION_INTERIOR) return split ? launch_fused2_rows(s, ty, bot0 - ty, wm, st) : LBM_OK;
    return fail(LBM_EINVAL, "bad region");
}

int lbm_swap2(lbm_handle_t s) {
    if (!s) return fail(LBM_EINVAL, "NULL handle");
    if (s->cfg.semantics == LBM_SEMANTICS_A || s->aa) return fail(LBM_ESTATE, "semantics A and the AA pattern have no region stepping: use lbm_step");
    s->cur ^= 1; s->side ^= 1; s->steps += 2;
    return LBM_OK;
}

// Stored rows and populations of the nine halo rows on one side (order of distributed.halo_plan(deep=True)):
// send: up {0,1,3,2,5,6} of local row 0 + {2,5,6} of local row 1; down {0,1,3,4,7,8} of row nyl-1 + {4,7,8} of nyl-2.
// recv: the same populations land in the ghost row of that side and in its three second ghost rows.
static HaloRows halo_rows(const lbm_solver* s, int dir, bool pack) {
    static const int up[9] = {0, 1, 3, 2, 5, 6, 2, 5, 6}, dn[9] = {0, 1, 3, 4, 7, 8, 4, 7, 8};
    HaloRows h{};
    for (int i = 0; i < 9; ++i) {
        if (pack) {
            h.pop[i] = dir == 0 ? up[i] : dn[i];
            h.row[i] = dir == 0 ? (i < 6 ? 1 : 2) : (i < 6 ? s->nyl : s->nyl - 1);       // stored row = local row + 1
        } else {        // what arrives from above are the neighbour's DOWN-going rows, and vice versa
            h.pop[i] = dir == 0 ? dn[i] : up[i];
            h.row[i] = dir == 0 ? 0 : s->nyl + 1;
        }
    }
    return h;
}

static int halo_move(lbm_solver* s, int dir, void* buf, bool pack, cudaStream_t st) {
    if (!buf) return fail(LBM_EINVAL, "NULL buffer");
    if (dir != 0 && dir != 1) return fail(LBM_EINVAL, "dir must be 0 (strip above) or 1 (strip below)");
    if (s->cfg.batch != 1 || s->nyl < 2 || s->aa) return fail(LBM_ESTATE, "packed halo rows need a single-cavity A/B strip of >= 2 rows");
    int rc = set_device(s);
    if (rc) return rc;
    char* f = (char*)s->f[s->cur ^ 1];                                  // the buffer written by the step in progress
    char* g2 = f + ((size_t)s->cfg.batch * s->cavity + (size_t)dir * 3 * s->pitch) * s->esz;
    const HaloRows h = halo_rows(s, dir, pack);
    dim3 grid((s->cfg.nx + 255) / 256, 9);
    if (s->esz == 8) {
        if (pack) lbm_halo_rows<double, true><<<grid, 256, 0, st>>>((double*)f, (double*)g2, (double*)buf, h, s->cfg.nx, s->pitch, s->plane);
        else lbm_halo_rows<double, false><<<grid, 256, 0, st>>>((double*)f, (double*)g2, (double*)buf, h, s->cfg.nx, s->pitch, s->plane);
    } else {
        if (pack) lbm_halo_rows<float, true><<<grid, 256, 0, st>>>((float*)f, (float*)g2, (float*)buf, h, s->cfg.nx, s->pitch, s->plane);
        else lbm_halo_rows<float, false><<<grid, 256, 0, st>>>((float*)f, (float*)g2, (float*)buf, h, s->cfg.nx, s->pitch, s->plane);
    }
    s->launches++;
    CK(cudaGetLastError());
    return LBM_OK;
}

int lbm_halo_pack(lbm_handle_t s, int dir, void* buf, void* stream) {
    if (!s) return fail(LBM_EINVAL, "NULL handle");
    return halo_move(s, dir, buf, true, (cudaStream_t)stream);
}

int lbm_halo_unpack(lbm_handle_t s, int dir, const void* buf, void* stream) {
    if (!s) return fail(LBM_EINVAL, "NULL handle");
    return halo_move(s, dir, const_cast<void*>(buf), false, (cudaStream_t)stream);
}

int lbm_buffer_ptr(lbm_handle_t s, int which, void** ptr) {
    if (!s || !ptr) return fail(LBM_EINVAL, "NULL argument");
    *ptr = s->aa ? s->f[0] : s->f[which ? (s->cur ^ 1) : s->cur];
    return LBM_OK;
}

// Steps per graph launch (even, so a graph leaves the A/B parity unchanged).  The per-launch CPU cost of a plain
// stream launch (~2.5 us) is what bounds small cavities such as 384^2 (kernel ~3 us); a graph of 32 steps amortises it.
#define LBM_GRAPH_STEPS 32
#define LBM_GRAPH_MAX_NODES (1 << 22)   // only launch-latency-bound sizes take the graph path



// Capture LBM_GRAPH_STEPS steady steps (fused two-step launches when usable) starting from the current (cur, side);
// an even number of launches of either kind returns to the same (cur, side), so the graph is re-launchable as is.
static int build_graph(lbm_solver* s, int key) {
    if (!s->capture_stream) CK(cudaStreamCreateWithFlags(&s->capture_stream, cudaStreamNonBlocking));
    cudaGraph_t g = nullptr;
    CK(cudaStreamBeginCapture(s->capture_stream, cudaStreamCaptureModeThreadLocal));
    int rc = LBM_OK;
    const int cur0 = s->cur, side0 = s->side;
    const int64_t launches0 = s->launches, steps0 = s->steps;
    const bool fused = fused2_usable(s);
    for (int i = 0; i < LBM_GRAPH_STEPS && rc == LBM_OK; i += fused ? 2 : 1) {
        if (fused) {
            rc = launch_fused2(s, false, s->capture_stream);
        } else {
            rc = launch_pass(s, s->f[s->cur], s->f[s->cur ^ 1], 0, s->nyl, 1, true, false, MODE_STEP, s->capture_stream);
            s->cur ^= 1;
        }
    }
    // counted when the graph is launched, not when it is captured
    s->cur = cur0; s->side = side0; s->launches = launches0; s->steps = steps0;
    cudaError_t e = cudaStreamEndCapture(s->capture_stream, &g);
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) return fail(LBM_ECUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(&s->graph[key], g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(LBM_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    return LBM_OK;
}

int lbm_step(lbm_handle_t s, int nsteps, int write_macros, void* stream) {
    if (!s) return fail(LBM_EINVAL, "NULL handle");
    if (nsteps < 0) return fail(LBM_EINVAL, "nsteps < 0");
    if (s->nyl != s->cfg.ny && nsteps > 1)
        return fail(LBM_ESTATE, "a y-strip handle needs a halo exchange between steps: use lbm_step_region/lbm_swap");
    if (nsteps == 0) return LBM_OK;
    int rc = set_device(s);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    rc = sync_params(s, st);
    if (rc) return rc;
    if (s->cfg.semantics == LBM_SEMANTICS_A) {
        for (int i = 0; i < nsteps; ++i) { rc = step_A(s, st); if (rc) return rc; }
        return LBM_OK;
    }
    if (s->aa) {
        for (int i = 0; i < nsteps; ++i) {
            rc = launch_aa_pass(s, s->aa_swapped, !s->pre, write_macros && i == nsteps - 1, MODE_STEP, nullptr, st);
            if (rc) return rc;
            s->aa_swapped = !s->aa_swapped; s->pre = false; s->steps++;
        }
        return LBM_OK;
    }
    int left = nsteps;
    const bool small = (long long)s->cfg.nx * s->nyl * s->cfg.batch <= LBM_GRAPH_MAX_NODES;
    while (left > 0) {
        if (!s->pre) {
            // steady state in blocks of LBM_GRAPH_STEPS: one graph launch (launch-latency-bound sizes only)
            if (s->use_graph && small && left > LBM_GRAPH_STEPS) {
                const int key = s->cur * 2 + s->side;
                const bool fused = fused2_usable(s);
                if (!s->graph[key]) { rc = build_graph(s, key); if (rc) return rc; }
                CK(cudaGraphLaunch(s->graph[key], st));
                s->launches += fused ? LBM_GRAPH_STEPS / 2 : LBM_GRAPH_STEPS;
                s->steps += LBM_GRAPH_STEPS;
                left -= LBM_GRAPH_STEPS;
                continue;
            }
            // temporal blocking: two steps per launch
            if (left >= 2 && fused2_usable(s)) {
                rc = launch_fused2(s, write_macros && left == 2, st);
                if (rc) return rc;
                left -= 2;
                continue;
            }
        }
        rc = lbm_step_region(s, LBM_REGION_ALL, (write_macros && left == 1) ? 1 : 0, stream);
        if (rc) return rc;
        lbm_swap(s);
        --left;
    }
    return LBM_OK;
}

static int macros_out(lbm_solver* s, void* rho, void* u, int on_device, cudaStream_t st) {
    const size_t pl = (size_t)s->cfg.nx * s->nyl * s->esz;   // one [nx][nyl] plane
    int rc;
    if (rho) {
        rc = move_planes(s, rho, pl, on_device != 0, false, s->cfg.batch, 1, s->rho, s->mplane, s->mplane, 0, st);
        if (rc) return rc;
    }
    if (u) {   // u[b][0] = ux, u[b][1] = uy
        rc = move_planes(s, u, 2 * pl, on_device != 0, false, s->cfg.batch, 1, s->ux, s->mplane, s->mplane, 0, st);
        if (rc) return rc;
        rc = move_planes(s, (char*)u + pl, 2 * pl, on_device != 0, false, s->cfg.batch, 1, s->uy, s->mplane, s->mplane, 0, st);
        if (rc) return rc;
    }
    return LBM_OK;
}

int lbm_get_macros(lbm_handle_t s, void* rho, void* u, int on_device, void* stream) {
    if (!s) return fail(LBM_EINVAL, "NULL handle");
    int rc = set_device(s);
    if (rc) return rc;
    return macros_out(s, rho, u, on_device, (cudaStream_t)stream);
}

int lbm_get_macros_current(lbm_handle_t s, void* rho, void* u, int on_device, void* stream) {
    if (!s) return fail(LBM_EINVAL, "NULL handle");
    int rc = set_device(s);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    rc = sync_params(s, st);
    if (rc) return rc;
    if (s->aa) rc = launch_aa_pass(s, s->aa_swapped, !s->pre, true, MODE_MACROS, nullptr, st);
    else rc = launch_pass(s, s->f[s->cur], s->f[s->cur ^ 1], 0, s->nyl, 1, !s->pre, true, MODE_MACROS, st);
    if (rc) return rc;
    return macros_out(s, rho, u, on_device, st);
}

int lbm_get_feq(lbm_handle_t s, void* feq, int on_device, void* stream) {
    if (!s || !feq) return fail(LBM_EINVAL, "NULL argument");
    int rc = set_device(s);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    // equilibrium of the stored rho, u in the device layout of the macro planes, then the usual layout change
    if (!s->scratch) CK(cudaMalloc(&s->scratch, s->state_bytes));            // >= 9 macro planes per cavity
    for (int b = 0; b < s->cfg.batch; ++b) {
        const size_t mo = (size_t)b * s->mplane * s->esz;
        char* out = (char*)s->scratch + 9 * mo;
        const int nb = (int)((s->mplane + 255) / 256 < 148 * 16 ? (s->mplane + 255) / 256 : 148 * 16);
        if (s->esz == 8) lbm_equ_kernel<double><<<nb, 256, 0, st>>>((const double*)((char*)s->rho + mo), (const double*)((char*)s->ux + mo), (const double*)((char*)s->uy + mo), (double*)out, s->mplane);
        else lbm_equ_kernel<float><<<nb, 256, 0, st>>>((const float*)((char*)s->rho + mo), (const float*)((char*)s->ux + mo), (const float*)((char*)s->uy + mo), (float*)out, s->mplane);
        s->launches++;
    }
    CK(cudaGetLastError());
    const size_t fcav = (size_t)9 * s->cfg.nx * s->nyl * s->esz;
    const int rc2 = move_planes(s, feq, fcav, on_device != 0, false, s->cfg.batch, 9, s->scratch, s->mplane, 9 * s->mplane, 0, st);
    if (s->aa) { cudaFree(s->scratch); s->scratch = nullptr; }     // an AA handle keeps no second buffer around
    return rc2;
}

int lbm_equilibrium(int dtype, int64_t n, const void* rho, const void* ux, const void* uy, void* feq, int on_device,
                    void* stream) {
    if (dtype != LBM_F32 && dtype != LBM_F64) return fail(LBM_EINVAL, "bad dtype");
    if (n < 0 || !rho || !ux || !uy || !feq) return fail(LBM_EINVAL, "bad argument");
    if (n == 0) return LBM_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(LBM_ECUDA, "no usable CUDA device (there is no CPU fallback)");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t esz = dtype == LBM_F64 ? 8 : 4;
    void* d = nullptr;
    const void *dr = rho, *dx = ux, *dy = uy;
    void* df = feq;
    if (!on_device) {
        CK(cudaMalloc(&d, 12 * n * esz));
        char* c = (char*)d;
        CK(cudaMemcpyAsync(c, rho, n * esz, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(c + n * esz, ux, n * esz, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(c + 2 * n * esz, uy, n * esz, cudaMemcpyHostToDevice, st));
        dr = c; dx = c + n * esz; dy = c + 2 * n * esz; df = c + 3 * n * esz;
    }
    const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    if (dtype == LBM_F64) lbm_equ_kernel<double><<<blocks, 256, 0, st>>>((const double*)dr, (const double*)dx, (const double*)dy, (double*)df, n);
    else lbm_equ_kernel<float><<<blocks, 256, 0, st>>>((const float*)dr, (const float*)dx, (const float*)dy, (float*)df, n);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && !on_device) {
        e = cudaMemcpyAsync(feq, df, 9 * n * esz, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    if (d) cudaFree(d);
    if (e != cudaSuccess) return fail(LBM_ECUDA, std::string("lbm_equilibrium: ") + cudaGetErrorString(e));
    return LBM_OK;
}

static int argmin_pass(lbm_solver* s, int cavity, int bc, const int box[4], long long* flat, cudaStream_t st) {
    const int nblk = 148 * 2;
    ArgMin* d = nullptr;
    CK(cudaMalloc(&d, sizeof(ArgMin) * nblk));
    const char* ux = (const char*)s->ux + (size_t)cavity * s->mplane * s->esz;
    const char* uy = (const char*)s->uy + (size_t)cavity * s->mplane * s->esz;
    if (s->esz == 8) lbm_argmin_usq<double><<<nblk, 256, 0, st>>>((const double*)ux, (const double*)uy, d, s->cfg.nx, s->cfg.ny, s->pitch, bc, box[0], box[1], box[2], box[3]);
    else lbm_argmin_usq<float><<<nblk, 256, 0, st>>>((const float*)ux, (const float*)uy, d, s->cfg.nx, s->cfg.ny, s->pitch, bc, box[0], box[1], box[2], box[3]);
    s->launches++;
    std::vector<ArgMin> h(nblk);
    cudaError_t e = cudaMemcpyAsync(h.data(), d, sizeof(ArgMin) * nblk, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (e != cudaSuccess) return fail(LBM_ECUDA, std::string("argmin: ") + cudaGetErrorString(e));
    double best = 1e300; long long bi = -1;
    for (const ArgMin& a : h)
        if (a.idx >= 0 && (bi < 0 || a.val < best || (a.val == best && a.idx < bi))) { best = a.val; bi = a.idx; }
    *flat = bi;
    return LBM_OK;
}

int lbm_diagnostics(lbm_handle_t s, int cavity, void* ux_col, void* uy_row, int32_t* vortex_xy, void* stream) {
    if (!s) return fail(LBM_EINVAL, "NULL handle");
    if (cavity < 0 || cavity >= s->cfg.batch) return fail(LBM_EINVAL, "cavity index out of range");
    int rc = set_device(s);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int nx = s->cfg.nx, ny = s->cfg.ny, nyl = s->nyl;
    if (ux_col || uy_row) {
        const int yc = ny / 2 - s->cfg.y0;                       // int(ysize/2), MRT_GPU.py:799
        const bool have_row = uy_row && yc >= 0 && yc < nyl;
        const size_t nbytes = (size_t)(nx + nyl) * s->esz;
        rc = ensure_staging(s, nbytes);
        if (rc) return rc;
        char* dcol = (char*)s->staging;
        char* drow = dcol + (size_t)nyl * s->esz;
        const char* ux = (const char*)s->ux + (size_t)cavity * s->mplane * s->esz;
        const char* uy = (const char*)s->uy + (size_t)cavity * s->mplane * s->esz;
        const int n = nx > nyl ? nx : nyl;
        if (s->esz == 8) lbm_centerlines<double><<<(n + 255) / 256, 256, 0, st>>>((const double*)ux, (const double*)uy, (double*)dcol, (double*)drow, nx, nyl, s->pitch, nx / 2, have_row ? yc : -1);
        else lbm_centerlines<float><<<(n + 255) / 256, 256, 0, st>>>((const float*)ux, (const float*)uy, (float*)dcol, (float*)drow, nx, nyl, s->pitch, nx / 2, have_row ? yc : -1);
        s->launches++;
        CK(cudaGetLastError());
        if (ux_col) CK(cudaMemcpyAsync(ux_col, dcol, (size_t)nyl * s->esz, cudaMemcpyDeviceToHost, st));
        if (have_row) CK(cudaMemcpyAsync(uy_row, drow, (size_t)nx * s->esz, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (uy_row && !have_row) return fail(LBM_ESTATE, "row y = ny/2 is not owned by this y-strip");
    }
    if (vortex_xy) {
        if (nyl != ny) return fail(LBM_ESTATE, "the vortex search needs the whole cavity in one handle");
        const int bc = nx / 40;                                   // BCoffset, MRT_GPU.py:767
        const int none[4] = {0, 0, 0, 0};
        long long f1 = -1, f2 = -1;
        rc = argmin_pass(s, cavity, bc, none, &f1, st);
        if (rc) return rc;
        if (f1 < 0) return fail(LBM_ESTATE, "cavity too small for the vortex search (everything masked)");
        const int x1 = (int)(f1 / ny), y1 = (int)(f1 % ny);
        const int box[4] = {x1 - bc, x1 + bc, y1 - bc, y1 + bc};  // MRT_GPU.py:774
        rc = argmin_pass(s, cavity, bc, box, &f2, st);
        if (rc) return rc;
        vortex_xy[0] = x1; vortex_xy[1] = y1;
        vortex_xy[2] = f2 < 0 ? -1 : (int)(f2 / ny);
        vortex_xy[3] = f2 < 0 ? -1 : (int)(f2 % ny);
    }
    return LBM_OK;
}

int lbm_mean_u(lbm_handle_t s, double* mean_out, void* stream) {
    if (!s || !mean_out) return fail(LBM_EINVAL, "NULL argument");
    int rc = set_device(s);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = s->cfg.batch;
    if (!s->usum) CK(cudaMalloc(&s->usum, sizeof(double) * nb));
    CK(cudaMemsetAsync(s->usum, 0, sizeof(double) * nb, st));
    dim3 grid(148 * 2, nb);
    if (s->esz == 8) lbm_sum_u<double><<<grid, 256, 0, st>>>((const double*)s->ux, (const double*)s->uy, s->usum, s->cfg.nx, s->nyl, s->pitch, s->mplane);
    else lbm_sum_u<float><<<grid, 256, 0, st>>>((const float*)s->ux, (const float*)s->uy, s->usum, s->cfg.nx, s->nyl, s->pitch, s->mplane);
    s->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(mean_out, s->usum, sizeof(double) * nb, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const double denom = 2.0 * (double)s->cfg.nx * (double)s->nyl;
    for (int b = 0; b < nb; ++b) mean_out[b] /= denom;
    return LBM_OK;
}

int lbm_set_active(lbm_handle_t s, const int32_t* active, void* stream) {
    if (!s || !active) return fail(LBM_EINVAL, "NULL argument");
    if (s->aa) return fail(LBM_ESTATE, "the AA pattern cannot freeze cavities (one buffer, one phase for the whole batch)");
    if (s->pre) return fail(LBM_ESTATE, "cavities can be frozen only after at least one step (post-collision state)");
    int rc = set_device(s);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = s->cfg.batch;
    if (s->active && s->conv_past) {   // lbm_converge_check may have retired cavities on the device since the last read-back
        CK(cudaMemcpyAsync(s->active_host.data(), s->active, sizeof(int) * nb, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    for (int b = 0; b < nb; ++b)       // validate everything before touching any state
        if (s->active && !s->active_host[b] && active[b]) return fail(LBM_ESTATE, "a frozen cavity cannot be re-activated");
    if (!s->active) {
        CK(cudaMalloc(&s->active, sizeof(int) * nb));
        s->active_host.assign(nb, 1);
        drop_graphs(s);                // kernel args change
    }
    const size_t cav_bytes = (size_t)s->cavity * s->esz;
    const size_t rl_bytes = (size_t)s->pitch * s->esz, rl_half = (size_t)nb * rl_bytes;
    const size_t ca_bytes = 4 * (size_t)s->esz, ca_half = (size_t)nb * ca_bytes;
    const size_t m_bytes = (size_t)s->mplane * s->esz, m_half = (size_t)nb * m_bytes;
    const int o = s->side ^ 1;
    for (int b = 0; b < nb; ++b) {
        const int now = active[b] ? 1 : 0;
        if (s->active_host[b] && !now) {
            // freeze: both A/B buffers and both halves of the side arrays must hold the cavity's current state,
            // whatever the parities later on
            CK(cudaMemcpyAsync((char*)s->f[s->cur ^ 1] + b * cav_bytes, (char*)s->f[s->cur] + b * cav_bytes, cav_bytes,
                               cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync((char*)s->rho_lid + o * rl_half + b * rl_bytes, (char*)s->rho_lid + s->side * rl_half + b * rl_bytes,
                               rl_bytes, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync((char*)s->carry + o * ca_half + b * ca_bytes, (char*)s->carry + s->side * ca_half + b * ca_bytes,
                               ca_bytes, cudaMemcpyDeviceToDevice, st));
            if (s->pi_eq) {
                CK(cudaMemcpyAsync((char*)s->pi_eq + o * m_half + b * m_bytes, (char*)s->pi_eq + s->side * m_half + b * m_bytes,
                                   m_bytes, cudaMemcpyDeviceToDevice, st));
                CK(cudaMemcpyAsync((char*)s->rho_prev + o * m_half + b * m_bytes,
                                   (char*)s->rho_prev + s->side * m_half + b * m_bytes, m_bytes, cudaMemcpyDeviceToDevice, st));
            }
        }
        s->active_host[b] = now;
    }
    CK(cudaMemcpyAsync(s->active, s->active_host.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    return LBM_OK;
}

int lbm_converge_check(lbm_handle_t s, double tol, int hits, int32_t* active_out, void* stream) {
    if (!s) return fail(LBM_EINVAL, "NULL handle");
    if (!(tol > 0.0) || hits < 1) return fail(LBM_EINVAL, "tol must be positive and hits >= 1");
    if (s->aa) return fail(LBM_ESTATE, "the AA pattern cannot retire cavities: use an A/B handle for the convergence rule");
    if (s->pre) return fail(LBM_ESTATE, "the convergence check needs at least one step (stored velocity field)");
    if (s->nyl != s->cfg.ny) return fail(LBM_ESTATE, "the convergence check needs whole cavities");
    int rc = set_device(s);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    rc = sync_params(s, st);
    if (rc) return rc;
    const int nb = s->cfg.batch;
    if (!s->active) {       // first use: every cavity active (kernel arguments change: drop captured graphs)
        CK(cudaMalloc(&s->active, sizeof(int) * nb));
        s->active_host.assign(nb, 1);
        CK(cudaMemcpyAsync(s->active, s->active_host.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
        drop_graphs(s);
    }
    if (!s->conv_past) {    // u_past = zeros, count = 0 (MRT_GPU_datagen.py:223, 705)
        CK(cudaMalloc(&s->conv_past, sizeof(double) * nb));
        CK(cudaMalloc(&s->conv_count, sizeof(int) * 2 * nb));
        CK(cudaMemsetAsync(s->conv_past, 0, sizeof(double) * nb, st));
        CK(cudaMemsetAsync(s->conv_count, 0, sizeof(int) * 2 * nb, st));
    }
    if (!s->usum) CK(cudaMalloc(&s->usum, sizeof(double) * nb));
    CK(cudaMemsetAsync(s->usum, 0, sizeof(double) * nb, st));
    dim3 grid(148 * 2, nb);
    if (s->esz == 8) lbm_sum_u<double><<<grid, 256, 0, st>>>((const double*)s->ux, (const double*)s->uy, s->usum, s->cfg.nx, s->nyl, s->pitch, s->mplane);
    else lbm_sum_u<float><<<grid, 256, 0, st>>>((const float*)s->ux, (const float*)s->uy, s->usum, s->cfg.nx, s->nyl, s->pitch, s->mplane);
    int* newly = s->conv_count + nb;
    lbm_converge_rule<<<(nb + 127) / 128, 128, 0, st>>>(s->usum, 2.0 * (double)s->cfg.nx * (double)s->nyl, s->cav, s->conv_past,
                                                       s->conv_count, s->active, newly, tol, hits, nb);
    // retired cavities: both A/B buffers and both halves of the side arrays must hold their final state
    const size_t cav_bytes = (size_t)s->cavity * s->esz;
    const size_t rl_bytes = (size_t)s->pitch * s->esz, rl_half = (size_t)nb * rl_bytes;
    const size_t ca_bytes = 4 * (size_t)s->esz, ca_half = (size_t)nb * ca_bytes;
    const size_t m_bytes = (size_t)s->mplane * s->esz, m_half = (size_t)nb * m_bytes;
    const int o = s->side ^ 1;
    dim3 cg(64, nb);
    lbm_freeze_copy<<<cg, 256, 0, st>>>(newly, (const char*)s->f[s->cur], (char*)s->f[s->cur ^ 1], cav_bytes);
    lbm_freeze_copy<<<dim3(1, nb), 256, 0, st>>>(newly, (const char*)s->rho_lid + s->side * rl_half, (char*)s->rho_lid + o * rl_half, rl_bytes);
    lbm_freeze_copy<<<dim3(1, nb), 32, 0, st>>>(newly, (const char*)s->carry + s->side * ca_half, (char*)s->carry + o * ca_half, ca_bytes);
    if (s->pi_eq) {
        lbm_freeze_copy<<<cg, 256, 0, st>>>(newly, (const char*)s->pi_eq + s->side * m_half, (char*)s->pi_eq + o * m_half, m_bytes);
        lbm_freeze_copy<<<cg, 256, 0, st>>>(newly, (const char*)s->rho_prev + s->side * m_half, (char*)s->rho_prev + o * m_half, m_bytes);
    }
    s->launches += s->pi_eq ? 7 : 5;
    CK(cudaGetLastError());
    if (active_out) {
        CK(cudaMemcpyAsync(s->active_host.data(), s->active, sizeof(int) * nb, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int b = 0; b < nb; ++b) active_out[b] = s->active_host[b];
    }
    return LBM_OK;
}

int lbm_sync(lbm_handle_t s) {
    if (!s) return fail(LBM_EINVAL, "NULL handle");
    int rc = set_device(s);
    if (rc) return rc;
    CK(cudaDeviceSynchronize());
    return LBM_OK;
}

int lbm_get_counters(lbm_handle_t s, int64_t* steps_done, int64_t* kernel_launches) {
    if (!s) return fail(LBM_EINVAL, "NULL handle");
    if (steps_done) *steps_done = s->steps;
    if (kernel_launches) *kernel_launches = s->launches;
    return LBM_OK;
}

const char* lbm_engine_name(lbm_handle_t s) {
    if (!s) return "";
    return s->aa ? "aa" : (s->engine == LBM_ENGINE_TMA ? "tma" : "ldg");
}

}  // extern "C"
