// "tma" kernel family: persistent, warp-specialised fused step with TMA-staged shifted neighbour reads (sm_100a).
//
// One producer warp per CTA drives the Tensor Memory Accelerator: for every tile of TX x TY nodes it issues nine
// 2-D bulk-tensor loads (cp.async.bulk.tensor.2d), one per population.  The row shift of the pull step (y + c_ky)
// is done by the copy engine's address generation (box origin row).  The column shift cannot be: measured on
// B200, a box whose innermost coordinate is not 16-byte aligned raises "illegal instruction"
// (tools/probes/tma_probe.cu), so populations with c_kx != 0 are fetched as a box one 128-byte line wider
// ([x0-EXT, x0+TX) for c_kx = +1, [x0, x0+TX+EXT) for c_kx = -1) and the one-element shift is applied when the compute
// warps read shared memory (two aligned 128-bit LDS + a compile-time select).  Boxes that hang over x < 0 or
// x >= pitch are zero-filled by the hardware; rows never leave the buffer because every population plane carries a
// ghost row above and below.  A STAGES-deep ring of {9 boxes, full mbarrier, empty
// mbarrier} keeps ~100 KB of loads in flight per SM independent of register pressure; compute warps release a slot
// as soon as its values are in registers, then apply the wall rule / moments / collision and store with 128-bit
// coalesced STG.  Tiles are assigned round-robin to gridDim.x persistent CTAs (a multiple of the SM count).
//
// The population buffer is described to TMA as a 2-D tensor [batch*9*(ny_local+2)][pitch] (planes are contiguous).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "lbm_device.cuh"

namespace lbm {

struct TileSched {
    int tiles_x, tiles_y;        // tiles per cavity
    long long tiles_total;       // tiles_x * tiles_y * batch
    int row_begin, row_count;    // local rows covered: [row_begin, row_begin + row_count)
    int rows_per_plane;          // ny_local + 2
    int adv_b, adv_y, adv_x;     // gridDim.x decomposed as adv_b * tiles_x*tiles_y + adv_y * tiles_x + adv_x
};

// Position of a persistent CTA in the tile sequence (x fastest, then y, then cavity), advanced by gridDim.x per
// iteration without integer division.
struct TileIter {
    int b, ty, tx;
    long long t;
    __device__ __forceinline__ TileIter(const TileSched& ts) {
        t = blockIdx.x;
        const int per = ts.tiles_x * ts.tiles_y;
        b = (int)(t / per);
        const int r = (int)(t - (long long)b * per);
        ty = r / ts.tiles_x;
        tx = r - ty * ts.tiles_x;
    }
    __device__ __forceinline__ bool valid(const TileSched& ts) const { return t < ts.tiles_total; }
    __device__ __forceinline__ void next(const TileSched& ts) {
        t += gridDim.x;
        tx += ts.adv_x;
        if (tx >= ts.tiles_x) { tx -= ts.tiles_x; ++ty; }
        ty += ts.adv_y;
        if (ty >= ts.tiles_y) { ty -= ts.tiles_y; ++b; }
        b += ts.adv_b;
    }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

template <typename T, int V> struct VecOf;
template <> struct VecOf<double, 2> { using type = double2; };
template <> struct VecOf<float, 4> { using type = float4; };
template <> struct VecOf<float, 2> { using type = float2; };
template <> struct VecOf<double, 1> { using type = double; };

template <typename T, int V>
__device__ __forceinline__ void vec_load_smem(const T* p, T out[V]) {
    using VT = typename VecOf<T, V>::type;
    const VT v = *reinterpret_cast<const VT*>(p);
    const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
    for (int i = 0; i < V; ++i) out[i] = e[i];
}
// 128-bit shared-memory load from a 32-bit shared address (keeps the access an LDS instead of a generic LD)
__device__ __forceinline__ void lds_vec(uint32_t addr, double out[2]) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(out[0]), "=d"(out[1]) : "r"(addr));
}
__device__ __forceinline__ void lds_vec(uint32_t addr, float out[4]) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(out[0]), "=f"(out[1]), "=f"(out[2]), "=f"(out[3]) : "r"(addr));
}

template <typename T, int V>
__device__ __forceinline__ void vec_store_global(T* p, const T in[V]) {
    using VT = typename VecOf<T, V>::type;
    VT v;
    T* e = reinterpret_cast<T*>(&v);
#pragma unroll
    for (int i = 0; i < V; ++i) e[i] = in[i];
    *reinterpret_cast<VT*>(p) = v;
}

// lattice velocities as compile-time tables (MRT.py:138)
__device__ __forceinline__ constexpr int lat_cx(int k) { return k == 1 || k == 5 || k == 8 ? 1 : (k == 3 || k == 6 || k == 7 ? -1 : 0); }
__device__ __forceinline__ constexpr int lat_cy(int k) { return k == 2 || k == 5 || k == 6 ? 1 : (k == 4 || k == 7 || k == 8 ? -1 : 0); }

template <typename T, int V, int TY, int STAGES>
struct TmaCfg {
    static constexpr int TXT = sizeof(T) == 8 ? 64 : 32;   // threads along x per tile row (box width TX+V <= 256)
    static constexpr int TX = TXT * V;             // nodes along x per tile
    static constexpr int EXT = 128 / sizeof(T);    // extra columns of a wide box: one full 128-byte line, so that every
                                                   // box row starts and ends on a line boundary (16-byte-aligned but
                                                   // line-misaligned rows ran the TMA path at a fraction of its rate)
    static constexpr int TXW = TX + EXT;           // columns of a wide box
    static constexpr int CONSUMERS = TXT * TY;     // compute threads
    static constexpr int THREADS = CONSUMERS + 32; // + one producer warp
    static constexpr uint32_t NARROW_BYTES = TX * TY * sizeof(T);
    static constexpr uint32_t WIDE_BYTES = TXW * TY * sizeof(T);
    static constexpr uint32_t WIDE_SLOT = (WIDE_BYTES + 127) / 128 * 128;       // TMA destinations are 128-byte aligned
    static constexpr uint32_t STAGE_BYTES = 3 * NARROW_BYTES + 6 * WIDE_SLOT;   // shared memory per stage
    static constexpr uint32_t TX_BYTES = 3 * NARROW_BYTES + 6 * WIDE_BYTES;     // bytes the nine loads deliver
    static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 2 * STAGES * sizeof(uint64_t) + 128;
    // byte offset of population k inside a stage: k = 0,2,4 narrow (slots 0..2), the six others wide
    __host__ __device__ static constexpr uint32_t slot_off(int k) {
        return k == 0 ? 0 : k == 2 ? NARROW_BYTES : k == 4 ? 2 * NARROW_BYTES
             : 3 * NARROW_BYTES + WIDE_SLOT * (k == 1 ? 0 : k == 3 ? 1 : k - 3);   // 1,3,5,6,7,8 -> 0..5
    }
};

template <typename T, int COLL, bool MACROS, int V, int TY, int STAGES, int MINB>
__global__ void __launch_bounds__(TmaCfg<T, V, TY, STAGES>::THREADS, MINB)
lbm_step_tma(const __grid_constant__ CUtensorMap tmap_narrow, const __grid_constant__ CUtensorMap tmap_wide,
             const StepArgs a, const TileSched ts) {
    using Cfg = TmaCfg<T, V, TY, STAGES>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty = full + STAGES;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    constexpr int NCW = Cfg::CONSUMERS / 32;       // consumer warps
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == NCW) {
        // ---------------- producer warp: one elected lane feeds the ring ----------------
        if ((tid & 31) == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (TileIter it(ts); it.valid(ts); it.next(ts)) {
                const int b = it.b;
                const int x0 = it.tx * Cfg::TX;
                const int row0 = ts.row_begin + it.ty * TY + 1;       // stored row of the tile's first node row
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_expect_tx(&full[stage], Cfg::TX_BYTES);
                unsigned char* dst = base + (size_t)stage * Cfg::STAGE_BYTES;
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const int c1 = (b * 9 + k) * ts.rows_per_plane + row0 + lat_cy(k);
                    if (lat_cx(k) == 0) tma_load_2d(dst + Cfg::slot_off(k), &tmap_narrow, x0, c1, &full[stage]);
                    else tma_load_2d(dst + Cfg::slot_off(k), &tmap_wide, lat_cx(k) > 0 ? x0 - Cfg::EXT : x0, c1, &full[stage]);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    // ---------------- consumer warps ----------------
    const int tx = tid % Cfg::TXT, ty = tid / Cfg::TXT;
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t smem_base = smem_u32(base);
    int cur_b = -1;
    Rates<T> rt(a.cav[0]);
    for (TileIter it(ts); it.valid(ts); it.next(ts)) {
        const int b = it.b;
        const int x = it.tx * Cfg::TX + tx * V;                     // first of this thread's V nodes
        const int yl = ts.row_begin + it.ty * TY + ty;
        const int y = a.y0 + yl;
        if (b != cur_b) { rt = Rates<T>(a.cav[b]); cur_b = b; }

        T f[V][9];
        mbar_wait(&full[stage], phase);
        {
            const uint32_t st = smem_base + (uint32_t)stage * Cfg::STAGE_BYTES;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                if (lat_cx(k) == 0) {
                    T tmp[V];
                    lds_vec(st + Cfg::slot_off(k) + (uint32_t)((ty * Cfg::TX + tx * V) * sizeof(T)), tmp);
#pragma unroll
                    for (int v = 0; v < V; ++v) f[v][k] = tmp[v];
                } else {
                    // wide box: columns [x0-EXT, x0+TX) for c_x = +1, [x0, x0+TX+EXT) for c_x = -1; node x needs x - c_x
                    T lo[V], hi[V];
                    const uint32_t row = st + Cfg::slot_off(k) +
                                         (uint32_t)((ty * Cfg::TXW + tx * V + (lat_cx(k) > 0 ? Cfg::EXT - V : 0)) * sizeof(T));
                    lds_vec(row, lo);
                    lds_vec(row + 16, hi);
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        if (lat_cx(k) > 0) f[v][k] = (v == 0) ? lo[V - 1] : hi[v - 1];
                        else f[v][k] = (v < V - 1) ? lo[v + 1] : hi[0];
                    }
                }
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[stage]);            // values are in registers: free the slot
        if (++stage == STAGES) { stage = 0; phase ^= 1; }

        const bool row_ok = (yl < ts.row_begin + ts.row_count) && (yl < a.nyl);
        if (!row_ok || x >= a.nx) continue;
        const bool lid = (y == 0), bot = (y == a.ny - 1);
        T* __restrict__ dstp = static_cast<T*>(a.dst) + (long long)b * a.cavity + (long long)(yl + 1) * a.pitch + x;
        T rho[V], ux[V], uy[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int xv = x + v;
            const bool left = (xv == 0), right = (xv == a.nx - 1);
            if (left || right || lid || bot) {
                if (xv < a.nx) {
                    const int slot = corner_slot(left, right, lid, bot);
                    T* carry = static_cast<T*>(a.carry) + b * 4;
                    const T stale = slot >= 0 ? carry[slot] : (T)0;
                    const T rl = lid ? static_cast<const T*>(a.rho_lid)[(long long)b * a.pitch + xv] : (T)1;
                    wall_rule<T>(f[v], left, right, lid, bot, rl, rt.uLB, stale);
                    if (slot >= 0) carry[slot] = corner_value<T>(f[v], slot);
                }
            }
            node_update<T, COLL, MACROS>(f[v], rt, left, right, lid, bot, rho[v], ux[v], uy[v]);
        }
        if (lid) {
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (x + v < a.nx) static_cast<T*>(a.rho_lid)[(long long)b * a.pitch + x + v] = rho[v];
        }
        if (x + V <= a.nx) {
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                T tmp[V];
#pragma unroll
                for (int v = 0; v < V; ++v) tmp[v] = f[v][k];
                vec_store_global<T, V>(dstp + k * a.plane, tmp);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k)
#pragma unroll
                for (int v = 0; v < V; ++v)
                    if (x + v < a.nx) dstp[k * a.plane + v] = f[v][k];
        }
        if (MACROS) {
            const long long m = (long long)b * a.mplane + (long long)yl * a.pitch + x;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                if (x + v < a.nx) {
                    static_cast<T*>(a.rho)[m + v] = rho[v];
                    static_cast<T*>(a.ux)[m + v] = ux[v];
                    static_cast<T*>(a.uy)[m + v] = uy[v];
                }
            }
        }
    }
}

}  // namespace lbm
