// C-ABI implementation (include/lbm_b200.h) and the "ldg" kernel family of the fused D2Q9 step.
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

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/lbm_b200.h"
#include "lbm_device.cuh"
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

// Cache-policy experiment hooks (compile with -DLBM_CACHE_HINTS=n; the shipped build uses 0 = default policy).
#ifndef LBM_CACHE_HINTS
#define LBM_CACHE_HINTS 0
#endif
template <typename T> __device__ __forceinline__ T ld_pop(const T* p) {
#if LBM_CACHE_HINTS == 2
    return __ldcs(p);
#elif LBM_CACHE_HINTS == 3
    return __ldg(p);
#elif LBM_CACHE_HINTS == 4
    return __ldcg(p);
#else
    return *p;
#endif
}
template <typename T> __device__ __forceinline__ void st_pop(T* p, T v) {
#if LBM_CACHE_HINTS == 1 || LBM_CACHE_HINTS == 2
    __stcs(p, v);
#elif LBM_CACHE_HINTS == 4
    __stcg(p, v);
#else
    *p = v;
#endif
}

// ------------------------------------------------------------------------------------------------------------
// "ldg" family: one thread per node, plain coalesced loads (x+-1 shifted reads are unaligned-but-contiguous per
// warp and are absorbed by L1/L2), aligned stores.  Template flags: dtype, collision, GATHER (false for the first
// launch after an upload: the buffer then holds pre-collision `fin`), MACROS (store rho,u), MODE.
// ------------------------------------------------------------------------------------------------------------
template <typename T, int COLL, bool GATHER, bool MACROS, int MODE, bool TURB = false>
__global__ void __launch_bounds__(256) lbm_step_ldg(const StepArgs a) {
    // programmatic dependent launch: let the next step's grid be scheduled, then wait for the previous grid
    // (no-ops for ordinary launches)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int lr = blockIdx.y * blockDim.y + threadIdx.y;          // launch row
    if (x >= a.nx || lr >= a.row_count) return;
    const int yl = a.row_begin + lr * a.row_stride;
    const int b = blockIdx.z;
    if (MODE == MODE_STEP && a.active && !a.active[b]) return;       // frozen (converged) cavity
    const int y = a.y0 + yl;
    const bool left = (x == 0), right = (x == a.nx - 1), lid = (y == 0), bot = (y == a.ny - 1);
    const T* __restrict__ src = static_cast<const T*>(a.src) + (long long)b * a.cavity;
    T* __restrict__ dst = static_cast<T*>(a.dst) + (long long)b * a.cavity;
    const long long P = a.plane;
    const long long rc = (long long)(yl + 1) * a.pitch + x;   // this node
    const long long ru = rc - a.pitch;                         // row y-1 (towards the lid)
    const long long rd = rc + a.pitch;                         // row y+1
    const Rates<T> r(a.cav[b]);

    T f[9];
    if (GATHER) {
        // pull: f_k arrives from (x - c_kx, y + c_ky)
        f[0] = ld_pop(src + rc);
        f[1] = left ? (T)0 : ld_pop(src + 1 * P + rc - 1);
        f[2] = bot ? (T)0 : ld_pop(src + 2 * P + rd);
        f[3] = right ? (T)0 : ld_pop(src + 3 * P + rc + 1);
        f[4] = lid ? (T)0 : ld_pop(src + 4 * P + ru);
        f[5] = (left || bot) ? (T)0 : ld_pop(src + 5 * P + rd - 1);
        f[6] = (right || bot) ? (T)0 : ld_pop(src + 6 * P + rd + 1);
        f[7] = (right || lid) ? (T)0 : ld_pop(src + 7 * P + ru + 1);
        f[8] = (left || lid) ? (T)0 : ld_pop(src + 8 * P + ru - 1);
        if (left || right || lid || bot) {
            const int slot = corner_slot(left, right, lid, bot);
            T* carry = static_cast<T*>(a.carry) + b * 4;
            const T stale = slot >= 0 ? carry[slot] : (T)0;
            const T rl = lid ? static_cast<const T*>(a.rho_lid)[(long long)b * a.pitch + x] : (T)1;
            wall_rule<T>(f, left, right, lid, bot, rl, r.uLB, stale);
            if (MODE == MODE_STEP && slot >= 0) carry[slot] = corner_value<T>(f, slot);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) f[k] = src[k * P + rc];
    }

    if (MODE == MODE_FINALIZE) {
#pragma unroll
        for (int k = 0; k < 9; ++k) dst[k * P + rc] = f[k];
        return;
    }

    T rho, ux, uy;
    if (MODE == MODE_MACROS) {
        // current-state moments with the reference's overrides, no collision
        T jx, jy;
        moments_ref<T>(f, rho, jx, jy);
        ux = jx / rho; uy = jy / rho;
        if (left || right || bot) { ux = (T)0; uy = (T)0; }
        if (lid) { rho = rho_lid_formula<T>(f); ux = r.uLB; uy = (T)0; }
    } else {
        if (TURB) {
            const long long m = (long long)b * a.mplane + (long long)yl * a.pitch + x;
            T* pi = static_cast<T*>(a.pi_eq) + m;
            T* rp = static_cast<T*>(a.rho_prev) + m;
            const T om = smagorinsky_omega<T>(f, *pi, *rp, r.tau0);
            T pi_new;
            node_update<T, COLL, MACROS, true>(f, r, left, right, lid, bot, rho, ux, uy, om, &pi_new);
            *pi = pi_new;
            *rp = rho;
        } else {
            node_update<T, COLL, MACROS>(f, r, left, right, lid, bot, rho, ux, uy);
        }
        if (lid) static_cast<T*>(a.rho_lid)[(long long)b * a.pitch + x] = rho;
#pragma unroll
        for (int k = 0; k < 9; ++k) st_pop(dst + k * P + rc, f[k]);
    }
    if (MACROS || MODE == MODE_MACROS) {
        const long long m = (long long)b * a.mplane + (long long)yl * a.pitch + x;
        static_cast<T*>(a.rho)[m] = rho;
        static_cast<T*>(a.ux)[m] = ux;
        static_cast<T*>(a.uy)[m] = uy;
    }
}

// ------------------------------------------------------------------------------------------------------------
// "ldg" family, vector form (hot path only: MODE_STEP with gather): one thread updates V consecutive nodes of a
// row.  Every population is fetched with ONE aligned V-wide load per thread (64/128-bit, coalesced along x); the
// x-1 / x+1 element that the pull step needs from the neighbouring thread's vector comes by warp shuffle, and only
// the first / last lane of a warp issues one extra scalar load.  Stores are aligned V-wide.  Per node this halves
// (V = 2) or quarters (V = 4) the load/store and address instructions of the scalar kernel.
// ------------------------------------------------------------------------------------------------------------
template <typename T, int V> struct GVec;
template <> struct GVec<float, 2> { using type = float2; };
template <> struct GVec<float, 4> { using type = float4; };
template <> struct GVec<double, 2> { using type = double2; };

template <typename T, int V>
__device__ __forceinline__ void gload(const T* p, T out[V]) {
    using VT = typename GVec<T, V>::type;
    const VT v = *reinterpret_cast<const VT*>(p);
    const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
    for (int i = 0; i < V; ++i) out[i] = e[i];
}
template <typename T, int V>
__device__ __forceinline__ void gstore(T* p, const T in[V]) {
    using VT = typename GVec<T, V>::type;
    VT v;
    T* e = reinterpret_cast<T*>(&v);
#pragma unroll
    for (int i = 0; i < V; ++i) e[i] = in[i];
    *reinterpret_cast<VT*>(p) = v;
}

template <typename T, int COLL, bool MACROS, int V, bool TURB = false>
__global__ void __launch_bounds__(256) lbm_step_vec(const StepArgs a) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int lane = threadIdx.x & 31;                                  // blockDim.x is a multiple of 32: a warp is one row
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * V;          // first node of this thread
    const int lr = blockIdx.y * blockDim.y + threadIdx.y;               // launch row (warp-uniform)
    if (lr >= a.row_count) return;
    const int yl = a.row_begin + lr * a.row_stride;
    const int b = blockIdx.z;
    if (a.active && !a.active[b]) return;                               // frozen (converged) cavity
    const int y = a.y0 + yl;
    const bool lid = (y == 0), bot = (y == a.ny - 1);
    const bool active = x < a.nx;                                       // whole warps may be partially outside
    const T* __restrict__ src = static_cast<const T*>(a.src) + (long long)b * a.cavity;
    T* __restrict__ dst = static_cast<T*>(a.dst) + (long long)b * a.cavity;
    const long long P = a.plane;
    // clamp the address of inactive threads to a valid one (they still take part in the shuffles)
    const int xc = active ? x : 0;
    const long long rc = (long long)(yl + 1) * a.pitch + xc;
    const long long ru = rc - a.pitch, rd = rc + a.pitch;
    const Rates<T> r(a.cav[b]);
    const unsigned full = 0xffffffffu;

    T f[V][9];
    // aligned vectors of every population at this thread's columns, from the row the population comes from
    T v0[V], v1[V], v2[V], v3[V], v4[V], v5[V], v6[V], v7[V], v8[V];
    gload<T, V>(src + rc, v0);
    gload<T, V>(src + 1 * P + rc, v1);
    gload<T, V>(src + 2 * P + rd, v2);
    gload<T, V>(src + 3 * P + rc, v3);
    gload<T, V>(src + 4 * P + ru, v4);
    gload<T, V>(src + 5 * P + rd, v5);
    gload<T, V>(src + 6 * P + rd, v6);
    gload<T, V>(src + 7 * P + ru, v7);
    gload<T, V>(src + 8 * P + ru, v8);
    // element x-1 for c_x = +1 (k = 1,5,8): previous lane's last element; lane 0 loads it (0 at the left wall)
    T l1 = __shfl_up_sync(full, v1[V - 1], 1), l5 = __shfl_up_sync(full, v5[V - 1], 1), l8 = __shfl_up_sync(full, v8[V - 1], 1);
    if (lane == 0) {
        const bool ok = active && x > 0;
        l1 = ok ? src[1 * P + rc - 1] : (T)0;
        l5 = ok ? src[5 * P + rd - 1] : (T)0;
        l8 = ok ? src[8 * P + ru - 1] : (T)0;
    }
    // element x+V for c_x = -1 (k = 3,6,7): next lane's first element; lane 31 loads it (0 beyond the right wall)
    T h3 = __shfl_down_sync(full, v3[0], 1), h6 = __shfl_down_sync(full, v6[0], 1), h7 = __shfl_down_sync(full, v7[0], 1);
    if (lane == 31) {
        const bool ok = active && (x + V) < a.nx;
        h3 = ok ? src[3 * P + rc + V] : (T)0;
        h6 = ok ? src[6 * P + rd + V] : (T)0;
        h7 = ok ? src[7 * P + ru + V] : (T)0;
    }
    if (!active) return;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        f[v][0] = v0[v];
        f[v][2] = v2[v];
        f[v][4] = v4[v];
        f[v][1] = v == 0 ? l1 : v1[v - 1];
        f[v][5] = v == 0 ? l5 : v5[v - 1];
        f[v][8] = v == 0 ? l8 : v8[v - 1];
        f[v][3] = v == V - 1 ? h3 : v3[v + 1];
        f[v][6] = v == V - 1 ? h6 : v6[v + 1];
        f[v][7] = v == V - 1 ? h7 : v7[v + 1];
    }
    T rho[V], ux[V], uy[V];
    T pi_old[V], rp_old[V], pi_new[V];
    if (TURB) {   // previous-step sum cx cy feq and rho of these nodes (pitch padding keeps the vector access in bounds)
        const long long m = (long long)b * a.mplane + (long long)yl * a.pitch + x;
        gload<T, V>(static_cast<const T*>(a.pi_eq) + m, pi_old);
        gload<T, V>(static_cast<const T*>(a.rho_prev) + m, rp_old);
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const int xv = x + v;
        const bool left = (xv == 0), right = (xv == a.nx - 1);
        if ((left || right || lid || bot) && xv < a.nx) {
            const int slot = corner_slot(left, right, lid, bot);
            T* carry = static_cast<T*>(a.carry) + b * 4;
            const T stale = slot >= 0 ? carry[slot] : (T)0;
            const T rl = lid ? static_cast<const T*>(a.rho_lid)[(long long)b * a.pitch + xv] : (T)1;
            wall_rule<T>(f[v], left, right, lid, bot, rl, r.uLB, stale);
            if (slot >= 0) carry[slot] = corner_value<T>(f[v], slot);
        }
        if (TURB) pi_new[v] = (T)0;
        if (TURB) {
            const T om = smagorinsky_omega<T>(f[v], pi_old[v], rp_old[v], r.tau0);
            node_update<T, COLL, MACROS, true>(f[v], r, left, right, lid, bot, rho[v], ux[v], uy[v], om, &pi_new[v]);
        } else {
            node_update<T, COLL, MACROS>(f[v], r, left, right, lid, bot, rho[v], ux[v], uy[v]);
        }
    }
    if (TURB) {
        const long long m = (long long)b * a.mplane + (long long)yl * a.pitch + x;
        if (x + V <= a.nx) {
            gstore<T, V>(static_cast<T*>(a.pi_eq) + m, pi_new);
            gstore<T, V>(static_cast<T*>(a.rho_prev) + m, rho);
        } else {
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (x + v < a.nx) { static_cast<T*>(a.pi_eq)[m + v] = pi_new[v]; static_cast<T*>(a.rho_prev)[m + v] = rho[v]; }
        }
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
            gstore<T, V>(dst + k * P + rc, tmp);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 9; ++k)
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (x + v < a.nx) dst[k * P + rc + v] = f[v][k];
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

// ------------------------------------------------------------------------------------------------------------
// Temporal blocking: TWO lattice steps per pass over memory.  A CTA owns a TX x TY tile.  Sub-step 1 advances the
// tile grown by one node on every side (pulling, like the scalar kernel, from the global post-collision buffer,
// wall rule on read) and keeps the resulting post-collision populations in shared memory; after one barrier,
// sub-step 2 advances the tile itself by pulling from shared memory and stores to the other global buffer.  Per
// node and TWO steps the kernel moves 9 loads (+ the tile halo, served by L2) and 9 stores: half the DRAM traffic
// per update of the one-step kernels, at the price of recomputing the one-node ring ((TX+2)(TY+2)/(TX TY) - 1 of
// sub-step 1) -- the step is HBM-bound with the fp64 pipe 24 % busy, so the arithmetic is available.
// The per-node arithmetic is the shared node_update()/wall_rule(), so results are bit-identical to two one-step
// launches.  The lid density and corner carries of the intermediate state live in shared memory; those of the
// final state go to the other half of the double-buffered side arrays (a neighbouring CTA may still need the old
// ones for its halo ring).  On a y-strip the ring of the first / last tile row lies in the ghost row, whose own
// sub-step 1 needs one more row from the neighbour: the second ghost rows (`ghost2`, three populations each).
// ------------------------------------------------------------------------------------------------------------
template <typename T, int TX_, int TY_> struct Fused2Cfg {
    static constexpr int TX = TX_, TY = TY_;
    static constexpr int RX = TX + 2, RY = TY + 2;                 // tile grown by one node
    static constexpr int PLANE = RX * RY;
    static constexpr size_t SMEM = (size_t)(9 * PLANE + RX + 4) * sizeof(T);
};

template <typename T, int COLL, bool MACROS, int TX_, int TY_, int MINB, bool GHOST2>
__global__ void __launch_bounds__(256, MINB) lbm_step_fused2(const StepArgs a) {
    using Cfg = Fused2Cfg<T, TX_, TY_>;
    extern __shared__ __align__(16) unsigned char fused_smem[];
    T* h1 = reinterpret_cast<T*>(fused_smem);                      // [9][RY][RX] post-collision after sub-step 1
    T* rl1 = h1 + 9 * Cfg::PLANE;                                  // [RX] lid density after sub-step 1
    T* c1 = rl1 + Cfg::RX;                                         // [4]  corner carries after sub-step 1
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * Cfg::TX;
    const int yl0 = (a.row_begin + blockIdx.y) * Cfg::TY;          // first LOCAL row of the tile (row_begin in tile rows)
    const T* __restrict__ src = static_cast<const T*>(a.src) + (long long)b * a.cavity;
    T* __restrict__ dst = static_cast<T*>(a.dst) + (long long)b * a.cavity;
    const long long P = a.plane;
    const Rates<T> r(a.cav[b]);
    const T* carry_in = static_cast<const T*>(a.carry) + b * 4;
    const T* rl_in = static_cast<const T*>(a.rho_lid) + (long long)b * a.pitch;
    // second ghost rows of a y-strip (neighbour's rows y0-2: populations 4,7,8 and y0+nyl+1: populations 2,5,6)
    const T* g2top = static_cast<const T*>(a.ghost2) + (long long)b * 6 * a.pitch;
    const T* g2bot = g2top + 3 * a.pitch;

    // ---- sub-step 1 on the grown tile: global (state t) -> shared (state t+1) ----
    for (int i = threadIdx.x; i < Cfg::PLANE; i += blockDim.x) {
        const int ly = i / Cfg::RX, lx = i - ly * Cfg::RX;
        const int x = x0 - 1 + lx, yl = yl0 - 1 + ly;               // yl in [-1, nyl]: the ring may sit in a ghost row
        const int y = a.y0 + yl;
        if (x < 0 || x >= a.nx || y < 0 || y >= a.ny || yl > a.nyl) continue;
        const bool left = (x == 0), right = (x == a.nx - 1), lid = (y == 0), bot = (y == a.ny - 1);
        const long long rc = (long long)(yl + 1) * a.pitch + x, ru = rc - a.pitch, rd = rc + a.pitch;
        // row above / below comes from the second ghost row (edge bands of a y-strip only: GHOST2)
        const bool up2 = GHOST2 && (yl == -1), dn2 = GHOST2 && (yl == a.nyl);
        T f[9];
        f[0] = src[rc];
        f[1] = left ? (T)0 : src[1 * P + rc - 1];
        f[3] = right ? (T)0 : src[3 * P + rc + 1];
        f[2] = bot ? (T)0 : (dn2 ? g2bot[x] : src[2 * P + rd]);
        f[5] = (left || bot) ? (T)0 : (dn2 ? g2bot[a.pitch + x - 1] : src[5 * P + rd - 1]);
        f[6] = (right || bot) ? (T)0 : (dn2 ? g2bot[2 * a.pitch + x + 1] : src[6 * P + rd + 1]);
        f[4] = lid ? (T)0 : (up2 ? g2top[x] : src[4 * P + ru]);
        f[7] = (right || lid) ? (T)0 : (up2 ? g2top[a.pitch + x + 1] : src[7 * P + ru + 1]);
        f[8] = (left || lid) ? (T)0 : (up2 ? g2top[2 * a.pitch + x - 1] : src[8 * P + ru - 1]);
        if (left || right || lid || bot) {
            const int slot = corner_slot(left, right, lid, bot);
            const T stale = slot >= 0 ? carry_in[slot] : (T)0;
            const T rl = lid ? rl_in[x] : (T)1;
            wall_rule<T>(f, left, right, lid, bot, rl, r.uLB, stale);
            if (slot >= 0) c1[slot] = corner_value<T>(f, slot);
        }
        T rho, ux, uy;
        node_update<T, COLL, false>(f, r, left, right, lid, bot, rho, ux, uy);
        if (lid) rl1[lx] = rho;
#pragma unroll
        for (int k = 0; k < 9; ++k) h1[k * Cfg::PLANE + i] = f[k];
    }
    __syncthreads();

    // ---- sub-step 2 on the tile: shared (state t+1) -> global (state t+2) ----
    for (int i = threadIdx.x; i < Cfg::TX * Cfg::TY; i += blockDim.x) {
        const int ty = i / Cfg::TX, tx = i - ty * Cfg::TX;
        const int x = x0 + tx, yl = yl0 + ty;
        const int y = a.y0 + yl;
        if (x >= a.nx || yl >= a.nyl) continue;
        const bool left = (x == 0), right = (x == a.nx - 1), lid = (y == 0), bot = (y == a.ny - 1);
        const int c = (ty + 1) * Cfg::RX + (tx + 1);               // this node inside the grown tile
        const int u = c - Cfg::RX, d = c + Cfg::RX;                // row y-1 / y+1
        T f[9];
        f[0] = h1[c];
        f[1] = left ? (T)0 : h1[1 * Cfg::PLANE + c - 1];
        f[2] = bot ? (T)0 : h1[2 * Cfg::PLANE + d];
        f[3] = right ? (T)0 : h1[3 * Cfg::PLANE + c + 1];
        f[4] = lid ? (T)0 : h1[4 * Cfg::PLANE + u];
        f[5] = (left || bot) ? (T)0 : h1[5 * Cfg::PLANE + d - 1];
        f[6] = (right || bot) ? (T)0 : h1[6 * Cfg::PLANE + d + 1];
        f[7] = (right || lid) ? (T)0 : h1[7 * Cfg::PLANE + u + 1];
        f[8] = (left || lid) ? (T)0 : h1[8 * Cfg::PLANE + u - 1];
        if (left || right || lid || bot) {
            const int slot = corner_slot(left, right, lid, bot);
            const T stale = slot >= 0 ? c1[slot] : (T)0;
            const T rl = lid ? rl1[tx + 1] : (T)1;
            wall_rule<T>(f, left, right, lid, bot, rl, r.uLB, stale);
            if (slot >= 0) static_cast<T*>(a.carry_out)[b * 4 + slot] = corner_value<T>(f, slot);
        }
        T rho, ux, uy;
        node_update<T, COLL, MACROS>(f, r, left, right, lid, bot, rho, ux, uy);
        if (lid) static_cast<T*>(a.rho_lid_out)[(long long)b * a.pitch + x] = rho;
        const long long rc = (long long)(yl + 1) * a.pitch + x;
#pragma unroll
        for (int k = 0; k < 9; ++k) dst[k * P + rc] = f[k];
        if (MACROS) {
            const long long m = (long long)b * a.mplane + (long long)yl * a.pitch + x;
            static_cast<T*>(a.rho)[m] = rho;
            static_cast<T*>(a.ux)[m] = ux;
            static_cast<T*>(a.uy)[m] = uy;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Semantics "A" (MRT.py:286-453), compatibility mode: two plain passes per step on pre-collision `fin`.
//   pass 1  moments + overrides (:292-342), SRT collision (:396)            fin -> fpost, rho, u
//   pass 2  slice streaming with xsize_max / ysize_max as EXCLUSIVE bounds (:404-414): slots outside the slices keep
//           their old value; then the four wall assignments in the script's order (:450-453), left wall "= feq"
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void lbm_A_collide(const StepArgs a) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= a.nx) return;
    const int y = blockIdx.y, b = blockIdx.z;
    const T* __restrict__ fin = static_cast<const T*>(a.src) + (long long)b * a.cavity;
    T* __restrict__ fpost = static_cast<T*>(a.dst) + (long long)b * a.cavity;
    const long long rc = (long long)(y + 1) * a.pitch + x;
    const Rates<T> r(a.cav[b]);
    T f[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) f[k] = fin[k * a.plane + rc];
    T rho, jx, jy;
    moments_ref<T>(f, rho, jx, jy);
    T ux = jx / rho, uy = jy / rho;
    if (y == 0) rho = rho_lid_formula<T>(f);                              // MRT.py:337
    if (x == 0 || x == a.nx - 1 || y == a.ny - 1) { ux = (T)0; uy = (T)0; }   // :341
    if (y == 0) { ux = r.uLB; uy = (T)0; }                                // :342
    collide_srt<T>(f, rho, ux, uy, r.omega);                              // :396
#pragma unroll
    for (int k = 0; k < 9; ++k) fpost[k * a.plane + rc] = f[k];
    const long long m = (long long)b * a.mplane + (long long)y * a.pitch + x;
    static_cast<T*>(a.rho)[m] = rho;
    static_cast<T*>(a.ux)[m] = ux;
    static_cast<T*>(a.uy)[m] = uy;
}

template <typename T>
__global__ void lbm_A_stream_bc(const StepArgs a) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= a.nx) return;
    const int y = blockIdx.y, b = blockIdx.z;
    const int nx = a.nx, ny = a.ny;
    const T* __restrict__ fpost = static_cast<const T*>(a.src) + (long long)b * a.cavity;
    T* __restrict__ fin = static_cast<T*>(a.dst) + (long long)b * a.cavity;
    const long long P = a.plane;
    const long long rc = (long long)(y + 1) * a.pitch + x, ru = rc - a.pitch, rd = rc + a.pitch;
    // target ranges of the slice assignments MRT.py:404-414
    const bool xr = x >= 1 && x <= nx - 2;       // c_x = +1 : fin[k, 1:xm]     <- fpost[k, 0:xm-1]
    const bool xl = x <= nx - 3;                 // c_x = -1 : fin[k, 0:xm-1]   <- fpost[k, 1:xm]
    const bool yu = y <= ny - 3;                 // c_y = +1 : fin[k, :, 0:ym-1] <- fpost[k, :, 1:ym]
    const bool yd = y >= 1 && y <= ny - 2;       // c_y = -1 : fin[k, :, 1:ym]   <- fpost[k, :, 0:ym-1]
    T f[9];
    f[0] = fpost[rc];
    f[1] = xr ? fpost[1 * P + rc - 1] : fin[1 * P + rc];
    f[2] = yu ? fpost[2 * P + rd] : fin[2 * P + rc];
    f[3] = xl ? fpost[3 * P + rc + 1] : fin[3 * P + rc];
    f[4] = yd ? fpost[4 * P + ru] : fin[4 * P + rc];
    f[5] = (xr && yu) ? fpost[5 * P + rd - 1] : fin[5 * P + rc];
    f[6] = (xl && yu) ? fpost[6 * P + rd + 1] : fin[6 * P + rc];
    f[7] = (xl && yd) ? fpost[7 * P + ru + 1] : fin[7 * P + rc];
    f[8] = (xr && yd) ? fpost[8 * P + ru - 1] : fin[8 * P + rc];
    if (x == 0 || x == nx - 1 || y == 0 || y == ny - 1) {
        const long long m = (long long)b * a.mplane + (long long)y * a.pitch + x;
        T fe[9];
        feq_all<T>(static_cast<const T*>(a.rho)[m], static_cast<const T*>(a.ux)[m], static_cast<const T*>(a.uy)[m], fe);
        if (x == 0) { f[1] = fe[1]; f[5] = fe[5]; f[8] = fe[8]; }                         // :450
        if (x == nx - 1) {                                                                // :451  (3,6,7) <- (1,5,8)
            f[3] = -fe[1] + (fe[3] + f[1]);
            f[6] = -fe[5] + (fe[6] + f[5]);
            f[7] = -fe[8] + (fe[7] + f[8]);
        }
        if (y == ny - 1) {                                                                // :452  (2,5,6) <- (4,7,8)
            f[2] = -fe[4] + (fe[2] + f[4]);
            f[5] = -fe[7] + (fe[5] + f[7]);
            f[6] = -fe[8] + (fe[6] + f[8]);
        }
        if (y == 0) {                                                                     // :453  (4,7,8) <- (2,5,6)
            f[4] = -fe[2] + (fe[4] + f[2]);
            f[7] = -fe[5] + (fe[7] + f[5]);
            f[8] = -fe[6] + (fe[8] + f[6]);
        }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) fin[k * P + rc] = f[k];
}

// Equilibrium start (MRT_GPU.py:259-267): rho = 1, u = (uLB, 0) on row y == 0, evaluated in fp64 then cast
// (the reference builds it in fp64 NumPy and casts to fp32, :298).  Also seeds the corner carries and rho = 1, u = 0.
template <typename T>
__global__ void lbm_init_eq(StepArgs a) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= a.nx) return;
    const int yl = blockIdx.y, b = blockIdx.z;
    const int y = a.y0 + yl;
    const double ux = (y == 0) ? a.cav[b].uLB : 0.0;
    double fe[9];
    feq_all<double>(1.0, ux, 0.0, fe);
    T* dst = static_cast<T*>(a.dst) + (long long)b * a.cavity;
    const long long rc = (long long)(yl + 1) * a.pitch + x;
#pragma unroll
    for (int k = 0; k < 9; ++k) dst[k * a.plane + rc] = (T)fe[k];
    const long long m = (long long)b * a.mplane + (long long)yl * a.pitch + x;
    static_cast<T*>(a.rho)[m] = (T)1;
    static_cast<T*>(a.ux)[m] = (T)0;
    static_cast<T*>(a.uy)[m] = (T)0;
    const bool left = (x == 0), right = (x == a.nx - 1), lid = (y == 0), bot = (y == a.ny - 1);
    const int slot = corner_slot(left, right, lid, bot);
    if (slot >= 0) static_cast<T*>(a.carry)[b * 4 + slot] = (T)corner_value<double>(fe, slot);
    if (a.pi_eq) {   // feq_g := fin, rho_g := 1 at start (MRT_GPU.py:325-326)
        static_cast<T*>(a.pi_eq)[m] = (T)fe[5] - (T)fe[6] + (T)fe[7] - (T)fe[8];
        static_cast<T*>(a.rho_prev)[m] = (T)1;
    }
}

// After an upload with turb = 1: feq_g := uploaded fin, rho_g := 1 (MRT_GPU.py:325-326).
template <typename T>
__global__ void lbm_seed_turb(StepArgs a) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= a.nx) return;
    const int yl = blockIdx.y, b = blockIdx.z;
    const T* src = static_cast<const T*>(a.src) + (long long)b * a.cavity;
    const long long rc = (long long)(yl + 1) * a.pitch + x;
    const long long m = (long long)b * a.mplane + (long long)yl * a.pitch + x;
    static_cast<T*>(a.pi_eq)[m] = src[5 * a.plane + rc] - src[6 * a.plane + rc] + src[7 * a.plane + rc] - src[8 * a.plane + rc];
    static_cast<T*>(a.rho_prev)[m] = (T)1;
}

// After an upload: seed the corner carries from the uploaded `fin` (stale ftemp slot == fin slot, MRT_GPU.py:324).
template <typename T>
__global__ void lbm_seed_carry(StepArgs a) {
    const int b = blockIdx.x, slot = threadIdx.x;
    if (slot >= 4) return;
    const bool lid = slot < 2, left = (slot == 0 || slot == 2);
    const int y = lid ? 0 : a.ny - 1;
    const int yl = y - a.y0;
    if (yl < 0 || yl >= a.nyl) return;
    const int x = left ? 0 : a.nx - 1;
    const T* src = static_cast<const T*>(a.src) + (long long)b * a.cavity;
    static_cast<T*>(a.carry)[b * 4 + slot] = src[corner_pop(slot) * a.plane + (long long)(yl + 1) * a.pitch + x];
}

template <typename T>
__global__ void lbm_fill(T* p, long long n, T v) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

// Layout change between the reference's host arrays [plane][nx][ny_local] (y fastest) and device planes
// [plane][row][pitch] (x fastest) -- the transposes of MRT_GPU.py:283-289 / 758-760, as a 32x32 shared-memory tile.
// dev_plane(p) = dev + p_off(p); TO_DEVICE: host layout -> device layout.
template <typename T, bool TO_DEVICE>
__global__ void lbm_transpose(T* __restrict__ dev, T* __restrict__ lin, int nx, int nyl, int pitch,
                              long long dev_plane_stride, long long dev_row0) {
    __shared__ T tile[32][33];
    const int p = blockIdx.z;
    T* d = dev + (long long)p * dev_plane_stride + dev_row0;
    T* l = lin + (long long)p * nx * nyl;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    if (TO_DEVICE) {
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {          // read lin[x][y], y fastest
            const int x = x0 + i, y = y0 + threadIdx.x;
            if (x < nx && y < nyl) tile[i][threadIdx.x] = l[(long long)x * nyl + y];
        }
        __syncthreads();
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {          // write dev[y][x], x fastest
            const int y = y0 + i, x = x0 + threadIdx.x;
            if (x < nx && y < nyl) d[(long long)y * pitch + x] = tile[threadIdx.x][i];
        }
    } else {
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            const int y = y0 + i, x = x0 + threadIdx.x;
            if (x < nx && y < nyl) tile[i][threadIdx.x] = d[(long long)y * pitch + x];
        }
        __syncthreads();
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            const int x = x0 + i, y = y0 + threadIdx.x;
            if (x < nx && y < nyl) l[(long long)x * nyl + y] = tile[threadIdx.x][i];
        }
    }
}

// functions.equ (functions.pyx:229-267): feq[k][i] from rho[i], ux[i], uy[i]
template <typename T>
__global__ void lbm_equ_kernel(const T* __restrict__ rho, const T* __restrict__ ux, const T* __restrict__ uy,
                               T* __restrict__ feq, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        T fe[9];
        feq_all<T>(rho[i], ux[i], uy[i], fe);
#pragma unroll
        for (int k = 0; k < 9; ++k) feq[k * n + i] = fe[k];
    }
}

// np.mean(u) of MRT_GPU_datagen.py:729 per cavity: sum of both stored velocity components over the valid nodes,
// accumulated in fp64 (block tree + one atomicAdd per block).
template <typename T>
__global__ void lbm_sum_u(const T* __restrict__ ux, const T* __restrict__ uy, double* __restrict__ out, int nx, int nyl,
                          int pitch, long long mplane) {
    const int b = blockIdx.y;
    const long long n = (long long)nyl * pitch;
    const T* px = ux + (long long)b * mplane;
    const T* py = uy + (long long)b * mplane;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % pitch);
        if (x < nx) acc += (double)px[i] + (double)py[i];
    }
    __shared__ double sh[32];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (threadIdx.x == 0) atomicAdd(&out[b], acc);
    }
}

// Diagnostics the reference scripts compute on the host after downloading the full fields (MRT_GPU.py:764-776,
// 793-800): centre-lines ux(x = nx/2, :) and uy(:, y = ny/2), and the vortex-centre search = argmin of |u|^2 with a
// border of BCoffset = nx/40 nodes (and optionally a box around the first centre) masked out.
template <typename T>
__global__ void lbm_centerlines(const T* __restrict__ ux, const T* __restrict__ uy, T* __restrict__ ux_col,
                                T* __restrict__ uy_row, int nx, int nyl, int pitch, int xc, int yc_local) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nyl) ux_col[i] = ux[(long long)i * pitch + xc];
    if (yc_local >= 0 && i < nx) uy_row[i] = uy[(long long)yc_local * pitch + i];
}

struct ArgMin { double val; long long idx; };

template <typename T>
__global__ void lbm_argmin_usq(const T* __restrict__ ux, const T* __restrict__ uy, ArgMin* __restrict__ out, int nx, int ny,
                               int pitch, int bc, int bx0, int bx1, int by0, int by1) {
    // flat index of the reference's [x][y] array = x * ny + y; first occurrence wins on ties (np.nanargmin)
    double best = 1e300;
    long long bidx = -1;
    const long long n = (long long)nx * ny;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / nx), x = (int)(i - (long long)y * nx);      // coalesced along x
        if (x < bc || y < bc || x >= nx - 1 - bc || y >= ny - 1 - bc) continue;
        if (x >= bx0 && x < bx1 && y >= by0 && y < by1) continue;
        const double a = (double)ux[(long long)y * pitch + x], b = (double)uy[(long long)y * pitch + x];
        const double v = a * a + b * b;
        const long long flat = (long long)x * ny + y;
        if (v < best || (v == best && flat < bidx)) { best = v; bidx = flat; }
    }
    __shared__ double sv[256];
    __shared__ long long si[256];
    sv[threadIdx.x] = best; si[threadIdx.x] = bidx;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double v = sv[threadIdx.x + o]; const long long j = si[threadIdx.x + o];
            if (j >= 0 && (si[threadIdx.x] < 0 || v < sv[threadIdx.x] || (v == sv[threadIdx.x] && j < si[threadIdx.x]))) {
                sv[threadIdx.x] = v; si[threadIdx.x] = j;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[blockIdx.x].val = sv[0]; out[blockIdx.x].idx = si[0]; }
}

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
    CavityParams* cav = nullptr;
    std::vector<CavityParams> cav_host;
    bool cav_dirty = true;
    void* staging = nullptr;   // host-layout staging for uploads/downloads
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
    // nodes per thread of the ldg family (1 = scalar kernel).  Measured at 4096^2 on B200 (tools/tma_sweep.py):
    // fp64 scalar 47 067 vs vec2 45 905 MLUPS; fp32 scalar 87 132, vec2 88 321, vec4 91 055 MLUPS.
    int vec_f64 = 1, vec_f32 = 4;
    // CUDA graphs of the steady step loop: graph[p] = 2*GRAPH_PAIRS launches starting with buffer p as source
    cudaGraphExec_t graph[4] = {nullptr, nullptr, nullptr, nullptr};   // index = cur * 2 + side
    cudaStream_t capture_stream = nullptr;
    int use_graph = 1;
    int use_pdl = 1;
    int use_fused2 = 1;        // temporal blocking (two steps per launch) for whole cavities
    int side = 0;              // which half of the double-buffered rho_lid / carry arrays is current
    int fused2_tile = -1;      // tile-shape variant of the fused kernel (-1 = per-dtype default)
};

// A fresh state (init / upload) un-freezes every cavity; graphs captured with the old flag pointer are dropped.
static void reset_active(lbm_solver* s) {
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
    a.pi_eq = s->pi_eq; a.rho_prev = s->rho_prev;
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

template <typename T, int COLL, int TY, int STAGES, int MINB>
static cudaError_t launch_tma_cfg(lbm_solver* s, const CUtensorMap* tm, const StepArgs& a, const TileSched& ts, cudaStream_t st) {
    constexpr int V = TmaVec<T>::V;
    using Cfg = TmaCfg<T, V, TY, STAGES>;
    auto kern = lbm_step_tma<T, COLL, false, V, TY, STAGES, MINB>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
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

// Below this many nodes the 64x8 / 32x16 tiles no longer fill the 148 SMs several times over and the one-step kernel
// (one wave of dependent L2 loads) is as fast or faster: 384^2 3.6 vs 3.9 us/step, 640^2 equal, 1024^2 46 -> 64 GLUPS.
#define LBM_FUSED2_MIN_NODES 600000

// Temporal blocking applies to this handle at all (whole cavity or y-strip of at least two rows)?
static bool fused2_capable(const lbm_solver* s) {
    // fp32 gains less from it (it is ALU- rather than HBM-bound) and loses on narrow cavities: 8 x 32 cavities of 384^2
    // ran at 666 537 MLUPS with it against 690 216 without
    if (s->esz == 4 && s->cfg.nx < 1024) return false;
    return s->use_fused2 && (long long)s->cfg.nx * s->cfg.ny * s->cfg.batch >= LBM_FUSED2_MIN_NODES && s->nyl >= 2 &&
           !s->cfg.turb && s->engine == LBM_ENGINE_LDG && !s->active && s->cfg.semantics == LBM_SEMANTICS_C;
}
// ... and to lbm_step, which owns whole cavities only
static bool fused2_usable(const lbm_solver* s) { return fused2_capable(s) && s->nyl == s->cfg.ny; }

// ---- fused two-step launch ------------------------------------------------------------------------------------
template <typename T, int COLL, int TX, int TY, int MINB, bool GHOST2>
static cudaError_t launch_fused2_cfg(lbm_solver* s, const StepArgs& a, bool macros, cudaStream_t st) {
    using Cfg = Fused2Cfg<T, TX, TY>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(lbm_step_fused2<T, COLL, false, TX, TY, MINB, GHOST2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(lbm_step_fused2<T, COLL, true, TX, TY, MINB, GHOST2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
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
    // defaults from tools/fused2_sweep.py at 4096^2: fp64 64x8 tiles at 4 CTAs/SM (73 610 MLUPS), fp32 32x16 at 4 CTAs/SM (105 000-113 000)
    const int variant = s->fused2_tile >= 0 ? s->fused2_tile : (sizeof(T) == 8 ? 3 : 4);
    switch (variant) {
        case 1: return launch_fused2_tile<T, COLL, 64, 8, 3>(s, a, macros, st);
        case 2: return launch_fused2_tile<T, COLL, 32, 16, 3>(s, a, macros, st);
        case 3: return launch_fused2_tile<T, COLL, 64, 8, 4>(s, a, macros, st);
        case 4: return launch_fused2_tile<T, COLL, 32, 16, 4>(s, a, macros, st);
        default: return launch_fused2_tile<T, COLL, 64, 16, 2>(s, a, macros, st);
    }
}

// Tile height of the fused kernel in use (needed to cut a strip into edge / interior bands of whole tile rows).
static int fused2_tile_height(const lbm_solver* s) {
    const int variant = s->fused2_tile >= 0 ? s->fused2_tile : (s->esz == 8 ? 3 : 4);
    switch (variant) {
        case 1: case 3: return 8;
        default: return 16;
    }
}

// Fused launch over local rows [row_begin, row_begin + row_count) (row_begin a multiple of the tile height); no
// state change -- the caller flips cur / side once every band of the double step has been launched.
static int launch_fused2_rows(lbm_solver* s, int row_begin, int row_count, bool macros, cudaStream_t st) {
    if (row_count <= 0) return LBM_OK;
    StepArgs a = make_args(s, s->f[s->cur], s->f[s->cur ^ 1]);
    a.row_begin = row_begin; a.row_count = row_count;
    a.ghost2 = (char*)s->f[s->cur] + (size_t)s->cfg.batch * s->cavity * s->esz;
    cudaError_t e;
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
    const int vw = s->cfg.dtype == LBM_F64 ? s->vec_f64 : s->vec_f32;
    const bool vec = mode == MODE_STEP && gather && vw > 1;
    Launch L{};
    L.st = st;
    L.pdl = s->use_pdl && mode == MODE_STEP;
    block_shape(vec ? (s->cfg.nx + vw - 1) / vw : s->cfg.nx, row_count, s->cfg.batch, &L);
    if (L.grid.y > 65535u || L.grid.z > 65535u) {
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

static int region_rows(lbm_solver* s, int region, int rows[2][3], int* n) {
    // rows[i] = {begin, count, stride}
    const int nyl = s->nyl;
    *n = 0;
    if (region == LBM_REGION_ALL) {
        rows[0][0] = 0; rows[0][1] = nyl; rows[0][2] = 1; *n = 1;
    } else if (region == LBM_REGION_EDGE) {
        if (nyl == 1) { rows[0][0] = 0; rows[0][1] = 1; rows[0][2] = 1; }
        else { rows[0][0] = 0; rows[0][1] = 2; rows[0][2] = nyl - 1; }
        *n = 1;
    } else if (region == LBM_REGION_INTERIOR) {
        rows[0][0] = 1; rows[0][1] = nyl - 2; rows[0][2] = 1; *n = 1;
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
    if (c->batch < 1) return fail(LBM_EINVAL, "batch must be >= 1");
    if (c->dtype != LBM_F32 && c->dtype != LBM_F64) return fail(LBM_EINVAL, "dtype must be LBM_F32 or LBM_F64");
    if (c->collision < LBM_SRT || c->collision > LBM_MRT) return fail(LBM_EINVAL, "bad collision");
    if (c->turb != 0 && c->turb != 1) return fail(LBM_EINVAL, "turb must be 0 or 1");
    int nyl = c->ny_local == 0 ? c->ny : c->ny_local;
    if (c->ny_local == 0 && c->y0 != 0) return fail(LBM_EINVAL, "y0 must be 0 when ny_local == 0");
    if (c->y0 < 0 || nyl < 1 || c->y0 + nyl > c->ny) return fail(LBM_EINVAL, "y-strip [y0, y0+ny_local) outside [0, ny)");
    if (c->engine < LBM_ENGINE_AUTO || c->engine > LBM_ENGINE_TMA) return fail(LBM_EINVAL, "bad engine");
    if (c->semantics != LBM_SEMANTICS_C && c->semantics != LBM_SEMANTICS_A) return fail(LBM_EINVAL, "bad semantics");
    if (c->reserved != 0) return fail(LBM_EINVAL, "reserved must be 0");
    if (c->semantics == LBM_SEMANTICS_A) {
        if (c->collision != LBM_SRT || c->turb) return fail(LBM_EINVAL, "semantics A (MRT.py) is SRT without turbulence model");
        if (nyl != c->ny) return fail(LBM_EINVAL, "semantics A does not support y-strips");
        if (c->ny > 65535) return fail(LBM_EINVAL, "semantics A supports ny <= 65535");
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
    cudaFree(s->active); cudaFree(s->usum);
    cudaFree(s->staging); cudaFree(s->scratch);
    for (int i = 0; i < 4; ++i) if (s->graph[i]) cudaGraphExecDestroy(s->graph[i]);
    if (s->capture_stream) cudaStreamDestroy(s->capture_stream);
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
    if (const char* ev = getenv("LBM_B200_ENGINE")) {         // development override
        if (!strcmp(ev, "tma")) s->engine = LBM_ENGINE_TMA;
        if (!strcmp(ev, "ldg")) s->engine = LBM_ENGINE_LDG;
    }
    if (const char* ev = getenv("LBM_B200_VEC_F64")) s->vec_f64 = atoi(ev) == 2 ? 2 : 1;
    if (const char* ev = getenv("LBM_B200_VEC_F32")) s->vec_f32 = (atoi(ev) == 2 || atoi(ev) == 4) ? atoi(ev) : 1;
    if (const char* ev = getenv("LBM_B200_GRAPH")) s->use_graph = atoi(ev) != 0;
    if (const char* ev = getenv("LBM_B200_PDL")) s->use_pdl = atoi(ev) != 0;
    if (const char* ev = getenv("LBM_B200_FUSED2")) s->use_fused2 = atoi(ev) != 0;
    if (const char* ev = getenv("LBM_B200_FUSED2_TILE")) s->fused2_tile = atoi(ev);
    if (const char* ev = getenv("LBM_B200_TMA_VARIANT")) s->tma_variant = atoi(ev) % LBM_TMA_VARIANTS;
    if (const char* ev = getenv("LBM_B200_TMA_CTAS")) s->tma_ctas_per_sm = atoi(ev) > 0 ? atoi(ev) : 1;
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
        CKD(cudaMalloc(&s->f[1], s->state_bytes));
    }
    // ghost rows and pitch padding must hold finite values: they are read (and discarded) by masked lanes only in
    // the TMA family, but zero them once for determinism.
    CKD(cudaMemset(s->f[0], 0, s->state_bytes));
    CKD(cudaMemset(s->f[1], 0, s->state_bytes));
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
    if (cfg->turb) {
        CKD(cudaMalloc(&s->pi_eq, mbytes));
        CKD(cudaMalloc(&s->rho_prev, mbytes));
        CKD(cudaMemset(s->pi_eq, 0, mbytes));
        CKD(cudaMemset(s->rho_prev, 0, mbytes));
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
    s->cur = 0; s->pre = true; s->steps = 0;
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

// Move `ncav` cavities of `planes` [nx][nyl] planes each between a reference-layout array (`lin`: host, or device
// when on_device) and device planes.  lin_cavity_bytes: distance between cavities in `lin`; dev_*_stride in elements;
// dev_row0: element offset of local row 0 inside a device plane.
static int move_planes(lbm_solver* s, void* lin_base, size_t lin_cavity_bytes, bool on_device, bool to_device,
                       int ncav, int planes, void* dev_base, long long dev_plane_stride, long long dev_cavity_stride,
                       long long dev_row0, cudaStream_t st) {
    const int nx = s->cfg.nx, nyl = s->nyl;
    const size_t cav_bytes = (size_t)planes * nx * nyl * s->esz;
    dim3 blk(32, 8), grid((nx + 31) / 32, (nyl + 31) / 32, planes);
    if (!on_device) { int rc = ensure_staging(s, cav_bytes); if (rc) return rc; }
    for (int b = 0; b < ncav; ++b) {
        char* h = (char*)lin_base + (size_t)b * lin_cavity_bytes;
        void* lin = on_device ? (void*)h : s->staging;
        char* dev = (char*)dev_base + (size_t)b * dev_cavity_stride * s->esz;
        if (to_device && !on_device) CK(cudaMemcpyAsync(lin, h, cav_bytes, cudaMemcpyHostToDevice, st));
        if (s->esz == 8) {
            if (to_device) lbm_transpose<double, true><<<grid, blk, 0, st>>>((double*)dev, (double*)lin, nx, nyl, s->pitch, dev_plane_stride, dev_row0);
            else lbm_transpose<double, false><<<grid, blk, 0, st>>>((double*)dev, (double*)lin, nx, nyl, s->pitch, dev_plane_stride, dev_row0);
        } else {
            if (to_device) lbm_transpose<float, true><<<grid, blk, 0, st>>>((float*)dev, (float*)lin, nx, nyl, s->pitch, dev_plane_stride, dev_row0);
            else lbm_transpose<float, false><<<grid, blk, 0, st>>>((float*)dev, (float*)lin, nx, nyl, s->pitch, dev_plane_stride, dev_row0);
        }
        s->launches++;
        CK(cudaGetLastError());
        if (!to_device && !on_device) CK(cudaMemcpyAsync(h, lin, cav_bytes, cudaMemcpyDeviceToHost, st));
        if (!on_device && b + 1 < ncav) CK(cudaStreamSynchronize(st));   // staging is reused by the next cavity
    }
    if (!on_device) CK(cudaStreamSynchronize(st));
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
    s->cur = 0; s->pre = true; s->steps = 0;
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
    void* fin = s->f[s->cur];
    if (!s->pre) {
        // gather + wall rule (no collision) into the buffer the next step will overwrite anyway
        void* dst = s->f[s->cur ^ 1];
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
    if (s->cfg.semantics == LBM_SEMANTICS_A) return fail(LBM_ESTATE, "semantics A has no region stepping: use lbm_step");
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
    const int ty = fused2_tile_height(s), nyl = s->nyl;
    const int ntr = (nyl + ty - 1) / ty;                              // tile rows of the strip
    const int nb = (nyl % ty == 1 && ntr > 1) ? 2 : 1;                // bottom band must contain rows nyl-2 and nyl-1
    const bool split = ntr > 1 + nb;                                  // otherwise the edge bands are the whole strip
    const int bot0 = (ntr - nb) * ty;                                 // first row of the bottom band
    const bool wm = write_macros != 0;
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
    if (s->cfg.semantics == LBM_SEMANTICS_A) return fail(LBM_ESTATE, "semantics A has no region stepping: use lbm_step");
    s->cur ^= 1; s->side ^= 1; s->steps += 2;
    return LBM_OK;
}

int lbm_buffer_ptr(lbm_handle_t s, int which, void** ptr) {
    if (!s || !ptr) return fail(LBM_EINVAL, "NULL argument");
    *ptr = s->f[which ? (s->cur ^ 1) : s->cur];
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
    rc = launch_pass(s, s->f[s->cur], s->f[s->cur ^ 1], 0, s->nyl, 1, !s->pre, true, MODE_MACROS, st);
    if (rc) return rc;
    return macros_out(s, rho, u, on_device, st);
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
    int rc = set_device(s);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = s->cfg.batch;
    if (!s->active) {
        CK(cudaMalloc(&s->active, sizeof(int) * nb));
        s->active_host.assign(nb, 1);
        for (int i = 0; i < 4; ++i)
            if (s->graph[i]) { cudaGraphExecDestroy(s->graph[i]); s->graph[i] = nullptr; }   // kernel args change
    }
    const size_t cav_bytes = (size_t)s->cavity * s->esz;
    for (int b = 0; b < nb; ++b) {
        const int now = active[b] ? 1 : 0;
        if (s->active_host[b] && !now) {
            // freeze: both A/B buffers must hold the cavity's current populations, whatever the parity later on
            CK(cudaMemcpyAsync((char*)s->f[s->cur ^ 1] + b * cav_bytes, (char*)s->f[s->cur] + b * cav_bytes, cav_bytes,
                               cudaMemcpyDeviceToDevice, st));
        } else if (!s->active_host[b] && now) {
            return fail(LBM_ESTATE, "a frozen cavity cannot be re-activated");
        }
        s->active_host[b] = now;
    }
    CK(cudaMemcpyAsync(s->active, s->active_host.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
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
    return s->engine == LBM_ENGINE_TMA ? "tma" : "ldg";
}

}  // extern "C"
