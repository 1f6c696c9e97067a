// Kernels of the "ldg" family (one-step fused pull kernels, scalar and vector forms), the semantics-A compatibility
// passes, and the auxiliary kernels (initial state, layout transposes, functions.equ, reductions, diagnostics).
// Host-side dispatch lives in lbm_b200.cu; the per-node arithmetic in lbm_device.cuh.
#pragma once
#include <cuda_runtime.h>

#include "lbm_device.cuh"

namespace lbm {

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
        const T inv = (T)1 / rho;                                  // same form as node_update
        ux = jx * inv; uy = jy * inv;
        if (left || right || bot) { ux = (T)0; uy = (T)0; }
        if (lid) { rho = rho_lid_formula<T>(f); ux = r.uLB; uy = (T)0; }
    } else {
        if (TURB) {
            const long long m = (long long)b * a.mplane + (long long)yl * a.pitch + x;
            T* pi = static_cast<T*>(a.pi_eq) + m;
            T* rp = static_cast<T*>(a.rho_prev) + m;
            const T om = smagorinsky_omega<T>(f, *pi, *rp, r.tau0);
            T pi_new, irho_new;
            node_update<T, COLL, MACROS, true>(f, r, left, right, lid, bot, rho, ux, uy, om, &pi_new, &irho_new);
            *pi = pi_new;
            *rp = irho_new;
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
    T pi_old[V], rp_old[V], pi_new[V], rp_new[V];
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
        if (TURB) { pi_new[v] = (T)0; rp_new[v] = (T)0; }
        if (TURB) {
            const T om = smagorinsky_omega<T>(f[v], pi_old[v], rp_old[v], r.tau0);
            node_update<T, COLL, MACROS, true>(f[v], r, left, right, lid, bot, rho[v], ux[v], uy[v], om, &pi_new[v], &rp_new[v]);
        } else {
            node_update<T, COLL, MACROS>(f[v], r, left, right, lid, bot, rho[v], ux[v], uy[v]);
        }
    }
    if (TURB) {
        const long long m = (long long)b * a.mplane + (long long)yl * a.pitch + x;
        if (x + V <= a.nx) {
            gstore<T, V>(static_cast<T*>(a.pi_eq) + m, pi_new);
            gstore<T, V>(static_cast<T*>(a.rho_prev) + m, rp_new);
        } else {
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (x + v < a.nx) { static_cast<T*>(a.pi_eq)[m + v] = pi_new[v]; static_cast<T*>(a.rho_prev)[m + v] = rp_new[v]; }
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
    {                                                                     // :396, the script's operation order
        T fe[9];
        feq_all_ref<T>(rho, ux, uy, fe);
#pragma unroll
        for (int k = 0; k < 9; ++k) f[k] = f[k] - r.omega * (f[k] - fe[k]);
    }
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
        feq_all_ref<T>(static_cast<const T*>(a.rho)[m], static_cast<const T*>(a.ux)[m], static_cast<const T*>(a.uy)[m], fe);
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
    feq_all_ref<double>(1.0, ux, 0.0, fe);
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
        feq_all_ref<T>(rho[i], ux[i], uy[i], fe);
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

// Halo rows of a y-strip <-> one contiguous buffer [9][nx] per neighbour (a single NCCL send / recv instead of nine).
// Row i of the buffer is row `row[i]` of population `pop[i]` of the strip buffer (PACK), or -- UNPACK -- the ghost row
// (i < 6) / the i-6-th second ghost row (i >= 6) on that side.  Order = distributed.halo_plan(deep=True).
struct HaloRows { int pop[9]; int row[9]; };
template <typename T, bool PACK>
__global__ void lbm_halo_rows(T* __restrict__ strip, T* __restrict__ g2side, T* __restrict__ buf, HaloRows h, int nx, int pitch,
                              long long plane) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= nx) return;
    const int i = blockIdx.y;
    T* p = (!PACK && i >= 6) ? g2side + (long long)(i - 6) * pitch : strip + h.pop[i] * plane + (long long)h.row[i] * pitch;
    if (PACK) buf[(long long)i * nx + x] = p[x];
    else p[x] = buf[(long long)i * nx + x];
}

// The stopping rule of MRT_GPU_datagen.py:726-733 for every cavity of a batch, on the device: one thread per cavity
// compares the mean of the stored velocity field with the one of the previous check, counts the hits (never reset, as
// in the reference) and retires the cavity when the count exceeds hits - 1.  newly[b] = 1 marks cavities retired by
// this call (their buffers are equalised by lbm_freeze_copy).
__global__ void lbm_converge_rule(const double* __restrict__ usum, double denom, const CavityParams* __restrict__ cav,
                                  double* __restrict__ past, int* __restrict__ count, int* __restrict__ active,
                                  int* __restrict__ newly, double tol, int hits, int nb) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    newly[b] = 0;
    if (!active[b]) return;
    const double mean = usum[b] / denom;
    if (fabs(mean - past[b]) / cav[b].uLB < tol) {
        const int c = ++count[b];
        if (c > hits - 1) { active[b] = 0; newly[b] = 1; }
    }
    past[b] = mean;
}

// Copy `n` bytes (a multiple of 16) per cavity from src to dst for the cavities flagged in newly[]: a retired cavity
// must hold the same state in both A/B buffers and both halves of the side arrays.
__global__ void lbm_freeze_copy(const int* __restrict__ newly, const char* __restrict__ src, char* __restrict__ dst,
                                size_t bytes_per_cavity) {
    const int b = blockIdx.y;
    if (!newly[b]) return;
    const uint4* s = reinterpret_cast<const uint4*>(src + (size_t)b * bytes_per_cavity);
    uint4* d = reinterpret_cast<uint4*>(dst + (size_t)b * bytes_per_cavity);
    const size_t n = bytes_per_cavity / 16;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
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

}  // namespace lbm
