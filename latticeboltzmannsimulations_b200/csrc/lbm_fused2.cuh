// The temporal-blocking kernel: two lattice steps per pass over memory.
#pragma once
#include <cuda_runtime.h>

#include "lbm_device.cuh"

namespace lbm {

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

}  // namespace lbm
