// AA-pattern one-step kernels: ONE population buffer instead of the A/B pair (north_star: "A/B or AA-pattern buffers").
//
// The buffer alternates between two layouts:
//   NATURAL   slot k of node z holds the population that ARRIVES at z along k -- the pre-collision `fin` of the
//             reference.  An uploaded / initialised state is in this layout and complete; after an ODD step the slots a
//             wall node would receive from outside the cavity are stale (nobody streams into them) and are rebuilt by
//             the wall rule on read, exactly as the A/B kernels rebuild a missing pull source.
//   SWAPPED   slot opp(k) of node z holds the post-collision f*_k(z): after every EVEN step.
// EVEN step (NATURAL -> SWAPPED): a thread reads its node's nine slots, collides, writes f*_k into slot opp(k) of the
//             same node.
// ODD step  (SWAPPED -> NATURAL): a thread pulls f_k from slot opp(k) of the neighbour the population comes from
//             (x - c_kx, y + c_ky), collides, and pushes f*_k into slot k of the neighbour it goes to (x + c_kx, y - c_ky).
// In both steps the set of addresses a thread reads is the set it writes, and no other thread touches it: one buffer,
// no ordering between threads.  The gathered populations of a node are the ones the A/B kernel lbm_step_ldg gathers,
// and everything after the gather is the same code (wall_rule / node_update of lbm_device.cuh): results are
// bit-identical to the A/B one-step kernels.  Traffic per node and step is the same 9 loads + 9 stores; what the
// pattern buys is memory -- a 32768^2 fp64 cavity needs 77 GB instead of 155 GB.
// Whole cavities only (no y-strips, no frozen cavities); MODE_FINALIZE writes the reference's `fin` into a separate
// scratch buffer (download), MODE_MACROS evaluates the current moments; neither changes the buffer.
#pragma once
#include <cuda_runtime.h>

#include "lbm_device.cuh"

namespace lbm {

#if defined(__CUDACC__)
#define LBM_AA_PDL_PROLOGUE()                                              \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");       \
    asm volatile("griddepcontrol.wait;" ::: "memory")
#else
#define LBM_AA_PDL_PROLOGUE() (void)0        // host compilation of this header by the CPU emulation tests
#endif

// ODD: the buffer is in the SWAPPED layout (else NATURAL).  WALLS: rebuild the populations a wall node cannot receive
// (false only for the first step after an upload / initialisation, whose NATURAL state is complete; ODD implies WALLS).
template <typename T, int COLL, bool ODD, bool WALLS, bool MACROS, int MODE, bool TURB = false>
__global__ void __launch_bounds__(256) lbm_step_aa(const StepArgs a) {
    LBM_AA_PDL_PROLOGUE();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int lr = blockIdx.y * blockDim.y + threadIdx.y;          // launch row
    if (x >= a.nx || lr >= a.row_count) return;
    const int yl = a.row_begin + lr * a.row_stride;
    const int b = blockIdx.z;
    const int y = a.y0 + yl;
    const bool left = (x == 0), right = (x == a.nx - 1), lid = (y == 0), bot = (y == a.ny - 1);
    // MODE_STEP works in place on a.dst; the read-only modes read a.src (the same buffer) and write elsewhere
    T* buf = (MODE == MODE_STEP ? static_cast<T*>(a.dst) : const_cast<T*>(static_cast<const T*>(a.src))) + (long long)b * a.cavity;
    const long long P = a.plane;
    const long long rc = (long long)(yl + 1) * a.pitch + x;   // this node
    const long long ru = rc - a.pitch;                         // row y-1 (towards the lid)
    const long long rd = rc + a.pitch;                         // row y+1
    const Rates<T> r(a.cav[b]);
    // population k has no source inside the cavity
    const bool m1 = left, m2 = bot, m3 = right, m4 = lid;
    const bool m5 = left || bot, m6 = right || bot, m7 = right || lid, m8 = left || lid;

    T f[9];
    if (ODD) {
        // pull from the neighbours' opposite slots: f_k = f*_k(x - c_kx, y + c_ky), stored there in slot opp(k)
        f[0] = buf[rc];
        f[1] = m1 ? (T)0 : buf[3 * P + rc - 1];
        f[2] = m2 ? (T)0 : buf[4 * P + rd];
        f[3] = m3 ? (T)0 : buf[1 * P + rc + 1];
        f[4] = m4 ? (T)0 : buf[2 * P + ru];
        f[5] = m5 ? (T)0 : buf[7 * P + rd - 1];
        f[6] = m6 ? (T)0 : buf[8 * P + rd + 1];
        f[7] = m7 ? (T)0 : buf[5 * P + ru + 1];
        f[8] = m8 ? (T)0 : buf[6 * P + ru - 1];
    } else {
        f[0] = buf[rc];
        f[1] = (WALLS && m1) ? (T)0 : buf[1 * P + rc];
        f[2] = (WALLS && m2) ? (T)0 : buf[2 * P + rc];
        f[3] = (WALLS && m3) ? (T)0 : buf[3 * P + rc];
        f[4] = (WALLS && m4) ? (T)0 : buf[4 * P + rc];
        f[5] = (WALLS && m5) ? (T)0 : buf[5 * P + rc];
        f[6] = (WALLS && m6) ? (T)0 : buf[6 * P + rc];
        f[7] = (WALLS && m7) ? (T)0 : buf[7 * P + rc];
        f[8] = (WALLS && m8) ? (T)0 : buf[8 * P + rc];
    }
    if ((ODD || WALLS) && (left || right || lid || bot)) {
        const int slot = corner_slot(left, right, lid, bot);
        T* carry = static_cast<T*>(a.carry) + b * 4;
        const T stale = slot >= 0 ? carry[slot] : (T)0;
        const T rl = lid ? static_cast<const T*>(a.rho_lid)[(long long)b * a.pitch + x] : (T)1;
        wall_rule<T>(f, left, right, lid, bot, rl, r.uLB, stale);
        if (MODE == MODE_STEP && slot >= 0) carry[slot] = corner_value<T>(f, slot);
    }

    if (MODE == MODE_FINALIZE) {          // the reference's `fin` in the natural layout, into the scratch buffer
        T* out = static_cast<T*>(a.dst) + (long long)b * a.cavity;
#pragma unroll
        for (int k = 0; k < 9; ++k) out[k * P + rc] = f[k];
        return;
    }

    T rho, ux, uy;
    if (MODE == MODE_MACROS) {
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
        if (ODD) {
            // push: f*_k goes to slot k of (x + c_kx, y - c_ky); a wall node keeps nothing of what would leave the cavity
            buf[rc] = f[0];
            if (!m3) buf[1 * P + rc + 1] = f[1];
            if (!m4) buf[2 * P + ru] = f[2];
            if (!m1) buf[3 * P + rc - 1] = f[3];
            if (!m2) buf[4 * P + rd] = f[4];
            if (!m7) buf[5 * P + ru + 1] = f[5];
            if (!m8) buf[6 * P + ru - 1] = f[6];
            if (!m5) buf[7 * P + rd - 1] = f[7];
            if (!m6) buf[8 * P + rd + 1] = f[8];
        } else {
            buf[rc] = f[0];
            buf[3 * P + rc] = f[1];
            buf[4 * P + rc] = f[2];
            buf[1 * P + rc] = f[3];
            buf[2 * P + rc] = f[4];
            buf[7 * P + rc] = f[5];
            buf[8 * P + rc] = f[6];
            buf[5 * P + rc] = f[7];
            buf[6 * P + rc] = f[8];
        }
    }
    if (MACROS || MODE == MODE_MACROS) {
        const long long m = (long long)b * a.mplane + (long long)yl * a.pitch + x;
        static_cast<T*>(a.rho)[m] = rho;
        static_cast<T*>(a.ux)[m] = ux;
        static_cast<T*>(a.uy)[m] = uy;
    }
}

}  // namespace lbm
