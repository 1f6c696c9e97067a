// The sliding-window temporal-blocking kernel: two lattice steps per pass over memory, one CTA per column strip segment.
//
// A CTA owns a strip of TX columns (512 bytes per row) and slides down a segment of `seg_h` rows, R = 4 rows per
// iteration:
//   stage    the source rows of the NEXT-BUT-ONE iteration are copied global -> shared by the bulk copy engine
//            (cp.async.bulk, one 544-byte row per population and row, completion on an mbarrier) into one of two
//            staging buffers [9][R][TX + 2A] (A = elements per 16 bytes: the halo chunk on either side): two
//            iterations of loads are in flight per CTA, no register and no LSU instruction is spent on them;
//   S1       sub-step 1 (state t -> t+1) on rows [s, s+R) x columns [x0-1, x0+TX]: every thread pulls its nine
//            populations from the staging buffer (the x / y shifts are plain shared-memory offsets), applies the wall
//            rule where needed, collides, and writes the post-collision populations into a rolling window of R+2 rows
//            [9][R+2][TX+2] in shared memory;
//   S2       sub-step 2 (t+1 -> t+2) on rows [s-1, s+R-1) x columns [x0, x0+TX) pulls from that window and stores
//            with aligned, coalesced 512-byte rows.
// Synchronisation per iteration: the "window free" barrier is split (every thread arrives on an mbarrier when its
// sub-step 2 is done and waits only just before its sub-step-1 results are written, i.e. after the arithmetic), the
// "window complete" barrier is an ordinary block barrier.  Redundant work: the two ring columns (2 / TX of sub-step 1)
// and one row above and below the segment (2 / seg_h) -- against 29 % for the 64x8 tiles of lbm_step_fused2.
// The per-node arithmetic is node_update()/wall_rule() of lbm_device.cuh: bit-identical to two one-step launches.
#pragma once
#include <cuda_runtime.h>

#include "lbm_device.cuh"

namespace lbm {

template <typename T> struct SlideCfg {
    static constexpr int TX = 512 / (int)sizeof(T);       // output columns of a strip: 64 fp64 / 128 fp32
    static constexpr int A = 16 / (int)sizeof(T);         // elements per 16-byte chunk: 2 / 4
    static constexpr int R = 4;                           // rows per iteration
    static constexpr int SW = TX + 2 * A;                 // staged columns [x0 - A, x0 + TX + A)
    static constexpr int CH = SW / A;                     // 16-byte chunks per staged row (34)
    static constexpr int WW = TX + 2;                     // window columns [x0 - 1, x0 + TX + 1)
    static constexpr int WR = R + 2;                      // window rows
    static constexpr int NT = 288;                        // threads: 9 warps >= R * WW = 264 sub-step-1 nodes (fp64)
    static constexpr int STAGE = 9 * R * SW;              // elements per staging buffer
    static constexpr int WINDOW = 9 * WR * WW;
    static constexpr int SIDE = WW + 4;                   // lid density after sub-step 1 [WW], corner carries [4]
    static constexpr int COPIERS = 9 * R;                 // one bulk copy per (population, row) of a stage
    static constexpr size_t DATA = (size_t)(2 * STAGE + WINDOW + SIDE) * sizeof(T);
    static constexpr size_t BAR_OFF = (DATA + 15) / 16 * 16;       // three mbarriers: staging full [2], window free
    static constexpr size_t SMEM = BAR_OFF + 3 * 8;
};

__device__ __forceinline__ unsigned slide_sa(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void slide_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void slide_mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void slide_mbar_expect(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void slide_mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SLIDE_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SLIDE_DONE;\n"
        "bra SLIDE_WAIT;\n"
        "SLIDE_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
// one row global -> shared through the bulk copy engine; bytes and both addresses are multiples of 16
__device__ __forceinline__ void slide_bulk(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// Apply the wall rule to the gathered populations of a (possibly) wall node; returns through f.
template <typename T>
__device__ __forceinline__ void slide_walls(T f[9], bool left, bool right, bool lid, bool bot, T rl, T uLB, const T* carry_in,
                                            T* carry_keep) {
    if (left) { f[1] = (T)0; f[5] = (T)0; f[8] = (T)0; }
    if (right) { f[3] = (T)0; f[6] = (T)0; f[7] = (T)0; }
    if (bot) { f[2] = (T)0; f[5] = (T)0; f[6] = (T)0; }
    if (lid) { f[4] = (T)0; f[7] = (T)0; f[8] = (T)0; }
    const int slot = corner_slot(left, right, lid, bot);
    const T stale = slot >= 0 ? carry_in[slot] : (T)0;
    wall_rule<T>(f, left, right, lid, bot, rl, uLB, stale);
    if (slot >= 0) carry_keep[slot] = corner_value<T>(f, slot);
}

template <typename T, int COLL, bool MACROS, int MINB>
__global__ void __launch_bounds__(SlideCfg<T>::NT, MINB) lbm_step_slide2(const StepArgs a) {
    using Cfg = SlideCfg<T>;
    constexpr int TX = Cfg::TX, A = Cfg::A, R = Cfg::R, SW = Cfg::SW, WW = Cfg::WW, WR = Cfg::WR, NT = Cfg::NT;
    constexpr int E = (int)sizeof(T);
    constexpr int P1 = (R * WW + NT - 1) / NT;                     // passes of sub-step 1 over the thread block
    constexpr int P2 = (R * TX + NT - 1) / NT;                     // passes of sub-step 2
    static_assert(R == 4, "the copy assignment below fixes the staged row per thread as (group & 3)");
    extern __shared__ __align__(16) unsigned char slide_smem[];
    T* stg = reinterpret_cast<T*>(slide_smem);                     // [2][9][R][SW]
    T* win = stg + 2 * Cfg::STAGE;                                 // [9][WR][WW]
    T* rl1 = win + Cfg::WINDOW;                                    // [WW] lid density after sub-step 1
    T* c1 = rl1 + WW;                                              // [4]  corner carries after sub-step 1
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int b = blockIdx.z;
    if (a.active && !a.active[b]) return;                          // frozen (converged) cavity
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TX;
    const int ya = a.row_begin + blockIdx.y * a.seg_h;             // segment [ya, yb) of local rows
    const int row_end = a.row_begin + a.row_count;
    const int yb = ya + a.seg_h < row_end ? ya + a.seg_h : row_end;
    const T* __restrict__ src = static_cast<const T*>(a.src) + (long long)b * a.cavity;
    T* __restrict__ dst = static_cast<T*>(a.dst) + (long long)b * a.cavity;
    const long long P = a.plane;
    const long long pitch = a.pitch;
    const Rates<T> rt(a.cav[b]);
    const T* carry_in = static_cast<const T*>(a.carry) + b * 4;
    const T* rl_in = static_cast<const T*>(a.rho_lid) + (long long)b * a.pitch;
    const unsigned stg_sa = slide_sa(stg);
    // strip without wall columns whose staged halo chunks lie inside the row
    const bool xin = x0 >= A && x0 + TX + A <= a.nx;

    // barriers: staging buffer full [2] (COPIERS arrivals + the bytes of their copies), window free (NT arrivals)
    const unsigned bar_full = slide_sa(slide_smem + Cfg::BAR_OFF);
    const unsigned bar_free = bar_full + 16;
    if (tid == 0) {
        slide_mbar_init(bar_full, Cfg::COPIERS);
        slide_mbar_init(bar_full + 8, Cfg::COPIERS);
        slide_mbar_init(bar_free, NT);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // copy assignment, fixed for the whole segment: thread t < 9 R stages row j = t & 3 of population k = t >> 2
    const int ck = tid >> 2, cj = tid & 3;
    const int cdy = (ck == 2 || ck == 5 || ck == 6) ? 1 : ((ck == 4 || ck == 7 || ck == 8) ? -1 : 0);

    // ---- stage the source rows of the iteration whose first sub-step-1 row is s into buffer `buf` -----------------
    auto issue = [&](int s, int buf) {
        if (s <= yb && tid < Cfg::COPIERS) {
            const unsigned bar = bar_full + 8 * buf;
            // interior block: rows s-1 .. s+R all exist in the main buffer (no wall row, no ghost row beyond)
            const int gy0 = a.y0 + s;
            const bool yin = gy0 - 1 >= 0 && gy0 + R <= a.ny - 1 && s - 1 >= -1 && s + R <= a.nyl;
            const unsigned d0 = stg_sa + (unsigned)((((buf * 9 + ck) * R + cj) * SW) * E);
            if (xin && yin) {
                slide_mbar_expect(bar, SW * E);
                slide_bulk(d0, src + ck * P + (long long)(s + cj + cdy + 1) * pitch + (x0 - A), SW * E, bar);
            } else {
                const int c0 = x0 - A < 0 ? 0 : x0 - A;                  // staged columns inside the stored row
                const int c1 = x0 + TX + A > a.pitch ? a.pitch : x0 + TX + A;
                int q = s + cj + cdy;                                    // local source row
                const int gq = a.y0 + q;
                if (gq < 0 || gq > a.ny - 1) q = s + cj;                // source row beyond a wall: never used, stay in bounds
                if (s + cj > a.nyl || c1 <= c0) {                        // row not computed
                    slide_mbar_arrive(bar);
                } else {
                    const T* g2top = static_cast<const T*>(a.ghost2) + (long long)b * 6 * a.pitch;   // second ghost rows
                    const T* g2bot = g2top + 3 * a.pitch;
                    const T* p;
                    if (q == -2) p = g2top + (ck == 4 ? 0 : ck == 7 ? 1 : 2) * pitch;
                    else if (q == a.nyl + 1) p = g2bot + (ck == 2 ? 0 : ck == 5 ? 1 : 2) * pitch;
                    else p = src + ck * P + (long long)(q + 1) * pitch;
                    const unsigned bytes = (unsigned)((c1 - c0) * E);
                    slide_mbar_expect(bar, bytes);
                    slide_bulk(d0 + (unsigned)((c0 - (x0 - A)) * E), p + c0, bytes, bar);
                }
            }
        }
    };

    // node assignment, fixed for the whole segment
    int j1[P1], lx1[P1], j2[P2], tx2[P2];
#pragma unroll
    for (int p = 0; p < P1; ++p) {
        const int n = tid + p * NT;
        j1[p] = n / WW; lx1[p] = n - j1[p] * WW;
    }
#pragma unroll
    for (int p = 0; p < P2; ++p) {
        const int n = tid + p * NT;
        j2[p] = n / TX; tx2[p] = n - j2[p] * TX;
    }

    const int s0 = ya - 1;
    issue(s0, 0);
    issue(s0 + R, 1);
    slide_mbar_arrive(bar_free);                                   // the window is free for the first iteration
    int wbase = 0;                                                 // window slot of sub-step-1 row s
    int buf = 0;
    unsigned it = 0;                                               // iteration count: parities of the barriers
    for (int s = s0; s <= yb; s += R, ++it) {
        slide_mbar_wait(bar_full + 8 * buf, (it >> 1) & 1);        // the staged rows of this iteration have landed
        bool window_free = false;                                  // waited for sub-step 2 of the previous iteration?
        const T* S = stg + buf * Cfg::STAGE;
        const int gys = a.y0 + s;
        // no wall node among the sub-step-1 nodes of this iteration (rows s .. s+R-1, columns x0-1 .. x0+TX), all exist
        const bool inner1 = x0 - 1 > 0 && x0 + TX < a.nx - 1 && gys > 0 && gys + R - 1 < a.ny - 1 && s + R - 1 <= yb;
        // ---- sub-step 1: rows [s, s+R) x columns [x0-1, x0+TX] -> window ----
#pragma unroll
        for (int p = 0; p < P1; ++p) {
            const int j = j1[p], lx = lx1[p];
            if (j >= R) continue;
            const T* c = S + j * SW + lx + (A - 1);                // this node in population 0's staged rows
            int ws = wbase + j;
            ws = ws >= WR ? ws - WR : ws;
            T* w = win + ws * WW + lx;
            T f[9];
            if (inner1) {
                f[0] = c[0];
                f[1] = c[1 * R * SW - 1];
                f[2] = c[2 * R * SW];
                f[3] = c[3 * R * SW + 1];
                f[4] = c[4 * R * SW];
                f[5] = c[5 * R * SW - 1];
                f[6] = c[6 * R * SW + 1];
                f[7] = c[7 * R * SW + 1];
                f[8] = c[8 * R * SW - 1];
                T rho, ux, uy;
                node_update<T, COLL, false>(f, rt, false, false, false, false, rho, ux, uy);
            } else {
                const int q = s + j, x = x0 - 1 + lx;
                const int y = a.y0 + q;
                if (q > yb || x < 0 || x >= a.nx || y < 0 || y >= a.ny) continue;
                const bool left = x == 0, right = x == a.nx - 1, lid = y == 0, bot = y == a.ny - 1;
                f[0] = c[0];
                f[1] = c[1 * R * SW - 1];
                f[2] = c[2 * R * SW];
                f[3] = c[3 * R * SW + 1];
                f[4] = c[4 * R * SW];
                f[5] = c[5 * R * SW - 1];
                f[6] = c[6 * R * SW + 1];
                f[7] = c[7 * R * SW + 1];
                f[8] = c[8 * R * SW - 1];
                if (left || right || lid || bot)
                    slide_walls<T>(f, left, right, lid, bot, lid ? rl_in[x] : (T)1, rt.uLB, carry_in, c1);
                T rho, ux, uy;
                node_update<T, COLL, false>(f, rt, left, right, lid, bot, rho, ux, uy);
                if (lid) rl1[lx] = rho;
            }
            if (!window_free) {                                    // everybody's sub-step 2 of the previous iteration
                slide_mbar_wait(bar_free, it & 1);                 // has read the window rows overwritten now
                window_free = true;
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) w[k * WR * WW] = f[k];
        }
        if (!window_free) slide_mbar_wait(bar_free, it & 1);       // (threads without a sub-step-1 node)
        __syncthreads();                                           // window rows [s, s+R) complete; staging buffer free
        issue(s + 2 * R, buf);
        // no wall node among the sub-step-2 nodes of this iteration (rows s-1 .. s+R-2, columns x0 .. x0+TX-1), all stored
        const bool inner2 = x0 > 0 && x0 + TX - 1 < a.nx - 1 && gys - 1 > 0 && gys + R - 2 < a.ny - 1 && s - 1 >= ya &&
                            s + R - 2 < yb;
        // ---- sub-step 2: rows [s-1, s+R-1) x columns [x0, x0+TX) <- window ----
#pragma unroll
        for (int p = 0; p < P2; ++p) {
            const int j = j2[p], tx = tx2[p];
            if (j >= R) continue;
            const int yl = s - 1 + j, x = x0 + tx;
            int wc = wbase + j - 1;                                // window slot of row yl, of yl - 1 and of yl + 1
            wc = wc < 0 ? wc + WR : (wc >= WR ? wc - WR : wc);
            int wu = wc - 1;
            wu = wu < 0 ? wu + WR : wu;
            int wd = wc + 1;
            wd = wd >= WR ? wd - WR : wd;
            const T* pc = win + wc * WW + tx + 1;
            const T* pu = win + wu * WW + tx + 1;
            const T* pd = win + wd * WW + tx + 1;
            T f[9];
            T rho, ux, uy;
            if (inner2) {
                f[0] = pc[0];
                f[1] = pc[1 * WR * WW - 1];
                f[3] = pc[3 * WR * WW + 1];
                f[2] = pd[2 * WR * WW];
                f[5] = pd[5 * WR * WW - 1];
                f[6] = pd[6 * WR * WW + 1];
                f[4] = pu[4 * WR * WW];
                f[7] = pu[7 * WR * WW + 1];
                f[8] = pu[8 * WR * WW - 1];
                node_update<T, COLL, MACROS>(f, rt, false, false, false, false, rho, ux, uy);
            } else {
                if (yl < ya || yl >= yb || x >= a.nx) continue;
                const int y = a.y0 + yl;
                const bool left = x == 0, right = x == a.nx - 1, lid = y == 0, bot = y == a.ny - 1;
                f[0] = pc[0];
                f[1] = pc[1 * WR * WW - 1];
                f[3] = pc[3 * WR * WW + 1];
                f[2] = pd[2 * WR * WW];
                f[5] = pd[5 * WR * WW - 1];
                f[6] = pd[6 * WR * WW + 1];
                f[4] = pu[4 * WR * WW];
                f[7] = pu[7 * WR * WW + 1];
                f[8] = pu[8 * WR * WW - 1];
                if (left || right || lid || bot)
                    slide_walls<T>(f, left, right, lid, bot, lid ? rl1[tx + 1] : (T)1, rt.uLB, c1,
                                   static_cast<T*>(a.carry_out) + b * 4);
                node_update<T, COLL, MACROS>(f, rt, left, right, lid, bot, rho, ux, uy);
                if (lid) static_cast<T*>(a.rho_lid_out)[(long long)b * a.pitch + x] = rho;
            }
            T* d = dst + (long long)(yl + 1) * pitch + x;
#pragma unroll
            for (int k = 0; k < 9; ++k) d[k * P] = f[k];
            if (MACROS) {
                const long long m = (long long)b * a.mplane + (long long)yl * pitch + x;
                static_cast<T*>(a.rho)[m] = rho;
                static_cast<T*>(a.ux)[m] = ux;
                static_cast<T*>(a.uy)[m] = uy;
            }
        }
        slide_mbar_arrive(bar_free);                               // this thread is done reading the window
        wbase += R;
        wbase = wbase >= WR ? wbase - WR : wbase;
        buf ^= 1;
    }
}

}  // namespace lbm
