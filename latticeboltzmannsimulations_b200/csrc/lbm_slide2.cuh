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
#include <cuda.h>
#include <cuda_runtime.h>

#include "lbm_device.cuh"

namespace lbm {

template <typename T, bool TURB = false> struct SlideCfg {
    static constexpr int TX = 512 / (int)sizeof(T);       // output columns of a strip: 64 fp64 / 128 fp32
    static constexpr int A = 16 / (int)sizeof(T);         // elements per 16-byte chunk: 2 / 4
    static constexpr int R = 4;                           // rows per iteration
    static constexpr int SW = TX + 2 * A;                 // staged columns [x0 - A, x0 + TX + A)
    static constexpr int CH = SW / A;                     // 16-byte chunks per staged row (34)
    static constexpr int NV = 8 / (int)sizeof(T);         // nodes per thread and sub-step: 1 fp64 / 2 fp32 (packed f32x2)
    static constexpr int WOFF = NV - 1;                   // window column of x0-1: keeps the pairs 8-byte aligned
    static constexpr int WW = TX + 2 + 2 * WOFF;          // window columns [x0 - 1, x0 + TX + 1) (+ alignment padding)
    static constexpr int WR = R + 2;                      // window rows
    static constexpr int ITEMS = R * 64;                  // main items per sub-step: R rows x 64 (pairs of) columns
    static constexpr int NT = 288;                        // threads: 9 warps >= ITEMS + 2 R ring nodes = 264
    static constexpr int STAGE = 9 * R * SW;              // elements per staging buffer
    static constexpr int NPL = TURB ? 11 : 9;             // window planes: populations (+ Smagorinsky pi, rho of state t+1)
    static constexpr int WINDOW = NPL * WR * WW;
    static constexpr int SIDE = TX + 2 + 4;               // lid density after sub-step 1 [TX + 2], corner carries [4]
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
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra SLIDE_DONE;\n"
        "bra SLIDE_WAIT;\n"
        "SLIDE_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity), "r"(2000u)     /* suspend-time hint (ns): sleep in hardware instead of spinning */
        : "memory");
}
// one row global -> shared through the bulk copy engine; bytes and both addresses are multiples of 16
// one 2-D box of the tensor map (columns c0 .., rows c1 ..) into shared memory, completing on the mbarrier
__device__ __forceinline__ void slide_tensor(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
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

// One item = NV x-adjacent nodes of a row handled by one thread: a double, or two floats as a packed f32x2.
template <typename T> struct SlideItem;
template <> struct SlideItem<double> {
    using AT = double;
    static __device__ __forceinline__ AT ld(const double* p) { return *p; }                    // nodes at p[0..NV)
    static __device__ __forceinline__ AT ldx(const double* p, int dx) { return p[dx]; }        // shifted by dx = +-1
    static __device__ __forceinline__ void st(double* p, AT v) { *p = v; }
    static __device__ __forceinline__ double get(AT v, int) { return v; }
};
template <> struct SlideItem<float> {
    using AT = f32x2;
    static __device__ __forceinline__ AT ld(const float* p) { return f32x2(*reinterpret_cast<const float2*>(p)); }
    static __device__ __forceinline__ AT ldx(const float* p, int dx) { return f32x2(p[dx], p[dx + 1]); }
    static __device__ __forceinline__ void st(float* p, AT v) { *reinterpret_cast<float2*>(p) = v.v; }
    static __device__ __forceinline__ float get(AT v, int i) { return i ? v.v.y : v.v.x; }
};

template <typename T, int COLL, bool MACROS, int MINB, bool TURB>
__global__ void __launch_bounds__(SlideCfg<T>::NT, MINB) lbm_step_slide2(const StepArgs a, const __grid_constant__ CUtensorMap tmap) {
    using Cfg = SlideCfg<T, TURB>;
    constexpr int TX = Cfg::TX, A = Cfg::A, R = Cfg::R, SW = Cfg::SW, WW = Cfg::WW, WR = Cfg::WR, NT = Cfg::NT;
    constexpr int NV = Cfg::NV, WOFF = Cfg::WOFF;
    constexpr int E = (int)sizeof(T);
    using Item = SlideItem<T>;
    using AT = typename Item::AT;
    static_assert(R == 4, "the copy assignment below fixes the staged row per thread as (tid & 3)");
    extern __shared__ __align__(128) unsigned char slide_smem[];
    T* stg = reinterpret_cast<T*>(slide_smem);                     // [2][9][R][SW]
    T* win = stg + 2 * Cfg::STAGE;                                 // [9][WR][WW]
    T* rl1 = win + Cfg::WINDOW;                                    // [TX + 2] lid density after sub-step 1
    T* c1 = rl1 + TX + 2;                                          // [4]  corner carries after sub-step 1
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int b = blockIdx.z;
    if (a.active && !a.active[b]) return;                          // frozen (converged) cavity
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TX;
    const int ya = a.row_begin + blockIdx.y * (a.seg_stride ? a.seg_stride : a.seg_h);   // segment [ya, yb) of local rows
    const int row_end = a.row_begin + a.row_count;
    const int yb = ya + a.seg_h < row_end ? ya + a.seg_h : row_end;
    const T* __restrict__ src = static_cast<const T*>(a.src) + (long long)b * a.cavity;
    T* __restrict__ dst = static_cast<T*>(a.dst) + (long long)b * a.cavity;
    const long long P = a.plane;
    const long long pitch = a.pitch;
    const Rates<T> rt(a.cav[b]);
    const Rates<AT> rta(a.cav[b]);                                 // the same rates for the item arithmetic
    const T* carry_in = static_cast<const T*>(a.carry) + b * 4;
    const T* rl_in = static_cast<const T*>(a.rho_lid) + (long long)b * a.pitch;
    const unsigned stg_sa = slide_sa(stg);
    // Smagorinsky state (whole cavities only): read side t-1, written side t+1
    const T* pi_in = TURB ? static_cast<const T*>(a.pi_eq) + (long long)b * a.mplane : nullptr;
    const T* rp_in = TURB ? static_cast<const T*>(a.rho_prev) + (long long)b * a.mplane : nullptr;
    T* pi_out = TURB ? static_cast<T*>(a.pi_eq_out) + (long long)b * a.mplane : nullptr;
    T* rp_out = TURB ? static_cast<T*>(a.rho_prev_out) + (long long)b * a.mplane : nullptr;
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

    // copy assignment, fixed for the whole segment: the first R lanes of warp k stage rows j = 0 .. R-1 of population k
    // (nine warps, nine populations).  A bulk copy is issued by one elected lane at a time, so spreading the 9 R copies
    // of a stage over all warps keeps the issue off any single warp's critical path.
    static_assert(NT == 9 * 32, "one warp per population");
    const int ck = tid >> 5, cj = tid & 31;
    const bool copier = cj < R;
    const int cdy = (ck == 2 || ck == 5 || ck == 6) ? 1 : ((ck == 4 || ck == 7 || ck == 8) ? -1 : 0);

    // ---- stage the source rows of the iteration whose first sub-step-1 row is s into buffer `buf` -----------------
    auto issue = [&](int s, int buf) {
        if (s <= yb && copier) {
            const unsigned bar = bar_full + 8 * buf;
            // interior block: rows s-1 .. s+R all exist in the main buffer (no wall row, no ghost row beyond)
            const int gy0 = a.y0 + s;
            const bool yin = gy0 - 1 >= 0 && gy0 + R <= a.ny - 1 && s - 1 >= -1 && s + R <= a.nyl;
            const unsigned d0 = stg_sa + (unsigned)((((buf * 9 + ck) * R + cj) * SW) * E);
            if (a.slide_tma && yin) {
                // one tensor copy per population: box = SW columns x R rows at (x0 - A, first source row); columns
                // outside the stored row (wall strips) arrive as zeros and are never used
                if (cj == 0) {
                    slide_mbar_expect(bar, R * SW * E);
                    slide_tensor(d0, &tmap, x0 - A, (b * 9 + ck) * (a.nyl + 2) + s + cdy + 1, bar);
                } else {
                    slide_mbar_arrive(bar);
                }
            } else if (xin && yin) {
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

    // item assignment, fixed for the whole segment: thread t < ITEMS owns row j = t / 64 and the NV nodes starting at
    // column x0 + NV (t % 64) in both sub-steps; threads ITEMS .. ITEMS + 2R - 1 own one ring node each in sub-step 1
    const bool main_item = tid < Cfg::ITEMS;
    const bool ring_item = !main_item && tid < Cfg::ITEMS + 2 * R;
    // wall lanes: the next 2R lanes of the ninth warp (idle otherwise) own the side-wall node of row j -- even lane the
    // left wall, odd lane the right one -- if this strip holds that column.  They run the general path for it in both
    // sub-steps next to the ring lanes, so that in rows between lid and bottom no main warp ever leaves the wall-free
    // path: a wall item computes all its nodes as if interior and stores only those that are (without the closure;
    // with it, wall items take the general path themselves)
    constexpr bool WALL_LANES = !TURB;
    const int wl = tid - (Cfg::ITEMS + 2 * R);
    const int wall_x = (wl & 1) ? a.nx - 1 : 0;
    const bool wall_lane = WALL_LANES && wl >= 0 && wl < 2 * R && wall_x >= x0 && wall_x < x0 + TX;
    const int ij = main_item ? tid >> 6 : (ring_item ? (tid - Cfg::ITEMS) >> 1 : (wl >> 1) & (R - 1));    // row within the block
    const int itx = main_item ? NV * (tid & 63) : wall_x - x0;                     // sub-step-2 column (x0 + itx)
    const int ilx = ring_item ? ((tid & 1) ? TX + 1 : 0) : itx + 1;                // sub-step-1 column index (x0 - 1 + ilx)
    // the item's NV nodes all exist and none of them lies on a side wall: with an interior row the item takes the
    // wall-free path.  Decided per item, not per block: a block that hangs over the segment's last row or holds the
    // lid / bottom row only sends those rows the general way (narrow cavities are mostly wall strips).
    const bool item_x_in = main_item && x0 + itx > 0 && x0 + itx + NV - 1 < a.nx - 1;
    // an item of several nodes that holds a wall column or hangs over the last column
    const bool item_x_edge = WALL_LANES && NV > 1 && main_item && !item_x_in && x0 + itx < a.nx;
    // local rows whose nodes are computed in this segment and lie on neither the lid nor the bottom wall
    const int lo1 = max(1 - a.y0, ya - 1), hi1 = min(a.ny - 2 - a.y0, yb);         // sub-step 1: rows ya-1 .. yb
    const int lo2 = max(1 - a.y0, ya), hi2 = min(a.ny - 2 - a.y0, yb - 1);         // sub-step 2: rows ya .. yb-1

    // one sub-step-1 node through the general (wall-aware) path: staged populations -> window
    auto s1_node = [&](const T* S, int s, int lx, T* w) {
        const int q = s + ij, x = x0 - 1 + lx;
        const int y = a.y0 + q;
        if (q > yb || x < 0 || x >= a.nx || y < 0 || y >= a.ny) return;
        const bool left = x == 0, right = x == a.nx - 1, lid = y == 0, bot = y == a.ny - 1;
        const T* c = S + ij * SW + lx + (A - 1);                   // this node in population 0's staged rows
        T f[9];
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
        if (TURB) {
            const long long m = (long long)q * pitch + x;
            T pi1, ir1;
            const T om = smagorinsky_omega<T>(f, pi_in[m], rp_in[m], rt.tau0);
            node_update<T, COLL, false, true>(f, rt, left, right, lid, bot, rho, ux, uy, om, &pi1, &ir1);
            w[9 * WR * WW] = pi1;
            w[10 * WR * WW] = ir1;
        } else {
            node_update<T, COLL, false>(f, rt, left, right, lid, bot, rho, ux, uy);
        }
        if (lid) rl1[lx] = rho;
#pragma unroll
        for (int k = 0; k < 9; ++k) w[k * WR * WW] = f[k];
    };
    // one sub-step-2 node through the general path: window -> global
    auto s2_node = [&](int s, int tx, const T* pc, const T* pu, const T* pd) {
        const int yl = s - 1 + ij, x = x0 + tx;
        if (yl < ya || yl >= yb || x >= a.nx) return;
        const int y = a.y0 + yl;
        const bool left = x == 0, right = x == a.nx - 1, lid = y == 0, bot = y == a.ny - 1;
        T f[9];
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
            slide_walls<T>(f, left, right, lid, bot, lid ? rl1[tx + 1] : (T)1, rt.uLB, c1, static_cast<T*>(a.carry_out) + b * 4);
        T rho, ux, uy;
        if (TURB) {
            T pi2, ir2;
            const T om = smagorinsky_omega<T>(f, pc[9 * WR * WW], pc[10 * WR * WW], rt.tau0);
            node_update<T, COLL, MACROS, true>(f, rt, left, right, lid, bot, rho, ux, uy, om, &pi2, &ir2);
            const long long mt = (long long)yl * pitch + x;
            pi_out[mt] = pi2;
            rp_out[mt] = ir2;
        } else {
            node_update<T, COLL, MACROS>(f, rt, left, right, lid, bot, rho, ux, uy);
        }
        if (lid) static_cast<T*>(a.rho_lid_out)[(long long)b * a.pitch + x] = rho;
        T* d = dst + (long long)(yl + 1) * pitch + x;
#pragma unroll
        for (int k = 0; k < 9; ++k) d[k * P] = f[k];
        if (MACROS) {
            const long long m = (long long)b * a.mplane + (long long)yl * pitch + x;
            static_cast<T*>(a.rho)[m] = rho;
            static_cast<T*>(a.ux)[m] = ux;
            static_cast<T*>(a.uy)[m] = uy;
        }
    };

    const int s0 = ya - 1;
    issue(s0, 0);
    issue(s0 + R, 1);
    slide_mbar_arrive(bar_free);                                   // the window is free for the first iteration
    int wbase = 0;                                                 // window slot of sub-step-1 row s
    int buf = 0;
    unsigned it = 0;                                               // iteration count: parities of the barriers
    for (int s = s0; s <= yb; s += R, ++it) {
        const T* S = stg + buf * Cfg::STAGE;
        // ---- sub-step 1: rows [s, s+R) x columns [x0-1, x0+TX] -> window ----
        // this item's row is computed in this segment and is neither the lid nor the bottom row
        const bool rin1 = s + ij >= lo1 && s + ij <= hi1;
        const bool inner1 = rin1 && item_x_in, edge1 = rin1 && item_x_edge;
        AT pi0, rp0;
        if (TURB && inner1) {                         // Smagorinsky state t-1 of this item's nodes: plain
            const long long m = (long long)(s + ij) * pitch + (x0 + itx);      // loads, in flight during the wait below
            pi0 = Item::ld(pi_in + m);
            rp0 = Item::ld(rp_in + m);
        }
        slide_mbar_wait(bar_full + 8 * buf, (it >> 1) & 1);        // the staged rows of this iteration have landed
        int ws = wbase + ij;
        ws = ws >= WR ? ws - WR : ws;
        T* w = win + ws * WW + ilx + WOFF;
        if (inner1 || edge1) {
            const T* c = S + ij * SW + ilx + (A - 1);              // first node of the item in population 0's staged rows
            AT f[9];
            f[0] = Item::ld(c);
            f[1] = Item::ldx(c + 1 * R * SW, -1);
            f[2] = Item::ld(c + 2 * R * SW);
            f[3] = Item::ldx(c + 3 * R * SW, 1);
            f[4] = Item::ld(c + 4 * R * SW);
            f[5] = Item::ldx(c + 5 * R * SW, -1);
            f[6] = Item::ldx(c + 6 * R * SW, 1);
            f[7] = Item::ldx(c + 7 * R * SW, 1);
            f[8] = Item::ldx(c + 8 * R * SW, -1);
            AT rho, ux, uy, pi1, ir1;
            if (TURB) {
                const AT om = smagorinsky_omega<AT>(f, pi0, rp0, rta.tau0);
                node_update<AT, COLL, false, true>(f, rta, false, false, false, false, rho, ux, uy, om, &pi1, &ir1);
            } else {
                node_update<AT, COLL, false>(f, rta, false, false, false, false, rho, ux, uy);
            }
            slide_mbar_wait(bar_free, it & 1);                     // everybody's sub-step 2 of the previous iteration has
            if (inner1) {                                          // read the window rows overwritten now
#pragma unroll
                for (int k = 0; k < 9; ++k) Item::st(w + k * WR * WW, f[k]);
                if (TURB) {
                    Item::st(w + 9 * WR * WW, pi1);
                    Item::st(w + 10 * WR * WW, ir1);
                }
            } else {                                               // wall item: keep the interior nodes only
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const int xv = x0 + itx + v;
                    if (xv > 0 && xv < a.nx - 1) {
#pragma unroll
                        for (int k = 0; k < 9; ++k) w[k * WR * WW + v] = Item::get(f[k], v);
                    }
                }
            }
        } else {
            slide_mbar_wait(bar_free, it & 1);
            if (main_item) {
                if (!WALL_LANES || !rin1) {                        // (else: the wall lane has this item's only node)
#pragma unroll
                    for (int v = 0; v < NV; ++v) s1_node(S, s, ilx + v, w + v);
                }
            } else if (ring_item || (wall_lane && rin1)) {
                s1_node(S, s, ilx, w);
            }
        }
        __syncthreads();                                           // window rows [s, s+R) complete; staging buffer free
        issue(s + 2 * R, buf);
        // ---- sub-step 2: rows [s-1, s+R-1) x columns [x0, x0+TX) <- window ----
        // this item's row belongs to the segment and is neither the lid nor the bottom row
        const bool rin2 = s - 1 + ij >= lo2 && s - 1 + ij <= hi2;
        const bool inner2 = rin2 && item_x_in, edge2 = rin2 && item_x_edge;
        if (main_item || (wall_lane && rin2)) {
            int wc = wbase + ij - 1;                               // window slot of row yl, of yl - 1 and of yl + 1
            wc = wc < 0 ? wc + WR : (wc >= WR ? wc - WR : wc);
            int wu = wc - 1;
            wu = wu < 0 ? wu + WR : wu;
            int wd = wc + 1;
            wd = wd >= WR ? wd - WR : wd;
            const T* pc = win + wc * WW + itx + 1 + WOFF;
            const T* pu = win + wu * WW + itx + 1 + WOFF;
            const T* pd = win + wd * WW + itx + 1 + WOFF;
            if (inner2 || edge2) {
                AT f[9];
                f[0] = Item::ld(pc);
                f[1] = Item::ldx(pc + 1 * WR * WW, -1);
                f[3] = Item::ldx(pc + 3 * WR * WW, 1);
                f[2] = Item::ld(pd + 2 * WR * WW);
                f[5] = Item::ldx(pd + 5 * WR * WW, -1);
                f[6] = Item::ldx(pd + 6 * WR * WW, 1);
                f[4] = Item::ld(pu + 4 * WR * WW);
                f[7] = Item::ldx(pu + 7 * WR * WW, 1);
                f[8] = Item::ldx(pu + 8 * WR * WW, -1);
                AT rho, ux, uy;
                const int yl = s - 1 + ij;
                if (TURB) {
                    AT pi2, ir2;
                    const AT om = smagorinsky_omega<AT>(f, Item::ld(pc + 9 * WR * WW), Item::ld(pc + 10 * WR * WW), rta.tau0);
                    node_update<AT, COLL, MACROS, true>(f, rta, false, false, false, false, rho, ux, uy, om, &pi2, &ir2);
                    const long long mt = (long long)yl * pitch + (x0 + itx);
                    Item::st(pi_out + mt, pi2);
                    Item::st(rp_out + mt, ir2);
                } else {
                    node_update<AT, COLL, MACROS>(f, rta, false, false, false, false, rho, ux, uy);
                }
                T* d = dst + (long long)(yl + 1) * pitch + (x0 + itx);
                const long long m = (long long)b * a.mplane + (long long)yl * pitch + (x0 + itx);
                if (inner2) {
#pragma unroll
                    for (int k = 0; k < 9; ++k) Item::st(d + k * P, f[k]);
                    if (MACROS) {
                        Item::st(static_cast<T*>(a.rho) + m, rho);
                        Item::st(static_cast<T*>(a.ux) + m, ux);
                        Item::st(static_cast<T*>(a.uy) + m, uy);
                    }
                } else {                                           // wall item: keep the interior nodes only
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        const int xv = x0 + itx + v;
                        if (xv > 0 && xv < a.nx - 1) {
#pragma unroll
                            for (int k = 0; k < 9; ++k) d[k * P + v] = Item::get(f[k], v);
                            if (MACROS) {
                                static_cast<T*>(a.rho)[m + v] = Item::get(rho, v);
                                static_cast<T*>(a.ux)[m + v] = Item::get(ux, v);
                                static_cast<T*>(a.uy)[m + v] = Item::get(uy, v);
                            }
                        }
                    }
                }
            } else if (main_item) {
                if (!WALL_LANES || !rin2) {                        // (else: the wall lane has this item's only node)
#pragma unroll
                    for (int v = 0; v < NV; ++v) s2_node(s, itx + v, pc + v, pu + v, pd + v);
                }
            } else {
                s2_node(s, itx, pc, pu, pd);                       // wall lane
            }
        }
        slide_mbar_arrive(bar_free);                               // this thread is done reading the window
        wbase += R;
        wbase = wbase >= WR ? wbase - WR : wbase;
        buf ^= 1;
    }
}

}  // namespace lbm
