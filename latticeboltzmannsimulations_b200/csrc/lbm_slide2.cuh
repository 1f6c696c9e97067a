// The sliding-window temporal-blocking kernel: two lattice steps per pass over memory, one CTA per column strip segment.
//
// A CTA owns a strip of TX columns (512 bytes per row) and slides down a segment of `seg_h` rows, R = 4 rows per
// iteration:
//   stage    the source rows of the NEXT-BUT-ONE iteration are copied global -> shared with 16-byte cp.async into one
//            of two staging buffers [9][R][TX + 2A] (A = elements per 16 bytes: the halo chunk on either side), so
//            that two iterations of loads are in flight per CTA without holding registers;
//   S1       sub-step 1 (state t -> t+1) on rows [s, s+R) x columns [x0-1, x0+TX]: every thread pulls its nine
//            populations from the staging buffer (the x / y shifts are plain shared-memory offsets), applies the wall
//            rule where needed, collides, and writes the post-collision populations into a rolling window of R+2 rows
//            [9][R+2][TX+2] in shared memory;
//   S2       sub-step 2 (t+1 -> t+2) on rows [s-1, s+R-1) x columns [x0, x0+TX) pulls from that window and stores
//            with aligned, coalesced 512-byte rows.
// Two block barriers per iteration (staging visible / window complete).  Redundant work: the two ring columns
// (2 / TX of sub-step 1) and one row above and below the segment (2 / seg_h) -- against 29 % for the 64x8 tiles of
// lbm_step_fused2 -- and no per-tile prologue: the loads of iteration i+2 overlap both sub-steps of iterations i, i+1.
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
    static constexpr size_t SMEM = (size_t)(2 * STAGE + WINDOW + SIDE) * sizeof(T);
};

__device__ __forceinline__ void slide_cp16(unsigned dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void slide_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void slide_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

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
    const unsigned stg_sa = (unsigned)__cvta_generic_to_shared(stg);
    // strip without wall columns whose staged halo chunks lie inside the row
    const bool xin = x0 >= A && x0 + TX + A <= a.nx;

    // copy assignment, fixed for the whole segment: thread = (group g, chunk ch); it stages chunk ch of row j = g & 3
    // of populations k = (g >> 2), (g >> 2) + 2, ...  (8 groups x 34 chunks = 272 of the 288 threads)
    const int cg = tid / Cfg::CH, cch = tid - cg * Cfg::CH;
    const int cj = cg & 3, ck0 = cg >> 2;
    const bool copier = cg < 8;
    const unsigned cdst = stg_sa + (unsigned)((cj * SW + cch * A) * E);      // + (buf * 9 + k) * R * SW * E

    // ---- stage the source rows of the iteration whose first sub-step-1 row is s into buffer `buf` -----------------
    auto issue = [&](int s, int buf) {
        if (s <= yb && copier) {
            // interior block: rows s-1 .. s+R all exist in the main buffer (no wall row, no ghost row beyond)
            const int gy0 = a.y0 + s;
            const bool yin = gy0 - 1 >= 0 && gy0 + R <= a.ny - 1 && s - 1 >= -1 && s + R <= a.nyl;
            const unsigned d0 = cdst + (unsigned)(buf * Cfg::STAGE * E);
            if (xin && yin) {
                const T* pb = src + (long long)(s + cj + 1) * pitch + (x0 - A + cch * A);     // population 0, row s + j
                if (ck0 == 0) {
                    slide_cp16(d0, pb);
                    slide_cp16(d0 + 2 * R * SW * E, pb + 2 * P + pitch);
                    slide_cp16(d0 + 4 * R * SW * E, pb + 4 * P - pitch);
                    slide_cp16(d0 + 6 * R * SW * E, pb + 6 * P + pitch);
                    slide_cp16(d0 + 8 * R * SW * E, pb + 8 * P - pitch);
                } else {
                    slide_cp16(d0 + 1 * R * SW * E, pb + P);
                    slide_cp16(d0 + 3 * R * SW * E, pb + 3 * P);
                    slide_cp16(d0 + 5 * R * SW * E, pb + 5 * P + pitch);
                    slide_cp16(d0 + 7 * R * SW * E, pb + 7 * P - pitch);
                }
            } else {
                const T* g2top = static_cast<const T*>(a.ghost2) + (long long)b * 6 * a.pitch;   // second ghost rows
                const T* g2bot = g2top + 3 * a.pitch;
                const int col = x0 - A + cch * A;
                if (col >= 0 && col < a.pitch && s + cj <= a.nyl) {      // chunk inside the stored row, row computed
                    for (int k = ck0; k < 9; k += 2) {
                        const int dy = (k == 2 || k == 5 || k == 6) ? 1 : ((k == 4 || k == 7 || k == 8) ? -1 : 0);
                        int q = s + cj + dy;                               // local source row
                        const int gq = a.y0 + q;
                        if (gq < 0 || gq > a.ny - 1) q = s + cj;          // source row beyond a wall: never used, stay in bounds
                        const T* p;
                        if (q == -2) p = g2top + (k == 4 ? 0 : k == 7 ? 1 : 2) * pitch + col;
                        else if (q == a.nyl + 1) p = g2bot + (k == 2 ? 0 : k == 5 ? 1 : 2) * pitch + col;
                        else p = src + k * P + (long long)(q + 1) * pitch + col;
                        slide_cp16(d0 + (unsigned)(k * R * SW * E), p);
                    }
                }
            }
        }
        slide_commit();
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
    int wbase = 0;                                                 // window slot of sub-step-1 row s
    int buf = 0;
    for (int s = s0; s <= yb; s += R) {
        slide_wait<1>();                                           // this thread's copies of the current buffer have landed
        __syncthreads();                                           // ... and everybody else's; sub-step 2 of the previous
                                                                   // iteration has finished reading the window
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
#pragma unroll
            for (int k = 0; k < 9; ++k) w[k * WR * WW] = f[k];
        }
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
        wbase += R;
        wbase = wbase >= WR ? wbase - WR : wbase;
        buf ^= 1;
    }
    slide_wait<0>();
}

}  // namespace lbm
