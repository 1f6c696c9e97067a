// The marching temporal-blocking kernel: two lattice steps per pass over memory, one independent warp per work item.
//
// A warp owns an x-strip of 32*V columns and marches down a segment of `seg_h` rows.  Per row it
//   (1) runs sub-step 1 (state t -> t+1) on row r: pull from the global post-collision buffer exactly like the
//       one-step kernels; the nine source rows are copied global -> shared D rows ahead with cp.async into a per-warp
//       ring of D stages, so that (D-1) rows per warp are in flight at any time without holding registers,
//   (2) hands the nine post-collision populations of row r to the nodes that will pull them in sub-step 2: populations
//       that move in x travel between lanes by warp shuffle, populations that move in y wait in a rolling register
//       window (row r feeds sub-step 2 of rows r-1, r and r+1),
//   (3) runs sub-step 2 (t+1 -> t+2) on row r-1 from that window and stores it with aligned, coalesced stores.
// Nothing but the strip's own edge columns is computed twice: sub-step 1 of the two ring columns next to the strip
// (x0-1 and x0+32V) is done for 16 rows at a time by one extra pass (lane = 16 rows x 2 sides) that leaves the three
// populations crossing into the strip in a per-warp shared-memory scratch, and a segment recomputes one row above and
// below itself.  That is (1/16V + 2/seg_h) of sub-step 1 against 29 % for the 64x8 shared-memory tiles of
// lbm_step_fused2, with no block-level barrier and no index arithmetic per node (pointers advance by one pitch per row).
// Per node and TWO steps: 9 loads + 9 stores (+2+2 with the Smagorinsky closure).
// The per-node arithmetic is node_update()/wall_rule() of lbm_device.cuh: results are bit-identical to two one-step
// launches.  Strips that touch a side wall and the lid / bottom rows take a general path with the wall predicates; all
// other (strip, row) pairs run a predicate-free instantiation of the same code.
#pragma once
#include <cuda_runtime.h>

#include "lbm_device.cuh"

namespace lbm {

template <typename T, int V> struct MarchVec;
template <> struct MarchVec<float, 1> { using type = float; };
template <> struct MarchVec<float, 2> { using type = float2; };
template <> struct MarchVec<float, 4> { using type = float4; };
template <> struct MarchVec<double, 1> { using type = double; };
template <> struct MarchVec<double, 2> { using type = double2; };

template <typename T, int V>
__device__ __forceinline__ void mload(const T* p, T out[V]) {
    using VT = typename MarchVec<T, V>::type;
    const VT v = *reinterpret_cast<const VT*>(p);
    const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
    for (int i = 0; i < V; ++i) out[i] = e[i];
}
template <typename T, int V>
__device__ __forceinline__ void mstore(T* p, const T in[V]) {
    using VT = typename MarchVec<T, V>::type;
    VT v;
    T* e = reinterpret_cast<T*>(&v);
#pragma unroll
    for (int i = 0; i < V; ++i) e[i] = in[i];
    *reinterpret_cast<VT*>(p) = v;
}

template <typename T, int V, bool TURB, int D> struct MarchCfg {
    static constexpr int COLS = 32 * V;           // columns of a warp's strip
    static constexpr int RING_ROWS = 16;          // rows per ring pass (x 2 sides = 32 lanes)
    static constexpr int SCRATCH = 6 * RING_ROWS + COLS + 4;   // per warp: ring[2][3][16], lid density [COLS], carries [4]
    // one stage of the cp.async ring = the nine source rows of one sub-step-1 row.
    //   V == 1: [k][lane] -- every lane copies exactly the element it pulls, nobody else reads it (no warp sync);
    //   V  > 1: [k][A + 32V + A] -- element x0+j at index A+j, the neighbouring strips' columns x0-1 and x0+32V at
    //           A-1 and A+32V (A = elements per 16 bytes keeps the vector slots 16-byte aligned).
    static constexpr int A = 16 / (int)sizeof(T);
    static constexpr int SROW = V == 1 ? 32 : COLS + 2 * A;
    static constexpr int STAGE = 9 * SROW + (TURB ? 2 * COLS : 0);
    static constexpr int SCRATCH_PAD = (SCRATCH * (int)sizeof(T) + 15) / 16 * 16 / (int)sizeof(T);   // stages start 16-byte aligned
    static constexpr size_t SMEM(int nw) { return (size_t)nw * (SCRATCH_PAD + D * STAGE) * sizeof(T); }
};

template <int BYTES>
__device__ __forceinline__ void cp_async(unsigned dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Raw loads of one sub-step-1 row, kept in registers from the iteration before it is computed.
//   V == 1: c[k][0] is the pulled value itself (the x-shifted populations are fetched at x-1 / x+1 directly);
//   V  > 1: c[k][.] is the aligned vector at this lane's columns; the element beyond the vector comes from the
//           neighbouring lane by shuffle at assembly time, and e[.] holds what lane 0 (left: k = 1,5,8) and
//           lane 31 (right: k = 3,6,7) fetch from the neighbouring strip's column.
template <typename T, int V, bool TURB> struct MarchRow {
    T c[9][V];
    T e[V > 1 ? 3 : 1];
    T pi[TURB ? V : 1], rp[TURB ? V : 1];
};

template <typename T>
struct MarchCtx {
    const T* src;          // cavity base of the source buffer
    T* dst;
    const T* g2top;        // second ghost rows (y-strips), [3][pitch] each
    const T* g2bot;
    long long P;           // plane stride (elements)
    int pitch, nx, ny, y0, nyl;
    int x;                 // first column of this lane (clamped to 0 for lanes beyond nx)
    int om, op;            // offsets of the x-1 / x+V element (0 where that column does not exist)
};

// Issue the loads of sub-step-1 row r (local row index, -1 .. nyl) for this lane: general form (any row, any strip).
template <typename T, int V, bool TURB>
__device__ __forceinline__ void march_load(const MarchCtx<T>& c, int r, int lane, const T* pi_in, const T* rp_in,
                                           MarchRow<T, V, TURB>& o) {
    const int gy = c.y0 + r;
    const bool lid = gy == 0, bot = gy == c.ny - 1;
    const T* ctr = c.src + (long long)(r + 1) * c.pitch;           // stored row r+1 = local row r, population 0
    // rows the y-moving populations are pulled from; a row that does not exist is replaced by the centre row (the
    // values are discarded by the wall path), the row beyond a ghost row by the second ghost rows
    const T *p2, *p5, *p6, *p4, *p7, *p8;
    if (bot) { p2 = ctr + 2 * c.P; p5 = ctr + 5 * c.P; p6 = ctr + 6 * c.P; }
    else if (r == c.nyl) { p2 = c.g2bot; p5 = c.g2bot + c.pitch; p6 = c.g2bot + 2 * c.pitch; }
    else { p2 = ctr + c.pitch + 2 * c.P; p5 = ctr + c.pitch + 5 * c.P; p6 = ctr + c.pitch + 6 * c.P; }
    if (lid) { p4 = ctr + 4 * c.P; p7 = ctr + 7 * c.P; p8 = ctr + 8 * c.P; }
    else if (r == -1) { p4 = c.g2top; p7 = c.g2top + c.pitch; p8 = c.g2top + 2 * c.pitch; }
    else { p4 = ctr - c.pitch + 4 * c.P; p7 = ctr - c.pitch + 7 * c.P; p8 = ctr - c.pitch + 8 * c.P; }
    const T* p1 = ctr + c.P;
    const T* p3 = ctr + 3 * c.P;
    if (V == 1) {
        o.c[0][0] = ctr[c.x];
        o.c[1][0] = p1[c.x + c.om];
        o.c[3][0] = p3[c.x + c.op];
        o.c[2][0] = p2[c.x];
        o.c[5][0] = p5[c.x + c.om];
        o.c[6][0] = p6[c.x + c.op];
        o.c[4][0] = p4[c.x];
        o.c[7][0] = p7[c.x + c.op];
        o.c[8][0] = p8[c.x + c.om];
    } else {
        mload<T, V>(ctr + c.x, o.c[0]);
        mload<T, V>(p1 + c.x, o.c[1]);
        mload<T, V>(p2 + c.x, o.c[2]);
        mload<T, V>(p3 + c.x, o.c[3]);
        mload<T, V>(p4 + c.x, o.c[4]);
        mload<T, V>(p5 + c.x, o.c[5]);
        mload<T, V>(p6 + c.x, o.c[6]);
        mload<T, V>(p7 + c.x, o.c[7]);
        mload<T, V>(p8 + c.x, o.c[8]);
        if (lane == 0) {            // column x-1 of the strip's first column
            o.e[0] = p1[c.x + c.om]; o.e[1] = p5[c.x + c.om]; o.e[2] = p8[c.x + c.om];
        } else if (lane == 31) {    // column x+V beyond the strip's last column
            o.e[0] = p3[c.x + c.op]; o.e[1] = p6[c.x + c.op]; o.e[2] = p7[c.x + c.op];
        }
    }
    if (TURB) {
        const long long m = (long long)r * c.pitch + c.x;            // r in [0, nyl): whole cavities only
        mload<T, V>(pi_in + m, o.pi);
        mload<T, V>(rp_in + m, o.rp);
    }
}

// The same for an interior row of a strip that touches no side wall (every source exists, all offsets are
// loop-invariant), asynchronously into one stage of the ring: pc = this lane's element of population 0 in row r,
// up / dn = -pitch / +pitch, st = shared-memory byte address of the stage.
template <typename T, int V, bool TURB, int D>
__device__ __forceinline__ void march_issue(const T* pc, long long P, long long up, long long dn, int lane, const T* pi_row,
                                            const T* rp_row, unsigned st) {
    using Cfg = MarchCfg<T, V, TURB, D>;
    constexpr int E = (int)sizeof(T);
    constexpr int RB = Cfg::SROW * E;                                  // bytes per population row of the stage
    if (V == 1) {
        const unsigned d = st + lane * E;
        cp_async<E>(d, pc);
        cp_async<E>(d + 1 * RB, pc + P - 1);
        cp_async<E>(d + 2 * RB, pc + 2 * P + dn);
        cp_async<E>(d + 3 * RB, pc + 3 * P + 1);
        cp_async<E>(d + 4 * RB, pc + 4 * P + up);
        cp_async<E>(d + 5 * RB, pc + 5 * P + dn - 1);
        cp_async<E>(d + 6 * RB, pc + 6 * P + dn + 1);
        cp_async<E>(d + 7 * RB, pc + 7 * P + up + 1);
        cp_async<E>(d + 8 * RB, pc + 8 * P + up - 1);
    } else {
        constexpr int VB = V * E;
        const unsigned d = st + (Cfg::A + lane * V) * E;
        cp_async<VB>(d, pc);
        cp_async<VB>(d + 1 * RB, pc + P);
        cp_async<VB>(d + 2 * RB, pc + 2 * P + dn);
        cp_async<VB>(d + 3 * RB, pc + 3 * P);
        cp_async<VB>(d + 4 * RB, pc + 4 * P + up);
        cp_async<VB>(d + 5 * RB, pc + 5 * P + dn);
        cp_async<VB>(d + 6 * RB, pc + 6 * P + dn);
        cp_async<VB>(d + 7 * RB, pc + 7 * P + up);
        cp_async<VB>(d + 8 * RB, pc + 8 * P + up);
        if (lane == 0) {                // column x0-1 -> index A-1
            cp_async<E>(d - E + 1 * RB, pc + P - 1);
            cp_async<E>(d - E + 5 * RB, pc + 5 * P + dn - 1);
            cp_async<E>(d - E + 8 * RB, pc + 8 * P + up - 1);
        } else if (lane == 31) {        // column x0+32V -> index A+32V
            cp_async<E>(d + VB + 3 * RB, pc + 3 * P + V);
            cp_async<E>(d + VB + 6 * RB, pc + 6 * P + dn + V);
            cp_async<E>(d + VB + 7 * RB, pc + 7 * P + up + V);
        }
    }
    if (TURB) {
        constexpr int VB = V * E;
        const unsigned d = st + (9 * Cfg::SROW + lane * V) * E;
        cp_async<VB>(d, pi_row);
        cp_async<VB>(d + Cfg::COLS * E, rp_row);
    }
}

// Sub-step-1 input populations of this lane's V nodes from a completed stage.
template <typename T, int V, bool TURB, int D>
__device__ __forceinline__ void march_read(const T* st, int lane, T f[V][9], T pi[], T rp[]) {
    using Cfg = MarchCfg<T, V, TURB, D>;
    if (V == 1) {
#pragma unroll
        for (int k = 0; k < 9; ++k) f[0][k] = st[k * Cfg::SROW + lane];
    } else {
        const T* q = st + Cfg::A + lane * V;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            T v[V];
            mload<T, V>(q + k * Cfg::SROW, v);
            const bool from_left = k == 1 || k == 5 || k == 8, from_right = k == 3 || k == 6 || k == 7;
            if (from_left) {
                f[0][k] = q[k * Cfg::SROW - 1];
#pragma unroll
                for (int i = 1; i < V; ++i) f[i][k] = v[i - 1];
            } else if (from_right) {
#pragma unroll
                for (int i = 0; i + 1 < V; ++i) f[i][k] = v[i + 1];
                f[V - 1][k] = q[k * Cfg::SROW + V];
            } else {
#pragma unroll
                for (int i = 0; i < V; ++i) f[i][k] = v[i];
            }
        }
    }
    if (TURB) {
        mload<T, V>(st + 9 * Cfg::SROW + lane * V, pi);
        mload<T, V>(st + 9 * Cfg::SROW + Cfg::COLS + lane * V, rp);
    }
}

// Sub-step-1 input populations of this lane's V nodes from the raw loads.
template <typename T, int V, bool TURB>
__device__ __forceinline__ void march_assemble(const MarchRow<T, V, TURB>& in, int lane, T f[V][9]) {
    if (V == 1) {
#pragma unroll
        for (int k = 0; k < 9; ++k) f[0][k] = in.c[k][0];
        return;
    }
    const unsigned full = 0xffffffffu;
    T l1 = __shfl_up_sync(full, in.c[1][V - 1], 1), l5 = __shfl_up_sync(full, in.c[5][V - 1], 1),
      l8 = __shfl_up_sync(full, in.c[8][V - 1], 1);
    T h3 = __shfl_down_sync(full, in.c[3][0], 1), h6 = __shfl_down_sync(full, in.c[6][0], 1),
      h7 = __shfl_down_sync(full, in.c[7][0], 1);
    if (lane == 0) { l1 = in.e[0]; l5 = in.e[1]; l8 = in.e[2]; }
    if (lane == 31) { h3 = in.e[0]; h6 = in.e[1]; h7 = in.e[2]; }
#pragma unroll
    for (int v = 0; v < V; ++v) {
        f[v][0] = in.c[0][v];
        f[v][2] = in.c[2][v];
        f[v][4] = in.c[4][v];
        f[v][1] = v == 0 ? l1 : in.c[1][v == 0 ? 0 : v - 1];
        f[v][5] = v == 0 ? l5 : in.c[5][v == 0 ? 0 : v - 1];
        f[v][8] = v == 0 ? l8 : in.c[8][v == 0 ? 0 : v - 1];
        f[v][3] = v == V - 1 ? h3 : in.c[3][v == V - 1 ? v : v + 1];
        f[v][6] = v == V - 1 ? h6 : in.c[6][v == V - 1 ? v : v + 1];
        f[v][7] = v == V - 1 ? h7 : in.c[7][v == V - 1 ? v : v + 1];
    }
}

// Rolling window between the sub-steps (per node): what sub-step 2 of the coming rows pulls from rows already advanced.
template <typename T, int V, bool TURB> struct MarchWindow {
    T a0[V], a1[V], a3[V];        // k = 0,1,3 of the row sub-step 2 handles next (already shifted in x)
    T bn4[V], bn7[V], bn8[V];     // k = 4,7,8 of that same row   -> pulled by the row after it
    T bo4[V], bo7[V], bo8[V];     // k = 4,7,8 of the row above it -> pulled by it
    T pi[TURB ? V : 1], rp[TURB ? V : 1];   // Smagorinsky state t of the row sub-step 2 handles next
};

// One sub-step on this lane's V nodes of one row.  WALL = false: interior nodes only (no predicates).
// flags: lid / bot are row properties, left / right are derived per node from its column.
template <typename T, int COLL, bool TURB, bool NEED_U, bool WALL, int V>
__device__ __forceinline__ void march_nodes(T f[V][9], const Rates<T>& rt, int x0, int nx, bool lid, bool bot,
                                            const T* rl_row, const T* carry_in, T* carry_keep, const T pi_old[],
                                            const T rp_old[], T rho[V], T ux[V], T uy[V], T pi_new[]) {
#pragma unroll
    for (int v = 0; v < V; ++v) {
        bool left = false, right = false;
        if (WALL) {
            const int xv = x0 + v;
            left = xv == 0; right = xv == nx - 1;
            if (left) { f[v][1] = (T)0; f[v][5] = (T)0; f[v][8] = (T)0; }
            if (right) { f[v][3] = (T)0; f[v][6] = (T)0; f[v][7] = (T)0; }
            if (bot) { f[v][2] = (T)0; f[v][5] = (T)0; f[v][6] = (T)0; }
            if (lid) { f[v][4] = (T)0; f[v][7] = (T)0; f[v][8] = (T)0; }
            if ((left || right || lid || bot) && xv < nx) {
                const int slot = corner_slot(left, right, lid, bot);
                const T stale = slot >= 0 ? carry_in[slot] : (T)0;
                const T rl = lid ? rl_row[v] : (T)1;
                wall_rule<T>(f[v], left, right, lid, bot, rl, rt.uLB, stale);
                if (slot >= 0) carry_keep[slot] = corner_value<T>(f[v], slot);
            }
        }
        if (TURB) {
            const T om = smagorinsky_omega<T>(f[v], pi_old[v], rp_old[v], rt.tau0);
            node_update<T, COLL, NEED_U, true>(f[v], rt, left, right, WALL && lid, WALL && bot, rho[v], ux[v], uy[v], om,
                                               &pi_new[v]);
        } else {
            node_update<T, COLL, NEED_U>(f[v], rt, left, right, WALL && lid, WALL && bot, rho[v], ux[v], uy[v]);
        }
    }
}

template <typename T, int COLL, bool TURB, bool MACROS, int V, int NW, int MINB, int D>
__global__ void __launch_bounds__(NW * 32, MINB) lbm_step_march2(const StepArgs a) {
    using Cfg = MarchCfg<T, V, TURB, D>;
    extern __shared__ __align__(16) unsigned char march_smem[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int sx = blockIdx.x * NW + w;                            // x-strip of this warp
    const int b = blockIdx.z;
    if (sx >= a.nsx) return;
    if (a.active && !a.active[b]) return;                          // frozen (converged) cavity
    T* ring = reinterpret_cast<T*>(march_smem) + (size_t)w * (Cfg::SCRATCH_PAD + D * Cfg::STAGE);   // [side][j][16]
    T* rl1 = ring + 6 * Cfg::RING_ROWS;                            // lid density after sub-step 1
    T* c1 = rl1 + Cfg::COLS;                                       // corner carries after sub-step 1
    T* stages = ring + Cfg::SCRATCH_PAD;                           // D stages of the cp.async ring
    const unsigned stages_sa = (unsigned)__cvta_generic_to_shared(stages);
    const int ya = a.row_begin + blockIdx.y * a.seg_h;             // segment [ya, yb) of local rows
    const int row_end = a.row_begin + a.row_count;
    const int yb = ya + a.seg_h < row_end ? ya + a.seg_h : row_end;
    const int xw = sx * Cfg::COLS;
    const int x = xw + lane * V;                                   // first node of this lane
    const bool act = x < a.nx;
    const Rates<T> rt(a.cav[b]);
    MarchCtx<T> c;
    c.src = static_cast<const T*>(a.src) + (long long)b * a.cavity;
    c.dst = static_cast<T*>(a.dst) + (long long)b * a.cavity;
    c.g2top = static_cast<const T*>(a.ghost2) + (long long)b * 6 * a.pitch;
    c.g2bot = c.g2top + 3 * a.pitch;
    c.P = a.plane; c.pitch = a.pitch; c.nx = a.nx; c.ny = a.ny; c.y0 = a.y0; c.nyl = a.nyl;
    c.x = act ? x : 0;
    c.om = (act && x > 0) ? -1 : 0;
    c.op = (act && x + V < a.nx) ? V : 0;
    const bool xwall = sx == 0 || xw + Cfg::COLS >= a.nx;          // strip touches the left / right wall
    const T* carry_in = static_cast<const T*>(a.carry) + b * 4;
    T* carry_out = static_cast<T*>(a.carry_out) + b * 4;
    const T* rl_in = static_cast<const T*>(a.rho_lid) + (long long)b * a.pitch;
    T* rl_out = static_cast<T*>(a.rho_lid_out) + (long long)b * a.pitch;
    const T* pi_in = TURB ? static_cast<const T*>(a.pi_eq) + (long long)b * a.mplane : nullptr;
    const T* rp_in = TURB ? static_cast<const T*>(a.rho_prev) + (long long)b * a.mplane : nullptr;
    const unsigned full = 0xffffffffu;

    // ---- sub-step 1 of the two ring columns for rows [r0, r0+16): lane = (side, row) ------------------------------
    auto ring_pass = [&](int r0) {
        const int side = lane >> 4, i = lane & 15;
        const int r = r0 + i;
        const int xr = side ? xw + Cfg::COLS : xw - 1;
        const int gy = a.y0 + r;
        __syncwarp();
        if (r <= yb && gy >= 0 && gy < a.ny && xr >= 0 && xr < a.nx) {
            const bool left = false, right = xr == a.nx - 1, lid = gy == 0, bot = gy == a.ny - 1;
            const long long rc = (long long)(r + 1) * a.pitch + xr, ru = rc - a.pitch, rd = rc + a.pitch;
            const bool up2 = r == -1, dn2 = r == a.nyl;             // ghost rows of a y-strip pull from the second ghost rows
            const long long P = a.plane;
            T f[9];
            f[0] = c.src[rc];
            f[1] = c.src[1 * P + rc - 1];
            f[3] = right ? (T)0 : c.src[3 * P + rc + 1];
            f[2] = bot ? (T)0 : (dn2 ? c.g2bot[xr] : c.src[2 * P + rd]);
            f[5] = bot ? (T)0 : (dn2 ? c.g2bot[a.pitch + xr - 1] : c.src[5 * P + rd - 1]);
            f[6] = (right || bot) ? (T)0 : (dn2 ? c.g2bot[2 * a.pitch + xr + 1] : c.src[6 * P + rd + 1]);
            f[4] = lid ? (T)0 : (up2 ? c.g2top[xr] : c.src[4 * P + ru]);
            f[7] = (right || lid) ? (T)0 : (up2 ? c.g2top[a.pitch + xr + 1] : c.src[7 * P + ru + 1]);
            f[8] = lid ? (T)0 : (up2 ? c.g2top[2 * a.pitch + xr - 1] : c.src[8 * P + ru - 1]);
            if (right || lid || bot) {
                const int slot = corner_slot(left, right, lid, bot);
                const T stale = slot >= 0 ? carry_in[slot] : (T)0;
                const T rl = lid ? rl_in[xr] : (T)1;
                wall_rule<T>(f, left, right, lid, bot, rl, rt.uLB, stale);
            }
            T rho, ux, uy;
            if (TURB) {
                const long long m = (long long)r * a.pitch + xr;
                T pn;
                const T om = smagorinsky_omega<T>(f, pi_in[m], rp_in[m], rt.tau0);
                node_update<T, COLL, false, true>(f, rt, left, right, lid, bot, rho, ux, uy, om, &pn);
            } else {
                node_update<T, COLL, false>(f, rt, left, right, lid, bot, rho, ux, uy);
            }
            T* q = ring + side * 3 * Cfg::RING_ROWS + i;
            q[0] = side ? f[3] : f[1];
            q[Cfg::RING_ROWS] = side ? f[6] : f[5];
            q[2 * Cfg::RING_ROWS] = side ? f[7] : f[8];
        }
        __syncwarp();
    };

    // a row whose nine sources all exist in the main buffer, in a strip without wall columns: staged through the ring
    auto is_fast = [&](int r) {
        const int gy = a.y0 + r;
        return !xwall && gy > 0 && gy < a.ny - 1 && r >= 0 && r < a.nyl;
    };
    // start the copies of row r into stage `st` (rows outside the segment / general rows: an empty group)
    auto issue = [&](int r, int st) {
        if (r <= yb && is_fast(r)) {
            const long long ro = (long long)(r + 1) * a.pitch + x;      // stored row r+1 = local row r
            const long long mo = TURB ? (long long)r * a.pitch + x : 0;
            march_issue<T, V, TURB, D>(c.src + ro, a.plane, -(long long)a.pitch, (long long)a.pitch, lane, pi_in + mo, rp_in + mo,
                                       stages_sa + (unsigned)(st * Cfg::STAGE * (int)sizeof(T)));
        }
        cp_async_commit();
    };

    MarchWindow<T, V, TURB> win;
    T c2[V], c5[V], c6[V];                                         // k = 2,5,6 of the row just advanced (pulled by the row above it)
#pragma unroll
    for (int v = 0; v < V; ++v) {
        win.a0[v] = win.a1[v] = win.a3[v] = (T)0;
        win.bn4[v] = win.bn7[v] = win.bn8[v] = win.bo4[v] = win.bo7[v] = win.bo8[v] = (T)0;
        c2[v] = c5[v] = c6[v] = (T)0;
    }
    const int r_first = ya - 1;
#pragma unroll
    for (int d = 0; d < D; ++d) issue(r_first + d, d);
    ring_pass(r_first);

    int st = 0;
    for (int r = r_first; r <= yb; ++r) {                          // r: row of sub-step 1; sub-step 2 handles row r-1
        const int gy = a.y0 + r;
        const bool exists = gy >= 0 && gy < a.ny;
        T f[V][9];
        T pi0[TURB ? V : 1], rp0[TURB ? V : 1];                    // Smagorinsky state t-1 of row r
        T n0[V], n1[V], n3[V], n4[V], n7[V], n8[V];               // row r's populations for the window
        T pi1[TURB ? V : 1], rho1[V];
        cp_async_wait<D - 1>();                                    // the group of row r has landed
        if (exists) {
            if (is_fast(r)) {
                if (V > 1) __syncwarp();                           // the neighbours' copies are visible
                march_read<T, V, TURB, D>(stages + st * Cfg::STAGE, lane, f, pi0, rp0);
                if (V > 1) __syncwarp();                           // everybody has read: the stage may be refilled
            } else {
                MarchRow<T, V, TURB> raw;
                march_load<T, V, TURB>(c, r, lane, pi_in, rp_in, raw);
                march_assemble<T, V, TURB>(raw, lane, f);
                if (TURB) {
#pragma unroll
                    for (int v = 0; v < V; ++v) { pi0[v] = raw.pi[v]; rp0[v] = raw.rp[v]; }
                }
            }
        }
        issue(r + D, st);
        st = st + 1 == D ? 0 : st + 1;
        if (exists) {
            // ---- sub-step 1 on row r ----
            const bool lid = gy == 0, bot = gy == a.ny - 1;
            T ux[V], uy[V];
            if (xwall || lid || bot) {
                T rl[V];
                if (lid) {
#pragma unroll
                    for (int v = 0; v < V; ++v) rl[v] = rl_in[c.x + v];
                }
                march_nodes<T, COLL, TURB, false, true, V>(f, rt, x, a.nx, lid, bot, rl, carry_in, c1, pi0, rp0, rho1, ux, uy, pi1);
                if (lid) {
#pragma unroll
                    for (int v = 0; v < V; ++v) rl1[lane * V + v] = rho1[v];
                }
            } else {
                march_nodes<T, COLL, TURB, false, false, V>(f, rt, x, a.nx, false, false, nullptr, nullptr, nullptr, pi0, rp0,
                                                             rho1, ux, uy, pi1);
            }
            // ---- hand the post-collision populations to the nodes that pull them in sub-step 2 ----
            const int ri = (r - r_first) & (Cfg::RING_ROWS - 1);
            T u1 = __shfl_up_sync(full, f[V - 1][1], 1), u5 = __shfl_up_sync(full, f[V - 1][5], 1),
              u8 = __shfl_up_sync(full, f[V - 1][8], 1);
            T d3 = __shfl_down_sync(full, f[0][3], 1), d6 = __shfl_down_sync(full, f[0][6], 1),
              d7 = __shfl_down_sync(full, f[0][7], 1);
            if (lane == 0) { u1 = ring[ri]; u5 = ring[Cfg::RING_ROWS + ri]; u8 = ring[2 * Cfg::RING_ROWS + ri]; }
            if (lane == 31) {
                d3 = ring[3 * Cfg::RING_ROWS + ri]; d6 = ring[4 * Cfg::RING_ROWS + ri]; d7 = ring[5 * Cfg::RING_ROWS + ri];
            }
#pragma unroll
            for (int v = 0; v < V; ++v) {
                c2[v] = f[v][2];
                c5[v] = v == 0 ? u5 : f[v == 0 ? 0 : v - 1][5];
                c6[v] = v == V - 1 ? d6 : f[v == V - 1 ? v : v + 1][6];
                // k = 0,1,3 / 4,7,8 enter the window below, after sub-step 2 has consumed the old values
                n0[v] = f[v][0];
                n1[v] = v == 0 ? u1 : f[v == 0 ? 0 : v - 1][1];
                n3[v] = v == V - 1 ? d3 : f[v == V - 1 ? v : v + 1][3];
                n4[v] = f[v][4];
                n8[v] = v == 0 ? u8 : f[v == 0 ? 0 : v - 1][8];
                n7[v] = v == V - 1 ? d7 : f[v == V - 1 ? v : v + 1][7];
            }
        }
        if (r - 1 >= ya) {
            // ---- sub-step 2 on row y = r-1 ----
            const int y = r - 1, gy2 = gy - 1;
            const bool lid = gy2 == 0, bot = gy2 == a.ny - 1;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                f[v][0] = win.a0[v]; f[v][1] = win.a1[v]; f[v][3] = win.a3[v];
                f[v][2] = c2[v]; f[v][5] = c5[v]; f[v][6] = c6[v];
                f[v][4] = win.bo4[v]; f[v][7] = win.bo7[v]; f[v][8] = win.bo8[v];
            }
            T rho[V], ux[V], uy[V], pi2[TURB ? V : 1];
            if (xwall || lid || bot) {
                T rl[V];
                if (lid) {
#pragma unroll
                    for (int v = 0; v < V; ++v) rl[v] = rl1[lane * V + v];
                }
                march_nodes<T, COLL, TURB, MACROS, true, V>(f, rt, x, a.nx, lid, bot, rl, c1, carry_out, win.pi, win.rp, rho, ux,
                                                             uy, pi2);
                if (lid) {
#pragma unroll
                    for (int v = 0; v < V; ++v)
                        if (x + v < a.nx) rl_out[x + v] = rho[v];
                }
            } else {
                march_nodes<T, COLL, TURB, MACROS, false, V>(f, rt, x, a.nx, false, false, nullptr, nullptr, nullptr, win.pi,
                                                              win.rp, rho, ux, uy, pi2);
            }
            if (act) {
                T* drow = c.dst + (long long)(y + 1) * a.pitch + x;
                const long long m = (long long)b * a.mplane + (long long)y * a.pitch + x;
                if (x + V <= a.nx) {
#pragma unroll
                    for (int k = 0; k < 9; ++k) {
                        T tmp[V];
#pragma unroll
                        for (int v = 0; v < V; ++v) tmp[v] = f[v][k];
                        mstore<T, V>(drow + k * c.P, tmp);
                    }
                    if (TURB) {
                        mstore<T, V>(static_cast<T*>(a.pi_eq_out) + m, pi2);
                        mstore<T, V>(static_cast<T*>(a.rho_prev_out) + m, rho);
                    }
                    if (MACROS) {
                        mstore<T, V>(static_cast<T*>(a.rho) + m, rho);
                        mstore<T, V>(static_cast<T*>(a.ux) + m, ux);
                        mstore<T, V>(static_cast<T*>(a.uy) + m, uy);
                    }
                } else {
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        if (x + v < a.nx) {
#pragma unroll
                            for (int k = 0; k < 9; ++k) drow[k * c.P + v] = f[v][k];
                            if (TURB) {
                                static_cast<T*>(a.pi_eq_out)[m + v] = pi2[v];
                                static_cast<T*>(a.rho_prev_out)[m + v] = rho[v];
                            }
                            if (MACROS) {
                                static_cast<T*>(a.rho)[m + v] = rho[v];
                                static_cast<T*>(a.ux)[m + v] = ux[v];
                                static_cast<T*>(a.uy)[m + v] = uy[v];
                            }
                        }
                    }
                }
            }
        }
        // ---- rotate the window: row r becomes "the row sub-step 2 handles next" ----
        if (exists) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                win.bo4[v] = win.bn4[v]; win.bo7[v] = win.bn7[v]; win.bo8[v] = win.bn8[v];
                win.a0[v] = n0[v]; win.a1[v] = n1[v]; win.a3[v] = n3[v];
                win.bn4[v] = n4[v]; win.bn7[v] = n7[v]; win.bn8[v] = n8[v];
                if (TURB) { win.pi[v] = pi1[v]; win.rp[v] = rho1[v]; }
            }
        }
        if (r + 1 <= yb && ((r + 1 - r_first) & (Cfg::RING_ROWS - 1)) == 0) ring_pass(r + 1);
    }
    cp_async_wait<0>();
}

}  // namespace lbm
