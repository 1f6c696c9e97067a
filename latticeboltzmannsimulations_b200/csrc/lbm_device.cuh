// Per-node arithmetic of the fused D2Q9 step (device code shared by every kernel family).
//
// Semantics "C" of the reference, executable spec = oracle/lbm_oracle.py step_C_pull:
//   gather (pull) + wall rule      <- MRT_GPU.py:654-656 (push) + funBC :664-692
//   moments + wall/lid overrides   <- MRT_GPU.py:615-631
//   collision SRT / TRT / MRT      <- MRT_GPU.py:413 / :455-462,514-527 / :633-648,655
// Index convention: y == 0 is the lid; a population with c_y = +1 moves to y-1 (MRT_GPU.py:413 "j-c").
//   k : 0      1      2      3       4       5      6       7        8
//   c : (0,0)  (1,0)  (0,1)  (-1,0)  (0,-1)  (1,1)  (-1,1)  (-1,-1)  (1,-1)      (MRT.py:138)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lbm {

enum { COLL_SRT = 0, COLL_TRT = 1, COLL_MRT = 2 };
enum { MODE_STEP = 0, MODE_FINALIZE = 1, MODE_MACROS = 2 };

// Per-cavity rates, kept in double on the device and converted by each thread (uniform, cached loads).
struct CavityParams {
    double uLB;
    double omega;      // SRT omega / TRT omega+ / MRT omega_nu   (MRT_GPU.py:65, 82, 88)
    double omegam;     // TRT omega-                              (MRT_GPU.py:84)
    double s_e, s_eps, s_q;                                    // (MRT_GPU.py:89-91)
    double tau0;       // 1 / omega, the molecular relaxation time entering the Smagorinsky closure (MRT_GPU.py:553)
    double pad;
};

// Kernel arguments shared by every kernel family (device pointers are untyped: dtype is a template parameter).
struct StepArgs {
    const void* src;
    void* dst;
    void* rho;
    void* ux;
    void* uy;
    void* rho_lid;             // [batch][pitch]
    void* carry;               // [batch][4]
    void* rho_lid_out;         // fused two-step kernel: side buffers written for the state two steps ahead (the other
    void* carry_out;           //   half of the double-buffered allocation); single-step kernels update in place
    const void* ghost2;        // fused kernel on a y-strip: [batch][top|bottom][3][pitch] second ghost rows of `src`
    void* pi_eq;               // [batch][nyl][pitch]  sum_k cx cy feq_k of the previous step (Smagorinsky only)
    void* rho_prev;            // [batch][nyl][pitch]  1 / rho of the previous step        (Smagorinsky only)
    void* pi_eq_out;           // two-step kernels: the other half of the double-buffered Smagorinsky state (one-step
    void* rho_prev_out;        //   kernels update in place)
    const CavityParams* cav;   // [batch]
    const int* active;         // [batch] or NULL: cavities with active[b] == 0 are frozen (skipped by step launches)
    int nx, ny, y0, nyl, pitch;
    long long plane, cavity;   // elements
    long long mplane;          // macro plane = nyl * pitch elements
    int row_begin, row_stride; // local row of launch row 0 and distance between consecutive launch rows
    int row_count;             // launch rows (blockIdx.y * blockDim.y + threadIdx.y < row_count)
    int seg_h;                 // sliding-window two-step kernel: rows per segment (one CTA each)
    int seg_stride;            //   0: segments follow each other; else exactly two segments whose first rows are this far
                               //   apart (the two edge bands of a y-strip in one launch)
    int slide_tma;             //   1: interior blocks are staged by tensor copies (one box per population), 0: row copies
};

template <typename T>
struct Rates {
    T uLB, omega, omegam, s_e, s_eps, s_q, tau0;
    // MRT rates with the constant factors of the inverse moment transform folded in once per thread (in double):
    // q_e = s_e / 36, q_eps = s_eps / 36, q_q = s_q / 12 -- what is left in the collision are powers of two
    T q_e, q_eps, q_q;
    __device__ __forceinline__ explicit Rates(const CavityParams& p)
        : uLB((T)p.uLB), omega((T)p.omega), omegam((T)p.omegam), s_e((T)p.s_e), s_eps((T)p.s_eps), s_q((T)p.s_q),
          tau0((T)p.tau0), q_e((T)(p.s_e / 36.0)), q_eps((T)(p.s_eps / 36.0)), q_q((T)(p.s_q / 12.0)) {}
};

// Lattice weights t_k (MRT.py:144-146).
template <typename T> __device__ __forceinline__ T w_rest() { return (T)(4.0 / 9.0); }
template <typename T> __device__ __forceinline__ T w_axis() { return (T)(1.0 / 9.0); }
template <typename T> __device__ __forceinline__ T w_diag() { return (T)(1.0 / 36.0); }

// ---- rounding discipline ----------------------------------------------------------------------------------------
// The library is compiled with -fmad=false: the compiler never contracts a*b+c on its own, every fused multiply-add
// below is spelled fm(a, b, c).  What a node computes is therefore fixed by this file alone and does not depend on
// which kernel inlines it -- the one-step, shared-memory-tile and sliding-window kernels are bit-identical by construction
// (with implicit contraction the fp64 SRT / TRT paths of two kernel families were seen to differ in the last bit).
__device__ __forceinline__ float fm(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fm(double a, double b, double c) { return __fma_rn(a, b, c); }

// ---- two fp32 nodes per thread: Blackwell packed arithmetic ---------------------------------------------------------
// sm_100 adds packed fp32 instructions (FADD2 / FMUL2 / FFMA2: two IEEE round-to-nearest results per lane and
// instruction).  f32x2 runs the SAME templated node arithmetic below on two x-adjacent nodes at once: every operator
// maps to one packed instruction whose two halves are exactly what the scalar code computes, so a kernel may mix
// packed and scalar nodes and stay bit-identical (a - b is fma(b, -1, a): b * -1 is exact; -a is a * -1).
struct f32x2 {
    float2 v;
    __device__ __forceinline__ f32x2() {}
    __device__ __forceinline__ explicit f32x2(float a) : v(make_float2(a, a)) {}
    __device__ __forceinline__ explicit f32x2(double a) : v(make_float2((float)a, (float)a)) {}
    __device__ __forceinline__ explicit f32x2(int a) : v(make_float2((float)a, (float)a)) {}
    __device__ __forceinline__ f32x2(float a, float b) : v(make_float2(a, b)) {}
    __device__ __forceinline__ explicit f32x2(float2 a) : v(a) {}
};
__device__ __forceinline__ f32x2 operator+(f32x2 a, f32x2 b) { return f32x2(__fadd2_rn(a.v, b.v)); }
__device__ __forceinline__ f32x2 operator*(f32x2 a, f32x2 b) { return f32x2(__fmul2_rn(a.v, b.v)); }
__device__ __forceinline__ f32x2 operator-(f32x2 a, f32x2 b) { return f32x2(__ffma2_rn(b.v, make_float2(-1.0f, -1.0f), a.v)); }
__device__ __forceinline__ f32x2 operator-(f32x2 a) { return f32x2(__fmul2_rn(a.v, make_float2(-1.0f, -1.0f))); }
__device__ __forceinline__ f32x2 operator/(f32x2 a, f32x2 b) { return f32x2(a.v.x / b.v.x, a.v.y / b.v.y); }
__device__ __forceinline__ f32x2 fm(f32x2 a, f32x2 b, f32x2 c) { return f32x2(__ffma2_rn(a.v, b.v, c.v)); }
// CAUTION (ptxas 12.9, seen in SASS): a packed product with a single use that feeds a packed ADDITION is contracted
// into one FFMA2 although both PTX instructions carry .rn and the build passes --fmad false (the scalar pair is left
// alone) -- the two halves then no longer equal the scalar code.  Subtractions (an FFMA2 with -1 here) and explicit
// fm() are not touched.  So in code shared with f32x2: never write `product + x` with a product used only there; use
// fm(), or arrange the expression as a difference.  tools/packed_check.cu compares the packed and the scalar node
// arithmetic bit for bit on random states (GPU test test_packed_arithmetic_equals_scalar).
using ::sqrt;       // keep the scalar overloads visible next to the packed ones
using ::fabs;
__device__ __forceinline__ f32x2 sqrt(f32x2 a) { return f32x2(sqrtf(a.v.x), sqrtf(a.v.y)); }
__device__ __forceinline__ f32x2 fabs(f32x2 a) { return f32x2(fabsf(a.v.x), fabsf(a.v.y)); }
// fp32 takes the deviation forms of SRT / TRT (see below); the packed type is fp32
template <typename T> struct is_fp32 { static constexpr bool value = sizeof(T) == 4; };
template <> struct is_fp32<f32x2> { static constexpr bool value = true; };

// feq_k = rho*t_k*(1. + 3.0*cu + 9*0.5*cu*cu - 3.0*0.5*usqr)   (MRT_GPU.py:651), Horner in cu; rt = rho * t_k.
template <typename T>
__device__ __forceinline__ T feq_one(T rt, T cu, T usqr) {
    return rt * fm((T)-1.5, usqr, fm(cu, fm((T)4.5, cu, (T)3.0), (T)1.0));
}

template <typename T>
__device__ __forceinline__ void feq_all(T rho, T ux, T uy, T fe[9]) {
    const T usqr = fm(ux, ux, uy * uy);
    const T r0 = rho * w_rest<T>(), r1 = rho * w_axis<T>(), r2 = rho * w_diag<T>();
    fe[0] = r0 * fm((T)-1.5, usqr, (T)1.0);
    fe[1] = feq_one(r1, ux, usqr);
    fe[2] = feq_one(r1, uy, usqr);
    fe[3] = feq_one(r1, -ux, usqr);
    fe[4] = feq_one(r1, -uy, usqr);
    fe[5] = feq_one(r2, ux + uy, usqr);
    fe[6] = feq_one(r2, uy - ux, usqr);
    fe[7] = feq_one(r2, -(ux + uy), usqr);
    fe[8] = feq_one(r2, ux - uy, usqr);
}

// The same equilibrium in the reference's own operation order and without any fused operation (MRT.py:213-231,
// functions.pyx:229-267): bit-identical to the NumPy / Cython expression.  Used off the hot path, where the result is
// handed to the caller as data: the initial state (lbm_init_eq) and functions.equ (lbm_equ_kernel).
template <typename T>
__device__ __forceinline__ void feq_all_ref(T rho, T ux, T uy, T fe[9]) {
    const T usqr = ux * ux + uy * uy;
    const T cu[9] = {(T)0, ux, uy, -ux, -uy, ux + uy, -ux + uy, -ux - uy, ux - uy};
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const T tk = k == 0 ? w_rest<T>() : (k < 5 ? w_axis<T>() : w_diag<T>());
        fe[k] = rho * tk * ((T)1.0 + (T)3.0 * cu[k] + (T)4.5 * cu[k] * cu[k] - (T)1.5 * usqr);
    }
}

// Wall rule on the post-stream populations f of a wall node (funBC, MRT_GPU.py:674-692): x-block, then y-block,
// the y-block seeing the x-block's results at corners.  feq is that of the node's own previous-step state:
// u = 0 on the three resting walls (all differences exactly 0 -> on-node bounce-back), (uLB,0) and rho_lid on y == 0.
// `stale` is the previous final value of the doubly-orphaned corner population (persistent ftemp in the reference).
template <typename T>
__device__ __forceinline__ void wall_rule(T f[9], bool left, bool right, bool lid, bool bot,
                                          T rho_lid, T uLB, T stale) {
    T fe[9];
    if (lid) {
        feq_all<T>(rho_lid, uLB, (T)0, fe);
    } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) fe[k] = (T)0;
    }
    if (left && lid) f[7] = stale;
    if (right && lid) f[8] = stale;
    if (left && bot) f[6] = stale;
    if (right && bot) f[5] = stale;
    if (left) {
        f[1] = fe[1] - fe[3] + f[3];
        f[5] = fe[5] - fe[7] + f[7];
        f[8] = fe[8] - fe[6] + f[6];
    } else if (right) {
        f[3] = -fe[1] + fe[3] + f[1];
        f[6] = -fe[8] + fe[6] + f[8];
        f[7] = -fe[5] + fe[7] + f[5];
    }
    if (bot) {
        f[2] = -fe[4] + fe[2] + f[4];
        f[5] = -fe[7] + fe[5] + f[7];
        f[6] = -fe[8] + fe[6] + f[8];
    } else if (lid) {
        f[4] = -fe[2] + fe[4] + f[2];
        f[7] = -fe[5] + fe[7] + f[5];
        f[8] = -fe[6] + fe[8] + f[6];
    }
}

// Which corner-carry slot a node owns (-1 = none): 0 f7@(0,0), 1 f8@(nx-1,0), 2 f6@(0,ny-1), 3 f5@(nx-1,ny-1).
__device__ __forceinline__ int corner_slot(bool left, bool right, bool lid, bool bot) {
    if (lid) return left ? 0 : (right ? 1 : -1);
    if (bot) return left ? 2 : (right ? 3 : -1);
    return -1;
}
__device__ __forceinline__ int corner_pop(int slot) { return slot == 0 ? 7 : slot == 1 ? 8 : slot == 2 ? 6 : 5; }
// f[corner_pop(slot)] without dynamic indexing (keeps f[] in registers)
template <typename T>
__device__ __forceinline__ T corner_value(const T f[9], int slot) {
    return slot == 0 ? f[7] : slot == 1 ? f[8] : slot == 2 ? f[6] : f[5];
}

// rho and momentum in the reference's summation order (MRT_GPU.py:615-619).
template <typename T>
__device__ __forceinline__ void moments_ref(const T f[9], T& rho, T& jx, T& jy) {
    rho = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8];
    jx = f[1] - f[3] + f[5] - f[6] - f[7] + f[8];
    jy = f[2] - f[4] + f[5] + f[6] - f[7] - f[8];
}

// Lid density, MRT_GPU.py:627.
template <typename T>
__device__ __forceinline__ T rho_lid_formula(const T f[9]) {
    return fm((T)2, f[2] + f[5] + f[6], f[0] + f[1] + f[3]);
}

// ---- collisions ---------------------------------------------------------------------------------------------
// SRT: f - omega (f - feq)   (MRT_GPU.py:413)
template <typename T>
__device__ __forceinline__ void collide_srt(T f[9], T rho, T ux, T uy, T omega) {
    T fe[9];
    feq_all<T>(rho, ux, uy, fe);
#pragma unroll
    for (int k = 0; k < 9; ++k) f[k] = fm(-omega, f[k] - fe[k], f[k]);
}

// TRT: f - omega+ (f+ - feq+) - omega- (f- - feq-)   (MRT_GPU.py:455-462, 514-527); the halves are folded into the rates
template <typename T>
__device__ __forceinline__ void collide_trt(T f[9], T rho, T ux, T uy, T omegap, T omegam) {
    T fe[9];
    feq_all<T>(rho, ux, uy, fe);
    const T hp = (T)0.5 * omegap, hm = (T)0.5 * omegam;
    f[0] = fm(-omegap, f[0] - fe[0], f[0]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int a = (i < 2) ? i + 1 : i + 3;     // a = 1, 2, 5, 6
        const int o = a + 2;                       // 1<->3, 2<->4, 5<->7, 6<->8
        const T sp = (f[a] + f[o]) - (fe[a] + fe[o]);     // 2 (f+ - feq+)
        const T sm = (f[a] - f[o]) - (fe[a] - fe[o]);     // 2 (f- - feq-)   (opposite sign for o)
        const T fa = fm(-hm, sm, fm(-hp, sp, f[a]));
        const T fo = fm(hm, sm, fm(-hp, sp, f[o]));
        f[a] = fa;
        f[o] = fo;
    }
}

// ---- fp32 accuracy: deviation form of SRT / TRT -----------------------------------------------------------------
// In fp32 the reference-order expression f - omega (f - feq) injects ~ulp(feq) = 3e-8 of noise per step into the
// conserved density (measured: 3e-6 in rho, 1e-5 of uLB in u after 150 steps at tau = 0.54, ten times the MRT path,
// whose correction form never rounds at the magnitude of f).  The same collision evaluated on the deviations
// g_k = f_k - t_k (exact by Sterbenz) and feq_k - t_k = t_k (drho + rho (3cu + 4.5cu^2 - 1.5u^2)) only rounds at the
// magnitude of the deviations.  Used for T = float only; fp64 evaluates the populations themselves.
// (The reference's own fp32 kernels evaluate these expressions with double literals, i.e. in mixed precision.)
template <typename T>
__device__ __forceinline__ T drho_of(const T f[9], bool lid) {
    const T g0 = f[0] - w_rest<T>(), g1 = f[1] - w_axis<T>(), g2 = f[2] - w_axis<T>(), g3 = f[3] - w_axis<T>(),
            g4 = f[4] - w_axis<T>(), g5 = f[5] - w_diag<T>(), g6 = f[6] - w_diag<T>(), g7 = f[7] - w_diag<T>(),
            g8 = f[8] - w_diag<T>();
    if (lid) return fm((T)2, g2 + g5 + g6, g0 + g1 + g3);        // lid formula: the weights sum to exactly 1
    return g0 + g1 + g2 + g3 + g4 + g5 + g6 + g7 + g8;
}
template <typename T>
__device__ __forceinline__ T weight_of(int k) { return k == 0 ? w_rest<T>() : (k < 5 ? w_axis<T>() : w_diag<T>()); }
template <typename T>
__device__ __forceinline__ void feq_dev_all(T drho, T rho, T ux, T uy, T fd[9]) {
    const T usqr = fm(ux, ux, uy * uy);
    const T cu[9] = {(T)0, ux, uy, -ux, -uy, ux + uy, uy - ux, -(ux + uy), ux - uy};
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const T q = fm((T)-1.5, usqr, cu[k] * fm((T)4.5, cu[k], (T)3.0));     // 3cu + 4.5cu^2 - 1.5u^2
        fd[k] = weight_of<T>(k) * fm(rho, q, drho);
    }
}

template <typename T>
__device__ __forceinline__ void collide_srt_dev(T f[9], T drho, T rho, T ux, T uy, T omega) {
    T fd[9];
    feq_dev_all<T>(drho, rho, ux, uy, fd);
#pragma unroll
    for (int k = 0; k < 9; ++k) f[k] = fm(-omega, (f[k] - weight_of<T>(k)) - fd[k], f[k]);
}
template <typename T>
__device__ __forceinline__ void collide_trt_dev(T f[9], T drho, T rho, T ux, T uy, T omegap, T omegam) {
    T fd[9];
    feq_dev_all<T>(drho, rho, ux, uy, fd);
    const T hp = (T)0.5 * omegap, hm = (T)0.5 * omegam;
    f[0] = fm(-omegap, (f[0] - w_rest<T>()) - fd[0], f[0]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int a = (i < 2) ? i + 1 : i + 3;
        const int o = a + 2;
        const T ga = f[a] - weight_of<T>(a), go = f[o] - weight_of<T>(o);
        const T sp = (ga + go) - (fd[a] + fd[o]);      // 2 (f+ - feq+)
        const T sm = (ga - go) - (fd[a] - fd[o]);      // 2 (f- - feq-)   (opposite sign for o)
        const T fa = fm(-hm, sm, fm(-hp, sp, f[a]));
        const T fo = fm(hm, sm, fm(-hp, sp, f[o]));
        f[a] = fa;
        f[o] = fo;
    }
}

// MRT in the Gram-Schmidt basis of MRT_GPU.py:593-612 with the reference's (non-standard) equilibrium moments
// (:636-644: momentum not velocity, +9 jx^2 jy^2, cubic heat-flux terms) and rates s = [0,s_e,s_eps,0,s_q,0,s_q,s_nu,s_nu].
// Evaluated as f* = f - Minv * S * (m - m_eq) with M and Minv hand-factored (entries 0,+-1,+-2,+-4 / 1/4..1/36):
// ~80 operations instead of two dense 9x9 products.  Rounding differs from the reference's "Minv * m*" form at the
// 1e-16 level per step (measured 5e-14 in rho after 1000 steps at 96^2, Re 3200; tolerance 1e-12).
// rho_given: the density entering the equilibrium moments when it is not the plain sum of the populations (lid row);
// otherwise (use_given = false) it is summed here from the partial sums the transform needs anyway -- always, so that a
// step gives the same bits whether or not it also writes rho / u (which carry the reference-order sum).
// q_e, q_eps, q_q: the rates pre-divided by 36, 36, 12 (Rates); with s_nu / 4 every remaining factor of M^-1 is a power
// of two: c0 = (de - dp)/9 = 4 (de' - dp'), c_axis = (de + 2 dp)/36 = de' + 2 dp', c_diag = -(2 de + dp)/36, q/6 = 2 q'.
template <typename T>
__device__ __forceinline__ void collide_mrt(T f[9], bool use_given, T rho_given, T q_e, T q_eps, T q_q, T s_nu) {
    const T p13 = f[1] + f[3], m13 = f[1] - f[3];
    const T p24 = f[2] + f[4], m24 = f[2] - f[4];
    const T s57 = f[5] + f[7], d57 = f[5] - f[7];
    const T s68 = f[6] + f[8], d68 = f[6] - f[8];
    const T dg = s57 + s68;
    const T jxd = d57 - d68, jyd = d57 + d68;
    const T pxy = s57 - s68;
    const T ab = p13 + p24;
    const T rho = use_given ? rho_given : f[0] + (ab + dg);
    const T jx = m13 + jxd, jy = m24 + jyd;
    // written without a single unary minus (free for scalars, one more packed instruction for f32x2): en = -e,
    // den = -de, can = -c_axis, t = -c_diag
    const T en = fm((T)4, f[0], fm((T)-2, dg, ab));
    const T eps = fm((T)4, f[0], fm((T)-2, ab, dg));
    const T qx = fm((T)-2, m13, jxd), qy = fm((T)-2, m24, jyd);
    const T pxx = p13 - p24;
    const T jx2 = jx * jx, jy2 = jy * jy, sq = jx2 + jy2;
    const T e_eq = fm((T)3, sq, (T)-2 * rho);
    const T eps_eq = fm((T)9, jx2 * jy2, fm((T)-3, sq, rho));
    const T qx_eq = jx * fm((T)3, jx2, (T)-1);
    const T qy_eq = jy * fm((T)3, jy2, (T)-1);
    const T pxx_eq = jx2 - jy2, pxy_eq = jx * jy;
    const T q_nu = (T)0.25 * s_nu;
    const T den = q_e * (en + e_eq);                                  // -s_e (e - e_eq) / 36
    const T dp = q_eps * (eps - eps_eq);                              //  s_eps (eps - eps_eq) / 36
    const T dqx = q_q * (qx - qx_eq), dqy = q_q * (qy - qy_eq);      //  / 12
    const T dxx = q_nu * (pxx - pxx_eq), dxy = q_nu * (pxy - pxy_eq);  //  / 4
    const T can = fm((T)-2, dp, den);                                 // -(de + 2 dp) / 36
    const T t = fm((T)-2, den, dp);                                   //  (2 de + dp) / 36
    const T am = can + dxx, ap = can - dxx;
    const T tm = t + dxy, tp = t - dxy;
    const T qs = dqx + dqy, qd = dqx - dqy;
    f[0] = fm((T)-4, den + dp, f[0]);                                 // + (de - dp) / 9
    f[1] = f[1] - fm((T)-2, dqx, am);
    f[3] = f[3] - fm((T)2, dqx, am);
    f[2] = f[2] - fm((T)-2, dqy, ap);
    f[4] = f[4] - fm((T)2, dqy, ap);
    f[5] = f[5] - (tm + qs);
    f[7] = f[7] - (tm - qs);
    f[6] = f[6] - (tp - qd);
    f[8] = f[8] - (tp + qd);
}

// Smagorinsky closure of MRT_GPU.py:570-589: effective relaxation rate from the non-equilibrium momentum flux
// Q = sum_k cx cy (f_k - feq_k^prev), with feq and rho of the PREVIOUS step (the reference reads feq_g / rho_g before
// overwriting them) and Cs2 hard-set to 0.025 (:578; the Van-Driest lines above it are dead code).
template <typename T>
__device__ __forceinline__ T smagorinsky_omega(const T f[9], T pi_prev, T irho_prev, T tau0) {
    // irho_prev = 1 / rho of the previous step: node_update() has that reciprocal anyway (u = j / rho) and hands it
    // out, so the closure costs one division (the final 1 / tau, written 2 / (tau0 + sqrt(..))) instead of two
    const T product1 = f[5] - f[6] + f[7] - f[8];
    const T Qmf = product1 - pi_prev;
    return (T)2.0 / (tau0 + sqrt(fm(tau0, tau0, ((T)(18 * 1.4142 * 0.025) * fabs(Qmf)) * irho_prev)));
}

// Everything after the gather for one node: overrides, optional macro output values, collision in place.
// Returns rho (lid-overridden) and the output velocity through the references; with TURB also the closure's state for
// the next step: sum_k cx cy feq_k and 1 / rho.
template <typename T, int COLL, bool NEED_U, bool TURB = false>
__device__ __forceinline__ void node_update(T f[9], const Rates<T>& r, bool left, bool right, bool lid, bool bot,
                                            T& rho_out, T& ux_out, T& uy_out, T omega_nu = (T)0, T* pi_out = nullptr,
                                            T* irho_out = nullptr) {
    T rho = (T)0, ux = (T)0, uy = (T)0, inv = (T)0;
    // MRT without output and without closure needs neither u nor (off the lid) the reference-order density: the
    // collision sums rho from its own partial sums; rho_out is then only defined on the lid row
    constexpr bool LEAN = COLL == COLL_MRT && !NEED_U && !TURB;
    if (!LEAN) {
        T jx, jy;
        moments_ref<T>(f, rho, jx, jy);
        inv = (T)1 / rho;                          // u = j / rho as one reciprocal and two multiplications
        ux = jx * inv;
        uy = jy * inv;
    }
    if (left || right || bot) { ux = (T)0; uy = (T)0; }
    if (lid) {
        rho = rho_lid_formula<T>(f);
        ux = r.uLB;
        uy = (T)0;
        if (TURB) inv = (T)1 / rho;
    }
    rho_out = rho; ux_out = ux; uy_out = uy;
    const T om = TURB ? omega_nu : r.omega;
    if (TURB) {
        // the next step's pi_prev = sum_k cx cy feq_k = feq5 - feq6 + feq7 - feq8 (MRT_GPU.py:575 on the stored feq).  For
        // the reference's equilibrium the constant and linear terms cancel and the squares leave 4.5/36 * 8 ux uy: the sum
        // IS rho ux uy -- two multiplications instead of four equilibria, and without their cancellation error
        *pi_out = rho * ux * uy;
        *irho_out = inv;
    }
    if (COLL == COLL_MRT) {
        collide_mrt<T>(f, lid, rho, r.q_e, r.q_eps, r.q_q, om);     // off the lid the same density with or without output
    } else if (is_fp32<T>::value) {
        const T drho = drho_of<T>(f, lid);
        if (COLL == COLL_SRT) collide_srt_dev<T>(f, drho, rho, ux, uy, om);
        else collide_trt_dev<T>(f, drho, rho, ux, uy, om, r.omegam);
    } else {
        if (COLL == COLL_SRT) collide_srt<T>(f, rho, ux, uy, om);
        else collide_trt<T>(f, rho, ux, uy, om, r.omegam);
    }
}

}  // namespace lbm
