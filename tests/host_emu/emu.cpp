// TEST INFRASTRUCTURE: runs the product's one-thread-per-node kernel SOURCE on the CPU, node by node (see stub/cuda_runtime.h).
// Built by tests/host_emu/build_emu.py with g++ (-ffp-contract=off: like the library's -fmad=false, every fused
// multiply-add is the explicit fm() of lbm_device.cuh).  Mirrors, for one whole cavity, what lbm_b200.cu does around the
// kernels: device layout [k][row + 1][pitch], side buffers, init / upload seeding, A/B ping-pong or the AA phases,
// finalize pass for the download.  family 0 = lbm_step_ldg (A/B), 1 = lbm_step_aa (single buffer), 2 = the semantics-A passes
// (lbm_A_collide + lbm_A_stream_bc, SRT only).
#include <cuda_runtime.h>      // the stub

#include <vector>

#include "lbm_device.cuh"
#include "lbm_kernels_nopdl.cuh"      // csrc/lbm_kernels.cuh minus the two griddepcontrol lines per kernel (build_emu.py)
#include "lbm_aa.cuh"

using namespace lbm;

template <typename K>
static void launch(K kern, const StepArgs& a, unsigned gx, unsigned gy, unsigned gz, unsigned tx = 1) {
    blockDim = dim3(tx, 1, 1);
    gridDim = dim3(gx, gy, gz);
    for (unsigned z = 0; z < gz; ++z)
        for (unsigned y = 0; y < gy; ++y)
            for (unsigned x = 0; x < gx; ++x)
                for (unsigned t = 0; t < tx; ++t) {
                    blockIdx.x = x; blockIdx.y = y; blockIdx.z = z;
                    threadIdx.x = t; threadIdx.y = 0; threadIdx.z = 0;
                    kern(a);
                }
}

template <typename T, int COLL, bool TURB>
static void step_ldg(const StepArgs& a, bool gather, bool macros, unsigned nx, unsigned ny) {
    if (gather) {
        if (macros) launch(lbm_step_ldg<T, COLL, true, true, MODE_STEP, TURB>, a, nx, ny, 1);
        else launch(lbm_step_ldg<T, COLL, true, false, MODE_STEP, TURB>, a, nx, ny, 1);
    } else {
        if (macros) launch(lbm_step_ldg<T, COLL, false, true, MODE_STEP, TURB>, a, nx, ny, 1);
        else launch(lbm_step_ldg<T, COLL, false, false, MODE_STEP, TURB>, a, nx, ny, 1);
    }
}

template <typename T, int COLL, bool TURB>
static void step_aa(const StepArgs& a, bool odd, bool walls, bool macros, unsigned nx, unsigned ny) {
    if (odd) {
        if (macros) launch(lbm_step_aa<T, COLL, true, true, true, MODE_STEP, TURB>, a, nx, ny, 1);
        else launch(lbm_step_aa<T, COLL, true, true, false, MODE_STEP, TURB>, a, nx, ny, 1);
    } else if (walls) {
        if (macros) launch(lbm_step_aa<T, COLL, false, true, true, MODE_STEP, TURB>, a, nx, ny, 1);
        else launch(lbm_step_aa<T, COLL, false, true, false, MODE_STEP, TURB>, a, nx, ny, 1);
    } else {
        if (macros) launch(lbm_step_aa<T, COLL, false, false, true, MODE_STEP, TURB>, a, nx, ny, 1);
        else launch(lbm_step_aa<T, COLL, false, false, false, MODE_STEP, TURB>, a, nx, ny, 1);
    }
}

template <typename T, int COLL, bool TURB>
static int run(int family, int nx, int ny, int steps, const CavityParams& cp, const double* f0, double* f_out,
               double* rho_out, double* u_out, double* rho_cur, double* u_cur) {
    const int pitch = (nx + 31) / 32 * 32, rows = ny + 2;
    const long long plane = (long long)rows * pitch, cavity = 9 * plane, mplane = (long long)ny * pitch;
    std::vector<T> buf[2], scratch(cavity, (T)0);
    buf[0].assign(cavity, (T)0);
    if (family == 0) buf[1].assign(cavity, (T)0);
    std::vector<T> rho(mplane, (T)0), ux(mplane, (T)0), uy(mplane, (T)0), rho_lid(pitch, (T)0), carry(4, (T)0);
    std::vector<T> pi_eq(TURB ? mplane : 0, (T)0), rho_prev(TURB ? mplane : 0, (T)0);
    CavityParams cav = cp;
    StepArgs a{};
    a.rho = rho.data(); a.ux = ux.data(); a.uy = uy.data();
    a.rho_lid = rho_lid.data(); a.carry = carry.data(); a.rho_lid_out = rho_lid.data(); a.carry_out = carry.data();
    if (TURB) { a.pi_eq = pi_eq.data(); a.rho_prev = rho_prev.data(); a.pi_eq_out = pi_eq.data(); a.rho_prev_out = rho_prev.data(); }
    a.cav = &cav; a.active = nullptr;
    a.nx = nx; a.ny = ny; a.y0 = 0; a.nyl = ny; a.pitch = pitch; a.plane = plane; a.cavity = cavity; a.mplane = mplane;
    a.row_begin = 0; a.row_stride = 1; a.row_count = ny;
    // initial state: lbm_init_equilibrium / lbm_upload_f
    if (!f0) {
        a.src = nullptr; a.dst = buf[0].data();
        launch(lbm_init_eq<T>, a, nx, ny, 1);
    } else {
        for (int k = 0; k < 9; ++k)
            for (int x = 0; x < nx; ++x)
                for (int y = 0; y < ny; ++y)
                    buf[0][k * plane + (long long)(y + 1) * pitch + x] = (T)f0[((long long)k * nx + x) * ny + y];
        a.src = buf[0].data(); a.dst = nullptr;
        launch(lbm_seed_carry<T>, a, 1, 1, 1, 4);
        for (long long i = 0; i < mplane; ++i) { rho[i] = (T)1; ux[i] = (T)0; uy[i] = (T)0; }
        if (TURB) launch(lbm_seed_turb<T>, a, nx, ny, 1);
    }
    int cur = 0;
    bool pre = true, swapped = false;
    if (family == 2) {
        // semantics A (the two plain passes of step_A in lbm_b200.cu): collide f[0] -> f[1], stream + walls f[1] -> f[0] in place
        buf[1].assign(cavity, (T)0);
        for (int i = 0; i < steps; ++i) {
            a.src = buf[0].data(); a.dst = buf[1].data();
            launch(lbm_A_collide<T>, a, nx, ny, 1);
            a.src = buf[1].data(); a.dst = buf[0].data();
            launch(lbm_A_stream_bc<T>, a, nx, ny, 1);
        }
        steps = 0;       // nothing left for the loops below; f[0] is `fin`, rho / u are the step's own moments
    }
    for (int i = 0; i < steps; ++i) {
        const bool macros = (i == steps - 1);
        if (family == 0) {
            a.src = buf[cur].data(); a.dst = buf[cur ^ 1].data();
            step_ldg<T, COLL, TURB>(a, !pre, macros, nx, ny);
            cur ^= 1;
        } else {
            a.src = buf[0].data(); a.dst = buf[0].data();
            step_aa<T, COLL, TURB>(a, swapped, !pre, macros, nx, ny);
            swapped = !swapped;
        }
        pre = false;
    }
    // current moments (lbm_get_macros_current) into copies of the macro arrays, then the download's finalize pass
    std::vector<T> rho_l = rho, ux_l = ux, uy_l = uy;
    if (rho_cur) {
        if (family == 0) {
            a.src = buf[cur].data(); a.dst = buf[cur ^ 1].data();
            if (pre) launch(lbm_step_ldg<T, COLL_MRT, false, true, MODE_MACROS>, a, nx, ny, 1);
            else launch(lbm_step_ldg<T, COLL_MRT, true, true, MODE_MACROS>, a, nx, ny, 1);
        } else {
            a.src = buf[0].data(); a.dst = nullptr;
            if (swapped) launch(lbm_step_aa<T, COLL_MRT, true, true, true, MODE_MACROS>, a, nx, ny, 1);
            else if (!pre) launch(lbm_step_aa<T, COLL_MRT, false, true, true, MODE_MACROS>, a, nx, ny, 1);
            else launch(lbm_step_aa<T, COLL_MRT, false, false, true, MODE_MACROS>, a, nx, ny, 1);
        }
        for (int x = 0; x < nx; ++x)
            for (int y = 0; y < ny; ++y) {
                rho_cur[(long long)x * ny + y] = (double)rho[(long long)y * pitch + x];
                u_cur[(long long)x * ny + y] = (double)ux[(long long)y * pitch + x];
                u_cur[((long long)nx + x) * ny + y] = (double)uy[(long long)y * pitch + x];
            }
    }
    const T* fin = buf[cur].data();
    if (!pre) {
        if (family == 0) {
            a.src = buf[cur].data(); a.dst = buf[cur ^ 1].data();
            launch(lbm_step_ldg<T, COLL_MRT, true, false, MODE_FINALIZE>, a, nx, ny, 1);
            fin = buf[cur ^ 1].data();
        } else {
            a.src = buf[0].data(); a.dst = scratch.data();
            if (swapped) launch(lbm_step_aa<T, COLL_MRT, true, true, false, MODE_FINALIZE>, a, nx, ny, 1);
            else launch(lbm_step_aa<T, COLL_MRT, false, true, false, MODE_FINALIZE>, a, nx, ny, 1);
            fin = scratch.data();
        }
    }
    for (int k = 0; k < 9; ++k)
        for (int x = 0; x < nx; ++x)
            for (int y = 0; y < ny; ++y)
                f_out[((long long)k * nx + x) * ny + y] = (double)fin[k * plane + (long long)(y + 1) * pitch + x];
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y) {
            rho_out[(long long)x * ny + y] = (double)rho_l[(long long)y * pitch + x];
            u_out[(long long)x * ny + y] = (double)ux_l[(long long)y * pitch + x];
            u_out[((long long)nx + x) * ny + y] = (double)uy_l[(long long)y * pitch + x];
        }
    return 0;
}

template <typename T, int COLL>
static int run_t(int family, int turb, int nx, int ny, int steps, const CavityParams& cp, const double* f0, double* f_out,
                 double* rho_out, double* u_out, double* rho_cur, double* u_cur) {
    return turb ? run<T, COLL, true>(family, nx, ny, steps, cp, f0, f_out, rho_out, u_out, rho_cur, u_cur)
                : run<T, COLL, false>(family, nx, ny, steps, cp, f0, f_out, rho_out, u_out, rho_cur, u_cur);
}

template <typename T>
static int run_c(int family, int coll, int turb, int nx, int ny, int steps, const CavityParams& cp, const double* f0,
                 double* f_out, double* rho_out, double* u_out, double* rho_cur, double* u_cur) {
    switch (coll) {
        case 0: return run_t<T, COLL_SRT>(family, turb, nx, ny, steps, cp, f0, f_out, rho_out, u_out, rho_cur, u_cur);
        case 1: return run_t<T, COLL_TRT>(family, turb, nx, ny, steps, cp, f0, f_out, rho_out, u_out, rho_cur, u_cur);
        default: return run_t<T, COLL_MRT>(family, turb, nx, ny, steps, cp, f0, f_out, rho_out, u_out, rho_cur, u_cur);
    }
}

// rates[7] = {uLB, omega, omega-, s_e, s_eps, s_q, tau0} (CavityParams); f0 = NULL starts from the equilibrium; outputs in
// the reference's host layout ([9][nx][ny], [nx][ny], [2][nx][ny]) as doubles; rho_cur / u_cur (may be NULL) receive the
// current-state moments.
extern "C" int emu_run(int family, int is_f64, int coll, int turb, int nx, int ny, int steps, const double* rates,
                       const double* f0, double* f_out, double* rho_out, double* u_out, double* rho_cur, double* u_cur) {
    CavityParams cp{};
    cp.uLB = rates[0]; cp.omega = rates[1]; cp.omegam = rates[2]; cp.s_e = rates[3]; cp.s_eps = rates[4]; cp.s_q = rates[5];
    cp.tau0 = rates[6];
    if (is_f64) return run_c<double>(family, coll, turb, nx, ny, steps, cp, f0, f_out, rho_out, u_out, rho_cur, u_cur);
    return run_c<float>(family, coll, turb, nx, ny, steps, cp, f0, f_out, rho_out, u_out, rho_cur, u_cur);
}

// ---- y-strips: the one-step kernel on caller-owned device-layout buffers -------------------------------------------
// One launch of lbm_step_ldg (fp64, MRT) over the local rows [row_begin, row_begin + row_count) of a strip that owns
// global rows [y0, y0 + nyl): src / dst are [9][nyl + 2][pitch] with one ghost row above and below (lbm_layout_t), the
// side buffers as in the library.  mode: 0 step (gather != 0 after the first step), 1 finalize (download), 3 initial
// equilibrium into dst (lbm_init_eq).  What tests/test_distributed_cpu.py runs under gloo between two halo exchanges.
extern "C" int emu_strip_pass(int mode, int gather, int macros, int nx, int ny, int y0, int nyl, const double* rates,
                              const double* src, double* dst, double* rho, double* ux, double* uy, double* rho_lid,
                              double* carry, int row_begin, int row_count) {
    CavityParams cav{};
    cav.uLB = rates[0]; cav.omega = rates[1]; cav.omegam = rates[2]; cav.s_e = rates[3]; cav.s_eps = rates[4]; cav.s_q = rates[5];
    cav.tau0 = rates[6];
    StepArgs a{};
    const int pitch = (nx + 31) / 32 * 32;
    a.src = src; a.dst = dst; a.rho = rho; a.ux = ux; a.uy = uy;
    a.rho_lid = rho_lid; a.carry = carry; a.rho_lid_out = rho_lid; a.carry_out = carry;
    a.cav = &cav;
    a.nx = nx; a.ny = ny; a.y0 = y0; a.nyl = nyl; a.pitch = pitch;
    a.plane = (long long)(nyl + 2) * pitch; a.cavity = 9 * a.plane; a.mplane = (long long)nyl * pitch;
    a.row_begin = row_begin; a.row_stride = 1; a.row_count = row_count;
    if (mode == 3) { launch(lbm_init_eq<double>, a, nx, nyl, 1); return 0; }
    if (mode == 1) { launch(lbm_step_ldg<double, COLL_MRT, true, false, MODE_FINALIZE>, a, nx, row_count, 1); return 0; }
    if (gather) {
        if (macros) launch(lbm_step_ldg<double, COLL_MRT, true, true, MODE_STEP>, a, nx, row_count, 1);
        else launch(lbm_step_ldg<double, COLL_MRT, true, false, MODE_STEP>, a, nx, row_count, 1);
    } else {
        if (macros) launch(lbm_step_ldg<double, COLL_MRT, false, true, MODE_STEP>, a, nx, row_count, 1);
        else launch(lbm_step_ldg<double, COLL_MRT, false, false, MODE_STEP>, a, nx, row_count, 1);
    }
    return 0;
}

// functions.equ through lbm_equ_kernel (fp64): feq[9][n] from rho[n], ux[n], uy[n], a grid-stride launch of 3 x 2 threads
extern "C" int emu_equ(long long n, const double* rho, const double* ux, const double* uy, double* feq) {
    blockDim = dim3(2, 1, 1);
    gridDim = dim3(3, 1, 1);
    for (unsigned b = 0; b < 3; ++b)
        for (unsigned t = 0; t < 2; ++t) {
            blockIdx.x = b; blockIdx.y = 0; blockIdx.z = 0;
            threadIdx.x = t; threadIdx.y = 0; threadIdx.z = 0;
            lbm_equ_kernel<double>(rho, ux, uy, feq, n);
        }
    return 0;
}
