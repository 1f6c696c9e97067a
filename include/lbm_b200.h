/*
 * lbm_b200.h -- C ABI of the B200-native D2Q9 lid-driven-cavity collide-and-stream step.
 *
 * This is the drop-in boundary for ONE hot path of RaghuvirJonnagiri/LatticeBoltzmannSimulations:
 * the per-time-step update "moments -> collision (SRT/TRT/MRT) -> streaming -> wall + moving-lid rule".
 * Every entry point names the reference interface it replaces (file:line into the upstream repo).
 * Plain C: opaque handle, plain pointers and sizes, int return codes (0 = ok) plus a thread-local
 * error string.  No torch / numpy types appear here; Python binds it with ctypes
 * (latticeboltzmannsimulations_b200/_capi.py), and INTEGRATION.md shows the reference-side stubs.
 *
 * Host array convention = the reference's: populations [9][nx][ny] (y fastest), rho [nx][ny],
 * u [2][nx][ny]; y == 0 is the moving lid, y == ny-1 the bottom wall (MRT.py:252-261).  The device
 * layout (SoA, x fastest, one ghost row above and below each y-strip) is private to the library and
 * described by lbm_get_layout() for callers that exchange halo rows themselves.
 *
 * Semantics: "C" of SURVEY.md 3.4 == MRT_GPU.py funRT + funBC (push + non-equilibrium bounce-back),
 * re-expressed as one fused pull pass (oracle/lbm_oracle.py step_C_pull is the executable spec).
 * There is NO CPU fallback: every call fails with LBM_ECUDA when no CUDA device is usable.
 */
#ifndef LBM_B200_H
#define LBM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBM_B200_ABI_VERSION 3

enum lbm_status { LBM_OK = 0, LBM_EINVAL = 1, LBM_ECUDA = 2, LBM_ENOMEM = 3, LBM_ESTATE = 4 };
enum lbm_dtype { LBM_F32 = 0, LBM_F64 = 1 };
/* `RT` of MRT_GPU.py:48 */
enum lbm_collision { LBM_SRT = 0, LBM_TRT = 1, LBM_MRT = 2 };
/* which rows of the local strip a launch covers (multi-GPU overlap of halo exchange and interior): EDGE = the first
 * two and last two rows (the rows a halo exchange ships), INTERIOR = the rest, ALL = both */
enum lbm_region { LBM_REGION_ALL = 0, LBM_REGION_EDGE = 1, LBM_REGION_INTERIOR = 2 };
/* which reference variant the step reproduces: C = MRT_GPU.py (push + NEBB in funBC; the product
 * path), A = MRT.py (NumPy solver: SRT only, slice streaming with the exclusive xsize_max bound that leaves the
 * last rows/columns stale, "= feq" left wall) -- a compatibility mode so that BASELINE config 1 can be compared on
 * identical inputs; two simple passes per step, not tuned */
enum lbm_semantics { LBM_SEMANTICS_C = 0, LBM_SEMANTICS_A = 1 };
/* kernel family: plain coalesced loads (with the two-step kernels on top, the default), TMA-staged persistent tiles, or
 * the AA pattern -- one-step kernels on ONE population buffer instead of the A/B pair (half the memory: a 32768^2 fp64
 * cavity takes 77 GB instead of 155 GB; results bit-identical to the A/B one-step kernels; one lattice step per pass
 * over memory, so about half the MLUPS of the two-step kernels).  AA handles hold whole cavities (no y-strips, no
 * caller-owned buffers, no frozen cavities / convergence rule); lbm_download_f borrows a second buffer for the call. */
enum lbm_engine { LBM_ENGINE_AUTO = 0, LBM_ENGINE_LDG = 1, LBM_ENGINE_TMA = 2, LBM_ENGINE_AA = 3 };

typedef struct lbm_solver* lbm_handle_t;

/* Replaces the module-level constants at the top of every solver script
 * (MRT_GPU.py:45-58: RT, turb, xsize, ysize; MRT_GPU_datagen.py:55-70 for the batch of cavities). */
typedef struct lbm_config {
    int32_t nx, ny;       /* global cavity size in nodes (no multiple-of-32 restriction, cf. MRT_GPU.py:53-54) */
    int32_t batch;        /* independent cavities held by this handle (MRT_GPU_datagen.py Re sweep) */
    int32_t dtype;        /* lbm_dtype: storage and arithmetic type */
    int32_t collision;    /* lbm_collision */
    int32_t turb;         /* 1 = Smagorinsky closure of MRT_GPU.py:570-589 */
    int32_t y0;           /* first global row owned by this handle (y-strip decomposition) */
    int32_t ny_local;     /* rows owned; 0 means the whole cavity (y0 must then be 0) */
    int32_t device;       /* CUDA device ordinal, -1 = current device */
    int32_t engine;       /* lbm_engine */
    int32_t semantics;    /* lbm_semantics */
    int32_t reserved;     /* must be 0 */
    void* ext_f[2];       /* optional caller-owned device buffers (e.g. torch tensors) for the A/B population
                             arrays, each lbm_state_bytes() long; both NULL = the library allocates
                             (one NULL and one not, or twice the same pointer: LBM_EINVAL) */
} lbm_config_t;

/* Private device layout, for callers that move halo rows themselves (NCCL send/recv on row views). */
typedef struct lbm_layout {
    int64_t elem_size;      /* 4 or 8 */
    int64_t pitch;          /* elements per stored row (>= nx, multiple of 32) */
    int64_t rows;           /* stored rows per population plane = ny_local + 2 (ghost row first and last) */
    int64_t plane;          /* elements per population plane = rows * pitch */
    int64_t cavity;         /* elements per cavity = 9 * plane */
    int64_t state_bytes;    /* bytes of one A/B buffer = (batch * cavity + batch * 6 * pitch) * elem_size */
    int64_t ghost2_offset;  /* elements from the buffer start to the tail [batch][top|bottom][3][pitch]: the SECOND
                               ghost rows used by the two-step kernel on a y-strip -- top: populations 4,7,8 of global
                               row y0-2, bottom: populations 2,5,6 of global row y0+ny_local+1 */
} lbm_layout_t;

const char* lbm_last_error(void);
int lbm_abi_version(void);
int lbm_device_count(int* count);

/* Size of one population buffer for this configuration (so the caller can allocate ext_f). */
int lbm_state_bytes(const lbm_config_t* cfg, size_t* bytes);

/* cuda.mem_alloc x8 + constants of MRT_GPU.py:309-328 -> one handle. */
int lbm_create(const lbm_config_t* cfg, lbm_handle_t* out);
int lbm_destroy(lbm_handle_t h);
int lbm_get_layout(lbm_handle_t h, lbm_layout_t* out);

/* Kernel-selection knobs.  None of them changes a result (every kernel family is bit-identical to every other,
 * tests/test_gpu_parity.py::test_kernel_variants_are_bit_identical); they exist for the tuning tools under tools/ and
 * for those tests.  The reference has no counterpart (its block shape is the literal of MRT_GPU.py:53-54).  Keys:
 *   "two_step"            0 = one lattice step per launch only (default 1: temporal blocking where it pays)
 *   "two_step_min_nodes"  smallest batch x nx x ny that uses a two-step kernel (default 10000)
 *   "slide"               0 = never use the sliding-window two-step kernel (default 1)
 *   "slide_min_nodes"     smallest batch x nx x ny that uses it instead of the shared-memory tiles (default 1500000)
 *   "slide_h"             rows per segment of the sliding-window kernel, 0 = automatic
 *   "slide_tma"           1 (default): interior blocks of the sliding-window kernel are staged by tensor copies (one
 *                         box per population and iteration), 0: one bulk copy per staged row
 *   "tile"                tile shape of the shared-memory two-step kernel, -1 = automatic
 *   "vec_f64", "vec_f32"  nodes per thread of the one-step kernels (1|2; 0|1|2|4 with 0 = by size, the default)
 *   "graph", "pdl"        CUDA graphs for small cavities / programmatic dependent launch (default 1, 1)
 *   "tma_variant", "tma_ctas"   tile configuration and CTAs per SM of the optional TMA engine
 * Unknown keys and out-of-range values return LBM_EINVAL.  The Python binding applies the comma-separated
 * "key=value" list in the environment variable LBM_B200_TUNING to every handle it creates. */
int lbm_set_tuning(lbm_handle_t h, const char* key, int64_t value);

/* functions.pyx:38-43 set_omega(uLB, Re, ysize) / MRT_GPU.py:63-65: omega = 2 / (6 uLB ny / Re + 1); the other
 * MRT rates take the GPU-script values (omega_e 1.0, omega_eps = omega_q = 1.2, MRT_GPU.py:88-91) and the TRT
 * omega- follows MRT_GPU.py:82-84.  cavity = -1 applies to every cavity of the batch. */
int lbm_set_reynolds(lbm_handle_t h, int cavity, double uLB, double Re);
/* Explicit rates (the literals spliced into the kernel source at MRT_GPU.py:422,531,662). */
int lbm_set_rates(lbm_handle_t h, int cavity, double uLB, double omega_nu, double omega_e,
                  double omega_eps, double omega_q, double omega_minus);

/* Host init of MRT_GPU.py:259-267 + uploads :323-328: rho = 1, u = (uLB,0) on row y = 0, f = feq.  Enqueued on the
 * default (NULL) stream: a caller that steps on a non-blocking stream orders it after this call itself (event or
 * lbm_sync). */
int lbm_init_equilibrium(lbm_handle_t h);
/* cuda.memcpy_htod(fin_g, fin) (MRT_GPU.py:323) incl. the [k,x,y]->[k,y,x] transposes of :283-289.
 * f: [batch][9][nx][ny_local] in the handle's dtype; on_device != 0 means `f` is a device pointer. */
int lbm_upload_f(lbm_handle_t h, const void* f, int on_device, void* stream);
/* cuda.memcpy_dtoh(fin, ftemp_g) + transposes (MRT_GPU.py:755, 758-760): the reference's `fin` after the
 * steps taken so far.  Does not disturb the state. */
int lbm_download_f(lbm_handle_t h, void* f, int on_device, void* stream);

/* The hot call.  Replaces the Python time loop body of MRT_GPU.py:707-732 (funRT + funBC launches) and the
 * per-step functions.allfunc of MRT_cython.py:453.  Runs `nsteps` fused steps asynchronously on `stream`
 * (a cudaStream_t, NULL = default stream).  With write_macros != 0 the LAST step also stores rho and u as the
 * reference does every step: the overridden moments of the state that entered that step. */
int lbm_step(lbm_handle_t h, int nsteps, int write_macros, void* stream);

/* One step restricted to a row region, without advancing the step counter; lbm_swap() completes the step.
 * Used by the y-strip driver: EDGE rows first, halo rows exchanged while INTERIOR runs. */
int lbm_step_region(lbm_handle_t h, int region, int write_macros, void* stream);
int lbm_swap(lbm_handle_t h);
/* Same for the temporal-blocking kernel: TWO steps of a row region in one launch (the state must be post-collision,
 * i.e. at least one ordinary step taken since the last upload, and ny_local >= 2); lbm_swap2() completes the double
 * step.  Before the next (double) step the ghost rows of the written buffer must hold, from the strip above, all of
 * populations {0,1,3,4,7,8} of its last row and {4,7,8} of the row before it (-> ghost2 top), and from the strip
 * below {0,1,3,2,5,6} of its first row and {2,5,6} of its second row (-> ghost2 bottom). */
int lbm_step2_region(lbm_handle_t h, int region, int write_macros, void* stream);
int lbm_swap2(lbm_handle_t h);
/* 1 if lbm_step2_region can be used right now (temporal blocking enabled, post-collision state, ny_local >= 2, ...). */
int lbm_step2_available(lbm_handle_t h);
/* Device pointers of the buffer being read (which = 0) / written (which = 1) by the next lbm_step_region. */
int lbm_buffer_ptr(lbm_handle_t h, int which, void** ptr);

/* The nine halo rows of one side as ONE contiguous device buffer buf[9][nx] (a single send / recv per neighbour
 * instead of nine), for the buffer written by the (double) step in progress, i.e. before lbm_swap / lbm_swap2:
 * pack gathers what goes to the strip above (dir = 0: populations {0,1,3,2,5,6} of the first row, {2,5,6} of the
 * second) or below (dir = 1: {0,1,3,4,7,8} of the last row, {4,7,8} of the row before it); unpack scatters what came
 * from the strip above (dir = 0) / below (dir = 1) into the ghost row and the second ghost rows of that side.
 * Single-cavity strips of >= 2 rows.  (No reference counterpart: the upstream solvers are single-GPU.) */
int lbm_halo_pack(lbm_handle_t h, int dir, void* buf, void* stream);
int lbm_halo_unpack(lbm_handle_t h, int dir, const void* buf, void* stream);

/* cuda.memcpy_dtoh(rho, rho_g) / (u, u_g) + transposes (MRT_GPU.py:756-760): rho [batch][nx][ny_local],
 * u [batch][2][nx][ny_local].  Either pointer may be NULL. */
int lbm_get_macros(lbm_handle_t h, void* rho, void* u, int on_device, void* stream);
/* Same fields evaluated from the CURRENT populations (no one-step lag). */
int lbm_get_macros_current(lbm_handle_t h, void* rho, void* u, int on_device, void* stream);

/* The equilibrium of the stored (lagged) rho, u -- the `feq` that functions.allfunc returns next to them
 * (functions.pyx:88, 222: equ(rho, u) of the step's own moments) -- computed on the device from the stored fields
 * instead of sending rho, u back up: feq [batch][9][nx][ny_local]. */
int lbm_get_feq(lbm_handle_t h, void* feq, int on_device, void* stream);

/* functions.equ(rho, ux, uy) (functions.pyx:229-267, == MRT.py:213-231): second-order equilibrium of arbitrary
 * fields, stateless.  rho, ux, uy: [n] values; feq: [9][n]; dtype = lbm_dtype; on_device != 0 for device pointers. */
int lbm_equilibrium(int dtype, int64_t n, const void* rho, const void* ux, const void* uy, void* feq,
                    int on_device, void* stream);

/* Per-cavity np.mean(u) over both components of the stored (lagged) velocity field -- the quantity of the
 * convergence test `abs(np.mean(u) - np.mean(u_past)) / uLB < 1e-7` in MRT_GPU_datagen.py:729 (MRT_GPU.py:883),
 * reduced on the device.  mean_out: [batch] doubles on the host. */
int lbm_mean_u(lbm_handle_t h, double* mean_out, void* stream);
/* Freeze cavities of a batch (active[b] == 0): the `break` of MRT_GPU_datagen.py:731-733 per cavity.  Frozen
 * cavities keep their populations and macros and no longer cost bandwidth; they cannot be re-activated. */
int lbm_set_active(lbm_handle_t h, const int32_t* active, void* stream);

/* The whole stopping rule of MRT_GPU_datagen.py:726-733 on the device, for every cavity of the batch at once: reduce
 * mean(u) of the stored velocity field, compare with the value of the previous call (zeros before the first, :223),
 * abs(mean - past) / uLB < tol increments the cavity's counter (never reset, as in the reference), a counter above
 * hits - 1 retires the cavity exactly like lbm_set_active would (`break`, :731-733).  No host round trip is needed
 * between checks; with active_out != NULL ([batch] on the host) the per-cavity flags are read back (one small copy +
 * stream synchronisation) so that the caller can stop when all are 0.  Whole-cavity handles, after at least one step. */
int lbm_converge_check(lbm_handle_t h, double tol, int hits, int32_t* active_out, void* stream);

/* Diagnostics of the stored (lagged) velocity field of one cavity, computed on the device instead of downloading the
 * full fields as the scripts do every Pinterval (MRT_GPU.py:764-776, 793-800 / MRT.py:504-516):
 *   ux_col    [ny_local]  u_x on the middle column x = nx/2                       (may be NULL)
 *   uy_row    [nx]        u_y on the middle row    y = ny/2 (must be owned by this strip)  (may be NULL)
 *   vortex_xy [4]         (x1, y1, x2, y2): argmin of |u|^2 with a border of nx/40 nodes masked, then again with a
 *                         box of +-nx/40 around the first centre masked (whole-cavity handles only; may be NULL).
 * Host pointers in the handle's dtype. */
int lbm_diagnostics(lbm_handle_t h, int cavity, void* ux_col, void* uy_row, int32_t* vortex_xy, void* stream);

int lbm_sync(lbm_handle_t h);
/* Steps completed, and number of kernels this handle has launched (for the bench's gpu_launches). */
int lbm_get_counters(lbm_handle_t h, int64_t* steps_done, int64_t* kernel_launches);
/* Name of the kernel family in use ("ldg" / "tma" / "aa"). */
const char* lbm_engine_name(lbm_handle_t h);

#ifdef __cplusplus
}
#endif
#endif /* LBM_B200_H */
