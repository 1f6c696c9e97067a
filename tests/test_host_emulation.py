"""The product's one-thread-per-node kernel SOURCE, executed on the CPU (tests/host_emu: g++ compiles
csrc/lbm_device.cuh, lbm_kernels.cuh and lbm_aa.cuh against a stub <cuda_runtime.h> and a loop calls the kernel function
once per node) -- so that the CPU-only suite checks the very statements of the kernels against the oracle, not only the
oracle against the reference:

* `lbm_step_ldg` (A/B one-step kernel; every other kernel family is compared bitwise with it on the GPU) against the
  oracle, fp64 <= 1e-12 and fp32 <= 1e-5, all collisions, with and without the closure, equilibrium and random starts,
  incl. the finalize pass of the download, the output lag and the current-state moments;
* `lbm_step_aa` (AA pattern, one population buffer) against `lbm_step_ldg` BIT FOR BIT, for even and odd step counts.

This is test infrastructure: nothing here ships, the product has no CPU path (tests/test_capi_symbols.py)."""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest

from oracle import lbm_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
COLL = {"SRT": 0, "TRT": 1, "MRT": 2}
TOL = {"float64": 1e-12, "float32": 1e-5}


@pytest.fixture(scope="module")
def emu():
    spec = importlib.util.spec_from_file_location("build_emu", os.path.join(HERE, "host_emu", "build_emu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = C.CDLL(mod.build())
    dp = C.POINTER(C.c_double)
    lib.emu_run.argtypes = [C.c_int] * 7 + [dp] * 7
    lib.emu_run.restype = C.c_int
    return lib


def run_emu(lib, family, dtype, p, steps, fin0=None, current=False):
    nx, ny = p.nx, p.ny
    rates = np.array([p.uLB, p.omega, p.omegam, p.omega_e, p.omega_eps, p.omega_q, 1.0 / p.omega])
    f = np.empty((9, nx, ny)); rho = np.empty((nx, ny)); u = np.empty((2, nx, ny))
    rc = np.empty((nx, ny)) if current else None
    uc = np.empty((2, nx, ny)) if current else None
    dp = C.POINTER(C.c_double)
    ptr = lambda a: a.ctypes.data_as(dp) if a is not None else None
    f0 = None if fin0 is None else np.ascontiguousarray(fin0, dtype=np.float64)
    assert lib.emu_run(family, int(dtype == "float64"), COLL[p.collision], int(p.turb), nx, ny, steps, ptr(rates), ptr(f0),
                       ptr(f), ptr(rho), ptr(u), ptr(rc), ptr(uc)) == 0
    return (rho, u, f) if not current else (rho, u, f, rc, uc)


def close(got, want, dtype, uLB=0.08):
    e = (np.abs(got[0] - want[0]).max(), np.abs(got[1] - want[1]).max() / uLB, np.abs(got[2] - want[2]).max())
    assert max(e) <= TOL[dtype], (dtype, e)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("coll,turb", [("MRT", 0), ("SRT", 0), ("TRT", 0), ("SRT", 1), ("MRT", 1)])
def test_one_step_kernel_source_against_the_oracle(emu, coll, turb, dtype):
    p = O.Params(40, 28, Re=400.0, collision=coll, turb=turb)
    for steps in (1, 2, 37):
        close(run_emu(emu, 0, dtype, p, steps), O.run_fast(p, steps), dtype)
    fin0 = O.random_state(40, 28, seed=5)
    close(run_emu(emu, 0, dtype, p, 12, fin0), O.run_fast(p, 12, fin0), dtype)


def test_one_step_kernel_source_every_boundary_class_and_minimum_size(emu):
    """3 x 3 is all walls and corners but one node; 5 x 4 has every class of wall node next to another."""
    for nx, ny in ((3, 3), (5, 4), (33, 3), (3, 34)):
        p = O.Params(nx, ny, Re=50.0, collision="MRT")
        fin0 = O.random_state(nx, ny, seed=nx * 100 + ny)
        for steps in (1, 2, 3, 8):
            close(run_emu(emu, 0, "float64", p, steps, fin0), O.run(p, steps, fin0=fin0, form="pull"), "float64")


def test_current_moments_and_output_lag(emu):
    p = O.Params(24, 20, Re=200.0, collision="MRT")
    rho, u, f, rho_c, u_c = run_emu(emu, 0, "float64", p, 9, current=True)
    want_lag = O.run(p, 9, form="pull")            # rho, u of the state that ENTERED step 9 ...
    assert np.abs(rho - want_lag[0]).max() < 1e-13 and np.abs(u - want_lag[1]).max() < 1e-13
    nxt = O.run(p, 10, form="pull")                # ... and the current ones are what a tenth step would report
    assert np.abs(rho_c - nxt[0]).max() < 1e-13 and np.abs(u_c - nxt[1]).max() < 1e-13


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("coll,turb", [("MRT", 0), ("SRT", 0), ("TRT", 0), ("SRT", 1), ("MRT", 1)])
def test_aa_pattern_equals_the_ab_kernel_bitwise(emu, coll, turb, dtype):
    """One buffer, alternating EVEN / ODD steps: the same bits as the A/B kernel after any number of steps -- populations
    (through the AA finalize pass), lagged and current moments -- from the equilibrium and from an uploaded state."""
    p = O.Params(37, 21, Re=400.0, collision=coll, turb=turb)
    fin0 = O.random_state(37, 21, seed=11)
    for start in (None, fin0):
        for steps in (0, 1, 2, 3, 10, 31):
            ab = run_emu(emu, 0, dtype, p, steps, start, current=True)
            aa = run_emu(emu, 1, dtype, p, steps, start, current=True)
            for x, y in zip(ab, aa):
                assert np.array_equal(x, y), (coll, turb, dtype, steps)


def test_aa_pattern_minimum_sizes_and_against_the_oracle(emu):
    for nx, ny in ((3, 3), (4, 3), (3, 5), (64, 7)):
        p = O.Params(nx, ny, Re=80.0, collision="MRT")
        fin0 = O.random_state(nx, ny, seed=nx + ny)
        for steps in (1, 2, 5, 6):
            close(run_emu(emu, 1, "float64", p, steps, fin0), O.run(p, steps, fin0=fin0, form="pull"), "float64")
    p = O.Params(48, 48, Re=1000.0, collision="SRT", turb=1)
    close(run_emu(emu, 1, "float64", p, 200), O.run_fast(p, 200), "float64")
    close(run_emu(emu, 1, "float32", p, 201), O.run_fast(p, 201), "float32")


@pytest.mark.parametrize("name", ["ref_A_32x32_Re100_N25.npz", "ref_A_40x24_Re400_N60.npz"])
def test_semantics_A_kernel_source_against_the_real_MRT_py(emu, name):
    """The compatibility mode's two passes (lbm_A_collide, lbm_A_stream_bc) against OUTPUTS OF THE REAL MRT.py (goldens
    written by executing the reference script, tests/golden/make_golden.py): BASELINE config 1 semantics, on the CPU."""
    d = np.load(os.path.join(HERE, "golden", name))
    nx, ny, Re, n, uLB = d["meta"]
    p = O.Params(int(nx), int(ny), uLB=float(uLB), Re=float(Re), collision="SRT")
    close(run_emu(emu, 2, "float64", p, int(n)), (d["rho"], d["u"], d["fin"]), "float64", float(uLB))
    close(run_emu(emu, 2, "float32", p, int(n)), (d["rho"], d["u"], d["fin"]), "float32", float(uLB))
    # from a random state too, against the oracle's restatement of MRT.py, incl. the rows / columns its slices never reach
    f0 = O.random_state(int(nx), int(ny), seed=8)
    got = run_emu(emu, 2, "float64", p, 40, f0)
    close(got, O.run(p, 40, semantics="A", fin0=f0), "float64", float(uLB))
    assert np.array_equal(got[2][3, int(nx) - 2, 1:-1], f0[3, int(nx) - 2, 1:-1])


def test_equ_kernel_source_is_the_reference_expression_bitwise(emu):
    """`functions.equ` (functions.pyx:229-267 == MRT.py:213-231) through lbm_equ_kernel: the reference's operation order
    without any fused operation, hence the same bits as the NumPy expression of the oracle -- and as the compiled
    reference module's output stored with the golden vectors."""
    rng = np.random.default_rng(3)
    n = 1000
    rho = 1 + 0.05 * rng.uniform(-1, 1, n); ux = 0.1 * rng.uniform(-1, 1, n); uy = 0.1 * rng.uniform(-1, 1, n)
    feq = np.empty((9, n))
    dp = C.POINTER(C.c_double)
    emu.emu_equ.argtypes = [C.c_longlong] + [dp] * 4
    assert emu.emu_equ(n, *[a.ctypes.data_as(dp) for a in (rho, ux, uy, feq)]) == 0
    want = O.equ(rho.reshape(n, 1), np.stack([ux, uy]).reshape(2, n, 1))[:, :, 0]
    assert np.array_equal(feq, want)
