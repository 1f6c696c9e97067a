"""CPU tests of the oracle: against the committed golden vectors (outputs of the real reference code),
internal consistency (push form == fused pull form), and -- when /root/reference is present -- live
against the reference itself."""
import glob
import os

import numpy as np
import pytest

from oracle import lbm_oracle as O
from oracle import ref_harness as R

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    d = np.load(os.path.join(GOLDEN, name))
    nx, ny, Re, n, uLB = d["meta"]
    return d, int(nx), int(ny), float(Re), int(n), float(uLB)


@pytest.mark.parametrize("name", ["ref_A_32x32_Re100_N25.npz", "ref_A_40x24_Re400_N60.npz"])
def test_semantics_A_equals_real_MRT_py(name):
    """Oracle 'A' must reproduce the exec'd upstream MRT.py bit for bit (golden = MRT.py's own output)."""
    d, nx, ny, Re, n, uLB = _load(name)
    p = O.Params(nx, ny, uLB=uLB, Re=Re, collision="SRT")
    rho, u, fin = O.run(p, n, semantics="A")
    assert np.array_equal(rho, d["rho"])
    assert np.array_equal(u, d["u"])
    assert np.array_equal(fin, d["fin"])


def test_semantics_C_srt_vs_compiled_allfunc():
    """One step from a seeded random state against the compiled functions.allfunc (functions.pyx:45-222):
    rho, u, feq everywhere and fin on every non-wall node; walls differ by design (B's boundary bugs)."""
    d, nx, ny, Re, n, uLB = _load("ref_allfunc_32x24_Re100.npz")
    p = O.Params(nx, ny, uLB=uLB, Re=Re, collision="SRT")
    st = O.StateC.initial(p, O.random_state(nx, ny, seed=1234))
    O.step_C(st, p)
    assert np.abs(st.rho - d["rho"]).max() <= 4.5e-16
    assert np.abs(st.u - d["u"]).max() <= 1e-16
    assert np.abs(st.feq - d["feq"]).max() <= 2.3e-16
    assert np.abs(st.fin - d["fin"])[:, 1:-1, 1:-1].max() <= 2.3e-16


@pytest.mark.parametrize("name", sorted(os.path.basename(f) for f in glob.glob(os.path.join(GOLDEN, "C_*_turb*.npz"))))
def test_semantics_C_regression_vectors(name):
    d, nx, ny, Re, n, uLB = _load(name)
    coll = name.split("_")[1]
    turb = int(name.split("_")[2][4:])
    p = O.Params(nx, ny, uLB=uLB, Re=Re, collision=coll, turb=turb)
    for form in ("push", "pull"):
        rho, u, fin = O.run(p, n, semantics="C", form=form)
        assert np.array_equal(rho, d["rho"]) and np.array_equal(u, d["u"]) and np.array_equal(fin, d["fin"]), form


@pytest.mark.parametrize("coll", ["SRT", "TRT", "MRT"])
@pytest.mark.parametrize("turb", [0, 1])
@pytest.mark.parametrize("shape", [(24, 24), (40, 16), (9, 33)])
def test_pull_form_is_the_push_form(coll, turb, shape):
    """The fused pull pass (spec of the CUDA kernel) equals funRT+funBC bit for bit, random and equilibrium start."""
    nx, ny = shape
    p = O.Params(nx, ny, Re=400, collision=coll, turb=turb)
    for f0 in (None, O.random_state(nx, ny, seed=7)):
        a = O.run(p, 40, fin0=f0, form="push")
        b = O.run(p, 40, fin0=f0, form="pull")
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("coll", ["SRT", "TRT", "MRT"])
@pytest.mark.parametrize("turb", [0, 1])
def test_c_restatement_equals_numpy_oracle(coll, turb):
    """oracle/lbm_oracle_c.c (threaded, used for the larger parity cases) reproduces the NumPy oracle bit for bit."""
    for (nx, ny, f0) in ((40, 28, O.random_state(40, 28, 3)), (33, 65, None)):
        p = O.Params(nx, ny, Re=400, collision=coll, turb=turb)
        a = O.run(p, 50, fin0=f0, form="push")
        b = O.run_fast(p, 50, fin0=f0)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def test_moment_basis_inverse():
    assert np.abs(O.M_GS @ O.M_GS_INV - np.eye(9)).max() <= 2.3e-16     # MRT.py:163-183


def test_resting_wall_nebb_is_exact_bounce_back():
    """feq_k - feq_opp(k) is exactly 0.0 where u_wall = 0, so C's NEBB is on-node bounce-back there."""
    rho = 1 + 0.3 * np.random.default_rng(0).uniform(-1, 1, (5, 5))
    fe = O._feq_kernel(rho, np.zeros((5, 5)), np.zeros((5, 5)))
    for k in range(1, 9):
        assert np.array_equal(fe[k], fe[O.BOUNCE[k]])


def test_corner_orphans_constant_bottom():
    """f6@(0,ny-1) and f5@(nx-1,ny-1) never change: the kernel may carry them as constants."""
    p = O.Params(24, 24, Re=400, collision="MRT")
    st = O.StateC.initial(p)
    a, b = st.fin[6, 0, -1], st.fin[5, -1, -1]
    for _ in range(50):
        O.step_C(st, p)
    assert st.fin[6, 0, -1] == a and st.fin[5, -1, -1] == b


def test_mass_drift_small():
    p = O.Params(32, 32, Re=100, collision="MRT")
    rho, u, fin = O.run(p, 200)
    assert abs(fin.sum() / (32 * 32) - 1.0) < 1e-3


@pytest.mark.skipif(not R.reference_available(), reason="/root/reference absent (GPU box)")
def test_live_reference_mrt_py():
    rho, u, fin = R.exec_reference_mrt_py(24, 28, 100, 12)
    p = O.Params(24, 28, Re=100, collision="SRT")
    r2, u2, f2 = O.run(p, 12, semantics="A")
    assert np.array_equal(rho, r2) and np.array_equal(u, u2) and np.array_equal(fin, f2)


@pytest.mark.skipif(not (R.ref_functions_built()), reason="oracle/_ref not built")
def test_live_compiled_allfunc():
    F = R.load_ref_functions("functions")
    nx, ny = 28, 36
    f0 = O.random_state(nx, ny, seed=99)
    F.set_omega(0.08, 100, ny)
    rho, u, fin, feq = F.allfunc(np.ones((nx, ny)), np.zeros((2, nx, ny)), f0.copy(), np.zeros((9, nx, ny)))
    p = O.Params(nx, ny, Re=100, collision="SRT")
    st = O.StateC.initial(p, f0)
    O.step_C(st, p)
    assert np.abs(st.rho - rho).max() <= 4.5e-16 and np.abs(st.u - u).max() <= 1e-16
    assert np.abs(st.fin - np.asarray(fin))[:, 1:-1, 1:-1].max() <= 2.3e-16


def test_oracle_against_ghia_re100():
    """The reference's only fixture for this path is GhiaData.csv (Re = 100 columns are the clean ones):
    the oracle's C-MRT cavity at Re 100, run to its steady state on the CPU, reproduces the centre-line velocities of
    Ghia et al. -- 96 x 96: 0.0093 / 0.0057 of uLB measured, the 128 x 128 figure of SURVEY.md 0-2 is 0.0091 / 0.0068."""
    p = O.Params(96, 96, Re=100, collision="MRT")
    rho, u, f = O.run_fast(p, 20000)
    again = O.run_fast(p, 20500)[1]
    assert np.abs(again - u).max() / 0.08 < 1e-4          # steady
    ex, ey = O.ghia_errors(u, 0.08, O.load_ghia())
    assert ex < 0.012 and ey < 0.009, (ex, ey)
