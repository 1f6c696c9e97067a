"""The reference's OWN CUDA kernels -- the funRT (SRT / TRT / MRT, with and without Smagorinsky) and funBC strings of
MRT_GPU.py:336-699, compiled unmodified apart from the parameter splice (oracle/build_ref_kernels.py) -- executed on
this GPU, against the fp64 oracle and against the product's fp32 path.

This is the executable pin for the two pieces of semantics "C" that nothing on the CPU can pin: the MRT relaxation
(MRT_GPU.py:633-655) and the wall rule funBC (:664-699, x-block then y-block, lid density, corner populations).
The reference kernels are fp32 with double-literal mixed arithmetic, so agreement is at fp32 round-off (measured a few
1e-6; asserted <= 5e-5 of the density / lid-speed scale), while any semantic slip -- a wrong corner order, lid formula,
push bound or equilibrium moment -- shows up at 1e-3 .. 1e-2.

A second build of the SAME kernel text with every `float` replaced by `double` (oracle/build_ref_kernels.py,
libref_kernels_f64.so) runs the reference's algorithm in the precision of the oracle and of the product's fp64 path:
there the pin is at fp64 round-off -- the north star's 1e-12.
"""
import numpy as np
import pytest

from oracle import lbm_oracle as O
from oracle import ref_harness as R

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not R.ref_kernels_built(), reason="oracle/_ref/libref_kernels.so not built")]
TOL = 5e-5


def _err(a, b, uLB=0.08):
    return (float(np.abs(a[0] - b[0]).max()), float(np.abs(a[1] - b[1]).max() / uLB), float(np.abs(a[2] - b[2]).max()))


@pytest.mark.parametrize("turb", [0, 1])
@pytest.mark.parametrize("coll", ["MRT", "SRT", "TRT"])
@pytest.mark.parametrize("nx,ny,Re,steps", [(64, 64, 100.0, 60), (96, 64, 1000.0, 120)])
def test_oracle_and_product_against_reference_cuda_kernels(coll, turb, nx, ny, Re, steps):
    import latticeboltzmannsimulations_b200 as L
    ref = R.run_reference_kernels(nx, ny, Re, steps, coll, turb)
    p = O.Params(nx, ny, Re=Re, collision=coll, turb=turb)
    want = O.run(p, steps, semantics="C", form="push")
    e_oracle = _err(ref, want)
    assert max(e_oracle) <= TOL, ("reference kernels vs oracle", coll, turb, e_oracle)
    got = L.run_cavity(nx, ny, Re, steps=steps, collision=coll, dtype="float32", turb=bool(turb), return_f=True)
    e_prod = _err(ref, got)
    assert max(e_prod) <= TOL, ("reference kernels vs product fp32", coll, turb, e_prod)


TOL64 = 1e-12            # product fp64 vs the reference kernels in double (measured <= 4.5e-14, profiles/r02_reference_kernels_check.txt)
TOL64_ORACLE = 1e-15     # oracle vs the reference kernels in double: measured 0.0 -- the same bits -- in all 18 cases


@pytest.mark.skipif(not R.ref_kernels_built("float64"), reason="oracle/_ref/libref_kernels_f64.so not built")
@pytest.mark.parametrize("turb", [0, 1])
@pytest.mark.parametrize("coll", ["MRT", "SRT", "TRT"])
@pytest.mark.parametrize("nx,ny,Re,steps", [(64, 64, 100.0, 60), (96, 64, 1000.0, 120), (128, 96, 3200.0, 400)])
def test_fp64_pin_against_reference_cuda_kernels_in_double(coll, turb, nx, ny, Re, steps):
    """MRT relaxation (MRT_GPU.py:633-655), funBC (:664-699), TRT and the Smagorinsky closure at fp64 round-off: the
    reference's own kernel text compiled in double against the oracle (literal two-kernel form) and against the product's
    fp64 path (fused pull form, hand-factored moment transform, explicit FMAs)."""
    import latticeboltzmannsimulations_b200 as L
    ref = R.run_reference_kernels(nx, ny, Re, steps, coll, turb, dtype="float64")
    p = O.Params(nx, ny, Re=Re, collision=coll, turb=turb)
    want = O.run(p, steps, semantics="C", form="push")
    e_oracle = _err(ref, want)
    assert max(e_oracle) <= TOL64_ORACLE, ("reference kernels (double) vs oracle", coll, turb, e_oracle)
    got = L.run_cavity(nx, ny, Re, steps=steps, collision=coll, dtype="float64", turb=bool(turb), return_f=True)
    e_prod = _err(ref, got)
    assert max(e_prod) <= TOL64, ("reference kernels (double) vs product fp64", coll, turb, e_prod)
