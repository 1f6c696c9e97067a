"""Print how closely the reference's own CUDA kernels (oracle/_ref/libref_kernels.so) agree with the oracle and with
the product (development aid; the assertion lives in tests/test_gpu_reference_kernels.py)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import latticeboltzmannsimulations_b200 as L
from oracle import lbm_oracle as O, ref_harness as R
for dtype in ("float32", "float64"):
  print("reference kernels compiled as", dtype)
  for (nx, ny, Re, steps) in ((64, 64, 100.0, 60), (96, 64, 1000.0, 120), (128, 96, 3200.0, 400)):
    for coll in ("MRT", "SRT", "TRT"):
        for turb in (0, 1):
            ref = R.run_reference_kernels(nx, ny, Re, steps, coll, turb, dtype=dtype)
            p = O.Params(nx, ny, Re=Re, collision=coll, turb=turb)
            want = O.run(p, steps, form="push")
            got = L.run_cavity(nx, ny, Re, steps=steps, collision=coll, dtype=dtype, turb=bool(turb), return_f=True)
            eo = [float(np.abs(a - b).max()) for a, b in zip(ref, want)]
            ep = [float(np.abs(a - b).max()) for a, b in zip(ref, got)]
            print("%dx%d Re=%g N=%d %s turb=%d: ref-vs-oracle rho %.2e u %.2e f %.2e | ref-vs-product rho %.2e u %.2e f %.2e" % (
                nx, ny, Re, steps, coll, turb, eo[0], eo[1] / 0.08, eo[2], ep[0], ep[1] / 0.08, ep[2]), flush=True)
