import sys
import numpy as np
sys.path.insert(0, ".")
import latticeboltzmannsimulations_b200 as L
from oracle import lbm_oracle as O
nx, ny, n = 96, 64, 150
for Re in (400.0, 100.0):
  for coll in ("SRT", "TRT", "MRT"):
    for turb in (0, 1):
        for start in ("eq", "rand"):
            f0 = None if start == "eq" else O.random_state(nx, ny, seed=11)
            p = O.Params(nx, ny, Re=Re, collision=coll, turb=turb)
            want = O.run(p, n, fin0=f0, form="push")
            for dt in ("float32", "float64"):
                got = L.run_cavity(nx, ny, Re, steps=n, collision=coll, dtype=dt, turb=bool(turb), return_f=True,
                                   f0=None if f0 is None else f0.astype(dt))
                e = (np.abs(got[0]-want[0]).max(), np.abs(got[1]-want[1]).max()/0.08, np.abs(got[2]-want[2]).max())
                print("Re=%g %s turb=%d %s %s: rho %.2e u/uLB %.2e f %.2e" % (Re, coll, turb, start, dt, *e), flush=True)
