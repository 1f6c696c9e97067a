"""Smallest run that touches every kernel family (for one compute-sanitizer memcheck pass)."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
import latticeboltzmannsimulations_b200 as L
for dt in ("float64", "float32"):
    for coll, turb in (("MRT", False), ("SRT", True), ("TRT", False)):
        for (nx, ny) in ((67, 35), (130, 9)):
            rho, u, f = L.run_cavity(nx, ny, 400, steps=6, dtype=dt, collision=coll, turb=turb, return_f=True)
            assert np.isfinite(f).all()
    rho, u, f = L.run_cavity(300, 40, 400, steps=6, dtype=dt, return_f=True, engine="tma")
    f_final, u_final, feq0, Re = L.datagen([100.0, 300.0, 900.0], 45, 37, steps=5, dtype=dt)
    with L.CavitySolver(64, 48, 2, dt) as s:
        s.set_reynolds([100, 200]); s.init_equilibrium(); s.step(40); s.mean_u(); s.set_active([1, 0]); s.step(3)
        s.diagnostics(); s.macros(current=True); s.download_f()
    with L.CavitySolver(50, 30, 1, dt, y0=10, ny_local=12) as s:      # a y-strip with both regions
        s.set_reynolds(100); s.init_equilibrium()
        for _ in range(3):
            s.step_region(1); s.step_region(2); s.swap()
        s.download_f()
os.environ["LBM_B200_VEC_F64"] = "2"; os.environ["LBM_B200_VEC_F32"] = "2"
for dt in ("float64", "float32"):
    L.run_cavity(67, 35, 400, steps=6, dtype=dt, return_f=True)
print("MEMCHECK_CASE_DONE")
