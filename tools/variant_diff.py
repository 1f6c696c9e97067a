"""Where does a two-step kernel differ from the one-step kernels (it must not)?  (development aid)"""
import sys
import numpy as np
sys.path.insert(0, ".")
import latticeboltzmannsimulations_b200 as L

def run(nx, ny, dt, coll, turb, steps, tuning):
    with L.CavitySolver(nx, ny, 1, dt, coll, turb, tuning=tuning) as s:
        s.set_reynolds(1000); s.init_equilibrium(); s.step(steps, write_macros=True)
        rho, u = s.macros()
        return rho, u, s.download_f()

for (nx, ny, dt, coll, turb, steps) in [(150, 70, "float32", "SRT", True, 2), (150, 70, "float32", "SRT", True, 40),
                                        (150, 70, "float32", "TRT", True, 40), (150, 70, "float32", "MRT", False, 40),
                                        (1100, 600, "float32", "SRT", False, 2), (1100, 600, "float32", "SRT", False, 70),
                                        (1001, 640, "float32", "MRT", False, 2), (1001, 640, "float32", "MRT", False, 70),
                                        (200, 90, "float64", "MRT", False, 2), (200, 90, "float64", "MRT", False, 70),
                                        (131, 77, "float32", "SRT", False, 2), (131, 77, "float32", "SRT", False, 70),
                                        (96, 64, "float32", "MRT", True, 2), (96, 64, "float32", "MRT", True, 70),
                                        (777, 801, "float64", "TRT", False, 3), (777, 801, "float64", "TRT", False, 11),
                                        (777, 801, "float64", "SRT", False, 11), (777, 801, "float64", "MRT", False, 11),
                                        (777, 801, "float32", "TRT", False, 11), (1000, 640, "float64", "TRT", False, 11),
                                        (930, 700, "float64", "SRT", True, 11), (1030, 610, "float32", "MRT", True, 11)]:
    a = run(nx, ny, dt, coll, turb, steps, {"two_step": 0})
    for tun in ({"slide_min_nodes": 0}, {"slide_min_nodes": 0, "slide_h": 37}, {"slide": 0, "two_step_min_nodes": 0}):
        b = run(nx, ny, dt, coll, turb, steps, tun)
        d = np.abs(a[2] - b[2]).max(axis=0)
        bad = np.argwhere(d > 0)
        msg = "equal" if len(bad) == 0 else "%d nodes differ, max %.3e, x in [%d,%d], y in [%d,%d], first %s" % (
            len(bad), d.max(), bad[:, 0].min(), bad[:, 0].max(), bad[:, 1].min(), bad[:, 1].max(), bad[:5].tolist())
        print(nx, ny, dt, coll, turb, steps, tun, "->", msg, "| macros equal:", np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), flush=True)
