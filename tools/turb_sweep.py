import sys
import torch
sys.path.insert(0, ".")
import latticeboltzmannsimulations_b200 as L
def run(nx, ny, dt, coll, turb, tuning, steps):
    with L.CavitySolver(nx, ny, 1, dt, coll, turb, tuning=tuning) as s:
        s.set_reynolds(1000); s.init_equilibrium(); s.step(65, write_macros=False); s.sync()
        best = 1e9
        st = torch.cuda.current_stream().cuda_stream
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); s.step(steps, write_macros=False, stream=st); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / steps)
    return nx * ny / best / 1e3, best
for (nx, steps) in ((4096, 200), (2048, 500), (1024, 2000), (640, 3000)):
    for dt in ("float64", "float32"):
        for coll, turb in (("SRT", True), ("MRT", True), ("SRT", False)):
            out = []
            for name, tun in (("one", {"two_step": 0}), ("slide", {"slide_min_nodes": 0})):
                m, ms = run(nx, nx, dt, coll, turb, tun, steps)
                out.append("%s %.0f (%.1f us)" % (name, m, ms * 1e3))
            print("%d^2 %s %s turb=%d: %s" % (nx, dt, coll, turb, " | ".join(out)), flush=True)
