"""Segment height of the sliding-window kernel on launches with few CTAs (mid-size cavities, small batches)."""
import sys
import torch
sys.path.insert(0, ".")
import latticeboltzmannsimulations_b200 as L

def run(nx, ny, batch, dt, tuning, steps):
    with L.CavitySolver(nx, ny, batch, dt, "MRT", False, tuning=tuning) as s:
        s.set_reynolds(1000); s.init_equilibrium(); s.step(33, write_macros=False); s.sync()
        best = 1e9
        st = torch.cuda.current_stream().cuda_stream
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); s.step(steps, write_macros=False, stream=st); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / steps)
    return batch * nx * ny / best / 1e3

for (nx, ny, batch) in [(1024, 1024, 1), (1280, 1280, 1), (1536, 1536, 1), (2048, 2048, 1), (3072, 3072, 1), (2048, 1024, 1),
                        (384, 384, 8), (384, 384, 16), (384, 384, 32), (640, 640, 4)]:
    steps = 1000 if nx * ny * batch < 8e6 else 300
    for dt in ("float64", "float32"):
        out = ["auto %.0f" % run(nx, ny, batch, dt, {"slide_min_nodes": 0}, steps)]
        for h in (10, 14, 18, 22, 26, 30, 34):
            out.append("%d: %.0f" % (h, run(nx, ny, batch, dt, {"slide_min_nodes": 0, "slide_h": h}, steps)))
        print("%dx%dx%d %s: %s" % (nx, ny, batch, dt, " | ".join(out)), flush=True)
