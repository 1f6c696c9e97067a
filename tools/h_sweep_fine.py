import sys
import torch
sys.path.insert(0, ".")
import latticeboltzmannsimulations_b200 as L
for (nx, ny, batch, steps) in ((4096, 4096, 1, 200), (32768, 4096, 1, 30), (2048, 2048, 1, 400), (384, 384, 32, 400)):
    for dt in ("float64", "float32"):
        out = []
        for h in (0, 14, 18, 22, 26, 30, 34, 38, 42, 46, 54):
            with L.CavitySolver(nx, ny, batch, dt, "MRT", tuning={"slide_h": h}) as s:
                s.set_reynolds(5000); s.init_equilibrium(); s.step(11, write_macros=False); s.sync()
                best = 1e9
                st = torch.cuda.current_stream().cuda_stream
                for rep in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); s.step(steps, write_macros=False, stream=st); e1.record()
                    torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1) / steps)
            out.append("%d:%.0f" % (h, batch * nx * ny / best / 1e3))
        print("%dx%dx%d %s: %s" % (nx, ny, batch, dt, "  ".join(out)), flush=True)
