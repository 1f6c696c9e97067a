import sys
import torch
sys.path.insert(0, ".")
import latticeboltzmannsimulations_b200 as L
for (nx, ny) in ((32768, 4096), (32768, 16384)):
    for dt in ("float64",):
        for tun in [{}] + [{"slide_h": h} for h in (30, 46, 62, 94, 126, 162, 254)]:
            with L.CavitySolver(nx, ny, 1, dt, "MRT", tuning=tun) as s:
                s.set_reynolds(5000); s.init_equilibrium(); s.step(5, write_macros=False); s.sync()
                best = 1e9
                st = torch.cuda.current_stream().cuda_stream
                steps = 20
                for rep in range(2):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); s.step(steps, write_macros=False, stream=st); e1.record()
                    torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1) / steps)
            print("%dx%d %s %s: %.4f ms/step %.0f MLUPS" % (nx, ny, dt, tun, best, nx * ny / best / 1e3), flush=True)
