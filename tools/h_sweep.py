import sys
import torch
sys.path.insert(0, ".")
import latticeboltzmannsimulations_b200 as L
nx = ny = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for dt in ("float64", "float32"):
    for tun in [{"two_step": 0}, {"slide": 0}] + [{"slide_h": h} for h in (14, 30, 46, 62, 94, 126, 254, 64, 128)]:
        with L.CavitySolver(nx, ny, 1, dt, "MRT", tuning=tun) as s:
            s.set_reynolds(5000); s.init_equilibrium(); s.step(11, write_macros=False); s.sync()
            best = 1e9
            st = torch.cuda.current_stream().cuda_stream
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); s.step(200, write_macros=False, stream=st); e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / 200)
        print("%s %s: %.4f ms/step %.0f MLUPS" % (dt, tun, best, nx * ny / best / 1e3), flush=True)
