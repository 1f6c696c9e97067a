"""Sweep kernel families / tile configurations at one grid size (development aid)."""
import os, sys
import torch
sys.path.insert(0, ".")
import latticeboltzmannsimulations_b200 as L
nx, ny = int(sys.argv[1]), int(sys.argv[2])
which = sys.argv[3] if len(sys.argv) > 3 else "all"
configs = [("ldg", 1, 0, 1), ("ldg", 2, 0, 1), ("ldg", 4, 0, 1)]
if which == "all":
    configs += [("tma", 1, 0, 1), ("tma", 1, 1, 2)]
for dt in ("float64", "float32"):
    for coll in ("MRT", "SRT"):
        for eng, vec, var, ctas in configs:
            if dt == "float64" and vec == 4:
                continue
            os.environ["LBM_B200_ENGINE"] = eng
            os.environ["LBM_B200_VEC_F64"] = str(vec)
            os.environ["LBM_B200_VEC_F32"] = str(vec)
            os.environ["LBM_B200_TMA_VARIANT"] = str(var)
            os.environ["LBM_B200_TMA_CTAS"] = str(ctas)
            with L.CavitySolver(nx, ny, 1, dt, coll) as s:
                s.set_reynolds(5000); s.init_equilibrium(); s.step(10, write_macros=False); s.sync()
                steps = max(10, int(3e9 / (nx * ny)))
                best = 1e9
                for rep in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); s.step(steps, write_macros=False, stream=torch.cuda.current_stream().cuda_stream); e1.record()
                    torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1) / steps)
                bpn = 144 if dt == "float64" else 72
                print("%s vec%d v%d c%d %5dx%-5d %s %s: %.4f ms  %.0f MLUPS  %.0f GB/s" % (
                    s.engine, vec, var, ctas, nx, ny, dt, coll, best, nx * ny / best / 1e3, nx * ny / best / 1e3 * bpn / 1e3), flush=True)
