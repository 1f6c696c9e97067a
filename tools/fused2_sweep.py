import os, sys
import torch
sys.path.insert(0, ".")
import latticeboltzmannsimulations_b200 as L
os.environ["LBM_B200_FUSED2"] = "1"
names = {0: "64x16/2", 1: "64x8/3", 2: "32x16/3", 3: "64x8/4", 4: "32x16/4", 5: "64x12/3", 6: "32x16/5", 7: "32x8/6", 8: "32x16/1"}
for dt in ("float64", "float32"):
    for tile in (3, 4, 1, 0):
        os.environ["LBM_B200_FUSED2_TILE"] = str(tile)
        try:
            with L.CavitySolver(4096, 4096, 1, dt, "MRT") as s:
                s.set_reynolds(5000); s.init_equilibrium(); s.step(11, write_macros=False); s.sync()
                best = 1e9
                for rep in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); s.step(400, write_macros=False, stream=torch.cuda.current_stream().cuda_stream); e1.record()
                    torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1) / 400)
                print("tile %-8s %s: %.4f ms/step %.0f MLUPS" % (names[tile], dt, best, 4096 * 4096 / best / 1e3), flush=True)
        except Exception as e:
            print("tile", names[tile], dt, "failed:", e, flush=True)
