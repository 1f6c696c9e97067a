"""Small-cavity (launch-latency-bound) throughput with and without CUDA graphs (development aid)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
import latticeboltzmannsimulations_b200 as L
for n in (128, 384, 640, 1024):
    ref = None
    for g in ("0", "1"):
        os.environ["LBM_B200_GRAPH"] = g
        for dt in ("float64", "float32"):
            with L.CavitySolver(n, n, 1, dt, "MRT") as s:
                s.set_reynolds(3200); s.init_equilibrium(); s.step(100, write_macros=False); s.sync()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                steps = 5000
                e0.record(); s.step(steps, write_macros=False, stream=torch.cuda.current_stream().cuda_stream); e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
                f = s.download_f()
                key = (n, dt)
                if g == "0": ref = ref or {}; ref[key] = f
                same = np.array_equal(f, ref[key]) if ref and key in ref else None
                print("n=%4d %s graph=%s: %.2f us/step %.0f MLUPS  bitwise==nograph: %s" % (n, dt, g, ms * 1e3, n * n / ms / 1e3, same), flush=True)
