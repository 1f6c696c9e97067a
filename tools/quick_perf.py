"""Quick device-resident throughput probe (development aid; bench.py is the contract)."""
import sys
import time

import torch

sys.path.insert(0, ".")
import latticeboltzmannsimulations_b200 as L

cases = [(4096, 4096, "float64"), (4096, 4096, "float32"), (384, 384, "float64"), (16384, 8192, "float64")]
if len(sys.argv) > 1:
    cases = [(int(sys.argv[1]), int(sys.argv[2]), sys.argv[3])]
engine = sys.argv[4] if len(sys.argv) > 4 else "auto"
for nx, ny, dt in cases:
    for coll in ("MRT", "SRT"):
        with L.CavitySolver(nx, ny, 1, dt, coll, engine=engine) as s:
            s.set_reynolds(5000)
            s.init_equilibrium()
            s.step(20, write_macros=False)
            s.sync()
            steps = 200 if nx * ny <= 4096 * 4096 else 40
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e9
            for rep in range(3):
                e0.record()
                s.step(steps, write_macros=False, stream=torch.cuda.current_stream().cuda_stream)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / steps)
            mlups = nx * ny / best / 1e3
            bpn = 144 if dt == "float64" else 72
            print("%s %5dx%-5d %s %s: %.4f ms/step  %.0f MLUPS  %.0f GB/s (%.1f%% of 6549)" % (
                s.engine, nx, ny, dt, coll, best, mlups, mlups * bpn / 1e3, mlups * bpn / 1e3 / 65.49), flush=True)
