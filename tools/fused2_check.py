"""Temporal-blocking kernel: bit-identity against the one-step path and throughput (development aid)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
import latticeboltzmannsimulations_b200 as L
ok = True
for dt in ("float64", "float32"):
    for coll in ("MRT", "SRT"):
        for (nx, ny, n) in ((2111, 2003, 9), (4096, 1100, 6)):
            res = []
            for fused in ("0", "1"):
                os.environ["LBM_B200_FUSED2"] = fused
                with L.CavitySolver(nx, ny, 1, dt, coll) as s:
                    s.set_reynolds(5000); s.init_equilibrium()
                    s.step(1); s.step(n, write_macros=True)
                    res.append((s.macros(), s.download_f(), s.counters()))
            same = np.array_equal(res[0][1], res[1][1]) and all(np.array_equal(a, b) for a, b in zip(res[0][0], res[1][0]))
            print(dt, coll, nx, ny, n, "bitwise equal:", same, "launches", res[0][2], res[1][2], flush=True)
            ok = ok and same
print("FUSED2_OK" if ok else "FUSED2_MISMATCH")
for dt, bpn in (("float64", 144), ("float32", 72)):
    for fused in ("0", "1"):
        os.environ["LBM_B200_FUSED2"] = fused
        for (nx, ny) in ((4096, 4096), (16384, 8192)):
            with L.CavitySolver(nx, ny, 1, dt, "MRT") as s:
                s.set_reynolds(5000); s.init_equilibrium(); s.step(11, write_macros=False); s.sync()
                steps = 400 if nx == 4096 else 60
                best = 1e9
                for rep in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); s.step(steps, write_macros=False, stream=torch.cuda.current_stream().cuda_stream); e1.record()
                    torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1) / steps)
                ml = nx * ny / best / 1e3
                print("fused2=%s %5dx%-5d %s MRT: %.4f ms/step %.0f MLUPS (%.0f GB/s at %d B/node-step)" % (fused, nx, ny, dt, best, ml, ml * bpn / 1e3, bpn), flush=True)
