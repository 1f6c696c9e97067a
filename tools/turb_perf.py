"""Throughput with the Smagorinsky closure (the reference's default GPU mode is SRT + turb, fp32)."""
import sys
import torch
sys.path.insert(0, ".")
import latticeboltzmannsimulations_b200 as L
for dt, bpn in (("float64", 176), ("float32", 88)):
    for coll in ("SRT", "MRT"):
        with L.CavitySolver(4096, 4096, 1, dt, coll, turb=True) as s:
            s.set_reynolds(5000); s.init_equilibrium(); s.step(10, write_macros=False); s.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); s.step(200, write_macros=False, stream=torch.cuda.current_stream().cuda_stream); e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 200
            ml = 4096 * 4096 / ms / 1e3
            print("turb=1 4096^2 %s %s: %.4f ms/step %.0f MLUPS %.0f GB/s (%d B/node)" % (dt, coll, ms, ml, ml * bpn / 1e3, bpn), flush=True)
