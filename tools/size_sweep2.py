"""Which kernel family wins at which size?  one-step / tiles / sliding window, fp64 and fp32, MRT and SRT+turb."""
import sys
import torch
sys.path.insert(0, ".")
import latticeboltzmannsimulations_b200 as L

def run(nx, ny, batch, dt, coll, turb, tuning, steps):
    with L.CavitySolver(nx, ny, batch, dt, coll, turb, tuning=tuning) as s:
        s.set_reynolds(1000); s.init_equilibrium(); s.step(65, write_macros=False); s.sync()
        best = 1e9
        st = torch.cuda.current_stream().cuda_stream
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); s.step(steps, write_macros=False, stream=st); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / steps)
    return batch * nx * ny / best / 1e3, best

SIZES = [(128, 128, 1), (192, 192, 1), (384, 384, 1), (512, 512, 1), (640, 640, 1), (768, 768, 1), (1024, 1024, 1), (2048, 2048, 1),
         (384, 384, 32), (384, 384, 256), (192, 192, 64)]
if len(sys.argv) > 1 and sys.argv[1] == "mid":       # around the sliding-window threshold
    SIZES = [(768, 768, 1), (1024, 1024, 1), (1280, 1280, 1), (1536, 1536, 1), (2048, 2048, 1), (1024, 512, 1), (384, 384, 8), (384, 384, 16)]
for (nx, ny, batch) in SIZES:
    steps = 4000 if nx * ny * batch < 1e6 else (1000 if nx * ny * batch < 8e6 else 200)
    for dt in ("float64", "float32"):
        out = []
        for name, tun in (("one", {"two_step": 0}), ("tile", {"slide": 0, "two_step_min_nodes": 0}),
                          ("slide", {"slide_min_nodes": 0}), ("slide14", {"slide_min_nodes": 0, "slide_h": 14})):
            try:
                m, ms = run(nx, ny, batch, dt, "MRT", False, tun, steps)
                out.append("%s %.0f (%.2f us)" % (name, m, ms * 1e3))
            except Exception as e:
                out.append("%s failed %s" % (name, e))
        print("%dx%dx%d %s MRT: %s" % (nx, ny, batch, dt, " | ".join(out)), flush=True)
