import sys
import torch
sys.path.insert(0, ".")
import latticeboltzmannsimulations_b200 as L
def run(nx, ny, dt, coll, turb, tuning, steps):
    with L.CavitySolver(nx, ny, 1, dt, coll, turb, tuning=tuning) as s:
        s.set_reynolds(1000); s.init_equilibrium(); s.step(65, write_macros=False); s.sync()
        best = 1e9
        st = torch.cuda.current_stream().cuda_stream
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); s.step(steps, write_macros=False, stream=st); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / steps)
    return best * 1e3
for n in (192, 384, 512, 640, 768, 1024, 1400):
    for dt, key, vals in (("float32", "vec_f32", (1, 2, 4)), ("float64", "vec_f64", (1, 2))):
        for coll, turb in (("MRT", False), ("SRT", True)):
            out = ["%s=%d: %.2f us" % (key, v, run(n, n, dt, coll, turb, {key: v, "two_step": 0}, 3000)) for v in vals]
            print("%d^2 %s %s turb=%d one-step: %s" % (n, dt, coll, turb, " | ".join(out)), flush=True)
