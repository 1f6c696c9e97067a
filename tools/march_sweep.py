"""Sweep of the two-step kernels at one cavity size: marching variants x segment heights, tiles, one-step.

usage: python tools/march_sweep.py [nx ny [batch]] [--quick]"""
import sys
import torch
sys.path.insert(0, ".")
import latticeboltzmannsimulations_b200 as L

args = [a for a in sys.argv[1:] if not a.startswith("--")]
nx = int(args[0]) if len(args) > 0 else 4096
ny = int(args[1]) if len(args) > 1 else nx
batch = int(args[2]) if len(args) > 2 else 1
quick = "--quick" in sys.argv
steps = 200 if nx * ny * batch > 4e6 else 2000


def run(dt, tuning, coll="MRT", turb=False):
    try:
        with L.CavitySolver(nx, ny, batch, dt, coll, turb, tuning=tuning) as s:
            s.set_reynolds(5000); s.init_equilibrium(); s.step(11, write_macros=False); s.sync()
            best = 1e9
            st = torch.cuda.current_stream().cuda_stream
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); s.step(steps, write_macros=False, stream=st); e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / steps)
        print("%-8s %-4s turb=%d %-36s %.4f ms/step %9.0f MLUPS" % (dt, coll, turb, tuning, best, batch * nx * ny / best / 1e3), flush=True)
    except Exception as e:
        print(dt, tuning, "failed:", e, flush=True)


print("size %d x %d x %d" % (nx, ny, batch))
for dt in ("float64", "float32"):
    run(dt, {"two_step": 0})
    run(dt, {"march": 0})
    nvar = 8
    for v in range(nvar):
        for h in ((0,) if quick else (16, 32, 64)):
            run(dt, {"march_variant": v, "march_h": h, "march_min_nodes": 0})
    for coll, turb in (("SRT", False), ("SRT", True), ("MRT", True)):
        run(dt, {"two_step": 0}, coll, turb)
        run(dt, {"march_min_nodes": 0}, coll, turb)
