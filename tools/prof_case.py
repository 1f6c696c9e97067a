"""One configuration, a few steps: the command profiled by ncu (tuning comes from LBM_B200_TUNING).
usage: python tools/prof_case.py nx ny dtype [steps [collision [turb]]]"""
import sys
import torch
sys.path.insert(0, ".")
import latticeboltzmannsimulations_b200 as L
nx, ny, dt = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
coll = sys.argv[5] if len(sys.argv) > 5 else "MRT"
turb = bool(int(sys.argv[6])) if len(sys.argv) > 6 else False
with L.CavitySolver(nx, ny, 1, dt, coll, turb) as s:
    s.set_reynolds(5000); s.init_equilibrium(); s.step(1, write_macros=False); s.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.step(steps, write_macros=False, stream=torch.cuda.current_stream().cuda_stream); e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print("%dx%d %s %s turb=%d: %.4f ms/step %.0f MLUPS launches=%d" % (nx, ny, dt, coll, turb, ms, nx * ny / ms / 1e3, s.counters()[1]))
