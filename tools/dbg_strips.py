import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import latticeboltzmannsimulations_b200 as L
from test_gpu_strips import _run_strips_one_gpu
for (nx, ny, world, steps, dt) in [(256, 192, 2, 80, "float64"), (256, 192, 2, 7, "float64"), (256, 192, 2, 3, "float64"), (256, 192, 4, 6, "float64"), (256, 192, 2, 80, "float32")]:
    want = L.run_cavity(nx, ny, 1000, steps=steps, dtype=dt, return_f=True)
    for split in (True, False):
        got = _run_strips_one_gpu(nx, ny, world, 1000, steps, dt, split)
        d = np.abs(got[2] - want[2])
        bad = np.argwhere(d.max(axis=0) > 0)
        print(nx, ny, world, steps, dt, "split" if split else "all", "two-step passes", got[3], "max diff", d.max(),
              "bad nodes", len(bad), "x range", (bad[:, 0].min(), bad[:, 0].max()) if len(bad) else None,
              "y range", (bad[:, 1].min(), bad[:, 1].max()) if len(bad) else None, flush=True)
