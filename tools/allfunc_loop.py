"""Time the time loop of MRT_cython.py:453 over the functions shim (384^2, the reference's Cython grid):
state kept on the device between calls (the returned fin handed back) against a fresh upload every call."""
import sys
import time
import numpy as np
sys.path.insert(0, ".")
import latticeboltzmannsimulations_b200.functions as F
n, calls = 384, 300
vel = np.zeros((2, n, n)); vel[0, :, 0] = 0.08
F.set_omega(0.08, 3200, n)
for resident in (True, False):
    fin = F.equ(np.ones((n, n)), vel[0], vel[1])
    rho = np.sum(fin, axis=0); u = np.zeros((2, n, n)); feq = fin.copy()
    for i in range(calls + 20):
        if i == 20:
            t = time.perf_counter()
        rho, u, fin, feq = F.allfunc(rho, u, fin if resident else fin.copy(), feq)
    dt = (time.perf_counter() - t) / calls
    print("allfunc loop %dx%d, %s: %.3f ms per call = %.1f MLUPS" % (
        n, n, "state resident (fin handed back)" if resident else "upload every call (fin.copy())", dt * 1e3, n * n / dt / 1e6), flush=True)
