"""Generate the committed golden fixtures in this directory.  Run from the repo root in the BUILD container:

    python oracle/build_ref.py && python tests/golden/make_golden.py

Sources of truth, strongest first:
  ref_A_*.npz        outputs of the REAL reference script /root/reference/MRT.py (exec'd under import stubs,
                     oracle/ref_harness.py) -- pins oracle semantics 'A' (expected difference: exactly 0.0).
  ref_allfunc_*.npz  outputs of the COMPILED reference functions.allfunc (functions.pyx:45-222, built by
                     oracle/build_ref.py) for one step from a seeded random state, run SERIALLY
                     (OMP_NUM_THREADS has no effect on the hard-coded 4 threads, but one step from a given
                     state is race-free on all nodes that the oracle is compared on: rho, u, feq everywhere,
                     fin on non-wall nodes) -- pins moments/overrides/equilibrium/SRT/push of semantics 'C'.
  ghia_re100.json    Re = 100 centre-line stations extracted from the reference fixture GhiaData.csv.
  C_*.npz            outputs of the oracle itself (semantics 'C', push form) -- regression vectors for the
                     pieces nothing executable pins (MRT relaxation, funBC); the CUDA parity tests and the
                     CPU tests compare against these on the GPU box, where /root/reference does not exist.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import lbm_oracle as O          # noqa: E402
from oracle import ref_harness as R         # noqa: E402


def save(name, **arrs):
    np.savez_compressed(os.path.join(HERE, name), **arrs)
    print("wrote", name, {k: np.asarray(v).shape for k, v in arrs.items()})


def main():
    assert R.reference_available(), "needs /root/reference"
    # --- Ghia Re = 100 ---------------------------------------------------------------------------
    raw = np.genfromtxt(os.path.join(R.REF_DIR, "GhiaData.csv"), delimiter=",")[6:23, 1:]   # MRT.py:104
    with open(os.path.join(HERE, "ghia_re100.json"), "w") as fh:
        json.dump({"source": "GhiaData.csv rows 7-23, columns Y, Ux(Re=100), X, Uy(Re=100)",
                   "Y": raw[:, 0].tolist(), "Ux": raw[:, 1].tolist(),
                   "X": raw[:, 9].tolist(), "Uy": raw[:, 10].tolist()}, fh, indent=1)
    # --- real MRT.py -----------------------------------------------------------------------------
    for nx, ny, Re, n in [(32, 32, 100, 25), (40, 24, 400, 60)]:
        rho, u, fin = R.exec_reference_mrt_py(nx, ny, Re, n)
        save("ref_A_%dx%d_Re%d_N%d.npz" % (nx, ny, Re, n), rho=rho, u=u, fin=fin,
             meta=np.array([nx, ny, Re, n, 0.08]))
    # --- compiled allfunc, one step ----------------------------------------------------------------
    F = R.load_ref_functions("functions")
    nx, ny, Re = 32, 24, 100
    f0 = O.random_state(nx, ny, seed=1234)
    F.set_omega(0.08, Re, ny)
    rho, u, fin, feq = F.allfunc(np.ones((nx, ny)), np.zeros((2, nx, ny)), f0.copy(), np.zeros((9, nx, ny)))
    save("ref_allfunc_%dx%d_Re%d.npz" % (nx, ny, Re), rho=np.array(rho), u=np.array(u), fin=np.array(fin),
         feq=np.array(feq), meta=np.array([nx, ny, Re, 1, 0.08]))
    # --- oracle regression vectors (semantics C) ---------------------------------------------------
    cases = [("MRT", 0, 32, 32, 100, 1), ("MRT", 0, 32, 32, 100, 10), ("MRT", 0, 32, 32, 100, 100),
             ("MRT", 0, 40, 24, 400, 100), ("SRT", 0, 32, 32, 100, 100), ("TRT", 0, 32, 32, 100, 100),
             ("MRT", 1, 32, 32, 1000, 100), ("SRT", 1, 32, 32, 1000, 100)]
    for coll, turb, nx, ny, Re, n in cases:
        p = O.Params(nx, ny, Re=Re, collision=coll, turb=turb)
        rho, u, fin = O.run(p, n, semantics="C", form="push")
        save("C_%s_turb%d_%dx%d_Re%d_N%d.npz" % (coll, turb, nx, ny, Re, n), rho=rho, u=u, fin=fin,
             meta=np.array([nx, ny, Re, n, 0.08]))
    # random-state single steps (every boundary class exercised with a generic state)
    for coll in ("MRT", "SRT"):
        nx, ny, Re = 24, 20, 100
        p = O.Params(nx, ny, Re=Re, collision=coll)
        f0 = O.random_state(nx, ny, seed=1234)
        rho, u, fin = O.run(p, 3, semantics="C", fin0=f0, form="push")
        save("C_%s_random_%dx%d_N3.npz" % (coll, nx, ny), rho=rho, u=u, fin=fin, f0=f0,
             meta=np.array([nx, ny, Re, 3, 0.08]))


if __name__ == "__main__":
    main()
