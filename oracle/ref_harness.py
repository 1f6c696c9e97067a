"""Run the REAL reference code (read where it lies under /root/reference) to pin the oracle.

TEST INFRASTRUCTURE ONLY (see ``lbm_oracle.py``).  Nothing here is imported by the product package.
``/root/reference`` exists only in the build container, never on the GPU box: callers must skip when
``reference_available()`` is False.

* ``exec_reference_mrt_py`` executes the upstream script ``MRT.py`` unmodified except for the run
  constants at its top (it has no callable entry point): the source is read with
  ``utf-8-sig`` (every upstream file starts with a BOM), the literals ``maxIt``, ``Re``,
  ``xsize, ysize`` (``MRT.py:41-45``) and ``SavePlot`` (``:34``) are replaced textually, stub modules
  stand in for the absent ``numexpr`` / ``matplotlib``, and the globals ``rho``, ``u``, ``fin`` are read
  back after ``exec``.
* ``load_ref_functions`` imports the compiled Cython module built by ``oracle/build_ref.py`` into
  ``oracle/_ref/`` (``functions`` = as shipped, 4 OpenMP threads hard-coded ``functions.pyx:69``;
  ``functions_allcores`` = same source with that literal removed so OMP picks every core).
"""
from __future__ import annotations

import importlib
import os
import re
import shutil
import sys
import tempfile
import types

import numpy as np

REF_DIR = os.environ.get("LBM_REFERENCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
REF_BUILD_DIR = os.path.join(HERE, "_ref")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "MRT.py"))


def _stub_modules():
    """numexpr.evaluate -> eval in the caller's frame (elementwise fp64, bit-identical); matplotlib -> no-ops."""
    ne = types.ModuleType("numexpr")

    def evaluate(expr, local_dict=None, global_dict=None):
        frame = sys._getframe(1)
        g = dict(frame.f_globals)
        g.update(frame.f_locals)
        return eval(expr, g)

    ne.evaluate = evaluate
    ne.detect_number_of_threads = lambda: 1
    ne.set_num_threads = lambda n: None

    class _Anything:
        def __getattr__(self, name):
            return _Anything()

        def __call__(self, *a, **k):
            return _Anything()

    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    pyplot = types.ModuleType("matplotlib.pyplot")
    pyplot.__getattr__ = lambda name: _Anything()
    mpl.pyplot = pyplot
    return {"numexpr": ne, "matplotlib": mpl, "matplotlib.pyplot": pyplot}


def exec_reference_mrt_py(nx: int, ny: int, Re: float, steps: int):
    """Execute /root/reference/MRT.py for ``steps`` iterations; returns (rho, u, fin) from its globals."""
    if not reference_available():
        raise FileNotFoundError(REF_DIR)
    with open(os.path.join(REF_DIR, "MRT.py"), encoding="utf-8-sig") as fh:
        src = fh.read()

    def sub(pattern, repl):
        nonlocal src
        src, n = re.subn(pattern, repl, src, count=1, flags=re.M)
        assert n == 1, pattern

    sub(r"^maxIt = \d+", "maxIt = %d" % steps)
    sub(r"^Re    = [\d.]+", "Re    = %r" % float(Re))
    sub(r"^xsize, ysize = \d+, \d+", "xsize, ysize = %d, %d" % (nx, ny))
    sub(r"^SavePlot = True", "SavePlot = False")
    saved = {k: sys.modules.get(k) for k in ("numexpr", "matplotlib", "matplotlib.pyplot")}
    sys.modules.update(_stub_modules())
    cwd = os.getcwd()
    scratch = tempfile.mkdtemp(prefix="mrt_ref_")
    try:
        shutil.copy(os.path.join(REF_DIR, "GhiaData.csv"), scratch)
        os.chdir(scratch)
        g = {"__name__": "__mrt_reference__"}
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            exec(compile(src, "MRT.py", "exec"), g)
        return np.array(g["rho"]), np.array(g["u"]), np.array(g["fin"])
    finally:
        os.chdir(cwd)
        shutil.rmtree(scratch, ignore_errors=True)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def ref_functions_built(variant: str = "functions") -> bool:
    if not os.path.isdir(REF_BUILD_DIR):
        return False
    return any(fn.startswith(variant + ".") and fn.endswith(".so") for fn in os.listdir(REF_BUILD_DIR))


def load_ref_functions(variant: str = "functions"):
    """Import the compiled reference Cython module (``set_omega``, ``equ``, ``allfunc`` ...) from oracle/_ref."""
    if not ref_functions_built(variant):
        raise ImportError("oracle/_ref/%s*.so not built -- run `python oracle/build_ref.py`" % variant)
    if REF_BUILD_DIR not in sys.path:
        sys.path.insert(0, REF_BUILD_DIR)
    return importlib.import_module(variant)


# --------------------------------------------------------------------------------------------
# The reference's own CUDA kernels (built by oracle/build_ref_kernels.py), run on the GPU box
# --------------------------------------------------------------------------------------------
REF_KERNELS_SO = os.path.join(REF_BUILD_DIR, "libref_kernels.so")
REF_KERNELS_F64_SO = os.path.join(REF_BUILD_DIR, "libref_kernels_f64.so")     # same text, float -> double


def ref_kernels_built(dtype: str = "float32") -> bool:
    return os.path.exists(REF_KERNELS_SO if dtype == "float32" else REF_KERNELS_F64_SO)


def run_reference_kernels(nx: int, ny: int, Re: float, steps: int, collision: str = "MRT", turb: int = 0,
                          uLB: float = 0.08, dtype: str = "float32"):
    """Run funRT (+ funBC) of MRT_GPU.py for ``steps`` iterations exactly as the script does (fp32 device arrays in
    the [k][y][x] layout, init of :259-267 / :323-328, parameters of :63-91) and return ``(rho[x,y], u[2,x,y],
    fin[9,x,y])`` as the script would after its downloads and transposes (:755-760)."""
    import ctypes as C_
    from . import lbm_oracle as O
    # dtype="float64": the same kernel text with float -> double (oracle/build_ref_kernels.py)
    npdt = np.float32 if dtype == "float32" else np.float64
    lib = C_.CDLL(REF_KERNELS_SO if dtype == "float32" else REF_KERNELS_F64_SO)
    fp = C_.POINTER(C_.c_float if dtype == "float32" else C_.c_double)
    lib.ref_run.argtypes = [C_.c_int, C_.c_int, C_.c_int, fp, C_.c_int, C_.c_int] + [fp] * 6
    nuLB = uLB * ny / Re                                       # MRT_GPU.py:63
    omega = 2.0 / (6. * nuLB + 1)                              # :65
    omegam = 1.0 / (0.5 + ((1.0 / 3.5) / ((1 / omega) - 0.5)))  # :82-84
    params = {"SRT": [uLB, omega, turb], "TRT": [uLB, omega, omegam, turb],
              "MRT": [uLB, omega, 1.0, 1.2, 1.2, turb]}[collision]      # :88-91 and the `%` tuples
    P = np.asarray(params, dtype=npdt)
    _, vel, feq = O.init_fields(nx, ny, uLB)                   # :259-267 (fp64), then cast to fp32 (:298)
    dev = lambda a: np.ascontiguousarray(np.swapaxes(a.astype(npdt), -1, -2))     # [.., x, y] -> [.., y, x]
    fin, ftemp, feq_g = dev(feq), dev(feq), dev(feq)
    rho = np.ones((ny, nx), npdt)
    u = np.zeros((2, ny, nx), npdt)
    taus = np.full((ny, nx), 1.0 / omega, npdt)
    ptr = lambda a: a.ctypes.data_as(fp)
    rc = lib.ref_run({"SRT": 0, "TRT": 1, "MRT": 2}[collision], nx, ny, ptr(P), len(P), int(steps),
                     ptr(fin), ptr(ftemp), ptr(feq_g), ptr(rho), ptr(u), ptr(taus))
    if rc != 0:
        raise RuntimeError("ref_run failed with %d" % rc)
    host = lambda a: np.ascontiguousarray(np.swapaxes(a, -1, -2))
    return host(rho), host(u), host(fin)
