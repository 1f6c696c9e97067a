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
