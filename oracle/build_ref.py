"""Build recipe: compile the reference's own Cython/OpenMP step (``functions.pyx``) into ``oracle/_ref/``.

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference sources are read where they lie under
``/root/reference`` (never copied into the tracked tree; ``oracle/_ref/`` is git-ignored and holds the
generated ``.pyx`` copy, the Cython-generated ``.c`` and the ``.so``, which travel to the GPU box).

Accommodations (recorded in DESIGN.md), none of which touches the arithmetic:
  1. ``ctypedef long int_t`` is inserted after ``from numpy cimport *`` (``functions.pyx:2``): Cython 3.x
     no longer exports ``int_t`` from ``numpy.pxd``.
  2. compiled with ``/usr/bin/gcc`` (the image's default ``$CC`` cannot find ``libgomp.spec``).
  3. ``-march=native`` (``setup.py:11``) becomes ``-march=x86-64-v3``: the ``.so`` is built in the CPU-only
     build container and executed on a different host (the GPU box), so it must not use ISA extensions
     the other host may lack.  ``-O3 -ffast-math -fopenmp`` stay as shipped.
Variants:
  ``functions``           as shipped: ``parallel(num_threads=4)`` hard-coded (``functions.pyx:69``).
  ``functions_allcores``  the same source with that literal dropped, so OpenMP uses every core
                          (``OMP_NUM_THREADS``); labelled "modified thread count" wherever it is reported.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_DIR = os.environ.get("LBM_REFERENCE_DIR", "/root/reference")
GCC = "/usr/bin/gcc"


def _variant_source(variant: str) -> str:
    with open(os.path.join(REF_DIR, "functions.pyx"), encoding="utf-8-sig") as fh:
        src = fh.read()
    marker = "from numpy cimport *\n"
    assert marker in src
    src = src.replace(marker, marker + "ctypedef long int_t\n", 1)
    if variant == "functions_allcores":
        live = "    with nogil, parallel(num_threads=4):"      # functions.pyx:69 (the :126 twin is a comment)
        assert src.count("\n" + live) == 1
        src = src.replace("\n" + live, "\n    with nogil, parallel():", 1)
    return src


def build(variant: str = "functions", force: bool = False) -> str:
    import numpy
    os.makedirs(OUT, exist_ok=True)
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    so = os.path.join(OUT, variant + ext)
    if os.path.exists(so) and not force:
        return so
    pyx = os.path.join(OUT, variant + ".pyx")
    with open(pyx, "w") as fh:
        fh.write(_variant_source(variant))
    c_file = os.path.join(OUT, variant + ".c")
    subprocess.check_call([sys.executable, "-m", "cython", "-3", pyx, "-o", c_file], cwd=OUT)
    inc = [sysconfig.get_paths()["include"], numpy.get_include()]
    cmd = [GCC, "-shared", "-fPIC", "-O3", "-ffast-math", "-march=x86-64-v3", "-fopenmp", "-w",
           "-DNPY_NO_DEPRECATED_API=0"] + ["-I" + i for i in inc] + [c_file, "-o", so, "-lm"]
    subprocess.check_call(cmd)
    for tmp in (pyx, c_file):        # keep only the binary: no copy of reference source stays in the tree
        os.remove(tmp)
    return so


def main() -> None:
    if not os.path.isfile(os.path.join(REF_DIR, "functions.pyx")):
        print("reference not present at %s -- using prebuilt oracle/_ref if any" % REF_DIR)
        return
    force = "--force" in sys.argv
    for v in ("functions", "functions_allcores"):
        print("built", build(v, force=force))


if __name__ == "__main__":
    main()
