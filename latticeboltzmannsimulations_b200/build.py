"""Build the C-ABI shared library in-tree:  python -m latticeboltzmannsimulations_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU; the resulting lib/liblbm_b200.so travels to the GPU box with the
gpurun snapshot (it is git-ignored, not gpurun-ignored)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "liblbm_b200.so")
SOURCES = ["lbm_b200.cu", "lbm_slide2_f64.cu", "lbm_slide2_f32.cu"]      # compiled in parallel, then linked
HEADERS = [os.path.join("..", "..", "include", "lbm_b200.h")]                # + every .cuh / .h in csrc
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# -fmad=false: no implicit contraction; every fused multiply-add is explicit in lbm_device.cuh, so that what a node
# computes does not depend on the kernel that inlines it (all kernel families are bit-identical by construction)
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
         "-Xcompiler", "-fPIC"]
OBJDIR = os.path.join(HERE, "lib", "obj")


def _deps():
    deps = [os.path.join(CSRC, f) for f in HEADERS] + [os.path.abspath(__file__)]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    return [d for d in deps if os.path.exists(d)]


STAMP = LIB + ".srchash"


def _source_hash() -> str:
    import hashlib
    h = hashlib.sha256()
    for d in sorted(_deps()):
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def stale() -> bool:
    """True if the library is missing or was built from other sources than the ones in the tree (content hash, not
    mtimes: the gpurun snapshot does not promise to preserve those)."""
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != _source_hash()


_stale = stale


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError("nvcc not found at %s" % NVCC)
    os.makedirs(OBJDIR, exist_ok=True)
    hdr_t = max(os.path.getmtime(d) for d in _deps() if not d.endswith(".cu"))     # object reuse is a local convenience
    procs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJDIR, src[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(hdr_t, os.path.getmtime(path)):
            continue
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
        procs.append((src, subprocess.Popen(cmd)))
    failed = [src for src, p in procs if p.wait() != 0]
    if failed:
        raise RuntimeError("nvcc failed for " + ", ".join(failed))
    objs = [os.path.join(OBJDIR, src[:-3] + ".o") for src in SOURCES]
    # link next to the target and rename: a process that loads the library meanwhile (several ranks of one job start
    # together) sees the old file or the new one, never a half-written one
    tmp = "%s.%d.tmp" % (LIB, os.getpid())
    try:
        subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs)
        os.replace(tmp, LIB)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
    with open(STAMP + ".tmp", "w") as fh:
        fh.write(_source_hash() + "\n")
    os.replace(STAMP + ".tmp", STAMP)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
