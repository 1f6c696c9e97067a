"""TEST INFRASTRUCTURE: build tests/host_emu/_build/libemu.so -- the product's one-thread-per-node kernel source
(csrc/lbm_device.cuh, lbm_kernels.cuh, lbm_aa.cuh) compiled by g++ as host code against stub/cuda_runtime.h and driven
node by node by emu.cpp.  The only edit to the source text: the `griddepcontrol` (programmatic dependent launch) inline
assembly lines of lbm_kernels.cuh are dropped in a scratch copy, the host assembler has no such instruction."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "..", "latticeboltzmannsimulations_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libemu.so")
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, "emu.cpp"), os.path.join(HERE, "stub", "cuda_runtime.h"), os.path.abspath(__file__)]
    srcs += [os.path.join(CSRC, f) for f in ("lbm_device.cuh", "lbm_kernels.cuh", "lbm_aa.cuh")]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) > max(os.path.getmtime(s) for s in srcs):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    import fcntl
    with open(os.path.join(OUT, ".lock"), "w") as lock:        # several test processes may get here together
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and os.path.exists(LIB) and os.path.getmtime(LIB) > max(os.path.getmtime(s) for s in srcs):
            return LIB                                         # built by another process meanwhile
        with open(os.path.join(CSRC, "lbm_kernels.cuh")) as fh:
            text = "".join(line for line in fh if "griddepcontrol" not in line)
        scratch = os.path.join(OUT, "lbm_kernels_nopdl.cuh")
        with open(scratch + ".tmp", "w") as fh:
            fh.write(text)
        os.replace(scratch + ".tmp", scratch)
        tmp = LIB + ".%d.tmp" % os.getpid()
        subprocess.check_call([GXX, "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                               "-I", os.path.join(HERE, "stub"), "-I", OUT, "-I", CSRC, os.path.join(HERE, "emu.cpp"), "-o", tmp])
        os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
