"""Build recipe: compile the REFERENCE'S OWN CUDA kernels (the `funRT` SRT / TRT / MRT strings and `funBC` of
/root/reference/MRT_GPU.py:336-699) with nvcc into oracle/_ref/libref_kernels.so, so that they can be executed on the
GPU box as a second opinion on semantics "C" -- in particular on the two pieces nothing else executable pins: the MRT
relaxation (MRT_GPU.py:633-655) and the wall rule funBC (:664-699).

TEST INFRASTRUCTURE ONLY.  The kernel text is read where it lies under /root/reference (never copied into the
tracked tree): the generated .cu is written to a temporary directory, only the .so lands in oracle/_ref/ (git-ignored;
it travels to the GPU box with the snapshot).  The reference compiles these strings at run time through PyCUDA after
splicing literals in with `%`; PyCUDA is not installed and the values differ per test, so the only edits are:
  * each `%s` becomes a read of a __constant__ float array P[] (same value the literal would have after the
    implicit double->float conversion of `float uLB = 0.08;`), in the order of the reference's own
    `funRT % (uLB, omega, turb)` / `(uLB, omegap, omegam, turb)` / `(uLB, omega_nu, omega_e, omega_eps, omega_q, turb)`;
  * the three variants, all called `funRT`, are renamed funRT_SRT / funRT_TRT / funRT_MRT;
  * a small host driver (ours, below) replaces the Python time loop MRT_GPU.py:707-732: launches funRT then funBC
    with block (32,32,1), grid (nx/32, ny/32), exactly like :724,732.
A second build, oracle/_ref/libref_kernels_f64.so, is the SAME text with every `float` token replaced by `double`
(kernel signatures, locals, tables, the parameter array and the driver's buffers): the reference's algorithm in the
precision the oracle works in, so that the MRT relaxation and funBC are pinned to round-off of fp64, not of fp32.
"""
from __future__ import annotations

import os
import re
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "libref_kernels.so")
OUT_F64 = os.path.join(HERE, "_ref", "libref_kernels_f64.so")
REF_DIR = os.environ.get("LBM_REFERENCE_DIR", "/root/reference")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

DRIVER = r'''
#include <cuda_runtime.h>
extern "C" int ref_run(int coll, int nx, int ny, const float* params, int nparams, int steps,
                       float* fin, float* ftemp, float* feq, float* rho, float* u, float* taus) {
    if (nx % 32 || ny % 32 || nparams > 8) return -1;             /* MRT_GPU.py:53-54: multiples of 32 */
    const size_t n = (size_t)nx * ny;
    float *d_fin, *d_ftemp, *d_feq, *d_rho, *d_u, *d_taus;
    if (cudaMalloc(&d_fin, 9 * n * 4) || cudaMalloc(&d_ftemp, 9 * n * 4) || cudaMalloc(&d_feq, 9 * n * 4) ||
        cudaMalloc(&d_rho, n * 4) || cudaMalloc(&d_u, 2 * n * 4) || cudaMalloc(&d_taus, n * 4)) return -2;
    cudaMemcpyToSymbol(P, params, nparams * sizeof(float));
    cudaMemcpy(d_fin, fin, 9 * n * 4, cudaMemcpyHostToDevice);     /* MRT_GPU.py:323-328 */
    cudaMemcpy(d_ftemp, ftemp, 9 * n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_feq, feq, 9 * n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_rho, rho, n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_u, u, 2 * n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_taus, taus, n * 4, cudaMemcpyHostToDevice);
    dim3 block(32, 32, 1), grid(nx / 32, ny / 32, 1);
    for (int it = 0; it < steps; ++it) {                            /* MRT_GPU.py:707, 724, 732 */
        if (coll == 0) funRT_SRT<<<grid, block>>>(d_fin, d_ftemp, d_feq, d_rho, d_u, d_taus);
        else if (coll == 1) funRT_TRT<<<grid, block>>>(d_fin, d_ftemp, d_feq, d_rho, d_u, d_taus);
        else funRT_MRT<<<grid, block>>>(d_fin, d_ftemp, d_feq, d_rho, d_u, d_taus);
        funBC<<<grid, block>>>(d_ftemp, d_feq, d_fin);
    }
    cudaError_t e = cudaGetLastError();                             /* a failed launch (resources) is not a sync error */
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaMemcpy(fin, d_fin, 9 * n * 4, cudaMemcpyDeviceToHost);     /* the scripts download ftemp_g == fin_g, :755 */
    cudaMemcpy(rho, d_rho, n * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(u, d_u, 2 * n * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(taus, d_taus, n * 4, cudaMemcpyDeviceToHost);
    cudaFree(d_fin); cudaFree(d_ftemp); cudaFree(d_feq); cudaFree(d_rho); cudaFree(d_u); cudaFree(d_taus);
    return e == cudaSuccess ? 0 : -3;
}
'''


def generate_source(real: str = "float") -> str:
    with open(os.path.join(REF_DIR, "MRT_GPU.py"), encoding="utf-8-sig") as fh:
        src = fh.read()
    rt = re.findall(r'funRT = """(.*?)"""', src, flags=re.S)
    bc = re.findall(r'funBC = """(.*?)"""', src, flags=re.S)
    assert len(rt) == 3 and len(bc) == 1, "unexpected layout of MRT_GPU.py"
    assert [k.count("%s") for k in rt] == [3, 4, 6] and bc[0].count("%") == 0
    out = ["// GENERATED from /root/reference/MRT_GPU.py by oracle/build_ref_kernels.py -- not tracked", "#include <math.h>",
           "__constant__ float P[8];"]
    for name, text in zip(("SRT", "TRT", "MRT"), rt):
        assert text.count("void funRT(") == 1
        text = text.replace("void funRT(", "void funRT_%s(" % name)
        i = 0
        while "%s" in text:
            text = text.replace("%s", "P[%d]" % i, 1)
            i += 1
        out.append(text)
    out.append(bc[0])
    out.append(DRIVER)
    text = "\n".join(out)
    if real != "float":
        text = re.sub(r"\bfloat\b", real, text).replace("* 4", "* sizeof(%s)" % real)
    return text


def build(force: bool = False) -> str:
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    for out, real in ((OUT, "float"), (OUT_F64, "double")):
        if os.path.exists(out) and not force:
            continue
        with tempfile.TemporaryDirectory(prefix="ref_kernels_") as tmp:
            cu = os.path.join(tmp, "ref_kernels.cu")
            with open(cu, "w") as fh:
                fh.write(generate_source(real))
            # fp64 build: -fmad=false (no contraction: the arithmetic the source text spells out) and 64 registers per
            # thread, without which the reference's 32 x 32 = 1024-thread blocks (MRT_GPU.py:53-54) cannot launch
            subprocess.check_call([NVCC, "-O2", "-w", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                                   "-shared"] + (["-fmad=false", "-maxrregcount=64"] if real == "double" else []) +
                                  ["-o", out, cu])
    return OUT


if __name__ == "__main__":
    import sys
    if not os.path.isfile(os.path.join(REF_DIR, "MRT_GPU.py")):
        print("reference not present -- using prebuilt oracle/_ref/libref_kernels.so if any")
    else:
        print(build(force="--force" in sys.argv))
