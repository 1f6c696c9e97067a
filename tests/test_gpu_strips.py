"""GPU tests of the y-strip decomposition.  On one GPU the strips are separate solver handles whose halo rows are
copied with torch row views following the same ``halo_plan`` the NCCL path uses; the decomposed result must equal
the single-handle run bit for bit (the update is local and order-independent).  With >= 2 GPUs the real NCCL path
(``StripCavity``) is launched under torchrun."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_strips_one_gpu(nx, ny, world, Re, steps, dtype, split_regions):
    import torch
    import latticeboltzmannsimulations_b200 as L
    from latticeboltzmannsimulations_b200 import _capi
    from latticeboltzmannsimulations_b200.distributed import halo_plan, partition_rows
    tdt = torch.float64 if dtype == "float64" else torch.float32
    parts = partition_rows(ny, world)
    solvers, bufs = [], []
    for (y0, nyl) in parts:
        nbytes = L.CavitySolver.state_bytes(nx, ny, 1, dtype, ny_local=nyl)
        raw = [torch.zeros(nbytes // tdt.itemsize, dtype=tdt, device="cuda") for _ in range(2)]
        s = L.CavitySolver(nx, ny, 1, dtype, "MRT", y0=y0, ny_local=nyl, ext_buffers=[t.data_ptr() for t in raw])
        s.set_reynolds(Re)
        s.init_equilibrium()
        lay = s.layout
        solvers.append(s)
        bufs.append(([t.view(9, int(lay.rows), int(lay.pitch)) for t in raw], {raw[0].data_ptr(): 0, raw[1].data_ptr(): 1}))
    for it in range(steps):
        dst = [bufs[r][1][solvers[r].buffer_ptr(1)] for r in range(world)]
        wm = it == steps - 1
        for r, s in enumerate(solvers):
            if split_regions and parts[r][1] >= 3:
                s.step_region(_capi.LBM_REGION_EDGE, wm)
                s.step_region(_capi.LBM_REGION_INTERIOR, wm)
            else:
                s.step_region(_capi.LBM_REGION_ALL, wm)
        torch.cuda.synchronize()
        for r in range(world):
            for kind, peer, k, row in halo_plan(r, world, parts[r][1]):
                if kind != "send":
                    continue
                ghost = parts[peer][1] + 1 if peer == r - 1 else 0
                bufs[peer][0][dst[peer]][k, ghost, :nx] = bufs[r][0][dst[r]][k, row, :nx]
        torch.cuda.synchronize()
        for s in solvers:
            s.swap()
    rho = np.concatenate([s.macros()[0] for s in solvers], axis=1)
    u = np.concatenate([s.macros()[1] for s in solvers], axis=2)
    f = np.concatenate([s.download_f() for s in solvers], axis=2)
    for s in solvers:
        s.close()
    return rho, u, f


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("nx,ny,world,split", [(96, 64, 2, True), (50, 37, 3, True), (64, 32, 8, False), (40, 9, 4, True)])
def test_strips_equal_single_domain_bitwise(nx, ny, world, split, dtype):
    import latticeboltzmannsimulations_b200 as L
    steps = 60
    want = L.run_cavity(nx, ny, 1000, steps=steps, dtype=dtype, return_f=True)
    got = _run_strips_one_gpu(nx, ny, world, 1000, steps, dtype, split)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


def test_nccl_strips_under_torchrun():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29577", os.path.join(ROOT, "tools", "check_strips_nccl.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "STRIPS_OK" in out.stdout


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_no_out_of_bounds_writes_canary(dtype):
    """Caller-owned population buffers with sentinel guard zones before and after (compute-sanitizer is closed on
    this pool): after stepping with every region / kernel form the guards and the pitch padding must be untouched."""
    import torch
    import latticeboltzmannsimulations_b200 as L
    from latticeboltzmannsimulations_b200 import _capi
    tdt = torch.float64 if dtype == "float64" else torch.float32
    nx, ny, y0, nyl = 70, 40, 12, 17          # nx not a multiple of the vector width or 32; a strip in the middle
    nbytes = L.CavitySolver.state_bytes(nx, ny, 1, dtype, ny_local=nyl)
    n, guard, sentinel = nbytes // tdt.itemsize, 4096, -7.25
    raw = [torch.full((n + 2 * guard,), sentinel, dtype=tdt, device="cuda") for _ in range(2)]
    for r in raw:
        r[guard:guard + n] = 0
    with L.CavitySolver(nx, ny, 1, dtype, "MRT", y0=y0, ny_local=nyl,
                        ext_buffers=[r[guard:].data_ptr() for r in raw]) as s:
        s.set_reynolds(400)
        s.init_equilibrium()
        lay = s.layout
        for it in range(6):
            if it % 2:
                s.step_region(_capi.LBM_REGION_EDGE); s.step_region(_capi.LBM_REGION_INTERIOR)
            else:
                s.step_region(_capi.LBM_REGION_ALL, write_macros=True)
            s.swap()
        s.download_f(); s.macros(current=True)
        torch.cuda.synchronize()
        for r in raw:
            assert bool((r[:guard] == sentinel).all()) and bool((r[guard + n:] == sentinel).all())
            body = r[guard:guard + n].view(9, int(lay.rows), int(lay.pitch))
            assert bool((body[:, :, nx:] == 0).all())            # pitch padding never written
            assert bool(torch.isfinite(body).all())
