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


def _run_strips_one_gpu(nx, ny, world, Re, steps, dtype, split_regions, two_step=True, tuning=None, packed=False):
    import torch
    import latticeboltzmannsimulations_b200 as L
    from latticeboltzmannsimulations_b200 import _capi
    from latticeboltzmannsimulations_b200.distributed import exchange_local, halo_plan, partition_rows, strip_views
    tdt = torch.float64 if dtype == "float64" else torch.float32
    parts = partition_rows(ny, world)
    deep = all(n >= 2 for _, n in parts)
    solvers, views, ptrs = [], [], []
    for (y0, nyl) in parts:
        nbytes = L.CavitySolver.state_bytes(nx, ny, 1, dtype, ny_local=nyl)
        raw = [torch.zeros(nbytes // tdt.itemsize, dtype=tdt, device="cuda") for _ in range(2)]
        s = L.CavitySolver(nx, ny, 1, dtype, "MRT", y0=y0, ny_local=nyl, ext_buffers=[t.data_ptr() for t in raw],
                           tuning=tuning)
        s.set_reynolds(Re)
        s.init_equilibrium()
        solvers.append(s)
        views.append([strip_views(t, s.layout) for t in raw])
        ptrs.append({raw[0].data_ptr(): 0, raw[1].data_ptr(): 1})
    plans = [halo_plan(r, world, parts[r][1], deep) for r in range(world)]
    it, used_two = 0, 0
    while it < steps:
        two = two_step and deep and steps - it >= 2 and all(s.step2_available() for s in solvers)
        n = 2 if two else 1
        wm = it + n == steps
        dst = [ptrs[r][solvers[r].buffer_ptr(1)] for r in range(world)]
        # worst-case ordering of the overlapped multi-GPU step: EDGE bands of every rank, then the halo exchange,
        # and only then the INTERIOR bands -- rows the exchange ships must already be final after EDGE
        do_split = split_regions and all(n >= 5 for _, n in parts)
        for s in solvers:
            region = s.step2_region if two else s.step_region
            region(_capi.LBM_REGION_EDGE if do_split else _capi.LBM_REGION_ALL, wm)
        torch.cuda.synchronize()
        if packed and deep:
            # the packed exchange of StripCavity: nine rows per neighbour through one contiguous buffer
            bufs = {(r, d): torch.empty(9, nx, dtype=tdt, device="cuda") for r in range(world) for d in (0, 1)}
            for r in range(world):
                for d in ((0,) if r > 0 else ()) + ((1,) if r < world - 1 else ()):
                    solvers[r].halo_pack(d, bufs[(r, d)].data_ptr())
            for r in range(world):
                if r > 0:
                    solvers[r].halo_unpack(0, bufs[(r - 1, 1)].data_ptr())      # what the strip above sent down
                if r < world - 1:
                    solvers[r].halo_unpack(1, bufs[(r + 1, 0)].data_ptr())      # what the strip below sent up
        else:
            exchange_local(plans, [views[r][dst[r]] for r in range(world)], nx)
        torch.cuda.synchronize()
        if do_split:
            for s in solvers:
                region = s.step2_region if two else s.step_region
                region(_capi.LBM_REGION_INTERIOR, wm)
            torch.cuda.synchronize()
        for s in solvers:
            s.swap2() if two else s.swap()
        it += n
        used_two += int(two)
    rho = np.concatenate([s.macros()[0] for s in solvers], axis=1)
    u = np.concatenate([s.macros()[1] for s in solvers], axis=2)
    f = np.concatenate([s.download_f() for s in solvers], axis=2)
    for s in solvers:
        s.close()
    return rho, u, f, used_two


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("nx,ny,world,split", [(96, 64, 2, True), (50, 37, 3, True), (64, 32, 8, False), (40, 9, 4, True),
                                               (256, 192, 2, True), (200, 120, 4, True)])
def test_strips_equal_single_domain_bitwise(nx, ny, world, split, dtype):
    import latticeboltzmannsimulations_b200 as L
    steps = 60
    want = L.run_cavity(nx, ny, 1000, steps=steps, dtype=dtype, return_f=True)
    got = _run_strips_one_gpu(nx, ny, world, 1000, steps, dtype, split)
    for a, b in zip(got[:3], want):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("tuning", [None, {"slide_min_nodes": 0}, {"slide_min_nodes": 0, "slide_h": 14},
                                    {"slide_min_nodes": 0, "slide_h": 50}])
@pytest.mark.parametrize("nx,ny,world,split,steps", [(1100, 640, 2, True, 9), (900, 700, 3, True, 12), (1500, 410, 4, False, 7),
                                                      (800, 800, 5, True, 10), (1300, 26, 8, True, 9)])
def test_two_step_kernel_on_strips_bitwise(nx, ny, world, split, steps, dtype, tuning):
    """Temporal blocking on y-strips (shared-memory tiles by default at these sizes, the sliding-window kernel with
    several segment heights through the tuning knob): edge / interior bands of whole tile rows / segments, nine-row
    halo exchange, odd step counts -- bit-identical to the undecomposed run and to the one-step kernels."""
    import latticeboltzmannsimulations_b200 as L
    want = L.run_cavity(nx, ny, 1000, steps=steps, dtype=dtype, return_f=True)
    got = _run_strips_one_gpu(nx, ny, world, 1000, steps, dtype, split, tuning=tuning)
    if (dtype == "float64" and nx * ny >= 600000) or tuning is not None:    # (fp32 takes no tiles; tiny strips neither)
        assert got[3] >= (steps - 1) // 2                  # the two-step kernel really ran
    for a, b in zip(got[:3], want):
        assert np.array_equal(a, b)
    packed = _run_strips_one_gpu(nx, ny, world, 1000, steps, dtype, split, tuning=tuning, packed=True)
    for a, b in zip(packed[:3], want):
        assert np.array_equal(a, b)
    if tuning is not None:
        return
    one = _run_strips_one_gpu(nx, ny, world, 1000, steps, dtype, split, two_step=False)
    assert one[3] == 0
    for a, b in zip(one[:3], want):
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
            from latticeboltzmannsimulations_b200.distributed import strip_views
            body, g2 = strip_views(r[guard:guard + n], lay)
            assert bool((body[:, :, nx:] == 0).all())            # pitch padding never written
            assert bool(torch.isfinite(body).all()) and bool((g2 == 0).all())
