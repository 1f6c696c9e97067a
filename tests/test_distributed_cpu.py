"""Host-side logic of the multi-GPU paths on CPU: row partition, halo plan, and a world_size-2/3 `gloo` run in
which every rank advances its strip with the oracle's strip form and exchanges halo rows through the SAME
HaloExchanger the GPU path uses (CPU tensors instead of CUDA tensors) -- the result must equal the single-domain
oracle bit for bit.  Also the cavity -> rank sharding of the batched sweep."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import lbm_oracle as O
from latticeboltzmannsimulations_b200.distributed import (HaloExchanger, halo_plan, partition_rows,
                                                           shard_indices, DOWN_POPS, UP_POPS)


def test_partition_rows():
    assert partition_rows(10, 3) == [(0, 4), (4, 3), (7, 3)]
    assert partition_rows(32768, 8) == [(i * 4096, 4096) for i in range(8)]
    for ny, w in [(7, 7), (100, 8), (33, 2)]:
        parts = partition_rows(ny, w)
        assert parts[0][0] == 0 and sum(n for _, n in parts) == ny
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(w - 1))
    with pytest.raises(ValueError):
        partition_rows(3, 4)


def test_halo_plan_is_symmetric():
    """Every send has exactly one matching recv on the peer: same population, edge row -> ghost row."""
    world, nyl = 4, 5
    plans = [halo_plan(r, world, nyl) for r in range(world)]
    assert len(plans[0]) == 6 and len(plans[1]) == 12 and len(plans[3]) == 6
    for r in range(world):
        for kind, peer, (what, k, row) in plans[r]:
            if kind != "send":
                continue
            going_up = peer == r - 1
            assert what == "row" and k in (UP_POPS if going_up else DOWN_POPS)
            assert row == (1 if going_up else nyl)
            want_row = nyl + 1 if going_up else 0
            assert ("recv", r, ("row", k, want_row)) in plans[peer]
    # order of sends on one side matches order of recvs on the other (NCCL/gloo match P2P ops in issue order)
    for r in range(world - 1):
        down = [spec[1] for kind, peer, spec in plans[r] if kind == "send" and peer == r + 1]
        up_recv = [spec[1] for kind, peer, spec in plans[r + 1] if kind == "recv" and peer == r]
        assert down == up_recv
    # deep plan (two-step kernel): nine rows per direction, the shallow rows are a subset, same send/recv order
    deep = [halo_plan(r, world, nyl, deep=True) for r in range(world)]
    assert len(deep[0]) == 18 and len(deep[1]) == 36
    for r in range(world):
        for item in plans[r]:
            assert item in deep[r]
    for r in range(world - 1):
        s_down = [spec for kind, peer, spec in deep[r] if kind == "send" and peer == r + 1]
        r_up = [spec for kind, peer, spec in deep[r + 1] if kind == "recv" and peer == r]
        assert [sp[1] for sp in s_down[:6]] == [sp[1] for sp in r_up[:6]] == [0, 1, 3, 4, 7, 8]
        assert [sp[2] for sp in s_down[6:]] == [nyl - 1] * 3 and [sp[0] for sp in r_up[6:]] == ["g2"] * 3
        s_up = [spec for kind, peer, spec in deep[r + 1] if kind == "send" and peer == r]
        r_down = [spec for kind, peer, spec in deep[r] if kind == "recv" and peer == r + 1]
        assert [sp[1] for sp in s_up[:6]] == [sp[1] for sp in r_down[:6]] == [0, 1, 3, 2, 5, 6]
        assert [sp[2] for sp in s_up[6:]] == [2] * 3 and [(sp[0], sp[1]) for sp in r_down[6:]] == [("g2", 1)] * 3


def test_shard_indices():
    assert shard_indices(10, 1, 4) == [1, 5, 9]
    allidx = sorted(i for r in range(8) for i in shard_indices(256, r, 8))
    assert allidx == list(range(256)) and len(shard_indices(256, 3, 8)) == 32


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, nx, ny, steps, q, deep=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = O.Params(nx, ny, Re=400, collision="MRT")
        fin0 = O.random_state(nx, ny, seed=3)
        y0, nyl = partition_rows(ny, world)[rank]
        ss = O.StripState.from_fin(fin0, p, y0, nyl)
        pitch = (nx + 31) // 32 * 32
        bufs = [torch.zeros(9, nyl + 2, pitch, dtype=torch.float64) for _ in range(2)]   # device layout [k][row][x]
        g2 = [torch.zeros(2, 3, pitch, dtype=torch.float64) for _ in range(2)]           # second ghost rows
        ex = HaloExchanger(list(zip(bufs, g2)), nx, rank, world, deep=deep)
        which = 0
        for _ in range(steps):
            ss.g[:, :, 0] = np.nan; ss.g[:, :, nyl + 1] = np.nan       # ghosts must come from the exchange only
            if rank > 0:
                ss.g[:, :, 0] = bufs[which][:, 0, :nx].numpy()
            if rank < world - 1:
                ss.g[:, :, nyl + 1] = bufs[which][:, nyl + 1, :nx].numpy()
            O.step_C_pull_strip(ss, p)
            which ^= 1
            bufs[which][:, 1:nyl + 1, :nx] = torch.from_numpy(ss.g[:, :, 1:nyl + 1].transpose(0, 2, 1).copy())
            for w in ex.exchange(which):
                w.wait()
        ss.g[:, :, 0] = bufs[which][:, 0, :nx].numpy()
        ss.g[:, :, nyl + 1] = bufs[which][:, nyl + 1, :nx].numpy()
        out = [None] * world if rank == 0 else None
        dist.gather_object((O.strip_fin(ss, p), ss.rho), out, dst=0)
        if rank == 0:
            q.put((np.concatenate([o[0] for o in out], axis=2), np.concatenate([o[1] for o in out], axis=1)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nx,ny,deep", [(2, 24, 30, False), (3, 20, 23, True)])
def test_strip_decomposition_gloo(world, nx, ny, deep):
    steps = 12
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nx, ny, steps, q, deep)) for r in range(world)]
    for pr in procs:
        pr.start()
    f, rho = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p = O.Params(nx, ny, Re=400, collision="MRT")
    want = O.run(p, steps, fin0=O.random_state(nx, ny, seed=3), form="pull")
    assert np.array_equal(f, want[2]) and np.array_equal(rho, want[0])


def _fake_datagen(Re_list, nx, ny, uLB, steps, collision, dtype, return_steps=False, **kw):
    """Stand-in for cavity.datagen on a box without a GPU: every array is filled with its cavity's Reynolds number."""
    n = len(Re_list)
    f = np.empty((n, 9, nx, ny), np.float32); u = np.empty((n, 2, nx, ny), np.float32)
    for i, re in enumerate(Re_list):
        f[i] = re; u[i] = -re
    feq = np.full((9, nx, ny), 7.0, np.float32)
    done = np.array([steps + int(re) for re in Re_list], np.int64)
    assert kw == {"turb": True, "converge": True, "Pinterval": 50}
    return f, u, feq, np.asarray(Re_list), done


def _sweep_worker(rank, world, port, out_dir, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from latticeboltzmannsimulations_b200 import cavity, distributed
        cavity.datagen = _fake_datagen
        Re = list(range(100, 150, 10))                       # 5 cavities over 2 ranks: 3 + 2; over 7 ranks: some get none
        res = distributed.datagen_sharded(Re, 6, 4, steps=1000, collision="SRT", out_dir=out_dir if rank == 0 else None,
                                          return_steps=True, turb=True, converge=True, Pinterval=50)
        mine, local = distributed.datagen_sharded(Re, 6, 4, steps=1000, gather=False, turb=True, converge=True, Pinterval=50)
        assert mine == list(range(rank, 5, world)) and (local is None) == (not mine)
        with pytest.raises(TypeError):
            distributed.datagen_sharded(Re, 6, 4, stepz=3)
        if rank == 0:
            q.put(res)
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 7])
def test_sharded_sweep_assembly_gloo(world, tmp_path):
    """datagen_sharded: cavity b -> rank b mod world, sweep options forwarded to every rank's datagen, results back in
    the order of Re_list on rank 0 (also with ranks that own no cavity), dataset files written by rank 0."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sweep_worker, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    for pr in procs:
        pr.start()
    f, u, feq, Re, done = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    want = np.arange(100, 150, 10)
    assert Re.dtype == np.int64 and np.array_equal(Re, want)
    assert np.array_equal(f[:, 0, 0, 0], want) and np.array_equal(u[:, 1, 5, 3], -want) and np.all(feq == 7.0)
    assert np.array_equal(done, 1000 + want)
    assert np.array_equal(np.load(tmp_path / "f_final.npy"), f) and np.load(tmp_path / "Re_range.npy").dtype == np.int64
    assert np.load(tmp_path / "u_final.npy").shape == (5, 2, 6, 4) and np.load(tmp_path / "feq_initial.npy").shape == (9, 6, 4)


def test_assemble_sweep_rejects_gaps_and_overlaps():
    from latticeboltzmannsimulations_b200.distributed import assemble_sweep
    part = _fake_datagen([100, 120], 3, 3, 0.08, 10, "MRT", "float32", turb=True, converge=True, Pinterval=50)
    with pytest.raises(ValueError, match="no rank"):
        assemble_sweep([([0, 2], part)], [100, 110, 120], 3, 3, "float32")
    with pytest.raises(ValueError, match="more than one"):
        assemble_sweep([([0, 2], part), ([2, 1], part)], [100, 110, 120], 3, 3, "float32")


def _load_emu():
    """tests/host_emu: the product's one-step kernel source compiled for the host (see tests/test_host_emulation.py)."""
    import ctypes as C
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("build_emu", os.path.join(here, "host_emu", "build_emu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = C.CDLL(mod.build())
    dp = C.POINTER(C.c_double)
    lib.emu_strip_pass.argtypes = [C.c_int] * 7 + [dp] * 8 + [C.c_int] * 2
    lib.emu_run.argtypes = [C.c_int] * 7 + [dp] * 7
    return lib, dp


def _kernel_strip_worker(rank, world, port, nx, ny, steps, q):
    """One rank of a y-strip run made of the PRODUCT's parts on the CPU: the one-step kernel source (emulated node by
    node) on the strip's device-layout buffers, the product's HaloExchanger over gloo between the steps, in the product's
    order -- edge rows, exchange, interior rows."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib, dp = _load_emu()
        p = O.Params(nx, ny, Re=400, collision="MRT")
        rates = np.array([p.uLB, p.omega, p.omegam, p.omega_e, p.omega_eps, p.omega_q, 1.0 / p.omega])
        y0, nyl = partition_rows(ny, world)[rank]
        pitch = (nx + 31) // 32 * 32
        bufs = [torch.zeros(9, nyl + 2, pitch, dtype=torch.float64) for _ in range(2)]
        rho, ux, uy = (np.zeros((nyl, pitch)) for _ in range(3))
        rho_lid, carry = np.zeros(pitch), np.zeros(4)
        import ctypes
        ptr = lambda a: a.ctypes.data_as(dp)
        tptr = lambda t: ctypes.cast(t.data_ptr(), dp)

        def launch(mode, gather, macros, src, dst, r0, n):
            assert lib.emu_strip_pass(mode, gather, macros, nx, ny, y0, nyl, ptr(rates), tptr(src), tptr(dst), ptr(rho), ptr(ux),
                                      ptr(uy), ptr(rho_lid), ptr(carry), r0, n) == 0

        ex = HaloExchanger(bufs, nx, rank, world)
        launch(3, 0, 0, bufs[0], bufs[0], 0, nyl)                       # lbm_init_equilibrium
        which = 0
        for i in range(steps):
            src, dst, last = bufs[which], bufs[which ^ 1], int(i == steps - 1)
            if nyl >= 5:                                               # LBM_REGION_EDGE, exchange, LBM_REGION_INTERIOR
                launch(0, int(i > 0), last, src, dst, 0, 2)
                launch(0, int(i > 0), last, src, dst, nyl - 2, 2)
                works = ex.exchange(which ^ 1)
                launch(0, int(i > 0), last, src, dst, 2, nyl - 4)
            else:
                launch(0, int(i > 0), last, src, dst, 0, nyl)
                works = ex.exchange(which ^ 1)
            for w in works:
                w.wait()
            which ^= 1
        launch(1, 1, 0, bufs[which], bufs[which ^ 1], 0, nyl)           # the download's finalize pass
        fin = bufs[which ^ 1][:, 1:nyl + 1, :nx].numpy().transpose(0, 2, 1).copy()
        out = [None] * world if rank == 0 else None
        dist.gather_object((fin, rho[:, :nx].T.copy()), out, dst=0)
        if rank == 0:
            q.put((np.concatenate([o[0] for o in out], axis=2), np.concatenate([o[1] for o in out], axis=1)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nx,ny", [(2, 40, 26), (3, 33, 13)])
def test_product_kernel_and_exchanger_in_strips_gloo(world, nx, ny):
    """Strips of the product's own kernel source + the product's halo exchanger equal the single-domain run of the same
    kernel BIT FOR BIT (and the oracle to 1e-12): the N > 1 data path, minus the GPU, on world_size 2 and 3."""
    steps = 15
    lib, dp = _load_emu()                 # built once here, before the ranks start
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_kernel_strip_worker, args=(r, world, port, nx, ny, steps, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    f, rho = q.get(timeout=180)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p = O.Params(nx, ny, Re=400, collision="MRT")
    rates = np.array([p.uLB, p.omega, p.omegam, p.omega_e, p.omega_eps, p.omega_q, 1.0 / p.omega])
    f1 = np.empty((9, nx, ny)); rho1 = np.empty((nx, ny)); u1 = np.empty((2, nx, ny))
    ptr = lambda a: a.ctypes.data_as(dp)
    assert lib.emu_run(0, 1, 2, 0, nx, ny, steps, ptr(rates), None, ptr(f1), ptr(rho1), ptr(u1), None, None) == 0
    assert np.array_equal(f, f1) and np.array_equal(rho, rho1)
    want = O.run(p, steps, form="pull")
    assert np.abs(f - want[2]).max() < 1e-12 and np.abs(rho - want[0]).max() < 1e-12
