"""Multi-GPU drivers: one process per GPU, ``torch.distributed`` (NCCL over NVLink) as plumbing.

* ``StripCavity`` -- ONE large cavity cut into contiguous y-strips (rows are x-contiguous in the device layout, so a
  strip boundary is a whole row).  The update of a row reads only rows y-1, y, y+1, hence exactly one
  nearest-neighbour exchange per step: across each interface the three populations that cross it
  (towards larger y: k in {4,7,8}; towards smaller y: k in {2,5,6}), ``3 * nx`` values per direction -- nine rows
  per direction when the two-step (temporal blocking) kernel is used, which recomputes the ghost row.  Per (double)
  step the edge bands of the strip are updated first on a high-priority halo stream, the nine rows for each neighbour
  are gathered into ONE contiguous buffer (``lbm_halo_pack``), exchanged with one grouped NCCL send + recv per
  interface and scattered into the ghost rows (``lbm_halo_unpack``), while the interior rows are updated concurrently
  on the main stream.  Strips of a single row (shallow plan) exchange row views of the population buffers directly.
  The reference has no multi-GPU path at all (single ``cuda.Device(0)``), so there is no upstream
  interface to mirror here.
* ``datagen_sharded`` -- the Reynolds sweep of ``MRT_GPU_datagen.py:55-57``: independent cavities, rank r takes the
  cavities ``r::world``; no data-path collective (results are gathered once at the end).

The partition / halo bookkeeping is plain Python over ``torch`` tensors and is exercised on CPU with the ``gloo``
backend in ``tests/test_distributed_cpu.py``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _capi

UP_POPS = (2, 5, 6)      # c_y = +1: move to y-1, i.e. to the strip above (smaller y)
DOWN_POPS = (4, 7, 8)    # c_y = -1: move to y+1, i.e. to the strip below


def nccl_options():
    """Process-group options for the y-strip driver: NCCL's internal stream at high priority.  torch runs the grouped
    send / recv on that stream, not on the caller's; at default priority its kernel queues behind every CTA of the
    interior launch (measured: the exchange then ends with the interior, a 45 us bubble per pass); at high priority it
    slips in between, like the edge launches.  Returns None where torch has no such option."""
    try:
        import torch.distributed as dist
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True
        return opts
    except Exception:     # pragma: no cover
        return None


def partition_rows(ny: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous strips [(y0, ny_local)] -- the first ``ny % world`` ranks get one extra row."""
    if world < 1 or ny < world:
        raise ValueError("need 1 <= world <= ny")
    base, rem = divmod(ny, world)
    out, y0 = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((y0, n))
        y0 += n
    return out


AXIS_POPS = (0, 1, 3)     # c_y = 0: stay in their row


def halo_plan(rank: int, world: int, nyl: int, deep: bool = False):
    """P2P plan of one (double) step for ``rank``: list of ``(kind, peer, spec)``.

    ``spec`` is ``("row", k, stored_row)`` for a row of population k in the main buffer (stored row r holds local
    row r-1; stored rows 0 and nyl+1 are the ghost rows) or ``("g2", which, j)`` for the j-th row of the second
    ghost rows (which = 0: above the strip, populations 4,7,8; 1: below, populations 2,5,6).  Sends read the freshly
    written edge rows, receives land in the ghost rows of the same (destination) buffer.  Matching sends and
    receives are listed in the same order on both sides.

    ``deep=False``: what one-step kernels need -- the three crossing populations of the adjacent row.
    ``deep=True``: what the two-step (temporal blocking) kernel needs -- it recomputes the ghost row's own first
    sub-step, so it also wants the ghost row's in-row populations and the crossing populations of the row beyond it:
    nine rows per direction and interface.  A superset of the shallow plan, valid for one-step kernels too.
    """
    plan = []
    if not deep:
        if rank > 0:                       # interface with the strip above
            plan += [("send", rank - 1, ("row", k, 1)) for k in UP_POPS]
            plan += [("recv", rank - 1, ("row", k, 0)) for k in DOWN_POPS]
        if rank < world - 1:               # interface with the strip below
            plan += [("send", rank + 1, ("row", k, nyl)) for k in DOWN_POPS]
            plan += [("recv", rank + 1, ("row", k, nyl + 1)) for k in UP_POPS]
        return plan
    if nyl < 2:
        raise ValueError("the deep halo plan needs at least two rows per strip")
    if rank > 0:
        plan += [("send", rank - 1, ("row", k, 1)) for k in AXIS_POPS + UP_POPS]
        plan += [("send", rank - 1, ("row", k, 2)) for k in UP_POPS]
        plan += [("recv", rank - 1, ("row", k, 0)) for k in AXIS_POPS + DOWN_POPS]
        plan += [("recv", rank - 1, ("g2", 0, j)) for j in range(3)]
    if rank < world - 1:
        plan += [("send", rank + 1, ("row", k, nyl)) for k in AXIS_POPS + DOWN_POPS]
        plan += [("send", rank + 1, ("row", k, nyl - 1)) for k in DOWN_POPS]
        plan += [("recv", rank + 1, ("row", k, nyl + 1)) for k in AXIS_POPS + UP_POPS]
        plan += [("recv", rank + 1, ("g2", 1, j)) for j in range(3)]
    return plan


def strip_views(raw, layout, batch: int = 1):
    """Views of one A/B allocation (a flat tensor of ``state_bytes``): main ``[9, rows, pitch]`` and second ghost
    rows ``[2, 3, pitch]`` (``lbm_layout_t.ghost2_offset``).  Single-cavity strips only."""
    if batch != 1:
        raise ValueError("y-strips hold one cavity")
    rows, pitch, off = int(layout.rows), int(layout.pitch), int(layout.ghost2_offset)
    return raw[:off].view(9, rows, pitch), raw[off:off + 6 * pitch].view(2, 3, pitch)


def plan_view(spec, main, g2, nx: int):
    kind, a, b = spec
    return main[a, b, :nx] if kind == "row" else g2[a, b, :nx]


def exchange_local(plans, views, nx: int) -> None:
    """Carry out the plans of all ranks inside one process (the single-GPU emulation used by the tests):
    ``views[r] = (main, g2)`` of rank r's destination buffer."""
    for r, plan in enumerate(plans):
        sends = {}
        for kind, peer, spec in plan:
            if kind == "send":
                sends.setdefault(peer, []).append(spec)
        for peer, specs in sends.items():
            recvs = [spec for kind, p2, spec in plans[peer] if kind == "recv" and p2 == r]
            assert len(recvs) == len(specs)
            for s_spec, r_spec in zip(specs, recvs):
                plan_view(r_spec, *views[peer], nx).copy_(plan_view(s_spec, *views[r], nx))


class HaloExchanger:
    """Executes ``halo_plan`` on a pair of A/B buffers given as ``(main [9, nyl+2, pitch], g2 [2, 3, pitch])`` views
    (any device/backend; ``g2`` may be None for the shallow plan)."""

    def __init__(self, buffers, nx: int, rank: int, world: int, group=None, deep: bool = False):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank, self.world = rank, world
        buffers = [b if isinstance(b, (tuple, list)) else (b, None) for b in buffers]
        nyl = buffers[0][0].shape[1] - 2
        self.plan = halo_plan(rank, world, nyl, deep)
        # ops are fixed per destination buffer: build them once (views of the persistent buffers)
        self._ops = []
        for main, g2 in buffers:
            ops = []
            for kind, peer, spec in self.plan:
                view = plan_view(spec, main, g2, nx)
                fn = dist.isend if kind == "send" else dist.irecv
                ops.append(dist.P2POp(fn, view, peer, group=group))
            self._ops.append(ops)

    def exchange(self, which: int):
        """Start the exchange on buffer ``which``; returns the work handles (call ``.wait()`` on each)."""
        if not self._ops[which]:
            return []
        return self.dist.batch_isend_irecv(self._ops[which])


class StripCavity:
    """One cavity decomposed into y-strips over the ranks of ``group`` (one CUDA device per rank)."""

    def __init__(self, nx: int, ny: int, Re: float, uLB: float = 0.08, dtype="float64", collision: str = "MRT",
                 group=None, device: Optional[int] = None, engine: str = "auto", overlap: bool = True,
                 tuning: Optional[dict] = None):
        import torch
        import torch.distributed as dist
        from .solver import CavitySolver, _dtype_name
        self.torch, self.dist = torch, dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.nx, self.ny = nx, ny
        self.parts = partition_rows(ny, self.world)
        self.y0, self.nyl = self.parts[self.rank]
        self.device = torch.cuda.current_device() if device is None else device
        tdt = torch.float64 if _dtype_name(dtype) == "float64" else torch.float32
        nbytes = CavitySolver.state_bytes(nx, ny, 1, dtype, ny_local=self.nyl)
        esz = 8 if tdt == torch.float64 else 4
        with torch.cuda.device(self.device):
            self._raw = [torch.zeros(nbytes // esz, dtype=tdt, device="cuda") for _ in range(2)]
        self.solver = CavitySolver(nx, ny, 1, dtype, collision, y0=self.y0, ny_local=self.nyl, device=self.device,
                                   engine=engine, ext_buffers=[t.data_ptr() for t in self._raw], tuning=tuning)
        lay = self.solver.layout
        self.buffers = [strip_views(t, lay) for t in self._raw]
        self.solver.set_reynolds(Re, uLB)
        self.solver.init_equilibrium()
        self._ptr = {self._raw[0].data_ptr(): 0, self._raw[1].data_ptr(): 1}
        # the deep plan (what the two-step kernel needs) whenever every strip has at least two rows
        self.deep = all(n >= 2 for _, n in self.parts)
        self.halo = HaloExchanger(self.buffers, nx, self.rank, self.world, group, deep=self.deep)
        # packed exchange (deep plan only): the nine rows per neighbour travel as ONE contiguous buffer, gathered /
        # scattered by lbm_halo_pack / lbm_halo_unpack -> one send + one recv per interface instead of nine each
        self.packed = self.deep and self.world > 1
        self._pack_ops = []
        if self.packed:
            with torch.cuda.device(self.device):
                self._sbuf = [torch.zeros(9, nx, dtype=tdt, device="cuda") for _ in range(2)]   # to above / to below
                self._rbuf = [torch.zeros(9, nx, dtype=tdt, device="cuda") for _ in range(2)]   # from above / from below
            self._sides = [d for d, ok in ((0, self.rank > 0), (1, self.rank < self.world - 1)) if ok]
            for d in self._sides:
                peer = self.rank - 1 if d == 0 else self.rank + 1
                self._pack_ops.append(dist.P2POp(dist.isend, self._sbuf[d], peer, group=group))
                self._pack_ops.append(dist.P2POp(dist.irecv, self._rbuf[d], peer, group=group))
        self.overlap = overlap and self.nyl >= 5
        with torch.cuda.device(self.device):
            self.s_main = torch.cuda.Stream()
            self.s_halo = torch.cuda.Stream(priority=-1)
            self.ev_main = torch.cuda.Event()
            self.ev_halo = torch.cuda.Event()
            self.ev_main.record(torch.cuda.current_stream())
            self.ev_halo.record(torch.cuda.current_stream())
            # two events per stream, used alternately (the wait of pass n+1 is enqueued before pass n+2 re-records)
            self._evs_main = [torch.cuda.Event(), torch.cuda.Event()]
            self._evs_halo = [torch.cuda.Event(), torch.cuda.Event()]
        self.steps_done = 0
        self.passes_done = 0          # launches of the whole strip (a two-step pass counts once)
        self.timeline = None          # set to [] to record CUDA events around the phases of every pass (tools/strip_timeline.py)

    def _dst_index(self) -> int:
        return self._ptr[self.solver.buffer_ptr(1)]

    def _exchange(self, dst: int, stream) -> None:
        """Halo exchange of the buffer being written, enqueued on ``stream`` (the host does not wait)."""
        s = self.solver
        if self.packed:
            for d in self._sides:
                s.halo_pack(d, self._sbuf[d].data_ptr(), stream.cuda_stream)
            for w in self.dist.batch_isend_irecv(self._pack_ops):
                w.wait()
            for d in self._sides:
                s.halo_unpack(d, self._rbuf[d].data_ptr(), stream.cuda_stream)
        else:
            for w in self.halo.exchange(dst):
                w.wait()

    def step(self, nsteps: int = 1, write_macros: bool = False) -> None:
        """Advance ``nsteps`` steps; two at a time with the temporal-blocking kernel when it is available (every
        rank takes the same decision: it depends only on the global size and on the state being post-collision)."""
        torch = self.torch
        s = self.solver
        i = 0
        while i < nsteps:
            two = self.deep and (nsteps - i) >= 2 and s.step2_available()
            n = 2 if two else 1
            wm = write_macros and i + n == nsteps
            region = s.step2_region if two else s.step_region
            dst = self._dst_index()
            if self.overlap:
                k = self.passes_done & 1
                tl = None
                if self.timeline is not None:
                    tl = {name: torch.cuda.Event(enable_timing=True) for name in
                          ("edge0", "edge1", "xchg1", "int0", "int1")}
                    tl["steps"] = n
                    self.timeline.append(tl)
                with torch.cuda.stream(self.s_halo):
                    self.s_halo.wait_event(self.ev_main)              # interior of the previous step
                    if tl: tl["edge0"].record(self.s_halo)
                    region(_capi.LBM_REGION_EDGE, wm, self.s_halo.cuda_stream)
                    if tl: tl["edge1"].record(self.s_halo)
                    self._exchange(dst, self.s_halo)                  # stream-side waits only, the host runs ahead
                    if tl: tl["xchg1"].record(self.s_halo)
                    new_halo = self._evs_halo[k]
                    new_halo.record(self.s_halo)
                with torch.cuda.stream(self.s_main):
                    self.s_main.wait_event(self.ev_halo)              # edge rows + halo of the previous step
                    if tl: tl["int0"].record(self.s_main)
                    region(_capi.LBM_REGION_INTERIOR, wm, self.s_main.cuda_stream)
                    if tl: tl["int1"].record(self.s_main)
                    new_main = self._evs_main[k]
                    new_main.record(self.s_main)
                self.ev_halo, self.ev_main = new_halo, new_main
            else:
                with torch.cuda.stream(self.s_main):
                    region(_capi.LBM_REGION_ALL, wm, self.s_main.cuda_stream)
                    self._exchange(dst, self.s_main)
            if two:
                s.swap2()
            else:
                s.swap()
            i += n
            self.steps_done += n
            self.passes_done += 1

    def sync(self) -> None:
        self.s_main.synchronize()
        self.s_halo.synchronize()

    def join_current_stream(self) -> None:
        """Make the caller's current stream wait for everything issued so far (for event timing on that stream)."""
        cur = self.torch.cuda.current_stream()
        if self.overlap:
            cur.wait_event(self.ev_main)
            cur.wait_event(self.ev_halo)
        else:
            cur.wait_stream(self.s_main)

    def fork_from_current_stream(self) -> None:
        cur = self.torch.cuda.current_stream()
        self.s_main.wait_stream(cur)
        self.s_halo.wait_stream(cur)

    # -- results (validation / output; not on the timed path) -----------------------------------------------------
    def local_fields(self, current: bool = False):
        self.sync()
        rho, u = self.solver.macros(current=current)
        return rho, u, self.solver.download_f()

    def gather_fields(self, current: bool = False):
        """Full-cavity (rho[nx,ny], u[2,nx,ny], f[9,nx,ny]) on rank 0 (None elsewhere)."""
        rho, u, f = self.local_fields(current)
        objs = [None] * self.world if self.rank == 0 else None
        self.dist.gather_object((rho, u, f), objs, dst=0, group=self.group)
        if self.rank != 0:
            return None
        return (np.concatenate([o[0] for o in objs], axis=1), np.concatenate([o[1] for o in objs], axis=2),
                np.concatenate([o[2] for o in objs], axis=2))

    def close(self) -> None:
        self.sync()
        self.solver.close()


def shard_indices(n: int, rank: int, world: int) -> List[int]:
    """Cavity b of a sweep goes to rank b mod world."""
    return list(range(rank, n, world))


def datagen_sharded(Re_list: Sequence[float], nx: int = 384, ny: int = 384, uLB: float = 0.08, steps: int = 10000,
                    collision: str = "MRT", dtype="float32", group=None, gather: bool = True,
                    out_dir: Optional[str] = None, return_steps: bool = False, **sweep):
    """Sharded sweep: each rank advances its own cavities (``cavity.datagen`` on the cavities ``rank::world``, no
    communication in the time loop); rank 0 assembles the reference's output arrays in the order of ``Re_list`` and,
    with ``out_dir``, writes the four ``.npy`` files of ``MRT_GPU_datagen.py:899-902``.

    ``sweep`` is forwarded to ``cavity.datagen``: ``turb``, ``converge``, ``Pinterval``, ``maxIt``, ``tol``, ``hits``
    (the reference's defaults are SRT + ``turb=True`` run to its stopping rule), ``chunk``, ``device``, ``engine``.
    Returns ``(f_final, u_final, feq_initial, Re_range[, steps_done])`` on rank 0 and ``None`` elsewhere; with
    ``gather=False`` every rank gets ``(its cavity indices, its local datagen result)``.
    """
    import torch.distributed as dist
    from . import cavity
    bad = set(sweep) - {"turb", "converge", "Pinterval", "maxIt", "tol", "hits", "chunk", "device", "engine"}
    if bad:
        raise TypeError("datagen_sharded() got unexpected keyword arguments %s" % sorted(bad))
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    Re_list = list(Re_list)
    mine = shard_indices(len(Re_list), rank, world)
    local = None
    if mine:
        local = cavity.datagen([Re_list[i] for i in mine], nx, ny, uLB, steps, collision, dtype, return_steps=True, **sweep)
    if not gather:
        return mine, local
    objs = [None] * world if rank == 0 else None
    dist.gather_object((mine, local), objs, dst=0, group=group)
    if rank != 0:
        return None
    out = assemble_sweep(objs, Re_list, nx, ny, dtype)
    if out_dir is not None:
        cavity.save_dataset(out_dir, *out[:4])
    return out if return_steps else out[:4]


def assemble_sweep(parts, Re_list, nx: int, ny: int, dtype):
    """Put the per-rank results ``[(cavity indices, datagen(..., return_steps=True) result or None), ...]`` back into
    the order of ``Re_list``: ``(f_final[N,9,nx,ny], u_final[N,2,nx,ny], feq_initial[9,nx,ny], Re_range[N],
    steps_done[N])``.  ``feq_initial`` is the same for every cavity (``rho = 1``, lid row at ``uLB``) and is taken from
    the rank that owns cavity 0."""
    from .cavity import _re_range_array
    n = len(Re_list)
    npdt = np.float32 if np.dtype(dtype) == np.float32 else np.float64
    f_final = np.empty((n, 9, nx, ny), npdt)
    u_final = np.empty((n, 2, nx, ny), npdt)
    steps_done = np.zeros(n, np.int64)
    feq0, seen = None, np.zeros(n, bool)
    for idx, loc in parts:
        if not idx:
            continue
        idx = list(idx)
        if seen[idx].any():
            raise ValueError("cavity assigned to more than one rank")
        seen[idx] = True
        f_final[idx] = loc[0]
        u_final[idx] = loc[1]
        steps_done[idx] = loc[4]
        if 0 in idx:
            feq0 = loc[2]
    if not seen.all():
        raise ValueError("cavities %s were computed by no rank" % np.flatnonzero(~seen).tolist())
    return f_final, u_final, feq0, _re_range_array(Re_list), steps_done
