"""torchrun entry: StripCavity over NCCL must equal the single-GPU run bit for bit (fp64 and fp32), with and
without the edge/interior overlap.  Prints STRIPS_OK on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import latticeboltzmannsimulations_b200 as L
from latticeboltzmannsimulations_b200.distributed import StripCavity, nccl_options, datagen_sharded

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=nccl_options())
rank, world = dist.get_rank(), dist.get_world_size()
ok = True
SLIDE = {"slide_min_nodes": 0, "slide_h": 14}            # the sliding two-step kernel forced onto small strips
for dtype in ("float64", "float32"):
    for (nx, ny, steps, overlap, tuning) in [(256, 192, 80, True, None), (130, 67, 40, True, None), (512, 512, 50, False, None),
                                             (512, 384, 41, True, SLIDE), (700, 300, 30, False, SLIDE),
                                             (1536, 1400, 21, True, None)]:
        sc = StripCavity(nx, ny, 1000.0, 0.08, dtype, "MRT", overlap=overlap, tuning=tuning)
        sc.step(steps, write_macros=True)
        got = sc.gather_fields()
        packed = sc.packed
        sc.close()
        if rank == 0:
            want = L.run_cavity(nx, ny, 1000.0, steps=steps, dtype=dtype, return_f=True)
            same = all(np.array_equal(a, b) for a, b in zip(got, want))
            print("strips", dtype, nx, ny, steps, "overlap" if overlap else "serial", "packed halo" if packed else "row views",
                  "slide forced" if tuning else "default kernels", "bitwise equal:", same, flush=True)
            ok = ok and same
# sharded sweep: each cavity equals its standalone run
Re = [100.0 + 50 * i for i in range(2 * world + 1)]
res = datagen_sharded(Re, 64, 64, steps=50, dtype="float32")
if rank == 0:
    f_final, u_final, feq0, Re_out = res
    for b in (0, len(Re) // 2, len(Re) - 1):
        rho, u, f = L.run_cavity(64, 64, Re[b], steps=50, dtype="float32", return_f=True)
        same = np.array_equal(f, f_final[b]) and np.array_equal(u, u_final[b])
        print("sweep cavity", b, "bitwise equal:", same, flush=True)
        ok = ok and same
# optional full-size check (BASELINE config 5): 32768^2 over the ranks, corner windows against a small-cavity oracle.
# Information travels one node per step, so after 40 steps the nodes within 60 of a corner depend only on the two
# walls meeting there and equal the same nodes of a 256^2 cavity with the same relaxation rate.
if "--full" in sys.argv:
    from oracle import lbm_oracle as O
    n, steps, w = 32768, 40, 60
    sc = StripCavity(n, n, 10000.0, 0.08, "float64", "MRT", overlap=True)
    sc.step(steps, write_macros=True)
    rho, u, f = sc.local_fields()
    p = O.Params(256, 256, Re=10000.0 * 256 / n, collision="MRT")
    want = O.run(p, steps, form="pull")[2]
    errs = []
    if rank == 0:
        errs += [np.abs(f[:, :w, :w] - want[:, :w, :w]).max(), np.abs(f[:, -w:, :w] - want[:, -w:, :w]).max()]
    if rank == world - 1:
        errs += [np.abs(f[:, :w, -w:] - want[:, :w, -w:]).max(), np.abs(f[:, -w:, -w:] - want[:, -w:, -w:]).max()]
    mass = torch.tensor([float(f.sum())], device="cuda", dtype=torch.float64)
    dist.all_reduce(mass)
    sc.close()
    good = all(e <= 1e-12 for e in errs) and bool(np.isfinite(f).all())
    print("rank", rank, "full-size corner windows max err", errs, "ok", good, flush=True)
    if rank == 0:
        drift = abs(float(mass.item()) / (float(n) * n) - 1.0)
        print("full-size mean density drift after %d steps: %.3e" % (steps, drift), flush=True)
        good = good and drift < 1e-6
    g = torch.tensor([1 if good else 0], device="cuda")
    dist.all_reduce(g, op=dist.ReduceOp.MIN)
    ok = ok and int(g.item()) == 1
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
dist.barrier()
dist.destroy_process_group()
if rank == 0 and ok:
    print("STRIPS_OK", flush=True)
sys.exit(0 if int(flag.item()) == 1 else 1)
