"""Diagnose multi-GPU step cost: kernel-only vs with exchange; CPU enqueue time (development aid)."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, ".")
from latticeboltzmannsimulations_b200.distributed import StripCavity
from latticeboltzmannsimulations_b200 import _capi
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
nx = ny = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
sc = StripCavity(nx, ny, 10000.0, 0.08, "float64", "MRT", overlap=True)
def timed(label, fn, n=10):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter(); fn(n); cpu = time.perf_counter() - t
    torch.cuda.synchronize(); tot = time.perf_counter() - t
    print("rank %d %-28s cpu-enqueue %.3f ms/step   total %.3f ms/step" % (rank, label, cpu / n * 1e3, tot / n * 1e3), flush=True)
sc.step(3); sc.sync()
def kernel_only(n):
    for _ in range(n):
        sc.solver.step_region(_capi.LBM_REGION_ALL, False, sc.s_main.cuda_stream); sc.solver.swap()
timed("kernel only (s_main)", kernel_only)
def kernel_default(n):
    for _ in range(n):
        sc.solver.step_region(_capi.LBM_REGION_ALL, False, 0); sc.solver.swap()
timed("kernel only (default stream)", kernel_default)
def comm_only(n):
    for _ in range(n):
        for w in sc.halo.exchange(0): w.wait()
timed("exchange only", comm_only)
sc.overlap = False
timed("step no-overlap", lambda n: sc.step(n))
sc.overlap = True
timed("step overlap", lambda n: sc.step(n))
def edge_only(n):
    for _ in range(n):
        sc.solver.step_region(_capi.LBM_REGION_EDGE, False, sc.s_main.cuda_stream)
timed("edge rows only", edge_only)
def interior_only(n):
    for _ in range(n):
        sc.solver.step_region(_capi.LBM_REGION_INTERIOR, False, sc.s_main.cuda_stream)
timed("interior rows only", interior_only)
print(rank, torch.cuda.memory_allocated() / 1e9, "GB torch;", os.environ.get("PYTORCH_CUDA_ALLOC_CONF"), flush=True)
sc.close(); dist.barrier(); dist.destroy_process_group()
