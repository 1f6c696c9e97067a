"""torchrun entry: CUDA-event timeline of the overlapped y-strip (double) step of the 32768^2 cavity -- what each phase of
a pass costs on the device and where the pass time goes.  Prints one line per rank: mean microseconds over the timed
passes of  edge launches | pack + NCCL + unpack | interior launch | whole pass (interior start -> next interior start),
and the gap between the end of one interior launch and the start of the next."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from latticeboltzmannsimulations_b200.distributed import StripCavity, nccl_options

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=nccl_options())
rank, world = dist.get_rank(), dist.get_world_size()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
sc = StripCavity(n, n, 10000.0, 0.08, "float64", "MRT")
sc.step(7)
sc.sync(); dist.barrier()
sc.timeline = []
import time
t_host = time.perf_counter()
sc.step(40)
t_host = (time.perf_counter() - t_host) / len(sc.timeline) * 1e6      # host time to ENQUEUE one pass (with the event records)
sc.sync()
tl = sc.timeline[2:]
edge = [t["edge0"].elapsed_time(t["edge1"]) * 1e3 for t in tl]
xchg = [t["edge1"].elapsed_time(t["xchg1"]) * 1e3 for t in tl]
inte = [t["int0"].elapsed_time(t["int1"]) * 1e3 for t in tl]
whole = [a["int0"].elapsed_time(b["int0"]) * 1e3 for a, b in zip(tl[:-1], tl[1:])]
gap = [a["int1"].elapsed_time(b["int0"]) * 1e3 for a, b in zip(tl[:-1], tl[1:])]
lag = [t["int0"].elapsed_time(t["edge0"]) * 1e3 for t in tl]
mean = lambda v: sum(v) / max(len(v), 1)
line = ("rank %d/%d strip %d rows, %d steps per pass: edge %.0f us | pack+nccl+unpack %.0f us | interior %.0f us | pass %.0f us "
        "| interior-to-interior gap %.0f us | edge starts %.0f us after interior | host enqueue time per pass %.0f us" % (
            rank, world, sc.nyl, tl[0]["steps"], mean(edge), mean(xchg), mean(inte), mean(whole), mean(gap), mean(lag), t_host))
out = [None] * world if rank == 0 else None
dist.gather_object(line, out, dst=0)
if rank == 0:
    print("\n".join(out), flush=True)
sc.close()
dist.barrier()
dist.destroy_process_group()
