"""Throughput vs grid shape / memory footprint (development aid)."""
import sys
import torch
sys.path.insert(0, ".")
import latticeboltzmannsimulations_b200 as L

cases = [(32768, 2048), (32768, 8192), (32768, 16384), (16384, 32768), (8192, 65536), (4096, 4096), (65536, 8192), (2048, 32768)]
dt = sys.argv[1] if len(sys.argv) > 1 else "float64"
for nx, ny in cases:
    try:
        with L.CavitySolver(nx, ny, 1, dt, "MRT") as s:
            s.set_reynolds(5000); s.init_equilibrium(); s.step(4, write_macros=False); s.sync()
            steps = max(4, int(2e9 / (nx * ny)))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); s.step(steps, write_macros=False, stream=torch.cuda.current_stream().cuda_stream); e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            bpn = 144 if dt == "float64" else 72
            print("%6d x %-6d %s state 2x%.1f GB: %.3f ms/step %.0f MLUPS %.0f GB/s" % (
                nx, ny, dt, nx * ny * bpn / 2 / 1e9, ms, nx * ny / ms / 1e3, nx * ny / ms / 1e3 * bpn / 1e3), flush=True)
    except Exception as e:
        print(nx, ny, "failed", e, flush=True)
