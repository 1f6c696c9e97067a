"""Summarise an .ncu-rep (read here on the CPU box) into profiles/<name>.txt and update profiles/traffic.json.

usage: python tools/ncu_summary.py gpurun_out/r01_f64.ncu-rep profiles/r01_ncu_f64_cavity4096 [traffic_key]
"""
import csv
import io
import json
import os
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "sm__cycles_elapsed.avg.per_second", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
    "lts__t_bytes.sum", "smsp__cycles_active.avg", "sm__cycles_active.avg",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    key = sys.argv[3] if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ["source: %s  (ncu --set full --clock-control none --import-source on)" % os.path.basename(rep), ""]
    traffic = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append("kernel: " + d.get("Kernel Name", "?"))
        for w in WANT:
            if w in d:
                lines.append("  %-75s %s %s" % (w, d[w], units[hdr.index(w)]))
        def tobytes(name):
            v, u = float(d[name]), units[hdr.index(name)].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        tr = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
        traffic.append(tr)
        lines.append("  dram traffic per launch (read+write)                                        %.0f bytes" % tr)
        lines.append("")
    with open(out + ".txt", "w") as fh:
        fh.write("\n".join(lines))
    print("\n".join(lines))
    if key:
        path = os.path.join(os.path.dirname(out), "traffic.json")
        tj = json.load(open(path)) if os.path.exists(path) else {}
        tj[key] = int(sum(traffic) / len(traffic))
        json.dump(tj, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
