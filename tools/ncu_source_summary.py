"""Where the warps of a captured kernel spend their time: stall samples and executed instructions per SASS opcode, and the
hottest instructions with the line before them (read here on the CPU box from an .ncu-rep taken with --import-source on).

usage: python tools/ncu_source_summary.py gpurun_out/r02c_f64.ncu-rep profiles/r02_ncu_f64_source_hotspots.txt
"""
import csv
import io
import subprocess
import sys
from collections import Counter


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    name, hdr, data = rows[0][1], rows[1], rows[2:]
    src, smp, exe = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    tot_s = sum(int(r[smp]) for r in data)
    tot_e = sum(int(r[exe]) for r in data)
    by_s, by_e = Counter(), Counter()
    for r in data:
        words = r[src].split()
        op = (words[1] if words[0].startswith("@") else words[0]).split(".")[0]
        by_s[op] += int(r[smp])
        by_e[op] += int(r[exe])
    lines = ["source: %s (ncu --page source --csv)" % rep.split("/")[-1], "kernel: " + name,
             "%d SASS instructions, %d warp-level instructions executed, %d stall samples" % (len(data), tot_e, tot_s), "",
             "%-10s %9s %7s %13s %7s" % ("opcode", "samples", "", "executed", "")]
    for op, _ in by_s.most_common(20):
        lines.append("%-10s %9d %6.1f%% %13d %6.1f%%" % (op, by_s[op], 100.0 * by_s[op] / tot_s, by_e[op], 100.0 * by_e[op] / tot_e))
    lines += ["", "hottest instructions (stall samples, share, instruction; the instruction before it in brackets)"]
    order = sorted(range(len(data)), key=lambda i: -int(data[i][smp]))[:12]
    for i in order:
        prev = data[i - 1][src].strip() if i else ""
        lines.append("%7d %5.1f%%  %-50s [%s]" % (int(data[i][smp]), 100.0 * int(data[i][smp]) / tot_s,
                                                  data[i][src].strip()[:50], prev[:60]))
    text = "\n".join(lines) + "\n"
    open(out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
