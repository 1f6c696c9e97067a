"""SASS evidence of the hot kernels in the shipped library (read on the CPU box with cuobjdump):
for each kernel the register count is not in SASS, so this lists the instruction mix and the Blackwell-specific
mnemonics (UBLKCP = cp.async.bulk, SYNCS = mbarrier, FFMA2 / FADD2 / FMUL2 = packed fp32, LDGSTS = cp.async,
UTMALDG = TMA tensor load) plus the first lines of the main loop.   usage: python tools/sass_excerpts.py > profiles/r02_sass_excerpts.txt"""
import collections
import re
import subprocess
import sys

LIB = "latticeboltzmannsimulations_b200/lib/liblbm_b200.so"
WANT = [("lbm_step_slide2<double, MRT, no macros, no closure>", r"lbm_step_slide2IdLi2ELb0ELi3ELb0E"),
        ("lbm_step_slide2<float, MRT, no macros, no closure>", r"lbm_step_slide2IfLi2ELb0ELi3ELb0E"),
        ("lbm_step_slide2<float, SRT, no macros, Smagorinsky> (the reference's default configuration)", r"lbm_step_slide2IfLi0ELb0ELi3ELb1E"),
        ("lbm_step_ldg<double, MRT, gather, no macros, step> (one-step fp64)", r"lbm_step_ldgIdLi2ELb1ELb0ELi0ELb0E"),
        ("lbm_step_vec<float, MRT, no macros, 4 nodes per thread> (one-step fp32)", r"lbm_step_vecIfLi2ELb0ELi4ELb0E"),
        ("lbm_step_fused2<double, MRT, 32x8 tiles> (small cavities)", r"lbm_step_fused2IdLi2ELb0ELi32ELi8ELi4ELb0E"),
        ("lbm_step_tma (optional TMA tensor engine)", r"lbm_step_tmaIdLi2ELb0ELi2ELi4ELi4ELi1E")]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", sass)
print("cuobjdump -sass", LIB, "(sm_100a)\n")
for title, pat in WANT:
    hit = [b for b in blocks if re.match(r"\S*" + pat, b)]
    if not hit:
        print("==", title, ": not found\n")
        continue
    b = hit[0]
    ins = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", b)
    mix = collections.Counter(i.split(".")[0] for i in ins)
    print("==", title)
    print("   symbol:", b.split("\n")[0].strip())
    print("   %d SASS instructions; mix: %s" % (len(ins), ", ".join("%s %d" % kv for kv in mix.most_common(14))))
    special = {k: sum(v for n, v in collections.Counter(ins).items() if n.startswith(k)) for k in
               ("UBLKCP", "SYNCS", "UTMALDG", "LDGSTS", "FFMA2", "FADD2", "FMUL2", "DFMA", "DADD", "DMUL", "LDG.E.128", "LDG.E.64",
                "STG.E.128", "STG.E.64", "LDS.64", "LDS.128", "STS.64", "SHFL", "BAR", "HMMA", "UTC")}
    print("   selected:", ", ".join("%s %d" % kv for kv in special.items() if kv[1]))
    lines = [l.strip() for l in b.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]
    keys = [l for l in lines if re.search(r"UBLKCP|SYNCS|FFMA2|UTMALDG|LDGSTS", l)][:6]
    for l in keys:
        print("     ", re.sub(r"\s*/\* 0x[0-9a-f]+ \*/", "", l))
    print()
