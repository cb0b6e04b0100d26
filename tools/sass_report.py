#!/usr/bin/env python
"""SASS evidence for the hot kernels of libmcb200.so: architecture, resource usage and opcode histogram
(whole function and the longest straight-line loop body), one file per kernel under profiles/.

    python tools/sass_report.py            # writes profiles/r2_sass_hist_<kernel>.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "monte-carlo-project-cuda_b200", "libmcb200.so")
KERNELS = {   # label -> substring of the mangled name
    "european_kernel": "european_kernelILi0ELi64ELi4E",
    "european_job_kernel": "european_job_kernelILi0ELi64E",
    "european_small_job_kernel": "european_small_job_kernelILi0ELi64E",
    "european_packed_kernel": "european_packed_kernelILi0ELi64E",
    "trajectory_long_kernel_prices_counts": "trajectory_long_kernelILb1ELb0ELi4E",
    "bullet_kernel": "bullet_kernelILi4E",
    "nested_kernel": "nested_kernelE",
    "sweep_kernel": "sweep_kernelILi0ELi64E",
    "trajectory_slab_kernel_prices": "trajectory_slab_kernelILi16ELi16ELi6ELi4ELb0ELb0ELb0ELb1ELb1E",
    "trajectory_slab_kernel_prices_counts_fast": "trajectory_slab_kernelILi16ELi16ELi4ELi4ELb1ELb0ELb0ELb1ELb1E",
    "segments_job_kernel": "segments_job_kernelE",
    "combine_job_kernel": "combine_job_kernelE",
}
INSTR = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    funcs = {}
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name:
            m = INSTR.match(line)
            if m:
                funcs[name].append((int(m.group(1), 16), m.group(2).strip()))
    usage = dict(re.findall(r"Function (\S+):\n\s+(REG:.*)", res))
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    for label, pat in KERNELS.items():
        hits = [n for n in funcs if pat in n]
        if not hits:
            print("missing", label)
            continue
        fn = hits[0]
        ins = funcs[fn]

        def opcode(text):
            parts = text.split()
            return parts[1] if parts[0].startswith("@") else parts[0]

        whole = collections.Counter(opcode(t) for _, t in ins)
        # loops = backward branches; report the one whose body holds the most multiply / MUFU work
        best = None
        for addr, text in ins:
            m = re.search(r"BRA(?:\.U)?\s+(?:\w+,\s*)?0x([0-9a-f]+)", text)
            if m and int(m.group(1), 16) < addr:
                body = [t for a, t in ins if int(m.group(1), 16) <= a <= addr]
                score = sum(1 for t in body if "IMAD.WIDE" in t or "MUFU" in t)
                if best is None or score > best[0]:
                    best = (score, int(m.group(1), 16), addr, body)
        out = [f"SASS report of {fn}", f"library arch: {', '.join(arch)}", f"resources: {usage.get(fn, 'n/a')}",
               f"instructions: {len(ins)}", "", "-- whole function --"]
        out += [f"{c:6d} {op}" for op, c in whole.most_common()]
        if best:
            hist = collections.Counter(opcode(t) for t in best[3])
            out += ["", f"-- hottest loop body 0x{best[1]:x}..0x{best[2]:x} ({len(best[3])} instructions) --"]
            out += [f"{c:6d} {op}" for op, c in hist.most_common()]
        marks = [k for k in ("UBLKCP", "UTMASTG", "UTMACMDFLUSH", "FFMA2", "FADD2", "FMUL2", "UIMAD.WIDE", "REDG", "ATOMG",
                             "MEMBAR", "ST.E", "STG", "STAS", "SYNCS.PHASECHK", "UCGABAR") if any(k in t for _, t in ins)]
        out += ["", "markers present: " + (", ".join(marks) or "none")]
        with open(os.path.join(ROOT, "profiles", f"r2_sass_hist_{label}.txt"), "w") as f:
            f.write("\n".join(out) + "\n")
        print(label, len(ins), "instructions;", "loop", len(best[3]) if best else 0)


if __name__ == "__main__":
    main()
