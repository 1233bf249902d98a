#!/usr/bin/env python
"""SASS opcode summary per kernel of libvq_b200.so (cuobjdump -sass): which kernels use the tensor cores (UTCHMMA /
UTCQMMA), TMA (UTMALDG / UTMASTG), tensor memory (LDTM / STTM / UTCBAR), 128-bit global loads, fp64 — the evidence the
profiling guide asks for.  Writes a markdown table to stdout:  python tools/sass_summary.py > profiles/rN_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "video_query_algorithms_b200", "lib", "libvq_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM", "STTM", "UTCCP", "SYNCS", "LDG.E.128", "LDG",
         "STG", "LDS", "STS", "FFMA", "DFMA", "DADD", "DMUL", "MUFU", "SHFL", "ATOM", "RED", "BAR", "HMMA", "NANOSLEEP"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    kernels, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and name:
            op = m.group(1)
            kernels[name]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + ".") or (w == "LDG.E.128" and op.startswith("LDG") and ".128" in op):
                    kernels[name][w] += 1
    demangled = subprocess.run(["c++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS opcode counts per kernel — libvq_b200.so (%s), `cuobjdump -sass`\n" % ", ".join(arch))
    print("Counts are static instruction counts in the cubin (not executed counts). Tensor-core path = UTCHMMA (tcgen05.mma kind::f16) / "
          "UTCQMMA (kind::f8f6f4), TMA = UTMALDG/UTMASTG, tensor memory = LDTM/STTM, mbarrier = SYNCS/UTCBAR.\n")
    cols = [w for w in WATCH if any(k[w] for k in kernels.values())]
    print("| kernel | instr | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for (mangled, c), nice in zip(kernels.items(), demangled):
        nice = nice.replace("(anonymous namespace)::", "")
        nice = re.sub(r"^void ", "", re.sub(r"\(.*", "", nice))
        print("| `%s` | %d | " % (nice, c["_total"]) + " | ".join(str(c[w]) if c[w] else "" for w in cols) + " |")


if __name__ == "__main__":
    main()
