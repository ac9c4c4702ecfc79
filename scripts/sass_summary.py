#!/usr/bin/env python
"""Per-kernel SASS summary of sc_b200/libscgpu.so (cuobjdump -sass, runs without a GPU): instruction count and the mnemonics that
matter for this path -- FP64 arithmetic (DFMA / DMUL / DADD / DSETP), FP32 gate arithmetic (FFMA / FMUL / FADD), MUFU, global and shared
memory traffic, barriers, warp votes / shuffles, atomics. Written to profiles/ so the instruction mix of the build under test is on
record next to the ncu numbers.    python scripts/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "sc_b200", "libscgpu.so")
GROUPS = [("DFMA", r"^DFMA"), ("DMUL", r"^DMUL"), ("DADD", r"^DADD"), ("DSETP", r"^DSETP"), ("FFMA", r"^FFMA"), ("FMUL/FADD", r"^F(MUL|ADD)"),
          ("FSETP", r"^FSETP"), ("MUFU", r"^MUFU"), ("LDG", r"^LDG"), ("STG", r"^STG"), ("LDS", r"^LDS"), ("STS", r"^STS"), ("LDC", r"^LDC"),
          ("BAR", r"^BAR"), ("VOTE/MATCH", r"^(VOTE|MATCH|VOTEU)"), ("SHFL", r"^SHFL"), ("ATOM/RED", r"^(ATOM|ATOMS|ATOMG|RED|REDG)"),
          ("LDL/STL", r"^(LDL|STL)"), ("BRA", r"^BRA")]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(2)
            cur["total"] += 1
            for name, pat in GROUPS:
                if re.match(pat, op):
                    cur[name] += 1
    demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.split("\n")
    print("# %s -- cuobjdump -sass, static instruction counts per kernel (sm_100a)" % os.path.relpath(LIB, ROOT))
    print("%-44s %7s " % ("kernel", "instr") + " ".join("%9s" % g[0] for g in GROUPS))
    for (k, c), d in zip(kernels.items(), demangle):
        name = re.sub(r"^void ", "", d) if d else k
        cut = name.rfind(">(")
        name = name[:cut + 1] if cut >= 0 else re.sub(r"\(.*", "", name)
        name = name.replace("(bool)", "").replace("(int)", "")
        print("%-44s %7d " % (name[:44], c["total"]) + " ".join("%9d" % c[g[0]] for g in GROUPS))


if __name__ == "__main__":
    main()
