#!/usr/bin/env python
"""Registers / spills / static shared memory of every kernel (ptxas -v), and the TMA / mbarrier / atomics SASS
mnemonics per kernel (cuobjdump -sass) of the sm_100a build. Runs without a GPU.

  python profiles/tools/res_usage.py            -> table on stdout (commit it as profiles/rNN_sass_summary.txt)
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, "cellranger_b200", "csrc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"]
MNEMONICS = ["UBLKCP", "SYNCS", "ATOMS", "ATOMG", "RED", "MATCH", "VOTE", "LDG.E.128", "STG.E.128", "NANOSLEEP"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    for src in sorted(f for f in os.listdir(CSRC) if f.endswith(".cu")):
        obj = f"/tmp/res_{src}.o"
        r = subprocess.run(["nvcc"] + FLAGS + ["-Xptxas", "-v", "-c", os.path.join(CSRC, src), "-o", obj],
                           capture_output=True, text=True)
        if r.returncode:
            print(r.stderr)
            sys.exit(1)
        rows, name, spill = [], None, ""
        for line in r.stderr.splitlines():
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                name, spill = m.group(1), ""
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and int(m.group(2)):
                spill = f"spill {m.group(2)}/{m.group(3)} B"
            m = re.search(r"Used (\d+) registers(.*)", line)
            if m and name:
                sm = re.search(r"(\d+) bytes smem", m.group(2))
                rows.append((name, int(m.group(1)), int(sm.group(1)) if sm else 0, spill))
                name = None
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        counts, cur = {}, None
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = m.group(1)
                counts[cur] = {k: 0 for k in MNEMONICS}
                continue
            if cur:
                for k in MNEMONICS:
                    if re.search(r"\b" + re.escape(k), line):
                        counts[cur][k] += 1
        dm = demangle([r_[0] for r_ in rows])
        print(f"== {src}")
        for nm, regs, smem, sp in rows:
            c = counts.get(nm, {})
            tags = " ".join(f"{k}={v}" for k, v in c.items() if v)
            short = re.sub(r"\(.*", "", dm.get(nm, nm).replace("(anonymous namespace)::", "")).replace("void ", "")
            print(f"  {short:<58} regs {regs:>3} smem {smem:>6} {sp:<18} {tags}")


if __name__ == "__main__":
    main()
