#!/usr/bin/env python
"""Attribute an ncu SASS-level source page to CUDA source lines.

ncu's CLI exports per-SASS-instruction metrics (`--page source --csv`) but not per-CUDA-line ones, so this
joins them with `nvdisasm -g` line info of the same cubin by instruction order.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep 'mesh_chunks_kernelILi1ELb0' [--top 40]
"""
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "despawnengine_b200", "lib", "libdespawn_b200.so")


def disasm_lines(mangled_substr):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=d, check=True, capture_output=True)
    cubin = glob.glob(os.path.join(d, "*.cubin"))[0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    out, active, cur = [], False, ("?", 0)
    for ln in txt:
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            active = mangled_substr in m.group(1)
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((cur, m.group(2).strip()))
    return out


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    csv_txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    # the page is a sequence of per-launch sections, each opened by a ["Kernel Name", <demangled name>] row
    sections, cur = [], None
    for r in csv.reader(csv_txt.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = [r[1] if len(r) > 1 else "", []]
            sections.append(cur)
        elif cur is not None and len(r) > 6 and r[2].isdigit():
            cur[1].append(r)
    dis = disasm_lines(kern)
    n = len(dis)
    rows = next((rs for name, rs in sections if len(rs) == n), None)  # first launch whose SASS length matches the cubin's
    assert n and rows is not None, (n, [(name[:40], len(rs)) for name, rs in sections])
    by_line = defaultdict(lambda: [0, 0, 0])  # samples, warp-instr, sass count
    for (loc, _), r in zip(dis, rows):
        b = by_line[loc]
        b[0] += int(r[2]); b[1] += int(r[5]); b[2] += 1
    tot_s = sum(b[0] for b in by_line.values()); tot_i = sum(b[1] for b in by_line.values())
    src = {}
    print(f"kernel {kern}: {n} SASS instr, {tot_i} warp-instr executed, {tot_s} stall samples")
    print(f"{'file:line':28s} {'samples':>8s} {'%':>6s} {'warp-instr':>11s} {'%':>6s}  source")
    for loc, b in sorted(by_line.items(), key=lambda kv: -max(kv[1][0] / max(tot_s, 1), kv[1][1] / max(tot_i, 1)))[:top]:
        f = os.path.join(ROOT, "despawnengine_b200", "csrc", loc[0])
        if f not in src and os.path.exists(f):
            src[f] = open(f).read().splitlines()
        text = src.get(f, [""] * (loc[1] + 1))[loc[1] - 1].strip()[:70] if f in src and loc[1] - 1 < len(src[f]) else ""
        print(f"{loc[0] + ':' + str(loc[1]):28s} {b[0]:8d} {100 * b[0] / max(tot_s, 1):6.1f} {b[1]:11d} {100 * b[1] / max(tot_i, 1):6.1f}  {text}")


if __name__ == "__main__":
    main()
