#!/usr/bin/env python
"""Join an ncu source-page dump (SASS rows with stall samples / instructions executed) with the line table of the shipped
cubin, and print the hottest CUDA source lines of one kernel (profiling aid).

    ncu -i rep.ncu-rep --page source --csv > src.csv
    python tools/ncu_lines.py src.csv whale.jl_b200/libwhalecuda.so _Z8k_dp_revILi128ELi4EEv7RevArgs [top]
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def line_table(so, mangled):
    d = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
    on, cur, tab = False, ("?", 0), []
    for ln in txt.splitlines():
        if ln.startswith(".text."):
            on = ln.strip() == f".text.{mangled}:"
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
        if m:
            tab.append((int(m.group(1), 16), cur, m.group(2)))
    return tab


def main():
    src, so, mangled = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    H = rows[hi]
    ci = {n: H.index(n) for n in ("Address", "# Samples", "Instructions Executed", "L1 Wavefronts Shared", "stall_barrier",
                                  "stall_long_sb", "stall_short_sb", "stall_wait")}
    body = [r for r in rows[hi + 1:] if len(r) > ci["stall_wait"]]
    base = int(body[0][ci["Address"]], 16)
    tab = {off: (loc, ins) for off, loc, ins in line_table(so, mangled)}
    agg = defaultdict(lambda: [0, 0, 0, 0, 0, 0, 0])
    tot = [0, 0, 0]
    for r in body:
        off = int(r[ci["Address"]], 16) - base
        loc = tab.get(off, (("?", 0), ""))[0]
        v = [int(float(r[ci[k]] or 0)) for k in ("# Samples", "Instructions Executed", "L1 Wavefronts Shared", "stall_barrier",
                                                 "stall_long_sb", "stall_short_sb", "stall_wait")]
        for j in range(7):
            agg[loc][j] += v[j]
        for j in range(3):
            tot[j] += v[j]
    print(f"total samples {tot[0]}  warp instructions {tot[1]}  shared wavefronts {tot[2]}")
    print(f"{'file:line':28s} {'samples':>8s} {'%':>6s} {'inst':>10s} {'%':>6s} {'sh.wave':>9s} {'barrier':>8s} {'long_sb':>8s} {'short_sb':>8s} {'wait':>6s}")
    for loc, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{loc[0] + ':' + str(loc[1]):28s} {v[0]:8d} {100 * v[0] / max(tot[0], 1):6.2f} {v[1]:10d} {100 * v[1] / max(tot[1], 1):6.2f} "
              f"{v[2]:9d} {v[3]:8d} {v[4]:8d} {v[5]:8d} {v[6]:6d}")


if __name__ == "__main__":
    main()
