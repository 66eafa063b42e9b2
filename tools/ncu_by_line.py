#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` SASS listing per CUDA source line.

    cuobjdump -xelf all libicpflow_b200.so ; nvdisasm -g -c icpf_icp.sm_100a.cubin > icp.sass
    ncu -i prof.ncu-rep --page source --csv > src.csv
    python tools/ncu_by_line.py src.csv icp.sass '<mangled kernel name>' [top]

Joins the two listings by instruction offset (the ncu CLI has no per-line metric view) and prints, per source line,
warp-level instructions executed, their share, average active threads, stall samples and shared-memory wavefronts.
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    src_csv, sass, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    # offsets -> (file, line) from nvdisasm
    loc = {}
    cur = None
    active = False
    for ln in open(sass):
        if ln.startswith("//---") and ".text." in ln:
            active = kernel in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m:
            loc[int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(src_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {n: i for i, n in enumerate(hdr)}
    data = rows[hdr_i + 1:]
    base = int(data[0][0], 16)
    agg = defaultdict(lambda: [0, 0, 0, 0, 0])
    tot = 0
    for r in data:
        off = int(r[0], 16) - base
        key = loc.get(off, ("?", 0))
        inst = int(r[col["Instructions Executed"]])
        thr = int(r[col["Thread Instructions Executed"]])
        smp = int(r[col["# Samples"]])
        wf = int(r[col["L1 Wavefronts Shared"]] or 0)
        wfi = int(r[col["L1 Wavefronts Shared Ideal"]] or 0)
        a = agg[key]
        a[0] += inst; a[1] += thr; a[2] += smp; a[3] += wf; a[4] += wfi
        tot += inst
    tot_s = sum(a[2] for a in agg.values())
    print(f"total warp instructions {tot}, samples {tot_s}")
    print(f"{'file:line':34s} {'inst':>10s} {'%inst':>6s} {'thr/inst':>8s} {'%smp':>6s} {'smem wf':>9s} {'ideal':>9s}")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{key[0] + ':' + str(key[1]):34s} {a[0]:10d} {100 * a[0] / tot:6.2f} {a[1] / max(a[0], 1):8.1f} "
              f"{100 * a[2] / max(tot_s, 1):6.2f} {a[3]:9d} {a[4]:9d}")


if __name__ == "__main__":
    main()
