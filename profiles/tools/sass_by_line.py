#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS export with `nvdisasm -g` line info and aggregate executed
warp instructions / stall samples per CUDA source line.

usage: sass_by_line.py <ncu_source.csv> <nvdisasm_-g_-c_output.sass> <mangled-function-substring> [top] [--outer]
"""
import csv
import re
import sys
from collections import defaultdict


def line_map(sass_path, func, outer=False):
    """offset -> (file, line).  With `nvdisasm -gi` output every instruction carries its inline
    chain; outer=True keeps the outermost frame (the call site in the kernel body)."""
    m, cur, infunc, group, in_group = {}, None, False, [], False
    for ln in open(sass_path, errors="ignore"):
        if ln.startswith("//---") and ".text." in ln:
            infunc = func in ln
            continue
        if not infunc:
            continue
        mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if mm:
            if not in_group:
                group, in_group = [], True
            group.append((mm.group(1).split("/")[-1], int(mm.group(2))))
            continue
        mm = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if mm:
            if in_group:
                cur = group[-1] if outer else group[0]
                in_group = False
            m[int(mm.group(1), 16)] = cur
    return m


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    outer = "--outer" in sys.argv
    argv = [a for a in sys.argv if a != "--outer"]
    lm = line_map(argv[2], argv[3], outer)
    top = int(argv[4]) if len(argv) > 4 else 40
    hi = next(i for i, r in enumerate(rows[:6]) if "Source" in r)
    ci = {h: i for i, h in enumerate(rows[hi])}
    base = None
    agg = defaultdict(lambda: [0, 0, defaultdict(int)])
    tot = smp = 0
    for r in rows[hi + 1:]:
        try:
            addr = int(r[ci["Address"]], 16)
            n = int(r[ci["Instructions Executed"]])
            s = int(r[ci["# Samples"]])
        except (ValueError, KeyError):
            continue
        if base is None:
            base = addr
        key = lm.get(addr - base)
        agg[key][0] += n
        agg[key][1] += s
        agg[key][2][r[ci["Source"]].split()[0].lstrip("@!P0123456789 ")] += n
        tot += n
        smp += s
    print(f"total warp instructions {tot}, samples {smp}")
    src = {}
    for key, (n, s, ops) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        txt = ""
        if key:
            if key[0] not in src:
                try:
                    import glob
                    path = glob.glob(f"**/{key[0]}", recursive=True)[0]
                    src[key[0]] = open(path).read().splitlines()
                except Exception:
                    src[key[0]] = []
            if key[1] - 1 < len(src[key[0]]):
                txt = src[key[0]][key[1] - 1].strip()[:90]
        opsd = ",".join(f"{k}:{v * 100 // max(1, n)}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:4])
        print(f"{n:>11} {100 * n / tot:5.1f}%  smp {100 * s / max(1, smp):5.1f}%  {key}  {txt}   [{opsd}]")


if __name__ == "__main__":
    main()
