#!/usr/bin/env python
"""Attribute ncu warp-stall samples to CUDA source lines without a GUI.

ncu's `--page source --csv` lists SASS instructions with their sample counts but no line numbers; nvdisasm -g lists
the same instructions with `//## File "...", line N` markers.  The two listings are matched by position inside the
kernel.  Usage: ncu_lines.py <report.ncu-rep> <kernel regex> <cubin> [top N]
"""
import csv
import io
import re
import subprocess
import sys


def main():
    rep, kern, cubin = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # first kernel instance only
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[start]
    si = hdr.index("# Samples")
    sass = []
    for r in rows[start + 1:]:
        if not r or r[0] == "Kernel Name":
            break
        if len(r) > si and r[0].startswith("0x"):
            sass.append((r[1].strip(), int(r[si] or 0)))
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    lines = []
    cur = None
    infn = False
    for l in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            infn = re.search(kern, m.group(1)) is not None
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", l)
        if m:
            lines.append(cur)
    if len(lines) != len(sass):
        print("warning: %d SASS rows in the report vs %d in the cubin" % (len(sass), len(lines)))
    agg = {}
    tot = 0
    for (ins, s), ln in zip(sass, lines):
        agg[ln] = agg.get(ln, 0) + s
        tot += s
    print("total samples", tot)
    src_cache = {}
    for ln, s in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
        text = ""
        if ln:
            import glob
            if ln[0] not in src_cache:
                c = glob.glob("/root/repo/ppg_slam_b200/csrc/" + ln[0])
                src_cache[ln[0]] = open(c[0]).read().splitlines() if c else []
            if 0 < ln[1] <= len(src_cache[ln[0]]):
                text = src_cache[ln[0]][ln[1] - 1].strip()
        print("%6d %5.1f%%  %s  %s" % (s, 100.0 * s / max(tot, 1), ln, text[:110]))


if __name__ == "__main__":
    main()
