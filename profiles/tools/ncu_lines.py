#!/usr/bin/env python
"""Join `ncu --page source --print-source sass --csv` (per-instruction counters) with
`nvdisasm -g -c` line info of the same cubin, and print executed warp-instructions per source
line (innermost inlined location and outermost kernel-file line).

usage: ncu_lines.py <sass.csv> <nvdisasm.txt> <mangled-kernel-substring> [top_n]
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    sass_csv, dis, kname = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    # --- nvdisasm: list of (opcode text, innermost (file,line), outermost kernel-file line)
    ins = []
    cur_inner, cur_outer, active = None, None, False
    for line in open(dis, errors="replace"):
        if line.startswith(".text."):
            active = kname in line
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', line)
        if m:
            cur_inner = (m.group(1).split("/")[-1], int(m.group(2)))
            chain = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            cur_outer = (chain[-1][0].split("/")[-1], int(chain[-1][1])) if chain else cur_inner
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((m.group(2).strip(), cur_inner, cur_outer))
    # --- ncu sass page (first kernel instance only)
    rows = list(csv.reader(open(sass_csv)))
    hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    start = hdr_i[0]
    end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
    hdr = rows[start]
    ci = hdr.index("Instructions Executed")
    cs = hdr.index("# Samples")
    body = [r for r in rows[start + 1:end] if len(r) > ci and r[0].startswith("0x")]
    if len(body) != len(ins):
        print(f"warning: ncu has {len(body)} instructions, nvdisasm {len(ins)}", file=sys.stderr)
    inner, outer, samples = defaultdict(int), defaultdict(int), defaultdict(int)
    total = 0
    for (op, li, lo), r in zip(ins, body):
        n = int(r[ci]); total += n
        inner[li] += n; outer[lo] += n; samples[lo] += int(r[cs] or 0)
    print(f"total warp-instructions executed: {total}")
    print("--- by outermost location (line in the kernel source)")
    for k, v in sorted(outer.items(), key=lambda kv: -kv[1])[:top]:
        print(f"{v:12d} {100.0 * v / total:5.1f}%  samples {samples[k]:6d}  {k}")
    print("--- by innermost (inlined) location")
    for k, v in sorted(inner.items(), key=lambda kv: -kv[1])[:top]:
        print(f"{v:12d} {100.0 * v / total:5.1f}%  {k}")


if __name__ == "__main__":
    main()
