#!/usr/bin/env python
"""Per-source-line share of executed warp instructions and stall samples of the first kernel in an ncu report.
usage: tools/ncu_lines.py <prof.ncu-rep> [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout.splitlines()))
cur = None; rows = []
for r in src:
    if r and r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if len(r) > 8 and r[0] not in ('Line No', ''):
        try:
            rows.append((cur, int(r[0]), r[1].strip(), int(r[6] or 0), int(r[7] or 0)))
        except ValueError:
            pass
tots = sum(x[3] for x in rows); toti = sum(x[4] for x in rows)
print('samples', tots, 'warp-inst', toti)
for f, l, s, sa, ie in sorted(rows, key=lambda x: -x[4])[:top]:
    print(f"{f}:{l} inst {100 * ie / toti:5.1f}% samp {100 * sa / tots:5.1f}%  {s[:110]}")
