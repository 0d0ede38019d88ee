#!/usr/bin/env python
"""Per-source-line instruction breakdown of the first kernel of an ncu report (needs -lineinfo and --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [top=40]"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
rows = list(csv.reader(raw.splitlines()))
hdr = rows[2]
iL, iS, iI, iT, iM = hdr.index("Line No"), 1, hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
tot, lines = 0, []
for r in rows[3:]:
    if len(r) <= iT or r[iL] == "":
        continue
    try:
        n, t, sm = int(r[iI]), int(r[iT]), int(r[iM])
    except ValueError:
        continue
    lines.append((int(r[iL]), r[iS], n, t, sm)); tot += n
print("warp instructions: %d" % tot)
for ln, src, n, t, sm in sorted(lines, key=lambda x: -x[2])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%5d %5.1f%% lanes %4.1f samples %5d  %s" % (ln, 100.0 * n / tot, t / max(n, 1), sm, src.strip()[:120]))
