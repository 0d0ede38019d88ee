#!/usr/bin/env python
"""Reads the A/B log of tools/sessions/gpu_r2q.sh (one tools/profile_run.py run per PGX_SWEEP_VARIANT) and prints the variant with the smallest median
device time per pass (first pass of every run dropped: cold caches and first-touch allocations)."""
import re
import statistics
import sys

best, cur, res = None, None, {}
for line in open(sys.argv[1]):
    m = re.match(r"== PGX_SWEEP_VARIANT=(\d+)", line)
    if m:
        cur = int(m.group(1))
    elif line.startswith("per pass") and cur is not None:
        t = [float(a) + float(b) for a, b in re.findall(r"([0-9.]+)\+([0-9.]+)", line)]
        s1 = [float(a) for a, b in re.findall(r"([0-9.]+)\+([0-9.]+)", line)]
        if len(t) > 1:
            res[cur] = (statistics.median(t[1:]), statistics.median(s1[1:]))
    elif line.startswith("hits ") and cur is not None and cur in res:
        m = re.search(r"sweep_ms ([0-9.]+)", line)
        res[cur] = res[cur] + (float(m.group(1)) if m else -1.0,)
for v in sorted(res):
    print("variant %d: median %.3f ms per pass (stage 1 %.3f), first sweep %.3f ms" % ((v,) + res[v]), file=sys.stderr)
if not res:
    sys.exit(1)
print(min(res, key=lambda v: res[v][0]))
