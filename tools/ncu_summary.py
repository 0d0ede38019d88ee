#!/usr/bin/env python
"""Turn ncu output (run under gpurun, brought back in gpurun_out/) into the small text summaries kept under profiles/.

  python tools/ncu_summary.py launches <launches.csv>            -> per-kernel table (count, total, share, average)
  python tools/ncu_summary.py kernel <report.ncu-rep> [out.json] -> key metrics of the first kernel in the report
  python tools/ncu_summary.py dram <metrics.csv> [peak_GB/s]     -> per-kernel DRAM bytes / duration against the HBM peak (csv of a run with
                                                                    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum)
"""
import collections
import csv
import json
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for row in r:
        if len(row) <= vi:
            continue
        name = re.sub(r"\(.*", "", row[ki]).replace("void ", "")
        v = float(row[vi].replace(",", ""))
        v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row[ui], 1)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, v in tot.most_common():
        print("| %s | %d | %.3f | %.1f%% | %.1f |" % (k, cnt[k], v / 1e6, 100 * v / T, v / cnt[k] / 1e3))
    print("| **sum** | %d | %.3f | | |" % (sum(cnt.values()), T / 1e6))


def dram(path, peak=6536.7):
    lines = [l for l in open(path) if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ii, ki, mi, vi, ui = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    per = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])      # launches, ns, bytes read, bytes written
    seen = set()
    scale_t = {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}
    scale_b = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "B": 1, "KB": 1e3, "MB": 1e6, "GB": 1e9}
    for row in r:
        if len(row) <= vi:
            continue
        name = re.sub(r"\(.*", "", row[ki]).replace("void ", "")
        v = float(row[vi].replace(",", ""))
        e = per[name]
        if (row[ii], name) not in seen:
            seen.add((row[ii], name)); e[0] += 1
        if row[mi] == "gpu__time_duration.sum":
            e[1] += v * scale_t.get(row[ui], 1)
        elif row[mi] == "dram__bytes_read.sum":
            e[2] += v * scale_b.get(row[ui], 1)
        elif row[mi] == "dram__bytes_write.sum":
            e[3] += v * scale_b.get(row[ui], 1)
    print("| kernel | launches (of the first %d) | total us | DRAM read MB | DRAM written MB | DRAM GB/s | of %.0f GB/s |\n|---|---:|---:|---:|---:|---:|---:|" % (len(seen), peak))
    for k, e in sorted(per.items(), key=lambda kv: -kv[1][1]):
        gbs = (e[2] + e[3]) / e[1] if e[1] else 0.0               # bytes per ns = GB/s
        print("| %s | %d | %.1f | %.1f | %.1f | %.0f | %.1f %% |" % (k, e[0], e[1] / 1e3, e[2] / 1e6, e[3] / 1e6, gbs, 100 * gbs / peak))


def kernel(rep, out=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    rows = list(csv.reader(raw.splitlines()))
    h, u, v = rows[0], rows[1], rows[2]
    res = {"kernel": v[h.index("Kernel Name")]}
    for i, n in enumerate(h):
        if n in KEYS:
            res[n] = {"value": v[i], "unit": u[i]}
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    srows = list(csv.reader(src.splitlines()))[2:]
    ops = collections.Counter()
    for r in srows:
        t = r[1].split()
        if t:
            ops[t[1] if t[0].startswith("@") and len(t) > 1 else t[0]] += int(r[5])
    res["sass_opcodes_executed"] = {k: ops[k] for k in ops if k.startswith(("UBLKCP", "SYNCS", "UTMA", "ATOMS", "LDS", "LDG", "STG", "DFMA", "DMUL"))}
    res["warp_instructions_total"] = sum(ops.values())
    txt = json.dumps(res, indent=1)
    if out:
        open(out, "w").write(txt + "\n")
    print(txt)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "dram":
        dram(sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 6536.7)
    else:
        kernel(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
