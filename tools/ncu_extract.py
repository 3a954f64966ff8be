#!/usr/bin/env python
"""Condenses Nsight Compute reports into the small text files kept under profiles/.

    python tools/ncu_extract.py raw   <report.ncu-rep> <out.csv> [i]  selected metrics, one row per metric (kernel i of the report; default: the last)
    python tools/ncu_extract.py hot   <report.ncu-rep> <out.txt> [N] stall totals + N hottest SASS lines
    python tools/ncu_extract.py launches <launches.csv> <out.txt>    per-kernel launch count / time / share
"""
import collections
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg",
    "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_active.avg",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def raw(rep, dst, which=-1):
    rows = ncu_csv(rep, "raw")
    hdr, units, vals = rows[0], rows[1], (rows[2:])[which]
    with open(dst, "w") as f:
        f.write("metric,unit,value\n")
        for k in ("Kernel Name", "Block Size", "Grid Size"):
            if k in hdr:
                f.write('%s,,"%s"\n' % (k, vals[hdr.index(k)]))
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                f.write("%s,%s,%s\n" % (k, units[i], vals[i]))


def hot(rep, dst, n=25):
    rows = ncu_csv(rep, "source")
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    recs, samples, insts = [], 0, 0
    for r in data:
        try:
            s, e = int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]])
        except (ValueError, IndexError):
            continue
        samples += s
        insts += e
        for st in stalls:
            try:
                tot[st] += int(r[ix[st]])
            except ValueError:
                pass
        recs.append((s, e, r[ix["Address"]][-6:], r[ix["Source"]].strip()))
    recs.sort(reverse=True)
    with open(dst, "w") as f:
        f.write("%s\n" % rows[0][1] if len(rows[0]) > 1 else "")
        f.write("warp-level samples %d, warp instructions executed %d\n" % (samples, insts))
        f.write("stall reasons (samples): " + ", ".join("%s %d" % (k[6:], v) for k, v in tot.most_common(8)) + "\n")
        f.write("samples  executed  addr    SASS\n")
        for s, e, a, src in recs[:n]:
            f.write("%7d %9d %s %s\n" % (s, e, a, src[:110]))


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if r and r[0].strip('"').isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("<unnamed>::", "")
        ns = float(r[-1])
        unit = r[-2]
        ms = ns / 1e6 if unit in ("ns", "nsecond") else (ns / 1e3 if unit in ("us", "usecond") else ns)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    step = sum(v[1] for k, v in agg.items() if "synth" not in k and "at::" not in k and "elementwise" not in k)
    with open(dst, "w") as f:
        f.write("kernel, launches, total_ms, mean_ms, share_of_step (excl. synthetic-data and torch fill kernels)\n")
        for k, (n, ms) in agg.items():
            own = "synth" not in k and "at::" not in k and "elementwise" not in k
            f.write("%s, %d, %.3f, %.4f, %s\n" % (k, n, ms, ms / n, ("%.3f" % (ms / step)) if own else "-"))


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "raw":
        raw(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else -1)
    elif cmd == "hot":
        hot(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 25)
    elif cmd == "launches":
        launches(sys.argv[2], sys.argv[3])
