#!/usr/bin/env python
"""Summarise ncu artefacts into the text files kept under profiles/.

    python profiles/ncu_summary.py launches <launches.csv>
    python profiles/ncu_summary.py kernel   <report.ncu-rep> [--top N]

`launches` aggregates the gpu__time_duration.sum launch list per kernel (cold-cache, serialised:
compare SHARES).  `kernel` prints, per profiled launch, the roofline-relevant raw metrics and the
top stall sites from the source page (needs -lineinfo)."""
import collections
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
       "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
       "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
       "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
       "sm__cycles_elapsed.max"]
STALLS = ["barrier", "long_scoreboard", "short_scoreboard", "mio_throttle", "math_pipe_throttle", "wait",
          "not_selected", "dispatch_stall", "lg_throttle", "no_instruction", "branch_resolving", "membar", "sleeping"]


def launches(path):
    rows = [l for l in open(path) if l.startswith('"')]
    r = list(csv.DictReader(io.StringIO("".join(rows))))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for x in r:
        v = float(x["Metric Value"].replace(",", ""))
        v = v / 1000 if x["Metric Unit"] == "ns" else v * 1000 if x["Metric Unit"] == "ms" else v
        k = x["Kernel Name"].split("(")[0][-60:]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(t for _, t in agg.values())
    print(f"{'share':>7s} {'total_us':>12s} {'n':>6s} {'avg_us':>10s}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / tot * 100:6.2f}% {t:12.1f} {n:6d} {t / n:10.2f}  {k}")


def kernel(path, top=14):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    h, u = r[0], r[1]
    for row in r[2:]:
        print("=" * 100)
        print("kernel:", row[h.index("Kernel Name")][:90])
        for i, n in enumerate(h):
            if n in RAW:
                print(f"  {n:82s} {row[i]:>16s} {u[i]}")
        for s in STALLS:
            n = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if n in h:
                print(f"  stall/{s:30s} {float(row[h.index(n)]):8.3f} warps per issue-active cycle")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks = src.split('"Kernel Name",')
    seen = set()
    for blk in blocks[1:]:
        if blk.split("\n")[0] in seen:
            continue
        seen.add(blk.split("\n")[0])
        lines = blk.split("\n")
        print("-" * 100)
        print("source page:", lines[0][:90])
        rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
        if len(rows) < 2:
            continue
        hh = rows[0]
        ix = {n: i for i, n in enumerate(hh)}

        def f(x):
            try:
                return float(x)
            except ValueError:
                return 0.0
        data = [x for x in rows[1:] if len(x) == len(hh)]
        tot = sum(f(x[ix["# Samples"]]) for x in data) or 1.0
        for x in sorted(data, key=lambda x: -f(x[ix["# Samples"]]))[:top]:
            st = {n[6:]: f(x[ix[n]]) for n in hh if n.startswith("stall_") and "Not Issued" not in n and f(x[ix[n]]) > 0}
            s = ", ".join(f"{k}={int(v)}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
            print(f"  {f(x[ix['# Samples']]) / tot * 100:5.2f}%  exec={x[ix['Instructions Executed']]:>9s}  {x[ix['Source']][:64]:64s} {s}")
        cls, ex = collections.Counter(), collections.Counter()
        for x in data:
            t = x[ix["Source"]].split()
            op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else t[0] if t else "?").split(".")[0]
            cls[op] += f(x[ix["# Samples"]])
            ex[op] += f(x[ix["Instructions Executed"]])
        te = sum(ex.values()) or 1.0
        print("  opcode      stall-samples%   executed%")
        for op, c in cls.most_common(12):
            print(f"  {op:10s} {c / tot * 100:12.2f}% {ex[op] / te * 100:10.2f}%")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        kernel(sys.argv[2], int(sys.argv[4]) if len(sys.argv) > 4 else 14)
