#!/usr/bin/env python
"""Summarise ncu artefacts brought back in gpurun_out/ into small tracked text files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_X.csv  profiles/launches_X.md
    python tools/ncu_summary.py report   gpurun_out/prof_X.ncu-rep  profiles/prof_X.md

`launches`: per-kernel totals / shares of one `--metrics gpu__time_duration.sum` launch list.
`report`  : key raw metrics of every launch in an `--set full` report (read with `ncu -i ... --page raw --csv`).
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "uniform pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard", "stall long scoreboard"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard /issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier /issue"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar /issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard /issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle /issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait /issue"),
    ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall sleeping /issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle /issue"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction /issue"),
]


def short(name):
    name = name.replace("void ", "").replace("rnamsm::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    return name.split("(")[0]


def launches(src, dst):
    rows = []
    with open(src) as f:
        text = f.read()
    text = text[text.index('"ID"'):]
    for r in csv.DictReader(io.StringIO(text)):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], float(r["Metric Value"].replace(",", ""))))
    tot = sum(r[3] for r in rows)
    agg = OrderedDict()
    for k, g, b, ns in rows:
        a = agg.setdefault(k, [0, 0.0, g, b])
        a[0] += 1
        a[1] += ns
    with open(dst, "w") as f:
        f.write(f"# ncu launch list: {src}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: "
                f"compare SHARES, not absolutes).  {len(rows)} launches, {tot / 1e6:.3f} ms total.\n\n")
        f.write("| kernel | launches | total us | share | avg us | grid | block |\n|---|---:|---:|---:|---:|---|---|\n")
        for k, (n, ns, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {ns / 1e3:.1f} | {100 * ns / tot:.1f}% | {ns / 1e3 / n:.1f} | {g} | {b} |\n")
        f.write("\n## in launch order (first 70)\n\n| # | kernel | us |\n|---:|---|---:|\n")
        for i, (k, g, b, ns) in enumerate(rows[:70]):
            f.write(f"| {i} | `{k}` | {ns / 1e3:.1f} |\n")
    print(dst)


def report(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full: {src}\n\nRead with `ncu -i <rep> --page raw --csv`; one block per captured launch.\n")
        for r in rows[2:]:
            f.write(f"\n## `{short(r[col['Kernel Name']])}`  grid {r[col['Grid Size']]} block {r[col['Block Size']]}\n\n")
            f.write("| metric | value | unit |\n|---|---:|---|\n")
            for key, label in KEYS:
                if key in col and r[col[key]] != "":
                    f.write(f"| {label} (`{key}`) | {r[col[key]]} | {units[col[key]]} |\n")
    print(dst)


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2], sys.argv[3])
