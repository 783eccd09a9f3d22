#!/usr/bin/env python
"""Turn an `ncu --set full` report (or a launch-list CSV) into the small text summaries kept under profiles/.

    python tools/summarize_ncu.py report  gpurun_out/prof.ncu-rep  profiles/r01_full_<name>.md
    python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches_<name>.md
Runs on the CPU box (ncu -i needs no GPU)."""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__warps_active.avg.per_cycle_active", "active warps / scheduler"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / scheduler"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__waves_per_multiprocessor", "waves / SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "stall long scoreboard %"),
]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def report(rep, dst, alg_bytes=None):
    hdr, units, rows = raw_rows(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu --set full summary of `{rep.split('/')[-1]}`", "",
             "Captured with `ncu --set full --clock-control none --import-source on` on one B200 (cold cache, serialised:",
             "compare shares and counters, not absolute times, with the CUDA-event numbers of bench.py).", ""]
    traffic = {}
    for r in rows:
        name = r[idx["Kernel Name"]]
        lines.append(f"## {name}")
        lines.append("")
        lines.append("| metric | value |")
        lines.append("|---|---|")
        vals = {}
        for key, label in KEYS:
            if key in idx:
                vals[key] = r[idx[key]]
                lines.append(f"| {label} (`{key}`) | {r[idx[key]]} {units[idx[key]]} |")
        stalls = []
        for h, i in idx.items():
            if "issue_stalled" in h and h.endswith("per_warp_active.pct"):
                try:
                    stalls.append((float(r[i]), h.split("issue_stalled_")[1].replace("_per_warp_active.pct", "")))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        if stalls:
            lines.append("| top warp stall reasons (% of warp-active cycles) | " + ", ".join(f"{n} {v:.1f}" for v, n in stalls[:5]) + " |")
        try:
            def to_bytes(key):
                v, u = float(r[idx[key]]), units[idx[key]].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
            t = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
            lines.append(f"| DRAM traffic read+write | {t/1e6:.1f} MB |")
            short = name.split("<")[0].replace("void ", "").replace("cplb::", "").strip()
            traffic.setdefault(short, t)
        except Exception:
            pass
        lines.append("")
    open(dst, "w").write("\n".join(lines) + "\n")
    return traffic


def launches(csv_path, dst):
    rows = [r for r in csv.reader(open(csv_path, errors="replace")) if r]
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    per = {}
    for r in rows[start + 1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        name = r[ki].split("(")[0]
        per.setdefault(name, []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in per.values())
    lines = [f"# ncu launch list of `{csv_path.split('/')[-1]}` (`--metrics gpu__time_duration.sum --clock-control none`)", "",
             "Per-launch times under ncu are cold-cache and serialised: the SHARE of the step is what is comparable.", "",
             "| kernel | launches | total ns | share | avg ns | min ns | max ns |", "|---|---|---|---|---|---|---|"]
    for name, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"| `{name}` | {len(v)} | {sum(v):.0f} | {100*sum(v)/tot:.1f}% | {sum(v)/len(v):.0f} | {min(v):.0f} | {max(v):.0f} |")
    open(dst, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    if mode == "report":
        t = report(src, dst)
        print(json.dumps(t))
    else:
        launches(src, dst)
