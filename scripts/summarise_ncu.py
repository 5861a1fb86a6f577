#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep and launch CSVs into the small tracked summaries under profiles/.

    python scripts/summarise_ncu.py <report.ncu-rep> <out.md> [title]
    python scripts/summarise_ncu.py --launches <launches.csv> <out.md> [title]
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
]


def report(path, out, title):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# {title}", "", f"source: `{path}` (ncu --set full --clock-control none)", ""]
    for data in rows[2:]:
        d = {h: (data[i], units[i]) for i, h in enumerate(hdr)}
        lines += [f"## {d.get('Kernel Name', ('?', ''))[0][:100]}", "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in d:
                lines.append(f"| {k} | {d[k][0]} | {d[k][1]} |")
        stalls = []
        for h, (v, _) in d.items():
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    stalls.append((float(v), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        lines += ["", "warp stall reasons (warps per issue-active cycle): "
                  + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:7]), ""]
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    if len(rows) > 2:
        hdr = rows[1]
        isrc, iex, ist = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
        ops, samp = collections.Counter(), collections.Counter()
        seen, skip = {rows[0][1] if len(rows[0]) > 1 else ""}, False
        for r in rows[2:]:
            if r and r[0] == "Kernel Name":
                # a report with several kernels lists one section per result (and may list a kernel
                # twice): the opcode mix is that of the FIRST kernel only
                skip = True
                continue
            if skip or len(r) <= max(isrc, iex, ist):
                continue
            toks = r[isrc].split()
            if not toks:
                continue
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            try:
                ops[op] += int(r[iex] or 0)
                samp[op] += int(r[ist] or 0)
            except ValueError:
                continue                   # a repeated header row
        lines += ["SASS opcodes by warp-instructions executed: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(10)), "",
                  "SASS opcodes by stall samples: " + ", ".join(f"{k} {v}" for k, v in samp.most_common(8)), ""]
    open(out, "w").write("\n".join(lines) + "\n")


def launches(path, out, title):
    lines_in = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines_in):
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6}.get(row["Metric Unit"], 1.0)
        agg[name][0] += 1
        agg[name][1] += v
    # the FP64 peak micro-benchmarks run once per process, outside every timed region
    outside = {k: v for k, v in agg.items() if "peak_kernel" in k or "dmma_fed" in k}
    inside = {k: v for k, v in agg.items() if k not in outside}
    tot = sum(v[1] for v in inside.values())
    lines = [f"# {title}", "", f"source: `{path}` (ncu --metrics gpu__time_duration.sum --clock-control none; "
             "cold-cache, serialised: compare SHARES)", "", "| kernel | launches | total ms | ms per launch | share |",
             "|---|---|---|---|---|"]
    for k, v in sorted(inside.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| {k[:70]} | {v[0]} | {v[1]:.3f} | {v[1] / v[0]:.4f} | {v[1] / tot:.3f} |")
    if outside:
        lines += ["", "outside the timed region (roofline denominators, once per process): "
                  + ", ".join(f"{k.split('::')[-1]} {v[1]:.1f} ms" for k, v in outside.items())]
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "kernel launch list")
    else:
        report(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "ncu summary")
