#!/usr/bin/env python
"""Turn ncu artefacts from gpurun_out/ into the small text/JSON summaries committed under profiles/.
usage: python profiles/export_ncu.py <tag> <full.ncu-rep> [launches.csv]"""
import csv, json, os, subprocess, sys

tag, rep = sys.argv[1], sys.argv[2]
launches = sys.argv[3] if len(sys.argv) > 3 else None
here = os.path.dirname(os.path.abspath(__file__))

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__icc_request_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum"]
out = []
summary = []
for r in rows[2:]:
    rec = {h: v for h, v in zip(hdr, r)}
    out.append(f"== {rec.get('Kernel Name', '?')}")
    keep = {}
    for k in KEEP[1:]:
        if k in rec:
            u = units[hdr.index(k)]
            out.append(f"  {k:90s} {rec[k]:>18s} {u}")
            keep[k] = [rec[k], u]
    keep["Kernel Name"] = rec.get("Kernel Name", "?")
    summary.append(keep)
open(os.path.join(here, f"{tag}_ncu_full_summary.txt"), "w").write("\n".join(out) + "\n")
json.dump(summary, open(os.path.join(here, f"{tag}_ncu_full_summary.json"), "w"), indent=1)
print(f"wrote {tag}_ncu_full_summary.txt/.json ({len(summary)} kernels)")

if launches:
    lr = [r for r in csv.reader(open(launches)) if r and not r[0].startswith("==")]
    h = lr[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = {}
    for r in lr[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(r[ui], 1e-6)
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    lines = [f"# ncu launch list ({os.path.basename(launches)}): per-kernel launches, total ms, share of all kernel time in the command",
             f"# NOTE: times under ncu are cold-cache and serialised; compare SHARES, not absolutes."]
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{100*ms/tot:6.2f}%  {ms:10.3f} ms  n={n:4d}  {k[:110]}")
    open(os.path.join(here, f"{tag}_launches_summary.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:14]))
