"""Measures the DRAM traffic of one config 2 step's demodulator launches under ncu and writes
profiles/r02_demod_traffic.json (read by bench.py, which ignores it once the kernel sources change).
usage (GPU box): python scripts/ncu_traffic.py   -> runs ncu itself"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

out_csv = os.path.join(ROOT, "gpurun_out", "r02_demod_traffic.csv")
os.makedirs(os.path.dirname(out_csv), exist_ok=True)
cmd = ["ncu", "--metrics", "gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none",
       "-k", "regex:fsk_demod|fast_", "--csv", "--log-file", out_csv,
       sys.executable, os.path.join(ROOT, "scripts", "run_config2_once.py"), "--flags", "0", "--iters", "1"]
subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
rows = [r for r in csv.reader(open(out_csv)) if r and not r[0].startswith("==")]
h = rows[0]
ki, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
tot, launches, per = 0.0, set(), {}
for r in rows[1:]:
    if len(r) <= vi or "dram__bytes" not in r[mi]:
        continue
    v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    tot += v
    launches.add(r[0])
    k = r[ki].split("(")[0]
    per[k] = per.get(k, 0.0) + v
d = {"source_hash": bench.kernel_source_hash(), "bytes_per_step": tot, "launches": len(launches),
     "by_kernel_bytes": per, "algorithmic_bytes_per_step": 65536 * 48000 * 4,
     "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum of one config 2 call (all kernels of the call)"}
json.dump(d, open(os.path.join(ROOT, "profiles", "r02_demod_traffic.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in d.items() if k != "by_kernel_bytes"}), {k: round(v / 1e9, 2) for k, v in per.items()})
