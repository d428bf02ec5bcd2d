#!/bin/bash
# One GPU pass over the fast demodulator: doubt-band soundness (tap vs oracle), full config 2 against the float64
# kernel with timings, and the instruction / pipe counters of one slab launch.  usage: scripts/gpu_fast_check.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
timeout 300 python scripts/exp_fast_tap.py --streams 512 --out gpurun_out/fast_${tag}_tap.jsonl > gpurun_out/fast_${tag}_tap.log 2>&1
tail -1 gpurun_out/fast_${tag}_tap.log | cut -c1-400
timeout 500 python scripts/exp_fast_vs_exact.py --skip-a --iters 4 --out gpurun_out/fast_${tag}_vs_exact.jsonl > gpurun_out/fast_${tag}.log 2>&1
tail -2 gpurun_out/fast_${tag}.log | cut -c1-420
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:fsk_demod_fast\|fsk_demod_duo -s 24 -c 24 --csv --log-file gpurun_out/fast_${tag}_ncu.csv python scripts/run_config2_once.py --flags 512 --iters 2 > gpurun_out/ncu_run.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/fast_${tag}_ncu.csv')))
hdr=None; vals={}
for r in rows:
    if len(r)>10 and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); vals.setdefault(d['ID'],{})[d['Metric Name']]=float(d['Metric Value'].replace(',',''))
n=len(vals)
if n:
    keys=list(next(iter(vals.values())).keys())
    tot={k:sum(v[k] for v in vals.values()) for k in keys}
    print('launches',n,'sum us',round(tot['gpu__time_duration.sum']/1e3,1),'instr/sample',round(tot['smsp__inst_executed.sum']/(65536/32*48000),2))
    for k in keys:
        if 'pct' in k: print(' ',k.split('.')[0],round(tot[k]/n,1))
PY
