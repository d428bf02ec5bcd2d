#!/bin/bash
# Time bench.py's device-resident step (config 2) for the default library and every A/B variant under build/ab/.
cd "$(dirname "$0")/.."
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-60s %.3f ms/step  kernel %.3f ms' % ('$1', d['ms_per_step'], d['roofline']['kernel_ms']))"; }
unset WAM_LIB; run default
for f in build/ab/libwam_*.so; do
  case "$f" in *timing*) continue;; esac
  [ -e "$f" ] || continue
  export WAM_LIB=$PWD/$f; run "$(basename $f)"
done
