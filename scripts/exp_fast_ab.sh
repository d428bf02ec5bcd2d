#!/bin/bash
# Time the fast kernel (config 2, device-resident) for the default library and every A/B variant under build/ab/.
cd "$(dirname "$0")/.."
run() { timeout 400 python scripts/exp_fast_vs_exact.py --skip-a --iters 4 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('%-44s %-10s %8.3f ms (exact %.3f) differs=%d flagged=%d doubt=%d' % ('$1', d['kernel'], d['ms_min'], d['exact_ms_min'], d['differs_from_exact'], d['fast_stats']['flagged_last_call'], d['fast_stats']['doubtful_samples']))"; }
unset WAM_LIB; run default
for f in build/ab/libwam_*.so; do
  [ -e "$f" ] || continue
  export WAM_LIB=$PWD/$f; run "$(basename $f)"
done
