#!/bin/bash
# four sample levels against three at 12.5M rows (GPANN_TC_LEVELS=a/b/c/d)
cd "$(dirname "$0")/../.."
for lv in default 1/4/18/72 1/4/16/64 1/5/25/100 0/4/18/72 default; do
  envs="GPANN_X=1"; [ $lv != default ] && envs="GPANN_TC_LEVELS=$lv"
  env $envs GPANN_BENCH_N=12500000 timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-configs --no-verify 2>/dev/null | python -c "
import json,sys
l=[x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1]; l=json.loads(l)
print('$lv', 'ms/step %.2f' % l['ms_per_step'], 'kernel_ms %.2f' % l['roofline']['kernel_ms'])"
done
