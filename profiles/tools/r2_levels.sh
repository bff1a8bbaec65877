#!/bin/bash
# Sample-level sweep (GPANN_TC_LEVELS = tiles per split of the three pre-passes) at a given size
cd "$(dirname "$0")/../.."
N=${1:-12500000}
for lv in default 1/6/72 1/6/0 1/12/0 1/24/0 1/36/0 2/24/0 1/8/48 0/8/64; do
  envs="GPANN_X=1"; [ $lv != default ] && envs="GPANN_TC_LEVELS=$lv"
  env $envs GPANN_BENCH_N=$N timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-configs --no-verify 2>/dev/null | python -c "
import json,sys
l=[x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1]; l=json.loads(l)
print('$lv', 'ms/step %.2f' % l['ms_per_step'], 'kernel_ms %.2f' % l['roofline']['kernel_ms'])"
done
