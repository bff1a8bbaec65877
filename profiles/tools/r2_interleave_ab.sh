#!/bin/bash
# contiguous vs interleaved point splits: headline 100M + 12.5M shard, C1, C2
cd "$(dirname "$0")/../.."
for mode in default interleave contiguous; do
  envs="GPANN_X=1"; [ $mode = interleave ] && envs="GPANN_TC_INTERLEAVE=1"; [ $mode = contiguous ] && envs="GPANN_TC_INTERLEAVE=0"
  echo "== $mode"
  for n in 100000000 12500000; do
    env $envs GPANN_BENCH_N=$n timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-configs --no-verify 2>/dev/null | python -c "
import json,sys
l=[x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1]; l=json.loads(l)
print($n, 'ms/step %.2f' % l['ms_per_step'], 'kernel_ms %.2f' % l['roofline']['kernel_ms'])"
  done
  env $envs bash profiles/tools/r2_c12.sh
done
