#!/bin/bash
# order-statistic bound for the small-k lists in the sample passes, on/off: 100M, 12.5M shard, C1
cd "$(dirname "$0")/../.."
for mode in off on; do
  envs="GPANN_X=1"; [ $mode = on ] && envs="GPANN_TC_QSTAT_SMALL=1"
  echo "== qstat_small $mode"
  for n in 100000000 12500000; do
    env $envs GPANN_BENCH_N=$n timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-configs --no-verify 2>/dev/null | python -c "
import json,sys
l=[x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1]; l=json.loads(l)
print($n, 'ms/step %.2f' % l['ms_per_step'], 'kernel_ms %.2f' % l['roofline']['kernel_ms'])"
  done
  env $envs bash profiles/tools/r2_c12.sh | head -1
  env $envs GPANN_TC_STATS=2 GPANN_BENCH_N=12500000 timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-configs --no-verify 2>&1 | grep "sample pass" | tail -3 | cut -c1-130
done
