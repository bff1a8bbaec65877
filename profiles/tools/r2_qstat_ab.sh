#!/bin/bash
# order-statistic bound + interleaved splits (k > 32) on/off: C2, plus the k = 100 parity tests
cd "$(dirname "$0")/../.."
for mode in qstat noqstat; do
  envs="GPANN_X=1"; [ $mode = noqstat ] && envs="GPANN_TC_NOQSTAT=1"
  echo "== $mode"; env $envs bash profiles/tools/r2_c12.sh
done
GPANN_TC_STATS=2 bash profiles/tools/r2_c2stats.sh 2>&1 | grep -A40 "===== C2" | grep "main pass\|sample pass\|C2 ms" | head -8 | cut -c1-260
