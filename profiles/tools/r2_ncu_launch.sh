#!/bin/bash
# ncu --set full of ONE launch of the tcgen05 top-k kernel on the 12.5M-row shard: 4 launches per step (3 sample passes + main),
# $1 = launch index to capture (14 = third sample pass of step 4, 15 = main pass), $2 = report name
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
SKIP=${1:-15}; OUT=${2:-r2_ncu_launch}
GPANN_BENCH_N=12500000 timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:topk_tc_kernel -s $SKIP -c 1 -o gpurun_out/$OUT -f \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-configs --no-verify > gpurun_out/$OUT.log 2>&1
tail -2 gpurun_out/$OUT.log | cut -c1-200
ls -la gpurun_out/$OUT.ncu-rep
