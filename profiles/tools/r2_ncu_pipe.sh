#!/bin/bash
# ncu --set full of ONE main-pass launch of the tcgen05 top-k kernel on the 12.5M-row shard (4 launches per step: 3 sample passes + main)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
LIB=${1:-libgpann_cuda.so}; OUT=${2:-r2_ncu_pipe}
GPANN_LIB=$PWD/gp-ann_b200/$LIB GPANN_BENCH_N=12500000 timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:topk_tc_kernel -s 15 -c 1 -o gpurun_out/$OUT -f \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-configs --no-verify > gpurun_out/$OUT.log 2>&1
tail -3 gpurun_out/$OUT.log | cut -c1-300
ls -la gpurun_out/$OUT.ncu-rep
