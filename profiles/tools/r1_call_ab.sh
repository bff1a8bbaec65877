#!/bin/bash
# One-box A/B of the tcgen05 top-k kernel's pipeline variants (see params.h GPANN_ACC_N / GPANN_SLOTS / GPANN_A_TMEM).
# Usage (under gpurun): bash profiles/tools/r1_call_ab.sh
OUT=gpurun_out; mkdir -p $OUT
L=gp-ann_b200
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/ab_gpu.txt 2>&1
probe() {  # lib-tag mode
  local lib=$L/libgpann_cuda${1:+_$1}.so
  echo "== lib=${1:-default} mode=$2"
  GPANN_LIB=$lib GPANN_TC_MODE=$2 timeout 240 python bench.py --no-e2e --no-cpu --steps 5 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('ms_per_step', round(j['ms_per_step'],2), 'kernel_ms', round(j['roofline']['kernel_ms'],2), 'TOPs', round(j['roofline']['achieved'],1), j['clocks'])
    elif 'rror' in l or 'assert' in l: print(l.strip())
"
}
{
echo "#### parity of the variants (tensor tests only)"
for tag in ts n128s4; do
  echo "== parity lib=$tag"
  GPANN_LIB=$L/libgpann_cuda_$tag.so timeout 600 python -m pytest tests/test_gpu_topk.py -m gpu -x -q -k "golden_bytes or tensor_bytes or heavy_ties or tensor_float_vs or large_tensor" 2>&1 | tail -4
done
echo "#### probes"
probe base 0
probe "" 0
probe ts 0
probe n128s4 0
probe ts 2
probe ts 5
probe ts 1
probe "" 2
probe "" 5
probe "" 1
probe "" 3
probe n128s4 2
probe n128s4 5
} > $OUT/ab_probes.txt 2>&1
echo "#### full GPU suite (default lib)" > $OUT/ab_tests.txt
timeout 900 python -m pytest tests -m gpu -x -q >> $OUT/ab_tests.txt 2>&1
tail -3 $OUT/ab_tests.txt
cat $OUT/ab_probes.txt
