#!/bin/bash
# quick A/B of the default build: parity of the tensor tests + probe modes (see r1_call_ab.sh)
OUT=gpurun_out; mkdir -p $OUT
L=gp-ann_b200
probe() {
  local lib=$L/libgpann_cuda${1:+_$1}.so
  echo "== lib=${1:-default} mode=$2 ${3:-}"
  env GPANN_LIB=$lib GPANN_TC_MODE=$2 ${3:-X=1} timeout 240 python bench.py --no-e2e --no-cpu --steps 5 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('ms_per_step', round(j['ms_per_step'],2), 'kernel_ms', round(j['roofline']['kernel_ms'],2), 'TOPs', round(j['roofline']['achieved'],1), j['clocks'])
    elif 'rror' in l or 'assert' in l: print(l.strip())
"
}
{
for tag in ${PARITY_TAGS:-"-"}; do
  tag=${tag#-}
  echo "== parity lib=${tag:-default}"
  GPANN_LIB=$L/libgpann_cuda${tag:+_$tag}.so timeout 600 python -m pytest tests/test_gpu_topk.py -m gpu -x -q 2>&1 | tail -4
done
for pm in ${PROBES:-":0 :1 :3"}; do probe "${pm%%:*}" "${pm##*:}"; done
} > $OUT/ab2_probes.txt 2>&1
cat $OUT/ab2_probes.txt
