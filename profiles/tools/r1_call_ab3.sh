#!/bin/bash
# probes with env knobs: each entry "ENV1=..,ENV2=..:mode"
OUT=gpurun_out; mkdir -p $OUT
{
[ -n "$SKIP_MMA" ] || timeout 120 profiles/tools/_bin/mma_issue_bench | tail -8
for e in ${PROBES3}; do
  envs=${e%%:*}; mode=${e##*:}
  echo "== $envs mode=$mode"
  env ${envs//,/ } GPANN_TC_MODE=$mode timeout 240 python bench.py --no-e2e --no-cpu --steps 5 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('ms_per_step', round(j['ms_per_step'],2), 'kernel_ms', round(j['roofline']['kernel_ms'],2), 'TOPs', round(j['roofline']['achieved'],1), j['clocks'])
    elif 'rror' in l or 'assert' in l: print(l.strip())
    elif 'clk per item' in l: last=l.strip()
try: print(last)
except NameError: pass
"
done
} > $OUT/ab3_probes.txt 2>&1
cat $OUT/ab3_probes.txt
