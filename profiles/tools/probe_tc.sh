# A/B probes of the tcgen05 top-k kernel on ONE box (bench.py, resident value only): GPANN_TC_MODE 1/2/3 are invalid-result probes
for m in ${PROBE_MODES:-"GPANN_TC_MODE=0 GPANN_TC_MODE=1 GPANN_TC_MODE=2 GPANN_TC_MODE=3 GPANN_TC_WARM=1 GPANN_TC_MODE=0"}; do
  echo "== $m"; env $m timeout 300 python bench.py --no-e2e --no-cpu --steps 5 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print(j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['achieved'], j['clocks'])
    elif 'rror' in l: print(l.strip())
"
done
