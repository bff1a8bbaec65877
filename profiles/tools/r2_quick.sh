#!/bin/bash
# quick look at the current build: 100M headline (verified), 12.5M-row shard, the shard's per-pass counters
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for n in 100000000 12500000; do
  extra="--no-verify"; [ $n = 100000000 ] && extra=""
  GPANN_BENCH_N=$n timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-configs $extra > gpurun_out/quick_$n.json 2> gpurun_out/quick_$n.err
  python - <<PY
import json
try:
    l = json.loads([x for x in open("gpurun_out/quick_$n.json").read().splitlines() if x.startswith("{")][-1])
    print($n, "ms/step %.2f" % l["ms_per_step"], "kernel_ms %.2f" % l["roofline"]["kernel_ms"], "frac %.3f" % l["roofline"]["frac"],
          "verified", (l.get("verified") or {}).get("ok"), "fp", (l.get("verified") or {}).get("result_fingerprint"), "clk", l["clocks"]["sm_mhz"])
except Exception as e:
    print($n, "FAILED", e); print(open("gpurun_out/quick_$n.err").read()[-1500:])
PY
done
GPANN_TC_STATS=2 GPANN_BENCH_N=12500000 timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-configs --no-verify 2>&1 | grep "gpann stats" | tail -5 | cut -c1-400
