#!/bin/bash
# Probe modes of the top-k tensor kernel at the headline size (GPANN_TC_MODE; results of probe runs are not distances):
#   0+stats = production loop with wait counters, 5 = free-running MMA + TMA (no slot hand-shake, no drain), 2 = MMA + hand-shake,
#   3 = drain + max only (no MMA), 1 = full pipeline without selection
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-configs --no-verify > gpurun_out/probe_$name.json 2> gpurun_out/probe_$name.err
  python - <<PY
import json
try:
    l = json.loads([x for x in open("gpurun_out/probe_$name.json").read().splitlines() if x.startswith("{")][-1])
    print("$name", "ms/step %.2f" % l["ms_per_step"], "kernel_ms %.2f" % l["roofline"]["kernel_ms"], "clk", l["clocks"]["sm_mhz"])
except Exception as e:
    print("$name", "FAILED", e)
PY
  grep "gpann stats" gpurun_out/probe_$name.err | tail -2
}
run prod GPANN_X=0
run stats GPANN_TC_STATS=1
run mode5 GPANN_TC_MODE=5
run mode2 GPANN_TC_MODE=2
run mode3 GPANN_TC_MODE=3
run mode1 GPANN_TC_MODE=1
