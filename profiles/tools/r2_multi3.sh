#!/bin/bash
# multi-GPU pass 3 (N = $1): C5 at reduced size (smoke), headline with e2e (NUMA binding + smaller chunks), all GPU tests
cd "$(dirname "$0")/../.."
N=${1:-2}
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
    bench.py --gpus $N --steps 5 --warmup 3 $BARGS > gpurun_out/r2_bench_${name}.json 2> gpurun_out/r2_bench_${name}.err
  python - <<PY
import json
try:
    l = json.loads([x for x in open("gpurun_out/r2_bench_${name}.json").read().splitlines() if x.startswith("{")][-1])
    print("${name}", "ms/step %.2f" % l["ms_per_step"], "kernel_ms %.2f" % l["roofline"]["kernel_ms"], "frac %.3f" % l["roofline"]["frac"], "whole-step frac %.3f" % l["roofline"]["whole_step_frac"],
          "verified", (l.get("verified") or {}).get("ok"), "e2e", (l.get("e2e") or {}).get("ms_per_step"), (l.get("e2e") or {}).get("numa_binding"))
except Exception as e:
    print("${name} FAILED", e); print(open("gpurun_out/r2_bench_${name}.err").read()[-2500:])
PY
}
BARGS="--workload c5" run n${N}_c5_small GPANN_C5_N=${C5N:-40000000}
BARGS="" run n${N}_e2e X=1
timeout 1500 python -m pytest tests -m gpu -q --timeout=1200 > gpurun_out/r2_multi_tests.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r2_multi_tests.txt; tail -8 gpurun_out/r2_multi_tests.txt
