#!/bin/bash
# 8-GPU pass after the find log: headline at N=8 (verified, e2e, sharded k-means round), multi-GPU tests
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() {
  name=$1; n=$2; shift; shift
  env "$@" timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29555 \
    bench.py --gpus $n --steps 5 --warmup 3 $BARGS > gpurun_out/r2_bench_${name}.json 2> gpurun_out/r2_bench_${name}.err
  python - <<PY
import json
try:
    l = json.loads([x for x in open("gpurun_out/r2_bench_${name}.json").read().splitlines() if x.startswith("{")][-1])
    print("${name}", "ms/step %.2f" % l["ms_per_step"], "kernel_ms %.2f" % l["roofline"]["kernel_ms"], "frac %.3f" % l["roofline"]["frac"], "whole-step frac %.3f" % l["roofline"]["whole_step_frac"],
          "verified", (l.get("verified") or {}).get("ok"), (l.get("verified") or {}).get("result_fingerprint"), "e2e", (l.get("e2e") or {}).get("ms_per_step"),
          "kmeans", (l.get("kmeans_sharded") or {}).get("ms_per_round"))
except Exception as e:
    print("${name} FAILED", e); print(open("gpurun_out/r2_bench_${name}.err").read()[-2500:])
PY
}
BARGS="--kmeans-multi" run n8b 8 X=1
BARGS="--no-e2e --no-verify" run n8b_nolog 8 GPANN_TC_NOLOG=1
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=1000 > gpurun_out/r2_n8b_tests.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r2_n8b_tests.txt; tail -4 gpurun_out/r2_n8b_tests.txt
