#!/bin/bash
# Final multi-GPU pass for N = $1 GPUs: headline (verified; e2e + sharded k-means round), and at N = 8 the C5 evidence workload
cd "$(dirname "$0")/../.."
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/r2_final_n${N}_gpus.txt
run() {
  name=$1; shift
  env "$@" timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
    bench.py --gpus $N --steps 5 --warmup 3 $BARGS > gpurun_out/r2_final_bench_${name}.json 2> gpurun_out/r2_final_bench_${name}.err
  python - <<PY
import json
try:
    l = json.loads([x for x in open("gpurun_out/r2_final_bench_${name}.json").read().splitlines() if x.startswith("{")][-1])
    print("${name}", "ms/step %.2f" % l["ms_per_step"], "kernel_ms %.2f" % l["roofline"]["kernel_ms"], "frac %.3f" % l["roofline"]["frac"], "whole-step frac %.3f" % l["roofline"]["whole_step_frac"],
          "verified", (l.get("verified") or {}).get("ok"), (l.get("verified") or {}).get("result_fingerprint"), "e2e", (l.get("e2e") or {}).get("ms_per_step"),
          "kmeans", (l.get("kmeans_sharded") or {}).get("ms_per_round"))
except Exception as e:
    print("${name} FAILED", e); print(open("gpurun_out/r2_final_bench_${name}.err").read()[-2500:])
PY
}
BARGS="--kmeans-multi" run n${N} X=1
if [ $N = 8 ]; then
  BARGS="--workload c5" run n8_c5 X=1
fi
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=800 > gpurun_out/r2_final_n${N}_tests.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r2_final_n${N}_tests.txt; tail -3 gpurun_out/r2_final_n${N}_tests.txt
