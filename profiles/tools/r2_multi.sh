#!/bin/bash
# Round-2 multi-GPU pass (N = $1 GPUs, default 2): multi-GPU tests, then the torchrun bench at N (headline, verified,
# sharded k-means round) and the 12.5M-rows-per-rank shape (what each of 8 GPUs scans) with different sample-level plans.
cd "$(dirname "$0")/../.."
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2_multi_gpus.txt
nvidia-smi topo -m >> gpurun_out/r2_multi_gpus.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_async_and_run.py -m gpu -q --timeout=1200 > gpurun_out/r2_multi_tests.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r2_multi_tests.txt; tail -15 gpurun_out/r2_multi_tests.txt
run() {  # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
    bench.py --gpus $N --steps 5 --warmup 3 $BARGS > gpurun_out/r2_bench_${name}.json 2> gpurun_out/r2_bench_${name}.err
  python - <<PY
import json
try:
    l = json.loads([x for x in open("gpurun_out/r2_bench_${name}.json").read().splitlines() if x.startswith("{")][-1])
    print("${name}", "ms/step %.2f" % l["ms_per_step"], "kernel_ms %.2f" % l["roofline"]["kernel_ms"], "whole-step frac %.3f" % l["roofline"]["whole_step_frac"],
          "verified", (l.get("verified") or {}).get("ok"), "e2e ms", (l.get("e2e") or {}).get("ms_per_step"), "kmeans", (l.get("kmeans_sharded") or {}).get("ms_per_round"))
except Exception as e:
    print("${name} FAILED", e); print(open("gpurun_out/r2_bench_${name}.err").read()[-2500:])
PY
}
BARGS="--kmeans-multi" run n${N} X=1
BARGS="--no-e2e --no-verify" run n${N}_shard12m GPANN_BENCH_N=$((12500000*N))
BARGS="--no-e2e --no-verify" run n${N}_shard12m_2lv GPANN_BENCH_N=$((12500000*N)) GPANN_TC_LEVELS=1/12/0
BARGS="--no-e2e --no-verify" run n${N}_shard12m_lv3small GPANN_BENCH_N=$((12500000*N)) GPANN_TC_LEVELS=1/8/48
