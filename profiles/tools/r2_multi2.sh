#!/bin/bash
# multi-GPU pass 2 (N = $1): multi-GPU tests + the 12.5M-rows-per-rank step with and without peer-shared bounds
cd "$(dirname "$0")/../.."
N=${1:-2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=1200 > gpurun_out/r2_multi_tests.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r2_multi_tests.txt; tail -6 gpurun_out/r2_multi_tests.txt
run() {
  name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
    bench.py --gpus $N --steps 5 --warmup 3 $BARGS > gpurun_out/r2_bench_${name}.json 2> gpurun_out/r2_bench_${name}.err
  python - <<PY
import json
try:
    l = json.loads([x for x in open("gpurun_out/r2_bench_${name}.json").read().splitlines() if x.startswith("{")][-1])
    print("${name}", "ms/step %.2f" % l["ms_per_step"], "kernel_ms %.2f" % l["roofline"]["kernel_ms"], "whole-step frac %.3f" % l["roofline"]["whole_step_frac"],
          "verified", (l.get("verified") or {}).get("ok"), (l.get("verified") or {}).get("result_fingerprint"))
except Exception as e:
    print("${name} FAILED", e); print(open("gpurun_out/r2_bench_${name}.err").read()[-2500:])
PY
}
BARGS="--no-e2e --no-verify" run n${N}_shard12m_peers GPANN_BENCH_N=$((12500000*N))
BARGS="--no-e2e --no-verify" run n${N}_shard12m_nopeers GPANN_BENCH_N=$((12500000*N)) GPANN_NO_PEER_BOUNDS=1
BARGS="--no-e2e" run n${N}_peers X=1
