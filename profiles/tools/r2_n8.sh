#!/bin/bash
# Round-2 8-GPU pass: headline at N=8 (verified, e2e, sharded k-means round), C5 (1B x 100 f32), N=4 headline, multi-GPU tests
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv,noheader > gpurun_out/r2_n8_gpus.txt
nvidia-smi topo -m >> gpurun_out/r2_n8_gpus.txt 2>&1
nproc >> gpurun_out/r2_n8_gpus.txt; free -g | head -2 >> gpurun_out/r2_n8_gpus.txt
run() {
  name=$1; n=$2; shift; shift
  env "$@" timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29555 \
    bench.py --gpus $n --steps 5 --warmup 3 $BARGS > gpurun_out/r2_bench_${name}.json 2> gpurun_out/r2_bench_${name}.err
  python - <<PY
import json
try:
    l = json.loads([x for x in open("gpurun_out/r2_bench_${name}.json").read().splitlines() if x.startswith("{")][-1])
    print("${name}", "ms/step %.2f" % l["ms_per_step"], "kernel_ms %.2f" % l["roofline"]["kernel_ms"], "frac %.3f" % l["roofline"]["frac"], "whole-step frac %.3f" % l["roofline"]["whole_step_frac"],
          "verified", (l.get("verified") or {}).get("ok"), (l.get("verified") or {}).get("result_fingerprint"), "e2e", (l.get("e2e") or {}).get("ms_per_step"), (l.get("e2e") or {}).get("numa_binding"),
          "kmeans", (l.get("kmeans_sharded") or {}).get("ms_per_round"))
except Exception as e:
    print("${name} FAILED", e); print(open("gpurun_out/r2_bench_${name}.err").read()[-2500:])
PY
}
BARGS="--kmeans-multi" run n8 8 X=1
BARGS="--no-e2e --no-verify" run n8_nopeers 8 GPANN_NO_PEER_BOUNDS=1
BARGS="--workload c5" run n8_c5 8 X=1
BARGS="--kmeans-multi --no-e2e" run n4 4 X=1
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=1000 > gpurun_out/r2_n8_tests.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r2_n8_tests.txt; tail -4 gpurun_out/r2_n8_tests.txt
