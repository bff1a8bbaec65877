#!/bin/bash
# last 8-GPU check of the round: headline only (verified), no e2e leg
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 \
  bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e > gpurun_out/r2_last_bench_n8.json 2> gpurun_out/r2_last_bench_n8.err
python - <<PY
import json
try:
    l = json.loads([x for x in open("gpurun_out/r2_last_bench_n8.json").read().splitlines() if x.startswith("{")][-1])
    print("n8", "ms/step %.2f" % l["ms_per_step"], "kernel_ms %.2f" % l["roofline"]["kernel_ms"], "frac %.3f" % l["roofline"]["frac"], "whole-step frac %.3f" % l["roofline"]["whole_step_frac"],
          "verified", (l.get("verified") or {}).get("ok"), (l.get("verified") or {}).get("result_fingerprint"), l["clocks"])
except Exception as e:
    print("n8 FAILED", e); print(open("gpurun_out/r2_last_bench_n8.err").read()[-2500:])
PY
