#!/bin/bash
# A/B of kernel build variants on one box: $1.. = build tags ("default" = libgpann_cuda.so).  Per variant: the 100M headline
# (device-resident, verified against the reference) and the 12.5M-row shard (what one of 8 GPUs scans).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for tag in "$@"; do
  lib=gp-ann_b200/libgpann_cuda.so; [ "$tag" != default ] && lib=gp-ann_b200/libgpann_cuda_$tag.so
  for n in 100000000 12500000; do
    extra="--no-verify"; [ $n = 100000000 ] && extra=""
    GPANN_LIB=$PWD/$lib GPANN_BENCH_N=$n timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-configs $extra \
      > gpurun_out/ab_${tag}_$n.json 2> gpurun_out/ab_${tag}_$n.err
    python - <<PY
import json
try:
    l = json.loads([x for x in open("gpurun_out/ab_${tag}_$n.json").read().splitlines() if x.startswith("{")][-1])
    print("$tag", $n, "ms/step %.2f" % l["ms_per_step"], "kernel_ms %.2f" % l["roofline"]["kernel_ms"], "frac %.3f" % l["roofline"]["frac"],
          "verified", (l.get("verified") or {}).get("ok"), "clk", l["clocks"]["sm_mhz"])
except Exception as e:
    print("$tag", $n, "FAILED", e); print(open("gpurun_out/ab_${tag}_$n.err").read()[-1500:])
PY
  done
done
