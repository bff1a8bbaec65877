#!/bin/bash
# ncu --set full of ONE kmeans_tc_kernel launch (20M x 96 -> 1024) for the source-level stall picture of its epilogue
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
cat > /tmp/km_one.py <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "gp-ann_b200")
import torch, gpann, bench
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
ctx = gpann.Context(0)
pk, src = bench.peaks()
r = bench.config_kmeans(torch, gpann, ctx, dev, 20_000_000, 96, 1024, pk, src, rounds=2, cpu=False)
print(r["ms_per_round"], r["roofline"]["kernel_ms"])
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:kmeans_tc_kernel -s 2 -c 1 -f -o gpurun_out/r2_km_tc python /tmp/km_one.py > gpurun_out/r2_km_ncu.log 2>&1
tail -2 gpurun_out/r2_km_ncu.log
