#!/bin/bash
# C3 timing (one Lloyd round at 100M x 96 -> 1024) + the k-means GPU tests
cd "$(dirname "$0")/../.."
python - <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "gp-ann_b200")
import torch, gpann, bench
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
ctx = gpann.Context(0)
pk, src = bench.peaks()
r = bench.config_kmeans(torch, gpann, ctx, dev, 100_000_000, 96, 1024, pk, src, rounds=3, cpu=False)
print("C3 ms/round %.2f kernel_ms %.2f frac %.3f reranked %d" % (r["ms_per_round"], r["roofline"]["kernel_ms"], r["roofline"]["frac"], r["exact_rerank_points"]))
PY
timeout 600 python -m pytest tests/test_gpu_kmeans.py tests/test_gpu_async_and_run.py -m gpu -q 2>&1 | tail -3
