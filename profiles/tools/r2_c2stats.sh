#!/bin/bash
# C1 / C2 (float top-k) with the kernel's wait counters and per-pass record counts (GPANN_TC_STATS=2)
cd "$(dirname "$0")/../.."
GPANN_TC_STATS=2 python - <<'PY' 2>&1 | grep -v "^$" | cut -c1-420
import sys
sys.path.insert(0, "."); sys.path.insert(0, "gp-ann_b200")
import torch, gpann, bench
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
ctx = gpann.Context(0)
pk, src = bench.peaks()
for name, n, d, k in (("C1", 1_000_000, 128, 10), ("C2", 10_000_000, 96, 100)):
    print("=====", name, flush=True)
    r = bench.config_float_topk(torch, gpann, ctx, dev, name, n, d, 10_000, k, pk, src, steps=1, warmup=1)
    print(name, "ms/step %.2f kernel_ms %.2f frac %.3f" % (r["ms_per_step"], r["roofline"]["kernel_ms"], r["roofline"]["frac"]), flush=True)
PY
