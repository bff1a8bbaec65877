#!/bin/bash
# C1 / C2 timing only (float top-k through the bf16 filter + exact re-rank); optional arg = tree root to run in
cd "$(dirname "$0")/../.."
[ -n "$1" ] && cd "$1"
python - <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "gp-ann_b200")
import torch, gpann, bench
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
ctx = gpann.Context(0)
pk, src = bench.peaks()
for name, n, d, k in (("C1", 1_000_000, 128, 10), ("C2", 10_000_000, 96, 100)):
    r = bench.config_float_topk(torch, gpann, ctx, dev, name, n, d, 10_000, k, pk, src)
    print(name, "ms/step %.2f kernel_ms %.2f frac %.3f" % (r["ms_per_step"], r["roofline"]["kernel_ms"], r["roofline"]["frac"]))
PY
