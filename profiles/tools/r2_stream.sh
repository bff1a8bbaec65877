#!/bin/bash
# streaming-kernel numbers only (bench.py config_streaming) for quick A/B of stream.cu changes
cd "$(dirname "$0")/../.."
python - <<'PY'
import json, sys
sys.path.insert(0, "."); sys.path.insert(0, "gp-ann_b200")
import torch, gpann, bench
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
ctx = gpann.Context(0)
pk, src = bench.peaks()
r = bench.config_streaming(torch, gpann, ctx, dev, pk, src)
for c in r["cases"]:
    print(c["storage"], "nq", c["nq"], "ms %.2f" % c["ms"], "GB/s %.0f" % c["hbm_gbs"], "frac %.3f" % c["frac_of_hbm_peak"])
PY
