import os, sys, subprocess
code = r'''
import sys, numpy as np
sys.path.insert(0, "gp-ann_b200"); sys.path.insert(0, ".")
import gpann
from oracle.oracle import Oracle
g = np.load("tests/golden/balanced_kmeans.npz")
o = Oracle()
with gpann.Context(0) as c:
    pts = c.upload(g["points"], storage=gpann.STORAGE_F32)
    a, cen, s = c.balanced_kmeans(pts, g["init"], int(g["max_cluster_size"]))
    print("seed", __import__("os").environ.get("GPANN_BKM_SEED"), "objective", o.objective(g["points"], cen, a), "sizes", s.tolist())
'''
for seed in range(8):
    env = dict(os.environ, GPANN_BKM_SEED=str(seed))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-500:])
