"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing in gpann.dist: row sharding with global ids, the
all-gather of per-shard keys and the merge give exactly the single-scan result; the k-means partial all-reduce
reproduces the unsharded sums.  The per-shard compute is played by the oracle here (tests may use it); on the
GPU box the same ShardedTopK drives gpann_topk_device / gpann_merge_keys_device (tests/test_gpu_topk.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def merge_keys_numpy(gathered, k):
    g = np.ascontiguousarray(gathered).view(np.uint64)
    parts, nq, kk = g.shape
    flat = np.transpose(g, (1, 0, 2)).reshape(nq, parts * kk)
    return np.sort(flat, axis=1)[:, :k]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "gp-ann_b200"))
    import gpann
    from gpann.dist import ShardedTopK, allreduce_kmeans_partials, shard_bounds
    from oracle.oracle import Oracle
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = Oracle()
    rng = np.random.default_rng(11)  # same data on every rank
    n, d, nq, k = 1001, 32, 9, 10
    P = rng.integers(0, 4, (n, d)).astype(np.float32)  # tiny alphabet: many ties across the shard boundary
    Q = rng.integers(0, 4, (nq, d)).astype(np.float32)
    lo, hi = shard_bounds(n, rank, world)

    def local_topk(queries, kk):
        dd, ii, cc = o.ground_truth(P[lo:hi], queries.numpy(), kk)
        keys = np.full((nq, kk), gpann.EMPTY_KEY, np.uint64)
        for q in range(nq):
            for r in range(cc[q]):
                keys[q, r] = gpann.pack_key(dd[q, r], ii[q, r] + lo)  # global ids
        return torch.from_numpy(keys.view(np.int64))

    def merge(gathered, kk):
        return merge_keys_numpy(gathered.numpy(), kk)

    merged = ShardedTopK(local_topk, merge)(torch.from_numpy(Q), k)
    md, mi = gpann.unpack_keys(merged)
    gd, gi, _ = o.ground_truth(P, Q, k)
    ok_topk = bool(np.array_equal(mi, gi) and np.array_equal(md, gd))

    # k-means partials
    kc = 5
    C = P[:kc].copy()
    assign, _ = o.nearest_centers(P, C)
    sums = np.zeros((kc, d), np.float32)
    counts = np.zeros(kc, np.int64)
    for i in range(lo, hi):
        sums[assign[i]] += P[i]
        counts[assign[i]] += 1
    ts, tc = torch.from_numpy(sums), torch.from_numpy(counts)
    allreduce_kmeans_partials(ts, tc)
    full, sizes = o.aggregate(P, assign, kc, normalize=False)
    ok_km = bool(np.array_equal(tc.numpy(), sizes.astype(np.int64)) and np.allclose(ts.numpy(), full, rtol=1e-6))
    open(os.path.join(out_dir, f"rank{rank}.txt"), "w").write(f"{ok_topk} {ok_km} {lo} {hi}")
    dist.destroy_process_group()


def test_sharded_topk_and_kmeans_world2(tmp_path, oracle):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        ok_topk, ok_km, lo, hi = open(tmp_path / f"rank{r}.txt").read().split()
        assert ok_topk == "True" and ok_km == "True"
    assert open(tmp_path / "rank0.txt").read().split()[2:] == ["0", "501"]
    assert open(tmp_path / "rank1.txt").read().split()[2:] == ["501", "1001"]


def test_shard_bounds_cover_and_are_disjoint():
    sys.path.insert(0, os.path.join(ROOT, "gp-ann_b200"))
    from gpann.dist import shard_bounds
    for n in (0, 1, 7, 8, 1001, 100_000_000):
        for world in (1, 2, 4, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert all(lo <= hi for lo, hi in b)
