"""GPU parity tests of the segmented exact scans (gp-ann_b200/csrc/seg.cu): the batched forms of InvertedIndex::Query
(src/inverted_index.h:43-75), ApproximateKNNGraphBuilder::CrunchBucket (src/knn_graph.h:165-185) and ClosestLeaders for
every point (src/knn_graph.h:75-81), against fixtures generated from the unmodified reference."""
import numpy as np
import pytest

from conftest import golden
from helpers import assert_topk_close, assert_topk_exact

pytestmark = pytest.mark.gpu


def test_ivf_batch_golden(ctx, gpann_mod):
    g = golden("ivf_batch")
    perm = g["permutation"]
    # (a) original rows + permutation: positions break ties, ids are remapped inside the library
    pts = ctx.upload(g["points"], storage=gpann_mod.STORAGE_F32)
    d, i, c = ctx.ivf_topk(pts, g["offsets"], g["queries"], g["probes"], 10, permutation=perm)
    assert ctx.last_algo == "segmented_scan"
    assert_topk_exact(d, i, g["dist"], g["ids"])
    assert np.array_equal(c, np.isfinite(g["dist"]).sum(1))
    # (b) the clustered copy as the reference builds it (byte storage this time): ids are positions
    cl = ctx.upload(g["points"][perm])
    assert cl.storage == gpann_mod.STORAGE_U8
    d2, i2, _ = ctx.ivf_topk(cl, g["offsets"], g["queries"], g["probes"], 10)
    fin = np.isfinite(g["dist"])
    assert np.array_equal(d2, g["dist"]) and np.array_equal(perm[i2[fin]], g["ids"][fin])
    # one probe per query = QueryBucket
    d3, i3, _ = ctx.ivf_topk(pts, g["offsets"], g["queries"][:4], g["probes"][:4, :1], 10, permutation=perm)
    d0, i0, _ = ctx.ivf_topk(pts, g["offsets"], g["queries"][:4], np.concatenate([g["probes"][:4, :1]] * 2, axis=1), 10, permutation=perm)
    assert np.array_equal(i3, i0) and np.array_equal(d3, d0)  # a bucket listed twice is probed once


def test_ivf_batch_large_buckets_split_across_tasks(ctx, gpann_mod, oracle):
    rng = np.random.default_rng(41)
    n, d, nb = 200_000, 32, 3
    P = rng.standard_normal((n, d)).astype(np.float32)
    sizes = [120_000, 79_000, 1_000]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    Q = rng.standard_normal((37, d)).astype(np.float32)
    probes = np.array([[0, 2]] * 20 + [[1, -1]] * 17, np.int32)
    pts = ctx.upload(P, storage=gpann_mod.STORAGE_F32)
    d1, i1, _ = ctx.ivf_topk(pts, off, Q, probes, 20)
    for q in (0, 19, 20, 36):
        rows = np.concatenate([np.arange(off[b], off[b + 1]) for b in probes[q] if b >= 0]).astype(np.int64)
        d0, i0, _ = oracle.ground_truth(P[rows], Q[q:q + 1], 20)
        assert_topk_close(d1[q:q + 1], i1[q:q + 1], d0, rows[i0].astype(np.uint32))


def test_knn_buckets_golden(ctx, gpann_mod):
    g = golden("crunch_buckets")
    k = int(g["k"])
    pts = ctx.upload(g["points"], storage=gpann_mod.STORAGE_F32)
    d, i, c = ctx.knn_buckets(pts, g["offsets"], g["members"], k)
    assert np.array_equal(c, g["count"])
    assert_topk_close(d, i, g["dist"], g["ids"])
    own = g["members"][:, None]
    assert not (i == own).any()  # a point is never its own neighbour
    g = golden("crunch_buckets_u8_ties")
    pts8 = ctx.upload_bytes(g["points"])
    d, i, c = ctx.knn_buckets(pts8, g["offsets"], g["members"], int(g["k"]))
    assert_topk_exact(d, i, g["dist"], g["ids"], c, g["count"])


def test_closest_leaders_batch_golden(ctx, gpann_mod, oracle):
    g = golden("closest_leaders_batch")
    P, lead = g["points"], g["leader_ids"]
    pts = ctx.upload(P, storage=gpann_mod.STORAGE_F32)
    leaders = ctx.upload(P[lead], storage=gpann_mod.STORAGE_F32)
    d, i, c = ctx.closest_leaders_batch(pts, leaders, 3)
    assert (c == 3).all()
    assert_topk_close(d, i, g["dist"], g["ids"])
    sub = np.array([5, 3999, 17, 17], np.uint32)
    d2, i2, _ = ctx.closest_leaders_batch(pts, leaders, 3, point_ids=sub)
    assert np.array_equal(i2, i[sub]) and np.array_equal(d2, d[sub])
    # MIPS build of the same call (distance = 1 - <p,q>), checked against the oracle
    d3, i3, _ = ctx.closest_leaders_batch(pts, leaders, 2, point_ids=sub, metric=gpann_mod.MIPS)
    d0, i0, _ = oracle.ground_truth(P[lead], P[sub], 2, metric=1)
    assert_topk_close(d3, i3, d0, i0)


def test_group_min_dist_matches_a_per_group_scan(ctx, gpann_mod, oracle):
    """BruteForceRouting (analyze_approximation_losses.cpp:42-56): per query the minimum distance to every shard's points."""
    rng = np.random.default_rng(61)
    n, d, shards = 30_000, 100, 13
    P = rng.standard_normal((n, d)).astype(np.float32)
    part = rng.integers(0, shards - 1, n).astype(np.int32)  # the last shard stays empty
    Q = rng.standard_normal((37, d)).astype(np.float32)
    pts = ctx.upload(P, storage=gpann_mod.STORAGE_F32)
    out = ctx.group_min_dist(pts, part, shards, Q)
    assert (out[:, shards - 1] == np.finfo(np.float32).max).all()
    for g in range(shards - 1):
        d0, _, _ = oracle.ground_truth(P[part == g], Q, 1)
        np.testing.assert_allclose(out[:, g], d0[:, 0], rtol=1e-6)
    # probe order = the reference's routing order
    probes = np.argsort(out, axis=1, kind="stable")
    assert (probes[:, -1] == shards - 1).all()
