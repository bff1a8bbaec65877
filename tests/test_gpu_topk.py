"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the oracle on
the same seeded inputs, against the committed golden fixtures of the unmodified reference, and through
size-independent properties at larger sizes.  Bit-exact for integer-valued data; 1e-5 relative for floats."""
import numpy as np
import pytest

from conftest import golden
from helpers import REL, assert_topk_close, assert_topk_exact

pytestmark = pytest.mark.gpu


def _run(ctx, gp, P, Q, k, algo, metric=0, **kw):
    ctx.set_algo(algo)
    pts = ctx.upload(P, **kw)
    try:
        out = ctx.topk(pts, Q, k, metric)
        return out + (ctx.last_algo,)
    finally:
        pts.free()
        ctx.set_algo(gp.ALGO_AUTO)


# ---- golden fixtures of the unmodified reference ------------------------------------------------------------
@pytest.mark.parametrize("name,k,tensor", [("gt_u8_d128_k10", 10, True), ("gt_s8_d100_k10", 10, True),
                                           ("gt_u8_ties_d32_k10", 10, True)])
@pytest.mark.parametrize("algo", ["stream", "tensor"])
def test_golden_bytes_bit_exact(ctx, gpann_mod, name, k, tensor, algo):
    g = golden(name)
    a = gpann_mod.ALGO_STREAM if algo == "stream" else gpann_mod.ALGO_TENSOR
    d, i, c, used = _run(ctx, gpann_mod, g["points"], g["queries"].astype(np.float32), k, a)
    assert used.startswith(algo)
    assert_topk_exact(d, i, g["dist"], g["ids"], c, g["count"])


@pytest.mark.parametrize("name,k,metric", [("gt_f32_d100_k10", 10, 0), ("gt_f32_d96_k100", 100, 0),
                                           ("gt_f32_n_lt_k", 10, 0), ("gt_mips_d64_k5", 5, 1)])
def test_golden_float(ctx, gpann_mod, name, k, metric):
    g = golden(name)
    d, i, c, used = _run(ctx, gpann_mod, g["points"], g["queries"], k, gpann_mod.ALGO_AUTO, metric)
    assert np.array_equal(c, g["count"])
    assert_topk_close(d, i, g["dist"], g["ids"])
    if "kth" in g.files:
        pts = ctx.upload(g["points"])
        kth = ctx.kth_distance(pts, g["queries"], k)
        np.testing.assert_allclose(kth, g["kth"], rtol=REL)


def test_golden_closest_leaders_and_ivf(ctx, gpann_mod):
    g = golden("closest_leaders")
    d, i = gpann_mod.closest_leaders(ctx, g["points"], g["leaders"], int(g["my_id"]), 5)
    assert np.array_equal(i, g["ids"])
    np.testing.assert_allclose(d, g["dist"], rtol=REL)
    # InvertedIndex::Query = range-restricted scans over the clustered copy, local positions as tie-break ids,
    # remapped through the permutation afterwards (inverted_index.h:43-60)
    g = golden("ivf_query")
    part = g["partition"]
    clusters = [np.nonzero(part == b)[0] for b in range(8)]
    offsets = np.concatenate([[0], np.cumsum([len(c) for c in clusters])])
    perm = np.concatenate(clusters).astype(np.uint32)
    pts = ctx.upload(g["points"][perm], storage=gpann_mod.STORAGE_F32)
    keys = []
    for b in g["buckets"]:
        d, i, c = ctx.topk(pts, g["query"][None, :], 10, row_range=(int(offsets[b]), int(offsets[b + 1])))
        keys += [(float(d[0, r]), int(i[0, r])) for r in range(c[0])]
    keys.sort()
    assert np.array_equal(perm[[j for _, j in keys[:10]]], g["ids"])


# ---- oracle on the same seeded inputs ------------------------------------------------------------------------
@pytest.mark.parametrize("n,d,nq,k,signed", [(20000, 128, 300, 10, False), (5000, 96, 130, 10, False),
                                             (70001, 128, 600, 10, False), (9000, 100, 257, 32, True),
                                             (3000, 128, 128, 100, False), (300, 64, 40, 10, False),
                                             (255, 128, 17, 1, False), (30000, 128, 129, 16, False),
                                             (30000, 128, 129, 17, False), (40000, 96, 700, 10, False),
                                             (40000, 128, 700, 10, True)])
def test_tensor_bytes_vs_oracle(ctx, gpann_mod, oracle, n, d, nq, k, signed):
    rng = np.random.default_rng(n + d)
    lo, hi = (-128, 128) if signed else (0, 256)
    P = rng.integers(lo, hi, (n, d)).astype(np.int8 if signed else np.uint8)
    Q = rng.integers(lo, hi, (nq, d)).astype(np.float32)
    d1, i1, c1, used = _run(ctx, gpann_mod, P, Q, k, gpann_mod.ALGO_TENSOR)
    assert used == ("tensor_s8" if signed else "tensor_u8")
    d0, i0, c0 = oracle.ground_truth(P.astype(np.float32), Q, k)
    assert_topk_exact(d1, i1, d0, i0, c1, c0)


def test_tensor_zero_queries_and_padding_slots(ctx, gpann_mod, oracle):
    """Real all-zero queries must be answered like any other query, while the zero vectors that pad a batch to a
    multiple of 128 query slots must stay silent (they are pinned above every accumulator value in the kernel)."""
    rng = np.random.default_rng(11)
    P = rng.integers(0, 256, (50000, 128)).astype(np.uint8)
    P[::997] = 0  # a few zero rows: distance 0 to the zero queries, ties broken by id
    Q = rng.integers(0, 256, (131, 128)).astype(np.float32)
    Q[[0, 64, 130]] = 0.0
    d1, i1, c1, used = _run(ctx, gpann_mod, P, Q, 10, gpann_mod.ALGO_TENSOR)
    assert used == "tensor_u8"
    d0, i0, c0 = oracle.ground_truth(P.astype(np.float32), Q, 10)
    assert_topk_exact(d1, i1, d0, i0, c1, c0)


def test_tensor_paced_splits_bit_exact(ctx, gpann_mod, oracle, monkeypatch):
    """The CTAs of a point split pace each other inside an L2 window (topk_tc.cu, TMA producer).  With a window of 8
    tiles the protocol (publish, wait for the slowest, finished CTAs release the others) runs constantly on a set small
    enough for the oracle; several query groups (nq > 512) are needed for a split to have more than one CTA."""
    monkeypatch.setenv("GPANN_TC_PACE_WINDOW", "8")
    rng = np.random.default_rng(21)
    P = rng.integers(0, 256, (1_500_000, 128)).astype(np.uint8)
    Q = rng.integers(0, 256, (700, 128)).astype(np.float32)
    d1, i1, c1, used = _run(ctx, gpann_mod, P, Q, 10, gpann_mod.ALGO_TENSOR)
    assert used == "tensor_u8"
    d0, i0, c0 = oracle.ground_truth(P.astype(np.float32), Q, 10)
    assert_topk_exact(d1, i1, d0, i0, c1, c0)


def test_tensor_find_log_shares_keys_between_splits_bit_exact(ctx, gpann_mod, oracle, monkeypatch):
    """k <= 16 with several point splits per query group: the splits append every accepted key to the group's find log
    and apply each other's entries as foreign keys (id field dropped, blanked before the CTA reports its list).  Heavy
    ties (binary coordinates, duplicated rows) make the tie-break on ids and the foreign-key ordering matter; the result
    must be bit-identical to the oracle with the log on and with GPANN_TC_NOLOG."""
    rng = np.random.default_rng(77)
    P = rng.integers(0, 2, (60_000, 64)).astype(np.uint8)
    P[5000:5040] = P[7]           # exact duplicates across different splits (rows are sorted by norm, ids break the ties)
    P[40_000:40_016] = P[7]
    Q = np.concatenate([P[:24], rng.integers(0, 2, (276, 64)).astype(np.uint8)]).astype(np.float32)
    d0, i0, c0 = oracle.ground_truth(P.astype(np.float32), Q, 16)
    for nolog in (False, True):
        if nolog:
            monkeypatch.setenv("GPANN_TC_NOLOG", "1")
        d1, i1, c1, used = _run(ctx, gpann_mod, P, Q, 16, gpann_mod.ALGO_TENSOR)
        assert used == "tensor_u8"
        assert_topk_exact(d1, i1, d0, i0, c1, c0)
    monkeypatch.delenv("GPANN_TC_NOLOG")
    # float rows: foreign keys carry exact re-ranked distances
    Pf = rng.standard_normal((50_000, 48)).astype(np.float32)
    Qf = rng.standard_normal((200, 48)).astype(np.float32)
    d1, i1, c1, used = _run(ctx, gpann_mod, Pf, Qf, 10, gpann_mod.ALGO_TENSOR)
    assert used == "tensor_bf16"
    d0, i0, c0 = oracle.ground_truth(Pf, Qf, 10)
    assert np.array_equal(c1, c0)
    assert_topk_close(d1, i1, d0, i0)


def test_tensor_heavy_ties_and_duplicates(ctx, gpann_mod, oracle):
    rng = np.random.default_rng(5)
    P = rng.integers(0, 2, (4000, 128)).astype(np.uint8)
    P[1000:1100] = P[0]  # exact duplicates of one row
    Q = np.concatenate([P[:40], rng.integers(0, 2, (100, 128)).astype(np.uint8)]).astype(np.float32)
    d1, i1, c1, _ = _run(ctx, gpann_mod, P, Q, 20, gpann_mod.ALGO_TENSOR)
    d0, i0, c0 = oracle.ground_truth(P.astype(np.float32), Q, 20)
    assert_topk_exact(d1, i1, d0, i0, c1, c0)


@pytest.mark.parametrize("storage,d,nq,k,metric", [("f32", 128, 5, 10, 0), ("f32", 100, 33, 10, 0), ("f32", 96, 3, 100, 0),
                                                    ("f32", 20, 1, 3, 0), ("f32", 64, 16, 5, 1), ("u8", 128, 4, 10, 0),
                                                    ("s8", 100, 2, 7, 0), ("u8", 200, 20, 10, 0)])
def test_stream_vs_oracle(ctx, gpann_mod, oracle, storage, d, nq, k, metric):
    rng = np.random.default_rng(d * 7 + nq)
    n = 6001
    if storage == "f32":
        P = rng.standard_normal((n, d)).astype(np.float32)
        Q = rng.standard_normal((nq, d)).astype(np.float32)
    else:
        lo, hi = (-128, 128) if storage == "s8" else (0, 256)
        P = rng.integers(lo, hi, (n, d)).astype(np.float32)
        Q = rng.integers(lo, hi, (nq, d)).astype(np.float32)
    d1, i1, c1, used = _run(ctx, gpann_mod, P, Q, k, gpann_mod.ALGO_STREAM, metric)
    assert used == "stream"
    d0, i0, c0 = oracle.ground_truth(P, Q, k, metric)
    assert np.array_equal(c1, c0)
    if storage == "f32":
        assert_topk_close(d1, i1, d0, i0)
    else:
        assert_topk_exact(d1, i1, d0, i0)


def test_auto_storage_detection_and_float_queries_on_byte_points(ctx, gpann_mod, oracle):
    rng = np.random.default_rng(9)
    P = rng.integers(0, 256, (3000, 128)).astype(np.float32)
    pts = ctx.upload(P)
    assert pts.storage == gpann_mod.STORAGE_U8
    assert ctx.upload(P - 128).storage == gpann_mod.STORAGE_S8
    assert ctx.upload(P + 0.5).storage == gpann_mod.STORAGE_F32
    # non-integer queries against byte points: the tensor kernel must decline and the exact float scan answer
    Q = (rng.integers(0, 256, (40, 128)) + 0.25).astype(np.float32)
    d1, i1, c1 = ctx.topk(pts, Q, 10)
    assert ctx.last_algo == "stream"
    d0, i0, c0 = oracle.ground_truth(P, Q, 10)
    assert_topk_close(d1, i1, d0, i0)


def test_edge_cases(ctx, gpann_mod, oracle):
    rng = np.random.default_rng(2)
    P = rng.standard_normal((50, 16)).astype(np.float32)
    pts = ctx.upload(P)
    # n < k
    d1, i1, c1 = ctx.topk(pts, P[:3], 64)
    assert (c1 == 50).all() and np.isinf(d1[:, 50:]).all() and (i1[:, 50:] == gpann_mod.INVALID_ID).all()
    # empty range, empty query batch
    d1, i1, c1 = ctx.topk(pts, P[:2], 5, row_range=(7, 7))
    assert (c1 == 0).all() and np.isinf(d1).all()
    d1, i1, c1 = ctx.topk(pts, np.zeros((0, 16), np.float32), 5)
    assert d1.shape == (0, 5)
    # bad arguments fail loudly
    with pytest.raises(gpann_mod.GpannError):
        ctx.topk(pts, P[:2], 0)
    with pytest.raises(gpann_mod.GpannError):
        ctx.topk(pts, P[:2], 5, row_range=(10, 100))
    ctx.set_algo(gpann_mod.ALGO_TENSOR)
    with pytest.raises(gpann_mod.GpannError):
        ctx.topk(pts, P[:2], 5, row_range=(0, 10))  # the tensor kernels scan whole sets: must refuse, not fall back
    d1, i1, c1 = ctx.topk(pts, P[:3], 5)  # tiny float set through the bf16 tensor path
    assert ctx.last_algo == "tensor_bf16"
    d0, i0, _ = oracle.ground_truth(P, P[:3], 5)
    assert np.array_equal(i1, i0)
    ctx.set_algo(gpann_mod.ALGO_AUTO)
    # id_base shifts reported ids (shard-local -> global)
    p2 = ctx.upload(P, id_base=1000)
    d2, i2, _ = ctx.topk(p2, P[:4], 3)
    d0, i0, _ = oracle.ground_truth(P, P[:4], 3)
    assert np.array_equal(i2, i0 + 1000)


def test_long_rows_use_shorter_tiles(ctx, gpann_mod, oracle):
    """Rows too long for 128 of them in shared memory (GIST-960, d = 2000, d = 4096 f32; long byte rows): the one-thread-per-row
    kernels stage fewer rows per tile instead of refusing -- top-k, nearest centroids and the objective against the oracle."""
    rng = np.random.default_rng(960)
    for d, n, nq in ((960, 3000, 9), (2000, 700, 5), (4096, 300, 3)):
        P = rng.standard_normal((n, d)).astype(np.float32)
        Q = rng.standard_normal((nq, d)).astype(np.float32)
        pts = ctx.upload(P)
        try:
            d1, i1, c1 = ctx.topk(pts, Q, 10)
            assert ctx.last_algo == "stream"
            d0, i0, c0 = oracle.ground_truth(P, Q, 10)
            assert np.array_equal(c1, c0)
            assert_topk_close(d1, i1, d0, i0)
            C = P[rng.choice(n, 17, replace=False)].copy()
            a1 = ctx.kmeans_assign(pts, C)
            a0, _ = oracle.nearest_centers(P, C)
            assert np.array_equal(a1, a0)
            o1 = ctx.objective(pts, C, a1)
            o0 = oracle.objective(P, C, a1)
            assert abs(o1 - o0) <= 1e-5 * abs(o0)
        finally:
            pts.free()
    Pb = rng.integers(0, 256, (2500, 3000)).astype(np.uint8)  # byte rows of 3000: exact integers
    Qb = rng.integers(0, 256, (6, 3000)).astype(np.float32)
    pts = ctx.upload(Pb)
    try:
        d1, i1, c1 = ctx.topk(pts, Qb, 5)
        d0, i0, c0 = oracle.ground_truth(Pb.astype(np.float32), Qb, 5)
        assert_topk_exact(d1, i1, d0, i0, c1, c0)
    finally:
        pts.free()


# ---- properties at sizes the oracle cannot reach quickly -----------------------------------------------------------
def test_large_tensor_properties(ctx, gpann_mod, oracle, ref):
    rng = np.random.default_rng(77)
    n, d, nq, k = 2_000_000, 128, 1000, 10
    P = rng.integers(0, 256, (n, d), dtype=np.uint8)
    Q = rng.integers(0, 256, (nq, d), dtype=np.uint8).astype(np.float32)
    Q[:50] = P[rng.integers(0, n, 50)]  # planted exact matches: distance 0 must come first
    pts = ctx.upload_bytes(P)
    ctx.set_algo(gpann_mod.ALGO_TENSOR)
    d1, i1, c1 = ctx.topk(pts, Q, k)
    ctx.set_algo(gpann_mod.ALGO_AUTO)
    assert (c1 == k).all()
    assert (d1[:50, 0] == 0).all()
    # sortedness by (dist, id) and id uniqueness
    assert (np.diff(d1, axis=1) >= 0).all()
    tie = np.diff(d1, axis=1) == 0
    assert (np.diff(i1.astype(np.int64), axis=1)[tie] > 0).all()
    assert all(len(set(r)) == k for r in i1)
    # reported distances are the true distances of the reported ids (integer exact)
    sel = P[i1[:64].ravel()].astype(np.int64).reshape(64, k, d)
    true = ((sel - Q[:64, None, :].astype(np.int64)) ** 2).sum(-1)
    assert np.array_equal(true, d1[:64].astype(np.int64))
    # sampled queries against the oracle, bit-exact
    d0, i0, _ = oracle.ground_truth(P.astype(np.float32), Q[:6], k)
    assert_topk_exact(d1[:6], i1[:6], d0, i0)
    # split invariance: top-k of the union == merge of the top-k of two halves (what the multi-GPU path relies on)
    h = n // 2
    pa, pb = ctx.upload_bytes(P[:h]), ctx.upload_bytes(P[h:], id_base=h)
    da, ia, _ = ctx.topk(pa, Q[:200], k)
    db, ib, _ = ctx.topk(pb, Q[:200], k)
    both = np.concatenate([da, db], 1), np.concatenate([ia, ib], 1)
    order = np.lexsort((both[1], both[0]), axis=1)[:, :k]
    assert np.array_equal(np.take_along_axis(both[1], order, 1), i1[:200])
    # the reference's own self-check accepts the GPU ground truth (recall.h:85-134): recomputed kth distances equal
    kth = ref.convert_ground_truth(P[:200000].astype(np.float32), Q[:20], k, *ctx.topk(ctx.upload_bytes(P[:200000]), Q[:20], k)[:2])
    d_small, _, _ = ctx.topk(ctx.upload_bytes(P[:200000]), Q[:20], k)
    assert np.array_equal(kth, d_small[:, -1])


def test_sharded_topk_device_path_matches_single_scan(ctx, gpann_mod, oracle):
    """The per-rank calls of the multi-GPU path (gpann_topk_device on shards with global ids +
    gpann_merge_keys_device), driven on one GPU with torch tensors."""
    import torch
    from gpann.dist import CudaShard, shard_bounds
    rng = np.random.default_rng(4)
    n, d, nq, k, world = 30000, 128, 70, 10, 4
    P = rng.integers(0, 8, (n, d), dtype=np.uint8)
    Q = rng.integers(0, 8, (nq, d)).astype(np.float32)
    dev = torch.device("cuda:0")
    q = torch.from_numpy(Q).to(dev)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n, r, world)
        shard = CudaShard(ctx, ctx.upload_bytes(P[lo:hi], id_base=lo), dev)
        parts.append(shard.local_topk(q, k))
    dist_out, id_out = shard.merge(torch.stack(parts), k)
    torch.cuda.synchronize()
    d0, i0, _ = oracle.ground_truth(P.astype(np.float32), Q, k)
    assert_topk_exact(dist_out.cpu().numpy(), id_out.cpu().numpy().view(np.uint32), d0, i0)


# ---- float data through the bf16 tensor filter + exact fp32 re-rank -----------------------------------------------
@pytest.mark.parametrize("n,d,nq,k,kind", [(20000, 128, 300, 10, "normal"), (30000, 96, 130, 100, "normal"),
                                           (50001, 100, 257, 10, "normal"), (8000, 64, 40, 10, "normal"),
                                           (20000, 128, 100, 10, "offset"), (20000, 96, 100, 10, "unit"),
                                           (6000, 20, 50, 5, "clustered"), (20000, 64, 600, 10, "normal")])
def test_tensor_float_vs_oracle(ctx, gpann_mod, oracle, n, d, nq, k, kind):
    rng = np.random.default_rng(n + d + k)
    P = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    if kind == "offset":      # far from the origin: worst case for the bf16 filter's error bound (still exact)
        P += 30.0
        Q += 30.0
    elif kind == "unit":      # DEEP-like normalised vectors
        P /= np.linalg.norm(P, axis=1, keepdims=True)
        Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    elif kind == "clustered":  # queries are perturbed points: tiny distances, strong cancellation
        Q = P[rng.integers(0, n, nq)] + 1e-3 * rng.standard_normal((nq, d)).astype(np.float32)
    d1, i1, c1, used = _run(ctx, gpann_mod, P, Q, k, gpann_mod.ALGO_TENSOR, storage=gpann_mod.STORAGE_F32)
    assert used == "tensor_bf16"
    d0, i0, c0 = oracle.ground_truth(P, Q, k)
    assert np.array_equal(c1, c0)
    assert_topk_close(d1, i1, d0, i0)
    # the exact re-rank uses the reference's summation order: distances agree to the last bits, not just 1e-5
    fin = np.isfinite(d0)
    assert np.abs(d1[fin] - d0[fin]).max() <= 4e-7 * max(1.0, float(d0[fin].max()))


def test_tensor_float_matches_stream_kernel_large(ctx, gpann_mod):
    rng = np.random.default_rng(123)
    n, d, nq, k = 400_000, 96, 512, 10
    P = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    pts = ctx.upload(P, storage=gpann_mod.STORAGE_F32)
    ctx.set_algo(gpann_mod.ALGO_TENSOR)
    d1, i1, c1 = ctx.topk(pts, Q, k)
    assert ctx.last_algo == "tensor_bf16"
    ctx.set_algo(gpann_mod.ALGO_STREAM)
    d0, i0, c0 = ctx.topk(pts, Q[:64], k)
    ctx.set_algo(gpann_mod.ALGO_AUTO)
    assert np.array_equal(i1[:64], i0) and np.array_equal(d1[:64].view(np.uint32), d0.view(np.uint32))
    assert (np.diff(d1, axis=1) >= 0).all() and all(len(set(r)) == k for r in i1)


# ---- streamed scan of a host-resident byte set (gpann_topk_stream_bytes): chunking must not change anything ----------
@pytest.mark.parametrize("n,d,nq,k,chunk,dtype", [
    (100_000, 128, 300, 10, 30_000, np.uint8), (70_001, 100, 129, 10, 8192, np.int8), (5000, 128, 40, 100, 0, np.uint8),
    (300, 64, 33, 10, 256, np.uint8), (40_000, 32, 200, 1, 10_000, np.uint8),
])
def test_stream_bytes_matches_oracle(ctx, oracle, n, d, nq, k, chunk, dtype):
    rng = np.random.default_rng(n + k)
    lo, hi = (0, 256) if dtype == np.uint8 else (-128, 128)
    P = rng.integers(lo, hi, (n, d)).astype(dtype)
    if n > 1000:
        P[n // 2:n // 2 + 50] = P[:50]  # exact duplicates across chunk boundaries: ties must keep the smaller id
    Q = rng.integers(lo, hi, (nq, d)).astype(np.float32)
    d1, i1, c1 = ctx.topk_stream_bytes(P, Q, k, id_base=7, chunk_rows=chunk)
    assert ctx.last_algo.endswith("_stream")
    d0, i0, c0 = oracle.ground_truth(P.astype(np.float32), Q, k)
    assert_topk_exact(d1, i1 - 7, d0, i0, c1, c0)


@pytest.mark.parametrize("n,d,nq,k,chunk", [(60_000, 96, 200, 10, 16_384), (20_000, 100, 64, 100, 0), (9000, 128, 300, 10, 2048)])
def test_stream_float_rows_matches_oracle(ctx, oracle, n, d, nq, k, chunk):
    rng = np.random.default_rng(n)
    P = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    d1, i1, c1, kth = ctx.topk_stream(P, Q, k, chunk_rows=chunk, want_kth=True)
    assert ctx.last_algo == "tensor_bf16_stream"
    d0, i0, c0 = oracle.ground_truth(P, Q, k)
    assert_topk_close(d1, i1, d0, i0)
    np.testing.assert_allclose(kth, d0[:, k - 1], rtol=REL)


def test_stream_float_rows_holding_bytes_take_the_integer_kernel(ctx, oracle):
    rng = np.random.default_rng(77)
    P = rng.integers(0, 256, (50_000, 128)).astype(np.float32)
    Q = rng.integers(0, 256, (150, 128)).astype(np.float32)
    d1, i1, c1 = ctx.topk_stream(P, Q, 10, chunk_rows=12_000)
    assert ctx.last_algo == "tensor_u8_stream"
    d0, i0, c0 = oracle.ground_truth(P, Q, 10)
    assert_topk_exact(d1, i1, d0, i0, c1, c0)
    # a later chunk that is not byte-valued refutes the first chunk's guess: the call restarts on the float path
    P[40_000, 3] = 0.5
    d1, i1, c1 = ctx.topk_stream(P, Q, 10, chunk_rows=12_000)
    assert ctx.last_algo == "tensor_bf16_stream"
    d0, i0, c0 = oracle.ground_truth(P, Q, 10)
    assert_topk_close(d1, i1, d0, i0)


def test_mips_on_the_tensor_path(gpann_mod, oracle):
    """The -DMIPS_DISTANCE build of the scan (distance = 1 - <q,x>, src/dist.cpp:50-52) on the tcgen05 kernels: exact integer
    dot products for byte data (negative dots included), bf16 filter + sequential-FMA re-rank for float data."""
    rng = np.random.default_rng(77)
    with gpann_mod.Context(0) as c:
        g = golden("gt_mips_d64_k5")
        c.set_algo(gpann_mod.ALGO_TENSOR)
        pts = c.upload(g["points"], storage=gpann_mod.STORAGE_F32)
        d1, i1, c1 = c.topk(pts, g["queries"], 5, metric=gpann_mod.MIPS)
        assert c.last_algo == "tensor_bf16"
        assert_topk_close(d1, i1, g["dist"], g["ids"])
        c.set_algo(gpann_mod.ALGO_AUTO)
        for dtype, lo, hi in ((np.uint8, 0, 256), (np.int8, -128, 128)):
            P = rng.integers(lo, hi, (150_000, 100)).astype(dtype)
            Q = rng.integers(lo, hi, (200, 100)).astype(np.float32)
            pb = c.upload_bytes(P)
            d1, i1, _ = c.topk(pb, Q, 10, metric=gpann_mod.MIPS)
            assert c.last_algo.startswith("tensor_")
            pick = np.linspace(0, 199, 24).astype(int)
            d0, i0, _ = oracle.ground_truth(P.astype(np.float32), Q[pick], 10, metric=1)
            assert_topk_exact(d1[pick], i1[pick], d0, i0)
            pb.free()
        Pf = (rng.standard_normal((300_000, 96)) * rng.uniform(0.2, 3.0, (300_000, 1))).astype(np.float32)
        Qf = rng.standard_normal((400, 96)).astype(np.float32)
        pf = c.upload(Pf, storage=gpann_mod.STORAGE_F32)
        d1, i1, _ = c.topk(pf, Qf, 20, metric=gpann_mod.MIPS)
        assert c.last_algo == "tensor_bf16"
        pick = np.linspace(0, 399, 32).astype(int)
        d0, i0, _ = oracle.ground_truth(Pf, Qf[pick], 20, metric=1)
        assert_topk_close(d1[pick], i1[pick], d0, i0, atol=1e-4)
