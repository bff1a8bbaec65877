"""GPU tests of the round-2 boundary work: asynchronous *_device entry points (no host wait inside a step, scratch ordered
across streams), the bench-scale configuration on a small set (three sample levels + pacing + shared-memory lists), batched
fall-back scans, KMeans as one call (gpann_kmeans_run) and group contexts (gpann_init_multi) on the devices present."""
import os

import numpy as np
import pytest

from conftest import golden
from helpers import assert_topk_close, assert_topk_exact

pytestmark = pytest.mark.gpu


@pytest.fixture()
def torch_dev():
    import torch
    return torch, torch.device("cuda", 0)


def test_bench_configuration_on_a_small_set(ctx, gpann_mod, oracle, monkeypatch):
    """The combination the 100M headline runs -- three sample levels, the L2 pacing protocol, per-query lists in shared
    memory, the unrolled issuer (R = 4) with more than four query groups -- forced onto a set the oracle finishes."""
    monkeypatch.setenv("GPANN_TC_LEVELS", "1/2/4")
    monkeypatch.setenv("GPANN_TC_PACE_WINDOW", "8")
    rng = np.random.default_rng(21)
    n, nq, k = 1_000_000, 2700, 10
    P = rng.integers(0, 256, (n, 128), dtype=np.uint8)
    Q = rng.integers(0, 256, (nq, 128)).astype(np.float32)
    pts = ctx.upload_bytes(P)
    launches0 = ctx.launch_count
    d1, i1, c1 = ctx.topk(pts, Q, k)
    assert ctx.last_algo == "tensor_u8"
    # prep + 3 x (sample pass + merge) + main pass + merge + the gated float fall-back (scan + merge per query batch)
    assert ctx.launch_count - launches0 >= 9
    pick = np.linspace(0, nq - 1, 48).astype(int)
    d0, i0, c0 = oracle.ground_truth(P.astype(np.float32), Q[pick], k)
    assert_topk_exact(d1[pick], i1[pick], d0, i0, c1[pick], c0)
    # size-independent properties over ALL queries: sorted keys, exact distances of the reported ids
    assert (np.diff(d1, axis=1) >= 0).all()
    rows = P[i1[::97, 0]].astype(np.float32)
    np.testing.assert_array_equal(((rows - Q[::97]) ** 2).sum(1), d1[::97, 0])
    pts.free()


def test_device_entry_points_are_ordered_across_streams(ctx, gpann_mod, oracle, torch_dev):
    """A returned topk_device call still owns the ctx scratch: a second call on ANOTHER stream and a host call right
    behind it must each see their own queries (ev_order)."""
    torch, dev = torch_dev
    rng = np.random.default_rng(5)
    P = rng.integers(0, 256, (300_000, 128), dtype=np.uint8)
    pts = ctx.upload_bytes(P)
    QA = rng.integers(0, 256, (700, 128)).astype(np.float32)
    QB = rng.integers(0, 256, (900, 128)).astype(np.float32)
    QC = rng.integers(0, 256, (64, 128)).astype(np.float32)
    qa, qb = torch.from_numpy(QA).to(dev), torch.from_numpy(QB).to(dev)
    ka = torch.empty((700, 10), dtype=torch.int64, device=dev)
    kb = torch.empty((900, 10), dtype=torch.int64, device=dev)
    sa, sb = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    torch.cuda.synchronize()
    for _ in range(3):
        ctx.topk_device(pts, qa.data_ptr(), 700, 10, ka.data_ptr(), stream=sa.cuda_stream)
        ctx.topk_device(pts, qb.data_ptr(), 900, 10, kb.data_ptr(), stream=sb.cuda_stream)
        dc, ic, _ = ctx.topk(pts, QC, 10)  # host entry point on the ctx's own stream
    torch.cuda.synchronize()
    Pf = P.astype(np.float32)
    for keys, Q in ((ka, QA), (kb, QB)):
        d1, i1 = gpann_mod.unpack_keys(keys.cpu().numpy().view(np.uint64))
        d0, i0, _ = oracle.ground_truth(Pf, Q, 10)
        assert_topk_exact(d1, i1, d0, i0)
    d0, i0, _ = oracle.ground_truth(Pf, QC, 10)
    assert_topk_exact(dc, ic, d0, i0)
    pts.free()


def test_topk_device_fractional_queries_without_host_sync(ctx, gpann_mod, oracle, torch_dev):
    """Byte points, fractional queries through the ASYNCHRONOUS entry point: the tensor kernels are gated off on the device
    and the exact float scan queued behind them produces the keys."""
    torch, dev = torch_dev
    rng = np.random.default_rng(6)
    P = rng.integers(0, 256, (20_000, 96), dtype=np.uint8)
    Q = (rng.integers(0, 256, (50, 96)) + 0.5).astype(np.float32)
    pts = ctx.upload_bytes(P)
    q = torch.from_numpy(Q).to(dev)
    keys = torch.empty((50, 10), dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    ctx.topk_device(pts, q.data_ptr(), 50, 10, keys.data_ptr(), stream=st)
    torch.cuda.synchronize()
    d1, i1 = gpann_mod.unpack_keys(keys.cpu().numpy().view(np.uint64))
    d0, i0, _ = oracle.ground_truth(P.astype(np.float32), Q, 10)
    assert_topk_close(d1, i1, d0, i0)
    # and integer queries right after take the tensor kernel again
    Qi = rng.integers(0, 256, (50, 96)).astype(np.float32)
    d2, i2, _ = ctx.topk(pts, Qi, 10)
    assert ctx.last_algo == "tensor_u8"
    d0, i0, _ = oracle.ground_truth(P.astype(np.float32), Qi, 10)
    assert_topk_exact(d2, i2, d0, i0)
    pts.free()


def test_streaming_scan_batches_large_query_sets(ctx, gpann_mod, oracle):
    """the CUDA-core scan reserved gx*nq*k keys at once; now it serves the queries in batches."""
    rng = np.random.default_rng(7)
    P = rng.standard_normal((60_000, 24)).astype(np.float32)
    Q = rng.standard_normal((9_000, 24)).astype(np.float32)
    pts = ctx.upload(P)
    ctx.set_algo(gpann_mod.ALGO_STREAM)
    try:
        d1, i1, c1 = ctx.topk(pts, Q, 100, metric=gpann_mod.MIPS)
    finally:
        ctx.set_algo(gpann_mod.ALGO_AUTO)
    assert ctx.last_algo == "stream"
    pick = np.linspace(0, len(Q) - 1, 40).astype(int)
    d0, i0, _ = oracle.ground_truth(P, Q[pick], 100, metric=1)
    assert_topk_close(d1[pick], i1[pick], d0, i0)
    pts.free()


@pytest.mark.parametrize("name", ["kmeans_d32_k24", "kmeans_empty_clusters"])
def test_kmeans_run_equals_round_by_round(ctx, gpann_mod, name):
    """gpann_kmeans_run (centroids device-resident for 20 rounds, removed clusters masked, one renumbering at the end) returns
    what the round-by-round form with the reference's host-side RemoveEmptyClusters returns, and both match the golden."""
    g = golden(name)
    a1, c1 = gpann_mod.kmeans(ctx, g["points"], g["init"])
    a2, c2 = gpann_mod.kmeans_by_steps(ctx, g["points"], g["init"])
    assert c1.shape == c2.shape == g["centroids"].shape
    assert np.array_equal(a1, a2)
    np.testing.assert_array_equal(c1, c2)
    np.testing.assert_allclose(c1, g["centroids"], rtol=1e-4, atol=1e-4)
    assert (a1 == g["assign"]).mean() > 0.999


def test_kmeans_run_tensor_path_with_dying_clusters(gpann_mod, oracle):
    """Tensor-path Lloyd rounds where clusters die on the way: duplicates and far-away seeds lose every member."""
    rng = np.random.default_rng(11)
    n, d, kc = 60_000, 96, 48
    centres = rng.standard_normal((12, d)).astype(np.float32) * 5
    P = (centres[rng.integers(0, 12, n)] + rng.standard_normal((n, d))).astype(np.float32)
    C = np.concatenate([P[:40], 200 + rng.standard_normal((8, d)).astype(np.float32)]).astype(np.float32)
    with gpann_mod.Context(0) as c:
        c.set_algo(gpann_mod.ALGO_TENSOR)
        pts = c.upload(P, storage=gpann_mod.STORAGE_F32)
        a1, c1, s1 = c.kmeans_run(pts, C, rounds=6)
        assert c.last_algo == "kmeans_tensor_bf16x3"
        a2, c2 = gpann_mod.kmeans_by_steps(c, pts, C, rounds=6)
    a0, c0 = oracle.kmeans(P, C, rounds=6)
    assert c1.shape == c0.shape and c1.shape[0] < kc
    assert np.array_equal(a1, a2)
    # same survivors, objective within 1e-4; a handful of near-tie points may flip over the rounds (the
    # reference's own sums are order-dependent), which moves a few centroid coordinates in the third digit
    assert (a1 == a0).mean() > 0.999
    np.testing.assert_allclose(c1, c0, rtol=1e-2, atol=1e-2)
    o1, o0 = oracle.objective(P, c1, a1), oracle.objective(P, c0, a0)
    assert abs(o1 - o0) <= 1e-4 * o0
    assert int(s1.sum()) == n and (np.bincount(a1, minlength=len(s1)) == s1).all()


def test_group_context_matches_single_device(gpann_mod, oracle):
    """gpann_init_multi over the devices present (1 on the test box; 2+ exercise the NCCL exchanges): every host entry point
    returns what the single-device ctx returns."""
    import torch
    ndev = min(torch.cuda.device_count(), int(os.environ.get("GPANN_TEST_DEVICES", "8")))
    rng = np.random.default_rng(13)
    P = rng.integers(0, 256, (200_003, 128), dtype=np.uint8)
    Q = rng.integers(0, 256, (300, 128)).astype(np.float32)
    Pf = (rng.standard_normal((50_000, 96)) * 3).astype(np.float32)
    C = Pf[:64].copy()
    with gpann_mod.Context(0) as one, gpann_mod.Context(list(range(ndev))) as grp:
        assert grp.num_devices == ndev
        p1, pg = one.upload_bytes(P, id_base=7), grp.upload_bytes(P, id_base=7)
        d1, i1, c1 = one.topk(p1, Q, 10)
        dg, ig, cg = grp.topk(pg, Q, 10)
        assert_topk_exact(dg, ig, d1, i1, cg, c1)
        assert np.array_equal(grp.kth_distance(pg, Q, 10), d1[:, 9])
        dr, ir, _ = grp.topk(pg, Q[:5], 10, row_range=(1000, 150_000))
        d0, i0, _ = oracle.ground_truth(P[1000:150_000].astype(np.float32), Q[:5], 10)
        assert_topk_exact(dr, ir - 1007, d0, i0)
        ds, is_, cs = grp.topk_stream_bytes(P, Q, 10, id_base=7, chunk_rows=40_000)
        assert_topk_exact(ds, is_, d1, i1)
        df, if_, _ = grp.topk_stream(P.astype(np.float32), Q, 10, id_base=7, chunk_rows=30_000)
        assert_topk_exact(df, if_, d1, i1)
        f1, fg = one.upload(Pf, storage=gpann_mod.STORAGE_F32), grp.upload(Pf, storage=gpann_mod.STORAGE_F32)
        a1, cen1, s1 = one.kmeans_run(f1, C, rounds=5)
        ag, ceng, sg = grp.kmeans_run(fg, C, rounds=5)
        assert np.array_equal(a1, ag) and np.array_equal(s1, sg)
        np.testing.assert_allclose(ceng, cen1, rtol=1e-5, atol=1e-5)
        assert np.array_equal(grp.kmeans_assign(fg, C), one.kmeans_assign(f1, C))
        ast, cst, sst = grp.kmeans_step(fg, C)
        as1, cs1, ss1 = one.kmeans_step(f1, C)
        assert np.array_equal(ast, as1) and np.array_equal(sst, ss1)
        np.testing.assert_allclose(cst, cs1, rtol=1e-5, atol=1e-5)
        o1, og = one.objective(f1, cen1, a1), grp.objective(fg, cen1, a1)
        assert abs(o1 - og) <= 1e-9 * abs(o1)
