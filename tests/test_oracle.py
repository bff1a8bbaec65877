"""CPU tests: the oracle (oracle/gpann_oracle.cpp) against the golden vectors generated from the unmodified
reference, against the compiled reference itself when it is present, and TopN's edge semantics."""
import numpy as np
import pytest

from conftest import golden
from helpers import REL, assert_topk_close, assert_topk_exact


def _check_gt(oracle, g, k, metric=0, exact=False):
    P, Q = g["points"].astype(np.float32), g["queries"].astype(np.float32)
    d, i, c = oracle.ground_truth(P, Q, k, metric)
    assert np.array_equal(c, g["count"])
    if exact:
        assert_topk_exact(d, i, g["dist"], g["ids"])
    else:
        assert_topk_close(d, i, g["dist"], g["ids"])


def test_golden_u8_bit_exact(oracle):
    _check_gt(oracle, golden("gt_u8_d128_k10"), 10, exact=True)


def test_golden_s8_d100_bit_exact(oracle):
    _check_gt(oracle, golden("gt_s8_d100_k10"), 10, exact=True)


def test_golden_ties_keep_smaller_id(oracle):
    g = golden("gt_u8_ties_d32_k10")
    _check_gt(oracle, g, 10, exact=True)
    # the fixture really has ties inside the lists, and ids ascend inside each tie group
    d, i = g["dist"], g["ids"]
    tied = d[:, 1:] == d[:, :-1]
    assert tied.any()
    assert (i[:, 1:][tied] > i[:, :-1][tied]).all()


def test_golden_f32(oracle):
    g = golden("gt_f32_d100_k10")
    _check_gt(oracle, g, 10)
    kth = oracle.kth_distance(g["points"], g["queries"], 10)
    np.testing.assert_allclose(kth, g["kth"], rtol=REL)
    _check_gt(oracle, golden("gt_f32_d96_k100"), 100)


def test_golden_n_less_than_k(oracle):
    g = golden("gt_f32_n_lt_k")
    _check_gt(oracle, g, 10)
    assert (g["count"] == 7).all()


def test_golden_mips(oracle):
    _check_gt(oracle, golden("gt_mips_d64_k5"), 5, metric=1)


def test_golden_kmeans(oracle):
    g = golden("kmeans_d32_k24")
    assign, cent = oracle.kmeans(g["points"], g["init"])
    assert cent.shape == g["centroids"].shape
    assert (assign == g["assign"]).mean() > 0.999
    np.testing.assert_allclose(cent, g["centroids"], rtol=1e-4, atol=1e-5)
    obj = oracle.objective(g["points"], cent, assign)
    assert abs(obj - float(g["objective"])) <= 1e-4 * float(g["objective"])
    idx, rows = oracle.random_sample(g["points"], 24, 555)
    assert np.array_equal(rows, g["init"])
    assert (np.diff(idx) > 0).all()  # std::sample keeps index order


def test_golden_kmeans_empty_clusters(oracle):
    g = golden("kmeans_empty_clusters")
    assign, cent = oracle.kmeans(g["points"], g["init"])
    assert cent.shape == g["centroids"].shape == (6, 32)
    assert np.array_equal(assign, g["assign"])
    np.testing.assert_allclose(cent, g["centroids"], rtol=1e-4, atol=1e-5)
    with pytest.raises(RuntimeError):
        oracle.kmeans(g["points"], np.zeros((0, 32), np.float32))  # kmeans.cpp:229-231


def test_golden_closest_leaders_and_ivf(oracle):
    g = golden("closest_leaders")
    d, i = oracle.closest_leaders(g["points"], g["leaders"], int(g["my_id"]), 5)
    assert np.array_equal(i, g["ids"])
    np.testing.assert_allclose(d, g["dist"], rtol=REL)
    g = golden("ivf_query")
    part = g["partition"]
    clusters = [np.nonzero(part == b)[0] for b in range(8)]
    offsets = np.concatenate([[0], np.cumsum([len(c) for c in clusters])]).astype(np.int32)
    perm = np.concatenate(clusters).astype(np.uint32)
    d, i = oracle.ivf_query(g["points"][perm], offsets, perm, g["query"], 10, g["buckets"])
    assert np.array_equal(i, g["ids"])
    np.testing.assert_allclose(d, g["dist"], rtol=REL)


def test_topn_duplicates_and_k1(oracle):
    # identical points: ids ascend; k=1 returns the lowest id among equal distances
    P = np.ones((50, 8), np.float32)
    Q = np.zeros((2, 8), np.float32)
    d, i, c = oracle.ground_truth(P, Q, 4)
    assert np.array_equal(i, np.tile(np.arange(4, dtype=np.uint32), (2, 1)))
    assert (d == 8.0).all()
    d, i, c = oracle.ground_truth(P, Q, 1)
    assert (i == 0).all()


def test_oracle_vs_compiled_reference(oracle, ref):
    """Pin the restatement against the reference itself on fresh seeded inputs (skipped off the authoring box)."""
    rng = np.random.default_rng(7)
    for d, k, intval in ((128, 10, True), (96, 100, False), (100, 10, False), (20, 3, False)):
        n, nq = 4000, 50
        if intval:
            P = rng.integers(0, 256, (n, d)).astype(np.float32)
            Q = rng.integers(0, 256, (nq, d)).astype(np.float32)
        else:
            P = rng.standard_normal((n, d)).astype(np.float32)
            Q = rng.standard_normal((nq, d)).astype(np.float32)
        d1, i1, c1 = oracle.ground_truth(P, Q, k)
        d2, i2, c2 = ref.ground_truth(P, Q, k)
        assert np.array_equal(c1, c2)
        if intval:
            assert np.array_equal(d1, d2) and np.array_equal(i1, i2)
        else:
            np.testing.assert_allclose(d1, d2, rtol=REL)
            # id sets may differ only inside 1e-5-relative ties at the kth distance
            for q in range(nq):
                diff = set(i1[q]) ^ set(i2[q])
                for j in diff:
                    dj = ref.distance(P[j], Q[q])
                    assert abs(dj - d2[q, -1]) <= REL * d2[q, -1]
    # the reference's own self-check (recall.h:74-137) accepts the oracle's ground truth
    P = rng.integers(0, 256, (2000, 128)).astype(np.float32)
    Q = rng.integers(0, 256, (20, 128)).astype(np.float32)
    d1, i1, _ = oracle.ground_truth(P, Q, 10)
    kth = ref.convert_ground_truth(P, Q, 10, d1, i1)
    assert np.array_equal(kth, d1[:, -1])
    assert ref.recall(d1, kth, 10) == 1.0


def test_crunch_bucket_matches_reference_golden(oracle):
    """ApproximateKNNGraphBuilder::CrunchBucket (src/knn_graph.h:165-185): float data within 1e-5, byte-valued data bit-exact
    including the (distance, global id) tie-break."""
    g = golden("crunch_buckets")
    k = int(g["k"])
    for b in range(len(g["offsets"]) - 1):
        lo, hi = int(g["offsets"][b]), int(g["offsets"][b + 1])
        d, i, c = oracle.crunch_bucket(g["points"], g["members"][lo:hi], k)
        assert np.array_equal(c, g["count"][lo:hi])
        assert_topk_close(d, i, g["dist"][lo:hi], g["ids"][lo:hi])
    g = golden("crunch_buckets_u8_ties")
    k = int(g["k"])
    for b in range(len(g["offsets"]) - 1):
        lo, hi = int(g["offsets"][b]), int(g["offsets"][b + 1])
        d, i, c = oracle.crunch_bucket(g["points"].astype(np.float32), g["members"][lo:hi], k)
        assert_topk_exact(d, i, g["dist"][lo:hi], g["ids"][lo:hi], c, g["count"][lo:hi])


def test_ivf_batch_golden_is_the_per_query_scan(oracle):
    """The batched fixture is the reference's InvertedIndex::Query run per query: ties break on the POSITION in the
    clustered copy, reported ids go through the permutation (src/inverted_index.h:43-60)."""
    g = golden("ivf_batch")
    perm, off = g["permutation"], g["offsets"].astype(np.int32)
    clustered = g["points"][perm]
    for q in range(len(g["queries"])):
        b = [int(x) for x in g["probes"][q] if x >= 0]
        d, i = oracle.ivf_query(clustered, off, perm, g["queries"][q], 10, b)
        n = len(d)
        assert np.array_equal(i, g["ids"][q, :n]) and np.array_equal(d, g["dist"][q, :n])
        assert np.isinf(g["dist"][q, n:]).all()
