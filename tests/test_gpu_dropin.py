"""GPU tests of the C++ drop-in layer (gp-ann_b200/cxx/: dist.h, topn.h, recall.h, kmeans.h with the reference's
signatures, forwarding to libgpann_cuda.so).  The layer is driven through gp-ann_b200/libgpann_dropin.so, whose C shim
mirrors the shim the golden fixtures were generated with over the unmodified reference (oracle/ref_shim.cpp), so
these read like tests of the reference's own entry points."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import golden
from helpers import REL, assert_topk_close, assert_topk_exact

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_f = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u32 = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_i32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_sz = C.c_size_t


@pytest.fixture(scope="module")
def dropin():
    lib = C.CDLL(os.path.join(ROOT, "gp-ann_b200", "libgpann_dropin.so"))
    lib.dropin_distance.restype = C.c_float
    lib.dropin_distance.argtypes = [_f, _f, C.c_uint]
    lib.dropin_compute_ground_truth.argtypes = [_f, _sz, _sz, _f, _sz, C.c_int, _f, _u32, _u32]
    lib.dropin_kth_distance.argtypes = [_f, _sz, _sz, _f, _sz, C.c_int, _f]
    lib.dropin_convert_ground_truth.argtypes = [_f, _sz, _sz, _f, _sz, C.c_int, _f, _u32, _f]
    lib.dropin_kmeans.restype = C.c_longlong
    lib.dropin_kmeans.argtypes = [_f, _sz, _sz, _f, _sz, _i32]
    lib.dropin_objective.restype = C.c_double
    lib.dropin_objective.argtypes = [_f, _sz, _sz, _f, _sz, _i32]
    lib.dropin_random_sample.argtypes = [_f, _sz, _sz, _sz, C.c_int, _f]
    lib.dropin_closest_leaders.argtypes = [_f, _sz, _sz, _f, _sz, C.c_uint32, C.c_int, _f, _u32, _u32]
    return lib


def _gt(lib, P, Q, k):
    P, Q = np.ascontiguousarray(P, np.float32), np.ascontiguousarray(Q, np.float32)
    d = np.full((len(Q), k), np.inf, np.float32)
    i = np.full((len(Q), k), 0xFFFFFFFF, np.uint32)
    c = np.zeros(len(Q), np.uint32)
    assert lib.dropin_compute_ground_truth(P, P.shape[0], P.shape[1], Q, len(Q), k, d, i, c) == 0
    return d, i, c


def test_compute_ground_truth_golden(dropin):
    for name, k in (("gt_u8_d128_k10", 10), ("gt_s8_d100_k10", 10), ("gt_u8_ties_d32_k10", 10)):
        g = golden(name)
        d, i, c = _gt(dropin, g["points"], g["queries"], k)
        assert_topk_exact(d, i, g["dist"], g["ids"], c, g["count"])
    for name, k in (("gt_f32_d100_k10", 10), ("gt_f32_d96_k100", 100), ("gt_f32_n_lt_k", 10)):
        g = golden(name)
        d, i, c = _gt(dropin, g["points"], g["queries"], k)
        assert np.array_equal(c, g["count"])
        assert_topk_close(d, i, g["dist"], g["ids"])
    g = golden("gt_f32_d100_k10")
    kth = np.empty(len(g["queries"]), np.float32)
    assert dropin.dropin_kth_distance(g["points"], *g["points"].shape, g["queries"], len(g["queries"]), 10, kth) == 0
    np.testing.assert_allclose(kth, g["kth"], rtol=REL)


def test_ground_truth_passes_the_self_check(dropin, oracle, capfd):
    """ConvertGroundTruthToDistanceToKthNeighbor (recall.h:74-137) must report 0 wrong distances / orders for
    GPU-produced ground truth on integer-valued data (SURVEY.md section 4 iii)."""
    rng = np.random.default_rng(21)
    P = rng.integers(0, 256, (30000, 128)).astype(np.float32)
    Q = rng.integers(0, 256, (100, 128)).astype(np.float32)
    d, i, c = _gt(dropin, P, Q, 10)
    kth = np.empty(100, np.float32)
    dropin.dropin_convert_ground_truth(P, *P.shape, Q, 100, 10, d, i, kth)
    out = capfd.readouterr().out
    assert "0 out of 1000 distances were wrong" in out
    assert "0 out of 100 neighbors lists were ordered incorrectly before recomputing distances. And 0 were" in out
    assert np.array_equal(kth, d[:, -1])
    d0, i0, _ = oracle.ground_truth(P, Q, 10)
    assert_topk_exact(d, i, d0, i0)


def test_scalar_distance_and_leaders(dropin, oracle):
    rng = np.random.default_rng(1)
    for d in (16, 20, 96, 100, 128):
        a, b = rng.standard_normal(d).astype(np.float32), rng.standard_normal(d).astype(np.float32)
        assert dropin.dropin_distance(a, b, d) == pytest.approx(oracle.distance(a, b), rel=1e-6)
    g = golden("closest_leaders")
    dist, ids, cnt = np.empty(5, np.float32), np.empty(5, np.uint32), np.zeros(1, np.uint32)
    dropin.dropin_closest_leaders(g["points"], *g["points"].shape, g["leaders"], len(g["leaders"]), int(g["my_id"]), 5, dist,
                                  ids, cnt)
    assert cnt[0] == 5 and np.array_equal(ids, g["ids"])
    np.testing.assert_allclose(dist, g["dist"], rtol=REL)


def test_kmeans_golden(dropin):
    g = golden("kmeans_d32_k24")
    P = g["points"]
    sample = np.empty_like(g["init"])
    dropin.dropin_random_sample(P, *P.shape, 24, 555, sample)
    assert np.array_equal(sample, g["init"])  # identical initial centroids (kmeans.cpp:207-226)
    cent = g["init"].copy()
    assign = np.empty(len(P), np.int32)
    kc = dropin.dropin_kmeans(P, *P.shape, cent, 24, assign)
    assert kc == g["centroids"].shape[0]
    assert (assign == g["assign"]).mean() > 0.999
    np.testing.assert_allclose(cent[:kc], g["centroids"], rtol=1e-4, atol=1e-4)
    obj = dropin.dropin_objective(P, *P.shape, cent[:kc].copy(), kc, assign)
    assert abs(obj - float(g["objective"])) <= 1e-4 * float(g["objective"])
    # empty clusters are compacted away, #centroids < 1 throws (-> -1 through the shim)
    g = golden("kmeans_empty_clusters")
    cent = g["init"].copy()
    assign = np.empty(len(g["points"]), np.int32)
    kc = dropin.dropin_kmeans(g["points"], *g["points"].shape, cent, len(cent), assign)
    assert kc == 6 and np.array_equal(assign, g["assign"])
    assert dropin.dropin_kmeans(g["points"], *g["points"].shape, cent, 0, assign) == -1
