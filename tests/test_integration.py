"""INTEGRATION.md's recipe, executed: the reference's UNMODIFIED kmeans_tree_router.cpp / defs.cpp / dist.cpp and its own
headers, compiled with `-I gp-ann_b200/cxx` ahead of `-I <gp-ann>/src` (so recall.h is the drop-in's), linked with
cxx/kmeans_gpu.cpp and libgpann_cuda.so (gp-ann_b200/build.py: build_integration).  CPU: it compiles, links and exports the
entry points.  GPU: KMeansTreeRouter::Train -- which calls KMeans concurrently from parlay workers,
src/kmeans_tree_router.cpp:22-37 -- builds the same tree as the reference's CPU KMeans; ComputeGroundTruth and KMeans
through a gp-ann translation unit match the golden fixtures."""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest

from conftest import ROOT, golden
from helpers import assert_topk_exact

_f = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u32 = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_i32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_sz = C.c_size_t


def _build():
    spec = importlib.util.spec_from_file_location("gpann_build", os.path.join(ROOT, "gp-ann_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    return mod.build_integration()


def _load(path):
    lib = C.CDLL(path)
    lib.integ_router_train.argtypes = [_f, _sz, _sz, C.c_int, C.c_int, C.c_long, _f, _i32, _sz, C.POINTER(_sz)]
    lib.integ_ground_truth.argtypes = [_f, _sz, _sz, _f, _sz, C.c_int, _f, _u32]
    lib.integ_kmeans.argtypes = [_f, _sz, _sz, _f, _sz, _i32, C.POINTER(_sz)]
    return lib


def test_recipe_compiles_and_links():
    path = _build()
    if path is None:
        pytest.skip("needs /root/reference (or the prebuilt tests/_build/libgpann_integration.so)")
    lib = _load(path)
    for name in ("integ_router_train", "integ_ground_truth", "integ_kmeans"):
        assert hasattr(lib, name)


def _router(lib, P, shards, num_centroids, budget):
    n, d = P.shape
    cap = 100_000
    cent = np.zeros((cap, d), np.float32)
    shard = np.zeros(cap, np.int32)
    out_n = _sz(0)
    rc = lib.integ_router_train(P, n, d, shards, num_centroids, budget, cent, shard, cap, C.byref(out_n))
    assert rc == 0
    return cent[:out_n.value].copy(), shard[:out_n.value].copy()


@pytest.mark.gpu
def test_router_train_through_the_dropin():
    path = _build()
    if path is None:
        pytest.skip("integration library not built")
    lib = _load(path)
    rng = np.random.default_rng(17)
    d, shards = 32, 6
    centres = rng.standard_normal((60, d)).astype(np.float32) * 6
    P = np.ascontiguousarray((centres[rng.integers(0, 60, 48_000)] + rng.standard_normal((48_000, d))).astype(np.float32))
    cent, shard = _router(lib, P, shards, 16, 400)
    assert len(cent) > shards and set(shard.tolist()) == set(range(shards))
    assert np.isfinite(cent).all()
    # the reference's own CPU KMeans behind the same router code (oracle/_ref, test infrastructure)
    ref_path = os.path.join(ROOT, "oracle", "_ref", "libgpann_router_ref.so")
    if os.path.exists(ref_path):
        cent0, shard0 = _router(_load(ref_path), P, shards, 16, 400)
        assert cent.shape == cent0.shape and np.array_equal(shard, shard0)
        np.testing.assert_allclose(cent, cent0, rtol=1e-3, atol=1e-3)
    # ComputeGroundTruth / KMeans from a gp-ann translation unit
    g = golden("gt_u8_d128_k10")
    Pg, Qg = np.ascontiguousarray(g["points"], np.float32), np.ascontiguousarray(g["queries"], np.float32)
    dd = np.zeros((len(Qg), 10), np.float32)
    ii = np.zeros((len(Qg), 10), np.uint32)
    assert lib.integ_ground_truth(Pg, len(Pg), Pg.shape[1], Qg, len(Qg), 10, dd, ii) == 0
    assert_topk_exact(dd, ii, g["dist"], g["ids"])
    k = golden("kmeans_empty_clusters")
    Pk, Ck = np.ascontiguousarray(k["points"], np.float32), np.ascontiguousarray(k["init"], np.float32).copy()
    assign = np.zeros(len(Pk), np.int32)
    kc = _sz(0)
    assert lib.integ_kmeans(Pk, len(Pk), Pk.shape[1], Ck, len(Ck), assign, C.byref(kc)) == 0
    assert kc.value == 6 and np.array_equal(assign, k["assign"])
