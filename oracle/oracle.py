"""ctypes bindings for the CPU oracle and the compiled reference.  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg; never from
the product package (gp-ann_b200/).  `Oracle` wraps oracle/_build/libgpann_oracle.so (the restatement),
`Ref` wraps oracle/_ref/libgpann_ref[_mips].so (the unmodified reference, see oracle/ref_shim.cpp).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
L2, MIPS = 0, 1

_f = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u32 = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_i32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u64 = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_sz = C.c_size_t


def _f32(x):
    return np.ascontiguousarray(x, dtype=np.float32)


class Oracle:
    def __init__(self, path=None):
        path = path or os.path.join(HERE, "_build", "libgpann_oracle.so")
        if not os.path.exists(path):
            from . import build  # noqa
            build.build_oracle()
        self.lib = lib = C.CDLL(path)
        lib.oracle_distance.restype = C.c_float
        lib.oracle_distance.argtypes = [_f, _f, C.c_uint, C.c_int]
        lib.oracle_ground_truth.argtypes = [_f, _sz, _sz, _f, _sz, C.c_int, C.c_int, _f, _u32, _u32]
        lib.oracle_kth_distance.argtypes = [_f, _sz, _sz, _f, _sz, C.c_int, C.c_int, _f]
        lib.oracle_closest_leaders.argtypes = [_f, _sz, _f, _sz, C.c_uint32, C.c_int, _f, _u32, _u32]
        lib.oracle_nearest_centers.argtypes = [_f, _sz, _sz, _f, _sz, C.c_int, _i32, C.c_void_p]
        lib.oracle_aggregate_clusters.argtypes = [_f, _sz, _sz, _i32, _f, _sz, _u64, C.c_int]
        lib.oracle_remove_empty_clusters.restype = _sz
        lib.oracle_remove_empty_clusters.argtypes = [_f, _sz, _sz, _i32, _sz, _u64]
        lib.oracle_kmeans.restype = C.c_longlong
        lib.oracle_kmeans.argtypes = [_f, _sz, _sz, _f, _sz, C.c_int, _i32]
        lib.oracle_objective.restype = C.c_double
        lib.oracle_objective.argtypes = [_f, _sz, _sz, _f, _i32, C.c_int]
        lib.oracle_random_sample.argtypes = [_f, _sz, _sz, _sz, C.c_int, _i32, C.c_void_p]
        lib.oracle_ivf_query.argtypes = [_f, _sz, _i32, _u32, _f, C.c_int, _i32, _sz, C.c_int, _f, _u32, _u32]
        lib.oracle_crunch_bucket.argtypes = [_f, _sz, _u32, _sz, C.c_int, C.c_int, _f, _u32, _u32]

    def distance(self, p, q, metric=L2):
        p, q = _f32(p), _f32(q)
        return float(self.lib.oracle_distance(p, q, p.shape[0], metric))

    def ground_truth(self, points, queries, k, metric=L2):
        points, queries = _f32(points), _f32(queries)
        n, d = points.shape
        nq = queries.shape[0]
        dist = np.empty((nq, k), np.float32)
        ids = np.empty((nq, k), np.uint32)
        cnt = np.empty(nq, np.uint32)
        self.lib.oracle_ground_truth(points, n, d, queries, nq, k, metric, dist, ids, cnt)
        return dist, ids, cnt

    def kth_distance(self, points, queries, k, metric=L2):
        points, queries = _f32(points), _f32(queries)
        out = np.empty(queries.shape[0], np.float32)
        self.lib.oracle_kth_distance(points, points.shape[0], points.shape[1], queries, queries.shape[0], k, metric, out)
        return out

    def closest_leaders(self, points, leaders, my_id, k):
        points, leaders = _f32(points), _f32(leaders)
        dist = np.empty(k, np.float32)
        ids = np.empty(k, np.uint32)
        cnt = np.zeros(1, np.uint32)
        self.lib.oracle_closest_leaders(points, points.shape[1], leaders, leaders.shape[0], my_id, k, dist, ids, cnt)
        return dist[:cnt[0]], ids[:cnt[0]]

    def nearest_centers(self, points, centroids, metric=L2):
        points, centroids = _f32(points), _f32(centroids)
        out = np.empty(points.shape[0], np.int32)
        md = np.empty(points.shape[0], np.float32)
        self.lib.oracle_nearest_centers(points, points.shape[0], points.shape[1], centroids, centroids.shape[0], metric,
                                        out, md.ctypes.data)
        return out, md

    def aggregate(self, points, assign, kc, normalize=True):
        points = _f32(points)
        assign = np.ascontiguousarray(assign, np.int32)
        cent = np.zeros((kc, points.shape[1]), np.float32)
        sizes = np.zeros(kc, np.uint64)
        self.lib.oracle_aggregate_clusters(points, points.shape[0], points.shape[1], assign, cent, kc, sizes,
                                           int(normalize))
        return cent, sizes

    def kmeans(self, points, centroids, rounds=20):
        points = _f32(points)
        cent = _f32(centroids).copy()
        assign = np.empty(points.shape[0], np.int32)
        kc = self.lib.oracle_kmeans(points, points.shape[0], points.shape[1], cent, cent.shape[0], rounds, assign)
        if kc < 0:
            raise RuntimeError("KMeans #centroids < 1")
        return assign, cent[:kc].copy()

    def objective(self, points, centroids, assign, metric=L2):
        points, centroids = _f32(points), _f32(centroids)
        assign = np.ascontiguousarray(assign, np.int32)
        return float(self.lib.oracle_objective(points, points.shape[0], points.shape[1], centroids, assign, metric))

    def random_sample(self, points, num_samples, seed):
        points = _f32(points)
        idx = np.empty(num_samples, np.int32)
        rows = np.empty((num_samples, points.shape[1]), np.float32)
        self.lib.oracle_random_sample(points, points.shape[0], points.shape[1], num_samples, seed, idx, rows.ctypes.data)
        return idx, rows

    def ivf_query(self, clustered, offsets, permutation, q, k, buckets, metric=L2):
        clustered, q = _f32(clustered), _f32(q)
        offsets = np.ascontiguousarray(offsets, np.int32)
        permutation = np.ascontiguousarray(permutation, np.uint32)
        buckets = np.ascontiguousarray(buckets, np.int32)
        dist = np.empty(k, np.float32)
        ids = np.empty(k, np.uint32)
        cnt = np.zeros(1, np.uint32)
        self.lib.oracle_ivf_query(clustered, clustered.shape[1], offsets, permutation, q, k, buckets, len(buckets), metric,
                                  dist, ids, cnt)
        return dist[:cnt[0]], ids[:cnt[0]]


    def crunch_bucket(self, points, bucket, k, metric=L2):
        """Per-member top-k inside one bucket (global ids); returns (dist [m,k], ids [m,k], count [m])."""
        points = _f32(points)
        bucket = np.ascontiguousarray(bucket, np.uint32)
        m = len(bucket)
        dist = np.full((m, k), np.inf, np.float32)
        ids = np.full((m, k), 0xFFFFFFFF, np.uint32)
        cnt = np.zeros(m, np.uint32)
        self.lib.oracle_crunch_bucket(points, points.shape[1], bucket, m, k, metric, dist, ids, cnt)
        return dist, ids, cnt


def ref_path(metric=L2):
    return os.path.join(HERE, "_ref", "libgpann_ref.so" if metric == L2 else "libgpann_ref_mips.so")


def have_ref(metric=L2):
    return os.path.exists(ref_path(metric))


class Ref:
    """The unmodified reference (compiled from /root/reference by oracle/build.py)."""

    def __init__(self, metric=L2):
        self.metric = metric
        self.lib = lib = C.CDLL(ref_path(metric))
        assert lib.ref_metric() == metric
        lib.ref_distance.restype = C.c_float
        lib.ref_distance.argtypes = [_f, _f, C.c_uint]
        lib.ref_compute_ground_truth.argtypes = [_f, _sz, _sz, _f, _sz, C.c_int, _f, _u32, _u32]
        lib.ref_kth_distance.argtypes = [_f, _sz, _sz, _f, _sz, C.c_int, _f]
        lib.ref_convert_ground_truth.argtypes = [_f, _sz, _sz, _f, _sz, C.c_int, _f, _u32, _f]
        lib.ref_recall.restype = C.c_double
        lib.ref_recall.argtypes = [_f, _sz, C.c_int, _f, C.c_int]
        lib.ref_kmeans.restype = C.c_longlong
        lib.ref_kmeans.argtypes = [_f, _sz, _sz, _f, _sz, _i32]
        lib.ref_objective.restype = C.c_double
        lib.ref_objective.argtypes = [_f, _sz, _sz, _f, _sz, _i32]
        lib.ref_random_sample.argtypes = [_f, _sz, _sz, _sz, C.c_int, _f]
        lib.ref_closest_leaders.argtypes = [_f, _sz, _sz, _f, _sz, C.c_uint32, C.c_int, _f, _u32, _u32]
        lib.ref_ivf_query.argtypes = [_f, _sz, _sz, _i32, _u32, _sz, _f, C.c_int, _i32, _sz, _f, _u32, _u32]
        if hasattr(lib, "ref_balanced_kmeans"):
            lib.ref_balanced_kmeans.restype = C.c_longlong
            lib.ref_balanced_kmeans.argtypes = [_f, _sz, _sz, _f, _sz, _sz, _i32]
        if hasattr(lib, "ref_crunch_bucket"):
            lib.ref_crunch_bucket.argtypes = [_f, _sz, _sz, _u32, _sz, C.c_int, _f, _u32, _u32]
        lib.ref_pointset_create.restype = C.c_void_p
        lib.ref_pointset_create.argtypes = [_f, _sz, _sz]
        lib.ref_pointset_free.argtypes = [C.c_void_p]
        lib.ref_ground_truth_handles.restype = _sz
        lib.ref_ground_truth_handles.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        lib.ref_kmeans_handles.restype = C.c_void_p
        lib.ref_kmeans_handles.argtypes = [C.c_void_p, C.c_void_p]
        lib.ref_intvec_free.argtypes = [C.c_void_p]
        lib.ref_pointset_n.restype = _sz
        lib.ref_pointset_n.argtypes = [C.c_void_p]

    def num_workers(self):
        return int(self.lib.ref_num_workers())

    def distance(self, p, q):
        p, q = _f32(p), _f32(q)
        return float(self.lib.ref_distance(p, q, p.shape[0]))

    def ground_truth(self, points, queries, k):
        points, queries = _f32(points), _f32(queries)
        n, d = points.shape
        nq = queries.shape[0]
        dist = np.full((nq, k), np.inf, np.float32)
        ids = np.full((nq, k), 0xFFFFFFFF, np.uint32)
        cnt = np.empty(nq, np.uint32)
        self.lib.ref_compute_ground_truth(points, n, d, queries, nq, k, dist, ids, cnt)
        return dist, ids, cnt

    def kth_distance(self, points, queries, k):
        points, queries = _f32(points), _f32(queries)
        out = np.empty(queries.shape[0], np.float32)
        self.lib.ref_kth_distance(points, points.shape[0], points.shape[1], queries, queries.shape[0], k, out)
        return out

    def convert_ground_truth(self, points, queries, k, gt_dist, gt_id):
        """The reference's self-check (recall.h:74-137); prints its mismatch counters, returns kth distances."""
        points, queries = _f32(points), _f32(queries)
        out = np.empty(queries.shape[0], np.float32)
        self.lib.ref_convert_ground_truth(points, points.shape[0], points.shape[1], queries, queries.shape[0], k,
                                          _f32(gt_dist), np.ascontiguousarray(gt_id, np.uint32), out)
        return out

    def recall(self, nb_dist, kth, k):
        nb_dist = _f32(nb_dist)
        return float(self.lib.ref_recall(nb_dist, nb_dist.shape[0], nb_dist.shape[1], _f32(kth), k))

    def kmeans(self, points, centroids):
        points = _f32(points)
        cent = _f32(centroids).copy()
        assign = np.empty(points.shape[0], np.int32)
        kc = self.lib.ref_kmeans(points, points.shape[0], points.shape[1], cent, cent.shape[0], assign)
        if kc < 0:
            raise RuntimeError("KMeans #centroids < 1")
        return assign, cent[:kc].copy()

    def objective(self, points, centroids, assign):
        points, centroids = _f32(points), _f32(centroids)
        return float(self.lib.ref_objective(points, points.shape[0], points.shape[1], centroids, centroids.shape[0],
                                            np.ascontiguousarray(assign, np.int32)))

    def random_sample(self, points, num_samples, seed):
        points = _f32(points)
        rows = np.empty((num_samples, points.shape[1]), np.float32)
        self.lib.ref_random_sample(points, points.shape[0], points.shape[1], num_samples, seed, rows)
        return rows

    def closest_leaders(self, points, leaders, my_id, k):
        points, leaders = _f32(points), _f32(leaders)
        dist = np.empty(k, np.float32)
        ids = np.empty(k, np.uint32)
        cnt = np.zeros(1, np.uint32)
        self.lib.ref_closest_leaders(points, points.shape[0], points.shape[1], leaders, leaders.shape[0], my_id, k, dist,
                                     ids, cnt)
        return dist[:cnt[0]], ids[:cnt[0]]

    def balanced_kmeans(self, points, centroids, max_cluster_size):
        points = _f32(points)
        cent = _f32(centroids).copy()
        assign = np.empty(points.shape[0], np.int32)
        kc = self.lib.ref_balanced_kmeans(points, points.shape[0], points.shape[1], cent, cent.shape[0], int(max_cluster_size), assign)
        return assign, cent[:kc].copy()

    def crunch_bucket(self, points, bucket, k):
        points = _f32(points)
        bucket = np.ascontiguousarray(bucket, np.uint32)
        m = len(bucket)
        dist = np.full((m, k), np.inf, np.float32)
        ids = np.full((m, k), 0xFFFFFFFF, np.uint32)
        cnt = np.zeros(m, np.uint32)
        self.lib.ref_crunch_bucket(points, points.shape[0], points.shape[1], bucket, m, k, dist, ids, cnt)
        return dist, ids, cnt

    def ivf_query(self, points, clusters, q, k, buckets):
        points, q = _f32(points), _f32(q)
        offs = np.zeros(len(clusters) + 1, np.int32)
        offs[1:] = np.cumsum([len(c) for c in clusters])
        members = np.ascontiguousarray(np.concatenate([np.asarray(c, np.uint32) for c in clusters]), np.uint32)
        buckets = np.ascontiguousarray(buckets, np.int32)
        dist = np.empty(k, np.float32)
        ids = np.empty(k, np.uint32)
        cnt = np.zeros(1, np.uint32)
        self.lib.ref_ivf_query(points, points.shape[0], points.shape[1], offs, members, len(clusters), q, k, buckets,
                               len(buckets), dist, ids, cnt)
        return dist[:cnt[0]], ids[:cnt[0]]
