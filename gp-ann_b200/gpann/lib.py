"""ctypes declarations for every symbol of include/gpann.h, and thin RAII wrappers (Context, Points)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

L2, MIPS = 0, 1
STORAGE_AUTO, STORAGE_F32, STORAGE_U8, STORAGE_S8 = 0, 1, 2, 3
ALGO_AUTO, ALGO_STREAM, ALGO_TENSOR = 0, 1, 2
INVALID_ID = 0xFFFFFFFF
EMPTY_KEY = 0xFFFFFFFFFFFFFFFF

# every exported symbol of include/gpann.h: name -> (restype, argtypes)
_vp, _sz, _i, _u32, _u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_uint64
SYMBOLS = {
    "gpann_init": (_i, [C.POINTER(_vp), _i]),
    "gpann_init_multi": (_i, [C.POINTER(_vp), C.POINTER(_i), _i]),
    "gpann_num_devices": (_i, [_vp]),
    "gpann_comm_unique_id": (_i, [_vp]),
    "gpann_comm_init_rank": (_i, [_vp, _i, _i, _vp]),
    "gpann_comm_size": (_i, [_vp]),
    "gpann_comm_rank": (_i, [_vp]),
    "gpann_destroy": (None, [_vp]),
    "gpann_last_error": (C.c_char_p, [_vp]),
    "gpann_device": (_i, [_vp]),
    "gpann_set_algo": (_i, [_vp, _i]),
    "gpann_launch_count": (_u64, [_vp]),
    "gpann_last_algo": (C.c_char_p, [_vp]),
    "gpann_set_timing": (_i, [_vp, _i]),
    "gpann_last_kernel_ms": (C.c_float, [_vp]),
    "gpann_last_ambiguous": (_u64, [_vp]),
    "gpann_points_upload": (_i, [_vp, _vp, _sz, _sz, _i, _u32, C.POINTER(_vp)]),
    "gpann_points_upload_bytes": (_i, [_vp, _vp, _sz, _sz, _i, _u32, C.POINTER(_vp)]),
    "gpann_points_from_device": (_i, [_vp, _vp, _sz, _sz, _i, _u32, _vp, C.POINTER(_vp)]),
    "gpann_points_free": (None, [_vp]),
    "gpann_points_n": (_sz, [_vp]),
    "gpann_points_d": (_sz, [_vp]),
    "gpann_points_storage": (_i, [_vp]),
    "gpann_topk": (_i, [_vp, _vp, _vp, _sz, _i, _i, _vp, _vp, _vp]),
    "gpann_kth_distance": (_i, [_vp, _vp, _vp, _sz, _i, _i, _vp]),
    "gpann_topk_device": (_i, [_vp, _vp, _vp, _sz, _i, _i, _vp, _vp]),
    "gpann_topk_sharded_device": (_i, [_vp, _vp, _vp, _sz, _i, _i, _vp, _vp, _vp, _vp]),
    "gpann_topk_stream_bytes": (_i, [_vp, _vp, _sz, _sz, _i, _u32, _vp, _sz, _i, _sz, _vp, _vp, _vp]),
    "gpann_topk_stream": (_i, [_vp, _vp, _sz, _sz, _i, _u32, _vp, _sz, _i, _sz, _vp, _vp, _vp, _vp]),
    "gpann_topk_stream_bytes_device": (_i, [_vp, _vp, _sz, _sz, _i, _u32, _vp, _sz, _i, _sz, _vp, _vp]),
    "gpann_topk_range": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, _i, _i, _vp, _vp, _vp]),
    "gpann_ivf_topk": (_i, [_vp, _vp, _vp, _sz, _vp, _vp, _sz, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "gpann_knn_buckets": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _vp, _vp, _vp]),
    "gpann_closest_leaders_batch": (_i, [_vp, _vp, _vp, _sz, _vp, _i, _i, _vp, _vp, _vp]),
    "gpann_group_min_dist": (_i, [_vp, _vp, _vp, _i, _vp, _sz, _i, _vp]),
    "gpann_merge_keys_device": (_i, [_vp, _vp, _sz, _sz, _i, _vp, _vp, _vp, _vp]),
    "gpann_pack_key": (_u64, [C.c_float, _u32]),
    "gpann_unpack_key": (None, [_u64, C.POINTER(C.c_float), C.POINTER(_u32)]),
    "gpann_kmeans_assign": (_i, [_vp, _vp, _vp, _sz, _i, _vp, _vp]),
    "gpann_kmeans_step": (_i, [_vp, _vp, _vp, _sz, _i, _vp, _vp]),
    "gpann_kmeans_partial_device": (_i, [_vp, _vp, _vp, _sz, _i, _vp, _vp, _vp, _vp]),
    "gpann_kmeans_run": (_i, [_vp, _vp, _vp, _sz, _i, _i, _vp, _vp, C.POINTER(_sz)]),
    "gpann_kmeans_run_device": (_i, [_vp, _vp, _vp, _sz, _i, _i, _vp, _vp, _vp, _vp]),
    "gpann_balanced_kmeans": (_i, [_vp, _vp, _vp, _sz, _sz, _i, _vp, _vp, C.POINTER(_sz)]),
    "gpann_objective": (_i, [_vp, _vp, _vp, _sz, _vp, _i, C.POINTER(C.c_double)]),
}


class GpannError(RuntimeError):
    """Raised for every non-zero return of the C ABI (the reference throws std::runtime_error, kmeans.cpp:229)."""


def lib_path():
    return os.environ.get("GPANN_LIB", os.path.join(os.path.dirname(HERE), "libgpann_cuda.so"))


_lib = None


def load_library():
    """Load libgpann_cuda.so and declare all prototypes.  Loading needs no GPU; gpann_init does."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise GpannError(f"{path} is missing: build it with `python gp-ann_b200/build.py` (there is no CPU fallback)")
        lib = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def pack_key(dist, idx):
    return int(load_library().gpann_pack_key(float(dist), int(idx)))


def unpack_keys(keys):
    """Vectorised inverse of the key packing: uint64 array -> (float32 distances, uint32 ids)."""
    keys = np.asarray(keys, dtype=np.uint64)
    o = (keys >> np.uint64(32)).astype(np.uint32)
    neg = (o & np.uint32(0x80000000)) == 0
    bits = np.where(neg, ~o, o & np.uint32(0x7FFFFFFF)).astype(np.uint32)
    return bits.view(np.float32), (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def comm_unique_id():
    """128 bytes identifying a new NCCL communicator (rank 0 creates it, the caller ships it to the other ranks)."""
    buf = C.create_string_buffer(128)
    rc = load_library().gpann_comm_unique_id(buf)
    if rc != 0:
        raise GpannError(f"gpann_comm_unique_id failed with code {rc} (libnccl.so.2 not loadable?)")
    return buf.raw


class Context:
    """One gpann_ctx: one CUDA device (`device` an int) or a GROUP over several devices of this process (`device` a
    list: gpann_init_multi -- point sets are sharded by rows, exchanges are NCCL collectives inside the library).
    Externally synchronised: one call at a time."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            devs = (C.c_int * len(device))(*[int(x) for x in device])
            rc = self.lib.gpann_init_multi(C.byref(h), devs, len(device))
        else:
            rc = self.lib.gpann_init(C.byref(h), int(device))
        if rc != 0:
            raise GpannError(f"gpann_init(device={device}) failed with code {rc}: no usable sm_100 CUDA device "
                             "(there is no CPU fallback)")
        self.h = h
        self.device = device

    @property
    def num_devices(self):
        return int(self.lib.gpann_num_devices(self.h))

    def comm_init(self, nranks, rank, unique_id):
        """Join an NCCL communicator of `nranks` single-device contexts (one per process).  Collective."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self.check(self.lib.gpann_comm_init_rank(self.h, int(nranks), int(rank), buf))

    @property
    def comm_size(self):
        return int(self.lib.gpann_comm_size(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.gpann_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def check(self, rc):
        if rc != 0:
            raise GpannError(f"gpann error {rc}: {self.lib.gpann_last_error(self.h).decode()}")

    def set_algo(self, algo):
        self.check(self.lib.gpann_set_algo(self.h, algo))

    @property
    def last_algo(self):
        return self.lib.gpann_last_algo(self.h).decode()

    def set_timing(self, on):
        self.check(self.lib.gpann_set_timing(self.h, int(on)))

    def last_kernel_ms(self):
        """Duration of the last bracketed dominant kernel (waits for it); needs set_timing(True)."""
        return float(self.lib.gpann_last_kernel_ms(self.h))

    @property
    def last_ambiguous(self):
        """Points the last tensor-path k-means assignment had to re-rank exactly (diagnostic)."""
        return int(self.lib.gpann_last_ambiguous(self.h))

    @property
    def launch_count(self):
        return int(self.lib.gpann_launch_count(self.h))

    # ---- point sets ----
    def upload(self, rows, storage=STORAGE_AUTO, id_base=0):
        rows = np.ascontiguousarray(rows)
        if rows.dtype in (np.uint8, np.int8):
            return self.upload_bytes(rows, id_base)
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        n, d = rows.shape
        h = C.c_void_p()
        self.check(self.lib.gpann_points_upload(self.h, _ptr(rows), n, d, storage, id_base, C.byref(h)))
        return Points(self, h)

    def upload_bytes(self, rows, id_base=0):
        rows = np.ascontiguousarray(rows)
        assert rows.dtype in (np.uint8, np.int8)
        n, d = rows.shape
        h = C.c_void_p()
        self.check(self.lib.gpann_points_upload_bytes(self.h, _ptr(rows), n, d, int(rows.dtype == np.int8), id_base,
                                                      C.byref(h)))
        return Points(self, h)

    def from_device(self, dev_ptr, n, d, storage, id_base=0, stream=None, keepalive=None):
        h = C.c_void_p()
        self.check(self.lib.gpann_points_from_device(self.h, C.c_void_p(dev_ptr), n, d, storage, id_base,
                                                     C.c_void_p(stream or 0), C.byref(h)))
        p = Points(self, h)
        p.keepalive = keepalive  # the rows are borrowed
        return p

    # ---- top-k ----
    @staticmethod
    def _rows(a, d, what):
        a = np.ascontiguousarray(a, dtype=np.float32)
        if a.ndim != 2 or a.shape[1] != d:
            raise GpannError(f"{what}: expected rows of {d} floats, got shape {a.shape}")
        return a

    def topk(self, pts, queries, k, metric=L2, row_range=None, want_dist=True, want_count=True):
        q = self._rows(queries, pts.d, "topk queries")
        nq = q.shape[0]
        dist = np.empty((nq, k), np.float32) if want_dist else None
        ids = np.empty((nq, k), np.uint32)
        cnt = np.empty(nq, np.uint32) if want_count else None
        if row_range is None:
            rc = self.lib.gpann_topk(self.h, pts.h, _ptr(q), nq, k, metric, _ptr(dist), _ptr(ids), _ptr(cnt))
        else:
            rc = self.lib.gpann_topk_range(self.h, pts.h, row_range[0], row_range[1], _ptr(q), nq, k, metric, _ptr(dist),
                                           _ptr(ids), _ptr(cnt))
        self.check(rc)
        return dist, ids, cnt

    def topk_stream_bytes(self, host_rows, queries, k, id_base=0, chunk_rows=0):
        """Exact top-k against a host-resident uint8/int8 array streamed through the device in chunks."""
        rows = np.ascontiguousarray(host_rows)
        assert rows.dtype in (np.uint8, np.int8)
        q = np.ascontiguousarray(queries, dtype=np.float32)
        nq = q.shape[0]
        dist, ids, cnt = np.empty((nq, k), np.float32), np.empty((nq, k), np.uint32), np.empty(nq, np.uint32)
        self.check(self.lib.gpann_topk_stream_bytes(self.h, _ptr(rows), rows.shape[0], rows.shape[1], int(rows.dtype == np.int8),
                                                    id_base, _ptr(q), nq, k, chunk_rows, _ptr(dist), _ptr(ids), _ptr(cnt)))
        return dist, ids, cnt

    def topk_stream(self, host_rows, queries, k, storage=STORAGE_AUTO, id_base=0, chunk_rows=0, want_kth=False):
        """Exact top-k against host-resident float rows streamed through the device in chunks."""
        rows = np.ascontiguousarray(host_rows, dtype=np.float32)
        q = np.ascontiguousarray(queries, dtype=np.float32)
        nq = q.shape[0]
        dist, ids, cnt = np.empty((nq, k), np.float32), np.empty((nq, k), np.uint32), np.empty(nq, np.uint32)
        kth = np.empty(nq, np.float32) if want_kth else None
        self.check(self.lib.gpann_topk_stream(self.h, _ptr(rows), rows.shape[0], rows.shape[1], storage, id_base, _ptr(q), nq, k,
                                              chunk_rows, _ptr(dist), _ptr(ids), _ptr(cnt), _ptr(kth)))
        return (dist, ids, cnt, kth) if want_kth else (dist, ids, cnt)

    def topk_stream_bytes_device(self, host_ptr, n, d, is_signed, dev_queries, nq, k, dev_out_keys, id_base=0, chunk_rows=0,
                                 stream=None):
        self.check(self.lib.gpann_topk_stream_bytes_device(self.h, C.c_void_p(host_ptr), n, d, int(is_signed), id_base,
                                                           C.c_void_p(dev_queries), nq, k, chunk_rows, C.c_void_p(dev_out_keys),
                                                           C.c_void_p(stream or 0)))

    # ---- segmented scans: batched IVF probes, per-bucket all-pairs, leaders for every point ----
    def ivf_topk(self, pts, bucket_offsets, queries, probes, k, permutation=None, metric=L2):
        """InvertedIndex::Query for a batch: probes [nq, nprobe] bucket ids per query (negative = unused)."""
        q = self._rows(queries, pts.d, "ivf_topk queries")
        off = np.ascontiguousarray(bucket_offsets, dtype=np.uint64)
        pr = np.ascontiguousarray(probes, dtype=np.int32)
        perm = None if permutation is None else np.ascontiguousarray(permutation, dtype=np.uint32)
        nq = q.shape[0]
        dist, ids, cnt = np.empty((nq, k), np.float32), np.empty((nq, k), np.uint32), np.empty(nq, np.uint32)
        self.check(self.lib.gpann_ivf_topk(self.h, pts.h, _ptr(off), len(off) - 1, _ptr(perm), _ptr(q), nq, _ptr(pr), pr.shape[1], k, metric,
                                           _ptr(dist), _ptr(ids), _ptr(cnt)))
        return dist, ids, cnt

    def knn_buckets(self, pts, bucket_offsets, bucket_members, k, metric=L2):
        """CrunchBucket for all buckets: per member position the k nearest other members of its bucket (global ids)."""
        off = np.ascontiguousarray(bucket_offsets, dtype=np.uint64)
        mem = np.ascontiguousarray(bucket_members, dtype=np.uint32)
        total = len(mem)
        dist, ids, cnt = np.empty((total, k), np.float32), np.empty((total, k), np.uint32), np.empty(total, np.uint32)
        self.check(self.lib.gpann_knn_buckets(self.h, pts.h, _ptr(off), _ptr(mem), len(off) - 1, k, metric, _ptr(dist), _ptr(ids), _ptr(cnt)))
        return dist, ids, cnt

    def closest_leaders_batch(self, pts, leaders, k, point_ids=None, metric=L2):
        """ClosestLeaders for every listed point (all points when point_ids is None): ids = leader index."""
        pid = None if point_ids is None else np.ascontiguousarray(point_ids, dtype=np.uint32)
        m = pts.n if pid is None else len(pid)
        dist, ids, cnt = np.empty((m, k), np.float32), np.empty((m, k), np.uint32), np.empty(m, np.uint32)
        self.check(self.lib.gpann_closest_leaders_batch(self.h, pts.h, _ptr(pid), m, leaders.h, k, metric, _ptr(dist), _ptr(ids), _ptr(cnt)))
        return dist, ids, cnt

    def group_min_dist(self, pts, group, ngroups, queries, metric=L2):
        """out[q, g] = min distance from query q to the points of group g (FLT_MAX for an empty group)."""
        q = self._rows(queries, pts.d, "group_min_dist queries")
        grp = np.ascontiguousarray(group, dtype=np.int32)
        if grp.shape != (pts.n,):
            raise GpannError(f"group_min_dist: expected {pts.n} group ids, got shape {grp.shape}")
        out = np.empty((q.shape[0], ngroups), np.float32)
        self.check(self.lib.gpann_group_min_dist(self.h, pts.h, _ptr(grp), int(ngroups), _ptr(q), q.shape[0], metric, _ptr(out)))
        return out

    def kth_distance(self, pts, queries, k, metric=L2):
        q = self._rows(queries, pts.d, "kth_distance queries")
        out = np.empty(q.shape[0], np.float32)
        self.check(self.lib.gpann_kth_distance(self.h, pts.h, _ptr(q), q.shape[0], k, metric, _ptr(out)))
        return out

    def topk_device(self, pts, dev_queries, nq, k, dev_out_keys, metric=L2, stream=None):
        self.check(self.lib.gpann_topk_device(self.h, pts.h, C.c_void_p(dev_queries), nq, k, metric,
                                              C.c_void_p(dev_out_keys), C.c_void_p(stream or 0)))

    def merge_keys_device(self, dev_keys, parts, nq, k, dev_out_dist=None, dev_out_id=None, dev_out_keys=None, stream=None):
        self.check(self.lib.gpann_merge_keys_device(self.h, C.c_void_p(dev_keys), parts, nq, k,
                                                    C.c_void_p(dev_out_dist or 0), C.c_void_p(dev_out_id or 0),
                                                    C.c_void_p(dev_out_keys or 0), C.c_void_p(stream or 0)))

    # ---- k-means ----
    def kmeans_assign(self, pts, centroids, metric=L2, want_mindist=False):
        c = self._rows(centroids, pts.d, "kmeans_assign centroids")
        out = np.empty(pts.n, np.int32)
        md = np.empty(pts.n, np.float32) if want_mindist else None
        self.check(self.lib.gpann_kmeans_assign(self.h, pts.h, _ptr(c), c.shape[0], metric, _ptr(out), _ptr(md)))
        return (out, md) if want_mindist else out

    def kmeans_step(self, pts, centroids, metric=L2):
        """One Lloyd round.  Returns (assignment, new centroids incl. all-zero rows of empty clusters, sizes)."""
        c = self._rows(centroids, pts.d, "kmeans_step centroids").copy()
        assign = np.empty(pts.n, np.int32)
        sizes = np.empty(c.shape[0], np.uint64)
        self.check(self.lib.gpann_kmeans_step(self.h, pts.h, _ptr(c), c.shape[0], metric, _ptr(assign), _ptr(sizes)))
        return assign, c, sizes

    def kmeans_run(self, pts, centroids, rounds=20, metric=L2, want_assign=True):
        """KMeans (src/kmeans.cpp:228-244) in one call, centroids resident on the device for all rounds.  Returns
        (assignment of the last round, renumbered; surviving centroids; their sizes)."""
        c = self._rows(centroids, pts.d, "kmeans_run centroids").copy()
        assign = np.empty(pts.n, np.int32) if want_assign else None
        sizes = np.empty(c.shape[0], np.uint64)
        kc = C.c_size_t(0)
        self.check(self.lib.gpann_kmeans_run(self.h, pts.h, _ptr(c), c.shape[0], int(rounds), metric, _ptr(assign), _ptr(sizes),
                                             C.byref(kc)))
        return assign, np.ascontiguousarray(c[:kc.value]), sizes[:kc.value]

    def balanced_kmeans(self, pts, centroids, max_cluster_size, metric=L2):
        """BalancedKMeans (src/kmeans.cpp:251-524).  Returns (assignment, centroids, sizes) with every size <= max_cluster_size."""
        c = self._rows(centroids, pts.d, "balanced_kmeans centroids").copy()
        assign = np.empty(pts.n, np.int32)
        sizes = np.empty(c.shape[0], np.uint64)
        kc = C.c_size_t(0)
        self.check(self.lib.gpann_balanced_kmeans(self.h, pts.h, _ptr(c), c.shape[0], int(max_cluster_size), metric, _ptr(assign), _ptr(sizes),
                                                  C.byref(kc)))
        return assign, np.ascontiguousarray(c[:kc.value]), sizes[:kc.value]

    def kmeans_run_device(self, pts, dev_centroids, kc, rounds, dev_alive, dev_counts, dev_assign=None, metric=L2, stream=None):
        self.check(self.lib.gpann_kmeans_run_device(self.h, pts.h, C.c_void_p(dev_centroids), kc, int(rounds), metric,
                                                    C.c_void_p(dev_assign or 0), C.c_void_p(dev_alive), C.c_void_p(dev_counts),
                                                    C.c_void_p(stream or 0)))

    def topk_sharded_device(self, pts, dev_queries, nq, k, dev_out_dist=None, dev_out_id=None, dev_out_keys=None, metric=L2,
                            stream=None):
        """The sharded step in one call: local scan + bounds all-reduce(min) + key all-gather + merge (collective)."""
        self.check(self.lib.gpann_topk_sharded_device(self.h, pts.h, C.c_void_p(dev_queries), nq, k, metric,
                                                      C.c_void_p(dev_out_dist or 0), C.c_void_p(dev_out_id or 0),
                                                      C.c_void_p(dev_out_keys or 0), C.c_void_p(stream or 0)))

    def kmeans_partial_device(self, pts, dev_centroids, kc, dev_sums, dev_counts, dev_assign=None, metric=L2, stream=None):
        self.check(self.lib.gpann_kmeans_partial_device(self.h, pts.h, C.c_void_p(dev_centroids), kc, metric,
                                                        C.c_void_p(dev_assign or 0), C.c_void_p(dev_sums),
                                                        C.c_void_p(dev_counts), C.c_void_p(stream or 0)))

    def objective(self, pts, centroids, assign, metric=L2):
        c = self._rows(centroids, pts.d, "objective centroids")
        a = np.ascontiguousarray(assign, dtype=np.int32)
        if a.shape != (pts.n,):
            raise GpannError(f"objective: expected {pts.n} assignments, got shape {a.shape}")
        out = C.c_double()
        self.check(self.lib.gpann_objective(self.h, pts.h, _ptr(c), c.shape[0], _ptr(a), metric, C.byref(out)))
        return out.value


class Points:
    """A device-resident point set (gpann_points)."""

    def __init__(self, ctx, h):
        self.ctx, self.h, self.keepalive = ctx, h, None

    @property
    def n(self):
        return int(self.ctx.lib.gpann_points_n(self.h))

    @property
    def d(self):
        return int(self.ctx.lib.gpann_points_d(self.h))

    @property
    def storage(self):
        return int(self.ctx.lib.gpann_points_storage(self.h))

    def free(self):
        if self.h:
            self.ctx.lib.gpann_points_free(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.ctx.h:
                self.free()
        except Exception:
            pass
