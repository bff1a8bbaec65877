"""Host-side mirror of the reference's entry points for the hot path, over the C ABI.

Names and argument meaning follow the reference (file:line relative to the reference repository):
  compute_ground_truth              ComputeGroundTruth            src/recall.h:25-39
  compute_distance_to_kth_neighbor  ComputeDistanceToKthNeighbor  src/recall.h:10-23
  kmeans                            KMeans                        src/kmeans.cpp:228-244 (mutates/shrinks centroids)
  remove_empty_clusters             RemoveEmptyClusters           src/kmeans.cpp:29-50  (host side, O(kc*d + n))
  objective_value                   ObjectiveValue                src/kmeans.cpp:246-249
  closest_leaders                   ClosestLeaders                src/defs.cpp:87-96
All distance work runs on the GPU; a missing CUDA device raises GpannError (no CPU fallback).
"""
import numpy as np

from .lib import L2, Context, GpannError, Points

NUM_ROUNDS = 20  # src/kmeans.cpp:238


def _as_points(ctx, points):
    if isinstance(points, Points):
        return points, False
    return ctx.upload(points), True


def compute_ground_truth(ctx, points, queries, k, metric=L2):
    """k lexicographically-smallest (distance, id) per query, ascending.  Returns (dist[nq,k], ids[nq,k], count[nq]);
    count = min(k, n) as in src/topn.h:21-29, unused slots hold +inf / 0xFFFFFFFF."""
    pts, own = _as_points(ctx, points)
    try:
        return ctx.topk(pts, queries, k, metric)
    finally:
        if own:
            pts.free()


def compute_distance_to_kth_neighbor(ctx, points, queries, k, metric=L2):
    pts, own = _as_points(ctx, points)
    try:
        return ctx.kth_distance(pts, queries, k, metric)
    finally:
        if own:
            pts.free()


def closest_leaders(ctx, points, leaders, my_id, k):
    """Top-k of points[my_id] against every leader (ids = leader index), ascending."""
    q = np.ascontiguousarray(points, dtype=np.float32)[my_id:my_id + 1]
    pts, own = _as_points(ctx, leaders)
    try:
        dist, ids, cnt = ctx.topk(pts, q, k, L2)
        return dist[0, :cnt[0]], ids[0, :cnt[0]]
    finally:
        if own:
            pts.free()


def remove_empty_clusters(centroids, assign, sizes):
    """Compact non-empty centroids to the front and remap assignments (src/kmeans.cpp:29-50)."""
    sizes = np.asarray(sizes)
    if not (sizes == 0).any():
        return centroids, assign
    keep = sizes != 0
    remap = np.full(len(sizes), -1, np.int32)
    remap[keep] = np.arange(int(keep.sum()), dtype=np.int32)
    return np.ascontiguousarray(centroids[keep]), remap[assign]


def kmeans(ctx, points, centroids, rounds=NUM_ROUNDS):
    """20 Lloyd rounds; returns (assignment of the LAST round against the pre-update centroids, remapped; centroids)."""
    centroids = np.ascontiguousarray(centroids, dtype=np.float32)
    if centroids.shape[0] < 1:
        raise GpannError("KMeans #centroids < 1")  # src/kmeans.cpp:229-231
    pts, own = _as_points(ctx, points)
    try:
        # one call: the centroids stay on the device for all rounds (gpann_kmeans_run); kmeans_by_steps below is the
        # round-by-round form with the reference's host-side RemoveEmptyClusters, kept for the parity tests
        assign, centroids, _ = ctx.kmeans_run(pts, centroids, rounds)
        return assign, centroids
    finally:
        if own:
            pts.free()


def kmeans_by_steps(ctx, points, centroids, rounds=NUM_ROUNDS):
    """KMeans as the reference structures it: per round one gpann_kmeans_step and the host-side empty-cluster compaction."""
    centroids = np.ascontiguousarray(centroids, dtype=np.float32)
    if centroids.shape[0] < 1:
        raise GpannError("KMeans #centroids < 1")
    pts, own = _as_points(ctx, points)
    try:
        assign = np.full(pts.n, -1, np.int32)
        for _ in range(rounds):
            assign, centroids, sizes = ctx.kmeans_step(pts, centroids)
            centroids, assign = remove_empty_clusters(centroids, assign, sizes)
        return assign, centroids
    finally:
        if own:
            pts.free()


def objective_value(ctx, points, centroids, closest_center, metric=L2):
    pts, own = _as_points(ctx, points)
    try:
        return ctx.objective(pts, centroids, closest_center, metric)
    finally:
        if own:
            pts.free()
