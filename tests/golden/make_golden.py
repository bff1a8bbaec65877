"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libgpann_ref*.so, built in place from
/root/reference by oracle/build.py).  Run in the authoring container:  python tests/golden/make_golden.py

The reference ships no golden vectors for this path, so these fixtures are outputs of the
reference's own ComputeGroundTruth / ComputeDistanceToKthNeighbor / KMeans / ObjectiveValue / RandomSample /
ClosestLeaders / InvertedIndex::Query on seeded inputs.  Inputs are stored beside the outputs, so the fixtures do
not depend on numpy's generator staying stable.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import build  # noqa: E402
from oracle.oracle import MIPS, Ref  # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(name, {k: v.shape for k, v in arrays.items()}, os.path.getsize(path) // 1024, "KiB")


def main():
    build.build_ref()
    ref, ref_mips = Ref(), Ref(MIPS)
    rng = np.random.default_rng(20240917)

    # BigANN-shape bytes, d=128: bit-exact (dist,id) lists
    P = rng.integers(0, 256, (3000, 128), dtype=np.uint8)
    Q = rng.integers(0, 256, (40, 128), dtype=np.uint8)
    d, i, c = ref.ground_truth(P.astype(np.float32), Q.astype(np.float32), 10)
    save("gt_u8_d128_k10", points=P, queries=Q, dist=d, ids=i, count=c)

    # int8 (.i8bin), d=100 -> K padding in the tensor layout and a 4-element scalar tail in the reference
    P = rng.integers(-128, 128, (2500, 100), dtype=np.int8)
    Q = rng.integers(-128, 128, (33, 100), dtype=np.int8)
    d, i, c = ref.ground_truth(P.astype(np.float32), Q.astype(np.float32), 10)
    save("gt_s8_d100_k10", points=P, queries=Q, dist=d, ids=i, count=c)

    # heavy exact ties: tiny alphabet, small d; the (dist,id) order must break ties on the smaller id
    P = rng.integers(0, 3, (600, 32), dtype=np.uint8)
    Q = rng.integers(0, 3, (20, 32), dtype=np.uint8)
    d, i, c = ref.ground_truth(P.astype(np.float32), Q.astype(np.float32), 10)
    save("gt_u8_ties_d32_k10", points=P, queries=Q, dist=d, ids=i, count=c)

    # float data, d=100 (tail), k=10, plus ComputeDistanceToKthNeighbor
    P = rng.standard_normal((1200, 100)).astype(np.float32)
    Q = rng.standard_normal((24, 100)).astype(np.float32)
    d, i, c = ref.ground_truth(P, Q, 10)
    kth = ref.kth_distance(P, Q, 10)
    save("gt_f32_d100_k10", points=P, queries=Q, dist=d, ids=i, count=c, kth=kth)

    # DEEP-shape float, d=96, k=100
    P = rng.standard_normal((1000, 96)).astype(np.float32)
    Q = rng.standard_normal((12, 96)).astype(np.float32)
    d, i, c = ref.ground_truth(P, Q, 100)
    save("gt_f32_d96_k100", points=P, queries=Q, dist=d, ids=i, count=c)

    # n < k: the reference returns n entries (topn.h:21-29)
    P = rng.standard_normal((7, 16)).astype(np.float32)
    Q = rng.standard_normal((3, 16)).astype(np.float32)
    d, i, c = ref.ground_truth(P, Q, 10)
    save("gt_f32_n_lt_k", points=P, queries=Q, dist=d, ids=i, count=c)

    # MIPS build (-DMIPS_DISTANCE): distance = 1 - <p,q>
    P = rng.standard_normal((900, 64)).astype(np.float32)
    Q = rng.standard_normal((16, 64)).astype(np.float32)
    d, i, c = ref_mips.ground_truth(P, Q, 5)
    save("gt_mips_d64_k5", points=P, queries=Q, dist=d, ids=i, count=c)

    # k-means: mixture of Gaussians, RandomSample(…, 24, 555) as in partitioning.cpp:19, 20 rounds
    centres = rng.standard_normal((24, 32)).astype(np.float32) * 4
    P = (centres[rng.integers(0, 24, 3000)] + rng.standard_normal((3000, 32))).astype(np.float32)
    C0 = ref.random_sample(P, 24, 555)
    assign, cent = ref.kmeans(P, C0)
    obj = ref.objective(P, cent, assign)
    save("kmeans_d32_k24", points=P, init=C0, assign=assign, centroids=cent, objective=np.float64(obj))

    # k-means with duplicate initial centroids -> empty clusters get removed (kmeans.cpp:29-50)
    P2 = P[:800]
    C0 = np.concatenate([P2[:6], P2[:3]]).astype(np.float32)  # 3 duplicates lose every tie
    assign, cent = ref.kmeans(P2, C0)
    save("kmeans_empty_clusters", points=P2, init=C0, assign=assign, centroids=cent)

    # ClosestLeaders (defs.cpp:87-96)
    L = P[100:160]
    d, i = ref.closest_leaders(P, L, 123, 5)
    save("closest_leaders", points=P[:200], leaders=L, my_id=np.int64(123), dist=d, ids=i)

    # InvertedIndex::Query (inverted_index.h:43-60)
    n = 500
    Pi = rng.standard_normal((n, 24)).astype(np.float32)
    part = rng.integers(0, 8, n)
    clusters = [np.nonzero(part == b)[0].astype(np.uint32) for b in range(8)]
    q = rng.standard_normal(24).astype(np.float32)
    d, i = ref.ivf_query(Pi, clusters, q, 10, [5, 1, 6])
    save("ivf_query", points=Pi, partition=part.astype(np.int32), query=q, buckets=np.array([5, 1, 6], np.int32), dist=d,
         ids=i)


if __name__ == "__main__":
    main()
