"""Round-2 fixtures from the UNMODIFIED reference (oracle/_ref, built in place from /root/reference): batched
InvertedIndex::Query, ApproximateKNNGraphBuilder::CrunchBucket, ClosestLeaders for many points, and (MIPS build) KMeans.
Run in the authoring container:  python tests/golden/make_golden_r2.py   (inputs are stored beside the outputs)."""
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
    print(name, {k: np.shape(v) for k, v in arrays.items()}, os.path.getsize(path) // 1024, "KiB")


def main():
    build.build_ref(force=False)
    ref, ref_mips = Ref(), Ref(MIPS)
    rng = np.random.default_rng(20261018)

    # InvertedIndex::Query for a batch of queries (inverted_index.h:43-60).  Small alphabet => exact distance ties, so the
    # (distance, POSITION) tie-break and the permutation remap are both exercised.
    n, d, nb = 1500, 24, 12
    P = rng.integers(0, 3, (n, d)).astype(np.float32)
    part = rng.integers(0, nb, n)
    clusters = [np.nonzero(part == b)[0].astype(np.uint32) for b in range(nb)]
    rng.shuffle(clusters[3])  # member order inside a bucket is the caller's: positions, not ids, break ties
    Q = rng.integers(0, 3, (30, d)).astype(np.float32)
    probes = np.stack([rng.choice(nb, 4, replace=False) for _ in range(len(Q))]).astype(np.int32)
    probes[5, 2:] = -1  # ragged probe lists
    dist = np.full((len(Q), 10), np.inf, np.float32)
    ids = np.full((len(Q), 10), 0xFFFFFFFF, np.uint32)
    for q in range(len(Q)):
        b = [int(x) for x in probes[q] if x >= 0]
        dd, ii = ref.ivf_query(P, clusters, Q[q], 10, b)
        dist[q, :len(dd)], ids[q, :len(ii)] = dd, ii
    offsets = np.zeros(nb + 1, np.uint64)
    offsets[1:] = np.cumsum([len(c) for c in clusters])
    save("ivf_batch", points=P, offsets=offsets, permutation=np.concatenate(clusters), queries=Q, probes=probes, dist=dist, ids=ids)

    # CrunchBucket (knn_graph.h:165-185) for several buckets, float data with a scalar tail (d = 100) and a tiny bucket (m <= k)
    n, d, k = 2000, 100, 8
    P = rng.standard_normal((n, d)).astype(np.float32)
    sizes = [700, 5, 300, 1, 650]
    members = rng.permutation(n)[:sum(sizes)].astype(np.uint32)
    offsets = np.zeros(len(sizes) + 1, np.uint64)
    offsets[1:] = np.cumsum(sizes)
    dist = np.full((len(members), k), np.inf, np.float32)
    ids = np.full((len(members), k), 0xFFFFFFFF, np.uint32)
    cnt = np.zeros(len(members), np.uint32)
    for b in range(len(sizes)):
        lo, hi = int(offsets[b]), int(offsets[b + 1])
        dd, ii, cc = ref.crunch_bucket(P, members[lo:hi], k)
        dist[lo:hi], ids[lo:hi], cnt[lo:hi] = dd, ii, cc
    save("crunch_buckets", points=P, offsets=offsets, members=members, k=np.int64(k), dist=dist, ids=ids, count=cnt)

    # the same on byte-valued data with heavy ties: (distance, global id) order
    P8 = rng.integers(0, 4, (900, 32), dtype=np.uint8)
    members = rng.permutation(900)[:800].astype(np.uint32)
    offsets = np.array([0, 500, 800], np.uint64)
    dist = np.full((800, 6), np.inf, np.float32)
    ids = np.full((800, 6), 0xFFFFFFFF, np.uint32)
    cnt = np.zeros(800, np.uint32)
    for b in range(2):
        lo, hi = int(offsets[b]), int(offsets[b + 1])
        dist[lo:hi], ids[lo:hi], cnt[lo:hi] = ref.crunch_bucket(P8.astype(np.float32), members[lo:hi], 6)
    save("crunch_buckets_u8_ties", points=P8, offsets=offsets, members=members, k=np.int64(6), dist=dist, ids=ids, count=cnt)

    # ClosestLeaders for every point (knn_graph.h:75-81): n points x 950 leaders sampled from them, fanout 3
    n, d = 4000, 96
    P = rng.standard_normal((n, d)).astype(np.float32)
    lead = np.sort(rng.choice(n, 950, replace=False)).astype(np.uint32)
    L = P[lead]
    dist = np.zeros((n, 3), np.float32)
    ids = np.zeros((n, 3), np.uint32)
    for i in range(n):
        dist[i], ids[i] = ref.closest_leaders(P, L, i, 3)
    save("closest_leaders_batch", points=P, leader_ids=lead, dist=dist, ids=ids)

    # KMeans in the MIPS build (kmeans.cpp:61-101: norm-scaled sums, NormalizeCentroidsIP) -- 20 rounds
    centres = rng.standard_normal((16, 48)).astype(np.float32) * 3
    P = (centres[rng.integers(0, 16, 4000)] * rng.uniform(0.5, 1.5, (4000, 1)) + 0.3 * rng.standard_normal((4000, 48))).astype(np.float32)
    C0 = ref_mips.random_sample(P, 16, 555)
    assign, cent = ref_mips.kmeans(P, C0)
    obj = ref_mips.objective(P, cent, assign)
    save("kmeans_mips_d48_k16", points=P, init=C0, assign=assign, centroids=cent, objective=np.float64(obj))

    # BalancedKMeans (kmeans.cpp:251-524) on a skewed mixture: the reference's moving phase is racy, so the fixture records what
    # parity means here -- every cluster within the bound, and the objective the reference reached
    rng = np.random.default_rng(99)
    sizes = [2600, 900, 300, 120, 60, 20]
    centres = rng.standard_normal((len(sizes), 16)).astype(np.float32) * 5
    P = np.concatenate([centres[c] + rng.standard_normal((m, 16)).astype(np.float32) for c, m in enumerate(sizes)]).astype(np.float32)
    P = P[rng.permutation(len(P))]
    C0 = ref.random_sample(P, 8, 555)
    bound = 600
    assign, cent = ref.balanced_kmeans(P, C0, bound)
    obj = ref.objective(P, cent, assign)
    save("balanced_kmeans", points=P, init=C0, max_cluster_size=np.int64(bound), assign=assign, centroids=cent, objective=np.float64(obj),
         sizes=np.bincount(assign, minlength=len(cent)))


if __name__ == "__main__":
    main()
