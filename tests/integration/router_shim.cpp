// router_shim.cpp -- TEST INFRASTRUCTURE.  A gp-ann translation unit as a maintainer would build it after following
// compiled with `-I gp-ann_b200/cxx -I <gp-ann>/src`, so "recall.h" is the drop-in header while
// "defs.h", "topn.h", "dist.h", "kmeans.h" and "kmeans_tree_router.h" are the REFERENCE's own; linked with the reference's
// unmodified kmeans_tree_router.cpp / defs.cpp / dist.cpp and the drop-in's kmeans_gpu.cpp + libgpann_cuda.so.
// tests/test_integration.py builds it (CPU test: the recipe compiles and links) and runs it on the GPU box:
// KMeansTreeRouter::Train calls KMeans concurrently from parlay workers (src/kmeans_tree_router.cpp:22-37).
#include <cstring>
#include <iostream>

#include "recall.h"
#include "kmeans.h"
#include "kmeans_tree_router.h"

static PointSet make_pointset(const float* x, size_t n, size_t d) {
    PointSet ps;
    ps.n = n;
    ps.d = d;
    ps.coordinates.assign(x, x + n * d);
    return ps;
}

extern "C" {

// Train a router over `num_shards` contiguous shards; returns the tree's centroids (ExtractPoints) and their shard ids.
int integ_router_train(const float* points, size_t n, size_t d, int num_shards, int num_centroids, long budget, float* out_centroids,
                       int* out_shard, size_t cap, size_t* out_n) {
    try {
        PointSet P = make_pointset(points, n, d);
        Clusters clusters(num_shards);
        for (size_t i = 0; i < n; ++i) clusters[i * num_shards / n].push_back((uint32_t) i);
        KMeansTreeRouterOptions opt;
        opt.num_centroids = num_centroids;
        opt.budget = budget;
        KMeansTreeRouter router;
        router.Train(P, clusters, opt);
        auto [cp, part] = router.ExtractPoints();
        *out_n = cp.n;
        if (cp.n > cap) return -2;
        std::memcpy(out_centroids, cp.coordinates.data(), cp.n * d * sizeof(float));
        std::memcpy(out_shard, part.data(), cp.n * sizeof(int));
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "integ_router_train: " << e.what() << std::endl;
        return -1;
    }
}

int integ_ground_truth(const float* points, size_t n, size_t d, const float* queries, size_t nq, int k, float* out_dist, uint32_t* out_id) {
    try {
        PointSet P = make_pointset(points, n, d), Q = make_pointset(queries, nq, d);
        std::vector<NNVec> gt = ComputeGroundTruth(P, Q, k);
        for (size_t q = 0; q < nq; ++q)
            for (size_t r = 0; r < gt[q].size(); ++r) {
                out_dist[q * k + r] = gt[q][r].first;
                out_id[q * k + r] = gt[q][r].second;
            }
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "integ_ground_truth: " << e.what() << std::endl;
        return -1;
    }
}

int integ_kmeans(const float* points, size_t n, size_t d, float* centroids_inout, size_t kc, int* out_assign, size_t* out_kc) {
    try {
        PointSet P = make_pointset(points, n, d), C = make_pointset(centroids_inout, kc, d);
        std::vector<int> a = KMeans(P, C);
        std::memcpy(out_assign, a.data(), n * sizeof(int));
        std::memcpy(centroids_inout, C.coordinates.data(), C.n * d * sizeof(float));
        *out_kc = C.n;
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "integ_kmeans: " << e.what() << std::endl;
        return -1;
    }
}

}  // extern "C"
