// kmeans_gpu.cpp -- KMeans / ObjectiveValue / RandomSample / ClosestLeaders of the drop-in layer.
//   KMeans (reference src/kmeans.cpp:228-244): 20 fixed Lloyd rounds; each round = gpann_kmeans_step (nearest
//   centroid with first-minimum ties + deterministic per-cluster sums + division on the GPU) followed by the
//   reference's host-side empty-cluster compaction (src/kmeans.cpp:29-50).  Points are uploaded once per call.
//   RandomSample (:207-226) stays on the host verbatim in behaviour: std::iota + std::mt19937 + std::sample.
#include <numeric>
#include <random>
#include <stdexcept>

#include "dist.h"
#include "gpann_dropin.h"
#include "kmeans.h"

PointSet RandomSample(PointSet& points, size_t num_samples, int seed) {
    PointSet sample_set;
    sample_set.n = num_samples;
    sample_set.d = points.d;
    std::vector<int> all(points.n);
    std::iota(all.begin(), all.end(), 0);
    std::mt19937 prng(seed);
    std::vector<int> chosen(num_samples);
    std::sample(all.begin(), all.end(), chosen.begin(), num_samples, prng);
    sample_set.coordinates.reserve(num_samples * points.d);
    for (int row : chosen) {
        const float* p = points.GetPoint(row);
        sample_set.coordinates.insert(sample_set.coordinates.end(), p, p + points.d);
    }
    return sample_set;
}

namespace {
// src/kmeans.cpp:29-50: drop centroids without members, renumber assignments
void CompactEmptyClusters(PointSet& centroids, std::vector<int>& closest_center, const std::vector<uint64_t>& sizes) {
    if (std::none_of(sizes.begin(), sizes.end(), [](uint64_t s) { return s == 0; })) return;
    std::vector<int> new_id(centroids.n, -1);
    size_t kept = 0;
    for (size_t c = 0; c < centroids.n; ++c) {
        if (sizes[c] == 0) continue;
        if (kept != c) std::copy_n(centroids.GetPoint(c), centroids.d, centroids.GetPoint(kept));
        new_id[c] = static_cast<int>(kept++);
    }
    centroids.n = kept;
    centroids.coordinates.resize(kept * centroids.d);
    for (int& a : closest_center) a = new_id[a];
}
}  // namespace

std::vector<int> KMeans(PointSet& P, PointSet& centroids) {
    if (centroids.n < 1) throw std::runtime_error("KMeans #centroids < 1");
    auto& s = gpann_dropin::session();
    std::lock_guard<std::mutex> lock(s.mutex);
    std::vector<int> closest_center(P.n, -1);
    // k-means needs the float rows (sums, exact distances): keep F32 storage
    gpann_dropin::DevicePoints<PointSet> dp(s, P, GPANN_STORAGE_F32);
    static constexpr size_t NUM_ROUNDS = 20;
    for (size_t r = 0; r < NUM_ROUNDS; ++r) {
        std::vector<uint64_t> sizes(centroids.n);
        s.check(gpann_kmeans_step(s.ctx, dp.h, centroids.coordinates.data(), centroids.n, gpann_dropin::kMetric,
                                  closest_center.data(), sizes.data()));
        CompactEmptyClusters(centroids, closest_center, sizes);
    }
    return closest_center;
}

double ObjectiveValue(PointSet& points, PointSet& centroids, const std::vector<int>& closest_center) {
    auto& s = gpann_dropin::session();
    std::lock_guard<std::mutex> lock(s.mutex);
    gpann_dropin::DevicePoints<PointSet> dp(s, points, GPANN_STORAGE_F32);
    double total = 0.0;
    s.check(gpann_objective(s.ctx, dp.h, centroids.coordinates.data(), centroids.n, closest_center.data(), gpann_dropin::kMetric,
                            &total));
    return total;
}

// src/defs.cpp:87-96.  One query against <= ~1500 leaders: a handful of microseconds of host work, and it is called
// per point from inside worker threads (knn_graph.h:75-81), so the single-query form stays lock-free host code; the
// batched form (all points against the leaders at once) is gpann_topk on the leader set.
TopN ClosestLeaders(PointSet& points, PointSet& leader_points, uint32_t my_id, int k) {
    TopN best(k);
    float* me = points.GetPoint(my_id);
    for (uint32_t j = 0; j < leader_points.n; ++j) best.Add({ distance(leader_points.GetPoint(j), me, points.d), j });
    return best;
}
