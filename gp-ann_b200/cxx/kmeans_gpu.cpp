// kmeans_gpu.cpp -- KMeans / BalancedKMeans / ObjectiveValue / RandomSample of the drop-in layer: everything gp-ann's
// src/kmeans.h declares, so this file REPLACES src/kmeans.cpp in a gp-ann build.
//   KMeans (reference src/kmeans.cpp:228-244): 20 fixed Lloyd rounds = ONE gpann_kmeans_run call (nearest centroid with
//   first-minimum ties + deterministic per-cluster sums + division + empty-cluster removal, all on the GPU(s); on a group
//   ctx the sums and counts are all-reduced over NCCL every round).  Points are uploaded once per call.
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

std::vector<int> KMeans(PointSet& P, PointSet& centroids) {
    if (centroids.n < 1) throw std::runtime_error("KMeans #centroids < 1");
    auto& s = gpann_dropin::session();
    std::lock_guard<std::mutex> lock(s.mutex);
    std::vector<int> closest_center(P.n, -1);
    // k-means needs the float rows (sums, exact distances): keep F32 storage
    gpann_dropin::DevicePoints<PointSet> dp(s, P, GPANN_STORAGE_F32);
    static constexpr int NUM_ROUNDS = 20;  // src/kmeans.cpp:238
    // One call for all rounds: the centroids stay on the device, clusters that lose every member are removed as
    // RemoveEmptyClusters does (src/kmeans.cpp:29-50) and renumbered once at the end; one D2H of the assignment.
    size_t kept = centroids.n;
    s.check(gpann_kmeans_run(s.ctx, dp.h, centroids.coordinates.data(), centroids.n, NUM_ROUNDS, gpann_dropin::kMetric,
                             closest_center.data(), nullptr, &kept));
    centroids.n = kept;
    centroids.coordinates.resize(kept * centroids.d);
    return closest_center;
}

// BalancedKMeans (src/kmeans.cpp:251-524): KMeans + the BKM+ penalty rounds, every distance on the GPU (gpann_balanced_kmeans).
// The reference's progress prints are not reproduced.  `centroids` is updated in place and may shrink (the final
// AggregateClustersParallel removes empty clusters, :477).
std::vector<int> BalancedKMeans(PointSet& points, PointSet& centroids, size_t max_cluster_size) {
    if (centroids.n < 1) throw std::runtime_error("KMeans #centroids < 1");
    auto& s = gpann_dropin::session();
    std::lock_guard<std::mutex> lock(s.mutex);
    gpann_ctx* one = s.single_device();
    gpann_points* h = nullptr;
    s.check_single(gpann_points_upload(one, points.coordinates.data(), points.n, points.d, GPANN_STORAGE_F32, 0, &h));
    std::vector<int> closest_center(points.n, -1);
    size_t kept = centroids.n;
    const int rc = gpann_balanced_kmeans(one, h, centroids.coordinates.data(), centroids.n, max_cluster_size, gpann_dropin::kMetric,
                                         closest_center.data(), nullptr, &kept);
    gpann_points_free(h);
    s.check_single(rc);
    centroids.n = kept;
    centroids.coordinates.resize(kept * centroids.d);
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
