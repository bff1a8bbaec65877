// recall.h -- drop-in for gp-ann's src/recall.h.  Like the reference's header it DEFINES its functions (include it
// from one translation unit per executable).
//   ComputeDistanceToKthNeighbor (:10-23) and ComputeGroundTruth (:25-39): the n*nq scan runs on the GPU through
//   gpann_topk_stream (host rows streamed in chunks) or gpann_points_upload + gpann_topk / gpann_kth_distance; results are the k lexicographically smallest (distance,id) per query,
//   ascending, min(k,n) entries (topn.h:21-29).
//   OracleRecall (:41-68), ConvertGroundTruthToDistanceToKthNeighbor (:74-137), Recall (:139-150): O(nq*k) host
//   bookkeeping, kept on the CPU with the reference's behaviour and printed text.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <iostream>
#include <vector>

#include "defs.h"
#include "dist.h"
#include "gpann_dropin.h"
#include "topn.h"

namespace gpann_dropin {
// L2, d <= 128: the host PointSet is streamed through the device chunk by chunk (gpann_topk_stream: H2D overlapped with
// the scan, byte-valued rows narrowed on the device); everything else uploads the set and scans it resident.  Both are
// GPU paths -- GPANN_E_UNSUPPORTED from the first only selects the second.
inline void topk_any(Session& s, PointSet& points, PointSet& queries, int k, float* dist, uint32_t* ids, uint32_t* count,
                     float* kth) {
    if (kMetric == GPANN_METRIC_L2 && points.d <= 128 && k <= 448) {
        const int rc = gpann_topk_stream(s.ctx, points.coordinates.data(), points.n, points.d, GPANN_STORAGE_AUTO, 0,
                                         queries.coordinates.data(), queries.n, k, 0, dist, ids, count, kth);
        if (rc != GPANN_E_UNSUPPORTED) {
            s.check(rc);
            return;
        }
    }
    DevicePoints<PointSet> dp(s, points);
    if (dist || ids || count)
        s.check(gpann_topk(s.ctx, dp.h, queries.coordinates.data(), queries.n, k, kMetric, dist, ids, count));
    if (kth) s.check(gpann_kth_distance(s.ctx, dp.h, queries.coordinates.data(), queries.n, k, kMetric, kth));
}
}  // namespace gpann_dropin

std::vector<float> ComputeDistanceToKthNeighbor(PointSet& points, PointSet& queries, int k) {
    auto& s = gpann_dropin::session();
    std::lock_guard<std::mutex> lock(s.mutex);
    std::vector<float> kth(queries.n);
    gpann_dropin::topk_any(s, points, queries, k, nullptr, nullptr, nullptr, kth.data());
    return kth;
}

std::vector<NNVec> ComputeGroundTruth(PointSet& points, PointSet& queries, int k) {
    auto& s = gpann_dropin::session();
    std::lock_guard<std::mutex> lock(s.mutex);
    const size_t nq = queries.n;
    std::vector<float> dist(nq * k);
    std::vector<uint32_t> ids(nq * k), count(nq);
    gpann_dropin::topk_any(s, points, queries, k, dist.data(), ids.data(), count.data(), nullptr);
    std::vector<NNVec> res(nq);
    for (size_t q = 0; q < nq; ++q) {
        res[q].resize(count[q]);
        for (uint32_t r = 0; r < count[q]; ++r) res[q][r] = { dist[q * k + r], ids[q * k + r] };
    }
    return res;
}

void OracleRecall(const std::vector<NNVec>& ground_truth, const std::vector<int>& partition, int num_neighbors) {
    const int num_shards = NumPartsInPartition(partition);
    std::vector<size_t> hits(num_shards, 0);
    std::vector<int> freq(num_shards);
    for (const NNVec& neigh : ground_truth) {
        std::fill(freq.begin(), freq.end(), 0);
        for (int i = 0; i < num_neighbors; ++i) freq[partition[neigh[i].second]]++;
        std::sort(freq.begin(), freq.end(), std::greater<int>());  // best shard first
        for (int i = 0; i < num_shards; ++i) hits[i] += freq[i];
    }
    size_t total = 0;
    for (int i = 0; i < num_shards && hits[i] > 0; ++i) {
        total += hits[i];
        const double recall = static_cast<double>(total) / ground_truth.size() / num_neighbors;
        std::cout << "nprobes = " << i + 1 << " oracle recall = " << recall << std::endl;
    }
}

// Re-derives every ground-truth distance on the host, reports how many stored distances / orders were off and
// returns the k-th distance per query (the reference's self-check).
std::vector<float> ConvertGroundTruthToDistanceToKthNeighbor(std::vector<NNVec>& ground_truth, int k, PointSet& points,
                                                             PointSet& queries) {
    if (ground_truth.size() != queries.n) {
        std::cout << "Ground truth size and number of queries don't match." << std::endl;
        std::exit(0);
    }
    std::cout << "Convert ground truth. k = " << k << " points.d " << points.d << std::endl;
    const std::vector<double> epss = { 1e-1, 1e-2, 1e-3, 1e-4, 1e-5, 1e-6, 1e-7, 1e-8, 1e-9, 1e-10 };
    std::vector<size_t> wrong(epss.size(), 0);
    size_t mismatches = 0, unsorted_before = 0, unsorted_after = 0;
    std::vector<float> kth(ground_truth.size());
    auto by_distance = [](const std::pair<float, uint32_t>& l, const std::pair<float, uint32_t>& r) { return l.first < r.first; };
    for (size_t q = 0; q < queries.n; ++q) {
        NNVec& neighs = ground_truth[q];
        if (!std::is_sorted(neighs.begin(), neighs.begin() + k, by_distance)) unsorted_before++;
        float* Q = queries.GetPoint(q);
        for (int j = 0; j < k; ++j) {
            const float stored = neighs[j].first;
            const float actual = distance(points.GetPoint(neighs[j].second), Q, points.d);
            const double err = std::abs(stored - actual);
            if (err > 1e-8) mismatches++;
            for (size_t r = 0; r < epss.size(); ++r) wrong[r] += err > epss[r];
            neighs[j].first = actual;
        }
        if (!std::is_sorted(neighs.begin(), neighs.begin() + k, by_distance)) {
            std::sort(neighs.begin(), neighs.begin() + k, by_distance);
            unsorted_after++;
        }
        kth[q] = neighs[k - 1].first;
    }
    std::cout << mismatches << " out of " << ground_truth.size() * ground_truth[0].size() << " distances were wrong" << std::endl;
    std::cout << unsorted_before << " out of " << ground_truth.size()
              << " neighbors lists were ordered incorrectly before recomputing distances. And " << unsorted_after
              << " were ordered incorrectly after recomputing" << std::endl;
    for (size_t r = 0; r < epss.size(); ++r)
        std::cout << "For eps = " << epss[r] << " there were " << wrong[r] << " many distances wrong, i.e., |d1 - d2| > eps"
                  << std::endl;
    return kth;
}

double Recall(const std::vector<NNVec>& neighbors_per_query, const std::vector<float>& distance_to_kth_neighbor, int k) {
    size_t hits = 0;
    for (size_t i = 0; i < neighbors_per_query.size(); ++i)
        for (const auto& x : neighbors_per_query[i]) hits += x.first <= distance_to_kth_neighbor[i];
    return static_cast<double>(hits) / static_cast<double>(neighbors_per_query.size() * k);
}
