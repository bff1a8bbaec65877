// dropin_shim.cpp -- C shim over the C++ drop-in layer, used by tests/test_gpu_dropin.py.  It exposes the same
// entry points as the shim the tests put over the compiled reference, so the parity tests of the drop-in headers read
// like the tests of the reference itself.
#include <cstring>

#include "recall.h"  // defines ComputeGroundTruth & co
#include "ground_truth_file.h"
#include "kmeans.h"
#include "points_io.h"

static PointSet make_pointset(const float* x, size_t n, size_t d) {
    PointSet ps;
    ps.n = n;
    ps.d = d;
    ps.coordinates.assign(x, x + n * d);
    return ps;
}

extern "C" {

float dropin_distance(const float* p, const float* q, unsigned d) { return distance(const_cast<float*>(p), const_cast<float*>(q), d); }

int dropin_compute_ground_truth(const float* points, size_t n, size_t d, const float* queries, size_t nq, int k, float* out_dist,
                                uint32_t* out_id, uint32_t* out_count) {
    try {
        PointSet P = make_pointset(points, n, d), Q = make_pointset(queries, nq, d);
        std::vector<NNVec> gt = ComputeGroundTruth(P, Q, k);
        for (size_t q = 0; q < nq; ++q) {
            out_count[q] = (uint32_t) gt[q].size();
            for (size_t r = 0; r < gt[q].size(); ++r) {
                out_dist[q * k + r] = gt[q][r].first;
                out_id[q * k + r] = gt[q][r].second;
            }
        }
        return 0;
    } catch (const std::exception& e) {
        std::cerr << e.what() << std::endl;
        return -1;
    }
}

int dropin_kth_distance(const float* points, size_t n, size_t d, const float* queries, size_t nq, int k, float* out) {
    try {
        PointSet P = make_pointset(points, n, d), Q = make_pointset(queries, nq, d);
        std::vector<float> r = ComputeDistanceToKthNeighbor(P, Q, k);
        std::memcpy(out, r.data(), nq * sizeof(float));
        return 0;
    } catch (const std::exception& e) {
        std::cerr << e.what() << std::endl;
        return -1;
    }
}

int dropin_convert_ground_truth(const float* points, size_t n, size_t d, const float* queries, size_t nq, int k, const float* gt_dist,
                                const uint32_t* gt_id, float* out_kth) {
    PointSet P = make_pointset(points, n, d), Q = make_pointset(queries, nq, d);
    std::vector<NNVec> gt(nq, NNVec(k));
    for (size_t q = 0; q < nq; ++q)
        for (int r = 0; r < k; ++r) gt[q][r] = { gt_dist[q * k + r], gt_id[q * k + r] };
    std::vector<float> kth = ConvertGroundTruthToDistanceToKthNeighbor(gt, k, P, Q);
    std::memcpy(out_kth, kth.data(), nq * sizeof(float));
    return 0;
}

long long dropin_kmeans(const float* points, size_t n, size_t d, float* centroids, size_t kc, int32_t* out_assign) {
    try {
        PointSet P = make_pointset(points, n, d), C = make_pointset(centroids, kc, d);
        std::vector<int> a = KMeans(P, C);
        for (size_t i = 0; i < n; ++i) out_assign[i] = a[i];
        std::memcpy(centroids, C.coordinates.data(), C.n * d * sizeof(float));
        return (long long) C.n;
    } catch (const std::runtime_error&) {
        return -1;
    }
}

double dropin_objective(const float* points, size_t n, size_t d, const float* centroids, size_t kc, const int32_t* assign) {
    PointSet P = make_pointset(points, n, d), C = make_pointset(centroids, kc, d);
    std::vector<int> a(assign, assign + n);
    return ObjectiveValue(P, C, a);
}

void dropin_random_sample(const float* points, size_t n, size_t d, size_t num_samples, int seed, float* out_rows) {
    PointSet P = make_pointset(points, n, d);
    PointSet S = RandomSample(P, num_samples, seed);
    std::memcpy(out_rows, S.coordinates.data(), S.coordinates.size() * sizeof(float));
}

void dropin_closest_leaders(const float* points, size_t np, size_t d, const float* leaders, size_t nl, uint32_t my_id, int k,
                            float* out_dist, uint32_t* out_id, uint32_t* out_count) {
    PointSet P = make_pointset(points, np, d), L = make_pointset(leaders, nl, d);
    TopN t = ClosestLeaders(P, L, my_id, k);
    std::vector<TopN::value_type> v = t.Take();
    *out_count = (uint32_t) v.size();
    for (size_t r = 0; r < v.size(); ++r) {
        out_dist[r] = v[r].first;
        out_id[r] = v[r].second;
    }
}


// ComputeGroundTruthFromFile (bytes stay bytes) + WriteGroundTruth in the reference's file format; then read it back
int dropin_ground_truth_file(const char* points_path, const float* queries, size_t nq, size_t d, int k, size_t chunk_rows, const char* gt_path,
                             float* out_dist, uint32_t* out_id) {
    try {
        PointSet Q = make_pointset(queries, nq, d);
        std::vector<NNVec> gt = ComputeGroundTruthFromFile(points_path, Q, k, chunk_rows);
        WriteGroundTruth(gt_path, gt);
        std::vector<NNVec> back = ReadGroundTruth(gt_path);
        for (size_t q = 0; q < nq; ++q)
            for (size_t r = 0; r < back[q].size(); ++r) {
                out_dist[q * k + r] = back[q][r].first;
                out_id[q * k + r] = back[q][r].second;
            }
        return 0;
    } catch (const std::exception& e) {
        std::cerr << e.what() << std::endl;
        return -1;
    }
}

// host-only round trips of the file formats (no GPU): ReadPoints widens bytes like the reference's reader
int dropin_read_points(const char* path, float* out, size_t cap, size_t* n, size_t* d) {
    try {
        PointSet P = ReadPoints(path);
        *n = P.n;
        *d = P.d;
        if (P.coordinates.size() > cap) return -2;
        std::memcpy(out, P.coordinates.data(), P.coordinates.size() * sizeof(float));
        return 0;
    } catch (const std::exception& e) {
        std::cerr << e.what() << std::endl;
        return -1;
    }
}
int dropin_write_ground_truth(const char* path, const float* dist, const uint32_t* ids, size_t nq, size_t k) {
    std::vector<NNVec> gt(nq, NNVec(k));
    for (size_t q = 0; q < nq; ++q)
        for (size_t r = 0; r < k; ++r) gt[q][r] = { dist[q * k + r], ids[q * k + r] };
    WriteGroundTruth(path, gt);
    return 0;
}

long long dropin_balanced_kmeans(const float* points, size_t n, size_t d, float* centroids, size_t kc, size_t max_cluster_size, int32_t* out_assign) {
    try {
        PointSet P = make_pointset(points, n, d), C = make_pointset(centroids, kc, d);
        std::vector<int> a = BalancedKMeans(P, C, max_cluster_size);
        for (size_t i = 0; i < n; ++i) out_assign[i] = a[i];
        std::memcpy(centroids, C.coordinates.data(), C.n * d * sizeof(float));
        return (long long) C.n;
    } catch (const std::exception& e) {
        std::cerr << e.what() << std::endl;
        return -1;
    }
}
}
