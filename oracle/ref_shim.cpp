// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE.  A C-ABI shim over the UNMODIFIED reference hot path.
//
// Compiled by oracle/build.py together with /root/reference/src/{dist,defs,kmeans}.cpp, taken from where
// they lie (no reference source is copied into this repository), into oracle/_ref/libgpann_ref.so
// (L2 build) and oracle/_ref/libgpann_ref_mips.so (-DMIPS_DISTANCE build, CMakeLists.txt:54-59).
// It only marshals flat arrays into the reference's PointSet / NNVec types and calls the reference's own
// ComputeGroundTruth, ComputeDistanceToKthNeighbor, KMeans, ObjectiveValue, RandomSample, ClosestLeaders,
// distance and InvertedIndex.  Used by tests (to pin oracle/gpann_oracle.cpp), by
// tests/golden/make_golden.py and by bench.py's cpu_baseline / --impl reference leg.
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

#include <parlay/parallel.h>
#include <parlay/primitives.h>

#include "recall.h"  // defines (non-inline) ComputeGroundTruth & co; pulls in defs.h, dist.h, topn.h
#include "kmeans.h"
#include "inverted_index.h"
#include "knn_graph.h"

static PointSet make_pointset(const float* x, size_t n, size_t d) {
    PointSet ps;
    ps.n = n;
    ps.d = d;
    ps.coordinates.assign(x, x + n * d);
    return ps;
}

extern "C" {

int ref_num_workers() { return (int) parlay::num_workers(); }

#ifdef MIPS_DISTANCE
int ref_metric() { return 1; }
#else
int ref_metric() { return 0; }
#endif

float ref_distance(const float* p, const float* q, unsigned d) {
    return distance(const_cast<float*>(p), const_cast<float*>(q), d);
}

void ref_compute_ground_truth(const float* points, size_t n, size_t d, const float* queries, size_t nq, int k,
                              float* out_dist, uint32_t* out_id, uint32_t* out_count) {
    PointSet P = make_pointset(points, n, d), Q = make_pointset(queries, nq, d);
    std::vector<NNVec> gt = ComputeGroundTruth(P, Q, k);
    for (size_t q = 0; q < nq; ++q) {
        if (out_count) out_count[q] = (uint32_t) gt[q].size();
        for (size_t r = 0; r < gt[q].size(); ++r) {
            out_dist[q * k + r] = gt[q][r].first;
            out_id[q * k + r] = gt[q][r].second;
        }
    }
}

// Timing-only variant: PointSets are built by the caller once (handles), so the timed region is exactly
// the reference's ComputeGroundTruth call.
void* ref_pointset_create(const float* x, size_t n, size_t d) { return new PointSet(make_pointset(x, n, d)); }
void ref_pointset_free(void* p) { delete static_cast<PointSet*>(p); }
size_t ref_ground_truth_handles(void* points, void* queries, int k, uint32_t* out_id_or_null) {
    std::vector<NNVec> gt = ComputeGroundTruth(*static_cast<PointSet*>(points), *static_cast<PointSet*>(queries), k);
    size_t total = 0;
    for (size_t q = 0; q < gt.size(); ++q) {
        if (out_id_or_null)
            for (size_t r = 0; r < gt[q].size(); ++r) out_id_or_null[q * k + r] = gt[q][r].second;
        total += gt[q].size();
    }
    return total;
}

void ref_kth_distance(const float* points, size_t n, size_t d, const float* queries, size_t nq, int k, float* out) {
    PointSet P = make_pointset(points, n, d), Q = make_pointset(queries, nq, d);
    std::vector<float> r = ComputeDistanceToKthNeighbor(P, Q, k);
    std::memcpy(out, r.data(), nq * sizeof(float));
}

// The reference's own self-check (recall.h:74-137): returns the kth distances after recomputation.
void ref_convert_ground_truth(const float* points, size_t n, size_t d, const float* queries, size_t nq, int k,
                              const float* gt_dist, const uint32_t* gt_id, float* out_kth) {
    PointSet P = make_pointset(points, n, d), Q = make_pointset(queries, nq, d);
    std::vector<NNVec> gt(nq, NNVec(k));
    for (size_t q = 0; q < nq; ++q)
        for (int r = 0; r < k; ++r) gt[q][r] = { gt_dist[q * k + r], gt_id[q * k + r] };
    std::vector<float> kth = ConvertGroundTruthToDistanceToKthNeighbor(gt, k, P, Q);
    std::memcpy(out_kth, kth.data(), nq * sizeof(float));
}

double ref_recall(const float* nb_dist, size_t nq, int per_query, const float* kth, int k) {
    std::vector<NNVec> nb(nq, NNVec(per_query));
    for (size_t q = 0; q < nq; ++q)
        for (int r = 0; r < per_query; ++r) nb[q][r] = { nb_dist[q * per_query + r], 0u };
    std::vector<float> kthv(kth, kth + nq);
    return Recall(nb, kthv, k);
}

// KMeans (kmeans.cpp:228-244).  Returns the surviving number of centroids, -1 if the reference threw.
long long ref_kmeans(const float* points, size_t n, size_t d, float* centroids, size_t kc, int32_t* out_assign) {
    PointSet P = make_pointset(points, n, d), C = make_pointset(centroids, kc, d);
    try {
        std::vector<int> a = KMeans(P, C);
        for (size_t i = 0; i < n; ++i) out_assign[i] = a[i];
    } catch (const std::runtime_error&) {
        return -1;
    }
    std::memcpy(centroids, C.coordinates.data(), C.n * d * sizeof(float));
    return (long long) C.n;
}

void* ref_kmeans_handles(void* points, void* centroids) {
    auto* a = new std::vector<int>(KMeans(*static_cast<PointSet*>(points), *static_cast<PointSet*>(centroids)));
    return a;
}
void ref_intvec_free(void* v) { delete static_cast<std::vector<int>*>(v); }
size_t ref_pointset_n(void* p) { return static_cast<PointSet*>(p)->n; }

double ref_objective(const float* points, size_t n, size_t d, const float* centroids, size_t kc,
                     const int32_t* assign) {
    PointSet P = make_pointset(points, n, d), C = make_pointset(centroids, kc, d);
    std::vector<int> a(assign, assign + n);
    return ObjectiveValue(P, C, a);
}

void ref_random_sample(const float* points, size_t n, size_t d, size_t num_samples, int seed, float* out_rows) {
    PointSet P = make_pointset(points, n, d);
    PointSet S = RandomSample(P, num_samples, seed);
    std::memcpy(out_rows, S.coordinates.data(), S.coordinates.size() * sizeof(float));
}

void ref_closest_leaders(const float* points, size_t np, size_t d, const float* leaders, size_t nl, uint32_t my_id,
                         int k, float* out_dist, uint32_t* out_id, uint32_t* out_count) {
    PointSet P = make_pointset(points, np, d), L = make_pointset(leaders, nl, d);
    TopN t = ClosestLeaders(P, L, my_id, k);
    std::vector<TopN::value_type> v = t.Take();
    *out_count = (uint32_t) v.size();
    for (size_t r = 0; r < v.size(); ++r) {
        out_dist[r] = v[r].first;
        out_id[r] = v[r].second;
    }
}

// InvertedIndex (inverted_index.h:12-75): clusters given as CSR (cluster_offsets[nc+1], cluster_members).
void ref_ivf_query(const float* points, size_t n, size_t d, const int32_t* cluster_offsets,
                   const uint32_t* cluster_members, size_t num_clusters, const float* Q, int k, const int32_t* buckets,
                   size_t num_buckets, float* out_dist, uint32_t* out_id, uint32_t* out_count) {
    PointSet P = make_pointset(points, n, d);
    Clusters clusters(num_clusters);
    for (size_t c = 0; c < num_clusters; ++c)
        clusters[c].assign(cluster_members + cluster_offsets[c], cluster_members + cluster_offsets[c + 1]);
    InvertedIndex ivf(P, clusters);
    std::vector<int> b(buckets, buckets + num_buckets);
    NNVec r = ivf.Query(const_cast<float*>(Q), k, b, num_buckets);
    *out_count = (uint32_t) r.size();
    for (size_t i = 0; i < r.size(); ++i) {
        out_dist[i] = r[i].first;
        out_id[i] = r[i].second;
    }
}

// BalancedKMeans (kmeans.cpp:251-524); prints its progress to stdout like the reference does.
long long ref_balanced_kmeans(const float* points, size_t n, size_t d, float* centroids, size_t kc, size_t max_cluster_size, int32_t* out_assign) {
    PointSet P = make_pointset(points, n, d), C = make_pointset(centroids, kc, d);
    std::vector<int> a = BalancedKMeans(P, C, max_cluster_size);
    for (size_t i = 0; i < n; ++i) out_assign[i] = a[i];
    std::memcpy(centroids, C.coordinates.data(), C.n * d * sizeof(float));
    return (long long) C.n;
}

// ApproximateKNNGraphBuilder::CrunchBucket (knn_graph.h:165-185): all pairs inside one bucket, per-member top-k of
// (distance, global id); member j's list goes to out_*[j*k ..].
void ref_crunch_bucket(const float* points, size_t n, size_t d, const uint32_t* bucket, size_t m, int k, float* out_dist,
                       uint32_t* out_id, uint32_t* out_count) {
    PointSet P = make_pointset(points, n, d);
    ApproximateKNNGraphBuilder builder;
    builder.quiet = true;
    std::vector<uint32_t> b(bucket, bucket + m);
    std::vector<NNVec> res = builder.CrunchBucket(P, b, k);
    for (size_t j = 0; j < m; ++j) {
        out_count[j] = (uint32_t) res[j].size();
        for (size_t r = 0; r < res[j].size(); ++r) {
            out_dist[j * k + r] = res[j][r].first;
            out_id[j * k + r] = res[j][r].second;
        }
    }
}

}  // extern "C"
