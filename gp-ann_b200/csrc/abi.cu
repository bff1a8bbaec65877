// abi.cu -- the extern "C" surface of libgpann_cuda.so (include/gpann.h).  Every entry point is a thin dispatcher: a plain
// ctx goes to the single-device implementation (sd_*, gpann_api.cu), a group ctx (gpann_init_multi) to the sharded one
// (mg_*, multi.cu).  No computation lives here and none runs on the host.
#include "api_internal.h"

using namespace gpimpl;

namespace {
inline bool is_group(const gpann_ctx* ctx) { return ctx && ctx->group != nullptr; }
inline bool is_sharded(const gpann_points* pts) { return pts && !pts->shards.empty(); }
int no_group(gpann_ctx* ctx, const char* what) {
    return fail(ctx, GPANN_E_UNSUPPORTED, std::string(what) + ": needs a single-device ctx (group ctxs take the host-buffer entry points)");
}
}  // namespace

extern "C" {

int gpann_init(gpann_ctx** out, int device) { return sd_init(out, device); }
int gpann_init_multi(gpann_ctx** out, const int* devices, int ndev) { return mg_init(out, devices, ndev); }
int gpann_num_devices(const gpann_ctx* ctx) { return !ctx ? 0 : is_group(ctx) ? mg_num_devices(ctx) : 1; }
void gpann_destroy(gpann_ctx* ctx) {
    if (!ctx) return;
    if (is_group(ctx)) mg_destroy(ctx);
    else sd_destroy(ctx);
}
int gpann_comm_unique_id(void* out_id128) { return comm_unique_id(out_id128, 128); }
int gpann_comm_init_rank(gpann_ctx* ctx, int nranks, int rank, const void* id128) { return comm_init_rank(ctx, nranks, rank, id128, 128); }
int gpann_comm_size(const gpann_ctx* ctx) { return !ctx ? 0 : is_group(ctx) ? mg_num_devices(ctx) : ctx->comm_size; }
int gpann_comm_rank(const gpann_ctx* ctx) { return ctx ? ctx->comm_rank : -1; }

const char* gpann_last_error(const gpann_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int gpann_device(const gpann_ctx* ctx) { return ctx ? ctx->device : -1; }
int gpann_set_algo(gpann_ctx* ctx, int algo) {
    if (!ctx || algo < GPANN_ALGO_AUTO || algo > GPANN_ALGO_TENSOR) return GPANN_E_ARG;
    ctx->algo = algo;
    return is_group(ctx) ? mg_for_each(ctx, [algo](gpann_ctx* c) { return sd_set_algo(c, algo); }) : 0;
}
uint64_t gpann_launch_count(const gpann_ctx* ctx) {
    if (!ctx) return 0;
    if (!is_group(ctx)) return ctx->launches;
    uint64_t total = 0;
    for (int i = 0; i < mg_num_devices(ctx); ++i) total += mg_rank_ctx(ctx, i)->launches;
    return total;
}
const char* gpann_last_algo(const gpann_ctx* ctx) { return ctx ? ctx->last_algo : "none"; }
int gpann_set_timing(gpann_ctx* ctx, int on) {
    if (!ctx) return GPANN_E_ARG;
    return is_group(ctx) ? mg_for_each(ctx, [on](gpann_ctx* c) { return sd_set_timing(c, on); }) : sd_set_timing(ctx, on);
}
float gpann_last_kernel_ms(gpann_ctx* ctx) {
    if (!ctx) return -1.f;
    if (!is_group(ctx)) return sd_last_kernel_ms(ctx);
    float worst = -1.f;  // the slowest rank's bracket
    for (int i = 0; i < mg_num_devices(ctx); ++i) worst = std::max(worst, sd_last_kernel_ms(mg_rank_ctx(ctx, i)));
    return worst;
}
uint64_t gpann_last_ambiguous(gpann_ctx* ctx) {
    if (!ctx) return 0;
    if (!is_group(ctx)) return sd_last_ambiguous(ctx);
    uint64_t total = 0;
    for (int i = 0; i < mg_num_devices(ctx); ++i) total += sd_last_ambiguous(mg_rank_ctx(ctx, i));
    return total;
}

int gpann_points_upload(gpann_ctx* ctx, const float* host_rows, size_t n, size_t d, int storage, uint32_t id_base, gpann_points** out) {
    if (!ctx) return GPANN_E_ARG;
    if (storage < GPANN_STORAGE_AUTO || storage > GPANN_STORAGE_S8) return fail(ctx, GPANN_E_ARG, "points_upload: bad storage");
    return is_group(ctx) ? mg_points_upload(ctx, host_rows, n, d, storage, id_base, out) : sd_points_upload(ctx, host_rows, n, d, storage, id_base, out);
}
int gpann_points_upload_bytes(gpann_ctx* ctx, const void* host_rows, size_t n, size_t d, int is_signed, uint32_t id_base, gpann_points** out) {
    if (!ctx) return GPANN_E_ARG;
    return is_group(ctx) ? mg_points_upload_bytes(ctx, host_rows, n, d, is_signed, id_base, out)
                         : sd_points_upload_bytes(ctx, host_rows, n, d, is_signed, id_base, out);
}
int gpann_points_from_device(gpann_ctx* ctx, const void* dev_rows, size_t n, size_t d, int storage, uint32_t id_base, void* stream,
                             gpann_points** out) {
    if (!ctx) return GPANN_E_ARG;
    if (is_group(ctx)) return no_group(ctx, "points_from_device");
    return sd_points_from_device(ctx, dev_rows, n, d, storage, id_base, stream, out);
}
void gpann_points_free(gpann_points* pts) {
    if (!pts) return;
    if (is_sharded(pts) || (pts->ctx && pts->ctx->group)) mg_points_free(pts);
    else sd_points_free(pts);
}
size_t gpann_points_n(const gpann_points* pts) { return pts ? pts->n : 0; }
size_t gpann_points_d(const gpann_points* pts) { return pts ? pts->d : 0; }
int gpann_points_storage(const gpann_points* pts) { return pts ? pts->storage : 0; }

int gpann_topk(gpann_ctx* ctx, const gpann_points* pts, const float* queries, size_t nq, int k, int metric, float* out_dist,
               uint32_t* out_id, uint32_t* out_count) {
    if (!ctx) return GPANN_E_ARG;
    if (!pts) return fail(ctx, GPANN_E_ARG, "topk: null point set");
    if (is_group(ctx)) return mg_topk_host(ctx, pts, 0, pts->n, queries, nq, k, metric, out_dist, out_id, out_count, nullptr);
    return sd_topk(ctx, pts, queries, nq, k, metric, out_dist, out_id, out_count);
}
int gpann_kth_distance(gpann_ctx* ctx, const gpann_points* pts, const float* queries, size_t nq, int k, int metric, float* out_kth) {
    if (!ctx) return GPANN_E_ARG;
    if (!pts || !out_kth) return fail(ctx, GPANN_E_ARG, "kth_distance: null argument");
    if (is_group(ctx)) return mg_topk_host(ctx, pts, 0, pts->n, queries, nq, k, metric, nullptr, nullptr, nullptr, out_kth);
    return sd_kth_distance(ctx, pts, queries, nq, k, metric, out_kth);
}
int gpann_topk_range(gpann_ctx* ctx, const gpann_points* pts, size_t row_begin, size_t row_end, const float* queries, size_t nq, int k,
                     int metric, float* out_dist, uint32_t* out_id, uint32_t* out_count) {
    if (!ctx) return GPANN_E_ARG;
    if (!pts) return fail(ctx, GPANN_E_ARG, "topk_range: null point set");
    if (is_group(ctx)) return mg_topk_host(ctx, pts, row_begin, row_end, queries, nq, k, metric, out_dist, out_id, out_count, nullptr);
    return sd_topk_range(ctx, pts, row_begin, row_end, queries, nq, k, metric, out_dist, out_id, out_count);
}
int gpann_topk_stream_bytes(gpann_ctx* ctx, const void* host_rows, size_t n, size_t d, int is_signed, uint32_t id_base, const float* queries,
                            size_t nq, int k, size_t chunk_rows, float* out_dist, uint32_t* out_id, uint32_t* out_count) {
    if (!ctx) return GPANN_E_ARG;
    if (is_group(ctx))
        return mg_topk_stream_host(ctx, host_rows, is_signed ? kHostS8 : kHostU8, 0, n, d, id_base, queries, nq, k, chunk_rows, out_dist, out_id,
                                   out_count, nullptr);
    return sd_topk_stream_bytes(ctx, host_rows, n, d, is_signed, id_base, queries, nq, k, chunk_rows, out_dist, out_id, out_count);
}
int gpann_topk_stream(gpann_ctx* ctx, const float* host_rows, size_t n, size_t d, int storage, uint32_t id_base, const float* queries, size_t nq,
                      int k, size_t chunk_rows, float* out_dist, uint32_t* out_id, uint32_t* out_count, float* out_kth) {
    if (!ctx) return GPANN_E_ARG;
    if (storage < GPANN_STORAGE_AUTO || storage > GPANN_STORAGE_S8) return fail(ctx, GPANN_E_ARG, "topk_stream: bad storage");
    if (is_group(ctx))
        return mg_topk_stream_host(ctx, host_rows, kHostF32, storage, n, d, id_base, queries, nq, k, chunk_rows, out_dist, out_id, out_count, out_kth);
    return sd_topk_stream(ctx, host_rows, n, d, storage, id_base, queries, nq, k, chunk_rows, out_dist, out_id, out_count, out_kth);
}
int gpann_topk_stream_bytes_device(gpann_ctx* ctx, const void* host_rows, size_t n, size_t d, int is_signed, uint32_t id_base,
                                   const float* dev_queries, size_t nq, int k, size_t chunk_rows, uint64_t* dev_out_keys, void* stream) {
    if (!ctx) return GPANN_E_ARG;
    if (is_group(ctx)) return no_group(ctx, "topk_stream_bytes_device");
    return sd_topk_stream_bytes_device(ctx, host_rows, n, d, is_signed, id_base, dev_queries, nq, k, chunk_rows, dev_out_keys, stream);
}
int gpann_topk_device(gpann_ctx* ctx, const gpann_points* pts, const float* dev_queries, size_t nq, int k, int metric, uint64_t* dev_out_keys,
                      void* stream) {
    if (!ctx) return GPANN_E_ARG;
    if (is_group(ctx)) return no_group(ctx, "topk_device");
    return sd_topk_device(ctx, pts, dev_queries, nq, k, metric, dev_out_keys, stream);
}
int gpann_topk_sharded_device(gpann_ctx* ctx, const gpann_points* pts, const float* dev_queries, size_t nq, int k, int metric,
                              float* dev_out_dist, uint32_t* dev_out_id, uint64_t* dev_out_keys, void* stream) {
    if (!ctx) return GPANN_E_ARG;
    if (is_group(ctx)) return no_group(ctx, "topk_sharded_device");
    return sd_topk_sharded_device(ctx, pts, dev_queries, nq, k, metric, dev_out_dist, dev_out_id, dev_out_keys, stream);
}
int gpann_ivf_topk(gpann_ctx* ctx, const gpann_points* pts, const uint64_t* bucket_offsets, size_t num_buckets, const uint32_t* permutation,
                   const float* queries, size_t nq, const int32_t* probes, int nprobe, int k, int metric, float* out_dist, uint32_t* out_id,
                   uint32_t* out_count) {
    if (!ctx) return GPANN_E_ARG;
    if (is_group(ctx)) return no_group(ctx, "ivf_topk");
    return sd_ivf_topk(ctx, pts, bucket_offsets, num_buckets, permutation, queries, nq, probes, nprobe, k, metric, out_dist, out_id, out_count);
}
int gpann_knn_buckets(gpann_ctx* ctx, const gpann_points* pts, const uint64_t* bucket_offsets, const uint32_t* bucket_members, size_t num_buckets,
                      int k, int metric, float* out_dist, uint32_t* out_id, uint32_t* out_count) {
    if (!ctx) return GPANN_E_ARG;
    if (is_group(ctx)) return no_group(ctx, "knn_buckets");
    return sd_knn_buckets(ctx, pts, bucket_offsets, bucket_members, num_buckets, k, metric, out_dist, out_id, out_count);
}
int gpann_closest_leaders_batch(gpann_ctx* ctx, const gpann_points* pts, const uint32_t* point_ids, size_t m, const gpann_points* leaders, int k,
                                int metric, float* out_dist, uint32_t* out_id, uint32_t* out_count) {
    if (!ctx) return GPANN_E_ARG;
    if (is_group(ctx)) return no_group(ctx, "closest_leaders_batch");
    return sd_closest_leaders_batch(ctx, pts, point_ids, m, leaders, k, metric, out_dist, out_id, out_count);
}
int gpann_group_min_dist(gpann_ctx* ctx, const gpann_points* pts, const int32_t* group, int ngroups, const float* queries, size_t nq, int metric,
                         float* out) {
    if (!ctx) return GPANN_E_ARG;
    if (is_group(ctx)) return no_group(ctx, "group_min_dist");
    return sd_group_min_dist(ctx, pts, group, ngroups, queries, nq, metric, out);
}
int gpann_merge_keys_device(gpann_ctx* ctx, const uint64_t* dev_keys, size_t parts, size_t nq, int k, float* dev_out_dist, uint32_t* dev_out_id,
                            uint64_t* dev_out_keys, void* stream) {
    if (!ctx) return GPANN_E_ARG;
    if (is_group(ctx)) return no_group(ctx, "merge_keys_device");
    return sd_merge_keys_device(ctx, dev_keys, parts, nq, k, dev_out_dist, dev_out_id, dev_out_keys, stream);
}
uint64_t gpann_pack_key(float dist, uint32_t id) { return pack_key(dist, id); }
void gpann_unpack_key(uint64_t key, float* dist, uint32_t* id) {
    if (dist) *dist = key_dist(key);
    if (id) *id = key_id(key);
}

int gpann_kmeans_assign(gpann_ctx* ctx, const gpann_points* pts, const float* centroids, size_t kc, int metric, int32_t* out_assign,
                        float* out_mindist) {
    if (!ctx) return GPANN_E_ARG;
    return is_group(ctx) ? mg_kmeans_assign(ctx, pts, centroids, kc, metric, out_assign, out_mindist)
                         : sd_kmeans_assign(ctx, pts, centroids, kc, metric, out_assign, out_mindist);
}
int gpann_kmeans_partial_device(gpann_ctx* ctx, const gpann_points* pts, const float* dev_centroids, size_t kc, int metric,
                                int32_t* dev_out_assign, float* dev_out_sums, uint64_t* dev_out_counts, void* stream) {
    if (!ctx) return GPANN_E_ARG;
    if (is_group(ctx)) return no_group(ctx, "kmeans_partial_device");
    return sd_kmeans_partial_device(ctx, pts, dev_centroids, kc, metric, dev_out_assign, dev_out_sums, dev_out_counts, stream);
}
int gpann_kmeans_step(gpann_ctx* ctx, const gpann_points* pts, float* centroids_inout, size_t kc, int metric, int32_t* out_assign,
                      uint64_t* out_sizes) {
    if (!ctx) return GPANN_E_ARG;
    return is_group(ctx) ? mg_kmeans_step(ctx, pts, centroids_inout, kc, metric, out_assign, out_sizes)
                         : sd_kmeans_step(ctx, pts, centroids_inout, kc, metric, out_assign, out_sizes);
}
int gpann_kmeans_run(gpann_ctx* ctx, const gpann_points* pts, float* centroids_inout, size_t kc, int rounds, int metric, int32_t* out_assign,
                     uint64_t* out_sizes, size_t* out_kc) {
    if (!ctx) return GPANN_E_ARG;
    return is_group(ctx) ? mg_kmeans_run(ctx, pts, centroids_inout, kc, rounds, metric, out_assign, out_sizes, out_kc)
                         : sd_kmeans_run(ctx, pts, centroids_inout, kc, rounds, metric, out_assign, out_sizes, out_kc);
}
int gpann_kmeans_run_device(gpann_ctx* ctx, const gpann_points* pts, float* dev_centroids_inout, size_t kc, int rounds, int metric,
                            int32_t* dev_out_assign, unsigned char* dev_alive_inout, uint64_t* dev_out_counts, void* stream) {
    if (!ctx) return GPANN_E_ARG;
    if (is_group(ctx)) return no_group(ctx, "kmeans_run_device");
    return sd_kmeans_run_device(ctx, pts, dev_centroids_inout, kc, rounds, metric, dev_out_assign, dev_alive_inout, dev_out_counts, stream);
}
int gpann_balanced_kmeans(gpann_ctx* ctx, const gpann_points* pts, float* centroids_inout, size_t kc, size_t max_cluster_size, int metric,
                          int32_t* out_assign, uint64_t* out_sizes, size_t* out_kc) {
    if (!ctx) return GPANN_E_ARG;
    if (is_group(ctx)) return no_group(ctx, "balanced_kmeans");
    return sd_balanced_kmeans(ctx, pts, centroids_inout, kc, max_cluster_size, metric, out_assign, out_sizes, out_kc);
}
int gpann_objective(gpann_ctx* ctx, const gpann_points* pts, const float* centroids, size_t kc, const int32_t* assign, int metric, double* out) {
    if (!ctx) return GPANN_E_ARG;
    return is_group(ctx) ? mg_objective(ctx, pts, centroids, kc, assign, metric, out) : sd_objective(ctx, pts, centroids, kc, assign, metric, out);
}

}  // extern "C"
