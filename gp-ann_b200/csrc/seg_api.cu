// seg_api.cu -- host side of the segmented exact scans (seg.cu): the batched forms of the callers either side of the big
// scans.  Each entry point turns its index structure (bucket offsets, probe lists, bucket member lists) into a
// flat task list on the host -- O(#queries x #probes) integer work -- uploads it, runs ONE scan launch + ONE merge launch
// and copies the [nq,k] result back.
//   gpann_ivf_topk               InvertedIndex::Query / QueryBucket for a batch of queries   src/inverted_index.h:43-75
//   gpann_knn_buckets            ApproximateKNNGraphBuilder::CrunchBucket for all buckets    src/knn_graph.h:165-185
//   gpann_closest_leaders_batch  ClosestLeaders for every point of a set                      src/knn_graph.h:75-81, src/defs.cpp:87-96
#include <numeric>

#include "api_internal.h"

namespace gpimpl {

namespace {
constexpr uint64_t kSegMaxPos = 32768;  // positions per task: longer ranges are split (load balance across the SMs)

struct SegJob {
    std::vector<SegTask> tasks;
    std::vector<uint32_t> task_query;  // slot -> query
    std::vector<uint32_t> slot_begin;  // [nq+1]
    std::vector<uint32_t> slot_list;   // slots of each query, CSR
};

// `query_batch` = the queries that scan positions [pos_begin,pos_end): emitted as tasks of <= kSegQB queries x <= kSegMaxPos positions
void add_tasks(SegJob& job, const uint32_t* queries, size_t nqueries, uint64_t pos_begin, uint64_t pos_end) {
    if (pos_begin >= pos_end) return;
    for (size_t q0 = 0; q0 < nqueries; q0 += kSegQB) {
        const uint32_t cnt = (uint32_t) std::min<size_t>(kSegQB, nqueries - q0);
        for (uint64_t p0 = pos_begin; p0 < pos_end; p0 += kSegMaxPos) {
            SegTask t;
            t.pos_begin = p0;
            t.pos_end = std::min(pos_end, p0 + kSegMaxPos);
            t.q_begin = (uint32_t) job.task_query.size();
            t.q_count = cnt;
            job.tasks.push_back(t);
            job.task_query.insert(job.task_query.end(), queries + q0, queries + q0 + cnt);
        }
    }
}

void finish_job(SegJob& job, size_t nq) {  // CSR of the slots of every query
    job.slot_begin.assign(nq + 1, 0);
    for (uint32_t q : job.task_query) job.slot_begin[q + 1]++;
    for (size_t q = 0; q < nq; ++q) job.slot_begin[q + 1] += job.slot_begin[q];
    job.slot_list.resize(job.task_query.size());
    std::vector<uint32_t> fill(job.slot_begin.begin(), job.slot_begin.end() - 1);
    for (uint32_t s = 0; s < job.task_query.size(); ++s) job.slot_list[fill[job.task_query[s]]++] = s;
}

struct SegSource {                 // where the queries come from
    const char* base;              // device
    size_t pitch;
    int storage;
    const uint32_t* rows;          // device, optional
};

template <class T>
int upload(gpann_ctx* ctx, ScopedDev& d, const std::vector<T>& v, cudaStream_t st) {
    CU_TRY(ctx, d.alloc(v.size() * sizeof(T)));
    if (!v.empty()) CU_TRY(ctx, cudaMemcpyAsync(d.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    return 0;
}

// scan + merge of one job; results to host arrays [nq,k] (+ [nq])
int seg_run(gpann_ctx* ctx, const gpann_points* scanned, const uint32_t* d_row_map, int id_is_row, uint32_t id_base, const SegSource& qsrc,
            SegJob& job, size_t nq, int k, int metric, int exclude_self, const uint32_t* d_id_remap, float* out_dist, uint32_t* out_id,
            uint32_t* out_count, cudaStream_t st) {
    if (nq == 0) return 0;
    if (job.task_query.size() > 0xFFFFFFF0ull) return fail(ctx, GPANN_E_ARG, "segmented scan: too many (query, segment) pairs for one call");
    finish_job(job, nq);
    ScopedDev d_tasks, d_tq, d_sb, d_sl, d_partial, d_dist, d_id, d_cnt;
    GP_TRY(upload(ctx, d_tasks, job.tasks, st));
    GP_TRY(upload(ctx, d_tq, job.task_query, st));
    GP_TRY(upload(ctx, d_sb, job.slot_begin, st));
    GP_TRY(upload(ctx, d_sl, job.slot_list, st));
    CU_TRY(ctx, d_partial.alloc(job.task_query.size() * (size_t) k * 8));
    CU_TRY(ctx, d_dist.alloc(nq * (size_t) k * 4));
    CU_TRY(ctx, d_id.alloc(nq * (size_t) k * 4));
    CU_TRY(ctx, d_cnt.alloc(nq * 4));
    SegParams p;
    std::memset(&p, 0, sizeof(p));
    p.rows = (const char*) scanned->rows;
    p.pitch = scanned->pitch;
    p.storage = scanned->storage;
    p.d = (int) scanned->d;
    p.row_map = d_row_map;
    p.id_is_row = id_is_row;
    p.id_base = id_base;
    p.pos_id_origin = 0;
    p.q_base = qsrc.base;
    p.q_pitch = qsrc.pitch;
    p.q_storage = qsrc.storage;
    p.query_rows = qsrc.rows;
    p.task_query = d_tq.as<uint32_t>();
    p.tasks = d_tasks.as<SegTask>();
    p.ntasks = (int) job.tasks.size();
    p.k = k;
    p.metric = metric;
    p.exclude_self = exclude_self;
    p.partial = d_partial.as<uint64_t>();
    const int elem = scanned->storage == GPANN_STORAGE_F32 ? 4 : 1;
    const int row_bytes = p.d * elem;
    p.spitch = odd16_pitch(row_bytes);
    p.qpitch = (p.d + 3) & ~3;
    p.cap = pow2_ceil(k + kStreamTile);
    p.vec_ok = ((uintptr_t) scanned->rows % 16 == 0) && (scanned->pitch % 16 == 0) && (row_bytes % 16 == 0);
    const size_t smem = (size_t) kStreamTile * p.spitch + (size_t) kSegQB * ((size_t) p.qpitch * 4 + (size_t) p.cap * 8 + 8 + 4 + 4) +
                        (size_t) kStreamTile * 4 + 64;
    if (smem > 200 * 1024) return fail(ctx, GPANN_E_UNSUPPORTED, "segmented scan: d or k too large for the shared-memory tile");
    if (p.ntasks > 0) {
        const int occ = (int) std::max<size_t>(1, std::min<size_t>(8, (220 * 1024) / (smem + 1024)));
        const unsigned grid = (unsigned) std::min<size_t>((size_t) p.ntasks, (size_t) ctx->num_sms * occ * 4);
        launch_seg_scan(p, grid, smem, st);
        GP_TRY(check_launch(ctx, "seg_scan_kernel"));
    }
    SegMergeParams m;
    std::memset(&m, 0, sizeof(m));
    m.keys = p.partial;
    m.slot_begin = d_sb.as<uint32_t>();
    m.slot_list = d_sl.as<uint32_t>();
    m.nq = nq;
    m.k = k;
    m.cap = std::max(64, pow2_ceil(2 * k));
    m.id_remap = d_id_remap;
    m.out_dist = d_dist.as<float>();
    m.out_id = d_id.as<uint32_t>();
    m.out_count = d_cnt.as<uint32_t>();
    launch_seg_merge(m, st);
    GP_TRY(check_launch(ctx, "seg_merge_kernel"));
    if (out_dist) CU_TRY(ctx, cudaMemcpyAsync(out_dist, m.out_dist, nq * (size_t) k * 4, cudaMemcpyDeviceToHost, st));
    if (out_id) CU_TRY(ctx, cudaMemcpyAsync(out_id, m.out_id, nq * (size_t) k * 4, cudaMemcpyDeviceToHost, st));
    if (out_count) CU_TRY(ctx, cudaMemcpyAsync(out_count, m.out_count, nq * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    ctx->last_algo = "segmented_scan";
    return 0;
}

int check_common(gpann_ctx* ctx, const gpann_points* pts, int k, int metric, const char* what) {
    if (!pts || pts->ctx != ctx || !pts->shards.empty()) return fail(ctx, GPANN_E_ARG, std::string(what) + ": point set does not belong to this (single-device) context");
    if (k < 1 || k > 512) return fail(ctx, GPANN_E_ARG, std::string(what) + ": k must be in [1,512]");
    if (metric != GPANN_METRIC_L2 && metric != GPANN_METRIC_MIPS) return fail(ctx, GPANN_E_ARG, std::string(what) + ": bad metric");
    return 0;
}
}  // namespace

int sd_ivf_topk(gpann_ctx* ctx, const gpann_points* pts, const uint64_t* bucket_offsets, size_t num_buckets, const uint32_t* permutation,
                const float* queries, size_t nq, const int32_t* probes, int nprobe, int k, int metric, float* out_dist, uint32_t* out_id,
                uint32_t* out_count) {
    GP_TRY(check_common(ctx, pts, k, metric, "ivf_topk"));
    if (!bucket_offsets || !queries || !probes || nprobe < 1) return fail(ctx, GPANN_E_ARG, "ivf_topk: null argument");
    if (nq == 0) return 0;
    const uint64_t npos = bucket_offsets[num_buckets];
    if (npos > 0xFFFFFFFFull) return fail(ctx, GPANN_E_ARG, "ivf_topk: positions exceed 32 bits");
    if (!permutation && npos > pts->n) return fail(ctx, GPANN_E_ARG, "ivf_topk: offsets exceed the (clustered) point set");
    for (size_t b = 0; b < num_buckets; ++b)
        if (bucket_offsets[b] > bucket_offsets[b + 1]) return fail(ctx, GPANN_E_ARG, "ivf_topk: offsets must be non-decreasing");
    // (bucket, query) pairs grouped by bucket: a counting sort; a bucket listed twice for a query is probed once
    std::vector<uint32_t> count(num_buckets + 1, 0);
    std::vector<std::pair<int32_t, uint32_t>> pairs;
    pairs.reserve(nq * (size_t) nprobe);
    std::vector<int32_t> mine((size_t) nprobe);
    for (size_t q = 0; q < nq; ++q) {
        std::copy(probes + q * nprobe, probes + (q + 1) * nprobe, mine.begin());
        std::sort(mine.begin(), mine.end());
        for (int j = 0; j < nprobe; ++j) {
            const int32_t b = mine[j];
            if (b < 0 || (j > 0 && b == mine[j - 1])) continue;
            if ((size_t) b >= num_buckets) return fail(ctx, GPANN_E_ARG, "ivf_topk: probe out of range");
            pairs.emplace_back(b, (uint32_t) q);
            count[(size_t) b + 1]++;
        }
    }
    for (size_t b = 0; b < num_buckets; ++b) count[b + 1] += count[b];
    std::vector<uint32_t> by_bucket(pairs.size());
    {
        std::vector<uint32_t> fill(count.begin(), count.end() - 1);
        for (const auto& pr : pairs) by_bucket[fill[(size_t) pr.first]++] = pr.second;
    }
    SegJob job;
    for (size_t b = 0; b < num_buckets; ++b)
        add_tasks(job, by_bucket.data() + count[b], count[b + 1] - count[b], bucket_offsets[b], bucket_offsets[b + 1]);
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    OrderScope order(ctx, st);
    ScopedDev d_perm, d_q;
    if (permutation) {
        for (uint64_t i = 0; i < npos; ++i)
            if (permutation[i] >= pts->n) return fail(ctx, GPANN_E_ARG, "ivf_topk: permutation entry out of range");
        CU_TRY(ctx, d_perm.alloc(npos * 4));
        CU_TRY(ctx, cudaMemcpyAsync(d_perm.p, permutation, npos * 4, cudaMemcpyHostToDevice, st));
    }
    CU_TRY(ctx, d_q.alloc(nq * pts->d * 4));
    CU_TRY(ctx, cudaMemcpyAsync(d_q.p, queries, nq * pts->d * 4, cudaMemcpyHostToDevice, st));
    const SegSource qsrc = { d_q.as<char>(), pts->d * 4, GPANN_STORAGE_F32, nullptr };
    // ranking id = POSITION (ties keep the smaller position, src/inverted_index.h:49-51); reported id = permutation[position]
    return seg_run(ctx, pts, d_perm.as<uint32_t>(), 0, 0, qsrc, job, nq, k, metric, 0, d_perm.as<uint32_t>(), out_dist, out_id, out_count, st);
}

int sd_knn_buckets(gpann_ctx* ctx, const gpann_points* pts, const uint64_t* bucket_offsets, const uint32_t* bucket_members, size_t num_buckets,
                   int k, int metric, float* out_dist, uint32_t* out_id, uint32_t* out_count) {
    GP_TRY(check_common(ctx, pts, k, metric, "knn_buckets"));
    if (!bucket_offsets || (!bucket_members && bucket_offsets[num_buckets])) return fail(ctx, GPANN_E_ARG, "knn_buckets: null argument");
    const uint64_t total = bucket_offsets[num_buckets];
    if (total == 0) return 0;
    if (total > 0xFFFFFFF0ull) return fail(ctx, GPANN_E_ARG, "knn_buckets: too many members for one call");
    for (uint64_t i = 0; i < total; ++i)
        if (bucket_members[i] >= pts->n) return fail(ctx, GPANN_E_ARG, "knn_buckets: member id out of range");
    SegJob job;
    std::vector<uint32_t> ids;
    for (size_t b = 0; b < num_buckets; ++b) {
        if (bucket_offsets[b] > bucket_offsets[b + 1]) return fail(ctx, GPANN_E_ARG, "knn_buckets: offsets must be non-decreasing");
        ids.resize(bucket_offsets[b + 1] - bucket_offsets[b]);
        std::iota(ids.begin(), ids.end(), (uint32_t) bucket_offsets[b]);  // query index = member position
        add_tasks(job, ids.data(), ids.size(), bucket_offsets[b], bucket_offsets[b + 1]);
    }
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    OrderScope order(ctx, st);
    ScopedDev d_members;
    CU_TRY(ctx, d_members.alloc(total * 4));
    CU_TRY(ctx, cudaMemcpyAsync(d_members.p, bucket_members, total * 4, cudaMemcpyHostToDevice, st));
    // queries are rows of the point set (member ids), candidates are ranked by (distance, global id), a point is not its own neighbour
    const SegSource qsrc = { (const char*) pts->rows, pts->pitch, pts->storage, d_members.as<uint32_t>() };
    return seg_run(ctx, pts, d_members.as<uint32_t>(), 1, pts->id_base, qsrc, job, (size_t) total, k, metric, 1, nullptr, out_dist, out_id,
                   out_count, st);
}

int sd_closest_leaders_batch(gpann_ctx* ctx, const gpann_points* pts, const uint32_t* point_ids, size_t m, const gpann_points* leaders, int k,
                             int metric, float* out_dist, uint32_t* out_id, uint32_t* out_count) {
    GP_TRY(check_common(ctx, pts, k, metric, "closest_leaders_batch"));
    GP_TRY(check_common(ctx, leaders, k, metric, "closest_leaders_batch"));
    if (pts->d != leaders->d) return fail(ctx, GPANN_E_ARG, "closest_leaders_batch: dimensions differ");
    if (!point_ids) m = pts->n;
    if (m == 0) return 0;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    OrderScope order(ctx, st);
    const size_t chunk = (size_t) 1 << 21;  // queries per round: bounds the task list (host + device)
    std::vector<uint32_t> ids;
    for (size_t q0 = 0; q0 < m; q0 += chunk) {
        const size_t cnt = std::min(chunk, m - q0);
        ids.resize(cnt);
        std::iota(ids.begin(), ids.end(), 0u);  // query index inside the round
        SegJob job;
        add_tasks(job, ids.data(), cnt, 0, leaders->n);
        ScopedDev d_rows;
        const uint32_t* d_qrows = nullptr;
        size_t row0 = q0;
        if (point_ids) {
            for (size_t i = 0; i < cnt; ++i)
                if (point_ids[q0 + i] >= pts->n) return fail(ctx, GPANN_E_ARG, "closest_leaders_batch: point id out of range");
            CU_TRY(ctx, d_rows.alloc(cnt * 4));
            CU_TRY(ctx, cudaMemcpyAsync(d_rows.p, point_ids + q0, cnt * 4, cudaMemcpyHostToDevice, st));
            d_qrows = d_rows.as<uint32_t>();
            row0 = 0;
        }
        // ids = leader index (src/defs.cpp:92-93); the leaders are a separate set, so a point may be its own leader
        const SegSource qsrc = { (const char*) pts->rows + row0 * pts->pitch, pts->pitch, pts->storage, d_qrows };
        GP_TRY(seg_run(ctx, leaders, nullptr, 1, 0, qsrc, job, cnt, k, metric, 0, nullptr, out_dist ? out_dist + q0 * k : nullptr,
                       out_id ? out_id + q0 * k : nullptr, out_count ? out_count + q0 : nullptr, st));
    }
    return 0;
}

int sd_group_min_dist(gpann_ctx* ctx, const gpann_points* pts, const int32_t* group, int ngroups, const float* queries, size_t nq, int metric,
                      float* out) {
    GP_TRY(check_common(ctx, pts, 1, metric, "group_min_dist"));
    if (!group || !queries || !out || ngroups < 1) return fail(ctx, GPANN_E_ARG, "group_min_dist: bad argument");
    if (nq == 0) return 0;
    if (ngroups > 12288) return fail(ctx, GPANN_E_UNSUPPORTED, "group_min_dist: at most 12288 groups (per-CTA table in shared memory)");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    OrderScope order(ctx, st);
    const size_t n = pts->n, d = pts->d;
    ScopedDev d_group, d_q, d_out;
    CU_TRY(ctx, d_group.alloc(n * 4));
    CU_TRY(ctx, d_q.alloc(nq * d * 4));
    CU_TRY(ctx, d_out.alloc(nq * (size_t) ngroups * 4));
    if (n) CU_TRY(ctx, cudaMemcpyAsync(d_group.p, group, n * 4, cudaMemcpyHostToDevice, st));
    CU_TRY(ctx, cudaMemcpyAsync(d_q.p, queries, nq * d * 4, cudaMemcpyHostToDevice, st));
    // "no point in this group": the order-preserving image of FLT_MAX, as the reference's min_dist starts (FLT_MAX)
    std::vector<uint32_t> init(nq * (size_t) ngroups, float_to_ordered(3.402823466e+38f));
    CU_TRY(ctx, cudaMemcpyAsync(d_out.p, init.data(), init.size() * 4, cudaMemcpyHostToDevice, st));
    GroupMinParams p;
    std::memset(&p, 0, sizeof(p));
    p.rows = (const char*) pts->rows;
    p.pitch = pts->pitch;
    p.n = n;
    p.storage = pts->storage;
    p.d = (int) d;
    p.group = d_group.as<int32_t>();
    p.ngroups = ngroups;
    p.queries = d_q.as<float>();
    p.out_ordered = d_out.as<uint32_t>();
    p.nq = (int) nq;
    p.metric = metric;
    const int elem = pts->storage == GPANN_STORAGE_F32 ? 4 : 1;
    const int row_bytes = p.d * elem;
    p.spitch = odd16_pitch(row_bytes);
    p.qpitch = (p.d + 3) & ~3;
    p.vec_ok = ((uintptr_t) pts->rows % 16 == 0) && (pts->pitch % 16 == 0) && (row_bytes % 16 == 0);
    const size_t tile_bytes = (size_t) kStreamTile * p.spitch;
    const size_t per_q = (size_t) p.qpitch * 4 + (size_t) ngroups * 4;
    if (tile_bytes + per_q > 190 * 1024) return fail(ctx, GPANN_E_UNSUPPORTED, "group_min_dist: d too large for the shared-memory tile");
    p.qb = (int) std::max<size_t>(1, std::min<size_t>(std::min<size_t>(nq, 16), (96 * 1024) / per_q));
    const size_t smem = tile_bytes + (size_t) p.qb * per_q;
    if (n) {
        const size_t ntiles = (n + kStreamTile - 1) / kStreamTile;
        const int occ = (int) std::max<size_t>(1, std::min<size_t>(8, (220 * 1024) / (smem + 1024)));
        const unsigned gy = (unsigned) std::min<size_t>((nq + p.qb - 1) / p.qb, 1024);
        const unsigned gx = (unsigned) std::max<size_t>(1, std::min<size_t>(ntiles, std::max<size_t>(1, (size_t) ctx->num_sms * occ / gy)));
        launch_group_min(p, dim3(gx, gy), smem, st);
        GP_TRY(check_launch(ctx, "group_min_kernel"));
    }
    launch_group_min_finish(d_out.as<uint32_t>(), nq * (size_t) ngroups, blocks_for(ctx, nq * (size_t) ngroups), st);
    GP_TRY(check_launch(ctx, "group_min_finish_kernel"));
    CU_TRY(ctx, cudaMemcpyAsync(out, d_out.p, nq * (size_t) ngroups * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    ctx->last_algo = "group_min";
    return 0;
}

}  // namespace gpimpl
