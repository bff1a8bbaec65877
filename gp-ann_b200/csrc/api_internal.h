// api_internal.h -- handles (gpann_ctx, gpann_points) and host-side helpers shared by the translation units that
// implement the C ABI of include/gpann.h: gpann_api.cu (single-device entry points) and multi.cu (NCCL: sharded
// contexts over several devices or several processes).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <new>
#include <string>
#include <vector>

#include "../../include/gpann.h"
#include "common.cuh"
#include "params.h"

// ---- handles ---------------------------------------------------------------------------------------------
namespace gpimpl { struct Group; }
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct gpann_ctx {
    int device = 0;
    int num_sms = 0;
    int algo = GPANN_ALGO_AUTO;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    const char* last_algo = "none";
    EncodeTiledFn encode_tiled = nullptr;
    int tc_mode = 0;     // GPANN_TC_MODE env: 1 = probe (MMA + TMEM loads + max), 2 = probe (MMA issue only)
    int tc_prepass = 1;  // GPANN_TC_PREPASS env: 0 disables the sample pre-pass (cold-start thresholds)
    int tc_warm = 0;     // GPANN_TC_WARM env (probe): keep the previous call's k-th distances as the next call's bounds
    bool tc_warm_valid = false;
    int timing = 0;
    bool timed = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // grow-only scratch, reused across calls (a ctx is externally synchronised)
    DevBuf pace;
    DevBuf flog;  // find logs of the tensor top-k kernel: [G] tails (256-byte header block per 64) + [G][cap] entries
    DevBuf qstat; // order-statistic bounds of the point splits: [G][S][R*128]
    DevBuf partial, cand, qrows, qnorm, qf32, bound, flags, dist_out, id_out, cnt_out, qdev, misc, sort_tmp, sort_a, sort_b, sort_c,
        sort_d, segsums, cent, sums, counts, assign, km_cent, km_hn, km_scal, km_partial, km_amb_list, km_amb_info;
    int km_mode = 0;     // GPANN_KM_MODE env: 2 = probe (MMA issue only, results invalid)
    // streaming entry points: second stream for the H2D copies, double-buffered raw chunks, one chunk layout
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = { nullptr, nullptr }, ev_consumed[2] = { nullptr, nullptr };
    DevBuf st_raw[2], st_rows, st_norms, st_ids, st_a, st_b, st_c, st_d, st_tmp, st_narrow;
    // Stream ordering of the shared scratch: every entry point makes its stream wait for ev_order (recorded at the end of
    // the previous entry point on whatever stream that one used), so a returned asynchronous call still owns the scratch
    // buffers until its work is done and the next call -- on any stream -- queues behind it.
    cudaEvent_t ev_order = nullptr;
    bool gate_used = false;        // the last top-k call queued a gated float fall-back behind the byte tensor path
    // ---- NCCL (multi.cu): this ctx is one rank of a communicator (several processes: gpann_comm_init_rank; several
    // devices of one process: a child of a gpann_init_multi group) ----
    void* comm = nullptr;          // ncclComm_t
    int comm_rank = 0, comm_size = 1;
    // peer-shared bounds of the sharded top-k: a fixed-size array (one slot per query of the largest batch) that the other
    // ranks' kernels tighten with system-scope atomicMin over NVLink (cudaIpc between processes, direct peer access inside
    // a group ctx); peer_bound[i] = the other ranks' arrays mapped here.  peer_share is raised by topk_sharded for the
    // calls that may use them (one query batch per call, see there).
    uint32_t* shared_bound = nullptr;
    size_t shared_bound_slots = 0;
    uint32_t* peer_bound[gpann::tc::kMaxPeers] = {};
    int npeers = 0;
    bool peer_ipc = false;         // peer_bound[] came from cudaIpcOpenMemHandle (close on destroy)
    bool peer_share = false;
    DevBuf gather_send, gather_recv, km_run_cent, km_run_sums, km_run_counts, km_run_alive, km_run_remap, km_run_norms;
    // ---- group root (gpann_init_multi): owns one child ctx + worker thread per device; the root itself launches nothing
    gpimpl::Group* group = nullptr;
    gpann_ctx* parent = nullptr;   // child of a group
};

struct gpann_points {
    gpann_ctx* ctx = nullptr;
    size_t n = 0, d = 0;
    int storage = GPANN_STORAGE_F32;
    uint32_t id_base = 0;
    // rows in original order: float [n,d] or bytes [n,d]
    void* rows = nullptr;
    size_t pitch = 0;
    bool owns_rows = true;
    // norm-sorted tensor layout for byte storage with d <= 128 (topk_tc.cu)
    unsigned char* t_rows = nullptr;
    int32_t* t_norms = nullptr;
    uint32_t* t_ids = nullptr;
    size_t n_pad = 0;
    CUtensorMap tmap;
    bool has_tensor = false;
    // float storage, d <= 128: norm-sorted bf16 rows [n_pad, dp] + fp32 norms + ids, built lazily by the first
    // tensor-path query (t_rows / t_norms / t_ids / tmap above are reused for it)
    int dp = 0;               // padded row length in elements (multiple of 64)
    bool tensor_tried = false;
    // float storage, d <= 128: split-precision rows for the tensor k-means assignment (kmeans_tc.cu), built lazily
    unsigned short* km_rows = nullptr;  // [km_npad, 2*km_dpad] bf16: hi block, lo block
    float* km_xnorm = nullptr;          // [km_npad] |x| rounded up; [km_npad] holds the maximum
    size_t km_npad = 0;
    int km_dpad = 0;
    CUtensorMap km_tmap;
    bool km_tried = false, km_ok = false;
    // a point set of a group ctx (gpann_init_multi): rows [shard_lo[i], shard_lo[i+1]) live on device i as shards[i]
    std::vector<gpann_points*> shards;
    std::vector<size_t> shard_lo;
};

namespace gpimpl {
using namespace gpann;

// Hook of the sharded entry points: called between the sample passes and the main pass of every query batch with the
// per-query bounds [count] (float bits, +inf = none) -- multi.cu all-reduces them (min) across the ranks.
typedef std::function<int(uint32_t* bound, size_t count, cudaStream_t st)> BetweenFn;

inline int fail(gpann_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

#define CU_TRY(ctx, expr)                                                                                      \
    do {                                                                                                       \
        cudaError_t e__ = (expr);                                                                              \
        if (e__ != cudaSuccess) {                                                                              \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? GPANN_E_OOM : GPANN_E_CUDA,                    \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));                                  \
        }                                                                                                      \
    } while (0)

#define GP_TRY(expr)             \
    do {                         \
        int r__ = (expr);        \
        if (r__ != 0) return r__; \
    } while (0)

inline int reserve(gpann_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return 0;
    if (b.p) CU_TRY(ctx, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    const size_t want = std::max(bytes, (size_t) 256);
    CU_TRY(ctx, cudaMalloc(&b.p, want));
    b.cap = want;
    return 0;
}
inline void release(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

inline int odd16_pitch(int row_bytes) {
    int p = (row_bytes + 15) & ~15;
    if (((p >> 4) & 1) == 0) p += 16;
    return p;
}

inline int check_launch(gpann_ctx* ctx, const char* what, int count = 1) {
    ctx->launches += count;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, GPANN_E_CUDA, std::string(what) + " launch: " + cudaGetErrorString(e));
    return 0;
}

// Orders an entry point's work on `st` behind the previous entry point of the same ctx (see gpann_ctx::ev_order).
struct OrderScope {
    gpann_ctx* ctx;
    cudaStream_t st;
    OrderScope(gpann_ctx* c, cudaStream_t s) : ctx(c), st(s) {
        if (ctx->ev_order) cudaStreamWaitEvent(st, ctx->ev_order, 0);
    }
    ~OrderScope() {
        if (ctx->ev_order) cudaEventRecord(ctx->ev_order, st);
    }
    OrderScope(const OrderScope&) = delete;
    OrderScope& operator=(const OrderScope&) = delete;
};

// Temporary device allocation freed on scope exit (error paths included).
struct ScopedDev {
    void* p = nullptr;
    ~ScopedDev() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes < 256 ? 256 : bytes); }
    template <class T> T* as() const { return (T*) p; }
    void* release() { void* r = p; p = nullptr; return r; }
};

inline int log2_ceil(size_t x) {
    int b = 0;
    while (((size_t) 1 << b) < x) ++b;
    return b;
}

inline int blocks_for(gpann_ctx* ctx, size_t items, int per_block = 256, int waves = 16) {
    return (int) std::max<size_t>(1, std::min<size_t>((items + per_block - 1) / per_block, (size_t) ctx->num_sms * waves));
}

// ---- shared between gpann_api.cu, multi.cu and abi.cu ------------------------------------------------------------------
struct TopkOut {
    uint64_t* keys = nullptr;  // device [nq,k]
    float* dist = nullptr;     // device [nq,k]
    uint32_t* id = nullptr;    // device [nq,k]
    uint32_t* count = nullptr; // device [nq]
    float* kth = nullptr;      // device [nq]
    int kth_full_only = 0;
};

enum HostKind { kHostU8 = 0, kHostS8 = 1, kHostF32 = 2 };

int launch_merge(gpann_ctx* ctx, const uint64_t* keys, size_t parts, size_t nq_stride, size_t nq, int k, const TopkOut& o,
                 size_t out_q_offset, cudaStream_t st, const int* gate, int gate_want);
int topk_sharded(gpann_ctx* ctx, const gpann_points* pts, size_t row_begin, size_t row_end, const float* dq, size_t nq, int k,
                 int metric, const TopkOut& out, cudaStream_t st);
int topk_stream_host(gpann_ctx* ctx, const void* host_rows, int host_kind, int storage, size_t n, size_t d, uint32_t id_base,
                     const float* dq, size_t nq, int k, size_t chunk_rows, const TopkOut& out, cudaStream_t st);
int kmeans_assign_device(gpann_ctx* ctx, const gpann_points* pts, const float* d_centroids, size_t kc, int metric, int32_t* d_assign,
                         float* d_mindist, const unsigned char* d_alive, cudaStream_t st);
int kmeans_rounds_device(gpann_ctx* ctx, const gpann_points* pts, float* d_cent, size_t kc, int rounds, int metric, int32_t* d_assign,
                         unsigned char* d_alive, uint64_t* d_counts, cudaStream_t st);
int kmeans_run_enqueue(gpann_ctx* ctx, const gpann_points* pts, const float* centroids, size_t kc, int rounds, int metric);
int kmeans_run_fetch(gpann_ctx* ctx, const gpann_points* pts, float* centroids_out, size_t kc, int32_t* out_assign, uint64_t* out_sizes,
                     size_t* out_kc);
int objective_device(gpann_ctx* ctx, const gpann_points* pts, const float* d_centroids, size_t kc, const int32_t* d_assign, int metric,
                     double** d_out_ptr, cudaStream_t st);
RowsView rows_view(const gpann_points* pts);
int cluster_sums_device(gpann_ctx* ctx, const gpann_points* pts, const int32_t* d_assign, size_t kc, int normalize, float* d_out,
                        uint64_t* d_counts, cudaStream_t st, const float* d_row_norm = nullptr, float* d_norm_sums = nullptr);
int kmeans_compact_device(gpann_ctx* ctx, float* d_cent, size_t kc, size_t d, int32_t* d_assign, size_t n, const unsigned char* d_alive,
                          uint64_t* d_counts, int* d_kc_out, cudaStream_t st);
int objective_enqueue(gpann_ctx* ctx, const gpann_points* pts, const float* centroids, size_t kc, const int32_t* assign, int metric,
                      double** d_out_ptr);

// single-device implementations of the entry points of include/gpann.h (gpann_api.cu)
int sd_init(gpann_ctx** out, int device);
void sd_destroy(gpann_ctx* ctx);
int sd_points_upload(gpann_ctx* ctx, const float* host_rows, size_t n, size_t d, int storage, uint32_t id_base, gpann_points** out);
int sd_points_upload_bytes(gpann_ctx* ctx, const void* host_rows, size_t n, size_t d, int is_signed, uint32_t id_base, gpann_points** out);
int sd_points_from_device(gpann_ctx* ctx, const void* dev_rows, size_t n, size_t d, int storage, uint32_t id_base, void* stream,
                          gpann_points** out);
void sd_points_free(gpann_points* pts);
int sd_topk(gpann_ctx* ctx, const gpann_points* pts, const float* queries, size_t nq, int k, int metric, float* out_dist, uint32_t* out_id,
            uint32_t* out_count);
int sd_kth_distance(gpann_ctx* ctx, const gpann_points* pts, const float* queries, size_t nq, int k, int metric, float* out_kth);
int sd_topk_range(gpann_ctx* ctx, const gpann_points* pts, size_t row_begin, size_t row_end, const float* queries, size_t nq, int k,
                  int metric, float* out_dist, uint32_t* out_id, uint32_t* out_count);
int sd_topk_stream_bytes(gpann_ctx* ctx, const void* host_rows, size_t n, size_t d, int is_signed, uint32_t id_base, const float* queries,
                         size_t nq, int k, size_t chunk_rows, float* out_dist, uint32_t* out_id, uint32_t* out_count);
int sd_topk_stream(gpann_ctx* ctx, const float* host_rows, size_t n, size_t d, int storage, uint32_t id_base, const float* queries, size_t nq,
                   int k, size_t chunk_rows, float* out_dist, uint32_t* out_id, uint32_t* out_count, float* out_kth);
int sd_topk_stream_bytes_device(gpann_ctx* ctx, const void* host_rows, size_t n, size_t d, int is_signed, uint32_t id_base,
                                const float* dev_queries, size_t nq, int k, size_t chunk_rows, uint64_t* dev_out_keys, void* stream);
int sd_topk_device(gpann_ctx* ctx, const gpann_points* pts, const float* dev_queries, size_t nq, int k, int metric, uint64_t* dev_out_keys,
                   void* stream);
int sd_topk_sharded_device(gpann_ctx* ctx, const gpann_points* pts, const float* dev_queries, size_t nq, int k, int metric,
                           float* dev_out_dist, uint32_t* dev_out_id, uint64_t* dev_out_keys, void* stream);
int sd_merge_keys_device(gpann_ctx* ctx, const uint64_t* dev_keys, size_t parts, size_t nq, int k, float* dev_out_dist, uint32_t* dev_out_id,
                         uint64_t* dev_out_keys, void* stream);
int sd_kmeans_assign(gpann_ctx* ctx, const gpann_points* pts, const float* centroids, size_t kc, int metric, int32_t* out_assign,
                     float* out_mindist);
int sd_kmeans_partial_device(gpann_ctx* ctx, const gpann_points* pts, const float* dev_centroids, size_t kc, int metric,
                             int32_t* dev_out_assign, float* dev_out_sums, uint64_t* dev_out_counts, void* stream);
int sd_kmeans_step(gpann_ctx* ctx, const gpann_points* pts, float* centroids_inout, size_t kc, int metric, int32_t* out_assign,
                   uint64_t* out_sizes);
int sd_kmeans_run(gpann_ctx* ctx, const gpann_points* pts, float* centroids_inout, size_t kc, int rounds, int metric, int32_t* out_assign,
                  uint64_t* out_sizes, size_t* out_kc);
int sd_kmeans_run_device(gpann_ctx* ctx, const gpann_points* pts, float* dev_centroids_inout, size_t kc, int rounds, int metric,
                         int32_t* dev_out_assign, unsigned char* dev_alive_inout, uint64_t* dev_out_counts, void* stream);
int sd_objective(gpann_ctx* ctx, const gpann_points* pts, const float* centroids, size_t kc, const int32_t* assign, int metric, double* out);
int sd_ivf_topk(gpann_ctx* ctx, const gpann_points* pts, const uint64_t* bucket_offsets, size_t num_buckets, const uint32_t* permutation,
                const float* queries, size_t nq, const int32_t* probes, int nprobe, int k, int metric, float* out_dist, uint32_t* out_id,
                uint32_t* out_count);
int sd_knn_buckets(gpann_ctx* ctx, const gpann_points* pts, const uint64_t* bucket_offsets, const uint32_t* bucket_members, size_t num_buckets,
                   int k, int metric, float* out_dist, uint32_t* out_id, uint32_t* out_count);
int sd_closest_leaders_batch(gpann_ctx* ctx, const gpann_points* pts, const uint32_t* point_ids, size_t m, const gpann_points* leaders, int k,
                             int metric, float* out_dist, uint32_t* out_id, uint32_t* out_count);
int sd_group_min_dist(gpann_ctx* ctx, const gpann_points* pts, const int32_t* group, int ngroups, const float* queries, size_t nq, int metric,
                      float* out);
int sd_balanced_kmeans(gpann_ctx* ctx, const gpann_points* pts, float* centroids_inout, size_t kc, size_t max_cluster_size, int metric,
                       int32_t* out_assign, uint64_t* out_sizes, size_t* out_kc);
int sd_set_algo(gpann_ctx* ctx, int algo);
int sd_set_timing(gpann_ctx* ctx, int on);
float sd_last_kernel_ms(gpann_ctx* ctx);
uint64_t sd_last_ambiguous(gpann_ctx* ctx);

// NCCL plumbing (multi.cu).  libnccl is loaded on first use (dlopen: the copy torch already mapped, else the system's).
void comm_destroy(gpann_ctx* ctx);
int comm_allreduce_min_u32(gpann_ctx* ctx, uint32_t* buf, size_t count, cudaStream_t st);
int comm_allgather_u64(gpann_ctx* ctx, const uint64_t* send, uint64_t* recv, size_t count, cudaStream_t st);
int comm_allreduce_kmeans(gpann_ctx* ctx, float* sums, size_t nsums, uint64_t* counts, size_t ncounts, cudaStream_t st);

int comm_unique_id(void* out, size_t bytes);
int comm_init_rank(gpann_ctx* ctx, int nranks, int rank, const void* id_bytes, size_t bytes);

// group contexts (gpann_init_multi): one child ctx + worker thread per device (multi.cu)
int mg_init(gpann_ctx** out, const int* devices, int ndev);
void mg_destroy(gpann_ctx* root);
int mg_num_devices(const gpann_ctx* root);
gpann_ctx* mg_rank_ctx(const gpann_ctx* root, int i);
int mg_for_each(gpann_ctx* root, const std::function<int(gpann_ctx*)>& fn);
int mg_points_upload(gpann_ctx* root, const float* host_rows, size_t n, size_t d, int storage, uint32_t id_base, gpann_points** out);
int mg_points_upload_bytes(gpann_ctx* root, const void* host_rows, size_t n, size_t d, int is_signed, uint32_t id_base, gpann_points** out);
void mg_points_free(gpann_points* pts);
int mg_topk_host(gpann_ctx* root, const gpann_points* pts, size_t row_begin, size_t row_end, const float* queries, size_t nq, int k,
                 int metric, float* out_dist, uint32_t* out_id, uint32_t* out_count, float* out_kth);
int mg_topk_stream_host(gpann_ctx* root, const void* host_rows, int host_kind, int storage, size_t n, size_t d, uint32_t id_base,
                        const float* queries, size_t nq, int k, size_t chunk_rows, float* out_dist, uint32_t* out_id, uint32_t* out_count,
                        float* out_kth);
int mg_kmeans_assign(gpann_ctx* root, const gpann_points* pts, const float* centroids, size_t kc, int metric, int32_t* out_assign,
                     float* out_mindist);
int mg_kmeans_step(gpann_ctx* root, const gpann_points* pts, float* centroids_inout, size_t kc, int metric, int32_t* out_assign,
                   uint64_t* out_sizes);
int mg_kmeans_run(gpann_ctx* root, const gpann_points* pts, float* centroids_inout, size_t kc, int rounds, int metric, int32_t* out_assign,
                  uint64_t* out_sizes, size_t* out_kc);
int mg_objective(gpann_ctx* root, const gpann_points* pts, const float* centroids, size_t kc, const int32_t* assign, int metric, double* out);

}  // namespace gpimpl
