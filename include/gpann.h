/* gpann.h -- C ABI of libgpann_cuda.so: gp-ann's brute-force distance + top-k hot path on B200 (sm_100a).
 *
 * The reference (larsgottesbueren/gp-ann) has no plugin / FFI layer: the path sits behind plain C++ free
 * functions (src/dist.h:3-15, src/topn.h:6-32, src/recall.h:10,25, src/kmeans.h:5-8, src/defs.h:50).  The
 * drop-in C++ headers in gp-ann_b200/cxx/ keep those signatures and forward the batched ones to the entry
 * points below; INTEGRATION.md shows the binding.  Each entry point names the reference interface it
 * replaces (file:line relative to the reference repository root).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no exceptions and no STL across the boundary.
 *  - One gpann_ctx per CUDA device (gpann_init), or one GROUP ctx over several devices of this process (gpann_init_multi:
 *    point sets are sharded by rows, every entry point below that takes host buffers works on it unchanged, the exchanges
 *    are NCCL collectives inside the library).  A ctx is externally synchronised: one call at a time.  A call that returned
 *    while its work is still queued (the *_device entry points) keeps owning the ctx's scratch buffers; the next call on
 *    the same ctx -- on any stream -- is ordered behind it by the library (a stream-wait on an event, no host wait).
 *  - Every function returns 0 on success or a negative GPANN_E_* code; gpann_last_error(ctx) has the text.
 *  - Host entry points borrow caller-owned host buffers for the duration of the call; the library owns all
 *    device memory behind the opaque handles.  The *_device entry points take device pointers and a
 *    cudaStream_t (as void*) and are asynchronous on that stream; the handle is used verbatim, so NULL is
 *    the legacy default stream (what torch.cuda.current_stream().cuda_stream reports by default).
 *  - There is NO CPU fallback: without a usable CUDA device gpann_init fails with GPANN_E_CUDA.
 *  - Results are "keys": a (distance, id) pair packed into 64 bits whose unsigned order is the reference's
 *    std::pair<float,uint32_t> lexicographic order (src/topn.h:12-19): ties on distance keep the smaller id.
 */
#ifndef GPANN_H
#define GPANN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpann_ctx gpann_ctx;
typedef struct gpann_points gpann_points;

enum { GPANN_OK = 0, GPANN_E_CUDA = -1, GPANN_E_ARG = -2, GPANN_E_OOM = -3, GPANN_E_UNSUPPORTED = -4, GPANN_E_NCCL = -5 };

/* Distance metric.  The reference selects it at compile time (-DMIPS_DISTANCE, src/dist.cpp:75-81). */
enum { GPANN_METRIC_L2 = 0, GPANN_METRIC_MIPS = 1 };

/* Point storage in HBM.  AUTO narrows float rows to U8 / S8 when every coordinate is an integer in range
 * (BigANN .u8bin / .i8bin data is widened to float by the reference's reader, src/points_io.cpp:63-111). */
enum { GPANN_STORAGE_AUTO = 0, GPANN_STORAGE_F32 = 1, GPANN_STORAGE_U8 = 2, GPANN_STORAGE_S8 = 3 };

/* Which kernel family serves gpann_topk*: AUTO picks by shape; STREAM = CUDA-core HBM-streaming scan;
 * TENSOR = tcgen05/TMEM tile kernel (fails with GPANN_E_UNSUPPORTED where it does not apply). */
enum { GPANN_ALGO_AUTO = 0, GPANN_ALGO_STREAM = 1, GPANN_ALGO_TENSOR = 2 };

#define GPANN_INVALID_ID 0xFFFFFFFFu
#define GPANN_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull

/* ---- context ------------------------------------------------------------------------------------------- */
int gpann_init(gpann_ctx** out, int device);
/* One ctx over `ndev` devices of this process (the reference's entry points are single-call,
 * single-process by signature -- ComputeGroundTruth src/recall.h:25-39, KMeans src/kmeans.cpp:228-244 -- so the shards live
 * behind one handle): a worker thread, a stream and an NCCL communicator rank (ncclCommInitAll) per device.
 * gpann_points_upload* shard the rows [r*ceil(n/ndev), ...) with global ids; gpann_topk / gpann_kth_distance /
 * gpann_topk_range / gpann_topk_stream* / gpann_kmeans_assign / gpann_kmeans_step / gpann_kmeans_run / gpann_objective run
 * on all devices at once (bounds all-reduce(min) + key all-gather + merge; sums/counts all-reduce per Lloyd round).
 * The *_device entry points and gpann_points_from_device need a single-device ctx (GPANN_E_UNSUPPORTED on a group). */
int gpann_init_multi(gpann_ctx** out, const int* devices, int ndev);
int gpann_num_devices(const gpann_ctx* ctx);
/* Several PROCESSES, one single-device ctx each (torchrun): rank 0 calls gpann_comm_unique_id (128 bytes), the caller
 * ships the bytes to the other ranks, every rank calls gpann_comm_init_rank (collective: ncclCommInitRank).  Afterwards
 * gpann_topk_sharded_device and gpann_kmeans_run_device treat the ctx's point set as ONE SHARD of the whole. */
int gpann_comm_unique_id(void* out_id128);
int gpann_comm_init_rank(gpann_ctx* ctx, int nranks, int rank, const void* id128);
int gpann_comm_size(const gpann_ctx* ctx);
int gpann_comm_rank(const gpann_ctx* ctx);
void gpann_destroy(gpann_ctx* ctx);
const char* gpann_last_error(const gpann_ctx* ctx);
int gpann_device(const gpann_ctx* ctx);
int gpann_set_algo(gpann_ctx* ctx, int algo);
/* Number of kernel launches issued by this ctx since init (bench.py's gpu_launches claim). */
uint64_t gpann_launch_count(const gpann_ctx* ctx);
/* Name of the kernel family the last gpann_topk* call used ("stream", "tensor_u8", ...). */
const char* gpann_last_algo(const gpann_ctx* ctx);
/* Measurement hook: when on, the dominant kernel of each gpann_topk* call (the tcgen05 tile kernel or the
 * streaming scan) is bracketed by CUDA events on the launching stream; gpann_last_kernel_ms waits for the
 * last bracket and returns its duration in milliseconds (negative if none was recorded). */
int gpann_set_timing(gpann_ctx* ctx, int on);
float gpann_last_kernel_ms(gpann_ctx* ctx);
/* Diagnostic: number of points of the last tensor-path k-means assignment whose winner the split-precision filter could
 * not certify and that were re-ranked exactly in fp32 (waits for the stream the call ran on; 0 if none ran). */
uint64_t gpann_last_ambiguous(gpann_ctx* ctx);

/* ---- point sets: replaces PointSet (src/defs.h:10-21) as the scanned operand --------------------------- */
/* Upload n rows of d floats (row-major, as PointSet::coordinates).  Ids reported by gpann_topk are
 * id_base + row index (id_base lets a caller shard rows across devices/processes and keep global ids).
 * 1 <= d <= 4096.  The tensor kernels serve d <= 128; longer rows take the exact CUDA-core kernels, which stage fewer rows
 * per shared-memory tile as d grows (f32 d = 960: 24 rows instead of 128 -- correct, and slower per row). */
int gpann_points_upload(gpann_ctx* ctx, const float* host_rows, size_t n, size_t d, int storage, uint32_t id_base,
                        gpann_points** out);
/* Byte rows as stored in .u8bin / .i8bin files (src/points_io.cpp:63-111 without the float widening). */
int gpann_points_upload_bytes(gpann_ctx* ctx, const void* host_rows, size_t n, size_t d, int is_signed,
                              uint32_t id_base, gpann_points** out);
/* Rows already resident in device memory: storage is F32 (float rows), U8 or S8 (byte rows). */
int gpann_points_from_device(gpann_ctx* ctx, const void* dev_rows, size_t n, size_t d, int storage, uint32_t id_base,
                             void* stream, gpann_points** out);
void gpann_points_free(gpann_points* pts);
size_t gpann_points_n(const gpann_points* pts);
size_t gpann_points_d(const gpann_points* pts);
int gpann_points_storage(const gpann_points* pts);

/* ---- exact top-k: replaces ComputeGroundTruth (src/recall.h:25-39), ComputeDistanceToKthNeighbor
 *      (src/recall.h:10-23), ClosestLeaders (src/defs.cpp:87-96) and the scans of
 *      InvertedIndex::Query/QueryBucket (src/inverted_index.h:43-75) ------------------------------------- */
/* queries: nq rows of d floats on the host.  out_dist/out_id are [nq,k], ascending by (distance,id);
 * out_count[q] = min(k, n) (src/topn.h:21-29); unused slots hold +inf / GPANN_INVALID_ID.
 * out_dist or out_count may be NULL. */
int gpann_topk(gpann_ctx* ctx, const gpann_points* pts, const float* queries, size_t nq, int k, int metric,
               float* out_dist, uint32_t* out_id, uint32_t* out_count);
/* Only the kth distance (Top().first of the reference's heap): out_kth[nq]. */
int gpann_kth_distance(gpann_ctx* ctx, const gpann_points* pts, const float* queries, size_t nq, int k, int metric,
                       float* out_kth);
/* ComputeGroundTruth (src/recall.h:25-39) over a HOST-resident uint8/int8 point set that is STREAMED through the device:
 * chunks of chunk_rows rows (0 = 4 Mi rows) are copied H2D on a second stream while the previous chunk is laid out and
 * scanned by the tcgen05 kernel; the running top-k and per-query bounds carry from chunk to chunk.  Same results as
 * gpann_points_upload_bytes + gpann_topk, at max(PCIe time, scan time) instead of their sum, and with O(chunk) device
 * memory.  Pin host_rows (cudaHostAlloc / cudaHostRegister) for the copies to overlap.  d <= 128, k <= 448, L2, queries
 * integer-valued in the storage range (else GPANN_E_UNSUPPORTED: use gpann_points_upload* + gpann_topk). */
int gpann_topk_stream_bytes(gpann_ctx* ctx, const void* host_rows, size_t n, size_t d, int is_signed, uint32_t id_base,
                            const float* queries, size_t nq, int k, size_t chunk_rows, float* out_dist, uint32_t* out_id,
                            uint32_t* out_count);
/* The same for FLOAT host rows (PointSet::coordinates, src/defs.h:10-21): with storage AUTO, rows holding byte values
 * (the reference's reader widens .u8bin/.i8bin to float, src/points_io.cpp:63-111) are narrowed on the device chunk by
 * chunk and scanned by the exact integer kernel; other rows take the bf16 filter + exact fp32 re-rank against the raw
 * chunk.  out_kth (optional) receives the k-th distance per query (ComputeDistanceToKthNeighbor, src/recall.h:10-23). */
int gpann_topk_stream(gpann_ctx* ctx, const float* host_rows, size_t n, size_t d, int storage, uint32_t id_base, const float* queries,
                      size_t nq, int k, size_t chunk_rows, float* out_dist, uint32_t* out_id, uint32_t* out_count, float* out_kth);
/* Device-output form for the multi-GPU path: queries and the [nq,k] key block live on the device, work is enqueued on
 * `stream` (H2D copies on the ctx's copy stream) and the call returns without synchronising. */
int gpann_topk_stream_bytes_device(gpann_ctx* ctx, const void* host_rows, size_t n, size_t d, int is_signed, uint32_t id_base,
                                   const float* dev_queries, size_t nq, int k, size_t chunk_rows, uint64_t* dev_out_keys,
                                   void* stream);

/* Device-resident form: dev_queries = nq x d floats in device memory; dev_out_keys = [nq,k] uint64 keys
 * ascending (GPANN_EMPTY_KEY padded).  Asynchronous on `stream`.  This is the per-shard step of the
 * multi-GPU path: every rank runs it on its shard, all-gathers the keys and calls gpann_merge_keys_device. */
int gpann_topk_device(gpann_ctx* ctx, const gpann_points* pts, const float* dev_queries, size_t nq, int k, int metric,
                      uint64_t* dev_out_keys, void* stream);
/* The sharded step in one call (north_star: "each GPU computes a local top-k over its shard, and the final merge is an NCCL
 * all-gather of (distance, id) candidates over NVLink"): local scan of this rank's shard with the per-query bounds
 * all-reduced (min) across the ranks after the sample passes, all-gather of the [nq,k] key blocks, merge.  Every rank
 * receives the full result (any of dev_out_dist / dev_out_id / dev_out_keys, [nq,k]).  Asynchronous on `stream`; collective:
 * all ranks of the communicator call it with the same queries, k and metric.  Without a communicator = gpann_topk_device. */
int gpann_topk_sharded_device(gpann_ctx* ctx, const gpann_points* pts, const float* dev_queries, size_t nq, int k, int metric,
                              float* dev_out_dist, uint32_t* dev_out_id, uint64_t* dev_out_keys, void* stream);
/* Range-restricted scan (InvertedIndex buckets, src/inverted_index.h:43-75): only rows in
 * [row_begin,row_end) are scanned; ids are id_base + row. */
int gpann_topk_range(gpann_ctx* ctx, const gpann_points* pts, size_t row_begin, size_t row_end, const float* queries,
                     size_t nq, int k, int metric, float* out_dist, uint32_t* out_id, uint32_t* out_count);
/* ---- segmented exact scans: the batched forms of the callers either side of the big scans.  One launch
 *      serves all (query, segment) pairs; exact distances in the reference's arithmetic order (CUDA cores).  Single-device
 *      contexts only. -------------------------------------------------------------------------------------------------- */
/* InvertedIndex::Query / QueryBucket (src/inverted_index.h:43-75) for a BATCH of queries.  Bucket b owns the scan positions
 * [bucket_offsets[b], bucket_offsets[b+1]).  permutation == NULL: `pts` is the clustered copy (position = row, reported ids
 * are positions).  permutation != NULL: `pts` holds the ORIGINAL rows and position i is row permutation[i]; candidates are
 * still ranked by (distance, POSITION) -- the reference's tie-break, :49-51 -- and the reported ids are permutation[position]
 * (:56-59).  probes: [nq, nprobe] bucket ids per query (negative = unused; a bucket listed twice is probed once).
 * out_* as gpann_topk. */
int gpann_ivf_topk(gpann_ctx* ctx, const gpann_points* pts, const uint64_t* bucket_offsets, size_t num_buckets,
                   const uint32_t* permutation, const float* queries, size_t nq, const int32_t* probes, int nprobe, int k,
                   int metric, float* out_dist, uint32_t* out_id, uint32_t* out_count);
/* ApproximateKNNGraphBuilder::CrunchBucket (src/knn_graph.h:165-185) for ALL buckets at once: bucket b holds the points
 * bucket_members[bucket_offsets[b] .. bucket_offsets[b+1]); for every member position p the k nearest OTHER members of its
 * bucket, ranked by (distance, global id), ascending: out_dist / out_id [total,k], out_count[p] = min(k, bucket size - 1). */
int gpann_knn_buckets(gpann_ctx* ctx, const gpann_points* pts, const uint64_t* bucket_offsets, const uint32_t* bucket_members,
                      size_t num_buckets, int k, int metric, float* out_dist, uint32_t* out_id, uint32_t* out_count);
/* ClosestLeaders (src/defs.cpp:87-96) for every listed point at once (src/knn_graph.h:75-81; point_ids == NULL: all m =
 * n points): the k nearest rows of `leaders`, ids = leader index.  out_dist / out_id [m,k]. */
int gpann_closest_leaders_batch(gpann_ctx* ctx, const gpann_points* pts, const uint32_t* point_ids, size_t m,
                                const gpann_points* leaders, int k, int metric, float* out_dist, uint32_t* out_id,
                                uint32_t* out_count);
/* Per-group minimum distance: out[q*ngroups + g] = min over the points i with group[i] == g of distance(P_i, Q_q), FLT_MAX for
 * a group without points.  BruteForceRouting (analyze_approximation_losses.cpp:42-56) and the per-cluster ratings of the
 * overlapping partitioner (src/overlapping_partitioning.cpp:170-179).  group: host int32 [n] (negative = ignored). */
int gpann_group_min_dist(gpann_ctx* ctx, const gpann_points* pts, const int32_t* group, int ngroups, const float* queries, size_t nq,
                         int metric, float* out);
/* Merge `parts` key lists per query ([parts,nq,k], e.g. the all-gathered per-GPU candidates) into the k
 * smallest keys per query; writes distances/ids (either may be NULL) and/or merged keys. */
int gpann_merge_keys_device(gpann_ctx* ctx, const uint64_t* dev_keys, size_t parts, size_t nq, int k,
                            float* dev_out_dist, uint32_t* dev_out_id, uint64_t* dev_out_keys, void* stream);
/* Key <-> (distance,id) on the host, for callers that exchange keys themselves. */
uint64_t gpann_pack_key(float dist, uint32_t id);
void gpann_unpack_key(uint64_t key, float* dist, uint32_t* id);

/* ---- k-means: replaces NearestCenters / AggregateClusters* / KMeans / ObjectiveValue
 *      (src/kmeans.cpp:12-27, 116-204, 228-244, 246-249), L2 build and -DMIPS_DISTANCE build (metric = GPANN_METRIC_MIPS:
 *      rows scaled by 1/|x| before they are summed, NormalizeCentroidsIP; src/kmeans.cpp:61-101).  The tcgen05 assignment
 *      kernel serves L2; MIPS assignments run on the exact CUDA-core kernel. ------------------------------------------- */
/* argmin over centroids with strict '<' (lowest index wins ties); out_mindist may be NULL. */
int gpann_kmeans_assign(gpann_ctx* ctx, const gpann_points* pts, const float* centroids, size_t kc, int metric,
                        int32_t* out_assign, float* out_mindist);
/* One Lloyd round: assign, sum points per cluster, divide by the cluster size.  centroids_inout [kc,d] is
 * overwritten with the means (rows of empty clusters are all-zero, as in the reference before
 * RemoveEmptyClusters, src/kmeans.cpp:29-50, which stays in the host wrapper).  out_assign may be NULL. */
int gpann_kmeans_step(gpann_ctx* ctx, const gpann_points* pts, float* centroids_inout, size_t kc, int metric,
                      int32_t* out_assign, uint64_t* out_sizes);
/* Device form used by the multi-GPU path: writes this shard's per-cluster coordinate SUMS [kc,d] (float)
 * and counts [kc] (uint64) and the assignment [n] (int32, may be NULL); the caller all-reduces sums and
 * counts (NCCL) and divides. */
int gpann_kmeans_partial_device(gpann_ctx* ctx, const gpann_points* pts, const float* dev_centroids, size_t kc,
                                int metric, int32_t* dev_out_assign, float* dev_out_sums, uint64_t* dev_out_counts,
                                void* stream);
/* KMeans (src/kmeans.cpp:228-244) as ONE call: `rounds` Lloyd rounds (the reference's NUM_ROUNDS is 20) with the centroids
 * resident on the device -- no per-round host round trip.  A cluster that loses all members is removed as
 * RemoveEmptyClusters does (src/kmeans.cpp:29-50); the renumbering is applied once at the end: on return the first *out_kc
 * rows of centroids_inout are the surviving centroids, out_sizes[0..*out_kc) their sizes and out_assign [n] the LAST round's
 * assignment (against the pre-update centroids, renumbered) -- exactly what the reference returns.  out_assign / out_sizes
 * may be NULL.  On a group ctx (or a ctx with a communicator) sums and counts are all-reduced over NCCL every round. */
int gpann_kmeans_run(gpann_ctx* ctx, const gpann_points* pts, float* centroids_inout, size_t kc, int rounds, int metric,
                     int32_t* out_assign, uint64_t* out_sizes, size_t* out_kc);
/* Device form (asynchronous on `stream`, nothing renumbered): centroids [kc,d], alive [kc] (1 = live; the caller sets it)
 * and counts [kc] are caller-owned device buffers updated in place; dev_out_assign [n] may be NULL.  With a communicator
 * every rank passes its shard and all ranks end with identical centroids. */
int gpann_kmeans_run_device(gpann_ctx* ctx, const gpann_points* pts, float* dev_centroids_inout, size_t kc, int rounds, int metric,
                            int32_t* dev_out_assign, unsigned char* dev_alive_inout, uint64_t* dev_out_counts, void* stream);
/* BalancedKMeans (src/kmeans.cpp:251-524): KMeans, then rounds of 1000 mini-batches in which every point scores all clusters by
 * pos_distance + round_penalty * cluster_size under the BKM+ admissibility rule, moves, and the centroids follow; the penalty
 * rises until no cluster holds more than max_cluster_size points (at most 150 rounds, then the greedy re-migration of
 * :493-517).  The reference's moving phase is a racy parallel_for, so neither side is reproducible bit for bit: parity is
 * "balanced, and the objective within a few per cent".  Outputs as gpann_kmeans_run.  Single-device contexts. */
int gpann_balanced_kmeans(gpann_ctx* ctx, const gpann_points* pts, float* centroids_inout, size_t kc, size_t max_cluster_size,
                          int metric, int32_t* out_assign, uint64_t* out_sizes, size_t* out_kc);
/* sum_i pos_distance(P_i, C_assign[i]) accumulated in double (src/kmeans.cpp:246-249). */
int gpann_objective(gpann_ctx* ctx, const gpann_points* pts, const float* centroids, size_t kc, const int32_t* assign,
                    int metric, double* out);

#ifdef __cplusplus
}
#endif
#endif /* GPANN_H */
