// bkm.cu -- BalancedKMeans (reference src/kmeans.cpp:251-524) with every distance evaluation on the GPU.
//
// The reference: KMeans; then rounds of 1000 mini-batches over a random permutation of the points -- every point of a
// batch scores all clusters by  pos_distance + round_penalty * cluster_size  under the BKM+ admissibility rule
// (:403-447), moves (cluster sizes and coordinate sums follow with atomics, :324-345) and the centroids are recomputed after
// each batch (:348-378); the penalty rises from round to round (:466-474) until no cluster exceeds max_cluster_size.
// The moving phase is a racy parallel_for in the reference (live cluster sizes, CAS-float sums), so its result is not
// reproducible run to run; this implementation follows the same control flow with one deliberate simplification that makes
// it deterministic up to the order of the float atomics: a mini-batch reads the cluster sizes as they were when the
// batch STARTED (a valid serialisation of the reference's racy reads).  Parity is therefore statistical (tests: a balanced
// partition whose objective matches the reference's within a few per cent), as SURVEY.md 8c anticipates for KMeans.
//
// Kernels: bkm_move_kernel (one thread per point of the batch, centroids through shared memory, exact fp32 distances in the
// reference's order), bkm_update_kernel (update_centroids), bkm_min_kernel (min over penalties_needed), bkm_rebalance_kernel
// (the final greedy re-migration of :493-517, one CTA because every move changes the sizes the next one sees).
#include <cfloat>

#include "api_internal.h"
#include "distance.cuh"

namespace gpann {

struct BkmMoveParams {
    RowsView pts;
    const uint32_t* perm;        // the round's permutation of the point ids
    size_t begin, end;           // this mini-batch: perm[begin .. end)
    const float* centroids;      // [kc, d]
    int kc, metric;
    double round_penalty;
    int32_t* assign;             // [n] in/out
    unsigned long long* sizes;   // [kc] live cluster sizes (updated by the moves)
    const unsigned long long* sizes_at_start;  // [kc] snapshot the batch decides on
    float* sums;                 // [kc, d] cluster coordinate sums (MIPS: of x/|x|)
    double* norm_sums;           // MIPS: [kc] sums of |x|^2
    const float* row_norm;       // MIPS: [n] |x|
    double* penalties;           // [n] penalties_needed
    int cchunk, qpitch, spitch;
};

__device__ __forceinline__ float pos_of(float dist, int metric) { return metric == kMIPS ? __fadd_rn(dist, 1.0f) : dist; }  // pos_distance (src/dist.cpp:83-88)

// One thread per point of the mini-batch: the batch's rows (a random sample of the set) are gathered into padded shared
// memory, the centroids stream through shared memory in chunks.
template <int STORAGE>
__global__ void __launch_bounds__(128) bkm_move_kernel(const BkmMoveParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char* s_rows = smem_raw;                                                    // [128][spitch]
    float* s_c = reinterpret_cast<float*>(smem_raw + (size_t) 128 * p.spitch);          // [cchunk][qpitch]
    const int tid = threadIdx.x, d = p.pts.d;
    const int row_bytes = d * (STORAGE == kF32 ? 4 : 1);
    const size_t nb = p.end - p.begin;
    for (size_t base = (size_t) blockIdx.x * blockDim.x; base < nb; base += (size_t) gridDim.x * blockDim.x) {
        const size_t i = base + tid;
        const bool live = i < nb;
        const uint32_t pid = live ? p.perm[p.begin + i] : 0u;
        __syncthreads();
        if (live) {  // every thread copies its own row (rows of a random batch share nothing)
            const char* src = p.pts.rows + (size_t) pid * p.pts.pitch;
            for (int b = 0; b < row_bytes; ++b) s_rows[(size_t) tid * p.spitch + b] = (unsigned char) src[b];
        }
        const char* row = reinterpret_cast<const char*>(s_rows) + (size_t) tid * p.spitch;
        const int old = live ? p.assign[pid] : 0;
        float old_dist = 0.f;
        unsigned long long old_size = 0;
        if (live) {
            old_dist = pos_of(ref_distance_scalar<STORAGE>(row, p.centroids + (size_t) old * d, d, p.metric), p.metric);
            old_size = p.sizes_at_start[old];
        }
        int best = old;
        double best_score = DBL_MAX, min_pen = DBL_MAX;
        for (int c0 = 0; c0 < p.kc; c0 += p.cchunk) {
            const int nc = min(p.cchunk, p.kc - c0);
            __syncthreads();
            for (int t = tid; t < nc * d; t += blockDim.x) {
                const int c = t / d, j = t - c * d;
                s_c[c * p.qpitch + j] = p.centroids[(size_t) (c0 + c) * d + j];
            }
            __syncthreads();
            if (live) {
                for (int c = 0; c < nc; ++c) {
                    const int j = c0 + c;
                    const unsigned long long cs = p.sizes_at_start[j];
                    const float dist = pos_of(ref_distance<STORAGE>(row, s_c + c * p.qpitch, d, p.metric), p.metric);
                    const double score = (double) dist + p.round_penalty * (double) cs;
                    int denom = (int) old_size - (int) cs;  // src/kmeans.cpp:421-424
                    if (denom == 0) denom = 1;
                    const double pen = (double) ((dist - old_dist) / (float) denom);
                    if (old_size > cs) {
                        if (p.round_penalty < pen) {
                            if (pen < min_pen) min_pen = pen;
                        } else if (score < best_score) {
                            best = j;
                            best_score = score;
                        }
                    } else if (p.round_penalty < pen && score < best_score) {
                        best = j;
                        best_score = score;
                    }
                }
            }
        }
        if (live) {
            p.penalties[pid] = min_pen;
            if (best != old) {  // atomic_move (src/kmeans.cpp:324-345)
                p.assign[pid] = best;
                atomicAdd(p.sizes + old, (unsigned long long) -1ll);
                atomicAdd(p.sizes + best, 1ull);
                float mult = 1.0f;
                if (p.metric == kMIPS) {
                    const float nr = p.row_norm[pid];
                    atomicAdd(p.norm_sums + old, -(double) nr * (double) nr);
                    atomicAdd(p.norm_sums + best, (double) nr * (double) nr);
                    mult = 1.0f / nr;
                }
                for (int j = 0; j < d; ++j) {
                    const float x = load1<STORAGE>(row, j) * mult;
                    atomicAdd(p.sums + (size_t) best * d + j, x);
                    atomicAdd(p.sums + (size_t) old * d + j, -x);
                }
            }
        }
    }
}

// update_centroids (src/kmeans.cpp:348-378): C = sums / size (L2) or sums * sqrt((norm_sums/size) / |sums|^2) (MIPS); zero for
// an empty cluster.  Also refreshes the size snapshot the next mini-batch decides on.
__global__ void bkm_update_kernel(float* centroids, const float* sums, const unsigned long long* sizes, unsigned long long* sizes_snapshot,
                                  const double* norm_sums, int kc, int d, int metric) {
    __shared__ float s_red[128];
    const int c = blockIdx.x;
    const unsigned long long cnt = sizes[c];
    if (threadIdx.x == 0) sizes_snapshot[c] = cnt;
    float mult = 0.f;
    if (metric == kMIPS) {
        float acc = 0.f;
        for (int j = threadIdx.x; j < d; j += blockDim.x) {
            const float x = sums[(size_t) c * d + j];
            acc += x * x;
        }
        s_red[threadIdx.x] = acc;
        __syncthreads();
        for (int s = 64; s > 0; s >>= 1) {
            if ((int) threadIdx.x < s) s_red[threadIdx.x] += s_red[threadIdx.x + s];
            __syncthreads();
        }
        if (cnt != 0) mult = sqrtf(((float) norm_sums[c] / (float) cnt) / s_red[0]);
    }
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
        const float x = sums[(size_t) c * d + j];
        centroids[(size_t) c * d + j] = cnt == 0 ? 0.f : (metric == kMIPS ? x * mult : x / (float) cnt);
    }
}

__global__ void bkm_min_kernel(const double* x, size_t n, double* block_min) {
    __shared__ double s_red[256];
    double m = DBL_MAX;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) m = fmin(m, x[i]);
    s_red[threadIdx.x] = m;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int) threadIdx.x < s) s_red[threadIdx.x] = fmin(s_red[threadIdx.x], s_red[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) block_min[blockIdx.x] = s_red[0];
}
__global__ void bkm_min_final_kernel(const double* block_min, int nblocks, double* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double m = DBL_MAX;
        for (int i = 0; i < nblocks; ++i) m = fmin(m, block_min[i]);
        *out = m;
    }
}

__global__ void bkm_perm_keys_kernel(uint32_t* keys, uint32_t* vals, size_t n, uint32_t seed) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        uint32_t x = (uint32_t) i * 0x9E3779B1u + seed * 0x85EBCA77u;  // a 32-bit mixer: the sort of these keys is the shuffle
        x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
        keys[i] = x;
        vals[i] = (uint32_t) i;
    }
}

__global__ void bkm_widen_sizes_kernel(const uint64_t* a, unsigned long long* b, unsigned long long* c, double* norm_sums, const float* norm_sums_f, int kc) {
    for (int i = threadIdx.x; i < kc; i += blockDim.x) {
        b[i] = a[i];
        c[i] = a[i];
        if (norm_sums) norm_sums[i] = (double) norm_sums_f[i];
    }
}

// Final re-migration (src/kmeans.cpp:493-517), taken only when the rounds end with overloaded clusters: for every cluster c in
// order, while it is too large, its LAST member (largest id: ConvertPartitionToClusters fills in id order) moves to the
// nearest cluster that still has room (strict '<': lowest index on ties); centroids are not updated.  One CTA: every move
// changes the sizes the next one sees.  members: point ids sorted by (cluster, id); member_begin [kc+1].
template <int STORAGE>
__global__ void __launch_bounds__(256) bkm_rebalance_kernel(RowsView pts, const float* centroids, int kc, int metric, const uint32_t* members,
                                                            const uint64_t* member_begin, unsigned long long* sizes, unsigned long long max_size,
                                                            const unsigned char* alive, int32_t* assign) {
    __shared__ float s_best[256];
    __shared__ int s_idx[256];
    const int d = pts.d;
    for (int c = 0; c < kc; ++c) {
        uint64_t back = member_begin[c + 1];
        if (!alive[c]) continue;
        while (true) {
            __syncthreads();
            if (sizes[c] <= max_size) break;
            const uint32_t v = members[--back];
            const char* row = pts.rows + (size_t) v * pts.pitch;
            float best = FLT_MAX;
            int idx = 0x7fffffff;
            for (int j = threadIdx.x; j < kc; j += blockDim.x) {
                if (alive[j] && sizes[j] < max_size) {  // removed (empty) clusters do not exist any more (:477 ran RemoveEmptyClusters)
                    const float dist = ref_distance_scalar<STORAGE>(row, centroids + (size_t) j * d, d, metric);
                    if (dist < best) {
                        best = dist;
                        idx = j;
                    }
                }
            }
            s_best[threadIdx.x] = best;
            s_idx[threadIdx.x] = idx;
            __syncthreads();
            for (int s = 128; s > 0; s >>= 1) {
                if ((int) threadIdx.x < s) {
                    const float ob = s_best[threadIdx.x + s];
                    const int oi = s_idx[threadIdx.x + s];
                    if (ob < s_best[threadIdx.x] || (ob == s_best[threadIdx.x] && oi < s_idx[threadIdx.x])) {
                        s_best[threadIdx.x] = ob;
                        s_idx[threadIdx.x] = oi;
                    }
                }
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                const int target = s_idx[0];
                if (target != 0x7fffffff) {
                    assign[v] = target;
                    sizes[target] += 1;
                }
                sizes[c] -= 1;  // the member leaves c either way (no room anywhere: it keeps its old label, as nothing can take it)
            }
        }
    }
}

}  // namespace gpann

namespace gpimpl {
using namespace gpann;

namespace {
template <class F>
void by_storage(int storage, F f) {
    if (storage == kF32) f(std::integral_constant<int, kF32>());
    else if (storage == kU8) f(std::integral_constant<int, kU8>());
    else f(std::integral_constant<int, kS8>());
}
double penalty_function_iter(int round) { return round > 100 ? 1.01 : 1.5009 - 0.0009 * round; }  // src/kmeans.cpp:295-301 (BKM+)
}  // namespace

int sd_balanced_kmeans(gpann_ctx* ctx, const gpann_points* pts, float* centroids_inout, size_t kc, size_t max_cluster_size, int metric,
                       int32_t* out_assign, uint64_t* out_sizes, size_t* out_kc) {
    if (!ctx) return GPANN_E_ARG;
    if (!pts || pts->ctx != ctx || !pts->shards.empty()) return fail(ctx, GPANN_E_ARG, "balanced_kmeans: point set does not belong to this (single-device) context");
    if (!centroids_inout || !out_kc) return fail(ctx, GPANN_E_ARG, "balanced_kmeans: null argument");
    if (kc < 1) return fail(ctx, GPANN_E_ARG, "KMeans #centroids < 1");
    if (pts->n > 0xFFFFFFF0ull) return fail(ctx, GPANN_E_UNSUPPORTED, "balanced_kmeans: n must be below 2^32");
    const size_t n = pts->n, d = pts->d;
    const bool mips = metric == GPANN_METRIC_MIPS;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    OrderScope order(ctx, st);
    // ---- KMeans (:252) -> compacted centroids, renumbered assignment, sizes; all device-resident ----
    GP_TRY(kmeans_run_enqueue(ctx, pts, centroids_inout, kc, 20, metric));
    size_t kcl = 0;
    GP_TRY(kmeans_run_fetch(ctx, pts, nullptr, kc, nullptr, nullptr, &kcl));
    float* d_cent = (float*) ctx->km_run_cent.p;
    int32_t* d_assign = (int32_t*) ctx->assign.p;
    uint64_t* d_counts = (uint64_t*) ctx->km_run_counts.p;
    auto finish = [&](size_t kc_final) -> int {  // copy the device state out
        CU_TRY(ctx, cudaMemcpyAsync(centroids_inout, d_cent, kc_final * d * 4, cudaMemcpyDeviceToHost, st));
        if (out_sizes) CU_TRY(ctx, cudaMemcpyAsync(out_sizes, d_counts, kc_final * 8, cudaMemcpyDeviceToHost, st));
        if (out_assign && n) CU_TRY(ctx, cudaMemcpyAsync(out_assign, d_assign, n * 4, cudaMemcpyDeviceToHost, st));
        CU_TRY(ctx, cudaStreamSynchronize(st));
        *out_kc = kc_final;
        ctx->last_algo = "balanced_kmeans";
        return 0;
    };
    if (n == 0 || kcl == 0) return finish(kcl);
    // ---- cluster_coordinate_sums / cluster_sizes (:260-261: AggregateClustersParallel without normalising) ----
    ScopedDev sums, sizes, sizes_snap, norm_sums, norm_sums_f, penalties, best_assign, scal, block_min;
    const float* d_row_norm = nullptr;
    if (mips) {  // vector_sqrt_norms (:255-257)
        GP_TRY(reserve(ctx, ctx->km_run_norms, n * 4));
        launch_row_norms(rows_view(pts), (float*) ctx->km_run_norms.p, blocks_for(ctx, n), st);
        GP_TRY(check_launch(ctx, "row_norms_kernel"));
        d_row_norm = (const float*) ctx->km_run_norms.p;
    }
    CU_TRY(ctx, sums.alloc(kcl * d * 4));
    CU_TRY(ctx, sizes.alloc(kcl * 8));
    CU_TRY(ctx, sizes_snap.alloc(kcl * 8));
    CU_TRY(ctx, norm_sums.alloc(kcl * 8));
    CU_TRY(ctx, norm_sums_f.alloc(kcl * 4));
    CU_TRY(ctx, penalties.alloc(n * 8));
    CU_TRY(ctx, best_assign.alloc(n * 4));
    CU_TRY(ctx, scal.alloc(64));
    const int min_blocks = blocks_for(ctx, n, 256, 4);
    CU_TRY(ctx, block_min.alloc((size_t) min_blocks * 8));
    GP_TRY(cluster_sums_device(ctx, pts, d_assign, kcl, 0, sums.as<float>(), d_counts, st, d_row_norm, mips ? norm_sums_f.as<float>() : nullptr));
    bkm_widen_sizes_kernel<<<1, 256, 0, st>>>(d_counts, sizes.as<unsigned long long>(), sizes_snap.as<unsigned long long>(),
                                              mips ? norm_sums.as<double>() : nullptr, norm_sums_f.as<float>(), (int) kcl);
    GP_TRY(check_launch(ctx, "bkm_widen_sizes_kernel"));
    std::vector<uint64_t> h_sizes(kcl);
    auto fetch_sizes = [&]() -> int {
        CU_TRY(ctx, cudaMemcpyAsync(h_sizes.data(), sizes.p, kcl * 8, cudaMemcpyDeviceToHost, st));
        CU_TRY(ctx, cudaStreamSynchronize(st));
        return 0;
    };
    auto balanced = [&]() { return std::all_of(h_sizes.begin(), h_sizes.end(), [&](uint64_t s) { return s <= max_cluster_size; }); };
    GP_TRY(fetch_sizes());
    if (balanced()) return finish(kcl);  // :283-284

    // ---- the rounds (:380-475) ----
    GP_TRY(reserve(ctx, ctx->sort_a, n * 4));
    GP_TRY(reserve(ctx, ctx->sort_b, n * 4));
    GP_TRY(reserve(ctx, ctx->sort_c, n * 4));
    GP_TRY(reserve(ctx, ctx->sort_d, n * 4));
    size_t tmp_bytes = 0;
    sort_pairs_u32(nullptr, tmp_bytes, (uint32_t*) ctx->sort_a.p, (uint32_t*) ctx->sort_b.p, (uint32_t*) ctx->sort_c.p, (uint32_t*) ctx->sort_d.p, n, 32, st);
    GP_TRY(reserve(ctx, ctx->sort_tmp, tmp_bytes));
    CU_TRY(ctx, cudaMemcpyAsync(best_assign.p, d_assign, n * 4, cudaMemcpyDeviceToDevice, st));  // best_partition = closest_center (:313)
    BkmMoveParams mp;
    std::memset(&mp, 0, sizeof(mp));
    mp.pts = rows_view(pts);
    mp.perm = (const uint32_t*) ctx->sort_d.p;
    mp.centroids = d_cent;
    mp.kc = (int) kcl;
    mp.metric = metric;
    mp.assign = d_assign;
    mp.sizes = sizes.as<unsigned long long>();
    mp.sizes_at_start = sizes_snap.as<unsigned long long>();
    mp.sums = sums.as<float>();
    mp.norm_sums = mips ? norm_sums.as<double>() : nullptr;
    mp.row_norm = d_row_norm;
    mp.penalties = penalties.as<double>();
    mp.qpitch = ((int) d + 3) & ~3;
    mp.spitch = odd16_pitch((int) d * (pts->storage == GPANN_STORAGE_F32 ? 4 : 1));
    mp.cchunk = (int) std::max<size_t>(1, std::min<size_t>(kcl, (64 * 1024) / ((size_t) mp.qpitch * 4)));
    const size_t move_smem = (size_t) 128 * mp.spitch + (size_t) mp.cchunk * mp.qpitch * 4;
    if (move_smem > 200 * 1024) return fail(ctx, GPANN_E_UNSUPPORTED, "balanced_kmeans: d too large");
    constexpr int kMaxRounds = 150;
    constexpr size_t kSubRounds = 1000;
    const size_t chunk = (n + kSubRounds - 1) / kSubRounds;  // idiv_ceil(n, num_subrounds)
    by_storage(pts->storage, [&](auto S) {
        cudaFuncSetAttribute(bkm_move_kernel<decltype(S)::value>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    });
    int round = 0;
    double round_penalty = 0.0, best_objective = DBL_MAX;
    while (round++ <= kMaxRounds) {
        // perm = random_shuffle(iota(n), random(round)) (:400): a sort of hashed keys on the device
        const char* e_seed = std::getenv("GPANN_BKM_SEED");  // tests / experiments: a different shuffle sequence
        const uint32_t seed = (uint32_t) round + (e_seed ? (uint32_t) std::atoi(e_seed) * 1000u : 0u);
        bkm_perm_keys_kernel<<<blocks_for(ctx, n), 256, 0, st>>>((uint32_t*) ctx->sort_a.p, (uint32_t*) ctx->sort_c.p, n, seed);
        size_t tb = ctx->sort_tmp.cap;
        if (sort_pairs_u32(ctx->sort_tmp.p, tb, (uint32_t*) ctx->sort_a.p, (uint32_t*) ctx->sort_b.p, (uint32_t*) ctx->sort_c.p,
                           (uint32_t*) ctx->sort_d.p, n, 32, st) != cudaSuccess)
            return fail(ctx, GPANN_E_CUDA, "balanced_kmeans: radix sort failed");
        ctx->launches += 5;
        mp.round_penalty = round_penalty;
        for (size_t sub = 0; sub < kSubRounds; ++sub) {
            mp.begin = std::min(n, sub * chunk);
            mp.end = std::min(n, (sub + 1) * chunk);
            if (mp.begin >= mp.end) break;
            const unsigned grid = (unsigned) std::min<size_t>((mp.end - mp.begin + 127) / 128, (size_t) ctx->num_sms * 8);
            by_storage(pts->storage, [&](auto S) { bkm_move_kernel<decltype(S)::value><<<grid, 128, move_smem, st>>>(mp); });
            bkm_update_kernel<<<(unsigned) kcl, 128, 0, st>>>(d_cent, sums.as<float>(), sizes.as<unsigned long long>(), sizes_snap.as<unsigned long long>(),
                                                            mips ? norm_sums.as<double>() : nullptr, (int) kcl, (int) d, metric);
            ctx->launches += 2;
        }
        GP_TRY(check_launch(ctx, "bkm_move_kernel", 0));
        bkm_min_kernel<<<min_blocks, 256, 0, st>>>(penalties.as<double>(), n, block_min.as<double>());
        bkm_min_final_kernel<<<1, 32, 0, st>>>(block_min.as<double>(), min_blocks, scal.as<double>());
        double* d_obj = nullptr;
        GP_TRY(objective_device(ctx, pts, d_cent, kcl, d_assign, metric, &d_obj, st));
        double next_penalty = 0.0, objective = 0.0;
        CU_TRY(ctx, cudaMemcpyAsync(&next_penalty, scal.p, 8, cudaMemcpyDeviceToHost, st));
        CU_TRY(ctx, cudaMemcpyAsync(&objective, d_obj, 8, cudaMemcpyDeviceToHost, st));
        GP_TRY(fetch_sizes());
        static const bool verbose = std::getenv("GPANN_BKM_VERBOSE") != nullptr;
        if (verbose)
            std::fprintf(stderr, "[gpann bkm] round = %d penalty %g largest cluster %llu constraint %zu objective %g next penalty %g\n", round,
                         round_penalty, (unsigned long long) *std::max_element(h_sizes.begin(), h_sizes.end()), max_cluster_size, objective, next_penalty);
        if (balanced()) {
            if (objective < best_objective) {
                best_objective = objective;
                CU_TRY(ctx, cudaMemcpyAsync(best_assign.p, d_assign, n * 4, cudaMemcpyDeviceToDevice, st));
            } else {
                break;  // no improvement but balanced (:468-470)
            }
        } else {
            round_penalty = penalty_function_iter(round) * next_penalty;
        }
    }
    // ---- final aggregate with the best partition, normalised, empty clusters removed (:477) ----
    CU_TRY(ctx, cudaMemcpyAsync(d_assign, best_assign.p, n * 4, cudaMemcpyDeviceToDevice, st));
    GP_TRY(cluster_sums_device(ctx, pts, d_assign, kcl, 0, sums.as<float>(), d_counts, st, d_row_norm, mips ? norm_sums_f.as<float>() : nullptr));
    GP_TRY(reserve(ctx, ctx->km_run_alive, kcl));
    CU_TRY(ctx, cudaMemsetAsync(ctx->km_run_alive.p, 1, kcl, st));
    if (mips) launch_kmeans_update_ip(d_cent, sums.as<float>(), d_counts, norm_sums_f.as<float>(), (unsigned char*) ctx->km_run_alive.p, (int) kcl, (int) d, st);
    else launch_kmeans_update(d_cent, sums.as<float>(), d_counts, (unsigned char*) ctx->km_run_alive.p, (int) kcl, (int) d, st);
    GP_TRY(check_launch(ctx, "kmeans_update_kernel"));
    // the re-migration (:493-517) works on the clusters BEFORE renumbering would matter: sizes > max only in surviving clusters
    CU_TRY(ctx, cudaMemcpyAsync(h_sizes.data(), d_counts, kcl * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    if (!balanced()) {
        // members sorted by (cluster, id) and their offsets are what cluster_sums_device just left in the sort scratch
        bkm_widen_sizes_kernel<<<1, 256, 0, st>>>(d_counts, sizes.as<unsigned long long>(), sizes_snap.as<unsigned long long>(), nullptr, nullptr, (int) kcl);
        const uint64_t* member_begin = (const uint64_t*) ctx->misc.p;
        by_storage(pts->storage, [&](auto S) {
            bkm_rebalance_kernel<decltype(S)::value><<<1, 256, 0, st>>>(rows_view(pts), d_cent, (int) kcl, metric, (const uint32_t*) ctx->sort_d.p, member_begin,
                                                                        sizes.as<unsigned long long>(), (unsigned long long) max_cluster_size,
                                                                        (const unsigned char*) ctx->km_run_alive.p, d_assign);
        });
        GP_TRY(check_launch(ctx, "bkm_rebalance_kernel", 2));
        CU_TRY(ctx, cudaMemcpyAsync(d_counts, sizes.p, kcl * 8, cudaMemcpyDeviceToDevice, st));
    }
    int* d_kc_out = (int*) (d_counts + kcl);
    GP_TRY(reserve(ctx, ctx->km_run_counts, kc * 8 + 64));
    GP_TRY(kmeans_compact_device(ctx, d_cent, kcl, d, d_assign, n, (const unsigned char*) ctx->km_run_alive.p, d_counts, d_kc_out, st));
    int kc_final = 0;
    CU_TRY(ctx, cudaMemcpyAsync(&kc_final, d_kc_out, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    return finish((size_t) kc_final);
}

}  // namespace gpimpl
