// multi.cu -- the multi-GPU form of the hot path behind the C ABI (north_star: "each GPU computes a local
// top-k over its shard, and the final merge is an NCCL all-gather of (distance, id) candidates over NVLink.  For k-means,
// per-GPU centroid partial sums are combined with an NCCL all-reduce").
//
// Two ways to get a communicator, one data path:
//   * several processes, one GPU each (torchrun, bench.py): every process creates a plain ctx and joins with
//     gpann_comm_init_rank(ctx, nranks, rank, id) -- id from gpann_comm_unique_id on rank 0, passed around by the caller;
//   * one process, several GPUs (the C++ drop-in: ComputeGroundTruth / KMeans keep their single-call signatures,
//     src/recall.h:25-39, src/kmeans.cpp:228-244): gpann_init_multi(ctx, devices, ndev) builds a GROUP ctx that owns one
//     child ctx + worker thread per device and an ncclCommInitAll communicator.  Point sets uploaded through a group ctx
//     are sharded by contiguous row ranges; every host entry point runs its per-rank part on all workers at once.
// In both the per-rank work is the same code (gpann_api.cu: topk_sharded, kmeans_rounds_device) and the exchanges are:
// all-reduce(min) of the per-query bounds after the sample passes (so a shard starts its main pass with the bound of the
// whole set), all-gather of the [nq,k] key blocks, all-reduce(sum) of the [kc,d] sums + [kc] counts per Lloyd round.
//
// libnccl is not a link-time dependency: it is loaded on first use (the copy already mapped into the process -- torch's --
// or the system's libnccl.so.2), so single-GPU users need no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>

#include <condition_variable>
#include <mutex>
#include <thread>

#include "api_internal.h"

namespace gpimpl {

// ---- libnccl, loaded lazily ---------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
    void* lib = nullptr;
    std::string err;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok() const { return lib != nullptr && err.empty(); }
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = { std::getenv("GPANN_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) {
            api.err = std::string("libnccl.so.2 could not be loaded: ") + (dlerror() ? dlerror() : "not found");
            return;
        }
        auto sym = [&](const char* name) -> void* {
            void* p = dlsym(api.lib, name);
            if (!p && api.err.empty()) api.err = std::string("libnccl lacks ") + name;
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId)) sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank)) sym("ncclCommInitRank");
        api.CommInitAll = (decltype(api.CommInitAll)) sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy)) sym("ncclCommDestroy");
        api.AllReduce = (decltype(api.AllReduce)) sym("ncclAllReduce");
        api.AllGather = (decltype(api.AllGather)) sym("ncclAllGather");
        api.GroupStart = (decltype(api.GroupStart)) sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd)) sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString)) sym("ncclGetErrorString");
    });
    return api;
}

int nccl_fail(gpann_ctx* ctx, const char* what, ncclResult_t r) {
    return fail(ctx, GPANN_E_NCCL, std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error"));
}
#define NCCL_TRY(ctx, expr)                                   \
    do {                                                      \
        ncclResult_t r__ = (expr);                            \
        if (r__ != ncclSuccess) return nccl_fail(ctx, #expr, r__); \
    } while (0)
}  // namespace

// ---- peer-shared bounds ------------------------------------------------------------------------------------------------------
namespace {
size_t shared_bound_slots(const gpann_ctx* ctx) { return (size_t) ctx->num_sms * tc::kMaxR * tc::kTileM; }  // the largest query batch

int alloc_shared_bound(gpann_ctx* ctx) {
    if (ctx->shared_bound) return 0;
    const size_t slots = shared_bound_slots(ctx);
    CU_TRY(ctx, cudaMalloc(&ctx->shared_bound, slots * 4));
    CU_TRY(ctx, cudaMemset(ctx->shared_bound, 0x7f, slots * 4));
    ctx->shared_bound_slots = slots;
    return 0;
}

void peer_destroy(gpann_ctx* ctx) {
    if (ctx->peer_ipc)
        for (int i = 0; i < ctx->npeers; ++i) cudaIpcCloseMemHandle(ctx->peer_bound[i]);
    ctx->npeers = 0;
    ctx->peer_ipc = false;
    if (ctx->shared_bound) cudaFree(ctx->shared_bound);
    ctx->shared_bound = nullptr;
    ctx->shared_bound_slots = 0;
}

// Several processes: every rank exports its array as a CUDA IPC handle, the handles travel through the communicator
// itself (one all-gather), every rank maps the others'.  A rank that cannot map a peer (no P2P path) simply does not push
// to it: the shared bounds are an accelerator, never needed for the result.
int setup_peer_bounds_ipc(gpann_ctx* ctx) {
    static const bool off = std::getenv("GPANN_NO_PEER_BOUNDS") != nullptr;
    if (off || ctx->comm_size <= 1) return 0;
    GP_TRY(alloc_shared_bound(ctx));
    cudaIpcMemHandle_t mine;
    if (cudaIpcGetMemHandle(&mine, ctx->shared_bound) != cudaSuccess) {
        cudaGetLastError();
        std::memset(&mine, 0, sizeof(mine));
    }
    const int n = ctx->comm_size;
    ScopedDev d;
    CU_TRY(ctx, d.alloc((size_t) n * sizeof(mine)));
    CU_TRY(ctx, cudaMemcpy(d.as<char>() + (size_t) ctx->comm_rank * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice));
    NCCL_TRY(ctx, nccl().AllGather(d.as<char>() + (size_t) ctx->comm_rank * sizeof(mine), d.p, sizeof(mine), ncclChar, (ncclComm_t) ctx->comm,
                                   ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<cudaIpcMemHandle_t> all(n);
    CU_TRY(ctx, cudaMemcpy(all.data(), d.p, (size_t) n * sizeof(mine), cudaMemcpyDeviceToHost));
    const cudaIpcMemHandle_t zero = {};
    for (int r = 0; r < n && ctx->npeers < tc::kMaxPeers; ++r) {
        if (r == ctx->comm_rank || std::memcmp(&all[r], &zero, sizeof(zero)) == 0) continue;
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) ctx->peer_bound[ctx->npeers++] = (uint32_t*) p;
        else cudaGetLastError();
    }
    ctx->peer_ipc = true;
    return 0;
}
}  // namespace

void comm_destroy(gpann_ctx* ctx) {
    peer_destroy(ctx);
    if (ctx->comm && nccl().ok()) nccl().CommDestroy((ncclComm_t) ctx->comm);
    ctx->comm = nullptr;
    ctx->comm_size = 1;
    ctx->comm_rank = 0;
}

// Non-negative floats order like their bit patterns, so min over uint32 is min over the bounds (+inf = "none yet").
int comm_allreduce_min_u32(gpann_ctx* ctx, uint32_t* buf, size_t count, cudaStream_t st) {
    if (ctx->comm_size <= 1 || count == 0) return 0;
    NCCL_TRY(ctx, nccl().AllReduce(buf, buf, count, ncclUint32, ncclMin, (ncclComm_t) ctx->comm, st));
    ctx->launches++;
    return 0;
}
int comm_allgather_u64(gpann_ctx* ctx, const uint64_t* send, uint64_t* recv, size_t count, cudaStream_t st) {
    if (ctx->comm_size <= 1) {
        if (send != recv) CU_TRY(ctx, cudaMemcpyAsync(recv, send, count * 8, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    NCCL_TRY(ctx, nccl().AllGather(send, recv, count, ncclUint64, (ncclComm_t) ctx->comm, st));
    ctx->launches++;
    return 0;
}
int comm_allreduce_kmeans(gpann_ctx* ctx, float* sums, size_t nsums, uint64_t* counts, size_t ncounts, cudaStream_t st) {
    if (ctx->comm_size <= 1) return 0;
    NCCL_TRY(ctx, nccl().GroupStart());
    ncclResult_t r1 = nccl().AllReduce(sums, sums, nsums, ncclFloat32, ncclSum, (ncclComm_t) ctx->comm, st);
    ncclResult_t r2 = nccl().AllReduce(counts, counts, ncounts, ncclUint64, ncclSum, (ncclComm_t) ctx->comm, st);
    ncclResult_t r3 = nccl().GroupEnd();
    if (r1 != ncclSuccess) return nccl_fail(ctx, "ncclAllReduce(sums)", r1);
    if (r2 != ncclSuccess) return nccl_fail(ctx, "ncclAllReduce(counts)", r2);
    if (r3 != ncclSuccess) return nccl_fail(ctx, "ncclGroupEnd", r3);
    ctx->launches += 2;
    return 0;
}

int comm_unique_id(void* out, size_t bytes) {
    if (!out || bytes < sizeof(ncclUniqueId)) return GPANN_E_ARG;
    if (!nccl().ok()) {
        std::fprintf(stderr, "gpann_comm_unique_id: %s\n", nccl().err.c_str());
        return GPANN_E_NCCL;
    }
    ncclUniqueId id;
    if (nccl().GetUniqueId(&id) != ncclSuccess) return GPANN_E_NCCL;
    std::memset(out, 0, bytes);
    std::memcpy(out, &id, sizeof(id));
    return 0;
}

int comm_init_rank(gpann_ctx* ctx, int nranks, int rank, const void* id_bytes, size_t bytes) {
    if (!ctx) return GPANN_E_ARG;
    if (ctx->group || ctx->parent) return fail(ctx, GPANN_E_ARG, "comm_init_rank: ctx already belongs to a group");
    if (nranks < 1 || rank < 0 || rank >= nranks || !id_bytes || bytes < sizeof(ncclUniqueId)) return fail(ctx, GPANN_E_ARG, "comm_init_rank: bad argument");
    if (!nccl().ok()) return fail(ctx, GPANN_E_NCCL, nccl().err);
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    comm_destroy(ctx);
    if (nranks == 1) return 0;
    ncclUniqueId id;
    std::memcpy(&id, id_bytes, sizeof(id));
    ncclComm_t comm = nullptr;
    NCCL_TRY(ctx, nccl().CommInitRank(&comm, nranks, id, rank));
    ctx->comm = comm;
    ctx->comm_size = nranks;
    ctx->comm_rank = rank;
    return setup_peer_bounds_ipc(ctx);
}

// ---- group ctx: one worker thread per device ------------------------------------------------------------------------------
struct Worker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<int()> job;
    bool pending = false, stop = false, done = true;
    int rc = 0;
    void loop(int device) {
        cudaSetDevice(device);
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cv.wait(lk, [&] { return pending || stop; });
            if (stop) return;
            std::function<int()> j = std::move(job);
            pending = false;
            lk.unlock();
            const int r = j();
            lk.lock();
            rc = r;
            done = true;
            cv.notify_all();
        }
    }
};

struct Group {
    std::vector<gpann_ctx*> ranks;
    std::vector<Worker*> workers;
};

namespace {
// Run fn(rank, child ctx) on every worker at once; returns the first failure (its message copied to the root ctx).
int run_all(gpann_ctx* root, const std::function<int(int, gpann_ctx*)>& fn) {
    Group* g = root->group;
    const int n = (int) g->ranks.size();
    for (int i = 0; i < n; ++i) {
        Worker* w = g->workers[i];
        gpann_ctx* c = g->ranks[i];
        std::lock_guard<std::mutex> lk(w->m);
        w->job = [fn, i, c] { return fn(i, c); };
        w->pending = true;
        w->done = false;
        w->cv.notify_all();
    }
    int rc = 0;
    for (int i = 0; i < n; ++i) {
        Worker* w = g->workers[i];
        std::unique_lock<std::mutex> lk(w->m);
        w->cv.wait(lk, [&] { return w->done; });
        if (w->rc != 0 && rc == 0) {
            rc = w->rc;
            root->err = "device " + std::to_string(g->ranks[i]->device) + ": " + g->ranks[i]->err;
        }
    }
    return rc;
}

void shard_range(size_t n, int rank, int world, size_t* lo, size_t* hi) {  // contiguous ceil(n/G) rows per GPU
    const size_t per = (n + world - 1) / world;
    *lo = std::min(n, (size_t) rank * per);
    *hi = std::min(n, *lo + per);
}

int check_group_points(gpann_ctx* root, const gpann_points* pts, const char* what) {
    if (!pts || pts->ctx != root || pts->shards.size() != root->group->ranks.size())
        return fail(root, GPANN_E_ARG, std::string(what) + ": point set does not belong to this (group) context");
    return 0;
}
}  // namespace

int mg_init(gpann_ctx** out, const int* devices, int ndev) {
    if (!out) return GPANN_E_ARG;
    *out = nullptr;
    if (!devices || ndev < 1 || ndev > 64) return GPANN_E_ARG;
    for (int i = 0; i < ndev; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return GPANN_E_ARG;
    if (ndev > 1 && !nccl().ok()) {
        std::fprintf(stderr, "gpann_init_multi: %s\n", nccl().err.c_str());
        return GPANN_E_NCCL;
    }
    gpann_ctx* root = new (std::nothrow) gpann_ctx();
    if (!root) return GPANN_E_OOM;
    Group* g = new Group();
    root->group = g;
    root->device = devices[0];
    int rc = 0;
    for (int i = 0; i < ndev && rc == 0; ++i) {
        gpann_ctx* c = nullptr;
        rc = sd_init(&c, devices[i]);
        if (rc == 0) {
            c->parent = root;
            g->ranks.push_back(c);
        }
    }
    if (rc == 0 && ndev > 1) {
        std::vector<ncclComm_t> comms(ndev);
        if (nccl().CommInitAll(comms.data(), ndev, devices) != ncclSuccess) {
            std::fprintf(stderr, "gpann_init_multi: ncclCommInitAll failed\n");
            rc = GPANN_E_NCCL;
        } else {
            for (int i = 0; i < ndev; ++i) {
                g->ranks[i]->comm = comms[i];
                g->ranks[i]->comm_size = ndev;
                g->ranks[i]->comm_rank = i;
            }
        }
    }
    if (rc != 0) {
        for (gpann_ctx* c : g->ranks) sd_destroy(c);
        delete g;
        delete root;
        return rc;
    }
    root->num_sms = g->ranks[0]->num_sms;
    // peer-shared bounds inside one process: direct peer access between the devices of the group
    static const bool no_peers = std::getenv("GPANN_NO_PEER_BOUNDS") != nullptr;
    if (ndev > 1 && !no_peers) {
        for (int i = 0; i < ndev; ++i) {
            cudaSetDevice(devices[i]);
            if (alloc_shared_bound(g->ranks[i]) != 0) cudaGetLastError();
        }
        for (int i = 0; i < ndev; ++i) {
            gpann_ctx* c = g->ranks[i];
            cudaSetDevice(devices[i]);
            for (int j = 0; j < ndev && c->npeers < tc::kMaxPeers; ++j) {
                int can = 0;
                if (j == i || !g->ranks[j]->shared_bound || !c->shared_bound) continue;
                if (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) != cudaSuccess || !can) continue;
                const cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    cudaGetLastError();
                    continue;
                }
                cudaGetLastError();
                c->peer_bound[c->npeers++] = g->ranks[j]->shared_bound;
            }
        }
    }
    for (int i = 0; i < ndev; ++i) {
        Worker* w = new Worker();
        w->th = std::thread([w, d = devices[i]] { w->loop(d); });
        g->workers.push_back(w);
    }
    *out = root;
    return 0;
}

void mg_destroy(gpann_ctx* root) {
    Group* g = root->group;
    for (Worker* w : g->workers) {
        {
            std::lock_guard<std::mutex> lk(w->m);
            w->stop = true;
            w->cv.notify_all();
        }
        w->th.join();
        delete w;
    }
    for (gpann_ctx* c : g->ranks) sd_destroy(c);
    delete g;
    delete root;
}

int mg_num_devices(const gpann_ctx* root) { return (int) root->group->ranks.size(); }
gpann_ctx* mg_rank_ctx(const gpann_ctx* root, int i) { return root->group->ranks[(size_t) i]; }

// ---- sharded point sets ---------------------------------------------------------------------------------------------------
namespace {
int make_sharded(gpann_ctx* root, size_t n, size_t d, gpann_points** out,
                 const std::function<int(int, gpann_ctx*, size_t lo, size_t hi, gpann_points** shard)>& make) {
    const int world = (int) root->group->ranks.size();
    gpann_points* pts = new (std::nothrow) gpann_points();
    if (!pts) return fail(root, GPANN_E_OOM, "points_upload: host allocation failed");
    pts->ctx = root;
    pts->n = n;
    pts->d = d;
    pts->owns_rows = false;
    pts->shards.assign(world, nullptr);
    pts->shard_lo.assign(world + 1, n);
    for (int r = 0; r < world; ++r) {
        size_t lo, hi;
        shard_range(n, r, world, &lo, &hi);
        pts->shard_lo[r] = lo;
    }
    const int rc = run_all(root, [&](int r, gpann_ctx* c) { return make(r, c, pts->shard_lo[r], pts->shard_lo[r + 1], &pts->shards[r]); });
    if (rc != 0) {
        for (gpann_points* s : pts->shards) sd_points_free(s);
        delete pts;
        return rc;
    }
    // one storage for the handle: the shards' common storage, F32 if they differ (each shard scans exactly in its own)
    pts->storage = pts->shards[0]->storage;
    for (gpann_points* s : pts->shards)
        if (s->storage != pts->storage) pts->storage = GPANN_STORAGE_F32;
    *out = pts;
    return 0;
}
}  // namespace

int mg_points_upload(gpann_ctx* root, const float* host_rows, size_t n, size_t d, int storage, uint32_t id_base, gpann_points** out) {
    if (!out || (!host_rows && n) || d == 0 || d > 4096) return fail(root, GPANN_E_ARG, "points_upload: bad argument");
    if ((uint64_t) id_base + n > 0xFFFFFFFFull) return fail(root, GPANN_E_ARG, "points_upload: ids exceed 32 bits (recall.h:30)");
    *out = nullptr;
    return make_sharded(root, n, d, out, [&](int, gpann_ctx* c, size_t lo, size_t hi, gpann_points** shard) {
        return sd_points_upload(c, host_rows + lo * d, hi - lo, d, storage, id_base + (uint32_t) lo, shard);
    });
}

int mg_points_upload_bytes(gpann_ctx* root, const void* host_rows, size_t n, size_t d, int is_signed, uint32_t id_base, gpann_points** out) {
    if (!out || (!host_rows && n) || d == 0 || d > 4096) return fail(root, GPANN_E_ARG, "points_upload_bytes: bad argument");
    if ((uint64_t) id_base + n > 0xFFFFFFFFull) return fail(root, GPANN_E_ARG, "points_upload_bytes: ids exceed 32 bits");
    *out = nullptr;
    return make_sharded(root, n, d, out, [&](int, gpann_ctx* c, size_t lo, size_t hi, gpann_points** shard) {
        return sd_points_upload_bytes(c, (const char*) host_rows + lo * d, hi - lo, d, is_signed, id_base + (uint32_t) lo, shard);
    });
}

void mg_points_free(gpann_points* pts) {
    for (gpann_points* s : pts->shards) sd_points_free(s);
    delete pts;
}

// ---- top-k ------------------------------------------------------------------------------------------------------------------
namespace {
struct HostOut {
    float* dist;
    uint32_t* id;
    uint32_t* count;
    float* kth;
};

// Rank-local device output buffers (rank 0 copies them out; the other ranks run the same merge and drop the result).
int reserve_out(gpann_ctx* c, size_t nq, int k, size_t d) {
    GP_TRY(reserve(c, c->qdev, nq * d * 4));
    GP_TRY(reserve(c, c->dist_out, nq * k * 4));
    GP_TRY(reserve(c, c->id_out, nq * k * 4));
    GP_TRY(reserve(c, c->cnt_out, nq * 8));
    return 0;
}
TopkOut device_out(gpann_ctx* c, size_t nq, const HostOut& h, bool root_rank) {
    TopkOut o;
    if (root_rank) {
        o.dist = h.dist ? (float*) c->dist_out.p : nullptr;
        o.id = h.id ? (uint32_t*) c->id_out.p : nullptr;
        o.count = h.count ? (uint32_t*) c->cnt_out.p : nullptr;
        o.kth = h.kth ? (float*) c->cnt_out.p + nq : nullptr;
    }
    return o;
}
int fetch_out(gpann_ctx* c, size_t nq, int k, const HostOut& h, const TopkOut& o, bool root_rank, cudaStream_t st) {
    if (root_rank) {
        if (h.dist) CU_TRY(c, cudaMemcpyAsync(h.dist, o.dist, nq * k * 4, cudaMemcpyDeviceToHost, st));
        if (h.id) CU_TRY(c, cudaMemcpyAsync(h.id, o.id, nq * k * 4, cudaMemcpyDeviceToHost, st));
        if (h.count) CU_TRY(c, cudaMemcpyAsync(h.count, o.count, nq * 4, cudaMemcpyDeviceToHost, st));
        if (h.kth) CU_TRY(c, cudaMemcpyAsync(h.kth, o.kth, nq * 4, cudaMemcpyDeviceToHost, st));
    }
    CU_TRY(c, cudaStreamSynchronize(st));
    return 0;
}
}  // namespace

int mg_topk_host(gpann_ctx* root, const gpann_points* pts, size_t row_begin, size_t row_end, const float* queries, size_t nq, int k,
                 int metric, float* out_dist, uint32_t* out_id, uint32_t* out_count, float* out_kth) {
    GP_TRY(check_group_points(root, pts, "topk"));
    if (!queries) return fail(root, GPANN_E_ARG, "topk: null argument");
    if (nq == 0) return 0;
    if (k < 1 || k > 512) return fail(root, GPANN_E_ARG, "topk: k must be in [1,512]");
    if (row_begin > row_end || row_end > pts->n) return fail(root, GPANN_E_ARG, "topk: bad row range");
    const HostOut h = { out_dist, out_id, out_count, out_kth };
    const size_t d = pts->d;
    const int rc = run_all(root, [&](int r, gpann_ctx* c) {
        CU_TRY(c, cudaSetDevice(c->device));
        cudaStream_t st = c->stream;
        OrderScope order(c, st);
        GP_TRY(reserve_out(c, nq, k, d));
        CU_TRY(c, cudaMemcpyAsync(c->qdev.p, queries, nq * d * 4, cudaMemcpyHostToDevice, st));
        const TopkOut o = device_out(c, nq, h, r == 0);
        // this rank's part of the requested row range, in shard-local rows (the whole shard for a full scan)
        const size_t lo = pts->shard_lo[r], hi = pts->shard_lo[r + 1];
        const size_t b = std::min(std::max(row_begin, lo), hi) - lo, e = std::min(std::max(row_end, lo), hi) - lo;
        const bool full = row_begin == 0 && row_end == pts->n;
        c->gate_used = false;
        GP_TRY(topk_sharded(c, pts->shards[r], full ? 0 : b, full ? hi - lo : std::max(b, e), (const float*) c->qdev.p, nq, k, metric, o, st));
        return fetch_out(c, nq, k, h, o, r == 0, st);
    });
    root->last_algo = root->group->ranks[0]->last_algo;
    return rc;
}

int mg_topk_stream_host(gpann_ctx* root, const void* host_rows, int host_kind, int storage, size_t n, size_t d, uint32_t id_base,
                        const float* queries, size_t nq, int k, size_t chunk_rows, float* out_dist, uint32_t* out_id, uint32_t* out_count,
                        float* out_kth) {
    if ((!host_rows && n) || !queries) return fail(root, GPANN_E_ARG, "topk_stream: null argument");
    if (nq == 0) return 0;
    if (k < 1 || k > 448) return fail(root, GPANN_E_ARG, "topk_stream: k must be in [1,448]");
    if (d == 0 || d > (size_t) tc::kRowBytes) return fail(root, GPANN_E_UNSUPPORTED, "topk_stream: d must be in [1,128]");
    if ((uint64_t) id_base + n > 0xFFFFFFFFull) return fail(root, GPANN_E_ARG, "topk_stream: ids exceed 32 bits (recall.h:30)");
    const HostOut h = { out_dist, out_id, out_count, out_kth };
    const int world = (int) root->group->ranks.size();
    const size_t esz = host_kind == kHostF32 ? 4 : 1;
    const size_t part = nq * (size_t) k;
    // phase 1: every rank streams ITS rows of the host set through its device (no collective yet: a rank that fails here
    // -- memory, unsupported shape -- must not leave the others waiting in one)
    int rc = run_all(root, [&](int r, gpann_ctx* c) {
        CU_TRY(c, cudaSetDevice(c->device));
        cudaStream_t st = c->stream;
        OrderScope order(c, st);
        GP_TRY(reserve_out(c, nq, k, d));
        GP_TRY(reserve(c, c->gather_send, part * 8));
        GP_TRY(reserve(c, c->gather_recv, part * 8 * world));
        CU_TRY(c, cudaMemcpyAsync(c->qdev.p, queries, nq * d * 4, cudaMemcpyHostToDevice, st));
        size_t lo, hi;
        shard_range(n, r, world, &lo, &hi);
        TopkOut local;
        local.keys = (uint64_t*) c->gather_send.p;
        return topk_stream_host(c, (const char*) host_rows + lo * d * esz, host_kind, storage, hi - lo, d, id_base + (uint32_t) lo,
                                (const float*) c->qdev.p, nq, k, chunk_rows, local, st);
    });
    if (rc != 0) return rc;
    // phase 2: all-gather of the key blocks, merge, result to the host from rank 0
    rc = run_all(root, [&](int r, gpann_ctx* c) {
        cudaStream_t st = c->stream;
        OrderScope order(c, st);
        GP_TRY(comm_allgather_u64(c, (const uint64_t*) c->gather_send.p, (uint64_t*) c->gather_recv.p, part, st));
        const TopkOut o = device_out(c, nq, h, r == 0);
        GP_TRY(launch_merge(c, (const uint64_t*) c->gather_recv.p, world, nq, nq, k, o, 0, st, nullptr, 0));
        return fetch_out(c, nq, k, h, o, r == 0, st);
    });
    root->last_algo = root->group->ranks[0]->last_algo;
    return rc;
}

// ---- k-means ------------------------------------------------------------------------------------------------------------------
int mg_kmeans_assign(gpann_ctx* root, const gpann_points* pts, const float* centroids, size_t kc, int metric, int32_t* out_assign,
                     float* out_mindist) {
    GP_TRY(check_group_points(root, pts, "kmeans_assign"));
    if (!centroids || !out_assign) return fail(root, GPANN_E_ARG, "kmeans_assign: null argument");
    if (kc < 1) return fail(root, GPANN_E_ARG, "KMeans #centroids < 1");
    const int rc = run_all(root, [&](int r, gpann_ctx* c) {
        const size_t lo = pts->shard_lo[r];
        return sd_kmeans_assign(c, pts->shards[r], centroids, kc, metric, out_assign + lo, out_mindist ? out_mindist + lo : nullptr);
    });
    root->last_algo = root->group->ranks[0]->last_algo;
    return rc;
}

int mg_kmeans_run(gpann_ctx* root, const gpann_points* pts, float* centroids_inout, size_t kc, int rounds, int metric, int32_t* out_assign,
                  uint64_t* out_sizes, size_t* out_kc) {
    GP_TRY(check_group_points(root, pts, "kmeans_run"));
    const int rc = run_all(root, [&](int r, gpann_ctx* c) {
        OrderScope order(c, c->stream);
        GP_TRY(kmeans_run_enqueue(c, pts->shards[r], centroids_inout, kc, rounds, metric));
        // every rank holds the same centroids and sizes after the all-reduces: rank 0 reports them, each rank its assignments
        return kmeans_run_fetch(c, pts->shards[r], r == 0 ? centroids_inout : nullptr, kc, out_assign ? out_assign + pts->shard_lo[r] : nullptr,
                                r == 0 ? out_sizes : nullptr, r == 0 ? out_kc : nullptr);
    });
    root->last_algo = root->group->ranks[0]->last_algo;
    return rc;
}

// One Lloyd round through host buffers (gpann_kmeans_step): rows of empty clusters all-zero, nothing renumbered.
int mg_kmeans_step(gpann_ctx* root, const gpann_points* pts, float* centroids_inout, size_t kc, int metric, int32_t* out_assign,
                   uint64_t* out_sizes) {
    GP_TRY(check_group_points(root, pts, "kmeans_step"));
    if (!centroids_inout || !out_sizes) return fail(root, GPANN_E_ARG, "kmeans_step: null argument");
    if (kc < 1) return fail(root, GPANN_E_ARG, "KMeans #centroids < 1");
    if (metric != GPANN_METRIC_L2 && metric != GPANN_METRIC_MIPS) return fail(root, GPANN_E_ARG, "kmeans_step: bad metric");
    const size_t d = pts->d;
    std::vector<float> cent_in(centroids_inout, centroids_inout + kc * d);  // rank 0 overwrites the caller's buffer
    const int rc = run_all(root, [&](int r, gpann_ctx* c) {
        const gpann_points* sh = pts->shards[r];
        CU_TRY(c, cudaSetDevice(c->device));
        cudaStream_t st = c->stream;
        OrderScope order(c, st);
        GP_TRY(reserve(c, c->km_run_cent, kc * d * 4));
        GP_TRY(reserve(c, c->km_run_counts, kc * 8 + 64));
        GP_TRY(reserve(c, c->km_run_alive, kc));
        GP_TRY(reserve(c, c->assign, std::max<size_t>(sh->n, 1) * 4));
        CU_TRY(c, cudaMemcpyAsync(c->km_run_cent.p, cent_in.data(), kc * d * 4, cudaMemcpyHostToDevice, st));
        CU_TRY(c, cudaMemsetAsync(c->km_run_alive.p, 1, kc, st));
        GP_TRY(kmeans_rounds_device(c, sh, (float*) c->km_run_cent.p, kc, 1, metric, (int32_t*) c->assign.p, (unsigned char*) c->km_run_alive.p,
                                    (uint64_t*) c->km_run_counts.p, st));
        if (r == 0) {
            CU_TRY(c, cudaMemcpyAsync(centroids_inout, c->km_run_cent.p, kc * d * 4, cudaMemcpyDeviceToHost, st));
            CU_TRY(c, cudaMemcpyAsync(out_sizes, c->km_run_counts.p, kc * 8, cudaMemcpyDeviceToHost, st));
        }
        if (out_assign && sh->n) CU_TRY(c, cudaMemcpyAsync(out_assign + pts->shard_lo[r], c->assign.p, sh->n * 4, cudaMemcpyDeviceToHost, st));
        CU_TRY(c, cudaStreamSynchronize(st));
        return 0;
    });
    root->last_algo = root->group->ranks[0]->last_algo;
    return rc;
}

int mg_objective(gpann_ctx* root, const gpann_points* pts, const float* centroids, size_t kc, const int32_t* assign, int metric, double* out) {
    GP_TRY(check_group_points(root, pts, "objective"));
    if (!centroids || !out || (!assign && pts->n)) return fail(root, GPANN_E_ARG, "objective: null argument");
    std::vector<double> part(root->group->ranks.size(), 0.0);
    const int rc = run_all(root, [&](int r, gpann_ctx* c) {
        return sd_objective(c, pts->shards[r], centroids, kc, assign ? assign + pts->shard_lo[r] : nullptr, metric, &part[r]);
    });
    double total = 0.0;
    for (double p : part) total += p;  // rank order: reproducible
    *out = total;
    return rc;
}

int mg_for_each(gpann_ctx* root, const std::function<int(gpann_ctx*)>& fn) {
    int rc = 0;
    for (gpann_ctx* c : root->group->ranks) {
        const int r = fn(c);
        if (r != 0 && rc == 0) rc = r;
    }
    return rc;
}

}  // namespace gpimpl
