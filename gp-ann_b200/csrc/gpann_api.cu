// gpann_api.cu -- the C ABI of libgpann_cuda.so (declared in include/gpann.h): host-side orchestration over the
// sm_100a kernels of stream.cu / topk_tc.cu / kmeans.cu / prep.cu.  No CPU fallback lives here: every compute
// entry point launches CUDA kernels or fails.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "api_internal.h"

using namespace gpann;

namespace gpimpl {

// Row-major [rows, row_elems] operand matrix, fetched as boxes of box_rows x one 128-byte K block, SWIZZLE_128B.
int encode_rows_map(gpann_ctx* ctx, CUtensorMap* map, void* base, size_t rows, int box_rows, const char* what,
                    bool bf16 = false, int row_elems = tc::kRowBytes) {
    const int esz = bf16 ? 2 : 1;
    const cuuint64_t gdim[2] = { (cuuint64_t) row_elems, (cuuint64_t) rows };
    const cuuint64_t gstride[1] = { (cuuint64_t) row_elems * esz };
    const cuuint32_t box[2] = { (cuuint32_t) (tc::kRowBytes / esz), (cuuint32_t) box_rows };
    const cuuint32_t estr[2] = { 1, 1 };
    CUresult r = ctx->encode_tiled(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, gdim, gstride, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(ctx, GPANN_E_CUDA, std::string("cuTensorMapEncodeTiled(") + what + ") failed: " + std::to_string((int) r));
    return 0;
}

// ---- tensor layout for byte rows ---------------------------------------------------------------------------
int build_tensor_bytes(gpann_ctx* ctx, gpann_points* pts, cudaStream_t st) {
    const size_t n = pts->n;
    const int d = (int) pts->d;
    if (n == 0 || d > tc::kRowBytes || n > 0x7FFFFF00ull) return 0;
    if (!ctx->encode_tiled) return 0;
    const size_t n_pad = ((n + tc::kTileN - 1) / tc::kTileN) * tc::kTileN;
    ScopedDev norms, index, norms_sorted, perm, tmp;  // freed on every exit path
    size_t tmp_bytes = 0;
    if (norms.alloc(n * 4) != cudaSuccess || index.alloc(n * 4) != cudaSuccess || norms_sorted.alloc(n * 4) != cudaSuccess ||
        perm.alloc(n * 4) != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, GPANN_E_OOM, "build_tensor_bytes: out of device memory");
    }
    const int is_signed = pts->storage == GPANN_STORAGE_S8;
    launch_byte_norms((const unsigned char*) pts->rows, n, d, pts->pitch, is_signed, norms.as<uint32_t>(), index.as<uint32_t>(),
                      ctx->num_sms, st);
    GP_TRY(check_launch(ctx, "byte_norms_kernel"));
    sort_pairs_u32(nullptr, tmp_bytes, norms.as<uint32_t>(), norms_sorted.as<uint32_t>(), index.as<uint32_t>(), perm.as<uint32_t>(), n, 25, st);
    if (tmp.alloc(tmp_bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, GPANN_E_OOM, "build_tensor_bytes: out of device memory (sort)");
    }
    if (sort_pairs_u32(tmp.p, tmp_bytes, norms.as<uint32_t>(), norms_sorted.as<uint32_t>(), index.as<uint32_t>(), perm.as<uint32_t>(), n, 25,
                       st) != cudaSuccess)
        return fail(ctx, GPANN_E_CUDA, "build_tensor_bytes: radix sort failed");
    ctx->launches += 4;
    if (cudaMalloc(&pts->t_rows, n_pad * tc::kRowBytes) != cudaSuccess || cudaMalloc(&pts->t_norms, n_pad * 4) != cudaSuccess ||
        cudaMalloc(&pts->t_ids, n_pad * 4) != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, GPANN_E_OOM, "build_tensor_bytes: out of device memory (layout)");  // gpann_points_free releases what was allocated
    }
    launch_gather_sorted_rows((const unsigned char*) pts->rows, n, d, pts->pitch, perm.as<uint32_t>(), norms_sorted.as<uint32_t>(),
                              pts->id_base, n_pad, pts->t_rows, pts->t_norms, pts->t_ids, blocks_for(ctx, n_pad * 32, 256, 32), st);
    GP_TRY(check_launch(ctx, "gather_sorted_rows_kernel"));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    pts->n_pad = n_pad;
    GP_TRY(encode_rows_map(ctx, &pts->tmap, pts->t_rows, n_pad, tc::kTileN, "points"));
    pts->has_tensor = true;
    return 0;
}

// Lazily built on the first tensor-path query over a float point set (d <= 128): norm-sorted bf16 rows.  Running out of
// memory here is not an error of the query: the layout is simply not built and the exact CUDA-core scan serves the set.
int build_tensor_float(gpann_ctx* ctx, gpann_points* pts, cudaStream_t st) {
    pts->tensor_tried = true;
    const size_t n = pts->n;
    const int d = (int) pts->d;
    if (n == 0 || d > 64 * tc::kMaxKBlocks || n > 0x7FFFFF00ull || !ctx->encode_tiled) return 0;
    const int dp = ((d + 63) / 64) * 64;
    const size_t n_pad = ((n + tc::kTileN - 1) / tc::kTileN) * tc::kTileN;
    ScopedDev norms, index, norms_sorted, perm, tmp;
    size_t tmp_bytes = 0;
    auto give_up = [&]() {  // not enough memory: drop what was allocated, keep the point set usable
        cudaGetLastError();
        if (pts->t_rows) cudaFree(pts->t_rows);
        if (pts->t_norms) cudaFree(pts->t_norms);
        if (pts->t_ids) cudaFree(pts->t_ids);
        pts->t_rows = nullptr, pts->t_norms = nullptr, pts->t_ids = nullptr;
        return 0;
    };
    if (norms.alloc(n * 4) != cudaSuccess || index.alloc(n * 4) != cudaSuccess || norms_sorted.alloc(n * 4) != cudaSuccess ||
        perm.alloc(n * 4) != cudaSuccess)
        return give_up();
    launch_float_norms((const float*) pts->rows, n, d, norms.as<uint32_t>(), index.as<uint32_t>(), blocks_for(ctx, n), st);
    GP_TRY(check_launch(ctx, "float_norms_kernel"));
    sort_pairs_u32(nullptr, tmp_bytes, norms.as<uint32_t>(), norms_sorted.as<uint32_t>(), index.as<uint32_t>(), perm.as<uint32_t>(), n, 32, st);
    if (tmp.alloc(tmp_bytes) != cudaSuccess) return give_up();
    if (sort_pairs_u32(tmp.p, tmp_bytes, norms.as<uint32_t>(), norms_sorted.as<uint32_t>(), index.as<uint32_t>(), perm.as<uint32_t>(), n, 32,
                       st) != cudaSuccess)
        return fail(ctx, GPANN_E_CUDA, "build_tensor_float: radix sort failed");
    ctx->launches += 5;
    if (cudaMalloc(&pts->t_rows, n_pad * (size_t) dp * 2) != cudaSuccess || cudaMalloc(&pts->t_norms, n_pad * 4) != cudaSuccess ||
        cudaMalloc(&pts->t_ids, n_pad * 4) != cudaSuccess)
        return give_up();
    launch_gather_sorted_rows_bf16((const float*) pts->rows, n, d, dp, perm.as<uint32_t>(), norms_sorted.as<uint32_t>(), pts->id_base,
                                   n_pad, (unsigned short*) pts->t_rows, (float*) pts->t_norms, pts->t_ids,
                                   blocks_for(ctx, n_pad * 32, 256, 32), st);
    GP_TRY(check_launch(ctx, "gather_sorted_rows_bf16_kernel"));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    pts->n_pad = n_pad;
    pts->dp = dp;
    GP_TRY(encode_rows_map(ctx, &pts->tmap, pts->t_rows, n_pad, tc::kTileN, "float points", true, dp));
    pts->has_tensor = true;
    return 0;
}

// ---- top-k dispatch ----------------------------------------------------------------------------------------

// gate/gate_want: optional device flag -- the merge runs only when (*gate != 0) == gate_want (asynchronous fall-backs).
int launch_merge(gpann_ctx* ctx, const uint64_t* keys, size_t parts, size_t nq_stride, size_t nq, int k, const TopkOut& o,
                 size_t out_q_offset, cudaStream_t st, const int* gate, int gate_want) {
    MergeParams mp;
    mp.keys = keys;
    mp.parts = parts;
    mp.nq = nq;
    mp.nq_stride = nq_stride;
    mp.k = k;
    mp.cap = std::max(64, pow2_ceil(2 * k));
    mp.out_dist = o.dist ? o.dist + out_q_offset * k : nullptr;
    mp.out_id = o.id ? o.id + out_q_offset * k : nullptr;
    mp.out_keys = o.keys ? o.keys + out_q_offset * k : nullptr;
    mp.out_count = o.count ? o.count + out_q_offset : nullptr;
    mp.out_kth = o.kth ? o.kth + out_q_offset : nullptr;
    mp.kth_full_only = o.kth_full_only;
    mp.gate = gate;
    mp.gate_want = gate_want;
    const size_t smem = (size_t) 4 * mp.cap * sizeof(uint64_t);
    launch_merge_keys(mp, (unsigned) ((nq + 3) / 4), smem, st);
    return check_launch(ctx, "merge_keys_kernel");
}

// The CUDA-core scan of rows [row_begin,row_end) for nq queries.  Queries are served in batches sized so that the per-CTA
// partial lists stay inside a fixed scratch budget (the reference's ComputeGroundTruth has no limit on nq or k).
// Rows per shared-memory tile of the one-thread-per-row kernels: kStreamTile (128), fewer when a row is so long that 128 of
// them would not leave room for the queries / centroids next to them (f32 d > ~190: GIST-960 stages 24 rows, d = 4096 five).
static int stream_tile_rows(int spitch) {
    return (int) std::max<size_t>(1, std::min<size_t>(kStreamTile, (size_t) (96 * 1024) / (size_t) spitch));
}

int topk_stream(gpann_ctx* ctx, const gpann_points* pts, size_t row_begin, size_t row_end, const float* dq, size_t nq, int k,
                int metric, const TopkOut& out, cudaStream_t st, const int* gate, int gate_want) {
    StreamParams p;
    std::memset(&p, 0, sizeof(p));
    p.rows = (const char*) pts->rows;
    p.pitch = pts->pitch;
    p.storage = pts->storage;
    p.d = (int) pts->d;
    p.row_begin = row_begin;
    p.row_end = row_end;
    p.ids = nullptr;
    p.id_base = pts->id_base;
    p.id_lo = 0;
    p.id_hi = 0xFFFFFFFFu;
    p.k = k;
    p.metric = metric;
    p.gate = gate;
    p.gate_want = gate_want;
    const int elem = pts->storage == GPANN_STORAGE_F32 ? 4 : 1;
    const int row_bytes = p.d * elem;
    p.spitch = odd16_pitch(row_bytes);
    p.qpitch = (p.d + 3) & ~3;
    p.cap = pow2_ceil(k + kStreamTile);
    p.vec_ok = (((uintptr_t) pts->rows + row_begin * pts->pitch) % 16 == 0) && (pts->pitch % 16 == 0) && (row_bytes % 16 == 0);
    // two row tiles (the next one in flight behind the scan of this one) when the rows allow 16-byte copies, two tiles leave
    // room for two CTAs per SM, and the scan is bound by HBM: few queries.  Measured (u8 100M x 128 / f32 10M x 96, one or two
    // buffers): nq = 1: 4.40 -> 3.09 ms / 1.07 -> 0.93 ms; nq = 4: 7.26 -> 5.48 / 1.69 -> 2.19; nq = 16: 19.0 -> 22.9 / 4.38 -> 5.19
    // (more queries make the tile compute-bound and the second buffer only costs occupancy).
    static const bool one_buf = std::getenv("GPANN_STREAM_NBUF1") != nullptr;
    const size_t nq_hbm_bound = pts->storage == GPANN_STORAGE_F32 ? 1 : 4;
    p.tile_rows = stream_tile_rows(p.spitch);
    p.nbuf = (p.vec_ok && !one_buf && nq <= nq_hbm_bound && (size_t) p.tile_rows * p.spitch * 2 <= 104 * 1024) ? 2 : 1;
    if (const char* e_nb = std::getenv("GPANN_STREAM_NBUF")) {  // experiment knob: ring depth 1..4 (needs 16-byte rows)
        const int nb = std::atoi(e_nb);
        if (p.vec_ok && nb >= 1 && nb <= 4 && (size_t) p.tile_rows * p.spitch * nb <= 160 * 1024) p.nbuf = nb;
    }
    const size_t tile_bytes = (size_t) p.nbuf * p.tile_rows * p.spitch;
    // per query: float row, candidate buffer, threshold + count; byte storage adds the query packed as bytes (qpitch bytes,
    // a multiple of 16) + its integer norm + the "integer-valued" flag of the dp4a path
    // (threshold 8 + count 4 + integer norm 4 + integer flag 4 bytes per query, whatever the storage: the kernel initialises all)
    const size_t per_q = (size_t) p.qpitch * 4 + (size_t) p.cap * 8 + 24 + (elem == 1 ? (size_t) ((p.qpitch + 15) & ~15) + 16 : 0);
    const size_t budget = 160 * 1024;
    if (tile_bytes + per_q > 200 * 1024)
        return fail(ctx, GPANN_E_UNSUPPORTED, "topk: k too large for the streaming kernel's shared-memory candidate buffers");
    const int qb = (int) std::min<size_t>(std::min<size_t>(nq, 16),
                                          std::max<size_t>(1, (budget > tile_bytes ? budget - tile_bytes : 0) / per_q));
    p.qb = qb;
    const size_t smem = tile_bytes + (size_t) qb * per_q + 16;  // + alignment slack of the packed byte queries
    const size_t nrows = row_end - row_begin;
    const size_t ntiles = (nrows + p.tile_rows - 1) / p.tile_rows;
    const int occ = (int) std::max<size_t>(1, std::min<size_t>(8, (220 * 1024) / (smem + 1024)));
    unsigned gx = (unsigned) std::max<size_t>(1, std::min<size_t>(ntiles, (size_t) ctx->num_sms * occ));
    // a gated launch is normally a no-op (every CTA reads the flag and returns): keep its grid, and its scratch, small
    if (gate) gx = std::min<unsigned>(gx, (unsigned) ctx->num_sms);
    // queries per launch: gx * nq_l * k keys of partial lists within ~256 MB (at least one shared-memory batch)
    const size_t scratch_budget = (size_t) 256 << 20;
    const size_t nq_launch = std::max<size_t>(qb, std::min<size_t>(nq, scratch_budget / ((size_t) gx * k * sizeof(uint64_t)) / qb * qb));
    GP_TRY(reserve(ctx, ctx->partial, (size_t) gx * nq_launch * k * sizeof(uint64_t)));
    p.partial = (uint64_t*) ctx->partial.p;
    for (size_t q0 = 0; q0 < nq; q0 += nq_launch) {
        const size_t nq_l = std::min(nq_launch, nq - q0);
        p.queries = dq + q0 * pts->d;
        p.nq = (int) nq_l;
        const unsigned gy = (unsigned) std::min<size_t>((nq_l + qb - 1) / qb, gate ? 4 : 32768);
        const bool time_it = ctx->timing && q0 == 0 && !gate;  // a gated fall-back must not overwrite the tensor kernel's bracket
        if (time_it) cudaEventRecord(ctx->ev0, st);
        launch_topk_stream(p, dim3(gx, gy), smem, st);
        if (time_it) {
            cudaEventRecord(ctx->ev1, st);
            ctx->timed = true;
        }
        GP_TRY(check_launch(ctx, "topk_stream_kernel"));
        GP_TRY(launch_merge(ctx, p.partial, gx, nq_l, nq_l, k, out, q0, st, gate, gate_want));
    }
    if (!gate) ctx->last_algo = "stream";
    return 0;
}

// ---- the tcgen05 path: byte storage (exact int32 distances) or float storage (bf16 filter + exact fp32 re-rank) ----
// A scan is split into query preparation (once per query batch) and the scan of one norm-sorted layout, so the
// streaming entry points can reuse the prepared queries for every chunk of a host-resident point set.
struct TcLayout {                 // one norm-sorted tensor layout: a whole point set or one streamed chunk
    const CUtensorMap* tmap;
    const void* norms;
    const uint32_t* ids;
    size_t n_pad;
    const float* rows_f32;        // float path: original rows for the exact re-rank (ids - id_base index them)
    uint32_t id_base;
};

struct TcPlan {
    bool is_float = false;
    int is_signed = 0, d = 0, dp = 0, kblocks = 1, max_r = tc::kMaxR, stages = 3, qcap = 1024;
    int R = 1, G = 1, k = 0, cap = 128, lists_smem = 0;
    size_t nqb = 0, nq_pad = 0;
    const int* gate = nullptr;  // device flag the batch's kernels are gated on (byte storage: queries integer-valued)
    uint32_t* bound = nullptr;  // [nq_pad] per-query bounds of this batch: ctx->bound, or the peer-shared array of a sharded call
    bool peers = false;         // the kernels also push tightened bounds to the other ranks' arrays
    int mips = 0;               // metric of the scan
    CUtensorMap tmap_q;
};

size_t tc_batch_max(const gpann_ctx* ctx, int kblocks) { return (size_t) ctx->num_sms * (kblocks == 1 ? tc::kMaxR : 2) * tc::kTileM; }

constexpr int kFlagQueriesOk = 4;  // ctx->flags[4]: non-zero = the prepared queries are integer-valued in the storage range

// Prepare one batch of nqb <= tc_batch_max queries (asynchronous).  For byte storage the kernel clears
// ctx->flags[kFlagQueriesOk] when a query coordinate is not an integer in the storage's range; every tensor-path kernel of
// the batch is gated on that flag (pl.gate) and the caller queues the exact float scan behind it, gated the other way.
void tc_plan_shape(TcPlan& pl, int storage, int d, int dp, size_t nqb, int k) {
    pl.is_float = storage == GPANN_STORAGE_F32;
    pl.is_signed = storage == GPANN_STORAGE_S8;
    pl.d = d;
    pl.dp = dp;
    pl.kblocks = pl.is_float ? dp / 64 : 1;
    // shared-memory plan: bytes / one K block: 4 Q tiles, 3 stages, 1024-deep queues; two K blocks: 2, 2, 512
    pl.max_r = pl.kblocks == 1 ? tc::kMaxR : 2;
    pl.stages = pl.kblocks == 1 ? 3 : 2;
    pl.qcap = pl.kblocks == 1 ? 1024 : 512;
    // small k, one K block: the per-query lists fit in shared memory next to a 2-deep ring (a 3rd stage measured no
    // faster) and 512-record queues; a candidate record then costs the selection warp ~0.3k clk instead of ~2.8k
    static const char* e_sl = std::getenv("GPANN_TC_SMEMLISTS");
    pl.lists_smem = pl.kblocks == 1 && k <= tc::kSmemListKeys && !(e_sl && std::atoi(e_sl) == 0);
    if (pl.lists_smem) {
        pl.stages = 2;
        pl.qcap = 512;
    }
    // experiment knobs (probes; the defaults above are the tuned plan)
    static const char* e_st = std::getenv("GPANN_TC_STAGES");
    static const char* e_qc = std::getenv("GPANN_TC_QCAP");
    static const char* e_r = std::getenv("GPANN_TC_R");
    if (e_st) pl.stages = std::max(2, std::min(tc::kMaxStages, std::atoi(e_st)));
    if (e_qc) pl.qcap = std::max(64, pow2_ceil(std::atoi(e_qc)));
    if (e_r) pl.max_r = std::max(1, std::min(pl.max_r, std::atoi(e_r)));
    pl.nqb = nqb;
    pl.k = k;
    pl.R = (int) std::min<size_t>(pl.max_r, (nqb + tc::kTileM - 1) / tc::kTileM);  // Q tiles resident per CTA
    pl.G = (int) ((nqb + (size_t) pl.R * tc::kTileM - 1) / ((size_t) pl.R * tc::kTileM));
    pl.nq_pad = (size_t) pl.G * pl.R * tc::kTileM;
    pl.cap = std::max(128, pow2_ceil(k + 64));
    pl.gate = nullptr;
}

int tc_prepare_queries(gpann_ctx* ctx, TcPlan& pl, int storage, int d, int dp, const float* dq, size_t nqb, int k, cudaStream_t st,
                       int metric = GPANN_METRIC_L2) {
    tc_plan_shape(pl, storage, d, dp, nqb, k);
    pl.mips = metric == GPANN_METRIC_MIPS;
    const size_t qrow_bytes = pl.is_float ? (size_t) dp * 2 : tc::kRowBytes;
    GP_TRY(reserve(ctx, ctx->qrows, pl.nq_pad * qrow_bytes));
    GP_TRY(reserve(ctx, ctx->qnorm, pl.nq_pad * 4));
    GP_TRY(reserve(ctx, ctx->bound, pl.nq_pad * 4));
    GP_TRY(reserve(ctx, ctx->flags, 64));
    pl.peers = ctx->peer_share && ctx->shared_bound && pl.nq_pad <= ctx->shared_bound_slots;
    pl.bound = pl.peers ? ctx->shared_bound : (uint32_t*) ctx->bound.p;
    if (pl.is_float) {
        GP_TRY(reserve(ctx, ctx->qf32, pl.nq_pad * d * sizeof(float)));
        launch_prep_queries_bf16(dq, (int) nqb, d, dp, (int) pl.nq_pad, (unsigned short*) ctx->qrows.p, (float*) ctx->qnorm.p,
                                 (float*) ctx->qf32.p, (ctx->tc_warm && ctx->tc_warm_valid) ? nullptr : pl.bound, st);
        GP_TRY(check_launch(ctx, "prep_queries_bf16_kernel"));
    } else {
        int* okflag = (int*) ctx->flags.p + kFlagQueriesOk;
        CU_TRY(ctx, cudaMemsetAsync(okflag, 1, sizeof(int), st));  // 0x01010101: "ok" until a query refutes it
        launch_prep_queries_bytes(dq, (int) nqb, d, (int) pl.nq_pad, pl.is_signed, (unsigned char*) ctx->qrows.p,
                                  (int32_t*) ctx->qnorm.p, okflag,
                                  (ctx->tc_warm && ctx->tc_warm_valid) ? nullptr : pl.bound, st);
        GP_TRY(check_launch(ctx, "prep_queries_bytes_kernel"));
        pl.gate = okflag;
    }
    return encode_rows_map(ctx, &pl.tmap_q, ctx->qrows.p, pl.nq_pad, tc::kTileM, "queries", pl.is_float,
                           pl.is_float ? dp : tc::kRowBytes);
}

// Launch geometry + kernel arguments shared by the sample passes and the main pass over one layout.
struct TcLaunch {
    tc::Args a;
    int S = 1, S_max = 1;
    unsigned grid = 1;
    size_t smem = 0, part_keys = 0;
    uint32_t tiles_total = 0;
    uint64_t* carry_keys = nullptr;  // ctx->partial part 0
    size_t flog_tail_bytes = 0, flog_bytes = 0;  // find log (tc::Args::flog): cleared before every launch
    size_t qstat_bytes = 0;                      // order-statistic bounds (tc::Args::qstat): 0xFF before every launch
};

// The find log starts every launch empty: tails 0, entries 0xFF (= not written).
int tc_flog_reset(gpann_ctx* ctx, const TcLaunch& L, cudaStream_t st) {
    if (L.a.qstat) CU_TRY(ctx, cudaMemsetAsync(L.a.qstat, 0xFF, L.qstat_bytes, st));
    if (!L.a.flog) return 0;
    CU_TRY(ctx, cudaMemsetAsync(L.a.flog_tail, 0, L.flog_tail_bytes, st));
    CU_TRY(ctx, cudaMemsetAsync(L.a.flog, 0xFF, L.flog_bytes, st));
    return 0;
}

int tc_setup(gpann_ctx* ctx, const TcPlan& pl, const TcLayout& lay, TcLaunch& L) {
    const int k = pl.k;
    L.tiles_total = (uint32_t) (lay.n_pad / tc::kTileN);
    L.S = (int) std::max<uint32_t>(1, std::min<uint32_t>((uint32_t) (ctx->num_sms / pl.G), L.tiles_total));
    L.S_max = std::max(1, ctx->num_sms / pl.G);
    L.part_keys = pl.nq_pad * (size_t) k;
    if (!pl.lists_smem) GP_TRY(reserve(ctx, ctx->cand, (size_t) pl.G * L.S * pl.R * tc::kTileM * pl.cap * sizeof(uint64_t)));
    GP_TRY(reserve(ctx, ctx->partial, (size_t) (L.S_max + 1) * L.part_keys * sizeof(uint64_t)));
    L.carry_keys = (uint64_t*) ctx->partial.p;  // part 0: the running result; the splits write parts 1..S
    tc::Args& a = L.a;
    std::memset(&a, 0, sizeof(a));
    a.pnorm = lay.norms;
    a.pid = lay.ids;
    a.qnorm = ctx->qnorm.p;
    a.qrows = (const unsigned char*) ctx->qrows.p;
    a.nq = (int) pl.nqb;
    a.R = pl.R;
    a.G = pl.G;
    a.S = L.S;
    a.nks = pl.is_float ? (pl.d + 15) / 16 : (pl.d + 31) / 32;
    a.kblocks = pl.kblocks;
    a.stages = pl.stages;
    a.qcap = pl.qcap;
    a.k = k;
    a.cap = pl.cap;
    a.lists_smem = pl.lists_smem;
    a.idesc = pl.is_float ? tc::make_idesc_bf16() : tc::make_idesc_i8(pl.is_signed);
    a.cand = (uint64_t*) ctx->cand.p;
    a.partial = L.carry_keys + L.part_keys;
    a.bound = pl.bound;
    a.npeers = pl.peers ? ctx->npeers : 0;
    for (int i = 0; i < a.npeers; ++i) a.peer_bound[i] = ctx->peer_bound[i];
    a.mode = ctx->tc_mode;
    a.rows_f32 = pl.is_float ? lay.rows_f32 : nullptr;
    a.q_f32 = pl.is_float ? (const float*) ctx->qf32.p : nullptr;
    a.d = pl.d;
    a.id_base = lay.id_base;
    a.eps = 0.0078125f + 0.00006103515625f;  // 2^-7 (bf16 rounding of both operands) + 2^-14 (fp32 accumulation, K <= 128)
    a.gate = pl.gate;
    a.mips = pl.mips;
    L.grid = (unsigned) (pl.G * L.S);
    L.smem = tc::smem_bytes(pl.R, pl.kblocks, pl.stages, pl.qcap, pl.lists_smem);
    // find log: several point splits serve each query group.  Small k only (lists in shared memory): at k = 100 every CTA
    // re-applying the whole group's finds to lists in L2 costs more than the tighter thresholds save (C2: 65 -> 78 ms).
    const bool no_flog = std::getenv("GPANN_TC_NOLOG") != nullptr;  // (not cached: the parity test toggles it)
    if (!no_flog && pl.lists_smem && L.S > 1 && ctx->tc_mode == 0) {
        // ~k (1 + ln(n / sample)) finds per query are expected; an overflowing log only loses sharing
        a.flog_cap = (uint32_t) (pl.R * tc::kTileM) * (uint32_t) std::min(1024, std::max(64, 8 * k));
        L.flog_tail_bytes = ((size_t) pl.G * sizeof(uint32_t) + 255) & ~(size_t) 255;
        L.flog_bytes = (size_t) pl.G * a.flog_cap * sizeof(uint4);
        GP_TRY(reserve(ctx, ctx->flog, L.flog_tail_bytes + L.flog_bytes));
        a.flog_tail = (uint32_t*) ctx->flog.p;
        a.flog = (uint4*) ((char*) ctx->flog.p + L.flog_tail_bytes);
    }
    // Interleaved splits (split s scans tiles s, s + S, ...): with contiguous norm ranges the splits whose norms sit near the
    // queries' take most of the candidate records and the others wait for them (12.5M-row shard 12.9 -> 12.4 ms, C1 7.4 -> 6.4 ms)
    const char* e_il = std::getenv("GPANN_TC_INTERLEAVE");  // experiment knob: 0 = contiguous ranges
    a.interleave = e_il ? std::atoi(e_il) != 0 : 1;
    // k > 32 (append + compaction lists): interleaved splits that publish their ceil(k/S)-th best distance
    const bool no_qstat = std::getenv("GPANN_TC_NOQSTAT") != nullptr;
    const bool qstat_small = std::getenv("GPANN_TC_QSTAT_SMALL") != nullptr;  // experiment: also for k <= 32 (where ceil(k/S) <= 2)
    if (!no_qstat && (k > 32 || (qstat_small && (k + L.S - 1) / L.S <= 2)) && L.S > 1 && ctx->tc_mode == 0) {
        L.qstat_bytes = (size_t) pl.G * L.S * pl.R * tc::kTileM * sizeof(uint32_t);
        GP_TRY(reserve(ctx, ctx->qstat, L.qstat_bytes));
        a.qstat = (uint32_t*) ctx->qstat.p;
    }
    return 0;
}

// Pre-passes over strided samples of tiles: the k-th distances found on a sample are valid upper bounds on the final
// ones and let every CTA of the next pass start with a tight filter threshold instead of a cold one (a cold start pushes
// EVERY distance through the selection warps).  Up to three levels: 1 tile per split (cold, but tiny), then ~12 and ~160
// tiles per split, each filtered by the previous level's bound.  The bounds land in ctx->bound [nq_pad] (float bits);
// a sharded caller all-reduces them (min) across its ranks before the main pass (multi.cu).
int tc_sample(gpann_ctx* ctx, const TcPlan& pl, const TcLayout& lay, cudaStream_t st) {
    if (!ctx->tc_prepass || ctx->tc_mode != 0 || (ctx->tc_warm && ctx->tc_warm_valid)) return 0;
    TcLaunch L;
    GP_TRY(tc_setup(ctx, pl, lay, L));
    tc::Args& a = L.a;
    const int S = L.S, k = pl.k;
    // level sizes follow the shard: a third level of 160 tiles per split is 8 % of a 12.5M-row shard's work
    const uint32_t per_split_tiles = L.tiles_total / (uint32_t) S;
    uint32_t l3 = std::min(160u, per_split_tiles / 96u);
    if (l3 < 24u) l3 = 0;  // small splits: two levels (measured: a third costs more than it saves at 1M points)
    uint32_t levels[4] = { 1, l3 ? std::max(4u, std::min(12u, l3 / 12u)) : 12u, l3, 0 };  // (a fourth level: experiments only)
    // float rows: every record of a cold pass costs an exact re-rank from HBM, and a cold second level costs no more than a
    // cold first one (C1: 7.3 -> 6.3 ms without the 1-tile level)
    if (pl.is_float) levels[0] = 0;
    const char* e_lv = std::getenv("GPANN_TC_LEVELS");  // experiment / test knob: "a/b/c" tiles per split (0 = skip the level)
    if (e_lv) std::sscanf(e_lv, "%u/%u/%u/%u", &levels[0], &levels[1], &levels[2], &levels[3]);
    for (uint32_t per_split : levels) {
        if (per_split == 0) continue;
        const uint32_t nsample = (uint32_t) S * per_split;
        if (L.tiles_total < 32u * nsample || (size_t) nsample * tc::kTileN < (size_t) 4 * k) continue;
        a.ntiles = nsample;
        a.tile_stride = L.tiles_total / nsample;
        a.tile_lo = a.tile_stride / 2;
        a.sample_pass = 1;
        // no find log in the sample passes: they are bound by the selection warps (a cold pass accepts nearly every record),
        // and sharing would multiply exactly that work by the number of splits (measured: 2.0M -> 5.9M clk for the three levels)
        static const bool log_samples = std::getenv("GPANN_TC_LOG_SAMPLE") != nullptr;
        if (!log_samples) a.flog = nullptr;
        static const char* e_sst = std::getenv("GPANN_TC_STATS");
        if (e_sst && std::atoi(e_sst) >= 2) {
            GP_TRY(reserve(ctx, ctx->misc, 256));
            CU_TRY(ctx, cudaMemsetAsync(ctx->misc.p, 0, 256, st));
            a.stats = (unsigned long long*) ctx->misc.p;
        }
        GP_TRY(tc_flog_reset(ctx, L, st));
        launch_topk_tc(pl.is_float, *lay.tmap, pl.tmap_q, a, L.grid, L.smem, st);
        GP_TRY(check_launch(ctx, "topk_tc_kernel(sample)"));
        if (a.stats) {
            unsigned long long h[32] = { 0 };
            cudaStreamSynchronize(st);
            cudaMemcpy(h, ctx->misc.p, 256, cudaMemcpyDeviceToHost);
            std::fprintf(stderr, "[gpann stats] sample pass %u tiles/split: CTA 0 clk: filter loop done %llu, selection drained %llu, final selection done %llu; "
                                 "records %llu; mma thread clk %llu | CTA 0 selection warp 0: %llu batches, %llu rounds, insert clk %llu, pop clk %llu, idle polls %llu | slowest CTA: set-up %llu, filter loop %llu, whole kernel %llu clk\n",
                         per_split, h[10], h[11], h[12], h[0], h[2] / L.grid, h[16], h[17], h[18], h[20], h[19], h[22], h[23], h[21]);
            a.stats = nullptr;
        }
        TopkOut b;
        b.kth = (float*) pl.bound;      // float bits: read back by the next pass as its bound
        b.kth_full_only = 1;            // min(previous bound, k-th of a full list)
        GP_TRY(launch_merge(ctx, a.partial, S, pl.nq_pad, pl.nqb, k, b, 0, st, pl.gate, 1));
    }
    return 0;
}

// The full pass over one layout for the prepared batch.  The per-split partial lists land in ctx->partial
// [1 + S, nq_pad, k] (parts 1..S); with `carry` part 0 holds the running result of earlier chunks (and receives the merged
// one, together with the tightened per-query bounds); otherwise the merged result of the S parts goes to `out` at query
// offset q_off.
int tc_main(gpann_ctx* ctx, const TcPlan& pl, const TcLayout& lay, bool carry, bool carry_init, const TopkOut& out, size_t q_off,
            cudaStream_t st) {
    TcLaunch L;
    GP_TRY(tc_setup(ctx, pl, lay, L));
    tc::Args& a = L.a;
    const int S = L.S, k = pl.k;
    if (carry && carry_init) CU_TRY(ctx, cudaMemsetAsync(L.carry_keys, 0xFF, L.part_keys * sizeof(uint64_t), st));
    a.tile_lo = 0;
    a.tile_stride = 1;
    a.ntiles = L.tiles_total;
    a.sample_pass = 0;
    static const bool no_pace = std::getenv("GPANN_TC_NOPACE") != nullptr;
    const char* e_pw = std::getenv("GPANN_TC_PACE_WINDOW");  // tests: a tight window exercises the pacing protocol on small sets
    a.pace_window = e_pw ? (uint32_t) std::max(1, std::atoi(e_pw)) : tc::kPaceWindow;
    if (!no_pace && L.tiles_total / S >= 4 * a.pace_window && pl.G > 1 && (int) L.grid <= ctx->num_sms) {  // all CTAs co-resident
        GP_TRY(reserve(ctx, ctx->pace, (size_t) L.grid * sizeof(uint32_t)));
        CU_TRY(ctx, cudaMemsetAsync(ctx->pace.p, 0, (size_t) L.grid * sizeof(uint32_t), st));
        a.pace = (uint32_t*) ctx->pace.p;
    }
    static const bool want_stats = std::getenv("GPANN_TC_STATS") != nullptr;
    if (want_stats) {
        GP_TRY(reserve(ctx, ctx->misc, 256));
        CU_TRY(ctx, cudaMemsetAsync(ctx->misc.p, 0, 256, st));
        a.stats = (unsigned long long*) ctx->misc.p;
    }
    GP_TRY(tc_flog_reset(ctx, L, st));
    if (ctx->timing) cudaEventRecord(ctx->ev0, st);
    launch_topk_tc(pl.is_float, *lay.tmap, pl.tmap_q, a, L.grid, L.smem, st);
    if (ctx->timing) {
        cudaEventRecord(ctx->ev1, st);
        ctx->timed = true;
    }
    GP_TRY(check_launch(ctx, "topk_tc_kernel"));
    if (want_stats) {
        unsigned long long h[32] = { 0 };
        cudaStreamSynchronize(st);
        cudaMemcpy(h, ctx->misc.p, 256, cudaMemcpyDeviceToHost);
        const double g = (double) L.grid, items = (double) L.tiles_total / S * pl.R * tc::kHalves;
        std::fprintf(stderr, "[gpann stats] main pass: %llu records pushed, %llu slow-path entries (warp x chunk pair), grid %u, %.0f items/CTA\n", h[0], h[1], L.grid, items);
        std::fprintf(stderr, "[gpann stats] main pass CTA 0 clk: filter loop done %llu, selection drained %llu, final selection done %llu | slowest CTA: set-up %llu, "
                             "filter loop %llu, whole kernel %llu clk\n", h[10], h[11], h[12], h[22], h[23], h[21]);
        std::fprintf(stderr, "[gpann stats] clk per item (CTA average): MMA thread total %.0f = wait operands %.0f + wait TMEM slot %.0f + issue %.0f | "
                             "filter warp 0 total %.0f = wait accumulator %.0f + wait tile %.0f + work %.0f | TMA thread total %.0f, wait free stage %.0f\n",
                     h[2] / g / items, h[3] / g / items, h[4] / g / items, (h[2] - h[3] - h[4]) / g / items, h[5] / g / items, h[6] / g / items,
                     h[7] / g / items, (h[5] - h[6] - h[7]) / g / items, h[9] / g / items, h[8] / g / items);
    }
    if (!carry) {
        if (ctx->tc_warm) {  // probe: the final k-th distances become the next call's bounds
            TopkOut b;
            b.kth = (float*) pl.bound;
            b.kth_full_only = 1;
            GP_TRY(launch_merge(ctx, a.partial, S, pl.nq_pad, pl.nqb, k, b, 0, st, pl.gate, 1));
            ctx->tc_warm_valid = true;
        }
        return launch_merge(ctx, a.partial, S, pl.nq_pad, pl.nqb, k, out, q_off, st, pl.gate, 1);
    }
    TopkOut c;
    c.keys = L.carry_keys;            // in place: a warp reads all parts of its query before it writes
    c.kth = (float*) pl.bound;
    c.kth_full_only = 1;
    return launch_merge(ctx, L.carry_keys, S + 1, pl.nq_pad, pl.nqb, k, c, 0, st, pl.gate, 1);
}

int tc_scan(gpann_ctx* ctx, const TcPlan& pl, const TcLayout& lay, bool prepass, bool carry, bool carry_init, const TopkOut& out,
            size_t q_off, cudaStream_t st) {
    if (prepass) GP_TRY(tc_sample(ctx, pl, lay, st));
    return tc_main(ctx, pl, lay, carry, carry_init, out, q_off, st);
}

TcLayout layout_of(const gpann_points* pts) {
    TcLayout lay;
    lay.tmap = &pts->tmap;
    lay.norms = pts->t_norms;
    lay.ids = pts->t_ids;
    lay.n_pad = pts->n_pad;
    lay.rows_f32 = pts->storage == GPANN_STORAGE_F32 ? (const float*) pts->rows : nullptr;
    lay.id_base = pts->id_base;
    return lay;
}

// All query batches of one call on the tensor path (asynchronous).  *gate receives the device flag the byte path's kernels
// were gated on (nullptr for float storage): the caller queues the exact float scan behind, gated the other way.
// `between` (optional) runs between the sample passes and the main pass of every batch: the sharded entry points
// all-reduce the per-query bounds across their ranks there.
int topk_tensor(gpann_ctx* ctx, const gpann_points* pts, const float* dq, size_t nq, int k, int metric, const TopkOut& out,
                cudaStream_t st, const int** gate, const BetweenFn& between) {
    const bool is_float = pts->storage == GPANN_STORAGE_F32;
    const int d = (int) pts->d;
    const size_t batch_max = tc_batch_max(ctx, is_float ? pts->dp / 64 : 1);
    const TcLayout lay = layout_of(pts);
    *gate = nullptr;
    for (size_t q0 = 0; q0 < nq; q0 += batch_max) {
        TcPlan pl;
        GP_TRY(tc_prepare_queries(ctx, pl, pts->storage, d, pts->dp, dq + q0 * d, std::min(batch_max, nq - q0), k, st, metric));
        *gate = pl.gate;
        GP_TRY(tc_sample(ctx, pl, lay, st));
        if (between) GP_TRY(between(pl.bound, pl.nqb, st));  // real queries only: the count every rank agrees on
        GP_TRY(tc_main(ctx, pl, lay, false, false, out, q0, st));
    }
    ctx->last_algo = is_float ? "tensor_bf16" : (pts->storage == GPANN_STORAGE_S8 ? "tensor_s8" : "tensor_u8");
    return 0;
}

// ---- streamed scan of a HOST-resident point set -----------------------------------------------------------------
// The set never sits in HBM as a whole: chunks of `chunk_rows` rows go H2D on a second stream into one of two raw
// buffers while the previous chunk is turned into a norm-sorted layout and scanned; the running top-k and the per-query
// bounds carry from chunk to chunk (tc_scan's carry part), so every chunk after the first starts with a tight filter.
// Host rows are bytes (u8/s8) or floats; float rows that hold byte values (the reference's reader widens .u8bin/.i8bin
// to float, src/points_io.cpp:63-111) are narrowed on the device chunk by chunk and take the exact integer kernel,
// other float rows take the bf16 filter + exact fp32 re-rank against the raw chunk.

int topk_stream_host(gpann_ctx* ctx, const void* host_rows, int host_kind, int storage, size_t n, size_t d, uint32_t id_base,
                     const float* dq, size_t nq, int k, size_t chunk_rows, const TopkOut& out, cudaStream_t st) {
    if (d == 0 || d > (size_t) tc::kRowBytes) return fail(ctx, GPANN_E_UNSUPPORTED, "topk_stream: d must be in [1,128]");
    if (k < 1 || k > 448) return fail(ctx, GPANN_E_ARG, "topk_stream: k must be in [1,448]");
    if ((uint64_t) id_base + n > 0xFFFFFFFFull) return fail(ctx, GPANN_E_ARG, "topk_stream: ids exceed 32 bits (recall.h:30)");
    if (nq == 0) return 0;
    if (!ctx->encode_tiled) return fail(ctx, GPANN_E_UNSUPPORTED, "topk_stream: cuTensorMapEncodeTiled unavailable");
    if (!ctx->copy_stream) {
        CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
            CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_consumed[i], cudaEventDisableTiming));
        }
    }
    const bool host_f32 = host_kind == kHostF32;
    const size_t esz = host_f32 ? 4 : 1;
    const size_t tile = tc::kTileN;
    // default chunk: ~1/8 of the set, between 64 MB and 512 MB of host rows -- large enough to amortise the per-chunk layout
    // build, small enough that a 1.6 GB shard (100M x 128 bytes over 8 GPUs) still has ~8 chunks to overlap H2D with the scan
    const size_t total_bytes = n * d * esz;
    const size_t chunk_bytes = std::min<size_t>((size_t) 512 << 20, std::max<size_t>((size_t) 64 << 20, total_bytes / 8));
    size_t C = chunk_rows ? chunk_rows : chunk_bytes / (d * esz);
    C = std::min(((C + tile - 1) / tile) * tile, ((std::max<size_t>(n, 1) + tile - 1) / tile) * tile);
    const size_t nchunks = std::max<size_t>(1, (n + C - 1) / C);
    const char* src = (const char*) host_rows;

    // storage of the scan: bytes for byte hosts; for float hosts AUTO looks at the first chunk (a later chunk that
    // breaks the guess restarts the call on the float path)
    if (!host_f32) storage = host_kind == kHostS8 ? GPANN_STORAGE_S8 : GPANN_STORAGE_U8;
    bool chunk0_resident = false;  // AUTO detection leaves chunk 0 in raw[0]: the first round does not copy it again
    for (int b = 0; b < 2; ++b) GP_TRY(reserve(ctx, ctx->st_raw[b], C * d * esz));
    GP_TRY(reserve(ctx, ctx->flags, 64));
    if (host_f32 && storage == GPANN_STORAGE_AUTO) {
        const size_t rows0 = std::min(C, n);
        int flags_h[2] = { 1, 1 };
        if (rows0) {
            CU_TRY(ctx, cudaStreamWaitEvent(st, ctx->ev_consumed[0], 0));
            CU_TRY(ctx, cudaMemcpyAsync(ctx->st_raw[0].p, src, rows0 * d * 4, cudaMemcpyHostToDevice, st));
            CU_TRY(ctx, cudaMemcpyAsync(ctx->flags.p, flags_h, 8, cudaMemcpyHostToDevice, st));
            launch_detect_bytes((const float*) ctx->st_raw[0].p, rows0 * d, (int*) ctx->flags.p, blocks_for(ctx, rows0 * d), st);
            GP_TRY(check_launch(ctx, "detect_bytes_kernel"));
            CU_TRY(ctx, cudaMemcpyAsync(flags_h, ctx->flags.p, 8, cudaMemcpyDeviceToHost, st));
            CU_TRY(ctx, cudaEventRecord(ctx->ev_consumed[0], st));
            CU_TRY(ctx, cudaStreamSynchronize(st));
            chunk0_resident = true;
        }
        storage = flags_h[0] ? GPANN_STORAGE_U8 : (flags_h[1] ? GPANN_STORAGE_S8 : GPANN_STORAGE_F32);
    }
    for (;;) {  // at most two rounds: a byte guess refuted by a later chunk falls back to the float path
        const bool is_float = storage == GPANN_STORAGE_F32;
        const int is_signed = storage == GPANN_STORAGE_S8;
        const int dp = is_float ? (int) ((d + 63) / 64) * 64 : 0;
        if (nq > tc_batch_max(ctx, is_float ? dp / 64 : 1))
            return fail(ctx, GPANN_E_UNSUPPORTED, "topk_stream: too many queries for one batch");
        GP_TRY(reserve(ctx, ctx->st_rows, C * (is_float ? (size_t) dp * 2 : (size_t) tc::kRowBytes)));
        GP_TRY(reserve(ctx, ctx->st_norms, C * 4));
        GP_TRY(reserve(ctx, ctx->st_ids, C * 4));
        GP_TRY(reserve(ctx, ctx->st_a, C * 4));
        GP_TRY(reserve(ctx, ctx->st_b, C * 4));
        GP_TRY(reserve(ctx, ctx->st_c, C * 4));
        GP_TRY(reserve(ctx, ctx->st_d, C * 4));
        if (host_f32 && !is_float) GP_TRY(reserve(ctx, ctx->st_narrow, C * d));
        const int sort_bits = is_float ? 32 : 25;
        size_t tmp_bytes = 0;
        sort_pairs_u32(nullptr, tmp_bytes, (uint32_t*) ctx->st_a.p, (uint32_t*) ctx->st_b.p, (uint32_t*) ctx->st_c.p,
                       (uint32_t*) ctx->st_d.p, C, sort_bits, st);
        GP_TRY(reserve(ctx, ctx->st_tmp, tmp_bytes));
        TcPlan pl;
        int ok = 1;
        GP_TRY(tc_prepare_queries(ctx, pl, storage, (int) d, dp, dq, nq, k, st));
        if (pl.gate) {  // byte scan: the one synchronisation of the streamed path decides between the integer and the float scan
            CU_TRY(ctx, cudaMemcpyAsync(&ok, pl.gate, sizeof(int), cudaMemcpyDeviceToHost, st));
            CU_TRY(ctx, cudaStreamSynchronize(st));
            pl.gate = nullptr;  // decided on the host: no gating inside the chunk loop
        }
        if (!ok) {
            if (!host_f32) return fail(ctx, GPANN_E_UNSUPPORTED, "topk_stream: queries are not integer-valued in the storage range");
            storage = GPANN_STORAGE_F32;  // byte-valued points, fractional queries: exact float path
            continue;
        }
        CUtensorMap tmap;
        GP_TRY(encode_rows_map(ctx, &tmap, ctx->st_rows.p, C, tc::kTileN, "chunk", is_float, is_float ? dp : tc::kRowBytes));
        TcLayout lay;
        lay.tmap = &tmap;
        lay.norms = ctx->st_norms.p;
        lay.ids = (const uint32_t*) ctx->st_ids.p;
        lay.rows_f32 = nullptr;
        if (host_f32 && !is_float) {
            int flags_h[2] = { 1, 1 };
            CU_TRY(ctx, cudaMemcpyAsync(ctx->flags.p, flags_h, 8, cudaMemcpyHostToDevice, st));
        }
        for (size_t c = 0; c < nchunks; ++c) {
            const int b = (int) (c & 1);
            const size_t r0 = c * C, rows_c = std::min(C, n - std::min(n, r0));
            const size_t n_pad = std::max(tile, ((rows_c + tile - 1) / tile) * tile);
            const uint32_t chunk_base = id_base + (uint32_t) r0;
            unsigned char* raw = (unsigned char*) ctx->st_raw[b].p;
            // raw[b] is free once the last consumer of the chunk that used it before (this or an earlier call) is done
            CU_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_consumed[b], 0));
            if (rows_c && !(c == 0 && chunk0_resident))
                CU_TRY(ctx, cudaMemcpyAsync(raw, src + r0 * d * esz, rows_c * d * esz, cudaMemcpyHostToDevice, ctx->copy_stream));
            if (c == 0) chunk0_resident = false;
            CU_TRY(ctx, cudaEventRecord(ctx->ev_copied[b], ctx->copy_stream));
            CU_TRY(ctx, cudaStreamWaitEvent(st, ctx->ev_copied[b], 0));
            uint32_t *ka = (uint32_t*) ctx->st_a.p, *kb = (uint32_t*) ctx->st_b.p, *va = (uint32_t*) ctx->st_c.p, *vb = (uint32_t*) ctx->st_d.p;
            const unsigned char* bytes = raw;
            if (rows_c && host_f32 && !is_float) {  // float host rows holding byte values: check and narrow this chunk
                launch_detect_bytes((const float*) raw, rows_c * d, (int*) ctx->flags.p, blocks_for(ctx, rows_c * d), st);
                launch_narrow_rows((const float*) raw, rows_c, (int) d, (int) d, (unsigned char*) ctx->st_narrow.p, blocks_for(ctx, rows_c * d), st);
                GP_TRY(check_launch(ctx, "detect_bytes_kernel+narrow_rows_kernel", 2));
                bytes = (const unsigned char*) ctx->st_narrow.p;
            }
            if (rows_c) {
                if (is_float) launch_float_norms((const float*) raw, rows_c, (int) d, ka, va, blocks_for(ctx, rows_c), st);
                else launch_byte_norms(bytes, rows_c, (int) d, d, is_signed, ka, va, ctx->num_sms, st);
                GP_TRY(check_launch(ctx, "norms_kernel"));
                size_t tb = ctx->st_tmp.cap;
                if (sort_pairs_u32(ctx->st_tmp.p, tb, ka, kb, va, vb, rows_c, sort_bits, st) != cudaSuccess)
                    return fail(ctx, GPANN_E_CUDA, "topk_stream: radix sort failed");
                ctx->launches += 4;
            }
            if (is_float)
                launch_gather_sorted_rows_bf16((const float*) raw, rows_c, (int) d, dp, vb, kb, chunk_base, n_pad, (unsigned short*) ctx->st_rows.p,
                                               (float*) ctx->st_norms.p, (uint32_t*) ctx->st_ids.p, blocks_for(ctx, n_pad * 32, 256, 32), st);
            else
                launch_gather_sorted_rows(bytes, rows_c, (int) d, d, vb, kb, chunk_base, n_pad, (unsigned char*) ctx->st_rows.p,
                                          (int32_t*) ctx->st_norms.p, (uint32_t*) ctx->st_ids.p, blocks_for(ctx, n_pad * 32, 256, 32), st);
            GP_TRY(check_launch(ctx, "gather_sorted_rows_kernel"));
            if (!is_float) CU_TRY(ctx, cudaEventRecord(ctx->ev_consumed[b], st));
            lay.n_pad = n_pad;
            lay.id_base = chunk_base;
            lay.rows_f32 = is_float ? (const float*) raw : nullptr;
            GP_TRY(tc_scan(ctx, pl, lay, c == 0, true, c == 0, out, 0, st));
            if (is_float) CU_TRY(ctx, cudaEventRecord(ctx->ev_consumed[b], st));  // the re-rank reads the raw rows
        }
        if (host_f32 && !is_float) {
            int flags_h[2] = { 1, 1 };
            CU_TRY(ctx, cudaMemcpyAsync(flags_h, ctx->flags.p, 8, cudaMemcpyDeviceToHost, st));
            CU_TRY(ctx, cudaStreamSynchronize(st));
            if (!flags_h[is_signed ? 1 : 0]) {  // some later chunk is not byte-valued after all
                storage = GPANN_STORAGE_F32;
                continue;
            }
        }
        ctx->last_algo = is_float ? "tensor_bf16_stream" : (is_signed ? "tensor_s8_stream" : "tensor_u8_stream");
        return launch_merge(ctx, (const uint64_t*) ctx->partial.p, 1, pl.nq_pad, nq, k, out, 0, st, nullptr, 0);
    }
}

int topk_stream_host_api(gpann_ctx* ctx, const void* host_rows, int host_kind, int storage, size_t n, size_t d, uint32_t id_base,
                         const float* queries, size_t nq, int k, size_t chunk_rows, float* out_dist, uint32_t* out_id,
                         uint32_t* out_count, float* out_kth) {
    if ((!host_rows && n) || !queries) return fail(ctx, GPANN_E_ARG, "topk_stream: null argument");
    if (nq == 0) return 0;
    if (k < 1 || k > 448) return fail(ctx, GPANN_E_ARG, "topk_stream: k must be in [1,448]");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    OrderScope order(ctx, st);
    GP_TRY(reserve(ctx, ctx->qdev, nq * d * 4));
    GP_TRY(reserve(ctx, ctx->dist_out, nq * k * 4));
    GP_TRY(reserve(ctx, ctx->id_out, nq * k * 4));
    GP_TRY(reserve(ctx, ctx->cnt_out, nq * 8));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->qdev.p, queries, nq * d * 4, cudaMemcpyHostToDevice, st));
    TopkOut o;
    o.dist = out_dist ? (float*) ctx->dist_out.p : nullptr;
    o.id = out_id ? (uint32_t*) ctx->id_out.p : nullptr;
    o.count = out_count ? (uint32_t*) ctx->cnt_out.p : nullptr;
    o.kth = out_kth ? (float*) ctx->cnt_out.p + nq : nullptr;
    GP_TRY(topk_stream_host(ctx, host_rows, host_kind, storage, n, d, id_base, (const float*) ctx->qdev.p, nq, k, chunk_rows, o, st));
    if (out_dist) CU_TRY(ctx, cudaMemcpyAsync(out_dist, o.dist, nq * k * 4, cudaMemcpyDeviceToHost, st));
    if (out_id) CU_TRY(ctx, cudaMemcpyAsync(out_id, o.id, nq * k * 4, cudaMemcpyDeviceToHost, st));
    if (out_count) CU_TRY(ctx, cudaMemcpyAsync(out_count, o.count, nq * 4, cudaMemcpyDeviceToHost, st));
    if (out_kth) CU_TRY(ctx, cudaMemcpyAsync(out_kth, o.kth, nq * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    return 0;
}

// Is the tensor path the one a call of this SHAPE takes?  Depends only on values every shard of a sharded call agrees on
// (storage, d, nq, k, metric, algo, full range) -- not on the local row count -- so that all ranks of a communicator
// issue the same sequence of collectives (multi.cu).
bool tensor_shape(const gpann_ctx* ctx, int storage, size_t d, size_t nq, int k, int metric, bool full) {
    const bool is_float = storage == GPANN_STORAGE_F32;
    // both metrics: 1 - <q,x> is the same contraction without the norms (byte data: exact integers; float data: bf16 filter
    // with the error bound eps*|q|*|x| + exact sequential-FMA re-rank)
    (void) metric;
    if (!full || k > 448 || d > (size_t) (is_float ? 64 * tc::kMaxKBlocks : tc::kRowBytes)) return false;
    return ctx->algo == GPANN_ALGO_TENSOR || (ctx->algo == GPANN_ALGO_AUTO && nq > 16);
}
int topk_dispatch(gpann_ctx* ctx, const gpann_points* pts, size_t row_begin, size_t row_end, const float* dq, size_t nq, int k,
                  int metric, const TopkOut& out, cudaStream_t st, const BetweenFn& between) {
    if (!pts || pts->ctx != ctx) return fail(ctx, GPANN_E_ARG, "topk: point set does not belong to this context");
    if (k < 1 || k > 512) return fail(ctx, GPANN_E_ARG, "topk: k must be in [1,512]");
    if (metric != GPANN_METRIC_L2 && metric != GPANN_METRIC_MIPS) return fail(ctx, GPANN_E_ARG, "topk: bad metric");
    if (row_begin > row_end || row_end > pts->n) return fail(ctx, GPANN_E_ARG, "topk: bad row range");
    if (nq == 0) return 0;
    const bool full = row_begin == 0 && row_end == pts->n;
    const bool is_float = pts->storage == GPANN_STORAGE_F32;
    const bool shape_ok = tensor_shape(ctx, pts->storage, pts->d, nq, k, metric, full);
    // small float sets are not worth a layout -- unless a sharded call needs every rank on the same path
    const bool want_tensor = shape_ok && (ctx->algo == GPANN_ALGO_TENSOR || !is_float || pts->n >= 4096 || between);
    if (want_tensor && is_float && !pts->tensor_tried) {  // lazily build the bf16 layout (handle state, not query state)
        const int rc = build_tensor_float(ctx, const_cast<gpann_points*>(pts), st);
        // AUTO: no memory for the layout means "the exact streaming kernel serves this set", on the first call as on later ones
        if (rc != 0 && !(ctx->algo == GPANN_ALGO_AUTO && (rc == GPANN_E_OOM || rc == GPANN_E_UNSUPPORTED))) return rc;
        if (rc != 0) cudaGetLastError();
    }
    const bool tensor_ok = want_tensor && pts->has_tensor;
    if (ctx->algo == GPANN_ALGO_TENSOR && !tensor_ok)
        return fail(ctx, GPANN_E_UNSUPPORTED, "topk: tensor kernel needs d<=128, full range, k<=448");
    const int* gate = nullptr;
    if (tensor_ok) {
        GP_TRY(topk_tensor(ctx, pts, dq, nq, k, metric, out, st, &gate, between));
        if (!gate) return 0;
        ctx->gate_used = true;  // float storage: nothing can refute the tensor path
        // byte storage: the exact float scan is queued behind, gated on "queries were NOT integer-valued" -- normally every
        // CTA of it returns at once.  Nothing synchronises: the call stays asynchronous on `st`.
    } else if (between && shape_ok) {
        // a rank of a sharded call without a local layout (empty shard, no memory for it): the same bound exchange as the
        // ranks on the tensor path (one per call: the sharded caller keeps nq within one batch), contributing "no bound"
        GP_TRY(reserve(ctx, ctx->bound, nq * 4));
        CU_TRY(ctx, cudaMemsetAsync(ctx->bound.p, 0x7f, nq * 4, st));  // 0x7f7f7f7f: a huge finite float
        GP_TRY(between((uint32_t*) ctx->bound.p, nq, st));
    }
    if (row_end == row_begin) {
        // empty scan: all-empty output
        GP_TRY(reserve(ctx, ctx->partial, nq * (size_t) k * 8));
        CU_TRY(ctx, cudaMemsetAsync(ctx->partial.p, 0xFF, nq * (size_t) k * 8, st));
        ctx->last_algo = "empty";
        return launch_merge(ctx, (const uint64_t*) ctx->partial.p, 1, nq, nq, k, out, 0, st, nullptr, 0);
    }
    return topk_stream(ctx, pts, row_begin, row_end, dq, nq, k, metric, out, st, gate, 0);
}

// Host-side check after a synchronised call: did byte points meet fractional queries (the gated float scan served them)?
int queries_were_integer(gpann_ctx* ctx, cudaStream_t st, int* ok) {
    *ok = 1;
    if (!ctx->flags.p) return 0;
    CU_TRY(ctx, cudaMemcpyAsync(ok, (const int*) ctx->flags.p + kFlagQueriesOk, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    return 0;
}

int topk_host(gpann_ctx* ctx, const gpann_points* pts, size_t row_begin, size_t row_end, const float* queries, size_t nq, int k,
              int metric, float* out_dist, uint32_t* out_id, uint32_t* out_count, float* out_kth) {
    if (!ctx) return GPANN_E_ARG;
    if (!pts || !queries) return fail(ctx, GPANN_E_ARG, "topk: null argument");
    if (nq == 0) return 0;
    if (k < 1 || k > 512) return fail(ctx, GPANN_E_ARG, "topk: k must be in [1,512]");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    OrderScope order(ctx, st);
    const size_t d = pts->d;
    GP_TRY(reserve(ctx, ctx->qdev, nq * d * 4));
    GP_TRY(reserve(ctx, ctx->dist_out, nq * k * 4));
    GP_TRY(reserve(ctx, ctx->id_out, nq * k * 4));
    GP_TRY(reserve(ctx, ctx->cnt_out, nq * 8));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->qdev.p, queries, nq * d * 4, cudaMemcpyHostToDevice, st));
    TopkOut o;
    o.dist = out_dist ? (float*) ctx->dist_out.p : nullptr;
    o.id = out_id ? (uint32_t*) ctx->id_out.p : nullptr;
    o.count = out_count ? (uint32_t*) ctx->cnt_out.p : nullptr;
    o.kth = out_kth ? (float*) ctx->cnt_out.p + nq : nullptr;
    ctx->gate_used = false;
    GP_TRY(topk_dispatch(ctx, pts, row_begin, row_end, (const float*) ctx->qdev.p, nq, k, metric, o, st, nullptr));
    if (out_dist) CU_TRY(ctx, cudaMemcpyAsync(out_dist, o.dist, nq * k * 4, cudaMemcpyDeviceToHost, st));
    if (out_id) CU_TRY(ctx, cudaMemcpyAsync(out_id, o.id, nq * k * 4, cudaMemcpyDeviceToHost, st));
    if (out_count) CU_TRY(ctx, cudaMemcpyAsync(out_count, o.count, nq * 4, cudaMemcpyDeviceToHost, st));
    if (out_kth) CU_TRY(ctx, cudaMemcpyAsync(out_kth, o.kth, nq * 4, cudaMemcpyDeviceToHost, st));
    int ok = 1;
    if (ctx->gate_used) GP_TRY(queries_were_integer(ctx, st, &ok));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    if (!ok) {  // byte points, fractional queries: the gated exact float scan produced the result
        if (ctx->algo == GPANN_ALGO_TENSOR) return fail(ctx, GPANN_E_UNSUPPORTED, "topk: queries are not integer-valued in the storage range");
        ctx->last_algo = "stream";
    }
    return 0;
}

// Row-major [rows, row_elems] bf16 matrix fetched as boxes of box_rows x 32 elements (64 bytes), SWIZZLE_64B.
int encode_rows_map_sw64(gpann_ctx* ctx, CUtensorMap* map, void* base, size_t rows, int row_elems, int box_rows, const char* what) {
    const cuuint64_t gdim[2] = { (cuuint64_t) row_elems, (cuuint64_t) rows };
    const cuuint64_t gstride[1] = { (cuuint64_t) row_elems * 2 };
    const cuuint32_t box[2] = { (cuuint32_t) kmtc::kKBlockElems, (cuuint32_t) box_rows };
    const cuuint32_t estr[2] = { 1, 1 };
    CUresult r = ctx->encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(ctx, GPANN_E_CUDA, std::string("cuTensorMapEncodeTiled(") + what + ") failed: " + std::to_string((int) r));
    return 0;
}

// Lazily built by the first tensor-path k-means call over a float point set: split-precision bf16 rows + |x|.
int build_kmeans_layout(gpann_ctx* ctx, gpann_points* pts, cudaStream_t st) {
    pts->km_tried = true;
    const size_t n = pts->n;
    const int d = (int) pts->d;
    if (n == 0 || d > 128 || pts->storage != GPANN_STORAGE_F32 || !ctx->encode_tiled || n > 0xFFFFFF00ull) return 0;
    const int dpad = ((d + 31) / 32) * 32;
    const size_t n_pad = ((n + kmtc::kTileM - 1) / kmtc::kTileM) * kmtc::kTileM;
    if (cudaMalloc(&pts->km_rows, n_pad * (size_t) dpad * 4) != cudaSuccess || cudaMalloc(&pts->km_xnorm, (n_pad + 4) * 4) != cudaSuccess) {
        cudaGetLastError();
        if (pts->km_rows) cudaFree(pts->km_rows);
        pts->km_rows = nullptr;
        return 0;  // not enough memory for the layout: the CUDA-core kernel serves the point set
    }
    float* xmax = pts->km_xnorm + n_pad;
    CU_TRY(ctx, cudaMemsetAsync(xmax, 0, 16, st));
    launch_split_rows_bf16((const float*) pts->rows, n, d, dpad, n_pad, pts->km_rows, pts->km_xnorm, nullptr, xmax, nullptr,
                           blocks_for(ctx, n_pad * 32, 256, 32), st);
    GP_TRY(check_launch(ctx, "split_rows_bf16_kernel"));
    pts->km_npad = n_pad;
    pts->km_dpad = dpad;
    GP_TRY(encode_rows_map_sw64(ctx, &pts->km_tmap, pts->km_rows, n_pad, 2 * dpad, kmtc::kTileM, "k-means points"));
    pts->km_ok = true;
    return 0;
}

// NearestCenters on the tensor cores (kmeans_tc.cu): split-precision filter + exact re-rank of ambiguous points.
int kmeans_assign_tensor(gpann_ctx* ctx, const gpann_points* pts, const float* d_centroids, size_t kc, int32_t* d_assign,
                         const unsigned char* d_alive, cudaStream_t st) {
    const int d = (int) pts->d, dpad = pts->km_dpad, nkb = dpad / 32;
    // shared memory: NB centroids resident (4*dpad bytes each) + a ring of point tiles (128 rows x 4*dpad bytes)
    int NB = 256, stages = 3;
    while (kmtc::smem_bytes(NB, nkb, stages) > 225 * 1024) {
        if (stages > 2) --stages;
        else if (NB > 128) { NB = 128; stages = 3; }
        else return fail(ctx, GPANN_E_UNSUPPORTED, "kmeans tensor path: d too large");
    }
    if (kc <= 128) NB = 128;
    const size_t kc_pad = ((kc + NB - 1) / NB) * NB;
    const int CS = (int) (kc_pad / NB);
    if (CS > ctx->num_sms) return fail(ctx, GPANN_E_UNSUPPORTED, "kmeans tensor path: too many centroids");
    const size_t n_pad = pts->km_npad;
    GP_TRY(reserve(ctx, ctx->km_cent, kc_pad * (size_t) dpad * 4));
    GP_TRY(reserve(ctx, ctx->km_hn, kc_pad * 4));
    GP_TRY(reserve(ctx, ctx->km_scal, 256));
    GP_TRY(reserve(ctx, ctx->km_partial, (size_t) CS * n_pad * sizeof(uint4)));
    GP_TRY(reserve(ctx, ctx->km_amb_list, pts->n * 4));
    GP_TRY(reserve(ctx, ctx->km_amb_info, pts->n * 8));
    float* scal = (float*) ctx->km_scal.p;  // [0] cmax, [1] hnmax, [2] ambiguous count, [8..] Consts
    kmtc::Consts* consts = (kmtc::Consts*) (scal + 8);
    CU_TRY(ctx, cudaMemsetAsync(scal, 0, 32, st));
    launch_split_rows_bf16(d_centroids, kc, d, dpad, kc_pad, (unsigned short*) ctx->km_cent.p, nullptr, (float*) ctx->km_hn.p, scal,
                           scal + 1, blocks_for(ctx, kc_pad * 32, 256, 4), st);
    GP_TRY(check_launch(ctx, "split_rows_bf16_kernel(centroids)"));
    launch_kmeans_centroid_consts(pts->km_xnorm + n_pad, scal, scal + 1, (float*) ctx->km_hn.p, (int) kc, (int) kc_pad, d_alive, consts, st);
    GP_TRY(check_launch(ctx, "kmeans_centroid_consts_kernel"));
    CUtensorMap tmap_b;
    GP_TRY(encode_rows_map_sw64(ctx, &tmap_b, ctx->km_cent.p, kc_pad, 2 * dpad, NB, "k-means centroids"));
    kmtc::Args a;
    std::memset(&a, 0, sizeof(a));
    a.NB = NB;
    a.CS = CS;
    a.nkb = nkb;
    a.nks = (d + 15) / 16;
    a.stages = stages;
    a.ntiles = (uint32_t) (n_pad / kmtc::kTileM);
    a.hnp = (const float*) ctx->km_hn.p;
    a.partial = (uint4*) ctx->km_partial.p;
    a.idesc = kmtc::make_idesc_bf16(NB);
    a.mode = ctx->km_mode;
    a.pack_mask = 0xFFFFFFE0u;
    const unsigned groups = (unsigned) std::max<size_t>(1, std::min<size_t>((size_t) (ctx->num_sms / CS), a.ntiles));
    if (ctx->timing) cudaEventRecord(ctx->ev0, st);
    launch_kmeans_tc(pts->km_tmap, tmap_b, a, groups * CS, kmtc::smem_bytes(NB, nkb, stages), st);
    if (ctx->timing) {
        cudaEventRecord(ctx->ev1, st);
        ctx->timed = true;
    }
    GP_TRY(check_launch(ctx, "kmeans_tc_kernel"));
    KmFinalizeParams fp;
    fp.partial = a.partial;
    fp.n = pts->n;
    fp.n_pad = n_pad;
    fp.CS = CS;
    fp.NB = NB;
    fp.kc = (int) kc;
    fp.d = d;
    fp.rows = (const float*) pts->rows;
    fp.centroids = d_centroids;
    fp.xnorm = pts->km_xnorm;
    fp.consts = consts;
    fp.out_assign = d_assign;
    fp.out_mindist = nullptr;
    fp.amb_list = (uint32_t*) ctx->km_amb_list.p;
    fp.amb_count = (uint32_t*) (scal + 2);
    fp.amb_info = (uint2*) ctx->km_amb_info.p;
    fp.alive = d_alive;
    launch_kmeans_finalize(fp, blocks_for(ctx, pts->n, 256, 8), st);
    GP_TRY(check_launch(ctx, "kmeans_finalize_kernel"));
    launch_kmeans_resolve(fp, ctx->num_sms * 8, st);
    GP_TRY(check_launch(ctx, "kmeans_resolve_kernel"));
    ctx->last_algo = "kmeans_tensor_bf16x3";
    return 0;
}

// ---- k-means helpers -----------------------------------------------------------------------------------------
RowsView rows_view(const gpann_points* pts) {
    RowsView v;
    v.rows = (const char*) pts->rows;
    v.pitch = pts->pitch;
    v.storage = pts->storage;
    v.d = (int) pts->d;
    v.n = pts->n;
    return v;
}

// Can the tcgen05 assignment kernel serve this shape?  (NB centroids resident per CTA, CS = kc_pad / NB splits <= SMs.)
bool kmeans_tensor_fits(const gpann_ctx* ctx, int dpad, size_t kc) {
    const int nkb = dpad / 32;
    int NB = 256, stages = 3;
    while (kmtc::smem_bytes(NB, nkb, stages) > 225 * 1024) {
        if (stages > 2) --stages;
        else if (NB > 128) { NB = 128; stages = 3; }
        else return false;
    }
    if (kc <= 128) NB = 128;
    return (kc + NB - 1) / NB <= (size_t) ctx->num_sms;
}

int kmeans_assign_device(gpann_ctx* ctx, const gpann_points* pts, const float* d_centroids, size_t kc, int metric,
                         int32_t* d_assign, float* d_mindist, const unsigned char* d_alive, cudaStream_t st) {
    // tensor path: float rows, L2, d <= 128, no per-point distance requested; AUTO takes it when the contraction is big
    // enough to pay for the layout (n*kc >= 2^24 and at least 32 centroids) and the shape fits the kernel -- otherwise,
    // and when the layout could not be built (memory), the CUDA-core kernel serves the call
    const bool tensor_shape = pts->storage == GPANN_STORAGE_F32 && metric == GPANN_METRIC_L2 && pts->d <= 128 && !d_mindist &&
                              pts->n > 0;
    bool want_tensor = tensor_shape && (ctx->algo == GPANN_ALGO_TENSOR ||
                                        (ctx->algo == GPANN_ALGO_AUTO && kc >= 32 && pts->n * kc >= ((size_t) 1 << 24)));
    if (want_tensor && ctx->algo == GPANN_ALGO_AUTO && !kmeans_tensor_fits(ctx, (((int) pts->d + 31) / 32) * 32, kc)) want_tensor = false;
    if (want_tensor && !pts->km_tried) GP_TRY(build_kmeans_layout(ctx, const_cast<gpann_points*>(pts), st));
    if (want_tensor && pts->km_ok) return kmeans_assign_tensor(ctx, pts, d_centroids, kc, d_assign, d_alive, st);
    if (ctx->algo == GPANN_ALGO_TENSOR && tensor_shape && !pts->km_ok)
        return fail(ctx, GPANN_E_OOM, "kmeans tensor path: the split-precision layout could not be built");
    ctx->last_algo = "kmeans_cuda_core";
    AssignParams p;
    p.pts = rows_view(pts);
    p.centroids = d_centroids;
    p.kc = (int) kc;
    p.metric = metric;
    const int elem = pts->storage == GPANN_STORAGE_F32 ? 4 : 1;
    const int row_bytes = (int) pts->d * elem;
    p.spitch = odd16_pitch(row_bytes);
    p.qpitch = ((int) pts->d + 3) & ~3;
    p.vec_ok = ((uintptr_t) pts->rows % 16 == 0) && (pts->pitch % 16 == 0) && (row_bytes % 16 == 0);
    p.tile_rows = stream_tile_rows(p.spitch);
    const size_t tile_bytes = (size_t) p.tile_rows * p.spitch;
    const size_t per_c = (size_t) p.qpitch * 4;
    if (tile_bytes + per_c > 200 * 1024) return fail(ctx, GPANN_E_UNSUPPORTED, "kmeans: d too large");
    const size_t budget = std::max<size_t>(tile_bytes + per_c, 100 * 1024);
    p.cchunk = (int) std::max<size_t>(1, std::min<size_t>(kc, (budget - tile_bytes) / per_c));
    p.out_assign = d_assign;
    p.out_mindist = d_mindist;
    p.alive = d_alive;
    const size_t smem = tile_bytes + (size_t) p.cchunk * per_c;
    const size_t ntiles = (pts->n + p.tile_rows - 1) / p.tile_rows;
    const int occ = (int) std::max<size_t>(1, std::min<size_t>(8, (220 * 1024) / (smem + 1024)));
    const unsigned grid = (unsigned) std::max<size_t>(1, std::min<size_t>(ntiles, (size_t) ctx->num_sms * occ));
    launch_kmeans_assign(p, grid, smem, st);
    return check_launch(ctx, "kmeans_assign_kernel");
}

// sums [kc,d] (optionally normalised) and counts [kc] of the points assigned to each cluster, fixed order.
// MIPS build: d_row_norm [n] = |x_i| scales every row by 1/|x_i| and d_norm_sums [kc] receives the per-cluster sums of |x_i|^2.
int cluster_sums_device(gpann_ctx* ctx, const gpann_points* pts, const int32_t* d_assign, size_t kc, int normalize,
                        float* d_out, uint64_t* d_counts, cudaStream_t st, const float* d_row_norm, float* d_norm_sums) {
    const size_t n = pts->n;
    const int d = (int) pts->d;
    if (n > 0x7FFFFFFFull) return fail(ctx, GPANN_E_UNSUPPORTED, "kmeans: n must be below 2^31 per device");
    CU_TRY(ctx, cudaMemsetAsync(d_counts, 0, kc * 8, st));
    GP_TRY(reserve(ctx, ctx->sort_a, n * 4));
    GP_TRY(reserve(ctx, ctx->sort_b, n * 4));
    GP_TRY(reserve(ctx, ctx->sort_c, n * 4));
    GP_TRY(reserve(ctx, ctx->sort_d, n * 4));
    uint32_t* keys_in = (uint32_t*) ctx->sort_a.p;
    uint32_t* keys_out = (uint32_t*) ctx->sort_b.p;
    uint32_t* idx_in = (uint32_t*) ctx->sort_c.p;
    uint32_t* idx_out = (uint32_t*) ctx->sort_d.p;
    launch_histogram_keys(d_assign, n, (int) kc, (unsigned long long*) d_counts, keys_in, idx_in, blocks_for(ctx, n), st);
    GP_TRY(check_launch(ctx, "histogram_keys_kernel"));
    const int bits = std::max(1, log2_ceil(kc + 1));
    size_t tmp_bytes = 0;
    sort_pairs_u32(nullptr, tmp_bytes, keys_in, keys_out, idx_in, idx_out, n, bits, st);
    GP_TRY(reserve(ctx, ctx->sort_tmp, tmp_bytes));
    if (sort_pairs_u32(ctx->sort_tmp.p, tmp_bytes, keys_in, keys_out, idx_in, idx_out, n, bits, st) != cudaSuccess)
        return fail(ctx, GPANN_E_CUDA, "kmeans: radix sort failed");
    ctx->launches += 3;
    GP_TRY(reserve(ctx, ctx->misc, (kc + 1) * 4 + (kc + 1) * 8 + 64));
    uint64_t* member_begin = (uint64_t*) ctx->misc.p;
    uint32_t* seg_begin = (uint32_t*) (member_begin + kc + 1);
    launch_cluster_segments(d_counts, (int) kc, seg_begin, member_begin, st);
    GP_TRY(check_launch(ctx, "cluster_segments_kernel"));
    const size_t max_segs = n / kSegRows + kc + 1;
    GP_TRY(reserve(ctx, ctx->segsums, max_segs * d * 4 + max_segs * 4));
    ReduceParams rp;
    rp.pts = rows_view(pts);
    rp.members = idx_out;
    rp.seg_begin = seg_begin;
    rp.member_begin = member_begin;
    rp.kc = (int) kc;
    rp.seg_sums = (float*) ctx->segsums.p;
    rp.row_norm = d_row_norm;
    rp.seg_norm_sums = d_row_norm && d_norm_sums ? rp.seg_sums + max_segs * d : nullptr;
    launch_cluster_reduce_segments(rp, (unsigned) max_segs, st);
    GP_TRY(check_launch(ctx, "cluster_reduce_segments_kernel", rp.seg_norm_sums ? 2 : 1));
    launch_cluster_finalize(rp.seg_sums, seg_begin, d_counts, (int) kc, d, normalize, d_out, rp.seg_norm_sums, d_norm_sums, st);
    return check_launch(ctx, "cluster_finalize_kernel");
}

}  // namespace gpimpl

// ---- single-device implementations of the C ABI (abi.cu dispatches: plain ctx -> sd_*, group ctx -> multi.cu) ----------
namespace gpimpl {

int sd_init(gpann_ctx** out, int device) {
    if (!out) return GPANN_E_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) return GPANN_E_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return GPANN_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return GPANN_E_CUDA;
    if (prop.major != 10) {
        std::fprintf(stderr, "gpann_init: device %d is sm_%d%d; libgpann_cuda is built for sm_100a only\n", device, prop.major,
                     prop.minor);
        return GPANN_E_UNSUPPORTED;
    }
    gpann_ctx* ctx = new (std::nothrow) gpann_ctx();
    if (!ctx) return GPANN_E_OOM;
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_order, cudaEventDisableTiming) != cudaSuccess) {
        delete ctx;
        return GPANN_E_CUDA;
    }
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
        ctx->encode_tiled = (EncodeTiledFn) fn;
    // opt every large-shared-memory kernel in for this device (function attributes are per device)
    if (stream_set_attributes() != cudaSuccess || seg_set_attributes() != cudaSuccess || kmeans_set_attributes() != cudaSuccess || tc_set_attributes() != cudaSuccess ||
        kmtc_set_attributes() != cudaSuccess) {
        std::fprintf(stderr, "gpann_init: cudaFuncSetAttribute failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return GPANN_E_CUDA;
    }
    const char* mode = std::getenv("GPANN_TC_MODE");
    ctx->tc_mode = mode ? std::atoi(mode) : 0;
    const char* kmm = std::getenv("GPANN_KM_MODE");
    ctx->km_mode = kmm ? std::atoi(kmm) : 0;
    const char* warm = std::getenv("GPANN_TC_WARM");
    ctx->tc_warm = warm ? std::atoi(warm) : 0;
    const char* pre = std::getenv("GPANN_TC_PREPASS");
    ctx->tc_prepass = pre ? std::atoi(pre) : 1;
    *out = ctx;
    return 0;
}

void sd_destroy(gpann_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();  // asynchronous entry points may still be running on caller streams
    comm_destroy(ctx);
    DevBuf* bufs[] = { &ctx->pace, &ctx->flog, &ctx->qstat, &ctx->partial, &ctx->cand, &ctx->qrows, &ctx->qnorm, &ctx->qf32, &ctx->bound, &ctx->flags, &ctx->dist_out,
                       &ctx->id_out, &ctx->cnt_out, &ctx->qdev, &ctx->misc, &ctx->sort_tmp, &ctx->sort_a, &ctx->sort_b,
                       &ctx->sort_c, &ctx->sort_d, &ctx->segsums, &ctx->cent, &ctx->sums, &ctx->counts, &ctx->assign,
                       &ctx->km_cent, &ctx->km_hn, &ctx->km_scal, &ctx->km_partial, &ctx->km_amb_list, &ctx->km_amb_info,
                       &ctx->gather_send, &ctx->gather_recv, &ctx->km_run_cent, &ctx->km_run_sums, &ctx->km_run_counts,
                       &ctx->km_run_alive, &ctx->km_run_remap, &ctx->km_run_norms };
    for (DevBuf* b : bufs) release(*b);
    DevBuf* sbufs[] = { &ctx->st_raw[0], &ctx->st_raw[1], &ctx->st_rows, &ctx->st_norms, &ctx->st_ids, &ctx->st_a, &ctx->st_b,
                        &ctx->st_c, &ctx->st_d, &ctx->st_tmp, &ctx->st_narrow };
    for (DevBuf* b : sbufs) release(*b);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (int i = 0; i < 2; ++i) {
        if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]);
        if (ctx->ev_consumed[i]) cudaEventDestroy(ctx->ev_consumed[i]);
    }
    cudaStreamDestroy(ctx->stream);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    if (ctx->ev_order) cudaEventDestroy(ctx->ev_order);
    delete ctx;
}

const char* sd_last_error(const gpann_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int sd_device(const gpann_ctx* ctx) { return ctx ? ctx->device : -1; }
int sd_set_algo(gpann_ctx* ctx, int algo) {
    if (!ctx || algo < GPANN_ALGO_AUTO || algo > GPANN_ALGO_TENSOR) return GPANN_E_ARG;
    ctx->algo = algo;
    return 0;
}
uint64_t sd_launch_count(const gpann_ctx* ctx) { return ctx ? ctx->launches : 0; }
int sd_set_timing(gpann_ctx* ctx, int on) {
    if (!ctx) return GPANN_E_ARG;
    ctx->timing = on ? 1 : 0;
    ctx->timed = false;
    return 0;
}
float sd_last_kernel_ms(gpann_ctx* ctx) {
    if (!ctx || !ctx->timed) return -1.f;
    float ms = -1.f;
    if (cudaEventSynchronize(ctx->ev1) != cudaSuccess || cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) != cudaSuccess) return -1.f;
    return ms;
}
const char* sd_last_algo(const gpann_ctx* ctx) { return ctx ? ctx->last_algo : "none"; }
uint64_t sd_last_ambiguous(gpann_ctx* ctx) {
    if (!ctx || !ctx->km_scal.p) return 0;
    uint32_t v = 0;
    cudaSetDevice(ctx->device);
    if (cudaDeviceSynchronize() != cudaSuccess) return 0;
    if (cudaMemcpy(&v, (const float*) ctx->km_scal.p + 2, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    return v;
}

static int finish_points(gpann_ctx* ctx, gpann_points* pts, cudaStream_t st, gpann_points** out) {
    if (pts->storage != GPANN_STORAGE_F32) {
        int r = build_tensor_bytes(ctx, pts, st);
        if (r != 0) {
            sd_points_free(pts);
            return r;
        }
    }
    *out = pts;
    return 0;
}

int sd_points_upload(gpann_ctx* ctx, const float* host_rows, size_t n, size_t d, int storage, uint32_t id_base,
                        gpann_points** out) {
    if (!ctx) return GPANN_E_ARG;
    if (!out || (!host_rows && n) || d == 0 || d > 4096) return fail(ctx, GPANN_E_ARG, "points_upload: bad argument");
    if (storage < GPANN_STORAGE_AUTO || storage > GPANN_STORAGE_S8) return fail(ctx, GPANN_E_ARG, "points_upload: bad storage");
    if ((uint64_t) id_base + n > 0xFFFFFFFFull) return fail(ctx, GPANN_E_ARG, "points_upload: ids exceed 32 bits (recall.h:30)");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    *out = nullptr;
    cudaStream_t st = ctx->stream;
    OrderScope order(ctx, st);
    ScopedDev f32;  // freed on every error path; ownership moves to the handle for F32 storage
    CU_TRY(ctx, f32.alloc(n * d * 4));
    if (n) CU_TRY(ctx, cudaMemcpyAsync(f32.p, host_rows, n * d * 4, cudaMemcpyHostToDevice, st));
    if (storage == GPANN_STORAGE_AUTO && n) {
        int flags_h[2] = { 1, 1 };
        GP_TRY(reserve(ctx, ctx->flags, 64));
        CU_TRY(ctx, cudaMemcpyAsync(ctx->flags.p, flags_h, 8, cudaMemcpyHostToDevice, st));
        launch_detect_bytes(f32.as<float>(), n * d, (int*) ctx->flags.p, blocks_for(ctx, n * d), st);
        GP_TRY(check_launch(ctx, "detect_bytes_kernel"));
        CU_TRY(ctx, cudaMemcpyAsync(flags_h, ctx->flags.p, 8, cudaMemcpyDeviceToHost, st));
        CU_TRY(ctx, cudaStreamSynchronize(st));
        storage = flags_h[0] ? GPANN_STORAGE_U8 : (flags_h[1] ? GPANN_STORAGE_S8 : GPANN_STORAGE_F32);
    } else if (storage == GPANN_STORAGE_AUTO) {
        storage = GPANN_STORAGE_F32;
    }
    gpann_points* pts = new (std::nothrow) gpann_points();
    if (!pts) return fail(ctx, GPANN_E_OOM, "points_upload: host allocation failed");
    pts->ctx = ctx;
    pts->n = n;
    pts->d = d;
    pts->storage = storage;
    pts->id_base = id_base;
    if (storage == GPANN_STORAGE_F32) {
        pts->rows = f32.release();
        pts->pitch = d * 4;
    } else {
        unsigned char* bytes = nullptr;
        if (cudaMalloc(&bytes, std::max<size_t>(n * d, 256)) != cudaSuccess) {
            cudaGetLastError();
            delete pts;
            return fail(ctx, GPANN_E_OOM, "points_upload: out of device memory");
        }
        pts->rows = bytes;
        pts->pitch = d;
        if (n) {
            launch_narrow_rows(f32.as<float>(), n, (int) d, (int) d, bytes, blocks_for(ctx, n * d), st);
            ctx->launches++;
        }
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) {
        sd_points_free(pts);
        return fail(ctx, GPANN_E_CUDA, std::string("points_upload: ") + cudaGetErrorString(cudaGetLastError()));
    }
    return finish_points(ctx, pts, st, out);
}

int sd_points_upload_bytes(gpann_ctx* ctx, const void* host_rows, size_t n, size_t d, int is_signed, uint32_t id_base,
                              gpann_points** out) {
    if (!ctx) return GPANN_E_ARG;
    if (!out || (!host_rows && n) || d == 0 || d > 4096) return fail(ctx, GPANN_E_ARG, "points_upload_bytes: bad argument");
    if ((uint64_t) id_base + n > 0xFFFFFFFFull) return fail(ctx, GPANN_E_ARG, "points_upload_bytes: ids exceed 32 bits");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    *out = nullptr;
    unsigned char* bytes = nullptr;
    CU_TRY(ctx, cudaMalloc(&bytes, std::max<size_t>(n * d, 256)));
    if (n) {
        cudaError_t e = cudaMemcpyAsync(bytes, host_rows, n * d, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            cudaFree(bytes);
            return fail(ctx, GPANN_E_CUDA, std::string("points_upload_bytes H2D: ") + cudaGetErrorString(e));
        }
    }
    gpann_points* pts = new (std::nothrow) gpann_points();
    if (!pts) {
        cudaFree(bytes);
        return fail(ctx, GPANN_E_OOM, "points_upload_bytes: host allocation failed");
    }
    pts->ctx = ctx;
    pts->n = n;
    pts->d = d;
    pts->storage = is_signed ? GPANN_STORAGE_S8 : GPANN_STORAGE_U8;
    pts->id_base = id_base;
    pts->rows = bytes;
    pts->pitch = d;
    return finish_points(ctx, pts, ctx->stream, out);
}

int sd_points_from_device(gpann_ctx* ctx, const void* dev_rows, size_t n, size_t d, int storage, uint32_t id_base, void* stream,
                             gpann_points** out) {
    if (!ctx) return GPANN_E_ARG;
    if (!out || (!dev_rows && n) || d == 0 || d > 4096) return fail(ctx, GPANN_E_ARG, "points_from_device: bad argument");
    if (storage < GPANN_STORAGE_F32 || storage > GPANN_STORAGE_S8) return fail(ctx, GPANN_E_ARG, "points_from_device: bad storage");
    if ((uint64_t) id_base + n > 0xFFFFFFFFull) return fail(ctx, GPANN_E_ARG, "points_from_device: ids exceed 32 bits");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    *out = nullptr;
    gpann_points* pts = new (std::nothrow) gpann_points();
    if (!pts) return fail(ctx, GPANN_E_OOM, "points_from_device: host allocation failed");
    pts->ctx = ctx;
    pts->n = n;
    pts->d = d;
    pts->storage = storage;
    pts->id_base = id_base;
    pts->rows = const_cast<void*>(dev_rows);  // borrowed: the caller keeps the rows alive
    pts->owns_rows = false;
    pts->pitch = storage == GPANN_STORAGE_F32 ? d * 4 : d;
    return finish_points(ctx, pts, (cudaStream_t) stream, out);
}

void sd_points_free(gpann_points* pts) {
    if (!pts) return;
    if (pts->ctx) cudaSetDevice(pts->ctx->device);
    if (pts->owns_rows && pts->rows) cudaFree(pts->rows);
    if (pts->t_rows) cudaFree(pts->t_rows);
    if (pts->t_norms) cudaFree(pts->t_norms);
    if (pts->t_ids) cudaFree(pts->t_ids);
    if (pts->km_rows) cudaFree(pts->km_rows);
    if (pts->km_xnorm) cudaFree(pts->km_xnorm);
    delete pts;
}
size_t sd_points_n(const gpann_points* pts) { return pts ? pts->n : 0; }
size_t sd_points_d(const gpann_points* pts) { return pts ? pts->d : 0; }
int sd_points_storage(const gpann_points* pts) { return pts ? pts->storage : 0; }

int sd_topk(gpann_ctx* ctx, const gpann_points* pts, const float* queries, size_t nq, int k, int metric, float* out_dist,
               uint32_t* out_id, uint32_t* out_count) {
    if (!ctx) return GPANN_E_ARG;
    if (!pts) return fail(ctx, GPANN_E_ARG, "topk: null point set");
    return topk_host(ctx, pts, 0, pts->n, queries, nq, k, metric, out_dist, out_id, out_count, nullptr);
}

int sd_kth_distance(gpann_ctx* ctx, const gpann_points* pts, const float* queries, size_t nq, int k, int metric,
                       float* out_kth) {
    if (!ctx) return GPANN_E_ARG;
    if (!pts || !out_kth) return fail(ctx, GPANN_E_ARG, "kth_distance: null argument");
    return topk_host(ctx, pts, 0, pts->n, queries, nq, k, metric, nullptr, nullptr, nullptr, out_kth);
}

int sd_topk_range(gpann_ctx* ctx, const gpann_points* pts, size_t row_begin, size_t row_end, const float* queries, size_t nq,
                     int k, int metric, float* out_dist, uint32_t* out_id, uint32_t* out_count) {
    if (!ctx) return GPANN_E_ARG;
    if (!pts) return fail(ctx, GPANN_E_ARG, "topk_range: null point set");
    return topk_host(ctx, pts, row_begin, row_end, queries, nq, k, metric, out_dist, out_id, out_count, nullptr);
}

int sd_topk_stream_bytes(gpann_ctx* ctx, const void* host_rows, size_t n, size_t d, int is_signed, uint32_t id_base,
                            const float* queries, size_t nq, int k, size_t chunk_rows, float* out_dist, uint32_t* out_id,
                            uint32_t* out_count) {
    if (!ctx) return GPANN_E_ARG;
    return topk_stream_host_api(ctx, host_rows, is_signed ? kHostS8 : kHostU8, 0, n, d, id_base, queries, nq, k, chunk_rows, out_dist,
                                out_id, out_count, nullptr);
}

int sd_topk_stream(gpann_ctx* ctx, const float* host_rows, size_t n, size_t d, int storage, uint32_t id_base, const float* queries,
                      size_t nq, int k, size_t chunk_rows, float* out_dist, uint32_t* out_id, uint32_t* out_count, float* out_kth) {
    if (!ctx) return GPANN_E_ARG;
    if (storage < GPANN_STORAGE_AUTO || storage > GPANN_STORAGE_S8) return fail(ctx, GPANN_E_ARG, "topk_stream: bad storage");
    return topk_stream_host_api(ctx, host_rows, kHostF32, storage, n, d, id_base, queries, nq, k, chunk_rows, out_dist, out_id,
                                out_count, out_kth);
}

int sd_topk_stream_bytes_device(gpann_ctx* ctx, const void* host_rows, size_t n, size_t d, int is_signed, uint32_t id_base,
                                const float* dev_queries, size_t nq, int k, size_t chunk_rows, uint64_t* dev_out_keys, void* stream) {
    if (!ctx) return GPANN_E_ARG;
    if ((!host_rows && n) || !dev_queries || !dev_out_keys) return fail(ctx, GPANN_E_ARG, "topk_stream_bytes_device: null argument");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    OrderScope order(ctx, (cudaStream_t) stream);
    TopkOut o;
    o.keys = dev_out_keys;
    return topk_stream_host(ctx, host_rows, is_signed ? kHostS8 : kHostU8, 0, n, d, id_base, dev_queries, nq, k, chunk_rows, o,
                            (cudaStream_t) stream);
}

int sd_topk_device(gpann_ctx* ctx, const gpann_points* pts, const float* dev_queries, size_t nq, int k, int metric,
                   uint64_t* dev_out_keys, void* stream) {
    if (!ctx) return GPANN_E_ARG;
    if (!pts || !dev_queries || !dev_out_keys) return fail(ctx, GPANN_E_ARG, "topk_device: null argument");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    OrderScope order(ctx, (cudaStream_t) stream);
    TopkOut o;
    o.keys = dev_out_keys;
    return topk_dispatch(ctx, pts, 0, pts->n, dev_queries, nq, k, metric, o, (cudaStream_t) stream, nullptr);
}

int sd_merge_keys_device(gpann_ctx* ctx, const uint64_t* dev_keys, size_t parts, size_t nq, int k, float* dev_out_dist,
                         uint32_t* dev_out_id, uint64_t* dev_out_keys, void* stream) {
    if (!ctx) return GPANN_E_ARG;
    if (!dev_keys || parts == 0 || k < 1 || k > 512) return fail(ctx, GPANN_E_ARG, "merge_keys: bad argument");
    if (nq == 0) return 0;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    TopkOut o;
    o.dist = dev_out_dist;
    o.id = dev_out_id;
    o.keys = dev_out_keys;
    return launch_merge(ctx, dev_keys, parts, nq, nq, k, o, 0, (cudaStream_t) stream, nullptr, 0);  // no ctx scratch involved
}

// The per-rank step of the sharded top-k (north_star: "each GPU computes a local top-k over its shard, and the final
// merge is an NCCL all-gather of (distance, id) candidates"): local scan of rows [row_begin,row_end) with the per-query
// bounds all-reduced (min) across the ranks between the sample passes and the main pass, one all-gather of the [nq,k] key
// blocks, merge.  Every rank ends up with the same result in `out`.  Without a communicator it is the plain local scan.
// Queries are served in batches no larger than the smallest tensor-path batch, so every rank issues one bound exchange
// per batch whatever its shard's storage or size.
int topk_sharded(gpann_ctx* ctx, const gpann_points* pts, size_t row_begin, size_t row_end, const float* dq, size_t nq, int k,
                 int metric, const TopkOut& out, cudaStream_t st) {
    if (ctx->comm_size <= 1) return topk_dispatch(ctx, pts, row_begin, row_end, dq, nq, k, metric, out, st, nullptr);
    if (k < 1 || k > 512) return fail(ctx, GPANN_E_ARG, "topk: k must be in [1,512]");
    if (nq == 0) return 0;
    const size_t part = nq * (size_t) k;
    GP_TRY(reserve(ctx, ctx->gather_send, part * 8));
    GP_TRY(reserve(ctx, ctx->gather_recv, part * 8 * ctx->comm_size));
    const BetweenFn between = [ctx](uint32_t* bound, size_t count, cudaStream_t s) { return comm_allreduce_min_u32(ctx, bound, count, s); };
    const size_t batch = tc_batch_max(ctx, 2);
    // Peer-shared bounds (the kernels of every rank tighten every other rank's bounds over NVLink while they run) need the
    // shared array to belong to ONE query batch between two collectives of the call: the all-gather below is what tells a
    // rank that no peer kernel of this call is still writing before the next call resets the array.
    static const bool no_peers = std::getenv("GPANN_NO_PEER_BOUNDS") != nullptr;
    ctx->peer_share = !no_peers && ctx->npeers > 0 && nq <= batch;
    int rc = 0;
    for (size_t q0 = 0; q0 < nq && rc == 0; q0 += batch) {
        TopkOut local;
        local.keys = (uint64_t*) ctx->gather_send.p + q0 * k;
        rc = topk_dispatch(ctx, pts, row_begin, row_end, dq + q0 * pts->d, std::min(batch, nq - q0), k, metric, local, st, between);
    }
    ctx->peer_share = false;
    GP_TRY(rc);
    GP_TRY(comm_allgather_u64(ctx, (const uint64_t*) ctx->gather_send.p, (uint64_t*) ctx->gather_recv.p, part, st));
    return launch_merge(ctx, (const uint64_t*) ctx->gather_recv.p, ctx->comm_size, nq, nq, k, out, 0, st, nullptr, 0);
}

int sd_topk_sharded_device(gpann_ctx* ctx, const gpann_points* pts, const float* dev_queries, size_t nq, int k, int metric,
                           float* dev_out_dist, uint32_t* dev_out_id, uint64_t* dev_out_keys, void* stream) {
    if (!ctx) return GPANN_E_ARG;
    if (!pts || !dev_queries) return fail(ctx, GPANN_E_ARG, "topk_sharded_device: null argument");
    if (pts->ctx != ctx) return fail(ctx, GPANN_E_ARG, "topk: point set does not belong to this context");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    OrderScope order(ctx, (cudaStream_t) stream);
    TopkOut o;
    o.dist = dev_out_dist;
    o.id = dev_out_id;
    o.keys = dev_out_keys;
    return topk_sharded(ctx, pts, 0, pts->n, dev_queries, nq, k, metric, o, (cudaStream_t) stream);
}

int sd_kmeans_assign(gpann_ctx* ctx, const gpann_points* pts, const float* centroids, size_t kc, int metric, int32_t* out_assign,
                     float* out_mindist) {
    if (!ctx) return GPANN_E_ARG;
    if (!pts || !centroids || !out_assign) return fail(ctx, GPANN_E_ARG, "kmeans_assign: null argument");
    if (kc < 1) return fail(ctx, GPANN_E_ARG, "KMeans #centroids < 1");  // src/kmeans.cpp:229-231
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    OrderScope order(ctx, st);
    const size_t n = pts->n, d = pts->d;
    GP_TRY(reserve(ctx, ctx->cent, kc * d * 4));
    GP_TRY(reserve(ctx, ctx->assign, std::max<size_t>(n, 1) * 4));
    if (out_mindist) GP_TRY(reserve(ctx, ctx->dist_out, std::max<size_t>(n, 1) * 4));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->cent.p, centroids, kc * d * 4, cudaMemcpyHostToDevice, st));
    if (n) {
        GP_TRY(kmeans_assign_device(ctx, pts, (const float*) ctx->cent.p, kc, metric, (int32_t*) ctx->assign.p,
                                    out_mindist ? (float*) ctx->dist_out.p : nullptr, nullptr, st));
        CU_TRY(ctx, cudaMemcpyAsync(out_assign, ctx->assign.p, n * 4, cudaMemcpyDeviceToHost, st));
        if (out_mindist) CU_TRY(ctx, cudaMemcpyAsync(out_mindist, ctx->dist_out.p, n * 4, cudaMemcpyDeviceToHost, st));
    }
    CU_TRY(ctx, cudaStreamSynchronize(st));
    return 0;
}

// assignment + this shard's per-cluster coordinate sums and counts (no division)
int kmeans_partial(gpann_ctx* ctx, const gpann_points* pts, const float* d_centroids, size_t kc, int metric, int32_t* d_assign,
                   const unsigned char* d_alive, float* d_sums, uint64_t* d_counts, cudaStream_t st, const float* d_row_norm = nullptr,
                   float* d_norm_sums = nullptr) {
    if (pts->n == 0) {
        CU_TRY(ctx, cudaMemsetAsync(d_sums, 0, kc * pts->d * 4, st));
        CU_TRY(ctx, cudaMemsetAsync(d_counts, 0, kc * 8, st));
        if (d_norm_sums) CU_TRY(ctx, cudaMemsetAsync(d_norm_sums, 0, kc * 4, st));
        return 0;
    }
    GP_TRY(kmeans_assign_device(ctx, pts, d_centroids, kc, metric, d_assign, nullptr, d_alive, st));
    return cluster_sums_device(ctx, pts, d_assign, kc, 0, d_sums, d_counts, st, d_row_norm, d_norm_sums);
}

int sd_kmeans_partial_device(gpann_ctx* ctx, const gpann_points* pts, const float* dev_centroids, size_t kc, int metric,
                             int32_t* dev_out_assign, float* dev_out_sums, uint64_t* dev_out_counts, void* stream) {
    if (!ctx) return GPANN_E_ARG;
    if (!pts || !dev_centroids || !dev_out_sums || !dev_out_counts) return fail(ctx, GPANN_E_ARG, "kmeans_partial: null argument");
    if (kc < 1) return fail(ctx, GPANN_E_ARG, "KMeans #centroids < 1");
    if (metric != GPANN_METRIC_L2) return fail(ctx, GPANN_E_UNSUPPORTED, "kmeans: only the L2 build is implemented");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t) stream;
    OrderScope order(ctx, st);
    int32_t* d_assign = dev_out_assign;
    if (!d_assign) {
        GP_TRY(reserve(ctx, ctx->assign, std::max<size_t>(pts->n, 1) * 4));
        d_assign = (int32_t*) ctx->assign.p;
    }
    return kmeans_partial(ctx, pts, dev_centroids, kc, metric, d_assign, nullptr, dev_out_sums, dev_out_counts, st);
}

int kmeans_rounds_device(gpann_ctx* ctx, const gpann_points* pts, float* d_cent, size_t kc, int rounds, int metric, int32_t* d_assign,
                         unsigned char* d_alive, uint64_t* d_counts, cudaStream_t st);

// One Lloyd round through host buffers: rows of empty clusters all-zero, nothing renumbered (the reference's host-side
// RemoveEmptyClusters, src/kmeans.cpp:29-50, stays with the caller).  L2 and MIPS builds.
int sd_kmeans_step(gpann_ctx* ctx, const gpann_points* pts, float* centroids_inout, size_t kc, int metric, int32_t* out_assign,
                   uint64_t* out_sizes) {
    if (!ctx) return GPANN_E_ARG;
    if (!pts || !centroids_inout || !out_sizes) return fail(ctx, GPANN_E_ARG, "kmeans_step: null argument");
    if (kc < 1) return fail(ctx, GPANN_E_ARG, "KMeans #centroids < 1");
    if (metric != GPANN_METRIC_L2 && metric != GPANN_METRIC_MIPS) return fail(ctx, GPANN_E_ARG, "kmeans_step: bad metric");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    OrderScope order(ctx, st);
    const size_t n = pts->n, d = pts->d;
    GP_TRY(reserve(ctx, ctx->km_run_cent, kc * d * 4));
    GP_TRY(reserve(ctx, ctx->km_run_counts, kc * 8 + 64));
    GP_TRY(reserve(ctx, ctx->km_run_alive, kc));
    GP_TRY(reserve(ctx, ctx->assign, std::max<size_t>(n, 1) * 4));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->km_run_cent.p, centroids_inout, kc * d * 4, cudaMemcpyHostToDevice, st));
    CU_TRY(ctx, cudaMemsetAsync(ctx->km_run_alive.p, 1, kc, st));
    GP_TRY(kmeans_rounds_device(ctx, pts, (float*) ctx->km_run_cent.p, kc, 1, metric, (int32_t*) ctx->assign.p,
                                (unsigned char*) ctx->km_run_alive.p, (uint64_t*) ctx->km_run_counts.p, st));
    CU_TRY(ctx, cudaMemcpyAsync(centroids_inout, ctx->km_run_cent.p, kc * d * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaMemcpyAsync(out_sizes, ctx->km_run_counts.p, kc * 8, cudaMemcpyDeviceToHost, st));
    if (out_assign && n) CU_TRY(ctx, cudaMemcpyAsync(out_assign, ctx->assign.p, n * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    return 0;
}

// ---- KMeans (src/kmeans.cpp:228-244) with the centroids resident on the device for all rounds ----------------------------
// Per round: nearest centroid (removed clusters skipped), deterministic per-cluster sums, an all-reduce(sum) of sums and
// counts when the ctx is a rank of a communicator, division, removal of clusters that lost every member
// (in place: alive[c] = 0).  d_cent [kc,d] and d_alive [kc] are updated in place; d_counts [kc] holds the last round's
// (global) cluster sizes; d_assign [n_local] the last round's assignment in ORIGINAL cluster numbers.
int kmeans_rounds_device(gpann_ctx* ctx, const gpann_points* pts, float* d_cent, size_t kc, int rounds, int metric, int32_t* d_assign,
                         unsigned char* d_alive, uint64_t* d_counts, cudaStream_t st) {
    const size_t d = pts->d;
    const bool mips = metric == GPANN_METRIC_MIPS;
    GP_TRY(reserve(ctx, ctx->km_run_sums, kc * d * 4 + kc * 4));
    float* d_sums = (float*) ctx->km_run_sums.p;
    float* d_norm_sums = mips ? d_sums + kc * d : nullptr;  // behind the sums: one all-reduce covers both
    float* d_row_norm = nullptr;
    if (mips && pts->n && rounds > 0) {  // vector_sqrt_norms, once per KMeans call (src/kmeans.cpp:233-236)
        GP_TRY(reserve(ctx, ctx->km_run_norms, pts->n * 4));
        d_row_norm = (float*) ctx->km_run_norms.p;
        launch_row_norms(rows_view(pts), d_row_norm, blocks_for(ctx, pts->n), st);
        GP_TRY(check_launch(ctx, "row_norms_kernel"));
    }
    for (int r = 0; r < rounds; ++r) {
        GP_TRY(kmeans_partial(ctx, pts, d_cent, kc, metric, d_assign, d_alive, d_sums, d_counts, st, d_row_norm, d_norm_sums));
        if (ctx->comm_size > 1) GP_TRY(comm_allreduce_kmeans(ctx, d_sums, kc * d + (mips ? kc : 0), d_counts, kc, st));
        if (mips) launch_kmeans_update_ip(d_cent, d_sums, d_counts, d_norm_sums, d_alive, (int) kc, (int) d, st);
        else launch_kmeans_update(d_cent, d_sums, d_counts, d_alive, (int) kc, (int) d, st);
        GP_TRY(check_launch(ctx, "kmeans_update_kernel"));
    }
    return 0;
}

// Final renumbering (RemoveEmptyClusters, src/kmeans.cpp:29-50, applied once): survivors compacted to the front of d_cent /
// d_counts, assignments renumbered, the number of survivors in d_kc_out.
int kmeans_compact_device(gpann_ctx* ctx, float* d_cent, size_t kc, size_t d, int32_t* d_assign, size_t n, const unsigned char* d_alive,
                          uint64_t* d_counts, int* d_kc_out, cudaStream_t st) {
    GP_TRY(reserve(ctx, ctx->km_run_remap, kc * 4));
    GP_TRY(reserve(ctx, ctx->km_run_sums, kc * d * 4));
    launch_kmeans_compact(d_cent, d_counts, d_alive, (int) kc, (int) d, (int32_t*) ctx->km_run_remap.p, d_kc_out, (float*) ctx->km_run_sums.p, st);
    GP_TRY(check_launch(ctx, "kmeans_compact_kernel"));
    if (n && d_assign) {
        launch_remap_assign(d_assign, n, (const int32_t*) ctx->km_run_remap.p, (int) kc, blocks_for(ctx, n), st);
        GP_TRY(check_launch(ctx, "remap_assign_kernel"));
    }
    return 0;
}

// Queue the whole KMeans call on ctx->stream: H2D of the centroids, `rounds` Lloyd rounds, compaction.  Results stay on the
// device (ctx->km_run_cent / km_run_counts / assign, survivors in km_run_remap-adjacent scalar); kmeans_run_fetch copies
// them out.  Split in two so that a group ctx (multi.cu) can queue every rank before any rank waits.
int kmeans_run_enqueue(gpann_ctx* ctx, const gpann_points* pts, const float* centroids, size_t kc, int rounds, int metric) {
    if (!pts || !centroids) return fail(ctx, GPANN_E_ARG, "kmeans_run: null argument");
    if (pts->ctx != ctx) return fail(ctx, GPANN_E_ARG, "kmeans_run: point set does not belong to this context");
    if (kc < 1) return fail(ctx, GPANN_E_ARG, "KMeans #centroids < 1");  // src/kmeans.cpp:229-231
    if (rounds < 0) return fail(ctx, GPANN_E_ARG, "kmeans_run: rounds < 0");
    if (metric != GPANN_METRIC_L2 && metric != GPANN_METRIC_MIPS) return fail(ctx, GPANN_E_ARG, "kmeans_run: bad metric");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t n = pts->n, d = pts->d;
    GP_TRY(reserve(ctx, ctx->km_run_cent, kc * d * 4));
    GP_TRY(reserve(ctx, ctx->km_run_counts, kc * 8 + 64));
    GP_TRY(reserve(ctx, ctx->km_run_alive, kc));
    GP_TRY(reserve(ctx, ctx->assign, std::max<size_t>(n, 1) * 4));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->km_run_cent.p, centroids, kc * d * 4, cudaMemcpyHostToDevice, st));
    CU_TRY(ctx, cudaMemsetAsync(ctx->km_run_alive.p, 1, kc, st));
    CU_TRY(ctx, cudaMemsetAsync(ctx->km_run_counts.p, 0, kc * 8 + 64, st));
    if (n) CU_TRY(ctx, cudaMemsetAsync(ctx->assign.p, 0xFF, n * 4, st));  // -1: the reference's initial closest_center
    uint64_t* d_counts = (uint64_t*) ctx->km_run_counts.p;
    int* d_kc_out = (int*) (d_counts + kc);
    GP_TRY(kmeans_rounds_device(ctx, pts, (float*) ctx->km_run_cent.p, kc, rounds, metric, (int32_t*) ctx->assign.p,
                                (unsigned char*) ctx->km_run_alive.p, d_counts, st));
    return kmeans_compact_device(ctx, (float*) ctx->km_run_cent.p, kc, d, (int32_t*) ctx->assign.p, n, (const unsigned char*) ctx->km_run_alive.p,
                                 d_counts, d_kc_out, st);
}

int kmeans_run_fetch(gpann_ctx* ctx, const gpann_points* pts, float* centroids_out, size_t kc, int32_t* out_assign, uint64_t* out_sizes,
                     size_t* out_kc) {
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t n = pts->n, d = pts->d;
    const uint64_t* d_counts = (const uint64_t*) ctx->km_run_counts.p;
    int kc_new = 0;
    if (centroids_out) CU_TRY(ctx, cudaMemcpyAsync(centroids_out, ctx->km_run_cent.p, kc * d * 4, cudaMemcpyDeviceToHost, st));
    if (out_sizes) CU_TRY(ctx, cudaMemcpyAsync(out_sizes, d_counts, kc * 8, cudaMemcpyDeviceToHost, st));
    if (out_assign && n) CU_TRY(ctx, cudaMemcpyAsync(out_assign, ctx->assign.p, n * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaMemcpyAsync(&kc_new, d_counts + kc, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    if (out_kc) *out_kc = (size_t) kc_new;
    return 0;
}

int sd_kmeans_run(gpann_ctx* ctx, const gpann_points* pts, float* centroids_inout, size_t kc, int rounds, int metric, int32_t* out_assign,
                  uint64_t* out_sizes, size_t* out_kc) {
    if (!ctx) return GPANN_E_ARG;
    OrderScope order(ctx, ctx->stream);
    GP_TRY(kmeans_run_enqueue(ctx, pts, centroids_inout, kc, rounds, metric));
    return kmeans_run_fetch(ctx, pts, centroids_inout, kc, out_assign, out_sizes, out_kc);
}

// Device form: centroids [kc,d], alive [kc] (caller initialises: 1 = live) and counts [kc] live in caller-owned device memory and
// are updated in place, nothing is renumbered and nothing synchronises -- bench.py and the torch.distributed driver time
// Lloyd rounds with it.  With a communicator every rank passes its shard and ends up with identical centroids.
int sd_kmeans_run_device(gpann_ctx* ctx, const gpann_points* pts, float* dev_centroids_inout, size_t kc, int rounds, int metric,
                         int32_t* dev_out_assign, unsigned char* dev_alive_inout, uint64_t* dev_out_counts, void* stream) {
    if (!ctx) return GPANN_E_ARG;
    if (!pts || !dev_centroids_inout || !dev_alive_inout || !dev_out_counts) return fail(ctx, GPANN_E_ARG, "kmeans_run_device: null argument");
    if (pts->ctx != ctx) return fail(ctx, GPANN_E_ARG, "kmeans_run: point set does not belong to this context");
    if (kc < 1) return fail(ctx, GPANN_E_ARG, "KMeans #centroids < 1");
    if (metric != GPANN_METRIC_L2 && metric != GPANN_METRIC_MIPS) return fail(ctx, GPANN_E_ARG, "kmeans_run: bad metric");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t) stream;
    OrderScope order(ctx, st);
    int32_t* d_assign = dev_out_assign;
    if (!d_assign) {
        GP_TRY(reserve(ctx, ctx->assign, std::max<size_t>(pts->n, 1) * 4));
        d_assign = (int32_t*) ctx->assign.p;
    }
    return kmeans_rounds_device(ctx, pts, dev_centroids_inout, kc, rounds, metric, d_assign, dev_alive_inout, dev_out_counts, st);
}

// sum_i pos_distance(P_i, C_assign[i]) of this ctx's rows, queued on ctx->stream; the double lands in *d_out_ptr (device).
int objective_enqueue(gpann_ctx* ctx, const gpann_points* pts, const float* centroids, size_t kc, const int32_t* assign, int metric,
                      double** d_out_ptr) {
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t n = pts->n, d = pts->d;
    GP_TRY(reserve(ctx, ctx->cent, kc * d * 4));
    GP_TRY(reserve(ctx, ctx->assign, std::max<size_t>(n, 1) * 4));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->cent.p, centroids, kc * d * 4, cudaMemcpyHostToDevice, st));
    if (n) CU_TRY(ctx, cudaMemcpyAsync(ctx->assign.p, assign, n * 4, cudaMemcpyHostToDevice, st));
    return objective_device(ctx, pts, (const float*) ctx->cent.p, kc, (const int32_t*) ctx->assign.p, metric, d_out_ptr, st);
}

// device-pointer form (centroids [kc,d], assign [n] already on the device); uses ctx->misc for the per-CTA partial sums
int objective_device(gpann_ctx* ctx, const gpann_points* pts, const float* d_centroids, size_t kc, const int32_t* d_assign, int metric,
                     double** d_out_ptr, cudaStream_t st) {
    const size_t n = pts->n, d = pts->d;
    ObjectiveParams p;
    p.pts = rows_view(pts);
    p.centroids = d_centroids;
    p.assign = d_assign;
    p.kc = (int) kc;
    p.metric = metric;
    const int elem = pts->storage == GPANN_STORAGE_F32 ? 4 : 1;
    const int row_bytes = (int) d * elem;
    p.spitch = odd16_pitch(row_bytes);
    p.vec_ok = ((uintptr_t) pts->rows % 16 == 0) && (pts->pitch % 16 == 0) && (row_bytes % 16 == 0);
    p.tile_rows = stream_tile_rows(p.spitch);
    const size_t smem = (size_t) p.tile_rows * p.spitch;
    if (smem > 200 * 1024) return fail(ctx, GPANN_E_UNSUPPORTED, "objective: d too large");
    const size_t ntiles = (n + p.tile_rows - 1) / p.tile_rows;
    const unsigned grid = (unsigned) std::max<size_t>(1, std::min<size_t>(ntiles, (size_t) ctx->num_sms * 4));
    GP_TRY(reserve(ctx, ctx->misc, (size_t) grid * 8 + 64));
    p.block_sums = (double*) ctx->misc.p;
    double* d_out = p.block_sums + grid;
    launch_objective(p, grid, smem, d_out, st);  // n == 0: every CTA contributes 0
    GP_TRY(check_launch(ctx, "objective_kernel", 2));
    *d_out_ptr = d_out;
    return 0;
}

int sd_objective(gpann_ctx* ctx, const gpann_points* pts, const float* centroids, size_t kc, const int32_t* assign, int metric,
                 double* out) {
    if (!ctx) return GPANN_E_ARG;
    if (!pts || !centroids || !out || (!assign && pts->n)) return fail(ctx, GPANN_E_ARG, "objective: null argument");
    OrderScope order(ctx, ctx->stream);
    double* d_out = nullptr;
    GP_TRY(objective_enqueue(ctx, pts, centroids, kc, assign, metric, &d_out));
    CU_TRY(ctx, cudaMemcpyAsync(out, d_out, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

}  // namespace gpimpl
