// topk_tc.cuh -- tcgen05 / TMEM / TMA tile kernel: batched byte-vector L2 distances fused with per-query top-k.
//
// Replaces the reference's `for query: for point: distance(); TopN.Add()` (src/recall.h:25-39) for uint8 /
// int8 data (BigANN-style, widened to float by the reference's reader, src/points_io.cpp:63-111).  For
// integer coordinates every partial sum of src/dist.cpp:9-16 is an exact integer below 2^24, so
//     dist(q,x) = |q|^2 + |x|^2 - 2 q.x          (all int32)
// is bit-identical to the reference's float result; q.x is the only O(n*nq*d) term and runs on the tensor
// cores as u8 x u8 -> s32 (tcgen05.mma.kind::i8).
//
// Layout in HBM (built once per point set by prep.cuh): point rows are SORTED BY NORM, padded to 128-byte rows
// and to a multiple of 256 rows; norms[] (int32) and ids[] (original id) ride along.  Because a 256-point
// tile then holds near-equal norms, the top-k test against the running k-th best distance kth_q
//     2*acc >= |q|^2 + |x|^2 - kth_q      <=     acc >= ceil((|q|^2 - kth_q + min_tile|x|^2) / 2) =: T
// needs no per-value norm: the epilogue takes a 3-input max over the accumulators straight out of TMEM
// (VIMNMX3, half an instruction per distance) and compares once per 32 values.  Values reaching T are rare
// (k / points-seen) and are re-tested exactly in a slow path that appends (distance,id) keys to a per-query
// candidate buffer; a full buffer is compacted by a warp-wide register bitonic select, which also tightens T.
//
// CTA roles (192 threads, 1 CTA/SM, persistent over its point range):
//   warp 0 lane 0 : TMA producer  -- Q tiles once, then 256x128B point tiles into a kStages ring (SWIZZLE_128B)
//   warp 1 lane 0 : MMA issuer    -- per (point tile, Q tile): ceil(d/32) x tcgen05.mma M128 N256 K32 into one of
//                                    two 256-column TMEM accumulators; tcgen05.commit -> mbarriers
//   warps 2..5    : epilogue      -- tcgen05.ld 32x32b.x32, max3 filter, rare slow path, tmem_empty arrive
// Grid: G query groups (R Q-tiles of 128 queries each, resident in shared memory) x S point splits.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace gpann {
namespace tc {

constexpr int kTileN = 256;       // points per tile  (UMMA N)
constexpr int kTileM = 128;       // queries per tile (UMMA M)
constexpr int kRowBytes = 128;    // operand row bytes = one SWIZZLE_128B row
constexpr int kStages = 3;
constexpr int kMaxR = 4;
constexpr int kThreads = 192;
constexpr int kQTileBytes = kTileM * kRowBytes;  // 16 KB
constexpr int kPTileBytes = kTileN * kRowBytes;  // 32 KB
constexpr int kTmemCols = 512;                   // 2 accumulators x 256 columns

struct Args {
    const int32_t* pnorm;    // [n_pad] ascending
    const uint32_t* pid;     // [n_pad] original ids, kInvalidId on padding rows
    const int32_t* qnorm;    // [G*R*128]
    int R, G, S;             // Q tiles per CTA, query groups, point splits
    int nks;                 // K steps of 32 bytes
    int k, cap;              // top-k and candidate buffer capacity (power of two >= k + 64)
    uint32_t tiles_total;    // n_pad / 256
    uint32_t idesc;          // tcgen05 instruction descriptor
    uint64_t* cand;          // [grid, R*128, cap]
    uint64_t* partial;       // [S, G*R*128, k]
    int mode;                // 0 = full; 1 = MMA + TMEM loads only (roofline probe, results invalid)
};

// ---- PTX wrappers ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Bounded wait: a protocol bug traps (launch failure reported to the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const uint64_t t0 = global_timer_ns();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0x3FFu) == 0 && global_timer_ns() - t0 > 4000000000ull) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t) ((saddr >> 4) & 0x3FFFu);  // start address, 16-byte units
    d |= (uint64_t) 1 << 16;                   // leading byte offset: unused for swizzled K-major
    d |= (uint64_t) (1024 >> 4) << 32;         // stride byte offset between 8-row groups
    d |= (uint64_t) 1 << 46;                   // descriptor version (sm_100)
    d |= (uint64_t) 2 << 61;                   // SWIZZLE_128B
    return d;
}
__host__ __device__ inline uint32_t make_idesc_i8(bool is_signed) {
    const uint32_t fmt = is_signed ? 1u : 0u;
    return (2u << 4)                    // D format: S32
           | (fmt << 7) | (fmt << 10)   // A, B format: 0 = u8, 1 = s8
           | ((uint32_t) (kTileN >> 3) << 17) | ((uint32_t) (kTileM >> 4) << 24);  // both operands K-major
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory footprint for R resident Q tiles (+1 KB so the operand area can be aligned to 1024 bytes).
__host__ __device__ inline size_t smem_bytes(int R) {
    return 1024 + (size_t) R * kQTileBytes + (size_t) kStages * kPTileBytes + 256 /*barriers*/ +
           (size_t) R * kTileM * (8 + 4 + 4 + 4);
}

__global__ void __launch_bounds__(kThreads, 1)
topk_tc_bytes_kernel(const __grid_constant__ CUtensorMap tmap_p, const __grid_constant__ CUtensorMap tmap_q, const Args a) {
    extern __shared__ unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int R = a.R;
    const int g = blockIdx.x % a.G, s = blockIdx.x / a.G;
    const uint32_t tile_begin = (uint32_t) (((uint64_t) a.tiles_total * s) / a.S);
    const uint32_t tile_end = (uint32_t) (((uint64_t) a.tiles_total * (s + 1)) / a.S);

    // ---- carve shared memory (operand tiles 1024-byte aligned for SWIZZLE_128B) ----
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
    unsigned char* base = smem_raw + pad;
    const uint32_t sQ = raw_addr + pad;
    const uint32_t sP = sQ + R * kQTileBytes;
    unsigned char* after = base + (size_t) R * kQTileBytes + (size_t) kStages * kPTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(after);  // full[kStages], empty[kStages], qfull, tfull[2], tempty[2]
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kStages, bar_q = bar_empty + 8 * kStages,
                   bar_tfull = bar_q + 8, bar_tempty = bar_tfull + 16;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(after + 8 * (2 * kStages + 5));
    uint64_t* s_kth = reinterpret_cast<uint64_t*>(after + 256);             // [R*128]
    int* s_U = reinterpret_cast<int*>(s_kth + R * kTileM);                  // [R*128]
    int* s_cnt = s_U + R * kTileM;
    int* s_nq = s_cnt + R * kTileM;

    const int nslots = R * kTileM;
    const size_t qbase = (size_t) g * nslots;  // first (padded) query handled by this CTA
    for (int i = tid; i < nslots; i += kThreads) {
        s_kth[i] = kEmptyKey;
        s_U[i] = -(1 << 30);
        s_cnt[i] = 0;
        s_nq[i] = a.qnorm[qbase + i];
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_p);
        prefetch_tmap(&tmap_q);
        for (int i = 0; i < kStages; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1);
        }
        mbar_init(bar_q, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 4);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                     "r"((uint32_t) kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            mbar_expect_tx(bar_q, (uint32_t) (R * kQTileBytes));
            for (int r = 0; r < R; ++r) tma_load_2d(sQ + r * kQTileBytes, &tmap_q, bar_q, 0, (int) ((g * R + r) * kTileM));
            for (uint32_t t = tile_begin; t < tile_end; ++t) {
                const uint32_t tl = t - tile_begin, st = tl % kStages, ph = (tl / kStages) & 1;
                mbar_wait(bar_empty + 8 * st, ph ^ 1);
                mbar_expect_tx(bar_full + 8 * st, kPTileBytes);
                tma_load_2d(sP + st * kPTileBytes, &tmap_p, bar_full + 8 * st, 0, (int) (t * kTileN));
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            mbar_wait(bar_q, 0);
            tcgen05_fence_after();
            uint32_t it = 0;
            for (uint32_t t = tile_begin; t < tile_end; ++t) {
                const uint32_t tl = t - tile_begin, st = tl % kStages, ph = (tl / kStages) & 1;
                mbar_wait(bar_full + 8 * st, ph);
                tcgen05_fence_after();
                const uint64_t bdesc = make_sw128_desc(sP + st * kPTileBytes);
                for (int r = 0; r < R; ++r, ++it) {
                    const uint32_t slot = it & 1, sph = (it >> 1) & 1;
                    mbar_wait(bar_tempty + 8 * slot, sph ^ 1);
                    tcgen05_fence_after();
                    const uint64_t adesc = make_sw128_desc(sQ + r * kQTileBytes);
                    const uint32_t dcol = tmem_base + slot * kTileN;
                    for (int ks = 0; ks < a.nks; ++ks)  // +32 bytes of K per step = +2 in 16-byte units
                        umma_i8(dcol, adesc + 2 * ks, bdesc + 2 * ks, a.idesc, ks > 0 ? 1u : 0u);
                    umma_commit(bar_tfull + 8 * slot);  // accumulator ready for the epilogue
                }
                umma_commit(bar_empty + 8 * st);  // point tile consumed -> slot back to the producer
            }
        }
    } else {
        // ===== epilogue: 4 warps, warp (w & 3) owns TMEM lanes [32*(w&3), +32) =====
        const int wq = warp & 3;
        const int qrow = wq * 32 + lane;
        const uint32_t lane_sel = (uint32_t) (wq * 32) << 16;
        const int k = a.k, cap = a.cap;
        uint32_t it = 0;
        int nmin_next = (tile_begin < tile_end) ? __ldg(&a.pnorm[(size_t) tile_begin * kTileN]) : 0;
        for (uint32_t t = tile_begin; t < tile_end; ++t) {
            const int nmin = nmin_next;
            if (t + 1 < tile_end) nmin_next = __ldg(&a.pnorm[(size_t) (t + 1) * kTileN]);
            for (int r = 0; r < R; ++r, ++it) {
                const uint32_t slot = it & 1, sph = (it >> 1) & 1;
                const int qs = r * kTileM + qrow;
                int U = s_U[qs];
                int T = (U + nmin + 1) >> 1;
                mbar_wait(bar_tfull + 8 * slot, sph);
                tcgen05_fence_after();
                const uint32_t tbase = tmem_base + slot * kTileN + lane_sel;
#pragma unroll 1
                for (int c = 0; c < kTileN / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld32(tbase + c * 32, v);
                    tmem_wait_ld();
                    int m = (int) v[0];
#pragma unroll
                    for (int j = 1; j < 32; ++j) m = max(m, (int) v[j]);
                    if (a.mode == 1) {  // roofline probe: keep the loads and the max alive, skip selection
                        if (m == 0x7fffffff) s_U[qs] = m;
                        continue;
                    }
                    if (__any_sync(0xffffffffu, m >= T)) {
                        // ---- slow path: exact re-test of the (rare) values that reached the threshold ----
                        int cnt = s_cnt[qs];
                        uint64_t kth = s_kth[qs];
                        const int nqv = s_nq[qs];
                        uint64_t* mybuf = a.cand + ((size_t) blockIdx.x * nslots + qs) * cap;
                        const size_t p0 = (size_t) t * kTileN + c * 32;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if ((int) v[j] >= T) {
                                const uint32_t id = __ldg(&a.pid[p0 + j]);
                                if (id != kInvalidId) {
                                    const int nx = __ldg(&a.pnorm[p0 + j]);
                                    const int dist = nqv + nx - 2 * (int) v[j];
                                    const uint64_t key = pack_key((float) dist, id);
                                    if (key < kth) mybuf[cnt++] = key;
                                }
                            }
                        }
                        __syncwarp();
                        unsigned need = __ballot_sync(0xffffffffu, cnt > cap - 32);
                        while (need) {
                            const int L = __ffs(need) - 1;
                            need &= need - 1;
                            uint64_t* b = reinterpret_cast<uint64_t*>(
                                __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(mybuf), L));
                            const int n = __shfl_sync(0xffffffffu, cnt, L);
                            const uint64_t newkth = warp_select_topk_dyn(b, n, k, b, lane, cap);
                            if (lane == L) {
                                cnt = min(n, k);
                                kth = newkth;
                                if (kth != kEmptyKey) {
                                    U = nqv - (int) key_dist(kth);
                                    T = (U + nmin + 1) >> 1;
                                }
                            }
                        }
                        s_cnt[qs] = cnt;
                        s_kth[qs] = kth;
                        s_U[qs] = U;
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * slot);
            }
        }
        // ---- final selection: every query's k best of this CTA's point range -> partial[s] ----
        __syncwarp();
        for (int r = 0; r < R; ++r) {
            const int qs = r * kTileM + qrow;
            const int mycnt = s_cnt[qs];
            for (int L = 0; L < 32; ++L) {
                const int qsL = r * kTileM + wq * 32 + L;
                const int n = __shfl_sync(0xffffffffu, mycnt, L);
                const uint64_t* b = a.cand + ((size_t) blockIdx.x * nslots + qsL) * cap;
                uint64_t* dst = a.partial + ((size_t) s * a.G * nslots + qbase + qsL) * k;
                warp_select_topk_dyn(b, n, k, dst, lane, cap);
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t) kTmemCols)
                     : "memory");
    }
}

}  // namespace tc
}  // namespace gpann
