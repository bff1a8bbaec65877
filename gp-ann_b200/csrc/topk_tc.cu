// topk_tc.cu -- tcgen05 / TMEM / TMA tile kernel: batched byte-vector L2 distances fused with per-query top-k.
//
// Replaces the reference's `for query: for point: distance(); TopN.Add()` (src/recall.h:25-39) for uint8 /
// int8 data (BigANN-style, widened to float by the reference's reader, src/points_io.cpp:63-111).  For
// integer coordinates every partial sum of src/dist.cpp:9-16 is an exact integer below 2^24, so
//     dist(q,x) = |q|^2 + |x|^2 - 2 q.x          (all int32)
// is bit-identical to the reference's float result; q.x is the only O(n*nq*d) term and runs on the tensor
// cores as u8 x u8 -> s32 (tcgen05.mma.kind::i8).
//
// Layout in HBM (built once per point set by prep.cu): point rows are SORTED BY NORM, padded to 128-byte rows
// and to a multiple of 256 rows; norms[] (int32) and ids[] (original id) ride along.  Because a 256-point
// tile then holds near-equal norms, the top-k test against the running k-th best distance kth_q
//     2*acc >= |q|^2 + |x|^2 - kth_q      <=     acc >= ceil((|q|^2 - kth_q + min_tile|x|^2) / 2) =: T
// needs no per-value norm: the epilogue takes a 3-input max over the accumulators straight out of TMEM
// (VIMNMX3, half an instruction per distance) and compares once per 32 values.  Values reaching T are rare
// (k / points-seen): the filter warps turn them into exact (distance,id) keys and push them into a shared-memory
// queue; dedicated SELECTION warps pop the records, keep every query's exact top-k list (replace-the-maximum for
// k <= 32, append + warp-wide bitonic compaction above) and publish the tightened threshold back through shared
// memory.  The tensor pipeline (TMA -> MMA -> TMEM drain) therefore never waits for selection work.
//
// FLOAT DATA (kFloat = true): the same pipeline with bf16 operands (tcgen05.mma.kind::f16, fp32 accumulate) as a
// conservative FILTER: |dot_bf16 - dot| <= eps*|q|*|x| (operand rounding 2^-8 each + fp32 accumulation), so every point
// whose fp32 distance can be <= kth_q satisfies  acc >= T_q(tile)  with the error bound folded into T.  Survivors are
// re-ranked EXACTLY: the selection warp recomputes the distance from the original fp32 rows in the reference's own
// summation order (src/dist.cpp:9-16), so reported distances and tie-breaks are those of the CPU code.
//
// CTA roles (384 threads, 1 CTA/SM, persistent over its point range):
//   warp 0 lane 0 : TMA producer -- Q tiles once, then per point tile 256x128B rows (SWIZZLE_128B) + its 256 norms
//                                   + 256 ids into a kStages ring, all on one mbarrier
//   warp 1 lane 0 : MMA issuer   -- per (point tile, Q tile) "item": ceil(d/32) x tcgen05.mma M128 N256 K32 into
//                                   TMEM accumulator slot (item & 1); tcgen05.commit -> mbarriers
//   warps 2..9    : filter       -- all 8 warps drain every item: warp (w & 3) owns TMEM lanes 32*(w&3).. and the
//                                   column half (w-2)/4 (128 columns): tcgen05.ld 32x32b.x32 x2, max3 tree, one
//                                   compare per 64 values.  They hold no per-query state (read-only thresholds).
//   warps 10..11  : selection    -- one candidate queue each; exact top-k lists in global memory (L2), thresholds in smem
// Grid: G query groups (R Q-tiles of 128 queries, resident in shared memory) x S point splits; split s scans the tiles
// s, s + S, ... (Args::interleave), so every split sees the whole norm range and the candidate records spread evenly.
//
// Sharing between the S CTAs of a query group (a CTA that only knows its own finds lets ~S times the candidates through,
// and with two accumulator slots every candidate record is a detour the MMA issuer ends up waiting for):
//   * every CTA publishes its k-th distance with atomicMin into Args::bound (and, in a sharded call, into the other GPUs'
//     bound arrays over NVLink);
//   * k <= 16: the FIND LOG (Args::flog) -- accepted keys are appended to the group's log and applied by the other CTAs as
//     foreign keys (distance kept, id dropped, blanked before the CTA reports), so each list tracks the group's running
//     k-th distance; main pass only (SelState::flog);
//   * k > 32: an ORDER-STATISTIC bound (Args::qstat) -- each CTA publishes its ceil(k/S)-th best distance at every
//     compaction, the largest of the S values bounds the union's k-th distance (SelState::qstat_*).
// For k <= 32 a CTA hands its (unsorted) lists to merge_keys_kernel; larger k selects its k best first.
#include "common.cuh"
#include "params.h"
#include "tc_ptx.cuh"

// critical-path hand-offs: 0 = mbarrier.try_wait (suspending), 1 = MMA issuer busy-polls, 2 = filter warps too
#ifndef GPANN_SPIN
#define GPANN_SPIN 0
#endif
// 1: a filter warp loads all of its columns of an item before it releases the accumulator slot (128 data registers)
#ifndef GPANN_DRAIN_ALL
#define GPANN_DRAIN_ALL 0
#endif
// (A software-pipelined drain -- the next chunk pair's TMEM loads queued behind the current pair's max trees, across items --
// was built and measured in round 2: 150 vs 96 ms at 100M, see profiles/r2_summary.md; commit 84d16f3 has the code.)
// 1: the lean drain (all 128 columns of a filter warp in registers, one max tree + one branch per item, out-of-line hit scan)
#ifndef GPANN_LEAN
#define GPANN_LEAN 0
#endif
// 1: the hit test of a chunk pair is a plain (divergent) branch instead of a warp vote + uniform branch
#ifndef GPANN_NO_VOTE
#define GPANN_NO_VOTE 0
#endif
#ifndef GPANN_SEL_SLEEP
#define GPANN_SEL_SLEEP 1000
#endif

namespace gpann {
namespace tc {

using namespace ptx;



__device__ __forceinline__ int max3(int a, int b, int c) { return max(max(a, b), c); }  // -> VIMNMX3
// max of 32 accumulators as a depth-4 tree of 3-input maxima (16 independent-ish VIMNMX3 instead of a chain)
__device__ __forceinline__ int max32(const uint32_t (&v)[32]) {
    int a[11];
#pragma unroll
    for (int i = 0; i < 10; ++i) a[i] = max3((int) v[3 * i], (int) v[3 * i + 1], (int) v[3 * i + 2]);
    a[10] = max((int) v[30], (int) v[31]);
    const int b0 = max3(a[0], a[1], a[2]), b1 = max3(a[3], a[4], a[5]), b2 = max3(a[6], a[7], a[8]), b3 = max(a[9], a[10]);
    return max(max3(b0, b1, b2), b3);
}

__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }  // -> FMNMX3
__device__ __forceinline__ float max32f(const uint32_t (&v)[32]) {
    float a[11];
#pragma unroll
    for (int i = 0; i < 10; ++i)
        a[i] = fmax3(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1]), __uint_as_float(v[3 * i + 2]));
    a[10] = fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]));
    const float b0 = fmax3(a[0], a[1], a[2]), b1 = fmax3(a[3], a[4], a[5]), b2 = fmax3(a[6], a[7], a[8]), b3 = fmaxf(a[9], a[10]);
    return fmaxf(fmax3(b0, b1, b2), b3);
}

// ---- candidate queues: epilogue lanes push (key, query slot) records, a selection warp pops them ------------------
// Single consumer per queue; a slot's qs word doubles as its ready flag (kQEmpty = free).
constexpr uint32_t kQEmpty = 0xFFFFFFFFu;

struct Queue {
    uint32_t cap;             // records (power of two)
    uint4* rec;               // [cap] {key lo, key hi, query slot (kQEmpty = free), -}: written and read as ONE 16-byte
                              // shared-memory access, so a record needs no fence between its payload and its ready flag
    uint32_t* tail;           // reserved by producers (atomicAdd)
    volatile uint32_t* head;  // consumed (written by the consumer's lane 0)
};

__device__ __forceinline__ void st_shared_v4(void* p, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4_volatile(const void* p) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)) : "memory");
    return v;
}

// The counters and records live in shared memory: explicit .shared instructions (ATOMS / LDS / STS) -- through the generic
// pointers the compiler emitted ATOM.E / LD.E.STRONG.SYS, the slowest instructions of a candidate record's detour.
__device__ __forceinline__ uint32_t ld_shared_u32_volatile(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __noinline__ void queue_push(const Queue& q, uint64_t key, uint32_t qs, uint32_t extra = 0u) {
    const uint32_t tail_s = smem_u32(q.tail), head_s = smem_u32(const_cast<uint32_t*>(q.head));
    uint32_t pos;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(pos) : "r"(tail_s) : "memory");
    if (pos - ld_shared_u32_volatile(head_s) >= q.cap) {  // full: wait for the selection warp (rare)
        const uint64_t t0 = global_timer_ns();
        while (pos - ld_shared_u32_volatile(head_s) >= q.cap) {
            __nanosleep(64);
            if (global_timer_ns() - t0 > 4000000000ull) __trap();
        }
    }
    st_shared_v4(q.rec + (pos & (q.cap - 1)), (uint32_t) key, (uint32_t) (key >> 32), qs, extra);
}

// v[j] for a run-time j without local memory: a 5-level tree of selects on the bits of j (31 SEL, depth 5).
__device__ __forceinline__ uint32_t pick32(const uint32_t (&v)[32], int j) {
    uint32_t a[16], b[8], c[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (j & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
    const uint32_t d0 = (j & 8) ? c[1] : c[0], d1 = (j & 8) ? c[3] : c[2];
    return (j & 16) ? d1 : d0;
}

// Exact re-test of the values of one 32-column chunk that reached the filter threshold: compute the integer
// distance, build the (distance,id) key and hand it to the selection warp that owns the query.  Branch-free scan:
// a chain of 32 compare+branch pairs costs a lone hit lane ~15 cycles per untaken branch (in-order issue), and every
// cycle a filter warp loses here delays the whole TMEM pipeline, so the hits are collected into a bit mask first.
__device__ __forceinline__ void enqueue_chunk(const uint32_t (&v)[32], int T, const int32_t* tn, const uint32_t* ti, int col0,
                                              int nqv, uint32_t qs, const Queue& q, int mips) {
    uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        m0 |= ((int) v[j] >= T) ? (1u << j) : 0u;
        m1 |= ((int) v[j + 8] >= T) ? (1u << (j + 8)) : 0u;
        m2 |= ((int) v[j + 16] >= T) ? (1u << (j + 16)) : 0u;
        m3 |= ((int) v[j + 24] >= T) ? (1u << (j + 24)) : 0u;
    }
    uint32_t mask = (m0 | m1) | (m2 | m3);
    while (mask) {
        const int j = __ffs(mask) - 1;
        mask &= mask - 1;
        const uint32_t id = ti[col0 + j];
        if (id != kInvalidId) {
            const int acc = (int) pick32(v, j);
            const int dist = mips ? 1 - acc : nqv + tn[col0 + j] - 2 * acc;  // MIPS build: 1 - <q,x> (src/dist.cpp:50-52)
            queue_push(q, pack_key((float) dist, id), qs);
        }
    }
}

// Float variant: the approximate distance is not good enough to rank, so the record carries the point id, the bf16
// accumulator and the point's squared norm; the selection warp re-tests the candidate against the query's CURRENT k-th
// distance (records age in the queue while the lists tighten) and only then recomputes the exact fp32 distance.
__device__ __forceinline__ void enqueue_chunk_f(const uint32_t (&v)[32], float T, const int32_t* tn, const uint32_t* ti, int col0,
                                                uint32_t qs, const Queue& q) {
    uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        m0 |= (__uint_as_float(v[j]) >= T) ? (1u << j) : 0u;
        m1 |= (__uint_as_float(v[j + 8]) >= T) ? (1u << (j + 8)) : 0u;
        m2 |= (__uint_as_float(v[j + 16]) >= T) ? (1u << (j + 16)) : 0u;
        m3 |= (__uint_as_float(v[j + 24]) >= T) ? (1u << (j + 24)) : 0u;
    }
    uint32_t mask = (m0 | m1) | (m2 | m3);
    while (mask) {
        const int j = __ffs(mask) - 1;
        mask &= mask - 1;
        const uint32_t id = ti[col0 + j];
        if (id != kInvalidId) queue_push(q, ((uint64_t) pick32(v, j) << 32) | id, qs, (uint32_t) tn[col0 + j]);
    }
}

// ---- lean drain (GPANN_LEAN, the production filter loop) -----------------------------------------------------------------
// A filter warp pulls ALL of its 128 columns of an item into registers with four x32 loads, hands the accumulator slot back
// and only then reduces: four 32-value max trees, ONE compare and ONE branch per item.  A lane that holds a value at or above
// its threshold passes the 32-column chunk(s) in question through local memory to the out-of-line copy of enqueue_chunk
// below (eight 16-byte stores and loads; the select trees then run on registers again), so the hot loop carries neither the
// hit-mask builders nor the select trees.  The hit path must stay SHORT: with two accumulator slots a warp that lingers
// stalls the MMA issuer after one more item (a serial scan of the lane's values from local memory cost 2x on the whole pass).
template <bool kFloat>
__device__ __noinline__ void slow_chunk(const uint32_t* vp, uint32_t Tbits, const int32_t* tn, const uint32_t* ti, int col0, int nqv,
                                        uint32_t qs, uint32_t* q_tail, volatile uint32_t* q_head, uint4* q_rec, uint32_t q_cap,
                                        int mips) {
    Queue q;
    q.cap = q_cap;
    q.rec = q_rec;
    q.tail = q_tail;
    q.head = q_head;
    uint32_t v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = vp[j];
    if (kFloat) enqueue_chunk_f(v, __uint_as_float(Tbits), tn, ti, col0, qs, q);
    else enqueue_chunk(v, (int) Tbits, tn, ti, col0, nqv, qs, q, mips);
}

// Per-query selection state owned by the selection warps (shared memory, indexed by query slot qs).
struct SelState {
    uint64_t* kth;   // current k-th best key (kEmptyKey until k keys are held)
    int* U;          // |q|^2 - kth distance: what the epilogue turns into its filter threshold
    int* cnt;        // keys held in the query's list
    int* maxpos;     // position of the k-th best (largest kept) key in the unsorted list (k <= 32)
    const int* nq;   // |q|^2
    uint32_t* gbound;  // global [nslots]: float bits of the best known upper bound on each query's k-th distance, shared by
                       // every CTA (point split) that serves the same queries -- atomicMin when a list tightens, re-read
                       // while the selection warp idles
    int nslots;
    // Sharded calls (multi.cu): the other ranks' bound arrays, mapped into this GPU's address space (NVLink peer memory).
    // A tightened bound is pushed to them with one system-scope atomicMin per peer, batched through a dirty bitmap while
    // the selection warp idles -- a shard then filters with the k-th distances of the WHOLE set as they form, instead of
    // rediscovering them (each shard on its own pushes ~k ln(n_shard) records per query through the slow path).
    uint32_t* peer[kMaxPeers];
    int npeers;
    uint32_t* dirty;   // shared-memory bitmap [nslots/32]: bound tightened since the last push to the peers
    uint64_t* lists; // [nslots][stride] keys of this CTA: global (L2), or shared memory when k <= kSmemListKeys
    int k, cap, stride;
    int mips;        // MIPS build: distance = 1 - <q,x>; the threshold source is U = 1 - kth instead of |q|^2 - kth
    // float path: exact re-rank inputs
    const float* rows_f32;  // [n, d]
    const float* q_f32;     // this CTA's queries [nslots, d]
    int d;
    uint32_t id_base;
    float eps;
    // find log of this CTA's query group (Args::flog): every accepted key of this CTA is appended, the other splits'
    // entries are applied to the lists as FOREIGN keys: they tighten the k-th distance but are not this CTA's to report, so
    // they are stored with the id field set to kInvalidId (>= the real key: the list's k-th stays a valid, at most
    // tie-conservative bound, and an own key of the true top-k still has fewer than k keys below it) and blanked at the end
    uint4* flog;
    uint32_t* flog_tail;
    uint32_t flog_cap, split;
    // Order-statistic bound (k > 32, Args::qstat): after every compaction this CTA publishes the distance of its m-th best
    // key, m = ceil(k / S).  If every split of the group holds m keys at or below t, the group holds at least k: the largest of
    // the S published values bounds the k-th distance of the union -- far tighter than the smallest per-split k-th distance
    // once the splits are interleaved (each then sees a 1/S sample of the whole set).
    unsigned char* claim;         // [nslots] shared memory: lane that applies this query's record in the current round
    uint32_t* qstat_mine;         // [nslots] this split's row (float bits; 0xFFFFFFFF = nothing yet)
    const uint32_t* qstat_group;  // [S][nslots] all rows of the group
    int S, qstat_m;
};

// Append the keys this warp accepted in one batch to the group's find log: one reservation (global atomic round trip) per
// batch, one 16-byte store per key.
__device__ __forceinline__ void log_finds(const SelState& s, bool accepted, uint32_t qs, uint64_t key, int lane) {
    const unsigned am = __ballot_sync(0xffffffffu, accepted);
    if (am == 0) return;
    const int leader = __ffs(am) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(s.flog_tail, (uint32_t) __popc(am));
    base = __shfl_sync(0xffffffffu, base, leader);
    const uint32_t pos = base + (uint32_t) __popc(am & ((1u << lane) - 1u));
    if (accepted && pos < s.flog_cap)
        asm volatile("st.global.cg.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(s.flog + pos), "r"((uint32_t) key), "r"((uint32_t) (key >> 32)),
                     "r"(qs | (s.split << 16)), "r"(0u)
                     : "memory");
}

// Squared L2 distance of two fp32 rows in the reference's summation order (src/dist.cpp:9-16 ->
// hnswlib space_l2.h:60-91, 6-20): 8 interleaved FMA lanes over the leading multiple of 16, lanes added in order, then
// a sequential FMA tail.
__device__ __forceinline__ float exact_sqr_l2(const float* __restrict__ x, const float* __restrict__ q, int d) {
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f, l4 = 0.f, l5 = 0.f, l6 = 0.f, l7 = 0.f;
    const int d16 = d & ~15;
    if ((d & 3) == 0) {  // rows are 16-byte aligned
        for (int i = 0; i < d16; i += 8) {
            const float4 xa = __ldg(reinterpret_cast<const float4*>(x + i)), xb = __ldg(reinterpret_cast<const float4*>(x + i + 4));
            const float4 qa = __ldg(reinterpret_cast<const float4*>(q + i)), qb = __ldg(reinterpret_cast<const float4*>(q + i + 4));
            float t;
            t = xa.x - qa.x; l0 = __fmaf_rn(t, t, l0);
            t = xa.y - qa.y; l1 = __fmaf_rn(t, t, l1);
            t = xa.z - qa.z; l2 = __fmaf_rn(t, t, l2);
            t = xa.w - qa.w; l3 = __fmaf_rn(t, t, l3);
            t = xb.x - qb.x; l4 = __fmaf_rn(t, t, l4);
            t = xb.y - qb.y; l5 = __fmaf_rn(t, t, l5);
            t = xb.z - qb.z; l6 = __fmaf_rn(t, t, l6);
            t = xb.w - qb.w; l7 = __fmaf_rn(t, t, l7);
        }
    } else {
        for (int i = 0; i < d16; i += 8) {
            float t;
            t = x[i + 0] - q[i + 0]; l0 = __fmaf_rn(t, t, l0);
            t = x[i + 1] - q[i + 1]; l1 = __fmaf_rn(t, t, l1);
            t = x[i + 2] - q[i + 2]; l2 = __fmaf_rn(t, t, l2);
            t = x[i + 3] - q[i + 3]; l3 = __fmaf_rn(t, t, l3);
            t = x[i + 4] - q[i + 4]; l4 = __fmaf_rn(t, t, l4);
            t = x[i + 5] - q[i + 5]; l5 = __fmaf_rn(t, t, l5);
            t = x[i + 6] - q[i + 6]; l6 = __fmaf_rn(t, t, l6);
            t = x[i + 7] - q[i + 7]; l7 = __fmaf_rn(t, t, l7);
        }
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(l0, l1), l2), l3), l4), l5), l6), l7);
    float tail = 0.f;
    for (int i = d16; i < d; ++i) {
        const float t = x[i] - q[i];
        tail = __fmaf_rn(t, t, tail);
    }
    return __fadd_rn(res, tail);
}

// Insert one record.  k <= 32: unsorted list of the k best with replace-the-maximum (exact threshold after every
// insert).  Larger k: append; the caller compacts a full list warp-wide.  Returns true if the list needs compaction.
template <bool kFloat, bool kMips>
__device__ __forceinline__ void apply_bound(const SelState& s, uint32_t qs, float kth_dist) {
    if (kFloat) {  // U = |q|^2 - kth (MIPS: 1 - kth), float bits in the int arrays; monotone non-decreasing
        const float u = (kMips ? 1.f : __int_as_float(s.nq[qs])) - kth_dist;
        s.U[qs] = __float_as_int(fmaxf(__int_as_float(s.U[qs]), u));
    } else {
        s.U[qs] = max(s.U[qs], (kMips ? 1 : s.nq[qs]) - (int) kth_dist);
    }
}
template <bool kFloat, bool kMips>
__device__ __forceinline__ void publish_threshold(const SelState& s, uint32_t qs, uint64_t kth) {
    const float kd = key_dist(kth);
    apply_bound<kFloat, kMips>(s, qs, kd);
    // distances are non-negative, so their bit patterns order like unsigned integers
    if (s.gbound && kd >= 0.f) {
        // fire and forget (RED.MIN): waiting for the old value cost a global round trip per accepted record.  The dirty bit is set
        // even when another split already published a smaller bound -- the push to the peers then repeats that value.
        atomicMin(s.gbound + qs, __float_as_uint(kd));
        if (s.npeers > 0) atomicOr(s.dirty + (qs >> 5), 1u << (qs & 31));
    }
}
// Pull the bounds other point splits have published for this selection warp's queries into the filter thresholds.
template <bool kFloat, bool kMips>
__device__ __forceinline__ void refresh_bounds(const SelState& s, int sel, int lane) {
    if (!s.gbound) return;
    for (int base = sel * 32; base < s.nslots; base += 32 * kSelWarps) {
        const uint32_t qs = (uint32_t) (base + lane);
        if (s.qstat_group) {  // the union's k-th distance is at most the largest m-th distance of the splits
            uint32_t worst = 0;
            for (int j = 0; j < s.S; ++j) worst = max(worst, __ldcg(s.qstat_group + (size_t) j * s.nslots + qs));
            if (worst < 0x7f800000u && atomicMin(s.gbound + qs, worst) > worst && s.npeers > 0)
                atomicOr(s.dirty + (qs >> 5), 1u << (qs & 31));
        }
        const uint32_t bits = __ldcg(s.gbound + qs);
        const float b = __uint_as_float(bits);
        if (b < 3.0e38f) apply_bound<kFloat, kMips>(s, qs, b);
        if (s.npeers > 0) {  // this selection warp owns the bitmap words of its slots: plain read-and-clear by lane 0
            uint32_t word = 0;
            if (lane == 0) word = atomicExch(s.dirty + (base >> 5), 0u);
            word = __shfl_sync(0xffffffffu, word, 0);
            if ((word >> lane) & 1u)
                for (int pi = 0; pi < s.npeers; ++pi) atomicMin_system(s.peer[pi] + qs, bits);
        }
    }
}

template <bool kFloat, bool kMips>
__device__ __forceinline__ bool select_insert(const SelState& s, uint32_t qs, uint64_t& key, uint32_t extra, bool& accepted,
                                              bool foreign = false) {
    if (foreign) key |= 0xFFFFFFFFull;  // keep the distance, drop the id (see SelState::flog)
    if (kFloat && !foreign) {  // the record holds (accumulator, id) and the point's norm
        const uint32_t id = (uint32_t) key;
        const uint64_t kth_now = s.kth[qs];
        if (kth_now != kEmptyKey) {  // same test as the filter's, with this point's own norm and the current k-th distance
            const float nqv = __int_as_float(s.nq[qs]), nx = __uint_as_float(extra), acc = __uint_as_float((uint32_t) (key >> 32));
            const float lb = kMips ? 1.f - acc - (s.eps * sqrtf(nqv * nx) + 1e-5f * (1.f + fabsf(acc)))
                                    : nqv + nx - 2.f * acc - (2.f * s.eps * sqrtf(nqv * nx) + 1e-5f * (nqv + nx));
            if (lb - 1e-6f * fabsf(lb) > key_dist(kth_now)) return false;
        }
        const float* xrow = s.rows_f32 + (size_t) (id - s.id_base) * s.d;
        // the row is a random 4*d bytes of HBM: request all of its lines at once instead of one miss per loop trip
        for (int off = 128; off < s.d * 4; off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(xrow) + off));
        float dist;
        if (kMips) {  // inner_product (src/dist.cpp:42-48): sequential FMA; mips_distance = 1 - <p,q>
            const float* qrow = s.q_f32 + (size_t) qs * s.d;
            float r = 0.f;
            for (int i = 0; i < s.d; ++i) r = __fmaf_rn(qrow[i], xrow[i], r);
            dist = __fsub_rn(1.0f, r);
        } else {
            dist = exact_sqr_l2(xrow, s.q_f32 + (size_t) qs * s.d, s.d);
        }
        key = pack_key(dist, id);
    }
    uint64_t* list = s.lists + (size_t) qs * s.stride;
    int cnt = s.cnt[qs];
    if (s.k <= 32) {
        if (cnt < s.k) {
            list[cnt++] = key;
            s.cnt[qs] = cnt;
            accepted = true;
            if (cnt < s.k) return false;
        } else {
            if (key >= s.kth[qs]) return false;
            list[s.maxpos[qs]] = key;
        }
        accepted = true;
        // new maximum of the list: the keys sit in global memory (L2), so the loads are issued eight at a time instead
        // of one dependent load per key (a record cost ~3 us of selection-warp time before, ~1M records per sample pass)
        uint64_t m = 0, s1 = kEmptyKey, s2 = kEmptyKey;  // the largest and (order-statistic bound) the two smallest keys
        int mp = 0;
        const bool want_small = s.qstat_mine != nullptr;
        if (!want_small) {
            // maximum + position as a tournament over 16 keys at a time: all loads in flight at once and a dependent chain of
            // 4 compare/select levels instead of 16 (this scan is most of what an accepted record costs a selection warp)
            for (int i0 = 0; i0 < s.k; i0 += 16) {
                uint64_t v[16];
                int pos[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    v[j] = (i0 + j < s.k) ? list[i0 + j] : 0ull;
                    pos[j] = i0 + j;
                }
#pragma unroll
                for (int w = 8; w >= 1; w >>= 1) {
#pragma unroll
                    for (int j = 0; j < w; ++j) {
                        const bool g = v[j + w] > v[j];
                        v[j] = g ? v[j + w] : v[j];
                        pos[j] = g ? pos[j + w] : pos[j];
                    }
                }
                if (v[0] >= m) {
                    m = v[0];
                    mp = pos[0];
                }
            }
        } else
        for (int i0 = 0; i0 < s.k; i0 += 8) {
            uint64_t x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = (i0 + j < s.k) ? list[i0 + j] : 0ull;  // generic: the list may be shared memory
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (i0 + j < s.k && x[j] >= m) {
                    m = x[j];
                    mp = i0 + j;
                }
            if (want_small) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (i0 + j < s.k) {
                        const uint64_t lo = x[j] < s1 ? x[j] : s1, hi = x[j] < s1 ? s1 : x[j];
                        s1 = lo;
                        s2 = hi < s2 ? hi : s2;
                    }
            }
        }
        s.kth[qs] = m;
        s.maxpos[qs] = mp;
        publish_threshold<kFloat, kMips>(s, qs, m);
        if (want_small) {  // k <= 32: qstat_m is 1 or 2 (see the kernel's set-up)
            const float dm = key_dist(s.qstat_m == 1 ? s1 : s2);
            if (dm >= 0.f) __stcg(s.qstat_mine + qs, __float_as_uint(dm));
        }
        return false;
    }
    if (key >= s.kth[qs]) return false;
    list[cnt++] = key;
    s.cnt[qs] = cnt;
    accepted = true;
    return cnt >= s.cap;
}

// Apply the other splits' new finds (the group's find log) to this selection warp's lists.  Never blocks: an entry that is
// reserved but not yet written ends the round, the next poll resumes there.
template <bool kFloat, bool kMips>
__device__ __forceinline__ void poll_log(const SelState& s, uint32_t& log_read, int sel, int lane) {
    if (!s.flog) return;
    const uint32_t tail = min(__ldcg(s.flog_tail), s.flog_cap);
    while (log_read < tail) {
        const uint32_t n = min(tail - log_read, 32u);
        uint4 e = make_uint4(0u, 0u, kQEmpty, 0u);
        if ((uint32_t) lane < n)
            asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(e.x), "=r"(e.y), "=r"(e.z), "=r"(e.w) : "l"(s.flog + log_read + lane) : "memory");
        const unsigned notready = __ballot_sync(0xffffffffu, (uint32_t) lane < n && e.z == kQEmpty);
        const uint32_t m = notready ? (uint32_t) (__ffs(notready) - 1) : n;
        const uint32_t qs = e.z & 0xFFFFu;
        bool pending = (uint32_t) lane < m && (e.z >> 16) != s.split && ((qs >> 5) & (uint32_t) (kSelWarps - 1)) == (uint32_t) sel;
        const uint64_t key = ((uint64_t) e.y << 32) | e.x;
        while (__any_sync(0xffffffffu, pending)) {  // entries for the same query: one per round, lowest lane first
            const unsigned peers = __match_any_sync(0xffffffffu, pending ? qs : 0xFFFF0000u + lane);
            if (pending && (__ffs(peers) - 1) == lane) {
                uint64_t fkey = key;
                bool acc = false;
                select_insert<kFloat, kMips>(s, qs, fkey, 0u, acc, true);
                pending = false;
            }
            __syncwarp();
        }
        log_read += m;
        if (m < n) break;
    }
}

template <bool kFloat, bool kMips>
__device__ __noinline__ void selection_warp_loop(const Queue q, const SelState s, volatile uint32_t* done, int lane, int sel,
                                                 unsigned long long* stats) {
    uint32_t head = 0, idle = 0, busy = 0, log_read = 0;
    unsigned long long n_batches = 0, n_rounds = 0, n_idle = 0, clk_insert = 0, clk_pop = 0;
    while (true) {
        const long long c_top = stats ? clock64() : 0;
        const uint32_t finished = *done;  // read BEFORE the tail: once all producers are done the tail is final
        __threadfence_block();
        const uint32_t tail = *reinterpret_cast<volatile uint32_t*>(q.tail);
        const uint32_t avail = tail - head;
        if (avail == 0) {
            if (finished == (uint32_t) kEpiWarps) break;
            if ((++idle & 7u) == 0) refresh_bounds<kFloat, kMips>(s, sel, lane);  // every ~8 us of idling
            poll_log<kFloat, kMips>(s, log_read, sel, lane);
            busy = 0;
            __nanosleep(GPANN_SEL_SLEEP);  // an idle selection warp should cost (almost) no issue slots or power
            ++n_idle;
            continue;
        }
        const uint32_t n = min(avail, 32u);
        uint32_t qs = 0xFFFF0000u + lane;  // unique dummy for idle lanes (never matches a real slot)
        uint64_t key = 0;
        uint32_t extra = 0;
        bool pending = (uint32_t) lane < n;
        if (pending) {
            uint4* slot = q.rec + ((head + lane) & (q.cap - 1));
            uint4 r = ld_shared_v4_volatile(slot);
            if (r.z == kQEmpty) {
                const uint64_t t0 = global_timer_ns();
                while ((r = ld_shared_v4_volatile(slot)).z == kQEmpty) {
                    if (global_timer_ns() - t0 > 4000000000ull) __trap();
                }
            }
            key = ((uint64_t) r.y << 32) | r.x;
            qs = r.z;
            extra = r.w;
            reinterpret_cast<volatile uint32_t*>(slot)[2] = kQEmpty;
        }
        __threadfence_block();
        __syncwarp();
        head += n;
        if (lane == 0) *q.head = head;  // slots reusable
        const long long c_ins = stats ? clock64() : 0;
        clk_pop += (unsigned long long) (c_ins - c_top);
        ++n_batches;
        // records for the same query are applied one per round (lowest lane first)
        bool accepted = false;  // this lane's record went into its query's list (key then holds the exact key)
        while (__any_sync(0xffffffffu, pending)) {
            ++n_rounds;
            // records of the same query are applied one per round: every pending lane writes its lane id into the query's claim
            // byte and the lane that reads its own id back goes (MATCH.ANY here cost ~18 % of a busy selection warp's time)
            volatile unsigned char* claim = s.claim;
            if (pending) claim[qs] = (unsigned char) lane;
            __syncwarp();
            const bool mine = pending && claim[qs] == (unsigned char) lane;
            bool compact = false;
            if (mine) {
                compact = select_insert<kFloat, kMips>(s, qs, key, extra, accepted);
                pending = false;
            }
            __syncwarp();
            unsigned need = __ballot_sync(0xffffffffu, compact);
            while (need) {  // k > 32 only: a full list is compacted by the whole warp
                const int L = __ffs(need) - 1;
                need &= need - 1;
                const uint32_t qsL = __shfl_sync(0xffffffffu, qs, L);
                uint64_t* b = s.lists + (size_t) qsL * s.stride;
                const uint64_t newkth = warp_select_topk_dyn<16>(b, s.cap, s.k, b, lane, s.cap);
                if (lane == L) {
                    s.cnt[qsL] = s.k;
                    s.kth[qsL] = newkth;
                    publish_threshold<kFloat, kMips>(s, qsL, newkth);
                    if (s.qstat_mine) {
                        const float dm = key_dist(b[s.qstat_m - 1]);  // the list is sorted ascending now
                        if (dm >= 0.f) __stcg(s.qstat_mine + qsL, __float_as_uint(dm));
                    }
                }
                __syncwarp();
            }
        }
        if (s.flog) log_finds(s, accepted, qs, key, lane);
        if (stats) clk_insert += (unsigned long long) (clock64() - c_ins);
        // a busy selection warp (cold start) still trades bounds with the other splits / ranks now and then
        ++busy;
        if ((s.npeers > 0 || s.qstat_group) && (busy & 15u) == 0) refresh_bounds<kFloat, kMips>(s, sel, lane);
        if ((busy & 3u) == 0) poll_log<kFloat, kMips>(s, log_read, sel, lane);
    }
    if (stats && lane == 0 && sel == 0 && blockIdx.x == 0) {  // diagnostics (GPANN_TC_STATS)
        stats[16] = n_batches;
        stats[17] = n_rounds;
        stats[18] = clk_insert;
        stats[19] = n_idle;
        stats[20] = clk_pop;
    }
}

// One tcgen05.mma of an item: A = query tile (shared-memory descriptor, or TMEM columns when kATmem), B = point tile.
template <bool kFloat>
__device__ __forceinline__ void issue_mma(uint32_t dcol, uint64_t adesc, uint64_t bdesc, int a_tmem_col, uint32_t tmem_base,
                                          uint32_t idesc, bool accumulate) {
    if (kATmem) {
        const uint32_t acol = tmem_base + kATmemCol + (uint32_t) a_tmem_col;
        if (kFloat) umma_bf16_ts(dcol, acol, bdesc, idesc, accumulate ? 1u : 0u);
        else umma_i8_ts(dcol, acol, bdesc, idesc, accumulate ? 1u : 0u);
    } else {
        if (kFloat) umma_bf16(dcol, adesc, bdesc, idesc, accumulate ? 1u : 0u);
        else umma_i8(dcol, adesc, bdesc, idesc, accumulate ? 1u : 0u);
    }
}

// kProbe = false is the production instantiation: the probe modes (Args::mode) and the diagnostics (Args::stats) are
// compiled out, because every uniform load + compare + branch they cost sits on the per-item critical path of the
// single MMA-issuing thread and of the filter warps (ncu: ~190 instructions per item in the issuer alone before).
template <bool kFloat, bool kProbe>
__global__ void __launch_bounds__(kThreads, 1)
topk_tc_kernel(const __grid_constant__ CUtensorMap tmap_p, const __grid_constant__ CUtensorMap tmap_q, const Args a) {
    extern __shared__ unsigned char smem_raw[];
    // byte points met queries that are not integer-valued: the asynchronous fall-back (exact float scan) serves the call
    if (a.gate && *reinterpret_cast<const volatile int*>(a.gate) == 0) return;
    const int mode = kProbe ? a.mode : 0;
    unsigned long long* const stats = kProbe ? a.stats : nullptr;
    const long long t_entry = stats ? clock64() : 0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int R = a.R, KB = a.kblocks, NST = a.stages;
    const int g = blockIdx.x % a.G, s = blockIdx.x / a.G;
    // logical tile indices [tile_begin, tile_end) of this split; physical tile = tile_lo + index * tile_stride
    // (a.interleave: split s takes the logical tiles s, s + S, s + 2S, ... instead of a contiguous range, so every split
    // sees the whole norm range -- what the order-statistic bound below needs; the loops then count from 0)
    const uint32_t tile_begin = a.interleave ? 0u : (uint32_t) (((uint64_t) a.ntiles * s) / a.S);
    const uint32_t tile_end = a.interleave ? (a.ntiles > (uint32_t) s ? (a.ntiles - (uint32_t) s + (uint32_t) a.S - 1u) / (uint32_t) a.S : 0u)
                                           : (uint32_t) (((uint64_t) a.ntiles * (s + 1)) / a.S);
    const uint32_t tile_first = a.interleave ? (uint32_t) s : tile_begin, tile_step = a.interleave ? (uint32_t) a.S : 1u;
    const uint32_t q_tile_bytes = (uint32_t) KB * kQBlockBytes, p_tile_bytes = (uint32_t) KB * kPBlockBytes;

    // ---- carve shared memory (operand tiles 1024-byte aligned for SWIZZLE_128B) ----
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
    unsigned char* base = smem_raw + pad;
    const uint32_t sQ = raw_addr + pad;
    const uint32_t sP = sQ + R * q_tile_bytes;
    unsigned char* after = base + (size_t) R * q_tile_bytes + (size_t) NST * p_tile_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(after);  // full[3], empty[3], qfull, tfull[kSlots], tempty[kSlots]
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kMaxStages, bar_q = bar_empty + 8 * kMaxStages,
                   bar_tfull = bar_q + 8, bar_tempty = bar_tfull + 8 * kSlots;
    static_assert(8 * (2 * kMaxStages + 1 + 2 * kSlots) <= 188, "barrier area");
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(after + 188);
    uint32_t* s_ctr = reinterpret_cast<uint32_t*>(after + 192);               // q_tail[2], q_head[2], done
    int32_t* s_norm = reinterpret_cast<int32_t*>(after + 256);                 // [NST][256] int32 or float bits
    uint32_t* s_id = reinterpret_cast<uint32_t*>(s_norm + NST * kTileN);       // [NST][256]
    uint64_t* s_kth = reinterpret_cast<uint64_t*>(s_id + NST * kTileN);        // [R*128]
    uint4* s_qrec = reinterpret_cast<uint4*>(s_kth + R * kTileM);              // [kSelWarps][qcap] 16-byte records
    int* s_U = reinterpret_cast<int*>(s_qrec + kSelWarps * a.qcap);            // [R*128] int or float bits
    int* s_cnt = s_U + R * kTileM;
    int* s_nq = s_cnt + R * kTileM;                                            // int or float bits
    int* s_maxpos = s_nq + R * kTileM;
    uint64_t* s_lists = reinterpret_cast<uint64_t*>(s_maxpos + R * kTileM);   // [R*128][kSmemListKeys] when a.lists_smem
    uint32_t* s_dirty = reinterpret_cast<uint32_t*>(s_lists + (a.lists_smem ? (size_t) R * kTileM * kSmemListKeys : 0));  // [R*4]
    unsigned char* s_claim = reinterpret_cast<unsigned char*>(s_dirty + 16);  // [R*128]
    const bool use_flog = a.flog != nullptr && a.k <= 32;  // poll_log applies foreign keys without the k > 32 compaction step

    const int nslots = R * kTileM;
    const size_t qbase = (size_t) g * nslots;  // first (padded) query handled by this CTA
    for (int i = tid; i < nslots; i += kThreads) {
        s_kth[i] = kEmptyKey;
        s_cnt[i] = 0;
        s_maxpos[i] = 0;
        // warm start: an upper bound on this query's kth distance from the sample passes (+inf = none)
        const float b = a.bound ? __uint_as_float(a.bound[qbase + i]) : __int_as_float(0x7f800000);
        // Padding slots (query index >= nq) hold zero vectors: left alone, every low-norm point "beats" their k-th
        // distance and one CTA per pass drowns in candidate records (measured: the slowest CTA of a sample pass took 35x
        // the others).  Their threshold is pinned above any accumulator value instead, so they never produce a record.
        const bool padding = qbase + i >= (size_t) a.nq;
        if (kFloat) {
            const float nqv = reinterpret_cast<const float*>(a.qnorm)[qbase + i];
            s_nq[i] = __float_as_int(nqv);
            s_U[i] = __float_as_int(padding ? 3.0e38f : (b < 3.0e38f) ? (a.mips ? 1.f : nqv) - b : -3.0e38f);
        } else {
            const int nqv = reinterpret_cast<const int32_t*>(a.qnorm)[qbase + i];
            s_nq[i] = nqv;
            s_U[i] = padding ? (1 << 29) : (b < 3.0e9f) ? (a.mips ? 1 : nqv) - (int) b : -(1 << 30);
        }
    }
    for (int i = tid; i < kSelWarps * a.qcap; i += kThreads) s_qrec[i] = make_uint4(0u, 0u, kQEmpty, 0u);
    if (tid < 8) s_ctr[tid] = 0;
    if (tid < R * (kTileM / 32)) s_dirty[tid] = 0;
    if (warp == kTmaWarp && lane == 0) {
        prefetch_tmap(&tmap_p);
        prefetch_tmap(&tmap_q);
        for (int i = 0; i < kMaxStages; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1 + kEpiWarps);  // MMA commit + every filter warp done with the tile
        }
        mbar_init(bar_q, kATmem ? kEpiWarps : 1);
        for (int i = 0; i < kSlots; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, kFilterByItem ? kEpiWarps / 2 : kEpiWarps);  // one arrive per filter warp on the item
        }
        fence_barrier_init();
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                     "r"((uint32_t) kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *s_tmem;
    const long long t_kernel0 = stats ? clock64() : 0;
    if (stats && tid == 0) atomicMax(stats + 22, (unsigned long long) (t_kernel0 - t_entry));  // slowest CTA: set-up phase
    const int lstride = a.lists_smem ? kSmemListKeys : a.cap;
    uint64_t* my_lists = a.lists_smem ? s_lists : a.cand + (size_t) blockIdx.x * nslots * a.cap;
    const int kb_elems = kFloat ? 64 : 128;  // elements per 128-byte K block

    if (kATmem && warp >= kFilterWarp0 && warp < kFilterWarp0 + kEpiWarps) {
        // Query tiles -> tensor memory, once: lane = query row, 32 columns per 128-byte K block (bytes of the row in
        // order), which is the layout tcgen05.mma expects for an A operand in TMEM.  Warp (w & 3) may touch TMEM lanes
        // [32*(w&3), +32) only; the two column-half groups of filter warps split the R tiles.
        const int wq = warp & 3;
        for (int r = (warp - kFilterWarp0) >> 2; r < R; r += kEpiWarps / 4) {
            const uint4* src = reinterpret_cast<const uint4*>(a.qrows + ((size_t) (g * R + r) * kTileM + wq * 32 + lane) * ((size_t) KB * kRowBytes));
            for (int kb = 0; kb < KB; ++kb) {
                uint32_t v[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint4 x = __ldg(src + kb * 8 + i);
                    v[4 * i] = x.x, v[4 * i + 1] = x.y, v[4 * i + 2] = x.z, v[4 * i + 3] = x.w;
                }
                tmem_st32(tmem_base + kATmemCol + (r * KB + kb) * 32 + ((uint32_t) (wq * 32) << 16), v);
            }
        }
        tmem_wait_st();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_q);
    }

    if (warp == kTmaWarp) {
        if (lane == 0) {
            // ===== TMA producer =====
            if (!kATmem) {
                mbar_expect_tx(bar_q, (uint32_t) R * q_tile_bytes);
                for (int r = 0; r < R; ++r)
                    for (int kb = 0; kb < KB; ++kb)
                        tma_load_2d(sQ + r * q_tile_bytes + kb * kQBlockBytes, &tmap_q, bar_q, kb * kb_elems, (int) ((g * R + r) * kTileM));
            }
            long long w_empty = 0;
            const long long t_begin = stats ? clock64() : 0;
            uint32_t st = 0, ph = 0;
            size_t row0 = ((size_t) a.tile_lo + (size_t) tile_first * a.tile_stride) * kTileN;
            const size_t row_step = (size_t) tile_step * a.tile_stride * kTileN;
            // Pacing: the G CTAs of a point split stream the same tiles; left alone they drift apart by more than the L2
            // holds and every tile is fetched from HBM ~3.6 times (ncu: 49 GB read for 13.6 GB of points).  Every 32 tiles
            // a CTA publishes its position and waits until the slowest CTA of its split is within kPaceWindow tiles.
            volatile uint32_t* pace = a.pace ? a.pace + (size_t) s * a.G : nullptr;
            for (uint32_t t = tile_begin; t < tile_end; ++t, row0 += row_step) {
                if (pace && ((t - tile_begin) & 31u) == 0) {
                    const uint32_t tl = t - tile_begin;
                    pace[g] = tl;
                    if (tl >= a.pace_window) {
                        const uint64_t t0 = global_timer_ns();
                        for (;;) {
                            uint32_t slowest = 0xFFFFFFFFu;
                            for (int j = 0; j < a.G; ++j) slowest = min(slowest, pace[j]);
                            if (slowest + a.pace_window >= tl) break;
                            __nanosleep(2000);
                            if (global_timer_ns() - t0 > 5000000ull) {  // a peer is not running (not co-resident?): stop pacing
                                pace = nullptr;
                                break;
                            }
                        }
                    }
                }
                const long long c0 = stats ? clock64() : 0;
                mbar_wait(bar_empty + 8 * st, ph ^ 1);
                if (stats) w_empty += clock64() - c0;
                mbar_expect_tx(bar_full + 8 * st, p_tile_bytes + 2 * kTileN * 4);
                for (int kb = 0; kb < KB; ++kb)
                    tma_load_2d(sP + st * p_tile_bytes + kb * kPBlockBytes, &tmap_p, bar_full + 8 * st, kb * kb_elems, (int) row0);
                bulk_load_1d(smem_u32(s_norm + st * kTileN), reinterpret_cast<const int32_t*>(a.pnorm) + row0, kTileN * 4,
                             bar_full + 8 * st);
                bulk_load_1d(smem_u32(s_id + st * kTileN), a.pid + row0, kTileN * 4, bar_full + 8 * st);
                if (++st == (uint32_t) NST) {
                    st = 0;
                    ph ^= 1;
                }
            }
            if (a.pace) a.pace[(size_t) s * a.G + g] = 0xFFFFFFFFu;  // done: never holds the others back
            if (stats) {
                atomicAdd(stats + 8, (unsigned long long) w_empty);
                atomicAdd(stats + 9, (unsigned long long) (clock64() - t_begin));
            }
        }
    } else if (warp == kMmaWarp) {
        // ===== MMA issuer: the whole warp walks the loop (so its control flow stays warp-uniform), one elected lane
        // issues the tcgen05.mma / tcgen05.commit of an item.  Shared-memory descriptors differ from a base descriptor
        // only in the 14-bit start-address field, so every operand descriptor is base + a small constant.
        mbar_wait(bar_q, 0);
        tcgen05_fence_after();
        const uint64_t a_base = make_sw128_desc(sQ), b_base = make_sw128_desc(sP);
        const uint32_t idesc = a.idesc;
        const int nks = a.nks;
        uint32_t st = 0, ph = 0, slot = 0, sph = 0;
        long long w_full = 0, w_tempty = 0;
        const long long t_begin = stats ? clock64() : 0;
        // Full-batch fast path (4 resident query tiles, one K block of 4 steps, two 256-column slots): item 4*tile + r
        // always lands in slot r & 1 with phase (r >> 1) & 1, so the unrolled loop has constant barrier addresses,
        // parities and A descriptors -- the issuer's share of the slot hand-shake is a few dozen instructions per item.
        const bool fast_issue = !kProbe && R == 4 && nks == 4 && KB == 1 && kHalves == 1 && kSlots == 2;
        for (uint32_t t = tile_begin; fast_issue && t < tile_end; ++t) {
            mbar_wait(bar_full + 8 * st, ph);
            tcgen05_fence_after();
            const uint64_t b_st = b_base + ((st * p_tile_bytes) >> 4);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                mbar_wait(bar_tempty + 8 * (r & 1), ((r >> 1) & 1) ^ 1);
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint32_t dcol = tmem_base + (r & 1) * kAccN;
                    const uint64_t a_r = a_base + (uint32_t) (r * (kQBlockBytes >> 4));
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) issue_mma<kFloat>(dcol, a_r + 2 * ks, b_st + 2 * ks, r * 32 + 8 * ks, tmem_base, idesc, ks > 0);
                    umma_commit(bar_tfull + 8 * (r & 1));
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(bar_empty + 8 * st);
            __syncwarp();
            if (++st == (uint32_t) NST) {
                st = 0;
                ph ^= 1;
            }
        }
        // The same for 128-column accumulators in four slots (GPANN_ACC_N=128 GPANN_SLOTS=4): 8 items per tile, item j = (query
        // tile j >> 1, point half j & 1) in slot j & 3; a slot is used twice per tile, so the parities are constants again.
        const bool fast_issue8 = !kProbe && R == 4 && nks == 4 && KB == 1 && kHalves == 2 && kSlots == 4;
        for (uint32_t t = tile_begin; fast_issue8 && t < tile_end; ++t) {
            mbar_wait(bar_full + 8 * st, ph);
            tcgen05_fence_after();
            const uint64_t b_st = b_base + ((st * p_tile_bytes) >> 4);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                mbar_wait(bar_tempty + 8 * (j & 3), ((j >> 2) & 1) ^ 1);
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint32_t dcol = tmem_base + (j & 3) * kAccN;
                    const uint64_t a_r = a_base + (uint32_t) ((j >> 1) * (kQBlockBytes >> 4));
                    const uint64_t b_h = b_st + (uint32_t) ((j & 1) * ((kAccN * kRowBytes) >> 4));
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) issue_mma<kFloat>(dcol, a_r + 2 * ks, b_h + 2 * ks, 0, tmem_base, idesc, ks > 0);
                    umma_commit(bar_tfull + 8 * (j & 3));
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(bar_empty + 8 * st);
            __syncwarp();
            if (++st == (uint32_t) NST) {
                st = 0;
                ph ^= 1;
            }
        }
        for (uint32_t t = tile_begin; !fast_issue && !fast_issue8 && t < tile_end; ++t) {
            const long long c0 = stats ? clock64() : 0;
            mbar_wait(bar_full + 8 * st, ph);
            if (stats) w_full += clock64() - c0;
            tcgen05_fence_after();
            const uint64_t b_st = b_base + ((st * p_tile_bytes) >> 4);
            for (int r = 0; r < R; ++r) {
                const uint64_t a_r = a_base + (((uint32_t) r * q_tile_bytes) >> 4);
#pragma unroll
                for (int h = 0; h < kHalves; ++h) {
                    const long long c1 = stats ? clock64() : 0;
                    if (mode != 5) {  // probe 5: free-running MMA stream
                        if (GPANN_SPIN >= 1) mbar_spin(bar_tempty + 8 * slot, sph ^ 1);
                        else mbar_wait(bar_tempty + 8 * slot, sph ^ 1);
                    }
                    if (stats) w_tempty += clock64() - c1;
                    tcgen05_fence_after();
                    if (elect_one()) {
                        const uint32_t dcol = tmem_base + slot * kAccN;
                        // K step ks = 32 bytes: block ks/4 of the row, +2 (16-byte units) per step inside the block;
                        // rows [h*kAccN, +kAccN) of the point tile start h*kAccN*128 bytes into each K block
                        const uint64_t b_h = b_st + (uint32_t) (h * ((kAccN * kRowBytes) >> 4));
                        if (mode != 3) {  // probe 3: no MMA at all, the drain rate alone
                            if (nks == 4) {  // d = 128 bytes / 64 bf16: one K block, fully unrolled
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) issue_mma<kFloat>(dcol, a_r + 2 * ks, b_h + 2 * ks, r * KB * 32 + 8 * ks, tmem_base, idesc, ks > 0);
                            } else {
                                for (int ks = 0; ks < nks; ++ks) {
                                    const int kb = ks >> 2, w = ks & 3;
                                    issue_mma<kFloat>(dcol, a_r + (uint32_t) (kb * (kQBlockBytes >> 4) + 2 * w),
                                                      b_h + (uint32_t) (kb * (kPBlockBytes >> 4) + 2 * w), r * KB * 32 + 8 * ks, tmem_base, idesc, ks > 0);
                                }
                            }
                        }
                        umma_commit(bar_tfull + 8 * slot);  // accumulator ready for the filter warps
                    }
                    __syncwarp();
                    if (++slot == (uint32_t) kSlots) {
                        slot = 0;
                        sph ^= 1;
                    }
                }
            }
            if (elect_one()) umma_commit(bar_empty + 8 * st);  // operands consumed (1 of 1+kEpiWarps arrivals)
            __syncwarp();
            if (++st == (uint32_t) NST) {
                st = 0;
                ph ^= 1;
            }
        }
        if (stats && lane == 0) {
            atomicAdd(stats + 2, (unsigned long long) (clock64() - t_begin));
            atomicAdd(stats + 3, (unsigned long long) w_full);
            atomicAdd(stats + 4, (unsigned long long) w_tempty);
        }
    } else if (warp >= kFilterWarp0 && warp < kFilterWarp0 + kEpiWarps) {
        // ===== filter: every warp drains every item; warp (w & 3) owns TMEM lanes [32*(w&3), +32) and the column
        // half (w-2)/4 of the 256-column accumulator =====
        const int ew = warp - kFilterWarp0;
        const int half = ew >> 2;  // column group of this warp: kEpiWarps/4 groups of kColsPerWarp columns
        const int wq = warp & 3;
        const int qrow = wq * 32 + lane;
        const uint32_t lane_sel = (uint32_t) (wq * 32) << 16;
        // kFilterByItem: the two groups of four filter warps take alternate items (group = accumulator slot) and drain all
        // kAccN columns of their lane quadrant; otherwise every warp drains its column group of every item.
        constexpr int kColsPerWarp = kFilterByItem ? kAccN : kAccN / (kEpiWarps / 4);
        static_assert(kColsPerWarp % 64 == 0, "a filter warp drains 64 columns at a time");
        const int col_base = kFilterByItem ? 0 : half * kColsPerWarp;  // this warp's columns inside an accumulator
        uint32_t slow_entries = 0;
        // the selection warp that owns this warp's queries: (qs >> 5) & 1 = (4 r + wq) & 1 = wq & 1 for every r
        const int sel = wq & (kSelWarps - 1);
        Queue q;
        q.cap = (uint32_t) a.qcap;
        q.rec = s_qrec + sel * a.qcap;
        q.tail = s_ctr + sel;
        q.head = s_ctr + kSelWarps + sel;
        long long w_tfull = 0, w_pfull = 0;
        const long long t_begin = stats ? clock64() : 0;
        uint32_t st = 0, ph = 0, slot = 0, sph = 0;
        const uint32_t tbase0 = tmem_base + col_base + lane_sel;
        const int sample_pass = a.sample_pass;
        const float eps = a.eps;
        const int mips = a.mips;
        // ===== lean drain: the production loop (see slow_chunk above) =====
        constexpr bool kLean = GPANN_LEAN && !kProbe && !kFilterByItem && kHalves == 1 && kColsPerWarp == 128 && kSlots == 2;
        for (uint32_t t = tile_begin; kLean && t < tile_end; ++t) {
            mbar_wait(bar_full + 8 * st, ph);
            const int32_t* tn = s_norm + st * kTileN;
            const uint32_t* ti = s_id + st * kTileN;
            const int nmin1 = tn[0] + 1;
            const float nmin_f = __int_as_float(tn[0]), nmax_f = __int_as_float(tn[kAccN - 1]);
            const float rmin_f = kFloat ? sqrtf(nmin_f) : 0.f, rmax_f = kFloat ? sqrtf(nmax_f) : 0.f;
            for (int r = 0; r < R; ++r) {
                const uint32_t qs = (uint32_t) (r * kTileM + qrow);
                const int Uraw = *reinterpret_cast<volatile int*>(s_U + qs);
                const int nqraw = s_nq[qs];
                int Ti = 0;
                float Tf = 0.f;
                if (kFloat) {  // derivation: the generic loop below
                    const float nqv = __int_as_float(nqraw), U = __int_as_float(Uraw);
                    const float sq = 2.f * eps * sqrtf(nqv);
                    const float g0 = nmin_f - sq * rmin_f - 1e-5f * (nqv + nmin_f);
                    const float g1 = nmax_f - sq * rmax_f - 1e-5f * (nqv + nmax_f);
                    float gmin = fminf(g0, g1);
                    const float nstar = eps * eps * nqv;
                    if (nmin_f <= 1.001f * nstar && nstar <= 1.001f * nmax_f) gmin = -1.001f * nstar - 2e-5f * nqv;
                    Tf = 0.5f * (U + gmin);
                    if (mips) {
                        const float qx = sqrtf(nqv) * rmax_f;
                        Tf = U - eps * qx - 1e-5f * (1.f + fabsf(U) + qx);
                    }
                    Tf -= 1e-6f * fabsf(Tf);
                } else {
                    Ti = mips ? Uraw : (Uraw + nmin1) >> 1;
                }
                mbar_wait(bar_tfull + 8 * slot, sph);
                tcgen05_fence_after();
                const uint32_t tbase = tbase0 + slot * kAccN;
                // kLeanGroups loads of (4 / kLeanGroups) x 32 columns: 1 = everything up front (128 data registers), 2 = two
                // 64-column phases like the generic loop (the slot goes back after the second load)
                constexpr int kLeanGroups = GPANN_LEAN == 2 ? 2 : 1, kPer = 4 / kLeanGroups;
                const uint32_t slot_now = slot;
                if (++slot == (uint32_t) kSlots) {
                    slot = 0;
                    sph ^= 1;
                }
#pragma unroll
                for (int gq = 0; gq < kLeanGroups; ++gq) {
                    uint32_t v[kPer][32];
#pragma unroll
                    for (int c = 0; c < kPer; ++c) tmem_ld32(tbase + (gq * kPer + c) * 32, v[c]);
                    tmem_wait_ld();            // warp-wide (.sync.aligned): every lane's columns are in registers
                    if (gq == kLeanGroups - 1) {
                        tcgen05_fence_before();
                        if (lane == 0) mbar_arrive(bar_tempty + 8 * slot_now);  // slot back to the MMA issuer before the arithmetic
                    }
                    // chunk maxima, one compare, one (divergent) branch
                    if (kFloat) {
                        float m[kPer], mm;
#pragma unroll
                        for (int c = 0; c < kPer; ++c) m[c] = max32f(v[c]);
                        mm = m[0];
#pragma unroll
                        for (int c = 1; c < kPer; ++c) mm = fmaxf(mm, m[c]);
                        if (mm >= Tf) {
                            const bool cold = sample_pass && Tf < -1.0e37f;
#pragma unroll
                            for (int c = 0; c < kPer; ++c) {
                                if (m[c] >= Tf) {
                                    uint32_t tmp[32];
#pragma unroll
                                    for (int j = 0; j < 32; ++j) tmp[j] = v[c][j];
                                    slow_chunk<true>(tmp, __float_as_uint(cold ? m[c] : Tf), tn, ti, col_base + 32 * (gq * kPer + c), nqraw, qs,
                                                     q.tail, q.head, q.rec, q.cap, mips);
                                }
                            }
                        }
                    } else {
                        int m[kPer], mm;
#pragma unroll
                        for (int c = 0; c < kPer; ++c) m[c] = max32(v[c]);
                        mm = m[0];
#pragma unroll
                        for (int c = 1; c < kPer; ++c) mm = max(mm, m[c]);
                        if (mm >= Ti) {
                            const bool cold = sample_pass && Ti < -(1 << 28);
#pragma unroll
                            for (int c = 0; c < kPer; ++c) {
                                if (m[c] >= Ti) {
                                    uint32_t tmp[32];
#pragma unroll
                                    for (int j = 0; j < 32; ++j) tmp[j] = v[c][j];
                                    slow_chunk<false>(tmp, (uint32_t) (cold ? m[c] : Ti), tn, ti, col_base + 32 * (gq * kPer + c), nqraw, qs,
                                                      q.tail, q.head, q.rec, q.cap, mips);
                                }
                            }
                        }
                    }
                    if (gq + 1 < kLeanGroups) __syncwarp();
                }
                __syncwarp();  // the next tcgen05.ld is .aligned
            }
            if (lane == 0) mbar_arrive(bar_empty + 8 * st);  // this warp no longer needs the tile's norms/ids
            if (++st == (uint32_t) NST) {
                st = 0;
                ph ^= 1;
            }
        }
        for (uint32_t t = tile_begin; !kLean && t < tile_end; ++t) {
            // the tile's norms/ids arrive with its operands; the MMA is ahead of us, so this is normally immediate
            const long long c0 = stats ? clock64() : 0;
            mbar_wait(bar_full + 8 * st, ph);
            if (stats) w_pfull += clock64() - c0;
            const int32_t* tn = s_norm + st * kTileN;
            const uint32_t* ti = s_id + st * kTileN;
            // norm range of each item's rows (rows are sorted by norm): first and last row
            int nmin_i[kHalves];
            float nmin_fa[kHalves], nmax_fa[kHalves], rmin_fa[kHalves], rmax_fa[kHalves];
#pragma unroll
            for (int h = 0; h < kHalves; ++h) {
                nmin_i[h] = tn[h * kAccN];
                nmin_fa[h] = __int_as_float(tn[h * kAccN]);
                nmax_fa[h] = __int_as_float(tn[h * kAccN + kAccN - 1]);
                rmin_fa[h] = kFloat ? sqrtf(nmin_fa[h]) : 0.f;
                rmax_fa[h] = kFloat ? sqrtf(nmax_fa[h]) : 0.f;
            }
            for (int r = 0; r < R; ++r) {
                const uint32_t qs = (uint32_t) (r * kTileM + qrow);
#pragma unroll
                for (int h = 0; h < kHalves; ++h) {
                    const uint32_t bar_tf = bar_tfull + 8 * slot, bar_te = bar_tempty + 8 * slot, par_tf = sph;
                    const uint32_t tbase = tbase0 + slot * kAccN;
                    const bool mine = !kFilterByItem || (slot & 1u) == (uint32_t) half;  // by item: group g takes the slots of parity g
                    if (++slot == (uint32_t) kSlots) {
                        slot = 0;
                        sph ^= 1;
                    }
                    if (!mine) continue;
                    const int Uraw = *reinterpret_cast<volatile int*>(s_U + qs);
                    const int nqraw = s_nq[qs];
                    const int row0 = h * kAccN;  // first tile row of this item
                    int Ti = 0;
                    float Tf = 0.f;
                    if (kFloat) {
                        // keep x if  nq + nx - 2*dot~ - slack(nq,nx) <= kth,  slack = 2*eps*sqrt(nq*nx) + 1e-5*(nq+nx);
                        // g(nx) = nx - slack is evaluated at both ends of the item's norm range (it is monotone in between)
                        const float nmin_f = nmin_fa[h], nmax_f = nmax_fa[h], rmin_f = rmin_fa[h], rmax_f = rmax_fa[h];
                        const float nqv = __int_as_float(nqraw), U = __int_as_float(Uraw);
                        const float sq = 2.f * eps * sqrtf(nqv);
                        const float g0 = nmin_f - sq * rmin_f - 1e-5f * (nqv + nmin_f);
                        const float g1 = nmax_f - sq * rmax_f - 1e-5f * (nqv + nmax_f);
                        float gmin = fminf(g0, g1);
                        // g is a convex parabola in sqrt(nx) with its minimum at nx* = (eps^2)*nq: if the item straddles it
                        // (points with norms far below the query's), use the global minimum instead of the end points
                        const float nstar = eps * eps * nqv;
                        if (nmin_f <= 1.001f * nstar && nstar <= 1.001f * nmax_f) gmin = -1.001f * nstar - 2e-5f * nqv;
                        Tf = 0.5f * (U + gmin);
                        if (mips) {  // keep x if 1 - dot <= kth: dot~ >= U - eps*|q|*|x|, |x| <= the item's largest norm
                            const float qx = sqrtf(nqv) * rmax_f;
                            Tf = U - eps * qx - 1e-5f * (1.f + fabsf(U) + qx);
                        }
                        Tf -= 1e-6f * fabsf(Tf);
                    } else {
                        Ti = mips ? Uraw : (Uraw + nmin_i[h] + 1) >> 1;  // MIPS: acc >= 1 - kth, exact in integers
                    }
                    if (mode == 5) continue;
                    const long long c1 = stats ? clock64() : 0;
                    if (GPANN_SPIN >= 2) mbar_spin(bar_tf, par_tf);
                    else mbar_wait(bar_tf, par_tf);
                    if (stats) w_tfull += clock64() - c1;
                    tcgen05_fence_after();
                    if (mode == 2) {  // probe: MMA issue rate only
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_te);
                        continue;
                    }
                    // kColsPerWarp columns per warp and item, 64 at a time (two x32 loads in flight).  A software pipeline
                    // of single x32 loads (next chunk in flight during the max tree) was measured SLOWER (120 -> 135 ms):
                    // the TMEM read path is not the limit (~1 KB/clk/SM, profiles/tools/tmem_ld_bench.cu); what matters is
                    // how soon after tcgen05.commit the slot goes back to the MMA issuer.
                    uint32_t vv[kColsPerWarp / 32][32];
                    if (GPANN_DRAIN_ALL) {  // variant: every column of the warp into registers first, slot back at once
#pragma unroll
                        for (int j = 0; j < kColsPerWarp / 32; ++j) tmem_ld32(tbase + j * 32, vv[j]);
                        tmem_wait_ld();
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_te);
                    }
#pragma unroll
                    for (int c = 0; c < kColsPerWarp / 64; ++c) {
                        uint32_t (&va)[32] = vv[2 * c];
                        uint32_t (&vb)[32] = vv[2 * c + 1];
                        if (!GPANN_DRAIN_ALL) {
                            tmem_ld32(tbase + c * 64, va);
                            tmem_ld32(tbase + c * 64 + 32, vb);
                            tmem_wait_ld();
                            if (c == kColsPerWarp / 64 - 1) {  // accumulator drained into registers: hand the slot back before
                                tcgen05_fence_before();        // the arithmetic, so a slow-path entry cannot stall the MMA
                                if (!GPANN_NO_VOTE) __syncwarp();  // (tcgen05.wait::ld is warp-wide: nothing to wait for)
                                if (lane == 0) mbar_arrive(bar_te);
                            }
                        }
                        const int col0 = row0 + col_base + c * 64;  // tile row of va[0]
                        if (kFloat) {
                            const float ma = max32f(va), mb = max32f(vb);
                            if (mode == 1 || mode == 3) {  // probe: keep the loads and the max alive, skip selection
                                if (fmaxf(ma, mb) == 3.0e38f) s_U[qs] = __float_as_int(ma);
                                continue;
                            }
                            if (GPANN_NO_VOTE ? fmaxf(ma, mb) >= Tf : __any_sync(0xffffffffu, fmaxf(ma, mb) >= Tf)) {
                                // cold lane of a SAMPLE pass (no bound yet): any subset of real distances bounds the k-th one,
                                // so only the best value of each chunk goes to the selection warp instead of all 32
                                const bool cold = sample_pass && Tf < -1.0e37f;
                                if (ma >= Tf) enqueue_chunk_f(va, cold ? ma : Tf, tn, ti, col0, qs, q);
                                if (mb >= Tf) enqueue_chunk_f(vb, cold ? mb : Tf, tn, ti, col0 + 32, qs, q);
                                if (!GPANN_NO_VOTE) __syncwarp();
                            }
                            if (GPANN_NO_VOTE) __syncwarp();  // divergent hit branch: reconverge before the next .aligned load
                        } else {
                            const int ma = max32(va), mb = max32(vb);
                            if (mode == 1 || mode == 3) {
                                if (max(ma, mb) == 0x7fffffff) s_U[qs] = ma;
                                continue;
                            }
                            if (GPANN_NO_VOTE ? max(ma, mb) >= Ti : __any_sync(0xffffffffu, max(ma, mb) >= Ti)) {
                                ++slow_entries;
                                const bool cold = sample_pass && Ti < -(1 << 28);  // see the float branch
                                if (ma >= Ti) enqueue_chunk(va, cold ? ma : Ti, tn, ti, col0, nqraw, qs, q, mips);
                                if (mb >= Ti) enqueue_chunk(vb, cold ? mb : Ti, tn, ti, col0 + 32, nqraw, qs, q, mips);
                                if (!GPANN_NO_VOTE) __syncwarp();
                            }
                            if (GPANN_NO_VOTE) __syncwarp();
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8 * st);  // this warp no longer needs the tile's norms/ids
            if (++st == (uint32_t) NST) {
                st = 0;
                ph ^= 1;
            }
        }
        __threadfence_block();
        __syncwarp();
        if (stats && lane == 0) atomicAdd(stats + 1, (unsigned long long) slow_entries);
        if (stats && lane == 0 && ew == 0 && blockIdx.x == 0) stats[10] = (unsigned long long) (clock64() - t_kernel0);  // filter loop done
        if (stats && lane == 0 && ew == 0) atomicMax(stats + 23, (unsigned long long) (clock64() - t_kernel0));  // slowest CTA
        if (stats && lane == 0 && ew == 0) {
            atomicAdd(stats + 5, (unsigned long long) (clock64() - t_begin));
            atomicAdd(stats + 6, (unsigned long long) w_tfull);
            atomicAdd(stats + 7, (unsigned long long) w_pfull);
        }
        if (lane == 0) atomicAdd(s_ctr + 2 * kSelWarps, 1u);  // done: every record of this warp is in a queue
    } else {
        // ===== selection warps: exact per-query top-k lists, thresholds published through s_U =====
        const int sel = warp - kSelWarp0;
        Queue q;
        q.cap = (uint32_t) a.qcap;
        q.rec = s_qrec + sel * a.qcap;
        q.tail = s_ctr + sel;
        q.head = s_ctr + kSelWarps + sel;
        SelState ss;
        ss.kth = s_kth;
        ss.U = s_U;
        ss.cnt = s_cnt;
        ss.maxpos = s_maxpos;
        ss.nq = s_nq;
        ss.lists = my_lists;
        ss.k = a.k;
        ss.cap = a.cap;
        ss.stride = lstride;
        ss.rows_f32 = a.rows_f32;
        ss.q_f32 = a.q_f32 ? a.q_f32 + qbase * a.d : nullptr;
        ss.d = a.d;
        ss.id_base = a.id_base;
        ss.eps = a.eps;
        ss.mips = a.mips;
        ss.gbound = a.bound ? a.bound + qbase : nullptr;
        ss.npeers = a.bound ? a.npeers : 0;
        for (int pi = 0; pi < kMaxPeers; ++pi) ss.peer[pi] = pi < a.npeers ? a.peer_bound[pi] + qbase : nullptr;
        ss.dirty = s_dirty;
        ss.claim = s_claim;
        ss.nslots = nslots;
        ss.flog = use_flog ? a.flog + (size_t) g * a.flog_cap : nullptr;
        ss.flog_tail = use_flog ? a.flog_tail + g : nullptr;
        ss.flog_cap = a.flog_cap;
        ss.split = (uint32_t) s;
        // order-statistic bound: the append + compaction lists (k > 32) publish at every compaction; the replace-the-maximum
        // lists (k <= 32) track their two smallest keys, so they take part when ceil(k / S) <= 2 -- and only where the find
        // log is off (the sample passes): with the log a list holds foreign keys and its order statistics are the group's
        const int qm = (a.k + a.S - 1) / a.S;
        const bool use_qstat = a.qstat != nullptr && a.bound != nullptr && (a.k > 32 || (qm <= 2 && !use_flog));
        ss.qstat_group = use_qstat ? a.qstat + (size_t) g * a.S * nslots : nullptr;
        ss.qstat_mine = use_qstat ? a.qstat + ((size_t) g * a.S + s) * nslots : nullptr;
        ss.S = a.S;
        ss.qstat_m = qm;
        if (a.mips) selection_warp_loop<kFloat, true>(q, ss, s_ctr + 2 * kSelWarps, lane, sel, stats);
        else selection_warp_loop<kFloat, false>(q, ss, s_ctr + 2 * kSelWarps, lane, sel, stats);
        if (stats && lane == 0) atomicAdd(stats, (unsigned long long) *q.tail);  // records pushed to this queue
    }

    // ---- final selection: every query's k best of this CTA's point range -> partial[s], by all non-role warps ----
    if (warp >= kFilterWarp0 && warp < kFilterWarp0 + kEpiWarps + kSelWarps) {
        asm volatile("bar.sync 1, %0;" ::"n"(32 * (kEpiWarps + kSelWarps)) : "memory");
        if (stats && tid == kFilterWarp0 * 32 && blockIdx.x == 0) stats[11] = (unsigned long long) (clock64() - t_kernel0);  // selection warps drained
        if (a.k <= 32) {
            // replace-the-maximum lists hold at most k keys: they go out as they are (unsorted, foreign keys and free positions
            // as kEmptyKey) -- merge_keys_kernel selects, and a sort per query here cost ~2k clk x 51 queries per warp, most
            // of what a sample pass spends after its last tile
            const int t0 = (warp - kFilterWarp0) * 32 + lane, nthr = 32 * (kEpiWarps + kSelWarps);
            for (int i = t0; i < nslots * a.k; i += nthr) {
                const int qs = i / a.k, j = i - qs * a.k;
                uint64_t v = j < s_cnt[qs] ? my_lists[(size_t) qs * lstride + j] : kEmptyKey;
                if (use_flog && (uint32_t) v == kInvalidId) v = kEmptyKey;  // the other splits' keys served the threshold only
                a.partial[((size_t) s * a.G * nslots + qbase + qs) * a.k + j] = v;
            }
        } else
        for (int qs = warp - kFilterWarp0; qs < nslots; qs += kEpiWarps + kSelWarps) {
            const uint64_t* b = my_lists + (size_t) qs * lstride;
            uint64_t* dst = a.partial + ((size_t) s * a.G * nslots + qbase + qs) * a.k;
            const int cnt = s_cnt[qs];
            if (use_flog) {  // the other splits' keys served the threshold only: they are not this CTA's to report
                uint64_t* wl = my_lists + (size_t) qs * lstride;
                for (int j = lane; j < cnt; j += 32)
                    if ((uint32_t) wl[j] == kInvalidId) wl[j] = kEmptyKey;
                __syncwarp();
            }
            warp_select_topk_dyn<16>(b, cnt, a.k, dst, lane, min(a.cap, max(32, pow2_ceil(max(cnt, a.k)))));
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (stats && tid == 0 && blockIdx.x == 0) stats[12] = (unsigned long long) (clock64() - t_kernel0);  // final selection done
    if (warp == kMmaWarp) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t) kTmemCols)
                     : "memory");
        if (stats && lane == 0) atomicMax(stats + 21, (unsigned long long) (clock64() - t_entry));  // slowest CTA: whole kernel
    }
}

}  // namespace tc

cudaError_t tc_set_attributes() {
    const int big = 227 * 1024;
    cudaError_t e = cudaFuncSetAttribute(tc::topk_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::topk_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::topk_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::topk_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    return e;
}

void launch_topk_tc(bool is_float, const CUtensorMap& tmap_p, const CUtensorMap& tmap_q, const tc::Args& a, unsigned grid,
                    size_t smem, cudaStream_t st) {
    const bool probe = a.mode != 0 || a.stats != nullptr;  // GPANN_TC_MODE / GPANN_TC_STATS: the instrumented instantiation
    if (is_float) {
        if (probe) tc::topk_tc_kernel<true, true><<<grid, tc::kThreads, smem, st>>>(tmap_p, tmap_q, a);
        else tc::topk_tc_kernel<true, false><<<grid, tc::kThreads, smem, st>>>(tmap_p, tmap_q, a);
    } else {
        if (probe) tc::topk_tc_kernel<false, true><<<grid, tc::kThreads, smem, st>>>(tmap_p, tmap_q, a);
        else tc::topk_tc_kernel<false, false><<<grid, tc::kThreads, smem, st>>>(tmap_p, tmap_q, a);
    }
}

}  // namespace gpann
