// common.cuh -- shared device helpers for libgpann_cuda (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpann {

constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;
constexpr uint32_t kInvalidId = 0xFFFFFFFFu;

enum Storage : int { kF32 = 1, kU8 = 2, kS8 = 3 };
enum Metric : int { kL2 = 0, kMIPS = 1 };

// ---- (distance,id) keys --------------------------------------------------------------------------------
// A key's unsigned 64-bit order equals std::pair<float,uint32_t>'s lexicographic order (reference
// src/topn.h:12-19): high word = float bits mapped monotonically to uint32, low word = id.
__host__ __device__ __forceinline__ uint32_t float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t pack_key(float dist, uint32_t id) {
    return ((uint64_t) float_to_ordered(dist) << 32) | id;
}
__host__ __device__ __forceinline__ float key_dist(uint64_t key) { return ordered_to_float((uint32_t) (key >> 32)); }
__host__ __device__ __forceinline__ uint32_t key_id(uint64_t key) { return (uint32_t) key; }

// ---- reference-order distances -------------------------------------------------------------------------
// Squared L2 in the reference's summation order (src/dist.cpp:9-16 -> hnswlib space_l2.h:60-91 + :6-20):
// 8 lane accumulators with FMA over the first floor(d/16)*16 elements, lanes added in index order, then a
// sequential FMA tail.  `ld(i)` returns element i of the point as float, `q` is the query row.
template <class LoadX>
__device__ __forceinline__ float sqr_l2_ref_order(LoadX ld, const float* __restrict__ q, int d) {
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f, l4 = 0.f, l5 = 0.f, l6 = 0.f, l7 = 0.f;
    const int d16 = d & ~15;
    for (int i = 0; i < d16; i += 8) {
        float t;
        t = ld(i + 0) - q[i + 0]; l0 = __fmaf_rn(t, t, l0);
        t = ld(i + 1) - q[i + 1]; l1 = __fmaf_rn(t, t, l1);
        t = ld(i + 2) - q[i + 2]; l2 = __fmaf_rn(t, t, l2);
        t = ld(i + 3) - q[i + 3]; l3 = __fmaf_rn(t, t, l3);
        t = ld(i + 4) - q[i + 4]; l4 = __fmaf_rn(t, t, l4);
        t = ld(i + 5) - q[i + 5]; l5 = __fmaf_rn(t, t, l5);
        t = ld(i + 6) - q[i + 6]; l6 = __fmaf_rn(t, t, l6);
        t = ld(i + 7) - q[i + 7]; l7 = __fmaf_rn(t, t, l7);
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(l0, l1), l2), l3), l4), l5), l6), l7);
    float tail = 0.f;
    for (int i = d16; i < d; ++i) {
        const float t = ld(i) - q[i];
        tail = __fmaf_rn(t, t, tail);
    }
    return __fadd_rn(res, tail);
}
// inner_product (src/dist.cpp:42-48): sequential FMA.
template <class LoadX>
__device__ __forceinline__ float inner_product_ref_order(LoadX ld, const float* __restrict__ q, int d) {
    float r = 0.f;
    for (int i = 0; i < d; ++i) r = __fmaf_rn(q[i], ld(i), r);
    return r;
}

// ---- warp-level bitonic sort of 32*E keys held E per lane (element index = e*32 + lane) -------------------
template <int E>
__device__ __forceinline__ void warp_bitonic_sort(uint64_t (&v)[E], int lane) {
    constexpr int N = 32 * E;
#pragma unroll
    for (int k2 = 2; k2 <= N; k2 <<= 1) {
#pragma unroll
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int je = j >> 5;  // partner register distance
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    if ((e & je) == 0) {
                        const int i = e * 32;  // lane bits do not matter for direction when k2 >= 64
                        const bool asc = ((i & k2) == 0);
                        uint64_t a = v[e], b = v[e | je];
                        const bool sw = asc ? (a > b) : (a < b);
                        v[e] = sw ? b : a;
                        v[e | je] = sw ? a : b;
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int i = e * 32 + lane;
                    const bool asc = ((i & k2) == 0);
                    const uint64_t mine = v[e];
                    const uint64_t other = __shfl_xor_sync(0xffffffffu, mine, j);
                    const bool lower = ((lane & j) == 0);
                    // the lower index of the pair keeps min when ascending, max when descending
                    const bool keep_min = (lower == asc);
                    v[e] = keep_min ? (mine < other ? mine : other) : (mine < other ? other : mine);
                }
            }
        }
    }
}

// Select the k smallest of `cnt` keys in buf[0..cnt) (cnt <= 32*E), write them ascending to dst[0..k) padded
// with kEmptyKey, and return the k-th smallest (kEmptyKey if cnt < k) in every lane.  Whole warp cooperates;
// buf/dst may be global or shared (generic pointers) and may alias.
template <int E>
__device__ __forceinline__ uint64_t warp_select_topk(const uint64_t* buf, int cnt, int k, uint64_t* dst, int lane) {
    uint64_t v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = e * 32 + lane;
        v[e] = (i < cnt) ? buf[i] : kEmptyKey;
    }
    __syncwarp();
    warp_bitonic_sort<E>(v, lane);
    uint64_t kth = kEmptyKey;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = e * 32 + lane;
        if (i < k) dst[i] = v[e];
        const uint64_t cand = __shfl_sync(0xffffffffu, v[e], (k - 1) & 31);
        if (e == ((k - 1) >> 5)) kth = cand;
    }
    __syncwarp();
    return kth;
}

// Runtime dispatch on the number of keys per lane (cap <= 32*MAXE, checked by the host API).
// __noinline__: one copy per translation unit (the unrolled sorting networks are large) and the callers' hot
// loops keep their own register budget; every call site is a rare path.
template <int MAXE>
__device__ __noinline__ uint64_t warp_select_topk_dyn(const uint64_t* buf, int cnt, int k, uint64_t* dst, int lane, int cap) {
    if (cap <= 32) return warp_select_topk<1>(buf, cnt, k, dst, lane);
    if (cap <= 64) return warp_select_topk<2>(buf, cnt, k, dst, lane);
    if (cap <= 128 || MAXE <= 4) return warp_select_topk<4>(buf, cnt, k, dst, lane);
    if (cap <= 256 || MAXE <= 8) return warp_select_topk<8>(buf, cnt, k, dst, lane);
    if (cap <= 512 || MAXE <= 16) return warp_select_topk<16>(buf, cnt, k, dst, lane);
    return warp_select_topk<32>(buf, cnt, k, dst, lane);
}

__host__ __device__ __forceinline__ int pow2_ceil(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

}  // namespace gpann
