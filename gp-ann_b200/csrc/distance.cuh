// distance.cuh -- point-row loaders and the reference-order distance shared by the CUDA-core kernels.
#pragma once
#include "common.cuh"

namespace gpann {

template <int STORAGE>
__device__ __forceinline__ void load8(const char* row, int i, float (&x)[8]) {
    if constexpr (STORAGE == kF32) {
        const float4 a = *reinterpret_cast<const float4*>(row + (size_t) i * 4);
        const float4 b = *reinterpret_cast<const float4*>(row + (size_t) i * 4 + 16);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
        const uint2 w = *reinterpret_cast<const uint2*>(row + i);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t word = j < 4 ? w.x : w.y;
            const uint32_t byte = (word >> (8 * (j & 3))) & 0xFFu;
            x[j] = STORAGE == kU8 ? (float) byte : (float) (int) (int8_t) byte;
        }
    }
}
template <int STORAGE>
__device__ __forceinline__ float load1(const char* row, int i) {
    if constexpr (STORAGE == kF32) return reinterpret_cast<const float*>(row)[i];
    if constexpr (STORAGE == kU8) return (float) reinterpret_cast<const uint8_t*>(row)[i];
    return (float) reinterpret_cast<const int8_t*>(row)[i];
}

// Distance of one shared-memory point row to one shared-memory query row, reference arithmetic order.
template <int STORAGE>
__device__ __forceinline__ float ref_distance(const char* row, const float* q, int d, int metric) {
    if (metric == kL2) {
        float l[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int d16 = d & ~15;
        for (int i = 0; i < d16; i += 8) {
            float x[8];
            load8<STORAGE>(row, i, x);
            const float4 qa = *reinterpret_cast<const float4*>(q + i);
            const float4 qb = *reinterpret_cast<const float4*>(q + i + 4);
            float t;
            t = x[0] - qa.x; l[0] = __fmaf_rn(t, t, l[0]);
            t = x[1] - qa.y; l[1] = __fmaf_rn(t, t, l[1]);
            t = x[2] - qa.z; l[2] = __fmaf_rn(t, t, l[2]);
            t = x[3] - qa.w; l[3] = __fmaf_rn(t, t, l[3]);
            t = x[4] - qb.x; l[4] = __fmaf_rn(t, t, l[4]);
            t = x[5] - qb.y; l[5] = __fmaf_rn(t, t, l[5]);
            t = x[6] - qb.z; l[6] = __fmaf_rn(t, t, l[6]);
            t = x[7] - qb.w; l[7] = __fmaf_rn(t, t, l[7]);
        }
        float res = __fadd_rn(l[0], l[1]);
#pragma unroll
        for (int j = 2; j < 8; ++j) res = __fadd_rn(res, l[j]);
        float tail = 0.f;
        for (int i = d16; i < d; ++i) {
            const float t = load1<STORAGE>(row, i) - q[i];
            tail = __fmaf_rn(t, t, tail);
        }
        return __fadd_rn(res, tail);
    } else {
        float r = 0.f;
        for (int i = 0; i < d; ++i) r = __fmaf_rn(q[i], load1<STORAGE>(row, i), r);
        return __fsub_rn(1.0f, r);
    }
}

// The same arithmetic with element-wise loads: for rows / queries in global memory whose alignment is not known.
template <int STORAGE>
__device__ __forceinline__ float ref_distance_scalar(const char* row, const float* q, int d, int metric) {
    if (metric == kL2) return sqr_l2_ref_order([row](int i) { return load1<STORAGE>(row, i); }, q, d);
    return __fsub_rn(1.0f, inner_product_ref_order([row](int i) { return load1<STORAGE>(row, i); }, q, d));
}

}  // namespace gpann
