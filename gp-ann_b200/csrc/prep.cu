// prep.cu -- one-time data preparation kernels: storage detection/narrowing (the reference's reader widens
// .u8bin/.i8bin to float, src/points_io.cpp:63-111; this undoes it on upload), norms, and the norm-sorted,
// 128-byte-row layout the tensor kernel streams (see topk_tc.cuh).
#include <algorithm>

#include "common.cuh"
#include "params.h"

namespace gpann {

// flags[0] &= all values are integers in [0,255]; flags[1] &= all values are integers in [-128,127]
__global__ void detect_bytes_kernel(const float* x, size_t count, int* flags) {
    bool u8 = true, s8 = true;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t) gridDim.x * blockDim.x) {
        const float v = x[i];
        const bool integral = (v == rintf(v));
        u8 &= integral && v >= 0.f && v <= 255.f;
        s8 &= integral && v >= -128.f && v <= 127.f;
    }
    if (!__all_sync(0xffffffffu, u8) && (threadIdx.x & 31) == 0) atomicAnd(&flags[0], 0);
    if (!__all_sync(0xffffffffu, s8) && (threadIdx.x & 31) == 0) atomicAnd(&flags[1], 0);
}

// float rows [n,d] -> byte rows [n, pitch] (zero padded), as u8 or s8 bit patterns
__global__ void narrow_rows_kernel(const float* x, size_t n, int d, int pitch, unsigned char* out) {
    const size_t total = n * (size_t) pitch;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t) gridDim.x * blockDim.x) {
        const size_t r = i / pitch;
        const int j = (int) (i - r * pitch);
        out[i] = j < d ? (unsigned char) (int) x[r * d + j] : 0;
    }
}

// Per-row squared norm of byte rows (exact int32; d <= 256 keeps it below 2^24).  Scalar form: any pitch / alignment.
__global__ void byte_norms_kernel(const unsigned char* rows, size_t n, int d, size_t pitch, int is_signed, uint32_t* norms,
                                  uint32_t* index) {
    for (size_t r = (size_t) blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (size_t) gridDim.x * blockDim.x) {
        const unsigned char* row = rows + r * pitch;
        int acc = 0;
        for (int j = 0; j < d; ++j) {
            const int v = is_signed ? (int) (int8_t) row[j] : (int) row[j];
            acc += v * v;
        }
        norms[r] = (uint32_t) acc;
        if (index) index[r] = (uint32_t) r;
    }
}

// HBM-rate form for 16-byte aligned rows (d % 16 == 0): `lpr` lanes (a power of two >= d/16) share a row, each fetching
// 16-byte pieces (coalesced: a warp reads 32/lpr consecutive rows per load) and squaring them with dp4a; four row groups
// are in flight per warp.  The scalar kernel above walked a 128-byte row byte by byte per thread and ran at 4 % of HBM
// (47 ms for 100M x 128: two thirds of a full scan).
template <bool kSigned>
__global__ void __launch_bounds__(256) byte_norms_vec_kernel(const unsigned char* __restrict__ rows, size_t n, int d, size_t pitch,
                                                             int lpr, uint32_t* __restrict__ norms, uint32_t* __restrict__ index) {
    const int lane = threadIdx.x & 31;
    const int rpw = 32 / lpr;                  // rows per warp and load
    const int sub = lane / lpr, l = lane - sub * lpr;
    const int pieces = d >> 4;
    const size_t warps = ((size_t) gridDim.x * blockDim.x) >> 5;
    const size_t gw = ((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    constexpr int kU = 4;
    for (size_t r0 = gw * (size_t) (rpw * kU); r0 < n; r0 += warps * (size_t) (rpw * kU)) {
        int acc[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            acc[u] = 0;
            const size_t r = r0 + (size_t) u * rpw + sub;
            if (r < n) {
                for (int c = l; c < pieces; c += lpr) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(rows + r * pitch) + c);
                    if (kSigned) {
                        acc[u] = __dp4a((int) v.x, (int) v.x, acc[u]);
                        acc[u] = __dp4a((int) v.y, (int) v.y, acc[u]);
                        acc[u] = __dp4a((int) v.z, (int) v.z, acc[u]);
                        acc[u] = __dp4a((int) v.w, (int) v.w, acc[u]);
                    } else {
                        acc[u] = (int) __dp4a(v.x, v.x, (unsigned) acc[u]);
                        acc[u] = (int) __dp4a(v.y, v.y, (unsigned) acc[u]);
                        acc[u] = (int) __dp4a(v.z, v.z, (unsigned) acc[u]);
                        acc[u] = (int) __dp4a(v.w, v.w, (unsigned) acc[u]);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            int a = acc[u];
            for (int o = lpr >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            const size_t r = r0 + (size_t) u * rpw + sub;
            if (l == 0 && r < n) {
                norms[r] = (uint32_t) a;
                if (index) index[r] = (uint32_t) r;
            }
        }
    }
}

// Gather rows into norm order, padded to 128-byte rows; rows >= n are zero padding with id = kInvalidId.
// One warp per output row (32 lanes x 4 bytes).
__global__ void gather_sorted_rows_kernel(const unsigned char* rows, size_t n, int d, size_t pitch, const uint32_t* perm,
                                          const uint32_t* sorted_norms, uint32_t id_base, size_t n_pad,
                                          unsigned char* out_rows /*[n_pad,128]*/, int32_t* out_norms, uint32_t* out_ids) {
    const int lane = threadIdx.x & 31;
    const size_t warps = ((size_t) gridDim.x * blockDim.x) >> 5;
    const bool words_ok = (d & 3) == 0 && (pitch & 3) == 0 && (reinterpret_cast<uintptr_t>(rows) & 3) == 0;
    for (size_t r = ((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_pad; r += warps) {
        uint32_t word = 0;
        if (r < n) {
            const unsigned char* src = rows + (size_t) perm[r] * pitch;
            if (words_ok) {  // 4-byte aligned rows: one coalesced 128-byte read per row
                if (lane * 4 < d) word = __ldg(reinterpret_cast<const uint32_t*>(src) + lane);
            } else {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int j = lane * 4 + b;
                    if (j < d) word |= (uint32_t) src[j] << (8 * b);
                }
            }
        }
        reinterpret_cast<uint32_t*>(out_rows + r * 128)[lane] = word;
        if (lane == 0) {
            out_norms[r] = r < n ? (int32_t) sorted_norms[r] : 0x3fffffff;
            out_ids[r] = r < n ? id_base + perm[r] : kInvalidId;
        }
    }
}

// Queries for the byte tensor kernel: float [nq,d] -> byte rows [nq_pad,128] + int32 norms; clears *ok when a
// coordinate is not an integer in the storage's range (the caller then takes the exact float path instead).
__global__ void prep_queries_bytes_kernel(const float* q, int nq, int d, int nq_pad, int is_signed, unsigned char* out_rows,
                                          int32_t* out_norms, int* ok, uint32_t* bound) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nq_pad; r += warps) {
        uint32_t word = 0;
        int nrm = 0;
        bool good = true;
        if (r < nq) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int j = lane * 4 + b;
                if (j < d) {
                    const float v = q[(size_t) r * d + j];
                    const float lo = is_signed ? -128.f : 0.f, hi = is_signed ? 127.f : 255.f;
                    good &= (v == rintf(v)) && v >= lo && v <= hi;
                    const int iv = (int) v;
                    nrm += iv * iv;
                    word |= (uint32_t) (iv & 0xFF) << (8 * b);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
        reinterpret_cast<uint32_t*>(out_rows + (size_t) r * 128)[lane] = word;
        if (lane == 0) {
            out_norms[r] = nrm;
            if (bound) bound[r] = 0x7f800000u;  // +inf: no bound yet
        }
        if (!__all_sync(0xffffffffu, good) && lane == 0) atomicAnd(ok, 0);
    }
}

// ---- float tensor layout (bf16 filter operands + fp32 norms), see topk_tc.cu --------------------------------------
// Squared norm of float rows in fp32 (plain sequential sum; the filter's slack covers its rounding), emitted as
// the float's bit pattern: norms are >= 0, so the unsigned bit order is the numeric order (radix-sort key).
__global__ void float_norms_kernel(const float* rows, size_t n, int d, uint32_t* norm_bits, uint32_t* index) {
    for (size_t r = (size_t) blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (size_t) gridDim.x * blockDim.x) {
        const float* row = rows + r * d;
        float acc = 0.f;
        for (int j = 0; j < d; ++j) acc = __fmaf_rn(row[j], row[j], acc);
        norm_bits[r] = __float_as_uint(acc);
        index[r] = (uint32_t) r;
    }
}

__device__ __forceinline__ unsigned short f32_to_bf16_rn(float f) {  // round to nearest even
    uint32_t u = __float_as_uint(f);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (unsigned short) (u >> 16);
}

// Gather float rows into norm order as bf16 rows of dp elements (zero padded); rows >= n are zero padding with
// id = kInvalidId and a huge norm.  One warp per output row.
__global__ void gather_sorted_rows_bf16_kernel(const float* rows, size_t n, int d, int dp, const uint32_t* perm,
                                               const uint32_t* sorted_norm_bits, uint32_t id_base, size_t n_pad,
                                               unsigned short* out_rows, float* out_norms, uint32_t* out_ids) {
    const int lane = threadIdx.x & 31;
    const size_t warps = ((size_t) gridDim.x * blockDim.x) >> 5;
    for (size_t r = ((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_pad; r += warps) {
        const float* src = r < n ? rows + (size_t) perm[r] * d : nullptr;
        for (int j = lane; j < dp; j += 32) out_rows[r * dp + j] = (src && j < d) ? f32_to_bf16_rn(src[j]) : (unsigned short) 0;
        if (lane == 0) {
            out_norms[r] = r < n ? __uint_as_float(sorted_norm_bits[r]) : 3.0e38f;
            out_ids[r] = r < n ? id_base + perm[r] : kInvalidId;
        }
    }
}

// Queries for the bf16 tensor kernel: float [nq,d] -> bf16 rows [nq_pad,dp], fp32 norms, a zero-padded fp32 copy
// for the exact re-rank, and the warm-start bounds reset to +inf.
__global__ void prep_queries_bf16_kernel(const float* q, int nq, int d, int dp, int nq_pad, unsigned short* out_rows,
                                         float* out_norms, float* out_q_f32, uint32_t* bound) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nq_pad; r += warps) {
        float nrm = 0.f;
        for (int j = lane; j < dp; j += 32) {
            const float v = (r < nq && j < d) ? q[(size_t) r * d + j] : 0.f;
            out_rows[(size_t) r * dp + j] = f32_to_bf16_rn(v);
            if (j < d) out_q_f32[(size_t) r * d + j] = v;
            nrm = __fmaf_rn(v, v, nrm);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
        if (lane == 0) {
            out_norms[r] = nrm;
            if (bound) bound[r] = 0x7f800000u;
        }
    }
}

void launch_detect_bytes(const float* x, size_t count, int* flags, int blocks, cudaStream_t st) {
    detect_bytes_kernel<<<blocks, 256, 0, st>>>(x, count, flags);
}
void launch_narrow_rows(const float* x, size_t n, int d, int pitch, unsigned char* out, int blocks, cudaStream_t st) {
    narrow_rows_kernel<<<blocks, 256, 0, st>>>(x, n, d, pitch, out);
}
void launch_byte_norms(const unsigned char* rows, size_t n, int d, size_t pitch, int is_signed, uint32_t* norms, uint32_t* index,
                       int num_sms, cudaStream_t st) {
    const bool vec = n > 0 && (d & 15) == 0 && (pitch & 15) == 0 && (reinterpret_cast<uintptr_t>(rows) & 15) == 0 && d <= 512;
    if (vec) {
        int lpr = 1;
        while (lpr < (d >> 4)) lpr <<= 1;
        const size_t rows_per_block = (size_t) 8 * (32 / lpr) * 4;
        const unsigned blocks = (unsigned) std::max<size_t>(1, std::min<size_t>((n + rows_per_block - 1) / rows_per_block, (size_t) num_sms * 16));
        if (is_signed) byte_norms_vec_kernel<true><<<blocks, 256, 0, st>>>(rows, n, d, pitch, lpr, norms, index);
        else byte_norms_vec_kernel<false><<<blocks, 256, 0, st>>>(rows, n, d, pitch, lpr, norms, index);
    } else {
        const unsigned blocks = (unsigned) std::max<size_t>(1, std::min<size_t>((n + 255) / 256, (size_t) num_sms * 16));
        byte_norms_kernel<<<blocks, 256, 0, st>>>(rows, n, d, pitch, is_signed, norms, index);
    }
}
void launch_gather_sorted_rows(const unsigned char* rows, size_t n, int d, size_t pitch, const uint32_t* perm,
                               const uint32_t* sorted_norms, uint32_t id_base, size_t n_pad, unsigned char* out_rows,
                               int32_t* out_norms, uint32_t* out_ids, int blocks, cudaStream_t st) {
    gather_sorted_rows_kernel<<<blocks, 256, 0, st>>>(rows, n, d, pitch, perm, sorted_norms, id_base, n_pad, out_rows, out_norms,
                                                      out_ids);
}
void launch_prep_queries_bytes(const float* q, int nq, int d, int nq_pad, int is_signed, unsigned char* out_rows, int32_t* out_norms,
                               int* ok, uint32_t* bound, cudaStream_t st) {
    const unsigned blocks = (unsigned) ((((size_t) nq_pad * 32 + 255) / 256) < 1024 ? (((size_t) nq_pad * 32 + 255) / 256) : 1024);
    prep_queries_bytes_kernel<<<blocks, 256, 0, st>>>(q, nq, d, nq_pad, is_signed, out_rows, out_norms, ok, bound);
}

void launch_float_norms(const float* rows, size_t n, int d, uint32_t* norm_bits, uint32_t* index, int blocks, cudaStream_t st) {
    float_norms_kernel<<<blocks, 256, 0, st>>>(rows, n, d, norm_bits, index);
}
void launch_gather_sorted_rows_bf16(const float* rows, size_t n, int d, int dp, const uint32_t* perm, const uint32_t* sorted_norm_bits,
                                    uint32_t id_base, size_t n_pad, unsigned short* out_rows, float* out_norms, uint32_t* out_ids,
                                    int blocks, cudaStream_t st) {
    gather_sorted_rows_bf16_kernel<<<blocks, 256, 0, st>>>(rows, n, d, dp, perm, sorted_norm_bits, id_base, n_pad, out_rows, out_norms,
                                                           out_ids);
}
void launch_prep_queries_bf16(const float* q, int nq, int d, int dp, int nq_pad, unsigned short* out_rows, float* out_norms,
                              float* out_q_f32, uint32_t* bound, cudaStream_t st) {
    const size_t want = ((size_t) nq_pad * 32 + 255) / 256;
    prep_queries_bf16_kernel<<<(unsigned) (want < 1024 ? want : 1024), 256, 0, st>>>(q, nq, d, dp, nq_pad, out_rows, out_norms, out_q_f32,
                                                                                     bound);
}

}  // namespace gpann
