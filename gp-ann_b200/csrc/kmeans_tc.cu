// kmeans_tc.cu -- tcgen05 / TMEM / TMA nearest-centroid assignment for float point sets (the k-means form of the hot
// path: reference NearestCenters / Top1Neighbor, src/kmeans.cpp:12-27, called 20x per KMeans, src/kmeans.cpp:238).
//
// argmin_j |x - c_j|^2 = argmax_j (x.c_j - 0.5*|c_j|^2).  The n x kc x d contraction runs on the tensor cores in SPLIT
// PRECISION: x = xh + xl, c = ch + cl with bf16 hi/lo parts (tcgen05.mma.kind::f16, fp32 accumulation in TMEM),
//     x.c ~= xh.ch + xh.cl + xl.ch            (three accumulating MMA chains per tile)
// with the rigorous bound |dot~ - x.c| <= eps*|x|*|c|, eps = 2^-16 (dropped terms) + 2^-13 (fp32 accumulation).
// The tensor result is only a FILTER: the epilogue finds, per point, the three best scores with exact knowledge of what
// it did not keep, and the finalize kernel accepts the winner only when no other centroid can be within the error
// bound; every other point is re-ranked EXACTLY (fp32, the reference's own summation order, lowest index on ties),
// so assignments equal the CPU code's.
//
// Layout (built once per point set): rows [n_pad, 2*dpad] bf16 = hi block then lo block, dpad = ceil(d/32)*32, fetched
// by TMA as 64-byte K blocks (SWIZZLE_64B); centroids are split the same way every round, plus hn'[j] = 0.5|c_j|^2 -
// offset, where offset makes every score a non-negative float (so scores compare as integers).
//
// CTA (320 threads, 1 per SM, persistent): blockIdx -> (centroid split cs, point-tile group).  The CTA's NB centroids
// (128 or 256: the UMMA N) stay resident in shared memory; point tiles (128 rows: UMMA M, TMEM lanes) stream through a
// TMA ring; TMEM holds two accumulators so the MMA of tile i+1 overlaps the drain of tile i.
//   warp 0 lane 0 : TMA producer       warp 1 lane 0 : MMA issuer (3 x ceil(d/16) tcgen05.mma per tile)
//   warps 2..9    : filter -- warp (w&3) owns TMEM lanes 32*(w&3).., column half (w-2)/4.  Per 32-column chunk:
//                   score = acc - hn' (FADD), low 5 mantissa bits := column (LOP3), max tree (VIMNMX3) -> c1,
//                   exact second maximum c2 via min over (c1 - s - 1) as unsigned; (c1,c2) go into the lane's
//                   top-3.  Halves merge through shared memory; one uint4 per point and centroid split goes to HBM.
#include "common.cuh"
#include "params.h"
#include "tc_ptx.cuh"

namespace gpann {
namespace kmtc {

using namespace ptx;

// ---- layout / per-round preparation ---------------------------------------------------------------------------------
__device__ __forceinline__ unsigned short bf16_rn(float f) {  // round to nearest even
    uint32_t u = __float_as_uint(f);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (unsigned short) (u >> 16);
}
__device__ __forceinline__ float bf16_to_f32(unsigned short h) { return __uint_as_float((uint32_t) h << 16); }

// One warp per row: hi/lo bf16 split into [hi(dpad) | lo(dpad)], |row| (rounded up), optionally 0.5*|row|^2, and the
// maxima of both over all rows (atomicMax on the bits of non-negative floats).
__global__ void split_rows_bf16_kernel(const float* rows, size_t n, int d, int dpad, size_t n_pad, unsigned short* out_rows,
                                       float* out_norm, float* out_hn, float* max_norm, float* max_hn) {
    const int lane = threadIdx.x & 31;
    const size_t warps = ((size_t) gridDim.x * blockDim.x) >> 5;
    float wmax_norm = 0.f, wmax_hn = 0.f;
    for (size_t r = ((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_pad; r += warps) {
        double n2 = 0.0;
        unsigned short* dst = out_rows + r * (size_t) (2 * dpad);
        for (int j = lane; j < dpad; j += 32) {
            const float v = (r < n && j < d) ? rows[r * (size_t) d + j] : 0.f;
            const unsigned short hi = bf16_rn(v);
            const unsigned short lo = bf16_rn(v - bf16_to_f32(hi));
            dst[j] = hi;
            dst[dpad + j] = lo;
            n2 += (double) v * (double) v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
        const float nrm = (float) sqrt(n2) * 1.000001f;
        const float hn = (float) (0.5 * n2);
        if (lane == 0) {
            if (out_norm) out_norm[r] = nrm;
            if (out_hn) out_hn[r] = hn;
        }
        wmax_norm = fmaxf(wmax_norm, nrm);
        wmax_hn = fmaxf(wmax_hn, hn * 1.000001f);
    }
    if (lane == 0) {
        if (max_norm) atomicMax(reinterpret_cast<int*>(max_norm), __float_as_int(wmax_norm));
        if (max_hn) atomicMax(reinterpret_cast<int*>(max_hn), __float_as_int(wmax_hn));
    }
}

// offset >= max |dot~| + max hn, so that score = dot~ - hn + offset >= 0 for every (point, centroid); hn'[j] = hn[j] -
// offset for real centroids and 0 for padding ones (their accumulators are exactly 0 -> score +0, the minimum).
// Removed centroids (alive[j] == 0: all-zero rows, see kmeans_update_kernel) are treated like padding.
__global__ void kmeans_centroid_consts_kernel(const float* xmax, const float* cmax, const float* hnmax, float* hn, int kc, int kc_pad,
                                              const unsigned char* alive, Consts* consts) {
    const float xm = *xmax, cm = *cmax;
    const float offset = 1.02f * xm * cm + *hnmax + 1e-30f;
    for (int j = threadIdx.x; j < kc_pad; j += blockDim.x) hn[j] = (j < kc && (!alive || alive[j])) ? hn[j] - offset : 0.f;
    if (threadIdx.x == 0) {
        consts->offset = offset;
        consts->cmax = cm;
        consts->xmax = xm;
        consts->eps = 1.52587890625e-05f + 1.220703125e-04f;  // 2^-16 + 2^-13
    }
}

// ---- filter arithmetic ----------------------------------------------------------------------------------------------
__device__ __forceinline__ int imax3(int a, int b, int c) { return max(max(a, b), c); }
__device__ __forceinline__ uint32_t umin3(uint32_t a, uint32_t b, uint32_t c) { return min(min(a, b), c); }

struct Top3 {
    int e1, e2, e3;   // packed scores, e1 >= e2 >= e3
    uint32_t ids;     // chunk ids: bits 0-3 of e1's chunk, 4-7 e2's, 8-11 e3's
};
__device__ __forceinline__ void top3_insert(Top3& t, int v, uint32_t cid) {
    if (v > t.e3) {
        if (v > t.e2) {
            t.e3 = t.e2;
            if (v > t.e1) {
                t.e2 = t.e1;
                t.e1 = v;
                t.ids = ((t.ids << 4) & 0xFF0u) | cid;
            } else {
                t.e2 = v;
                t.ids = (t.ids & 0xFu) | (cid << 4) | ((t.ids << 4) & 0xF00u);
            }
        } else {
            t.e3 = v;
            t.ids = (t.ids & 0xFFu) | (cid << 8);
        }
    }
}

// One 32-column chunk of accumulators -> (c1, c2): the two largest packed scores of the chunk, exactly.
// (x & mask) | j in ONE LOP3: the mask comes in a register (a kernel argument), the column index is the immediate.  Written
// with both as literals the compiler emits two LOP3 per score (ncu, round 2: 149 of the 530 instructions per warp and tile).
template <int J>
__device__ __forceinline__ int pack_col(float score, uint32_t mask) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xEC;" : "=r"(r) : "r"(__float_as_uint(score)), "n"(J), "r"(mask));  // (a & c) | b
    return (int) r;
}

__device__ __forceinline__ void chunk_top2(const uint32_t (&v)[32], const float* __restrict__ hn, uint32_t mask, int& c1, int& c2) {
    int s[32];
    {
        const float4* h4 = reinterpret_cast<const float4*>(hn);
#define GPANN_PACK4(J)                                                   \
        {                                                                \
            const float4 h = h4[(J) / 4];                                \
            s[(J) + 0] = pack_col<(J) + 0>(__uint_as_float(v[(J) + 0]) - h.x, mask); \
            s[(J) + 1] = pack_col<(J) + 1>(__uint_as_float(v[(J) + 1]) - h.y, mask); \
            s[(J) + 2] = pack_col<(J) + 2>(__uint_as_float(v[(J) + 2]) - h.z, mask); \
            s[(J) + 3] = pack_col<(J) + 3>(__uint_as_float(v[(J) + 3]) - h.w, mask); \
        }
        GPANN_PACK4(0) GPANN_PACK4(4) GPANN_PACK4(8) GPANN_PACK4(12) GPANN_PACK4(16) GPANN_PACK4(20) GPANN_PACK4(24) GPANN_PACK4(28)
#undef GPANN_PACK4
    }
    int a[11];
#pragma unroll
    for (int i = 0; i < 10; ++i) a[i] = imax3(s[3 * i], s[3 * i + 1], s[3 * i + 2]);
    a[10] = max(s[30], s[31]);
    c1 = max(imax3(imax3(a[0], a[1], a[2]), imax3(a[3], a[4], a[5]), imax3(a[6], a[7], a[8])), max(a[9], a[10]));
    // second maximum: the winner maps to 0xFFFFFFFF, every other value to c1 - s - 1 >= 0 (values are distinct)
    uint32_t u[11];
    const int c1m1 = c1 - 1;
#pragma unroll
    for (int i = 0; i < 10; ++i)
        u[i] = umin3((uint32_t) (c1m1 - s[3 * i]), (uint32_t) (c1m1 - s[3 * i + 1]), (uint32_t) (c1m1 - s[3 * i + 2]));
    u[10] = min((uint32_t) (c1m1 - s[30]), (uint32_t) (c1m1 - s[31]));
    const uint32_t um = min(umin3(umin3(u[0], u[1], u[2]), umin3(u[3], u[4], u[5]), umin3(u[6], u[7], u[8])), min(u[9], u[10]));
    c2 = c1m1 - (int) um;
}

__global__ void __launch_bounds__(kThreads, 1)
kmeans_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const Args a) {
    extern __shared__ unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NB = a.NB, NKB2 = 2 * a.nkb, NST = a.stages;
    const int cs = blockIdx.x % a.CS, pg = blockIdx.x / a.CS, PG = gridDim.x / a.CS;
    const uint32_t b_block_bytes = (uint32_t) NB * 64, a_tile_bytes = (uint32_t) NKB2 * kABlockBytes;

    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
    unsigned char* base = smem_raw + pad;
    const uint32_t sB = raw_addr + pad;
    const uint32_t sA = sB + (uint32_t) NKB2 * b_block_bytes;
    unsigned char* after = base + (size_t) NKB2 * b_block_bytes + (size_t) NST * a_tile_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(after);  // full[3], empty[3], bfull, tfull[2], tempty[2]
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kMaxStages, bar_b = bar_empty + 8 * kMaxStages,
                   bar_tfull = bar_b + 8, bar_tempty = bar_tfull + 16;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(after + 8 * (2 * kMaxStages + 5));
    float* s_hn = reinterpret_cast<float*>(after + 256);             // [NB]
    uint4* s_part = reinterpret_cast<uint4*>(s_hn + NB);             // [2][kFilterWarps/4 - 1][128]
    // column groups of 64 (or more) columns: NB = 256 with 16 filter warps -> 4 groups; warps beyond the last group idle
    constexpr int kMaxGroups = kFilterWarps / 4;
    const int ngroups = min(kMaxGroups, NB / 64);

    for (int i = tid; i < NB; i += kThreads) s_hn[i] = a.hnp[(size_t) cs * NB + i];
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        for (int i = 0; i < kMaxStages; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1);  // MMA commit: operands consumed
        }
        mbar_init(bar_b, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 4 * ngroups);  // one arrive per working filter warp
        }
        fence_barrier_init();
    }
    const uint32_t tmem_cols = (uint32_t) (2 * NB);
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer: this split's centroids once, then the point tiles =====
            mbar_expect_tx(bar_b, (uint32_t) NKB2 * b_block_bytes);
            for (int kb = 0; kb < NKB2; ++kb) tma_load_2d(sB + kb * b_block_bytes, &tmap_b, bar_b, kb * kKBlockElems, cs * NB);
            uint32_t st = 0, ph = 0;
            for (uint32_t t = pg; t < a.ntiles; t += PG) {
                mbar_wait(bar_empty + 8 * st, ph ^ 1);
                mbar_expect_tx(bar_full + 8 * st, a_tile_bytes);
                for (int kb = 0; kb < NKB2; ++kb)
                    tma_load_2d(sA + st * a_tile_bytes + kb * kABlockBytes, &tmap_a, bar_full + 8 * st, kb * kKBlockElems,
                                (int) (t * kTileM));
                if (++st == (uint32_t) NST) {
                    st = 0;
                    ph ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: acc = xh.cl + xl.ch + xh.ch (small terms first).  The whole warp walks the loop and one
        // elected lane issues a tile's MMAs + commits (see topk_tc.cu: a lane-0-only loop costs an ELECT/branch wrapper
        // and fresh descriptor arithmetic per tcgen05.mma and ends up issue-bound); descriptors are base + constant.
        mbar_wait(bar_b, 0);
        tcgen05_fence_after();
        const uint64_t a_base = make_sw64_desc(sA), b_base = make_sw64_desc(sB);
        const uint32_t idesc = a.idesc;
        const int nks = a.nks, nkb = a.nkb;
        uint32_t st = 0, ph = 0, slot = 0, sph = 0;
        for (uint32_t t = pg; t < a.ntiles; t += PG) {
            mbar_wait(bar_full + 8 * st, ph);
            mbar_wait(bar_tempty + 8 * slot, sph ^ 1);
            tcgen05_fence_after();
            if (elect_one()) {
                const uint32_t dcol = tmem_base + slot * NB;
                const uint64_t a_st = a_base + ((st * a_tile_bytes) >> 4);
                uint32_t acc = 0;
#pragma unroll
                for (int term = 0; term < 3; ++term) {
                    const int ta = term == 1 ? 1 : 0;  // A part: hi, lo, hi
                    const int tb = term == 0 ? 1 : 0;  // B part: lo, hi, hi
                    const uint64_t a_t = a_st + (uint32_t) ((ta * nkb * kABlockBytes) >> 4);
                    const uint64_t b_t = b_base + (uint32_t) ((tb * nkb * (int) b_block_bytes) >> 4);
                    for (int ks = 0; ks < nks; ++ks) {
                        // K step ks = 16 bf16 = 32 bytes: block ks/2 of the part, +2 (16-byte units) for the second half
                        const int kb = ks >> 1, w = ks & 1;
                        umma_bf16(dcol, a_t + (uint32_t) (kb * (kABlockBytes >> 4) + 2 * w), b_t + (uint32_t) (kb * (int) (b_block_bytes >> 4) + 2 * w),
                                  idesc, acc);
                        acc = 1;
                    }
                }
                umma_commit(bar_tfull + 8 * slot);
                umma_commit(bar_empty + 8 * st);
            }
            __syncwarp();
            if (++st == (uint32_t) NST) {
                st = 0;
                ph ^= 1;
            }
            slot ^= 1;
            sph ^= (slot == 0);
        }
    } else {
        // ===== filter warps =====
        const int ew = warp - 2;
        const int grp = ew >> 2;              // column group of this warp
        const int wq = warp & 3;
        const int qrow = wq * 32 + lane;
        const uint32_t lane_sel = (uint32_t) (wq * 32) << 16;
        const int ncol = NB / ngroups;        // columns of this warp: [grp*ncol, +ncol)
        const int col_base = grp * ncol;
        const uint32_t cid_base = (uint32_t) (col_base >> 5);
        const uint32_t pack_mask = a.pack_mask;
        uint32_t tl = 0;
        for (uint32_t t = pg; grp < ngroups && t < a.ntiles; t += PG, ++tl) {
            const uint32_t slot = tl & 1, sph = (tl >> 1) & 1;
            Top3 top;
            top.e1 = top.e2 = top.e3 = -1;
            top.ids = 0;
            mbar_wait(bar_tfull + 8 * slot, sph);
            tcgen05_fence_after();
            const uint32_t tbase = tmem_base + slot * NB + col_base + lane_sel;
            if (a.mode == 2) {
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * slot);
                continue;
            }
#pragma unroll 1
            for (int c = 0; c < ncol; c += 64) {
                uint32_t va[32], vb[32];
                tmem_ld32(tbase + c, va);
                tmem_ld32(tbase + c + 32, vb);
                tmem_wait_ld();
                if (c + 64 >= ncol) {  // accumulator drained: hand the slot back before the arithmetic
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8 * slot);
                }
                int c1, c2;
                chunk_top2(va, s_hn + col_base + c, pack_mask, c1, c2);
                uint32_t cid = cid_base + (uint32_t) (c >> 5);
                top3_insert(top, c1, cid);
                top3_insert(top, c2, cid);
                chunk_top2(vb, s_hn + col_base + c + 32, pack_mask, c1, c2);
                cid += 1;
                top3_insert(top, c1, cid);
                top3_insert(top, c2, cid);
            }
            // merge the column groups of a point through shared memory; group 0 writes the result
            uint4* part = s_part + (size_t) (tl & 1) * (kMaxGroups - 1) * kTileM;
            if (grp > 0) part[(grp - 1) * kTileM + qrow] = make_uint4((uint32_t) top.e1, (uint32_t) top.e2, (uint32_t) top.e3, top.ids);
            asm volatile("bar.sync %0, %1;" ::"r"(2 + wq), "r"(32 * ngroups) : "memory");
            if (grp == 0) {
                for (int og = 1; og < ngroups; ++og) {
                    const uint4 o = part[(og - 1) * kTileM + qrow];
                    top3_insert(top, (int) o.x, o.w & 0xFu);
                    top3_insert(top, (int) o.y, (o.w >> 4) & 0xFu);
                    top3_insert(top, (int) o.z, (o.w >> 8) & 0xFu);
                }
                a.partial[(size_t) cs * ((size_t) a.ntiles * kTileM) + (size_t) t * kTileM + qrow] =
                    make_uint4((uint32_t) top.e1, (uint32_t) top.e2, (uint32_t) top.e3, top.ids);
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ---- finalize: merge the centroid splits, accept unambiguous winners, list the rest ----------------------------------
struct Merged {
    int e[3];
    int j[3];  // centroid index
};

__device__ __forceinline__ void merged_insert(Merged& m, int v, int idx) {
    if (v > m.e[2]) {
        if (v > m.e[1]) {
            m.e[2] = m.e[1]; m.j[2] = m.j[1];
            if (v > m.e[0]) {
                m.e[1] = m.e[0]; m.j[1] = m.j[0];
                m.e[0] = v; m.j[0] = idx;
            } else {
                m.e[1] = v; m.j[1] = idx;
            }
        } else {
            m.e[2] = v; m.j[2] = idx;
        }
    }
}

constexpr uint32_t kModePair = 0u, kModeChunk = 1u, kModeFull = 2u;

__global__ void kmeans_finalize_kernel(const KmFinalizeParams p) {
    const Consts c = *p.consts;
    const int chunks_per_split = p.NB >> 5;
    // relative rounding error of the reference's fp32 distance (8 FMA lanes + tail, src/dist.cpp:9-16), with margin
    const float gamma_ref = (float) (p.d / 8 + 16) * 1.1920929e-07f;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (size_t) gridDim.x * blockDim.x) {
        Merged m;
        m.e[0] = m.e[1] = m.e[2] = -1;
        m.j[0] = m.j[1] = m.j[2] = 0x7fffffff;
        for (int cs = 0; cs < p.CS; ++cs) {
            const uint4 q = p.partial[(size_t) cs * p.n_pad + i];
            const int cb = cs * chunks_per_split;
            merged_insert(m, (int) q.x, ((cb + (int) (q.w & 0xFu)) << 5) | (int) (q.x & 31u));
            merged_insert(m, (int) q.y, ((cb + (int) ((q.w >> 4) & 0xFu)) << 5) | (int) (q.y & 31u));
            merged_insert(m, (int) q.z, ((cb + (int) ((q.w >> 8) & 0xFu)) << 5) | (int) (q.z & 31u));
        }
        const float xn = p.xnorm[i];
        const float s1 = __int_as_float(m.e[0]), s2 = __int_as_float(m.e[1]), s3 = __int_as_float(m.e[2]);
        // per-score error: split-precision dot + packing truncation (5 mantissa bits of a score <= 2*offset) + the
        // roundings of hn' and of the subtraction
        const float err = c.eps * xn * c.cmax + c.offset * (7.62939453125e-06f + 9.5367431640625e-07f);
        const float dist1 = fmaxf(0.f, xn * xn - 2.f * (s1 - c.offset)) + 4.f * err;
        const float slack = 2.f * err + gamma_ref * dist1;
        const float thr = s1 - slack;
        uint32_t mode;
        const bool bad0 = m.j[0] >= p.kc || (p.alive && !p.alive[m.j[0]]);  // padding or removed centroid
        const bool bad1 = m.j[1] >= p.kc || (p.alive && !p.alive[m.j[1]]);
        if (bad0) {
            mode = kModeFull;
        } else if (s2 < thr) {
            p.out_assign[i] = m.j[0];
            continue;
        } else if (s3 >= thr || bad1) {
            mode = kModeFull;
        } else if ((m.j[0] >> 5) == (m.j[1] >> 5)) {
            mode = kModeChunk;
        } else {
            mode = kModePair;
        }
        const uint32_t pos = atomicAdd(p.amb_count, 1u);
        p.amb_list[pos] = (uint32_t) i;
        p.amb_info[pos] = make_uint2((uint32_t) m.j[0], ((uint32_t) m.j[1] & 0x3FFFFFFFu) | (mode << 30));
    }
}

// Squared L2 distance of two global fp32 rows in the reference's summation order (src/dist.cpp:9-16 -> hnswlib
// space_l2.h:60-91, 6-20).
__device__ __forceinline__ float exact_l2(const float* __restrict__ x, const float* __restrict__ q, int d) {
    if ((d & 3) == 0 && (((uintptr_t) x | (uintptr_t) q) & 15) == 0) {
        // same arithmetic, 16-byte loads: lane l of the reference's 8 accumulators takes elements i = l (mod 8)
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f, l4 = 0.f, l5 = 0.f, l6 = 0.f, l7 = 0.f;
        const int d16 = d & ~15;
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
        const float res = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(l0, l1), l2), l3), l4), l5), l6), l7);
        float tail = 0.f;
        for (int i = d16; i < d; ++i) {
            const float t = x[i] - q[i];
            tail = __fmaf_rn(t, t, tail);
        }
        return __fadd_rn(res, tail);
    }
    return sqr_l2_ref_order([x](int i) { return x[i]; }, q, d);
}

// Exact re-rank of the ambiguous points: distances to the candidate centroids (a pair, one 32-centroid chunk, or all of
// them), minimum (distance, index) key = the reference's strict-< scan (src/kmeans.cpp:16-21).  A warp takes 16 list
// entries at a time: the common case, a pair of candidates, is settled by the entry's own two lanes (16 points per warp
// pass instead of one); chunk / full entries are then re-ranked by the whole warp, one after the other.
__global__ void kmeans_resolve_kernel(const KmFinalizeParams p) {
    const int lane = threadIdx.x & 31;
    const uint32_t total = *p.amb_count;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t groups = (total + 15u) / 16u;
    for (uint32_t gidx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; gidx < groups; gidx += warps) {
        const uint32_t w = gidx * 16u + (uint32_t) (lane >> 1);
        const bool valid = w < total;
        const uint32_t i = valid ? p.amb_list[w] : 0u;
        const uint2 raw = valid ? p.amb_info[w] : make_uint2(0u, 0u);
        const uint32_t mode = raw.y >> 30;
        uint64_t best = kEmptyKey;
        if (valid && mode == kModePair) {
            const int j = (lane & 1) == 0 ? (int) raw.x : (int) (raw.y & 0x3FFFFFFFu);
            best = pack_key(exact_l2(p.rows + (size_t) i * p.d, p.centroids + (size_t) j * p.d, p.d), (uint32_t) j);
        }
        const uint64_t other = __shfl_xor_sync(0xffffffffu, best, 1);
        best = other < best ? other : best;
        if (valid && mode == kModePair && (lane & 1) == 0) p.out_assign[i] = key_dist(best) < 3.402823466e+38f ? (int32_t) key_id(best) : -1;
        // entries that need more than a pair: one at a time, warp-wide
        unsigned pending = __ballot_sync(0xffffffffu, valid && mode != kModePair && (lane & 1) == 0);
        while (pending) {
            const int src = __ffs(pending) - 1;
            pending &= pending - 1;
            const uint32_t pi = __shfl_sync(0xffffffffu, i, src);
            const uint32_t px = __shfl_sync(0xffffffffu, raw.x, src);
            const uint32_t pm = __shfl_sync(0xffffffffu, mode, src);
            const float* x = p.rows + (size_t) pi * p.d;
            uint64_t b = kEmptyKey;
            if (pm == kModeChunk) {
                const int j = (int) (px & ~31u) + lane;
                if (j < p.kc && (!p.alive || p.alive[j])) b = pack_key(exact_l2(x, p.centroids + (size_t) j * p.d, p.d), (uint32_t) j);
            } else {
                for (int j = lane; j < p.kc; j += 32) {
                    if (p.alive && !p.alive[j]) continue;
                    const uint64_t key = pack_key(exact_l2(x, p.centroids + (size_t) j * p.d, p.d), (uint32_t) j);
                    b = key < b ? key : b;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const uint64_t ob = __shfl_xor_sync(0xffffffffu, b, o);
                b = ob < b ? ob : b;
            }
            if (lane == 0) p.out_assign[pi] = key_dist(b) < 3.402823466e+38f ? (int32_t) key_id(b) : -1;
        }
    }
}

}  // namespace kmtc

cudaError_t kmtc_set_attributes() {
    return cudaFuncSetAttribute(kmtc::kmeans_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}
void launch_split_rows_bf16(const float* rows, size_t n, int d, int dpad, size_t n_pad, unsigned short* out_rows, float* out_norm,
                            float* out_hn, float* max_norm, float* max_hn, int blocks, cudaStream_t st) {
    kmtc::split_rows_bf16_kernel<<<blocks, 256, 0, st>>>(rows, n, d, dpad, n_pad, out_rows, out_norm, out_hn, max_norm, max_hn);
}
void launch_kmeans_centroid_consts(const float* xmax, const float* cmax, const float* hnmax, float* hn, int kc, int kc_pad,
                                   const unsigned char* alive, kmtc::Consts* consts, cudaStream_t st) {
    kmtc::kmeans_centroid_consts_kernel<<<1, 256, 0, st>>>(xmax, cmax, hnmax, hn, kc, kc_pad, alive, consts);
}
void launch_kmeans_tc(const CUtensorMap& tmap_a, const CUtensorMap& tmap_b, const kmtc::Args& a, unsigned grid, size_t smem,
                      cudaStream_t st) {
    kmtc::kmeans_tc_kernel<<<grid, kmtc::kThreads, smem, st>>>(tmap_a, tmap_b, a);
}
void launch_kmeans_finalize(const KmFinalizeParams& p, int blocks, cudaStream_t st) {
    kmtc::kmeans_finalize_kernel<<<blocks, 256, 0, st>>>(p);
}
void launch_kmeans_resolve(const KmFinalizeParams& p, int blocks, cudaStream_t st) {
    kmtc::kmeans_resolve_kernel<<<blocks, 256, 0, st>>>(p);
}

}  // namespace gpann
