// kmeans.cu -- k-means form of the hot path, exact (reference arithmetic order) CUDA-core kernels.
//
//   kmeans_assign_kernel   NearestCenters / Top1Neighbor      (reference src/kmeans.cpp:12-27)
//   cluster_reduce_*       SumPointsInClustersL2 / Aggregate*  (src/kmeans.cpp:116-204), deterministic order
//   objective_kernel       ObjectiveValue                      (src/kmeans.cpp:246-249)
//
// The reference sums points per cluster sequentially in point order (and, for n > 19.66M at kc=1024,d=96,
// in racy CAS-float block merges, src/kmeans.cpp:179-191, i.e. nondeterministically).  Here the members of a
// cluster are brought together by a STABLE radix sort of (assignment, point index) pairs, cut into segments
// of kSegRows rows, summed by one CTA per segment in a fixed order, and the segment sums are added per
// cluster in segment order -- bitwise reproducible run to run and independent of the grid.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "distance.cuh"
#include "params.h"

namespace gpann {

template <int STORAGE>
__device__ __forceinline__ void stage_rows(const RowsView& v, size_t r0, int rows_here, unsigned char* s_rows, int spitch,
                                           int vec_ok, int tid, int nthreads) {
    const int elem = STORAGE == kF32 ? 4 : 1;
    const int row_bytes = v.d * elem;
    const char* src = v.rows + r0 * v.pitch;
    if (vec_ok) {
        const int cpr = row_bytes >> 4;
        for (int idx = tid; idx < rows_here * cpr; idx += nthreads) {
            const int r = idx / cpr, c = idx - r * cpr;
            const int4 val = __ldg(reinterpret_cast<const int4*>(src + (size_t) r * v.pitch) + c);
            *reinterpret_cast<int4*>(s_rows + (size_t) r * spitch + c * 16) = val;
        }
    } else {
        for (int idx = tid; idx < rows_here * row_bytes; idx += nthreads) {
            const int r = idx / row_bytes, c = idx - r * row_bytes;
            s_rows[(size_t) r * spitch + c] = (unsigned char) src[(size_t) r * v.pitch + c];
        }
    }
}

// One thread per point, centroids streamed through shared memory in chunks; strict '<' keeps the lowest
// centroid index on ties, -1 if nothing is below FLT_MAX (src/kmeans.cpp:12-23).
template <int STORAGE>
__global__ void __launch_bounds__(kStreamTile) kmeans_assign_kernel(const AssignParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    unsigned char* s_rows = smem_raw;
    const int trows = p.tile_rows;
    float* s_c = reinterpret_cast<float*>(s_rows + (size_t) trows * p.spitch);  // cchunk * qpitch
    const size_t ntiles = (p.pts.n + trows - 1) / trows;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t r0 = tile * trows;
        const int rows_here = (int) min((size_t) trows, p.pts.n - r0);
        __syncthreads();
        stage_rows<STORAGE>(p.pts, r0, rows_here, s_rows, p.spitch, p.vec_ok, tid, kStreamTile);
        int best = -1;
        float best_dist = 3.402823466e+38f;
        const char* myrow = reinterpret_cast<const char*>(s_rows) + (size_t) tid * p.spitch;
        for (int c0 = 0; c0 < p.kc; c0 += p.cchunk) {
            const int nc = min(p.cchunk, p.kc - c0);
            __syncthreads();
            for (int i = tid; i < nc * p.pts.d; i += kStreamTile) {
                const int c = i / p.pts.d, j = i - c * p.pts.d;
                s_c[c * p.qpitch + j] = p.centroids[(size_t) (c0 + c) * p.pts.d + j];
            }
            __syncthreads();
            if (tid < rows_here) {
                for (int c = 0; c < nc; ++c) {
                    if (p.alive && !p.alive[c0 + c]) continue;  // removed cluster
                    const float dist = ref_distance<STORAGE>(myrow, s_c + c * p.qpitch, p.pts.d, p.metric);
                    if (dist < best_dist) {
                        best_dist = dist;
                        best = c0 + c;
                    }
                }
            }
        }
        if (tid < rows_here) {
            const size_t row = r0 + tid;
            p.out_assign[row] = best;
            if (p.out_mindist) p.out_mindist[row] = best_dist;
        }
    }
}

// ---- deterministic per-cluster sums --------------------------------------------------------------------

// counts[c] = number of members; seg_begin[c] = first segment index of cluster c (exclusive scan of
// ceil(count/kSegRows)); seg_begin[kc] = total.  Single CTA, kc <= a few thousand.
__global__ void cluster_segments_kernel(const uint64_t* counts, int kc, uint32_t* seg_begin, uint64_t* member_begin) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        uint32_t s = 0;
        uint64_t m = 0;
        for (int c = 0; c < kc; ++c) {
            seg_begin[c] = s;
            member_begin[c] = m;
            s += (uint32_t) ((counts[c] + kSegRows - 1) / kSegRows);
            m += counts[c];
        }
        seg_begin[kc] = s;
        member_begin[kc] = m;
    }
}

// counts[c]++ for every point; also emits the radix-sort input: key = cluster (kc for unassigned -1), value = row.
// Counts go through a per-CTA shared-memory histogram when kc fits (100M global atomics on 1024 hot counters cost 7 ms
// of a 100 ms Lloyd round).
constexpr int kHistSmemClusters = 8192;
__global__ void histogram_keys_kernel(const int32_t* assign, size_t n, int kc, unsigned long long* counts, uint32_t* keys,
                                      uint32_t* index) {
    extern __shared__ uint32_t s_hist[];
    const bool in_smem = kc <= kHistSmemClusters;
    if (in_smem) {
        for (int c = threadIdx.x; c < kc; c += blockDim.x) s_hist[c] = 0u;
        __syncthreads();
    }
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        const int c = assign[i];
        const bool ok = c >= 0 && c < kc;
        if (ok) {
            if (in_smem) atomicAdd(&s_hist[c], 1u);
            else atomicAdd(&counts[c], 1ull);
        }
        keys[i] = ok ? (uint32_t) c : (uint32_t) kc;
        index[i] = (uint32_t) i;
    }
    if (in_smem) {
        __syncthreads();
        for (int c = threadIdx.x; c < kc; c += blockDim.x)
            if (s_hist[c]) atomicAdd(&counts[c], (unsigned long long) s_hist[c]);
    }
}

// One CTA per segment: thread t owns dimension j = t % dpad and row slot t / dpad; slots are added in order.
template <int STORAGE>
__global__ void __launch_bounds__(256) cluster_reduce_segments_kernel(const ReduceParams p) {
    __shared__ float s_part[256];
    const uint32_t seg = blockIdx.x;
    if (seg >= p.seg_begin[p.kc]) return;
    // binary search the cluster owning this segment
    int lo = 0, hi = p.kc;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (p.seg_begin[mid] <= seg) lo = mid; else hi = mid;
    }
    const int c = lo;
    const uint64_t cb = p.member_begin[c], ce = p.member_begin[c + 1];
    const uint64_t sb = cb + (uint64_t) (seg - p.seg_begin[c]) * kSegRows;
    const uint64_t se = min(ce, sb + kSegRows);
    const int d = p.pts.d;
    const int dchunk = min(256, ((d + 31) / 32) * 32);
    const int nslots = 256 / dchunk;
    const int j_in = threadIdx.x % dchunk, slot = threadIdx.x / dchunk;
    for (int j0 = 0; j0 < d; j0 += dchunk) {
        const int j = j0 + j_in;
        float acc = 0.f;
        if (j < d && slot < nslots) {
            for (uint64_t m = sb + slot; m < se; m += nslots) {
                const uint32_t r = p.members[m];
                const char* row = p.pts.rows + (size_t) r * p.pts.pitch;
                const float x = load1<STORAGE>(row, j);
                acc += p.row_norm ? x * (1.0f / p.row_norm[r]) : x;  // MIPS: P_i[j] * multiplier (src/kmeans.cpp:96-99)
            }
        }
        s_part[threadIdx.x] = acc;
        __syncthreads();
        if (slot == 0 && j < d) {
            float tot = s_part[j_in];
            for (int s = 1; s < nslots; ++s) tot += s_part[s * dchunk + j_in];
            p.seg_sums[(size_t) seg * d + j] = tot;
        }
        __syncthreads();
    }
}

// float rows with d % 4 == 0 (16-byte aligned): d/4 lanes fetch one row as float4, 256/(d/4) row slots per CTA (10 rows
// in flight for d = 96 instead of 2 scalar ones).  Slot s adds the rows sb+s, sb+s+nslots, ... in order and the slots
// are added in slot order, so the sums stay deterministic.
__global__ void __launch_bounds__(256) cluster_reduce_segments_vec4_kernel(const ReduceParams p) {
    __shared__ float4 s_part[256];
    const uint32_t seg = blockIdx.x;
    if (seg >= p.seg_begin[p.kc]) return;
    int lo = 0, hi = p.kc;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (p.seg_begin[mid] <= seg) lo = mid; else hi = mid;
    }
    const int c = lo;
    const uint64_t cb = p.member_begin[c], ce = p.member_begin[c + 1];
    const uint64_t sb = cb + (uint64_t) (seg - p.seg_begin[c]) * kSegRows;
    const uint64_t se = min(ce, sb + kSegRows);
    const int d4 = p.pts.d >> 2;            // float4 lanes per row (<= 64: d <= 256)
    const int nslots = 256 / d4;
    const int j = threadIdx.x % d4, slot = threadIdx.x / d4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (slot < nslots) {
        uint64_t m = sb + slot;
        // two independent rows per iteration keep more loads in flight; the adds stay in row order
        for (; m + nslots < se; m += 2 * (uint64_t) nslots) {
            const uint32_t r0 = p.members[m], r1 = p.members[m + nslots];
            float4 x0 = __ldg(reinterpret_cast<const float4*>(p.pts.rows + (size_t) r0 * p.pts.pitch) + j);
            float4 x1 = __ldg(reinterpret_cast<const float4*>(p.pts.rows + (size_t) r1 * p.pts.pitch) + j);
            if (p.row_norm) {
                const float m0 = 1.0f / p.row_norm[r0], m1 = 1.0f / p.row_norm[r1];
                x0.x *= m0; x0.y *= m0; x0.z *= m0; x0.w *= m0;
                x1.x *= m1; x1.y *= m1; x1.z *= m1; x1.w *= m1;
            }
            acc.x += x0.x; acc.y += x0.y; acc.z += x0.z; acc.w += x0.w;
            acc.x += x1.x; acc.y += x1.y; acc.z += x1.z; acc.w += x1.w;
        }
        if (m < se) {
            const uint32_t r0 = p.members[m];
            float4 x0 = __ldg(reinterpret_cast<const float4*>(p.pts.rows + (size_t) r0 * p.pts.pitch) + j);
            if (p.row_norm) {
                const float m0 = 1.0f / p.row_norm[r0];
                x0.x *= m0; x0.y *= m0; x0.z *= m0; x0.w *= m0;
            }
            acc.x += x0.x; acc.y += x0.y; acc.z += x0.z; acc.w += x0.w;
        }
    }
    s_part[threadIdx.x] = acc;
    __syncthreads();
    if (slot == 0) {
        float4 tot = s_part[j];
        for (int s = 1; s < nslots; ++s) {
            const float4 o = s_part[s * d4 + j];
            tot.x += o.x; tot.y += o.y; tot.z += o.z; tot.w += o.w;
        }
        reinterpret_cast<float4*>(p.seg_sums + (size_t) seg * p.pts.d)[j] = tot;
    }
}

// MIPS build: sum of |x_i|^2 over a segment's members (norm_sums of src/kmeans.cpp:95), fixed-order tree per segment.
__global__ void __launch_bounds__(256) segment_norm_sums_kernel(const ReduceParams p) {
    __shared__ float s_red[256];
    const uint32_t seg = blockIdx.x;
    if (seg >= p.seg_begin[p.kc]) return;
    int lo = 0, hi = p.kc;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (p.seg_begin[mid] <= seg) lo = mid; else hi = mid;
    }
    const uint64_t cb = p.member_begin[lo], ce = p.member_begin[lo + 1];
    const uint64_t sb = cb + (uint64_t) (seg - p.seg_begin[lo]) * kSegRows;
    const uint64_t se = min(ce, sb + kSegRows);
    float acc = 0.f;
    for (uint64_t m = sb + threadIdx.x; m < se; m += 256) {
        const float nr = p.row_norm[p.members[m]];
        acc += nr * nr;
    }
    s_red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int) threadIdx.x < s) s_red[threadIdx.x] += s_red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) p.seg_norm_sums[seg] = s_red[0];
}

// |x_i| = sqrt(vec_norm(P_i)) (src/kmeans.cpp:235, src/dist.cpp:54-58: sequential sum of squares), one thread per row
template <int STORAGE>
__global__ void row_norms_kernel(const RowsView v, float* out_norm) {
    for (size_t r = (size_t) blockIdx.x * blockDim.x + threadIdx.x; r < v.n; r += (size_t) gridDim.x * blockDim.x) {
        const char* row = v.rows + r * v.pitch;
        float acc = 0.f;
        for (int j = 0; j < v.d; ++j) {
            const float x = load1<STORAGE>(row, j);
            acc = __fmaf_rn(x, x, acc);
        }
        out_norm[r] = sqrtf(acc);
    }
}

// NormalizeCentroidsIP (src/kmeans.cpp:61-76): C = sums * sqrt((norm_sums / size) / |sums|^2); clusters without members
// are removed in place as in kmeans_update_kernel.
__global__ void kmeans_update_ip_kernel(float* centroids, const float* sums, const uint64_t* counts, const float* norm_sums,
                                        unsigned char* alive, int kc, int d) {
    __shared__ float s_red[128];
    const int c = blockIdx.x;
    const uint64_t cnt = counts[c];
    const bool live = alive[c] != 0 && cnt != 0;
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
    const float current = s_red[0];
    const float desired = live ? norm_sums[c] / (float) cnt : 0.f;
    const float mult = live ? sqrtf(desired / current) : 0.f;
    for (int j = threadIdx.x; j < d; j += blockDim.x) centroids[(size_t) c * d + j] = live ? sums[(size_t) c * d + j] * mult : 0.f;
    if (threadIdx.x == 0 && !live) alive[c] = 0;
}

// Per cluster: add its segment sums in segment order; optionally divide by the member count
// (NormalizeCentroidsL2, src/kmeans.cpp:104-114).  Empty clusters give all-zero rows.
__global__ void cluster_finalize_kernel(const float* seg_sums, const uint32_t* seg_begin, const uint64_t* counts, int kc,
                                        int d, int normalize, float* out /*[kc,d]*/, const float* seg_norm_sums, float* out_norm_sums) {
    const int c = blockIdx.x;
    if (seg_norm_sums && threadIdx.x == 0) {
        float tot = 0.f;
        for (uint32_t s = seg_begin[c]; s < seg_begin[c + 1]; ++s) tot += seg_norm_sums[s];
        out_norm_sums[c] = tot;
    }
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
        float tot = 0.f;
        for (uint32_t s = seg_begin[c]; s < seg_begin[c + 1]; ++s) tot += seg_sums[(size_t) s * d + j];
        if (normalize && counts[c] != 0) tot = tot / (float) counts[c];
        out[(size_t) c * d + j] = tot;
    }
}

// Divide summed coordinates by counts (after a cross-GPU all-reduce of sums and counts).
__global__ void normalize_centroids_kernel(float* sums, const uint64_t* counts, int kc, int d) {
    const int c = blockIdx.x;
    if (counts[c] == 0) return;
    for (int j = threadIdx.x; j < d; j += blockDim.x) sums[(size_t) c * d + j] = sums[(size_t) c * d + j] / (float) counts[c];
}

// ---- gpann_kmeans_run: Lloyd rounds without a host round trip --------------------------------------------------------
// centroids[c] = sums[c] / counts[c] (NormalizeCentroidsL2, src/kmeans.cpp:104-114); a cluster without members is REMOVED
// as RemoveEmptyClusters does (src/kmeans.cpp:29-50), but in place: alive[c] = 0 and its row zeroed.  Later rounds skip it,
// which selects the same winner as the reference's renumbered scan (compaction keeps the order of the survivors, so
// "lowest index wins ties" is unchanged); kmeans_compact_kernel renumbers once, at the end.
__global__ void kmeans_update_kernel(float* centroids, const float* sums, const uint64_t* counts, unsigned char* alive, int kc, int d) {
    const int c = blockIdx.x;
    const uint64_t cnt = counts[c];
    const bool live = alive[c] != 0 && cnt != 0;
    for (int j = threadIdx.x; j < d; j += blockDim.x)
        centroids[(size_t) c * d + j] = live ? sums[(size_t) c * d + j] / (float) cnt : 0.f;
    if (threadIdx.x == 0 && !live) alive[c] = 0;
}

// remap[c] = new index of a surviving cluster (-1 for removed ones), *kc_out = survivors; centroid rows and counts are
// compacted to the front (through cent_tmp [kc,d]).  One CTA.
__global__ void kmeans_compact_kernel(float* centroids, uint64_t* counts, const unsigned char* alive, int kc, int d, int32_t* remap,
                                      int* kc_out, float* cent_tmp) {
    __shared__ int s_total;
    if (threadIdx.x == 0) {
        int l = 0;
        for (int c = 0; c < kc; ++c) remap[c] = alive[c] ? l++ : -1;
        s_total = l;
        *kc_out = l;
    }
    __syncthreads();
    if (s_total == kc) return;
    for (size_t i = threadIdx.x; i < (size_t) kc * d; i += blockDim.x) cent_tmp[i] = centroids[i];
    __syncthreads();
    for (size_t i = threadIdx.x; i < (size_t) kc * d; i += blockDim.x) {
        const int c = (int) (i / d), j = (int) (i - (size_t) c * d);
        if (remap[c] >= 0) centroids[(size_t) remap[c] * d + j] = cent_tmp[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int c = 0; c < kc; ++c)
            if (remap[c] >= 0) counts[remap[c]] = counts[c];  // remap[c] <= c: in-place forward copy is safe
    }
}

__global__ void remap_assign_kernel(int32_t* assign, size_t n, const int32_t* remap, int kc) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        const int a = assign[i];
        if (a >= 0 && a < kc) assign[i] = remap[a];
    }
}

// sum_i pos_distance(P_i, C_{a_i}) with every f32 term widened to double (src/kmeans.cpp:246-249); fixed-order
// two-level reduction (per-CTA partials, summed in CTA order by objective_final_kernel).
template <int STORAGE>
__global__ void __launch_bounds__(kStreamTile) objective_kernel(const ObjectiveParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_red[kStreamTile];
    const int tid = threadIdx.x;
    unsigned char* s_rows = smem_raw;
    double acc = 0.0;
    const int trows = p.tile_rows;
    const size_t ntiles = (p.pts.n + trows - 1) / trows;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t r0 = tile * trows;
        const int rows_here = (int) min((size_t) trows, p.pts.n - r0);
        __syncthreads();
        stage_rows<STORAGE>(p.pts, r0, rows_here, s_rows, p.spitch, p.vec_ok, tid, kStreamTile);
        __syncthreads();
        if (tid < rows_here) {
            const size_t row = r0 + tid;
            {
                const int c = p.assign[row];
                const char* myrow = reinterpret_cast<const char*>(s_rows) + (size_t) tid * p.spitch;
                const float* cen = p.centroids + (size_t) c * p.pts.d;
                float dist;
                if (p.metric == kL2) {
                    dist = sqr_l2_ref_order([&](int i) { return load1<STORAGE>(myrow, i); }, cen, p.pts.d);
                } else {
                    const float ip = inner_product_ref_order([&](int i) { return load1<STORAGE>(myrow, i); }, cen, p.pts.d);
                    dist = __fadd_rn(__fsub_rn(1.0f, ip), 1.0f);  // pos_distance under MIPS (src/dist.cpp:83-88)
                }
                acc += (double) dist;
            }
        }
    }
    s_red[tid] = acc;
    __syncthreads();
    for (int s = kStreamTile / 2; s > 0; s >>= 1) {
        if (tid < s) s_red[tid] += s_red[tid + s];
        __syncthreads();
    }
    if (tid == 0) p.block_sums[blockIdx.x] = s_red[0];
}

__global__ void objective_final_kernel(const double* block_sums, int nblocks, double* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < nblocks; ++i) t += block_sums[i];
        *out = t;
    }
}

cudaError_t kmeans_set_attributes() {
    const int big = 200 * 1024;
    cudaError_t e = cudaFuncSetAttribute(kmeans_assign_kernel<kF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kmeans_assign_kernel<kU8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kmeans_assign_kernel<kS8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(objective_kernel<kF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(objective_kernel<kU8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(objective_kernel<kS8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    return e;
}

void launch_kmeans_assign(const AssignParams& p, unsigned grid, size_t smem, cudaStream_t st) {
    if (p.pts.storage == kF32) kmeans_assign_kernel<kF32><<<grid, kStreamTile, smem, st>>>(p);
    else if (p.pts.storage == kU8) kmeans_assign_kernel<kU8><<<grid, kStreamTile, smem, st>>>(p);
    else kmeans_assign_kernel<kS8><<<grid, kStreamTile, smem, st>>>(p);
}
void launch_histogram_keys(const int32_t* assign, size_t n, int kc, unsigned long long* counts, uint32_t* keys, uint32_t* index,
                           int blocks, cudaStream_t st) {
    const size_t smem = kc <= kHistSmemClusters ? (size_t) kc * sizeof(uint32_t) : 0;
    histogram_keys_kernel<<<blocks, 256, smem, st>>>(assign, n, kc, counts, keys, index);
}
void launch_cluster_segments(const uint64_t* counts, int kc, uint32_t* seg_begin, uint64_t* member_begin, cudaStream_t st) {
    cluster_segments_kernel<<<1, 32, 0, st>>>(counts, kc, seg_begin, member_begin);
}
void launch_cluster_reduce_segments(const ReduceParams& p, unsigned grid, cudaStream_t st) {
    const bool vec4 = p.pts.storage == kF32 && p.pts.d % 4 == 0 && p.pts.d >= 4 && p.pts.d <= 256 && p.pts.pitch % 16 == 0 &&
                      (reinterpret_cast<uintptr_t>(p.pts.rows) & 15) == 0;
    if (vec4) cluster_reduce_segments_vec4_kernel<<<grid, 256, 0, st>>>(p);
    else if (p.pts.storage == kF32) cluster_reduce_segments_kernel<kF32><<<grid, 256, 0, st>>>(p);
    else if (p.pts.storage == kU8) cluster_reduce_segments_kernel<kU8><<<grid, 256, 0, st>>>(p);
    else cluster_reduce_segments_kernel<kS8><<<grid, 256, 0, st>>>(p);
    if (p.row_norm && p.seg_norm_sums) segment_norm_sums_kernel<<<grid, 256, 0, st>>>(p);
}
void launch_row_norms(const RowsView& pts, float* out_norm, int blocks, cudaStream_t st) {
    if (pts.storage == kF32) row_norms_kernel<kF32><<<blocks, 256, 0, st>>>(pts, out_norm);
    else if (pts.storage == kU8) row_norms_kernel<kU8><<<blocks, 256, 0, st>>>(pts, out_norm);
    else row_norms_kernel<kS8><<<blocks, 256, 0, st>>>(pts, out_norm);
}
void launch_kmeans_update_ip(float* centroids, const float* sums, const uint64_t* counts, const float* norm_sums, unsigned char* alive, int kc,
                             int d, cudaStream_t st) {
    kmeans_update_ip_kernel<<<(unsigned) kc, 128, 0, st>>>(centroids, sums, counts, norm_sums, alive, kc, d);
}
void launch_cluster_finalize(const float* seg_sums, const uint32_t* seg_begin, const uint64_t* counts, int kc, int d, int normalize,
                             float* out, const float* seg_norm_sums, float* out_norm_sums, cudaStream_t st) {
    cluster_finalize_kernel<<<(unsigned) kc, 128, 0, st>>>(seg_sums, seg_begin, counts, kc, d, normalize, out, seg_norm_sums, out_norm_sums);
}
void launch_kmeans_update(float* centroids, const float* sums, const uint64_t* counts, unsigned char* alive, int kc, int d, cudaStream_t st) {
    kmeans_update_kernel<<<(unsigned) kc, 128, 0, st>>>(centroids, sums, counts, alive, kc, d);
}
void launch_kmeans_compact(float* centroids, uint64_t* counts, const unsigned char* alive, int kc, int d, int32_t* remap, int* kc_out,
                           float* cent_tmp, cudaStream_t st) {
    kmeans_compact_kernel<<<1, 1024, 0, st>>>(centroids, counts, alive, kc, d, remap, kc_out, cent_tmp);
}
void launch_remap_assign(int32_t* assign, size_t n, const int32_t* remap, int kc, int blocks, cudaStream_t st) {
    remap_assign_kernel<<<blocks, 256, 0, st>>>(assign, n, remap, kc);
}
void launch_objective(const ObjectiveParams& p, unsigned grid, size_t smem, double* d_out, cudaStream_t st) {
    if (p.pts.storage == kF32) objective_kernel<kF32><<<grid, kStreamTile, smem, st>>>(p);
    else if (p.pts.storage == kU8) objective_kernel<kU8><<<grid, kStreamTile, smem, st>>>(p);
    else objective_kernel<kS8><<<grid, kStreamTile, smem, st>>>(p);
    objective_final_kernel<<<1, 32, 0, st>>>(p.block_sums, (int) grid, d_out);
}
cudaError_t sort_pairs_u32(void* tmp, size_t& tmp_bytes, const uint32_t* keys_in, uint32_t* keys_out, const uint32_t* vals_in,
                           uint32_t* vals_out, size_t n, int bits, cudaStream_t st) {
    return cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, vals_in, vals_out, (int) n, 0, bits, st);
}

}  // namespace gpann
