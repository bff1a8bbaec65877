// seg.cu -- segmented exact scans (CUDA cores, the reference's own arithmetic order): many small brute-force scans in ONE
// launch.  The callers either side of the big scans have this shape:
//   * InvertedIndex::Query / QueryBucket (src/inverted_index.h:43-75): a query scans the rows of the buckets it probes;
//     candidates are ranked by (distance, POSITION in the clustered copy) and the ids are remapped through `permutation`
//     afterwards -- here for a whole batch of queries at once, each with its own list of buckets;
//   * ApproximateKNNGraphBuilder::CrunchBucket (src/knn_graph.h:165-185): all pairs inside a bucket of <= 5000 points,
//     per-point top-k of (distance, global id), a point never its own neighbour (:18-19) -- here for all buckets at once;
//   * ClosestLeaders for every point of a set (src/knn_graph.h:75-81, src/defs.cpp:87-96): n points x <= 1500 leaders,
//     top-fanout by (distance, leader index).
// All three are "tasks": a contiguous range of scan positions x a batch of <= kSegQB queries.  One CTA per task walks the
// range in tiles of 128 positions (one thread per position; rows gathered through an optional position -> row map and
// staged through padded shared memory with 16-byte loads) and keeps, per query, a candidate buffer that a warp compacts
// before it can overflow -- the machinery of topk_stream_kernel (stream.cu), driven by a task list instead of the grid.
// Every (task, query) pair owns one slot of k keys; seg_merge_kernel folds a query's slots into its final list.
#include "common.cuh"
#include "distance.cuh"
#include "params.h"

namespace gpann {

// float view of element i of a row of any storage (query rows may be rows of a byte point set)
__device__ __forceinline__ float elem_as_float(const char* row, int storage, int i) {
    if (storage == kF32) return reinterpret_cast<const float*>(row)[i];
    if (storage == kU8) return (float) reinterpret_cast<const uint8_t*>(row)[i];
    return (float) reinterpret_cast<const int8_t*>(row)[i];
}

template <int STORAGE>
__global__ void __launch_bounds__(kStreamTile, 5) seg_scan_kernel(const SegParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int elem = STORAGE == kF32 ? 4 : 1;
    const int row_bytes = p.d * elem;
    unsigned char* s_rows = smem_raw;                                                   // kStreamTile * spitch
    float* s_q = reinterpret_cast<float*>(s_rows + (size_t) kStreamTile * p.spitch);   // kSegQB * qpitch
    uint64_t* s_buf = reinterpret_cast<uint64_t*>(s_q + (size_t) kSegQB * p.qpitch);   // kSegQB * cap
    uint64_t* s_thr = s_buf + (size_t) kSegQB * p.cap;                                  // kSegQB
    int* s_cnt = reinterpret_cast<int*>(s_thr + kSegQB);                                // kSegQB
    uint32_t* s_self = reinterpret_cast<uint32_t*>(s_cnt + kSegQB);                     // kSegQB: the query's own row (exclude_self)
    uint32_t* s_rowid = s_self + kSegQB;                                                // kStreamTile: row of each staged position

    for (int task_i = blockIdx.x; task_i < p.ntasks; task_i += gridDim.x) {
        const SegTask task = p.tasks[task_i];
        const int nqb = (int) task.q_count;
        __syncthreads();  // the previous task's buffers are done with
        // ---- this task's queries -> shared memory (float), from the query matrix or from rows of a point set ----
        for (int i = tid; i < nqb * p.d; i += kStreamTile) {
            const int q = i / p.d, j = i - q * p.d;
            const uint32_t qi = p.task_query[task.q_begin + q];
            const uint32_t qrow = p.query_rows ? p.query_rows[qi] : qi;
            s_q[q * p.qpitch + j] = elem_as_float(p.q_base + (size_t) qrow * p.q_pitch, p.q_storage, j);
        }
        for (int q = tid; q < nqb; q += kStreamTile) {
            s_thr[q] = kEmptyKey;
            s_cnt[q] = 0;
            const uint32_t qi = p.task_query[task.q_begin + q];
            s_self[q] = p.exclude_self ? (p.query_rows ? p.query_rows[qi] : qi) : kInvalidId;
        }
        __syncthreads();
        for (uint64_t pos0 = task.pos_begin; pos0 < task.pos_end; pos0 += kStreamTile) {
            const int rows_here = (int) min((uint64_t) kStreamTile, task.pos_end - pos0);
            if (tid < rows_here) s_rowid[tid] = p.row_map ? p.row_map[pos0 + tid] : (uint32_t) (pos0 + tid);
            __syncthreads();
            if (p.vec_ok) {
                const int cpr = row_bytes >> 4;
                for (int idx = tid; idx < rows_here * cpr; idx += kStreamTile) {
                    const int r = idx / cpr, c = idx - r * cpr;
                    const int4 v = __ldg(reinterpret_cast<const int4*>(p.rows + (size_t) s_rowid[r] * p.pitch) + c);
                    *reinterpret_cast<int4*>(s_rows + (size_t) r * p.spitch + c * 16) = v;
                }
            } else {
                for (int idx = tid; idx < rows_here * row_bytes; idx += kStreamTile) {
                    const int r = idx / row_bytes, c = idx - r * row_bytes;
                    s_rows[(size_t) r * p.spitch + c] = (unsigned char) p.rows[(size_t) s_rowid[r] * p.pitch + c];
                }
            }
            __syncthreads();
            if (tid < rows_here) {
                const uint32_t row = s_rowid[tid];
                // ranking id: the scan position (IVF: ties by position in the clustered copy) or the global row
                const uint32_t id = p.id_is_row ? row + p.id_base : (uint32_t) (pos0 + tid - p.pos_id_origin);
                const char* myrow = reinterpret_cast<const char*>(s_rows) + (size_t) tid * p.spitch;
                for (int q = 0; q < nqb; ++q) {
                    if (row == s_self[q]) continue;  // a point is never its own neighbour (src/knn_graph.h:18-19, :170-171)
                    const float dist = ref_distance<STORAGE>(myrow, s_q + q * p.qpitch, p.d, p.metric);
                    const uint64_t key = pack_key(dist, id);
                    if (key < s_thr[q]) {
                        const int at = atomicAdd(&s_cnt[q], 1);
                        s_buf[(size_t) q * p.cap + at] = key;
                    }
                }
            }
            __syncthreads();
            for (int q = warp; q < nqb; q += kStreamTile / 32) {
                const int cnt = s_cnt[q];
                if (cnt > p.cap - kStreamTile) {
                    uint64_t* b = s_buf + (size_t) q * p.cap;
                    const uint64_t kth = warp_select_topk_dyn<32>(b, cnt, p.k, b, lane, p.cap);
                    if (lane == 0) {
                        s_cnt[q] = min(cnt, p.k);
                        s_thr[q] = kth;
                    }
                }
            }
            __syncthreads();
        }
        for (int q = warp; q < nqb; q += kStreamTile / 32) {
            uint64_t* b = s_buf + (size_t) q * p.cap;
            warp_select_topk_dyn<32>(b, s_cnt[q], p.k, p.partial + (size_t) (task.q_begin + q) * p.k, lane, p.cap);
        }
    }
}

// One warp per query: fold the query's slots (slot_list[slot_begin[q] .. slot_begin[q+1])) of k ascending keys each into
// the k smallest; ids optionally remapped (InvertedIndex: permutation[position], src/inverted_index.h:56-59).
__global__ void __launch_bounds__(128) seg_merge_kernel(const SegMergeParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t q = (size_t) blockIdx.x * 4 + warp;
    if (q >= p.nq) return;
    uint64_t* win = reinterpret_cast<uint64_t*>(smem_raw) + (size_t) warp * p.cap;
    int cnt = 0;
    const int per = (p.cap - p.k) / p.k;  // slots per selection round
    const uint32_t sb = p.slot_begin[q], se = p.slot_begin[q + 1];
    for (uint32_t s0 = sb; s0 < se; s0 += per) {
        const int take = (int) min((uint32_t) per, se - s0) * p.k;
        for (int i = lane; i < take; i += 32) {
            const uint32_t slot = p.slot_list ? p.slot_list[s0 + i / p.k] : s0 + i / p.k;
            win[cnt + i] = p.keys[(size_t) slot * p.k + i % p.k];
        }
        __syncwarp();
        cnt += take;
        warp_select_topk_dyn<32>(win, cnt, p.k, win, lane, p.cap);
        cnt = min(cnt, p.k);
    }
    if (sb == se)
        for (int i = lane; i < p.k; i += 32) win[i] = kEmptyKey;
    __syncwarp();
    uint32_t valid = 0;
    for (int i = lane; i < p.k; i += 32) {
        const uint64_t key = win[i];
        const bool ok = key != kEmptyKey;
        valid += ok;
        uint32_t id = key_id(key);
        if (ok && p.id_remap) id = p.id_remap[id];
        if (p.out_dist) p.out_dist[q * p.k + i] = ok ? key_dist(key) : __int_as_float(0x7f800000);
        if (p.out_id) p.out_id[q * p.k + i] = ok ? id : kInvalidId;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) valid += __shfl_xor_sync(0xffffffffu, valid, o);
    if (lane == 0 && p.out_count) p.out_count[q] = valid;
}

// ---- per-group minimum distance ------------------------------------------------------------------------------------------
// out[q][g] = min over the points of group g of distance(P_i, Q_q): BruteForceRouting (analyze_approximation_losses.cpp:42-56)
// and the per-cluster ratings of the overlapping partitioner (src/overlapping_partitioning.cpp:170-179).  One thread per
// point, query batches through shared memory, minima kept per CTA in shared memory as order-preserving uint32 images and
// folded into the global table with atomicMin -- min is order-independent, so the result is deterministic.
template <int STORAGE>
__global__ void __launch_bounds__(kStreamTile) group_min_kernel(const GroupMinParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int elem = STORAGE == kF32 ? 4 : 1;
    const int row_bytes = p.d * elem;
    unsigned char* s_rows = smem_raw;
    float* s_q = reinterpret_cast<float*>(s_rows + (size_t) kStreamTile * p.spitch);
    uint32_t* s_min = reinterpret_cast<uint32_t*>(s_q + (size_t) p.qb * p.qpitch);  // [qb][ngroups]
    const uint32_t kMax = float_to_ordered(3.402823466e+38f);
    for (int q0 = blockIdx.y * p.qb; q0 < p.nq; q0 += gridDim.y * p.qb) {
        const int nqb = min(p.qb, p.nq - q0);
        __syncthreads();
        for (int i = tid; i < nqb * p.d; i += kStreamTile) {
            const int q = i / p.d, j = i - q * p.d;
            s_q[q * p.qpitch + j] = p.queries[(size_t) (q0 + q) * p.d + j];
        }
        for (int i = tid; i < nqb * p.ngroups; i += kStreamTile) s_min[i] = kMax;
        __syncthreads();
        const size_t ntiles = (p.n + kStreamTile - 1) / kStreamTile;
        for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const size_t r0 = tile * kStreamTile;
            const int rows_here = (int) min((size_t) kStreamTile, p.n - r0);
            const char* src = p.rows + r0 * p.pitch;
            if (p.vec_ok) {
                const int cpr = row_bytes >> 4;
                for (int idx = tid; idx < rows_here * cpr; idx += kStreamTile) {
                    const int r = idx / cpr, c = idx - r * cpr;
                    *reinterpret_cast<int4*>(s_rows + (size_t) r * p.spitch + c * 16) =
                        __ldg(reinterpret_cast<const int4*>(src + (size_t) r * p.pitch) + c);
                }
            } else {
                for (int idx = tid; idx < rows_here * row_bytes; idx += kStreamTile) {
                    const int r = idx / row_bytes, c = idx - r * row_bytes;
                    s_rows[(size_t) r * p.spitch + c] = (unsigned char) src[(size_t) r * p.pitch + c];
                }
            }
            __syncthreads();
            if (tid < rows_here) {
                const int g = p.group[r0 + tid];
                if (g >= 0 && g < p.ngroups) {
                    const char* myrow = reinterpret_cast<const char*>(s_rows) + (size_t) tid * p.spitch;
                    for (int q = 0; q < nqb; ++q) {
                        const float dist = ref_distance<STORAGE>(myrow, s_q + q * p.qpitch, p.d, p.metric);
                        const uint32_t o = float_to_ordered(dist);
                        if (o < s_min[q * p.ngroups + g]) atomicMin(&s_min[q * p.ngroups + g], o);
                    }
                }
            }
            __syncthreads();
        }
        for (int i = tid; i < nqb * p.ngroups; i += kStreamTile)
            if (s_min[i] != kMax) atomicMin(p.out_ordered + (size_t) q0 * p.ngroups + i, s_min[i]);
    }
}

__global__ void group_min_finish_kernel(uint32_t* x, size_t count) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t) gridDim.x * blockDim.x)
        x[i] = __float_as_uint(ordered_to_float(x[i]));
}

cudaError_t seg_set_attributes() {
    const int big = 200 * 1024;
    cudaError_t e = cudaFuncSetAttribute(seg_scan_kernel<kF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(seg_scan_kernel<kU8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(seg_scan_kernel<kS8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(group_min_kernel<kF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(group_min_kernel<kU8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(group_min_kernel<kS8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    return e;
}

void launch_seg_scan(const SegParams& p, unsigned grid, size_t smem, cudaStream_t st) {
    if (p.storage == kF32) seg_scan_kernel<kF32><<<grid, kStreamTile, smem, st>>>(p);
    else if (p.storage == kU8) seg_scan_kernel<kU8><<<grid, kStreamTile, smem, st>>>(p);
    else seg_scan_kernel<kS8><<<grid, kStreamTile, smem, st>>>(p);
}

void launch_group_min(const GroupMinParams& p, dim3 grid, size_t smem, cudaStream_t st) {
    if (p.storage == kF32) group_min_kernel<kF32><<<grid, kStreamTile, smem, st>>>(p);
    else if (p.storage == kU8) group_min_kernel<kU8><<<grid, kStreamTile, smem, st>>>(p);
    else group_min_kernel<kS8><<<grid, kStreamTile, smem, st>>>(p);
}
void launch_group_min_finish(uint32_t* x, size_t count, int blocks, cudaStream_t st) { group_min_finish_kernel<<<blocks, 256, 0, st>>>(x, count); }

void launch_seg_merge(const SegMergeParams& p, cudaStream_t st) {
    const size_t smem = (size_t) 4 * p.cap * sizeof(uint64_t);
    seg_merge_kernel<<<(unsigned) ((p.nq + 3) / 4), 128, smem, st>>>(p);
}

}  // namespace gpann
