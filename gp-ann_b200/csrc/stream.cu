// stream.cu -- HBM-streaming exact scan (CUDA cores) + key merge.
//
// Replaces the reference's scalar triple loop `for query: for point: distance(); TopN.Add()`
// (src/recall.h:27-37, src/defs.cpp:87-96, src/inverted_index.h:43-75) for SMALL query batches, where the path
// is bound by streaming the point rows from HBM, and serves as the exact (reference summation order) kernel
// for every shape the tensor kernel does not cover.
//
// Work decomposition: a CTA walks point tiles of kTile rows (coalesced 16-byte loads into padded shared
// memory), one thread per point.  For each query of the current batch the thread computes the distance in
// the reference's own arithmetic order and, if the (distance,id) key beats the CTA's current k-th best for
// that query, appends it to a shared-memory candidate buffer; a warp compacts a buffer (register bitonic
// select) before it could overflow, which also tightens the threshold.  Each CTA finally emits its k best
// keys per query and merge_keys_kernel reduces over CTAs.
#include "common.cuh"
#include "distance.cuh"
#include "params.h"

namespace gpann {

template <int STORAGE>
__global__ void __launch_bounds__(kStreamTile, 5) topk_stream_kernel(const StreamParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (p.gate && ((*reinterpret_cast<const volatile int*>(p.gate) != 0) != (p.gate_want != 0))) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // query batches of qb: blockIdx.y strides over them (the host keeps gridDim.y <= 65535)
    for (int q0 = blockIdx.y * p.qb; q0 < p.nq; q0 += gridDim.y * p.qb) {
    const int nqb = min(p.qb, p.nq - q0);
    const int elem = STORAGE == kF32 ? 4 : 1;
    const int row_bytes = p.d * elem;

    unsigned char* s_rows0 = smem_raw;                                           // nbuf * kStreamTile * spitch
    const int trows = p.tile_rows;  // rows per tile: kStreamTile unless a row is too long for 128 of them in shared memory
    const size_t tile_smem = (size_t) trows * p.spitch;
    const int nbuf = max(1, min(p.nbuf, 4));
    float* s_q = reinterpret_cast<float*>(s_rows0 + (size_t) nbuf * tile_smem);   // qb * qpitch
    uint64_t* s_buf = reinterpret_cast<uint64_t*>(s_q + (size_t) p.qb * p.qpitch);    // qb * cap
    uint64_t* s_thr = s_buf + (size_t) p.qb * p.cap;                              // qb
    int* s_cnt = reinterpret_cast<int*>(s_thr + p.qb);                            // qb
    // byte storage, L2: queries whose coordinates are integers in the storage's range take an exact integer path
    // (dist = |x|^2 + |q|^2 - 2 x.q with dp4a: ~50 instead of ~520 instructions per point and query); their packed bytes:
    const int qbpitch = (p.qpitch + 15) & ~15;                                    // bytes per packed query (multiple of 16)
    int* s_qn = s_cnt + p.qb;                                                     // qb: |q|^2
    int* s_qint = s_qn + p.qb;                                                    // qb: 1 = integer-valued in range
    unsigned char* s_qb = reinterpret_cast<unsigned char*>(                       // qb * qbpitch, 16-byte aligned
        (reinterpret_cast<uintptr_t>(s_qint + p.qb) + 15) & ~(uintptr_t) 15);
    // ... as long as every partial sum of the reference's float arithmetic is an integer below 2^24 (d * 255^2, d * 128^2):
    // longer rows round inside the reference's accumulators, and only its own float order reproduces that
    const bool int_path = STORAGE != kF32 && p.metric == kL2 && p.d <= (STORAGE == kS8 ? 1023 : 258);

    for (int i = tid; i < nqb * p.d; i += kStreamTile) {
        const int q = i / p.d, j = i - q * p.d;
        s_q[q * p.qpitch + j] = p.queries[(size_t) (q0 + q) * p.d + j];
    }
    for (int q = tid; q < nqb; q += kStreamTile) {
        s_thr[q] = kEmptyKey;
        s_cnt[q] = 0;
        s_qn[q] = 0;
        s_qint[q] = 1;
    }
    __syncthreads();
    if (int_path) {
        for (int i = tid; i < nqb * qbpitch; i += kStreamTile) {
            const int q = i / qbpitch, j = i - q * qbpitch;
            int iv = 0;
            if (j < p.d) {
                const float v = s_q[q * p.qpitch + j];
                const float lo = STORAGE == kS8 ? -128.f : 0.f, hi = STORAGE == kS8 ? 127.f : 255.f;
                if (!(v == rintf(v) && v >= lo && v <= hi)) s_qint[q] = 0;  // benign race: every writer stores 0
                else iv = (int) v;
                atomicAdd(&s_qn[q], iv * iv);
            }
            s_qb[i] = (unsigned char) (iv & 0xFF);
        }
        __syncthreads();
    }

    const size_t nrows = p.row_end - p.row_begin;
    const size_t ntiles = (nrows + trows - 1) / trows;
    // Double-buffered ring (nbuf == 2): the rows of the CTA's NEXT tile travel HBM -> shared memory with cp.async (16 bytes
    // each, no register staging) while this tile is scanned; a single-buffered CTA alternates between a load phase and a
    // compute phase and leaves the memory system idle half of the time (u8 nq = 1: 0.43 of the copy bandwidth).
    auto fetch_async = [&](size_t tile, unsigned char* dst) {
        const size_t r0 = p.row_begin + tile * trows;
        const int rows_here = (int) min((size_t) trows, p.row_end - r0);
        const char* src = p.rows + r0 * p.pitch;
        const int cpr = row_bytes >> 4;
        for (int idx = tid; idx < rows_here * cpr; idx += kStreamTile) {
            const int r = idx / cpr, c = idx - r * cpr;
            const uint32_t d32 = (uint32_t) __cvta_generic_to_shared(dst + (size_t) r * p.spitch + c * 16);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d32), "l"(src + (size_t) r * p.pitch + (size_t) c * 16) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // ring of nbuf tiles, nbuf - 1 of them in flight: one commit group per iteration (an empty one past the end)
    for (int b = 0; b + 1 < nbuf; ++b) {
        const size_t t = blockIdx.x + (size_t) b * gridDim.x;
        if (t < ntiles) fetch_async(t, s_rows0 + (size_t) b * tile_smem);
        else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    int it = 0;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const size_t r0 = p.row_begin + tile * trows;
        const int rows_here = (int) min((size_t) trows, p.row_end - r0);
        const char* src = p.rows + r0 * p.pitch;
        unsigned char* s_rows = s_rows0 + (size_t) (it % nbuf) * tile_smem;
        if (nbuf > 1) {
            // the buffer refilled now was last read in the previous iteration (a __syncthreads() closes every iteration)
            const size_t tnext = tile + (size_t) (nbuf - 1) * gridDim.x;
            if (tnext < ntiles) fetch_async(tnext, s_rows0 + (size_t) ((it + nbuf - 1) % nbuf) * tile_smem);
            else asm volatile("cp.async.commit_group;" ::: "memory");
            // this tile's group has landed when at most nbuf - 1 younger groups are pending
            if (nbuf == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else if (nbuf == 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
            else asm volatile("cp.async.wait_group 3;" ::: "memory");
        } else if (p.vec_ok) {
            const int cpr = row_bytes >> 4;  // 16-byte chunks per row (row_bytes % 16 == 0 when vec_ok)
            for (int idx = tid; idx < rows_here * cpr; idx += kStreamTile) {
                const int r = idx / cpr, c = idx - r * cpr;
                const int4 v = __ldg(reinterpret_cast<const int4*>(src + (size_t) r * p.pitch) + c);
                *reinterpret_cast<int4*>(s_rows + (size_t) r * p.spitch + c * 16) = v;
            }
        } else {
            for (int idx = tid; idx < rows_here * row_bytes; idx += kStreamTile) {
                const int r = idx / row_bytes, c = idx - r * row_bytes;
                s_rows[(size_t) r * p.spitch + c] = (unsigned char) src[(size_t) r * p.pitch + c];
            }
        }
        __syncthreads();
        if (tid < rows_here) {
            const size_t row = r0 + tid;
            const uint32_t id = p.ids ? p.ids[row] : p.id_base + (uint32_t) row;
            if (id != kInvalidId && id >= p.id_lo && id < p.id_hi) {
                const char* myrow = reinterpret_cast<const char*>(s_rows) + (size_t) tid * p.spitch;
                int xx = 0;  // |x|^2, once per point (byte storage)
                if (int_path) {
                    for (int j = 0; j < p.d; j += 16) {
                        uint4 x = *reinterpret_cast<const uint4*>(myrow + j);
                        if (j + 16 > p.d) {  // the staged row is valid up to d bytes only: clear the tail
                            uint32_t w[4] = { x.x, x.y, x.z, x.w };
                            for (int b = 0; b < 16; ++b)
                                if (j + b >= p.d) w[b >> 2] &= ~(0xFFu << (8 * (b & 3)));
                            x = make_uint4(w[0], w[1], w[2], w[3]);
                        }
                        if (STORAGE == kS8) {
                            xx = __dp4a((int) x.x, (int) x.x, xx); xx = __dp4a((int) x.y, (int) x.y, xx);
                            xx = __dp4a((int) x.z, (int) x.z, xx); xx = __dp4a((int) x.w, (int) x.w, xx);
                        } else {
                            xx = (int) __dp4a(x.x, x.x, (unsigned) xx); xx = (int) __dp4a(x.y, x.y, (unsigned) xx);
                            xx = (int) __dp4a(x.z, x.z, (unsigned) xx); xx = (int) __dp4a(x.w, x.w, (unsigned) xx);
                        }
                    }
                }
                for (int q = 0; q < nqb; ++q) {
                    float dist;
                    if (int_path && s_qint[q]) {
                        // exact integers below 2^24: bit-identical to the reference's float result (src/dist.cpp:9-16)
                        const unsigned char* qb_row = s_qb + q * qbpitch;
                        int xq = 0;
                        for (int j = 0; j < p.d; j += 16) {  // the packed query is zero beyond d, so the row's tail bytes cancel
                            const uint4 x = *reinterpret_cast<const uint4*>(myrow + j);
                            const uint4 y = *reinterpret_cast<const uint4*>(qb_row + j);
                            if (STORAGE == kS8) {
                                xq = __dp4a((int) x.x, (int) y.x, xq); xq = __dp4a((int) x.y, (int) y.y, xq);
                                xq = __dp4a((int) x.z, (int) y.z, xq); xq = __dp4a((int) x.w, (int) y.w, xq);
                            } else {
                                xq = (int) __dp4a(x.x, y.x, (unsigned) xq); xq = (int) __dp4a(x.y, y.y, (unsigned) xq);
                                xq = (int) __dp4a(x.z, y.z, (unsigned) xq); xq = (int) __dp4a(x.w, y.w, (unsigned) xq);
                            }
                        }
                        dist = (float) (xx + s_qn[q] - 2 * xq);
                    } else {
                        dist = ref_distance<STORAGE>(myrow, s_q + q * p.qpitch, p.d, p.metric);
                    }
                    const uint64_t key = pack_key(dist, id);
                    if (key < s_thr[q]) {
                        const int pos = atomicAdd(&s_cnt[q], 1);
                        s_buf[(size_t) q * p.cap + pos] = key;
                    }
                }
            }
        }
        __syncthreads();
        // compact any buffer that could overflow during the next tile
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
    // final selection -> this CTA's partial result
    for (int q = warp; q < nqb; q += kStreamTile / 32) {
        uint64_t* b = s_buf + (size_t) q * p.cap;
        uint64_t* dst = p.partial + ((size_t) blockIdx.x * p.nq + (q0 + q)) * p.k;
        warp_select_topk_dyn<32>(b, s_cnt[q], p.k, dst, lane, p.cap);
    }
    if (nbuf > 1) asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // the next query batch reuses the shared-memory buffers
    }
}

// Merge `parts` key lists per query (k keys each, any order, kEmptyKey = none) into the k smallest.  One warp per query; keys [parts, nq, k].
// Output: any of out_dist [nq,k] (float), out_id [nq,k], out_keys [nq,k], out_count [nq], out_kth [nq] (the
// reference's Top().first: the largest kept distance).
__global__ void __launch_bounds__(128) merge_keys_kernel(const MergeParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t q = (size_t) blockIdx.x * 4 + warp;
    if (q >= p.nq) return;
    if (p.gate && ((*reinterpret_cast<const volatile int*>(p.gate) != 0) != (p.gate_want != 0))) return;
    uint64_t* win = reinterpret_cast<uint64_t*>(smem_raw) + (size_t) warp * p.cap;  // [cap]
    // running best k in win[0..k), incoming keys appended behind, select when the window is full
    int cnt = 0;
    const int chunk = p.cap - p.k;  // free slots after keeping k
    const size_t total = p.parts * (size_t) p.k;
    for (size_t base = 0; base < total; base += chunk) {
        const int take = (int) min((size_t) chunk, total - base);
        for (int i = lane; i < take; i += 32) {
            const size_t g = base + i;
            const size_t part = g / p.k, r = g - part * p.k;
            win[cnt + i] = p.keys[(part * p.nq_stride + q) * p.k + r];
        }
        __syncwarp();
        cnt += take;
        warp_select_topk_dyn<32>(win, cnt, p.k, win, lane, p.cap);
        cnt = min(cnt, p.k);
    }
    __syncwarp();
    uint32_t valid = 0;
    for (int i = lane; i < p.k; i += 32) {
        const uint64_t key = win[i];
        const bool ok = key != kEmptyKey;
        valid += ok;
        if (p.out_keys) p.out_keys[q * p.k + i] = key;
        if (p.out_dist) p.out_dist[q * p.k + i] = ok ? key_dist(key) : __int_as_float(0x7f800000);
        if (p.out_id) p.out_id[q * p.k + i] = ok ? key_id(key) : kInvalidId;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) valid += __shfl_xor_sync(0xffffffffu, valid, o);
    if (lane == 0) {
        if (p.out_count) p.out_count[q] = valid;
        const bool have = p.kth_full_only ? (valid == (uint32_t) p.k) : (valid > 0);
        if (p.out_kth) {
            float v = have ? key_dist(win[valid - 1]) : __int_as_float(0x7f800000);
            if (p.kth_full_only) v = fminf(v, p.out_kth[q]);  // bounds only ever tighten
            p.out_kth[q] = v;
        }
    }
}

cudaError_t stream_set_attributes() {
    const int big = 200 * 1024;
    cudaError_t e = cudaFuncSetAttribute(topk_stream_kernel<kF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(topk_stream_kernel<kU8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(topk_stream_kernel<kS8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    return e;
}

void launch_topk_stream(const StreamParams& p, dim3 grid, size_t smem, cudaStream_t st) {
    if (p.storage == kF32) topk_stream_kernel<kF32><<<grid, kStreamTile, smem, st>>>(p);
    else if (p.storage == kU8) topk_stream_kernel<kU8><<<grid, kStreamTile, smem, st>>>(p);
    else topk_stream_kernel<kS8><<<grid, kStreamTile, smem, st>>>(p);
}

void launch_merge_keys(const MergeParams& p, unsigned grid, size_t smem, cudaStream_t st) {
    merge_keys_kernel<<<grid, 128, smem, st>>>(p);
}

}  // namespace gpann
