// mma_issue_bench.cu -- microbenchmark: back-to-back tcgen05.mma throughput per SM on sm_100a for the instruction
// patterns the top-k kernel could use (garbage operands, no TMA, no drain): clocks per MMA instruction.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_issue_bench mma_issue_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t) ((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t) 1 << 16;
    d |= (uint64_t) (1024 >> 4) << 32;
    d |= (uint64_t) 1 << 46;
    d |= (uint64_t) 2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_i8(int n) { return (2u << 4) | ((uint32_t) (n >> 3) << 17) | ((uint32_t) (128 >> 4) << 24); }
__host__ __device__ constexpr uint32_t idesc_bf16(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (n >> 3) << 17) | ((uint32_t) (128 >> 4) << 24); }

#define MMA(OP, d, a, b, idesc, acc)                                                                            \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t" OP " [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), \
                 "l"(b), "r"(idesc), "r"(acc)                                                                   \
                 : "memory")

enum { SS256_CHAIN = 0, SS256_ALT, SS256_COLL, WS256, WS128, SS128_CHAIN, BF16_256_CHAIN, SS256_NOACC, SS256_CHAIN2, NPAT };
static const char* kNames[NPAT] = { "i8 SS N256, 4-chain per accumulator (the kernel today)", "i8 SS N256, alternate accumulators every MMA",
                                    "i8 SS N256, collector::a fill + 3x use (same A)", "i8 WS N256, B stationary in collectors b0-b3, A streams",
                                    "i8 WS N128, B stationary in collectors b0-b3, A streams", "i8 SS N128, 4-chain", "bf16 SS N256, 4-chain",
                                    "i8 SS N256, never accumulates (no D read)", "i8 SS N256, 4-chain, two interleaved accumulators" };

template <int P>
__global__ void __launch_bounds__(128, 1) bench(int iters, unsigned long long* out, int random_fill) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint32_t s_tmem;
    __shared__ uint64_t s_bar;
    const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - raw))[i] = random_fill ? (i * 2654435761u + blockIdx.x * 40503u) ^ ((i * 2246822519u) >> 13) : 0x01010101u * (i & 3);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = s_tmem;
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        // tiles: "small" operand tiles of 128 rows (16 KB) at base + 16K*j, j = 0..3; "large" tiles of 256 rows (32 KB) at base + 64K + 32K*j, j = 0..2
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t small = base + 16384u * (uint32_t) (i & 3), large = base + 65536u + 32768u * (uint32_t) (i % 3);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint32_t acc = ks > 0;
                if (P == SS256_CHAIN) MMA("tcgen05.mma.cta_group::1.kind::i8", tm + 256 * (i & 1), sw128_desc(small) + 2 * ks, sw128_desc(large) + 2 * ks, idesc_i8(256), acc);
                if (P == SS256_ALT) MMA("tcgen05.mma.cta_group::1.kind::i8", tm + 256 * (ks & 1), sw128_desc(small) + 2 * ks, sw128_desc(large) + 2 * ks, idesc_i8(256), 1u);
                if (P == SS256_NOACC) MMA("tcgen05.mma.cta_group::1.kind::i8", tm + 256 * (i & 1), sw128_desc(small) + 2 * ks, sw128_desc(large) + 2 * ks, idesc_i8(256), 0u);
                if (P == SS256_COLL) {
                    if (ks == 0) MMA("tcgen05.mma.cta_group::1.kind::i8.collector::a::fill", tm + 256 * (i & 1), sw128_desc(small), sw128_desc(large) + 2 * ks, idesc_i8(256), acc);
                    else if (ks < 3) MMA("tcgen05.mma.cta_group::1.kind::i8.collector::a::use", tm + 256 * (i & 1), sw128_desc(small), sw128_desc(large) + 2 * ks, idesc_i8(256), acc);
                    else MMA("tcgen05.mma.cta_group::1.kind::i8.collector::a::lastuse", tm + 256 * (i & 1), sw128_desc(small), sw128_desc(large) + 2 * ks, idesc_i8(256), acc);
                }
                if (P == WS256 || P == WS128) {
                    // B (stationary) = the first large tile, K slice ks lives in collector b<ks>; A (streams) = small tile i
                    const uint32_t d = tm + (P == WS256 ? 256 : 128) * (i & 1), id = idesc_i8(P == WS256 ? 256 : 128);
                    const uint64_t bd = sw128_desc(base + 65536u) + 2 * ks, ad = sw128_desc(small) + 2 * ks;
                    if (i == 0) {
                        if (ks == 0) MMA("tcgen05.mma.ws.cta_group::1.kind::i8.collector::b0::fill", d, ad, bd, id, acc);
                        if (ks == 1) MMA("tcgen05.mma.ws.cta_group::1.kind::i8.collector::b1::fill", d, ad, bd, id, acc);
                        if (ks == 2) MMA("tcgen05.mma.ws.cta_group::1.kind::i8.collector::b2::fill", d, ad, bd, id, acc);
                        if (ks == 3) MMA("tcgen05.mma.ws.cta_group::1.kind::i8.collector::b3::fill", d, ad, bd, id, acc);
                    } else {
                        if (ks == 0) MMA("tcgen05.mma.ws.cta_group::1.kind::i8.collector::b0::use", d, ad, bd, id, acc);
                        if (ks == 1) MMA("tcgen05.mma.ws.cta_group::1.kind::i8.collector::b1::use", d, ad, bd, id, acc);
                        if (ks == 2) MMA("tcgen05.mma.ws.cta_group::1.kind::i8.collector::b2::use", d, ad, bd, id, acc);
                        if (ks == 3) MMA("tcgen05.mma.ws.cta_group::1.kind::i8.collector::b3::use", d, ad, bd, id, acc);
                    }
                }
                if (P == SS128_CHAIN) MMA("tcgen05.mma.cta_group::1.kind::i8", tm + 128 * (i & 3), sw128_desc(small) + 2 * ks, sw128_desc(large) + 2 * ks, idesc_i8(128), acc);
                if (P == BF16_256_CHAIN) MMA("tcgen05.mma.cta_group::1.kind::f16", tm + 256 * (i & 1), sw128_desc(small) + 2 * ks, sw128_desc(large) + 2 * ks, idesc_bf16(256), acc);
                if (P == SS256_CHAIN2) {  // two accumulators advance together: consecutive MMAs never share D
                    MMA("tcgen05.mma.cta_group::1.kind::i8", tm, sw128_desc(small) + 2 * ks, sw128_desc(large) + 2 * ks, idesc_i8(256), acc);
                    MMA("tcgen05.mma.cta_group::1.kind::i8", tm + 256, sw128_desc(small ^ 16384u) + 2 * ks, sw128_desc(large) + 2 * ks, idesc_i8(256), acc);
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(ok) : "r"(smem_u32(&s_bar)) : "memory");
        t1 = clock64();
        out[blockIdx.x] = (unsigned long long) (t1 - t0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

template <int P>
void run(unsigned long long* d_out, int grid, int iters = 20000, int random_fill = 0) {
    const int smem = 162 * 1024;
    cudaFuncSetAttribute(bench<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    bench<P><<<grid, 128, smem>>>(200, d_out, random_fill);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<P><<<grid, 128, smem>>>(iters, d_out, random_fill);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[148];
    cudaMemcpy(h, d_out, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost);
    unsigned long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    const double mmas = (double) iters * 4 * (P == SS256_CHAIN2 ? 2 : 1);
    const int n = (P == WS128 || P == SS128_CHAIN) ? 128 : 256;
    const double ops = mmas * 2.0 * 128 * n * (P == BF16_256_CHAIN ? 16 : 32) * grid;
    printf("[grid %3d, %6d iters, %s data] %-62s: %7.1f clk/MMA, %8.1f T(FL)OP/s (%.2f ms)  %s\n", grid, iters, random_fill ? "random" : "const", kNames[P], (double) mx / mmas, ops / (ms * 1e-3) / 1e12, ms,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}
int main() {
    unsigned long long* d_out;
    cudaMalloc(&d_out, 148 * 8);
    for (int grid : { 1, 148 }) {
        run<SS256_CHAIN>(d_out, grid);
        run<SS256_CHAIN2>(d_out, grid);
        run<SS256_ALT>(d_out, grid);
        run<SS256_NOACC>(d_out, grid);
        run<SS256_COLL>(d_out, grid);
        run<SS128_CHAIN>(d_out, grid);
        run<BF16_256_CHAIN>(d_out, grid);
        run<WS256>(d_out, grid);
        run<WS128>(d_out, grid);
    }
    // sustained (power-capped) rates: ~0.1-0.3 s of back-to-back MMAs on every SM, random operand bits
    for (int rep = 0; rep < 2; ++rep) {
        run<SS256_CHAIN>(d_out, 148, 500000, 1);
        run<BF16_256_CHAIN>(d_out, 148, 500000, 1);
        run<SS256_CHAIN>(d_out, 148, 500000, 0);
    }
    return 0;
}
