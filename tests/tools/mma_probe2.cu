// Round-2 microbenchmark: what paces small tcgen05.mma (kind::f16, SS mode) instructions?
//   * M = 128 / 64, N = 64 / 128 / 256, one issuing lane per CTA, back-to-back stream into one accumulator
//   * 1 or 2 co-resident CTAs per SM (256 TMEM columns each): does a second issuer overlap the first one's
//     operand-read latency (aggregate rate doubles) or is the tensor pipe's operand fetch serial per SM?
//   * the sparse kernel's real per-K16 pattern: (A0 x B[N=128]) then (A1 x B[N=64])
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I sa-ssd_b200/csrc tests/tools/mma_probe2.cu -o tests/tools/mma_probe2
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>

#include "tc_common.cuh"

using namespace tc;

// pattern 0: one MMA shape (M x N) repeated; pattern 1: alternate (M x 2*N from A0) and (M x N from A1)
template <int M, int N, int PATTERN>
__global__ void __launch_bounds__(128, 2) probe(int iters, unsigned tmem_cols, int random_data, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
    // A0, A1: 16 KB each; B: up to 256 rows x 128 B = 32 KB
    const uint32_t a0 = base, a1 = base + 16384, b_s = base + 32768, bar = base + 65536, slot = bar + 16;
    for (int i = threadIdx.x; i < 65536 / 16; i += blockDim.x) {
        uint4 v = make_uint4(0x3c003c00u, 0x3c003c00u, 0, 0);
        if (random_data) {        // fp16 values in (-2, 2) with random mantissas (operand toggling like real activations)
            uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
            auto nxt = [&]() { h ^= h << 13; h ^= h >> 17; h ^= h << 5; return (h & 0x83FF83FFu) | 0x3C003C00u; };
            v = make_uint4(nxt(), nxt(), nxt(), nxt());
        }
        ((uint4*)bp)[i] = v;
    }
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *(volatile uint32_t*)(bp + (slot - base));
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = make_idesc(M, N, 0u), idesc2 = make_idesc(M, (2 * N <= 256 ? 2 * N : N), 0u);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it)
            for (int k16 = 0; k16 < 4; ++k16) {
                const uint64_t db = make_desc(b_s + k16 * 32);
                if (PATTERN == 0) {
                    mma_f16(tmem, make_desc(a0 + k16 * 32), db, idesc, 1u);
                } else {
                    mma_f16(tmem, make_desc(a0 + k16 * 32), db, idesc2, 1u);                 // ah x [bh|bl]
                    mma_f16(tmem + 2 * N, make_desc(a1 + k16 * 32), db, idesc, 1u);          // al x bh
                }
            }
        mma_commit(bar);
        mbar_wait(bar, 0);
        const long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
    }
}

template <int M, int N, int PATTERN>
static void run(int ctas_per_sm, int random_data) {
    const int grid = 148 * ctas_per_sm, iters = 300;
    const int smem = 65536 + 1024 + 64;
    cudaFuncSetAttribute(probe<M, N, PATTERN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long* dc;
    cudaMalloc(&dc, grid * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<M, N, PATTERN><<<grid, 128, smem>>>(iters, 256u, random_data, dc);   // warm
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<M, N, PATTERN><<<grid, 128, smem>>>(iters, 256u, random_data, dc);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("M=%d N=%d: %s\n", M, N, cudaGetErrorString(e)); return; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), dc, grid * 8, cudaMemcpyDeviceToHost);
    double mean = 0, mx = 0;
    for (auto v : h) { mean += v; if (v > mx) mx = v; }
    mean /= grid;
    const int mmas = iters * 4 * (PATTERN ? 2 : 1);
    printf("%s data M=%3d N=%3d pattern %d ctas/SM %d: per-CTA %.1f clk/MMA (max %.1f)  -> SM-aggregate %.1f clk/MMA   kernel %.3f ms\n", random_data ? "random" : "const ", M, N,
           PATTERN, ctas_per_sm, mean / mmas, mx / mmas, mean / mmas / ctas_per_sm, ms);
    cudaFree(dc);
}

int main() {
    for (int rnd = 0; rnd <= 1; ++rnd)
        for (int c = 1; c <= 2; ++c) {
            run<128, 64, 0>(c, rnd);
            run<128, 128, 0>(c, rnd);
            run<128, 256, 0>(c, rnd);
            run<64, 64, 0>(c, rnd);
            run<64, 128, 0>(c, rnd);
            run<64, 256, 0>(c, rnd);
            run<128, 64, 1>(c, rnd);      // the sparse kernel's BN=64 pattern: N=128 then N=64
            run<128, 32, 1>(c, rnd);      // BN=32
        }
    return 0;
}
