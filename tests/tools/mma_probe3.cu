// Round-2 microbenchmark: the dense neck's per-K16 MMA pattern in isolation (operands resident in shared memory, no
// loads, one elected lane issuing from a warp-convergent loop exactly like conv2d_tma_kernel): how many clocks per
// 128x256x16 MMA do the variants of the three-product split cost on the tensor pipe itself?
//   0  one accumulator, plain SS MMAs (reference: 128 clk)
//   1  big/small accumulators, plain:            (d_big, ah, bh) (d_small, al, bh) (d_small, ah, bl)
//   2  weight-stationary, B collector (shipped): ws.fill(d_big, ah, bh) ws.lastuse(d_small, al, bh) ws(d_small, ah, bl)
//   3  A collector:                              (d_small, al, bh) a.fill(d_big, ah, bh) a.lastuse(d_small, ah, bl)
//   4  plain, grouped by accumulator over the 4 K steps: 4 x (d_big, ah, bh) then 8 x d_small
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I sa-ssd_b200/csrc tests/tools/mma_probe3.cu -o tests/tools/mma_probe3
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>

#include "tc_common.cuh"
using namespace tc;

template <int PATTERN, int N>
__global__ void __launch_bounds__(128, 1) probe(int iters, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
    constexpr uint32_t A_BYTES = 16384, B_BYTES = N * 128;
    const uint32_t bar = base + 2 * A_BYTES + 2 * B_BYTES, slot = bar + 16;
    for (int i = threadIdx.x; i < (int)(2 * A_BYTES + 2 * B_BYTES) / 16; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
        auto nxt = [&]() { h ^= h << 13; h ^= h >> 17; h ^= h << 5; return (h & 0x83FF83FFu) | 0x3C003C00u; };
        ((uint4*)bp)[i] = make_uint4(nxt(), nxt(), nxt(), nxt());
    }
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *(volatile uint32_t*)(bp + (slot - base));
    if (warp == 0) {
        constexpr uint32_t idesc = make_idesc(128, N, 0u);
        const bool leader = elect_one();
        const uint32_t ah = desc_lo(base), al = ah + (A_BYTES >> 4), bh = al + (A_BYTES >> 4), bl = bh + (B_BYTES >> 4);
        const uint32_t d_big = tmem, d_small = tmem + (N <= 256 ? N : 256) % 512;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (leader) {
                if constexpr (PATTERN == 4) {
#pragma unroll
                    for (uint32_t k = 0; k < 4; ++k) mma_f16_lo(d_big, ah + 2 * k, bh + 2 * k, idesc, 1u);
#pragma unroll
                    for (uint32_t k = 0; k < 4; ++k) { mma_f16_lo(d_small, al + 2 * k, bh + 2 * k, idesc, 1u); mma_f16_lo(d_small, ah + 2 * k, bl + 2 * k, idesc, 1u); }
                } else {
#pragma unroll
                    for (uint32_t k = 0; k < 4; ++k) {
                        const uint32_t ko = 2 * k;
                        if constexpr (PATTERN == 0) {
                            mma_f16_lo(d_big, ah + ko, bh + ko, idesc, 1u); mma_f16_lo(d_big, al + ko, bh + ko, idesc, 1u); mma_f16_lo(d_big, ah + ko, bl + ko, idesc, 1u);
                        } else if constexpr (PATTERN == 1) {
                            mma_f16_lo(d_big, ah + ko, bh + ko, idesc, 1u); mma_f16_lo(d_small, al + ko, bh + ko, idesc, 1u); mma_f16_lo(d_small, ah + ko, bl + ko, idesc, 1u);
                        } else if constexpr (PATTERN == 2) {
                            mma_f16_ws_lo<1>(d_big, ah + ko, bh + ko, idesc, 1u); mma_f16_ws_lo<2>(d_small, al + ko, bh + ko, idesc, 1u); mma_f16_ws_lo<0>(d_small, ah + ko, bl + ko, idesc, 1u);
                        } else {
                            mma_f16_lo(d_small, al + ko, bh + ko, idesc, 1u); mma_f16_acoll_lo<1>(d_big, ah + ko, bh + ko, idesc, 1u); mma_f16_acoll_lo<2>(d_small, ah + ko, bl + ko, idesc, 1u);
                        }
                    }
                }
            }
            __syncwarp();
        }
        if (leader) { mma_commit(bar); mbar_wait(bar, 0); cycles[blockIdx.x] = clock64() - t0; }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

template <int PATTERN, int N>
static void run() {
    const int iters = 200, smem = 2 * 16384 + 2 * N * 128 + 1024 + 64;
    cudaFuncSetAttribute(probe<PATTERN, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long* dc; cudaMalloc(&dc, 148 * 8);
    probe<PATTERN, N><<<148, 128, smem>>>(iters, dc);
    cudaDeviceSynchronize();
    probe<PATTERN, N><<<148, 128, smem>>>(iters, dc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("pattern %d N %d: %s\n", PATTERN, N, cudaGetErrorString(e)); return; }
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), dc, 148 * 8, cudaMemcpyDeviceToHost);
    double mean = 0; for (auto v : h) mean += v; mean /= 148;
    printf("pattern %d N=%3d: %.1f clk per MMA (%.0f clk per K=64 chunk of 12 MMAs)\n", PATTERN, N, mean / (iters * 12), mean / iters);
    cudaFree(dc);
}

int main() {
    run<0, 256>(); run<1, 256>(); run<2, 256>(); run<3, 256>(); run<4, 256>();
    run<1, 128>(); run<2, 128>(); run<3, 128>();
    return 0;
}
