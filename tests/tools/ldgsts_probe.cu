// Round-2 microbenchmark: what does the cp.async (LDGSTS) row gather of the sparse conv cost per 128-row chunk
// (2 planes x 128 B per row = 32 KB), as a function of the lane -> (row, piece) mapping, the fraction of rows present
// and whether instructions whose rows are all absent are skipped?  8 producer warps per CTA, 148 CTAs, 4 chunks in
// flight (cp.async groups), indices from shared memory like the kernel's neighbour tile.
//   map 0: lane = row            (an instruction = one 16-byte piece of 32 rows  -> 32 cache lines)
//   map 1: octet = row           (8 lanes x 16 B = one row; 4 rows per instruction ->  4 lines)
//   map 2: half-warp = row       (8 pieces x {hi, lo}; 2 rows per instruction     ->  2 lines x 2 planes)
//   skip 1: (map 1/2) the instruction is not issued when none of its rows is present (warp vote)
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tests/tools/ldgsts_probe.cu -o tests/tools/ldgsts_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cstdint>

__device__ __forceinline__ void cp16(uint32_t dst, const void* src, uint32_t nbytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
}

template <int MAP, int SKIP>
__global__ void __launch_bounds__(256, 1) probe(const __half* __restrict__ feat, size_t plane, const int* __restrict__ idx_g, int chunks,
                                                long long* cycles) {
    extern __shared__ uint8_t smem[];
    const uint32_t base = ((uint32_t)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
    int* idx_s = (int*)(smem + (base - (uint32_t)__cvta_generic_to_shared(smem)) + 4 * 32768);   // [27][128] after the stages
    for (int i = threadIdx.x; i < 27 * 128; i += 256) idx_s[i] = idx_g[(blockIdx.x % 64) * 27 * 128 + i];
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t0 = clock64();
    for (int ch = 0; ch < chunks; ++ch) {
        const uint32_t st = base + (ch & 3) * 32768;
        const int* ip = idx_s + (ch % 27) * 128;
        if (MAP == 0) {
            const int r = threadIdx.x & 127, hf = threadIdx.x >> 7;
            const int src = ip[r];
            const uint32_t nb = src >= 0 ? 16u : 0u;
            const __half* sp = feat + (size_t)(src < 0 ? 0 : src) * 64;
            const uint32_t row_off = (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int q = hf * 4 + c;
                const uint32_t off = row_off + ((q ^ (r & 7)) << 4);
                cp16(st + off, sp + q * 8, nb);
                cp16(st + 16384 + off, sp + plane + q * 8, nb);
            }
        } else if (MAP == 1) {
            const int o = lane >> 3, piece = lane & 7;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = w * 16 + j * 4 + o;
                const int src = ip[r];
                const bool pres = src >= 0;
                if (SKIP && !__any_sync(0xffffffffu, pres)) continue;
                const uint32_t off = (r >> 3) * 1024 + (r & 7) * 128 + ((piece ^ (r & 7)) << 4);
                const __half* sp = feat + (size_t)(pres ? src : 0) * 64 + piece * 8;
                cp16(st + off, sp, pres ? 16u : 0u);
                cp16(st + 16384 + off, sp + plane, pres ? 16u : 0u);
            }
        } else {
            const int h = lane >> 4, q = lane & 15, pl = q >> 3, piece = q & 7;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int r = w * 16 + j * 2 + h;
                const int src = ip[r];
                const bool pres = src >= 0;
                if (SKIP && !__any_sync(0xffffffffu, pres)) continue;
                const uint32_t off = (r >> 3) * 1024 + (r & 7) * 128 + ((piece ^ (r & 7)) << 4) + pl * 16384;
                cp16(st + off, feat + (pl ? plane : 0) + (size_t)(pres ? src : 0) * 64 + piece * 8, pres ? 16u : 0u);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 3;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

template <int MAP, int SKIP>
static void run(const __half* feat, size_t plane, const int* idx, const char* tag) {
    const int chunks = 540, smem = 4 * 32768 + 27 * 128 * 4 + 1024;
    cudaFuncSetAttribute(probe<MAP, SKIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long* dc; cudaMalloc(&dc, 148 * 8);
    probe<MAP, SKIP><<<148, 256, smem>>>(feat, plane, idx, chunks, dc);
    cudaDeviceSynchronize();
    probe<MAP, SKIP><<<148, 256, smem>>>(feat, plane, idx, chunks, dc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", tag, cudaGetErrorString(e)); return; }
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), dc, 148 * 8, cudaMemcpyDeviceToHost);
    double mean = 0; for (auto v : h) mean += v; mean /= 148;
    printf("map %d skip %d %-14s: %6.0f clk per 128-row chunk\n", MAP, SKIP, tag, mean / chunks);
    cudaFree(dc);
}

int main() {
    const int rows = 120000;
    __half* feat; cudaMalloc(&feat, (size_t)2 * rows * 64 * 2); cudaMemset(feat, 0, (size_t)2 * rows * 64 * 2);
    for (int dens = 0; dens < 3; ++dens) {
        const int pct = dens == 0 ? 35 : (dens == 1 ? 100 : 0);
        std::vector<int> idx((size_t)64 * 27 * 128);
        srand(5);
        for (size_t i = 0; i < idx.size(); ++i) {
            const int row = (int)(i % 128), tile = (int)(i / (27 * 128));
            // neighbours of adjacent rows are adjacent rows; presence is correlated over runs of ~4 rows like a sorted rulebook
            const bool pres = (rand() % 100 < pct) ? true : false;
            static bool last = false; static int run = 0;
            if (run-- <= 0) { last = pres; run = rand() % 6; }
            idx[i] = (pct == 100 || (pct && last)) ? (tile * 1800 + row + (rand() % 300)) % rows : -1;
        }
        int* d_idx; cudaMalloc(&d_idx, idx.size() * 4);
        cudaMemcpy(d_idx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice);
        char tag[32]; snprintf(tag, 32, "%d%% present", pct);
        run<0, 0>(feat, (size_t)rows * 64, d_idx, tag);
        run<1, 0>(feat, (size_t)rows * 64, d_idx, tag);
        run<1, 1>(feat, (size_t)rows * 64, d_idx, tag);
        run<2, 0>(feat, (size_t)rows * 64, d_idx, tag);
        run<2, 1>(feat, (size_t)rows * 64, d_idx, tag);
        cudaFree(d_idx);
    }
    return 0;
}
