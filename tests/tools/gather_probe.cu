// Round-2 microbenchmark: TMA row gather (cp.async.bulk.tensor.2d.tile::gather4) as the sparse conv's A-operand feed.
//   * correctness: do 4 gathered 128-byte rows land as rows r..r+3 of a K-major SWIZZLE_128B UMMA tile, do
//     out-of-range row indices zero-fill?  (tries boxDim[1] = 1 and 4 for the tensor map)
//   * throughput: one warp per CTA issues 64 gather4 (32 KB = one 128-row hi+lo chunk) per stage into a 4-stage ring,
//     148 CTAs; clk per chunk and bytes/clk/SM.  Compare: the cp.async path needs 2048 16-byte LDGSTS per chunk.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I sa-ssd_b200/csrc tests/tools/gather_probe.cu -o tests/tools/gather_probe -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"
using namespace tc;

__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(dst), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}

constexpr int STAGES = 4, STAGE_BYTES = 32768;

// idx [chunks][128] row indices (>= rows means "missing": expect zero fill); planes: row i of plane 1 = rows + i
__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap map, const int* __restrict__ idx, int rows,
                                                int chunks, int verify, long long* cycles, uint4* dump) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar0 = base + STAGES * STAGE_BYTES;
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (STAGES + s); };
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        fence_barrier_init();
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        int stage = 0; uint32_t phase = 0;
        const long long t0 = clock64();
        for (int ch = 0; ch < chunks; ++ch) {
            mbar_wait(empty(stage), phase ^ 1u);
            if (lane == 0) mbar_expect_tx(full(stage), STAGE_BYTES);
            __syncwarp();
            const int* ip = idx + ((size_t)(blockIdx.x * chunks + ch) % 4096) * 128 + lane * 4;
            const int4 r = *(const int4*)ip;
            const uint32_t dst = base + stage * STAGE_BYTES + lane * 512;
            gather4(dst, &map, 0, r.x, r.y, r.z, r.w, full(stage));                                   // hi plane rows
            gather4(dst + 16384, &map, 0, r.x + rows, r.y + rows, r.z + rows, r.w + rows, full(stage));   // lo plane
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (lane == 0) cycles[blockIdx.x] = clock64() - t0;
    } else if (warp == 1) {
        int stage = 0; uint32_t phase = 0;
        for (int ch = 0; ch < chunks; ++ch) {
            mbar_wait(full(stage), phase);
            if (verify && ch == 0 && blockIdx.x == 0)
                for (int i = lane; i < STAGE_BYTES / 16; i += 32) dump[i] = ((const uint4*)bp)[i];
            __syncwarp();
            if (lane == 0) mbar_arrive(empty(stage));
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int rows = 120000, C = 64, chunks = 400;
    std::vector<__half> feat((size_t)2 * rows * C);
    for (int p = 0; p < 2; ++p)
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < C; ++c) feat[((size_t)p * rows + r) * C + c] = __float2half((float)((r * 7 + c + p * 1000) % 2039));
    std::vector<int> idx((size_t)4096 * 128);
    srand(3);
    for (size_t i = 0; i < idx.size(); ++i) {
        const int near = (int)((i / 128) * 29 % rows);
        idx[i] = (rand() % 100 < 35) ? (near + rand() % 600) % rows : 2 * rows + 5;      // 35 % present, the rest missing
    }
    for (int i = 0; i < 128; ++i) idx[i] = (i % 5 == 4) ? 2 * rows + 5 : (i * 937) % rows;       // chunk 0: known pattern
    __half* d_feat; int* d_idx; long long* d_cyc; uint4* d_dump;
    cudaMalloc(&d_feat, feat.size() * 2); cudaMalloc(&d_idx, idx.size() * 4); cudaMalloc(&d_cyc, 148 * 8);
    cudaMalloc(&d_dump, STAGE_BYTES);
    cudaMemcpy(d_feat, feat.data(), feat.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(d_idx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) { printf("no encode fn\n"); return 1; }
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const int smem = STAGES * STAGE_BYTES + 1024 + 256;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int boxrows = 1; boxrows <= 4; boxrows += 3) {
        CUtensorMap map;
        cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)(2 * rows)};
        cuuint64_t strides[1] = {(cuuint64_t)C * 2};
        cuuint32_t box[2] = {(cuuint32_t)C, (cuuint32_t)boxrows};
        cuuint32_t estr[2] = {1, 1};
        CUresult rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d_feat, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("boxDim[1]=%d: encode rc=%d\n", boxrows, (int)rc);
        if (rc != CUDA_SUCCESS) continue;
        cudaMemset(d_dump, 0xff, STAGE_BYTES);
        probe<<<148, 128, smem>>>(map, d_idx, rows, chunks, 1, d_cyc, d_dump);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("  launch: %s\n", cudaGetErrorString(e)); cudaGetLastError(); continue; }
        // verify chunk 0 of CTA 0: row r of plane p at p*16384 + r*128, 16-byte piece c at ((c ^ (r & 7)) << 4)
        std::vector<__half> got(STAGE_BYTES / 2);
        cudaMemcpy(got.data(), d_dump, STAGE_BYTES, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int p = 0; p < 2 && bad < 5; ++p)
            for (int r = 0; r < 128 && bad < 5; ++r) {
                const int src = idx[r];
                for (int c = 0; c < C; ++c) {
                    const int piece = c / 8, within = c % 8;
                    const float v = __half2float(got[(size_t)p * 8192 + r * 64 + ((piece ^ (r & 7)) * 8) + within]);
                    const float exp = src < rows ? (float)((src * 7 + c + p * 1000) % 2039) : 0.f;
                    if (v != exp) { if (bad < 5) printf("  MISMATCH plane %d row %d ch %d: got %g want %g (src %d)\n", p, r, c, v, exp, src); ++bad; break; }
                }
            }
        printf("  layout/zero-fill check: %s\n", bad ? "FAILED" : "OK (rows land as a K-major SWIZZLE_128B tile, missing rows are zero)");
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        probe<<<148, 128, smem>>>(map, d_idx, rows, chunks, 0, d_cyc, d_dump);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        std::vector<long long> h(148);
        cudaMemcpy(h.data(), d_cyc, 148 * 8, cudaMemcpyDeviceToHost);
        double mean = 0; for (auto v : h) mean += v; mean /= 148;
        printf("  148 CTAs x %d chunks (128 rows x 2 planes x 128 B, 35 %% present): %.0f clk per chunk, %.1f B/clk/SM landed, kernel %.3f ms\n",
               chunks, mean / chunks, 32768.0 * chunks / mean, ms);
    }
    return 0;
}
