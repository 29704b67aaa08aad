// Microbenchmark (round-2 groundwork): how long does one tcgen05.mma 128xNx16 (kind::f16) take
//   * SS mode (A and B from shared memory),
//   * TS mode (A from tensor memory, copied there by tcgen05.cp.128x256b from the same swizzled tile),
//   * TS mode with the smem->TMEM copy of the next A tile issued in the same stream,
// and do SS and TS produce the same accumulator?  One CTA per SM, one issuing lane, clock64 around the stream.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I sa-ssd_b200/csrc tests/tools/ts_probe.cu -o tests/tools/ts_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>

#include "tc_common.cuh"

using namespace tc;

__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}

// mode 0: SS, 1: TS (A copied once per chunk before its MMAs), 2: TS + copy of the "next" tile interleaved
template <int N>
__global__ void __launch_bounds__(128, 1) probe(const __half* __restrict__ a_g, const __half* __restrict__ b_g, int mode,
                                                int chunks, long long* cycles, float* d_out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t a_s = base, b_s = base + 16384, bar = base + 16384 + N * 128, slot = bar + 16;
    // fill the swizzled tiles: row r, 16-byte piece c -> (r/8)*1024 + (r%8)*128 + ((c ^ r%8) * 16)
    for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
        const int r = i >> 3, c = i & 7;
        *(uint4*)(bp + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) = ((const uint4*)a_g)[i];
    }
    for (int i = threadIdx.x; i < N * 8; i += blockDim.x) {
        const int r = i >> 3, c = i & 7;
        *(uint4*)(bp + 16384 + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) = ((const uint4*)b_g)[i];
    }
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *(volatile uint32_t*)(bp + (slot - base));
    const uint32_t d_acc = tmem, a_t0 = tmem + 256, a_t1 = tmem + 320;     // D: 256 cols; A: 32 cols per K=64 tile
    constexpr uint32_t idesc = make_idesc(128, N, 0u);
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        t0 = clock64();
        for (int ch = 0; ch < chunks; ++ch) {
            const uint32_t a_t = (ch & 1) ? a_t1 : a_t0;
            if (mode >= 1 && (mode == 1 || ch == 0))
                for (int k16 = 0; k16 < 4; ++k16) tmem_cp_128x256b(a_t + k16 * 8, make_desc(a_s + k16 * 32));
            if (mode == 2)      // next chunk's tile while this chunk's MMAs run
                for (int k16 = 0; k16 < 4; ++k16) tmem_cp_128x256b(((ch & 1) ? a_t0 : a_t1) + k16 * 8, make_desc(a_s + k16 * 32));
            for (int rep = 0; rep < 3; ++rep)
                for (int k16 = 0; k16 < 4; ++k16) {
                    const uint64_t db = make_desc(b_s + k16 * 32);
                    const uint32_t accum = (ch | rep | k16) ? 1u : 0u;
                    if (mode == 0) mma_f16(d_acc, make_desc(a_s + k16 * 32), db, idesc, accum);
                    else mma_f16_ts(d_acc, a_t + k16 * 8, db, idesc, accum);
                }
        }
        mma_commit(bar);
        mbar_wait(bar, 0);
        t1 = clock64();
        if (blockIdx.x == 0) cycles[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (blockIdx.x == 0 && d_out) {      // accumulator row r, columns 0..N-1
        const int r = threadIdx.x;
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t v[16];
            tmem_ld<16>(v, tmem + ((uint32_t)((r >> 5) * 32) << 16) + c0);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 16; ++j) d_out[r * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

template <int N>
static void run(int grid) {
    std::vector<__half> a(128 * 64), b((size_t)N * 64);
    srand(1);
    for (auto& v : a) v = __float2half((rand() % 2001 - 1000) / 1000.f);
    for (auto& v : b) v = __float2half((rand() % 2001 - 1000) / 1000.f);
    __half *da, *db;
    long long* dc;
    float* dd;
    cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dc, 8); cudaMalloc(&dd, 128 * N * 4);
    cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    const int smem = 16384 + N * 128 + 1024 + 64;
    cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> ref(128 * N), got(128 * N);
    for (int mode = 0; mode < 3; ++mode) {
        const int chunks = 200;
        probe<N><<<grid, 128, smem>>>(da, db, mode, chunks, dc, dd);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d mode %d: %s\n", N, mode, cudaGetErrorString(e)); return; }
        long long cyc = 0;
        cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(got.data(), dd, got.size() * 4, cudaMemcpyDeviceToHost);
        if (mode == 0) ref = got;
        double maxd = 0, maxv = 0;
        for (size_t i = 0; i < got.size(); ++i) {
            maxd = fmax(maxd, fabs((double)got[i] - ref[i]));
            maxv = fmax(maxv, fabs((double)ref[i]));
        }
        // host check of one element of the SS result: D[0][0] = chunks*3 * sum_k a[0][k] b[0][k]
        double s = 0;
        for (int k = 0; k < 64; ++k) s += (double)__half2float(a[k]) * __half2float(b[k]);
        printf("N=%3d grid=%3d mode %d (%s): %.1f clk per MMA (%d MMAs)  max|D - D_ss| %.3g (|D|max %.3g)  D00 %.4f expect %.4f\n", N,
               grid, mode, mode == 0 ? "SS" : (mode == 1 ? "TS, copy per chunk" : "TS, copy of next tile interleaved"),
               (double)cyc / (chunks * 12), chunks * 12, maxd, maxv, got[0], s * chunks * 3);
    }
    cudaFree(da); cudaFree(db); cudaFree(dc); cudaFree(dd);
}

int main() {
    run<256>(1);
    run<256>(148);
    run<128>(148);
    run<64>(148);
    return 0;
}
