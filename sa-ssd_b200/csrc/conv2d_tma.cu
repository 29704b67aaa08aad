// Dense NHWC 3x3 / 1x1 convolution + folded BatchNorm + ReLU for the BEV neck and the heads
// (BEVNet cmn.py:233-282, SSDRotateHead ssd_rotate_head.py:218-231, PSWarpHead.convs :424-429),
// tcgen05 FP16x3, with the activation operand moved by TMA.
//
// Why a second tensor-core kernel: in gconv_tc.cu the A operand is gathered by producer warps
// (LDG -> split -> STS), nine times per element for a 3x3 conv; ncu shows that kernel bound by the
// L1TEX/shared-memory pipe (85 % of peak) with the tensor pipe at 67 %.  Here
//   * activations live in HBM already split: two fp16 planes [2][B][H][W][C] (hi, lo*2048), written
//     once by the epilogue of the producing layer (or by the sparse->BEV scatter);
//   * a tile is an 8x16-pixel patch; for tap (dy,dx) and a 64-channel chunk the A operand is ONE
//     cp.async.bulk.tensor.4d box {64 ch, 16 x, 8 y, 1} at (y0+dy, x0+dx) — TMA writes it 128B-swizzled
//     straight into the UMMA layout and zero-fills out-of-image pixels (= the conv's zero padding);
//   * no producer warps: warp 4 lane 0 issues the TMA boxes (A hi, A lo) and the weight bulk copies,
//     warp 5 lane 0 issues the MMAs, warps 0-3 drain TMEM.
// The LSU / shared-memory store path is out of the main loop entirely.
#include <cuda.h>

#include "tc_common.cuh"

namespace tma {

using namespace tc;

constexpr int TILE_H = 8, TILE_W = 16;          // 128 output pixels per tile
constexpr int BKC = 64;                         // channels per chunk (one 128-byte fp16 row)
constexpr int EPI_WARPS = 4;
constexpr int THREADS2 = (EPI_WARPS + 2) * 32;  // 192
constexpr int WARP_LOAD = EPI_WARPS, WARP_ISSUE = EPI_WARPS + 1;

template <int BN>
struct Cfg2 {
    static constexpr int B_TILE_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
    static constexpr int STAGES = (BN >= 256) ? 2 : (BN >= 128 ? 3 : 4);
    static constexpr int ACC_BUFS = (4 * BN <= 512) ? 2 : 1;
    static constexpr int TMEM_COLS = (ACC_BUFS * 2 * BN < 32) ? 32 : ACC_BUFS * 2 * BN;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}

struct Conv2dArgs {
    const void* wpack;
    const float* scale;
    const float* shift;
    float* out_f32;       // [B,H,W,out_f32_stride] or null
    __half* out_split;    // [2,B,H,W,out_split_ch] or null
    int batch, H, W, cin, cout, taps, relu, out_f32_stride, out_split_ch;
};

template <int BN>
__global__ void __launch_bounds__(THREADS2, 1)
conv2d_tma_kernel(const __grid_constant__ CUtensorMap amap, const Conv2dArgs p) {
    using C = Cfg2<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar_base = base + C::STAGES * C::STAGE_BYTES;
    auto full = [&](int s) { return bar_base + 8u * s; };
    auto empty = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
    auto tmem_full = [&](int a) { return bar_base + 8u * (2 * C::STAGES + a); };
    auto tmem_empty = [&](int a) { return bar_base + 8u * (2 * C::STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 4);
    volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(base_ptr + C::STAGES * C::STAGE_BYTES + 8 * (2 * C::STAGES + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = (p.W + TILE_W - 1) / TILE_W, tiles_y = (p.H + TILE_H - 1) / TILE_H;
    const int ntiles = p.batch * tiles_y * tiles_x;
    const int kchunks = (p.cin + BKC - 1) / BKC;
    const int nchunks = p.taps * kchunks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tmem_full(a), 1); mbar_init(tmem_empty(a), EPI_WARPS); }
        fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&amap) : "memory");
    }
    if (warp == WARP_ISSUE) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == WARP_LOAD) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int b = tile / (tiles_y * tiles_x);
                const int ty = (tile / tiles_x) % tiles_y, tx = tile % tiles_x;
                const int y0 = ty * TILE_H, x0 = tx * TILE_W;
                for (int t = 0; t < p.taps; ++t) {
                    const int dy = p.taps == 9 ? t / 3 - 1 : 0, dx = p.taps == 9 ? t % 3 - 1 : 0;
                    for (int kc = 0; kc < kchunks; ++kc) {
                        mbar_wait(empty(stage), phase ^ 1u);
                        const uint32_t a_hi = base + stage * C::STAGE_BYTES, a_lo = a_hi + A_TILE_BYTES;
                        const uint32_t b_dst = a_hi + 2 * A_TILE_BYTES;
                        mbar_expect_tx(full(stage), 2 * A_TILE_BYTES + 2 * C::B_TILE_BYTES);
                        // coordinates innermost first: {channel, x, y, plane*B + b}; out-of-image pixels arrive as zeros
                        tma_load_4d(a_hi, &amap, kc * BKC, x0 + dx, y0 + dy, b, full(stage));
                        tma_load_4d(a_lo, &amap, kc * BKC, x0 + dx, y0 + dy, p.batch + b, full(stage));
                        const uint8_t* src = (const uint8_t*)p.wpack + (size_t)(t * kchunks + kc) * (2 * C::B_TILE_BYTES);
                        constexpr uint32_t kPiece = (2 * C::B_TILE_BYTES >= 16384) ? 16384u : (uint32_t)(2 * C::B_TILE_BYTES);
#pragma unroll 1
                        for (uint32_t o = 0; o < 2u * C::B_TILE_BYTES; o += kPiece)
                            bulk_g2s(b_dst + o, src + o, kPiece, full(stage));
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == WARP_ISSUE) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BM, BN, 0u /*F16*/);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                mbar_wait(tmem_empty(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_big = tmem_base + (uint32_t)(acc * 2 * BN), d_small = d_big + (uint32_t)BN;
                for (int ch = 0; ch < nchunks; ++ch) {
                    mbar_wait(full(stage), phase);
                    tc_fence_after();
                    const uint32_t a_hi = base + stage * C::STAGE_BYTES, a_lo = a_hi + A_TILE_BYTES;
                    const uint32_t b_hi = a_hi + 2 * A_TILE_BYTES, b_lo = b_hi + C::B_TILE_BYTES;
#pragma unroll
                    for (int k16 = 0; k16 < 4; ++k16) {
                        const uint32_t ko = (uint32_t)k16 * 32u;
                        const uint64_t dah = make_desc(a_hi + ko), dal = make_desc(a_lo + ko);
                        const uint64_t dbh = make_desc(b_hi + ko), dbl = make_desc(b_lo + ko);
                        mma_f16(d_small, dal, dbh, idesc, (ch | k16) ? 1u : 0u);
                        mma_f16(d_small, dah, dbl, idesc, 1u);
                        mma_f16(d_big, dah, dbh, idesc, (ch | k16) ? 1u : 0u);
                    }
                    mma_commit(empty(stage));
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
                mma_commit(tmem_full(acc));
                if (++acc == C::ACC_BUFS) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else if (warp < EPI_WARPS) {
        int acc = 0;
        uint32_t acc_phase = 0;
        const int r = warp * 32 + lane;
        const int py = r / TILE_W, px = r % TILE_W;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int b = tile / (tiles_y * tiles_x);
            const int ty = (tile / tiles_x) % tiles_y, tx = tile % tiles_x;
            const int y = ty * TILE_H + py, x = tx * TILE_W + px;
            const bool valid = y < p.H && x < p.W;
            const size_t pix = ((size_t)b * p.H + y) * p.W + x;
            if (lane == 0) mbar_wait(tmem_full(acc), acc_phase);
            __syncwarp();
            tc_fence_after();
            constexpr int CW = (BN >= 32) ? 32 : 16;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += CW) {
                uint32_t v[CW], u[CW];
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * 2 * BN + c0);
                tmem_ld<CW>(v, taddr);
                tmem_ld<CW>(u, taddr + (uint32_t)BN);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (valid) {
                    float o[CW];
#pragma unroll
                    for (int j = 0; j < CW; ++j) {
                        const int n = c0 + j;
                        const float sc = (p.scale && n < p.cout) ? __ldg(&p.scale[n]) : 1.f;
                        const float sh = (p.shift && n < p.cout) ? __ldg(&p.shift[n]) : 0.f;
                        float val = fmaf(__fadd_rn(__uint_as_float(v[j]), __uint_as_float(u[j]) * (1.f / kF16LoScale)), sc, sh);
                        if (p.relu) val = fmaxf(val, 0.f);
                        o[j] = n < p.cout ? val : 0.f;
                    }
                    if (p.out_f32) {
                        float* orow = p.out_f32 + pix * p.out_f32_stride;
#pragma unroll
                        for (int j = 0; j < CW; j += 4) {
                            const int n = c0 + j;
                            if (n + 3 < p.out_f32_stride) *(float4*)(orow + n) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
                            else
                                for (int e = 0; e < 4; ++e)
                                    if (n + e < p.out_f32_stride) orow[n + e] = o[j + e];
                        }
                    }
                    if (p.out_split) {
                        const size_t plane = (size_t)p.batch * p.H * p.W * p.out_split_ch;
                        __half* ohi = p.out_split + pix * p.out_split_ch;
                        __half* olo = ohi + plane;
#pragma unroll
                        for (int j = 0; j < CW; j += 8) {
                            const int n = c0 + j;
                            if (n + 7 < p.out_split_ch) {
                                uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
                                split_f16x2(o[j + 0], o[j + 1], h0, l0);
                                split_f16x2(o[j + 2], o[j + 3], h1, l1);
                                split_f16x2(o[j + 4], o[j + 5], h2, l2);
                                split_f16x2(o[j + 6], o[j + 7], h3, l3);
                                *(uint4*)(ohi + n) = make_uint4(h0, h1, h2, h3);
                                *(uint4*)(olo + n) = make_uint4(l0, l1, l2, l3);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty(acc));
            if (++acc == C::ACC_BUFS) { acc = 0; acc_phase ^= 1u; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == WARP_ISSUE) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

template <int BN>
static int launch2(const CUtensorMap& map, const Conv2dArgs& a, cudaStream_t stream) {
    using C = Cfg2<BN>;
    auto kern = conv2d_tma_kernel<BN>;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess)
            return SASSD_ERR_LAUNCH;
        configured = true;
    }
    const int tiles = a.batch * sassd_div_up(a.H, TILE_H) * sassd_div_up(a.W, TILE_W);
    const int grid = tiles < 148 ? tiles : 148;
    kern<<<grid, THREADS2, C::SMEM_BYTES, stream>>>(map, a);
    return sassd_check_launch();
}

}  // namespace tma

extern "C" int sassd_conv2d_f16x3(const sassd_conv2d_desc* d, const void* in_split, const void* wpack,
                                  const float* scale, const float* shift, float* out_f32, void* out_split,
                                  sassd_stream_t stream_) {
    using namespace tma;
    if (!d || !in_split || !wpack || (!out_f32 && !out_split)) return SASSD_ERR_ARG;
    if (d->batch < 1 || d->H < 1 || d->W < 1 || d->cin < 1 || d->cout < 1 || d->cout > 256) return SASSD_ERR_ARG;
    if (!(d->taps == 9 || d->taps == 1) || (d->cin_stored % 64) != 0 || d->cin_stored < d->cin) return SASSD_ERR_ARG;
    if (out_split && (d->out_split_ch % 8) != 0) return SASSD_ERR_ARG;
    if (out_f32 && (d->out_f32_stride % 4) != 0) return SASSD_ERR_ARG;
    EncodeTiledFn enc = get_encode();
    if (!enc) return SASSD_ERR_UNSUPPORTED;
    CUtensorMap map;
    // dims innermost first: channels, x, y, plane*batch ; fp16 elements
    cuuint64_t dims[4] = {(cuuint64_t)d->cin_stored, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)(2 * d->batch)};
    cuuint64_t strides[3] = {(cuuint64_t)d->cin_stored * 2, (cuuint64_t)d->W * d->cin_stored * 2,
                             (cuuint64_t)d->H * d->W * d->cin_stored * 2};
    cuuint32_t box[4] = {(cuuint32_t)BKC, (cuuint32_t)TILE_W, (cuuint32_t)TILE_H, 1u};
    cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    CUresult rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(in_split), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return SASSD_ERR_LAUNCH;
    Conv2dArgs a;
    a.wpack = wpack; a.scale = scale; a.shift = shift; a.out_f32 = out_f32; a.out_split = (__half*)out_split;
    a.batch = d->batch; a.H = d->H; a.W = d->W; a.cin = d->cin; a.cout = d->cout; a.taps = d->taps; a.relu = d->relu;
    a.out_f32_stride = d->out_f32_stride; a.out_split_ch = d->out_split_ch;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (d->cout <= 32) return launch2<32>(map, a, stream);
    if (d->cout <= 64) return launch2<64>(map, a, stream);
    if (d->cout <= 128) return launch2<128>(map, a, stream);
    return launch2<256>(map, a, stream);
}

// SparseConvTensor.dense() into the split BEV map: hi / lo*2048 fp16 planes [2,B,H,W,D*C] (channel d*C + c).
__global__ void sparse_to_bev_split_kernel(const float4* __restrict__ feat, const int4* __restrict__ coors,
                                           const int* __restrict__ d_rows, int rows_cap, int C4, int D, int H, int W,
                                           size_t plane, __half* __restrict__ bev) {
    const int rows = min(*d_rows, rows_cap);
    const long long total = (long long)rows * C4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / C4), q = (int)(i % C4);
        const int4 c = __ldg(&coors[r]);
        const float4 v = __ldg(&feat[i]);
        uint32_t h0, h1, l0, l1;
        tc::split_f16x2(v.x, v.y, h0, l0);
        tc::split_f16x2(v.z, v.w, h1, l1);
        __half* dst = bev + ((((size_t)c.x * H + c.z) * W + c.w) * (size_t)(D * C4) + (size_t)c.y * C4 + q) * 4;
        *(uint2*)dst = make_uint2(h0, h1);
        *(uint2*)(dst + plane) = make_uint2(l0, l1);
    }
}

extern "C" int sassd_sparse_to_bev_split(const float* feat, const int32_t* coors, const int32_t* d_rows, int rows_cap,
                                         int C, int D, int H, int W, int batch, void* bev_split, sassd_stream_t stream_) {
    if (!feat || !coors || !d_rows || !bev_split || (C & 3) || batch < 1) return SASSD_ERR_ARG;
    if (rows_cap <= 0) return SASSD_OK;
    const size_t plane = (size_t)batch * H * W * D * C;
    sparse_to_bev_split_kernel<<<sassd_grid((long long)rows_cap * (C / 4), 256), 256, 0, (cudaStream_t)stream_>>>(
        (const float4*)feat, (const int4*)coors, d_rows, rows_cap, C / 4, D, H, W, plane, (__half*)bev_split);
    return sassd_check_launch();
}
