// Ruled sparse convolution (SubMConv3d / SparseConv3d / 1x1x1; spconv v1.0 indice_conv semantics, call sites
// mmdet/models/necks/cmn.py:145-173,192-231) on tcgen05 FP16x3 with the features kept in "split rows":
// two fp16 planes [2][rows_cap][C] (hi = half(x), lo = half((x - hi) * 2048)), C a multiple of 8.
//
// Why: ncu on gconv_tc.cu's TABLE mode (fp32 rows gathered through registers, split on the fly) shows the
// producers — not the tensor pipe (28 %), not L2 (14 %) — as the limit: ~250 instructions per warp and chunk at
// IPC 0.36 with long-scoreboard stalls.  Here a producer thread only *issues* eight 16-byte cp.async copies per
// chunk (zero-fill for missing neighbours and for channels beyond C) straight into the 128B-swizzled operand
// tiles; no registers are staged and no split math runs in the main loop (the producing layer's epilogue wrote
// the planes).  cp.async groups give a 3-chunk-deep gather pipeline; completion -> fence.proxy.async -> one
// mbarrier arrive per warp hands the tile to the MMA lane.
#include "tc_common.cuh"

namespace sps {

using namespace tc;

constexpr int BKC = 64;                 // channels per chunk
constexpr int EPI_WARPS = 4, PROD_WARPS = 8;
constexpr int THREADS3 = (EPI_WARPS + PROD_WARPS + 2) * 32;   // 448
constexpr int W_MMA = EPI_WARPS + PROD_WARPS, W_BLOAD = W_MMA + 1;
constexpr int DEPTH = 3;                // cp.async groups in flight per producer thread

template <int BN>
struct Cfg3 {
    static constexpr int B_TILE_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
    static constexpr int STAGES = 4;
    static constexpr int ACC_BUFS = 2;                                    // BN <= 64: 4*BN <= 512 columns
    static constexpr int TMEM_COLS = (ACC_BUFS * 2 * BN < 32) ? 32 : ACC_BUFS * 2 * BN;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct Args {
    const __half* in;       // [2][in_rows_cap][cin]
    size_t in_plane;        // elements between the hi and lo planes
    const void* wpack;
    const float* scale;
    const float* shift;
    const int* nbr;
    const int* d_rows;
    __half* out_split;      // [2][rows_cap][out_ch] or null
    size_t out_plane;
    float* out_f32;         // [rows_cap][out_f32_stride] or null
    int cin, cout, taps, rows_cap, relu, out_ch, out_f32_stride;
};

template <int TABLE, int BN>
__global__ void __launch_bounds__(THREADS3, 1) spconv_split_kernel(const Args p) {
    using C = Cfg3<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = base + C::STAGES * C::STAGE_BYTES;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    auto full_a = [&](int s) { return bar_base + 8u * s; };
    auto full_b = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
    auto empty = [&](int s) { return bar_base + 8u * (2 * C::STAGES + s); };
    auto tmem_full = [&](int a) { return bar_base + 8u * (3 * C::STAGES + a); };
    auto tmem_empty = [&](int a) { return bar_base + 8u * (3 * C::STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (3 * C::STAGES + 4);
    volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(base_ptr + C::STAGES * C::STAGE_BYTES + 8 * (3 * C::STAGES + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int M = p.d_rows ? min(__ldg(p.d_rows), p.rows_cap) : p.rows_cap;
    const int ntiles = (M + BM - 1) / BM;
    const int kchunks = (p.cin + BKC - 1) / BKC;
    const int nchunks = p.taps * kchunks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(full_a(s), PROD_WARPS);
            mbar_init(full_b(s), 1);
            mbar_init(empty(s), 1);
        }
        for (int a = 0; a < 2; ++a) { mbar_init(tmem_full(a), 1); mbar_init(tmem_empty(a), EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp >= EPI_WARPS && warp < EPI_WARPS + PROD_WARPS) {
        // ===================== A producers: cp.async gather of split rows =====================
        const int pt = threadIdx.x - EPI_WARPS * 32;
        const int r = pt & 127, hf = pt >> 7;
        const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        const uint32_t sw = (uint32_t)(r & 7);
        int stage = 0;                 // stage being issued
        uint32_t phase = 0;
        int done_stage = 0;            // oldest stage whose copies are still unsignalled
        int inflight = 0;              // committed, unsignalled groups
        auto signal_oldest = [&]() {   // caller guarantees the oldest group has completed (cp.async.wait_group)
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_a(done_stage));
            if (++done_stage == C::STAGES) done_stage = 0;
            --inflight;
        };
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int m = tile * BM + r;
            int src_next = -1;
            if (m < M) src_next = TABLE ? __ldg(&p.nbr[(size_t)m * p.taps]) : m;
            for (int t = 0; t < p.taps; ++t) {
                const int src = src_next;
                if (t + 1 < p.taps) src_next = (m < M) ? __ldg(&p.nbr[(size_t)m * p.taps + t + 1]) : -1;
                const __half* rowp = p.in + (size_t)(src < 0 ? 0 : src) * p.cin;
                for (int kc = 0; kc < kchunks; ++kc) {
                    // one lane polls the mbarrier, the warp follows (256 threads spinning on one shared-memory
                    // word slow every other barrier operation of the CTA)
                    if (lane == 0) mbar_wait(empty(stage), phase ^ 1u);
                    __syncwarp();
                    const uint32_t a_hi = base + stage * C::STAGE_BYTES + row_off, a_lo = a_hi + A_TILE_BYTES;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int q = hf * 4 + c;                        // 16-byte piece (8 channels) of the row
                        const int k = kc * BKC + q * 8;
                        const uint32_t nbytes = (src >= 0 && k < p.cin) ? 16u : 0u;   // 0 -> hardware zero fill
                        const __half* sp = rowp + (nbytes ? k : 0);
                        const uint32_t off = ((uint32_t)q ^ sw) << 4;
                        cp_async16(a_hi + off, sp, nbytes);
                        cp_async16(a_lo + off, sp + p.in_plane, nbytes);
                    }
                    cp_async_commit();
                    ++inflight;
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                    if (inflight == DEPTH) { cp_async_wait<DEPTH - 1>(); signal_oldest(); }
                }
            }
        }
        // drain
        while (inflight > 0) {
            if (inflight >= 3) cp_async_wait<2>(); else if (inflight == 2) cp_async_wait<1>(); else cp_async_wait<0>();
            signal_oldest();
        }
    } else if (warp == W_BLOAD) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int ch = 0; ch < nchunks; ++ch) {
                    mbar_wait(empty(stage), phase ^ 1u);
                    const uint32_t dst = base + stage * C::STAGE_BYTES + 2 * A_TILE_BYTES;
                    const uint8_t* src = (const uint8_t*)p.wpack + (size_t)ch * (2 * C::B_TILE_BYTES);
                    mbar_expect_tx(full_b(stage), 2 * C::B_TILE_BYTES);
                    bulk_g2s(dst, src, 2 * C::B_TILE_BYTES, full_b(stage));
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == W_MMA) {
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
                    mbar_wait(full_a(stage), phase);
                    mbar_wait(full_b(stage), phase);
                    tc_fence_after();
                    const uint32_t a_hi = base + stage * C::STAGE_BYTES, a_lo = a_hi + A_TILE_BYTES;
                    const uint32_t b_hi = a_hi + 2 * A_TILE_BYTES, b_lo = b_hi + C::B_TILE_BYTES;
                    // channels beyond cin are zero in both operands: issue only the K=16 steps that carry data
                    const int ksteps = min(4, (p.cin - (ch % kchunks) * BKC + 15) / 16);
                    for (int k16 = 0; k16 < ksteps; ++k16) {
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
    } else {
        // ===================== epilogue =====================
        int acc = 0;
        uint32_t acc_phase = 0;
        const int r = warp * 32 + lane;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            if (lane == 0) mbar_wait(tmem_full(acc), acc_phase);
            __syncwarp();
            tc_fence_after();
            const int m = tile * BM + r;
            constexpr int CW = (BN >= 32) ? 32 : 16;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += CW) {
                uint32_t v[CW], u[CW];
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * 2 * BN + c0);
                tmem_ld<CW>(v, taddr);
                tmem_ld<CW>(u, taddr + (uint32_t)BN);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (m < M) {
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
                        float* orow = p.out_f32 + (size_t)m * p.out_f32_stride;
#pragma unroll
                        for (int j = 0; j < CW; j += 4)
                            if (c0 + j + 3 < p.out_f32_stride)
                                *(float4*)(orow + c0 + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
                    }
                    if (p.out_split) {
                        __half* ohi = p.out_split + (size_t)m * p.out_ch;
                        __half* olo = ohi + p.out_plane;
#pragma unroll
                        for (int j = 0; j < CW; j += 8) {
                            const int n = c0 + j;
                            if (n + 7 < p.out_ch) {
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
    if (warp == W_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
    }
}

template <int TABLE, int BN>
static int launch3(const Args& a, cudaStream_t stream) {
    using C = Cfg3<BN>;
    auto kern = spconv_split_kernel<TABLE, BN>;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess)
            return SASSD_ERR_LAUNCH;
        configured = true;
    }
    int grid = sassd_div_up(a.rows_cap, BM);
    if (grid > 148) grid = 148;
    kern<<<grid, THREADS3, C::SMEM_BYTES, stream>>>(a);
    return sassd_check_launch();
}

template <int TABLE>
static int dispatch3(const Args& a, cudaStream_t s) {
    if (a.cout <= 16) return launch3<TABLE, 16>(a, s);
    if (a.cout <= 32) return launch3<TABLE, 32>(a, s);
    if (a.cout <= 64) return launch3<TABLE, 64>(a, s);
    return SASSD_ERR_UNSUPPORTED;
}

}  // namespace sps

extern "C" int sassd_spconv_f16x3(const sassd_spconv_desc* d, const void* in_split, const void* wpack,
                                  const float* scale, const float* shift, const int32_t* nbr, const int32_t* d_rows,
                                  void* out_split, float* out_f32, sassd_stream_t stream_) {
    if (!d || !in_split || !wpack || (!out_split && !out_f32)) return SASSD_ERR_ARG;
    if (d->cin < 8 || (d->cin & 7) || d->cout < 1 || d->cout > 64 || d->taps < 1 || d->rows_cap < 0) return SASSD_ERR_ARG;
    if (d->taps > 1 && !nbr) return SASSD_ERR_ARG;
    if (out_split && ((d->out_ch & 7) || d->out_ch < d->cout)) return SASSD_ERR_ARG;
    if (out_f32 && (d->out_f32_stride & 3)) return SASSD_ERR_ARG;
    if (d->rows_cap == 0) return SASSD_OK;
    sps::Args a;
    a.in = (const __half*)in_split; a.in_plane = (size_t)d->in_rows_cap * d->cin;
    a.wpack = wpack; a.scale = scale; a.shift = shift; a.nbr = nbr; a.d_rows = d_rows;
    a.out_split = (__half*)out_split; a.out_plane = (size_t)d->rows_cap * d->out_ch; a.out_f32 = out_f32;
    a.cin = d->cin; a.cout = d->cout; a.taps = d->taps; a.rows_cap = d->rows_cap; a.relu = d->relu;
    a.out_ch = d->out_ch; a.out_f32_stride = d->out_f32_stride;
    return d->taps > 1 ? sps::dispatch3<1>(a, (cudaStream_t)stream_) : sps::dispatch3<0>(a, (cudaStream_t)stream_);
}

// fp32 rows [rows, cin] -> split rows [2][rows_cap][cs] (cs >= cin, multiple of 8; padding channels zero)
__global__ void features_to_split_kernel(const float* __restrict__ feat, const int* __restrict__ d_rows, int rows_cap,
                                         int cin, int cs, __half* __restrict__ out) {
    const int rows = d_rows ? min(*d_rows, rows_cap) : rows_cap;
    const long long total = (long long)rows * cs;
    const size_t plane = (size_t)rows_cap * cs;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / cs), c = (int)(i % cs);
        float hi = 0.f, lo = 0.f;
        if (c < cin) tc::split_f16(feat[(size_t)r * cin + c], hi, lo);
        out[i] = __float2half_rn(hi);
        out[i + plane] = __float2half_rn(lo);
    }
}

extern "C" int sassd_features_to_split(const float* feat, const int32_t* d_rows, int rows_cap, int cin, int cs,
                                       void* out_split, sassd_stream_t stream_) {
    if (!feat || !out_split || cin < 1 || cs < cin || (cs & 7)) return SASSD_ERR_ARG;
    if (rows_cap <= 0) return SASSD_OK;
    features_to_split_kernel<<<sassd_grid((long long)rows_cap * cs, 256), 256, 0, (cudaStream_t)stream_>>>(
        feat, d_rows, rows_cap, cin, cs, (__half*)out_split);
    return sassd_check_launch();
}

// dense(): split rows [2][rows_cap][C] -> split BEV map [2][B][H][W][D*C] (channel d*C + c), 16 bytes per thread
__global__ void split_rows_to_bev_kernel(const uint4* __restrict__ feat, size_t in_plane16, const int4* __restrict__ coors,
                                         const int* __restrict__ d_rows, int rows_cap, int C8, int D, int H, int W,
                                         size_t out_plane16, uint4* __restrict__ bev) {
    const int rows = min(*d_rows, rows_cap);
    const long long total = (long long)rows * C8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / C8), q = (int)(i % C8);
        const int4 c = __ldg(&coors[r]);
        const size_t o = (((size_t)c.x * H + c.z) * W + c.w) * (size_t)(D * C8) + (size_t)c.y * C8 + q;
        bev[o] = __ldg(&feat[i]);
        bev[o + out_plane16] = __ldg(&feat[i + in_plane16]);
    }
}

extern "C" int sassd_split_rows_to_bev(const void* feat_split, const int32_t* coors, const int32_t* d_rows, int rows_cap,
                                       int C, int D, int H, int W, int batch, void* bev_split, sassd_stream_t stream_) {
    if (!feat_split || !coors || !d_rows || !bev_split || (C & 7) || batch < 1) return SASSD_ERR_ARG;
    if (rows_cap <= 0) return SASSD_OK;
    const size_t in_plane16 = (size_t)rows_cap * C / 8, out_plane16 = (size_t)batch * H * W * D * C / 8;
    split_rows_to_bev_kernel<<<sassd_grid((long long)rows_cap * (C / 8), 256), 256, 0, (cudaStream_t)stream_>>>(
        (const uint4*)feat_split, in_plane16, (const int4*)coors, d_rows, rows_cap, C / 8, D, H, W, out_plane16,
        (uint4*)bev_split);
    return sassd_check_launch();
}
