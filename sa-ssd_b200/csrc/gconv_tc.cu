// Gathered implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05, sm_100a),
// fp32-accurate through a 3xTF32 split (SASSD_PREC_TF32X3).
//
//   out[m, :] = act( (sum_t  in[row(m,t), :] @ W[t]) * scale + shift )
//
// One persistent CTA per SM, warp-specialised:
//   warps 0-3  epilogue  : tcgen05.ld accumulator (TMEM) -> BN scale/shift, ReLU -> global rows
//   warps 4-7  A producer: each thread owns one of the tile's 128 output rows; per (tap, 32-channel
//                          chunk) it gathers the 128 B of its input row (neighbour table / 3x3
//                          window / identity), splits every fp32 into tf32 hi + lo and stores both
//                          into 128B-swizzled K-major shared-memory tiles (generic proxy ->
//                          fence.proxy.async -> mbarrier)
//   warp  8    MMA issuer: one elected lane issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8):
//                          lo*hi + hi*lo + hi*hi per K step, accumulating in TMEM across all taps
//   warp  9    B loader  : cp.async.bulk (TMA 1-D) of the pre-split, pre-swizzled weight block
// Pipelines: smem stages (full_a/full_b/empty mbarriers) and two TMEM accumulators
// (tmem_full/tmem_empty) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Why 3xTF32: tcgen05 has no fp32-input MMA and one TF32 pass (10-bit mantissa) cannot hold the
// 1e-4 parity bar across 22 layers; hi = tf32_rn(x), lo = tf32_rn(x - hi), and
// a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi leaves ~2^-21 relative error per product.
#include <cstdlib>

#include "common.cuh"

#include "tc_common.cuh"

namespace tc {

template <int MODE>
struct RowMapTC {
    const int* nbr;
    int taps, M, H, W;
    __device__ __forceinline__ int operator()(int m, int t, int x, int y) const {
        if (m >= M) return -1;
        if (MODE == SASSD_GCONV_TABLE) return __ldg(&nbr[(size_t)m * taps + t]);
        if (MODE == SASSD_GCONV_ROWS || taps == 1) return m;
        const int dy = t / 3 - 1, dx = t % 3 - 1;
        const int yy = y + dy, xx = x + dx;
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) return -1;
        return m + dy * W + dx;
    }
};

template <int BN>
struct Cfg {
    static constexpr int B_TILE_BYTES = BN * 128;                         // per hi / lo
    static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
    static constexpr int STAGES = (BN >= 256) ? 2 : (BN >= 128 ? 3 : 4);
    // per tile two fp32 accumulators of BN columns each: "big" (hi*hi) and "small" (lo*hi + hi*lo).
    // Tensor-core accumulation truncates (measured: error grows linearly with the number of
    // accumulate steps, ~2^-24 each); keeping the 2^-11-smaller terms out of the big accumulator
    // cuts its step count by 3x, and the two are summed once, in fp32 RN, by the epilogue.
    static constexpr int ACC_BUFS = (4 * BN <= 512) ? 2 : 1;             // double-buffer when TMEM allows
    static constexpr int TMEM_COLS = (ACC_BUFS * 2 * BN < 32) ? 32 : ACC_BUFS * 2 * BN;   // power of two
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// CL = thread-block cluster size (1, 2 or 4).  With CL > 1 the CTAs of a cluster walk the same
// (tap, chunk) sequence on different row tiles and share the weight stream: CTA r fetches the r-th
// 1/CL slice of every weight block and multicasts it to all CTAs of the cluster, which divides the
// L2 -> SM weight traffic (the dominant term at M=128 tiles: 64 KB of B per 16 KB of A) by CL.
template <int MODE, int BN, int CL, int PREC>
__global__ void __launch_bounds__(THREADS, 1)
gconv_tc_kernel(const float* __restrict__ in, const float* __restrict__ wpack, const float* __restrict__ scale,
                const float* __restrict__ shift, const int* __restrict__ nbr, const int* __restrict__ d_rows,
                float* __restrict__ out, int cin, int cout, int taps, int in_stride, int out_stride, int rows_cap,
                int H, int W, int relu) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128B swizzle
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar_base = base + C::STAGES * C::STAGE_BYTES;
    // barriers: full_a[S], full_b[S], empty[S], tmem_full[2], tmem_empty[2], then the TMEM address word
    auto full_a = [&](int s) { return bar_base + 8u * s; };
    auto full_b = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
    auto empty = [&](int s) { return bar_base + 8u * (2 * C::STAGES + s); };
    auto empty_a = [&](int s) { return bar_base + 8u * (3 * C::STAGES + s); };   // CTA-local release (A producers)
    auto tmem_full = [&](int a) { return bar_base + 8u * (4 * C::STAGES + a); };
    auto tmem_empty = [&](int a) { return bar_base + 8u * (4 * C::STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (4 * C::STAGES + 4);
    volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(base_ptr + C::STAGES * C::STAGE_BYTES + 8 * (4 * C::STAGES + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int M = d_rows ? min(__ldg(d_rows), rows_cap) : rows_cap;
    const int ntiles_real = (M + BM - 1) / BM;
    // every CTA of a cluster must run the same number of tile iterations (they hand each other weight
    // slices and stage-release signals); tiles past the end are computed on zero rows and never stored
    const int ntiles = CL > 1 ? ((ntiles_real + (int)gridDim.x - 1) / (int)gridDim.x) * (int)gridDim.x : ntiles_real;
    using PR = Prec<PREC>;
    const int kchunks = (cin + PR::BKC - 1) / PR::BKC;
    const int nchunks = taps * kchunks;
    const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0u;
    constexpr uint16_t kClusterMask = (uint16_t)((1u << CL) - 1u);

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(full_a(s), NUM_PROD_WARPS);   // one elected arrive per producer warp
            mbar_init(full_b(s), 1);
            mbar_init(empty(s), CL);      // one tcgen05.commit per CTA of the cluster (weight slices)
            mbar_init(empty_a(s), 1);     // this CTA's own commit (A tiles are private)
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full(a), 1);
            mbar_init(tmem_empty(a), NUM_EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();      // peers' barriers must exist before anything is multicast to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp >= NUM_EPI_WARPS && warp < NUM_EPI_WARPS + NUM_PROD_WARPS) {
        // ===================== A producers =====================
        const int pt = threadIdx.x - NUM_EPI_WARPS * 32;
        const int r = pt & 127;                           // tile row 0..127
        const int hf = pt >> 7;                           // which half of the row's 128-byte chunk this thread fills
        RowMapTC<MODE> rowmap{nbr, taps, M, H, W};
        const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        const uint32_t sw = (uint32_t)(r & 7);
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int m = tile * BM + r;
            int x = 0, y = 0;
            if (MODE == SASSD_GCONV_CONV2D) { x = m % W; y = (m / W) % H; }
            // software pipeline: the global loads of the next PF chunks are in flight (register ring) while
            // the current chunk is split, waits for its stage and is stored — the gather is latency-bound
            // (random rows out of L2), so bytes in flight are what buys throughput
            constexpr int NF4H = PR::NF4 / 2;             // float4 loads per thread per chunk
            constexpr int PF = PREC == 0 ? 3 : 2;         // prefetch distance in chunks
            float4 ring[PF][NF4H];
            bool ring_live[PF];                           // false = the slot stands for a missing neighbour (zeros)
            int f_t = 0, f_kc = 0;                        // fetch cursor (tap, channel chunk)
            int f_src = rowmap(m, 0, x, y);
            int f_src_nt = taps > 1 ? rowmap(m, 1, x, y) : -1;   // row index one tap ahead of the cursor
            auto fetch = [&](float4 (&dst)[NF4H], bool& live) {
                const float* rowp = in + (size_t)(f_src < 0 ? 0 : f_src) * in_stride;
                live = f_src >= 0;
#pragma unroll
                for (int c = 0; c < NF4H; ++c) {
                    const int k = f_kc * PR::BKC + (hf * NF4H + c) * 4;
                    dst[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (live && k < cin) dst[c] = __ldg((const float4*)(rowp + k));
                }
                if (++f_kc == kchunks) {
                    f_kc = 0;
                    ++f_t;
                    f_src = f_src_nt;
                    f_src_nt = (f_t + 1 < taps) ? rowmap(m, f_t + 1, x, y) : -1;
                }
            };
#pragma unroll
            for (int u = 0; u < PF; ++u)
                if (u < nchunks) fetch(ring[u], ring_live[u]);
            for (int ch0 = 0; ch0 < nchunks; ch0 += PF) {
#pragma unroll
                for (int u = 0; u < PF; ++u) {
                    const int ch = ch0 + u;
                    if (ch >= nchunks) break;
                    float4 (&vn)[NF4H] = ring[u];
                    uint4 ph[4], pl[4];       // this thread's 4 of the 8 16-byte chunks of the hi / lo rows
                    if (!ring_live[u]) {      // missing neighbour (65 % of the sparse (row, offset) slots): no split math
#pragma unroll
                        for (int c = 0; c < 4; ++c) { ph[c] = make_uint4(0u, 0u, 0u, 0u); pl[c] = ph[c]; }
                    } else if constexpr (PREC == 0) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            float4 hi, lo;
                            split_tf32(vn[c].x, hi.x, lo.x);
                            split_tf32(vn[c].y, hi.y, lo.y);
                            split_tf32(vn[c].z, hi.z, lo.z);
                            split_tf32(vn[c].w, hi.w, lo.w);
                            ph[c] = *reinterpret_cast<const uint4*>(&hi);
                            pl[c] = *reinterpret_cast<const uint4*>(&lo);
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {      // 8 channels (two float4) -> one 16-byte chunk of halfs
                            const float4 a = vn[2 * c], b = vn[2 * c + 1];
                            uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
                            split_f16x2(a.x, a.y, h0, l0);
                            split_f16x2(a.z, a.w, h1, l1);
                            split_f16x2(b.x, b.y, h2, l2);
                            split_f16x2(b.z, b.w, h3, l3);
                            ph[c] = make_uint4(h0, h1, h2, h3);
                            pl[c] = make_uint4(l0, l1, l2, l3);
                        }
                    }
                    if (ch + PF < nchunks) fetch(vn, ring_live[u]);        // refill this ring slot
                    if (lane == 0) mbar_wait(empty_a(stage), phase ^ 1u);   // one polling lane per warp
                    __syncwarp();
                    uint8_t* a_hi = base_ptr + stage * C::STAGE_BYTES;
                    uint8_t* a_lo = a_hi + A_TILE_BYTES;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t off = row_off + (((uint32_t)(hf * 4 + c) ^ sw) << 4);
                        *(uint4*)(a_hi + off) = ph[c];
                        *(uint4*)(a_lo + off) = pl[c];
                    }
                    // no proxy fence here (it lowers to MEMBAR.ALL.CTA and would wait for the prefetch ring's
                    // loads still in flight); the MMA lane fences after its mbarrier wait
                    __syncwarp();               // one arrive per warp (256 arrivals on one mbarrier per chunk
                    if (lane == 0) mbar_arrive(full_a(stage));   // serialise for ~1-2k cycles)
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == WARP_BLOAD) {
        // ===================== B loader (weights) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int ch = 0; ch < nchunks; ++ch) {
                    mbar_wait(empty(stage), phase ^ 1u);
                    const uint32_t dst = base + stage * C::STAGE_BYTES + 2 * A_TILE_BYTES;
                    const uint8_t* src = (const uint8_t*)wpack + (size_t)ch * (2 * C::B_TILE_BYTES);
                    mbar_expect_tx(full_b(stage), 2 * C::B_TILE_BYTES);      // all slices, own and peers'
                    if (CL == 1) {
                        // several medium-sized bulk copies in flight move a block faster than one big one
                        constexpr uint32_t kPiece = (2 * C::B_TILE_BYTES >= 8192) ? 8192u : (uint32_t)(2 * C::B_TILE_BYTES);
#pragma unroll 1
                        for (uint32_t o = 0; o < 2u * C::B_TILE_BYTES; o += kPiece)
                            bulk_g2s(dst + o, src + o, kPiece, full_b(stage));
                    } else {
                        constexpr uint32_t kSlice = 2 * C::B_TILE_BYTES / CL;
                        bulk_g2s_mc(dst + cta_rank * kSlice, src + cta_rank * kSlice, kSlice, full_b(stage), kClusterMask);
                    }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == WARP_MMA) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BM, BN, PR::FMT);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                mbar_wait(tmem_empty(acc), acc_phase ^ 1u);     // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_big = tmem_base + (uint32_t)(acc * 2 * BN), d_small = d_big + (uint32_t)BN;
                for (int ch = 0; ch < nchunks; ++ch) {
                    mbar_wait(full_a(stage), phase);
                    fence_proxy_async();        // producers' generic-proxy stores -> visible to the async proxy
                    mbar_wait(full_b(stage), phase);
                    tc_fence_after();
                    const uint32_t a_hi = base + stage * C::STAGE_BYTES, a_lo = a_hi + A_TILE_BYTES;
                    const uint32_t b_hi = a_hi + 2 * A_TILE_BYTES, b_lo = b_hi + C::B_TILE_BYTES;
#pragma unroll
                    for (int k8 = 0; k8 < 4; ++k8) {
                        const uint32_t ko = (uint32_t)k8 * 32u;     // 8 tf32 / 16 fp16 = 32 bytes along K in the swizzle row
                        const uint64_t dah = make_desc(a_hi + ko), dal = make_desc(a_lo + ko);
                        const uint64_t dbh = make_desc(b_hi + ko), dbl = make_desc(b_lo + ko);
                        mma_any<PREC>(d_small, dal, dbh, idesc, (ch | k8) ? 1u : 0u);
                        mma_any<PREC>(d_small, dah, dbl, idesc, 1u);
                        mma_any<PREC>(d_big, dah, dbh, idesc, (ch | k8) ? 1u : 0u);
                    }
                    // frees the smem stage (in every CTA of the cluster) when the MMAs retire
                    mma_commit(empty_a(stage));
                    if (CL == 1) mma_commit(empty(stage)); else mma_commit_mc(empty(stage), kClusterMask);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
                mma_commit(tmem_full(acc));                    // accumulator complete -> epilogue
                if (++acc == C::ACC_BUFS) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ===================== epilogue (warps 0-3; TMEM lane quarter = warp id) =====================
        int acc = 0;
        uint32_t acc_phase = 0;
        const int r = warp * 32 + lane;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            if (lane == 0) mbar_wait(tmem_full(acc), acc_phase);
            __syncwarp();
            tc_fence_after();
            const int m = tile * BM + r;
            float* orow = out + (size_t)m * out_stride;
            constexpr int CW = (BN >= 32) ? 32 : 16;           // columns per tcgen05.ld
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += CW) {
                uint32_t v[CW], u[CW];
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * 2 * BN + c0);
                tmem_ld<CW>(v, taddr);
                tmem_ld<CW>(u, taddr + (uint32_t)BN);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (m < M) {
#pragma unroll
                    for (int j = 0; j < CW; j += 4) {
                        const int n = c0 + j;
                        if (n >= cout) break;
                        float o[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int ne = n + e;
                            const float sc = (scale && ne < cout) ? __ldg(&scale[ne]) : 1.f;
                            const float sh = (shift && ne < cout) ? __ldg(&shift[ne]) : 0.f;
                            const float small = PREC == 0 ? __uint_as_float(u[j + e])
                                                          : __uint_as_float(u[j + e]) * (1.f / kF16LoScale);
                            float val = fmaf(__fadd_rn(__uint_as_float(v[j + e]), small), sc, sh);
                            if (relu) val = fmaxf(val, 0.f);
                            o[e] = val;
                        }
                        if (n + 3 < cout && (out_stride & 3) == 0) {
                            *(float4*)(orow + n) = make_float4(o[0], o[1], o[2], o[3]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (n + e < cout) orow[n + e] = o[e];
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
    if (CL > 1) cluster_sync_all();      // no CTA may exit while a peer can still signal its barriers
    if (warp == WARP_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
    }
}

template <int MODE, int BN, int CL, int PREC>
static int launch_cl(const sassd_gconv_desc* d, const float* in, const float* w, const float* scale, const float* shift,
                     const int* nbr, const int* d_rows, float* out, cudaStream_t stream, int grid) {
    using C = Cfg<BN>;
    auto kern = gconv_tc_kernel<MODE, BN, CL, PREC>;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess)
            return SASSD_ERR_LAUNCH;
        configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, in, w, scale, shift, nbr, d_rows, out, d->cin, d->cout, d->taps, d->in_stride,
                           d->out_stride, d->rows_cap, d->H, d->W, d->relu) != cudaSuccess)
        return SASSD_ERR_LAUNCH;
    return sassd_check_launch();
}

template <int MODE, int BN, int PREC>
static int launch(const sassd_gconv_desc* d, const float* in, const float* w, const float* scale, const float* shift,
                  const int* nbr, const int* d_rows, float* out, cudaStream_t stream) {
    const int tiles = sassd_div_up(d->rows_cap, BM);
    // Dense maps with at least one tile per SM can share the weight stream inside a cluster (CTA r multicasts the
    // r-th slice of every weight block).  Measured on B200 (round 1, 3xTF32, BEV 3x3 256->256): cluster 2 = same
    // time as independent CTAs, cluster 4 = 1.8x slower (lock-step stage release with 2 stages) — the kernel is
    // MMA/shared-memory bound, not L2 bound — so it stays opt-in: SASSD_TC_CLUSTER=2.
    if constexpr (MODE == SASSD_GCONV_CONV2D && BN == 256) {
        static int cl = -1;
        if (cl < 0) { const char* e = getenv("SASSD_TC_CLUSTER"); cl = e ? atoi(e) : 1; }
        if (cl == 2 && tiles >= 148)
            return launch_cl<MODE, BN, 2, PREC>(d, in, w, scale, shift, nbr, d_rows, out, stream, 148);
    }
    return launch_cl<MODE, BN, 1, PREC>(d, in, w, scale, shift, nbr, d_rows, out, stream, tiles < 148 ? tiles : 148);
}

template <int MODE, int PREC>
static int dispatch(const sassd_gconv_desc* d, const float* in, const float* w, const float* scale, const float* shift,
                    const int* nbr, const int* d_rows, float* out, cudaStream_t s) {
    if (d->cout <= 16) return launch<MODE, 16, PREC>(d, in, w, scale, shift, nbr, d_rows, out, s);
    if (d->cout <= 32) return launch<MODE, 32, PREC>(d, in, w, scale, shift, nbr, d_rows, out, s);
    if (d->cout <= 64) return launch<MODE, 64, PREC>(d, in, w, scale, shift, nbr, d_rows, out, s);
    if (d->cout <= 128) return launch<MODE, 128, PREC>(d, in, w, scale, shift, nbr, d_rows, out, s);
    if (d->cout <= 256) return launch<MODE, 256, PREC>(d, in, w, scale, shift, nbr, d_rows, out, s);
    return SASSD_ERR_UNSUPPORTED;
}

template <int PREC>
static int dispatch_mode(const sassd_gconv_desc* d, const float* in, const float* w, const float* scale,
                         const float* shift, const int* nbr, const int* d_rows, float* out, cudaStream_t s) {
    switch (d->mode) {
        case SASSD_GCONV_TABLE: return dispatch<SASSD_GCONV_TABLE, PREC>(d, in, w, scale, shift, nbr, d_rows, out, s);
        case SASSD_GCONV_CONV2D: return dispatch<SASSD_GCONV_CONV2D, PREC>(d, in, w, scale, shift, nbr, d_rows, out, s);
        case SASSD_GCONV_ROWS: return dispatch<SASSD_GCONV_ROWS, PREC>(d, in, w, scale, shift, nbr, d_rows, out, s);
    }
    return SASSD_ERR_ARG;
}

}  // namespace tc

// `weight` for these paths is the pre-split, pre-swizzled pack produced by sassd_gconv_pack:
//   TF32X3: [taps*ceil(cin/32)][hi|lo][BN rows (n)][32 fp32]   F16X3: [taps*ceil(cin/64)][hi|lo][BN][64 fp16]
// every row is 128 bytes, its 16-byte chunks XOR-swizzled by (n & 7).
int sassd_gconv_tc(const sassd_gconv_desc* d, const float* in, const float* weight, const float* scale,
                   const float* shift, const int32_t* nbr, const int32_t* d_rows, float* out, cudaStream_t stream) {
    if (d->precision == SASSD_PREC_TF32X3) return tc::dispatch_mode<0>(d, in, weight, scale, shift, nbr, d_rows, out, stream);
    if (d->precision == SASSD_PREC_F16X3) return tc::dispatch_mode<1>(d, in, weight, scale, shift, nbr, d_rows, out, stream);
    return SASSD_ERR_ARG;
}

static inline int tc_bn(int cout) { return cout <= 16 ? 16 : cout <= 32 ? 32 : cout <= 64 ? 64 : cout <= 128 ? 128 : 256; }

// Weight packer (device): W [taps, cin, cout] fp32 -> the layouts above.  One thread per packed element.
template <int PREC>
__global__ void pack_kernel(const float* __restrict__ w, int taps, int cin, int cout, int bn, void* __restrict__ out_) {
    constexpr int BKC = tc::Prec<PREC>::BKC;
    constexpr int EPC = PREC == 0 ? 4 : 8;                // elements per 16-byte chunk
    const int kchunks = (cin + BKC - 1) / BKC;
    const long long per_chunk = 2LL * bn * BKC;
    const long long total = (long long)taps * kchunks * per_chunk;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long ch = i / per_chunk;
        long long rem = i % per_chunk;
        const int part = (int)(rem / (bn * BKC));         // 0 = hi, 1 = lo
        rem %= (bn * BKC);
        const int n = (int)(rem / BKC);
        const int pos = (int)(rem % BKC);                 // physical element slot inside the 128-byte row
        const int chunk16 = pos / EPC, e = pos % EPC;
        const int kk = ((chunk16 ^ (n & 7)) * EPC) + e;   // logical k stored at this physical slot
        const int t = (int)(ch / kchunks), kc = (int)(ch % kchunks);
        const int k = kc * BKC + kk;
        float v = 0.f;
        if (k < cin && n < cout) v = w[((size_t)t * cin + k) * cout + n];
        if constexpr (PREC == 0) {
            float hi, lo;
            tc::split_tf32(v, hi, lo);
            ((float*)out_)[i] = part == 0 ? hi : lo;
        } else {
            float hi, lo;
            tc::split_f16(v, hi, lo);
            ((__half*)out_)[i] = __float2half_rn(part == 0 ? hi : lo);
        }
    }
}

extern "C" size_t sassd_gconv_pack_bytes(int taps, int cin, int cout, int precision) {
    const int bn = tc_bn(cout);
    if (precision == SASSD_PREC_TF32X3) return (size_t)taps * ((cin + 31) / 32) * 2 * bn * 128;
    if (precision == SASSD_PREC_F16X3) return (size_t)taps * ((cin + 63) / 64) * 2 * bn * 128;
    return 0;
}

extern "C" int sassd_gconv_pack(const float* weight, int taps, int cin, int cout, int precision, void* packed,
                                sassd_stream_t stream_) {
    if (!weight || !packed || taps <= 0 || cin <= 0 || cout <= 0 || cout > 256) return SASSD_ERR_ARG;
    const int bn = tc_bn(cout);
    if (precision == SASSD_PREC_TF32X3) {
        const long long total = (long long)taps * ((cin + 31) / 32) * 2 * bn * 32;
        pack_kernel<0><<<sassd_grid(total, 256), 256, 0, (cudaStream_t)stream_>>>(weight, taps, cin, cout, bn, packed);
    } else if (precision == SASSD_PREC_F16X3) {
        const long long total = (long long)taps * ((cin + 63) / 64) * 2 * bn * 64;
        pack_kernel<1><<<sassd_grid(total, 256), 256, 0, (cudaStream_t)stream_>>>(weight, taps, cin, cout, bn, packed);
    } else {
        return SASSD_ERR_ARG;
    }
    return sassd_check_launch();
}
