// Device-side helpers shared by the tcgen05 kernels (gconv_tc.cu, conv2d_tma.cu): mbarrier / bulk-copy /
// tcgen05 PTX wrappers, UMMA descriptors, TMEM loads and the hi/lo operand splits.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

#include <cstdlib>
#include <utility>

namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;                   // fp32 elements per chunk = one 128-byte swizzle row
constexpr int A_TILE_BYTES = BM * 128;   // 16 KB (hi) ; same for lo
constexpr int NUM_EPI_WARPS = 4, NUM_PROD_WARPS = 8;   // two producer threads per tile row
constexpr int THREADS = (NUM_EPI_WARPS + NUM_PROD_WARPS + 2) * 32;   // 448
constexpr int WARP_MMA = NUM_EPI_WARPS + NUM_PROD_WARPS, WARP_BLOAD = WARP_MMA + 1;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute may start
// while its predecessor in the stream is still running.  launch_dependents lets the NEXT grid be scheduled as soon
// as this grid's CTAs free their resources; wait blocks until every prerequisite grid has completed and its memory
// is visible - nothing that reads or writes global data may run before it.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// multicast variant: the same bytes land at the same shared-memory offset of every CTA in `mask` and
// complete_tx on the mbarrier at the same offset in each of them
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void mma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"(mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused for swizzled K-major) | SBO>>4 [32,46) (=1024 B between
// 8-row groups) | version=1 [46,48) | layout_type=2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// The MMA-issuing lane is an instruction-issue bottleneck long before the tensor pipe is (round-2 probe,
// profiles/r2_mma_issue_probe.md: ~6-8 clk per SASS instruction in that single-thread stream, so a loop that rebuilds
// three 64-bit descriptors per K step - ~20 instructions per UTCHMMA - runs at 100-130 clk per MMA against a 32-64
// clk tensor floor).  Hence: the constant high word and the per-tile low word of a descriptor are split, a K=16 step
// is "+2" on the low word (32 bytes >> 4; the 14-bit address field cannot carry, shared memory is < 256 KB), and the
// issue loops are fully unrolled around these.
constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);    // SBO | version 1 | SWIZZLE_128B
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
constexpr uint32_t kDescK16 = 2u;      // low-word advance per K=16 step of fp16 operands (32 bytes)

// whole-warp convergent election (CUTLASS' elect_one_sync): exactly one lane gets true, always the same one
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n" : "=r"(pred));
    return pred != 0;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, both K-major
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, uint32_t fmt) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
// kind::f16 with descriptors given as low words (high word = kDescHi); ACCUM is a compile-time flag where known
__device__ __forceinline__ void mma_f16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accum), "r"(kDescHi)
        : "memory");
}
// MODE 0 plain, 1 = fill collector b0 with B, 2 = last use of b0 (weight-stationary form, see mma_f16_ws)
template <int MODE>
__device__ __forceinline__ void mma_f16_ws_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accum) {
    if constexpr (MODE == 1)
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 da, {%1, %5};\n\t"
                     "mov.b64 db, {%2, %5};\n\t"
                     "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::fill [%0], da, db, %3, p;\n\t}\n" ::"r"(tmem_d),
                     "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accum), "r"(kDescHi) : "memory");
    else if constexpr (MODE == 2)
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 da, {%1, %5};\n\t"
                     "mov.b64 db, {%2, %5};\n\t"
                     "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::lastuse [%0], da, db, %3, p;\n\t}\n" ::"r"(tmem_d),
                     "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accum), "r"(kDescHi) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 da, {%1, %5};\n\t"
                     "mov.b64 db, {%2, %5};\n\t"
                     "tcgen05.mma.ws.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(tmem_d),
                     "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accum), "r"(kDescHi) : "memory");
}
// A-collector forms: MODE 1 = keep A after this MMA (fill), 2 = take A from the collector (lastuse)
template <int MODE>
__device__ __forceinline__ void mma_f16_acoll_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accum) {
    if constexpr (MODE == 1)
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 da, {%1, %5};\n\t"
                     "mov.b64 db, {%2, %5};\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], da, db, %3, p;\n\t}\n" ::"r"(tmem_d),
                     "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accum), "r"(kDescHi) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 da, {%1, %5};\n\t"
                     "mov.b64 db, {%2, %5};\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], da, db, %3, p;\n\t}\n" ::"r"(tmem_d),
                     "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accum), "r"(kDescHi) : "memory");
}
// Same instruction with the A-operand collector: `keep` leaves A in the tensor core's collector buffer after this
// MMA, `reuse` takes A from there instead of reading shared memory again (SASS: UTCHMMA ... A_KEEP / A_REUSE).
__device__ __forceinline__ void mma_f16_akeep(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void mma_f16_areuse(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
// Weight-stationary form: the B operand can stay in collector buffer b0 (SASS: UTCHMMA.WS ... B_KEEP / B_REUSE).
// mode 0 = plain, 1 = fill b0, 2 = last use of b0.
template <int MODE>
__device__ __forceinline__ void mma_f16_ws(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    if constexpr (MODE == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::fill [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
                     "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
    else if constexpr (MODE == 2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::lastuse [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
                     "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.ws.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
                     "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}
template <int PREC>
__device__ __forceinline__ void mma_any(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    if constexpr (PREC == 0) mma_tf32(tmem_d, da, db, idesc, accum); else mma_f16(tmem_d, da, db, idesc, accum);
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// fp32 -> tf32 with round-to-nearest (the tensor core would otherwise just drop the 13 low bits, which
// biases every product the same way); lo = tf32_rn(x - hi) is then a signed residual of <= 2^-11 |x|.
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = tf32_rn(x);
    lo = tf32_rn(x - hi);
}

template <int CW>
__device__ __forceinline__ void tmem_ld(uint32_t (&v)[CW], uint32_t taddr) {
    if constexpr (CW == 32) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
    } else {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
    }
}

// PREC 0 = 3xTF32 (kind::tf32, 32 channels per 128-byte row, K=8 per MMA),
// PREC 1 = 3xFP16 (kind::f16, 64 channels per row, K=16 per MMA, twice the MMA rate and half the operand bytes).
// FP16 split: hi = half_rn(x), lo = half_rn((x - hi) * 2048); the residual is scaled into the normal fp16 range,
// the "small" accumulator therefore carries a factor 2048 that the epilogue removes.  22 significand bits survive
// (vs 21 for the tf32 split); |x| must stay below 65504 (fp16 range) — activations of this network are O(1..100).
constexpr float kF16LoScale = 2048.f;
template <int PREC>
struct Prec {
    static constexpr int BKC = PREC == 0 ? 32 : 64;   // input channels per pipeline chunk
    static constexpr int NF4 = BKC / 4;               // float4 loads per row per chunk
    static constexpr uint32_t FMT = PREC == 0 ? 2u : 0u;   // UMMA operand format: TF32 = 2, F16 = 0
};

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void split_f16(float x, float& hi_as_float, float& lo_scaled) {
    const __half h = __float2half_rn(x);
    hi_as_float = __half2float(h);
    lo_scaled = (x - hi_as_float) * kF16LoScale;
}

// two values at once: hi pair and scaled-residual pair as packed half2 words
__device__ __forceinline__ void split_f16x2(float x, float y, uint32_t& hi2, uint32_t& lo2) {
    const __half2 h = __floats2half2_rn(x, y);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn((x - hf.x) * kF16LoScale, (y - hf.y) * kF16LoScale);
    hi2 = *reinterpret_cast<const uint32_t*>(&h);
    lo2 = *reinterpret_cast<const uint32_t*>(&l);
}

// Host: launch `kern`, optionally as a programmatic dependent of the previous kernel in the stream (every kernel launched
// through here calls pdl_wait() before touching global data, so only its prologue overlaps the predecessor's tail).
// Measured (round 2, B=1): +2 % with one step on the GPU at a time, -8 % with four captured steps in flight, where the
// early-resident CTAs of the next layer sit on SMs another frame's kernels could have used.  So it is a per-capture
// choice: sassd_set_pdl(1) while the latency graph is captured (detectors._GraphedStep), off otherwise; the environment
// variable SASSD_PDL=0/1 is the default when sassd_set_pdl was never called (experiments).
extern int g_sassd_pdl;      // -1: not set (voxelize.cu)
inline bool pdl_enabled() {
    if (g_sassd_pdl >= 0) return g_sassd_pdl != 0;
    static const bool on = [] { const char* e = getenv("SASSD_PDL"); return e && atoi(e) != 0; }();
    return on;
}
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

}  // namespace tc
