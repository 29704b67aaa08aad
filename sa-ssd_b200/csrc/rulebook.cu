// Rulebook construction for the spconv-v1 style sparse convolutions
// (SubMConv3d / SparseConv3d call sites: mmdet/models/necks/cmn.py:139-173,197-212).
//
// Hot-path representation: neighbour table nbr[n_out, 27] (output-stationary
// gather list).  The spconv-v1 pair tables are a re-indexing of it
// (sassd_rulebook_pairs) used at the API boundary and by the parity tests.
//
//  * hash index   : open addressing on the 31-bit flattened (b,z,y,x) key; one
//                   table per resolution level, shared by the SubM rulebook of
//                   that level and by the strided conv that consumes it.
//  * strided conv : every input marks the <= 8 output cells it feeds in a
//                   bitmap over the output grid (atomicOr); a popcount scan over
//                   the bitmap words yields the output rows already sorted by
//                   flattened index (the canonical order) — no sort, no
//                   thrust::unique, no dense int32 grid as in spconv v1.
//  * nbr fill     : one thread per (output row, offset) probes the input hash;
//                   a CTA stages one 128-row tile of the table in shared memory,
//                   writes it with one bulk (TMA) store and records which of the
//                   27 offsets occur in the tile (tile mask for tap skipping).
#include "common.cuh"

__device__ __forceinline__ int flat_key(int b, int z, int y, int x, int D, int H, int W) {
    return ((b * D + z) * H + y) * W + x;
}

__global__ void hash_build_kernel(const int4* __restrict__ coors, const int* __restrict__ d_rows, int rows_cap,
                                  int D, int H, int W, int* __restrict__ keys, int* __restrict__ vals, int slots,
                                  int* __restrict__ status) {
    const int rows = min(*d_rows, rows_cap);
    const uint32_t mask = (uint32_t)slots - 1u;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
        const int4 c = __ldg(&coors[r]);
        const int key = flat_key(c.x, c.y, c.z, c.w, D, H, W);
        uint32_t s = sassd_hash32((uint32_t)key) & mask;
        int probes = 0;
        while (true) {
            int prev = atomicCAS(&keys[s], SASSD_EMPTY_KEY, key);
            if (prev == SASSD_EMPTY_KEY) { vals[s] = r; break; }
            if (prev == key) break;  // duplicate coordinate: first writer wins
            s = (s + 1) & mask;
            if (++probes >= slots) { atomicOr(status, SASSD_FLAG_HASH_FULL); break; }
        }
    }
}

extern "C" int sassd_hash_build(const int32_t* coors, const int32_t* d_rows, int rows_cap, int batch, int D, int H,
                                int W, int32_t* keys, int32_t* vals, int slots, int32_t* d_status,
                                sassd_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!coors || !d_rows || !keys || !vals || !d_status) return SASSD_ERR_ARG;
    if (slots < 2 || (slots & (slots - 1)) || slots < 2 * rows_cap) return SASSD_ERR_ARG;
    if ((long long)batch * D * H * W >= 2147483647LL) return SASSD_ERR_UNSUPPORTED;
    cudaMemsetAsync(keys, 0xff, (size_t)slots * sizeof(int), stream);
    hash_build_kernel<<<sassd_grid(rows_cap > 0 ? rows_cap : 1, 256), 256, 0, stream>>>(
        (const int4*)coors, d_rows, rows_cap, D, H, W, keys, vals, slots, d_status);
    return sassd_check_launch();
}

// nbr[o*27 + k] = row of input cell  stride*o - pad + k  (stride=1,pad=1: SubM; stride=2,pad=1: strided).
// One CTA builds the table rows of one 128-row tile (SASSD_SPCONV_TILE_ROWS) in shared memory - one thread per
// (row, offset) probe of the input hash - and writes them with ONE bulk (TMA) store: the tile's rows are a contiguous
// 13.8 KB block of the table, so the write is fully coalesced and off the LSU.  While the entries are in flight the
// warps OR together which of the 27 offsets occur in the tile at all (ballot per offset is overkill: a warp-wide
// __reduce_or of the per-lane bit, then one shared atomicOr per warp); that 27-bit tile mask is what lets the sparse
// conv skip absent taps.
#define NBR_TILE_ROWS SASSD_SPCONV_TILE_ROWS
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
#define NBR_THREADS 1024      // 3456 probes per tile: <= 4 independent (latency-bound) hash probes in flight per thread
template <int STRIDE>
__global__ void __launch_bounds__(NBR_THREADS)
nbr_fill_kernel(const int4* __restrict__ coors_out, const int* __restrict__ d_rows, int rows_cap, int D, int H, int W,
                const int* __restrict__ keys, const int* __restrict__ vals, int slots, int* __restrict__ nbr,
                int* __restrict__ tile_mask) {
    __shared__ __align__(128) int s_tile[NBR_TILE_ROWS * 27];
    __shared__ unsigned int s_mask;
    const int rows = min(*d_rows, rows_cap);
    const int ntiles = (rows + NBR_TILE_ROWS - 1) / NBR_TILE_ROWS;
    const uint32_t mask = (uint32_t)slots - 1u;
    const uint32_t s_addr = (uint32_t)__cvta_generic_to_shared(s_tile);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (threadIdx.x == 0) s_mask = 0u;
        __syncthreads();
        const int row0 = tile * NBR_TILE_ROWS;
        const int rows_here = min(NBR_TILE_ROWS, rows_cap - row0);       // table rows this tile owns (capacity)
        unsigned int my_bits = 0u;
#pragma unroll 4
        for (int t = threadIdx.x; t < rows_here * 27; t += NBR_THREADS) {
            const int o = t / 27, k = t - o * 27;
            int res = -1;
            if (row0 + o < rows) {
                const int4 c = __ldg(&coors_out[row0 + o]);
                const int kz = k / 9, ky = (k / 3) % 3, kx = k % 3;
                const int z = c.y * STRIDE - 1 + kz, y = c.z * STRIDE - 1 + ky, x = c.w * STRIDE - 1 + kx;
                if (z >= 0 && z < D && y >= 0 && y < H && x >= 0 && x < W) {
                    const int slot = sassd_hash_find(keys, mask, flat_key(c.x, z, y, x, D, H, W));
                    if (slot >= 0) res = __ldg(&vals[slot]);
                }
            }
            s_tile[t] = res;
            if (res >= 0) my_bits |= 1u << k;
        }
        my_bits = __reduce_or_sync(0xffffffffu, my_bits);
        if ((threadIdx.x & 31) == 0 && my_bits) atomicOr(&s_mask, my_bits);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic writes -> visible to the bulk copy
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t bytes = (uint32_t)(rows_here * 27 * 4);
            const uint32_t bulk = bytes & ~15u;
            int* dst = nbr + (size_t)row0 * 27;
            if (bulk) bulk_s2g(dst, s_addr, bulk);
            for (uint32_t i = bulk / 4; i < bytes / 4; ++i) dst[i] = s_tile[i];       // < 16-byte tail of a ragged table
            if (tile_mask) tile_mask[tile] = (int)s_mask;
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");            // s_tile may be overwritten
        }
        __syncthreads();
    }
}

static int nbr_grid(int rows_cap) {
    const int tiles = sassd_div_up(rows_cap, NBR_TILE_ROWS);
    return tiles < 148 * 2 ? (tiles > 0 ? tiles : 1) : 148 * 2;
}

extern "C" int sassd_rulebook_subm(const int32_t* coors, const int32_t* d_rows, int rows_cap, int D, int H, int W,
                                   const int32_t* keys, const int32_t* vals, int slots, int32_t* nbr,
                                   int32_t* tile_mask, sassd_stream_t stream_) {
    if (!coors || !d_rows || !keys || !vals || !nbr) return SASSD_ERR_ARG;
    if (rows_cap <= 0) return SASSD_OK;
    nbr_fill_kernel<1><<<nbr_grid(rows_cap), NBR_THREADS, 0, (cudaStream_t)stream_>>>(
        (const int4*)coors, d_rows, rows_cap, D, H, W, keys, vals, slots, nbr, tile_mask);
    return sassd_check_launch();
}

extern "C" int sassd_rulebook_conv_nbr(const int32_t* coors_out, const int32_t* d_rows_out, int rows_cap_out, int D,
                                       int H, int W, const int32_t* keys_in, const int32_t* vals_in, int slots_in,
                                       int32_t* nbr, int32_t* tile_mask, sassd_stream_t stream_) {
    if (!coors_out || !d_rows_out || !keys_in || !vals_in || !nbr) return SASSD_ERR_ARG;
    if (rows_cap_out <= 0) return SASSD_OK;
    nbr_fill_kernel<2><<<nbr_grid(rows_cap_out), NBR_THREADS, 0, (cudaStream_t)stream_>>>(
        (const int4*)coors_out, d_rows_out, rows_cap_out, D, H, W, keys_in, vals_in, slots_in, nbr, tile_mask);
    return sassd_check_launch();
}

// ---------------------------------------------------------------------------
// strided conv output set (k=3, s=2, p=1)
// ---------------------------------------------------------------------------
#define BM_CHUNK 1024  // bitmap words per scan chunk

// Warp-aggregated marking: a warp covers 4 input rows x 8 candidate outputs; neighbouring inputs feed the same
// output cells, so lanes that hit the same bitmap word find each other (__match_any_sync), OR their bits in the warp
// and one lane issues the atomicOr.
__global__ void __launch_bounds__(256)
conv_mark_kernel(const int4* __restrict__ coors_in, const int* __restrict__ d_rows, int rows_cap,
                 int Do, int Ho, int Wo, uint32_t* __restrict__ bitmap) {
    const int rows = min(*d_rows, rows_cap);
    const long long total = (long long)rows * 8;
    const int lane = threadIdx.x & 31;
    for (long long t0 = blockIdx.x * (long long)blockDim.x + (threadIdx.x & ~31); t0 < total;
         t0 += (long long)gridDim.x * blockDim.x) {
        const long long t = t0 + lane;
        bool ok = t < total;
        int cell = 0;
        if (ok) {
            const int r = (int)(t >> 3), v = (int)(t & 7);
            const int4 c = __ldg(&coors_in[r]);
            // i = 2*o - 1 + k, k in {0,1,2}: odd i -> o in {(i+1)/2 (k=0), (i-1)/2 (k=2)}; even i -> o = i/2 (k=1)
            int oz, oy, ox;
            {
                const int i = c.y, sel = (v >> 2) & 1;
                if (i & 1) oz = sel ? (i - 1) >> 1 : (i + 1) >> 1; else { oz = i >> 1; ok &= (sel == 0); }
                ok &= (oz >= 0 && oz < Do);
            }
            {
                const int i = c.z, sel = (v >> 1) & 1;
                if (i & 1) oy = sel ? (i - 1) >> 1 : (i + 1) >> 1; else { oy = i >> 1; ok &= (sel == 0); }
                ok &= (oy >= 0 && oy < Ho);
            }
            {
                const int i = c.w, sel = v & 1;
                if (i & 1) ox = sel ? (i - 1) >> 1 : (i + 1) >> 1; else { ox = i >> 1; ok &= (sel == 0); }
                ok &= (ox >= 0 && ox < Wo);
            }
            if (ok) cell = flat_key(c.x, oz, oy, ox, Do, Ho, Wo);
        }
        const unsigned okmask = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            const int word = cell >> 5;
            const unsigned peers = __match_any_sync(okmask, word);
            const uint32_t bits = __reduce_or_sync(peers, 1u << (cell & 31));
            if (lane == __ffs(peers) - 1) atomicOr(&bitmap[word], bits);
        }
    }
}

__global__ void __launch_bounds__(256)
bm_count_kernel(const uint32_t* __restrict__ bitmap, int nwords, int* __restrict__ chunk_count) {
    __shared__ int s_red[8];
    const int c = blockIdx.x;
    int acc = 0;
    for (int w = threadIdx.x; w < BM_CHUNK; w += 256) {
        const int idx = c * BM_CHUNK + w;
        if (idx < nwords) acc += __popc(bitmap[idx]);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < 8; ++i) t += s_red[i];
        chunk_count[c] = t;
    }
}

// single CTA: exclusive scan of the chunk counts, total -> d_rows_out
__global__ void __launch_bounds__(1024)
bm_scan_kernel(int* __restrict__ chunk_count, int nchunks, int rows_cap, int* __restrict__ d_rows_out,
               int* __restrict__ status) {
    __shared__ int s_scan[33];
    int base = 0;
    for (int c0 = 0; c0 < nchunks; c0 += 1024) {
        const int c = c0 + threadIdx.x;
        const int v = c < nchunks ? chunk_count[c] : 0;
        int total;
        const int ex = sassd_block_exscan(v, s_scan, &total);
        if (c < nchunks) chunk_count[c] = base + ex;
        base += total;
    }
    if (threadIdx.x == 0) {
        if (base > rows_cap) { atomicOr(status, SASSD_FLAG_ROWS_CAP); base = rows_cap; }
        *d_rows_out = base;
    }
}

__global__ void __launch_bounds__(256)
bm_emit_kernel(const uint32_t* __restrict__ bitmap, int nwords, const int* __restrict__ chunk_off, int Do, int Ho,
               int Wo, int rows_cap, int4* __restrict__ coors_out) {
    __shared__ int s_scan[33];
    const int c = blockIdx.x;
    int base = chunk_off[c];
    // 4 rounds of 256 consecutive words keep the output order = ascending cell index
    for (int w0 = 0; w0 < BM_CHUNK; w0 += 256) {
        const int idx = c * BM_CHUNK + w0 + threadIdx.x;
        uint32_t bits = idx < nwords ? bitmap[idx] : 0u;
        int total;
        int row = base + sassd_block_exscan(__popc(bits), s_scan, &total);
        while (bits) {
            const int bit = __ffs(bits) - 1;
            bits &= bits - 1;
            if (row < rows_cap) {
                int cell = idx * 32 + bit;
                const int x = cell % Wo; cell /= Wo;
                const int y = cell % Ho; cell /= Ho;
                const int z = cell % Do; cell /= Do;
                coors_out[row] = make_int4(cell, z, y, x);
            }
            ++row;
        }
        base += total;
    }
}

static inline size_t rb_align(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" size_t sassd_rulebook_conv_workspace_bytes(int batch, int Do, int Ho, int Wo) {
    const long long cells = (long long)batch * Do * Ho * Wo;
    const long long nwords = (cells + 31) / 32;
    const long long nchunks = (nwords + BM_CHUNK - 1) / BM_CHUNK;
    return rb_align((size_t)nwords * 4) + rb_align((size_t)nchunks * 4);
}

extern "C" int sassd_rulebook_conv_outputs(const int32_t* coors_in, const int32_t* d_rows_in, int rows_cap_in,
                                           int batch, int D, int H, int W, int32_t* coors_out, int32_t* d_rows_out,
                                           int rows_cap_out, int32_t* d_status, void* ws, size_t ws_bytes,
                                           sassd_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!coors_in || !d_rows_in || !coors_out || !d_rows_out || !d_status || !ws) return SASSD_ERR_ARG;
    const int Do = (D + 2 - 3) / 2 + 1, Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long cells = (long long)batch * Do * Ho * Wo;
    if (cells >= 2147483647LL) return SASSD_ERR_UNSUPPORTED;
    if (ws_bytes < sassd_rulebook_conv_workspace_bytes(batch, Do, Ho, Wo)) return SASSD_ERR_WORKSPACE;
    const int nwords = (int)((cells + 31) / 32);
    const int nchunks = (nwords + BM_CHUNK - 1) / BM_CHUNK;
    uint32_t* bitmap = (uint32_t*)ws;
    int* chunk = (int*)((char*)ws + rb_align((size_t)nwords * 4));
    cudaMemsetAsync(bitmap, 0, (size_t)nwords * 4, stream);
    conv_mark_kernel<<<sassd_grid((long long)(rows_cap_in > 0 ? rows_cap_in : 1) * 8, 256), 256, 0, stream>>>(
        (const int4*)coors_in, d_rows_in, rows_cap_in, Do, Ho, Wo, bitmap);
    bm_count_kernel<<<nchunks, 256, 0, stream>>>(bitmap, nwords, chunk);
    bm_scan_kernel<<<1, 1024, 0, stream>>>(chunk, nchunks, rows_cap_out, d_rows_out, d_status);
    bm_emit_kernel<<<nchunks, 256, 0, stream>>>(bitmap, nwords, chunk, Do, Ho, Wo, rows_cap_out, (int4*)coors_out);
    return sassd_check_launch();
}

// ---------------------------------------------------------------------------
// nbr -> spconv-v1 tables.  One CTA per kernel offset: ordered compaction over
// the output rows (canonical order: ascending output row).
// indice_pairs [2, 27, rows_cap] (-1 padded), indice_pair_num [27].
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
pairs_kernel(const int* __restrict__ nbr, const int* __restrict__ d_rows, int rows_cap, int* __restrict__ pairs,
             int* __restrict__ pair_num) {
    __shared__ int s_scan[33];
    const int k = blockIdx.x;
    const int rows = min(*d_rows, rows_cap);
    int* pin = pairs + (size_t)k * rows_cap;
    int* pout = pairs + (size_t)(27 + k) * rows_cap;
    int base = 0;
    for (int o0 = 0; o0 < rows; o0 += 1024) {
        const int o = o0 + threadIdx.x;
        const int i = o < rows ? nbr[(size_t)o * 27 + k] : -1;
        int total;
        const int pos = base + sassd_block_exscan(i >= 0 ? 1 : 0, s_scan, &total);
        if (i >= 0) { pin[pos] = i; pout[pos] = o; }
        base += total;
    }
    for (int p = base + threadIdx.x; p < rows_cap; p += 1024) { pin[p] = -1; pout[p] = -1; }
    if (threadIdx.x == 0) pair_num[k] = base;
}

extern "C" int sassd_rulebook_pairs(const int32_t* nbr, const int32_t* d_rows_out, int rows_cap,
                                    int32_t* indice_pairs, int32_t* indice_pair_num, sassd_stream_t stream_) {
    if (!nbr || !d_rows_out || !indice_pairs || !indice_pair_num || rows_cap <= 0) return SASSD_ERR_ARG;
    pairs_kernel<<<27, 1024, 0, (cudaStream_t)stream_>>>(nbr, d_rows_out, rows_cap, indice_pairs, indice_pair_num);
    return sassd_check_launch();
}
