// Tile-bucketed binning: the B200-first replacement of the
//   InclusiveSum -> duplicateWithKeys -> 6-pass DeviceRadixSort -> identifyTileRanges
// chain.  Result contract is unchanged and bit-exact: per tile, the
// Gaussian ids in ascending (depth bits, id) order - which is exactly what the stable sort of
// the 64-bit (tile << 32 | depth) keys in emission order produces - plus the [start,end) ranges.
//
//   1. preprocess already counted how many Gaussians touch every tile (geom->tile_count);
//   2. tile_scan_kernel: exclusive scan of the <= 64K tile counts -> tile segments.  The
//      segments ARE the ranges; the duplicate count D is the last element (stays on device);
//   3. tile_place_kernel: every (Gaussian, tile) instance takes a slot in its tile's segment
//      with one returning atomic and stores (depth bits << 32 | id) there - the only pass
//      that writes the duplicates to HBM (8 B each);
//   4. tile_sort_kernel: ONE CTA per tile sorts its segment with a stable LSD radix sort that
//      lives entirely in shared memory (keys ping-pong between two smem buffers; ranks from
//      __match_any_sync / popc in item order), then writes the sorted ids (4 B) and, when asked,
//      the sorted 64-bit keys.  Segments larger than the smem capacity run the same passes
//      through global memory (L2 resident) - any size is handled, nothing is truncated.
// HBM traffic: 8 B write + 8 B read + 4 B write per duplicate, versus 5 x (12 + 12 + 8) B for
// the global LSD sort.  Compiled with -fmad=false (the tile rectangle must match preprocess).
#include "common.cuh"

#define TS_THREADS 256
#define TS_WARPS (TS_THREADS / 32)
#define TS_SMEM_ITEMS 4096           // 2 x 32 KB key buffers + 8 KB of per-warp digit counters
#define TS_BINS 256

// ------------------------------------------------------------ tile segments ----
__global__ void __launch_bounds__(1024) tile_scan_kernel(const uint32_t* __restrict__ tile_count, int tiles,
                                                         uint32_t* __restrict__ tile_off, uint32_t* __restrict__ num_rendered,
                                                         uint32_t cap) {
    __shared__ uint32_t ws[33];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < tiles; base += 1024) {
        const int i = base + threadIdx.x;
        uint32_t v = i < tiles ? tile_count[i] : 0;
        const uint32_t orig = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (lane == 31) ws[w] = v;
        __syncthreads();
        if (w == 0) {
            const uint32_t x = ws[lane];
            uint32_t y = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, y, o);
                if (lane >= o) y += t;
            }
            ws[lane] = y - x;
            if (lane == 31) ws[32] = y;
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        if (i < tiles) tile_off[i] = carry + ws[w] + v - orig;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + ws[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        tile_off[tiles] = carry_s;
        num_rendered[0] = carry_s;
        num_rendered[1] = carry_s > cap ? 1u : 0u;
    }
}

// ------------------------------------------------------------------ placement ----
__global__ void __launch_bounds__(RDG_BLOCK) tile_place_kernel(int64_t n, const int32_t* __restrict__ radii,
                                                               const float4* __restrict__ p0, const float2* __restrict__ p2,
                                                               int gx, int gy, const uint32_t* __restrict__ tile_off,
                                                               uint32_t* __restrict__ tile_fill, uint64_t* __restrict__ pairs,
                                                               uint32_t cap) {
    const int64_t stride = (int64_t)gridDim.x * RDG_BLOCK;
    for (int64_t i = (int64_t)blockIdx.x * RDG_BLOCK + threadIdx.x; i < n; i += stride) {
        const int r = radii[i];
        if (r <= 0) continue;
        const float4 a = p0[i];
        const float rad_f = (float)r;
        const int rminx = min(gx, max(0, (int)((a.x - rad_f) / 16.0f)));
        const int rminy = min(gy, max(0, (int)((a.y - rad_f) / 16.0f)));
        const int rmaxx = min(gx, max(0, (int)((a.x + rad_f + 15.0f) / 16.0f)));
        const int rmaxy = min(gy, max(0, (int)((a.y + rad_f + 15.0f) / 16.0f)));
        const uint64_t key = ((uint64_t)__float_as_uint(p2[i].y) << 32) | (uint32_t)i;
        for (int y = rminy; y < rmaxy; ++y) {
#pragma unroll 4
            for (int x = rminx; x < rmaxx; ++x) {
                const int t = y * gx + x;
                const uint32_t pos = tile_off[t] + atomicAdd(&tile_fill[t], 1u);
                if (pos < cap) pairs[pos] = key;
            }
        }
    }
}

// ---------------------------------------------------------------- per-tile sort ----
// One stable LSD pass over n keys, src -> dst (shared or global memory), digit = (key >> shift) & mask.
// Warp w owns the contiguous item range [w*chunk, (w+1)*chunk); items are visited in index order, so
// rank(item) = #same-digit items before it = (earlier warps) + (earlier rounds of this warp) + (lower lanes).
__device__ __forceinline__ void ts_radix_pass(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst, int n, int shift,
                                              uint32_t mask, uint32_t (*wh)[TS_BINS], uint32_t* scan_ws) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    for (int b = threadIdx.x; b < TS_WARPS * TS_BINS; b += TS_THREADS) (&wh[0][0])[b] = 0;
    __syncthreads();
    const int chunk = (((n + TS_WARPS - 1) / TS_WARPS) + 31) & ~31;
    const int beg = w * chunk, end = min(n, beg + chunk);
    // phase 1: per-warp digit counts
    for (int base = beg; base < end; base += 32) {
        const int i = base + lane;
        const bool valid = i < end;
        const uint32_t d = valid ? ((uint32_t)(src[i] >> shift) & mask) : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (valid && lane == __ffs(peers) - 1) wh[w][d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // phase 2: offset[d][w] = sum_{d'<d} total[d'] + sum_{w'<w} count[w'][d]   (TS_THREADS == TS_BINS)
    {
        const int d = threadIdx.x;
        uint32_t tot = 0;
#pragma unroll
        for (int k = 0; k < TS_WARPS; ++k) tot += wh[k][d];
        uint32_t v = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (lane == 31) scan_ws[w] = v;
        __syncthreads();
        uint32_t wbase = 0;
#pragma unroll
        for (int k = 0; k < TS_WARPS; ++k) if (k < w) wbase += scan_ws[k];
        uint32_t run = wbase + v - tot;
#pragma unroll
        for (int k = 0; k < TS_WARPS; ++k) { const uint32_t t = wh[k][d]; wh[k][d] = run; run += t; }
    }
    __syncthreads();
    // phase 3: stable scatter
    for (int base = beg; base < end; base += 32) {
        const int i = base + lane;
        const bool valid = i < end;
        const uint64_t key = valid ? src[i] : 0;
        const uint32_t d = valid ? ((uint32_t)(key >> shift) & mask) : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (valid) {
            const uint32_t off = wh[w][d];
            dst[off + __popc(peers & lt)] = key;
        }
        __syncwarp();
        if (valid && lane == __ffs(peers) - 1) wh[w][d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(TS_THREADS) tile_sort_kernel(const uint32_t* __restrict__ tile_off, int tiles, uint32_t cap,
                                                               uint64_t* __restrict__ pairs, uint64_t* __restrict__ pairs_tmp,
                                                               int id_bits, uint32_t* __restrict__ vals_sorted,
                                                               uint64_t* __restrict__ keys_sorted, uint2* __restrict__ ranges) {
    extern __shared__ __align__(16) uint64_t ts_smem[];
    uint64_t* buf0 = ts_smem;
    uint64_t* buf1 = ts_smem + TS_SMEM_ITEMS;
    uint32_t (*wh)[TS_BINS] = reinterpret_cast<uint32_t (*)[TS_BINS]>(ts_smem + 2 * TS_SMEM_ITEMS);
    __shared__ uint32_t scan_ws[TS_WARPS];
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const uint32_t beg = min(tile_off[t], cap), end = min(tile_off[t + 1], cap);
        const int n = (int)(end - beg);
        if (threadIdx.x == 0) ranges[t] = n > 0 ? make_uint2(beg, end) : make_uint2(0u, 0u);
        if (n == 0) continue;
        const bool in_smem = n <= TS_SMEM_ITEMS;
        uint64_t* a = in_smem ? buf0 : pairs + beg;
        uint64_t* b = in_smem ? buf1 : pairs_tmp + beg;
        __syncthreads();   // smem buffers free (previous tile fully written out)
        if (in_smem)
            for (int i = threadIdx.x; i < n; i += TS_THREADS) buf0[i] = pairs[beg + i];
        __syncthreads();
        if (n > 1) {
            // id bits first (low 32 bits hold the Gaussian id), then the 32 depth bits
            for (int shift = 0; shift < id_bits; shift += 8) {
                const int bits = min(8, id_bits - shift);
                ts_radix_pass(a, b, n, shift, (1u << bits) - 1u, wh, scan_ws);
                uint64_t* tmp = a; a = b; b = tmp;
            }
            for (int shift = 32; shift < 64; shift += 8) {
                ts_radix_pass(a, b, n, shift, 0xffu, wh, scan_ws);
                uint64_t* tmp = a; a = b; b = tmp;
            }
        }
        const uint64_t tile_hi = (uint64_t)t << 32;
        for (int i = threadIdx.x; i < n; i += TS_THREADS) {
            const uint64_t k = a[i];
            vals_sorted[beg + i] = (uint32_t)k;
            if (keys_sorted) keys_sorted[beg + i] = tile_hi | (k >> 32);
        }
    }
}

// ------------------------------------------------------------------------ host ----
struct TileLayout {
    int64_t tile_off, tile_fill, pairs, pairs_tmp, total;
    int tiles;
};

static TileLayout tile_layout(int64_t d_cap, int32_t height, int32_t width) {
    TileLayout L;
    L.tiles = ((width + RDG_TILE - 1) / RDG_TILE) * ((height + RDG_TILE - 1) / RDG_TILE);
    int64_t off = 0;
    L.tile_off = off;  off += rdg_align_up((int64_t)(L.tiles + 1) * 4, 256);
    L.tile_fill = off; off += rdg_align_up((int64_t)L.tiles * 4, 256);
    L.pairs = off;     off += rdg_align_up(d_cap * 8, 256);
    L.pairs_tmp = off; off += rdg_align_up(d_cap * 8, 256);
    L.total = off;
    return L;
}

extern "C" int64_t rdg_bin_tiles_workspace_bytes(int64_t n, int64_t d_cap, int32_t height, int32_t width) {
    if (n < 0 || d_cap < 0 || height <= 0 || width <= 0) return RDG_E_ARG;
    return tile_layout(d_cap, height, width).total;
}

extern "C" int rdg_bin_tiles(int64_t n, const RdgGeom* geom, int32_t height, int32_t width, int64_t d_cap,
                             const RdgBins* bins, void* workspace, int64_t workspace_bytes, void* stream) {
    RDG_CHECK_ARG(geom && bins && workspace, "null argument");
    RDG_CHECK_ARG(n >= 0 && d_cap > 0 && d_cap < (int64_t)0xffffffffLL, "bad sizes");
    RDG_CHECK_ARG(bins->vals_sorted && bins->ranges && bins->num_rendered, "null bin buffer");
    const TileLayout L = tile_layout(d_cap, height, width);
    if (workspace_bytes < L.total) {
        rdg_set_error("rdg_bin_tiles: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)L.total);
        return RDG_E_CAPACITY;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        RDG_CUDA(cudaMemsetAsync(bins->ranges, 0, (size_t)L.tiles * 2 * sizeof(uint32_t), s));
        RDG_CUDA(cudaMemsetAsync(bins->num_rendered, 0, 2 * sizeof(uint32_t), s));
        return RDG_OK;
    }
    RDG_CHECK_ARG(geom->tile_count && geom->radii && geom->p0 && geom->p2, "geom->tile_count (from rdg_preprocess_fwd) is required");
    char* ws = (char*)workspace;
    uint32_t* tile_off = (uint32_t*)(ws + L.tile_off);
    uint32_t* tile_fill = (uint32_t*)(ws + L.tile_fill);
    uint64_t* pairs = (uint64_t*)(ws + L.pairs);
    uint64_t* pairs_tmp = (uint64_t*)(ws + L.pairs_tmp);
    const int gx = (width + RDG_TILE - 1) / RDG_TILE, gy = (height + RDG_TILE - 1) / RDG_TILE;
    RDG_CUDA(cudaMemsetAsync(tile_fill, 0, (size_t)L.tiles * sizeof(uint32_t), s));
    tile_scan_kernel<<<1, 1024, 0, s>>>(geom->tile_count, L.tiles, tile_off, bins->num_rendered, (uint32_t)d_cap);
    {
        const int64_t want = (n + RDG_BLOCK - 1) / RDG_BLOCK;
        const int grid = (int)(want < (int64_t)RDG_SM_COUNT * 16 ? want : (int64_t)RDG_SM_COUNT * 16);
        tile_place_kernel<<<grid, RDG_BLOCK, 0, s>>>(n, geom->radii, (const float4*)geom->p0, (const float2*)geom->p2, gx, gy,
                                                    tile_off, tile_fill, pairs, (uint32_t)d_cap);
    }
    int id_bits = 0;
    while (id_bits < 32 && ((int64_t)1 << id_bits) < n) ++id_bits;
    if (id_bits == 0) id_bits = 1;
    const size_t smem = (size_t)2 * TS_SMEM_ITEMS * sizeof(uint64_t) + (size_t)TS_WARPS * TS_BINS * sizeof(uint32_t);
    RDG_CUDA(cudaFuncSetAttribute(tile_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = L.tiles < RDG_SM_COUNT * 32 ? L.tiles : RDG_SM_COUNT * 32;
    tile_sort_kernel<<<grid, TS_THREADS, smem, s>>>(tile_off, L.tiles, (uint32_t)d_cap, pairs, pairs_tmp, id_bits,
                                                   bins->vals_sorted, bins->keys_sorted, (uint2*)bins->ranges);
    RDG_CHECK_LAUNCH();
    rdg_count_launches(3);
    return RDG_OK;
}
