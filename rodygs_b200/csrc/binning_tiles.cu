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
//      that writes the duplicates to HBM (8 B each).  Warp-cooperative: the lanes of a warp
//      walk the concatenated instance list of its 32 Gaussians, so 32 atomics are in flight
//      per instruction regardless of how uneven the tile rectangles are;
//   4. tile_sort_kernel: ONE CTA per tile sorts its segment with a stable LSD radix sort that
//      lives entirely in shared memory (keys ping-pong between two smem buffers; ranks from
//      __match_any_sync / popc in item order).  Four 8-bit passes over the depth bits; ties in
//      depth (rare) are then put in ascending-id order by a fix-up, or - if a tile has many of
//      them - by running the id passes as well.  Two size classes (<= 1024 items: 4 warps,
//      20 KB; larger: 8 warps, 72 KB, and segments beyond 4096 items run the same passes through
//      global memory) keep ~10 tiles resident per SM.  Nothing is ever truncated.
// HBM traffic: 8 B write + 8 B read + 4 B write per duplicate, versus 5 x (12 + 12 + 8) B for
// the global LSD sort.  Compiled with -fmad=false (the tile rectangle must match preprocess).
#include "common.cuh"

#define TS_BINS 256
// One fill counter per 128-byte line.  Packed (32 counters per line) the 8160 counters of a 1080p frame are 255 lines, which
// the address hash spreads unevenly over the L2 slices: ncu r01 showed the busiest slice at 96 % while the average was
// 55 % - the placement pass was bound by its hottest slice, not by atomic latency (four requests in flight per lane
// instead of one changed nothing).
#define TP_FILL_STRIDE 32
#define TS_TIE_FIXUP_MAX 64

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
                                                               const uint32_t* __restrict__ tiles_touched,
                                                               const float4* __restrict__ p0, const float2* __restrict__ p2,
                                                               int gx, int gy, const uint32_t* __restrict__ tile_off,
                                                               uint32_t* __restrict__ tile_fill, uint64_t* __restrict__ pairs,
                                                               uint32_t cap) {
    const int lane = threadIdx.x & 31;
    const int64_t n_groups = (n + 31) / 32;
    const int64_t warps_total = (int64_t)gridDim.x * (RDG_BLOCK / 32);
    for (int64_t grp = (int64_t)blockIdx.x * (RDG_BLOCK / 32) + (threadIdx.x >> 5); grp < n_groups; grp += warps_total) {
        const int64_t i = grp * 32 + lane;
        uint32_t cnt = 0, dbits = 0;
        int rminx = 0, rminy = 0, w = 1;
        if (i < n) {
            cnt = tiles_touched[i];
            if (cnt > 0) {
                const float4 a = p0[i];
                const float rad_f = (float)radii[i];
                rminx = min(gx, max(0, (int)((a.x - rad_f) / 16.0f)));
                rminy = min(gy, max(0, (int)((a.y - rad_f) / 16.0f)));
                const int rmaxx = min(gx, max(0, (int)((a.x + rad_f + 15.0f) / 16.0f)));
                w = max(rmaxx - rminx, 1);
                dbits = __float_as_uint(p2[i].y);
            }
        }
        // warp-local inclusive scan of the instance counts
        uint32_t end = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, end, o);
            if (lane >= o) end += t;
        }
        const uint32_t start = end - cnt;
        const uint32_t total = __shfl_sync(0xffffffffu, end, 31);
        // Batches of TP_UNROLL x 32 instances: all their slot requests (returning atomics) and segment-offset loads are
        // issued before the first result is consumed.  ncu r01: with one request per lane in flight the kernel sat in
        // long-scoreboard stalls (40 cycles per issue) at 87 % occupancy - it waits for L2 atomic round trips.
        constexpr int TP_UNROLL = 4;
        for (uint32_t eb = 0; eb < total; eb += 32 * TP_UNROLL) {
            int tile[TP_UNROLL];
            uint64_t key[TP_UNROLL];
            bool act[TP_UNROLL];
#pragma unroll
            for (int u = 0; u < TP_UNROLL; ++u) {
                const uint32_t e = eb + 32 * u + lane;
                act[u] = e < total;
                const uint32_t eq = act[u] ? e : total - 1;
                int lo = 0;  // smallest lane whose end > eq
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const uint32_t probe = __shfl_sync(0xffffffffu, end, lo + step - 1);
                    if (probe <= eq) lo += step;
                }
                const uint32_t o_start = __shfl_sync(0xffffffffu, start, lo);
                const int o_minx = __shfl_sync(0xffffffffu, rminx, lo);
                const int o_miny = __shfl_sync(0xffffffffu, rminy, lo);
                const int o_w = __shfl_sync(0xffffffffu, w, lo);
                const uint32_t o_bits = __shfl_sync(0xffffffffu, dbits, lo);
                const uint32_t local = eq - o_start;
                tile[u] = (o_miny + (int)(local / o_w)) * gx + o_minx + (int)(local % o_w);
                key[u] = ((uint64_t)o_bits << 32) | (uint32_t)(grp * 32 + lo);
            }
            uint32_t slot[TP_UNROLL], base[TP_UNROLL];
#pragma unroll
            for (int u = 0; u < TP_UNROLL; ++u) {
                slot[u] = act[u] ? atomicAdd(&tile_fill[(size_t)tile[u] * TP_FILL_STRIDE], 1u) : 0u;
                base[u] = act[u] ? __ldg(tile_off + tile[u]) : 0u;
            }
#pragma unroll
            for (int u = 0; u < TP_UNROLL; ++u) {
                const uint32_t pos = base[u] + slot[u];
                if (act[u] && pos < cap) pairs[pos] = key[u];
            }
        }
    }
}

// ---------------------------------------------------------------- per-tile sort ----
// One stable LSD pass over n keys, src -> dst (shared or global memory), digit = (key >> shift) & mask.
// Warp w owns the contiguous item range [w*chunk, (w+1)*chunk); items are visited in index order, so
// rank(item) = #same-digit items before it = (earlier warps) + (earlier rounds of this warp) + (lower lanes).
template <int THREADS>
__device__ __forceinline__ void ts_radix_pass(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst, int n, int shift,
                                              uint32_t mask, uint32_t (*wh)[TS_BINS], uint32_t* scan_ws) {
    constexpr int WARPS = THREADS / 32;
    constexpr int BPT = TS_BINS / THREADS;   // bins per thread in the scan (1 or 2)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    for (int b = threadIdx.x; b < WARPS * TS_BINS; b += THREADS) (&wh[0][0])[b] = 0;
    __syncthreads();
    const int chunk = (((n + WARPS - 1) / WARPS) + 31) & ~31;
    const int beg = w * chunk, end = min(n, beg + chunk);
    const int bits = 32 - __clz(mask);
    // phase 1: per-warp digit counts (native integer shared-memory atomics; order does not matter here)
    for (int i = beg + lane; i < end; i += 32) atomicAdd(&wh[w][(uint32_t)(src[i] >> shift) & mask], 1u);
    __syncthreads();
    // phase 2: offset[d][w] = sum_{d'<d} total[d'] + sum_{w'<w} count[w'][d]; thread t owns bins BPT*t .. BPT*t+BPT-1
    {
        uint32_t tot[BPT], sum = 0;
#pragma unroll
        for (int q = 0; q < BPT; ++q) {
            tot[q] = 0;
#pragma unroll
            for (int k = 0; k < WARPS; ++k) tot[q] += wh[k][BPT * threadIdx.x + q];
            sum += tot[q];
        }
        uint32_t v = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (lane == 31) scan_ws[w] = v;
        __syncthreads();
        uint32_t run = v - sum;
#pragma unroll
        for (int k = 0; k < WARPS; ++k) if (k < w) run += scan_ws[k];
#pragma unroll
        for (int q = 0; q < BPT; ++q) {
#pragma unroll
            for (int k = 0; k < WARPS; ++k) {
                const uint32_t t = wh[k][BPT * threadIdx.x + q];
                wh[k][BPT * threadIdx.x + q] = run;
                run += t;
            }
        }
    }
    __syncthreads();
    // phase 3: stable scatter
    for (int base = beg; base < end; base += 32) {
        const int i = base + lane;
        const bool valid = i < end;
        const uint64_t key = valid ? src[i] : 0;
        const uint32_t d = (uint32_t)(key >> shift) & mask;
        const unsigned peers = rdg_match_digit(d, bits, __ballot_sync(0xffffffffu, valid));
        if (valid) dst[wh[w][d] + __popc(peers & lt)] = key;
        __syncwarp();
        if (valid && lane == __ffs(peers) - 1) wh[w][d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
}

// THREADS threads per tile; handles tiles with n_lo < n <= n_hi; segments up to SMEM_ITEMS are sorted in
// shared memory, larger ones through global memory (pairs <-> pairs_tmp).
template <int THREADS, int SMEM_ITEMS>
__global__ void __launch_bounds__(THREADS) tile_sort_kernel(const uint32_t* __restrict__ tile_off, int tiles, uint32_t cap,
                                                            int n_lo, int n_hi, uint64_t* __restrict__ pairs,
                                                            uint64_t* __restrict__ pairs_tmp, int id_bits,
                                                            uint32_t* __restrict__ vals_sorted, uint64_t* __restrict__ keys_sorted,
                                                            uint2* __restrict__ ranges) {
    constexpr int WARPS = THREADS / 32;
    extern __shared__ __align__(16) uint64_t ts_smem[];
    uint64_t* buf0 = ts_smem;
    uint64_t* buf1 = ts_smem + SMEM_ITEMS;
    uint32_t (*wh)[TS_BINS] = reinterpret_cast<uint32_t (*)[TS_BINS]>(ts_smem + 2 * SMEM_ITEMS);
    __shared__ uint32_t scan_ws[WARPS];
    __shared__ uint32_t n_ties, n_bad;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const uint32_t beg = min(tile_off[t], cap), end = min(tile_off[t + 1], cap);
        const int n = (int)(end - beg);
        if (n <= n_lo || n > n_hi) continue;
        if (threadIdx.x == 0) ranges[t] = make_uint2(beg, end);
        const bool in_smem = n <= SMEM_ITEMS;
        uint64_t* a = in_smem ? buf0 : pairs + beg;
        uint64_t* b = in_smem ? buf1 : pairs_tmp + beg;
        __syncthreads();   // smem buffers free (previous tile fully written out)
        if (threadIdx.x == 0) { n_ties = 0; n_bad = 0; }
        if (in_smem)
            for (int i = threadIdx.x; i < n; i += THREADS) buf0[i] = pairs[beg + i];
        __syncthreads();
        if (n > 1) {
            for (int shift = 32; shift < 64; shift += 8) {
                ts_radix_pass<THREADS>(a, b, n, shift, 0xffu, wh, scan_ws);
                uint64_t* tmp = a; a = b; b = tmp;
            }
            // equal depth bits: the contract (stable sort of emission order) wants ascending Gaussian id
            uint32_t my = 0, bad = 0;
            for (int i = threadIdx.x; i + 1 < n; i += THREADS)
                if ((a[i] >> 32) == (a[i + 1] >> 32)) { ++my; if ((uint32_t)a[i] > (uint32_t)a[i + 1]) ++bad; }
            if (my) atomicAdd(&n_ties, my);
            if (bad) atomicAdd(&n_bad, bad);
            __syncthreads();
            // ties == number of adjacent equal-depth pairs (bounds every run length), n_bad == out-of-order ones
            const uint32_t ties = n_bad == 0 ? 0u : n_ties;
            if (ties > TS_TIE_FIXUP_MAX) {
                // many ties (e.g. a fronto-parallel plane of Gaussians): full (depth, id) LSD sort
                for (int shift = 0; shift < id_bits; shift += 8) {
                    const int bits = min(8, id_bits - shift);
                    ts_radix_pass<THREADS>(a, b, n, shift, (1u << bits) - 1u, wh, scan_ws);
                    uint64_t* tmp = a; a = b; b = tmp;
                }
                for (int shift = 32; shift < 64; shift += 8) {
                    ts_radix_pass<THREADS>(a, b, n, shift, 0xffu, wh, scan_ws);
                    uint64_t* tmp = a; a = b; b = tmp;
                }
            } else if (ties > 0) {
                // few ties: the thread at the head of each equal-depth run insertion-sorts it by id
                for (int i = threadIdx.x; i + 1 < n; i += THREADS) {
                    const uint32_t dep = (uint32_t)(a[i] >> 32);
                    if ((i == 0 || (uint32_t)(a[i - 1] >> 32) != dep) && (uint32_t)(a[i + 1] >> 32) == dep) {
                        int j = i + 1;
                        while (j < n && (uint32_t)(a[j] >> 32) == dep) ++j;
                        for (int p = i + 1; p < j; ++p) {
                            const uint64_t k = a[p];
                            int q = p - 1;
                            while (q >= i && a[q] > k) { a[q + 1] = a[q]; --q; }
                            a[q + 1] = k;
                        }
                    }
                }
                __syncthreads();
            }
        }
        const uint64_t tile_hi = (uint64_t)t << 32;
        for (int i = threadIdx.x; i < n; i += THREADS) {
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
    L.tile_fill = off; off += rdg_align_up((int64_t)L.tiles * TP_FILL_STRIDE * 4, 256);
    L.pairs = off;     off += rdg_align_up(d_cap * 8, 256);
    L.pairs_tmp = off; off += rdg_align_up(d_cap * 8, 256);
    L.total = off;
    return L;
}

extern "C" int64_t rdg_bin_tiles_workspace_bytes(int64_t n, int64_t d_cap, int32_t height, int32_t width) {
    if (n < 0 || d_cap < 0 || height <= 0 || width <= 0) return RDG_E_ARG;
    return tile_layout(d_cap, height, width).total;
}

#define TS_SMALL_THREADS 128
#define TS_SMALL_ITEMS 1024
#define TS_BIG_THREADS 256
#define TS_BIG_ITEMS 4096

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
    RDG_CUDA(cudaMemsetAsync(bins->ranges, 0, (size_t)L.tiles * 2 * sizeof(uint32_t), s));
    if (n == 0) {
        RDG_CUDA(cudaMemsetAsync(bins->num_rendered, 0, 2 * sizeof(uint32_t), s));
        return RDG_OK;
    }
    RDG_CHECK_ARG(geom->tile_count && geom->radii && geom->tiles_touched && geom->p0 && geom->p2,
                  "geom->tile_count (from rdg_preprocess_fwd) is required");
    char* ws = (char*)workspace;
    uint32_t* tile_off = (uint32_t*)(ws + L.tile_off);
    uint32_t* tile_fill = (uint32_t*)(ws + L.tile_fill);
    uint64_t* pairs = (uint64_t*)(ws + L.pairs);
    uint64_t* pairs_tmp = (uint64_t*)(ws + L.pairs_tmp);
    const int gx = (width + RDG_TILE - 1) / RDG_TILE, gy = (height + RDG_TILE - 1) / RDG_TILE;
    RDG_CUDA(cudaMemsetAsync(tile_fill, 0, (size_t)L.tiles * TP_FILL_STRIDE * sizeof(uint32_t), s));
    tile_scan_kernel<<<1, 1024, 0, s>>>(geom->tile_count, L.tiles, tile_off, bins->num_rendered, (uint32_t)d_cap);
    {
        const int64_t groups = (n + 31) / 32;
        const int64_t want = (groups + RDG_BLOCK / 32 - 1) / (RDG_BLOCK / 32);
        const int grid = (int)(want < (int64_t)RDG_SM_COUNT * 16 ? want : (int64_t)RDG_SM_COUNT * 16);
        tile_place_kernel<<<grid, RDG_BLOCK, 0, s>>>(n, geom->radii, geom->tiles_touched, (const float4*)geom->p0,
                                                    (const float2*)geom->p2, gx, gy, tile_off, tile_fill, pairs, (uint32_t)d_cap);
    }
    int id_bits = 0;
    while (id_bits < 32 && ((int64_t)1 << id_bits) < n) ++id_bits;
    if (id_bits == 0) id_bits = 1;
    {
        const size_t smem = (size_t)2 * TS_SMALL_ITEMS * sizeof(uint64_t) + (size_t)(TS_SMALL_THREADS / 32) * TS_BINS * sizeof(uint32_t);
        const int grid = L.tiles < RDG_SM_COUNT * 16 ? L.tiles : RDG_SM_COUNT * 16;
        tile_sort_kernel<TS_SMALL_THREADS, TS_SMALL_ITEMS><<<grid, TS_SMALL_THREADS, smem, s>>>(
            tile_off, L.tiles, (uint32_t)d_cap, 0, TS_SMALL_ITEMS, pairs, pairs_tmp, id_bits, bins->vals_sorted, bins->keys_sorted,
            (uint2*)bins->ranges);
    }
    {
        const size_t smem = (size_t)2 * TS_BIG_ITEMS * sizeof(uint64_t) + (size_t)(TS_BIG_THREADS / 32) * TS_BINS * sizeof(uint32_t);
        RDG_CUDA(cudaFuncSetAttribute(tile_sort_kernel<TS_BIG_THREADS, TS_BIG_ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = L.tiles < RDG_SM_COUNT * 3 ? L.tiles : RDG_SM_COUNT * 3;
        tile_sort_kernel<TS_BIG_THREADS, TS_BIG_ITEMS><<<grid, TS_BIG_THREADS, smem, s>>>(
            tile_off, L.tiles, (uint32_t)d_cap, TS_SMALL_ITEMS, 0x7fffffff, pairs, pairs_tmp, id_bits, bins->vals_sorted,
            bins->keys_sorted, (uint2*)bins->ranges);
    }
    RDG_CHECK_LAUNCH();
    rdg_count_launches(4);
    return RDG_OK;
}
