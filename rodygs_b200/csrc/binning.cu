// Tile binning: inclusive scan of tiles_touched, duplicateWithKeys with 64-bit
// (tile << 32 | depth bits) keys, stable LSD radix sort, identifyTileRanges.
// SURVEY.md §8 row a8; spec: SURVEY.md App. A.3 == oracle/splat_oracle.py::bin_tiles.
// Integer work, HBM-bound; the contract is bit-exact keys / order / ranges.
//
// Design notes (B200):
//  * no host synchronisation: the duplicate count D stays on the device
//    (bins.num_rendered); every kernel is launched for the capacity d_cap and
//    reads D itself, so the whole forward is CUDA-graph capturable.
//  * duplicateWithKeys is warp-cooperative: the 32 Gaussians of a warp own one
//    contiguous output span (their offsets are a prefix sum), which the lanes fill
//    with fully coalesced 8-byte / 4-byte stores after a shuffle binary search.
//  * the radix sort uses up to 9-bit digits so that 32 + ceil(log2(tiles)) <= 45
//    significant bits take 5 passes instead of cub's 6 at 1080p.
// Compiled with -fmad=false (the tile rectangle is recomputed here and must match
// preprocess bit for bit).
#include "common.cuh"

#define SCAN_ITEMS 8
#define SCAN_TILE (RDG_BLOCK * SCAN_ITEMS)
#define SORT_ITEMS 16
#define SORT_TILE (RDG_BLOCK * SORT_ITEMS)
#define SORT_MAX_BITS 9
#define SORT_MAX_BINS (1 << SORT_MAX_BITS)
#define NWARPS (RDG_BLOCK / 32)

// ---------------------------------------------------------------- scan ----
__device__ __forceinline__ uint32_t block_scan_incl(uint32_t v, uint32_t* ws /*[NWARPS+1]*/, uint32_t& total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    __syncthreads();
    if (lane == 31) ws[w] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int k = 0; k < NWARPS; ++k) { uint32_t t = ws[k]; ws[k] = run; run += t; }
        ws[NWARPS] = run;
    }
    __syncthreads();
    total = ws[NWARPS];
    return v + ws[w];
}

__global__ void __launch_bounds__(RDG_BLOCK) scan_reduce_kernel(const uint32_t* __restrict__ in, int64_t n,
                                                                uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t ws[NWARPS + 1];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + (int64_t)k * RDG_BLOCK + threadIdx.x;
        if (i < n) s += in[i];
    }
    uint32_t total;
    block_scan_incl(s, ws, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of the block sums (in place); optional total / overflow flag
__global__ void __launch_bounds__(1024) scan_sums_kernel(uint32_t* __restrict__ sums, int nb,
                                                         uint32_t* __restrict__ total_out, uint32_t cap) {
    __shared__ uint32_t ws[33];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 1024) {
        int i = base + threadIdx.x;
        uint32_t v = i < nb ? sums[i] : 0, orig = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (lane == 31) ws[w] = v;
        __syncthreads();
        if (w == 0) {
            uint32_t x = ws[lane], y = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, y, o);
                if (lane >= o) y += t;
            }
            ws[lane] = y - x;
            if (lane == 31) ws[32] = y;
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        if (i < nb) sums[i] = carry + ws[w] + v - orig;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + ws[32];
        __syncthreads();
    }
    if (total_out && threadIdx.x == 0) {
        total_out[0] = carry_s;
        total_out[1] = carry_s > cap ? 1u : 0u;
    }
}

template <bool INCLUSIVE>
__global__ void __launch_bounds__(RDG_BLOCK) scan_apply_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                               int64_t n, const uint32_t* __restrict__ block_offsets) {
    __shared__ uint32_t ws[NWARPS + 1];
    // blocked arrangement: thread t owns SCAN_ITEMS consecutive items
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    uint32_t total;
    uint32_t incl = block_scan_incl(s, ws, total);
    uint32_t run = block_offsets[blockIdx.x] + incl - s;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (INCLUSIVE) { run += v[k]; if (base + k < n) out[base + k] = run; }
        else { if (base + k < n) out[base + k] = run; run += v[k]; }
    }
}

// out may alias in.  sums_ws needs ceil(n / SCAN_TILE) uint32.
template <bool INCLUSIVE>
static int scan_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* sums_ws, uint32_t* total_out,
                    uint32_t cap, cudaStream_t s) {
    if (n <= 0) return RDG_OK;
    const int nb = rdg_div_up(n, SCAN_TILE);
    scan_reduce_kernel<<<nb, RDG_BLOCK, 0, s>>>(in, n, sums_ws);
    scan_sums_kernel<<<1, 1024, 0, s>>>(sums_ws, nb, total_out, cap);
    scan_apply_kernel<INCLUSIVE><<<nb, RDG_BLOCK, 0, s>>>(in, out, n, sums_ws);
    RDG_CHECK_LAUNCH();
    rdg_count_launches(3);
    return RDG_OK;
}

// ------------------------------------------------------ duplicateWithKeys ----
__global__ void __launch_bounds__(RDG_BLOCK) duplicate_kernel(int64_t n, const int32_t* __restrict__ radii,
                                                              const uint32_t* __restrict__ tiles_touched,
                                                              const uint32_t* __restrict__ offsets,
                                                              const float4* __restrict__ p0, const float2* __restrict__ p2,
                                                              int gx, int gy, uint64_t* __restrict__ keys,
                                                              uint32_t* __restrict__ vals, uint32_t d_cap) {
    const int lane = threadIdx.x & 31;
    const int64_t n_groups = (n + 31) / 32;
    const int64_t warps_total = (int64_t)gridDim.x * NWARPS;
    for (int64_t grp = (int64_t)blockIdx.x * NWARPS + (threadIdx.x >> 5); grp < n_groups; grp += warps_total) {
        const int64_t i = grp * 32 + lane;
        uint32_t end = 0, cnt = 0, dbits = 0;
        int rminx = 0, rminy = 0, w = 1;
        if (i < n) {
            end = offsets[i];
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
        // lanes past n inherit the last valid end so that `end` stays non-decreasing
        const uint32_t last_end = __reduce_max_sync(0xffffffffu, end);
        if (i >= n) end = last_end;
        const uint32_t start = end - cnt;
        const uint32_t span_lo = __shfl_sync(0xffffffffu, start, 0);
        const uint32_t span_hi = last_end;
        for (uint32_t eb = span_lo; eb < span_hi; eb += 32) {
            const uint32_t e = eb + lane;
            const bool act = e < span_hi;
            const uint32_t eq = act ? e : span_hi - 1;
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
            if (act && e < d_cap) {
                const uint32_t local = e - o_start;
                const uint32_t ty = o_miny + local / o_w, tx = o_minx + local % o_w;
                keys[e] = ((uint64_t)(ty * gx + tx) << 32) | o_bits;
                vals[e] = (uint32_t)(grp * 32 + lo);
            }
        }
    }
}

// ----------------------------------------------------------- radix sort ----
// Pass structure: per-block digit histogram -> exclusive scan over (digit-major)
// counters -> stable scatter.  Stability inside a block comes from processing
// the block's items in order: warp w owns SORT_ITEMS rounds of 32 consecutive
// items, ranks inside a round come from __match_any_sync + popc.
__device__ __forceinline__ uint32_t clamp_count(const uint32_t* num, uint32_t cap) { return min(num[0], cap); }

__global__ void __launch_bounds__(RDG_BLOCK) radix_hist_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ num,
                                                               uint32_t cap, int shift, int bits, int nblk,
                                                               uint32_t* __restrict__ hist) {
    __shared__ uint32_t h[SORT_MAX_BINS];
    const int bins = 1 << bits;
    for (int b = threadIdx.x; b < bins; b += RDG_BLOCK) h[b] = 0;
    __syncthreads();
    const uint32_t n = clamp_count(num, cap);
    const uint32_t base = blockIdx.x * SORT_TILE;
    if (base < n) {
        const uint32_t mask = bins - 1;
#pragma unroll 4
        for (int k = 0; k < SORT_ITEMS; ++k) {
            const uint32_t i = base + k * RDG_BLOCK + threadIdx.x;
            if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & mask], 1u);
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < bins; b += RDG_BLOCK) hist[(size_t)b * nblk + blockIdx.x] = h[b];
}

__global__ void __launch_bounds__(RDG_BLOCK) radix_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                                  uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                                  const uint32_t* __restrict__ num, uint32_t cap, int shift, int bits,
                                                                  int nblk, const uint32_t* __restrict__ hist_scanned) {
    __shared__ uint32_t wh[NWARPS][SORT_MAX_BINS];
    const uint32_t n = clamp_count(num, cap);
    const uint32_t base = blockIdx.x * SORT_TILE;
    if (base >= n) return;
    const int bins = 1 << bits;
    const uint32_t mask = bins - 1;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int b = threadIdx.x; b < NWARPS * SORT_MAX_BINS; b += RDG_BLOCK) (&wh[0][0])[b] = 0;
    __syncthreads();

    uint64_t key[SORT_ITEMS];
    uint32_t rank[SORT_ITEMS];
    const uint32_t wbase = base + w * (32 * SORT_ITEMS);
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        const bool valid = i < n;
        key[r] = valid ? keys_in[i] : 0;
        const uint32_t d = (uint32_t)(key[r] >> shift) & mask;
        const uint32_t peers = rdg_match_digit(d, bits, __ballot_sync(0xffffffffu, valid));
        const int leader = __ffs(peers) - 1;
        const uint32_t below = __popc(peers & ((1u << lane) - 1u));
        uint32_t old = 0;
        if (valid && lane == leader) {
            old = wh[w][d];
            wh[w][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[r] = old + below;
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive scan across the warps, seeded with this block's global offset
    for (int b = threadIdx.x; b < bins; b += RDG_BLOCK) {
        uint32_t run = hist_scanned[(size_t)b * nblk + blockIdx.x];
#pragma unroll
        for (int k = 0; k < NWARPS; ++k) { const uint32_t t = wh[k][b]; wh[k][b] = run; run += t; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        if (i < n) {
            const uint32_t d = (uint32_t)(key[r] >> shift) & mask;
            const uint32_t dst = wh[w][d] + rank[r];
            keys_out[dst] = key[r];
            vals_out[dst] = vals_in[i];
        }
    }
}

// ----------------------------------------------------- identifyTileRanges ----
__global__ void __launch_bounds__(RDG_BLOCK) tile_ranges_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ num,
                                                                uint32_t cap, uint2* __restrict__ ranges) {
    const uint32_t n = clamp_count(num, cap);
    const uint32_t stride = gridDim.x * RDG_BLOCK;
    for (uint32_t i = blockIdx.x * RDG_BLOCK + threadIdx.x; i < n; i += stride) {
        const uint32_t t = (uint32_t)(keys[i] >> 32);
        if (i == 0) ranges[t].x = 0;
        else {
            const uint32_t tp = (uint32_t)(keys[i - 1] >> 32);
            if (tp != t) { ranges[tp].y = i; ranges[t].x = i; }
        }
        if (i == n - 1) ranges[t].y = n;
    }
}

// ------------------------------------------------------------------- host ----
static int bits_for(uint32_t v) {  // bits needed to represent values 0..v
    int b = 0;
    while (v) { ++b; v >>= 1; }
    return b;
}

struct BinLayout {
    int64_t scan_sums, keys_b, vals_b, hist, hist_sums, total;
    int nblk;
};

static BinLayout bin_layout(int64_t n, int64_t d_cap) {
    BinLayout L;
    int64_t off = 0;
    L.nblk = rdg_div_up(d_cap > 0 ? d_cap : 1, SORT_TILE);
    L.scan_sums = off; off += rdg_align_up((int64_t)rdg_div_up(n > 0 ? n : 1, SCAN_TILE) * 4, 256);
    L.keys_b = off;    off += rdg_align_up(d_cap * 8, 256);
    L.vals_b = off;    off += rdg_align_up(d_cap * 4, 256);
    const int64_t hist_n = (int64_t)SORT_MAX_BINS * L.nblk;
    L.hist = off;      off += rdg_align_up(hist_n * 4, 256);
    L.hist_sums = off; off += rdg_align_up((int64_t)rdg_div_up(hist_n, SCAN_TILE) * 4, 256);
    L.total = off;
    return L;
}

extern "C" int64_t rdg_bin_workspace_bytes(int64_t n, int64_t d_cap, int32_t height, int32_t width) {
    (void)height; (void)width;
    if (n < 0 || d_cap < 0) return RDG_E_ARG;
    return bin_layout(n, d_cap).total;
}

extern "C" int rdg_bin(int64_t n, const RdgGeom* geom, int32_t height, int32_t width, int64_t d_cap,
                       const RdgBins* bins, void* workspace, int64_t workspace_bytes, void* stream) {
    RDG_CHECK_ARG(geom && bins && workspace, "null argument");
    RDG_CHECK_ARG(n >= 0 && d_cap > 0 && d_cap < (int64_t)0xffffffffLL, "bad sizes");
    RDG_CHECK_ARG(bins->keys_sorted && bins->vals_sorted && bins->ranges && bins->num_rendered && (n == 0 || bins->point_offsets),
                  "null bin buffer");
    const BinLayout L = bin_layout(n, d_cap);
    if (workspace_bytes < L.total) {
        rdg_set_error("rdg_bin: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)L.total);
        return RDG_E_CAPACITY;
    }
    cudaStream_t s = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    const int gx = (width + RDG_TILE - 1) / RDG_TILE, gy = (height + RDG_TILE - 1) / RDG_TILE;
    const int tiles = gx * gy;
    RDG_CUDA(cudaMemsetAsync(bins->ranges, 0, (size_t)tiles * 2 * sizeof(uint32_t), s));
    if (n == 0) {
        RDG_CUDA(cudaMemsetAsync(bins->num_rendered, 0, 2 * sizeof(uint32_t), s));
        return RDG_OK;
    }
    int rc = scan_u32<true>(geom->tiles_touched, bins->point_offsets, n, (uint32_t*)(ws + L.scan_sums),
                            bins->num_rendered, (uint32_t)d_cap, s);
    if (rc) return rc;

    const int total_bits = 32 + bits_for((uint32_t)(tiles - 1));
    const int passes = (total_bits + SORT_MAX_BITS - 1) / SORT_MAX_BITS;
    const int digit = (total_bits + passes - 1) / passes;
    uint64_t* kx = bins->keys_sorted;  uint32_t* vx = bins->vals_sorted;
    uint64_t* ky = (uint64_t*)(ws + L.keys_b);  uint32_t* vy = (uint32_t*)(ws + L.vals_b);
    // pass k (1-based) writes X when (passes - k) is even, so the last pass lands in keys_sorted
    uint64_t* k_src = bins->keys_unsorted ? bins->keys_unsorted : ((passes % 2 == 0) ? kx : ky);
    uint32_t* v_src = bins->vals_unsorted ? bins->vals_unsorted : ((passes % 2 == 0) ? vx : vy);

    {
        const int64_t groups = (n + 31) / 32;
        const int64_t want = (groups + NWARPS - 1) / NWARPS;
        const int grid = (int)(want < (int64_t)RDG_SM_COUNT * 8 ? want : (int64_t)RDG_SM_COUNT * 8);
        duplicate_kernel<<<grid, RDG_BLOCK, 0, s>>>(n, geom->radii, geom->tiles_touched, bins->point_offsets,
                                                   (const float4*)geom->p0, (const float2*)geom->p2, gx, gy,
                                                   k_src, v_src, (uint32_t)d_cap);
        RDG_CHECK_LAUNCH();
        rdg_count_launches(1);
    }
    uint32_t* hist = (uint32_t*)(ws + L.hist);
    uint32_t* hist_sums = (uint32_t*)(ws + L.hist_sums);
    for (int k = 1; k <= passes; ++k) {
        const int shift = (k - 1) * digit;
        const int bits = (shift + digit <= total_bits) ? digit : (total_bits - shift);
        uint64_t* k_dst = ((passes - k) % 2 == 0) ? kx : ky;
        uint32_t* v_dst = ((passes - k) % 2 == 0) ? vx : vy;
        radix_hist_kernel<<<L.nblk, RDG_BLOCK, 0, s>>>(k_src, bins->num_rendered, (uint32_t)d_cap, shift, bits, L.nblk, hist);
        rc = scan_u32<false>(hist, hist, (int64_t)(1 << bits) * L.nblk, hist_sums, nullptr, 0, s);
        if (rc) return rc;
        radix_scatter_kernel<<<L.nblk, RDG_BLOCK, 0, s>>>(k_src, v_src, k_dst, v_dst, bins->num_rendered, (uint32_t)d_cap,
                                                        shift, bits, L.nblk, hist);
        RDG_CHECK_LAUNCH();
        rdg_count_launches(2);
        k_src = k_dst;
        v_src = v_dst;
    }
    {
        const int grid = (int)(L.nblk * (SORT_TILE / RDG_BLOCK) < RDG_SM_COUNT * 8 ? L.nblk * (SORT_TILE / RDG_BLOCK) : RDG_SM_COUNT * 8);
        tile_ranges_kernel<<<grid, RDG_BLOCK, 0, s>>>(kx, bins->num_rendered, (uint32_t)d_cap, (uint2*)bins->ranges);
        RDG_CHECK_LAUNCH();
        rdg_count_launches(1);
    }
    return RDG_OK;
}
