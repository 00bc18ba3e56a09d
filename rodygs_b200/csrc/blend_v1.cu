// Front-to-back alpha blend (forward) and reverse-order backward over the
// depth-sorted per-tile lists.  SURVEY.md §8 rows a9 / a10; spec: SURVEY.md
// App. A.4-A.6 == oracle/splat_oracle.py::blend (+ autograd).
//
// B200 mapping
//  * one CTA per 16x16 tile, 8 warps; warp w owns an 8x4 pixel sub-tile.
//  * per-tile batches of 256 Gaussians are staged in shared memory as three packed
//    records (16 + 16 + 8 bytes).  The staging thread also computes the exact
//    alpha >= 1/255 ellipse bound of its Gaussian and turns it into an 8-bit mask of the
//    sub-tiles it can touch; each warp then compacts the batch into its own list with
//    ballot/popc, so the inner loop only ever visits Gaussians that can contribute to this
//    warp (28 % of the (warp, Gaussian) pairs at the bench config) - no per-iteration
//    cull test on the ALU pipe, which was the limiter of the first version (ncu r01).
//  * backward: the ten per-Gaussian partial gradients are reduced across the warp with a
//    12-shuffle reduce-scatter butterfly (5+3+2+1+1) instead of 10 x 5 butterflies; the
//    ten lanes that end up owning a total add it to the CTA's shared accumulators, which
//    leave the CTA as three 16-byte vector atomics (red.global.add.v4.f32, sm_90+) per
//    (Gaussian, tile) instead of upstream's ten scalar atomics per (Gaussian, pixel).
// Bound: FP32/ALU issue + MUFU (ex2) + shuffle crossbar; charged against the HBM
// roofline as north_star asks (algorithmic bytes: 44 B per duplicate + 28 B per pixel
// forward; 44 B per duplicate + 44 B per pixel + 48 B per visible Gaussian backward).
#include "common.cuh"

#define BATCH 256
#define NWARP (BATCH / 32)

struct __align__(16) Staged {
    float4 a[BATCH];            // px, py, A, B
    float4 b[BATCH];            // C, opacity, r, g
    float2 c[BATCH];            // b, depth
    uint32_t id[BATCH];
    uint8_t wm[BATCH];          // bit w: may touch warp w's 8x4 sub-tile
    uint8_t list[NWARP][BATCH]; // per-warp compacted batch slots, in list order
};

// alpha = min(0.99, o * exp(power)); identical instruction sequence in both passes so
// that the skip decisions replayed by the backward pass match the forward ones bit for bit.
__device__ __forceinline__ bool rdg_alpha(float dx, float dy, float A, float B, float C, float o, float& G, float& alpha) {
    const float q = __fmaf_rn(A * dx, dx, (C * dy) * dy);
    const float power = __fmaf_rn(-0.5f, q, -(B * dx) * dy);
    if (power > 0.0f) return false;
    G = __expf(power);
    alpha = fminf(RDG_ALPHA_MAX, o * G);
    return alpha >= RDG_ALPHA_MIN;
}

// 8-bit mask of the 8x4 sub-tiles (bit = sx + 2*sy) that the alpha >= 1/255 ellipse of
// this Gaussian can reach; conservative (0.1 % + 0.01 px margins), exact rule stays per pixel.
__device__ __forceinline__ uint32_t rdg_warp_mask(float4 a, float4 b, float tile_x0, float tile_y0) {
    const float A = a.z, B = a.w, C = b.x, o = b.y;
    if (!(o >= RDG_ALPHA_MIN)) return 0u;   // o * exp(power <= 0) can never reach 1/255
    const float tau = 2.0f * __logf(255.0f * o) * 1.001f + 1e-3f;
    const float det = A * C - B * B;
    if (!(det > 0.0f)) return 0xffu;
    const float ex = sqrtf(tau * C / det) * 1.001f + 0.01f;
    const float ey = sqrtf(tau * A / det) * 1.001f + 0.01f;
    if (!(ex == ex) || !(ey == ey)) return 0xffu;
    const float x0 = a.x - ex - tile_x0, x1 = a.x + ex - tile_x0;   // relative to the tile origin
    const float y0 = a.y - ey - tile_y0, y1 = a.y + ey - tile_y0;
    uint32_t xb = 0, m = 0;
    if (x1 >= 0.f && x0 <= 7.f) xb |= 1u;
    if (x1 >= 8.f && x0 <= 15.f) xb |= 2u;
#pragma unroll
    for (int sy = 0; sy < 4; ++sy)
        if (y1 >= (float)(4 * sy) && y0 <= (float)(4 * sy + 3)) m |= xb << (2 * sy);
    return m;
}

// Build this warp's compacted list of batch slots (order preserved). Returns its length.
__device__ __forceinline__ int rdg_compact(Staged& sm, int cnt, int warp, int lane) {
    int n_w = 0;
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int g = 0; g < BATCH / 32; ++g) {
        const int j = g * 32 + lane;
        const bool bit = (j < cnt) && ((sm.wm[j] >> warp) & 1u);
        const unsigned m = __ballot_sync(0xffffffffu, bit);
        if (bit) sm.list[warp][n_w + __popc(m & lt)] = (uint8_t)j;
        n_w += __popc(m);
    }
    __syncwarp();
    return n_w;
}

__global__ void __launch_bounds__(BATCH) blend_fwd_kernel_v1(const uint2* __restrict__ ranges, const uint32_t* __restrict__ vals,
                                                          const float4* __restrict__ p0, const float4* __restrict__ p1,
                                                          const float2* __restrict__ p2, const float* __restrict__ bg,
                                                          int W, int H, int gx, float* __restrict__ out_color,
                                                          float* __restrict__ out_depth, float* __restrict__ out_alpha,
                                                          float* __restrict__ out_T, uint32_t* __restrict__ out_ncontrib) {
    __shared__ Staged sm;
    const int tile = blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pxi = tx * RDG_TILE + (warp & 1) * 8 + (lane & 7), pyi = ty * RDG_TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = pxi < W && pyi < H;
    const float pixx = (float)pxi, pixy = (float)pyi;
    const float tile_x0 = (float)(tx * RDG_TILE), tile_y0 = (float)(ty * RDG_TILE);

    const uint2 range = ranges[tile];
    const int n_g = (int)(range.y - range.x);
    const int rounds = (n_g + BATCH - 1) / BATCH;

    bool done = !inside;
    float T = 1.0f, Cr = 0.f, Cg = 0.f, Cb = 0.f, Dp = 0.f;
    uint32_t last = 0;

    for (int r = 0; r < rounds; ++r) {
        if (__syncthreads_count(done) == BATCH) break;
        const int idx = r * BATCH + threadIdx.x;
        if (idx < n_g) {
            const uint32_t id = vals[range.x + idx];
            const float4 a = p0[id], b = p1[id];
            sm.a[threadIdx.x] = a;
            sm.b[threadIdx.x] = b;
            sm.c[threadIdx.x] = p2[id];
            sm.wm[threadIdx.x] = (uint8_t)rdg_warp_mask(a, b, tile_x0, tile_y0);
        }
        __syncthreads();
        const int cnt = min(BATCH, n_g - r * BATCH);
        if (__all_sync(0xffffffffu, done)) continue;     // this warp's 32 pixels are saturated
        const int n_w = rdg_compact(sm, cnt, warp, lane);
        for (int i = 0; i < n_w; ++i) {
            const int j = sm.list[warp][i];
            const float4 a = sm.a[j];
            const float4 b = sm.b[j];
            float G, alpha;
            if (done || !rdg_alpha(a.x - pixx, a.y - pixy, a.z, a.w, b.x, b.y, G, alpha)) continue;
            const float test_T = T * (1.0f - alpha);
            if (test_T < RDG_T_STOP) { done = true; continue; }
            const float2 c = sm.c[j];
            const float wgt = alpha * T;
            Cr = __fmaf_rn(b.z, wgt, Cr);
            Cg = __fmaf_rn(b.w, wgt, Cg);
            Cb = __fmaf_rn(c.x, wgt, Cb);
            Dp = __fmaf_rn(c.y, wgt, Dp);
            T = test_T;
            last = (uint32_t)(r * BATCH + j + 1);
        }
    }
    if (inside) {
        const size_t pix = (size_t)pyi * W + pxi, hw = (size_t)H * W;
        out_color[pix] = Cr + T * bg[0];
        out_color[hw + pix] = Cg + T * bg[1];
        out_color[2 * hw + pix] = Cb + T * bg[2];
        out_depth[pix] = Dp;
        out_alpha[pix] = 1.0f - T;
        out_T[pix] = T;
        out_ncontrib[pix] = last;
    }
}

// ---------------------------------------------------------------- backward ----
#define NACC 12   // dpx dpy dA dB dC dop dr dg db ddepth pad pad

// Reduce-scatter of ten per-lane values over the warp in 12 shuffles.  On return the lane
// rdg_rs10_owner(k) holds the warp total of v[k] in the returned register.
__device__ __forceinline__ float rdg_reduce_scatter10(const float* v, int lane) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
    float a[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const float keep = b4 ? v[i + 5] : v[i], send = b4 ? v[i] : v[i + 5];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    float q0, q1, q2;
    {
        const float k0 = b3 ? a[1] : a[0], s0 = b3 ? a[0] : a[1];
        const float k1 = b3 ? a[3] : a[2], s1 = b3 ? a[2] : a[3];
        q0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 8);
        q1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 8);
        q2 = a[4] + __shfl_xor_sync(0xffffffffu, a[4], 8);
    }
    float c0, c1;
    {
        const float k = b2 ? q1 : q0, s = b2 ? q0 : q1;
        c0 = k + __shfl_xor_sync(0xffffffffu, s, 4);
        c1 = q2 + __shfl_xor_sync(0xffffffffu, q2, 4);
    }
    float d;
    {
        const float k = b1 ? c1 : c0, s = b1 ? c0 : c1;
        d = k + __shfl_xor_sync(0xffffffffu, s, 2);
    }
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;
}

// value index owned by a lane after rdg_reduce_scatter10, or -1 (only lanes with bit0 == 0 own)
__device__ __forceinline__ int rdg_rs10_index(int lane) {
    if (lane & 1) return -1;
    const int hi = (lane & 16) ? 5 : 0;
    if (lane & 2) return ((lane & 12) == 0) ? hi + 4 : -1;          // lanes 2, 18
    if (lane & 4) return hi + ((lane & 8) ? 3 : 2);                 // lanes 4, 12, 20, 28
    return hi + ((lane & 8) ? 1 : 0);                               // lanes 0, 8, 16, 24
}

__global__ void __launch_bounds__(BATCH) blend_bwd_kernel_v1(const uint2* __restrict__ ranges, const uint32_t* __restrict__ vals,
                                                          const float4* __restrict__ p0, const float4* __restrict__ p1,
                                                          const float2* __restrict__ p2, const float* __restrict__ bg,
                                                          int W, int H, int gx, const float* __restrict__ final_T,
                                                          const uint32_t* __restrict__ n_contrib,
                                                          const float* __restrict__ dL_dcolor, const float* __restrict__ dL_ddepth,
                                                          const float* __restrict__ dL_dalpha, float* __restrict__ acc) {
    __shared__ Staged sm;
    __shared__ __align__(16) float sacc[BATCH][NACC];
    __shared__ uint32_t smax[NWARP];
    const int tile = blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pxi = tx * RDG_TILE + (warp & 1) * 8 + (lane & 7), pyi = ty * RDG_TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = pxi < W && pyi < H;
    const float pixx = (float)pxi, pixy = (float)pyi;
    const float tile_x0 = (float)(tx * RDG_TILE), tile_y0 = (float)(ty * RDG_TILE);
    const size_t pix = (size_t)pyi * W + pxi, hw = (size_t)H * W;

    const uint2 range = ranges[tile];
    const int n_g = (int)(range.y - range.x);

    const float T_final = inside ? final_T[pix] : 0.f;
    const uint32_t my_last = inside ? n_contrib[pix] : 0u;
    float gr = 0.f, gg = 0.f, gb = 0.f, gd = 0.f, ga = 0.f;
    if (inside) {
        if (dL_dcolor) { gr = dL_dcolor[pix]; gg = dL_dcolor[hw + pix]; gb = dL_dcolor[2 * hw + pix]; }
        if (dL_ddepth) gd = dL_ddepth[pix];
        if (dL_dalpha) ga = dL_dalpha[pix];
    }
    const float bg_dot = bg[0] * gr + bg[1] * gg + bg[2] * gb;

    // the CTA only has to walk back from the deepest contributor of any of its pixels
    const uint32_t warp_last = __reduce_max_sync(0xffffffffu, my_last);
    if (lane == 0) smax[warp] = warp_last;
    for (int k = threadIdx.x; k < BATCH * NACC; k += BATCH) (&sacc[0][0])[k] = 0.f;
    __syncthreads();
    uint32_t max_last = 0;
#pragma unroll
    for (int k = 0; k < NWARP; ++k) max_last = max(max_last, smax[k]);
    max_last = min(max_last, (uint32_t)n_g);
    if (max_last == 0) return;

    // Running state per pixel.  With P_i = <g, (r,g,b,depth,1)_i> the alpha gradient is
    //   dL/dalpha_i = T_i P_i - (A_dot_i + T_final <g_rgb, bg>) / (1 - alpha_i),  A_dot_i = sum_{k>i} P_k alpha_k T_k,
    // so one scalar recursion replaces the five per-channel "accumulated behind" recursions.
    float T = T_final;
    float A_dot = T_final * bg_dot;
    const int own_k = rdg_rs10_index(lane);

    const int rounds = ((int)max_last + BATCH - 1) / BATCH;
    for (int r = 0; r < rounds; ++r) {
        // batch r covers list positions pos = max_last-1 - (r*BATCH + slot), slot = 0..cnt-1 (back to front)
        const int cnt = min(BATCH, (int)max_last - r * BATCH);
        const int pos0 = (int)max_last - 1 - r * BATCH;       // list position of slot 0
        __syncthreads();
        if ((int)threadIdx.x < cnt) {
            const uint32_t id = vals[range.x + pos0 - (int)threadIdx.x];
            const float4 a = p0[id], b = p1[id];
            sm.id[threadIdx.x] = id;
            sm.a[threadIdx.x] = a;
            sm.b[threadIdx.x] = b;
            sm.c[threadIdx.x] = p2[id];
            sm.wm[threadIdx.x] = (uint8_t)rdg_warp_mask(a, b, tile_x0, tile_y0);
        }
        __syncthreads();
        // slots whose position is beyond this warp's deepest contributor cannot contribute
        const int first_slot = max(0, pos0 - (int)warp_last + 1);
        if (first_slot < cnt) {
            const int n_w = rdg_compact(sm, cnt, warp, lane);
            for (int i = 0; i < n_w; ++i) {
                const int j = sm.list[warp][i];
                if (j < first_slot) continue;
                const uint32_t pos = (uint32_t)(pos0 - j);
                const float4 a = sm.a[j];
                const float4 b = sm.b[j];
                const float dx = a.x - pixx, dy = a.y - pixy;
                float G = 0.f, alpha = 0.f;
                bool on = pos < my_last;
                if (on) on = rdg_alpha(dx, dy, a.z, a.w, b.x, b.y, G, alpha);
                if (!__any_sync(0xffffffffu, on)) continue;
                float v[10];
#pragma unroll
                for (int k = 0; k < 10; ++k) v[k] = 0.f;
                if (on) {
                    const float2 c = sm.c[j];
                    const float inv = __fdividef(1.0f, 1.0f - alpha);
                    T *= inv;                                     // transmittance in front of this Gaussian
                    const float wgt = alpha * T;
                    const float P = fmaf(b.z, gr, fmaf(b.w, gg, fmaf(c.x, gb, fmaf(c.y, gd, ga))));
                    const float dL_da = fmaf(T, P, -inv * A_dot);
                    A_dot = fmaf(P, wgt, A_dot);
                    // raw moments; the conic / sign factors are applied once per (Gaussian, tile) at the flush
                    const float t5 = G * dL_da;                   // d/dopacity
                    const float w = b.y * t5;                     // dL/dG * G
                    const float wx = w * dx, wy = w * dy;
                    v[0] = wx; v[1] = wy;
                    v[2] = wx * dx; v[3] = wx * dy; v[4] = wy * dy;
                    v[5] = t5;
                    v[6] = wgt * gr; v[7] = wgt * gg; v[8] = wgt * gb;   // drgb
                    v[9] = wgt * gd;                              // ddepth
                }
                const float tot = rdg_reduce_scatter10(v, lane);
                if (own_k >= 0) atomicAdd(&sacc[j][own_k], tot);
            }
        }
        __syncthreads();
        if ((int)threadIdx.x < cnt) {
            float4* row = reinterpret_cast<float4*>(&sacc[threadIdx.x][0]);
            float4 r0 = row[0], r1 = row[1], r2 = row[2];
            const bool nz = r0.x != 0.f || r0.y != 0.f || r0.z != 0.f || r0.w != 0.f || r1.x != 0.f || r1.y != 0.f ||
                            r1.z != 0.f || r1.w != 0.f || r2.x != 0.f || r2.y != 0.f;
            if (nz) {
                // moments -> gradients: dpx = -(A Sx + B Sy), dpy = -(C Sy + B Sx), dA = -Sxx/2, dB = -Sxy, dC = -Syy/2
                const float4 ga4 = sm.a[threadIdx.x];
                const float cA = ga4.z, cB = ga4.w, cC = sm.b[threadIdx.x].x;
                const float sx = r0.x, sy = r0.y;
                r0.x = -(cA * sx + cB * sy);
                r0.y = -(cC * sy + cB * sx);
                r0.z *= -0.5f;
                r0.w = -r0.w;
                r1.x *= -0.5f;
                float4* dst = reinterpret_cast<float4*>(acc + (size_t)sm.id[threadIdx.x] * NACC);
                atomicAdd(dst + 0, r0);
                atomicAdd(dst + 1, r1);
                atomicAdd(dst + 2, r2);
                row[0] = make_float4(0.f, 0.f, 0.f, 0.f);
                row[1] = make_float4(0.f, 0.f, 0.f, 0.f);
                row[2] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
}

extern "C" int rdg_blend_fwd_v1(int64_t n, const RdgGeom* geom, const RdgBins* bins, const RdgView* view,
                             const RdgImage* out, void* stream) {
    (void)n;
    RDG_CHECK_ARG(geom && bins && view && out, "null argument");
    RDG_CHECK_ARG(out->color && out->depth && out->alpha && out->final_T && out->n_contrib, "null image buffer");
    RDG_CHECK_ARG(view->bg, "null background");
    const int W = view->width, H = view->height;
    const int gx = (W + RDG_TILE - 1) / RDG_TILE, gy = (H + RDG_TILE - 1) / RDG_TILE;
    blend_fwd_kernel_v1<<<gx * gy, BATCH, 0, (cudaStream_t)stream>>>(
        (const uint2*)bins->ranges, bins->vals_sorted, (const float4*)geom->p0, (const float4*)geom->p1,
        (const float2*)geom->p2, view->bg, W, H, gx, out->color, out->depth, out->alpha, out->final_T, out->n_contrib);
    RDG_CHECK_LAUNCH();
    rdg_count_launches(1);
    return RDG_OK;
}

extern "C" int rdg_blend_bwd_v1(int64_t n, const RdgGeom* geom, const RdgBins* bins, const RdgView* view,
                             const RdgImage* fwd, const float* dL_dcolor, const float* dL_ddepth,
                             const float* dL_dalpha, float* acc, void* stream) {
    (void)n;
    RDG_CHECK_ARG(geom && bins && view && fwd && acc, "null argument");
    RDG_CHECK_ARG(fwd->final_T && fwd->n_contrib, "null forward state");
    const int W = view->width, H = view->height;
    const int gx = (W + RDG_TILE - 1) / RDG_TILE, gy = (H + RDG_TILE - 1) / RDG_TILE;
    blend_bwd_kernel_v1<<<gx * gy, BATCH, 0, (cudaStream_t)stream>>>(
        (const uint2*)bins->ranges, bins->vals_sorted, (const float4*)geom->p0, (const float4*)geom->p1,
        (const float2*)geom->p2, view->bg, W, H, gx, fwd->final_T, fwd->n_contrib, dL_dcolor, dL_ddepth, dL_dalpha, acc);
    RDG_CHECK_LAUNCH();
    rdg_count_launches(1);
    return RDG_OK;
}
