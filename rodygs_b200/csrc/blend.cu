// Front-to-back alpha blend (forward) and reverse-order backward over the
// depth-sorted per-tile lists.  SURVEY.md §8 rows a9 / a10; spec: SURVEY.md
// App. A.4-A.6 == oracle/splat_oracle.py::blend (+ autograd).
//
// B200 mapping
//  * one CTA per 16x16 tile, 8 warps; warp w owns an 8x4 pixel sub-tile so that a
//    Gaussian's footprint can be rejected per warp with one broadcast LDS.128 and four
//    compares (exact alpha >= 1/255 ellipse bound, conservative margin) before any
//    FP32 / MUFU work is spent;
//  * per-tile batches of 256 Gaussians are staged in shared memory as three packed
//    records (16 + 16 + 8 bytes) + the bounding box;
//  * backward: per-Gaussian gradients are reduced across the warp with shuffles only
//    when some lane contributed, combined across the 8 warps in shared memory, and
//    leave the CTA as three 16-byte vector atomics (red.global.add.v4.f32, sm_90+) per
//    (Gaussian, tile) instead of upstream's ten scalar atomics per (Gaussian, pixel).
// Bound: FP32 pipe + MUFU (ex2) + shared-memory broadcast; charged against the HBM
// roofline as north_star asks (algorithmic bytes: 48 B per duplicate + 28 B per pixel
// forward; 48 B per duplicate + 44 B per pixel backward).
#include "common.cuh"

#define BATCH 256

struct __align__(16) Staged {
    float4 a[BATCH];    // px, py, A, B
    float4 b[BATCH];    // C, opacity, r, g
    float2 c[BATCH];    // b, depth
    float4 bb[BATCH];   // xmin, xmax, ymin, ymax of the alpha >= 1/255 ellipse
    uint32_t id[BATCH];
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

__device__ __forceinline__ float4 rdg_bbox(float4 a, float4 b) {
    const float A = a.z, B = a.w, C = b.x, o = b.y;
    const float huge = 3.0e38f;
    if (!(o >= RDG_ALPHA_MIN)) return make_float4(huge, -huge, huge, -huge);  // can never reach 1/255
    const float tau = 2.0f * __logf(255.0f * o) * 1.001f + 1e-3f;
    const float det = A * C - B * B;
    if (!(det > 0.0f)) return make_float4(-huge, huge, -huge, huge);
    const float ex = sqrtf(tau * C / det) * 1.001f + 0.01f;
    const float ey = sqrtf(tau * A / det) * 1.001f + 0.01f;
    if (!(ex == ex) || !(ey == ey)) return make_float4(-huge, huge, -huge, huge);
    return make_float4(a.x - ex, a.x + ex, a.y - ey, a.y + ey);
}

__global__ void __launch_bounds__(BATCH) blend_fwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ vals,
                                                          const float4* __restrict__ p0, const float4* __restrict__ p1,
                                                          const float2* __restrict__ p2, const float* __restrict__ bg,
                                                          int W, int H, int gx, float* __restrict__ out_color,
                                                          float* __restrict__ out_depth, float* __restrict__ out_alpha,
                                                          float* __restrict__ out_T, uint32_t* __restrict__ out_ncontrib) {
    __shared__ Staged sm;
    const int tile = blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wx0 = tx * RDG_TILE + (warp & 1) * 8, wy0 = ty * RDG_TILE + (warp >> 1) * 4;
    const int pxi = wx0 + (lane & 7), pyi = wy0 + (lane >> 3);
    const bool inside = pxi < W && pyi < H;
    const float pixx = (float)pxi, pixy = (float)pyi;
    const float fx0 = (float)wx0, fx1 = (float)(wx0 + 7), fy0 = (float)wy0, fy1 = (float)(wy0 + 3);

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
            sm.bb[threadIdx.x] = rdg_bbox(a, b);
        }
        __syncthreads();
        const int cnt = min(BATCH, n_g - r * BATCH);
        for (int j = 0; j < cnt; ++j) {
            const float4 bb = sm.bb[j];
            if (bb.y < fx0 || bb.x > fx1 || bb.w < fy0 || bb.z > fy1) continue;  // warp-uniform reject
            if (done) continue;
            const float4 a = sm.a[j];
            const float4 b = sm.b[j];
            float G, alpha;
            if (!rdg_alpha(a.x - pixx, a.y - pixy, a.z, a.w, b.x, b.y, G, alpha)) continue;
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

__global__ void __launch_bounds__(BATCH) blend_bwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ vals,
                                                          const float4* __restrict__ p0, const float4* __restrict__ p1,
                                                          const float2* __restrict__ p2, const float* __restrict__ bg,
                                                          int W, int H, int gx, const float* __restrict__ final_T,
                                                          const uint32_t* __restrict__ n_contrib,
                                                          const float* __restrict__ dL_dcolor, const float* __restrict__ dL_ddepth,
                                                          const float* __restrict__ dL_dalpha, float* __restrict__ acc) {
    __shared__ Staged sm;
    __shared__ __align__(16) float sacc[BATCH][NACC];
    __shared__ uint32_t smax[BATCH / 32];
    const int tile = blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wx0 = tx * RDG_TILE + (warp & 1) * 8, wy0 = ty * RDG_TILE + (warp >> 1) * 4;
    const int pxi = wx0 + (lane & 7), pyi = wy0 + (lane >> 3);
    const bool inside = pxi < W && pyi < H;
    const float pixx = (float)pxi, pixy = (float)pyi;
    const float fx0 = (float)wx0, fx1 = (float)(wx0 + 7), fy0 = (float)wy0, fy1 = (float)(wy0 + 3);
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
    uint32_t m = __reduce_max_sync(0xffffffffu, my_last);
    if (lane == 0) smax[warp] = m;
    for (int k = threadIdx.x; k < BATCH * NACC; k += BATCH) (&sacc[0][0])[k] = 0.f;
    __syncthreads();
    uint32_t max_last = 0;
#pragma unroll
    for (int k = 0; k < BATCH / 32; ++k) max_last = max(max_last, smax[k]);
    max_last = min(max_last, (uint32_t)n_g);
    if (max_last == 0) return;

    float T = T_final;
    float rec_r = 0.f, rec_g = 0.f, rec_b = 0.f, rec_d = 0.f, rec_a = 0.f;
    float last_alpha = 0.f, last_r = 0.f, last_g = 0.f, last_b = 0.f, last_d = 0.f;

    const int rounds = ((int)max_last + BATCH - 1) / BATCH;
    for (int r = 0; r < rounds; ++r) {
        // batch r covers list positions pos = max_last-1 - (r*BATCH + slot), slot = 0..cnt-1 (back to front)
        const int cnt = min(BATCH, (int)max_last - r * BATCH);
        __syncthreads();
        if ((int)threadIdx.x < cnt) {
            const int pos = (int)max_last - 1 - (r * BATCH + (int)threadIdx.x);
            const uint32_t id = vals[range.x + pos];
            const float4 a = p0[id], b = p1[id];
            sm.id[threadIdx.x] = id;
            sm.a[threadIdx.x] = a;
            sm.b[threadIdx.x] = b;
            sm.c[threadIdx.x] = p2[id];
            sm.bb[threadIdx.x] = rdg_bbox(a, b);
        }
        __syncthreads();
        for (int j = 0; j < cnt; ++j) {
            const float4 bb = sm.bb[j];
            if (bb.y < fx0 || bb.x > fx1 || bb.w < fy0 || bb.z > fy1) continue;  // warp-uniform reject
            const uint32_t pos = max_last - 1 - (uint32_t)(r * BATCH + j);
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
                T = T / (1.0f - alpha);
                const float wgt = alpha * T;
                rec_r = last_alpha * last_r + (1.f - last_alpha) * rec_r;
                rec_g = last_alpha * last_g + (1.f - last_alpha) * rec_g;
                rec_b = last_alpha * last_b + (1.f - last_alpha) * rec_b;
                rec_d = last_alpha * last_d + (1.f - last_alpha) * rec_d;
                rec_a = last_alpha + (1.f - last_alpha) * rec_a;
                last_r = b.z; last_g = b.w; last_b = c.x; last_d = c.y;
                float dL_da = (b.z - rec_r) * gr + (b.w - rec_g) * gg + (c.x - rec_b) * gb + (c.y - rec_d) * gd + (1.f - rec_a) * ga;
                dL_da *= T;
                last_alpha = alpha;
                dL_da += (-T_final / (1.f - alpha)) * bg_dot;
                const float dL_dG = b.y * dL_da;
                const float gdx = G * dx, gdy = G * dy;
                v[0] = dL_dG * (-gdx * a.z - gdy * a.w);      // d/dpx (pixel units)
                v[1] = dL_dG * (-gdy * b.x - gdx * a.w);      // d/dpy
                v[2] = -0.5f * gdx * dx * dL_dG;              // dA
                v[3] = -gdx * dy * dL_dG;                     // dB (B enters `power` once)
                v[4] = -0.5f * gdy * dy * dL_dG;              // dC
                v[5] = G * dL_da;                             // dopacity
                v[6] = wgt * gr; v[7] = wgt * gg; v[8] = wgt * gb;   // drgb
                v[9] = wgt * gd;                              // ddepth
            }
#pragma unroll
            for (int k = 0; k < 10; ++k) v[k] = warp_sum(v[k]);
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 10; ++k) atomicAdd(&sacc[j][k], v[k]);
            }
        }
        __syncthreads();
        if ((int)threadIdx.x < cnt) {
            float4* row = reinterpret_cast<float4*>(&sacc[threadIdx.x][0]);
            const float4 r0 = row[0], r1 = row[1], r2 = row[2];
            const bool nz = r0.x != 0.f || r0.y != 0.f || r0.z != 0.f || r0.w != 0.f || r1.x != 0.f || r1.y != 0.f ||
                            r1.z != 0.f || r1.w != 0.f || r2.x != 0.f || r2.y != 0.f;
            if (nz) {
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

extern "C" int rdg_blend_fwd(int64_t n, const RdgGeom* geom, const RdgBins* bins, const RdgView* view,
                             const RdgImage* out, void* stream) {
    (void)n;
    RDG_CHECK_ARG(geom && bins && view && out, "null argument");
    RDG_CHECK_ARG(out->color && out->depth && out->alpha && out->final_T && out->n_contrib, "null image buffer");
    RDG_CHECK_ARG(view->bg, "null background");
    const int W = view->width, H = view->height;
    const int gx = (W + RDG_TILE - 1) / RDG_TILE, gy = (H + RDG_TILE - 1) / RDG_TILE;
    blend_fwd_kernel<<<gx * gy, BATCH, 0, (cudaStream_t)stream>>>(
        (const uint2*)bins->ranges, bins->vals_sorted, (const float4*)geom->p0, (const float4*)geom->p1,
        (const float2*)geom->p2, view->bg, W, H, gx, out->color, out->depth, out->alpha, out->final_T, out->n_contrib);
    RDG_CHECK_LAUNCH();
    rdg_count_launches(1);
    return RDG_OK;
}

extern "C" int rdg_blend_bwd(int64_t n, const RdgGeom* geom, const RdgBins* bins, const RdgView* view,
                             const RdgImage* fwd, const float* dL_dcolor, const float* dL_ddepth,
                             const float* dL_dalpha, float* acc, void* stream) {
    (void)n;
    RDG_CHECK_ARG(geom && bins && view && fwd && acc, "null argument");
    RDG_CHECK_ARG(fwd->final_T && fwd->n_contrib, "null forward state");
    const int W = view->width, H = view->height;
    const int gx = (W + RDG_TILE - 1) / RDG_TILE, gy = (H + RDG_TILE - 1) / RDG_TILE;
    blend_bwd_kernel<<<gx * gy, BATCH, 0, (cudaStream_t)stream>>>(
        (const uint2*)bins->ranges, bins->vals_sorted, (const float4*)geom->p0, (const float4*)geom->p1,
        (const float2*)geom->p2, view->bg, W, H, gx, fwd->final_T, fwd->n_contrib, dL_dcolor, dL_ddepth, dL_dalpha, acc);
    RDG_CHECK_LAUNCH();
    rdg_count_launches(1);
    return RDG_OK;
}
