// Front-to-back alpha blend (forward) and reverse-order backward over the
// depth-sorted per-tile lists.  SURVEY.md §8 rows a9 / a10; spec: SURVEY.md
// App. A.4-A.6 == oracle/splat_oracle.py::blend (+ autograd).
//
// B200 mapping (v2, profiles/README.md has the measured history)
//  * one CTA of 4 warps per 16x16 tile.  The tile is cut into sixteen 4x4-pixel sub-tiles; a
//    quarter-warp (8 lanes) owns one sub-tile and every lane owns TWO vertically adjacent
//    pixels, so the Gaussian record a lane fetches from shared memory (the crossbar is the
//    co-limiter of this loop, ncu r01) and the x-dependent half of the conic are used twice.
//  * per-tile batches of 128 list entries are staged in shared memory as three packed
//    records (16 + 16 + 8 bytes).  The staging thread solves the alpha >= 1/255 ellipse of its
//    Gaussian row by row and turns it into a 16-bit mask of the sub-tiles it can reach
//    (conservative margins; the exact rule stays per pixel).  Each warp compacts the batch
//    into one list per quarter with ballot/popc; the four quarters then walk their own lists
//    in lock step (3.0 (sub-tile, Gaussian) visits per list entry at the bench config,
//    0.84 warp iterations per entry against 2.3 for the 8x4-per-warp bounding-box version).
//  * backward: a lane first adds the ten partial gradients of its two pixels, the quarter
//    reduce-scatters them in 10 shuffles (5+3+2) and every lane stores the total it owns into
//    a per-(sub-tile, Gaussian) record in shared memory - a plain store, no atomics in the
//    loop: every record is written exactly once.  At the end of the batch one thread per list
//    entry sums its records and issues three 16-byte vector atomics
//    (red.global.add.v4.f32) per (Gaussian, tile).
// Bound: FP32/ALU issue + shared-memory crossbar (LDS broadcast + SHFL); charged against the
// HBM roofline as north_star asks (algorithmic bytes: 44 B per duplicate + 28 B per pixel
// forward; 44 B per duplicate + 44 B per pixel + 48 B per visible Gaussian backward).
#include <stdlib.h>
#include "common.cuh"

#define SUBS 16                 // 4x4-pixel sub-tiles per tile
#define BATCH 128               // list entries staged per round == threads per CTA
#define NWARP (BATCH / 32)
#define LROW (BATCH + 2)        // list row stride (u16): rows of neighbouring sub-tiles start in different banks
#define POOL 512                // (sub-tile, entry) gradient records per round (backward)
#define PREC 12                 // floats per record (10 used; 48-byte rows for vector loads)
#define FULL 0xffffffffu

extern "C" int rdg_blend_fwd_v1(int64_t n, const RdgGeom* geom, const RdgBins* bins, const RdgView* view, const RdgImage* out, void* stream);
extern "C" int rdg_blend_bwd_v1(int64_t n, const RdgGeom* geom, const RdgBins* bins, const RdgView* view, const RdgImage* fwd,
                     const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha, float* acc, void* stream);

struct __align__(16) Staged {
    float4 a[BATCH];              // px, py, A, B
    float4 b[BATCH];              // C, opacity, r, g
    float2 c[BATCH];              // b, depth
    uint32_t id[BATCH];
    uint16_t mask[BATCH];         // bit s: may touch sub-tile s (s = 4 * sub_y + sub_x)
    uint16_t ebase[BATCH];        // backward: first gradient record of this entry
    uint16_t list[SUBS][LROW];    // per sub-tile compacted entries: slot | (record << 7)
};

__device__ __forceinline__ float rdg_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rdg_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rdg_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// alpha = min(0.99, o * exp(power)), power = -(A dx^2 + C dy^2)/2 - B dx dy.  Identical
// instruction sequence in both passes so that the skip decisions replayed by the backward
// pass match the forward ones bit for bit.  Adx2 = (A dx) dx and Bdx = B dx are shared by the
// two pixels of a lane (same column).
__device__ __forceinline__ bool rdg_alpha(float Adx2, float Bdx, float C, float dy, float o, float& G, float& alpha) {
    const float q = __fmaf_rn(C * dy, dy, Adx2);
    const float power = __fmaf_rn(-0.5f, q, -Bdx * dy);
    G = rdg_ex2(power * 1.4426950408889634f);
    alpha = fminf(RDG_ALPHA_MAX, o * G);
    return (power <= 0.0f) && (alpha >= RDG_ALPHA_MIN);
}

// 16-bit mask of the 4x4 sub-tiles that the alpha >= 1/255 ellipse of this Gaussian reaches:
// with u = x - px, v = y - py the rule is  A u^2 + 2 B u v + C v^2 <= tau = 2 ln(255 o);
// for pixel row v that is  u in [(-B v - sqrt(tau A - v^2 det)) / A, (-B v + sqrt(..)) / A].
// Conservative (0.5 % on tau, 0.02 px on every bound); the exact rule stays per pixel.
__device__ __forceinline__ unsigned rdg_sub_mask(const float4 a, const float4 b, float tile_x0, float tile_y0) {
    const float A = a.z, B = a.w, C = b.x, o = b.y;
    if (!(o >= RDG_ALPHA_MIN)) return 0u;        // o * exp(power <= 0) can never reach 1/255
    const float tau = 2.0f * __logf(255.0f * o) * 1.005f + 2e-3f;
    const float det = A * C - B * B;
    if (!(det > 0.0f) || !(A > 0.0f)) return 0xffffu;
    const float tA = tau * A;
    const float ey = rdg_sqrt(__fdividef(tA, det)) * 1.001f + 0.02f;
    const float cx = a.x - tile_x0, cy = a.y - tile_y0;
    if (!(ey < 1e6f) || !(fabsf(cx) < 1e6f) || !(fabsf(cy) < 1e6f)) return 0xffffu;
    const int r0 = max(0, (int)ceilf(cy - ey)), r1 = min(RDG_TILE - 1, (int)floorf(cy + ey));
    const float invA = rdg_rcp(A);
    unsigned m = 0;
    for (int r = r0; r <= r1; ++r) {
        const float v = (float)r - cy;
        const float D = fmaxf(tA - v * v * det, 0.0f);
        const float h = rdg_sqrt(D) * invA * 1.001f + 0.02f;
        const float mid = cx - B * v * invA;
        const float lo = fmaxf(mid - h, -1.0f), hi = fminf(mid + h, (float)RDG_TILE);   // NaN -> whole row
        const int c0 = max(0, (int)ceilf(lo)), c1 = min(RDG_TILE - 1, (int)floorf(hi));
        if (c0 <= c1) m |= (((2u << (c1 >> 2)) - 1u) & ~((1u << (c0 >> 2)) - 1u)) << (4 * (r >> 2));
    }
    return m;
}

// sub-tile of quarter q of warp w: the warp owns a 2x2 block of sub-tiles (an 8x8 pixel region)
__device__ __forceinline__ int rdg_sub_of(int warp, int q) { return (2 * (warp >> 1) + (q >> 1)) * 4 + 2 * (warp & 1) + (q & 1); }

// Build the lists of this warp's four quarters from mask[0..cnt) (order preserved).
// live: bit q set = quarter q still has work.  WITH_E: append the gradient record index.
template <bool WITH_E>
__device__ __forceinline__ void rdg_compact4(Staged& sm, int cnt, int warp, int lane, unsigned live, int (&n)[4]) {
    n[0] = n[1] = n[2] = n[3] = 0;
    const unsigned lt = (1u << lane) - 1u;
    for (int g = 0; g * 32 < cnt; ++g) {
        const int j = g * 32 + lane;
        const unsigned m = (j < cnt) ? (unsigned)sm.mask[j] : 0u;
        const unsigned eb = WITH_E ? (unsigned)sm.ebase[j] : 0u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int s = rdg_sub_of(warp, q);
            const bool bit = ((m >> s) & 1u) && ((live >> q) & 1u);
            const unsigned bal = __ballot_sync(FULL, bit);
            if (bit) {
                unsigned e = (unsigned)j;
                if (WITH_E) e |= (eb + __popc(m & ((1u << s) - 1u))) << 7;
                sm.list[s][n[q] + __popc(bal & lt)] = (uint16_t)e;
            }
            n[q] += __popc(bal);
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(BATCH, 6) blend_fwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ vals,
                                                             const float4* __restrict__ p0, const float4* __restrict__ p1,
                                                             const float2* __restrict__ p2, const float* __restrict__ bg,
                                                             int W, int H, int gx, float* __restrict__ out_color,
                                                             float* __restrict__ out_depth, float* __restrict__ out_alpha,
                                                             float* __restrict__ out_T, uint32_t* __restrict__ out_ncontrib) {
    __shared__ Staged sm;
    const int tile = blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = lane >> 3, l8 = lane & 7;
    const int sub = rdg_sub_of(warp, q);
    const int pxi = tx * RDG_TILE + 4 * (sub & 3) + (l8 & 3);
    const int py0 = ty * RDG_TILE + 4 * (sub >> 2) + 2 * (l8 >> 2), py1 = py0 + 1;
    const bool in0 = pxi < W && py0 < H, in1 = pxi < W && py1 < H;
    const float pixx = (float)pxi, pixy0 = (float)py0, pixy1 = (float)py1;
    const float tile_x0 = (float)(tx * RDG_TILE), tile_y0 = (float)(ty * RDG_TILE);

    const uint2 range = ranges[tile];
    const int n_g = (int)(range.y - range.x);

    bool done0 = !in0, done1 = !in1;
    float T0 = 1.0f, r0 = 0.f, g0 = 0.f, b0 = 0.f, d0 = 0.f;
    float T1 = 1.0f, r1 = 0.f, g1 = 0.f, b1 = 0.f, d1 = 0.f;
    uint32_t last0 = 0, last1 = 0;
    const uint16_t* my_list = sm.list[sub];

    for (int base = 0; base < n_g; base += BATCH) {
        if (__syncthreads_count(done0 && done1) == BATCH) break;   // also orders the reuse of sm
        const int idx = base + (int)threadIdx.x;
        unsigned m = 0u;
        if (idx < n_g) {
            const uint32_t id = vals[range.x + idx];
            const float4 a = p0[id], b = p1[id];
            sm.a[threadIdx.x] = a;
            sm.b[threadIdx.x] = b;
            sm.c[threadIdx.x] = p2[id];
            m = rdg_sub_mask(a, b, tile_x0, tile_y0);
        }
        sm.mask[threadIdx.x] = (uint16_t)m;
        __syncthreads();
        const int cnt = min(BATCH, n_g - base);
        const unsigned act = __ballot_sync(FULL, !(done0 && done1));
        const unsigned live = ((act & 0xffu) ? 1u : 0u) | ((act & 0xff00u) ? 2u : 0u) | ((act & 0xff0000u) ? 4u : 0u) |
                              ((act & 0xff000000u) ? 8u : 0u);
        if (live == 0u) continue;                                 // this warp's 64 pixels are saturated
        int n[4];
        rdg_compact4<false>(sm, cnt, warp, lane, live, n);
        const int my_n = q == 0 ? n[0] : (q == 1 ? n[1] : (q == 2 ? n[2] : n[3]));
        const int nmax = max(max(n[0], n[1]), max(n[2], n[3]));
        for (int i = 0; i < nmax; ++i) {
            const bool valid = i < my_n;
            const int j = valid ? (int)my_list[i] : 0;
            const float4 a = sm.a[j];
            const float4 b = sm.b[j];
            const float2 c = sm.c[j];
            const float dx = a.x - pixx, dy0 = a.y - pixy0, dy1 = a.y - pixy1;
            const float Adx2 = (a.z * dx) * dx, Bdx = a.w * dx;
            const uint32_t here = (uint32_t)(base + j + 1);
            float G, alpha;
            {
                const bool on = rdg_alpha(Adx2, Bdx, b.x, dy0, b.y, G, alpha) && valid && !done0;
                const float test_T = T0 * (1.0f - alpha);
                const bool stop = on && (test_T < RDG_T_STOP);    // this Gaussian is not blended; the pixel is finished
                const bool upd = on && !stop;
                done0 = done0 || stop;
                const float wgt = upd ? alpha * T0 : 0.0f;
                r0 = __fmaf_rn(b.z, wgt, r0);
                g0 = __fmaf_rn(b.w, wgt, g0);
                b0 = __fmaf_rn(c.x, wgt, b0);
                d0 = __fmaf_rn(c.y, wgt, d0);
                T0 = upd ? test_T : T0;
                last0 = upd ? here : last0;
            }
            {
                const bool on = rdg_alpha(Adx2, Bdx, b.x, dy1, b.y, G, alpha) && valid && !done1;
                const float test_T = T1 * (1.0f - alpha);
                const bool stop = on && (test_T < RDG_T_STOP);
                const bool upd = on && !stop;
                done1 = done1 || stop;
                const float wgt = upd ? alpha * T1 : 0.0f;
                r1 = __fmaf_rn(b.z, wgt, r1);
                g1 = __fmaf_rn(b.w, wgt, g1);
                b1 = __fmaf_rn(c.x, wgt, b1);
                d1 = __fmaf_rn(c.y, wgt, d1);
                T1 = upd ? test_T : T1;
                last1 = upd ? here : last1;
            }
        }
    }
    const size_t hw = (size_t)H * W;
    const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
    if (in0) {
        const size_t pix = (size_t)py0 * W + pxi;
        out_color[pix] = r0 + T0 * bg0;
        out_color[hw + pix] = g0 + T0 * bg1;
        out_color[2 * hw + pix] = b0 + T0 * bg2;
        out_depth[pix] = d0;
        out_alpha[pix] = 1.0f - T0;
        out_T[pix] = T0;
        out_ncontrib[pix] = last0;
    }
    if (in1) {
        const size_t pix = (size_t)py1 * W + pxi;
        out_color[pix] = r1 + T1 * bg0;
        out_color[hw + pix] = g1 + T1 * bg1;
        out_color[2 * hw + pix] = b1 + T1 * bg2;
        out_depth[pix] = d1;
        out_alpha[pix] = 1.0f - T1;
        out_T[pix] = T1;
        out_ncontrib[pix] = last1;
    }
}

// ---------------------------------------------------------------- backward ----
#define NACC 12   // acc row: dpx dpy dA dB dC dop dr dg db ddepth pad pad

// Reduce-scatter of ten per-lane values over a quarter-warp (8 lanes) in 10 shuffles.
// On return lane l holds the quarter total of v[5*b2 + b1 + 2*b0] in `r_main` (b2 b1 b0 = bits
// of l & 7) and the lanes with (l & 3) == 0 hold the total of v[5*b2 + 4] in `r_extra`.
__device__ __forceinline__ void rdg_reduce_q10(const float (&v)[10], int lane, float& r_main, float& r_extra) {
    const bool b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
    float a[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const float keep = b2 ? v[i + 5] : v[i], send = b2 ? v[i] : v[i + 5];
        a[i] = keep + __shfl_xor_sync(FULL, send, 4);
    }
    const float k0 = b1 ? a[1] : a[0], s0 = b1 ? a[0] : a[1];
    const float k1 = b1 ? a[3] : a[2], s1 = b1 ? a[2] : a[3];
    const float q0 = k0 + __shfl_xor_sync(FULL, s0, 2);
    const float q1 = k1 + __shfl_xor_sync(FULL, s1, 2);
    const float q2 = a[4] + __shfl_xor_sync(FULL, a[4], 2);
    const float k = b0 ? q1 : q0, s = b0 ? q0 : q1;
    r_main = k + __shfl_xor_sync(FULL, s, 1);
    r_extra = q2 + __shfl_xor_sync(FULL, q2, 1);
}

__global__ void __launch_bounds__(BATCH, 6) blend_bwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ vals,
                                                             const float4* __restrict__ p0, const float4* __restrict__ p1,
                                                             const float2* __restrict__ p2, const float* __restrict__ bg,
                                                             int W, int H, int gx, const float* __restrict__ final_T,
                                                             const uint32_t* __restrict__ n_contrib,
                                                             const float* __restrict__ dL_dcolor, const float* __restrict__ dL_ddepth,
                                                             const float* __restrict__ dL_dalpha, float* __restrict__ acc) {
    __shared__ Staged sm;
    __shared__ __align__(16) float pool[POOL * PREC];
    __shared__ uint32_t qlast[SUBS];
    __shared__ int wsum[NWARP];
    const int tile = blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = lane >> 3, l8 = lane & 7;
    const int sub = rdg_sub_of(warp, q);
    const int pxi = tx * RDG_TILE + 4 * (sub & 3) + (l8 & 3);
    const int py0 = ty * RDG_TILE + 4 * (sub >> 2) + 2 * (l8 >> 2), py1 = py0 + 1;
    const bool in0 = pxi < W && py0 < H, in1 = pxi < W && py1 < H;
    const float pixx = (float)pxi, pixy0 = (float)py0, pixy1 = (float)py1;
    const float tile_x0 = (float)(tx * RDG_TILE), tile_y0 = (float)(ty * RDG_TILE);
    const size_t pix0 = (size_t)py0 * W + pxi, pix1 = (size_t)py1 * W + pxi, hw = (size_t)H * W;

    const uint2 range = ranges[tile];
    const int n_g = (int)(range.y - range.x);

    const float Tf0 = in0 ? final_T[pix0] : 0.f, Tf1 = in1 ? final_T[pix1] : 0.f;
    const uint32_t last0 = in0 ? n_contrib[pix0] : 0u, last1 = in1 ? n_contrib[pix1] : 0u;
    float gr0 = 0.f, gg0 = 0.f, gb0 = 0.f, gd0 = 0.f, ga0 = 0.f;
    float gr1 = 0.f, gg1 = 0.f, gb1 = 0.f, gd1 = 0.f, ga1 = 0.f;
    if (in0) {
        if (dL_dcolor) { gr0 = dL_dcolor[pix0]; gg0 = dL_dcolor[hw + pix0]; gb0 = dL_dcolor[2 * hw + pix0]; }
        if (dL_ddepth) gd0 = dL_ddepth[pix0];
        if (dL_dalpha) ga0 = dL_dalpha[pix0];
    }
    if (in1) {
        if (dL_dcolor) { gr1 = dL_dcolor[pix1]; gg1 = dL_dcolor[hw + pix1]; gb1 = dL_dcolor[2 * hw + pix1]; }
        if (dL_ddepth) gd1 = dL_ddepth[pix1];
        if (dL_dalpha) ga1 = dL_dalpha[pix1];
    }
    const float bgr = bg[0], bgg = bg[1], bgb = bg[2];

    // deepest contributor of every sub-tile: the CTA only walks back from the deepest one, and an
    // entry is dropped from the lists of the sub-tiles whose pixels all stopped in front of it
    {
        uint32_t ql = max(last0, last1);
        ql = max(ql, __shfl_xor_sync(FULL, ql, 1));
        ql = max(ql, __shfl_xor_sync(FULL, ql, 2));
        ql = max(ql, __shfl_xor_sync(FULL, ql, 4));
        if (l8 == 0) qlast[sub] = ql;
    }
    __syncthreads();
    uint32_t max_last = 0;
#pragma unroll
    for (int s = 0; s < SUBS; ++s) max_last = max(max_last, qlast[s]);
    max_last = min(max_last, (uint32_t)n_g);
    if (max_last == 0) return;

    // Running state per pixel.  With P_i = <g, (r,g,b,depth,1)_i> the alpha gradient is
    //   dL/dalpha_i = T_i P_i - (A_dot_i + T_final <g_rgb, bg>) / (1 - alpha_i),  A_dot_i = sum_{k>i} P_k alpha_k T_k,
    // so one scalar recursion replaces the five per-channel "accumulated behind" recursions.
    float T0 = Tf0, A0 = Tf0 * (bgr * gr0 + bgg * gg0 + bgb * gb0);
    float T1 = Tf1, A1 = Tf1 * (bgr * gr1 + bgg * gg1 + bgb * gb1);
    const int k_main = 5 * ((lane >> 2) & 1) + ((lane >> 1) & 1) + 2 * (lane & 1);
    const int k_extra = 5 * ((lane >> 2) & 1) + 4;
    const bool own_extra = (lane & 3) == 0;
    const uint16_t* my_list = sm.list[sub];

    int done_slots = 0;
    while (done_slots < (int)max_last) {
        // this round covers list positions pos = pos0 - slot, slot = 0..cnt-1 (back to front)
        const int cnt = min(BATCH, (int)max_last - done_slots);
        const int pos0 = (int)max_last - 1 - done_slots;
        __syncthreads();                                           // previous flush is done with sm / pool
        unsigned m = 0u;
        if ((int)threadIdx.x < cnt) {
            const uint32_t id = vals[range.x + pos0 - (int)threadIdx.x];
            const float4 a = p0[id], b = p1[id];
            sm.id[threadIdx.x] = id;
            sm.a[threadIdx.x] = a;
            sm.b[threadIdx.x] = b;
            sm.c[threadIdx.x] = p2[id];
            m = rdg_sub_mask(a, b, tile_x0, tile_y0);
            const uint32_t pos = (uint32_t)(pos0 - (int)threadIdx.x);
#pragma unroll
            for (int s = 0; s < SUBS; ++s)
                if (pos >= qlast[s]) m &= ~(1u << s);
        }
        // records are allotted entry-major: ebase = exclusive prefix of popc(mask) over the slots; the
        // round is cut where the pool would overflow (the rest is staged again by the next round)
        const int np = __popc(m);
        int incl = np;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < NWARP - 1; ++w)
            if (w < warp) incl += wsum[w];
        const bool keep = ((int)threadIdx.x < cnt) && (incl <= POOL);
        const int cnt2 = __syncthreads_count(keep);                // keep is a prefix: incl is non-decreasing
        if (!keep) m = 0u;
        sm.mask[threadIdx.x] = (uint16_t)m;
        sm.ebase[threadIdx.x] = (uint16_t)(incl - np);
        __syncthreads();

        int n[4];
        rdg_compact4<true>(sm, cnt2, warp, lane, 0xfu, n);
        const int my_n = q == 0 ? n[0] : (q == 1 ? n[1] : (q == 2 ? n[2] : n[3]));
        const int nmax = max(max(n[0], n[1]), max(n[2], n[3]));
        for (int i = 0; i < nmax; ++i) {
            const bool valid = i < my_n;
            const unsigned ent = valid ? (unsigned)my_list[i] : 0u;
            const int j = (int)(ent & 127u);
            const float4 a = sm.a[j];
            const float4 b = sm.b[j];
            const float2 c = sm.c[j];
            const uint32_t pos = (uint32_t)(pos0 - j);
            const float dx = a.x - pixx, dy0 = a.y - pixy0, dy1 = a.y - pixy1;
            const float Adx2 = (a.z * dx) * dx, Bdx = a.w * dx;
            float v[10];
            {
                float G, alpha;
                const bool on = rdg_alpha(Adx2, Bdx, b.x, dy0, b.y, G, alpha) && valid && (pos < last0);
                G = on ? G : 0.0f;
                alpha = on ? alpha : 0.0f;
                const float inv = on ? rdg_rcp(1.0f - alpha) : 1.0f;
                T0 *= inv;                                         // transmittance in front of this Gaussian
                const float wgt = alpha * T0;
                const float P = fmaf(b.z, gr0, fmaf(b.w, gg0, fmaf(c.x, gb0, fmaf(c.y, gd0, ga0))));
                const float dL_da = fmaf(T0, P, -inv * A0);
                A0 = fmaf(P, wgt, A0);
                // raw moments; the conic / sign factors are applied once per (Gaussian, tile) at the flush
                const float t5 = G * dL_da;                        // d/dopacity
                const float w = b.y * t5;                          // dL/dG * G
                const float wx = w * dx, wy = w * dy0;
                v[0] = wx; v[1] = wy;
                v[2] = wx * dx; v[3] = wx * dy0; v[4] = wy * dy0;
                v[5] = t5;
                v[6] = wgt * gr0; v[7] = wgt * gg0; v[8] = wgt * gb0;   // drgb
                v[9] = wgt * gd0;                                  // ddepth
            }
            {
                float G, alpha;
                const bool on = rdg_alpha(Adx2, Bdx, b.x, dy1, b.y, G, alpha) && valid && (pos < last1);
                G = on ? G : 0.0f;
                alpha = on ? alpha : 0.0f;
                const float inv = on ? rdg_rcp(1.0f - alpha) : 1.0f;
                T1 *= inv;
                const float wgt = alpha * T1;
                const float P = fmaf(b.z, gr1, fmaf(b.w, gg1, fmaf(c.x, gb1, fmaf(c.y, gd1, ga1))));
                const float dL_da = fmaf(T1, P, -inv * A1);
                A1 = fmaf(P, wgt, A1);
                const float t5 = G * dL_da;
                const float w = b.y * t5;
                const float wx = w * dx, wy = w * dy1;
                v[0] += wx; v[1] += wy;
                v[2] = fmaf(wx, dx, v[2]); v[3] = fmaf(wx, dy1, v[3]); v[4] = fmaf(wy, dy1, v[4]);
                v[5] += t5;
                v[6] = fmaf(wgt, gr1, v[6]); v[7] = fmaf(wgt, gg1, v[7]); v[8] = fmaf(wgt, gb1, v[8]);
                v[9] = fmaf(wgt, gd1, v[9]);
            }
            float r_main, r_extra;
            rdg_reduce_q10(v, lane, r_main, r_extra);
            if (valid) {
                float* rec = pool + (ent >> 7) * PREC;
                rec[k_main] = r_main;
                if (own_extra) rec[k_extra] = r_extra;
            }
        }
        __syncthreads();
        if ((int)threadIdx.x < cnt2) {
            const int ne = __popc((unsigned)sm.mask[threadIdx.x]);
            if (ne > 0) {
                const float4* row = reinterpret_cast<const float4*>(pool + (int)sm.ebase[threadIdx.x] * PREC);
                float4 s0 = row[0], s1 = row[1], s2 = row[2];
                for (int k = 1; k < ne; ++k) {
                    const float4 t0 = row[3 * k], t1 = row[3 * k + 1], t2 = row[3 * k + 2];
                    s0.x += t0.x; s0.y += t0.y; s0.z += t0.z; s0.w += t0.w;
                    s1.x += t1.x; s1.y += t1.y; s1.z += t1.z; s1.w += t1.w;
                    s2.x += t2.x; s2.y += t2.y;
                }
                // moments -> gradients: dpx = -(A Sx + B Sy), dpy = -(C Sy + B Sx), dA = -Sxx/2, dB = -Sxy, dC = -Syy/2
                const float4 ga4 = sm.a[threadIdx.x];
                const float cA = ga4.z, cB = ga4.w, cC = sm.b[threadIdx.x].x;
                const float sx = s0.x, sy = s0.y;
                s0.x = -(cA * sx + cB * sy);
                s0.y = -(cC * sy + cB * sx);
                s0.z *= -0.5f;
                s0.w = -s0.w;
                s1.x *= -0.5f;
                s2.z = 0.f;
                s2.w = 0.f;
                float4* dst = reinterpret_cast<float4*>(acc + (size_t)sm.id[threadIdx.x] * NACC);
                atomicAdd(dst + 0, s0);
                atomicAdd(dst + 1, s1);
                atomicAdd(dst + 2, s2);
            }
        }
        done_slots += cnt2;
    }
}

static bool rdg_use_v1() {
    static const bool v = [] { const char* e = getenv("RDG_BLEND_V1"); return e && e[0] == '1'; }();
    return v;
}

extern "C" int rdg_blend_fwd(int64_t n, const RdgGeom* geom, const RdgBins* bins, const RdgView* view,
                             const RdgImage* out, void* stream) {
    if (rdg_use_v1()) return rdg_blend_fwd_v1(n, geom, bins, view, out, stream);
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
    if (rdg_use_v1()) return rdg_blend_bwd_v1(n, geom, bins, view, fwd, dL_dcolor, dL_ddepth, dL_dalpha, acc, stream);
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
