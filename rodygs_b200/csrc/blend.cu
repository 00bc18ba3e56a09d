// Front-to-back alpha blend (forward) and reverse-order backward over the depth-sorted per-tile lists.
// SURVEY.md §8 rows a9 / a10; spec: SURVEY.md App. A.4-A.6 == oracle/splat_oracle.py::blend (+ autograd).
//
// B200 mapping, round 2 (profiles/README.md has the measured history of both rounds):
//  * The 16x16 tile is cut into sixteen 4x4-pixel sub-tiles.  A WARP owns one REGION = a 16x8 half tile = eight
//    sub-tiles; four lanes own one sub-tile and every lane owns one COLUMN of four pixels, so the record a lane
//    fetches from shared memory, the column offset dx and the x-half of the conic are used four times, and the
//    per-Gaussian gradient sums are pre-added over four pixels before they cross lanes (round 1: two pixels per
//    lane, eight lanes per sub-tile - 127 SASS instructions per two pixels in the backward loop against 68 here).
//  * tile_split_kernel (one pass per frame, shared by both directions): solves the alpha >= 1/255 ellipse of every
//    list entry against the four 4-row bands of its tile (rdg_sub_mask, conservative; the exact rule stays per
//    pixel) and writes, per region, the depth-ordered list of the entries that can reach it with the 8-bit mask of
//    its sub-tiles.  Round 1 recomputed the masks and re-compacted the whole tile list in every warp of both blend
//    kernels (56 % / 43 % of their time went to staging, masks, compaction and CTA barriers - ncu source view).
//  * the blend kernels have NO block-level barrier: a warp walks its own region list in chunks of 64 entries,
//    stages their packed records (16 + 16 + 8 bytes) with cp.async (LDGSTS) into a private double buffer, compacts
//    the chunk into one list per sub-tile with ballot/popc and pads the eight lists to a common length with a null
//    entry (opacity 0), so the inner loop has no validity test; the eight lane groups walk their own lists in lock
//    step.
//  * backward: a lane adds the ten partial gradients of its four pixels, the four lanes of the sub-tile
//    reduce-scatter them in 8 shuffles (5 + 3) and every lane stores the totals it owns into a per-(sub-tile,
//    Gaussian) record in shared memory - plain stores, every record written exactly once.  At the end of the chunk
//    one lane per entry sums its records and issues three 16-byte vector atomics (red.global.add.v4.f32) per
//    (Gaussian, region).
//  * tensor cores were evaluated for the pixel reduction (mma.sync m16n8k8 tf32 with a hi/lo operand split, the
//    moments expressed against a constant pixel-coordinate matrix): tools/microbench_mma.cu measures 8.6 cycles per
//    HMMA per scheduler and 4.7 lost issue slots per HMMA inside an FP32 loop - six of them cost as much as the
//    shuffle reduce-scatter they would replace, so the reduction stays on shuffles (profiles/README.md, round 2).
// Bound: FP32/ALU issue + shared-memory crossbar; charged against the HBM roofline as north_star asks
// (algorithmic bytes: 44 B per duplicate + 28 B per pixel forward; 44 B per duplicate + 44 B per pixel + 48 B per
// visible Gaussian backward).
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"

#define RSUBS 8                 // 4x4-pixel sub-tiles per region (= lane groups of 4 per warp)
#define REGIONS 2               // regions (16x8 half tiles) per tile
#ifndef CH
#define CH 64                   // region-list entries staged per chunk (up to two per lane; 32 < CH <= 64)
#endif
#define NSLOT (CH + 1)          // + the null slot (opacity 0) that pads the lists
#define LROW (CH + 4)           // list row stride (u16): CH entries + the two null entries the record prefetch may touch
#ifndef POOL
#define POOL 192                // (sub-tile, entry) gradient records per chunk (backward); 64 entries average 154
#endif
#ifndef FWD_MINB
#define FWD_MINB 12             // __launch_bounds__ minimum CTAs per SM, forward / backward
#endif
#ifndef BWD_MINB
#define BWD_MINB 8
#endif
#define PREC 10                 // floats per record (40-byte rows)
#define RDG_PW_FLAG 0x80000000u // region-list id bit: this entry's power may round to a positive value, test it
#define NULL_ENTRY ((uint16_t)(CH | (POOL << 7)))   // list entry = slot | record index << 7; null: slot CH, scratch record POOL
#define FULL 0xffffffffu
#ifndef BLEND_WARPS
#define BLEND_WARPS 2           // warps per CTA; every warp blends one region on its own (no block-level synchronisation)
#endif

struct __align__(16) StagedW {      // one per warp
    float4 a[2][NSLOT];             // px, py, A, B          (double buffered: chunk k+1 lands while chunk k is blended)
    float4 b[2][NSLOT];             // C, opacity, r, g
    float2 c[2][NSLOT];             // b, depth
    uint16_t list[RSUBS][LROW];     // per sub-tile compacted entries: slot | (backward: gradient record of the visit << 7)
};

__device__ __forceinline__ float rdg_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rdg_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rdg_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// shared-window accesses by 32-bit address: the inner loops index three arrays with one offset
__device__ __forceinline__ uint32_t rdg_saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t rdg_lds16(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float2 rdg_lds64(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float4 rdg_lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void rdg_sts64(uint32_t addr, float x, float y) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void rdg_sts32_if(bool on, uint32_t addr, float v) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.f32 [%0], %1;\n\t}" ::"r"(addr), "f"(v), "r"((uint32_t)on) : "memory");
}
__device__ __forceinline__ void rdg_cp16(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void rdg_cp8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void rdg_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void rdg_cp_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Packed FP32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: one issue slot for the same operation on two pixels of a
// lane; a scalar operand is broadcast for free).  Individually rounded IEEE operations.
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t rdg_pk(float lo, float hi) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2_t rdg_bc(float x) { return rdg_pk(x, x); }
__device__ __forceinline__ void rdg_unpk(f2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2_t rdg_fma2(f2_t a, f2_t b, f2_t c) { f2_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2_t rdg_mul2(f2_t a, f2_t b) { f2_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2_t rdg_add2(f2_t a, f2_t b) { f2_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float rdg_hsum(f2_t v) { float lo, hi; rdg_unpk(v, lo, hi); return lo + hi; }

// o * exp(power) of two pixels of a lane (same column, rows y and y + 1; npixy = (-y, -(y+1))), with
//   power = -(A dx^2 + C dy^2)/2 - B dx dy,  Adx2 = (A dx) dx and Bdx = B dx shared by the whole column.
// ONE instruction sequence for both passes so that the skip decisions replayed by the backward pass match the
// forward ones bit for bit.  Returns o G (packed); power and dy are handed back for the tests / the gradients.
__device__ __forceinline__ f2_t rdg_og2(float Adx2, float Bdx, float C, float py, float o, f2_t npixy, f2_t& power, f2_t& dy) {
    dy = rdg_add2(rdg_bc(py), npixy);
    const f2_t q = rdg_fma2(rdg_mul2(rdg_bc(C), dy), dy, rdg_bc(Adx2));
    power = rdg_fma2(rdg_bc(-0.5f), q, rdg_mul2(rdg_bc(-Bdx), dy));
    float e0, e1;
    rdg_unpk(rdg_mul2(power, rdg_bc(1.4426950408889634f)), e0, e1);
    return rdg_mul2(rdg_bc(o), rdg_pk(rdg_ex2(e0), rdg_ex2(e1)));
}

// 16-bit mask of the 4x4 sub-tiles that the alpha >= 1/255 ellipse of this Gaussian reaches.
// With u = x - px, v = y - py the rule is  A u^2 + 2 B u v + C v^2 <= tau = 2 ln(255 o).  At height v
// the ellipse spans u-(v) .. u+(v) = (-B v -+ sqrt(tau A - v^2 det)) / A; u+ is concave and peaks at
// v* = -B ex / C with u+ = ex = sqrt(tau C / det), u- mirrors it.  So over the rows of one band the
// span is bounded by the values at the band's two edges (clamped to the ellipse's own height) and by
// +-ex when v* (-v*) falls inside the band: five edge evaluations for the four bands, no loop over
// rows.  Conservative (0.5 % on tau, half a pixel row on each band, 0.03 px on every bound); the exact
// rule stays per pixel.  Bit s = 4 * sub_y + sub_x.
__device__ __forceinline__ unsigned rdg_sub_mask(const float4 a, const float4 b, float tile_x0, float tile_y0) {
    const float A = a.z, B = a.w, C = b.x, o = b.y;
    if (!(o >= RDG_ALPHA_MIN)) return 0u;        // o * exp(power <= 0) can never reach 1/255
    const float tau = 2.0f * __logf(255.0f * o) * 1.005f + 2e-3f;
    const float det = A * C - B * B;
    if (!(det > 0.0f) || !(A > 0.0f)) return 0xffffu;
    const float tA = tau * A, inv_det = rdg_rcp(det);
    const float ey = rdg_sqrt(tA * inv_det) * 1.001f, ex = rdg_sqrt(tau * C * inv_det) * 1.001f;
    const float cx = a.x - tile_x0, cy = a.y - tile_y0;
    if (!(ey < 1e6f) || !(ex < 1e6f) || !(fabsf(cx) < 1e6f) || !(fabsf(cy) < 1e6f)) return 0xffffu;
    const float invA = rdg_rcp(A);
    const float vstar = -B * ex * rdg_rcp(C);
    const float mg = 0.03f + 1e-3f * ex;
    float up[5], um[5], vk[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        vk[k] = (float)(4 * k) - 0.5f - cy;
        const float vc = fminf(fmaxf(vk[k], -ey), ey);
        const float h = rdg_sqrt(fmaxf(tA - vc * vc * det, 0.0f)) * invA;
        const float mid = -B * vc * invA;
        up[k] = mid + h;
        um[k] = mid - h;
    }
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool rows = (vk[k] <= ey + 0.03f) && (vk[k + 1] >= -ey - 0.03f);
        const bool peak_r = (vstar >= vk[k] - 0.03f) && (vstar <= vk[k + 1] + 0.03f);
        const bool peak_l = (-vstar >= vk[k] - 0.03f) && (-vstar <= vk[k + 1] + 0.03f);
        const float hi = peak_r ? ex : fmaxf(up[k], up[k + 1]);
        const float lo = peak_l ? -ex : fminf(um[k], um[k + 1]);
        const float xl = fmaxf(cx + lo - mg, -1.0f), xh = fminf(cx + hi + mg, (float)RDG_TILE);
        const int c0 = max(0, (int)ceilf(xl)), c1 = min(RDG_TILE - 1, (int)floorf(xh));
        if (rows && c0 <= c1) m |= (((2u << (c1 >> 2)) - 1u) & ~((1u << (c0 >> 2)) - 1u)) << (4 * k);
    }
    return m;
}

// ------------------------------------------------------------- region lists ----
// One CTA per tile.  Region r of the tile (rows 8r .. 8r+7) gets, in list order, the entries whose mask has a bit in
// byte r: ids[r * stride + range.x + k], masks[...] = that byte; rcount[2 tile + r] = their number.
__global__ void __launch_bounds__(128) tile_split_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ vals,
                                                         const float4* __restrict__ p0, const float4* __restrict__ p1, int gx,
                                                         uint32_t* __restrict__ rl_ids, uint8_t* __restrict__ rl_masks,
                                                         int64_t stride, uint32_t* __restrict__ rcount) {
    __shared__ uint32_t cnt[2][4][REGIONS];
    const int tile = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tx = tile % gx, ty = tile / gx;
    const float tile_x0 = (float)(tx * RDG_TILE), tile_y0 = (float)(ty * RDG_TILE);
    const uint2 range = ranges[tile];
    const int n_g = (int)(range.y - range.x);
    const unsigned lt = (1u << lane) - 1u;
    uint32_t run0 = 0, run1 = 0;
    uint32_t id_next = tid < n_g ? vals[range.x + tid] : 0u;
    for (int base = 0, par = 0; base < n_g; base += 128, par ^= 1) {
        const int idx = base + tid;
        const uint32_t id = id_next;
        unsigned m = 0u;
        uint32_t flag = 0u;
        if (idx < n_g) {
            const float4 a = p0[id], b = p1[id];
            m = rdg_sub_mask(a, b, tile_x0, tile_y0);
            // power = -q(d)/2 with q the conic form.  Its evaluation (rdg_og2) carries at most 3 roundings per term, so the
            // computed value can only exceed 0 when lambda_min / lambda_max of the conic is below ~4e-7; 1e-5 leaves a 25x
            // margin.  lambda_min lambda_max = det, lambda_min + lambda_max = A + C.  Everything else skips the power test.
            const float A = a.z, B = a.w, C = b.x, tr = A + C;
            if (!(A * C - B * B > 1e-5f * tr * tr) || !(tr < 1e30f)) flag = RDG_PW_FLAG;
        }
        if (idx + 128 < n_g) id_next = vals[range.x + idx + 128];
        const unsigned m0 = m & 0xffu, m1 = m >> 8;
        const unsigned bal0 = __ballot_sync(FULL, m0 != 0u), bal1 = __ballot_sync(FULL, m1 != 0u);
        if (lane == 0) { cnt[par][warp][0] = __popc(bal0); cnt[par][warp][1] = __popc(bal1); }
        __syncthreads();
        uint32_t before0 = 0, before1 = 0, tot0 = 0, tot1 = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint32_t c0 = cnt[par][w][0], c1 = cnt[par][w][1];
            tot0 += c0; tot1 += c1;
            if (w < warp) { before0 += c0; before1 += c1; }
        }
        if (m0) {
            const size_t pos = (size_t)range.x + run0 + before0 + __popc(bal0 & lt);
            rl_ids[pos] = id | flag;
            rl_masks[pos] = (uint8_t)m0;
        }
        if (m1) {
            const size_t pos = (size_t)stride + range.x + run1 + before1 + __popc(bal1 & lt);
            rl_ids[pos] = id | flag;
            rl_masks[pos] = (uint8_t)m1;
        }
        run0 += tot0;
        run1 += tot1;
    }
    if (tid == 0) { rcount[2 * tile] = run0; rcount[2 * tile + 1] = run1; }
}

// Build the lists of this warp's eight sub-tiles from the masks of the chunk's 64 slots (two per lane: slot = lane
// and lane + 32; order preserved) and pad them to a common length with the null entry.  WITH_E: the entry also carries
// the index of the visit's gradient record, eb_* (first record of the slot's entry) + rank of the sub-tile among the
// entry's sub-tiles.  Returns the common length.
template <bool WITH_E>
__device__ __forceinline__ int rdg_compact8(StagedW& sm, unsigned m_lo, unsigned m_hi, int lane, unsigned eb_lo = 0, unsigned eb_hi = 0) {
    const unsigned lt = (1u << lane) - 1u;
    const int my_s = lane >> 2;
    int nmax = 0, my_n = 0;
#pragma unroll
    for (int s = 0; s < RSUBS; ++s) {
        const bool b_lo = (m_lo >> s) & 1u, b_hi = (m_hi >> s) & 1u;
        const unsigned bal_lo = __ballot_sync(FULL, b_lo), bal_hi = __ballot_sync(FULL, b_hi);
        const int n_lo = __popc(bal_lo), n_s = n_lo + __popc(bal_hi);
        if (b_lo) {
            unsigned e = (unsigned)lane;
            if (WITH_E) e |= (eb_lo + (unsigned)__popc(m_lo & ((1u << s) - 1u))) << 7;
            sm.list[s][__popc(bal_lo & lt)] = (uint16_t)e;
        }
        if (b_hi) {
            unsigned e = (unsigned)(lane + 32);
            if (WITH_E) e |= (eb_hi + (unsigned)__popc(m_hi & ((1u << s) - 1u))) << 7;
            sm.list[s][n_lo + __popc(bal_hi & lt)] = (uint16_t)e;
        }
        nmax = max(nmax, n_s);
        my_n = (s == my_s) ? n_s : my_n;
    }
    uint16_t* row = sm.list[my_s];
    for (int k = my_n + (lane & 3); k < nmax + 2; k += 4) row[k] = NULL_ENTRY;   // + 2: the loops fetch the record one entry ahead
    __syncwarp();
    return nmax;
}

// null slot: conic (0,0,1), opacity 0 -> o G = 0 < 1/255, never contributes
__device__ __forceinline__ void rdg_init_null(StagedW& sm, int lane) {
    if (lane < 2) {
        sm.a[lane][CH] = make_float4(0.f, 0.f, 0.f, 0.f);
        sm.b[lane][CH] = make_float4(1.f, 0.f, 0.f, 0.f);
        sm.c[lane][CH] = make_float2(0.f, 0.f);
    }
}

// copies of one slot's packed record
__device__ __forceinline__ void rdg_issue(uint32_t sa, uint32_t sb, uint32_t sc, int buf, int slot, uint32_t id,
                                          const float4* p0, const float4* p1, const float2* p2) {
    const uint32_t o16 = (uint32_t)(buf * NSLOT + slot) * 16u;
    id &= ~RDG_PW_FLAG;
    rdg_cp16(sa + o16, p0 + id);
    rdg_cp16(sb + o16, p1 + id);
    rdg_cp8(sc + (o16 >> 1), p2 + id);
}

// bit s (0..7) of the result = some lane of group s (lanes 4s .. 4s+3) has its bit set in `act`
__device__ __forceinline__ unsigned rdg_groups_of(unsigned act) {
    unsigned t = act | (act >> 1);
    t |= t >> 2;
    t &= 0x11111111u;
    t = (t | (t >> 3)) & 0x03030303u;
    t = (t | (t >> 6)) & 0x000f000fu;
    return (t | (t >> 12)) & 0xffu;
}

struct BlendGeo {
    int tile, region, warp, tx, ty, lane, s, x, pxi, py0;
    bool in[4];
};
__device__ __forceinline__ BlendGeo rdg_geo(int gx, int W, int H) {
    BlendGeo g;
    g.warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * BLEND_WARPS + g.warp;              // region index: 2 * tile + half
    g.tile = gw >> 1;
    g.region = gw & 1;
    g.lane = threadIdx.x & 31;
    g.tx = g.tile % gx;
    g.ty = g.tile / gx;
    g.s = g.lane >> 2;
    g.x = g.lane & 3;
    g.pxi = g.tx * RDG_TILE + 4 * (g.s & 3) + g.x;
    g.py0 = g.ty * RDG_TILE + 8 * g.region + 4 * (g.s >> 2);
#pragma unroll
    for (int k = 0; k < 4; ++k) g.in[k] = g.pxi < W && g.py0 + k < H;
    return g;
}

// ------------------------------------------------------------------ forward ----
// One pixel's skip / stop rules and state update (App. A.4) in six or seven predicated instructions:
//   on   = [power <= 0 &&] o G >= thr        thr = 1/255 while the pixel is live, NaN once it is finished (every compare fails);
//                                             the power test only runs in chunks that hold a flagged entry (RDG_PW_FLAG)
//   stop = on && T (1 - alpha) < 1e-4         -> thr := NaN carrying the entry's list position (the pixel's n_contrib)
//   upd  = on && !stop                        -> T := T (1 - alpha), weight alpha T; otherwise T stays and the weight is 0
template <bool PW>
__device__ __forceinline__ void rdg_fwd_rules(float pw, float oG, float tt, float w_in, float nanpos, float& thr, float& T, float& w) {
    if (PW)
        asm("{\n\t.reg .pred on, st;\n\t"
            "setp.ge.f32 on, %3, %0;\n\t"
            "setp.le.and.f32 on, %4, 0f00000000, on;\n\t"
            "setp.lt.and.f32 st, %5, 0f38D1B717, on;\n\t"
            "setp.ge.and.f32 on, %5, 0f38D1B717, on;\n\t"
            "selp.f32 %2, %6, 0f00000000, on;\n\t"
            "selp.f32 %1, %5, %1, on;\n\t"
            "selp.f32 %0, %7, %0, st;\n\t}"
            : "+f"(thr), "+f"(T), "=f"(w) : "f"(oG), "f"(pw), "f"(tt), "f"(w_in), "f"(nanpos));
    else
        asm("{\n\t.reg .pred on, st;\n\t"
            "setp.ge.f32 on, %3, %0;\n\t"
            "setp.lt.and.f32 st, %4, 0f38D1B717, on;\n\t"
            "setp.ge.and.f32 on, %4, 0f38D1B717, on;\n\t"
            "selp.f32 %2, %5, 0f00000000, on;\n\t"
            "selp.f32 %1, %4, %1, on;\n\t"
            "selp.f32 %0, %6, %0, st;\n\t}"
            : "+f"(thr), "+f"(T), "=f"(w) : "f"(oG), "f"(tt), "f"(w_in), "f"(nanpos));
}

// one pair of pixels (rows y, y+1 of the lane's column)
struct FwdPair {
    f2_t T, r, g, b, d, npixy;
    float thr0, thr1;
};
template <bool PW>
__device__ __forceinline__ void rdg_fwd_pair(FwdPair& p, float Adx2, float Bdx, const float4& b, const float2& c, float py, float nanpos) {
    f2_t power, dy;
    const f2_t og = rdg_og2(Adx2, Bdx, b.x, py, b.y, p.npixy, power, dy);
    float pw0, pw1, oG0, oG1;
    rdg_unpk(power, pw0, pw1);
    rdg_unpk(og, oG0, oG1);
    const f2_t al2 = rdg_pk(fminf(RDG_ALPHA_MAX, oG0), fminf(RDG_ALPHA_MAX, oG1));
    float tt0, tt1, w0, w1, T0, T1, wg0, wg1;
    rdg_unpk(rdg_mul2(p.T, rdg_fma2(al2, rdg_bc(-1.0f), rdg_bc(1.0f))), tt0, tt1);
    rdg_unpk(rdg_mul2(al2, p.T), w0, w1);
    rdg_unpk(p.T, T0, T1);
    rdg_fwd_rules<PW>(pw0, oG0, tt0, w0, nanpos, p.thr0, T0, wg0);
    rdg_fwd_rules<PW>(pw1, oG1, tt1, w1, nanpos, p.thr1, T1, wg1);
    const f2_t wgt2 = rdg_pk(wg0, wg1);
    p.r = rdg_fma2(rdg_bc(b.z), wgt2, p.r);
    p.g = rdg_fma2(rdg_bc(b.w), wgt2, p.g);
    p.b = rdg_fma2(rdg_bc(c.x), wgt2, p.b);
    p.d = rdg_fma2(rdg_bc(c.y), wgt2, p.d);
    p.T = rdg_pk(T0, T1);
}
#define RDG_NAN_POS 0x7fc00000u      // quiet NaN; the low 22 bits carry a list position
__device__ __forceinline__ bool rdg_finished(float thr) { return thr != thr; }

__global__ void __launch_bounds__(32 * BLEND_WARPS, FWD_MINB) blend_fwd_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ rl_ids, const uint8_t* __restrict__ rl_masks, int64_t stride,
    const uint32_t* __restrict__ rcount, const float4* __restrict__ p0, const float4* __restrict__ p1,
    const float2* __restrict__ p2, const float* __restrict__ bg, int W, int H, int gx, int n_tiles, float* __restrict__ out_color,
    float* __restrict__ out_depth, float* __restrict__ out_alpha, float* __restrict__ out_T, uint32_t* __restrict__ out_ncontrib) {
    __shared__ StagedW sm_all[BLEND_WARPS];
    const BlendGeo geo = rdg_geo(gx, W, H);
    if (geo.tile >= n_tiles) return;
    StagedW& sm = sm_all[geo.warp];
    const int lane = geo.lane;
    const bool hi_ok = lane + 32 < CH;                            // slot lane + 32 exists
    const float pixx = (float)geo.pxi;
    const int n_g = (int)rcount[REGIONS * geo.tile + geo.region];
    const size_t lbase = (size_t)geo.region * stride + ranges[geo.tile].x;
    const uint32_t* ids = rl_ids + lbase;
    const uint8_t* masks = rl_masks + lbase;

    FwdPair pa, pb;
    pa.T = pb.T = rdg_bc(1.0f);
    pa.r = pa.g = pa.b = pa.d = pb.r = pb.g = pb.b = pb.d = rdg_bc(0.0f);
    pa.npixy = rdg_pk(-(float)geo.py0, -(float)(geo.py0 + 1));
    pb.npixy = rdg_pk(-(float)(geo.py0 + 2), -(float)(geo.py0 + 3));
    // pixels outside the image start finished (position 0: the backward pass has nothing to replay for them)
    const float nan0 = __uint_as_float(RDG_NAN_POS);
    pa.thr0 = geo.in[0] ? RDG_ALPHA_MIN : nan0; pa.thr1 = geo.in[1] ? RDG_ALPHA_MIN : nan0;
    pb.thr0 = geo.in[2] ? RDG_ALPHA_MIN : nan0; pb.thr1 = geo.in[3] ? RDG_ALPHA_MIN : nan0;

    rdg_init_null(sm, lane);
    const uint32_t sa = rdg_saddr(&sm.a[0][0]), sb = rdg_saddr(&sm.b[0][0]), sc = rdg_saddr(&sm.c[0][0]);
    const uint32_t my_list = rdg_saddr(&sm.list[geo.s][0]);
    // chunk 0 in flight; ids / masks of chunk 1 in registers
    uint32_t id_lo = 0, id_hi = 0, mk_lo = 0, mk_hi = 0;
    if (lane < n_g) { id_lo = ids[lane]; mk_lo = masks[lane]; rdg_issue(sa, sb, sc, 0, lane, id_lo, p0, p1, p2); }
    if (hi_ok && lane + 32 < n_g) { id_hi = ids[lane + 32]; mk_hi = masks[lane + 32]; rdg_issue(sa, sb, sc, 0, lane + 32, id_hi, p0, p1, p2); }
    rdg_cp_commit();
    uint32_t idn_lo = 0, idn_hi = 0, mkn_lo = 0, mkn_hi = 0;
    if (CH + lane < n_g) { idn_lo = ids[CH + lane]; mkn_lo = masks[CH + lane]; }
    if (hi_ok && CH + lane + 32 < n_g) { idn_hi = ids[CH + lane + 32]; mkn_hi = masks[CH + lane + 32]; }

    int buf = 0;
    for (int base = 0; base < n_g; base += CH, buf ^= 1) {
        rdg_cp_wait_all();
        __syncwarp();                                              // the chunk's copies are visible, everybody is done with the other buffer
        const unsigned act = __ballot_sync(FULL, !(rdg_finished(pa.thr0) && rdg_finished(pa.thr1) && rdg_finished(pb.thr0) && rdg_finished(pb.thr1)));
        if (act == 0u) break;                                      // the region is saturated
        const unsigned live = rdg_groups_of(act);
        const unsigned m_lo = mk_lo & live, m_hi = mk_hi & live;
        const bool test_pw = __any_sync(FULL, ((id_lo | id_hi) & RDG_PW_FLAG) != 0u);
        // next chunk's copies fly during the blend
        if (base + CH + lane < n_g) rdg_issue(sa, sb, sc, buf ^ 1, lane, idn_lo, p0, p1, p2);
        if (hi_ok && base + CH + lane + 32 < n_g) rdg_issue(sa, sb, sc, buf ^ 1, lane + 32, idn_hi, p0, p1, p2);
        rdg_cp_commit();
        mk_lo = mkn_lo; mk_hi = mkn_hi; id_lo = idn_lo; id_hi = idn_hi;
        mkn_lo = mkn_hi = idn_lo = idn_hi = 0u;
        if (base + 2 * CH + lane < n_g) { idn_lo = ids[base + 2 * CH + lane]; mkn_lo = masks[base + 2 * CH + lane]; }
        if (hi_ok && base + 2 * CH + lane + 32 < n_g) { idn_hi = ids[base + 2 * CH + lane + 32]; mkn_hi = masks[base + 2 * CH + lane + 32]; }

        const int nmax = rdg_compact8<false>(sm, m_lo, m_hi, lane);
        const uint32_t boff = (uint32_t)(buf * NSLOT) * 16u;
        auto blend_lists = [&](auto pw_tag) {
            constexpr bool PW = decltype(pw_tag)::value;
            // software pipeline: the record of entry i + 1 is requested before entry i is blended (rows are padded with nulls)
            uint32_t j = rdg_lds16(my_list) & 0x7fu, j_next = rdg_lds16(my_list + 2u) & 0x7fu;
            float4 a = rdg_lds128(sa + boff + (j << 4)), b = rdg_lds128(sb + boff + (j << 4));
            float2 c = rdg_lds64(sc + ((boff + (j << 4)) >> 1));
            for (int i = 0; i < nmax; ++i) {
                const uint32_t o16n = boff + (j_next << 4);
                const float4 a_n = rdg_lds128(sa + o16n), b_n = rdg_lds128(sb + o16n);
                const float2 c_n = rdg_lds64(sc + (o16n >> 1));
                const uint32_t j_nn = rdg_lds16(my_list + 2u * (i + 2)) & 0x7fu;
                const float dx = a.x - pixx;
                const float Adx2 = (a.z * dx) * dx, Bdx = a.w * dx;
                const float nanpos = __uint_as_float(((uint32_t)base + j) | RDG_NAN_POS);   // position of this entry, should a pixel stop at it
                rdg_fwd_pair<PW>(pa, Adx2, Bdx, b, c, a.y, nanpos);
                rdg_fwd_pair<PW>(pb, Adx2, Bdx, b, c, a.y, nanpos);
                a = a_n; b = b_n; c = c_n; j = j_next; j_next = j_nn;
            }
        };
        if (test_pw) blend_lists(std::true_type{});
        else blend_lists(std::false_type{});
    }
    rdg_cp_wait_all();
    const size_t hw = (size_t)H * W;
    const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
    float T[4], r[4], g[4], bl[4], d[4];
    rdg_unpk(pa.T, T[0], T[1]); rdg_unpk(pb.T, T[2], T[3]);
    rdg_unpk(pa.r, r[0], r[1]); rdg_unpk(pb.r, r[2], r[3]);
    rdg_unpk(pa.g, g[0], g[1]); rdg_unpk(pb.g, g[2], g[3]);
    rdg_unpk(pa.b, bl[0], bl[1]); rdg_unpk(pb.b, bl[2], bl[3]);
    rdg_unpk(pa.d, d[0], d[1]); rdg_unpk(pb.d, d[2], d[3]);
    // n_contrib: the backward pass replays the region-list positions below it - the position of the entry the pixel
    // stopped at, or the whole list (entries that a rule skipped are skipped again by the same bits)
    const float thr[4] = {pa.thr0, pa.thr1, pb.thr0, pb.thr1};
    uint32_t last[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) last[k] = rdg_finished(thr[k]) ? (__float_as_uint(thr[k]) & 0x3fffffu) : (uint32_t)n_g;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (geo.in[k]) {
            const size_t pix = (size_t)(geo.py0 + k) * W + geo.pxi;
            out_color[pix] = r[k] + T[k] * bg0;
            out_color[hw + pix] = g[k] + T[k] * bg1;
            out_color[2 * hw + pix] = bl[k] + T[k] * bg2;
            out_depth[pix] = d[k];
            out_alpha[pix] = 1.0f - T[k];
            out_T[pix] = T[k];
            out_ncontrib[pix] = last[k];
        }
    }
}

// ---------------------------------------------------------------- backward ----
#define NACC 12   // acc row: dpx dpy dA dB dC dop dr dg db ddepth pad pad

// Reduce-scatter of ten per-lane values over the four lanes of a sub-tile in 8 shuffles.  On return the lane with
// (b1 b0) = bits 1, 0 of its index holds the group totals of
//   (0,0): v0 v1 v2   (0,1): v3 v4   (1,0): v5 v6 v7   (1,1): v8 v9      in r0, r1 (, r2 on the b0 = 0 lanes).
// Record layout (10 floats): every lane stores (r0, r1) as one 8-byte word at float 2 * (2 b1 + b0), the b0 = 0 lanes
// their r2 at float 8 + b1:   0 v0  1 v1 | 2 v3  3 v4 | 4 v5  5 v6 | 6 v8  7 v9 | 8 v2  9 v7
__device__ __forceinline__ void rdg_reduce_q4(const float (&v)[10], bool b1, bool b0, float& r0, float& r1, float& r2) {
    float k[5], t[5], a[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        k[i] = b1 ? v[i + 5] : v[i];
        t[i] = __shfl_xor_sync(FULL, b1 ? v[i] : v[i + 5], 2);
    }
    rdg_unpk(rdg_add2(rdg_pk(k[0], k[1]), rdg_pk(t[0], t[1])), a[0], a[1]);
    rdg_unpk(rdg_add2(rdg_pk(k[2], k[3]), rdg_pk(t[2], t[3])), a[2], a[3]);
    a[4] = k[4] + t[4];
    const float k0 = b0 ? a[3] : a[0], s0 = b0 ? a[0] : a[3];
    const float k1 = b0 ? a[4] : a[1], s1 = b0 ? a[1] : a[4];
    rdg_unpk(rdg_add2(rdg_pk(k0, k1), rdg_pk(__shfl_xor_sync(FULL, s0, 1), __shfl_xor_sync(FULL, s1, 1))), r0, r1);
    r2 = a[2] + __shfl_xor_sync(FULL, a[2], 1);
}

// Running state of one pair of pixels.  With P_i = <g, (r,g,b,depth,1)_i> the alpha gradient is
//   dL/dalpha_i = T_i P_i - (A_dot_i + T_final <g_rgb, bg>) / (1 - alpha_i),  A_dot_i = sum_{k>i} P_k alpha_k T_k,
// so one scalar recursion replaces the five per-channel "accumulated behind" recursions.
struct BwdPair {
    f2_t T, A, gr, gg, gb, gd, ga, npixy;
    uint32_t last0, last1;
};
// o G of a pixel that takes the entry, else 0:  [power <= 0 &&] o G >= 1/255 [&& position below the pixel's n_contrib].
// POS: rounds that lie entirely below every n_contrib of the warp run without the position test; PW: the power test
// only runs in rounds that hold a flagged entry (RDG_PW_FLAG).
template <bool POS, bool PW>
__device__ __forceinline__ float rdg_bwd_rules(float pw, float oG, uint32_t pos, uint32_t last) {
    float r;
    if (POS && PW)
        asm("{\n\t.reg .pred on;\n\t"
            "setp.lt.u32 on, %3, %4;\n\t"
            "setp.ge.and.f32 on, %1, 0f3B808081, on;\n\t"
            "setp.le.and.f32 on, %2, 0f00000000, on;\n\t"
            "selp.f32 %0, %1, 0f00000000, on;\n\t}"
            : "=f"(r) : "f"(oG), "f"(pw), "r"(pos), "r"(last));
    else if (POS)
        asm("{\n\t.reg .pred on;\n\t"
            "setp.lt.u32 on, %2, %3;\n\t"
            "setp.ge.and.f32 on, %1, 0f3B808081, on;\n\t"
            "selp.f32 %0, %1, 0f00000000, on;\n\t}"
            : "=f"(r) : "f"(oG), "r"(pos), "r"(last));
    else if (PW)
        asm("{\n\t.reg .pred on;\n\t"
            "setp.ge.f32 on, %1, 0f3B808081;\n\t"
            "setp.le.and.f32 on, %2, 0f00000000, on;\n\t"
            "selp.f32 %0, %1, 0f00000000, on;\n\t}"
            : "=f"(r) : "f"(oG), "f"(pw));
    else
        asm("{\n\t.reg .pred on;\n\t"
            "setp.ge.f32 on, %1, 0f3B808081;\n\t"
            "selp.f32 %0, %1, 0f00000000, on;\n\t}"
            : "=f"(r) : "f"(oG));
    return r;
}

// One entry on the lane's four pixels (two pairs): advances T and A_dot, returns per pair w = dL/dG * G, the blend weight
// alpha T and dy.  A pixel that skips the entry (rule or position) gets o G := 0, which makes alpha = 0, 1 / (1 - alpha)
// = 1 (exactly: rcp.approx(1) = 1), w = 0 and leaves T and A_dot untouched - one select instead of masking every product.
// The two pairs go through the stages side by side so that their MUFU / select chains overlap.
template <bool POS, bool PW>
__device__ __forceinline__ void rdg_bwd_quad(BwdPair& pa, BwdPair& pb, float Adx2, float Bdx, const float4& b, const float2& c, float py,
                                             uint32_t pos, f2_t& wa, f2_t& wb, f2_t& wga, f2_t& wgb, f2_t& dya, f2_t& dyb) {
    f2_t pwa, pwb;
    const f2_t oga = rdg_og2(Adx2, Bdx, b.x, py, b.y, pa.npixy, pwa, dya);
    const f2_t ogb = rdg_og2(Adx2, Bdx, b.x, py, b.y, pb.npixy, pwb, dyb);
    const f2_t Pa = rdg_fma2(rdg_bc(b.z), pa.gr, rdg_fma2(rdg_bc(b.w), pa.gg, rdg_fma2(rdg_bc(c.x), pa.gb, rdg_fma2(rdg_bc(c.y), pa.gd, pa.ga))));
    const f2_t Pb = rdg_fma2(rdg_bc(b.z), pb.gr, rdg_fma2(rdg_bc(b.w), pb.gg, rdg_fma2(rdg_bc(c.x), pb.gb, rdg_fma2(rdg_bc(c.y), pb.gd, pb.ga))));
    float pw[4], oG[4];
    rdg_unpk(pwa, pw[0], pw[1]); rdg_unpk(pwb, pw[2], pw[3]);
    rdg_unpk(oga, oG[0], oG[1]); rdg_unpk(ogb, oG[2], oG[3]);
    oG[0] = rdg_bwd_rules<POS, PW>(pw[0], oG[0], pos, pa.last0);
    oG[1] = rdg_bwd_rules<POS, PW>(pw[1], oG[1], pos, pa.last1);
    oG[2] = rdg_bwd_rules<POS, PW>(pw[2], oG[2], pos, pb.last0);
    oG[3] = rdg_bwd_rules<POS, PW>(pw[3], oG[3], pos, pb.last1);
    const f2_t ala = rdg_pk(fminf(RDG_ALPHA_MAX, oG[0]), fminf(RDG_ALPHA_MAX, oG[1]));
    const f2_t alb = rdg_pk(fminf(RDG_ALPHA_MAX, oG[2]), fminf(RDG_ALPHA_MAX, oG[3]));
    float om[4];
    rdg_unpk(rdg_fma2(ala, rdg_bc(-1.0f), rdg_bc(1.0f)), om[0], om[1]);
    rdg_unpk(rdg_fma2(alb, rdg_bc(-1.0f), rdg_bc(1.0f)), om[2], om[3]);
    const f2_t inva = rdg_pk(rdg_rcp(om[0]), rdg_rcp(om[1])), invb = rdg_pk(rdg_rcp(om[2]), rdg_rcp(om[3]));
    pa.T = rdg_mul2(pa.T, inva);                                   // transmittance in front of this Gaussian
    pb.T = rdg_mul2(pb.T, invb);
    wga = rdg_mul2(ala, pa.T);
    wgb = rdg_mul2(alb, pb.T);
    float nia[4];
    rdg_unpk(rdg_mul2(inva, pa.A), nia[0], nia[1]);
    rdg_unpk(rdg_mul2(invb, pb.A), nia[2], nia[3]);
    const f2_t dLa = rdg_fma2(pa.T, Pa, rdg_pk(-nia[0], -nia[1])), dLb = rdg_fma2(pb.T, Pb, rdg_pk(-nia[2], -nia[3]));
    pa.A = rdg_fma2(Pa, wga, pa.A);
    pb.A = rdg_fma2(Pb, wgb, pb.A);
    wa = rdg_mul2(rdg_pk(oG[0], oG[1]), dLa);                      // straight-through clamp: o G, not alpha (App. A.6 i)
    wb = rdg_mul2(rdg_pk(oG[2], oG[3]), dLb);
}

// DET: 0 = float atomics (the product path); 1 / 2 = the two passes of rdg_blend_bwd_deterministic: largest magnitude per
// accumulator, then exact integer accumulation scaled by it.
template <int DET>
__global__ void __launch_bounds__(32 * BLEND_WARPS, BWD_MINB) blend_bwd_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ rl_ids, const uint8_t* __restrict__ rl_masks, int64_t stride,
    const uint32_t* __restrict__ rcount, const float4* __restrict__ p0, const float4* __restrict__ p1,
    const float2* __restrict__ p2, const float* __restrict__ bg, int W, int H, int gx, int n_tiles, const float* __restrict__ final_T,
    const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dcolor, const float* __restrict__ dL_ddepth,
    const float* __restrict__ dL_dalpha, float* __restrict__ acc, uint32_t* __restrict__ det_max, long long* __restrict__ det_sum) {
    __shared__ StagedW sm_all[BLEND_WARPS];
    __shared__ __align__(16) float pool_all[BLEND_WARPS][(POOL + 1) * PREC];   // + the scratch record of the null entry
    const BlendGeo geo = rdg_geo(gx, W, H);
    if (geo.tile >= n_tiles) return;
    StagedW& sm = sm_all[geo.warp];
    float* pool = pool_all[geo.warp];
    const int lane = geo.lane;
    const bool hi_ok = lane + 32 < CH;                            // slot lane + 32 exists
    const float pixx = (float)geo.pxi;
    const int n_g = (int)rcount[REGIONS * geo.tile + geo.region];
    const size_t lbase = (size_t)geo.region * stride + ranges[geo.tile].x;
    const uint32_t* ids = rl_ids + lbase;
    const uint8_t* masks = rl_masks + lbase;
    const size_t hw = (size_t)H * W;

    float Tf[4], gr[4], gg[4], gb[4], gd[4], ga[4];
    uint32_t last[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        Tf[k] = gr[k] = gg[k] = gb[k] = gd[k] = ga[k] = 0.f;
        last[k] = 0u;
        if (geo.in[k]) {
            const size_t pix = (size_t)(geo.py0 + k) * W + geo.pxi;
            Tf[k] = final_T[pix];
            last[k] = n_contrib[pix];
            if (dL_dcolor) { gr[k] = dL_dcolor[pix]; gg[k] = dL_dcolor[hw + pix]; gb[k] = dL_dcolor[2 * hw + pix]; }
            if (dL_ddepth) gd[k] = dL_ddepth[pix];
            if (dL_dalpha) ga[k] = dL_dalpha[pix];
        }
    }
    const float bgr = bg[0], bgg = bg[1], bgb = bg[2];

    // deepest contributor of every sub-tile: the warp only walks back from the deepest one, and an entry is
    // dropped from the lists of the sub-tiles whose pixels all stopped in front of it
    uint32_t ql = max(max(last[0], last[1]), max(last[2], last[3]));
    ql = max(ql, __shfl_xor_sync(FULL, ql, 1));
    ql = max(ql, __shfl_xor_sync(FULL, ql, 2));
    uint32_t qlast[RSUBS], max_last = 0;
#pragma unroll
    for (int s = 0; s < RSUBS; ++s) {
        qlast[s] = __shfl_sync(FULL, ql, 4 * s);
        max_last = max(max_last, qlast[s]);
    }
    max_last = min(max_last, (uint32_t)n_g);
    if (max_last == 0) return;
    uint32_t min_last = 0xffffffffu;                               // smallest n_contrib over the warp's pixels inside the image
#pragma unroll
    for (int k = 0; k < 4; ++k) min_last = geo.in[k] ? min(min_last, last[k]) : min_last;
    min_last = __reduce_min_sync(FULL, min_last);
    const int n_l = (int)max_last;

    BwdPair pa, pb;
    pa.T = rdg_pk(Tf[0], Tf[1]);
    pb.T = rdg_pk(Tf[2], Tf[3]);
    pa.A = rdg_pk(Tf[0] * (bgr * gr[0] + bgg * gg[0] + bgb * gb[0]), Tf[1] * (bgr * gr[1] + bgg * gg[1] + bgb * gb[1]));
    pb.A = rdg_pk(Tf[2] * (bgr * gr[2] + bgg * gg[2] + bgb * gb[2]), Tf[3] * (bgr * gr[3] + bgg * gg[3] + bgb * gb[3]));
    pa.gr = rdg_pk(gr[0], gr[1]); pb.gr = rdg_pk(gr[2], gr[3]);
    pa.gg = rdg_pk(gg[0], gg[1]); pb.gg = rdg_pk(gg[2], gg[3]);
    pa.gb = rdg_pk(gb[0], gb[1]); pb.gb = rdg_pk(gb[2], gb[3]);
    pa.gd = rdg_pk(gd[0], gd[1]); pb.gd = rdg_pk(gd[2], gd[3]);
    pa.ga = rdg_pk(ga[0], ga[1]); pb.ga = rdg_pk(ga[2], ga[3]);
    pa.npixy = rdg_pk(-(float)geo.py0, -(float)(geo.py0 + 1));
    pb.npixy = rdg_pk(-(float)(geo.py0 + 2), -(float)(geo.py0 + 3));
    pa.last0 = last[0]; pa.last1 = last[1]; pb.last0 = last[2]; pb.last1 = last[3];

    rdg_init_null(sm, lane);
    const uint32_t sa = rdg_saddr(&sm.a[0][0]), sb = rdg_saddr(&sm.b[0][0]), sc = rdg_saddr(&sm.c[0][0]);
    const uint32_t my_list = rdg_saddr(&sm.list[geo.s][0]);
    const bool b1 = geo.x & 2, b0 = geo.x & 1;
    const uint32_t pool_mine = rdg_saddr(pool) + 8u * (uint32_t)geo.x, pool_third = rdg_saddr(pool) + 4u * (uint32_t)(8 + (b1 ? 1 : 0));
    uint32_t min_qlast = 0xffffffffu;
#pragma unroll
    for (int s = 0; s < RSUBS; ++s) min_qlast = min(min_qlast, qlast[s]);

    // A round covers list positions pos = pos0 - slot, slot = 0..cnt-1 (back to front), pos0 = n_l - 1 - done_slots; this
    // lane stages the slots `lane` and `lane + 32`.  The copies of the round that starts at done_slots + CH are issued
    // while this round is blended; if the pool cut this round short they are simply issued again for the right positions.
    int done_slots = 0, buf = 0, pf_start = 0;
    uint32_t idc_lo = 0, idc_hi = 0, mkc_lo = 0, mkc_hi = 0;       // this round's entries
    if (lane < n_l) { idc_lo = ids[n_l - 1 - lane]; mkc_lo = masks[n_l - 1 - lane]; rdg_issue(sa, sb, sc, 0, lane, idc_lo, p0, p1, p2); }
    if (hi_ok && lane + 32 < n_l) { idc_hi = ids[n_l - 33 - lane]; mkc_hi = masks[n_l - 33 - lane]; rdg_issue(sa, sb, sc, 0, lane + 32, idc_hi, p0, p1, p2); }
    rdg_cp_commit();
    uint32_t idn_lo = 0, idn_hi = 0, mkn_lo = 0, mkn_hi = 0;       // the next round's, assuming no cut
    if (CH + lane < n_l) { idn_lo = ids[n_l - 1 - CH - lane]; mkn_lo = masks[n_l - 1 - CH - lane]; }
    if (hi_ok && CH + lane + 32 < n_l) { idn_hi = ids[n_l - 33 - CH - lane]; mkn_hi = masks[n_l - 33 - CH - lane]; }

    while (done_slots < n_l) {
        const int cnt = min(CH, n_l - done_slots);
        const int pos0 = n_l - 1 - done_slots;
        if (pf_start != done_slots) {                              // the previous round was cut short (uniform branch)
            idc_lo = idc_hi = mkc_lo = mkc_hi = 0u;
            if (lane < cnt) { idc_lo = ids[pos0 - lane]; mkc_lo = masks[pos0 - lane]; rdg_issue(sa, sb, sc, buf, lane, idc_lo, p0, p1, p2); }
            if (hi_ok && lane + 32 < cnt) { idc_hi = ids[pos0 - 32 - lane]; mkc_hi = masks[pos0 - 32 - lane]; rdg_issue(sa, sb, sc, buf, lane + 32, idc_hi, p0, p1, p2); }
            rdg_cp_commit();
            pf_start = done_slots;
            idn_lo = idn_hi = mkn_lo = mkn_hi = 0u;
            if (CH + lane < n_l - done_slots) { idn_lo = ids[pos0 - CH - lane]; mkn_lo = masks[pos0 - CH - lane]; }
            if (hi_ok && CH + lane + 32 < n_l - done_slots) { idn_hi = ids[pos0 - CH - 32 - lane]; mkn_hi = masks[pos0 - CH - 32 - lane]; }
        }
        rdg_cp_wait_all();
        __syncwarp();                                              // copies visible; everybody is done with the previous flush
        unsigned m_lo = lane < cnt ? mkc_lo : 0u, m_hi = lane + 32 < cnt ? mkc_hi : 0u;   // (cnt <= CH)
        if ((uint32_t)pos0 >= min_qlast) {                         // some sub-tile stopped in front of this round (uniform branch)
            const uint32_t pos_lo = (uint32_t)(pos0 - lane), pos_hi = (uint32_t)(pos0 - 32 - lane);
#pragma unroll
            for (int s = 0; s < RSUBS; ++s) {
                if (pos_lo >= qlast[s]) m_lo &= ~(1u << s);
                if (pos_hi >= qlast[s]) m_hi &= ~(1u << s);
            }
        }
        const bool test_pw = __any_sync(FULL, ((idc_lo | idc_hi) & RDG_PW_FLAG) != 0u);
        // records are allotted entry-major: ebase = exclusive prefix of popc(mask) over the slots (both halves in one
        // scan: low half in bits 0..15, high half in bits 16..31); the round is cut where the pool would overflow (the
        // rest is staged again by the next round - rare: POOL holds the records of an average round with room to spare)
        const int np_lo = __popc(m_lo), np_hi = __popc(m_hi);
        int incl = np_lo | (np_hi << 16);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
        }
        const int incl_lo = incl & 0xffff;
        const int incl_hi = (incl >> 16) + (__shfl_sync(FULL, incl, 31) & 0xffff);
        const bool keep_lo = (lane < cnt) && (incl_lo <= POOL), keep_hi = (lane + 32 < cnt) && (incl_hi <= POOL);
        if (!keep_lo) m_lo = 0u;
        if (!keep_hi) m_hi = 0u;
        const unsigned eb_lo = (unsigned)(incl_lo - np_lo), eb_hi = (unsigned)(incl_hi - np_hi);   // first record of my two entries
        const int cnt2 = __popc(__ballot_sync(FULL, keep_lo)) + __popc(__ballot_sync(FULL, keep_hi));   // keep is a prefix
        // next round's copies (assuming no cut) fly during the blend
        if (done_slots + CH + lane < n_l) rdg_issue(sa, sb, sc, buf ^ 1, lane, idn_lo, p0, p1, p2);
        if (hi_ok && done_slots + CH + lane + 32 < n_l) rdg_issue(sa, sb, sc, buf ^ 1, lane + 32, idn_hi, p0, p1, p2);
        rdg_cp_commit();
        const uint32_t idp_lo = idn_lo, idp_hi = idn_hi, mkp_lo = mkn_lo, mkp_hi = mkn_hi;
        idn_lo = idn_hi = mkn_lo = mkn_hi = 0u;
        if (done_slots + 2 * CH + lane < n_l) { idn_lo = ids[pos0 - 2 * CH - lane]; mkn_lo = masks[pos0 - 2 * CH - lane]; }
        if (hi_ok && done_slots + 2 * CH + lane + 32 < n_l) { idn_hi = ids[pos0 - 2 * CH - 32 - lane]; mkn_hi = masks[pos0 - 2 * CH - 32 - lane]; }

        const int nmax = rdg_compact8<true>(sm, m_lo, m_hi, lane, eb_lo, eb_hi);
        const uint32_t boff = (uint32_t)(buf * NSLOT) * 16u;
        auto blend_lists = [&](auto pos_tag, auto pw_tag) {
            constexpr bool POS = decltype(pos_tag)::value, PW = decltype(pw_tag)::value;
            // software pipeline: the record of entry i + 1 is requested before entry i is blended (rows are padded with nulls)
            uint32_t ent = rdg_lds16(my_list), ent_next = rdg_lds16(my_list + 2u);
            float4 a = rdg_lds128(sa + boff + ((ent & 0x7fu) << 4)), b = rdg_lds128(sb + boff + ((ent & 0x7fu) << 4));
            float2 c = rdg_lds64(sc + ((boff + ((ent & 0x7fu) << 4)) >> 1));
            for (int i = 0; i < nmax; ++i) {
                const uint32_t o16n = boff + ((ent_next & 0x7fu) << 4);
                const float4 a_n = rdg_lds128(sa + o16n), b_n = rdg_lds128(sb + o16n);
                const float2 c_n = rdg_lds64(sc + (o16n >> 1));
                const uint32_t ent_nn = rdg_lds16(my_list + 2u * (i + 2));
                const uint32_t j = ent & 0x7fu;
                const uint32_t rec = (ent >> 7) * (uint32_t)(PREC * 4);
                const uint32_t pos = (uint32_t)pos0 - j;           // null slot: garbage, but its alpha test fails
                const float dx = a.x - pixx;
                const float Adx2 = (a.z * dx) * dx, Bdx = a.w * dx;
                f2_t wa, wb, wga, wgb, dya, dyb;
                rdg_bwd_quad<POS, PW>(pa, pb, Adx2, Bdx, b, c, a.y, pos, wa, wb, wga, wgb, dya, dyb);
                // raw moments about the Gaussian centre, pre-added over the column (dx is common to its four pixels);
                // the conic / sign factors are applied once per (Gaussian, region) at the flush
                const f2_t wya = rdg_mul2(wa, dya), wyb = rdg_mul2(wb, dyb);
                float v[10];
                v[5] = rdg_hsum(rdg_add2(wa, wb));                 // sum w = o * dL/dopacity
                v[1] = rdg_hsum(rdg_add2(wya, wyb));               // sum w dy
                v[4] = rdg_hsum(rdg_fma2(wyb, dyb, rdg_mul2(wya, dya)));   // sum w dy^2
                v[0] = v[5] * dx;                                  // sum w dx
                v[2] = v[0] * dx;                                  // sum w dx^2
                v[3] = v[1] * dx;                                  // sum w dx dy
                v[6] = rdg_hsum(rdg_fma2(wgb, pb.gr, rdg_mul2(wga, pa.gr)));
                v[7] = rdg_hsum(rdg_fma2(wgb, pb.gg, rdg_mul2(wga, pa.gg)));
                v[8] = rdg_hsum(rdg_fma2(wgb, pb.gb, rdg_mul2(wga, pa.gb)));
                v[9] = rdg_hsum(rdg_fma2(wgb, pb.gd, rdg_mul2(wga, pa.gd)));
                float r0, r1, r2;
                rdg_reduce_q4(v, b1, b0, r0, r1, r2);
                rdg_sts64(pool_mine + rec, r0, r1);
                rdg_sts32_if(!b0, pool_third + rec, r2);
                a = a_n; b = b_n; c = c_n; ent = ent_next; ent_next = ent_nn;
            }
        };
        // every position of this round lies below every n_contrib of the warp: no per-entry position test
        const bool test_pos = (uint32_t)pos0 >= min_last;
        if (test_pos) { if (test_pw) blend_lists(std::true_type{}, std::true_type{}); else blend_lists(std::true_type{}, std::false_type{}); }
        else { if (test_pw) blend_lists(std::false_type{}, std::true_type{}); else blend_lists(std::false_type{}, std::false_type{}); }
        __syncwarp();
        // flush: one lane per entry sums its records (floats: 0 Sx 1 Sy | 2 Sxy 3 Syy | 4 Sw 5 dr | 6 db 7 dd | 8 Sxx 9 dg)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const unsigned m = half ? m_hi : m_lo;
            const int ne = __popc(m), slot = lane + 32 * half;
            if (ne > 0) {
                const float2* row = reinterpret_cast<const float2*>(pool + (int)(half ? eb_hi : eb_lo) * PREC);
                float2 s0 = row[0], s1 = row[1], s2 = row[2], s3 = row[3], s4 = row[4];
                for (int k = 1; k < ne; ++k) {
                    const float2 t0 = row[5 * k], t1 = row[5 * k + 1], t2 = row[5 * k + 2], t3 = row[5 * k + 3], t4 = row[5 * k + 4];
                    s0.x += t0.x; s0.y += t0.y; s1.x += t1.x; s1.y += t1.y; s2.x += t2.x; s2.y += t2.y;
                    s3.x += t3.x; s3.y += t3.y; s4.x += t4.x; s4.y += t4.y;
                }
                // moments -> gradients: dpx = -(A Sx + B Sy), dpy = -(C Sy + B Sx), dA = -Sxx/2, dB = -Sxy, dC = -Syy/2,
                // dopacity = Sw / o (o >= 1/255 for every entry that got here)
                const float4 ga4 = sm.a[buf][slot], gb4 = sm.b[buf][slot];
                const float cA = ga4.z, cB = ga4.w, cC = gb4.x;
                const float sx = s0.x, sy = s0.y;
                const float4 o0 = make_float4(-(cA * sx + cB * sy), -(cC * sy + cB * sx), -0.5f * s4.x, -s1.x);
                const float4 o1 = make_float4(-0.5f * s1.y, s2.x / gb4.y, s2.y, s4.y);
                const float4 o2 = make_float4(s3.x, s3.y, 0.f, 0.f);
                const size_t arow = (size_t)((half ? idc_hi : idc_lo) & ~RDG_PW_FLAG) * NACC;
                if (DET == 0) {
                    float4* dst = reinterpret_cast<float4*>(acc + arow);
                    atomicAdd(dst + 0, o0);
                    atomicAdd(dst + 1, o1);
                    atomicAdd(dst + 2, o2);
                } else {
                    const float vals[10] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w, o2.x, o2.y};
#pragma unroll
                    for (int c = 0; c < 10; ++c) {
                        if (DET == 1) {
                            atomicMax(det_max + arow + c, __float_as_uint(fabsf(vals[c])));
                        } else {
                            const float mx = __uint_as_float(det_max[arow + c]);
                            if (mx > 0.0f && mx < 3.0e38f) {
                                int ex;
                                frexpf(mx, &ex);                   // mx = m 2^ex, m in [0.5, 1)
                                const long long q = __double2ll_rn((double)vals[c] * ldexp(1.0, 40 - ex));
                                atomicAdd(reinterpret_cast<unsigned long long*>(det_sum + arow + c), (unsigned long long)q);
                            }
                        }
                    }
                }
            }
        }
        done_slots += cnt2;
        if (cnt2 == cnt) {                                         // the prefetched round is the next one
            pf_start = done_slots;
            buf ^= 1;
            idc_lo = idp_lo; idc_hi = idp_hi; mkc_lo = mkp_lo; mkc_hi = mkp_hi;
        }
    }
    rdg_cp_wait_all();
}

extern "C" int rdg_blend_fwd(int64_t n, const RdgGeom* geom, const RdgBins* bins, const RdgView* view,
                             const RdgImage* out, void* stream) {
    RDG_CHECK_ARG(geom && bins && view && out, "null argument");
    RDG_CHECK_ARG(out->color && out->depth && out->alpha && out->final_T && out->n_contrib, "null image buffer");
    RDG_CHECK_ARG(view->bg, "null background");
    RDG_CHECK_ARG(bins->region_ids && bins->region_masks && bins->region_count && bins->region_stride > 0, "null region-list buffer");
    const int W = view->width, H = view->height;
    const int gx = (W + RDG_TILE - 1) / RDG_TILE, gy = (H + RDG_TILE - 1) / RDG_TILE;
    cudaStream_t s = (cudaStream_t)stream;
    tile_split_kernel<<<gx * gy, 128, 0, s>>>((const uint2*)bins->ranges, bins->vals_sorted, (const float4*)geom->p0,
                                              (const float4*)geom->p1, gx, bins->region_ids, bins->region_masks,
                                              bins->region_stride, bins->region_count);
    blend_fwd_kernel<<<(REGIONS * gx * gy + BLEND_WARPS - 1) / BLEND_WARPS, 32 * BLEND_WARPS, 0, s>>>(
        (const uint2*)bins->ranges, bins->region_ids, bins->region_masks, bins->region_stride, bins->region_count,
        (const float4*)geom->p0, (const float4*)geom->p1, (const float2*)geom->p2, view->bg, W, H, gx, gx * gy, out->color, out->depth,
        out->alpha, out->final_T, out->n_contrib);
    RDG_CHECK_LAUNCH();
    rdg_count_launches(2);
    return RDG_OK;
}

extern "C" int rdg_blend_bwd(int64_t n, const RdgGeom* geom, const RdgBins* bins, const RdgView* view,
                             const RdgImage* fwd, const float* dL_dcolor, const float* dL_ddepth,
                             const float* dL_dalpha, float* acc, void* stream) {
    RDG_CHECK_ARG(geom && bins && view && fwd && acc, "null argument");
    RDG_CHECK_ARG(fwd->final_T && fwd->n_contrib, "null forward state");
    RDG_CHECK_ARG(bins->region_ids && bins->region_masks && bins->region_count && bins->region_stride > 0, "null region-list buffer");
    const int W = view->width, H = view->height;
    const int gx = (W + RDG_TILE - 1) / RDG_TILE, gy = (H + RDG_TILE - 1) / RDG_TILE;
    blend_bwd_kernel<0><<<(REGIONS * gx * gy + BLEND_WARPS - 1) / BLEND_WARPS, 32 * BLEND_WARPS, 0, (cudaStream_t)stream>>>(
        (const uint2*)bins->ranges, bins->region_ids, bins->region_masks, bins->region_stride, bins->region_count,
        (const float4*)geom->p0, (const float4*)geom->p1, (const float2*)geom->p2, view->bg, W, H, gx, gx * gy, fwd->final_T,
        fwd->n_contrib, dL_dcolor, dL_ddepth, dL_dalpha, acc, nullptr, nullptr);
    RDG_CHECK_LAUNCH();
    rdg_count_launches(1);
    return RDG_OK;
}

// ---- deterministic variant (test mode) ----
__global__ void det_finalize_kernel(int64_t n_vals, const uint32_t* __restrict__ det_max, const long long* __restrict__ det_sum,
                                    float* __restrict__ acc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_vals) return;
    const float mx = __uint_as_float(det_max[i]);
    float v = 0.0f;
    if (mx > 0.0f && mx < 3.0e38f) {
        int ex;
        frexpf(mx, &ex);
        v = (float)((double)det_sum[i] * ldexp(1.0, ex - 40));
    }
    acc[i] = v;
}

extern "C" int64_t rdg_blend_bwd_deterministic_scratch_bytes(int64_t n) {
    if (n < 0) return RDG_E_ARG;
    return rdg_align_up(n * NACC * (int64_t)sizeof(long long), 256) + rdg_align_up(n * NACC * (int64_t)sizeof(uint32_t), 256);
}

extern "C" int rdg_blend_bwd_deterministic(int64_t n, const RdgGeom* geom, const RdgBins* bins, const RdgView* view,
                                           const RdgImage* fwd, const float* dL_dcolor, const float* dL_ddepth,
                                           const float* dL_dalpha, float* acc, void* scratch, int64_t scratch_bytes, void* stream) {
    RDG_CHECK_ARG(geom && bins && view && fwd && acc && scratch, "null argument");
    RDG_CHECK_ARG(fwd->final_T && fwd->n_contrib, "null forward state");
    RDG_CHECK_ARG(bins->region_ids && bins->region_masks && bins->region_count && bins->region_stride > 0, "null region-list buffer");
    if (scratch_bytes < rdg_blend_bwd_deterministic_scratch_bytes(n)) {
        rdg_set_error("rdg_blend_bwd_deterministic: scratch too small");
        return RDG_E_CAPACITY;
    }
    if (n == 0) return RDG_OK;
    const int W = view->width, H = view->height;
    const int gx = (W + RDG_TILE - 1) / RDG_TILE, gy = (H + RDG_TILE - 1) / RDG_TILE;
    cudaStream_t s = (cudaStream_t)stream;
    long long* det_sum = (long long*)scratch;
    uint32_t* det_max = (uint32_t*)((char*)scratch + rdg_align_up(n * NACC * (int64_t)sizeof(long long), 256));
    RDG_CUDA(cudaMemsetAsync(scratch, 0, (size_t)rdg_blend_bwd_deterministic_scratch_bytes(n), s));
    const int grid = (REGIONS * gx * gy + BLEND_WARPS - 1) / BLEND_WARPS;
    blend_bwd_kernel<1><<<grid, 32 * BLEND_WARPS, 0, s>>>(
        (const uint2*)bins->ranges, bins->region_ids, bins->region_masks, bins->region_stride, bins->region_count,
        (const float4*)geom->p0, (const float4*)geom->p1, (const float2*)geom->p2, view->bg, W, H, gx, gx * gy, fwd->final_T,
        fwd->n_contrib, dL_dcolor, dL_ddepth, dL_dalpha, acc, det_max, det_sum);
    blend_bwd_kernel<2><<<grid, 32 * BLEND_WARPS, 0, s>>>(
        (const uint2*)bins->ranges, bins->region_ids, bins->region_masks, bins->region_stride, bins->region_count,
        (const float4*)geom->p0, (const float4*)geom->p1, (const float2*)geom->p2, view->bg, W, H, gx, gx * gy, fwd->final_T,
        fwd->n_contrib, dL_dcolor, dL_ddepth, dL_dalpha, acc, det_max, det_sum);
    const int64_t n_vals = n * NACC;
    det_finalize_kernel<<<rdg_div_up(n_vals, 256), 256, 0, s>>>(n_vals, det_max, det_sum, acc);
    RDG_CHECK_LAUNCH();
    rdg_count_launches(3);
    return RDG_OK;
}
