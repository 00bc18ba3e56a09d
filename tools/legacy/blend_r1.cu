// ROUND-1 VERSION of the blend kernels, kept for reference only: NOT part of the product library (rodygs_b200/build.py does not compile it).
// Front-to-back alpha blend (forward) and reverse-order backward over the
// depth-sorted per-tile lists.  SURVEY.md §8 rows a9 / a10; spec: SURVEY.md
// App. A.4-A.6 == oracle/splat_oracle.py::blend (+ autograd).
//
// B200 mapping (v3, profiles/README.md has the measured history)
//  * one CTA of 4 warps per 16x16 tile.  The tile is cut into sixteen 4x4-pixel sub-tiles; a
//    quarter-warp (8 lanes) owns one sub-tile and every lane owns TWO vertically adjacent
//    pixels, so the Gaussian record a lane fetches from shared memory (the crossbar is the
//    co-limiter of this loop, ncu r01) and the x-dependent half of the conic are used twice.
//  * per-tile batches of 128 list entries are staged in shared memory as three packed
//    records (16 + 16 + 8 bytes) by cp.async (LDGSTS) into a double buffer: batch k+1 is in
//    flight while batch k is blended, one barrier per batch in the forward pass.
//  * the staging thread solves the alpha >= 1/255 ellipse of its Gaussian against the four
//    4-row bands of the tile and turns it into a 16-bit mask of the sub-tiles it can reach
//    (conservative margins; the exact rule stays per pixel).  The forward pass stores the mask
//    per list entry (2 B) and the backward pass reads it back.  Each warp compacts the batch
//    into one list per quarter with ballot/popc and pads the four lists to a common length
//    with a null entry (opacity 0), so the inner loop has no per-quarter validity test; the
//    four quarters walk their own lists in lock step (3.0 (sub-tile, Gaussian) visits per list
//    entry at the bench config, 0.84 warp iterations per entry against 2.3 for the
//    8x4-per-warp bounding-box version).
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
namespace blend_r1 {

#define SUBS 16                 // 4x4-pixel sub-tiles per tile
#define BATCH 128               // list entries staged per round == threads per CTA
#define NWARP (BATCH / 32)
#define NSLOT (BATCH + 1)       // + the null slot (opacity 0) that pads the lists
#define LROW (BATCH + 2)        // list row stride (u16): rows of neighbouring sub-tiles start in different banks
#define POOL 512                // (sub-tile, entry) gradient records per round (backward)
#define PREC 10                 // floats per record
#define NULL_ENTRY ((uint16_t)BATCH)   // null slot (its ebase points at the scratch record), rank 0
#define FULL 0xffffffffu


struct __align__(16) Staged {
    float4 a[2][NSLOT];           // px, py, A, B          (double buffered: batch k+1 lands while batch k is blended)
    float4 b[2][NSLOT];           // C, opacity, r, g
    float2 c[2][NSLOT];           // b, depth
    uint16_t mask[2][BATCH];      // bit s: may touch sub-tile s (s = 4 * sub_y + sub_x)
    uint16_t ebase[NSLOT];        // backward: first gradient record of this entry
    uint16_t list[SUBS][LROW];    // per sub-tile compacted entries: slot | (rank of the sub-tile among the entry's << 8)
};

__device__ __forceinline__ float rdg_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rdg_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rdg_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// shared-window accesses by 32-bit address: the inner loops index three arrays with one offset
// and keep the window bases in registers
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
__device__ __forceinline__ void rdg_sts32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
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

// Packed FP32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: one issue slot for the same operation on the two
// pixels of a lane; a scalar operand is broadcast for free).  Individually rounded IEEE operations, so
// the packed and the scalar inner loops produce the same alpha bit for bit.
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t rdg_pk(float lo, float hi) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2_t rdg_bc(float x) { return rdg_pk(x, x); }
__device__ __forceinline__ void rdg_unpk(f2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2_t rdg_fma2(f2_t a, f2_t b, f2_t c) { f2_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2_t rdg_mul2(f2_t a, f2_t b) { f2_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2_t rdg_add2(f2_t a, f2_t b) { f2_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// power of the two pixels of a lane (same column, rows y0 and y0 + 1); npixy = (-y0, -y1).  Same operation
// order and roundings as rdg_alpha.
__device__ __forceinline__ f2_t rdg_power2(float Adx2, float Bdx, float C, float py, f2_t npixy, f2_t& dy) {
    dy = rdg_add2(rdg_bc(py), npixy);
    const f2_t q = rdg_fma2(rdg_mul2(rdg_bc(C), dy), dy, rdg_bc(Adx2));
    return rdg_fma2(rdg_bc(-0.5f), q, rdg_mul2(rdg_bc(-Bdx), dy));
}

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

// 16-bit mask of the 4x4 sub-tiles that the alpha >= 1/255 ellipse of this Gaussian reaches.
// With u = x - px, v = y - py the rule is  A u^2 + 2 B u v + C v^2 <= tau = 2 ln(255 o).  At height v
// the ellipse spans u-(v) .. u+(v) = (-B v -+ sqrt(tau A - v^2 det)) / A; u+ is concave and peaks at
// v* = -B ex / C with u+ = ex = sqrt(tau C / det), u- mirrors it.  So over the rows of one band the
// span is bounded by the values at the band's two edges (clamped to the ellipse's own height) and by
// +-ex when v* (-v*) falls inside the band: five edge evaluations for the four bands, no loop over
// rows.  Conservative (0.5 % on tau, half a pixel row on each band, 0.03 px on every bound); the exact
// rule stays per pixel.
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

// sub-tile of quarter q of warp w: the warp owns a 2x2 block of sub-tiles (an 8x8 pixel region)
__device__ __forceinline__ int rdg_sub_of(int warp, int q) { return (2 * (warp >> 1) + (q >> 1)) * 4 + 2 * (warp & 1) + (q & 1); }

// Build the lists of this warp's four quarters from mask[0..cnt) (order preserved) and pad them to a
// common length with the null entry.  live: bit q set = quarter q still has work.  WITH_E: append the
// rank of the sub-tile among the entry's sub-tiles (its gradient record is ebase[slot] + rank).
// Returns the common length.
template <bool WITH_E>
__device__ __forceinline__ int rdg_compact4(Staged& sm, const uint16_t* mask, int cnt, int warp, int lane, unsigned live) {
    int n[4] = {0, 0, 0, 0};
    const unsigned lt = (1u << lane) - 1u;
    for (int g = 0; g * 32 < cnt; ++g) {
        const int j = g * 32 + lane;
        const unsigned m = (j < cnt) ? (unsigned)mask[j] : 0u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int s = rdg_sub_of(warp, q);
            const bool bit = ((m >> s) & 1u) && ((live >> q) & 1u);
            const unsigned bal = __ballot_sync(FULL, bit);
            if (bit) {
                unsigned e = (unsigned)j;
                if (WITH_E) e |= (unsigned)__popc(m & ((1u << s) - 1u)) << 8;
                sm.list[s][n[q] + __popc(bal & lt)] = (uint16_t)e;
            }
            n[q] += __popc(bal);
        }
    }
    const int nmax = max(max(n[0], n[1]), max(n[2], n[3]));
    const int q = lane >> 3;
    const int my_n = q == 0 ? n[0] : (q == 1 ? n[1] : (q == 2 ? n[2] : n[3]));
    uint16_t* row = sm.list[rdg_sub_of(warp, q)];
    for (int k = my_n + (lane & 7); k < nmax; k += 8) row[k] = NULL_ENTRY;
    __syncwarp();
    return nmax;
}

// null slot: conic (0,0,1), opacity 0 -> alpha = 0 < 1/255, never contributes
__device__ __forceinline__ void rdg_init_null(Staged& sm) {
    if (threadIdx.x < 2) {
        sm.a[threadIdx.x][BATCH] = make_float4(0.f, 0.f, 0.f, 0.f);
        sm.b[threadIdx.x][BATCH] = make_float4(1.f, 0.f, 0.f, 0.f);
        sm.c[threadIdx.x][BATCH] = make_float2(0.f, 0.f);
        sm.ebase[BATCH] = (uint16_t)POOL;
    }
}

template <bool PK>
__global__ void __launch_bounds__(BATCH, 6) blend_fwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ vals,
                                                             const float4* __restrict__ p0, const float4* __restrict__ p1,
                                                             const float2* __restrict__ p2, const float* __restrict__ bg,
                                                             int W, int H, int gx, float* __restrict__ out_color,
                                                             float* __restrict__ out_depth, float* __restrict__ out_alpha,
                                                             float* __restrict__ out_T, uint32_t* __restrict__ out_ncontrib,
                                                             uint16_t* __restrict__ sub_masks,
                                                             const uint32_t* __restrict__ tile_order) {
    __shared__ Staged sm;
    const int tile = tile_order ? (int)tile_order[blockIdx.x] : (int)blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, q = lane >> 3, l8 = lane & 7;
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

    rdg_init_null(sm);
    const uint32_t sa = rdg_saddr(&sm.a[0][0]), sb = rdg_saddr(&sm.b[0][0]), sc = rdg_saddr(&sm.c[0][0]);
    const uint32_t my_list = rdg_saddr(&sm.list[sub][0]);
    // batch 0 in flight; the id of this thread's entry of batch 1 in a register
    uint32_t id_next = 0;
    if (tid < n_g) {
        const uint32_t id = vals[range.x + tid];
        rdg_cp16(sa + tid * 16, p0 + id);
        rdg_cp16(sb + tid * 16, p1 + id);
        rdg_cp8(sc + tid * 8, p2 + id);
    }
    rdg_cp_commit();
    if (BATCH + tid < n_g) id_next = vals[range.x + BATCH + tid];

    int buf = 0;
    for (int base = 0; base < n_g; base += BATCH, buf ^= 1) {
        rdg_cp_wait_all();
        const int idx = base + tid;
        unsigned m = 0u;
        if (idx < n_g) {
            m = rdg_sub_mask(sm.a[buf][tid], sm.b[buf][tid], tile_x0, tile_y0);
            if (sub_masks) sub_masks[range.x + idx] = (uint16_t)m;
        }
        sm.mask[buf][tid] = (uint16_t)m;
        // one barrier per batch: the copies and masks of this batch are visible, and every warp is done
        // with the other buffer
        if (__syncthreads_count(done0 && done1) == BATCH) {
            // every pixel is saturated: the remaining entries still need their masks for the backward pass
            if (sub_masks)
                for (int k = base + BATCH + tid; k < n_g; k += BATCH) sub_masks[range.x + k] = 0;
            break;
        }
        if (idx + BATCH < n_g) {
            const uint32_t o16 = (uint32_t)((buf ^ 1) * NSLOT + tid) * 16u;
            rdg_cp16(sa + o16, p0 + id_next);
            rdg_cp16(sb + o16, p1 + id_next);
            rdg_cp8(sc + (o16 >> 1), p2 + id_next);
        }
        rdg_cp_commit();
        if (idx + 2 * BATCH < n_g) id_next = vals[range.x + idx + 2 * BATCH];

        const int cnt = min(BATCH, n_g - base);
        const unsigned act = __ballot_sync(FULL, !(done0 && done1));
        const unsigned live = ((act & 0xffu) ? 1u : 0u) | ((act & 0xff00u) ? 2u : 0u) | ((act & 0xff0000u) ? 4u : 0u) |
                              ((act & 0xff000000u) ? 8u : 0u);
        if (live == 0u) continue;                                 // this warp's 64 pixels are saturated
        const int nmax = rdg_compact4<false>(sm, sm.mask[buf], cnt, warp, lane, live);
        const uint32_t boff = (uint32_t)(buf * NSLOT) * 16u;
        uint32_t j_next = rdg_lds16(my_list);
        if constexpr (PK) {
            const f2_t npixy = rdg_pk(-pixy0, -pixy1);
            f2_t T2 = rdg_pk(T0, T1), r2 = rdg_pk(r0, r1), g2 = rdg_pk(g0, g1), b2 = rdg_pk(b0, b1), d2 = rdg_pk(d0, d1);
            for (int i = 0; i < nmax; ++i) {
                const uint32_t j = j_next;
                j_next = rdg_lds16(my_list + 2u * (i + 1));
                const uint32_t o16 = boff + (j << 4);
                const float4 a = rdg_lds128(sa + o16);
                const float4 b = rdg_lds128(sb + o16);
                const float2 c = rdg_lds64(sc + (o16 >> 1));
                const float dx = a.x - pixx;
                const float Adx2 = (a.z * dx) * dx, Bdx = a.w * dx;
                const uint32_t here = (uint32_t)base + j + 1u;
                f2_t dy;
                const f2_t power = rdg_power2(Adx2, Bdx, b.x, a.y, npixy, dy);
                float pw0, pw1, e0, e1;
                rdg_unpk(power, pw0, pw1);
                rdg_unpk(rdg_mul2(power, rdg_bc(1.4426950408889634f)), e0, e1);
                float oG0, oG1;
                rdg_unpk(rdg_mul2(rdg_bc(b.y), rdg_pk(rdg_ex2(e0), rdg_ex2(e1))), oG0, oG1);
                const float al0 = fminf(RDG_ALPHA_MAX, oG0), al1 = fminf(RDG_ALPHA_MAX, oG1);
                const bool on0 = (pw0 <= 0.0f) && (al0 >= RDG_ALPHA_MIN) && !done0;
                const bool on1 = (pw1 <= 0.0f) && (al1 >= RDG_ALPHA_MIN) && !done1;
                const f2_t al2 = rdg_pk(al0, al1);
                const f2_t test2 = rdg_mul2(T2, rdg_fma2(al2, rdg_bc(-1.0f), rdg_bc(1.0f)));
                float tt0, tt1, w0, w1, Tc0, Tc1;
                rdg_unpk(test2, tt0, tt1);
                rdg_unpk(rdg_mul2(al2, T2), w0, w1);
                rdg_unpk(T2, Tc0, Tc1);
                const bool stop0 = on0 && (tt0 < RDG_T_STOP), stop1 = on1 && (tt1 < RDG_T_STOP);
                const bool upd0 = on0 && !stop0, upd1 = on1 && !stop1;
                done0 = done0 || stop0;
                done1 = done1 || stop1;
                const f2_t wgt2 = rdg_pk(upd0 ? w0 : 0.0f, upd1 ? w1 : 0.0f);
                r2 = rdg_fma2(rdg_bc(b.z), wgt2, r2);
                g2 = rdg_fma2(rdg_bc(b.w), wgt2, g2);
                b2 = rdg_fma2(rdg_bc(c.x), wgt2, b2);
                d2 = rdg_fma2(rdg_bc(c.y), wgt2, d2);
                T2 = rdg_pk(upd0 ? tt0 : Tc0, upd1 ? tt1 : Tc1);
                last0 = upd0 ? here : last0;
                last1 = upd1 ? here : last1;
            }
            rdg_unpk(T2, T0, T1);
            rdg_unpk(r2, r0, r1);
            rdg_unpk(g2, g0, g1);
            rdg_unpk(b2, b0, b1);
            rdg_unpk(d2, d0, d1);
        } else {
        for (int i = 0; i < nmax; ++i) {
            const uint32_t j = j_next;
            j_next = rdg_lds16(my_list + 2u * (i + 1));            // one entry ahead (rows are padded): off the critical path
            const uint32_t o16 = boff + (j << 4);
            const float4 a = rdg_lds128(sa + o16);
            const float4 b = rdg_lds128(sb + o16);
            const float2 c = rdg_lds64(sc + (o16 >> 1));
            const float dx = a.x - pixx, dy0 = a.y - pixy0, dy1 = a.y - pixy1;
            const float Adx2 = (a.z * dx) * dx, Bdx = a.w * dx;
            const uint32_t here = (uint32_t)base + j + 1u;
            float G, alpha;
            {
                const bool on = rdg_alpha(Adx2, Bdx, b.x, dy0, b.y, G, alpha) && !done0;
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
                const bool on = rdg_alpha(Adx2, Bdx, b.x, dy1, b.y, G, alpha) && !done1;
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
    }
    rdg_cp_wait_all();
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
template <bool PK = false>
__device__ __forceinline__ void rdg_reduce_q10(const float (&v)[10], int lane, float& r_main, float& r_extra) {
    const bool b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
    float a[5];
    if constexpr (PK) {
        float k[5], t[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            k[i] = b2 ? v[i + 5] : v[i];
            t[i] = __shfl_xor_sync(FULL, b2 ? v[i] : v[i + 5], 4);
        }
        rdg_unpk(rdg_add2(rdg_pk(k[0], k[1]), rdg_pk(t[0], t[1])), a[0], a[1]);
        rdg_unpk(rdg_add2(rdg_pk(k[2], k[3]), rdg_pk(t[2], t[3])), a[2], a[3]);
        a[4] = k[4] + t[4];
    } else {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const float keep = b2 ? v[i + 5] : v[i], send = b2 ? v[i] : v[i + 5];
            a[i] = keep + __shfl_xor_sync(FULL, send, 4);
        }
    }
    const float k0 = b1 ? a[1] : a[0], s0 = b1 ? a[0] : a[1];
    const float k1 = b1 ? a[3] : a[2], s1 = b1 ? a[2] : a[3];
    float q0, q1;
    if constexpr (PK) {
        rdg_unpk(rdg_add2(rdg_pk(k0, k1), rdg_pk(__shfl_xor_sync(FULL, s0, 2), __shfl_xor_sync(FULL, s1, 2))), q0, q1);
    } else {
        q0 = k0 + __shfl_xor_sync(FULL, s0, 2);
        q1 = k1 + __shfl_xor_sync(FULL, s1, 2);
    }
    const float q2 = a[4] + __shfl_xor_sync(FULL, a[4], 2);
    const float k = b0 ? q1 : q0, s = b0 ? q0 : q1;
    r_main = k + __shfl_xor_sync(FULL, s, 1);
    r_extra = q2 + __shfl_xor_sync(FULL, q2, 1);
}

// Issue the copies of list positions pos0 - slot (slot = tid) of one backward round into buffer `buf`.
__device__ __forceinline__ void rdg_bwd_issue(uint32_t sa, uint32_t sb, uint32_t sc, int buf, int tid, uint32_t id,
                                              const float4* p0, const float4* p1, const float2* p2) {
    const uint32_t o16 = (uint32_t)(buf * NSLOT + tid) * 16u;
    rdg_cp16(sa + o16, p0 + id);
    rdg_cp16(sb + o16, p1 + id);
    rdg_cp8(sc + (o16 >> 1), p2 + id);
}

template <bool PK>
__global__ void __launch_bounds__(BATCH, 6) blend_bwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ vals,
                                                             const float4* __restrict__ p0, const float4* __restrict__ p1,
                                                             const float2* __restrict__ p2, const float* __restrict__ bg,
                                                             int W, int H, int gx, const float* __restrict__ final_T,
                                                             const uint32_t* __restrict__ n_contrib,
                                                             const float* __restrict__ dL_dcolor, const float* __restrict__ dL_ddepth,
                                                             const float* __restrict__ dL_dalpha, float* __restrict__ acc,
                                                             const uint16_t* __restrict__ sub_masks,
                                                             const uint32_t* __restrict__ tile_order) {
    __shared__ Staged sm;
    __shared__ __align__(16) float pool[(POOL + 1) * PREC];       // + the scratch record of the null entry
    __shared__ uint32_t qlast[SUBS];
    __shared__ int wsum[NWARP];
    const int tile = tile_order ? (int)tile_order[blockIdx.x] : (int)blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, q = lane >> 3, l8 = lane & 7;
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
    rdg_init_null(sm);
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
    const uint32_t sa = rdg_saddr(&sm.a[0][0]), sb = rdg_saddr(&sm.b[0][0]), sc = rdg_saddr(&sm.c[0][0]);
    const uint32_t my_list = rdg_saddr(&sm.list[sub][0]);
    const uint32_t pool_main = rdg_saddr(pool) + 4u * (uint32_t)(5 * ((lane >> 2) & 1) + ((lane >> 1) & 1) + 2 * (lane & 1));
    const bool own_extra = (lane & 3) == 0;
    const uint32_t sebase = rdg_saddr(&sm.ebase[0]);

    // round r covers list positions pos = pos0 - slot, slot = 0..cnt-1 (back to front), pos0 = max_last-1-done_slots.
    // The copies of the round that starts at done_slots + BATCH are issued while this round is blended; if
    // the pool cut this round short they are simply issued again for the right positions.
    int done_slots = 0, buf = 0;
    uint32_t id_cur = (tid < (int)max_last) ? vals[range.x + (int)max_last - 1 - tid] : 0u;   // this thread's entry of the round
    if (tid < (int)max_last) rdg_bwd_issue(sa, sb, sc, 0, tid, id_cur, p0, p1, p2);
    rdg_cp_commit();
    int pf_start = 0;                                              // first slot of the round sitting in (or flying into) buffer `buf`
    uint32_t id_next = (BATCH + tid < (int)max_last) ? vals[range.x + (int)max_last - 1 - BATCH - tid] : 0u;
    while (done_slots < (int)max_last) {
        const int cnt = min(BATCH, (int)max_last - done_slots);
        const int pos0 = (int)max_last - 1 - done_slots;
        if (pf_start != done_slots) {                              // the previous round was cut short (uniform branch)
            id_cur = (tid < cnt) ? vals[range.x + pos0 - tid] : 0u;
            if (tid < cnt) rdg_bwd_issue(sa, sb, sc, buf, tid, id_cur, p0, p1, p2);
            rdg_cp_commit();
            pf_start = done_slots;
            id_next = (BATCH + tid < (int)max_last - done_slots) ? vals[range.x + pos0 - BATCH - tid] : 0u;
        }
        rdg_cp_wait_all();
        unsigned m = 0u;
        if (tid < cnt) {
            m = sub_masks ? (unsigned)sub_masks[range.x + pos0 - tid]
                          : rdg_sub_mask(sm.a[buf][tid], sm.b[buf][tid], tile_x0, tile_y0);
            const uint32_t pos = (uint32_t)(pos0 - tid);
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
        __syncthreads();                                           // also: every warp is done with the previous flush
#pragma unroll
        for (int w = 0; w < NWARP - 1; ++w)
            if (w < warp) incl += wsum[w];
        const bool keep = (tid < cnt) && (incl <= POOL);
        if (!keep) m = 0u;
        sm.mask[buf][tid] = (uint16_t)m;
        sm.ebase[tid] = (uint16_t)(incl - np);
        const int cnt2 = __syncthreads_count(keep);                // keep is a prefix: incl is non-decreasing
        // next round's copies (assuming no cut) fly during the blend
        if (done_slots + BATCH + tid < (int)max_last) rdg_bwd_issue(sa, sb, sc, buf ^ 1, tid, id_next, p0, p1, p2);
        rdg_cp_commit();
        const uint32_t id_pf = id_next;
        if (done_slots + 2 * BATCH + tid < (int)max_last) id_next = vals[range.x + pos0 - 2 * BATCH - tid];

        const int nmax = rdg_compact4<true>(sm, sm.mask[buf], cnt2, warp, lane, 0xfu);
        const uint32_t boff = (uint32_t)(buf * NSLOT) * 16u;
        uint32_t ent_next = rdg_lds16(my_list);
        if constexpr (PK) {
            const f2_t npixy = rdg_pk(-pixy0, -pixy1);
            const f2_t gr2 = rdg_pk(gr0, gr1), gg2 = rdg_pk(gg0, gg1), gb2 = rdg_pk(gb0, gb1), gd2 = rdg_pk(gd0, gd1),
                       ga2 = rdg_pk(ga0, ga1);
            f2_t T2 = rdg_pk(T0, T1), A2 = rdg_pk(A0, A1);
            for (int i = 0; i < nmax; ++i) {
                const uint32_t ent = ent_next;
                ent_next = rdg_lds16(my_list + 2u * (i + 1));
                const uint32_t j = ent & 0xffu;
                const uint32_t o16 = boff + (j << 4);
                const float4 a = rdg_lds128(sa + o16);
                const float4 b = rdg_lds128(sb + o16);
                const float2 c = rdg_lds64(sc + (o16 >> 1));
                const uint32_t rec = (rdg_lds16(sebase + 2u * j) + (ent >> 8)) * (uint32_t)(PREC * 4);
                const uint32_t pos = (uint32_t)pos0 - j;
                const float dx = a.x - pixx;
                const float Adx2 = (a.z * dx) * dx, Bdx = a.w * dx;
                f2_t dy2;
                const f2_t power = rdg_power2(Adx2, Bdx, b.x, a.y, npixy, dy2);
                float pw0, pw1, e0, e1;
                rdg_unpk(power, pw0, pw1);
                rdg_unpk(rdg_mul2(power, rdg_bc(1.4426950408889634f)), e0, e1);
                float G0 = rdg_ex2(e0), G1 = rdg_ex2(e1);
                float oG0, oG1;
                rdg_unpk(rdg_mul2(rdg_bc(b.y), rdg_pk(G0, G1)), oG0, oG1);
                float al0 = fminf(RDG_ALPHA_MAX, oG0), al1 = fminf(RDG_ALPHA_MAX, oG1);
                const bool on0 = (pw0 <= 0.0f) && (al0 >= RDG_ALPHA_MIN) && (pos < last0);
                const bool on1 = (pw1 <= 0.0f) && (al1 >= RDG_ALPHA_MIN) && (pos < last1);
                G0 = on0 ? G0 : 0.0f;
                G1 = on1 ? G1 : 0.0f;
                al0 = on0 ? al0 : 0.0f;
                al1 = on1 ? al1 : 0.0f;
                const f2_t al2 = rdg_pk(al0, al1);
                float om0, om1;
                rdg_unpk(rdg_fma2(al2, rdg_bc(-1.0f), rdg_bc(1.0f)), om0, om1);
                const float inv0 = on0 ? rdg_rcp(om0) : 1.0f, inv1 = on1 ? rdg_rcp(om1) : 1.0f;
                const f2_t inv2 = rdg_pk(inv0, inv1);
                T2 = rdg_mul2(T2, inv2);                            // transmittance in front of this Gaussian
                const f2_t wgt2 = rdg_mul2(al2, T2);
                const f2_t P2 = rdg_fma2(rdg_bc(b.z), gr2, rdg_fma2(rdg_bc(b.w), gg2, rdg_fma2(rdg_bc(c.x), gb2, rdg_fma2(rdg_bc(c.y), gd2, ga2))));
                const f2_t nIA = rdg_mul2(inv2, A2);
                float nia0, nia1;
                rdg_unpk(nIA, nia0, nia1);
                const f2_t dLda2 = rdg_fma2(T2, P2, rdg_pk(-nia0, -nia1));
                A2 = rdg_fma2(P2, wgt2, A2);
                const f2_t t52 = rdg_mul2(rdg_pk(G0, G1), dLda2);   // d/dopacity
                const f2_t w2 = rdg_mul2(rdg_bc(b.y), t52);         // dL/dG * G
                const f2_t wx2 = rdg_mul2(w2, rdg_bc(dx)), wy2 = rdg_mul2(w2, dy2);
                float wxa, wxb, wya, wyb, dya, dyb, t5a, t5b, wga, wgb;
                rdg_unpk(wx2, wxa, wxb);
                rdg_unpk(wy2, wya, wyb);
                rdg_unpk(dy2, dya, dyb);
                rdg_unpk(t52, t5a, t5b);
                rdg_unpk(wgt2, wga, wgb);
                float v[10];
                v[0] = wxa + wxb;
                v[1] = wya + wyb;
                v[2] = v[0] * dx;
                v[3] = fmaf(wxb, dyb, wxa * dya);
                v[4] = fmaf(wyb, dyb, wya * dya);
                v[5] = t5a + t5b;
                v[6] = fmaf(wgb, gr1, wga * gr0);
                v[7] = fmaf(wgb, gg1, wga * gg0);
                v[8] = fmaf(wgb, gb1, wga * gb0);
                v[9] = fmaf(wgb, gd1, wga * gd0);
                float r_main, r_extra;
                rdg_reduce_q10<true>(v, lane, r_main, r_extra);
                rdg_sts32(pool_main + rec, r_main);
                rdg_sts32_if(own_extra, pool_main + rec + 16u, r_extra);   // owner lanes (l & 3) == 0: main slot 5*b2, extra slot 5*b2 + 4
            }
            rdg_unpk(T2, T0, T1);
            rdg_unpk(A2, A0, A1);
        } else {
        for (int i = 0; i < nmax; ++i) {
            const uint32_t ent = ent_next;
            ent_next = rdg_lds16(my_list + 2u * (i + 1));          // one entry ahead (rows are padded): off the critical path
            const uint32_t j = ent & 0xffu;
            const uint32_t o16 = boff + (j << 4);
            const float4 a = rdg_lds128(sa + o16);
            const float4 b = rdg_lds128(sb + o16);
            const float2 c = rdg_lds64(sc + (o16 >> 1));
            const uint32_t rec = (rdg_lds16(sebase + 2u * j) + (ent >> 8)) * (uint32_t)(PREC * 4);
            const uint32_t pos = (uint32_t)pos0 - j;               // null slot: garbage, but its alpha test fails
            const float dx = a.x - pixx, dy0 = a.y - pixy0, dy1 = a.y - pixy1;
            const float Adx2 = (a.z * dx) * dx, Bdx = a.w * dx;
            float v[10];
            {
                float G, alpha;
                const bool on = rdg_alpha(Adx2, Bdx, b.x, dy0, b.y, G, alpha) && (pos < last0);
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
                const bool on = rdg_alpha(Adx2, Bdx, b.x, dy1, b.y, G, alpha) && (pos < last1);
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
            rdg_sts32(pool_main + rec, r_main);
            rdg_sts32_if(own_extra, pool_main + rec + 16u, r_extra);
        }
        }
        __syncthreads();
        if (tid < cnt2) {
            const int ne = __popc((unsigned)sm.mask[buf][tid]);
            if (ne > 0) {
                const float2* row = reinterpret_cast<const float2*>(pool + (int)sm.ebase[tid] * PREC);
                float2 s0 = row[0], s1 = row[1], s2 = row[2], s3 = row[3], s4 = row[4];
                for (int k = 1; k < ne; ++k) {
                    const float2 t0 = row[5 * k], t1 = row[5 * k + 1], t2 = row[5 * k + 2], t3 = row[5 * k + 3], t4 = row[5 * k + 4];
                    s0.x += t0.x; s0.y += t0.y; s1.x += t1.x; s1.y += t1.y; s2.x += t2.x; s2.y += t2.y;
                    s3.x += t3.x; s3.y += t3.y; s4.x += t4.x; s4.y += t4.y;
                }
                // moments -> gradients: dpx = -(A Sx + B Sy), dpy = -(C Sy + B Sx), dA = -Sxx/2, dB = -Sxy, dC = -Syy/2
                const float4 ga4 = sm.a[buf][tid];
                const float cA = ga4.z, cB = ga4.w, cC = sm.b[buf][tid].x;
                const float sx = s0.x, sy = s0.y;
                const float4 o0 = make_float4(-(cA * sx + cB * sy), -(cC * sy + cB * sx), -0.5f * s1.x, -s1.y);
                const float4 o1 = make_float4(-0.5f * s2.x, s2.y, s3.x, s3.y);
                const float4 o2 = make_float4(s4.x, s4.y, 0.f, 0.f);
                float4* dst = reinterpret_cast<float4*>(acc + (size_t)id_cur * NACC);
                atomicAdd(dst + 0, o0);
                atomicAdd(dst + 1, o1);
                atomicAdd(dst + 2, o2);
            }
        }
        done_slots += cnt2;
        if (cnt2 == cnt) {                                         // the prefetched round is the next one
            pf_start = done_slots;
            buf ^= 1;
            id_cur = id_pf;
        }
    }
    rdg_cp_wait_all();
}

// RDG_BLEND_PACKED=0 selects the scalar-FP32 inner loops (A/B switch; both produce the same alpha bits)
static bool rdg_use_packed() {
    static const bool v = [] { const char* e = getenv("RDG_BLEND_PACKED"); return !(e && e[0] == '0'); }();
    return v;
}

}  // namespace blend_r1
using namespace blend_r1;
int rdg_blend_fwd_r1(int64_t n, const RdgGeom* geom, const RdgBins* bins, const RdgView* view,
                             const RdgImage* out, void* stream) {
    RDG_CHECK_ARG(geom && bins && view && out, "null argument");
    RDG_CHECK_ARG(out->color && out->depth && out->alpha && out->final_T && out->n_contrib, "null image buffer");
    RDG_CHECK_ARG(view->bg, "null background");
    const int W = view->width, H = view->height;
    const int gx = (W + RDG_TILE - 1) / RDG_TILE, gy = (H + RDG_TILE - 1) / RDG_TILE;
    auto kern = rdg_use_packed() ? blend_fwd_kernel<true> : blend_fwd_kernel<false>;
    kern<<<gx * gy, BATCH, 0, (cudaStream_t)stream>>>(
        (const uint2*)bins->ranges, bins->vals_sorted, (const float4*)geom->p0, (const float4*)geom->p1,
        (const float2*)geom->p2, view->bg, W, H, gx, out->color, out->depth, out->alpha, out->final_T, out->n_contrib,
        (uint16_t*)nullptr, (const uint32_t*)nullptr);
    RDG_CHECK_LAUNCH();
    rdg_count_launches(1);
    return RDG_OK;
}

int rdg_blend_bwd_r1(int64_t n, const RdgGeom* geom, const RdgBins* bins, const RdgView* view,
                             const RdgImage* fwd, const float* dL_dcolor, const float* dL_ddepth,
                             const float* dL_dalpha, float* acc, void* stream) {
    RDG_CHECK_ARG(geom && bins && view && fwd && acc, "null argument");
    RDG_CHECK_ARG(fwd->final_T && fwd->n_contrib, "null forward state");
    const int W = view->width, H = view->height;
    const int gx = (W + RDG_TILE - 1) / RDG_TILE, gy = (H + RDG_TILE - 1) / RDG_TILE;
    auto kern = rdg_use_packed() ? blend_bwd_kernel<true> : blend_bwd_kernel<false>;
    kern<<<gx * gy, BATCH, 0, (cudaStream_t)stream>>>(
        (const uint2*)bins->ranges, bins->vals_sorted, (const float4*)geom->p0, (const float4*)geom->p1,
        (const float2*)geom->p2, view->bg, W, H, gx, fwd->final_T, fwd->n_contrib, dL_dcolor, dL_ddepth, dL_dalpha, acc,
        (uint16_t*)nullptr, (const uint32_t*)nullptr);
    RDG_CHECK_LAUNCH();
    rdg_count_launches(1);
    return RDG_OK;
}
