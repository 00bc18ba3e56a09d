// Shared helpers for the rodygs_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/rodygs_b200.h"

#define RDG_BLOCK 256
#define RDG_SM_COUNT 148   // B200: 2 dies x 74 SMs; persistent grids are sized in multiples of this

void rdg_set_error(const char* fmt, ...);
void rdg_count_launches(int n);
enum { RDG_TUN_PRE_GRID_CAP = 0, RDG_TUN_DTABLE_V1 = 1, RDG_TUN_DIFF_SMEM = 2, RDG_TUN_DETERMINISTIC = 3, RDG_TUN_SM_RESERVE = 4, RDG_TUN_AR_UNROLL = 5, RDG_TUN_L2_PREFETCH = 6, RDG_TUN_COUNT = 7 };
// dynamic shared memory a preprocess CTA may ask for and still run 2 CTAs per SM (227 KB per SM, 1 KB reserved per CTA)
#define RDG_PRE_SMEM_MAX ((size_t)112 * 1024)
int rdg_tunable(int id);   // api.cu: environment default, rdg_set_tunable() override

#define RDG_CHECK_ARG(cond, msg)                                   \
    do {                                                           \
        if (!(cond)) {                                             \
            rdg_set_error("%s: %s", __func__, msg);                \
            return RDG_E_ARG;                                      \
        }                                                          \
    } while (0)

#define RDG_CHECK_LAUNCH()                                                              \
    do {                                                                                \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess) {                                                       \
            rdg_set_error("%s: CUDA error: %s", __func__, cudaGetErrorString(e__));     \
            return RDG_E_CUDA;                                                          \
        }                                                                               \
    } while (0)

#define RDG_CUDA(call)                                                                  \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            rdg_set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e__)); \
            return RDG_E_CUDA;                                                          \
        }                                                                               \
    } while (0)

static inline int64_t rdg_align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
static inline int rdg_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// SH constants: /root/reference/src/utils/sh_utils.py:24-41
#define RDG_SH_C0 0.28209479177387814f
#define RDG_SH_C1 0.4886025119029199f
#define RDG_SH_C2_0 1.0925484305920792f
#define RDG_SH_C2_1 -1.0925484305920792f
#define RDG_SH_C2_2 0.31539156525252005f
#define RDG_SH_C2_3 -1.0925484305920792f
#define RDG_SH_C2_4 0.5462742152960396f
#define RDG_SH_C3_0 -0.5900435899266435f
#define RDG_SH_C3_1 2.890611442640554f
#define RDG_SH_C3_2 -0.4570457994644658f
#define RDG_SH_C3_3 0.3731763325901154f
#define RDG_SH_C3_4 -0.4570457994644658f
#define RDG_SH_C3_5 1.445305721320277f
#define RDG_SH_C3_6 -0.5900435899266435f

#define RDG_NEAR_Z 0.2f
#define RDG_LOWPASS 0.3f
#define RDG_ALPHA_MAX 0.99f
#define RDG_ALPHA_MIN (1.0f / 255.0f)
#define RDG_T_STOP 0.0001f

#ifdef __CUDACC__
// The camera as the kernels see it: mathematical row-major V and P, unpacked
// from glm storage once per block.
struct RdgCam {
    float V[12];   // rows 0..2 of V (row 3 is 0,0,0,1)
    float P[16];
    float tanx, tany, fx, fy, limx, limy;
    float W, H;
    int gx, gy;
};

__device__ __forceinline__ void rdg_load_cam(RdgCam& c, const float* __restrict__ vm,
                                             const float* __restrict__ pm, float tanx, float tany,
                                             int width, int height) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) c.V[r * 4 + k] = vm[k * 4 + r];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) c.P[r * 4 + k] = pm[k * 4 + r];
    c.tanx = tanx;
    c.tany = tany;
    c.W = (float)width;
    c.H = (float)height;
    c.fx = c.W / (2.0f * tanx);
    c.fy = c.H / (2.0f * tany);
    c.limx = 1.3f * tanx;
    c.limy = 1.3f * tany;
    c.gx = (width + RDG_TILE - 1) / RDG_TILE;
    c.gy = (height + RDG_TILE - 1) / RDG_TILE;
}

// Lanes of `valid` that hold the same `bits`-bit digit as the caller.  One VOTE per digit bit:
// on sm_100 this is several times faster than match.any.sync (MATCH serialises over the distinct
// values, and radix digits are almost all distinct inside a warp) - ncu r01, tile_sort_kernel.
__device__ __forceinline__ unsigned rdg_match_digit(uint32_t d, int bits, unsigned valid) {
    unsigned peers = valid;
    for (int b = 0; b < bits; ++b) {
        const bool bit = (d >> b) & 1u;
        const unsigned m = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? m : ~m;
    }
    return peers;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

// SM-partitioned persistent kernels (preprocess_bwd.cu, sh_grad_views.cu): q[0] is the chunk counter, q[1] counts the CTAs
// that are done with it.  Called by one thread of EVERY CTA of the grid once it will not touch the queue again; the last one
// zeroes both words, so the next launch finds a clean queue without a memset between the kernels (a stream-ordered memset
// costs ~10 us of idle GPU per launch - the pipelined exchange launches six of these kernels back to back).
#ifdef __CUDACC__
__device__ __forceinline__ void rdg_queue_release(uint32_t* q, unsigned n_ctas) {
    __threadfence();
    if (atomicAdd(q + 1, 1u) == n_ctas - 1u) {
        q[0] = 0u;
        q[1] = 0u;
        __threadfence();
    }
}
#endif
