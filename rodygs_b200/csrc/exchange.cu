// In-switch all-reduce of a gradient range over an NVLink-switch MULTICAST mapping.
// The range lives in symmetric memory (every rank's copy at the same offset of one multicast address).  Rank r owns the
// r-th of `world` equal slices: for every 16 bytes of its slice ONE multimem.ld_reduce makes the switch fetch the vector
// from all ranks and return the sum (no rank ever sees the other ranks' partial data), the mean's 1 / world rides on the
// result, and ONE multimem.st sends it back to every rank's copy.  Per rank: size / world in, size / world out over its
// NVLink port - against 2 (world - 1) / world x size each way for a ring.  The caller brackets the launch with cross-rank
// barriers (everybody's gradients written before / everybody's slice broadcast after).
// Bound: NVLink / NVSwitch bandwidth; SMs are only address generators here, so the grid is small (`ctas` x 512 threads)
// and leaves the machine to the per-Gaussian backward kernel that runs beside it.
#include "common.cuh"

__device__ __forceinline__ float4 mm_ld_reduce_v4(const float* addr) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void mm_st_v4(float* addr, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

#define AR_MAX_RANGES 8
struct ArRanges {
    int64_t v4_begin[AR_MAX_RANGES], v4_end[AR_MAX_RANGES];   // this rank's slice of every range, in 16-byte vectors from `mc`
    int n;
};

template <int AR_UNROLL>
__global__ void __launch_bounds__(512) allreduce_multimem_kernel(float* __restrict__ mc, const __grid_constant__ ArRanges r, float scale) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int k = 0; k < r.n; ++k) {
        const int64_t v4_end = r.v4_end[k];
        int64_t i = r.v4_begin[k] + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + (AR_UNROLL - 1) * stride < v4_end; i += AR_UNROLL * stride) {
            float4 v[AR_UNROLL];
#pragma unroll
            for (int u = 0; u < AR_UNROLL; ++u) v[u] = mm_ld_reduce_v4(mc + 4 * (i + u * stride));
#pragma unroll
            for (int u = 0; u < AR_UNROLL; ++u) {
                v[u].x *= scale; v[u].y *= scale; v[u].z *= scale; v[u].w *= scale;
                mm_st_v4(mc + 4 * (i + u * stride), v[u]);
            }
        }
        for (; i < v4_end; i += stride) {
            float4 v = mm_ld_reduce_v4(mc + 4 * i);
            v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
            mm_st_v4(mc + 4 * i, v);
        }
    }
    __threadfence_system();
}

// Several ranges of one symmetric buffer in ONE launch (the structure-of-arrays fields of a slice of Gaussians: xyz, scaling,
// rotation, opacity, motion coefficients): offsets / lens in floats from mc_base, each 16-byte aligned and a multiple of 4.
extern "C" int rdg_allreduce_multimem_ranges(float* mc_base, const int64_t* offsets, const int64_t* lens, int32_t n_ranges,
                                             int32_t rank, int32_t world, float scale, int32_t ctas, void* stream) {
    RDG_CHECK_ARG(mc_base && offsets && lens && world > 0 && rank >= 0 && rank < world, "bad argument");
    RDG_CHECK_ARG(n_ranges >= 0 && n_ranges <= AR_MAX_RANGES, "at most 8 ranges per launch");
    RDG_CHECK_ARG(((uintptr_t)mc_base & 15u) == 0, "the buffer must be 16-byte aligned");
    ArRanges r;
    r.n = 0;
    for (int k = 0; k < n_ranges; ++k) {
        RDG_CHECK_ARG(offsets[k] >= 0 && lens[k] >= 0 && (offsets[k] & 3) == 0 && (lens[k] & 3) == 0,
                      "ranges must start on a 16-byte boundary and hold a multiple of 4 floats");
        const int64_t n4 = lens[k] >> 2, per = (n4 + world - 1) / world;
        const int64_t b = per * rank < n4 ? per * rank : n4, e = b + per < n4 ? b + per : n4;
        if (e > b) { r.v4_begin[r.n] = (offsets[k] >> 2) + b; r.v4_end[r.n] = (offsets[k] >> 2) + e; ++r.n; }
    }
    if (r.n == 0) return RDG_OK;
    if (ctas <= 0) ctas = 32;
    if (ctas > RDG_SM_COUNT * 2) ctas = RDG_SM_COUNT * 2;
    cudaStream_t s = (cudaStream_t)stream;
    switch (rdg_tunable(RDG_TUN_AR_UNROLL)) {      // 16-byte vectors in flight per thread
        case 8: allreduce_multimem_kernel<8><<<ctas, 512, 0, s>>>(mc_base, r, scale); break;
        case 2: allreduce_multimem_kernel<2><<<ctas, 512, 0, s>>>(mc_base, r, scale); break;
        default: allreduce_multimem_kernel<4><<<ctas, 512, 0, s>>>(mc_base, r, scale); break;
    }
    RDG_CHECK_LAUNCH();
    rdg_count_launches(1);
    return RDG_OK;
}

extern "C" int rdg_allreduce_multimem(float* mc_range, int64_t n_floats, int32_t rank, int32_t world, float scale, int32_t ctas,
                                      void* stream) {
    RDG_CHECK_ARG(mc_range != nullptr, "null argument");
    const int64_t off = 0;
    return rdg_allreduce_multimem_ranges(mc_range, &off, &n_floats, 1, rank, world, scale, ctas, stream);
}
