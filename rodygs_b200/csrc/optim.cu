// Adam over the flat parameter buffer with one learning rate per parameter group, ONE launch
// per optimiser step.  Reference: torch.optim.Adam(l, lr=0.0, eps=1e-15) with the
// groups xyz / f_dc / f_rest / opacity / scaling / rotation (src/trainer/rodygs_static.py:106-141)
// + motion_coeff (src/trainer/rodygs_dynamic.py:92-123), stepped at src/trainer/rodygs.py:364-365 -
// there one multi-tensor pass per group and per moment (6-8 groups x ~6 kernels).
//
// B200 mapping: a persistent grid (RDG_SM_COUNT x 8 CTAs) walks the groups one after the other and
// streams each as float4: 16 B of param + grad + 2 moments in, param + 2 moments out = 28 B per
// float, nothing else.  The moments are read and written with evict-first hints (they are not touched
// again before the next step and the flat buffer is 4x the L2), the parameters are left cacheable for
// the next preprocess.  Bound: HBM.
#include <math.h>
#include "common.cuh"

#define ADAM_MAX_GROUPS 16

struct AdamArgs {
    int64_t begin[ADAM_MAX_GROUPS];
    int64_t end[ADAM_MAX_GROUPS];
    float step_size[ADAM_MAX_GROUPS];   // lr / (1 - beta1^t)
    int n_groups;
    float b1, b2, eps, bc2_sqrt, gscale;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float step_size, const AdamArgs& a) {
    g *= a.gscale;
    m = a.b1 * m + (1.f - a.b1) * g;
    v = a.b2 * v + (1.f - a.b2) * g * g;
    // torch.optim.Adam: denom = sqrt(v) / sqrt(bias2) + eps; param -= lr / bias1 * m / denom
    const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
    p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(RDG_BLOCK) adam_groups_kernel(float* __restrict__ param, const float* __restrict__ grad,
                                                                float* __restrict__ m, float* __restrict__ v,
                                                                const __grid_constant__ AdamArgs a) {
    const int64_t tid = (int64_t)blockIdx.x * RDG_BLOCK + threadIdx.x, nthr = (int64_t)gridDim.x * RDG_BLOCK;
    for (int k = 0; k < a.n_groups; ++k) {
        const int64_t b = a.begin[k], e = a.end[k];
        const float ss = a.step_size[k];
        const int64_t n4 = (e - b) >> 2;
        float4* p4 = reinterpret_cast<float4*>(param + b);
        const float4* g4 = reinterpret_cast<const float4*>(grad + b);
        float4* m4 = reinterpret_cast<float4*>(m + b);
        float4* v4 = reinterpret_cast<float4*>(v + b);
        for (int64_t i = tid; i < n4; i += nthr) {
            float4 p = p4[i], mm = __ldcs(m4 + i), vv = __ldcs(v4 + i);
            const float4 g = __ldcs(g4 + i);
            adam_one(p.x, g.x, mm.x, vv.x, ss, a);
            adam_one(p.y, g.y, mm.y, vv.y, ss, a);
            adam_one(p.z, g.z, mm.z, vv.z, ss, a);
            adam_one(p.w, g.w, mm.w, vv.w, ss, a);
            p4[i] = p;
            __stcs(m4 + i, mm);
            __stcs(v4 + i, vv);
        }
        const int64_t i = b + (n4 << 2) + tid;                    // ragged tail (< 4 floats)
        if (i < e) adam_one(param[i], grad[i], m[i], v[i], ss, a);
    }
}

extern "C" int rdg_adam_groups(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, const RdgAdamGroup* groups,
                               int32_t n_groups, float beta1, float beta2, float eps, int32_t step, float grad_scale,
                               void* stream) {
    RDG_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && groups, "null argument");
    RDG_CHECK_ARG(n_groups > 0 && n_groups <= ADAM_MAX_GROUPS, "1..16 groups");
    RDG_CHECK_ARG(step >= 1, "step must be >= 1");
    RDG_CHECK_ARG((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
                  "buffers must be 16-byte aligned");
    AdamArgs a;
    const float bc1 = 1.0f - (float)pow((double)beta1, (double)step);
    const float bc2 = 1.0f - (float)pow((double)beta2, (double)step);
    int64_t total = 0;
    for (int k = 0; k < n_groups; ++k) {
        RDG_CHECK_ARG(groups[k].begin >= 0 && groups[k].end >= groups[k].begin, "bad group range");
        RDG_CHECK_ARG((groups[k].begin & 3) == 0, "group offsets must be multiples of 4 floats");
        a.begin[k] = groups[k].begin;
        a.end[k] = groups[k].end;
        a.step_size[k] = groups[k].lr / bc1;
        total += groups[k].end - groups[k].begin;
    }
    if (total == 0) return RDG_OK;
    a.n_groups = n_groups;
    a.b1 = beta1; a.b2 = beta2; a.eps = eps; a.bc2_sqrt = sqrtf(bc2); a.gscale = grad_scale;
    const int64_t want = (total / 4 + RDG_BLOCK - 1) / RDG_BLOCK + 1;
    const int grid = (int)(want < (int64_t)RDG_SM_COUNT * 8 ? want : (int64_t)RDG_SM_COUNT * 8);
    adam_groups_kernel<<<grid, RDG_BLOCK, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, a);
    RDG_CHECK_LAUNCH();
    rdg_count_launches(1);
    return RDG_OK;
}
