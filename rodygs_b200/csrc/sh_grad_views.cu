// dL/dSH of ALL camera-time views of a data-parallel step from its rank-1 factors.
// the step's views are sharded over the GPUs and the Gaussian-parameter gradients are
// summed.  81 % of that message is dL/dSH [N,16,3], and it is an outer product per view:
//   dL/dSH_i[k][c] = sum_v  Y_k(dir_i,v) * g_i,v[c],     dir_i,v = normalize(x_i(t_v) - campos_v),
// where g_i,v = dL/d(rgb) of Gaussian i in view v after the SH clamp mask (12 bytes).  Every rank holds
// all parameters and all cameras, so it can evaluate Y_k itself: the ranks exchange the 12-byte factors
// (all-gather, written by rdg_preprocess_bwd into RdgSceneGrad.dcolor) instead of all-reducing the
// 192-byte blocks, and this kernel rebuilds the sum over every view.  Same gradient as the sum of the
// per-view dSH that rdg_preprocess_bwd writes on one GPU, up to float summation order.
//
// HBM-bound: reads 12 B (xyz) + 12 B per view [+ 68 B coefficients / birth frame for a dynamic Gaussian],
// writes 12*K B.  One thread per Gaussian, 256-Gaussian chunks of one model; the chunk's 46 KB of dSH
// rows leave through shared memory by one TMA bulk store (coalesced, no LSU store instructions).
#include <stdlib.h>
#include "scene.cuh"
#include "tma.cuh"

#define SH_ROW 45
#define SGV_MAX_VIEWS 16

struct ShGradParams {
    RdgScene sc;
    RdgSetGrad st, dy;
    const float* viewmats;   // [V,16] glm storage
    const float* basis_ts;   // [V,K,7] B(t_v) (dynamic + deform only)
    const float* dcolor;     // [V,N,3]
    int n_views;
    int sh_degree;
    float scale;
    int use_tma;
    int tab_smem;            // 1: the translation columns of the whole table are staged in shared memory (num_times <= SGV_TAB_MAX_T)
};
#define SGV_TAB_MAX_T 144

template <int DEG>
__global__ void __launch_bounds__(RDG_BLOCK, 2) sh_grad_views_kernel(const ShGradParams p) {
    extern __shared__ __align__(128) float smem[];
    constexpr int K = (DEG + 1) * (DEG + 1);
    const RdgScene& sc = p.sc;
    float* sh_s = smem;                                   // [256][SH_ROW]
    float* coef_s = smem + RDG_BLOCK * SH_ROW;            // [256][17]: the chunk's motion coefficients (coalesced load, padded rows)
    float* tab3_s = coef_s + RDG_BLOCK * 17;              // [T][K][3]: translation columns of the table (tab_smem)
    __shared__ float campos_s[SGV_MAX_VIEWS][3];
    __shared__ float bt3_s[SGV_MAX_VIEWS][RDG_NUM_BASIS_MAX][3];   // B(t_v)[k][0:3]
    const bool deform = sc.raw && sc.use_deform && sc.n_dynamic > 0;
    const int nv = p.n_views;
    if ((int)threadIdx.x < nv) {
        // campos = -R^T T of the mathematical V; glm storage: V[r][k] = vm[k*4 + r]
        const float* vm = p.viewmats + threadIdx.x * 16;
#pragma unroll
        for (int j = 0; j < 3; ++j)
            campos_s[threadIdx.x][j] = -(vm[j * 4 + 0] * vm[12 + 0] + vm[j * 4 + 1] * vm[12 + 1] + vm[j * 4 + 2] * vm[12 + 2]);
    }
    if (deform) {
        for (int e = threadIdx.x; e < nv * sc.num_basis * 3; e += RDG_BLOCK) {
            const int v = e / (sc.num_basis * 3), r = e - v * sc.num_basis * 3, k = r / 3, j = r - k * 3;
            bt3_s[v][k][j] = p.basis_ts[((int64_t)v * sc.num_basis + k) * 7 + j];
        }
        // every dynamic Gaussian needs the translation part of its birth-frame row: 48 scattered floats of a 448-byte row per
        // thread from global memory (the L1TEX-bound gather the preprocess kernels got rid of in round 1) - or 19 KB staged once
        if (p.tab_smem)
            for (int e = threadIdx.x; e < sc.num_times * sc.num_basis * 3; e += RDG_BLOCK) {
                const int row = e / 3, j = e - row * 3;
                tab3_s[e] = p.sc.table[(int64_t)row * 7 + j];
            }
    }
    __syncthreads();

    const int64_t N = sc.n_static + sc.n_dynamic;
    const int64_t cs = (sc.n_static + RDG_BLOCK - 1) / RDG_BLOCK, cd = (sc.n_dynamic + RDG_BLOCK - 1) / RDG_BLOCK;
    bool store_pending = false;
    for (int64_t chunk = blockIdx.x; chunk < cs + cd; chunk += gridDim.x) {
        const bool dyn = chunk >= cs;
        const RdgSet& set = dyn ? sc.dy : sc.st;
        const RdgSetGrad& gs = dyn ? p.dy : p.st;
        const int64_t lbase = (dyn ? chunk - cs : chunk) * RDG_BLOCK;
        const int64_t n_set = dyn ? sc.n_dynamic : sc.n_static;
        const int cnt = (int)min((int64_t)RDG_BLOCK, n_set - lbase);
        const bool valid = (int)threadIdx.x < cnt;
        const int64_t local = lbase + threadIdx.x;
        const int64_t i = (dyn ? sc.n_static : 0) + local;

        const bool def_chunk = deform && dyn;
        if (def_chunk) {
            // coalesced: the chunk's coefficient rows are contiguous in memory; row pitch 17 keeps the per-thread reads conflict-free
            __syncthreads();                               // the previous chunk is done with coef_s
            const float* src = sc.motion_coeff + lbase * sc.num_basis;
            for (int e = threadIdx.x; e < cnt * sc.num_basis; e += RDG_BLOCK) {
                const int gi = e / sc.num_basis, k = e - gi * sc.num_basis;
                coef_s[gi * 17 + k] = __ldg(src + e);
            }
            __syncthreads();
        }
        float g[3 * K];
#pragma unroll
        for (int e = 0; e < 3 * K; ++e) g[e] = 0.f;
        if (valid) {
            const float* px = set.xyz + local * 3;
            float x = __ldg(px), y = __ldg(px + 1), z = __ldg(px + 2);
            float c[RDG_NUM_BASIS_MAX];
            const bool def = deform && dyn;
            if (def) {
                // x(t_v) = x - s sum_k c_k B_k(t_i) + s sum_k c_k B_k(t_v): the birth-frame part once, the view part per view
                const int ti = __ldg(sc.time_ind + local);
                const float* pc = coef_s + threadIdx.x * 17;
                float bx = 0.f, by = 0.f, bz = 0.f;
                if (p.tab_smem) {
                    const float* row = tab3_s + ti * sc.num_basis * 3;
#pragma unroll
                    for (int k = 0; k < RDG_NUM_BASIS_MAX; ++k) {
                        c[k] = k < sc.num_basis ? pc[k] : 0.f;
                        if (k < sc.num_basis) {
                            bx = fmaf(c[k], row[k * 3 + 0], bx);
                            by = fmaf(c[k], row[k * 3 + 1], by);
                            bz = fmaf(c[k], row[k * 3 + 2], bz);
                        }
                    }
                } else {
                    const float* row = sc.table + (int64_t)ti * sc.num_basis * 7;
#pragma unroll
                    for (int k = 0; k < RDG_NUM_BASIS_MAX; ++k) {
                        c[k] = k < sc.num_basis ? pc[k] : 0.f;
                        if (k < sc.num_basis) {
                            bx = fmaf(c[k], __ldg(row + k * 7 + 0), bx);
                            by = fmaf(c[k], __ldg(row + k * 7 + 1), by);
                            bz = fmaf(c[k], __ldg(row + k * 7 + 2), bz);
                        }
                    }
                }
                x -= sc.spatial_lr_scale * bx; y -= sc.spatial_lr_scale * by; z -= sc.spatial_lr_scale * bz;
            }
            // views in groups of four: the twelve factor loads of a group are in flight together (one thread per Gaussian and
            // 16 warps per SM: with one view's loads at a time the kernel sat in long-scoreboard stalls, ncu r02)
            for (int v0 = 0; v0 < nv; v0 += 4) {
                float dcv[4][3];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const bool on = v0 + u < nv;
                    const float* d = p.dcolor + ((int64_t)(on ? v0 + u : v0) * N + i) * 3;
                    dcv[u][0] = on ? __ldg(d) : 0.f;
                    dcv[u][1] = on ? __ldg(d + 1) : 0.f;
                    dcv[u][2] = on ? __ldg(d + 2) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int v = min(v0 + u, nv - 1);
                    const float d0 = dcv[u][0], d1 = dcv[u][1], d2 = dcv[u][2];
                    if (d0 == 0.f && d1 == 0.f && d2 == 0.f) continue;     // not visible in view v (or all channels clamped)
                    float vx = x, vy = y, vz = z;
                    if (def) {
                        float bx = 0.f, by = 0.f, bz = 0.f;
#pragma unroll
                        for (int k = 0; k < RDG_NUM_BASIS_MAX; ++k) {
                            bx = fmaf(c[k], bt3_s[v][k][0], bx);
                            by = fmaf(c[k], bt3_s[v][k][1], by);
                            bz = fmaf(c[k], bt3_s[v][k][2], bz);
                        }
                        vx += sc.spatial_lr_scale * bx; vy += sc.spatial_lr_scale * by; vz += sc.spatial_lr_scale * bz;
                    }
                    float dx = vx - campos_s[v][0], dy = vy - campos_s[v][1], dz = vz - campos_s[v][2];
                    const float inv = rsqrtf(dx * dx + dy * dy + dz * dz);
                    dx *= inv; dy *= inv; dz *= inv;
                    float b[K];
                    rdg_sh_basis<DEG>(dx, dy, dz, b);
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        g[k * 3 + 0] = fmaf(b[k], d0, g[k * 3 + 0]);
                        g[k * 3 + 1] = fmaf(b[k], d1, g[k * 3 + 1]);
                        g[k * 3 + 2] = fmaf(b[k], d2, g[k * 3 + 2]);
                    }
                }
            }
            if (gs.sh_dc) {
                float* o = gs.sh_dc + local * set.sh_dc_stride;
                o[0] = g[0] * p.scale; o[1] = g[1] * p.scale; o[2] = g[2] * p.scale;
            }
        }
        if (gs.sh_rest) {
            const bool tma_out = p.use_tma && set.sh_rest_stride == SH_ROW && (cnt & 3) == 0;
            if (threadIdx.x == 0 && store_pending) { rdg_bulk_store_wait_read(); store_pending = false; }
            __syncthreads();   // the previous chunk's rows have left shared memory
            if (valid) {
                float* row = sh_s + threadIdx.x * SH_ROW;
#pragma unroll
                for (int e = 3; e < 3 * K; ++e) row[e - 3] = g[e] * p.scale;
#pragma unroll
                for (int e = 3 * K; e < 48; ++e) row[e - 3] = 0.f;
            }
            if (tma_out) {
                rdg_fence_proxy_async();
                __syncthreads();
                if (threadIdx.x == 0) {
                    rdg_bulk_store(gs.sh_rest + lbase * SH_ROW, sh_s, (uint32_t)(cnt * SH_ROW * sizeof(float)));
                    store_pending = true;
                }
            } else {
                __syncthreads();
                const int stride = set.sh_rest_stride;
                float* dst = gs.sh_rest + lbase * stride;
                for (int e = threadIdx.x; e < cnt * SH_ROW; e += RDG_BLOCK) {
                    const int gi = e / SH_ROW, k = e - gi * SH_ROW;
                    dst[(int64_t)gi * stride + k] = sh_s[e];
                }
            }
        }
    }
    if (threadIdx.x == 0 && store_pending) rdg_bulk_store_wait_read();
}

// dcolor[i] = acc[i][6..8] with the channels the forward pass clamped at zero masked out: the factors of dL/dSH,
// available as soon as the blend backward is done (so their all-gather overlaps the per-Gaussian backward)
__global__ void __launch_bounds__(RDG_BLOCK) dcolor_from_acc_kernel(int64_t n, const float* __restrict__ acc,
                                                                    const uint8_t* __restrict__ clamped, float* __restrict__ dcolor) {
    for (int64_t i = (int64_t)blockIdx.x * RDG_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * RDG_BLOCK) {
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(acc + i * 12 + 4));   // dC, dop, dr, dg
        const float gb = __ldg(acc + i * 12 + 8);
        const bool any = g1.z != 0.f || g1.w != 0.f || gb != 0.f;
        const unsigned cl = any ? clamped[i] : 0u;     // only written for visible Gaussians; the others have zero rows
        float* o = dcolor + i * 3;
        o[0] = (cl & 1u) ? 0.f : g1.z;
        o[1] = (cl & 2u) ? 0.f : g1.w;
        o[2] = (cl & 4u) ? 0.f : gb;
    }
}

// The same factors written straight into EVERY rank's gathered buffer: `dcolor_mc` is the NVLink-switch MULTICAST address
// (torch symmetric memory, handle.multicast_ptr) of this rank's block, and multimem.st makes the switch replicate each
// 16-byte store to all GPUs of the group - the all-gather of the data-parallel step fused into the kernel that produces its
// input: one pass, 24 MB of egress per rank instead of seven 24 MB peer copies queued on the copy engines (0.39 ms at 8
// GPUs, profiles/r02_timeline_n8_*.txt).  Four Gaussians per thread = three 16-byte multicast stores.
__device__ __forceinline__ void mc_st_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void mc_st_f32(float* addr, float a) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}
__global__ void __launch_bounds__(RDG_BLOCK) dcolor_multicast_kernel(int64_t n, const float* __restrict__ acc,
                                                                     const uint8_t* __restrict__ clamped, float* __restrict__ dcolor_mc) {
    const int64_t groups = (n + 3) / 4;
    for (int64_t gq = (int64_t)blockIdx.x * RDG_BLOCK + threadIdx.x; gq < groups; gq += (int64_t)gridDim.x * RDG_BLOCK) {
        float o[12];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = gq * 4 + u;
            o[3 * u] = o[3 * u + 1] = o[3 * u + 2] = 0.f;
            if (i < n) {
                const float4 g1 = __ldg(reinterpret_cast<const float4*>(acc + i * 12 + 4));   // dC, dop, dr, dg
                const float gb = __ldg(acc + i * 12 + 8);
                const bool any = g1.z != 0.f || g1.w != 0.f || gb != 0.f;
                const unsigned cl = any ? clamped[i] : 0u;
                o[3 * u] = (cl & 1u) ? 0.f : g1.z;
                o[3 * u + 1] = (cl & 2u) ? 0.f : g1.w;
                o[3 * u + 2] = (cl & 4u) ? 0.f : gb;
            }
        }
        float* dst = dcolor_mc + gq * 12;
        if (gq * 4 + 4 <= n) {
            mc_st_v4(dst, o[0], o[1], o[2], o[3]);
            mc_st_v4(dst + 4, o[4], o[5], o[6], o[7]);
            mc_st_v4(dst + 8, o[8], o[9], o[10], o[11]);
        } else {
            for (int64_t e = 0; e < (n - gq * 4) * 3; ++e) mc_st_f32(dst + e, o[e]);
        }
    }
    __threadfence_system();
}

extern "C" int rdg_dcolor_multicast(int64_t n, const float* acc, const uint8_t* clamped, float* dcolor_mc, void* stream) {
    RDG_CHECK_ARG(acc && clamped && dcolor_mc, "null argument");
    RDG_CHECK_ARG(((uintptr_t)dcolor_mc & 15u) == 0, "the multicast block must be 16-byte aligned");
    if (n <= 0) return RDG_OK;
    const int64_t want = ((n + 3) / 4 + RDG_BLOCK - 1) / RDG_BLOCK;
    const int grid = (int)(want < (int64_t)RDG_SM_COUNT * 8 ? want : (int64_t)RDG_SM_COUNT * 8);
    dcolor_multicast_kernel<<<grid, RDG_BLOCK, 0, (cudaStream_t)stream>>>(n, acc, clamped, dcolor_mc);
    RDG_CHECK_LAUNCH();
    rdg_count_launches(1);
    return RDG_OK;
}

extern "C" int rdg_dcolor_from_acc(int64_t n, const float* acc, const uint8_t* clamped, float* dcolor, void* stream) {
    RDG_CHECK_ARG(acc && clamped && dcolor, "null argument");
    if (n <= 0) return RDG_OK;
    const int64_t want = (n + RDG_BLOCK - 1) / RDG_BLOCK;
    const int grid = (int)(want < (int64_t)RDG_SM_COUNT * 8 ? want : (int64_t)RDG_SM_COUNT * 8);
    dcolor_from_acc_kernel<<<grid, RDG_BLOCK, 0, (cudaStream_t)stream>>>(n, acc, clamped, dcolor);
    RDG_CHECK_LAUNCH();
    rdg_count_launches(1);
    return RDG_OK;
}

template <int DEG>
static int launch_sgv(const ShGradParams& p, int grid, size_t smem, cudaStream_t s) {
    RDG_CUDA(cudaFuncSetAttribute(sh_grad_views_kernel<DEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sh_grad_views_kernel<DEG><<<grid, RDG_BLOCK, smem, s>>>(p);
    RDG_CHECK_LAUNCH();
    return RDG_OK;
}

extern "C" int rdg_sh_grad_views(const RdgScene* scene, int32_t sh_degree, int32_t n_views, const float* viewmatrices,
                                 const float* basis_ts, const float* dcolor, float scale, const RdgSetGrad* grad_static,
                                 const RdgSetGrad* grad_dynamic, void* stream) {
    RDG_CHECK_ARG(scene && viewmatrices && dcolor && grad_static && grad_dynamic, "null argument");
    RDG_CHECK_ARG(n_views > 0 && n_views <= SGV_MAX_VIEWS, "n_views must be 1..16");
    RDG_CHECK_ARG(sh_degree >= 0 && sh_degree <= 3, "sh_degree must be 0..3");
    const int64_t N = scene->n_static + scene->n_dynamic;
    if (N == 0) return RDG_OK;
    const bool deform = scene->raw && scene->use_deform && scene->n_dynamic > 0;
    if (deform) {
        RDG_CHECK_ARG(scene->num_basis > 0 && scene->num_basis <= RDG_NUM_BASIS_MAX, "num_basis out of range");
        RDG_CHECK_ARG(scene->motion_coeff && scene->time_ind && scene->table && basis_ts, "null deformation input");
    }
    RDG_CHECK_ARG((scene->n_static == 0 || scene->st.xyz) && (scene->n_dynamic == 0 || scene->dy.xyz), "null means");
    ShGradParams p;
    p.sc = *scene; p.st = *grad_static; p.dy = *grad_dynamic;
    p.viewmats = viewmatrices; p.basis_ts = basis_ts; p.dcolor = dcolor;
    p.n_views = n_views; p.sh_degree = sh_degree; p.scale = scale;
    p.use_tma = ((((uintptr_t)grad_static->sh_rest | (uintptr_t)grad_dynamic->sh_rest) & 15u) == 0) ? 1 : 0;
    p.tab_smem = (deform && scene->num_times <= SGV_TAB_MAX_T) ? 1 : 0;
    const size_t smem = ((size_t)RDG_BLOCK * SH_ROW + (size_t)RDG_BLOCK * 17 +
                         (p.tab_smem ? (size_t)scene->num_times * scene->num_basis * 3 : 0)) * sizeof(float);
    const int64_t chunks = (scene->n_static + RDG_BLOCK - 1) / RDG_BLOCK + (scene->n_dynamic + RDG_BLOCK - 1) / RDG_BLOCK;
    // persistent grid, two CTAs per SM.  "sm_reserve" leaves SMs to a collective that runs at the same time - these CTAs
    // take the whole register file of the SMs they sit on.
    int sms = RDG_SM_COUNT - rdg_tunable(RDG_TUN_SM_RESERVE);
    if (sms < 8) sms = 8;
    const int64_t cap = (int64_t)sms * 2;
    const int grid = (int)(chunks < cap ? chunks : cap);
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    switch (sh_degree) {
        case 0: rc = launch_sgv<0>(p, grid, smem, s); break;
        case 1: rc = launch_sgv<1>(p, grid, smem, s); break;
        case 2: rc = launch_sgv<2>(p, grid, smem, s); break;
        default: rc = launch_sgv<3>(p, grid, smem, s); break;
    }
    if (rc) return rc;
    rdg_count_launches(1);
    return RDG_OK;
}
