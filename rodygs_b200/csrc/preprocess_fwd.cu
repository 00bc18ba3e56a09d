// Fused per-Gaussian forward: static||dynamic fetch -> activations -> time-conditioned
// deformation -> EWA projection -> conic/radius/tile rectangle -> SH->RGB.
// SURVEY.md §8 rows a2-a7; spec: SURVEY.md App. A.2 == oracle/splat_oracle.py::preprocess.
//
// THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false: depth bits, pixel centre,
// radius and tile rectangle must be bit-identical to the oracle's individually
// rounded float32 operations.  The kernel is HBM-bound, so the lost FMA contraction
// costs nothing.
//
// Roofline: HBM.  Algorithmic bytes per Gaussian (degree 3): read 44 (xyz, scale,
// quat, opacity) + 192 (SH) [+ 68 dynamic: coeff + birth index], write 8 (radii,
// tiles_touched) + 41 for visible ones (p0, p1, p2, clamped).
//
// B200 mapping: one thread per Gaussian, 256-Gaussian chunks that never straddle the
// static/dynamic boundary, so a chunk's higher-order SH rows are one contiguous,
// 16-byte aligned 46 KB span: it is fetched by ONE cp.async.bulk (TMA) into shared memory
// while the threads load their 44 B of geometry and run the projection; each thread then
// reads its own row (stride 45 words: conflict-free).  46 KB per CTA -> 4 CTAs per SM.
#include "scene.cuh"
#include "tma.cuh"

#define SH_ROW 45  // odd row stride in shared memory: conflict-free per-thread row reads

struct PreFwdParams {
    RdgScene sc;
    RdgView view;
    RdgGeom geom;
    int use_tma;
    int diff_smem;   // B(t) - table rows staged in shared memory (rdg_stage_diff); the launcher sized the buffer
    int l2_prefetch; // next chunk's parameter rows prefetched into L2 (cp.async.bulk.prefetch.L2)
};

// Measured alternatives that did NOT help (B200, C4, profiles/r01_ab_v12_*.json, r01_ab_v13_*.json): double-buffering the SH
// rows (next chunk's bulk load issued one chunk ahead; 0.266 vs 0.258 ms - the bulk copy is not what the warps wait for)
// and forcing 3 CTAs per SM with __launch_bounds__(256, 3) (80 registers + spills; 0.302 vs 0.259 ms).
template <bool RAW, int DEG>
__global__ void __launch_bounds__(RDG_BLOCK) preprocess_fwd_kernel(const PreFwdParams p) {
    extern __shared__ __align__(128) float smem[];
    constexpr int K = (DEG + 1) * (DEG + 1);
    constexpr int NREST = 3 * (K - 1);
    const RdgScene& sc = p.sc;
    float* sh_s = smem;                       // [256][SH_ROW]
    float* bt_s = smem + RDG_BLOCK * SH_ROW;  // [16*7] B(t)
    __shared__ __align__(8) uint64_t bar;

    RdgCam cam;
    rdg_load_cam(cam, p.view.viewmatrix, p.view.projmatrix, p.view.tanfovx, p.view.tanfovy,
                 p.view.width, p.view.height);
    float campos[3];
    rdg_campos(cam, campos);

    const bool deform = RAW && sc.use_deform && sc.n_dynamic > 0;
    float* diff_s = (deform && p.diff_smem) ? bt_s + RDG_NUM_BASIS_MAX * 7 : nullptr;
    if (deform) {
        for (int e = threadIdx.x; e < sc.num_basis * 7; e += RDG_BLOCK) bt_s[e] = sc.basis_t[e];
        if (diff_s) rdg_stage_diff(sc, sc.basis_t, diff_s, RDG_BLOCK);
    }
    if (threadIdx.x == 0) rdg_mbar_init(&bar, 1);
    __syncthreads();

    const bool use_sh = (sc.colors_precomp == nullptr) && NREST > 0;
    const int64_t cs = (sc.n_static + RDG_BLOCK - 1) / RDG_BLOCK, cd = (sc.n_dynamic + RDG_BLOCK - 1) / RDG_BLOCK;
    uint32_t phase = 0;
    for (int64_t chunk = blockIdx.x; chunk < cs + cd; chunk += gridDim.x) {
        const bool dyn = chunk >= cs;
        const RdgSet& set = dyn ? sc.dy : sc.st;
        const int64_t lbase = (dyn ? chunk - cs : chunk) * RDG_BLOCK;
        const int64_t n_set = dyn ? sc.n_dynamic : sc.n_static;
        const int cnt = (int)min((int64_t)RDG_BLOCK, n_set - lbase);
        if (p.l2_prefetch && chunk + gridDim.x < cs + cd && threadIdx.x < 8)
            rdg_prefetch_chunk_field(sc, chunk + gridDim.x, cs, threadIdx.x, nullptr);

        // ---- stage this chunk's higher-order SH rows in shared memory ----
        bool tma = false;
        if (use_sh) {
            __syncthreads();  // previous chunk's readers are done with sh_s
            tma = p.use_tma && NREST == SH_ROW && set.sh_rest_stride == SH_ROW && (cnt & 3) == 0;
            if (tma) {
                if (threadIdx.x == 0) {
                    rdg_fence_proxy_async();
                    rdg_bulk_load(sh_s, set.sh_rest + lbase * SH_ROW, (uint32_t)(cnt * SH_ROW * sizeof(float)), &bar);
                }
            } else if (set.sh_rest_stride == NREST) {
                const float* src = set.sh_rest + lbase * NREST;
                for (int e = threadIdx.x; e < cnt * NREST; e += RDG_BLOCK) sh_s[(e / NREST) * SH_ROW + (e % NREST)] = src[e];
            } else {
                const int stride = set.sh_rest_stride;   // e.g. 48: cat'ed [n,16,3] rows of the drop-in boundary
                const float* src = set.sh_rest + lbase * stride;
                for (int e = threadIdx.x; e < cnt * NREST; e += RDG_BLOCK) {
                    const int g = e / NREST, k = e - g * NREST;
                    sh_s[g * SH_ROW + k] = src[(int64_t)g * stride + k];
                }
            }
        }

        const bool valid = (int)threadIdx.x < cnt;
        const int64_t local = lbase + threadIdx.x;
        const int64_t i = (dyn ? sc.n_static : 0) + local;
        int radius = 0;
        unsigned tiles = 0;
        bool vis = false;
        RdgAct a;
        float px = 0.f, py = 0.f, conA = 0.f, conB = 0.f, conC = 0.f, tz = 0.f;
        // degree-0 SH coefficients: requested here with the geometry, not after the SH rows have arrived - otherwise they
        // are one more exposed DRAM round trip at the end of every chunk (the kernel is latency bound)
        float dc0 = 0.f, dc1 = 0.f, dc2 = 0.f;
        if (valid && !sc.colors_precomp) {
            const float* dc = set.sh_dc + local * set.sh_dc_stride;
            dc0 = __ldg(dc); dc1 = __ldg(dc + 1); dc2 = __ldg(dc + 2);
        }
        if (valid) {
            rdg_fetch<RAW>(sc, dyn, local, bt_s, a, diff_s);
            if (p.geom.dbg_activated) {
                float* d = p.geom.dbg_activated + i * 11;
                d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.s[0]; d[4] = a.s[1]; d[5] = a.s[2];
                d[6] = a.q[0]; d[7] = a.q[1]; d[8] = a.q[2]; d[9] = a.q[3]; d[10] = a.op;
            }
            RdgProj pr;
            rdg_project(cam, a, p.view.scale_modifier, pr);
            if (pr.tz > RDG_NEAR_Z && pr.det != 0.0f) {
                const float det_inv = 1.0f / pr.det;
                conA = pr.cc * det_inv; conB = -pr.cb * det_inv; conC = pr.ca * det_inv;
                const float mid = 0.5f * (pr.ca + pr.cc);
                const float disc = sqrtf(fmaxf(0.1f, mid * mid - pr.det));
                const float lam1 = mid + disc, lam2 = mid - disc;
                const float rad_f = ceilf(3.0f * sqrtf(fmaxf(lam1, lam2)));
                const float ndcx = pr.hx * pr.pw, ndcy = pr.hy * pr.pw;
                px = ((ndcx + 1.0f) * cam.W - 1.0f) * 0.5f;
                py = ((ndcy + 1.0f) * cam.H - 1.0f) * 0.5f;
                const int rminx = min(cam.gx, max(0, (int)((px - rad_f) / 16.0f)));
                const int rminy = min(cam.gy, max(0, (int)((py - rad_f) / 16.0f)));
                const int rmaxx = min(cam.gx, max(0, (int)((px + rad_f + 15.0f) / 16.0f)));
                const int rmaxy = min(cam.gy, max(0, (int)((py + rad_f + 15.0f) / 16.0f)));
                const int cntt = (rmaxx - rminx) * (rmaxy - rminy);
                if (cntt > 0) {
                    radius = (int)rad_f; tiles = (unsigned)cntt; vis = true; tz = pr.tz;
                    if (p.geom.tile_count) {   // per-tile population for the tile-bucketed binning
                        for (int yy = rminy; yy < rmaxy; ++yy)
                            for (int xx = rminx; xx < rmaxx; ++xx) atomicAdd(&p.geom.tile_count[yy * cam.gx + xx], 1u);
                    }
                }
            }
        }
        if (use_sh) {
            if (tma) rdg_mbar_wait(&bar, phase & 1u);
            else __syncthreads();
        }
        if (tma) ++phase;
        if (vis) {
            float rgb[3];
            unsigned clamped = 0;
            if (sc.colors_precomp) {
                rgb[0] = sc.colors_precomp[i * 3 + 0];
                rgb[1] = sc.colors_precomp[i * 3 + 1];
                rgb[2] = sc.colors_precomp[i * 3 + 2];
            } else {
                float dx = a.x - campos[0], dy = a.y - campos[1], dz = a.z - campos[2];
                const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
                dx *= inv; dy *= inv; dz *= inv;
                float b[K];
                rdg_sh_basis<DEG>(dx, dy, dz, b);
                rgb[0] = b[0] * dc0; rgb[1] = b[0] * dc1; rgb[2] = b[0] * dc2;
                const float* rest = sh_s + threadIdx.x * SH_ROW;
#pragma unroll
                for (int k = 1; k < K; ++k) {   // explicit FMAs (-fmad=false TU): the colour feeds no integer output
                    rgb[0] = fmaf(b[k], rest[(k - 1) * 3 + 0], rgb[0]);
                    rgb[1] = fmaf(b[k], rest[(k - 1) * 3 + 1], rgb[1]);
                    rgb[2] = fmaf(b[k], rest[(k - 1) * 3 + 2], rgb[2]);
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    rgb[c] += 0.5f;
                    if (rgb[c] < 0.0f) { clamped |= 1u << c; rgb[c] = 0.0f; }
                }
            }
            reinterpret_cast<float4*>(p.geom.p0)[i] = make_float4(px, py, conA, conB);
            reinterpret_cast<float4*>(p.geom.p1)[i] = make_float4(conC, a.op, rgb[0], rgb[1]);
            reinterpret_cast<float2*>(p.geom.p2)[i] = make_float2(rgb[2], tz);
            p.geom.clamped[i] = (uint8_t)clamped;
        }
        if (valid) {
            p.geom.radii[i] = radius;
            p.geom.tiles_touched[i] = tiles;
        }
    }
}

template <bool RAW, int DEG>
static int launch_fwd(const PreFwdParams& p, int grid, size_t smem, cudaStream_t s) {
    const int cap = rdg_tunable(RDG_TUN_PRE_GRID_CAP);
    if (cap > 0 && grid > cap) grid = cap;
    RDG_CUDA(cudaFuncSetAttribute(preprocess_fwd_kernel<RAW, DEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    preprocess_fwd_kernel<RAW, DEG><<<grid, RDG_BLOCK, smem, s>>>(p);
    RDG_CHECK_LAUNCH();
    return RDG_OK;
}

template <bool RAW>
static int launch_fwd_deg(const PreFwdParams& p, int deg, int grid, size_t smem, cudaStream_t s) {
    switch (deg) {
        case 0: return launch_fwd<RAW, 0>(p, grid, smem, s);
        case 1: return launch_fwd<RAW, 1>(p, grid, smem, s);
        case 2: return launch_fwd<RAW, 2>(p, grid, smem, s);
        default: return launch_fwd<RAW, 3>(p, grid, smem, s);
    }
}

extern "C" int rdg_preprocess_fwd(const RdgScene* scene, const RdgView* view, const RdgGeom* geom, void* stream) {
    RDG_CHECK_ARG(scene && view && geom, "null argument");
    const int64_t N = scene->n_static + scene->n_dynamic;
    RDG_CHECK_ARG(N >= 0 && scene->n_static >= 0 && scene->n_dynamic >= 0, "negative count");
    RDG_CHECK_ARG(view->sh_degree >= 0 && view->sh_degree <= 3, "sh_degree must be 0..3");
    RDG_CHECK_ARG(view->width > 0 && view->height > 0, "empty image");
    RDG_CHECK_ARG(view->viewmatrix && view->projmatrix, "null camera matrix");
    if (N == 0) return RDG_OK;
    RDG_CHECK_ARG(geom->radii && geom->tiles_touched && geom->p0 && geom->p1 && geom->p2 && geom->clamped,
                  "null geometry buffer");
    RDG_CHECK_ARG(scene->n_static == 0 || (scene->st.xyz && scene->st.scaling && scene->st.rotation && scene->st.opacity),
                  "null static parameter");
    RDG_CHECK_ARG(scene->n_dynamic == 0 || (scene->dy.xyz && scene->dy.scaling && scene->dy.rotation && scene->dy.opacity),
                  "null dynamic parameter");
    RDG_CHECK_ARG(scene->colors_precomp || ((scene->n_static == 0 || scene->st.sh_dc) && (scene->n_dynamic == 0 || scene->dy.sh_dc)),
                  "exactly one of SHs / precomputed colours is required");
    RDG_CHECK_ARG(scene->colors_precomp || view->sh_degree == 0 ||
                      ((scene->n_static == 0 || scene->st.sh_rest) && (scene->n_dynamic == 0 || scene->dy.sh_rest)),
                  "null higher-order SH pointer");
    const bool deform = scene->raw && scene->use_deform && scene->n_dynamic > 0;
    if (deform) {
        RDG_CHECK_ARG(scene->num_basis > 0 && scene->num_basis <= RDG_NUM_BASIS_MAX, "num_basis out of range");
        RDG_CHECK_ARG(scene->motion_coeff && scene->time_ind && scene->basis_t && scene->table && scene->num_times > 0,
                      "null deformation input");
    }
    PreFwdParams p;
    p.sc = *scene;
    p.view = *view;
    p.geom = *geom;
    // TMA needs 16-byte aligned global sources; torch allocations are, arbitrary views may not be
    const bool aligned = (((uintptr_t)scene->st.sh_rest | (uintptr_t)scene->dy.sh_rest) & 15u) == 0;
    p.use_tma = aligned ? 1 : 0;
    p.l2_prefetch = rdg_tunable(RDG_TUN_L2_PREFETCH) != 0 ? 1 : 0;
    size_t smem = (RDG_BLOCK * SH_ROW + RDG_NUM_BASIS_MAX * 7) * sizeof(float);
    // B(t) - table rows in shared memory when the whole table fits next to the SH rows with 2 CTAs per SM (T <= 140)
    p.diff_smem = 0;
    if (deform && scene->num_basis == RDG_NUM_BASIS_MAX && rdg_tunable(RDG_TUN_DIFF_SMEM) != 0 &&
        (((uintptr_t)scene->basis_t | (uintptr_t)scene->table) & 15u) == 0) {
        const size_t extra = (size_t)scene->num_times * RDG_DIFF_STRIDE * sizeof(float);
        if (smem + extra <= RDG_PRE_SMEM_MAX) { p.diff_smem = 1; smem += extra; }
    }
    const int64_t chunks = (scene->n_static + RDG_BLOCK - 1) / RDG_BLOCK + (scene->n_dynamic + RDG_BLOCK - 1) / RDG_BLOCK;
    // persistent grid: a multiple of the SM count.  2 CTAs are resident per SM (registers); with the staged difference
    // table every CTA pays for filling it once, so the grid is exactly the resident set
    const int64_t per_sm = p.diff_smem ? 2 : 4;
    const int grid = (int)(chunks < (int64_t)RDG_SM_COUNT * per_sm ? chunks : (int64_t)RDG_SM_COUNT * per_sm);
    cudaStream_t s = (cudaStream_t)stream;
    if (geom->tile_count) {
        const int tiles = ((view->width + RDG_TILE - 1) / RDG_TILE) * ((view->height + RDG_TILE - 1) / RDG_TILE);
        RDG_CUDA(cudaMemsetAsync(geom->tile_count, 0, (size_t)(tiles + 1) * sizeof(uint32_t), s));
    }
    const int rc = scene->raw ? launch_fwd_deg<true>(p, view->sh_degree, grid, smem, s)
                              : launch_fwd_deg<false>(p, view->sh_degree, grid, smem, s);
    if (rc) return rc;
    rdg_count_launches(1);
    return RDG_OK;
}
