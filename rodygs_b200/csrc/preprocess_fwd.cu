// Fused per-Gaussian forward: static||dynamic fetch -> activations -> time-conditioned
// deformation -> EWA projection -> conic/radius/tile rectangle -> SH->RGB.
// SURVEY.md §8 rows a2-a7; spec: SURVEY.md App. A.2 == oracle/splat_oracle.py::preprocess.
//
// THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false: depth bits, pixel centre,
// radius and tile rectangle must be bit-identical to the oracle's individually
// rounded float32 operations.  The kernel is HBM-bound (>= 236 B/Gaussian in,
// 52 B out), so the lost FMA contraction costs nothing.
//
// Roofline: HBM.  Algorithmic bytes per Gaussian (degree 3): read 44 (xyz, scale,
// quat, opacity) + 192 (SH) [+ 68 dynamic: coeff + birth index], write 8 (radii,
// tiles_touched) + 41 for visible ones (p0, p1, p2, clamped).
#include "scene.cuh"

#define SH_ROW 45  // odd row stride in shared memory: conflict-free per-thread row reads

struct PreFwdParams {
    RdgScene sc;
    RdgView view;
    RdgGeom geom;
    int diff_in_smem;   // stage D[t][k][j] = B(t) - table[t] in shared memory
    int sh_in_smem;
};

template <bool RAW>
__global__ void __launch_bounds__(RDG_BLOCK) preprocess_fwd_kernel(const PreFwdParams p) {
    extern __shared__ float smem[];
    const RdgScene& sc = p.sc;
    const int64_t N = sc.n_static + sc.n_dynamic;
    const int deg = p.view.sh_degree;
    const int K = (deg + 1) * (deg + 1);
    const int nrest = 3 * (K - 1);
    float* sh_s = smem;                                             // [256][SH_ROW]
    float* diff_s = smem + (p.sh_in_smem ? RDG_BLOCK * SH_ROW : 0);  // [T][num_basis][7]

    RdgCam cam;
    rdg_load_cam(cam, p.view.viewmatrix, p.view.projmatrix, p.view.tanfovx, p.view.tanfovy,
                 p.view.width, p.view.height);
    float campos[3];
    rdg_campos(cam, campos);

    const float* diff = nullptr;
    if (RAW && sc.use_deform && sc.n_dynamic > 0 && p.diff_in_smem) {
        const int tot = sc.num_times * sc.num_basis * 7;
        const int per_t = sc.num_basis * 7;
        for (int e = threadIdx.x; e < tot; e += RDG_BLOCK) diff_s[e] = sc.basis_t[e % per_t] - sc.table[e];
        diff = diff_s;
    }
    __syncthreads();

    const int64_t n_chunks = (N + RDG_BLOCK - 1) / RDG_BLOCK;
    for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int64_t base = chunk * RDG_BLOCK;
        const int cnt = (int)min((int64_t)RDG_BLOCK, N - base);
        const bool use_sh = (p.sc.colors_precomp == nullptr) && nrest > 0;

        // ---- stage this chunk's higher-order SH rows in shared memory (coalesced) ----
        if (use_sh && p.sh_in_smem) {
            __syncthreads();  // previous chunk's readers are done
            const bool one_set = (base >= sc.n_static) || (base + cnt <= sc.n_static);
            const RdgSet& set0 = (base >= sc.n_static) ? sc.dy : sc.st;
            if (one_set && set0.sh_rest_stride == nrest && nrest == SH_ROW) {
                const int64_t lbase = (base >= sc.n_static) ? base - sc.n_static : base;
                const float* src = set0.sh_rest + lbase * SH_ROW;
                const int tot = cnt * SH_ROW;
                for (int e = threadIdx.x; e < tot; e += RDG_BLOCK) sh_s[e] = src[e];
            } else if (one_set && set0.sh_rest_stride == 48) {
                // cat'ed [n,16,3] rows (the drop-in boundary): contiguous 192-byte rows, dc first
                const int64_t lbase = (base >= sc.n_static) ? base - sc.n_static : base;
                const float* src = set0.sh_rest + lbase * 48;
                const int tot = cnt * 48 - 3;  // the pointer already skips the first row's dc
                for (int e = threadIdx.x; e < tot; e += RDG_BLOCK) {
                    const int g = e / 48, k = e - g * 48;
                    if (k < nrest) sh_s[g * SH_ROW + k] = src[e];
                }
            } else {
                const int tot = cnt * nrest;
                for (int e = threadIdx.x; e < tot; e += RDG_BLOCK) {
                    const int g = e / nrest, k = e - g * nrest;
                    const int64_t gi = base + g;
                    const bool dy = gi >= sc.n_static;
                    const RdgSet& set = dy ? sc.dy : sc.st;
                    const int64_t l = dy ? gi - sc.n_static : gi;
                    sh_s[g * SH_ROW + k] = set.sh_rest[l * set.sh_rest_stride + k];
                }
            }
            __syncthreads();
        }

        const int64_t i = base + threadIdx.x;
        if (i >= N) continue;

        RdgAct a;
        rdg_fetch<RAW>(sc, i, diff, a);
        if (p.geom.dbg_activated) {
            float* d = p.geom.dbg_activated + i * 11;
            d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.s[0]; d[4] = a.s[1]; d[5] = a.s[2];
            d[6] = a.q[0]; d[7] = a.q[1]; d[8] = a.q[2]; d[9] = a.q[3]; d[10] = a.op;
        }

        int radius = 0;
        unsigned tiles = 0;
        // near-plane test needs only t.z, but the projection is cheap next to the loads
        RdgProj pr;
        rdg_project(cam, a, p.view.scale_modifier, pr);
        if (pr.tz > RDG_NEAR_Z && pr.det != 0.0f) {
            const float det_inv = 1.0f / pr.det;
            const float conA = pr.cc * det_inv, conB = -pr.cb * det_inv, conC = pr.ca * det_inv;
            const float mid = 0.5f * (pr.ca + pr.cc);
            const float disc = sqrtf(fmaxf(0.1f, mid * mid - pr.det));
            const float lam1 = mid + disc, lam2 = mid - disc;
            const float rad_f = ceilf(3.0f * sqrtf(fmaxf(lam1, lam2)));
            const float ndcx = pr.hx * pr.pw, ndcy = pr.hy * pr.pw;
            const float px = ((ndcx + 1.0f) * cam.W - 1.0f) * 0.5f;
            const float py = ((ndcy + 1.0f) * cam.H - 1.0f) * 0.5f;
            const int rminx = min(cam.gx, max(0, (int)((px - rad_f) / 16.0f)));
            const int rminy = min(cam.gy, max(0, (int)((py - rad_f) / 16.0f)));
            const int rmaxx = min(cam.gx, max(0, (int)((px + rad_f + 15.0f) / 16.0f)));
            const int rmaxy = min(cam.gy, max(0, (int)((py + rad_f + 15.0f) / 16.0f)));
            const int cntt = (rmaxx - rminx) * (rmaxy - rminy);
            if (cntt > 0) {
                radius = (int)rad_f;
                tiles = (unsigned)cntt;
                float rgb[3];
                unsigned clamped = 0;
                if (sc.colors_precomp) {
                    rgb[0] = sc.colors_precomp[i * 3 + 0];
                    rgb[1] = sc.colors_precomp[i * 3 + 1];
                    rgb[2] = sc.colors_precomp[i * 3 + 2];
                } else {
                    const RdgSet& set = a.dyn ? sc.dy : sc.st;
                    const float* dc = set.sh_dc + a.local * set.sh_dc_stride;
                    float dx = a.x - campos[0], dy = a.y - campos[1], dz = a.z - campos[2];
                    const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
                    dx *= inv; dy *= inv; dz *= inv;
                    float b[16];
                    rdg_sh_basis(deg, dx, dy, dz, b);
                    rgb[0] = b[0] * dc[0]; rgb[1] = b[0] * dc[1]; rgb[2] = b[0] * dc[2];
                    if (nrest > 0) {
                        const float* rest = p.sh_in_smem ? (sh_s + threadIdx.x * SH_ROW)
                                                         : (set.sh_rest + a.local * set.sh_rest_stride);
                        for (int k = 1; k < K; ++k) {
                            rgb[0] += b[k] * rest[(k - 1) * 3 + 0];
                            rgb[1] += b[k] * rest[(k - 1) * 3 + 1];
                            rgb[2] += b[k] * rest[(k - 1) * 3 + 2];
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        rgb[c] += 0.5f;
                        if (rgb[c] < 0.0f) { clamped |= 1u << c; rgb[c] = 0.0f; }
                    }
                }
                reinterpret_cast<float4*>(p.geom.p0)[i] = make_float4(px, py, conA, conB);
                reinterpret_cast<float4*>(p.geom.p1)[i] = make_float4(conC, a.op, rgb[0], rgb[1]);
                reinterpret_cast<float2*>(p.geom.p2)[i] = make_float2(rgb[2], pr.tz);
                p.geom.clamped[i] = (uint8_t)clamped;
            }
        }
        p.geom.radii[i] = radius;
        p.geom.tiles_touched[i] = tiles;
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
    const bool need_sh = scene->colors_precomp == nullptr && view->sh_degree > 0;
    p.sh_in_smem = need_sh ? 1 : 0;
    size_t smem = p.sh_in_smem ? RDG_BLOCK * SH_ROW * sizeof(float) : 0;
    const size_t diff_bytes = deform ? (size_t)scene->num_times * scene->num_basis * 7 * sizeof(float) : 0;
    p.diff_in_smem = (deform && smem + diff_bytes <= 160 * 1024) ? 1 : 0;
    if (p.diff_in_smem) smem += diff_bytes;
    const int64_t chunks = (N + RDG_BLOCK - 1) / RDG_BLOCK;
    // persistent-ish grid: a multiple of the SM count so the staged table is amortised
    const int grid = (int)(chunks < (int64_t)RDG_SM_COUNT * 4 ? chunks : (int64_t)RDG_SM_COUNT * 4);
    cudaStream_t s = (cudaStream_t)stream;
    if (scene->raw) {
        RDG_CUDA(cudaFuncSetAttribute(preprocess_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        preprocess_fwd_kernel<true><<<grid, RDG_BLOCK, smem, s>>>(p);
    } else {
        RDG_CUDA(cudaFuncSetAttribute(preprocess_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        preprocess_fwd_kernel<false><<<grid, RDG_BLOCK, smem, s>>>(p);
    }
    RDG_CHECK_LAUNCH();
    rdg_count_launches(1);
    return RDG_OK;
}
