// Per-Gaussian backward: screen-space gradients (acc, from blend backward) ->
// gradients of means / scales / quaternions / opacity / SH (+ through the fused
// activations and the time-conditioned deformation when scene.raw) and of the pose.
// SURVEY.md §8 row a10, App. A.6-A.7; checked against autograd of the oracle.
//
// HBM-bound: reads 48 B acc + the forward inputs again, writes 44 B + 12*K B (dSH)
// + 12 B (means2D sink) [+ 64 B dcoeff] per Gaussian.  dSH rows are staged through
// shared memory so that both the SH read and the dSH write are fully coalesced.
// dL/dtable[T,K,7] is accumulated in shared memory per (persistent) CTA and flushed
// once; dL/dB(t) = -sum_t dL/dtable[t] falls out of the same flush.
#include "scene.cuh"

#define SH_ROW 45
#define NACC 12

struct PreBwdParams {
    RdgScene sc;
    RdgView view;
    RdgGeom geom;
    RdgSceneGrad gr;
    const float* acc;
    int diff_in_smem;
    int dtab_in_smem;
};

__device__ __forceinline__ void sh_basis_grad(int deg, float x, float y, float z, float* bx, float* by, float* bz) {
    bx[0] = by[0] = bz[0] = 0.f;
    if (deg > 0) {
        bx[1] = 0.f;         by[1] = -RDG_SH_C1; bz[1] = 0.f;
        bx[2] = 0.f;         by[2] = 0.f;        bz[2] = RDG_SH_C1;
        bx[3] = -RDG_SH_C1;  by[3] = 0.f;        bz[3] = 0.f;
        if (deg > 1) {
            bx[4] = RDG_SH_C2_0 * y;          by[4] = RDG_SH_C2_0 * x;          bz[4] = 0.f;
            bx[5] = 0.f;                      by[5] = RDG_SH_C2_1 * z;          bz[5] = RDG_SH_C2_1 * y;
            bx[6] = RDG_SH_C2_2 * -2.f * x;   by[6] = RDG_SH_C2_2 * -2.f * y;   bz[6] = RDG_SH_C2_2 * 4.f * z;
            bx[7] = RDG_SH_C2_3 * z;          by[7] = 0.f;                      bz[7] = RDG_SH_C2_3 * x;
            bx[8] = RDG_SH_C2_4 * 2.f * x;    by[8] = RDG_SH_C2_4 * -2.f * y;   bz[8] = 0.f;
            if (deg > 2) {
                const float xx = x * x, yy = y * y, zz = z * z;
                bx[9] = RDG_SH_C3_0 * 6.f * x * y;       by[9] = RDG_SH_C3_0 * (3.f * xx - 3.f * yy);      bz[9] = 0.f;
                bx[10] = RDG_SH_C3_1 * y * z;            by[10] = RDG_SH_C3_1 * x * z;                     bz[10] = RDG_SH_C3_1 * x * y;
                bx[11] = RDG_SH_C3_2 * -2.f * x * y;     by[11] = RDG_SH_C3_2 * (4.f * zz - xx - 3.f * yy); bz[11] = RDG_SH_C3_2 * 8.f * y * z;
                bx[12] = RDG_SH_C3_3 * -6.f * x * z;     by[12] = RDG_SH_C3_3 * -6.f * y * z;              bz[12] = RDG_SH_C3_3 * (6.f * zz - 3.f * xx - 3.f * yy);
                bx[13] = RDG_SH_C3_4 * (4.f * zz - 3.f * xx - yy); by[13] = RDG_SH_C3_4 * -2.f * x * y;    bz[13] = RDG_SH_C3_4 * 8.f * x * z;
                bx[14] = RDG_SH_C3_5 * 2.f * x * z;      by[14] = RDG_SH_C3_5 * -2.f * y * z;              bz[14] = RDG_SH_C3_5 * (xx - yy);
                bx[15] = RDG_SH_C3_6 * (3.f * xx - 3.f * yy); by[15] = RDG_SH_C3_6 * -6.f * x * y;         bz[15] = 0.f;
            }
        }
    }
}

template <bool RAW>
__global__ void __launch_bounds__(RDG_BLOCK) preprocess_bwd_kernel(const PreBwdParams p) {
    extern __shared__ float smem[];
    const RdgScene& sc = p.sc;
    const int64_t N = sc.n_static + sc.n_dynamic;
    const int deg = p.view.sh_degree;
    const int K = (deg + 1) * (deg + 1);
    const bool use_sh = sc.colors_precomp == nullptr;
    const bool deform = RAW && sc.use_deform && sc.n_dynamic > 0;
    const int per_t = sc.num_basis * 7;
    const int tab_n = deform ? sc.num_times * per_t : 0;
    float* sh_s = smem;                                    // [256][SH_ROW] in: SH rest coefficients, out: their gradients
    float* diff_s = sh_s + RDG_BLOCK * SH_ROW;             // [T][K][7]
    float* dtab_s = diff_s + (p.diff_in_smem ? tab_n : 0);  // [T][K][7]
    __shared__ float red_s[RDG_BLOCK / 32][16];

    RdgCam cam;
    rdg_load_cam(cam, p.view.viewmatrix, p.view.projmatrix, p.view.tanfovx, p.view.tanfovy, p.view.width, p.view.height);
    float campos[3];
    rdg_campos(cam, campos);
    const float* V = cam.V;
    const float* P = cam.P;

    const float* diff = nullptr;
    if (deform) {
        if (p.diff_in_smem) {
            for (int e = threadIdx.x; e < tab_n; e += RDG_BLOCK) diff_s[e] = sc.basis_t[e % per_t] - sc.table[e];
            diff = diff_s;
        }
        if (p.dtab_in_smem)
            for (int e = threadIdx.x; e < tab_n; e += RDG_BLOCK) dtab_s[e] = 0.f;
    }
    __syncthreads();

    float poseV[12];   // dL/dV rows 0..2 (row-major), this thread's partial sum
    float dcam[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 12; ++k) poseV[k] = 0.f;

    const int lane = threadIdx.x & 31;
    const int64_t n_chunks = (N + RDG_BLOCK - 1) / RDG_BLOCK;
    for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int64_t base = chunk * RDG_BLOCK;
        const int cnt = (int)min((int64_t)RDG_BLOCK, N - base);
        const int64_t i = base + threadIdx.x;
        const bool valid = i < N;
        const int radius = valid ? p.geom.radii[i] : 0;
        const bool vis = radius > 0;

        // ---- stage SH rest rows (visible Gaussians only need them, but the copy is coalesced) ----
        __syncthreads();
        if (use_sh && K > 1) {
            const int nrest = 3 * (K - 1);
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

        RdgAct a;
        a.dyn = false; a.local = 0; a.ti = 0;
        float dmean[3] = {0.f, 0.f, 0.f}, dscale[3] = {0.f, 0.f, 0.f}, dquat[4] = {0.f, 0.f, 0.f, 0.f};
        float dop = 0.f, ddc[3] = {0.f, 0.f, 0.f}, dm2[2] = {0.f, 0.f};
        float* my_sh = sh_s + threadIdx.x * SH_ROW;
        if (valid) {
            a.dyn = i >= sc.n_static;
            a.local = a.dyn ? i - sc.n_static : i;
        }
        if (vis) {
            rdg_fetch<RAW>(sc, i, diff, a);
            RdgProj pr;
            rdg_project(cam, a, p.view.scale_modifier, pr);
            const float4 g0 = reinterpret_cast<const float4*>(p.acc)[i * 3 + 0];
            const float4 g1 = reinterpret_cast<const float4*>(p.acc)[i * 3 + 1];
            const float4 g2 = reinterpret_cast<const float4*>(p.acc)[i * 3 + 2];
            const float gA = g0.z, gB = g0.w, gC = g1.x;
            float grgb[3] = {g1.z, g1.w, g2.x};
            const float gdepth = g2.y;

            // ---- projection path ----
            const float dndcx = g0.x * 0.5f * cam.W, dndcy = g0.y * 0.5f * cam.H;
            dm2[0] = dndcx; dm2[1] = dndcy;
            const float dhx = dndcx * pr.pw, dhy = dndcy * pr.pw;
            const float dhw = -(dndcx * pr.hx + dndcy * pr.hy) * pr.pw * pr.pw;
            float dt_pose[3], dt_cov[3];
            dt_pose[0] = P[0] * dhx + P[4] * dhy + P[12] * dhw;
            dt_pose[1] = P[1] * dhx + P[5] * dhy + P[13] * dhw;
            dt_pose[2] = P[2] * dhx + P[6] * dhy + P[14] * dhw + gdepth;

            // ---- conic -> cov2D ----
            const float ca = pr.ca, cb = pr.cb, cc = pr.cc, det = pr.det;
            const float d2 = 1.0f / (det * det + 1e-7f);
            const float da = d2 * (-cc * cc * gA + cb * cc * gB - cb * cb * gC);
            const float db = d2 * (2.f * cb * cc * gA - (det + 2.f * cb * cb) * gB + 2.f * ca * cb * gC);
            const float dc = d2 * (-cb * cb * gA + ca * cb * gB - ca * ca * gC);
            const float* S = pr.S;
            const float* T = pr.T;
            const float v0[3] = {S[0] * T[0] + S[1] * T[1] + S[2] * T[2], S[1] * T[0] + S[3] * T[1] + S[4] * T[2],
                                 S[2] * T[0] + S[4] * T[1] + S[5] * T[2]};
            const float v1[3] = {S[0] * T[3] + S[1] * T[4] + S[2] * T[5], S[1] * T[3] + S[3] * T[4] + S[4] * T[5],
                                 S[2] * T[3] + S[4] * T[4] + S[5] * T[5]};
            float dT0[3], dT1[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                dT0[j] = 2.f * da * v0[j] + db * v1[j];
                dT1[j] = 2.f * dc * v1[j] + db * v0[j];
            }
            // F = symmetric "full" gradient of Sigma
            const float F00 = da * T[0] * T[0] + db * T[0] * T[3] + dc * T[3] * T[3];
            const float F11 = da * T[1] * T[1] + db * T[1] * T[4] + dc * T[4] * T[4];
            const float F22 = da * T[2] * T[2] + db * T[2] * T[5] + dc * T[5] * T[5];
            const float F01 = da * T[0] * T[1] + 0.5f * db * (T[0] * T[4] + T[1] * T[3]) + dc * T[3] * T[4];
            const float F02 = da * T[0] * T[2] + 0.5f * db * (T[0] * T[5] + T[2] * T[3]) + dc * T[3] * T[5];
            const float F12 = da * T[1] * T[2] + 0.5f * db * (T[1] * T[5] + T[2] * T[4]) + dc * T[4] * T[5];
            const float* M = pr.M;
            float dM[9];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                dM[0 + k] = 2.f * (F00 * M[0 + k] + F01 * M[3 + k] + F02 * M[6 + k]);
                dM[3 + k] = 2.f * (F01 * M[0 + k] + F11 * M[3 + k] + F12 * M[6 + k]);
                dM[6 + k] = 2.f * (F02 * M[0 + k] + F12 * M[3 + k] + F22 * M[6 + k]);
            }
            const float* R = pr.R;
            float dR[9];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float sk = a.s[k] * p.view.scale_modifier;
                dscale[k] = (dM[0 + k] * R[0 + k] + dM[3 + k] * R[3 + k] + dM[6 + k] * R[6 + k]) * p.view.scale_modifier;
                dR[0 + k] = dM[0 + k] * sk; dR[3 + k] = dM[3 + k] * sk; dR[6 + k] = dM[6 + k] * sk;
            }
            {
                const float qr = a.q[0], qx = a.q[1], qy = a.q[2], qz = a.q[3];
                dquat[0] = 2.f * (-qz * dR[1] + qy * dR[2] + qz * dR[3] - qx * dR[5] - qy * dR[6] + qx * dR[7]);
                dquat[1] = 2.f * (qy * dR[1] + qz * dR[2] + qy * dR[3] - 2.f * qx * dR[4] - qr * dR[5] + qz * dR[6] + qr * dR[7] - 2.f * qx * dR[8]);
                dquat[2] = 2.f * (-2.f * qy * dR[0] + qx * dR[1] + qr * dR[2] + qx * dR[3] + qz * dR[5] - qr * dR[6] + qz * dR[7] - 2.f * qy * dR[8]);
                dquat[3] = 2.f * (-2.f * qz * dR[0] - qr * dR[1] + qx * dR[2] + qr * dR[3] - 2.f * qz * dR[4] + qy * dR[5] + qx * dR[6] + qy * dR[7]);
            }
            // ---- J, t ----
            const float dJ00 = dT0[0] * V[0] + dT0[1] * V[1] + dT0[2] * V[2];
            const float dJ02 = dT0[0] * V[8] + dT0[1] * V[9] + dT0[2] * V[10];
            const float dJ11 = dT1[0] * V[4] + dT1[1] * V[5] + dT1[2] * V[6];
            const float dJ12 = dT1[0] * V[8] + dT1[1] * V[9] + dT1[2] * V[10];
            const float tzi = 1.0f / pr.tz, tz2 = tzi * tzi, tz3 = tz2 * tzi;
            dt_cov[0] = pr.in_x ? -cam.fx * tz2 * dJ02 : 0.f;
            dt_cov[1] = pr.in_y ? -cam.fy * tz2 * dJ12 : 0.f;
            dt_cov[2] = -cam.fx * tz2 * dJ00 - cam.fy * tz2 * dJ11 + 2.f * cam.fx * pr.ctx * tz3 * dJ02 + 2.f * cam.fy * pr.cty * tz3 * dJ12;
            const float cov_pose = p.view.enable_cov_grad ? 1.f : 0.f;
            if (p.view.enable_cov_grad) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    poseV[0 + j] += dT0[j] * pr.J00;
                    poseV[4 + j] += dT1[j] * pr.J11;
                    poseV[8 + j] += dT0[j] * pr.J02 + dT1[j] * pr.J12;
                }
            }
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float dtt = dt_pose[r] + dt_cov[r];
                dmean[0] += V[r * 4 + 0] * dtt; dmean[1] += V[r * 4 + 1] * dtt; dmean[2] += V[r * 4 + 2] * dtt;
                const float dtp = dt_pose[r] + cov_pose * dt_cov[r];
                poseV[r * 4 + 0] += dtp * a.x; poseV[r * 4 + 1] += dtp * a.y; poseV[r * 4 + 2] += dtp * a.z; poseV[r * 4 + 3] += dtp;
            }

            // ---- colour ----
            if (use_sh) {
                const unsigned cl = p.geom.clamped[i];
#pragma unroll
                for (int c = 0; c < 3; ++c) if (cl & (1u << c)) grgb[c] = 0.f;
                float dx = a.x - campos[0], dy = a.y - campos[1], dz = a.z - campos[2];
                const float len2 = dx * dx + dy * dy + dz * dz;
                const float inv = 1.0f / sqrtf(len2);
                dx *= inv; dy *= inv; dz *= inv;
                float b[16], bx[16], by[16], bz[16];
                rdg_sh_basis(deg, dx, dy, dz, b);
                sh_basis_grad(deg, dx, dy, dz, bx, by, bz);
                ddc[0] = b[0] * grgb[0]; ddc[1] = b[0] * grgb[1]; ddc[2] = b[0] * grgb[2];
                float ddx = 0.f, ddy = 0.f, ddz = 0.f;
                for (int k = 1; k < K; ++k) {
                    float* row = my_sh + (k - 1) * 3;
                    const float w = row[0] * grgb[0] + row[1] * grgb[1] + row[2] * grgb[2];
                    ddx += bx[k] * w; ddy += by[k] * w; ddz += bz[k] * w;
                    row[0] = b[k] * grgb[0]; row[1] = b[k] * grgb[1]; row[2] = b[k] * grgb[2];
                }
                for (int k = K; k < 16; ++k) { float* row = my_sh + (k - 1) * 3; row[0] = row[1] = row[2] = 0.f; }
                const float dot = dx * ddx + dy * ddy + dz * ddz;
                const float gx_ = (ddx - dx * dot) * inv, gy_ = (ddy - dy * dot) * inv, gz_ = (ddz - dz * dot) * inv;
                dmean[0] += gx_; dmean[1] += gy_; dmean[2] += gz_;
                if (p.view.enable_sh_grad) { dcam[0] -= gx_; dcam[1] -= gy_; dcam[2] -= gz_; }
            } else {
                ddc[0] = grgb[0]; ddc[1] = grgb[1]; ddc[2] = grgb[2];
            }
            dop = g1.y;
        } else if (valid && use_sh) {
            for (int k = 0; k < SH_ROW; ++k) my_sh[k] = 0.f;
        }

        // ---- through the activations / deformation, and write out ----
        const RdgSetGrad& gs = a.dyn ? p.gr.dy : p.gr.st;
        float g7[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        bool has_def = false;
        if (valid) {
            if (p.gr.means2D) { float* o = p.gr.means2D + i * 3; o[0] = dm2[0]; o[1] = dm2[1]; o[2] = 0.f; }
            if (gs.xyz) { float* o = gs.xyz + a.local * 3; o[0] = dmean[0]; o[1] = dmean[1]; o[2] = dmean[2]; }
            if (RAW) {
                if (vis) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) dscale[k] *= a.s[k];
                    dop *= a.op * (1.0f - a.op);
                    has_def = deform && a.dyn;
                    if (has_def) {
                        g7[0] = dmean[0] * sc.spatial_lr_scale; g7[1] = dmean[1] * sc.spatial_lr_scale; g7[2] = dmean[2] * sc.spatial_lr_scale;
                        g7[3] = dquat[0]; g7[4] = dquat[1]; g7[5] = dquat[2]; g7[6] = dquat[3];
                    }
                    const float dot = a.qn[0] * dquat[0] + a.qn[1] * dquat[1] + a.qn[2] * dquat[2] + a.qn[3] * dquat[3];
#pragma unroll
                    for (int k = 0; k < 4; ++k) dquat[k] = (dquat[k] - a.qn[k] * dot) * a.qinv;
                }
            }
            if (gs.scaling) { float* o = gs.scaling + a.local * 3; o[0] = dscale[0]; o[1] = dscale[1]; o[2] = dscale[2]; }
            if (gs.rotation) { float* o = gs.rotation + a.local * 4; o[0] = dquat[0]; o[1] = dquat[1]; o[2] = dquat[2]; o[3] = dquat[3]; }
            if (gs.opacity) gs.opacity[a.local] = dop;
            if (use_sh) {
                if (gs.sh_dc) { float* o = gs.sh_dc + a.local * (a.dyn ? sc.dy.sh_dc_stride : sc.st.sh_dc_stride); o[0] = ddc[0]; o[1] = ddc[1]; o[2] = ddc[2]; }
            } else if (p.gr.colors_precomp) {
                float* o = p.gr.colors_precomp + i * 3; o[0] = ddc[0]; o[1] = ddc[1]; o[2] = ddc[2];
            }
            if (RAW && deform && a.dyn && p.gr.motion_coeff) {
                float* o = p.gr.motion_coeff + a.local * sc.num_basis;
                for (int k = 0; k < sc.num_basis; ++k) {
                    float s = 0.f;
                    if (has_def) {
#pragma unroll
                        for (int j = 0; j < 7; ++j) s += rdg_basis_diff(sc, diff, a.ti, k, j) * g7[j];
                    }
                    o[k] = s;
                }
            }
        }
        // dL/dtable[ti][k][j] -= c_k * g7[j]
        if (deform && p.gr.table) {
            const unsigned any = __ballot_sync(0xffffffffu, has_def);
            if (any) {
                const int src = __ffs(any) - 1;
                const int ti0 = __shfl_sync(0xffffffffu, a.ti, src);
                const bool uniform = __all_sync(0xffffffffu, !has_def || a.ti == ti0);
                float* tab = p.dtab_in_smem ? dtab_s : p.gr.table;
                if (uniform) {
                    for (int k = 0; k < sc.num_basis; ++k) {
                        const float ck = has_def ? a.c[k] : 0.f;
#pragma unroll
                        for (int j = 0; j < 7; ++j) {
                            const float s = warp_sum(ck * g7[j]);
                            if (lane == 0) atomicAdd(&tab[(ti0 * sc.num_basis + k) * 7 + j], -s);
                        }
                    }
                } else if (has_def) {
                    for (int k = 0; k < sc.num_basis; ++k) {
                        const float ck = a.c[k];
#pragma unroll
                        for (int j = 0; j < 7; ++j) atomicAdd(&tab[(a.ti * sc.num_basis + k) * 7 + j], -ck * g7[j]);
                    }
                }
            }
        }

        // ---- coalesced write of the dSH rest rows ----
        __syncthreads();
        if (use_sh) {
            const int tot = cnt * SH_ROW;
            for (int e = threadIdx.x; e < tot; e += RDG_BLOCK) {
                const int g = e / SH_ROW, k = e - g * SH_ROW;
                const int64_t gi = base + g;
                const bool dy = gi >= sc.n_static;
                const RdgSetGrad& gset = dy ? p.gr.dy : p.gr.st;
                if (gset.sh_rest) {
                    const int64_t l = dy ? gi - sc.n_static : gi;
                    const int stride = dy ? sc.dy.sh_rest_stride : sc.st.sh_rest_stride;
                    gset.sh_rest[l * stride + k] = sh_s[e];
                }
            }
        }
    }

    // ---- pose gradient: warp shuffle -> shared -> one atomic set per CTA ----
    if (p.gr.viewmatrix) {
        float vals[15];
#pragma unroll
        for (int k = 0; k < 12; ++k) vals[k] = warp_sum(poseV[k]);
#pragma unroll
        for (int k = 0; k < 3; ++k) vals[12 + k] = warp_sum(dcam[k]);
        __syncthreads();
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 15; ++k) red_s[threadIdx.x >> 5][k] = vals[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float t[15];
            for (int k = 0; k < 15; ++k) {
                t[k] = 0.f;
                for (int w = 0; w < RDG_BLOCK / 32; ++w) t[k] += red_s[w][k];
            }
            // campos_j = -sum_i V[i][j] V[i][3]
            for (int r = 0; r < 3; ++r) {
                for (int j = 0; j < 3; ++j) {
                    t[r * 4 + j] += -t[12 + j] * V[r * 4 + 3];
                    t[r * 4 + 3] += -t[12 + j] * V[r * 4 + j];
                }
            }
            for (int r = 0; r < 3; ++r)
                for (int k = 0; k < 4; ++k) atomicAdd(&p.gr.viewmatrix[k * 4 + r], t[r * 4 + k]);
        }
    }
    // ---- flush dL/dtable, derive dL/dB(t) ----
    if (deform && p.gr.table && p.dtab_in_smem) {
        __syncthreads();
        for (int e = threadIdx.x; e < tab_n; e += RDG_BLOCK) {
            const float v = dtab_s[e];
            if (v != 0.f) atomicAdd(&p.gr.table[e], v);
        }
        if (p.gr.basis_t) {
            for (int e = threadIdx.x; e < per_t; e += RDG_BLOCK) {
                float s = 0.f;
                for (int t = 0; t < sc.num_times; ++t) s += dtab_s[t * per_t + e];
                if (s != 0.f) atomicAdd(&p.gr.basis_t[e], -s);
            }
        }
    }
}

extern "C" int rdg_preprocess_bwd(const RdgScene* scene, const RdgView* view, const RdgGeom* geom,
                                  const float* acc, const RdgSceneGrad* grads, void* stream) {
    RDG_CHECK_ARG(scene && view && geom && acc && grads, "null argument");
    const int64_t N = scene->n_static + scene->n_dynamic;
    if (N == 0) return RDG_OK;
    RDG_CHECK_ARG(view->sh_degree >= 0 && view->sh_degree <= 3, "sh_degree must be 0..3");
    const bool deform = scene->raw && scene->use_deform && scene->n_dynamic > 0;
    PreBwdParams p;
    p.sc = *scene; p.view = *view; p.geom = *geom; p.gr = *grads; p.acc = acc;
    // shared memory: dSH staging (45 KB) + the dL/dtable accumulators; the B(t)-table difference joins
    // them only while two CTAs still fit per SM, otherwise it is read through L1.
    size_t smem = RDG_BLOCK * SH_ROW * sizeof(float);
    const size_t tab_bytes = deform ? (size_t)scene->num_times * scene->num_basis * 7 * sizeof(float) : 0;
    p.dtab_in_smem = (deform && grads->table && smem + tab_bytes <= 200 * 1024) ? 1 : 0;
    p.diff_in_smem = (deform && smem + (p.dtab_in_smem ? 2 : 1) * tab_bytes <= 100 * 1024) ? 1 : 0;
    smem += (size_t)(p.dtab_in_smem + p.diff_in_smem) * tab_bytes;
    if (deform && grads->table && !p.dtab_in_smem && grads->basis_t) {
        rdg_set_error("rdg_preprocess_bwd: motion table with %d times does not fit in shared memory; "
                      "pass basis_t = NULL and reduce dL/dtable on the host side", scene->num_times);
        return RDG_E_ARG;
    }
    const int64_t chunks = (N + RDG_BLOCK - 1) / RDG_BLOCK;
    const int per_sm = smem > 100 * 1024 ? 1 : 2;
    const int64_t cap = (int64_t)RDG_SM_COUNT * per_sm * (deform ? 1 : 4);
    const int grid = (int)(chunks < cap ? chunks : cap);
    cudaStream_t s = (cudaStream_t)stream;
    if (scene->raw) {
        RDG_CUDA(cudaFuncSetAttribute(preprocess_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        preprocess_bwd_kernel<true><<<grid, RDG_BLOCK, smem, s>>>(p);
    } else {
        RDG_CUDA(cudaFuncSetAttribute(preprocess_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        preprocess_bwd_kernel<false><<<grid, RDG_BLOCK, smem, s>>>(p);
    }
    RDG_CHECK_LAUNCH();
    rdg_count_launches(1);
    return RDG_OK;
}
