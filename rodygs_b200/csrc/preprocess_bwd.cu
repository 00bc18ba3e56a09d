// Per-Gaussian backward: screen-space gradients (acc, from blend backward) ->
// gradients of means / scales / quaternions / opacity / SH (+ through the fused
// activations and the time-conditioned deformation when scene.raw) and of the pose.
// SURVEY.md §8 row a10, App. A.6-A.7; checked against autograd of the oracle.
//
// HBM-bound: reads 48 B acc + the forward inputs again, writes 44 B + 12*K B (dSH)
// + 12 B (means2D sink) [+ 64 B dcoeff + 32 B g7] per Gaussian.
//
// B200 mapping: same chunking as the forward kernel (256 Gaussians of ONE model per
// chunk).  The chunk's SH rows arrive by one TMA bulk load, every thread turns its row
// into the dSH row in place, and the 46 KB of gradients leave by one TMA bulk store -
// both directions fully coalesced without LSU instructions.  The motion-table gradient
// dL/dtable[t] = -sum_{i born at t} c_i (x) g_i is NOT accumulated with atomics here:
// the kernel writes g_i (7 floats) per dynamic Gaussian and `dtable_kernel` reduces them
// per birth frame over a CSR built once by the host (frame_order / frame_offsets).
#include <stdlib.h>
#include "scene.cuh"
#include "tma.cuh"

#define SH_ROW 45
#define NACC 12

struct PreBwdParams {
    RdgScene sc;
    RdgView view;
    RdgGeom geom;
    RdgSceneGrad gr;
    const float* acc;
    int use_tma;
    int diff_smem;     // B(t) - table rows staged in shared memory (rdg_stage_diff)
    int dtab_atomic;   // no CSR: accumulate dL/dtable with global atomics from this kernel
    int64_t c_begin, c_end;   // chunk range of this launch (RdgSceneGrad.models: one model at a time under data parallelism)
    int l2_prefetch;          // next chunk's parameter / accumulator rows prefetched into L2 (cp.async.bulk.prefetch.L2)
    uint32_t* sm_queue;       // non-NULL: SM-partitioned mode - CTAs that land on an SM >= sm_limit exit at once, the others take
    int sm_limit;             // chunks from this counter, so that SMs sm_limit.. stay FREE for a collective on another stream
};

template <int DEG>
__device__ __forceinline__ void sh_basis_grad(float x, float y, float z, float* bx, float* by, float* bz) {
    bx[0] = by[0] = bz[0] = 0.f;
    if (DEG > 0) {
        bx[1] = 0.f;         by[1] = -RDG_SH_C1; bz[1] = 0.f;
        bx[2] = 0.f;         by[2] = 0.f;        bz[2] = RDG_SH_C1;
        bx[3] = -RDG_SH_C1;  by[3] = 0.f;        bz[3] = 0.f;
        if (DEG > 1) {
            bx[4] = RDG_SH_C2_0 * y;          by[4] = RDG_SH_C2_0 * x;          bz[4] = 0.f;
            bx[5] = 0.f;                      by[5] = RDG_SH_C2_1 * z;          bz[5] = RDG_SH_C2_1 * y;
            bx[6] = RDG_SH_C2_2 * -2.f * x;   by[6] = RDG_SH_C2_2 * -2.f * y;   bz[6] = RDG_SH_C2_2 * 4.f * z;
            bx[7] = RDG_SH_C2_3 * z;          by[7] = 0.f;                      bz[7] = RDG_SH_C2_3 * x;
            bx[8] = RDG_SH_C2_4 * 2.f * x;    by[8] = RDG_SH_C2_4 * -2.f * y;   bz[8] = 0.f;
            if (DEG > 2) {
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

// Measured alternatives that did NOT help (B200, C4, profiles/r01_ab_v12_*.json, r01_ab_v13_*.json): double-buffering the SH
// rows (0.489 vs 0.464 ms) and 3 CTAs per SM via __launch_bounds__(256, 3) (80 registers, 210 B of spills; 0.564 vs 0.467 ms).
template <bool RAW, int DEG>
__global__ void __launch_bounds__(RDG_BLOCK, 2) preprocess_bwd_kernel(const PreBwdParams p) {
    extern __shared__ __align__(128) float smem[];
    constexpr int K = (DEG + 1) * (DEG + 1);
    constexpr int NREST = 3 * (K - 1);
    const RdgScene& sc = p.sc;
    if (p.sm_queue) {                         // leave the upper SMs to NCCL (a persistent grid cannot be pinned, but it can decline)
        unsigned smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        if ((int)smid >= p.sm_limit) {
            if (threadIdx.x == 0) rdg_queue_release(p.sm_queue, gridDim.x);
            return;
        }
    }
    const bool use_sh = sc.colors_precomp == nullptr;
    const bool deform = RAW && sc.use_deform && sc.n_dynamic > 0;
    float* sh_s = smem;                       // [256][SH_ROW] in: SH rest coefficients, out: their gradients
    float* bt_s = smem + RDG_BLOCK * SH_ROW;  // [16*7] B(t)
    __shared__ float red_s[RDG_BLOCK / 32][16];
    __shared__ __align__(8) uint64_t bar;
    __shared__ long long q_s[2];

    RdgCam cam;
    rdg_load_cam(cam, p.view.viewmatrix, p.view.projmatrix, p.view.tanfovx, p.view.tanfovy, p.view.width, p.view.height);
    float campos[3];
    rdg_campos(cam, campos);
    const float* V = cam.V;
    const float* P = cam.P;

    float* diff_s = (deform && p.diff_smem) ? bt_s + RDG_NUM_BASIS_MAX * 7 : nullptr;
    if (deform) {
        for (int e = threadIdx.x; e < sc.num_basis * 7; e += RDG_BLOCK) bt_s[e] = sc.basis_t[e];
        if (diff_s) rdg_stage_diff(sc, sc.basis_t, diff_s, RDG_BLOCK);
    }
    if (threadIdx.x == 0) {
        rdg_mbar_init(&bar, 1);
        if (p.sm_queue) {                     // this CTA's first two chunks
            q_s[0] = p.c_begin + (long long)atomicAdd(p.sm_queue, 1u);
            q_s[1] = p.c_begin + (long long)atomicAdd(p.sm_queue, 1u);
        }
    }
    __syncthreads();
    const bool queued = p.sm_queue != nullptr;
    long long chunk_nx = queued ? q_s[1] : p.c_begin + (long long)blockIdx.x + gridDim.x;
    const long long chunk_first = queued ? q_s[0] : p.c_begin + (long long)blockIdx.x;

    float poseV[12];   // dL/dV rows 0..2 (row-major), this thread's partial sum
    float dcam[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 12; ++k) poseV[k] = 0.f;

    const int lane = threadIdx.x & 31;
    const int64_t cs = (sc.n_static + RDG_BLOCK - 1) / RDG_BLOCK, cd = (sc.n_dynamic + RDG_BLOCK - 1) / RDG_BLOCK;
    uint32_t phase = 0;
    bool store_pending = false;
    // radius (visibility) and clamp flags of this thread's Gaussian are fetched ONE CHUNK AHEAD: everything else a thread
    // loads hangs off `radius > 0`, so reading it at the top of its own chunk puts a DRAM round trip in front of the rest
    auto chunk_slot = [&](int64_t chunk, int64_t& gi) -> bool {
        const bool dyn_c = chunk >= cs;
        const int64_t lb = (dyn_c ? chunk - cs : chunk) * RDG_BLOCK;
        const int64_t n_c = dyn_c ? sc.n_dynamic : sc.n_static;
        gi = (dyn_c ? sc.n_static : 0) + lb + threadIdx.x;
        return lb + (int64_t)threadIdx.x < n_c;
    };
    int radius_pf = 0;
    unsigned clamped_pf = 0;
    if (chunk_first < p.c_end) {
        int64_t gi;
        if (chunk_slot(chunk_first, gi)) { radius_pf = p.geom.radii[gi]; clamped_pf = p.geom.clamped[gi]; }
    }
    int q_par = 0;
    for (int64_t chunk = chunk_first; chunk < p.c_end; ) {
        const bool dyn = chunk >= cs;
        const RdgSet& set = dyn ? sc.dy : sc.st;
        const RdgSetGrad& gs = dyn ? p.gr.dy : p.gr.st;
        const int64_t lbase = (dyn ? chunk - cs : chunk) * RDG_BLOCK;
        const int64_t n_set = dyn ? sc.n_dynamic : sc.n_static;
        const int cnt = (int)min((int64_t)RDG_BLOCK, n_set - lbase);
        const bool valid = (int)threadIdx.x < cnt;
        const int64_t local = lbase + threadIdx.x;
        const int64_t i = (dyn ? sc.n_static : 0) + local;
        const int radius = radius_pf;
        const unsigned cl_pf = clamped_pf;
        const bool vis = radius > 0;
        radius_pf = 0;
        clamped_pf = 0;
        if (chunk_nx < p.c_end) {
            int64_t gi;
            if (chunk_slot(chunk_nx, gi)) { radius_pf = p.geom.radii[gi]; clamped_pf = p.geom.clamped[gi]; }
        }
        if (p.l2_prefetch && chunk_nx < p.c_end && threadIdx.x < 8) rdg_prefetch_chunk_field(sc, chunk_nx, cs, threadIdx.x, p.acc);
        // the chunk after the next one (claimed by thread 0, published by the barrier below, read at the end of the iteration)
        if (queued && threadIdx.x == 0) q_s[q_par] = p.c_begin + (long long)atomicAdd(p.sm_queue, 1u);

        // ---- stage the SH rest rows ----
        const bool full_rows = use_sh && set.sh_rest_stride == SH_ROW;
        const bool tma_in = full_rows && p.use_tma && NREST == SH_ROW && (cnt & 3) == 0;
        const bool tma_out = full_rows && p.use_tma && gs.sh_rest != nullptr && (cnt & 3) == 0;
        if (threadIdx.x == 0 && store_pending) { rdg_bulk_store_wait_read(); store_pending = false; }
        __syncthreads();   // previous chunk fully written out / read
        // (two slots: thread 0 may write the next iteration's claim while slower threads still read this one)
        const long long chunk_after = queued ? q_s[q_par] : chunk_nx + (long long)gridDim.x;
        q_par ^= 1;
        if (use_sh && NREST > 0) {
            if (tma_in) {
                if (threadIdx.x == 0) {
                    rdg_fence_proxy_async();
                    rdg_bulk_load(sh_s, set.sh_rest + lbase * SH_ROW, (uint32_t)(cnt * SH_ROW * sizeof(float)), &bar);
                }
            } else {
                const int stride = set.sh_rest_stride;
                const float* src = set.sh_rest + lbase * stride;
                for (int e = threadIdx.x; e < cnt * NREST; e += RDG_BLOCK) {
                    const int g = e / NREST, k = e - g * NREST;
                    sh_s[g * SH_ROW + k] = src[(int64_t)g * stride + k];
                }
            }
        }

        RdgAct a;
        a.dyn = dyn; a.local = local; a.ti = 0;
        float dmean[3] = {0.f, 0.f, 0.f}, dscale[3] = {0.f, 0.f, 0.f}, dquat[4] = {0.f, 0.f, 0.f, 0.f};
        float dop = 0.f, ddc[3] = {0.f, 0.f, 0.f}, dm2[2] = {0.f, 0.f};
        float grgb[3] = {0.f, 0.f, 0.f};
        float* my_sh = sh_s + threadIdx.x * SH_ROW;
        if (vis) {
            rdg_fetch<RAW>(sc, dyn, local, bt_s, a, diff_s);
            RdgProj pr;
            rdg_project(cam, a, p.view.scale_modifier, pr);
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.acc) + i * 3 + 0);
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.acc) + i * 3 + 1);
            const float4 g2 = __ldg(reinterpret_cast<const float4*>(p.acc) + i * 3 + 2);
            const float gA = g0.z, gB = g0.w, gC = g1.x;
            grgb[0] = g1.z; grgb[1] = g1.w; grgb[2] = g2.x;
            const float gdepth = g2.y;
            dop = g1.y;

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
        }

        // ---- colour: needs the staged SH rows ----
        if (use_sh && NREST > 0) {
            if (tma_in) { rdg_mbar_wait(&bar, phase & 1u); ++phase; }
            else __syncthreads();
        }
        if (vis) {
            if (use_sh) {
                const unsigned cl = cl_pf;
#pragma unroll
                for (int c = 0; c < 3; ++c) if (cl & (1u << c)) grgb[c] = 0.f;
                float dx = a.x - campos[0], dy = a.y - campos[1], dz = a.z - campos[2];
                const float len2 = dx * dx + dy * dy + dz * dz;
                const float inv = 1.0f / sqrtf(len2);
                dx *= inv; dy *= inv; dz *= inv;
                float b[K], bx[K], by[K], bz[K];
                rdg_sh_basis<DEG>(dx, dy, dz, b);
                sh_basis_grad<DEG>(dx, dy, dz, bx, by, bz);
                ddc[0] = b[0] * grgb[0]; ddc[1] = b[0] * grgb[1]; ddc[2] = b[0] * grgb[2];
                float ddx = 0.f, ddy = 0.f, ddz = 0.f;
#pragma unroll
                for (int k = 1; k < K; ++k) {
                    float* row = my_sh + (k - 1) * 3;
                    const float w = row[0] * grgb[0] + row[1] * grgb[1] + row[2] * grgb[2];
                    ddx += bx[k] * w; ddy += by[k] * w; ddz += bz[k] * w;
                    row[0] = b[k] * grgb[0]; row[1] = b[k] * grgb[1]; row[2] = b[k] * grgb[2];
                }
#pragma unroll
                for (int k = K; k < 16; ++k) { float* row = my_sh + (k - 1) * 3; row[0] = row[1] = row[2] = 0.f; }
                const float dot = dx * ddx + dy * ddy + dz * ddz;
                const float gx_ = (ddx - dx * dot) * inv, gy_ = (ddy - dy * dot) * inv, gz_ = (ddz - dz * dot) * inv;
                dmean[0] += gx_; dmean[1] += gy_; dmean[2] += gz_;
                if (p.view.enable_sh_grad) { dcam[0] -= gx_; dcam[1] -= gy_; dcam[2] -= gz_; }
            } else {
                ddc[0] = grgb[0]; ddc[1] = grgb[1]; ddc[2] = grgb[2];
            }
        } else if (valid && use_sh) {
#pragma unroll
            for (int k = 0; k < SH_ROW; ++k) my_sh[k] = 0.f;
        }

        // factors of dL/dSH for the data-parallel exchange (rdg_sh_grad_views): dL/d(rgb) after the clamp mask
        if (valid && p.gr.dcolor) {
            float* o = p.gr.dcolor + i * 3;
            o[0] = vis ? grgb[0] : 0.f; o[1] = vis ? grgb[1] : 0.f; o[2] = vis ? grgb[2] : 0.f;
        }

        // ---- through the activations / deformation, and write out ----
        float g7[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        bool has_def = false;
        if (valid) {
            if (p.gr.means2D) { float* o = p.gr.means2D + i * 3; o[0] = dm2[0]; o[1] = dm2[1]; o[2] = 0.f; }
            if (gs.xyz) { float* o = gs.xyz + local * 3; o[0] = dmean[0]; o[1] = dmean[1]; o[2] = dmean[2]; }
            if (RAW && vis) {
#pragma unroll
                for (int k = 0; k < 3; ++k) dscale[k] *= a.s[k];
                dop *= a.op * (1.0f - a.op);
                has_def = deform && dyn;
                if (has_def) {
                    g7[0] = dmean[0] * sc.spatial_lr_scale; g7[1] = dmean[1] * sc.spatial_lr_scale; g7[2] = dmean[2] * sc.spatial_lr_scale;
                    g7[3] = dquat[0]; g7[4] = dquat[1]; g7[5] = dquat[2]; g7[6] = dquat[3];
                }
                const float dot = a.qn[0] * dquat[0] + a.qn[1] * dquat[1] + a.qn[2] * dquat[2] + a.qn[3] * dquat[3];
#pragma unroll
                for (int k = 0; k < 4; ++k) dquat[k] = (dquat[k] - a.qn[k] * dot) * a.qinv;
            }
            if (gs.scaling) { float* o = gs.scaling + local * 3; o[0] = dscale[0]; o[1] = dscale[1]; o[2] = dscale[2]; }
            if (gs.rotation) reinterpret_cast<float4*>(gs.rotation)[local] = make_float4(dquat[0], dquat[1], dquat[2], dquat[3]);
            if (gs.opacity) gs.opacity[local] = dop;
            if (use_sh) {
                if (gs.sh_dc) { float* o = gs.sh_dc + local * set.sh_dc_stride; o[0] = ddc[0]; o[1] = ddc[1]; o[2] = ddc[2]; }
            } else if (p.gr.colors_precomp) {
                float* o = p.gr.colors_precomp + i * 3; o[0] = ddc[0]; o[1] = ddc[1]; o[2] = ddc[2];
            }
            if (deform && dyn) {
                if (p.gr.motion_coeff) {
                    float dcf[RDG_NUM_BASIS_MAX];
#pragma unroll
                    for (int k = 0; k < RDG_NUM_BASIS_MAX; ++k) dcf[k] = 0.f;
                    if (has_def) {
                        if (sc.num_basis == RDG_NUM_BASIS_MAX && diff_s) {
                            const float4* row4 = reinterpret_cast<const float4*>(diff_s + a.ti * RDG_DIFF_STRIDE);
#pragma unroll
                            for (int q = 0; q < 28; ++q) {
                                const float4 v = row4[q];
                                const float r[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                                for (int m = 0; m < 4; ++m) {
                                    const int e = 4 * q + m;
                                    dcf[e / 7] += r[m] * g7[e % 7];
                                }
                            }
                        } else if (sc.num_basis == RDG_NUM_BASIS_MAX) {
                            const float4* row4 = reinterpret_cast<const float4*>(sc.table + (int64_t)a.ti * 112);
#pragma unroll
                            for (int q = 0; q < 28; ++q) {
                                const float4 v = __ldg(row4 + q);
                                const float r[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                                for (int m = 0; m < 4; ++m) {
                                    const int e = 4 * q + m;
                                    dcf[e / 7] += (bt_s[e] - r[m]) * g7[e % 7];
                                }
                            }
                        } else {
                            for (int k = 0; k < sc.num_basis; ++k)
#pragma unroll
                                for (int j = 0; j < 7; ++j) dcf[k] += rdg_basis_diff(sc, bt_s, a.ti, k, j) * g7[j];
                        }
                    }
                    float* o = p.gr.motion_coeff + local * sc.num_basis;
                    if (sc.num_basis == RDG_NUM_BASIS_MAX) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            reinterpret_cast<float4*>(o)[q] = make_float4(dcf[4 * q], dcf[4 * q + 1], dcf[4 * q + 2], dcf[4 * q + 3]);
                    } else {
                        for (int k = 0; k < sc.num_basis; ++k) o[k] = dcf[k];
                    }
                }
                if (p.gr.g7_scratch) {
                    float4* o = reinterpret_cast<float4*>(p.gr.g7_scratch + local * 8);
                    o[0] = make_float4(g7[0], g7[1], g7[2], g7[3]);
                    o[1] = make_float4(g7[4], g7[5], g7[6], 0.f);
                }
            }
        }
        // fallback without the birth-frame CSR: dL/dtable[ti][k][j] -= c_k g7[j] with global atomics
        if (deform && dyn && p.dtab_atomic && p.gr.table && has_def) {
            for (int k = 0; k < sc.num_basis; ++k) {
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    const float v = a.c[k] * g7[j];
                    atomicAdd(&p.gr.table[(a.ti * sc.num_basis + k) * 7 + j], -v);
                    if (p.gr.basis_t) atomicAdd(&p.gr.basis_t[k * 7 + j], v);
                }
            }
        }

        // ---- write the dSH rest rows ----
        if (use_sh && gs.sh_rest) {
            if (tma_out) {
                rdg_fence_proxy_async();   // my generic-proxy writes -> visible to the async proxy
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
                    const int g = e / SH_ROW, k = e - g * SH_ROW;
                    dst[(int64_t)g * stride + k] = sh_s[e];
                }
            }
        }
        chunk = chunk_nx;
        chunk_nx = chunk_after;
    }
    if (threadIdx.x == 0 && store_pending) rdg_bulk_store_wait_read();
    if (queued && threadIdx.x == 0) rdg_queue_release(p.sm_queue, gridDim.x);

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
}

// dL/dtable[t][k][j] = -sum_{i : birth frame t} c_i[k] g_i[j];  dL/dB(t)[k][j] = +sum_i c_i[k] g_i[j].
// grid (T, DT_SLICES); thread e < K*7 owns one (k, j) and accumulates in a register over the
// frame's Gaussians, staged 64 at a time in shared memory - no atomics until the final add.
#define DT_SLICES 32
#define DT_TILE 64
__global__ void __launch_bounds__(128) dtable_kernel(const int32_t* __restrict__ order, const int32_t* __restrict__ offsets,
                                                     const float* __restrict__ coeff, const float* __restrict__ g7, int num_basis,
                                                     float* __restrict__ dtable, float* __restrict__ dbasis) {
    __shared__ float c_s[DT_TILE][RDG_NUM_BASIS_MAX + 1];
    __shared__ float g_s[DT_TILE][8];
    const int t = blockIdx.x;
    const int beg = offsets[t], end = offsets[t + 1];
    const int per_t = num_basis * 7;
    const int e = threadIdx.x;
    const int k = e / 7, j = e - k * 7;
    float accv = 0.f;
    for (int base = beg + blockIdx.y * DT_TILE; base < end; base += DT_SLICES * DT_TILE) {
        const int cnt = min(DT_TILE, end - base);
        __syncthreads();
        for (int x = threadIdx.x; x < cnt * num_basis; x += blockDim.x) {
            const int g = x / num_basis, kk = x - g * num_basis;
            c_s[g][kk] = coeff[(int64_t)order[base + g] * num_basis + kk];
        }
        for (int x = threadIdx.x; x < cnt * 8; x += blockDim.x) {
            const int g = x >> 3, jj = x & 7;
            g_s[g][jj] = g7[(int64_t)order[base + g] * 8 + jj];
        }
        __syncthreads();
        if (e < per_t)
            for (int g = 0; g < cnt; ++g) accv = fmaf(c_s[g][k], g_s[g][j], accv);
    }
    if (e < per_t && accv != 0.f) {
        atomicAdd(&dtable[(int64_t)t * per_t + e], -accv);
        if (dbasis) atomicAdd(&dbasis[e], accv);
    }
}

// v2 of the same reduction (default; RDG_DTABLE_V1=1 selects the kernel above for A/B).  ncu r01 on v1: 37.7 M warp
// instructions (two LDS per FMA, runtime-divisor index math in the gather), and 3200 CTAs x 112 RED on the SAME 112
// dL/dB(t) addresses - four cache lines whose L2 atomic unit serialises them (0.85 cycles per op, B300_MICROARCH.md).
//   * persistent grid (RDG_SM_COUNT x 4 CTAs) over (frame, slice) items; dL/dB(t) is accumulated in registers over
//     ALL items of a CTA and added once per CTA at the end: 5x fewer same-address REDs;
//   * register tiling: warp w owns bases 4w..4w+3, lane l the Gaussians l, l+32, ... of the 128 staged in shared
//     memory -> 28 FMA per 3 conflict-free LDS.128 (rows padded to 20 / 12 floats), one butterfly per item;
//   * one Gaussian per thread in the gather: 1 index load, then six independent 16-byte loads.
#define DT2_TILE 128
#define DT2_CS 20     // padded row strides (floats): 16-byte reads of 8 consecutive rows hit 8 distinct bank groups
#define DT2_GS 12
__global__ void __launch_bounds__(128) dtable2_kernel(const int32_t* __restrict__ order, const int32_t* __restrict__ offsets,
                                                      const float* __restrict__ coeff, const float* __restrict__ g7,
                                                      int num_basis, int num_times, int slices, float* __restrict__ dtable,
                                                      float* __restrict__ dbasis) {
    __shared__ __align__(16) float c_s[DT2_TILE * DT2_CS];
    __shared__ __align__(16) float g_s[DT2_TILE * DT2_GS];
    const int lane = threadIdx.x & 31, kq = threadIdx.x >> 5;
    const int items = num_times * slices;
    float bas[4][7];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int j = 0; j < 7; ++j) bas[a][j] = 0.f;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int t = item / slices, sl = item - t * slices;
        const int beg = offsets[t], end = offsets[t + 1];
        float acc[4][7];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int j = 0; j < 7; ++j) acc[a][j] = 0.f;
        for (int base = beg + sl * DT2_TILE; base < end; base += slices * DT2_TILE) {
            const int cnt = min(DT2_TILE, end - base);
            __syncthreads();                       // the previous round's readers are done
            {
                const int i = threadIdx.x;
                float4 c[4], ga, gb;
                c[0] = c[1] = c[2] = c[3] = ga = gb = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < cnt) {
                    const int64_t id = order[base + i];
                    if (num_basis == 16) {
                        const float4* cr = reinterpret_cast<const float4*>(coeff + id * 16);
                        c[0] = __ldg(cr); c[1] = __ldg(cr + 1); c[2] = __ldg(cr + 2); c[3] = __ldg(cr + 3);
                    } else {
                        float tmp[16];
#pragma unroll
                        for (int k = 0; k < 16; ++k) tmp[k] = k < num_basis ? __ldg(coeff + id * num_basis + k) : 0.f;
#pragma unroll
                        for (int q = 0; q < 4; ++q) c[q] = make_float4(tmp[4 * q], tmp[4 * q + 1], tmp[4 * q + 2], tmp[4 * q + 3]);
                    }
                    const float4* gr = reinterpret_cast<const float4*>(g7 + id * 8);
                    ga = __ldg(gr); gb = __ldg(gr + 1);
                }
                float4* cd = reinterpret_cast<float4*>(c_s + i * DT2_CS);
                cd[0] = c[0]; cd[1] = c[1]; cd[2] = c[2]; cd[3] = c[3];
                float4* gd = reinterpret_cast<float4*>(g_s + i * DT2_GS);
                gd[0] = ga; gd[1] = gb;
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < DT2_TILE / 32; ++r) {
                const int g = lane + 32 * r;
                const float4 c4 = *reinterpret_cast<const float4*>(c_s + g * DT2_CS + 4 * kq);
                const float4 ga = *reinterpret_cast<const float4*>(g_s + g * DT2_GS);
                const float4 gb = *reinterpret_cast<const float4*>(g_s + g * DT2_GS + 4);
                const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
                const float gg[7] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int j = 0; j < 7; ++j) acc[a][j] = fmaf(cc[a], gg[j], acc[a][j]);
            }
        }
        // item total over the 32 lanes (butterfly: every lane ends with the sum); lane 0 adds it to dL/dtable[t]
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                float v = acc[a][j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                bas[a][j] += v;
                const int k = 4 * kq + a;
                if (lane == 0 && k < num_basis && v != 0.f) atomicAdd(&dtable[((int64_t)t * num_basis + k) * 7 + j], -v);
            }
    }
    if (dbasis && lane == 0) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const int k = 4 * kq + a;
                if (k < num_basis && bas[a][j] != 0.f) atomicAdd(&dbasis[k * 7 + j], bas[a][j]);
            }
    }
}

template <bool RAW, int DEG>
static int launch_bwd(const PreBwdParams& p, int grid, size_t smem, cudaStream_t s) {
    const int cap = rdg_tunable(RDG_TUN_PRE_GRID_CAP);
    if (cap > 0 && grid > cap) grid = cap;
    // deterministic mode: ONE persistent CTA walks every chunk, so the per-CTA atomics of dL/dV land in a fixed order
    if (rdg_tunable(RDG_TUN_DETERMINISTIC) != 0) grid = 1;
    RDG_CUDA(cudaFuncSetAttribute(preprocess_bwd_kernel<RAW, DEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    preprocess_bwd_kernel<RAW, DEG><<<grid, RDG_BLOCK, smem, s>>>(p);
    RDG_CHECK_LAUNCH();
    return RDG_OK;
}

template <bool RAW>
static int launch_bwd_deg(const PreBwdParams& p, int deg, int grid, size_t smem, cudaStream_t s) {
    switch (deg) {
        case 0: return launch_bwd<RAW, 0>(p, grid, smem, s);
        case 1: return launch_bwd<RAW, 1>(p, grid, smem, s);
        case 2: return launch_bwd<RAW, 2>(p, grid, smem, s);
        default: return launch_bwd<RAW, 3>(p, grid, smem, s);
    }
}

extern "C" int rdg_preprocess_bwd(const RdgScene* scene, const RdgView* view, const RdgGeom* geom,
                                  const float* acc, const RdgSceneGrad* grads, void* stream) {
    RDG_CHECK_ARG(scene && view && geom && acc && grads, "null argument");
    const int64_t N = scene->n_static + scene->n_dynamic;
    if (N == 0) return RDG_OK;
    RDG_CHECK_ARG(view->sh_degree >= 0 && view->sh_degree <= 3, "sh_degree must be 0..3");
    const bool deform = scene->raw && scene->use_deform && scene->n_dynamic > 0;
    const bool csr = deform && scene->frame_order && scene->frame_offsets;
    RDG_CHECK_ARG(!(csr && grads->table) || grads->g7_scratch, "g7_scratch is required with frame_order");
    PreBwdParams p;
    p.sc = *scene; p.view = *view; p.geom = *geom; p.gr = *grads; p.acc = acc;
    if (!(csr && grads->table)) p.gr.g7_scratch = nullptr;
    p.dtab_atomic = (deform && grads->table && !csr) ? 1 : 0;
    const uintptr_t al = (uintptr_t)scene->st.sh_rest | (uintptr_t)scene->dy.sh_rest | (uintptr_t)grads->st.sh_rest |
                         (uintptr_t)grads->dy.sh_rest;
    p.use_tma = (al & 15u) == 0 ? 1 : 0;
    size_t smem = (RDG_BLOCK * SH_ROW + RDG_NUM_BASIS_MAX * 7) * sizeof(float);
    p.diff_smem = 0;
    if (deform && scene->num_basis == RDG_NUM_BASIS_MAX && rdg_tunable(RDG_TUN_DIFF_SMEM) != 0 &&
        (((uintptr_t)scene->basis_t | (uintptr_t)scene->table) & 15u) == 0) {
        const size_t extra = (size_t)scene->num_times * RDG_DIFF_STRIDE * sizeof(float);
        if (smem + extra <= RDG_PRE_SMEM_MAX) { p.diff_smem = 1; smem += extra; }
    }
    const int64_t cs_h = (scene->n_static + RDG_BLOCK - 1) / RDG_BLOCK, cd_h = (scene->n_dynamic + RDG_BLOCK - 1) / RDG_BLOCK;
    const bool do_static = grads->models != 2, do_dynamic = grads->models != 1;
    p.sm_queue = nullptr;
    p.sm_limit = 0;
    p.l2_prefetch = rdg_tunable(RDG_TUN_L2_PREFETCH) != 0 ? 1 : 0;
    p.c_begin = do_static ? 0 : cs_h;
    p.c_end = do_dynamic ? cs_h + cd_h : cs_h;
    RDG_CHECK_ARG(grads->parts >= 0 && grads->part >= 0 && (grads->parts == 0 ? grads->part == 0 : grads->part < grads->parts), "bad part / parts");
    RDG_CHECK_ARG(grads->dtable_mode >= 0 && grads->dtable_mode <= 2, "dtable_mode must be 0..2");
    const bool last_part = grads->parts <= 1 || grads->part == grads->parts - 1;
    if (grads->parts > 1) {
        const int64_t all = p.c_end - p.c_begin, per = (all + grads->parts - 1) / grads->parts;
        const int64_t b = p.c_begin + per * grads->part, e = b + per;
        p.c_begin = b < p.c_end ? b : p.c_end;
        p.c_end = e < p.c_end ? e : p.c_end;
    }
    const bool run_dtable = csr && grads->table && ((grads->dtable_mode == 0 && do_dynamic && last_part) || grads->dtable_mode == 2);
    const int64_t chunks = grads->dtable_mode == 2 ? 0 : p.c_end - p.c_begin;
    if (chunks <= 0 && !run_dtable) return RDG_OK;
    // 2 CTAs per SM (register-limited), persistent; "sm_reserve" leaves SMs to a collective running beside this kernel
    int sms = RDG_SM_COUNT - rdg_tunable(RDG_TUN_SM_RESERVE);
    if (sms < 8) sms = 8;
    const int64_t cap = (int64_t)sms * 2;
    int grid = (int)(chunks < cap ? chunks : cap);
    if (grid < 1) grid = 1;
    cudaStream_t s = (cudaStream_t)stream;
    if (grads->sm_queue && sms < RDG_SM_COUNT && chunks > cap && rdg_tunable(RDG_TUN_DETERMINISTIC) == 0) {
        // SM-partitioned mode: a full-machine grid whose CTAs on SMs >= sms exit at once; the others share a chunk queue
        p.sm_queue = grads->sm_queue;      // zero on entry, left zero by the kernel (rdg_queue_release)
        p.sm_limit = sms;
        grid = RDG_SM_COUNT * 2;
    }
    if (chunks > 0) {
        const int rc = scene->raw ? launch_bwd_deg<true>(p, view->sh_degree, grid, smem, s)
                                  : launch_bwd_deg<false>(p, view->sh_degree, grid, smem, s);
        if (rc) return rc;
        rdg_count_launches(1);
    }
    if (run_dtable) {
        if (rdg_tunable(RDG_TUN_DTABLE_V1) != 0) {
            const dim3 g(scene->num_times, DT_SLICES);
            dtable_kernel<<<g, 128, 0, s>>>(scene->frame_order, scene->frame_offsets, scene->motion_coeff, grads->g7_scratch,
                                            scene->num_basis, grads->table, grads->basis_t);
        } else {
            // ~3200 (frame, slice) items so that a persistent grid of 4 CTAs per SM stays balanced (5.4 items per CTA)
            int slices = (3200 + scene->num_times - 1) / scene->num_times;
            slices = slices < 1 ? 1 : (slices > 32 ? 32 : slices);
            const int64_t items = (int64_t)scene->num_times * slices;
            int g = (int)(items < sms * 4 ? items : sms * 4);
            if (rdg_tunable(RDG_TUN_DETERMINISTIC) != 0) g = 1;   // one CTA: the slices of a frame add to dL/dtable in a fixed order
            dtable2_kernel<<<g, 128, 0, s>>>(scene->frame_order, scene->frame_offsets, scene->motion_coeff,
                                             grads->g7_scratch, scene->num_basis, scene->num_times, slices, grads->table,
                                             grads->basis_t);
        }
        RDG_CHECK_LAUNCH();
        rdg_count_launches(1);
    }
    return RDG_OK;
}
