// Motion regularisers that act on the coefficients and on the basis table.
//
//   rdg_motion_coeff_reg : MotionL1Loss + MotionSparsityLoss     (/root/reference/src/trainer/losses.py:364-379)
//   rdg_motion_basis_reg : MotionBasisRegularizaiton, degree 0   (/root/reference/src/trainer/losses.py:382-525)
//
// The reference builds each of them from 4-10 element-wise PyTorch kernels plus their autograd backward; here each
// is ONE launch that returns the value and adds the weighted gradient to the caller's gradient buffer (the flat
// gradient buffer of the trainer, so the optimiser needs no extra pass).
//
// B200 mapping: the coefficient pass is a pure stream - 64 B of coefficients in, 64 B of gradient read-modify-write,
// float4 accesses, a persistent grid of RDG_SM_COUNT x 8 CTAs; bound: HBM (128-192 B per dynamic Gaussian).
// The table pass touches T x 16 x 7 floats (45 KB at T = 100) - launch-latency bound, one thread per (time, basis).
#include <math.h>
#include "common.cuh"

__device__ __forceinline__ double block_sum_double(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = l < (blockDim.x >> 5) ? sh[l] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    return r;   // valid in thread 0
}

__device__ __forceinline__ float sgn(float v) { return (float)(v > 0.f) - (float)(v < 0.f); }

template <int B>
__global__ void __launch_bounds__(RDG_BLOCK) coeff_reg_kernel(int64_t n, const float* __restrict__ coeff, float g_l1,
                                                              float g_sp, float* __restrict__ d_coeff, int accumulate,
                                                              double* __restrict__ sums) {
    __shared__ double sh[RDG_BLOCK / 32];
    double s_l1 = 0.0, s_sp = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * RDG_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * RDG_BLOCK) {
        float c[B];
        const float4* src = reinterpret_cast<const float4*>(coeff + i * B);
#pragma unroll
        for (int k = 0; k < B / 4; ++k) {
            const float4 v = src[k];
            c[4 * k] = v.x; c[4 * k + 1] = v.y; c[4 * k + 2] = v.z; c[4 * k + 3] = v.w;
        }
        float mx = 0.f, sum = 0.f;
        int arg = 0;
#pragma unroll
        for (int k = 0; k < B; ++k) {
            const float a = fabsf(c[k]);
            sum += a;
            if (a > mx) { mx = a; arg = k; }      // torch.max: first occurrence of the maximum
        }
        const float inv = 1.0f / (mx + 1e-7f);
        s_l1 += (double)sum;
        s_sp += (double)(sum * inv);
        if (d_coeff) {
            // d/da_b [ sum_b a_b / (m + e) ] = 1 / (m + e)  - [b == argmax] * sum_b a_b / (m + e)^2
            const float at_max = -sum * inv * inv;
            float4* dst = reinterpret_cast<float4*>(d_coeff + i * B);
#pragma unroll
            for (int k = 0; k < B / 4; ++k) {
                float g[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int b = 4 * k + e;
                    g[e] = sgn(c[b]) * (g_l1 + g_sp * (inv + (b == arg ? at_max : 0.f)));
                }
                float4 o = accumulate ? dst[k] : make_float4(0.f, 0.f, 0.f, 0.f);
                o.x += g[0]; o.y += g[1]; o.z += g[2]; o.w += g[3];
                dst[k] = o;
            }
        }
    }
    const double a = block_sum_double(s_l1, sh);
    const double b = block_sum_double(s_sp, sh);
    if (threadIdx.x == 0) {
        atomicAdd(sums, a);
        atomicAdd(sums + 1, b);
    }
}

__global__ void coeff_reg_finalize_kernel(const double* __restrict__ sums, double inv_count, float* __restrict__ out) {
    out[0] = (float)(sums[0] * inv_count);
    out[1] = (float)(sums[1] * inv_count);
}

extern "C" int rdg_motion_coeff_reg(int64_t n, int32_t num_basis, const float* coeff, float w_l1, float w_sparsity,
                                    float* loss_parts, float* d_coeff, int32_t accumulate, void* workspace,
                                    void* stream) {
    RDG_CHECK_ARG(n > 0 && coeff && loss_parts && workspace, "null argument or empty model");
    RDG_CHECK_ARG(num_basis == 16, "only num_basis == 16 (the reference's configs) is built");
    RDG_CHECK_ARG((((uintptr_t)coeff | (uintptr_t)d_coeff) & 15) == 0, "coeff / d_coeff must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    double* sums = (double*)workspace;
    RDG_CUDA(cudaMemsetAsync(sums, 0, 2 * sizeof(double), st));
    const double inv_count = 1.0 / ((double)n * num_basis);
    const int64_t want = (n + RDG_BLOCK - 1) / RDG_BLOCK;
    const int grid = (int)(want < (int64_t)RDG_SM_COUNT * 8 ? want : (int64_t)RDG_SM_COUNT * 8);
    coeff_reg_kernel<16><<<grid, RDG_BLOCK, 0, st>>>(n, coeff, (float)(w_l1 * inv_count), (float)(w_sparsity * inv_count),
                                                     d_coeff, accumulate, sums);
    RDG_CHECK_LAUNCH();
    coeff_reg_finalize_kernel<<<1, 1, 0, st>>>(sums, inv_count, loss_parts);
    RDG_CHECK_LAUNCH();
    rdg_count_launches(2);
    return RDG_OK;
}

// ---- basis table ------------------------------------------------------------------------------------------------

struct Quat { float r, i, j, k; };

// graphic_utils.py:76-102: R = I + (2 / |q|^2) U(q), no normalisation beforehand
__device__ __forceinline__ void quat_to_matrix(const Quat q, float R[9]) {
    const float s = 2.0f / (q.r * q.r + q.i * q.i + q.j * q.j + q.k * q.k);
    R[0] = 1.f - s * (q.j * q.j + q.k * q.k); R[1] = s * (q.i * q.j - q.k * q.r); R[2] = s * (q.i * q.k + q.j * q.r);
    R[3] = s * (q.i * q.j + q.k * q.r); R[4] = 1.f - s * (q.i * q.i + q.k * q.k); R[5] = s * (q.j * q.k - q.i * q.r);
    R[6] = s * (q.i * q.k - q.j * q.r); R[7] = s * (q.j * q.k + q.i * q.r); R[8] = 1.f - s * (q.i * q.i + q.j * q.j);
}

// dL/dq from G = dL/dR
__device__ __forceinline__ void quat_to_matrix_bwd(const Quat q, const float G[9], float dq[4]) {
    const float n2 = q.r * q.r + q.i * q.i + q.j * q.j + q.k * q.k;
    const float s = 2.0f / n2;
    // U = (R - I) / s
    const float U[9] = {-(q.j * q.j + q.k * q.k), q.i * q.j - q.k * q.r, q.i * q.k + q.j * q.r,
                        q.i * q.j + q.k * q.r, -(q.i * q.i + q.k * q.k), q.j * q.k - q.i * q.r,
                        q.i * q.k - q.j * q.r, q.j * q.k + q.i * q.r, -(q.i * q.i + q.j * q.j)};
    float ds = 0.f;
#pragma unroll
    for (int m = 0; m < 9; ++m) ds += G[m] * U[m];
    const float dsq = -2.0f * s / n2 * ds;     // ds/dq = -4 q / n2^2 = (-2 s / n2) q
    dq[0] = s * (-q.k * G[1] + q.j * G[2] + q.k * G[3] - q.i * G[5] - q.j * G[6] + q.i * G[7]) + dsq * q.r;
    dq[1] = s * (q.j * (G[1] + G[3]) + q.k * (G[2] + G[6]) - 2.f * q.i * (G[4] + G[8]) + q.r * (G[7] - G[5])) + dsq * q.i;
    dq[2] = s * (-2.f * q.j * (G[0] + G[8]) + q.i * (G[1] + G[3]) + q.r * (G[2] - G[6]) + q.k * (G[5] + G[7])) + dsq * q.j;
    dq[3] = s * (-2.f * q.k * (G[0] + G[4]) + q.r * (G[3] - G[1]) + q.i * (G[2] + G[6]) + q.j * (G[5] + G[7])) + dsq * q.k;
}

// One edge (t0 -> t1) of basis b: value of reg * (|dtr| + ||I - (R1 - R0)||_F) and its gradient w.r.t. row `which`
// (0: the earlier row, 1: the later row).  Returns the value; grad[7] += scale * d(value)/d(row).
__device__ __forceinline__ void basis_edge(const float* __restrict__ r0, const float* __restrict__ r1, float reg, int which,
                                           float scale, float& v_tr, float& v_rot, float grad[7], bool want_grad) {
    const float dx = r1[0] - r0[0], dy = r1[1] - r0[1], dz = r1[2] - r0[2];
    const float nt = sqrtf(dx * dx + dy * dy + dz * dz);
    v_tr = reg * nt;
    const Quat q0 = {r0[3], r0[4], r0[5], r0[6]}, q1 = {r1[3], r1[4], r1[5], r1[6]};
    float R0[9], R1[9], M[9];
    quat_to_matrix(q0, R0);
    quat_to_matrix(q1, R1);
    float f2 = 0.f;
#pragma unroll
    for (int m = 0; m < 9; ++m) {
        M[m] = ((m == 0 || m == 4 || m == 8) ? 1.f : 0.f) - (R1[m] - R0[m]);
        f2 += M[m] * M[m];
    }
    const float f = sqrtf(f2);
    v_rot = reg * f;
    if (!want_grad) return;
    const float sg = which ? 1.f : -1.f;            // d(dtr)/d(row): +1 for the later row, -1 for the earlier one
    if (nt > 0.f) {                                 // torch.norm backward: zero sub-gradient at 0
        const float k = scale * reg * sg / nt;
        grad[0] += k * dx; grad[1] += k * dy; grad[2] += k * dz;
    }
    if (f > 0.f) {
        // dL/dM = M / f ; M = I - R1 + R0  =>  dL/dR1 = -M / f, dL/dR0 = +M / f
        float G[9], dq[4];
        const float k = -sg * scale * reg / f;
#pragma unroll
        for (int m = 0; m < 9; ++m) G[m] = k * M[m];
        quat_to_matrix_bwd(which ? q1 : q0, G, dq);
        grad[3] += dq[0]; grad[4] += dq[1]; grad[5] += dq[2]; grad[6] += dq[3];
    }
}

__global__ void __launch_bounds__(RDG_BLOCK) basis_reg_kernel(int T, int B, const float* __restrict__ table,
                                                              const float* __restrict__ reg_coeff, float w_tr, float w_rot,
                                                              float scale, float* __restrict__ d_table,
                                                              double* __restrict__ sums) {
    __shared__ double sh[RDG_BLOCK / 32];
    const int idx = blockIdx.x * RDG_BLOCK + threadIdx.x;
    double s_tr = 0.0, s_rot = 0.0;
    if (idx < T * B) {
        const int t = idx / B, b = idx % B;
        const float reg = reg_coeff[b];
        const float* row = table + ((int64_t)t * B + b) * 7;
        float grad[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float vt, vr;
        if (t + 1 < T) {   // edge (t, t+1): this thread owns its value; row t is the earlier row
            float gtr[7] = {0, 0, 0, 0, 0, 0, 0};
            basis_edge(row, row + (int64_t)B * 7, reg, 0, scale, vt, vr, gtr, d_table != nullptr);
            s_tr = vt; s_rot = vr;
#pragma unroll
            for (int m = 0; m < 3; ++m) grad[m] += w_tr * gtr[m];
#pragma unroll
            for (int m = 3; m < 7; ++m) grad[m] += w_rot * gtr[m];
        }
        if (t > 0 && d_table) {   // edge (t-1, t): row t is the later row (value owned by thread t-1)
            float gtr[7] = {0, 0, 0, 0, 0, 0, 0};
            basis_edge(row - (int64_t)B * 7, row, reg, 1, scale, vt, vr, gtr, true);
#pragma unroll
            for (int m = 0; m < 3; ++m) grad[m] += w_tr * gtr[m];
#pragma unroll
            for (int m = 3; m < 7; ++m) grad[m] += w_rot * gtr[m];
        }
        if (d_table) {
            float* dst = d_table + ((int64_t)t * B + b) * 7;
#pragma unroll
            for (int m = 0; m < 7; ++m) dst[m] += grad[m];
        }
    }
    const double a = block_sum_double(s_tr, sh);
    const double c = block_sum_double(s_rot, sh);
    if (threadIdx.x == 0) {
        atomicAdd(sums, a);
        atomicAdd(sums + 1, c);
    }
}

extern "C" int rdg_motion_basis_reg(int32_t num_times, int32_t num_basis, const float* table, const float* reg_coeff,
                                    int32_t transl_degree, int32_t rot_degree, float grad_scale, float* loss_parts,
                                    float* d_table, void* workspace, void* stream) {
    RDG_CHECK_ARG(table && reg_coeff && loss_parts && workspace, "null argument");
    RDG_CHECK_ARG(num_times >= 2 && num_basis >= 1, "needs at least two time steps");
    RDG_CHECK_ARG(transl_degree <= 0 && rot_degree <= 0,
                  "only degree 0 (velocity; every reference config) and negative (disabled) degrees are built");
    cudaStream_t st = (cudaStream_t)stream;
    double* sums = (double*)workspace;
    RDG_CUDA(cudaMemsetAsync(sums, 0, 2 * sizeof(double), st));
    const double inv_count = 1.0 / ((double)(num_times - 1) * num_basis);
    const float w_tr = transl_degree < 0 ? 0.f : 1.f, w_rot = rot_degree < 0 ? 0.f : 1.f;
    const int total = num_times * num_basis;
    basis_reg_kernel<<<rdg_div_up(total, RDG_BLOCK), RDG_BLOCK, 0, st>>>(num_times, num_basis, table, reg_coeff, w_tr, w_rot,
                                                                        (float)(grad_scale * inv_count), d_table, sums);
    RDG_CHECK_LAUNCH();
    coeff_reg_finalize_kernel<<<1, 1, 0, st>>>(sums, inv_count, loss_parts);
    RDG_CHECK_LAUNCH();
    rdg_count_launches(2);
    return RDG_OK;
}
