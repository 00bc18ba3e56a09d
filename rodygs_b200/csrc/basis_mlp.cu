// Time embedding + motion-basis MLP, forward and backward.
//
//   rdg_basis_mlp_fwd : TimestepEmbedder.forward                (/root/reference/src/model/rodygs_dynamic.py:202-220)
//                       MLPBasisNetwork.batch_inference / forward (:296-327): timenet E->W->W->W/2, then num_basis
//                       heads W/2 -> W/4 -> 7, for every row of a batch of times (row 0 = the query time t, rows
//                       1..T = the training times, so B(t) and the table B(t_i) come out of ONE launch).
//   rdg_basis_mlp_bwd : the autograd backward of the same, from dL/dB [rows, num_basis, 7] to the packed
//                       parameter gradient.
//
// The reference evaluates the 16 heads in a Python loop (:300-303, :314-317): ~40 launches forward and ~100 backward
// per call, twice per step.  Here: one forward launch, two backward launches, no atomics, deterministic.
//
// B200 mapping.  68 656 parameters (275 KB) and at most a few hundred rows: the work (7 M FMA at T = 100) is three
// orders of magnitude below anything that could load an SM, so this is a LATENCY problem, not a roofline one; the
// parameters stay in L2 (126 MB) and every CTA streams them once.
//   forward  : one CTA per row, the row's activations live in shared memory, one thread per output neuron
//              (float4 row reads when the row is 16-byte aligned).
//   backward : pass A, one CTA per row, back-propagates the row through the layers (transposed weight reads are
//              the coalesced direction here) and stores the per-layer activations and pre-activation gradients;
//              pass B, one thread per PARAMETER, contracts those two [rows, .] panels over the rows - a fixed
//              summation order instead of rows x 68 656 float atomics.
#include <math.h>
#include "common.cuh"

#define MLP_BLOCK 128
#define MLP_MAX_WIDTH 256
#define MLP_MAX_HEADS 2048   // num_basis * (width / 4)
#define MLP_MAX_OUT 256      // num_basis * out_dim

struct MlpDims {
    int E, W, H, Q, nb, O;      // embedding, width, width/2, width/4, bases, outputs per basis (7)
    int NQ, NO;                 // nb*Q, nb*O
    // parameter offsets (floats) in the packed buffer
    int w1, b1, w2, b2, w3, b3, h0, hb0, h2, hb2, total;
    // saved-row offsets (floats): x, z1, z2, z3, u   (pre-activations)
    int sx, s1, s2, s3, su, S;
};

static inline MlpDims mlp_dims(int E, int W, int nb, int O) {
    MlpDims d;
    d.E = E; d.W = W; d.H = W / 2; d.Q = d.H / 2; d.nb = nb; d.O = O;
    d.NQ = nb * d.Q; d.NO = nb * O;
    int o = 0;
    d.w1 = o; o += W * E;
    d.b1 = o; o += W;
    d.w2 = o; o += W * W;
    d.b2 = o; o += W;
    d.w3 = o; o += d.H * W;
    d.b3 = o; o += d.H;
    d.h0 = o; o += nb * d.Q * d.H;
    d.hb0 = o; o += nb * d.Q;
    d.h2 = o; o += nb * O * d.Q;
    d.hb2 = o; o += nb * O;
    d.total = o;
    int s = 0;
    d.sx = s; s += E;
    d.s1 = s; s += W;
    d.s2 = s; s += W;
    d.s3 = s; s += d.H;
    d.su = s; s += d.NQ;
    d.S = s;
    return d;
}

// nn.GELU() (exact, erf form) / nn.ReLU, and their derivatives.
template <int ACT>
__device__ __forceinline__ float act_f(float z) {
    if (ACT == 1) return fmaxf(z, 0.f);
    return 0.5f * z * (1.0f + erff(z * 0.70710678118654752f));
}
template <int ACT>
__device__ __forceinline__ float act_d(float z) {
    if (ACT == 1) return z > 0.f ? 1.f : 0.f;
    return 0.5f * (1.0f + erff(z * 0.70710678118654752f)) + z * 0.3989422804014327f * expf(-0.5f * z * z);
}

// bias + <row, x>, row in global memory (L2-resident parameters), x in shared memory.
__device__ __forceinline__ float dot_row(const float* __restrict__ row, const float* x, int n, float acc) {
    if (((reinterpret_cast<uintptr_t>(row) | (uintptr_t)n * 4u) & 15u) == 0) {
        const float4* r4 = reinterpret_cast<const float4*>(row);
        float a0 = acc, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int i = 0; i < n / 4; ++i) {
            const float4 w = __ldg(r4 + i);
            a0 = fmaf(w.x, x[4 * i], a0);
            a1 = fmaf(w.y, x[4 * i + 1], a1);
            a2 = fmaf(w.z, x[4 * i + 2], a2);
            a3 = fmaf(w.w, x[4 * i + 3], a3);
        }
        return (a0 + a1) + (a2 + a3);
    }
    for (int i = 0; i < n; ++i) acc = fmaf(__ldg(row + i), x[i], acc);
    return acc;
}

// One row of the batch: embedding (from `emb` or from `times` x `freqs_pi`) into x.
__device__ __forceinline__ void load_embedding(const MlpDims& d, int m, const float* __restrict__ emb,
                                               const float* __restrict__ times, const float* __restrict__ freqs_pi,
                                               float* x) {
    if (emb) {
        for (int i = threadIdx.x; i < d.E; i += MLP_BLOCK) x[i] = emb[(size_t)m * d.E + i];
    } else {
        // [t, sin(f0 pi t), cos(f0 pi t), sin(f1 pi t), ...]  (rodygs_dynamic.py:213-218); sinf/cosf take the
        // accurate (Payne-Hanek) path for the large arguments f_k pi t reaches (up to 2^25 pi).
        const float t = times[m];
        for (int i = threadIdx.x; i < d.E; i += MLP_BLOCK) {
            float v = t;
            if (i > 0) {
                const float a = t * freqs_pi[(i - 1) >> 1];
                v = ((i - 1) & 1) ? cosf(a) : sinf(a);
            }
            x[i] = v;
        }
    }
}

template <int ACT>
__global__ void __launch_bounds__(MLP_BLOCK) basis_mlp_fwd_kernel(MlpDims d, const float* __restrict__ P,
                                                                  const float* __restrict__ emb,
                                                                  const float* __restrict__ times,
                                                                  const float* __restrict__ freqs_pi,
                                                                  float* __restrict__ basis, float* __restrict__ basis_row0,
                                                                  float* __restrict__ saved) {
    __shared__ __align__(16) float x[MLP_MAX_WIDTH];
    __shared__ __align__(16) float a1[MLP_MAX_WIDTH];
    __shared__ __align__(16) float a2[MLP_MAX_WIDTH];
    __shared__ __align__(16) float a3[MLP_MAX_WIDTH / 2];
    __shared__ __align__(16) float au[MLP_MAX_HEADS];
    const int m = blockIdx.x;
    float* sv = saved ? saved + (size_t)m * d.S : nullptr;
    load_embedding(d, m, emb, times, freqs_pi, x);
    __syncthreads();
    if (sv) for (int i = threadIdx.x; i < d.E; i += MLP_BLOCK) sv[d.sx + i] = x[i];
    for (int j = threadIdx.x; j < d.W; j += MLP_BLOCK) {
        const float z = dot_row(P + d.w1 + (size_t)j * d.E, x, d.E, P[d.b1 + j]);
        if (sv) sv[d.s1 + j] = z;
        a1[j] = act_f<ACT>(z);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < d.W; j += MLP_BLOCK) {
        const float z = dot_row(P + d.w2 + (size_t)j * d.W, a1, d.W, P[d.b2 + j]);
        if (sv) sv[d.s2 + j] = z;
        a2[j] = act_f<ACT>(z);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < d.H; j += MLP_BLOCK) {
        const float z = dot_row(P + d.w3 + (size_t)j * d.W, a2, d.W, P[d.b3 + j]);
        if (sv) sv[d.s3 + j] = z;
        a3[j] = act_f<ACT>(z);
    }
    __syncthreads();
    for (int r = threadIdx.x; r < d.NQ; r += MLP_BLOCK) {      // r = basis k * Q + hidden unit
        const float z = dot_row(P + d.h0 + (size_t)r * d.H, a3, d.H, P[d.hb0 + r]);
        if (sv) sv[d.su + r] = z;
        au[r] = act_f<ACT>(z);
    }
    __syncthreads();
    // row 0 (the query time) may have its own destination, so that B(t) and the table land where the trainer keeps them
    float* out_row = basis_row0 ? (m == 0 ? basis_row0 : basis + (size_t)(m - 1) * d.NO) : basis + (size_t)m * d.NO;
    for (int r = threadIdx.x; r < d.NO; r += MLP_BLOCK) {      // r = basis k * O + output
        const int k = r / d.O;
        out_row[r] = dot_row(P + d.h2 + (size_t)r * d.Q, au + k * d.Q, d.Q, P[d.hb2 + r]);
    }
}

// Backward pass A: one CTA per row.  ws_act [rows, S] receives the layer INPUT activations (x, a1, a2, a3, au) and
// ws_dz [rows, S] the pre-activation gradients (slot x unused; dz1, dz2, dz3, du), both in the saved-row layout.
template <int ACT>
__global__ void __launch_bounds__(MLP_BLOCK) basis_mlp_bwd_rows_kernel(MlpDims d, const float* __restrict__ P,
                                                                       const float* __restrict__ saved,
                                                                       const float* __restrict__ d_basis,
                                                                       const float* __restrict__ d_basis_row0,
                                                                       float* __restrict__ ws_act,
                                                                       float* __restrict__ ws_dz) {
    __shared__ float g_out[MLP_MAX_OUT];
    __shared__ float du[MLP_MAX_HEADS];
    __shared__ float dzs[MLP_MAX_WIDTH];     // dz of the layer being propagated
    __shared__ float dzn[MLP_MAX_WIDTH];
    __shared__ float part[MLP_BLOCK];
    const int m = blockIdx.x;
    const float* sv = saved + (size_t)m * d.S;
    float* act = ws_act + (size_t)m * d.S;
    float* dz = ws_dz + (size_t)m * d.S;

    const float* g_row = d_basis_row0 ? (m == 0 ? d_basis_row0 : d_basis + (size_t)(m - 1) * d.NO) : d_basis + (size_t)m * d.NO;
    for (int r = threadIdx.x; r < d.NO; r += MLP_BLOCK) g_out[r] = g_row[r];
    // activations for pass B
    for (int i = threadIdx.x; i < d.E; i += MLP_BLOCK) act[d.sx + i] = sv[d.sx + i];
    for (int i = threadIdx.x; i < d.S - d.E; i += MLP_BLOCK) act[d.s1 + i] = act_f<ACT>(sv[d.s1 + i]);
    __syncthreads();
    // heads, second linear: da_u[k,i] = sum_o H2[k,o,i] dB[k,o]   (coalesced over i)
    for (int r = threadIdx.x; r < d.NQ; r += MLP_BLOCK) {
        const int k = r / d.Q, i = r - k * d.Q;
        float s = 0.f;
        for (int o = 0; o < d.O; ++o) s = fmaf(__ldg(P + d.h2 + (size_t)(k * d.O + o) * d.Q + i), g_out[k * d.O + o], s);
        const float g = s * act_d<ACT>(sv[d.su + r]);
        du[r] = g;
        dz[d.su + r] = g;
    }
    __syncthreads();
    // heads, first linear: da3[j] = sum_r H0[r, j] du[r]   (coalesced over j; the r range is split over MLP_BLOCK/H groups)
    {
        const int groups = d.H <= MLP_BLOCK ? MLP_BLOCK / d.H : 1;
        for (int j0 = 0; j0 < d.H; j0 += MLP_BLOCK) {
            const int j = j0 + (groups > 1 ? threadIdx.x % d.H : threadIdx.x);
            const int grp = groups > 1 ? threadIdx.x / d.H : 0;
            float s = 0.f;
            if (j < d.H && grp < groups)
                for (int r = grp; r < d.NQ; r += groups) s = fmaf(__ldg(P + d.h0 + (size_t)r * d.H + j), du[r], s);
            part[threadIdx.x] = s;
            __syncthreads();
            if (grp == 0 && j < d.H) {
                for (int g2 = 1; g2 < groups; ++g2) s += part[g2 * d.H + j];
                const float g = s * act_d<ACT>(sv[d.s3 + j]);
                dzs[j] = g;
                dz[d.s3 + j] = g;
            }
            __syncthreads();
        }
    }
    // timenet layer 3 -> 2: da2[i] = sum_j W3[j, i] dz3[j]
    for (int i = threadIdx.x; i < d.W; i += MLP_BLOCK) {
        float s = 0.f;
        for (int j = 0; j < d.H; ++j) s = fmaf(__ldg(P + d.w3 + (size_t)j * d.W + i), dzs[j], s);
        const float g = s * act_d<ACT>(sv[d.s2 + i]);
        dzn[i] = g;
        dz[d.s2 + i] = g;
    }
    __syncthreads();
    // layer 2 -> 1: da1[i] = sum_j W2[j, i] dz2[j]
    for (int i = threadIdx.x; i < d.W; i += MLP_BLOCK) {
        float s = 0.f;
        for (int j = 0; j < d.W; ++j) s = fmaf(__ldg(P + d.w2 + (size_t)j * d.W + i), dzn[j], s);
        dz[d.s1 + i] = s * act_d<ACT>(sv[d.s1 + i]);
    }
}

// Backward pass B: one thread per parameter; grad[p] (+)= sum_m dz[m, out(p)] * act[m, in(p)].
__global__ void __launch_bounds__(256) basis_mlp_bwd_params_kernel(MlpDims d, int rows, const float* __restrict__ ws_act,
                                                                   const float* __restrict__ ws_dz,
                                                                   const float* __restrict__ d_basis,
                                                                   const float* __restrict__ d_basis_row0,
                                                                   float* __restrict__ d_params, int accumulate) {
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= d.total) return;
    // (panel of output gradients, its row stride, column) x (panel of inputs, column); in_col < 0: bias
    const float* go = ws_dz;
    int go_stride = d.S, go_col, in_col = -1;
    if (p < d.b1)        { const int q = p - d.w1;  go_col = d.s1 + q / d.E; in_col = d.sx + q % d.E; }
    else if (p < d.w2)   { go_col = d.s1 + (p - d.b1); }
    else if (p < d.b2)   { const int q = p - d.w2;  go_col = d.s2 + q / d.W; in_col = d.s1 + q % d.W; }
    else if (p < d.w3)   { go_col = d.s2 + (p - d.b2); }
    else if (p < d.b3)   { const int q = p - d.w3;  go_col = d.s3 + q / d.W; in_col = d.s2 + q % d.W; }
    else if (p < d.h0)   { go_col = d.s3 + (p - d.b3); }
    else if (p < d.hb0)  { const int q = p - d.h0;  go_col = d.su + q / d.H; in_col = d.s3 + q % d.H; }
    else if (p < d.h2)   { go_col = d.su + (p - d.hb0); }
    else if (p < d.hb2)  {
        const int q = p - d.h2, r = q / d.Q;          // r = k*O + o
        go = d_basis; go_stride = d.NO; go_col = r;
        in_col = d.su + (r / d.O) * d.Q + q % d.Q;
    } else               { go = d_basis; go_stride = d.NO; go_col = p - d.hb2; }
    float s = 0.f;
    int m = 0, shift = 0;
    if (go == d_basis && d_basis_row0 && rows > 0) {      // row 0 of the upstream gradient lives in its own buffer
        const float g0 = d_basis_row0[go_col];
        s = in_col >= 0 ? g0 * ws_act[in_col] : g0;
        m = 1;
        shift = 1;
    }
    if (in_col >= 0) {
        for (; m < rows; ++m) s = fmaf(go[(size_t)(m - shift) * go_stride + go_col], ws_act[(size_t)m * d.S + in_col], s);
    } else {
        for (; m < rows; ++m) s += go[(size_t)(m - shift) * go_stride + go_col];
    }
    d_params[p] = accumulate ? d_params[p] + s : s;
}

static int check_dims(const RdgBasisMlp* a, MlpDims* d) {
    RDG_CHECK_ARG(a != nullptr, "null argument struct");
    RDG_CHECK_ARG(a->emb_dim >= 1 && a->emb_dim <= MLP_MAX_WIDTH, "emb_dim out of range (1..256)");
    RDG_CHECK_ARG(a->width >= 4 && a->width <= MLP_MAX_WIDTH && a->width % 4 == 0, "width must be a multiple of 4, <= 256");
    RDG_CHECK_ARG(a->num_basis >= 1 && a->out_dim >= 1, "num_basis / out_dim must be positive");
    RDG_CHECK_ARG(a->num_basis * (a->width / 4) <= MLP_MAX_HEADS, "num_basis * width/4 exceeds 2048");
    RDG_CHECK_ARG(a->num_basis * a->out_dim <= MLP_MAX_OUT, "num_basis * out_dim exceeds 256");
    RDG_CHECK_ARG(a->activation == 0 || a->activation == 1, "activation: 0 = GELU, 1 = ReLU");
    RDG_CHECK_ARG(a->rows >= 0, "negative row count");
    *d = mlp_dims(a->emb_dim, a->width, a->num_basis, a->out_dim);
    return RDG_OK;
}

extern "C" int64_t rdg_basis_mlp_param_count(int32_t emb_dim, int32_t width, int32_t num_basis, int32_t out_dim) {
    if (emb_dim < 1 || width < 4 || width % 4 || num_basis < 1 || out_dim < 1) return -1;
    return mlp_dims(emb_dim, width, num_basis, out_dim).total;
}

extern "C" int64_t rdg_basis_mlp_saved_floats(int32_t emb_dim, int32_t width, int32_t num_basis) {
    if (emb_dim < 1 || width < 4 || width % 4 || num_basis < 1) return -1;
    return mlp_dims(emb_dim, width, num_basis, 1).S;
}

extern "C" int64_t rdg_basis_mlp_bwd_workspace_bytes(int32_t rows, int32_t emb_dim, int32_t width, int32_t num_basis) {
    if (rows < 0 || emb_dim < 1 || width < 4 || width % 4 || num_basis < 1) return -1;
    return 2 * (int64_t)rows * mlp_dims(emb_dim, width, num_basis, 1).S * (int64_t)sizeof(float);
}

extern "C" int rdg_basis_mlp_fwd(const RdgBasisMlp* a, void* stream) {
    MlpDims d;
    if (int rc = check_dims(a, &d)) return rc;
    RDG_CHECK_ARG(a->params && (a->basis || (a->basis_row0 && a->rows <= 1)), "params / basis must not be NULL");
    RDG_CHECK_ARG(a->emb || (a->times && a->freqs_pi), "give emb, or times and freqs_pi");
    RDG_CHECK_ARG(a->emb || a->emb_dim % 2 == 1, "computed embedding needs emb_dim = 2 * multires + 1");
    if (a->rows == 0) return RDG_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (a->activation == 0)
        basis_mlp_fwd_kernel<0><<<a->rows, MLP_BLOCK, 0, s>>>(d, a->params, a->emb, a->times, a->freqs_pi, a->basis, a->basis_row0,
                                                               a->saved);
    else
        basis_mlp_fwd_kernel<1><<<a->rows, MLP_BLOCK, 0, s>>>(d, a->params, a->emb, a->times, a->freqs_pi, a->basis, a->basis_row0,
                                                               a->saved);
    RDG_CHECK_LAUNCH();
    rdg_count_launches(1);
    return RDG_OK;
}

extern "C" int rdg_basis_mlp_bwd(const RdgBasisMlp* a, const float* d_basis, const float* d_basis_row0, float* d_params,
                                 int32_t accumulate,
                                 void* workspace, int64_t workspace_bytes, void* stream) {
    MlpDims d;
    if (int rc = check_dims(a, &d)) return rc;
    RDG_CHECK_ARG(a->params && a->saved && d_params, "params / saved / d_params must not be NULL");
    RDG_CHECK_ARG(d_basis || (d_basis_row0 && a->rows <= 1), "d_basis must not be NULL");
    const int64_t need = 2 * (int64_t)a->rows * d.S * (int64_t)sizeof(float);
    if (workspace_bytes < need || (need > 0 && !workspace)) {
        rdg_set_error("rdg_basis_mlp_bwd: workspace of %lld bytes, need %lld", (long long)workspace_bytes, (long long)need);
        return RDG_E_CAPACITY;
    }
    cudaStream_t s = (cudaStream_t)stream;
    float* ws_act = (float*)workspace;
    float* ws_dz = ws_act + (size_t)a->rows * d.S;
    if (a->rows > 0) {
        if (a->activation == 0)
            basis_mlp_bwd_rows_kernel<0><<<a->rows, MLP_BLOCK, 0, s>>>(d, a->params, a->saved, d_basis, d_basis_row0, ws_act, ws_dz);
        else
            basis_mlp_bwd_rows_kernel<1><<<a->rows, MLP_BLOCK, 0, s>>>(d, a->params, a->saved, d_basis, d_basis_row0, ws_act, ws_dz);
        RDG_CHECK_LAUNCH();
    }
    basis_mlp_bwd_params_kernel<<<rdg_div_up(d.total, 256), 256, 0, s>>>(d, a->rows, ws_act, ws_dz, d_basis, d_basis_row0, d_params,
                                                                           accumulate);
    RDG_CHECK_LAUNCH();
    rdg_count_launches(a->rows > 0 ? 2 : 1);
    return RDG_OK;
}
