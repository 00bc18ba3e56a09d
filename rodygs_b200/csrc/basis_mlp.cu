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
//   forward  : one CTA (8 warps) per row, the row's activations live in shared memory; a group of n_in/4 lanes owns
//              one output neuron and reads its weight row with coalesced 16-byte loads, 8 row groups in flight per warp
//              before the first shuffle (first version: one thread per neuron walking its row - 27 us, 4 % issue-active).
//   backward : pass A, one CTA per row, back-propagates the row through the layers; the transposed products read the
//              weights along the coalesced direction (4 columns per thread, the row range split over the thread groups,
//              partial sums joined in shared memory; first version: one thread per column, 121 us) and the per-layer
//              activations and pre-activation gradients are stored;
//              pass B, one thread per PARAMETER, contracts those two [rows, .] panels over the rows - a fixed
//              summation order instead of rows x 68 656 float atomics.
#include <math.h>
#include "common.cuh"

#define MLP_BLOCK 256
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

// y[r] = bias[r] + <W[r, :], x> for r < n_out; W row-major [n_out, n_in] in global memory (L2-resident), x in shared
// memory.  Warp-cooperative: a group of G lanes owns one row and reads it with coalesced 16-byte loads (G = n_in / 4
// lanes, at most 32), 32 / G rows per warp instruction, UNROLL row-groups in flight per warp before the first shuffle -
// the whole problem is load latency (a few warps per SM), so independent loads in flight are what matters.  Rows that
// are not 16-byte aligned (the 53-wide first layer) take the scalar variant: 32 lanes per row.  The activation is NOT
// applied here (one lane per row would run erff divergently): the caller does it densely afterwards.
// BLOCKED: block-diagonal layer - row r reads x + (r / xblock) * n_in (the second head layers: basis k has its own
// hidden units); otherwise every row reads the same x.
template <bool BLOCKED>
__device__ __forceinline__ void layer_rows(const float* __restrict__ Wm, const float* __restrict__ bias, const float* x,
                                           int n_in, int n_out, float* y, int xblock) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = MLP_BLOCK / 32;
    const bool vec = ((reinterpret_cast<uintptr_t>(Wm) & 15u) == 0) && (n_in % 4 == 0) && (n_in <= 128) &&
                     ((n_in / 4) & (n_in / 4 - 1)) == 0;
    constexpr int UNROLL = 4;
    if (vec) {
        const int G = n_in / 4;                 // lanes per row (power of two, <= 32)
        const int rpw = 32 / G;                 // rows per warp instruction
        const int sub = lane / G, gl = lane - sub * G;
        for (int r0 = warp * rpw * UNROLL; r0 < n_out; r0 += nwarp * rpw * UNROLL) {
            float acc[UNROLL];
            float4 w[UNROLL];
            float bz[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {        // rows past the end are clamped (loaded, never written): no branches
                const int r = min(r0 + u * rpw + sub, n_out - 1);
                w[u] = __ldg(reinterpret_cast<const float4*>(Wm + (size_t)r * n_in) + gl);
                bz[u] = __ldg(bias + r);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int r = min(r0 + u * rpw + sub, n_out - 1);
                const float4 xv = *reinterpret_cast<const float4*>(x + (BLOCKED ? (r / xblock) * n_in : 0) + 4 * gl);
                acc[u] = (w[u].x * xv.x + w[u].y * xv.y) + (w[u].z * xv.z + w[u].w * xv.w);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                float v = acc[u];
                for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                const int r = r0 + u * rpw + sub;
                if (gl == 0 && r < n_out) y[r] = v + bz[u];
            }
        }
    } else {
        for (int r0 = warp * UNROLL; r0 < n_out; r0 += nwarp * UNROLL) {
            float acc[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int r = min(r0 + u, n_out - 1);
                float a = lane == 0 ? __ldg(bias + r) : 0.f;      // the bias rides in lane 0's partial sum
                const float* xr = x + (BLOCKED ? (r / xblock) * n_in : 0);
                for (int i = lane; i < n_in; i += 32) a = fmaf(__ldg(Wm + (size_t)r * n_in + i), xr[i], a);
                acc[u] = a;
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                float v = acc[u];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                const int r = r0 + u;
                if (lane == 0 && r < n_out) y[r] = v;
            }
        }
    }
}

// y <- act(y) in place (shared memory), the pre-activation kept in z_out (global, may be NULL).  Ends with a barrier.
template <int ACT>
__device__ __forceinline__ void activate(float* y, int n, float* z_out) {
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += MLP_BLOCK) {
        const float z = y[i];
        if (z_out) z_out[i] = z;
        y[i] = act_f<ACT>(z);
    }
    __syncthreads();
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
    layer_rows<false>(P + d.w1, P + d.b1, x, d.E, d.W, a1, 0);
    activate<ACT>(a1, d.W, sv ? sv + d.s1 : nullptr);
    layer_rows<false>(P + d.w2, P + d.b2, a1, d.W, d.W, a2, 0);
    activate<ACT>(a2, d.W, sv ? sv + d.s2 : nullptr);
    layer_rows<false>(P + d.w3, P + d.b3, a2, d.W, d.H, a3, 0);
    activate<ACT>(a3, d.H, sv ? sv + d.s3 : nullptr);
    // the num_basis first head layers share their input: one [nb*Q, H] matrix
    layer_rows<false>(P + d.h0, P + d.hb0, a3, d.H, d.NQ, au, 0);
    activate<ACT>(au, d.NQ, sv ? sv + d.su : nullptr);
    // row 0 (the query time) may have its own destination, so that B(t) and the table land where the trainer keeps them
    float* out_row = basis_row0 ? (m == 0 ? basis_row0 : basis + (size_t)(m - 1) * d.NO) : basis + (size_t)m * d.NO;
    // second head layers: block-diagonal, basis k reads its own Q hidden units; no activation on the output
    layer_rows<true>(P + d.h2, P + d.hb2, au, d.Q, d.NO, out_row, d.O);
}

// out[i] = act'(z[i]) * sum_j W[j, i] g[j]  (i < n_cols, j < n_rows): the transposed product of the backward pass.  The
// columns are the coalesced direction; a thread owns 4 adjacent columns (16-byte loads) of every `parts`-th row, so the
// j range is split over MLP_BLOCK / (n_cols / 4) thread groups and each thread has up to 8 independent loads in
// flight; the partial sums meet in shared memory.  out_s (shared) and out_g (global, the dz panel) both receive it.
template <int ACT>
__device__ __forceinline__ void layer_cols(const float* __restrict__ Wm, const float* g, int n_rows, int n_cols,
                                           const float* __restrict__ z, float* out_s, float* out_g, float4* scratch) {
    const int c4n = n_cols / 4;
    const bool vec = ((reinterpret_cast<uintptr_t>(Wm) & 15u) == 0) && (n_cols % 4 == 0) && c4n <= MLP_BLOCK && (MLP_BLOCK % c4n) == 0;
    if (vec) {
        const int parts = MLP_BLOCK / c4n, part = threadIdx.x / c4n, c4 = threadIdx.x - part * c4n;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* col = reinterpret_cast<const float4*>(Wm) + c4;
        int j = part;
        for (; j + 7 * parts < n_rows; j += 8 * parts) {      // 8 independent 16-byte loads in flight, then the FMAs
            float4 w[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) w[u] = __ldg(col + (size_t)(j + u * parts) * c4n);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float gj = g[j + u * parts];
                acc.x = fmaf(w[u].x, gj, acc.x); acc.y = fmaf(w[u].y, gj, acc.y);
                acc.z = fmaf(w[u].z, gj, acc.z); acc.w = fmaf(w[u].w, gj, acc.w);
            }
        }
        for (; j < n_rows; j += parts) {
            const float4 w = __ldg(col + (size_t)j * c4n);
            const float gj = g[j];
            acc.x = fmaf(w.x, gj, acc.x); acc.y = fmaf(w.y, gj, acc.y); acc.z = fmaf(w.z, gj, acc.z); acc.w = fmaf(w.w, gj, acc.w);
        }
        scratch[threadIdx.x] = acc;
        __syncthreads();
        if (threadIdx.x < c4n) {
            float4 t = scratch[threadIdx.x];
            for (int q = 1; q < parts; ++q) {
                const float4 u = scratch[q * c4n + threadIdx.x];
                t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
            }
            const int i = 4 * threadIdx.x;
            t.x *= act_d<ACT>(z[i]); t.y *= act_d<ACT>(z[i + 1]); t.z *= act_d<ACT>(z[i + 2]); t.w *= act_d<ACT>(z[i + 3]);
            out_s[i] = t.x; out_s[i + 1] = t.y; out_s[i + 2] = t.z; out_s[i + 3] = t.w;
            out_g[i] = t.x; out_g[i + 1] = t.y; out_g[i + 2] = t.z; out_g[i + 3] = t.w;
        }
        __syncthreads();
    } else {
        for (int i = threadIdx.x; i < n_cols; i += MLP_BLOCK) {
            float a = 0.f;
            for (int j = 0; j < n_rows; ++j) a = fmaf(__ldg(Wm + (size_t)j * n_cols + i), g[j], a);
            a *= act_d<ACT>(z[i]);
            out_s[i] = a;
            out_g[i] = a;
        }
        __syncthreads();
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
    __shared__ float dz_a[MLP_MAX_WIDTH];
    __shared__ float dz_b[MLP_MAX_WIDTH];
    __shared__ __align__(16) float4 scratch[MLP_BLOCK];
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
    // heads, second linear (block-diagonal, 7 terms): du[k,i] = act'(u[k,i]) * sum_o H2[k,o,i] dB[k,o]   (coalesced over i)
    for (int r = threadIdx.x; r < d.NQ; r += MLP_BLOCK) {
        const int k = r / d.Q, i = r - k * d.Q;
        float s = 0.f;
        for (int o = 0; o < d.O; ++o) s = fmaf(__ldg(P + d.h2 + (size_t)(k * d.O + o) * d.Q + i), g_out[k * d.O + o], s);
        const float g = s * act_d<ACT>(sv[d.su + r]);
        du[r] = g;
        dz[d.su + r] = g;
    }
    __syncthreads();
    // heads, first linear: dz3[j] = act'(z3[j]) * sum_r H0[r, j] du[r]
    layer_cols<ACT>(P + d.h0, du, d.NQ, d.H, sv + d.s3, dz_a, dz + d.s3, scratch);
    // timenet 3 -> 2: dz2[i] = act'(z2[i]) * sum_j W3[j, i] dz3[j]
    layer_cols<ACT>(P + d.w3, dz_a, d.H, d.W, sv + d.s2, dz_b, dz + d.s2, scratch);
    // timenet 2 -> 1: dz1[i] = act'(z1[i]) * sum_j W2[j, i] dz2[j]
    layer_cols<ACT>(P + d.w2, dz_b, d.W, d.W, sv + d.s1, dz_a, dz + d.s1, scratch);
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
    if (a->rows == 0) return RDG_OK;
    RDG_CHECK_ARG(a->params && (a->basis || (a->basis_row0 && a->rows <= 1)), "params / basis must not be NULL");
    RDG_CHECK_ARG(a->emb || (a->times && a->freqs_pi), "give emb, or times and freqs_pi");
    RDG_CHECK_ARG(a->emb || a->emb_dim % 2 == 1, "computed embedding needs emb_dim = 2 * multires + 1");
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
