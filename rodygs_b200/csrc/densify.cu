// Densification bookkeeping on the SoA parameter layout: the statistics the
// reference gathers after every backward pass, clone / split / prune as ONE stream compaction,
// and the opacity reset.  Reference (file:line under /root/reference):
//   stats            src/trainer/rodygs.py:316-341, src/trainer/rodygs_static.py:317-319
//   densify_and_prune src/trainer/rodygs_static.py:170-301, src/trainer/rodygs_dynamic.py:150-197
//   Adam-state surgery src/trainer/utils.py:15-95
//   reset_opacity    src/trainer/rodygs_static.py:150-159
//
// The reference runs clone -> (cat every tensor + Adam state) -> split -> (cat again) -> prune the
// split parents (boolean-mask copy of everything) -> prune by opacity / size (another copy): four
// full passes over every parameter and both Adam moments, each through `nn.Parameter` re-creation.
// Here the result of that whole sequence is computed in one go.  Every original Gaussian i gets a
// 4-bit code from the reference's predicates, evaluated on the ORIGINAL tensors (the reference's
// later stages only ever look at rows that are copies of original rows, so their predicates are
// functions of the original row):
//   bit 0  survivor   : not split-selected, not pruned             -> block A (original order)
//   bit 1  clone      : clone-selected and its copy is not pruned  -> block B
//   bit 2  split sel. : split-selected (rank k among them indexes the noise rows k and S + k)
//   bit 3  children   : split-selected and the children are not pruned -> blocks C (copy 0), D (copy 1)
// A three-kernel scan (flags + per-block counts, one-block scan, emit) writes the source map of the
// new rows [A | B | C | D] - exactly the row order the reference ends up with - and one gather
// kernel (a warp per new row) moves every parameter and both Adam moments of that row in a single
// pass: 2 x 300 B per surviving row instead of 4 x (read + write) of everything.
// Bound: HBM (pure row gather); runs every densification_interval (100) iterations.
#include "common.cuh"

#define DN_BLOCK 256
#define DN_SCAN_THREADS 1024

struct DensifyCrit {
    float grad_threshold;   // densify_grad_threshold
    float dense_limit;      // percent_dense * scene_extent
    float min_opacity;      // 0.005
    float ws_limit;         // 0.1 * extent (only when a screen-size threshold is given)
    int use_size;           // max_screen_size is not None
    int scaling_width;      // 3, or 1 for isotropic models (get_scaling repeats the column)
};

__device__ __forceinline__ float dn_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(DN_BLOCK) densify_flag_kernel(int64_t n, const float* __restrict__ scaling,
                                                                const float* __restrict__ opacity,
                                                                const float* __restrict__ grad_accum,
                                                                const float* __restrict__ denom, DensifyCrit c,
                                                                uint8_t* __restrict__ code, int32_t* __restrict__ block_counts) {
    const int64_t i = (int64_t)blockIdx.x * DN_BLOCK + threadIdx.x;
    unsigned bits = 0u;
    if (i < n) {
        // grads = xyz_gradient_accum / denom; grads[isnan] = 0            (rodygs_static.py:286-287)
        float g = grad_accum[i] / denom[i];
        if (isnan(g)) g = 0.0f;
        float ms, ms_child;
        if (c.scaling_width == 1) {
            const float s = expf(scaling[i]);
            ms = s;
            ms_child = expf(logf(s / 1.6f));
        } else {
            const float s0 = expf(scaling[3 * i]), s1 = expf(scaling[3 * i + 1]), s2 = expf(scaling[3 * i + 2]);
            ms = fmaxf(s0, fmaxf(s1, s2));
            // the children store log(scale / (0.8 N)) and are activated again before the prune test
            ms_child = fmaxf(expf(logf(s0 / 1.6f)), fmaxf(expf(logf(s1 / 1.6f)), expf(logf(s2 / 1.6f))));
        }
        const bool hot = g >= c.grad_threshold;
        const bool clone_sel = hot && (ms <= c.dense_limit);               // :247-253
        const bool split_sel = hot && (ms > c.dense_limit);                // :176-181 (clones carry grad 0)
        const bool low_op = dn_sigmoid(opacity[i]) < c.min_opacity;        // :292
        // big_points_vs (max_radii2D > max_screen_size) can never fire: densification_postfix zeroes
        // max_radii2D (:170) before it is read (:294); big_points_ws is live                       (:293-299)
        const bool pruned = low_op || (c.use_size && ms > c.ws_limit);
        const bool pruned_child = low_op || (c.use_size && ms_child > c.ws_limit);
        bits = ((!split_sel && !pruned) ? 1u : 0u) | ((clone_sel && !pruned) ? 2u : 0u) | (split_sel ? 4u : 0u) |
               ((split_sel && !pruned_child) ? 8u : 0u);
        code[i] = (uint8_t)bits;
    }
    const int c0 = __syncthreads_count(bits & 1u), c1 = __syncthreads_count(bits & 2u);
    const int c2 = __syncthreads_count(bits & 4u), c3 = __syncthreads_count(bits & 8u);
    if (threadIdx.x == 0) {
        int32_t* o = block_counts + 4 * (int64_t)blockIdx.x;
        o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
    }
}

// exclusive scan of the per-block counts (4 interleaved sequences), one CTA; totals -> counts[0..3]
__global__ void __launch_bounds__(DN_SCAN_THREADS) densify_scan_kernel(int nb, int32_t* __restrict__ block_counts,
                                                                       int32_t* __restrict__ counts) {
    __shared__ int4 part[DN_SCAN_THREADS];
    const int t = threadIdx.x;
    const int per = (nb + DN_SCAN_THREADS - 1) / DN_SCAN_THREADS;
    const int lo = min(nb, t * per), hi = min(nb, lo + per);
    int4* bc = reinterpret_cast<int4*>(block_counts);
    int4 s = make_int4(0, 0, 0, 0);
    for (int b = lo; b < hi; ++b) {
        const int4 v = bc[b];
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < DN_SCAN_THREADS; o <<= 1) {
        int4 v = make_int4(0, 0, 0, 0);
        if (t >= o) v = part[t - o];
        __syncthreads();
        int4 p = part[t];
        p.x += v.x; p.y += v.y; p.z += v.z; p.w += v.w;
        part[t] = p;
        __syncthreads();
    }
    int4 run = (t == 0) ? make_int4(0, 0, 0, 0) : part[t - 1];
    for (int b = lo; b < hi; ++b) {
        const int4 v = bc[b];
        bc[b] = run;
        run.x += v.x; run.y += v.y; run.z += v.z; run.w += v.w;
    }
    if (t == DN_SCAN_THREADS - 1) {
        const int4 tot = part[t];
        counts[0] = tot.x; counts[1] = tot.y; counts[2] = tot.w; counts[3] = tot.z;   // {A, B, C (kept children), S (split-selected)}
    }
}

__global__ void __launch_bounds__(DN_BLOCK) densify_emit_kernel(int64_t n, const uint8_t* __restrict__ code,
                                                                const int32_t* __restrict__ block_offsets,
                                                                const int32_t* __restrict__ counts, uint32_t* __restrict__ map,
                                                                int32_t* __restrict__ split_rank) {
    __shared__ int wcount[DN_BLOCK / 32][4];
    const int64_t i = (int64_t)blockIdx.x * DN_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bits = (i < n) ? (unsigned)code[i] : 0u;
    const unsigned lt = (1u << lane) - 1u;
    int rank[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned bal = __ballot_sync(0xffffffffu, (bits >> k) & 1u);
        rank[k] = __popc(bal & lt);
        if (lane == 0) wcount[warp][k] = __popc(bal);
    }
    __syncthreads();
    const int32_t* bo = block_offsets + 4 * (int64_t)blockIdx.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int base = bo[k];
        for (int w = 0; w < warp; ++w) base += wcount[w][k];
        rank[k] += base;
    }
    if (i >= n) return;
    const uint32_t nA = (uint32_t)counts[0], nB = (uint32_t)counts[1], nC = (uint32_t)counts[2];
    const uint32_t src = (uint32_t)i;
    if (bits & 1u) map[rank[0]] = src;
    if (bits & 2u) map[nA + rank[1]] = src | (1u << 30);
    if (bits & 8u) {
        map[nA + nB + rank[3]] = src | (2u << 30);
        map[nA + nB + nC + rank[3]] = src | (3u << 30);
    }
    split_rank[i] = (bits & 4u) ? rank[2] : -1;
}

#define DN_MAX_FIELDS 32
struct DensifyFields {
    RdgDensifyField f[DN_MAX_FIELDS];
    int n_fields;
    const float* xyz;        // sources of the split formula
    const float* scaling;
    const float* rotation;
    int scaling_width;
};

// new xyz of a split child: R(q / |q|) (z * scale) + xyz                 (rodygs_static.py:182-195, general_utils.py:92-115)
__device__ __forceinline__ float dn_child_xyz(const DensifyFields& p, uint32_t i, int comp, const float* __restrict__ z) {
    const float* q4 = p.rotation + 4 * (size_t)i;
    float r = q4[0], x = q4[1], y = q4[2], zq = q4[3];
    const float norm = sqrtf(r * r + x * x + y * y + zq * zq);
    r /= norm; x /= norm; y /= norm; zq /= norm;
    float s[3];
    if (p.scaling_width == 1) s[0] = s[1] = s[2] = expf(p.scaling[i]);
    else { s[0] = expf(p.scaling[3 * (size_t)i]); s[1] = expf(p.scaling[3 * (size_t)i + 1]); s[2] = expf(p.scaling[3 * (size_t)i + 2]); }
    const float v0 = z[0] * s[0], v1 = z[1] * s[1], v2 = z[2] * s[2];
    float R0, R1, R2;
    if (comp == 0) { R0 = 1.f - 2.f * (y * y + zq * zq); R1 = 2.f * (x * y - r * zq); R2 = 2.f * (x * zq + r * y); }
    else if (comp == 1) { R0 = 2.f * (x * y + r * zq); R1 = 1.f - 2.f * (x * x + zq * zq); R2 = 2.f * (y * zq - r * x); }
    else { R0 = 2.f * (x * zq - r * y); R1 = 2.f * (y * zq + r * x); R2 = 1.f - 2.f * (x * x + y * y); }
    return R0 * v0 + R1 * v1 + R2 * v2 + p.xyz[3 * (size_t)i + comp];
}

// one warp per new row: every field of the row in one pass
__global__ void __launch_bounds__(DN_BLOCK) densify_gather_kernel(int64_t n_new, const uint32_t* __restrict__ map,
                                                                  const int32_t* __restrict__ split_rank,
                                                                  const int32_t* __restrict__ counts,
                                                                  const float* __restrict__ noise, const DensifyFields p) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * DN_BLOCK) >> 5;
    for (int64_t j = (((int64_t)blockIdx.x * DN_BLOCK + threadIdx.x) >> 5); j < n_new; j += warps) {
        const uint32_t e = map[j];
        const uint32_t i = e & 0x3fffffffu, kind = e >> 30;
        for (int k = 0; k < p.n_fields; ++k) {
            const RdgDensifyField f = p.f[k];
            const int w = f.width;
            const uint32_t* src = reinterpret_cast<const uint32_t*>(f.src) + (size_t)i * w;
            uint32_t* dst = reinterpret_cast<uint32_t*>(f.dst) + (size_t)j * w;
            if (f.mode == RDG_DENSIFY_MOMENT) {
                for (int l = lane; l < w; l += 32) dst[l] = (kind == 0u) ? src[l] : 0u;
            } else if (f.mode == RDG_DENSIFY_XYZ && kind >= 2u) {
                if (lane < 3) {
                    const int S = counts[3];
                    const float* z = noise + 3 * ((size_t)(kind - 2u) * S + split_rank[i]);
                    reinterpret_cast<float*>(dst)[lane] = dn_child_xyz(p, i, lane, z);
                }
            } else if (f.mode == RDG_DENSIFY_SCALING && kind >= 2u) {
                if (lane < w) reinterpret_cast<float*>(dst)[lane] = logf(expf(__uint_as_float(src[lane])) / 1.6f);   // log(scale / (0.8 * 2))
            } else {
                for (int l = lane; l < w; l += 32) dst[l] = src[l];
            }
        }
    }
}

extern "C" int64_t rdg_densify_workspace_bytes(int64_t n) {
    if (n < 0) return -1;
    const int64_t nb = (n + DN_BLOCK - 1) / DN_BLOCK;
    return rdg_align_up(n, 256) + rdg_align_up(16 * (nb + 1), 256);
}

extern "C" int rdg_densify_plan(int64_t n, const float* scaling, int32_t scaling_width, const float* opacity,
                                const float* grad_accum, const float* denom, float grad_threshold, float percent_dense,
                                float extent, float min_opacity, int32_t use_screen_size, uint32_t* map,
                                int32_t* split_rank, int32_t* counts, void* workspace, int64_t workspace_bytes, void* stream) {
    RDG_CHECK_ARG(n > 0 && n < (1ll << 30), "n out of range");
    RDG_CHECK_ARG(scaling && opacity && grad_accum && denom && map && split_rank && counts && workspace, "null argument");
    RDG_CHECK_ARG(scaling_width == 1 || scaling_width == 3, "scaling width must be 1 (isotropic) or 3");
    RDG_CHECK_ARG(workspace_bytes >= rdg_densify_workspace_bytes(n), "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = rdg_div_up(n, DN_BLOCK);
    uint8_t* code = reinterpret_cast<uint8_t*>(workspace);
    int32_t* block_counts = reinterpret_cast<int32_t*>(code + rdg_align_up(n, 256));
    DensifyCrit c;
    c.grad_threshold = grad_threshold;
    c.dense_limit = percent_dense * extent;
    c.min_opacity = min_opacity;
    c.ws_limit = 0.1f * extent;
    c.use_size = use_screen_size != 0;
    c.scaling_width = scaling_width;
    densify_flag_kernel<<<nb, DN_BLOCK, 0, st>>>(n, scaling, opacity, grad_accum, denom, c, code, block_counts);
    RDG_CHECK_LAUNCH();
    densify_scan_kernel<<<1, DN_SCAN_THREADS, 0, st>>>(nb, block_counts, counts);
    RDG_CHECK_LAUNCH();
    densify_emit_kernel<<<nb, DN_BLOCK, 0, st>>>(n, code, block_counts, counts, map, split_rank);
    RDG_CHECK_LAUNCH();
    rdg_count_launches(3);
    return RDG_OK;
}

extern "C" int rdg_densify_apply(int64_t n_new, const uint32_t* map, const int32_t* split_rank, const int32_t* counts,
                                 const float* noise, const RdgDensifyField* fields, int32_t n_fields, const float* xyz,
                                 const float* scaling, int32_t scaling_width, const float* rotation, void* stream) {
    RDG_CHECK_ARG(n_new >= 0, "negative row count");
    RDG_CHECK_ARG(map && split_rank && counts && fields, "null argument");
    RDG_CHECK_ARG(n_fields > 0 && n_fields <= DN_MAX_FIELDS, "1..32 fields");
    if (n_new == 0) return RDG_OK;
    DensifyFields p;
    p.n_fields = n_fields;
    bool needs_split_src = false;
    for (int k = 0; k < n_fields; ++k) {
        p.f[k] = fields[k];
        RDG_CHECK_ARG(fields[k].src && fields[k].dst && fields[k].width > 0, "bad field");
        RDG_CHECK_ARG(fields[k].mode >= RDG_DENSIFY_COPY && fields[k].mode <= RDG_DENSIFY_SCALING, "bad field mode");
        RDG_CHECK_ARG(fields[k].mode != RDG_DENSIFY_XYZ || fields[k].width == 3, "xyz field must have width 3");
        needs_split_src |= fields[k].mode == RDG_DENSIFY_XYZ;
    }
    RDG_CHECK_ARG(!needs_split_src || (xyz && scaling && rotation && noise), "split needs xyz, scaling, rotation and noise");
    RDG_CHECK_ARG(scaling_width == 1 || scaling_width == 3, "scaling width must be 1 (isotropic) or 3");
    p.xyz = xyz; p.scaling = scaling; p.rotation = rotation; p.scaling_width = scaling_width;
    const int64_t blocks_needed = (n_new * 32 + DN_BLOCK - 1) / DN_BLOCK;
    const int grid = (int)(blocks_needed < (int64_t)RDG_SM_COUNT * 16 ? blocks_needed : (int64_t)RDG_SM_COUNT * 16);
    densify_gather_kernel<<<grid, DN_BLOCK, 0, (cudaStream_t)stream>>>(n_new, map, split_rank, counts, noise, p);
    RDG_CHECK_LAUNCH();
    rdg_count_launches(1);
    return RDG_OK;
}

// ---- per-iteration statistics (rodygs.py:316-341, rodygs_static.py:317-319) ----------------------
// For the n Gaussians of the model being trained (rows offset .. offset + n of the concatenated scene):
//   visible = radii > 0;  max_radii2D = max(max_radii2D, radii);  accum += |means2D.grad[:, :2]|;  denom += 1
__global__ void __launch_bounds__(DN_BLOCK) densify_stats_kernel(int64_t n, const int32_t* __restrict__ radii,
                                                                 const float* __restrict__ means2D_grad,
                                                                 float* __restrict__ max_radii2D, float* __restrict__ grad_accum,
                                                                 float* __restrict__ denom) {
    const int64_t i = (int64_t)blockIdx.x * DN_BLOCK + threadIdx.x;
    if (i >= n) return;
    const int r = radii[i];
    if (r <= 0) return;
    max_radii2D[i] = fmaxf(max_radii2D[i], (float)r);
    const float gx = means2D_grad[3 * i], gy = means2D_grad[3 * i + 1];
    grad_accum[i] += sqrtf(gx * gx + gy * gy);
    denom[i] += 1.0f;
}

extern "C" int rdg_densify_stats(int64_t n, const int32_t* radii, const float* means2D_grad, float* max_radii2D,
                                 float* grad_accum, float* denom, void* stream) {
    RDG_CHECK_ARG(n >= 0, "negative n");
    if (n == 0) return RDG_OK;
    RDG_CHECK_ARG(radii && means2D_grad && max_radii2D && grad_accum && denom, "null argument");
    densify_stats_kernel<<<rdg_div_up(n, DN_BLOCK), DN_BLOCK, 0, (cudaStream_t)stream>>>(n, radii, means2D_grad, max_radii2D,
                                                                                         grad_accum, denom);
    RDG_CHECK_LAUNCH();
    rdg_count_launches(1);
    return RDG_OK;
}

// ---- opacity reset (rodygs_static.py:150-159, utils.py:15-32) ----------------------------------------
// opacity <- inverse_sigmoid(min(sigmoid(opacity), cap)); both Adam moments of the group <- 0
__global__ void __launch_bounds__(DN_BLOCK) reset_opacity_kernel(int64_t n, float* __restrict__ opacity, float cap,
                                                                 float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq) {
    const int64_t i = (int64_t)blockIdx.x * DN_BLOCK + threadIdx.x;
    if (i >= n) return;
    const float o = fminf(dn_sigmoid(opacity[i]), cap);
    opacity[i] = logf(o / (1.0f - o));
    if (exp_avg) exp_avg[i] = 0.0f;
    if (exp_avg_sq) exp_avg_sq[i] = 0.0f;
}

extern "C" int rdg_reset_opacity(int64_t n, float* opacity, float cap, float* exp_avg, float* exp_avg_sq, void* stream) {
    RDG_CHECK_ARG(n >= 0, "negative n");
    if (n == 0) return RDG_OK;
    RDG_CHECK_ARG(opacity, "null opacity");
    RDG_CHECK_ARG(cap > 0.0f && cap < 1.0f, "cap must be in (0, 1)");
    reset_opacity_kernel<<<rdg_div_up(n, DN_BLOCK), DN_BLOCK, 0, (cudaStream_t)stream>>>(n, opacity, cap, exp_avg, exp_avg_sq);
    RDG_CHECK_LAUNCH();
    rdg_count_launches(1);
    return RDG_OK;
}
