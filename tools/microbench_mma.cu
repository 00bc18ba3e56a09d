// Microbenchmark for the tensor-core reduction of the blend backward pass (mma.sync.m16n8k8 tf32, SASS HMMA.1684.F32.TF32):
//   1. issue rate per SM sub-partition with 1..8 warps per scheduler, independent accumulators
//   2. latency of a dependent accumulate chain
//   3. cost of 6 HMMAs inside a loop of ~90 FP32/ALU instructions (the shape of the backward inner loop)
//   4. numerics of the hi/lo operand split (hi = x & 0xffffe000, lo = x - hi passed as raw fp32)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_mma.bin tools/microbench_mma.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int NACC>
__global__ void __launch_bounds__(256) k_tput(float* out, int iters, long long* cyc) {
    uint32_t a[4] = {__float_as_uint(1.0f), __float_as_uint(2.0f), __float_as_uint(1.0f), __float_as_uint(0.5f)};
    float d[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
    const uint32_t b0 = __float_as_uint(threadIdx.x * 1e-3f), b1 = __float_as_uint(1.0f);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) mma_tf32(d[i], a, b0, b1);
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pk(float lo, float hi) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk(f2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) { f2_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ unsigned alu1(unsigned a, unsigned b) { unsigned r; asm volatile("lop3.b32 %0, %1, %2, %1, 0x96;" : "=r"(r) : "r"(a), "r"(b)); return r; }

// loop body: 40 FFMA2 + 40 LOP3 (+ NMMA HMMAs in three independent chains)
template <int NMMA>
__global__ void __launch_bounds__(128) k_mix(float* out, int iters, long long* cyc) {
    f2_t p[8];
    unsigned u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { p[i] = pk(threadIdx.x * 1e-3f + i, 1.0f); u[i] = threadIdx.x + i; }
    const f2_t m2 = pk(0.999f, 0.999f), c2 = pk(0.001f, 0.001f);
    uint32_t a[4] = {__float_as_uint(1.0f), __float_as_uint(2.0f), __float_as_uint(1.0f), __float_as_uint(0.5f)};
    float d[3][4] = {};
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 5; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], m2, c2);
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = alu1(u[i], u[(i + 1) & 7]);
        }
        float x0, x1;
        unpk(p[0], x0, x1);
#pragma unroll
        for (int i = 0; i < NMMA; ++i) mma_tf32(d[i % 3], a, __float_as_uint(x0), __float_as_uint(x1));
        if (NMMA > 0) p[1] = pk(d[0][0] + d[1][0] + d[2][0], d[0][3]);
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float x, y; unpk(p[i], x, y); s += x + y + (float)u[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + d[0][1] + d[1][2] + d[2][3];
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// numerics: D[f][n] = sum_k A[f][k] * B[k][n] with A small integers and B arbitrary fp32, via hi/lo split
__global__ void k_num(const float* bvals, float* dout, int mode) {
    const int lane = threadIdx.x, g = lane >> 2, kk = lane & 3;
    // A[row][k]: row 0: 1, row 1: k & 3, row 2: k >> 2, row 3: (k&3)^2, row 4: (k&3)*(k>>2), row 5: (k>>2)^2, rows 6/7: k<4 / k>=4
    auto A = [](int row, int k) -> float {
        const int r = row & 7, x = k & 3, y = (k >> 2) + ((row >= 8) ? 2 : 0);
        switch (r) {
            case 0: return 1.f; case 1: return (float)x; case 2: return (float)y; case 3: return (float)(x * x);
            case 4: return (float)(x * y); case 5: return (float)(y * y); case 6: return k < 4 ? 1.f : 0.f; default: return k >= 4 ? 1.f : 0.f;
        }
    };
    uint32_t a[4] = {__float_as_uint(A(g, kk)), __float_as_uint(A(g + 8, kk)), __float_as_uint(A(g, kk + 4)), __float_as_uint(A(g + 8, kk + 4))};
    const float b0 = bvals[kk * 8 + g], b1 = bvals[(kk + 4) * 8 + g];
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    if (mode == 0) {            // raw fp32 bits only (hardware decides what to do with the low 13 bits)
        mma_tf32(d, a, __float_as_uint(b0), __float_as_uint(b1));
    } else {                    // hi/lo split
        const uint32_t h0 = __float_as_uint(b0) & 0xffffe000u, h1 = __float_as_uint(b1) & 0xffffe000u;
        const float l0 = b0 - __uint_as_float(h0), l1 = b1 - __uint_as_float(h1);
        mma_tf32(d, a, h0, h1);
        mma_tf32(d, a, __float_as_uint(l0), __float_as_uint(l1));
    }
    dout[g * 8 + 2 * kk] = d[0];
    dout[g * 8 + 2 * kk + 1] = d[1];
    dout[(g + 8) * 8 + 2 * kk] = d[2];
    dout[(g + 8) * 8 + 2 * kk + 1] = d[3];
}

static float elapsed(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main() {
    float* d; cudaMalloc(&d, 148 * 64 * 256 * 4);
    long long* cyc; cudaMallocManaged(&cyc, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    printf("== HMMA.1684.F32.TF32 issue rate (8 independent accumulators per warp) ==\n");
    for (int wps = 1; wps <= 8; wps *= 2) {   // warps per scheduler: CTAs of 256 threads = 2 warps per SMSP each
        const int ctas_per_sm = (wps + 1) / 2, threads = wps == 1 ? 128 : 256;
        k_tput<8><<<148 * ctas_per_sm, threads>>>(d, 100, cyc);
        cudaEventRecord(e0);
        k_tput<8><<<148 * ctas_per_sm, threads>>>(d, iters, cyc);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        const double n_mma_per_smsp = (double)iters * 8 * wps;
        printf("  %d warps/SMSP: %.3f ms, clock64 %.2f cycles per HMMA per SMSP (wall: %.2f at 1.965 GHz)\n", wps, elapsed(e0, e1),
               (double)cyc[0] / n_mma_per_smsp, elapsed(e0, e1) * 1e-3 * 1.965e9 / n_mma_per_smsp);
    }
    printf("== dependent chain (1 accumulator, 1 warp/SMSP) ==\n");
    k_tput<1><<<148, 128>>>(d, 100, cyc);
    k_tput<1><<<148, 128>>>(d, iters, cyc); cudaDeviceSynchronize();
    printf("  latency %.1f cycles per dependent HMMA\n", (double)cyc[0] / iters);
    printf("== 40 FFMA2 + 40 LOP3 per iteration, with 0 / 6 HMMAs (6 warps per SMSP: 6 CTAs x 128 threads per SM) ==\n");
    float ms0, ms6;
    k_mix<0><<<148 * 6, 128>>>(d, 100, cyc);
    cudaEventRecord(e0); k_mix<0><<<148 * 6, 128>>>(d, iters, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1); ms0 = elapsed(e0, e1);
    const long long c0 = cyc[0];
    k_mix<6><<<148 * 6, 128>>>(d, 100, cyc);
    cudaEventRecord(e0); k_mix<6><<<148 * 6, 128>>>(d, iters, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1); ms6 = elapsed(e0, e1);
    printf("  0 HMMA: %.3f ms (%.1f cycles/iter/warp)   6 HMMA: %.3f ms (%.1f cycles/iter/warp)  -> %.2f issue cycles per HMMA per SMSP\n",
           ms0, (double)c0 / iters, ms6, (double)cyc[0] / iters, (ms6 - ms0) * 1e-3 * 1.965e9 / ((double)iters * 6 * 6));
    printf("== numerics of the operand split ==\n");
    float hb[64], *db, *dd, hd[128];
    srand(1);
    for (int i = 0; i < 64; ++i) hb[i] = ((float)rand() / RAND_MAX - 0.5f) * expf(10.f * ((float)rand() / RAND_MAX - 0.5f));
    cudaMalloc(&db, 256); cudaMalloc(&dd, 512);
    cudaMemcpy(db, hb, 256, cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 2; ++mode) {
        k_num<<<1, 32>>>(db, dd, mode);
        cudaMemcpy(hd, dd, 512, cudaMemcpyDeviceToHost);
        double worst = 0, worst_abs = 0;
        for (int row = 0; row < 16; ++row)
            for (int n = 0; n < 8; ++n) {
                double ref = 0, mag = 0;
                for (int k = 0; k < 8; ++k) {
                    const int r = row & 7, x = k & 3, y = (k >> 2) + ((row >= 8) ? 2 : 0);
                    const double a = r == 0 ? 1 : r == 1 ? x : r == 2 ? y : r == 3 ? x * x : r == 4 ? x * y : r == 5 ? y * y : r == 6 ? (k < 4) : (k >= 4);
                    ref += a * (double)hb[k * 8 + n];
                    mag += fabs(a * (double)hb[k * 8 + n]);
                }
                if (mag > 0) worst = fmax(worst, fabs(hd[row * 8 + n] - ref) / mag);
                worst_abs = fmax(worst_abs, fabs(hd[row * 8 + n] - ref));
            }
        printf("  %s: max |err| / sum|terms| = %.3e\n", mode == 0 ? "raw fp32 operand   " : "hi/lo split (2 MMA)", worst);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
