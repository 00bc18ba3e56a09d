// Microbenchmark: does FFMA2 (fma.rn.f32x2) free issue slots on sm_100a?  Four loops with the same
// FP32 work per thread: scalar FFMA, packed FFMA2, and each of them interleaved with integer ALU work.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_ffma2.bin tools/microbench_ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pk(float lo, float hi) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk(f2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) { f2_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ unsigned alu1(unsigned a, unsigned b) { unsigned r; asm volatile("lop3.b32 %0, %1, %2, %1, 0x96;" : "=r"(r) : "r"(a), "r"(b)); return r; }

template <int MODE>   // 0 scalar, 1 packed, 2 scalar + ALU, 3 packed + ALU
__global__ void __launch_bounds__(256) k(float* out, int iters, float m, float c) {
    float x[16];
    unsigned u[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) u[i] = threadIdx.x + i;
    f2_t p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = pk(x[2 * i], x[2 * i + 1]);
    const f2_t m2 = pk(m, m), c2 = pk(c, c);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fma1(x[i], m, c);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], m2, c2);
        }
        if (MODE >= 2) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) u[i] = alu1(u[i], u[(i + 1) & 7]);
        }
    }
    float s = 0.f;
    if (MODE == 1 || MODE == 3) {
#pragma unroll
        for (int i = 0; i < 8; ++i) unpk(p[i], x[2 * i], x[2 * i + 1]);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += (float)u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void rcp_check(float* out) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(1.0f + 0.0f * threadIdx.x));
    out[0] = y;
}

template <int MODE>
static void run(const char* name, float* d) {
    const int grid = 148 * 8, iters = 20000;
    k<MODE><<<grid, 256>>>(d, 100, 0.999f, 0.001f);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<grid, 256>>>(d, iters, 0.999f, 0.001f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double fma = (double)grid * 256 * iters * 16;
    printf("%-22s %8.3f ms  %7.2f TFMA/s  (%.1f FMA lanes/clk/SM at 1.965 GHz)\n", name, ms, fma / ms * 1e-9,
           fma / (ms * 1e-3) / 148 / 1.965e9);
}

int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("16 FFMA", d);
    run<1>("8 FFMA2", d);
    run<2>("16 FFMA + 16 LOP3", d);
    run<3>("8 FFMA2 + 16 LOP3", d);
    rcp_check<<<1, 1>>>(d);
    float h; cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
    printf("rcp.approx.ftz(1.0f) == 1.0f : %s (%.9g)\n", h == 1.0f ? "yes" : "NO", h);
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
