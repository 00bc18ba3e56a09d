// 1-D bulk asynchronous copies (TMA, cp.async.bulk) with mbarrier completion - sm_90+/sm_100a.
// Used to stage contiguous per-chunk SH rows (46 KB) global -> shared with one instruction
// and to write the dSH rows back shared -> global, keeping many KB in flight per CTA without
// spending registers or LSU issue slots.  SASS: UBLKCP / SYNCS.
#pragma once
#include <stdint.h>

__device__ __forceinline__ uint32_t rdg_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void rdg_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rdg_smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// make prior generic-proxy shared-memory accesses visible to / ordered with the async proxy
__device__ __forceinline__ void rdg_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void rdg_bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    const uint32_t b = rdg_smem_addr(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     rdg_smem_addr(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(b)
                 : "memory");
}

__device__ __forceinline__ void rdg_mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t b = rdg_smem_addr(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(b),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void rdg_bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(rdg_smem_addr(smem_src)),
                 "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// wait until the bulk stores of this thread have finished READING their shared-memory source
__device__ __forceinline__ void rdg_bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// L2 prefetch of a contiguous global range by ONE instruction (no registers, no shared memory, no completion to wait for):
// gmem must be 16-byte aligned and bytes a multiple of 16.  The persistent per-Gaussian kernels issue it for the NEXT chunk's
// parameter rows while the current chunk is computed, so that the dependent scalar loads of the next chunk find their lines
// in L2 (~1/3 of the DRAM latency they were exposed to: ncu r02, 47 % of the warp stalls of preprocess_bwd were long-scoreboard).
__device__ __forceinline__ void rdg_bulk_prefetch_l2(const void* gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem), "r"(bytes) : "memory");
}
