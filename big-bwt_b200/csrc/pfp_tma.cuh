// pfp_tma.cuh -- bulk asynchronous copies global -> shared memory (TMA, `cp.async.bulk`, SASS
// UBLKCP) completing on an mbarrier.  One thread arms the barrier with the byte count and issues
// the copies; every thread then waits on the barrier's phase.  No registers, no LDG/STS issue
// slots, and the whole tile is in flight at once.
#pragma once
#include "pfp_common.cuh"

__device__ __forceinline__ u32 smem_addr_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64 *bar, u32 arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(arrivals) : "memory");
}
// make the initialised barrier visible to the async proxy (the copy engine)
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(bytes) : "memory");
}
// bytes % 16 == 0, src and dst 16-byte aligned
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr_u32(dst)), "l"(src), "r"(bytes), "r"(smem_addr_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    u32 ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_addr_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
