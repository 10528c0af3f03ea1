// cp.async.bulk (TMA 1-D bulk copy) + mbarrier helpers shared by the kernels that stage their input tiles in shared memory.
// SASS: UBLKCP.S.G + SYNCS.*.  Compiled out on the emulated device (tools/emu), whose builds never take the TMA paths.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200q {

#ifndef B200Q_EMULATED_DEVICE
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
#else
__device__ __forceinline__ void mbar_init(unsigned long long*, unsigned) {}
__device__ __forceinline__ void mbar_fence_init() {}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long*, unsigned) {}
__device__ __forceinline__ void bulk_g2s(void*, const void*, unsigned, unsigned long long*) {}
__device__ __forceinline__ void mbar_wait(unsigned long long*, unsigned) {}
#endif

}  // namespace b200q
