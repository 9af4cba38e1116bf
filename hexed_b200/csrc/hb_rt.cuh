/* hb_rt.cuh -- thin runtime shim so that the SAME kernel source builds two ways:
 *
 *   nvcc (product):      real CUDA, sm_100a. This is the only thing hexed_b200/ ever loads.
 *   g++ -DHB_EMULATE:    every CUDA thread of one block runs as a host thread with a real barrier behind
 *                        __syncthreads(). Built ONLY by tests/ (tests/emu/) to exercise the indexing logic of
 *                        the kernels in this GPU-less container before GPU minutes are spent. It is a debugging
 *                        aid, not a fallback: the package never loads it and nothing is timed on it.
 */
#ifndef HB_RT_CUH_
#define HB_RT_CUH_

#ifndef HB_EMULATE
#include <cuda_runtime.h>
#define HB_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define HB_DYN_SMEM(type, name) extern __shared__ __align__(128) unsigned char hb_dyn_smem_raw[]; type* name = reinterpret_cast<type*>(hb_dyn_smem_raw)

/* TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier: the element-major layout makes every input of an
 * element one contiguous run, so no tensor map is needed. Sizes and both addresses must be multiples of 16 bytes. */
namespace hb {
typedef unsigned long long mbar_t;
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(mbar_t* bar, unsigned count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(mbar_t* bar, unsigned bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
/* plain arrival (release semantics at CTA scope): the consumer side of a producer / consumer hand-over, no transaction bytes */
__device__ __forceinline__ void mbar_arrive(mbar_t* bar)
{ asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, mbar_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(mbar_t* bar, unsigned parity)
{
  unsigned done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
/* orders earlier generic-proxy accesses to shared memory before later async-proxy (TMA) writes to the same bytes */
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
} // namespace hb
#else
#include "../../tests/emu/cuda_emu.hpp"
#endif

#endif
