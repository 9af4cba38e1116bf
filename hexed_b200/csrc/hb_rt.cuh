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
#define HB_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char hb_dyn_smem_raw[]; type* name = reinterpret_cast<type*>(hb_dyn_smem_raw)
#else
#include "../../tests/emu/cuda_emu.hpp"
#endif

#endif
