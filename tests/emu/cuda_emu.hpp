/* cuda_emu.hpp -- minimal host emulation of the CUDA subset used by hexed_b200/csrc (TEST INFRASTRUCTURE).
 *
 * Lets g++ compile the product's .cu sources unchanged (-DHB_EMULATE -x c++) so the kernels' indexing logic can be
 * checked against the oracle in this GPU-less container. One block runs at a time; each CUDA thread of the block is
 * a std::thread and __syncthreads() is a std::barrier. Shuffles require the whole block to take part.
 * Never loaded by the hexed_b200 package, never timed.
 */
#ifndef HB_CUDA_EMU_HPP_
#define HB_CUDA_EMU_HPP_
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __constant__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)
#define __restrict__
#define __align__(n) alignas(n)

struct dim3 { unsigned x = 1, y = 1, z = 1; dim3() {} dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x{a}, y{b}, z{c} {} };
struct double2 { double x, y; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };

namespace hb_emu {
inline thread_local dim3 tl_threadIdx;
inline dim3 g_blockIdx, g_blockDim, g_gridDim;
inline std::barrier<>* g_barrier = nullptr;
alignas(128) inline unsigned char g_dyn_smem[256*1024];
inline double g_shfl[2048];
inline std::mutex g_atomic_mutex;

template <class F>
void launch(dim3 grid, dim3 block, F&& body)
{
  g_gridDim = grid; g_blockDim = block;
  const unsigned nt = block.x*block.y*block.z;
  for (unsigned bz = 0; bz < grid.z; ++bz) for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; ++bx) {
    g_blockIdx = dim3(bx, by, bz);
    std::barrier<> bar(nt);
    g_barrier = &bar;
    std::vector<std::thread> threads;
    threads.reserve(nt);
    for (unsigned t = 0; t < nt; ++t) {
      threads.emplace_back([&, t] {
        tl_threadIdx = dim3(t % block.x, (t/block.x) % block.y, t/(block.x*block.y));
        body();
      });
    }
    for (auto& th : threads) th.join();
  }
}
} // namespace hb_emu

#define threadIdx (hb_emu::tl_threadIdx)
#define blockIdx (hb_emu::g_blockIdx)
#define blockDim (hb_emu::g_blockDim)
#define gridDim (hb_emu::g_gridDim)

inline void __syncthreads() { hb_emu::g_barrier->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) {}
inline void __threadfence() {}

inline unsigned hb_emu_tid() { return threadIdx.x + blockDim.x*(threadIdx.y + blockDim.y*threadIdx.z); }
template <class T> T hb_emu_exchange(T v, int src_lane_in_warp)
{
  const unsigned tid = hb_emu_tid();
  const unsigned nt = blockDim.x*blockDim.y*blockDim.z;
  static_assert(sizeof(T) <= sizeof(double), "");
  std::memcpy(&hb_emu::g_shfl[tid], &v, sizeof(T));
  __syncthreads();
  const unsigned src = (tid/32)*32 + (unsigned)src_lane_in_warp;
  T r = v;
  if (src_lane_in_warp >= 0 && src_lane_in_warp < 32 && src < nt) std::memcpy(&r, &hb_emu::g_shfl[src], sizeof(T));
  __syncthreads();
  return r;
}
template <class T> T __shfl_xor_sync(unsigned, T v, int mask) { return hb_emu_exchange(v, (int)(hb_emu_tid() % 32) ^ mask); }
template <class T> T __shfl_down_sync(unsigned, T v, int delta) { return hb_emu_exchange(v, (int)(hb_emu_tid() % 32) + delta); }
template <class T> T __shfl_sync(unsigned, T v, int lane) { return hb_emu_exchange(v, lane); }

inline int __any_sync(unsigned, int pred)
{
  const unsigned tid = hb_emu_tid();
  const unsigned nt = blockDim.x*blockDim.y*blockDim.z;
  const double mine = pred ? 1. : 0.;
  std::memcpy(&hb_emu::g_shfl[tid], &mine, sizeof(double));
  __syncthreads();
  int any = 0;
  for (unsigned i = (tid/32)*32; i < (tid/32)*32 + 32 && i < nt; ++i) { double v; std::memcpy(&v, &hb_emu::g_shfl[i], sizeof(double)); any |= v != 0.; }
  __syncthreads();
  return any;
}
inline int __syncthreads_or(int pred)
{
  const unsigned tid = hb_emu_tid();
  const unsigned nt = blockDim.x*blockDim.y*blockDim.z;
  const double mine = pred ? 1. : 0.;
  std::memcpy(&hb_emu::g_shfl[tid], &mine, sizeof(double));
  __syncthreads();
  int any = 0;
  for (unsigned i = 0; i < nt; ++i) { double v; std::memcpy(&v, &hb_emu::g_shfl[i], sizeof(double)); any |= v != 0.; }
  __syncthreads();
  return any;
}
template <class T> T atomicCAS(T* p, T expected, T desired) { std::lock_guard<std::mutex> g(hb_emu::g_atomic_mutex); T o = *p; if (o == expected) *p = desired; return o; }
template <class T> T atomicOr(T* p, T v) { std::lock_guard<std::mutex> g(hb_emu::g_atomic_mutex); T o = *p; *p = o | v; return o; }
template <class T> T atomicAdd(T* p, T v) { std::lock_guard<std::mutex> g(hb_emu::g_atomic_mutex); T o = *p; *p = o + v; return o; }
template <class T> T atomicMin(T* p, T v) { std::lock_guard<std::mutex> g(hb_emu::g_atomic_mutex); T o = *p; if (v < o) *p = v; return o; }
template <class T> T atomicMax(T* p, T v) { std::lock_guard<std::mutex> g(hb_emu::g_atomic_mutex); T o = *p; if (v > o) *p = v; return o; }
inline unsigned atomicInc(unsigned* p, unsigned lim) { std::lock_guard<std::mutex> g(hb_emu::g_atomic_mutex); unsigned o = *p; *p = (o >= lim) ? 0 : o + 1; return o; }

template <class T> T __ldg(const T* p) { return *p; }
template <class T> T __ldca(const T* p) { std::lock_guard<std::mutex> g(hb_emu::g_atomic_mutex); return *p; }
inline float __fsqrt_rn(float x) { return std::sqrt(x); }
inline long long __double_as_longlong(double d) { long long r; std::memcpy(&r, &d, 8); return r; }
inline double __longlong_as_double(long long l) { double r; std::memcpy(&r, &l, 8); return r; }
namespace hb { inline void prefetch_l1(const void*) {} inline void prefetch_l2(const void*) {} }
inline float rsqrtf(float x) { return 1.f/std::sqrt(x); }
inline float __fdividef(float a, float b) { return a/b; }
inline int __float_as_int(float f) { int r; std::memcpy(&r, &f, 4); return r; }
inline float __int_as_float(int i) { float r; std::memcpy(&r, &i, 4); return r; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline double __drcp_rn(double a) { return 1./a; }
using std::fabs; using std::fmax; using std::fmin; using std::sqrt; using std::isfinite;

/* ---- host API subset ---- */
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef struct hb_emu_event { int dummy; }* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDefault = 0, cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaHostAllocDefault = 0,
       cudaFuncAttributePreferredSharedMemoryCarveout = 9, cudaSharedmemCarveoutMaxShared = 100 };
struct cudaDeviceProp { int multiProcessorCount = 148; char name[64] = "host-emulation"; int major = 10, minor = 0; };
inline const char* cudaGetErrorString(cudaError_t) { return "emulated cuda error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { *p = cudaDeviceProp(); return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::aligned_alloc(256, (n + 255)/256*256 + 256); return *p ? cudaSuccess : 1; }
template <class T> cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
template <class T> cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMalloc((void**)p, n); }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = nullptr)
{ for (size_t i = 0; i < h; ++i) std::memcpy((char*)d + i*dp, (const char*)s + i*sp, w); return cudaSuccess; }
inline cudaError_t cudaMemset2DAsync(void* d, size_t dp, int v, size_t w, size_t h, cudaStream_t = nullptr)
{ for (size_t i = 0; i < h; ++i) std::memset((char*)d + i*dp, v, w); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new hb_emu_event(); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new hb_emu_event(); return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0; return cudaSuccess; }
template <class F> cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type = cudaMemoryTypeUnregistered; };
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) { *a = cudaPointerAttributes(); return cudaSuccess; }

struct cudaFuncAttributes { int numRegs = 0; };
inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 1; return cudaSuccess; } // one "SM": persistent kernels iterate
enum { cudaDevAttrMultiProcessorCount = 16 };
template <class F> cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 2; return cudaSuccess; }

/* emulated mbarrier + bulk copy: the copy is done synchronously by the issuing thread; waiters spin on the phase counter */
namespace hb {
struct mbar_t { std::atomic<long long> pending{0}; std::atomic<unsigned> phase{0}; std::atomic<unsigned> arrived{0}; unsigned count = 1; };
inline void mbar_init(mbar_t* bar, unsigned count) { bar->pending = 0; bar->phase = 0; bar->arrived = 0; bar->count = count; }
/* plain arrivals (no transaction bytes): the phase completes when `count` threads have arrived */
inline void mbar_arrive(mbar_t* bar) { if (++bar->arrived == bar->count) { bar->arrived = 0; bar->phase++; } }
inline void mbar_init_fence() {}
inline void mbar_arrive_expect_tx(mbar_t* bar, unsigned bytes) { if ((bar->pending += (long long)bytes) == 0) bar->phase++; } // (nothing expected: the arrival alone completes the phase)
inline void bulk_g2s(void* dst, const void* src, unsigned bytes, mbar_t* bar)
{
  std::memcpy(dst, src, bytes);
  if ((bar->pending -= (long long)bytes) == 0) bar->phase++;
}
inline void mbar_wait(mbar_t* bar, unsigned parity) { while ((bar->phase.load() & 1u) == parity) std::this_thread::yield(); }
inline void fence_proxy_async() {}
}

#define HB_LAUNCH(kern, grid, block, smem, stream, ...) hb_emu::launch(dim3(grid), dim3(block), [&] { kern(__VA_ARGS__); })
#define HB_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(hb_emu::g_dyn_smem)

#endif
