// emul_lib_prelude.hpp — force-included (-include) in front of every translation unit of the HOST-EMULATED library
// build (tests/host_emul/build_emul_lib.py): CUDA keywords become plain C++, the fp64 intrinsics plain IEEE operations
// (the build uses -ffp-contract=off), and kernel launches — rewritten from `k<..><<<grid, threads, 0, stream>>>(args)` to
// emul_launch_serial / emul_launch_threaded calls — run synchronously on the calling thread: one emulated CUDA thread after another, or,
// for the two kernels whose threads interact (the shared-memory Lorenz-96 stage kernel and the cooperative device loop,
// as a one-CTA grid), one host thread per CUDA thread with real barriers. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstring>
#include <memory>
#include <thread>
#include <tuple>
#include <vector>

#include "cuda_runtime.h"   // the fake one (tests/host_emul/fake_cuda)

struct EmulDim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local EmulDim3 threadIdx, blockIdx, gridDim, blockDim;

struct EmulBlock {
  std::barrier<> block, cta_end;   // __syncthreads; end of a CTA (the team then plays the next one)
  std::vector<std::unique_ptr<std::barrier<>>> warp;
  explicit EmulBlock(unsigned threads) : block(threads), cta_end(threads) {
    for (unsigned w = 0; w * 32 < threads; ++w) warp.push_back(std::make_unique<std::barrier<>>(std::min(32u, threads - w * 32)));
  }
};
extern EmulBlock* g_emul_block;   // non-null only while a threaded launch is running (defined in emul_lib_support.cpp)
// One emulated "device" per process: function-static __shared__ memory and g_emul_block are process-wide, so when several
// ranks run as threads of one process (multi_rank_emul.py) their kernels take turns.
// A semaphore, not a mutex: a kernel waiting for its peers' partial sums (kernels.cuh: emul_peer_exchange) hands the
// device over from whichever emulated thread does the waiting.
#include <semaphore>
extern std::binary_semaphore g_emul_device;
struct EmulDeviceTurn { EmulDeviceTurn() { g_emul_device.acquire(); } ~EmulDeviceTurn() { g_emul_device.release(); } };
#define B200RK_EMUL_DEVICE_RELEASE() g_emul_device.release()
#define B200RK_EMUL_DEVICE_ACQUIRE() g_emul_device.acquire()

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static

inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline void __syncthreads() { if (g_emul_block) g_emul_block->block.arrive_and_wait(); }
template <class T>
inline T __shfl_down_sync(unsigned, T v, int off) {
  if (!g_emul_block) return v;
  static T lanes[1024];
  const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  lanes[tid] = v;
  g_emul_block->warp[w]->arrive_and_wait();
  const T r = (lane + (unsigned)off < 32u && tid + (unsigned)off < blockDim.x) ? lanes[tid + off] : v;
  g_emul_block->warp[w]->arrive_and_wait();
  return r;
}
template <class T>
inline T __shfl_up_sync(unsigned, T v, int off) {
  if (!g_emul_block) return v;
  static T lanes_up[1024];
  const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  lanes_up[tid] = v;
  g_emul_block->warp[w]->arrive_and_wait();
  const T r = (lane >= (unsigned)off) ? lanes_up[tid - off] : v;
  g_emul_block->warp[w]->arrive_and_wait();
  return r;
}
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T __ldcg(const T* p) { return *p; }
template <class T> inline T __ldg(const T* p) { return *p; }
inline long long clock64() { return 0; }
inline long long __double_as_longlong(double d) { long long r; std::memcpy(&r, &d, 8); return r; }
inline double __longlong_as_double(long long v) { double r; std::memcpy(&r, &v, 8); return r; }
struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }

#define __CUDACC_RTC__ 1               // kernels.cuh: no <cuda_runtime.h> / <stdint.h> of its own
#define B200RK_HOST_EMULATION 1        // kernels.cuh: host form of the inline-PTX access helpers
#define B200RK_EMULATE_SERIAL_SUM 1    // kernels.cuh: grid_sum_finish as a running sum that publishes like the last CTA

template <class F>
inline void emul_launch_serial(unsigned grid, unsigned threads, F&& body) {
  EmulDeviceTurn device;
  for (unsigned b = 0; b < grid; ++b)
    for (unsigned t = 0; t < threads; ++t) {
      gridDim.x = grid; blockDim.x = threads; blockIdx.x = b; threadIdx.x = t;
      body();
    }
}
template <class F>
inline void emul_launch_threaded(unsigned grid, unsigned threads, F&& body) {
  // one team of host threads per launch; it plays the CTAs one after another (function-static "shared memory" is one
  // copy), with a team-wide barrier between CTAs
  EmulDeviceTurn device;
  EmulBlock blk(threads);
  g_emul_block = &blk;
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < threads; ++t)
    pool.emplace_back([&, t] {
      gridDim.x = grid; blockDim.x = threads; threadIdx.x = t;
      for (unsigned b = 0; b < grid; ++b) {
        blockIdx.x = b;
        body();
        blk.cta_end.arrive_and_wait();
      }
    });
  for (auto& th : pool) th.join();
  g_emul_block = nullptr;
}
// cooperative launch of the device-resident loop: a one-CTA grid (fake device: 1 SM, 1 CTA per SM), threaded
template <class A>
inline cudaError_t cudaLaunchCooperativeKernel(void (*kernel)(A), dim3 grid, dim3 block, void** args, size_t, cudaStream_t) {
  if (grid.x != 1) return cudaErrorNotSupported;
  A a = *static_cast<A*>(args[0]);
  emul_launch_threaded(1, block.x, [&] { kernel(a); });
  return cudaSuccess;
}
