// emul.hpp — compile the product's CUDA kernels (numericalnim_b200/csrc/kernels.cuh, quad_kernels.cuh) with the HOST
// compiler and run them one emulated thread at a time, so the CPU test-suite can compare the kernels' arithmetic and
// indexing with the oracle bit for bit without a GPU. TEST INFRASTRUCTURE ONLY.
//
// What is emulated: thread/block indices, the fp64 intrinsics (plain IEEE operations: the translation unit is built
// with -ffp-contract=off, as the device code is built with -fmad=false), global loads/stores (kernels.cuh gives its
// inline-PTX access helpers a host form under B200RK_HOST_EMULATION) and the grid-wide sum (a running sum in
// emulation order). What is not: anything that needs threads to run concurrently (block reductions, the shared-memory
// Lorenz-96 tile, the cooperative device loop) — those kernels are parsed but only the GPU suite runs them.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

// the oracle first, before any CUDA keyword is turned into a macro
#include "quad_oracle.hpp"
#include "rk_oracle.hpp"

// Two modes. Default: ONE host thread plays every CUDA thread in turn (fast; kernels whose threads do not interact).
// EMUL_MT: one host thread per CUDA thread of a block, blocks one after another; __syncthreads / warp shuffles /
// atomics are real synchronisation, so block reductions, shared-memory tiles and (as a one-block grid) the cooperative
// loop run too — and ThreadSanitizer can look for data races between the emulated threads.
struct EmulDim3 { unsigned x = 0, y = 0, z = 0; };
#ifdef EMUL_MT
#include <barrier>
#include <memory>
#include <thread>
static thread_local EmulDim3 threadIdx, blockIdx, gridDim, blockDim;
struct EmulBlock {
  std::barrier<> block, cta_end;   // __syncthreads; end of a CTA (the team then plays the next one)
  std::vector<std::unique_ptr<std::barrier<>>> warp;
  explicit EmulBlock(unsigned threads) : block(threads), cta_end(threads) {
    for (unsigned w = 0; w * 32 < threads; ++w) warp.push_back(std::make_unique<std::barrier<>>(std::min(32u, threads - w * 32)));
  }
};
static EmulBlock* g_emul_block = nullptr;
inline void __syncthreads() { g_emul_block->block.arrive_and_wait(); }
void emul_grid_sync() { g_emul_block->block.arrive_and_wait(); }   // one-block grids only
template <class T>
inline T __shfl_down_sync(unsigned, T v, int off) {
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
#else
static EmulDim3 threadIdx, blockIdx, gridDim, blockDim;
void emul_grid_sync() {}
template <class T> inline T __shfl_down_sync(unsigned, T v, int) { return v; }
template <class T> inline T __shfl_up_sync(unsigned, T v, int) { return v; }
inline void __syncthreads() {}
inline void __threadfence() {}
inline void __threadfence_system() {}
inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned o = *p; *p += v; return o; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p += v; return o; }
inline int atomicExch(int* p, int v) { const int o = *p; *p = v; return o; }
#define B200RK_EMULATE_SERIAL_SUM 1   // kernels.cuh: grid_sum_finish becomes a running sum
#endif

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static

inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
template <class T> inline T __ldcg(const T* p) { return *p; }
template <class T> inline T __ldg(const T* p) { return *p; }
inline long long clock64() { return 0; }
inline long long __double_as_longlong(double d) { long long r; std::memcpy(&r, &d, 8); return r; }
inline double __longlong_as_double(long long v) { double r; std::memcpy(&r, &v, 8); return r; }
struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }

#define __CUDACC_RTC__ 1           // kernels.cuh: no <cuda_runtime.h>
#define B200RK_HOST_EMULATION 1

// a right-hand side "given as source", as jit.cu would generate it (PW_USER instantiations)
#define B200RK_JIT 1
#define B200RK_USER_NP 1
namespace b200rk {
inline double user_rhs(double t, double y, const double* p, const double* c) { return c[0] * y * (1.0 - y / p[0]) + c[1] * t; }
}

#include "kernels.cuh"
#include "finish_pf.cuh"
#include "stencil_attempt.cuh"
#include "quad_kernels.cuh"
#include "methods.h"

// run `body` once per emulated thread of a (grid x threads) launch
template <class F>
inline void emul_launch(unsigned grid, unsigned threads, F&& body) {
#ifdef EMUL_MT
  // one team of host threads per launch: blocks one after another (team-wide barrier in between), the threads of a
  // block concurrently
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
#else
  gridDim.x = grid; blockDim.x = threads;
  for (unsigned b = 0; b < grid; ++b)
    for (unsigned t = 0; t < threads; ++t) { blockIdx.x = b; threadIdx.x = t; body(); }
#endif
}
