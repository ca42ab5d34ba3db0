// Micro-benchmark: the arithmetic of one Tsit54 + Lorenz-96 attempt on registers only (no memory, no exchange): what fp64-pipe
// utilisation does this instruction stream reach with B blocks of 256 threads per SM and E = 4 elements per thread?
#include <cstdio>
#include <cuda_runtime.h>
#include "../../numericalnim_b200/csrc/kernels.cuh"
using namespace b200rk;
template <int PAT, int s, int E, int EXCH, int T>
struct St {
  template <int S>
  __device__ __forceinline__ static void run(const double (&y)[E], double (&k)[E][S], double (&in)[E], const FusedArgs<S>& a, double F, double (*buf)[E * T + 4]) {
    if constexpr (s > 2) St<PAT, s - 1, E, EXCH, T>::run(y, k, in, a, F, buf);
#pragma unroll
    for (int e = 0; e < E; ++e) in[e] = __dadd_rn(y[e], __dmul_rn(const_wsum<S, Pattern<PAT>::a(s - 2)>(k[e], a.a[s - 2]), a.dt));
    if constexpr (EXCH == 0) {
#pragma unroll
      for (int e = 0; e < E; ++e) {   // neighbours taken from the thread's own elements (rotation): same operations, no exchange
        const double m2 = in[(e + E - 2) % E], m1 = in[(e + E - 1) % E], p1 = in[(e + 1) % E];
        k[e][s - 1] = __dadd_rn(__dadd_rn(__dmul_rn(__dadd_rn(p1, -m2), m1), -in[e]), F);
      }
    } else {                          // the real kernel's exchange: pairs at p = 2*(tid + j*256) through shared memory, one barrier per stage
      double* sh = buf[s & 1] + 2;
#pragma unroll
      for (int j = 0; j < E / 2; ++j) { const int p = 2 * (threadIdx.x + j * T); sh[p] = in[2 * j]; sh[p + 1] = in[2 * j + 1]; }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < E / 2; ++j) {
        const int p = 2 * (threadIdx.x + j * T);
        const double m2 = sh[p - 2], m1 = sh[p - 1], p2 = sh[p + 2], c0 = in[2 * j], c1 = in[2 * j + 1];
        k[2 * j][s - 1] = __dadd_rn(__dadd_rn(__dmul_rn(__dadd_rn(c1, -m2), m1), -c0), F);
        k[2 * j + 1][s - 1] = __dadd_rn(__dadd_rn(__dmul_rn(__dadd_rn(p2, -m1), c0), -c1), F);
      }
    }
  }
};
template <int E, int EXCH, int T>
__global__ void __launch_bounds__(T) attempt_regs(const FusedArgs<7> a, double F, int iters, double* out) {
  constexpr int S = 7;
  __shared__ double buf[2][E * T + 4];
  if (threadIdx.x < 4) { buf[0][threadIdx.x < 2 ? threadIdx.x : E * T + threadIdx.x] = 8.0; buf[1][threadIdx.x < 2 ? threadIdx.x : E * T + threadIdx.x] = 8.0; }
  double y[E], k[E][S], in[E];
#pragma unroll
  for (int e = 0; e < E; ++e) { y[e] = 8.0 + 1e-3 * (threadIdx.x + e); k[e][0] = 0.1 * e; for (int j = 1; j < S; ++j) k[e][j] = 0.0; }
  double acc = 0.0;
  for (int it = 0; it < iters; ++it) {
    St<PAT_TSIT54, S, E, EXCH, T>::run(y, k, in, a, F, buf);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const double lo = __dmul_rn(const_wsum<S, Pattern<PAT_TSIT54>::bh()>(k[e], a.bh), a.cbh);
      const double tol = __dadd_rn(a.absTol, __dmul_rn(fabs(in[e]), a.relTol));
      const double r = err_ratio(lo, tol);
      acc = __dadd_rn(acc, __dmul_rn(r, r));
      y[e] = __dadd_rn(__dmul_rn(in[e], 0.5), 4.0);   // keep values bounded and dependent (2 extra fp64 ops per element)
      k[e][0] = k[e][S - 1];
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + y[0];
}
template <int E, int EXCH = 0, int T = 256>
void run(int blocks_per_sm) {
  FusedArgs<7> a;
  memset(&a, 0, sizeof(a));
  for (int s = 0; s < 6; ++s) for (int j = 0; j < 6; ++j) a.a[s][j] = 0.01 * (1 + s + j);
  for (int j = 0; j < 7; ++j) a.bh[j] = 0.001 * (j + 1);
  a.dt = 1e-3; a.cbh = 1e-3; a.absTol = 1e-6; a.relTol = 1e-6;
  double* out; cudaMalloc(&out, 148 * 8 * 256 * 8);
  const int iters = 2000, grid = 148 * blocks_per_sm;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  attempt_regs<E, EXCH, T><<<grid, T>>>(a, 8.0, iters, out);
  cudaEventRecord(e0);
  attempt_regs<E, EXCH, T><<<grid, T>>>(a, 8.0, iters, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fp64_per_elem = 12 + 36 + 24 + 14 + 2 + 6 + 2 + 2;   // stage inputs, rows, stencil, bh row, tol, ratio, r*r + acc, carry
  const double warp_instr = fp64_per_elem * E * iters * (grid * (T / 32.0));
  const double pipe_s = warp_instr * 2.0 / (148 * 4) / 1.965e9;
  printf("exchange=%d E=%d threads=%d blocks/SM=%d: %.3f ms; fp64 pipe time %.3f ms -> utilisation %.2f\n", E, blocks_per_sm, ms, pipe_s * 1e3, pipe_s * 1e3 / ms);
}

// The same arithmetic + exchange, with the real kernel's memory phases: per tile read y and k1 (4 elements per thread), write yNew and k_S.
// MEM = 1: loads at the tile top; MEM = 2: loads of the next tile issued before the current tile's stages (registers).
template <int T, int MEM, int ST = 1>
__global__ void __launch_bounds__(T) attempt_mem(const FusedArgs<7> a, double F, const double* __restrict__ yin, const double* __restrict__ kin,
                                                 double* __restrict__ yout, double* __restrict__ kout, size_t n, double* out) {
  constexpr int S = 7, E = 4, TW = E * T;
  __shared__ double buf[2][TW + 4];
  if (threadIdx.x < 4) { buf[0][threadIdx.x < 2 ? threadIdx.x : TW + threadIdx.x] = 8.0; buf[1][threadIdx.x < 2 ? threadIdx.x : TW + threadIdx.x] = 8.0; }
  double acc = 0.0;
  const size_t n_tiles = n / TW;
  Pk<2> ny[2], nk[2];
  size_t tile = blockIdx.x;
  if (MEM == 2 && tile < n_tiles)
    for (int j = 0; j < 2; ++j) { const size_t g = tile * TW + 2 * (threadIdx.x + j * T); ny[j] = ld_stream<2>(yin + g); nk[j] = ld_stream<2>(kin + g); }
  for (; tile < n_tiles; tile += gridDim.x) {
    double y[E], k[E][S], in[E];
    for (int j = 0; j < 2; ++j) {
      const size_t g = tile * TW + 2 * (threadIdx.x + j * T);
      Pk<2> yv, kv;
      if (MEM == 2) { yv = ny[j]; kv = nk[j]; } else { yv = ld_stream<2>(yin + g); kv = ld_stream<2>(kin + g); }
      y[2 * j] = yv.v[0]; y[2 * j + 1] = yv.v[1]; k[2 * j][0] = kv.v[0]; k[2 * j + 1][0] = kv.v[1];
    }
    if (MEM == 2 && tile + gridDim.x < n_tiles)
      for (int j = 0; j < 2; ++j) { const size_t g = (tile + gridDim.x) * TW + 2 * (threadIdx.x + j * T); ny[j] = ld_stream<2>(yin + g); nk[j] = ld_stream<2>(kin + g); }
#pragma unroll
    for (int e = 0; e < E; ++e) for (int j = 1; j < S; ++j) k[e][j] = 0.0;
    St<PAT_TSIT54, S, E, 1, T>::run(y, k, in, a, F, buf);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const size_t g = tile * TW + 2 * (threadIdx.x + j * T);
      Pk<2> yo, ko;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int e = 2 * j + h;
        const double lo = __dmul_rn(const_wsum<S, Pattern<PAT_TSIT54>::bh()>(k[e], a.bh), a.cbh);
        const double tol = __dadd_rn(a.absTol, __dmul_rn(fabs(in[e]), a.relTol));
        const double r = err_ratio(lo, tol);
        acc = __dadd_rn(acc, __dmul_rn(r, r));
        yo.v[h] = in[e]; ko.v[h] = k[e][S - 1];
      }
      if (ST == 1) { st_stream<2>(yout + g, yo); st_stream<2>(kout + g, ko); }
      else if (ST == 0) { acc = __dadd_rn(acc, yo.v[0] + ko.v[1]); }
      else { Pk<4> o4; o4.v[0] = yo.v[0]; o4.v[1] = yo.v[1]; o4.v[2] = ko.v[0]; o4.v[3] = ko.v[1]; st_stream<4>(yout + 2 * g, o4); }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int T, int MEM, int ST = 1>
void run_mem(int blocks_per_sm) {
  FusedArgs<7> a;
  memset(&a, 0, sizeof(a));
  for (int s = 0; s < 6; ++s) for (int j = 0; j < 6; ++j) a.a[s][j] = 0.01 * (1 + s + j);
  for (int j = 0; j < 7; ++j) a.bh[j] = 0.001 * (j + 1);
  a.dt = 1e-3; a.cbh = 1e-3; a.absTol = 1e-6; a.relTol = 1e-6;
  const size_t n = (size_t)1 << 24;
  double *yin, *kin, *yout, *kout, *out;
  cudaMalloc(&yin, n * 8); cudaMalloc(&kin, n * 8); cudaMalloc(&yout, 2 * n * 8); cudaMalloc(&kout, n * 8); cudaMalloc(&out, 148 * 16 * 512 * 8);
  cudaMemset(yin, 0, n * 8); cudaMemset(kin, 0, n * 8);
  const int grid = 148 * blocks_per_sm;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  attempt_mem<T, MEM, ST><<<grid, T>>>(a, 8.0, yin, kin, yout, kout, n, out);
  cudaEventRecord(e0);
  for (int r = 0; r < 10; ++r) attempt_mem<T, MEM, ST><<<grid, T>>>(a, 8.0, yin, kin, yout, kout, n, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
  const double warp_instr = (12 + 36 + 24 + 14 + 2 + 6 + 2) * (double)n / 32.0;
  const double pipe_ms = warp_instr * 2.0 / (148 * 4) / 1.965e9 * 1e3;
  printf("memory phases ST=%d MEM=%d threads=%d blocks/SM=%d: %.1f us per attempt at 2^24; fp64 pipe %.1f us -> utilisation %.2f; %.0f GB/s\n", ST, MEM, T, blocks_per_sm,
         ms * 1e3, pipe_ms * 1e3, pipe_ms / ms, 4.0 * n * 8 / (ms * 1e-3) / 1e9);
  cudaFree(yin); cudaFree(kin); cudaFree(yout); cudaFree(kout); cudaFree(out);
}
int main() {
  run_mem<256, 2, 1>(3);
  run_mem<256, 2, 0>(3);
  run_mem<256, 2, 2>(3);
  run_mem<128, 2, 0>(4);
  run_mem<128, 2, 1>(4);
  return 0;
}
