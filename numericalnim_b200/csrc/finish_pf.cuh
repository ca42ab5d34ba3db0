// finish_pf.cuh — software-pipelined form of finish_kernel (kernels.cuh). Knob "finish_prefetch", OFF: measured on the B200 in
// round 2 — 90.1-91.8 us against finish_kernel's 88.5 us inside a step (profiles/README.md) — and kept as the documented alternative. finish_kernel is the one kernel of the general pipeline that sits
// below the HBM limit inside a step (profiles/r01_kernel_model.md: 28 us of fp64 pipe vs 73 us of HBM, 88 us measured):
// its persistent loop loads a tile, then runs ~60 dependent fp64 issues per element (one IEEE division each) with no
// load in flight. Here the next tile's NK+1 loads are issued before the current tile's arithmetic, as in
// fused_attempt_kernel. Same arguments, same per-element arithmetic (finish_elem), same per-thread traversal as
// finish_kernel with U = 1, hence the same partial sums bit for bit on the same grid.
#pragma once
#include "kernels.cuh"

namespace b200rk {

template <int NK, int W, bool DIRECT, int YNEW_MODE, int THREADS>
__global__ void __launch_bounds__(THREADS) finish_pf_kernel(const FinishArgs<NK> a) {
  const size_t nvec = a.n / W;
  const size_t stride = (size_t)gridDim.x * THREADS;
  double acc = 0.0;
  size_t v = (size_t)blockIdx.x * THREADS + threadIdx.x;
  Pk<W> yv, kv[NK];
  if (v < nvec) {
    yv = ld_stream<W>(a.y + v * W);
#pragma unroll
    for (int j = 0; j < NK; ++j) kv[j] = ld_stream<W>(a.k[j] + v * W);
  }
  while (v < nvec) {
    const size_t vn = v + stride;
    Pk<W> yn_, kn_[NK];
    if (vn < nvec) {
      yn_ = ld_stream<W>(a.y + vn * W);
#pragma unroll
      for (int j = 0; j < NK; ++j) kn_[j] = ld_stream<W>(a.k[j] + vn * W);
    }
    Pk<W> yo, eo;
#pragma unroll
    for (int e = 0; e < W; ++e) {
      double ke[NK];
#pragma unroll
      for (int j = 0; j < NK; ++j) ke[j] = kv[j].v[e];
      acc = __dadd_rn(acc, finish_elem<NK, DIRECT, YNEW_MODE>(yv.v[e], ke, a, yo.v[e], eo.v[e]));
    }
    if (YNEW_MODE == 1) st_stream<W>(a.ynew_out + v * W, yo);
    if (a.err_out) st_stream<W>(a.err_out + v * W, eo);
    yv = yn_;
#pragma unroll
    for (int j = 0; j < NK; ++j) kv[j] = kn_[j];
    v = vn;
  }
  if (blockIdx.x == 0) {  // ragged tail (n % W elements), as in finish_kernel
    const size_t i = nvec * W + threadIdx.x;
    if (i < a.n) {
      double ke[NK], yn, ee;
#pragma unroll
      for (int j = 0; j < NK; ++j) ke[j] = a.k[j][i];
      acc = __dadd_rn(acc, finish_elem<NK, DIRECT, YNEW_MODE>(a.y[i], ke, a, yn, ee));
      if (YNEW_MODE == 1) a.ynew_out[i] = yn;
      if (a.err_out) a.err_out[i] = ee;
    }
  }
  grid_sum_finish<THREADS>(acc, a.rs);
}

}  // namespace b200rk
