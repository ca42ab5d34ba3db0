// quad_kernels.cuh — kernels for the consumers of the solver's output on the other side of the hot path
// (SURVEY.md §8f rank 4): Hermite interpolation of a trajectory and cumulative quadrature over it.
//
// What each kernel replaces in the reference (numericalnim, src/numericalnim/), T = Vector[float]:
//   cumtrapz_kernel       the loop of cumtrapz(Y, X)       integrate.nim:131-135   (3 allocating Vector ops per point)
//   simpson_scan_kernel   the loops of cumsimpson(Y, X)    integrate.nim:354-376   (6 allocating Vector ops per pair)
//   hermite_many_kernel   hermiteInterpolate               utils.nim:282-312       (7 allocating Vector ops per sample)
//   neq_count_kernel      `!=` of removeDuplicates         utils.nim:371-373, 51-55
//
// A trajectory is a LIST of device vectors (one per time point, as solveODE returns it), so the kernels take
// pointer tables and per-step scalar tables from device memory. The time direction is a sequential recurrence,
// the state direction is embarrassingly parallel: every thread owns W consecutive state elements and walks the
// whole time axis with the running integral in registers — each input vector is read once and each output
// vector written once (2 passes per time point instead of the reference's ~12), with the loads of the next
// UNROLL time points in flight while the current ones are consumed.
//
// Parity: __dmul_rn / __dadd_rn in the reference's association (scalar * Vector multiplies the component BY the
// scalar; sums left to right), so results are bit-identical to the CPU path.
#pragma once
#include "kernels.cuh"

namespace b200rk {

// ---------------------------------------------------------------------------------------------------
// cumtrapz(Y, X): out[0] = y0 - y0;  I_{k+1} = I_k + (y_{k+1} + y_k) * h_k,  h_k = 0.5*(x_{k+1} - x_k)
// ---------------------------------------------------------------------------------------------------
struct CumTrapzArgs {
  const double* const* y;   // m pointers (device table)
  double* const* out;       // m pointers
  const double* h;          // m-1 scalars
  int m;
  size_t n;
};

template <int W, int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS) cumtrapz_kernel(const CumTrapzArgs a) {
  const size_t nvec = a.n / W;
  for (size_t v = (size_t)blockIdx.x * THREADS + threadIdx.x; v < nvec; v += (size_t)gridDim.x * THREADS) {
    Pk<W> yk = ld_stream<W>(a.y[0] + v * W), I;
#pragma unroll
    for (int e = 0; e < W; ++e) I.v[e] = __dadd_rn(yk.v[e], -yk.v[e]);  // "the right kind of zero" (integrate.nim:129)
    st_stream<W>(a.out[0] + v * W, I);
    int k = 0;
    for (; k + UNROLL < a.m; k += UNROLL) {
      Pk<W> nx[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) nx[u] = ld_stream<W>(a.y[k + 1 + u] + v * W);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const double h = __ldg(a.h + k + u);
#pragma unroll
        for (int e = 0; e < W; ++e) I.v[e] = __dadd_rn(I.v[e], __dmul_rn(__dadd_rn(nx[u].v[e], yk.v[e]), h));
        st_stream<W>(a.out[k + 1 + u] + v * W, I);
        yk = nx[u];
      }
    }
    for (; k + 1 < a.m; ++k) {
      const Pk<W> nx = ld_stream<W>(a.y[k + 1] + v * W);
      const double h = __ldg(a.h + k);
#pragma unroll
      for (int e = 0; e < W; ++e) I.v[e] = __dadd_rn(I.v[e], __dmul_rn(__dadd_rn(nx.v[e], yk.v[e]), h));
      st_stream<W>(a.out[k + 1] + v * W, I);
      yk = nx;
    }
  }
  if (blockIdx.x == 0) {  // ragged tail (n % W elements)
    const size_t i = nvec * W + threadIdx.x;
    if (i < a.n) {
      double yk = a.y[0][i], I = __dadd_rn(yk, -yk);
      a.out[0][i] = I;
      for (int k = 0; k + 1 < a.m; ++k) {
        const double nx = a.y[k + 1][i];
        I = __dadd_rn(I, __dmul_rn(__dadd_rn(nx, yk), a.h[k]));
        a.out[k + 1][i] = I;
        yk = nx;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Simpson scan over unequal intervals (integrate.nim:354-376): node_out[0] = y_first - y_first;
//   I_{j+1} = I_j + ((y[ia]*ca + y[ib]*cb) + y[ic]*cc)
// regular pair i: (ia, ib, ic) = (2i+2, 2i+1, 2i), (ca, cb, cc) = (alpha, beta, eta);
// last interval of an even-length data set: (last-2, last-1, last), (eta, beta, alpha).
// ---------------------------------------------------------------------------------------------------
struct SimpsonStep {
  int ia, ib, ic;
  int reuse;  // 1: y[ic] is the previous step's y[ia] (kept in registers)
  double ca, cb, cc;
};
struct SimpsonScanArgs {
  const double* const* y;
  double* const* node_out;  // steps + 1 pointers
  const SimpsonStep* step;
  int steps;
  int first;                // index of the first data point (zero of the right kind)
  size_t n;
};

template <int W, int THREADS>
__global__ void __launch_bounds__(THREADS) simpson_scan_kernel(const SimpsonScanArgs a) {
  const size_t nvec = a.n / W;
  for (size_t v = (size_t)blockIdx.x * THREADS + threadIdx.x; v < nvec; v += (size_t)gridDim.x * THREADS) {
    Pk<W> prev_a = ld_stream<W>(a.y[a.first] + v * W), I;
#pragma unroll
    for (int e = 0; e < W; ++e) I.v[e] = __dadd_rn(prev_a.v[e], -prev_a.v[e]);
    st_stream<W>(a.node_out[0] + v * W, I);
    for (int j = 0; j < a.steps; ++j) {
      const SimpsonStep s = a.step[j];  // uniform: one broadcast read per step
      const Pk<W> ya = ld_stream<W>(a.y[s.ia] + v * W);
      const Pk<W> yb = ld_stream<W>(a.y[s.ib] + v * W);
      Pk<W> yc;
      if (s.reuse) yc = prev_a;
      else yc = ld_stream<W>(a.y[s.ic] + v * W);
#pragma unroll
      for (int e = 0; e < W; ++e) {
        const double t = __dadd_rn(__dadd_rn(__dmul_rn(ya.v[e], s.ca), __dmul_rn(yb.v[e], s.cb)), __dmul_rn(yc.v[e], s.cc));
        I.v[e] = __dadd_rn(I.v[e], t);
      }
      st_stream<W>(a.node_out[j + 1] + v * W, I);
      prev_a = ya;
    }
  }
  if (blockIdx.x == 0) {
    const size_t i = nvec * W + threadIdx.x;
    if (i < a.n) {
      const double y0 = a.y[a.first][i];
      double I = __dadd_rn(y0, -y0);
      a.node_out[0][i] = I;
      for (int j = 0; j < a.steps; ++j) {
        const SimpsonStep s = a.step[j];
        const double t = __dadd_rn(__dadd_rn(__dmul_rn(a.y[s.ia][i], s.ca), __dmul_rn(a.y[s.ib][i], s.cb)), __dmul_rn(a.y[s.ic][i], s.cc));
        I = __dadd_rn(I, t);
        a.node_out[j + 1][i] = I;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// hermiteInterpolate (utils.nim:282-312): every output is one hermiteSpline (utils.nim:273-279) on an interval
// [t_j, t_j+1] of the data set, or a copy of the last data point (x == t[high]). The host resolves which
// interval each x falls into (the reference's sorted / unsorted search, quirks included) and the spline's
// scalar factors; the kernel evaluates ALL outputs in one launch, keeping the interval's four vectors in
// registers while consecutive outputs stay in it and shifting two of them when the next output moves on by one.
//   out = ((y_j*h00 + dy_j*hA) + y_j1*h01) + dy_j1*hB,   hA = h10*(x2-x1), hB = h11*(x2-x1)
// ---------------------------------------------------------------------------------------------------
struct HermiteOut {
  int j;      // interval index, or the data point to copy when kind == 1
  int kind;   // 0 spline, 1 copy of y[j]
  double h00, hA, h01, hB;
};
struct HermiteManyArgs {
  const double* const* y;
  const double* const* dy;
  double* const* out;
  const HermiteOut* plan;
  int n_out;
  size_t n;
};

__device__ __forceinline__ double hermite_elem(double y1, double d1, double y2, double d2, const HermiteOut& p) {
  return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(y1, p.h00), __dmul_rn(d1, p.hA)), __dmul_rn(y2, p.h01)), __dmul_rn(d2, p.hB));
}

template <int W, int THREADS>
__global__ void __launch_bounds__(THREADS) hermite_many_kernel(const HermiteManyArgs a) {
  const size_t nvec = a.n / W;
  for (size_t v = (size_t)blockIdx.x * THREADS + threadIdx.x; v < nvec; v += (size_t)gridDim.x * THREADS) {
    int cur = -2;  // interval whose four vectors are in registers
    Pk<W> y1, d1, y2, d2;
    for (int o = 0; o < a.n_out; ++o) {
      const HermiteOut p = a.plan[o];  // uniform across the grid
      Pk<W> r;
      if (p.kind == 1) {
        r = ld_stream<W>(a.y[p.j] + v * W);
      } else {
        if (p.j == cur + 1) {  // moved on by one interval: two of the four vectors are already here
          y1 = y2; d1 = d2;
          y2 = ld_stream<W>(a.y[p.j + 1] + v * W);
          d2 = ld_stream<W>(a.dy[p.j + 1] + v * W);
        } else if (p.j != cur) {
          y1 = ld_stream<W>(a.y[p.j] + v * W);
          d1 = ld_stream<W>(a.dy[p.j] + v * W);
          y2 = ld_stream<W>(a.y[p.j + 1] + v * W);
          d2 = ld_stream<W>(a.dy[p.j + 1] + v * W);
        }
        cur = p.j;
#pragma unroll
        for (int e = 0; e < W; ++e) r.v[e] = hermite_elem(y1.v[e], d1.v[e], y2.v[e], d2.v[e], p);
      }
      st_stream<W>(a.out[o] + v * W, r);
    }
  }
  if (blockIdx.x == 0) {
    const size_t i = nvec * W + threadIdx.x;
    if (i < a.n) {
      for (int o = 0; o < a.n_out; ++o) {
        const HermiteOut p = a.plan[o];
        a.out[o][i] = (p.kind == 1) ? a.y[p.j][i] : hermite_elem(a.y[p.j][i], a.dy[p.j][i], a.y[p.j + 1][i], a.dy[p.j + 1][i], p);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// One step of the streaming cumtrapz(f, X, ctx, dx) (integrate.nim:165-174): I1 = I + (dyPrev + dyTemp) * h
// ---------------------------------------------------------------------------------------------------
template <int W, int THREADS>
__global__ void __launch_bounds__(THREADS)
    trapz_step_kernel(const double* __restrict__ I, const double* __restrict__ dy_prev, const double* __restrict__ dy_next, double h,
                      double* __restrict__ out, size_t n) {
  const size_t nvec = n / W;
  for (size_t v = (size_t)blockIdx.x * THREADS + threadIdx.x; v < nvec; v += (size_t)gridDim.x * THREADS) {
    const Pk<W> a = ld_stream<W>(I + v * W), b = ld_stream<W>(dy_prev + v * W), c = ld_stream<W>(dy_next + v * W);
    Pk<W> o;
#pragma unroll
    for (int e = 0; e < W; ++e) o.v[e] = __dadd_rn(a.v[e], __dmul_rn(__dadd_rn(b.v[e], c.v[e]), h));
    st_stream<W>(out + v * W, o);
  }
  if (blockIdx.x == 0) {
    const size_t i = nvec * W + threadIdx.x;
    if (i < n) out[i] = __dadd_rn(I[i], __dmul_rn(__dadd_rn(dy_prev[i], dy_next[i]), h));
  }
}

// Number of positions where a[i] != b[i] (NaN != NaN counts, like Nim's `!=` on floats), as a double through the
// library's deterministic reduction (all-reduced across shards like every other scalar of the path).
template <int W, int THREADS>
__global__ void __launch_bounds__(THREADS) neq_count_kernel(const double* __restrict__ a, const double* __restrict__ b, size_t n, ReduceScratch rs) {
  const size_t nvec = n / W;
  double acc = 0.0;
  for (size_t v = (size_t)blockIdx.x * THREADS + threadIdx.x; v < nvec; v += (size_t)gridDim.x * THREADS) {
    const Pk<W> x = ld_stream<W>(a + v * W), y = ld_stream<W>(b + v * W);
#pragma unroll
    for (int e = 0; e < W; ++e) acc = __dadd_rn(acc, (x.v[e] != y.v[e]) ? 1.0 : 0.0);
  }
  if (blockIdx.x == 0) {
    const size_t i = nvec * W + threadIdx.x;
    if (i < n) acc = __dadd_rn(acc, (a[i] != b[i]) ? 1.0 : 0.0);
  }
  grid_sum_finish<THREADS>(acc, rs);
}

}  // namespace b200rk

// ---------------------------------------------------------------------------------------------------
// cumsimpson(Y, X) in ONE pass (knob "fuse_simpson", default since round 2: 1.166 -> 0.674 ms at 2^23 x 33 points):
// the Simpson scan and the Hermite interpolation back onto X fused. While a thread walks the knot intervals it holds
// exactly what every sample inside the current interval needs — the integrals at both knots (I_j, I_{j+1}) and the
// data at both knots (the spline's slopes) — so the samples are emitted from registers and the knot integrals never
// travel to HBM: every data point read once, every result written once (2M vector passes instead of the 3.5M of
// simpson_scan_kernel + hermite_many_kernel). The host groups the samples by the interval that completes them.
// ---------------------------------------------------------------------------------------------------
namespace b200rk {

struct SimpsonFusedArgs {
  const double* const* y;     // sorted, trimmed data points
  double* const* out;         // one vector per returned sample
  const SimpsonStep* step;    // as for simpson_scan_kernel
  const int* tail;            // per step: 0 = knots are the data points (ic, ia) of a regular pair, 1 = (ib, ic) of the tail rule
  const int* emit_begin;      // steps + 1 offsets into emit / emit_out
  const HermiteOut* emit;     // samples grouped by knot interval; kind 1 = copy of the integral at the interval's right knot
  const int* emit_out;        // which output vector each grouped sample is
  int steps, first;
  size_t n;
};

template <int W, int THREADS>
__global__ void __launch_bounds__(THREADS) simpson_fused_kernel(const SimpsonFusedArgs a) {
  const size_t nvec = a.n / W;
  for (size_t v = (size_t)blockIdx.x * THREADS + threadIdx.x; v < nvec; v += (size_t)gridDim.x * THREADS) {
    Pk<W> prev_a = ld_stream<W>(a.y[a.first] + v * W), I;
#pragma unroll
    for (int e = 0; e < W; ++e) I.v[e] = __dadd_rn(prev_a.v[e], -prev_a.v[e]);
    for (int j = 0; j < a.steps; ++j) {
      const SimpsonStep s = a.step[j];
      const int tail = a.tail[j];
      const Pk<W> ya = ld_stream<W>(a.y[s.ia] + v * W);
      const Pk<W> yb = ld_stream<W>(a.y[s.ib] + v * W);
      Pk<W> yc;
      if (s.reuse) yc = prev_a;
      else yc = ld_stream<W>(a.y[s.ic] + v * W);
      Pk<W> I1;
#pragma unroll
      for (int e = 0; e < W; ++e) {
        const double t = __dadd_rn(__dadd_rn(__dmul_rn(ya.v[e], s.ca), __dmul_rn(yb.v[e], s.cb)), __dmul_rn(yc.v[e], s.cc));
        I1.v[e] = __dadd_rn(I.v[e], t);
      }
      const Pk<W> d0 = tail ? yb : yc, d1 = tail ? yc : ya;   // the data at the interval's two knots = the spline's slopes
      for (int q = a.emit_begin[j]; q < a.emit_begin[j + 1]; ++q) {
        const HermiteOut p = a.emit[q];
        Pk<W> r;
        if (p.kind == 1) r = I1;
        else {
#pragma unroll
          for (int e = 0; e < W; ++e) r.v[e] = hermite_elem(I.v[e], d0.v[e], I1.v[e], d1.v[e], p);
        }
        st_stream<W>(a.out[a.emit_out[q]] + v * W, r);
      }
      I = I1;
      prev_a = ya;
    }
  }
  if (blockIdx.x == 0) {  // ragged tail (n % W elements)
    const size_t i = nvec * W + threadIdx.x;
    if (i < a.n) {
      const double y0 = a.y[a.first][i];
      double I = __dadd_rn(y0, -y0);
      for (int j = 0; j < a.steps; ++j) {
        const SimpsonStep s = a.step[j];
        const double ya = a.y[s.ia][i], yb = a.y[s.ib][i], yc = a.y[s.ic][i];
        const double I1 = __dadd_rn(I, __dadd_rn(__dadd_rn(__dmul_rn(ya, s.ca), __dmul_rn(yb, s.cb)), __dmul_rn(yc, s.cc)));
        const double d0 = a.tail[j] ? yb : yc, d1 = a.tail[j] ? yc : ya;
        for (int q = a.emit_begin[j]; q < a.emit_begin[j + 1]; ++q) {
          const HermiteOut p = a.emit[q];
          a.out[a.emit_out[q]][i] = (p.kind == 1) ? I1 : hermite_elem(I, d0, I1, d1, p);
        }
        I = I1;
      }
    }
  }
}

}  // namespace b200rk
