// stencil_attempt.cuh — a whole adaptive attempt of an FSAL pair for the built-in Lorenz-96 right-hand side in ONE
// kernel (cyclic; one GPU, or one contiguous shard per GPU with a halo from the ring neighbours). EXPERIMENTAL: knob "fuse_stencil_attempt", off by default until it has been run and
// measured on the GPU (bit-identical to the stage_l96_kernel / finish_kernel pipeline under host emulation).
//
// The element-local fused attempt (kernels.cuh: fused_attempt_kernel) keeps k_1..k_S of an element in registers because
// nothing crosses elements. Lorenz-96 reads y[i-2], y[i-1], y[i+1] (ode.nim's ODEProc is the caller's; this one is
// b200rk's built-in), so stage s at element i needs the stage input at three neighbours — which needs their k's, and so
// on down the stages. Overlapped tiling removes the dependence between CTAs: a CTA loads a tile of TW elements of y and
// k1 (FSAL) that overlaps its neighbours' tiles by HL elements on the left and HR on the right, and evaluates every
// stage on the whole tile. Each right-hand-side evaluation makes 2 more elements on the left edge and 1 more on the
// right edge depend on data outside the tile; after the S-1 evaluations of an attempt positions [2(S-1), TW-(S-1)) are
// still exact. Only positions [HL, TW-HR) are stored (HL >= 2(S-1), HR >= S-1, both multiples of 4 so every tile stays
// 32-byte aligned): 2 % of the arithmetic is redundant, and HBM traffic per attempt drops from 53 vector passes
// (Tsit54: 41 stage/finish + 12 stencil) to 4 — read y and k1, write yNew and k_S.
// Sharded over GPUs the same overlap reaches HL / HR elements into the neighbouring shards: ONE halo exchange of y and k1
// per IntegratorProc call (they do not change between the retries of an attempt) replaces the 3-element exchange in front
// of each of the S-1 right-hand-side evaluations.
//
// Within the CTA the stage input travels through shared memory (two buffers, so one barrier per stage): a thread
// computes the stage input at its own E = 2*J positions from registers, publishes it, and after the barrier applies the
// stencil to its positions. Per element the operations are exactly those of stage_elem / stage_l96_kernel /
// finish_elem in the same order (__dmul_rn / __dadd_rn; err_ratio for r), so yNew, k_S and every r*r term are bit-identical
// to the unfused pipeline; the sum of the r*r terms is taken in a different order (tile by tile), like every other
// grid geometry of the reducing kernels.
#pragma once
#include "kernels.cuh"

namespace b200rk {

template <int S>
struct StencilTile {
  static constexpr int HL = ((2 * (S - 1) + 3) / 4) * 4;   // left overlap: 2 per right-hand-side evaluation, rounded to 32 bytes
  static constexpr int HR = (((S - 1) + 3) / 4) * 4;       // right overlap: 1 per evaluation
};

// Sharded state (one contiguous block of the cyclic global vector per GPU): the HL elements before the block (left_*[0]
// is element -HL) and the HR elements after it (right_*[0] is element n), of y and of k1. They are either a small buffer
// filled by one grouped ncclSend/ncclRecv per IntegratorProc call, or — inside a solver, when the ring neighbours'
// vectors are peer-mapped (CUDA IPC over NVLink) — the neighbours' own vectors: the edge tiles then read the left
// neighbour's tail and the right neighbour's head in place and the stencil path has no collective at all (executor.cu).
// All null on a single GPU: the block is the whole ring and the edge tiles index it cyclically.
struct L96Halo {
  const double *left_y, *right_y, *left_k, *right_k;
};

template <int S>
struct L96AttemptArgs {
  FusedArgs<S> f;     // y, k1, a/b/bh rows, dt, cb, cbh, tolerances, ynew, ks_out, n, rs (the parameter fields are unused)
  double F;           // forcing (the backward pass g = -f(-t, y), ode.nim:545, is the kernel's NEG template argument)
  L96Halo halo;       // sharded: where the elements around the block live (all null on a single GPU)
};

// lorenz96_kernel's ((y[i+1] - y[i-2]) * y[i-1] - y[i]) + F at the two adjacent positions p, p+1 a thread owns: c0, c1 are
// its own stage inputs (registers), sh[p-2], sh[p-1] (one 128-bit shared-memory load) and sh[p+2] its neighbours'.
// NEG: the backward pass g = -f(-t, y) (ode.nim:545). v * (+1.0) == v and v * (-1.0) == -v exactly, so instead of
// stage_l96_kernel's multiplication by sgn the sign is a compile-time negation (an operand modifier, no fp64 issue).
template <bool NEG>
__device__ __forceinline__ void l96_pair(const double* sh, int p, double c0, double c1, double F, double& k0, double& k1) {
  const double m2 = sh[p - 2], m1 = sh[p - 1], p2 = sh[p + 2];
  const double v0 = __dadd_rn(__dadd_rn(__dmul_rn(__dadd_rn(c1, -m2), m1), -c0), F);
  const double v1 = __dadd_rn(__dadd_rn(__dmul_rn(__dadd_rn(p2, -m1), c0), -c1), F);
  k0 = NEG ? -v0 : v0;
  k1 = NEG ? -v1 : v1;
}

// stage s (compile-time) of every element the thread owns; recursion keeps the row masks template constants.
// A thread owns J pairs of adjacent positions (p, p+1).
template <int PAT, int s, int J, int TW, bool NEG>
struct L96Stages {
  template <int S>
  __device__ __forceinline__ static void run(const double (&y)[2 * J], double (&k)[2 * J][S], double (&in)[2 * J], const int (&pos)[J],
                                             double (*buf)[TW + 4], const L96AttemptArgs<S>& a) {
    if constexpr (s > 2) L96Stages<PAT, s - 1, J, TW, NEG>::run(y, k, in, pos, buf, a);
    double* sh = buf[s & 1] + 2;   // sh[-2], sh[-1] and sh[TW] are zero pads (their consumers are outside the stored range)
#pragma unroll
    for (int e = 0; e < 2 * J; ++e) {
      const double acc = const_wsum<S, Pattern<PAT>::a(s - 2)>(k[e], a.f.a[s - 2]);
      in[e] = __dadd_rn(y[e], __dmul_rn(acc, a.f.dt));                                   // stage_elem
    }
#pragma unroll
    for (int j = 0; j < J; ++j) { sh[pos[j]] = in[2 * j]; sh[pos[j] + 1] = in[2 * j + 1]; }
    __syncthreads();   // the other buffer was last read before the previous stage's barrier: one barrier per stage
#pragma unroll
    for (int j = 0; j < J; ++j) l96_pair<NEG>(sh, pos[j], in[2 * j], in[2 * j + 1], a.F, k[2 * j][s - 1], k[2 * j + 1][s - 1]);
  }
};

// Position p of the tile that stores [tile0, tile0 + OUT): element tile0 - HL + p of the block — from the block itself,
// cyclically (single GPU, left == nullptr), or from around the block (sharded: left[0] = element -HL, right[0] = element n).
template <int HL, int HR>
__device__ __forceinline__ double l96_edge_load(const double* v, const double* left, const double* right, size_t n, size_t tile0, int p) {
  if (left) {
    const long long idx = (long long)tile0 - HL + p;
    if (idx < 0) return left[HL + idx];
    if ((size_t)idx < n) return v[idx];
    if ((size_t)idx - n < (size_t)HR) return right[(size_t)idx - n];
    return 0.0;   // further right: beyond what any stored position depends on
  }
  return v[(tile0 + (size_t)p + (n - (size_t)HL % n)) % n];
}

template <int PAT, int J, int THREADS, bool NEG>
__global__ void __launch_bounds__(THREADS) l96_attempt_kernel(const L96AttemptArgs<Pattern<PAT>::S> a) {
  constexpr int S = Pattern<PAT>::S;
  constexpr int E = 2 * J, TW = E * THREADS;
  constexpr int HL = StencilTile<S>::HL, HR = StencilTile<S>::HR, OUT = TW - HL - HR;
  __shared__ double buf[2][TW + 4];
  const size_t n = a.f.n;
  const size_t tile0 = (size_t)blockIdx.x * OUT;          // first stored element of this tile
  if (threadIdx.x < 3) {
    const int q = threadIdx.x < 2 ? (int)threadIdx.x : TW + 2;
    buf[0][q] = 0.0; buf[1][q] = 0.0;
  }
  double y[E], k[E][S], in[E];
  int pos[J];   // first position of each pair the thread owns
  const bool interior = tile0 >= (size_t)HL && tile0 - HL + TW <= n;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int p = 2 * ((int)threadIdx.x + j * THREADS);
    pos[j] = p;
    if (interior) {
      const size_t g = tile0 - HL + p;
      const Pk<2> yv = ld_stream<2>(a.f.y + g), kv = ld_stream<2>(a.f.k1 + g);
      y[2 * j] = yv.v[0]; y[2 * j + 1] = yv.v[1];
      k[2 * j][0] = kv.v[0]; k[2 * j + 1][0] = kv.v[1];
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        y[2 * j + h] = l96_edge_load<HL, HR>(a.f.y, a.halo.left_y, a.halo.right_y, n, tile0, p + h);
        k[2 * j + h][0] = l96_edge_load<HL, HR>(a.f.k1, a.halo.left_k, a.halo.right_k, n, tile0, p + h);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < E; ++e)
#pragma unroll
    for (int j = 1; j < S; ++j) k[e][j] = 0.0;
  L96Stages<PAT, S, J, TW, NEG>::run(y, k, in, pos, buf, a);

  double acc = 0.0;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int p = pos[j];
    const size_t g = tile0 + (size_t)(p - HL);            // p and HL even: a pair is stored whole or not at all (up to n)
    const bool stored = p >= HL && p < HL + OUT && g < n;
    double yn[2], ks[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int e = 2 * j + h;
      ks[h] = k[e][S - 1];
      if (Pattern<PAT>::last) yn[h] = in[e];
      else yn[h] = __dadd_rn(y[e], __dmul_rn(const_wsum<S, Pattern<PAT>::b()>(k[e], a.f.b), a.f.cb));
      const double lo = __dmul_rn(const_wsum<S, Pattern<PAT>::bh()>(k[e], a.f.bh), a.f.cbh);
      double err;
      if (Pattern<PAT>::direct) err = lo;
      else err = __dadd_rn(yn[h], -__dadd_rn(y[e], lo));
      const double tol = __dadd_rn(a.f.absTol, __dmul_rn(fabs(yn[h]), a.f.relTol));
      const double r = err_ratio(err, tol);
      if (stored && g + h < n) acc = __dadd_rn(acc, __dmul_rn(r, r));
    }
    if (stored) {
      if (g + 1 < n) {
        Pk<2> o;
        o.v[0] = yn[0]; o.v[1] = yn[1];
        st_stream<2>(a.f.ynew + g, o);
        o.v[0] = ks[0]; o.v[1] = ks[1];
        st_stream<2>(a.f.ks_out + g, o);
      } else {
        a.f.ynew[g] = yn[0];
        a.f.ks_out[g] = ks[0];
      }
    }
  }
#ifdef B200RK_EMULATE_SERIAL_SUM
  // host emulation (threads of a CTA run concurrently, CTAs one after another): thread 0 adds the CTA's terms in thread
  // order to the running sum and the last CTA publishes it the way grid_sum_finish's last CTA does
  __shared__ double red[THREADS];
  red[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = (blockIdx.x == 0) ? 0.0 : *a.f.rs.result;
    for (int i = 0; i < THREADS; ++i) t = __dadd_rn(t, red[i]);
    *a.f.rs.result = t;
    if (blockIdx.x == gridDim.x - 1 && a.f.rs.result_host) {
      if (a.f.rs.mail.world > 1) { t = emul_peer_exchange(t, a.f.rs); *a.f.rs.result = t; }
      *a.f.rs.result_host = t;
      *a.f.rs.seq_host = a.f.rs.seq;
    }
  }
#else
  grid_sum_finish<THREADS>(acc, a.f.rs);
#endif
}

// ---------------------------------------------------------------------------------------------------
// A whole RK4 step (ode.nim:180-189) for the built-in Lorenz-96 right-hand side in one kernel, same tiling: four
// stencil evaluations (k1 = f(y) included: RK4 has no FSAL) -> overlap 8 left / 4 right; reads y, writes yNew. Stage
// inputs y + k*hdt (hdt = 0.5*dt from the host), y + k3*dt and the final combine are fused_rk4_elem / rk4_elem.
// ---------------------------------------------------------------------------------------------------
struct L96Rk4Args {
  const double* y;
  double* ynew;
  size_t n;
  double F, hdt, dt, c6;
  L96Halo halo;           // sharded: left_y = the 8 elements before the block, right_y = the 4 after it (null on a single GPU)
};

template <int J, int THREADS, bool NEG>
__global__ void __launch_bounds__(THREADS) l96_rk4_kernel(const L96Rk4Args a) {
  constexpr int E = 2 * J, TW = E * THREADS, HL = 8, HR = 4, OUT = TW - HL - HR;
  __shared__ double buf[2][TW + 4];
  const size_t n = a.n;
  const size_t tile0 = (size_t)blockIdx.x * OUT;
  if (threadIdx.x < 3) {
    const int q = threadIdx.x < 2 ? (int)threadIdx.x : TW + 2;
    buf[0][q] = 0.0; buf[1][q] = 0.0;
  }
  double y[E], k[4][E], in[E];
  int pos[J];
  const bool interior = tile0 >= (size_t)HL && tile0 - HL + TW <= n;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int p = 2 * ((int)threadIdx.x + j * THREADS);
    pos[j] = p;
    if (interior) {
      const Pk<2> yv = ld_stream<2>(a.y + (tile0 - HL + p));
      y[2 * j] = yv.v[0]; y[2 * j + 1] = yv.v[1];
    } else {
      y[2 * j] = l96_edge_load<HL, HR>(a.y, a.halo.left_y, a.halo.right_y, n, tile0, p);
      y[2 * j + 1] = l96_edge_load<HL, HR>(a.y, a.halo.left_y, a.halo.right_y, n, tile0, p + 1);
    }
  }
#pragma unroll
  for (int s = 0; s < 4; ++s) {   // stage s + 1: its input, published; then k_{s+1} from the neighbours' inputs
    double* sh = buf[s & 1] + 2;
    const double c = (s == 3) ? a.dt : a.hdt;
#pragma unroll
    for (int e = 0; e < E; ++e) in[e] = (s == 0) ? y[e] : __dadd_rn(y[e], __dmul_rn(k[s - 1 < 0 ? 0 : s - 1][e], c));
#pragma unroll
    for (int j = 0; j < J; ++j) { sh[pos[j]] = in[2 * j]; sh[pos[j] + 1] = in[2 * j + 1]; }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < J; ++j) l96_pair<NEG>(sh, pos[j], in[2 * j], in[2 * j + 1], a.F, k[s][2 * j], k[s][2 * j + 1]);
  }
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int p = pos[j];
    const size_t g = tile0 + (size_t)(p - HL);
    if (p >= HL && p < HL + OUT && g < n) {
      Pk<2> o;
#pragma unroll
      for (int h = 0; h < 2; ++h) o.v[h] = rk4_elem(y[2 * j + h], k[0][2 * j + h], k[1][2 * j + h], k[2][2 * j + h], k[3][2 * j + h], a.c6);
      if (g + 1 < n) st_stream<2>(a.ynew + g, o);
      else a.ynew[g] = o.v[0];
    }
  }
}

}  // namespace b200rk
