// stencil_attempt.cuh — a whole adaptive attempt of an FSAL pair for the built-in Lorenz-96 right-hand side in ONE
// kernel (cyclic; one GPU, or one contiguous shard per GPU with a halo from the ring neighbours). Knob "fuse_stencil_attempt", the default since its first
// B200 runs in round 2 (bit-identical to the stage_l96_kernel / finish_kernel pipeline on the GPU and under host emulation).
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
// its own stage inputs (registers); m2 = in[p-2], m1 = in[p-1] (one 128-bit shared-memory load) and p2 = in[p+2] its neighbours'.
// NEG: the backward pass g = -f(-t, y) (ode.nim:545). v * (+1.0) == v and v * (-1.0) == -v exactly, so instead of
// stage_l96_kernel's multiplication by sgn the sign is a compile-time negation (an operand modifier, no fp64 issue).
template <bool NEG>
__device__ __forceinline__ void l96_pair(double m2, double m1, double p2, double c0, double c1, double F, double& k0, double& k1) {
  const double v0 = __dadd_rn(__dadd_rn(__dmul_rn(__dadd_rn(c1, -m2), m1), -c0), F);
  const double v1 = __dadd_rn(__dadd_rn(__dmul_rn(__dadd_rn(p2, -m1), c0), -c1), F);
  k0 = NEG ? -v0 : v0;
  k1 = NEG ? -v1 : v1;
}

// const_wsum's left-associated chain split at term NT: the prefix over the kept terms j < NT, and its continuation with
// term NT. prefix followed by finish performs exactly const_wsum's operations in const_wsum's order (bit-identical); the
// split exists so that the prefix — which needs only k_1..k_NT — can be evaluated while the operands of k_{NT+1} (the
// neighbours' stage inputs) are still on their way from shared memory.
template <int NK, uint32_t MASK, int NT>
__device__ __forceinline__ double const_wsum_prefix(const double (&k)[NK], const double* w) {
  return const_wsum<NK, (MASK & ((1u << NT) - 1u))>(k, w);
}
template <int NK, uint32_t MASK, int NT>
__device__ __forceinline__ double const_wsum_finish(double prefix, const double (&k)[NK], const double* w) {
  if constexpr (((MASK >> NT) & 1u) == 0u) return prefix;
  else {
    const double p = __dmul_rn(k[NT], w[NT]);
    if constexpr ((MASK & ((1u << NT) - 1u)) != 0u) return __dadd_rn(prefix, p);
    else return p;
  }
}

// stage s (compile-time) of every element the thread owns; recursion keeps the row masks template constants.
// A thread owns J pairs of adjacent positions (p, p+1).
// Software pipelining across stages: on entry part[e] holds row s's sum over k_1..k_{s-2}; the stage adds the k_{s-1} term,
// publishes its stage input, and — between issuing the shared-memory loads of the neighbours' inputs and consuming them in
// the stencil — evaluates the NEXT row's prefix over k_1..k_{s-1} (after the last stage: the prefixes of the b and bHat
// rows), so the ~30-cycle shared-memory latency right after the barrier hides under independent fp64 work instead of
// stalling the warp (first persistent form: 17 % of all stall samples sat on the stencil's first DADD).
template <int PAT, int s, int J, int TW, bool NEG>
struct L96Stages {
  template <int S>
  __device__ __forceinline__ static void run(const double (&y)[2 * J], double (&k)[2 * J][S], double (&in)[2 * J], double (&part)[2 * J],
                                             double (&part_bh)[2 * J], const int (&pos)[J], double (*buf)[TW + 4], const L96AttemptArgs<S>& a) {
    if constexpr (s > 2) L96Stages<PAT, s - 1, J, TW, NEG>::run(y, k, in, part, part_bh, pos, buf, a);
    double* sh = buf[s & 1] + 2;   // sh[-2], sh[-1] and sh[TW] are zero pads (their consumers are outside the stored range)
#pragma unroll
    for (int e = 0; e < 2 * J; ++e) {
      const double acc = const_wsum_finish<S, Pattern<PAT>::a(s - 2), s - 2>(part[e], k[e], a.f.a[s - 2]);
      in[e] = __dadd_rn(y[e], __dmul_rn(acc, a.f.dt));                                   // stage_elem
    }
#pragma unroll
    for (int j = 0; j < J; ++j) { sh[pos[j]] = in[2 * j]; sh[pos[j] + 1] = in[2 * j + 1]; }
    __syncthreads();   // the other buffer was last read before the previous stage's barrier: one barrier per stage
    double m2[J], m1[J], p2[J];
#pragma unroll
    for (int j = 0; j < J; ++j) { m2[j] = sh[pos[j] - 2]; m1[j] = sh[pos[j] - 1]; p2[j] = sh[pos[j] + 2]; }
#pragma unroll
    for (int e = 0; e < 2 * J; ++e) {
      if constexpr (s < S) part[e] = const_wsum_prefix<S, Pattern<PAT>::a(s - 1), s - 1>(k[e], a.f.a[s - 1]);
      else {
        if constexpr (!Pattern<PAT>::last) part[e] = const_wsum_prefix<S, Pattern<PAT>::b(), S - 1>(k[e], a.f.b);
        part_bh[e] = const_wsum_prefix<S, Pattern<PAT>::bh(), S - 1>(k[e], a.f.bh);
      }
    }
#pragma unroll
    for (int j = 0; j < J; ++j) l96_pair<NEG>(m2[j], m1[j], p2[j], in[2 * j], in[2 * j + 1], a.F, k[2 * j][s - 1], k[2 * j + 1][s - 1]);
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

// ---- tile prefetch: TMA 1-D bulk copies into shared memory, completion on an mbarrier ---------------------------------
// One elected thread arms the barrier with the byte count and issues two cp.async.bulk copies (the tile's y and k1, 8 KB
// each); the copies run in the async proxy while the CTA computes the previous tile, cost no registers, and every thread
// picks its four elements up from shared memory with two conflict-free 128-bit loads once the barrier's phase flips.
#ifndef B200RK_HOST_EMULATION
__device__ __forceinline__ unsigned int smem_addr(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tile_barrier_init(unsigned long long* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tile_prefetch(double* dst_y, const double* src_y, double* dst_k, const double* src_k, unsigned int bytes_each,
                                              unsigned long long* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(2u * bytes_each) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_y)), "l"(src_y),
               "r"(bytes_each), "r"(smem_addr(bar))
               : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_k)), "l"(src_k),
               "r"(bytes_each), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void tile_wait(unsigned long long* bar, unsigned int parity) {
  unsigned int done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done)
                 : "r"(smem_addr(bar)), "r"(parity)
                 : "memory");
  } while (!done);
}
#else
// host emulation: the elected thread copies synchronously; the stage barriers that follow order it before the readers
inline void tile_barrier_init(unsigned long long*) {}
inline void tile_prefetch(double* dst_y, const double* src_y, double* dst_k, const double* src_k, unsigned int bytes_each, unsigned long long*) {
  for (unsigned int i = 0; i < bytes_each / 8; ++i) { dst_y[i] = src_y[i]; dst_k[i] = src_k[i]; }
}
inline void tile_wait(unsigned long long*, unsigned int) {}
#endif

#ifndef L96_PREFETCH_DEPTH
#define L96_PREFETCH_DEPTH 1
#endif
// Persistent grid: a CTA walks the tiles blockIdx.x, blockIdx.x + gridDim.x, ...; while it evaluates the S - 1 stages of
// one tile the next tile's y and k1 are already on their way into shared memory (tile_prefetch), so the HBM latency of a
// tile hides under the ~500 fp64 instructions per warp of the previous one instead of idling the CTA (first measured form,
// one tile per CTA and register loads: 188 us per attempt at 2^24 with the fp64 pipe 55 % busy and 17 % of the warp
// stalls on the tile's loads).
template <int PAT, int J, int THREADS, bool NEG>
__global__ void __launch_bounds__(THREADS, ((Pattern<PAT>::S > 7 || J > 2) ? 2 : 3) * (256 / THREADS)) l96_attempt_kernel(const L96AttemptArgs<Pattern<PAT>::S> a) {
  constexpr int S = Pattern<PAT>::S;
  constexpr int E = 2 * J, TW = E * THREADS;
  constexpr int HL = StencilTile<S>::HL, HR = StencilTile<S>::HR, OUT = TW - HL - HR;
  constexpr int NST = (TW <= 512) ? L96_PREFETCH_DEPTH : 1;   // tiles in flight ahead of the one being evaluated (static shared memory: 48 KB)
  __shared__ double buf[2][TW + 4];
  alignas(128) __shared__ double staged[NST][2][TW];   // ring: [.][0] a tile's y, [.][1] its k1 (FSAL) — written by the bulk copies
  alignas(8) __shared__ unsigned long long tile_bar[NST];
  const size_t n = a.f.n;
  const size_t n_tiles = (n + OUT - 1) / OUT;
  if (threadIdx.x < 3) {
    const int q = threadIdx.x < 2 ? (int)threadIdx.x : TW + 2;
    buf[0][q] = 0.0; buf[1][q] = 0.0;
  }
  // a tile whose TW positions all lie inside the block is read with the bulk copies; the first and last tile(s) reach
  // around the ring (or into the neighbouring shards) and take the element-wise path
  auto interior_tile = [&](size_t t) { const size_t t0 = t * (size_t)OUT; return t0 >= (size_t)HL && t0 - HL + TW <= n; };
  size_t tile = blockIdx.x;
  if (threadIdx.x == 0) {
    for (int q = 0; q < NST; ++q) {
      tile_barrier_init(&tile_bar[q]);
      const size_t t = tile + (size_t)q * gridDim.x;
      if (t < n_tiles && interior_tile(t))
        tile_prefetch(staged[q][0], a.f.y + (t * OUT - HL), staged[q][1], a.f.k1 + (t * OUT - HL), TW * 8u, &tile_bar[q]);
    }
  }
  __syncthreads();
  unsigned int phase_bits = 0;   // bit q: parity of the next completion of ring slot q
  int slot = 0;
  double acc = 0.0;
  int pos[J];   // first position of each pair the thread owns
#pragma unroll
  for (int j = 0; j < J; ++j) pos[j] = 2 * ((int)threadIdx.x + j * THREADS);
  for (; tile < n_tiles; tile += gridDim.x) {
  const size_t tile0 = tile * (size_t)OUT;          // first stored element of this tile
  double y[E], k[E][S], in[E];
  const bool interior = interior_tile(tile);
  if (interior) { tile_wait(&tile_bar[slot], (phase_bits >> slot) & 1u); phase_bits ^= 1u << slot; }
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int p = pos[j];
    if (interior) {
      const double2 yv = *reinterpret_cast<const double2*>(&staged[slot][0][p]), kv = *reinterpret_cast<const double2*>(&staged[slot][1][p]);   // LDS.128, conflict-free
      y[2 * j] = yv.x; y[2 * j + 1] = yv.y;
      k[2 * j][0] = kv.x; k[2 * j + 1][0] = kv.y;
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        y[2 * j + h] = l96_edge_load<HL, HR>(a.f.y, a.halo.left_y, a.halo.right_y, n, tile0, p + h);
        k[2 * j + h][0] = l96_edge_load<HL, HR>(a.f.k1, a.halo.left_k, a.halo.right_k, n, tile0, p + h);
      }
    }
  }
  __syncthreads();   // every thread holds its part of the tile: the staging buffers (and the previous tile's last stage buffer) are free
  {
    const size_t nxt = tile + (size_t)NST * gridDim.x;   // the ring slot just emptied receives the tile NST rounds ahead
    if (threadIdx.x == 0 && nxt < n_tiles && interior_tile(nxt))
      tile_prefetch(staged[slot][0], a.f.y + (nxt * OUT - HL), staged[slot][1], a.f.k1 + (nxt * OUT - HL), TW * 8u, &tile_bar[slot]);
    slot = (slot + 1 == NST) ? 0 : slot + 1;
  }
#pragma unroll
  for (int e = 0; e < E; ++e)
#pragma unroll
    for (int j = 1; j < S; ++j) k[e][j] = 0.0;
  double part[E], part_bh[E];
#pragma unroll
  for (int e = 0; e < E; ++e) { part[e] = 0.0; part_bh[e] = 0.0; }
  L96Stages<PAT, S, J, TW, NEG>::run(y, k, in, part, part_bh, pos, buf, a);   // leaves the b / bHat rows' prefixes in part / part_bh

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
      else yn[h] = __dadd_rn(y[e], __dmul_rn(const_wsum_finish<S, Pattern<PAT>::b(), S - 1>(part[e], k[e], a.f.b), a.f.cb));
      const double lo = __dmul_rn(const_wsum_finish<S, Pattern<PAT>::bh(), S - 1>(part_bh[e], k[e], a.f.bh), a.f.cbh);
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
  }  // tiles
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
// The same attempt with WARP-sized overlapped tiles: no shared memory, no block barrier.
// In the CTA-tile kernel above every stage is  publish -> __syncthreads -> shared-memory loads -> stencil, and the warps of a
// CTA move through it in lockstep: measured on the B200 (profiles/r02_ncu_full_l96_attempt_*), the fp64 pipe stays 60 % busy
// with a quarter of all warp stalls on the barrier and on the shared-memory latency right behind it. Here a tile is what ONE
// warp holds: lane l owns the 4 adjacent positions 4l .. 4l+3 of a 128-position tile (one 256-bit load each of y and k1),
// and the three neighbouring stage inputs an evaluation needs — in[-2], in[-1] from lane l-1, in[+4] from lane l+1 — travel
// by warp shuffle. Warps never wait for each other, so the scheduler hides one warp's shuffle and dependency latency under
// the arithmetic of the others, as in the element-local fused kernel. The price is a larger overlap: HL + HR of 128
// positions (16 % for the 7-stage pairs, 19 % for Vern65) instead of 2 % of a 1024-wide tile. Per-element arithmetic is the
// CTA-tile kernel's (const_wsum rows, l96_pair's operation order, err_ratio): yNew and k_S are bit-identical.
// MEASURED (B200, Tsit54, 2^24; profiles/r02_l96_attempt_variants.md): 4 elements per lane — 164 us per attempt with plain register
// loads (fp64 pipe 68 % busy, a quarter of the stall samples on the tile's own loads), 182-188 us with the cp.async prefetch at 80
// registers (spills), 185 us at 124 registers / 2 CTAs per SM; 8 elements per lane (211 registers, 1 CTA per SM, half the overlap and
// shuffles) — 170 us; against 154 us for the CTA-tile kernel. The kernel therefore stays behind knob "l96_warp_tiles" (default 0;
// 8 or 4 = elements per lane) as the measured alternative.
// Positions whose neighbours lie outside the tile read another lane's value through the shuffle's clamping; they are
// in the overlap and never stored (lane 0's left inputs and lane 31's right input go bad first: 2 resp. 1 position per stage).
// ---------------------------------------------------------------------------------------------------
template <int S, int E>
struct WarpTile {
  static constexpr int TW = 32 * E;
  static constexpr int HL = StencilTile<S>::HL, HR = StencilTile<S>::HR, OUT = TW - HL - HR;
};

// Stencil at the lane's E adjacent positions: k[i] = ((in[i+1] - in[i-2]) * in[i-1] - in[i]) + F in l96_pair's operation order; in[-2],
// in[-1] come from lane l-1 (its last two inputs), in[E] from lane l+1 (its first).
template <int PAT, int s, int E, bool NEG>
struct L96WarpStages {
  template <int S>
  __device__ __forceinline__ static void run(const double (&y)[E], double (&k)[E][S], double (&in)[E], const L96AttemptArgs<S>& a) {
    if constexpr (s > 2) L96WarpStages<PAT, s - 1, E, NEG>::run(y, k, in, a);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const double acc = const_wsum<S, Pattern<PAT>::a(s - 2)>(k[e], a.f.a[s - 2]);
      in[e] = __dadd_rn(y[e], __dmul_rn(acc, a.f.dt));                                   // stage_elem
    }
    double w[E + 3];   // w[i] = in[i - 2]
    w[0] = __shfl_up_sync(0xffffffffu, in[E - 2], 1);
    w[1] = __shfl_up_sync(0xffffffffu, in[E - 1], 1);
    w[E + 2] = __shfl_down_sync(0xffffffffu, in[0], 1);
#pragma unroll
    for (int e = 0; e < E; ++e) w[e + 2] = in[e];
#pragma unroll
    for (int e = 0; e < E; e += 2)
      l96_pair<NEG>(w[e], w[e + 1], w[e + 4], w[e + 2], w[e + 3], a.F, k[e][s - 1], k[e + 1][s - 1]);
  }
};

// Register-free prefetch of a lane's own bytes of the next tile: cp.async (LDGSTS) into a per-thread staging area.
// A lane later reads back exactly the bytes it copied itself, so no warp- or block-level synchronisation is involved —
// cp.async.wait_group on the issuing thread is all the ordering there is.
#ifndef B200RK_HOST_EMULATION
__device__ __forceinline__ void lane_prefetch16(double* dst, const double* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned int)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void lane_prefetch_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void lane_prefetch_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#else
inline void lane_prefetch16(double* dst, const double* src) { dst[0] = src[0]; dst[1] = src[1]; }
inline void lane_prefetch_commit() {}
inline void lane_prefetch_wait() {}
#endif

// E = 8 elements per lane (256-position tiles, 8-9 % overlap, ~170-200 registers: ONE CTA per SM — the fp64 pipe of a scheduler is
// saturated by a single warp once it has 4 independent chains in flight (measured: 8-cycle dependent latency, 2 cycles per issue), so
// two warps of 8 chains per scheduler are enough, and wider lanes halve both the overlap and the shuffles per element).
// E = 4: 128-position tiles, 16-19 % overlap, 80 registers, 3 CTAs per SM.
template <int PAT, int E, int THREADS, bool NEG>
__global__ void __launch_bounds__(THREADS, E >= 8 ? 1 : (Pattern<PAT>::S > 7 ? 2 : 3)) l96_warp_attempt_kernel(const L96AttemptArgs<Pattern<PAT>::S> a) {
  constexpr int S = Pattern<PAT>::S;
  using T = WarpTile<S, E>;
  constexpr int HL = T::HL, HR = T::HR, OUT = T::OUT, TW = T::TW;
  static_assert(E % 4 == 0, "a lane stores whole 32-byte groups");
  const size_t n = a.f.n;
  const size_t n_tiles = (n + OUT - 1) / OUT;
  const int lane = threadIdx.x & 31;
  const size_t warp = (size_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5), n_warps = (size_t)gridDim.x * (THREADS / 32);
  const int p = E * lane;                                   // first of the lane's E positions
  double acc = 0.0;
  alignas(16) __shared__ double staged[2][E / 2][THREADS][2];   // [y | k1][pair][thread][2]: a lane's pairs lie 16 bytes apart across lanes (conflict-free); written by cp.async
  auto interior_tile = [&](size_t t) { const size_t t0 = t * (size_t)OUT; return t0 >= (size_t)HL && t0 - HL + TW <= n; };
  auto prefetch = [&](size_t t) {
    const size_t g = t * (size_t)OUT - HL + p;
#pragma unroll
    for (int q = 0; q < E / 2; ++q) {
      lane_prefetch16(staged[0][q][threadIdx.x], a.f.y + g + 2 * q);
      lane_prefetch16(staged[1][q][threadIdx.x], a.f.k1 + g + 2 * q);
    }
  };
  if (warp < n_tiles && interior_tile(warp)) prefetch(warp);
  lane_prefetch_commit();
  for (size_t tile = warp; tile < n_tiles; tile += n_warps) {
    const size_t tile0 = tile * (size_t)OUT;                // first stored element of this tile
    double y[E], k[E][S], in[E];
    const bool interior = interior_tile(tile);
    if (interior) {                                         // the lane's bytes were prefetched while the previous tile was evaluated
      lane_prefetch_wait();
#pragma unroll
      for (int e = 0; e < E; ++e) { y[e] = staged[0][e >> 1][threadIdx.x][e & 1]; k[e][0] = staged[1][e >> 1][threadIdx.x][e & 1]; }
    }
    {
      const size_t nxt = tile + n_warps;
      if (nxt < n_tiles && interior_tile(nxt)) prefetch(nxt);
      lane_prefetch_commit();
    }
    if (!interior) {                                        // first / last tiles: around the ring, or into the neighbouring shards
#pragma unroll
      for (int e = 0; e < E; ++e) {
        y[e] = l96_edge_load<HL, HR>(a.f.y, a.halo.left_y, a.halo.right_y, n, tile0, p + e);
        k[e][0] = l96_edge_load<HL, HR>(a.f.k1, a.halo.left_k, a.halo.right_k, n, tile0, p + e);
      }
    }
#pragma unroll
    for (int e = 0; e < E; ++e)
#pragma unroll
      for (int j = 1; j < S; ++j) k[e][j] = 0.0;
    L96WarpStages<PAT, S, E, NEG>::run(y, k, in, a);
#pragma unroll
    for (int h = 0; h < E / 4; ++h) {                       // HL, OUT multiples of 4: a group of 4 positions is stored whole or not at all (up to n)
      const int ph = p + 4 * h;
      const size_t g = tile0 + (size_t)(ph - HL);           // meaningful when the group is stored
      const bool stored = ph >= HL && ph < HL + OUT && g < n;
      Pk<4> yo, ko;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int e = 4 * h + q;
        ko.v[q] = k[e][S - 1];
        if (Pattern<PAT>::last) yo.v[q] = in[e];
        else yo.v[q] = __dadd_rn(y[e], __dmul_rn(const_wsum<S, Pattern<PAT>::b()>(k[e], a.f.b), a.f.cb));
        const double lo = __dmul_rn(const_wsum<S, Pattern<PAT>::bh()>(k[e], a.f.bh), a.f.cbh);
        double err;
        if (Pattern<PAT>::direct) err = lo;
        else err = __dadd_rn(yo.v[q], -__dadd_rn(y[e], lo));
        const double tol = __dadd_rn(a.f.absTol, __dmul_rn(fabs(yo.v[q]), a.f.relTol));
        const double r = err_ratio(err, tol);
        if (stored && g + q < n) acc = __dadd_rn(acc, __dmul_rn(r, r));
      }
      if (stored) {
        if (g + 4 <= n) {
          st_stream<4>(a.f.ynew + g, yo);
          st_stream<4>(a.f.ks_out + g, ko);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (g + q < n) { a.f.ynew[g + q] = yo.v[q]; a.f.ks_out[g + q] = ko.v[q]; }
        }
      }
    }
  }
#ifdef B200RK_EMULATE_SERIAL_SUM
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
    for (int j = 0; j < J; ++j) l96_pair<NEG>(sh[pos[j] - 2], sh[pos[j] - 1], sh[pos[j] + 2], in[2 * j], in[2 * j + 1], a.F, k[s][2 * j], k[s][2 * j + 1]);
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

// ===================================================================================================================
// STENCIL RIGHT-HAND SIDES FROM SOURCE (compiled at run time only — jit.cu)
//   dydt[i] = expr(t, Y(-RL) .. Y(+RR), p0[i] .., c0 ..)      on the cyclic index space of the state vector,
// Y(d) = y[(i + d) mod N]. jit.cu generates
//   __device__ double b200rk::user_stencil(double t, const double* b200rk_y_, const double* p, const double* c)
// (with `#define Y(d) b200rk_y_[(d)]` around the caller's expression) plus B200RK_STENCIL_RL / _RR / B200RK_USER_NP and
// compiles the two kernels below around it: the plain ODEProc evaluation every method can call through the stage / RHS /
// finish pipeline, and the whole-attempt kernel of the FSAL pairs over overlapped tiles — the tiling of l96_attempt_kernel
// with the overlap the stencil's radii ask for: HL = RL*(S-1), HR = RR*(S-1), rounded to 32 bytes.
// ===================================================================================================================
constexpr int kMaxStencilRadius = 8;   // per side; (RL + RR) * 8 evaluations must leave most of a 1024-wide tile to store

struct UStencilRhsArgs {
  const double* y;
  const double* p[kMaxUserVecs];
  double cs[kMaxUserScalars];
  double t;            // the time handed to f
  double* out;
  size_t n;            // this rank's block of the cyclic vector
  const double *left, *right;   // sharded: the RL elements before the block (left[0] = element -RL) and the RR after it; null on one GPU
};

template <int S>
struct UStencilAttemptArgs {
  FusedArgs<S> f;     // y, k1, p[], cs[], t, tsign, rhs_sign, cnode, rows, dt, tolerances, ynew, ks_out, n, rs
  L96Halo halo;       // sharded: HL elements before / HR after the block, of y and of k1 (all null on a single GPU)
};

struct UStencilRk4Args {   // one RK4 step (ode.nim:180-189) of a stencil right-hand side from source in one kernel
  const double* y;
  double* ynew;
  size_t n;
  const double* p[kMaxUserVecs];
  double cs[kMaxUserScalars];
  double t, tsign, rsign;   // step start time (solver coordinates); -1 / -1 for the backward pass g(t, y) = -f(-t, y)
  double hdt, dt, c6;       // 0.5*dt, dt, dt/6.0 computed by the host
  L96Halo halo;             // sharded: left_y = the 4*RL (rounded) elements before the block, right_y = the 4*RR after it
};

// overlap of the whole-attempt tiles for radii (rl, rr) and S stages, rounded to 32 bytes (host and device use the same rule)
__host__ __device__ constexpr int stencil_halo(int radius, int S) { return ((radius * (S - 1) + 3) / 4) * 4; }

#ifdef B200RK_JIT_STENCIL
constexpr int kStencilRL = B200RK_STENCIL_RL, kStencilRR = B200RK_STENCIL_RR;
constexpr int kStencilPadL = ((kStencilRL + 1) / 2) * 2, kStencilPadR = ((kStencilRR + 1) / 2) * 2;   // keeps every pair 16-byte aligned

template <int S>
struct UStencilTile {
  static constexpr int HL = stencil_halo(kStencilRL, S);
  static constexpr int HR = stencil_halo(kStencilRR, S);
};

// dydt = f(t, y): a tile of 4 * THREADS elements plus its RL + RR neighbours staged in shared memory, then one evaluation per element.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) user_stencil_rhs_kernel(const UStencilRhsArgs a) {
  constexpr int W = 4, TILE = THREADS * W, NP = PwTraits<PW_USER>::NP, NPX = PwTraits<PW_USER>::NPX;
  __shared__ double sh_[kStencilPadL + TILE + kStencilPadR];
  double* sh = sh_ + kStencilPadL;
  const size_t n = a.n;
  for (size_t tile0 = (size_t)blockIdx.x * TILE; tile0 < n; tile0 += (size_t)gridDim.x * TILE) {
    const size_t len = (tile0 + TILE <= n) ? (size_t)TILE : n - tile0;
    auto at = [&](long long g) -> double {   // element g of the block, g in [-RL, n + RR)
      if (g >= 0 && (size_t)g < n) return a.y[g];
      if (a.left) return g < 0 ? a.left[kStencilRL + g] : a.right[(size_t)g - n];
      return a.y[(size_t)((g % (long long)n + (long long)n) % (long long)n)];
    };
    const size_t i0 = tile0 + (size_t)threadIdx.x * W;
    if (i0 + W <= n) {
      const Pk<W> v = ld_stream<W>(a.y + i0);
#pragma unroll
      for (int e = 0; e < W; ++e) sh[threadIdx.x * W + e] = v.v[e];
    } else {
#pragma unroll
      for (int e = 0; e < W; ++e) sh[threadIdx.x * W + e] = (i0 + e < n) ? a.y[i0 + e] : 0.0;
    }
    if ((int)threadIdx.x < kStencilRL) sh[-kStencilRL + (int)threadIdx.x] = at((long long)tile0 - kStencilRL + threadIdx.x);
    __syncthreads();
    // the right halo goes directly behind the last valid element (a ragged last tile); written after the barrier so that it cannot
    // race with a thread storing a padded zero into the same slot
    if ((int)threadIdx.x < kStencilRR) sh[len + threadIdx.x] = at((long long)(tile0 + len) + threadIdx.x);
    __syncthreads();
    if (i0 < n) {
      double o[W];
#pragma unroll
      for (int e = 0; e < W; ++e) {
        double pe[NPX] = {};
        if (i0 + e < n) {
#pragma unroll
          for (int j = 0; j < NP; ++j) pe[j] = a.p[j][i0 + e];
          o[e] = user_stencil(a.t, sh + threadIdx.x * W + e, pe, a.cs);
        } else o[e] = 0.0;
      }
      if (i0 + W <= n) {
        Pk<W> ov;
#pragma unroll
        for (int e = 0; e < W; ++e) ov.v[e] = o[e];
        st_stream<W>(a.out + i0, ov);
      } else {
#pragma unroll
        for (int e = 0; e < W; ++e)
          if (i0 + e < n) a.out[i0 + e] = o[e];
      }
    }
    __syncthreads();   // the tile is rewritten by the next iteration
  }
}

template <int PAT, int s, int J, int TW>
struct UStencilStages {
  template <int S>
  __device__ __forceinline__ static void run(const double (&y)[2 * J], double (&k)[2 * J][S], double (&in)[2 * J], double (&part)[2 * J],
                                             double (&part_bh)[2 * J], const double (&pe)[2 * J][PwTraits<PW_USER>::NPX], const int (&pos)[J],
                                             double (*buf)[kStencilPadL + TW + kStencilPadR], const UStencilAttemptArgs<S>& a) {
    if constexpr (s > 2) UStencilStages<PAT, s - 1, J, TW>::run(y, k, in, part, part_bh, pe, pos, buf, a);
    double* sh = buf[s & 1] + kStencilPadL;   // the pads stay zero: their consumers are outside the stored range
#pragma unroll
    for (int e = 0; e < 2 * J; ++e) {
      const double acc = const_wsum_finish<S, Pattern<PAT>::a(s - 2), s - 2>(part[e], k[e], a.f.a[s - 2]);
      in[e] = __dadd_rn(y[e], __dmul_rn(acc, a.f.dt));                                   // stage_elem
    }
#pragma unroll
    for (int j = 0; j < J; ++j) { sh[pos[j]] = in[2 * j]; sh[pos[j] + 1] = in[2 * j + 1]; }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 2 * J; ++e) {
      if constexpr (s < S) part[e] = const_wsum_prefix<S, Pattern<PAT>::a(s - 1), s - 1>(k[e], a.f.a[s - 1]);
      else {
        if constexpr (!Pattern<PAT>::last) part[e] = const_wsum_prefix<S, Pattern<PAT>::b(), S - 1>(k[e], a.f.b);
        part_bh[e] = const_wsum_prefix<S, Pattern<PAT>::bh(), S - 1>(k[e], a.f.bh);
      }
    }
    // k_s = g(t + dt*c_s, stage input) with g = f forward and g(t, y) = -f(-t, y) backward (ode.nim:545): the signs are exact flips
    const double ts = flip_sign_by(__dadd_rn(a.f.t, __dmul_rn(a.f.dt, a.f.cnode[s - 1])), a.f.tsign);
    // the neighbourhood of a pair — positions p - RL .. p + 1 + RR — is pulled into registers once with 128-bit loads (the window
    // starts at an even offset, every pair is 16-byte aligned) and both evaluations index it with compile-time offsets
    constexpr int WL = kStencilPadL, WLEN = kStencilPadL + 2 + kStencilPadR;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      double w[WLEN];
#pragma unroll
      for (int q = 0; q < WLEN; q += 2) {
        const double2 v = *reinterpret_cast<const double2*>(sh + pos[j] - WL + q);
        w[q] = v.x; w[q + 1] = v.y;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h)
        k[2 * j + h][s - 1] = flip_sign_by(user_stencil(ts, w + WL + h, pe[2 * j + h], a.f.cs), a.f.rhs_sign);
    }
  }
};

template <int PAT, int J, int THREADS>
__global__ void __launch_bounds__(THREADS, (Pattern<PAT>::S > 7 || PwTraits<PW_USER>::NP > 0) ? 2 : 3) ustencil_attempt_kernel(const UStencilAttemptArgs<Pattern<PAT>::S> a) {
  constexpr int S = Pattern<PAT>::S;
  constexpr int E = 2 * J, TW = E * THREADS;
  constexpr int HL = UStencilTile<S>::HL, HR = UStencilTile<S>::HR, OUT = TW - HL - HR;
  constexpr int NP = PwTraits<PW_USER>::NP, NPX = PwTraits<PW_USER>::NPX;
  static_assert(OUT >= TW / 2, "stencil radii too large for the tile");
  __shared__ double buf[2][kStencilPadL + TW + kStencilPadR];
  alignas(128) __shared__ double staged[2][TW];
  alignas(8) __shared__ unsigned long long tile_bar;
  const size_t n = a.f.n;
  const size_t n_tiles = (n + OUT - 1) / OUT;
  for (int q = threadIdx.x; q < kStencilPadL + kStencilPadR; q += THREADS) {
    const int at = q < kStencilPadL ? q : TW + q;
    buf[0][at] = 0.0; buf[1][at] = 0.0;
  }
  auto interior_tile = [&](size_t t) { const size_t t0 = t * (size_t)OUT; return t0 >= (size_t)HL && t0 - HL + TW <= n; };
  size_t tile = blockIdx.x;
  if (threadIdx.x == 0) {
    tile_barrier_init(&tile_bar);
    if (tile < n_tiles && interior_tile(tile))
      tile_prefetch(staged[0], a.f.y + (tile * OUT - HL), staged[1], a.f.k1 + (tile * OUT - HL), TW * 8u, &tile_bar);
  }
  __syncthreads();
  unsigned int phase = 0;
  double acc = 0.0;
  int pos[J];
#pragma unroll
  for (int j = 0; j < J; ++j) pos[j] = 2 * ((int)threadIdx.x + j * THREADS);
  for (; tile < n_tiles; tile += gridDim.x) {
    const size_t tile0 = tile * (size_t)OUT;
    double y[E], k[E][S], in[E], pe[E][NPX];
    const bool interior = interior_tile(tile);
    if (interior) { tile_wait(&tile_bar, phase); phase ^= 1u; }
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int p = pos[j];
      if (interior) {
        const double2 yv = *reinterpret_cast<const double2*>(&staged[0][p]), kv = *reinterpret_cast<const double2*>(&staged[1][p]);
        y[2 * j] = yv.x; y[2 * j + 1] = yv.y;
        k[2 * j][0] = kv.x; k[2 * j + 1][0] = kv.y;
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          y[2 * j + h] = l96_edge_load<HL, HR>(a.f.y, a.halo.left_y, a.halo.right_y, n, tile0, p + h);
          k[2 * j + h][0] = l96_edge_load<HL, HR>(a.f.k1, a.halo.left_k, a.halo.right_k, n, tile0, p + h);
        }
      }
      // per-element parameters of the thread's own positions (positions outside the block are never stored: any in-range value will do)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        long long g = (long long)tile0 - HL + p + h;
        if (a.halo.left_y) g = g < 0 ? 0 : ((size_t)g >= n ? (long long)n - 1 : g);
        else g = ((g % (long long)n) + (long long)n) % (long long)n;
#pragma unroll
        for (int q = 0; q < NPX; ++q) pe[2 * j + h][q] = q < NP ? a.f.p[q][g] : 0.0;
      }
    }
    __syncthreads();
    {
      const size_t nxt = tile + gridDim.x;
      if (threadIdx.x == 0 && nxt < n_tiles && interior_tile(nxt))
        tile_prefetch(staged[0], a.f.y + (nxt * OUT - HL), staged[1], a.f.k1 + (nxt * OUT - HL), TW * 8u, &tile_bar);
    }
#pragma unroll
    for (int e = 0; e < E; ++e)
#pragma unroll
      for (int j = 1; j < S; ++j) k[e][j] = 0.0;
    double part[E], part_bh[E];
#pragma unroll
    for (int e = 0; e < E; ++e) { part[e] = 0.0; part_bh[e] = 0.0; }
    UStencilStages<PAT, S, J, TW>::run(y, k, in, part, part_bh, pe, pos, buf, a);
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int p = pos[j];
      const size_t g = tile0 + (size_t)(p - HL);
      const bool stored = p >= HL && p < HL + OUT && g < n;
      double yn[2], ks[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int e = 2 * j + h;
        ks[h] = k[e][S - 1];
        if (Pattern<PAT>::last) yn[h] = in[e];
        else yn[h] = __dadd_rn(y[e], __dmul_rn(const_wsum_finish<S, Pattern<PAT>::b(), S - 1>(part[e], k[e], a.f.b), a.f.cb));
        const double lo = __dmul_rn(const_wsum_finish<S, Pattern<PAT>::bh(), S - 1>(part_bh[e], k[e], a.f.bh), a.f.cbh);
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
  }
  grid_sum_finish<THREADS>(acc, a.f.rs);
}
// A whole RK4 step in one kernel, same tiling: four evaluations (k1 = f(t, y) included: RK4 has no FSAL) -> overlap 4*RL left /
// 4*RR right; reads y (+ parameters), writes yNew. Stage inputs and the final combine are fused_rk4_elem / rk4_elem, stage times
// those of user_rk4_elem.
template <int J, int THREADS>
__global__ void __launch_bounds__(THREADS) ustencil_rk4_kernel(const UStencilRk4Args a) {
  constexpr int E = 2 * J, TW = E * THREADS, HL = stencil_halo(kStencilRL, 5), HR = stencil_halo(kStencilRR, 5), OUT = TW - HL - HR;
  constexpr int NP = PwTraits<PW_USER>::NP, NPX = PwTraits<PW_USER>::NPX;
  __shared__ double buf[2][kStencilPadL + TW + kStencilPadR];
  const size_t n = a.n;
  const size_t tile0 = (size_t)blockIdx.x * OUT;
  for (int q = threadIdx.x; q < kStencilPadL + kStencilPadR; q += THREADS) {
    const int at = q < kStencilPadL ? q : TW + q;
    buf[0][at] = 0.0; buf[1][at] = 0.0;
  }
  double y[E], k[4][E], in[E], pe[E][NPX];
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
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      long long g = (long long)tile0 - HL + p + h;
      if (a.halo.left_y) g = g < 0 ? 0 : ((size_t)g >= n ? (long long)n - 1 : g);
      else g = ((g % (long long)n) + (long long)n) % (long long)n;
#pragma unroll
      for (int q = 0; q < NPX; ++q) pe[2 * j + h][q] = q < NP ? a.p[q][g] : 0.0;
    }
  }
  const double tm = __dadd_rn(a.t, __dmul_rn(a.dt, 0.5)), te = __dadd_rn(a.t, __dmul_rn(a.dt, 1.0));   // ode.nim:185-187
  constexpr int WL = kStencilPadL, WLEN = kStencilPadL + 2 + kStencilPadR;
#pragma unroll
  for (int s = 0; s < 4; ++s) {   // stage s + 1: its input, published; then k_{s+1} from the neighbourhood
    double* sh = buf[s & 1] + kStencilPadL;
    const double c = (s == 3) ? a.dt : a.hdt;
    const double ts = flip_sign_by(s == 0 ? a.t : (s == 3 ? te : tm), a.tsign);
#pragma unroll
    for (int e = 0; e < E; ++e) in[e] = (s == 0) ? y[e] : __dadd_rn(y[e], __dmul_rn(k[s - 1 < 0 ? 0 : s - 1][e], c));
#pragma unroll
    for (int j = 0; j < J; ++j) { sh[pos[j]] = in[2 * j]; sh[pos[j] + 1] = in[2 * j + 1]; }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < J; ++j) {
      double w[WLEN];
#pragma unroll
      for (int q = 0; q < WLEN; q += 2) {
        const double2 v = *reinterpret_cast<const double2*>(sh + pos[j] - WL + q);
        w[q] = v.x; w[q + 1] = v.y;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) k[s][2 * j + h] = flip_sign_by(user_stencil(ts, w + WL + h, pe[2 * j + h], a.cs), a.rsign);
    }
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
#endif  // B200RK_JIT_STENCIL

}  // namespace b200rk
