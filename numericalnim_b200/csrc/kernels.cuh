// kernels.cuh — hand-written sm_100a kernels of the explicit Runge–Kutta hot path (fp64, BLAS-1,
// HBM-bandwidth-bound; no tensor cores on purpose).
//
// What each kernel replaces in the reference (numericalnim, src/numericalnim/):
//   stage_kernel    y + c*(w1*k1 + ... + wm*km)            ode.nim:185-187, 294-299, 364-369, 455-462
//                   (2m-1 allocating Vector passes there: `*` utils.nim:176-180, `+` utils.nim:59-64)
//   finish_kernel   yNew, yLow/error_y, tolerance, scaled square and its sum
//                   ode.nim:301-303, 371-372, 464-466 + ode.nim:61-65 (34..41 Vector passes there)
//   rk4_final       y + dt/6*(k1 + 2*(k2+k3) + k4)         ode.nim:188
//   hermite_kernel  hermiteSpline                           utils.nim:273-279
//   ewise kernels   the Vector[T] operators themselves      utils.nim:59-223
//
// Parity: every product / sum uses __dmul_rn / __dadd_rn (never contracted into FMA) in exactly the
// reference's left-to-right association, so element-wise results are bit-identical to the CPU path.
// Only the error-norm reduction differs in summation order (deterministic tree vs sequential).
//
// Memory access: W doubles per thread per access (W = 2 -> 128-bit LDG/STG, W = 4 -> 256-bit
// LDG.E.256 / STG.E.256, new on sm_100), U independent accesses per input stream issued back to back
// before first use, fully coalesced, read-only non-coherent path with L1 no-allocate (every element is
// touched exactly once per launch). The Butcher row travels as kernel parameters, i.e. in the constant
// bank: one uniform broadcast read per weight, no shared-memory staging or barrier needed.
//
// This header is compiled twice: by nvcc into libb200rk.so, and at run time by NVRTC (jit.cu) when a user hands the
// library an element-local right-hand side as a source expression (PW_USER below) — the same fused kernels, with the
// user's expression inlined where the built-in right-hand sides sit. It therefore includes no host headers.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#else
typedef unsigned int uint32_t;  // NVRTC: no host headers
#endif

namespace b200rk {

constexpr int kMaxTerms = 9;  // Vern65 has 9 stages

// ---------------------------------------------------------------------------------------------------
// W-wide global memory access
// ---------------------------------------------------------------------------------------------------
template <int W>
struct Pk {
  double v[W];
};

#ifndef B200RK_HOST_EMULATION
template <int W>
__device__ __forceinline__ Pk<W> ld_stream(const double* p);
template <>
__device__ __forceinline__ Pk<1> ld_stream<1>(const double* p) {
  Pk<1> r;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r.v[0]) : "l"(p));
  return r;
}
template <>
__device__ __forceinline__ Pk<2> ld_stream<2>(const double* p) {
  Pk<2> r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.v[0]), "=d"(r.v[1]) : "l"(p));
  return r;
}
template <>
__device__ __forceinline__ Pk<4> ld_stream<4>(const double* p) {
  Pk<4> r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3])
               : "l"(p));
  return r;
}
// coherent variant for kernels whose output may alias an input (in-place Vector ops) or that re-read buffers
// they rewrite themselves (the device-resident loop): plain ld.global (no .nc), still one 128/256-bit access
template <int W>
__device__ __forceinline__ Pk<W> ld_plain(const double* p) {
  Pk<W> r = {};
  if constexpr (W == 4) {
    asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p) : "memory");
  } else if constexpr (W == 2) {
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.v[0]), "=d"(r.v[1]) : "l"(p) : "memory");
  } else {
    r.v[0] = *p;
  }
  return r;
}
template <int W>
__device__ __forceinline__ void st_stream(double* p, const Pk<W>& x);
template <>
__device__ __forceinline__ void st_stream<1>(double* p, const Pk<1>& x) {
  asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(x.v[0]) : "memory");
}
template <>
__device__ __forceinline__ void st_stream<2>(double* p, const Pk<2>& x) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(x.v[0]), "d"(x.v[1]) : "memory");
}
template <>
__device__ __forceinline__ void st_stream<4>(double* p, const Pk<4>& x) {
  asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(x.v[0]), "d"(x.v[1]),
               "d"(x.v[2]), "d"(x.v[3])
               : "memory");
}

// L2 eviction-priority variants (SASS: LDG.E.NA.EFL2 / STG.E.NA.ELL2). The 126 MB L2 can hold one 64 MiB
// vector: a producer stores its output with evict_last while streaming its inputs with evict_first, so the
// consumer kernel launched next finds that vector in L2 instead of HBM (stage input -> RHS, k_s -> next stage).
enum L2Policy : int { L2_NORMAL = 0, L2_EVICT_FIRST = 1, L2_EVICT_LAST = 2 };

template <int W, int POL>
__device__ __forceinline__ Pk<W> ld_pol(const double* p) {
  Pk<W> r = {};
  if (POL == L2_NORMAL) return ld_stream<W>(p);
  if constexpr (W == 4) {
    if (POL == L2_EVICT_FIRST)
      asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];"
                   : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[W > 2 ? 2 : 0]), "=d"(r.v[W > 3 ? 3 : 0]) : "l"(p));
    else
      asm volatile("ld.global.nc.L1::no_allocate.L2::evict_last.v4.f64 {%0,%1,%2,%3}, [%4];"
                   : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[W > 2 ? 2 : 0]), "=d"(r.v[W > 3 ? 3 : 0]) : "l"(p));
  } else {
    return ld_stream<W>(p);  // ptxas accepts the .L2:: eviction modifiers on 256-bit accesses only
  }
  return r;
}
template <int W, int POL>
__device__ __forceinline__ void st_pol(double* p, const Pk<W>& x) {
  if (POL == L2_NORMAL) { st_stream<W>(p, x); return; }
  if constexpr (W == 4) {
    if (POL == L2_EVICT_FIRST)
      asm volatile("st.global.L1::no_allocate.L2::evict_first.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(x.v[0]), "d"(x.v[1]),
                   "d"(x.v[W > 2 ? 2 : 0]), "d"(x.v[W > 3 ? 3 : 0]) : "memory");
    else
      asm volatile("st.global.L1::no_allocate.L2::evict_last.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(x.v[0]), "d"(x.v[1]),
                   "d"(x.v[W > 2 ? 2 : 0]), "d"(x.v[W > 3 ? 3 : 0]) : "memory");
  } else {
    st_stream<W>(p, x);
  }
}
#else
// Host emulation (tests/host_emul): the same kernels compiled by the host compiler and run one emulated thread at a
// time, so that the CPU test-suite can check the kernels' arithmetic and indexing against the oracle without a GPU.
// Only the memory-access helpers (inline PTX above) need a host form (plus, for the single-threaded emulation, the
// grid-wide reduction: B200RK_EMULATE_SERIAL_SUM).
enum L2Policy : int { L2_NORMAL = 0, L2_EVICT_FIRST = 1, L2_EVICT_LAST = 2 };
template <int W> inline Pk<W> ld_stream(const double* p) { Pk<W> r; for (int e = 0; e < W; ++e) r.v[e] = p[e]; return r; }
template <int W> inline Pk<W> ld_plain(const double* p) { return ld_stream<W>(p); }
template <int W> inline void st_stream(double* p, const Pk<W>& x) { for (int e = 0; e < W; ++e) p[e] = x.v[e]; }
template <int W, int POL> inline Pk<W> ld_pol(const double* p) { return ld_stream<W>(p); }
template <int W, int POL> inline void st_pol(double* p, const Pk<W>& x) { st_stream<W>(p, x); }
#endif

// ---------------------------------------------------------------------------------------------------
// Stage accumulate:  out = y + c * (w0*k0 + w1*k1 + ... )      (left-associated, no FMA)
//   CHAIN form (Kutta3's third stage, ode.nim:128):  out = ((y + w0*k0) + w1*k1) + ...   (c unused)
// ---------------------------------------------------------------------------------------------------
template <int M>
struct StageArgs {
  const double* y;
  const double* k[M];
  double w[M];
  double c;
  double* out;
  size_t n;
};

template <int M, bool CHAIN>
__device__ __forceinline__ double stage_elem(double y, const double (&k)[M], const double (&w)[M], double c) {
  if (CHAIN) {
    double acc = y;
#pragma unroll
    for (int j = 0; j < M; ++j) acc = __dadd_rn(acc, __dmul_rn(k[j], w[j]));
    return acc;
  }
  double acc = __dmul_rn(k[0], w[0]);
#pragma unroll
  for (int j = 1; j < M; ++j) acc = __dadd_rn(acc, __dmul_rn(k[j], w[j]));
  return __dadd_rn(y, __dmul_rn(acc, c));
}

template <int M, int W, int U, bool CHAIN, int THREADS, int L2 = 0>
__global__ void __launch_bounds__(THREADS) stage_kernel(const StageArgs<M> a) {
  constexpr int LDP = L2 ? L2_EVICT_FIRST : L2_NORMAL, STP = L2 ? L2_EVICT_LAST : L2_NORMAL;
  const size_t nvec = a.n / W;
  const size_t tile = (size_t)THREADS * U;
  for (size_t base = (size_t)blockIdx.x * tile; base < nvec; base += (size_t)gridDim.x * tile) {
    Pk<W> yv[U], kv[U][M];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t v = base + (size_t)u * THREADS + threadIdx.x;
      if (v < nvec) {
        yv[u] = ld_pol<W, LDP>(a.y + v * W);
#pragma unroll
        for (int j = 0; j < M; ++j) kv[u][j] = ld_pol<W, LDP>(a.k[j] + v * W);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t v = base + (size_t)u * THREADS + threadIdx.x;
      if (v < nvec) {
        Pk<W> o;
#pragma unroll
        for (int e = 0; e < W; ++e) {
          double ke[M];
#pragma unroll
          for (int j = 0; j < M; ++j) ke[j] = kv[u][j].v[e];
          o.v[e] = stage_elem<M, CHAIN>(yv[u].v[e], ke, a.w, a.c);
        }
        st_pol<W, STP>(a.out + v * W, o);
      }
    }
  }
  // ragged tail (n % W elements)
  if (blockIdx.x == 0) {
    const size_t i = nvec * W + threadIdx.x;
    if (i < a.n) {
      double ke[M];
#pragma unroll
      for (int j = 0; j < M; ++j) ke[j] = a.k[j][i];
      a.out[i] = stage_elem<M, CHAIN>(a.y[i], ke, a.w, a.c);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// RK4 final combine (ode.nim:188):  y + (dt/6) * ((k1 + 2*(k2 + k3)) + k4), c6 = dt/6.0 from the host
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double rk4_elem(double y, double k1, double k2, double k3, double k4, double c6) {
  const double s = __dadd_rn(__dadd_rn(k1, __dmul_rn(__dadd_rn(k2, k3), 2.0)), k4);
  return __dadd_rn(y, __dmul_rn(s, c6));
}
template <int W, int U, int THREADS>
__global__ void __launch_bounds__(THREADS)
    rk4_final_kernel(const double* __restrict__ y, const double* __restrict__ k1, const double* __restrict__ k2,
                     const double* __restrict__ k3, const double* __restrict__ k4, double c6, double* __restrict__ out,
                     size_t n) {
  const size_t nvec = n / W;
  const size_t tile = (size_t)THREADS * U;
  for (size_t base = (size_t)blockIdx.x * tile; base < nvec; base += (size_t)gridDim.x * tile) {
    Pk<W> a[U][5];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t v = base + (size_t)u * THREADS + threadIdx.x;
      if (v < nvec) {
        a[u][0] = ld_stream<W>(y + v * W);
        a[u][1] = ld_stream<W>(k1 + v * W);
        a[u][2] = ld_stream<W>(k2 + v * W);
        a[u][3] = ld_stream<W>(k3 + v * W);
        a[u][4] = ld_stream<W>(k4 + v * W);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t v = base + (size_t)u * THREADS + threadIdx.x;
      if (v < nvec) {
        Pk<W> o;
#pragma unroll
        for (int e = 0; e < W; ++e)
          o.v[e] = rk4_elem(a[u][0].v[e], a[u][1].v[e], a[u][2].v[e], a[u][3].v[e], a[u][4].v[e], c6);
        st_stream<W>(out + v * W, o);
      }
    }
  }
  if (blockIdx.x == 0) {
    const size_t i = nvec * W + threadIdx.x;
    if (i < n) out[i] = rk4_elem(y[i], k1[i], k2[i], k3[i], k4[i], c6);
  }
}

// ---------------------------------------------------------------------------------------------------
// Deterministic sum reduction: warp shuffle -> shared -> one partial per block -> the last block to
// finish (atomic ticket) adds the partials in index order. No floating-point atomics anywhere.
// ---------------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 16;

// Peer mailboxes for the sharded error norm: the collective is fused INTO the reducing kernel. Every rank
// owns a small device buffer that all peers have mapped (CUDA IPC, NVLink/NVSwitch peer stores); the last
// CTA of a rank's kernel stores its shard's partial sum and the launch sequence number into slot [rank] of
// every peer's mailbox, then waits until the `world` slots of its own mailbox carry this sequence number
// and adds them in rank order — the same order on every rank, so all ranks obtain the identical bits and
// take the identical accept/reject decision. 8 bytes per peer per attempt: pure latency, no NCCL launch,
// no separate kernel, no host round trip beyond the read-back the single-GPU path already has.
struct PeerMail {
  int world;   // <= 1: no exchange
  int rank;
  long long timeout_cycles;            // give up waiting for a peer after this many SM clocks (context knob "peer_timeout_s")
  unsigned long long* box[kMaxPeers];  // box[p] = base of rank p's mailbox: [2 parities][kMaxPeers][2] words
};
constexpr unsigned long long kPeerTimeoutFlag = 1ull << 63;

// A mailbox slot is two 8-byte words; each carries 32 bits of the fp64 payload under a 32-bit tag derived from the
// attempt's sequence number: word_h = (tag << 32) | half_h. An aligned 8-byte store is single-copy atomic, so a reader
// that finds the current tag in BOTH words holds the complete value whatever order the two stores arrive in — there is no
// separate flag and therefore no fence between payload and flag: one 16-byte store per peer, one NVLink traversal per
// exchange (the previous protocol — value, system fence, flag — paid a store round trip before the flag could leave).
// A slot is reused two attempts later (parity double buffer), when its old tag (seq - 2) can no longer match.
__device__ __forceinline__ unsigned int mail_tag(unsigned long long seq) { return (unsigned int)(seq & 0x7fffffffull) | 0x80000000u; }
__device__ __forceinline__ unsigned long long* mail_slot(unsigned long long* box, unsigned long long seq, int src_rank) {
  return box + ((((seq & 1ull) * kMaxPeers) + (unsigned)src_rank) << 1);
}
#ifndef B200RK_HOST_EMULATION
__device__ __forceinline__ void mail_put(unsigned long long* slot, double v, unsigned int tag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v), t = (unsigned long long)tag << 32;
  asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1,%2};" ::"l"(slot), "l"(t | (b & 0xffffffffull)), "l"(t | (b >> 32)) : "memory");
}
__device__ __forceinline__ bool mail_try_get(const unsigned long long* slot, unsigned int tag, double* v) {
  unsigned long long w0, w1;
  asm volatile("ld.relaxed.sys.global.v2.u64 {%0,%1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
  if ((unsigned int)(w0 >> 32) != tag || (unsigned int)(w1 >> 32) != tag) return false;
  *v = __longlong_as_double((long long)((w1 << 32) | (w0 & 0xffffffffull)));
  return true;
}
#else
inline void mail_put(unsigned long long* slot, double v, unsigned int tag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v), t = (unsigned long long)tag << 32;
  __atomic_store_n(slot, t | (b & 0xffffffffull), __ATOMIC_RELAXED);
  __atomic_store_n(slot + 1, t | (b >> 32), __ATOMIC_RELAXED);
}
inline bool mail_try_get(const unsigned long long* slot, unsigned int tag, double* v) {
  const unsigned long long w0 = __atomic_load_n(slot, __ATOMIC_RELAXED), w1 = __atomic_load_n(slot + 1, __ATOMIC_RELAXED);
  if ((unsigned int)(w0 >> 32) != tag || (unsigned int)(w1 >> 32) != tag) return false;
  *v = __longlong_as_double((long long)((w1 << 32) | (w0 & 0xffffffffull)));
  return true;
}
#endif
// threads q < world of a CTA: wait for rank q's slot of the LOCAL mailbox; false on timeout
__device__ __forceinline__ bool mail_wait(const PeerMail& mail, unsigned long long seq, int q, double* v) {
  const unsigned long long* src = mail_slot(mail.box[mail.rank], seq, q);
  const unsigned int tag = mail_tag(seq);
  if (mail_try_get(src, tag, v)) return true;
  const long long t0 = clock64();
  while (!mail_try_get(src, tag, v)) {
    if (clock64() - t0 > mail.timeout_cycles) return false;   // a peer died or never launched: report instead of hanging
  }
  return true;
}

struct ReduceScratch {
  double* partials;        // gridDim.x doubles
  unsigned int* ticket;    // zero before launch; reset to zero by the last block
  double* result;          // device scalar
  double* result_host;     // optional zero-copy mirror (mapped pinned host memory), may be null
  unsigned long long* seq_host;  // mapped pinned host word: set to `seq` after result_host is visible
  unsigned long long seq;        // launch sequence number the host spins on (no cudaStreamSynchronize)
  PeerMail mail;
};

// Runs in the last CTA only (all threads): exchange `local` with the peers, return the global sum in thread 0.
template <int THREADS>
__device__ __forceinline__ double peer_allreduce(double local, const ReduceScratch& rs, bool* timed_out) {
  __shared__ double peer_vals[kMaxPeers];
  __shared__ int peer_bad;
  const int q = threadIdx.x;
  const int world = rs.mail.world;
  if (q == 0) peer_bad = 0;
  __syncthreads();
  if (q < world) {
    // Everything this rank's kernel stored (a ring neighbour may read yNew / k_S in place in its next attempt) is
    // visible system-wide before a peer can learn the sum: one fence, in front of the publication.
    __threadfence_system();
    mail_put(mail_slot(rs.mail.box[q], rs.seq, rs.mail.rank), local, mail_tag(rs.seq));   // q == rank: a local store
    double v = 0.0;
    const bool ok = mail_wait(rs.mail, rs.seq, q, &v);
    peer_vals[q] = ok ? v : 0.0;
    if (!ok) atomicExch(&peer_bad, 1);
  }
  __syncthreads();
  double total = 0.0;
  if (q == 0) {
    for (int p = 0; p < world; ++p) total = __dadd_rn(total, peer_vals[p]);  // rank order: identical on every rank
    *timed_out = peer_bad != 0;
  }
  return total;
}

template <int THREADS>
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double warp_part[THREADS / 32];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = __dadd_rn(v, __shfl_down_sync(0xffffffffu, v, off));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();  // protect warp_part against a previous use
  if (lane == 0) warp_part[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) {
    r = (lane < THREADS / 32) ? warp_part[lane] : 0.0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) r = __dadd_rn(r, __shfl_down_sync(0xffffffffu, r, off));
  }
  return r;  // valid in thread 0
}

#ifdef B200RK_EMULATE_SERIAL_SUM
// Host emulation only (tests/host_emul): what peer_allreduce does, by one host thread — ranks are threads of one process
// there, and the emulated device is handed to the other ranks' kernels while this one waits for their partials.
#ifndef B200RK_EMUL_DEVICE_RELEASE
#define B200RK_EMUL_DEVICE_RELEASE()
#define B200RK_EMUL_DEVICE_ACQUIRE()
#endif
inline double emul_peer_exchange(double local, const ReduceScratch& rs) {
  const int world = rs.mail.world, rank = rs.mail.rank;
  __threadfence_system();
  for (int q = 0; q < world; ++q) mail_put(mail_slot(rs.mail.box[q], rs.seq, rank), local, mail_tag(rs.seq));
  B200RK_EMUL_DEVICE_RELEASE();
  double total = 0.0;
  for (int q = 0; q < world; ++q) {
    double v = 0.0;
    while (!mail_try_get(mail_slot(rs.mail.box[rank], rs.seq, q), mail_tag(rs.seq), &v)) __threadfence_system();
    total = __dadd_rn(total, v);   // rank order, as in peer_allreduce
  }
  __threadfence_system();
  B200RK_EMUL_DEVICE_ACQUIRE();
  return total;
}
#endif

template <int THREADS>
__device__ __forceinline__ void grid_sum_finish(double thread_val, const ReduceScratch& rs) {
#ifdef B200RK_EMULATE_SERIAL_SUM
  // serial host emulation (threads run one after another, CTA 0 / thread 0 first): a plain running sum; the last
  // thread publishes it the way the last CTA does
  if (blockIdx.x == 0 && threadIdx.x == 0) *rs.result = 0.0;
  *rs.result = __dadd_rn(*rs.result, thread_val);
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == THREADS - 1 && rs.result_host) {
    if (rs.mail.world > 1) *rs.result = emul_peer_exchange(*rs.result, rs);
    *rs.result_host = *rs.result;
    *rs.seq_host = rs.seq;
  }
  return;
#endif
  __shared__ bool is_last;
  const double bsum = block_sum<THREADS>(thread_val);
  if (threadIdx.x == 0) {
    rs.partials[blockIdx.x] = bsum;
    __threadfence();
    const unsigned int t = atomicAdd(rs.ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double acc = 0.0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += THREADS) acc = __dadd_rn(acc, __ldcg(rs.partials + i));
    double total = block_sum<THREADS>(acc);
    bool timed_out = false;
    if (rs.mail.world > 1) {  // sharded: all-reduce the scalar over the peers' mailboxes, inside this kernel
      __shared__ double local_total;
      if (threadIdx.x == 0) local_total = total;
      __syncthreads();
      total = peer_allreduce<THREADS>(local_total, rs, &timed_out);
    }
    if (threadIdx.x == 0) {
      *rs.result = total;
      *rs.ticket = 0u;
      if (rs.result_host) {
        *(volatile double*)rs.result_host = total;
        __threadfence_system();  // the value is visible to the host before the sequence word
        *(volatile unsigned long long*)rs.seq_host = timed_out ? (rs.seq | kPeerTimeoutFlag) : rs.seq;
      }
      __threadfence_system();
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Finish: solution combine + embedded error + scaled RMS accumulation of one adaptive attempt.
//   yNew  = y + cb  * (Σ_{j in mask_b}  wb[j]*k[j])                         ode.nim:301/371/464
//   e     = DIRECT ? cbh * (Σ wbh[j]*k[j])                                   ode.nim:372
//                  : yNew - (y + cbh * (Σ_{j in mask_bh} wbh[j]*k[j]))       ode.nim:302-303/465-466
//   tol   = absTol + relTol*|yNew| ; r = e/tol ; S += r*r                    ode.nim:61-65
// NK = number of distinct k streams loaded. YNEW_MODE: 0 = recompute in registers, do not store (the
// stage-S input buffer already holds the identical bits: DOPRI54/Tsit54/BS32), 1 = compute and store
// (Vern65, RK21), 2 = load yNew instead of y (Tsit54: y itself is not needed once yNew is known).
// ---------------------------------------------------------------------------------------------------
template <int NK>
struct FinishArgs {
  const double* y;       // y, or yNew when YNEW_MODE == 2
  const double* k[NK];
  double wb[NK], wbh[NK];
  uint32_t mask_b, mask_bh;
  double cb, cbh, absTol, relTol;
  double* ynew_out;      // YNEW_MODE == 1
  double* err_out;       // optional: element-wise error_y for tests (may be null)
  size_t n;
  ReduceScratch rs;
};

// r = e / tol of the scaled error norm (ode.nim:62-63). r feeds ONLY the sum of squares, which is compared with the
// reference at a stated tolerance (its summation order differs anyway) — never bit for bit; yNew, error_y and k_S do not
// pass through here. The IEEE division costs ~14 fp64-pipe issue slots per element in kernels that are fp64-issue bound
// (no FMA allowed elsewhere), so r is formed as e * (1/tol) with the reciprocal from MUFU.RCP64H refined by two Newton
// steps (error 2^-20 -> 2^-40 -> below 2^-53): 6 slots, |relative error| <= ~2 ulp. tol = absTol + relTol*|yNew| is a
// positive normal number in any sane configuration; everything else (zero tolerances, overflow, NaN) takes the exact
// division so that inf / NaN outcomes are those of the reference expression.
__device__ __forceinline__ double err_ratio(double e, double tol) {
#ifdef B200RK_HOST_EMULATION
  return __ddiv_rn(e, tol);
#else
  const unsigned int hi = (unsigned int)__double2hiint(tol);
  if ((hi - 0x00300000u) < 0x7fa00000u) {   // 2^-1020 <= tol < 2^1021 and the sign bit clear: 1/tol is normal too
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(tol));
    double t = __fma_rn(-tol, x, 1.0);
    x = __fma_rn(x, t, x);
    t = __fma_rn(-tol, x, 1.0);
    x = __fma_rn(x, t, x);
    return __dmul_rn(e, x);
  }
  return __ddiv_rn(e, tol);
#endif
}

// v with its sign flipped when sgn is negative (sgn = +-1): an integer XOR on the high word, no fp64 issue slot.
__device__ __forceinline__ double flip_sign_by(double v, double sgn) {
  return __longlong_as_double(__double_as_longlong(v) ^ (__double_as_longlong(sgn) & (long long)0x8000000000000000ull));
}

template <int NK>
__device__ __forceinline__ double masked_wsum(const double (&k)[NK], const double (&w)[NK], uint32_t mask) {
  double acc = 0.0;
  bool have = false;
#pragma unroll
  for (int j = 0; j < NK; ++j) {
    if ((mask >> j) & 1u) {
      const double p = __dmul_rn(k[j], w[j]);
      acc = have ? __dadd_rn(acc, p) : p;
      have = true;
    }
  }
  return acc;
}

template <int NK, bool DIRECT, int YNEW_MODE>
__device__ __forceinline__ double finish_elem(double y, const double (&k)[NK], const FinishArgs<NK>& a,
                                              double& ynew, double& e) {
  if (YNEW_MODE == 2) ynew = y;
  else ynew = __dadd_rn(y, __dmul_rn(masked_wsum<NK>(k, a.wb, a.mask_b), a.cb));
  const double lo = __dmul_rn(masked_wsum<NK>(k, a.wbh, a.mask_bh), a.cbh);
  if (DIRECT) e = lo;
  else e = __dadd_rn(ynew, -__dadd_rn(y, lo));  // a - b == a + (-b) exactly
  const double tol = __dadd_rn(a.absTol, __dmul_rn(fabs(ynew), a.relTol));
  const double r = err_ratio(e, tol);
  return __dmul_rn(r, r);
}

template <int NK, int W, int U, bool DIRECT, int YNEW_MODE, int THREADS>
__global__ void __launch_bounds__(THREADS) finish_kernel(const FinishArgs<NK> a) {
  const size_t nvec = a.n / W;
  const size_t tile = (size_t)THREADS * U;
  double acc = 0.0;
  for (size_t base = (size_t)blockIdx.x * tile; base < nvec; base += (size_t)gridDim.x * tile) {
    Pk<W> yv[U], kv[U][NK];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t v = base + (size_t)u * THREADS + threadIdx.x;
      if (v < nvec) {
        yv[u] = ld_stream<W>(a.y + v * W);
#pragma unroll
        for (int j = 0; j < NK; ++j) kv[u][j] = ld_stream<W>(a.k[j] + v * W);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t v = base + (size_t)u * THREADS + threadIdx.x;
      if (v < nvec) {
        Pk<W> yo, eo;
#pragma unroll
        for (int e = 0; e < W; ++e) {
          double ke[NK];
#pragma unroll
          for (int j = 0; j < NK; ++j) ke[j] = kv[u][j].v[e];
          acc = __dadd_rn(acc, finish_elem<NK, DIRECT, YNEW_MODE>(yv[u].v[e], ke, a, yo.v[e], eo.v[e]));
        }
        if (YNEW_MODE == 1) st_stream<W>(a.ynew_out + v * W, yo);
        if (a.err_out) st_stream<W>(a.err_out + v * W, eo);
      }
    }
  }
  if (blockIdx.x == 0) {
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

// ---------------------------------------------------------------------------------------------------
// Fused attempt for ELEMENT-LOCAL right-hand sides.
// When f(t, y)[i] depends on y[i] only (the built-in `c*y` and `-(lambda .* y)`), a whole attempt of an
// FSAL pair is independent per element: all S-1 stage inputs, all S-1 right-hand-side evaluations, yNew,
// error_y and the scaled square are evaluated in registers. HBM traffic drops from 57 (DOPRI54) / 78
// (Vern65) vector passes to 5: read y, k1 (FSAL), lambda; write yNew and k_S (the next FSAL).
// Every operation is the same __dmul_rn/__dadd_rn in the same order as in stage_kernel / ewise_kernel /
// finish_kernel, so the results are bit-identical to the unfused pipeline (tested).
// ---------------------------------------------------------------------------------------------------
// PW_USER: an element-local right-hand side supplied by the caller as a CUDA C++ expression of
//   t, y, p0..p3 (per-element parameter vectors), c0..c7 (scalars)
// and compiled at run time (NVRTC, jit.cu) into these same kernels: the generated translation unit defines
//   __device__ double b200rk::user_rhs(double t, double y, const double* p, const double* c)
// and B200RK_USER_NP (number of parameter vectors) before including this header.
enum PointwiseRhs : int { PW_SCALE = 0, PW_DIAG = 1, PW_USER = 2 };
constexpr int kMaxUserVecs = 4, kMaxUserScalars = 8;
#ifndef B200RK_USER_NP
#define B200RK_USER_NP 0
#endif
template <int KIND>
struct PwTraits {  // per-element parameter streams the right-hand side reads next to y
  static constexpr int NP = (KIND == PW_DIAG) ? 1 : (KIND == PW_USER ? B200RK_USER_NP : 0);
  static constexpr int NPX = NP > 0 ? NP : 1;  // array extent (no zero-length arrays)
};

// Signs: PW_SCALE folds the backward-pass negation (g = -f(-t, y), ode.nim:545) into the scalar on the host
// (-(y*c) == y*(-c) exactly). PW_DIAG: k = -(lam*y) forward, +(lam*y) backward; since -(lam*y) == (-lam)*y exactly under
// round-to-nearest, the sign goes onto lambda ONCE per element when it is loaded (pw_param: an integer XOR) instead of
// costing a multiplication by -1 in every one of the S-1 evaluations. No select instructions on the fp64 path.
template <int KIND>
__device__ __forceinline__ double pointwise_rhs(double y, double lam_signed, double c) {
  if (KIND == PW_SCALE) return __dmul_rn(y, c);           // EW_SCALE
  return __dmul_rn(lam_signed, y);                        // EW_NEG_HMUL with the sign already on lambda
}
// parameter j of an element as the right-hand side wants it (PW_DIAG: lambda carrying the sign of `sgn`)
template <int KIND>
__device__ __forceinline__ double pw_param(double p, double sgn) {
  if (KIND == PW_DIAG) return flip_sign_by(p, sgn);
  return p;
}

// Sparsity patterns of the three FSAL pairs, as COMPILE-TIME constants: which terms of each row are kept
// (bit j <-> k_{j+1}). With runtime masks every term costs shifts, predicates and 64-bit selects and the
// kernel becomes ALU-bound (measured: 354 instructions per element, alu pipe 71 %); with constexpr masks the
// unrolled loops fold to exactly the reference's multiply/add sequence. The host checks that the masks it
// derives from the tableau (methods.h) equal the pattern before taking this path.
enum FusedPattern : int { PAT_DOPRI54 = 0, PAT_DOPRI54_STRICT = 1, PAT_TSIT54 = 2, PAT_VERN65 = 3, PAT_VERN65_STRICT = 4 };

template <int PAT>
struct Pattern;
template <>
struct Pattern<PAT_DOPRI54> {  // a72 = bHat2 = 0 dropped (ode.nim:263, 277)
  static constexpr int S = 7;
  static constexpr bool direct = false, last = true;
  __host__ __device__ static constexpr uint32_t a(int r) { return r == 0 ? 0x1u : r == 1 ? 0x3u : r == 2 ? 0x7u : r == 3 ? 0xFu : r == 4 ? 0x1Fu : 0x3Du; }
  __host__ __device__ static constexpr uint32_t b() { return 0x3Du; }
  __host__ __device__ static constexpr uint32_t bh() { return 0x7Du; }
};
template <>
struct Pattern<PAT_DOPRI54_STRICT> {
  static constexpr int S = 7;
  static constexpr bool direct = false, last = true;
  __host__ __device__ static constexpr uint32_t a(int r) { return (2u << r) - 1u; }
  __host__ __device__ static constexpr uint32_t b() { return 0x3Fu; }
  __host__ __device__ static constexpr uint32_t bh() { return 0x7Fu; }
};
template <>
struct Pattern<PAT_TSIT54> {  // dense: no zero weights (ode.nim:317-352); error row is direct (ode.nim:372)
  static constexpr int S = 7;
  static constexpr bool direct = true, last = true;
  __host__ __device__ static constexpr uint32_t a(int r) { return (2u << r) - 1u; }
  __host__ __device__ static constexpr uint32_t b() { return 0x3Fu; }
  __host__ __device__ static constexpr uint32_t bh() { return 0x7Fu; }
};
template <>
struct Pattern<PAT_VERN65> {  // zeros a42 a52 a62 a72 a82 a92 a93 b2 b3 bHat2 bHat3 bHat7 dropped (ode.nim:393-441)
  static constexpr int S = 9;
  static constexpr bool direct = false, last = false;
  __host__ __device__ static constexpr uint32_t a(int r) {
    return r == 0 ? 0x1u : r == 1 ? 0x3u : r == 2 ? 0x5u : r == 3 ? 0xDu : r == 4 ? 0x1Du : r == 5 ? 0x3Du : r == 6 ? 0x7Du : 0xF9u;
  }
  __host__ __device__ static constexpr uint32_t b() { return 0xF9u; }
  __host__ __device__ static constexpr uint32_t bh() { return 0x1B9u; }
};
template <>
struct Pattern<PAT_VERN65_STRICT> {
  static constexpr int S = 9;
  static constexpr bool direct = false, last = false;
  __host__ __device__ static constexpr uint32_t a(int r) { return (2u << r) - 1u; }
  __host__ __device__ static constexpr uint32_t b() { return 0xFFu; }
  __host__ __device__ static constexpr uint32_t bh() { return 0x1FFu; }
};

template <int S>
struct FusedArgs {
  const double* y;
  const double* k1;    // FSAL
  const double* p[kMaxUserVecs];  // per-element parameter streams: p[0] = lambda (PW_DIAG); p0..p3 (PW_USER)
  double rhs_scalar;   // PW_SCALE (already negated for the backward pass)
  double rhs_sign;     // PW_DIAG: -1 forward, +1 backward; PW_USER: +1 forward, -1 backward (g = -f(-t, y))
  double cs[kMaxUserScalars];     // PW_USER: c0..c7
  double t, tsign;     // PW_USER: step start time (solver coordinates) and -1 for the backward pass (f sees -t)
  double cnode[S];     // PW_USER: c_s of the tableau; stage s is evaluated at t + dt*c_s (ode.nim:294-299)
  double a[S - 1][S - 1];   // a[s-2][j] = a_{s,j+1}
  double b[S], bh[S];
  double dt, cb, cbh, absTol, relTol;
  double* ynew;
  double* ks_out;
  size_t n;
  ReduceScratch rs;
};

// left-associated sum of the kept terms; MASK is a compile-time constant after unrolling
template <int NK, uint32_t MASK>
__device__ __forceinline__ double const_wsum(const double (&k)[NK], const double* w) {
  double acc = 0.0;
  bool have = false;
#pragma unroll
  for (int j = 0; j < NK; ++j) {
    if ((MASK >> j) & 1u) {
      const double p = __dmul_rn(k[j], w[j]);
      acc = have ? __dadd_rn(acc, p) : p;
      have = true;
    }
  }
  return acc;
}

// k_s = f(t + dt*c_s, stage input) for the element-local right-hand sides; s is a compile-time constant
template <int KIND, int s, int S>
__device__ __forceinline__ double pw_eval(double in, const double (&pe)[PwTraits<KIND>::NPX], const FusedArgs<S>& a) {
#ifdef B200RK_JIT
  if constexpr (KIND == PW_USER) {
    const double ts = __dadd_rn(a.t, __dmul_rn(a.dt, a.cnode[s - 1]));   // same rounding as the host's t + dt*c[s]
    return flip_sign_by(user_rhs(flip_sign_by(ts, a.tsign), in, pe, a.cs), a.rhs_sign);   // x * (+-1) == x with the sign flipped, exactly
  }
#endif
  return pointwise_rhs<KIND>(in, pe[0], a.rhs_scalar);
}

template <int PAT, int KIND, int s>
struct FusedStages {  // stages 2..s, recursively, so that every row mask is a template constant
  template <int S>
  __device__ __forceinline__ static double run(double y, const double (&pe)[PwTraits<KIND>::NPX], const FusedArgs<S>& a,
                                               double (&k)[S]) {
    if constexpr (s > 2) FusedStages<PAT, KIND, s - 1>::run(y, pe, a, k);
    const double acc = const_wsum<S, Pattern<PAT>::a(s - 2)>(k, a.a[s - 2]);
    const double in = __dadd_rn(y, __dmul_rn(acc, a.dt));
    k[s - 1] = pw_eval<KIND, s, S>(in, pe, a);
    return in;
  }
};

template <int PAT, int KIND>
__device__ __forceinline__ double fused_elem(double y, double k1, const double (&pe)[PwTraits<KIND>::NPX],
                                             const FusedArgs<Pattern<PAT>::S>& a, double& ynew, double& ks) {
  constexpr int S = Pattern<PAT>::S;
  double k[S];
  k[0] = k1;
#pragma unroll
  for (int j = 1; j < S; ++j) k[j] = 0.0;
  const double in = FusedStages<PAT, KIND, S>::run(y, pe, a, k);
  ks = k[S - 1];
  if (Pattern<PAT>::last) ynew = in;
  else ynew = __dadd_rn(y, __dmul_rn(const_wsum<S, Pattern<PAT>::b()>(k, a.b), a.cb));
  const double lo = __dmul_rn(const_wsum<S, Pattern<PAT>::bh()>(k, a.bh), a.cbh);
  double e;
  if (Pattern<PAT>::direct) e = lo;
  else e = __dadd_rn(ynew, -__dadd_rn(y, lo));
  const double tol = __dadd_rn(a.absTol, __dmul_rn(fabs(ynew), a.relTol));
  const double r = err_ratio(e, tol);
  return __dmul_rn(r, r);
}

// L2 != 0: one of the five vectors of an attempt is kept in the 126 MB L2 across launches and everything else
// streams through with evict_first. PW_DIAG keeps lambda (read by EVERY attempt, accepted or not, never
// written); PW_SCALE keeps yNew (the next attempt's y after an accepted step).
template <int PAT, int KIND, int W, int THREADS, int L2 = 0>
__global__ void __launch_bounds__(THREADS) fused_attempt_kernel(const FusedArgs<Pattern<PAT>::S> a) {
  constexpr int LDP = L2 ? L2_EVICT_FIRST : L2_NORMAL;
  constexpr int LAMP = L2 ? L2_EVICT_LAST : L2_NORMAL;
  constexpr int YSTP = (L2 && KIND != PW_DIAG) ? L2_EVICT_LAST : (L2 ? L2_EVICT_FIRST : L2_NORMAL);
  const size_t nvec = a.n / W;
  const size_t stride = (size_t)gridDim.x * THREADS;
  double acc = 0.0;
  // Software-pipelined grid-stride loop: the loads of the next tile are issued before the ~100 dependent
  // fp64 operations of the current one, so HBM latency hides under arithmetic even at 16-32 warps/SM
  // (the kernel is close to balanced: ~46 us of fp64 pipe vs ~51 us of HBM traffic per 2^23 elements).
  constexpr int NP = PwTraits<KIND>::NP, NPX = PwTraits<KIND>::NPX;
  size_t v = (size_t)blockIdx.x * THREADS + threadIdx.x;
  Pk<W> yv, kv, pv[NPX];
  if (v < nvec) {
    yv = ld_pol<W, LDP>(a.y + v * W);
    kv = ld_pol<W, LDP>(a.k1 + v * W);
#pragma unroll
    for (int j = 0; j < NP; ++j) pv[j] = ld_pol<W, LAMP>(a.p[j] + v * W);
  }
  while (v < nvec) {
    const size_t vn = v + stride;
    Pk<W> yn_, kn_, pn_[NPX];
    if (vn < nvec) {
      yn_ = ld_pol<W, LDP>(a.y + vn * W);
      kn_ = ld_pol<W, LDP>(a.k1 + vn * W);
#pragma unroll
      for (int j = 0; j < NP; ++j) pn_[j] = ld_pol<W, LAMP>(a.p[j] + vn * W);
    }
    Pk<W> yo, ko;
#pragma unroll
    for (int e = 0; e < W; ++e) {
      double pe[NPX] = {};
#pragma unroll
      for (int j = 0; j < NP; ++j) pe[j] = pw_param<KIND>(pv[j].v[e], a.rhs_sign);
      acc = __dadd_rn(acc, fused_elem<PAT, KIND>(yv.v[e], kv.v[e], pe, a, yo.v[e], ko.v[e]));
    }
    st_pol<W, YSTP>(a.ynew + v * W, yo);
    st_pol<W, L2 ? L2_EVICT_FIRST : L2_NORMAL>(a.ks_out + v * W, ko);
    yv = yn_; kv = kn_;
#pragma unroll
    for (int j = 0; j < NP; ++j) pv[j] = pn_[j];
    v = vn;
  }
  if (blockIdx.x == 0) {
    const size_t i = nvec * W + threadIdx.x;
    if (i < a.n) {
      double yn, ks, pe[NPX] = {};
#pragma unroll
      for (int j = 0; j < NP; ++j) pe[j] = pw_param<KIND>(a.p[j][i], a.rhs_sign);
      acc = __dadd_rn(acc, fused_elem<PAT, KIND>(a.y[i], a.k1[i], pe, a, yn, ks));
      a.ynew[i] = yn;
      a.ks_out[i] = ks;
    }
  }
  grid_sum_finish<THREADS>(acc, a.rs);
}

// ---------------------------------------------------------------------------------------------------
// Device-resident driver loop for element-local right-hand sides (small N):
// ONE persistent cooperative kernel runs up to `max_steps` accepted steps of the `while t < tEnd` loop
// (ode.nim:511-541) including the retry loop (ode.nim:57-76) and both controllers, with a grid-wide barrier
// per attempt instead of a kernel launch + host read-back per attempt. Every CTA adds the per-CTA partials in
// index order after the barrier, so all CTAs compute the same error and take the same branch. The state
// vectors ping-pong between Y[0]/Y[1] and F[0]/F[1] exactly like the host-driven path. Element-wise
// arithmetic is the same fused_elem as fused_attempt_kernel (bit-identical per attempt for a given dt); the
// controller uses the device pow(), which may differ from glibc's in the last ulp, so step sequences agree
// with the host-driven path within the tolerances stated in DESIGN.md §5, not bit for bit.
// ---------------------------------------------------------------------------------------------------
struct RunState {
  double t, dt, t_end, error;
  long long steps, attempts, rejected, limiter_hits;
  int cur;     // Y[cur], F[cur] hold the current state / FSAL
  int status;  // 0 ok, 1 error norm is NaN
};

template <int S>
struct RunArgs {
  FusedArgs<S> f;  // weights, tolerances, RHS parameters (pointers and dt are overridden per attempt)
  double* Y[2];
  double* F[2];
  double dtMin, dtMax, inv_order_inner, inv_order_outer, n_global;
  long long max_steps;
  double* partials;  // [2][gridDim.x]
  RunState* state;        // device
  RunState* state_host;   // mapped pinned mirror, written once at exit
  unsigned long long* seq_host;
  unsigned long long seq;        // sequence number of attempt 0 of this launch (attempt i uses seq + i)
  PeerMail mail;                 // the error norm travels through mailboxes every attempt: the peers' (world > 1, NVLink) or
                                 // a local one (world == 1: box[0] is this GPU's own, the same code path)
  unsigned long long* arrive;    // device counter, zero at launch: CTA arrivals (attempt i of the launch is complete at (i+1)*gridDim.x)
};

__device__ __forceinline__ double dev_nim_min(double x, double y) { return (x <= y) ? x : y; }
__device__ __forceinline__ double dev_nim_max(double x, double y) { return (y <= x) ? x : y; }

template <int PAT, int KIND, int W, int THREADS>
__global__ void __launch_bounds__(THREADS, 2) fused_run_kernel(const RunArgs<Pattern<PAT>::S> a) {
  // launched cooperatively for the co-residency guarantee only (CTAs wait for each other's tickets): no grid.sync()
  __shared__ double bcast;
  RunState st = *a.state;
  FusedArgs<Pattern<PAT>::S> f = a.f;
  const size_t nvec = f.n / W;
  const size_t stride = (size_t)gridDim.x * THREADS;
  constexpr int NP = PwTraits<KIND>::NP, NPX = PwTraits<KIND>::NPX;
  int parity = 0;
  long long done = 0, launch_attempts = 0;   // launch_attempts: attempts made by THIS launch (the arrival counter starts at 0)
  while (st.t < st.t_end && done < a.max_steps && st.status == 0) {
    double dt = dev_nim_min(st.dt, st.t_end - st.t);                      // ode.nim:525
    f.t = st.t;                                                           // PW_USER: stage times t + dt*c_s
    int limit = 0;
    double error = 0.0;
    while (true) {                                                        // ode.nim:58
      f.dt = dt; f.cb = dt; f.cbh = dt;
      const double* y = a.Y[st.cur];
      const double* k1 = a.F[st.cur];
      double* yn = a.Y[1 - st.cur];
      double* ks = a.F[1 - st.cur];
      double acc = 0.0;
      {  // software-pipelined grid-stride loop (see fused_attempt_kernel); coherent loads: Y/F are rewritten inside this kernel
        size_t v = (size_t)blockIdx.x * THREADS + threadIdx.x;
        Pk<W> yv, kv, pv[NPX];
        if (v < nvec) {
          yv = ld_plain<W>(y + v * W);
          kv = ld_plain<W>(k1 + v * W);
#pragma unroll
          for (int j = 0; j < NP; ++j) pv[j] = ld_stream<W>(f.p[j] + v * W);
        }
        while (v < nvec) {
          const size_t vn = v + stride;
          Pk<W> yn_, kn_, pn_[NPX];
          if (vn < nvec) {
            yn_ = ld_plain<W>(y + vn * W);
            kn_ = ld_plain<W>(k1 + vn * W);
#pragma unroll
            for (int j = 0; j < NP; ++j) pn_[j] = ld_stream<W>(f.p[j] + vn * W);
          }
          Pk<W> yo, ko;
#pragma unroll
          for (int e = 0; e < W; ++e) {
            double pe[NPX] = {};
#pragma unroll
            for (int j = 0; j < NP; ++j) pe[j] = pw_param<KIND>(pv[j].v[e], f.rhs_sign);
            acc = __dadd_rn(acc, fused_elem<PAT, KIND>(yv.v[e], kv.v[e], pe, f, yo.v[e], ko.v[e]));
          }
          st_stream<W>(yn + v * W, yo);
          st_stream<W>(ks + v * W, ko);
          yv = yn_; kv = kn_;
#pragma unroll
          for (int j = 0; j < NP; ++j) pv[j] = pn_[j];
          v = vn;
        }
      }
      if (blockIdx.x == 0) {
        const size_t i = nvec * W + threadIdx.x;
        if (i < f.n) {
          double y1, k1o, pe[NPX] = {};
#pragma unroll
          for (int j = 0; j < NP; ++j) pe[j] = pw_param<KIND>(f.p[j][i], f.rhs_sign);
          acc = __dadd_rn(acc, fused_elem<PAT, KIND>(y[i], k1[i], pe, f, y1, k1o));
          yn[i] = y1;
          ks[i] = k1o;
        }
      }
      // ---- grid-wide (and, sharded, machine-wide) sum of the attempt's r*r terms: ticket + mailbox, no grid barrier ----
      // Every CTA deposits its partial and takes a ticket; the LAST CTA to arrive adds the partials in index order (same
      // bits every run) and publishes the shard's sum into slot [rank] of every rank's mailbox — its own included — with one
      // 16-byte store each (mail_put). EVERY CTA of every rank then waits for the `world` slots of its local mailbox and adds
      // them in rank order: all CTAs of all ranks obtain identical bits and take the identical accept / reject branch.
      // There is no grid barrier, no second hop from a leader CTA to the others and no fence on the critical path: nothing
      // but the sum crosses CTAs in this kernel (a thread re-reads only the elements of yNew / k_S it wrote itself).
      const double bsum = block_sum<THREADS>(acc);
      const unsigned long long seq = a.seq + (unsigned long long)st.attempts;
      double* part = a.partials + (size_t)parity * gridDim.x;
      __shared__ bool is_last;
      __shared__ double peer_vals[kMaxPeers];
      __shared__ int peer_bad;
      if (threadIdx.x == 0) {
        part[blockIdx.x] = bsum;
        peer_bad = 0;
        __threadfence();
        const unsigned long long ticket = atomicAdd(a.arrive, 1ull);
        is_last = (ticket == (unsigned long long)(launch_attempts + 1) * gridDim.x - 1ull);
      }
      __syncthreads();
      if (is_last) {
        __threadfence();
        double p = 0.0;
        for (unsigned int i = threadIdx.x; i < gridDim.x; i += THREADS) p = __dadd_rn(p, __ldcg(part + i));
        const double total = block_sum<THREADS>(p);
        if (threadIdx.x == 0) bcast = total;
        __syncthreads();
        if ((int)threadIdx.x < a.mail.world) mail_put(mail_slot(a.mail.box[threadIdx.x], seq, a.mail.rank), bcast, mail_tag(seq));
      }
      if ((int)threadIdx.x < a.mail.world) {
        double v = 0.0;
        const bool ok = mail_wait(a.mail, seq, (int)threadIdx.x, &v);
        peer_vals[threadIdx.x] = ok ? v : 0.0;
        if (!ok) atomicExch(&peer_bad, 1);
      }
      __syncthreads();
      double S2 = 0.0;
      for (int p2 = 0; p2 < a.mail.world; ++p2) S2 = __dadd_rn(S2, peer_vals[p2]);   // rank order: identical everywhere
      const int bad = peer_bad;
      __syncthreads();   // peer_vals / peer_bad / is_last are rewritten by the next attempt
      ++launch_attempts;
      if (bad) { st.status = 2; st.attempts++; break; }
      parity ^= 1;
      st.attempts++;
      error = sqrt(1.0 / a.n_global * S2);                                // ode.nim:64-65
      if (error <= 1) break;                                              // ode.nim:69-70
      if (error != error) { st.status = 1; break; }
      st.rejected++;
      dt = dt * dev_nim_min(4, dev_nim_max(0.125, 0.9 * pow(1.0 / error, a.inv_order_inner)));  // ode.nim:71
      if (fabs(dt) < a.dtMin) { dt = a.dtMin; limit += 1; st.limiter_hits++; }                  // ode.nim:72-74
      else if (a.dtMax < fabs(dt)) dt = a.dtMax;                                                // ode.nim:75-76
      if (!(limit < 2)) break;
    }
    st.error = error;
    if (st.status != 0) break;
    st.cur = 1 - st.cur;
    st.t += dt;                                                           // ode.nim:532
    st.steps++;
    ++done;
    if (error == 0.0) dt *= 5;                                            // ode.nim:533-541
    else dt = dt * dev_nim_min(4, dev_nim_max(0.125, 0.9 * pow(1.0 / error, a.inv_order_outer)));
    if (dt < a.dtMin) dt = a.dtMin;
    else if (a.dtMax < dt) dt = a.dtMax;
    st.dt = dt;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {   // identical in every CTA; kernel completion orders all stores before the host's read
    *a.state = st;
    *a.state_host = st;  // the host waits for the launch with a stream synchronisation (once per many steps)
    __threadfence_system();
  }
}

// Fused RK4 step for element-local right-hand sides (ode.nim:180-189): reads y (+ lambda), writes yNew.
template <int KIND>
__device__ __forceinline__ double fused_rk4_elem(double y, double lam_raw, double c, double neg, double hdt, double dt, double c6) {
  const double lam = pw_param<KIND>(lam_raw, neg);
  const double k1 = pointwise_rhs<KIND>(y, lam, c);
  const double k2 = pointwise_rhs<KIND>(__dadd_rn(y, __dmul_rn(k1, hdt)), lam, c);
  const double k3 = pointwise_rhs<KIND>(__dadd_rn(y, __dmul_rn(k2, hdt)), lam, c);
  const double k4 = pointwise_rhs<KIND>(__dadd_rn(y, __dmul_rn(k3, dt)), lam, c);
  return rk4_elem(y, k1, k2, k3, k4, c6);
}
template <int KIND, int W, int THREADS>
__global__ void __launch_bounds__(THREADS)
    fused_rk4_kernel(const double* __restrict__ y, const double* __restrict__ lam, double c, double sgn, double hdt,
                     double dt, double c6, double* __restrict__ out, size_t n) {
  const size_t nvec = n / W;
  for (size_t v = (size_t)blockIdx.x * THREADS + threadIdx.x; v < nvec; v += (size_t)gridDim.x * THREADS) {
    const Pk<W> yv = ld_stream<W>(y + v * W);
    Pk<W> lv;
    if (KIND == PW_DIAG) lv = ld_stream<W>(lam + v * W);
    Pk<W> o;
#pragma unroll
    for (int e = 0; e < W; ++e) o.v[e] = fused_rk4_elem<KIND>(yv.v[e], KIND == PW_DIAG ? lv.v[e] : 0.0, c, sgn, hdt, dt, c6);
    st_stream<W>(out + v * W, o);
  }
  if (blockIdx.x == 0) {
    const size_t i = nvec * W + threadIdx.x;
    if (i < n) out[i] = fused_rk4_elem<KIND>(y[i], KIND == PW_DIAG ? lam[i] : 0.0, c, sgn, hdt, dt, c6);
  }
}

// ---------------------------------------------------------------------------------------------------
// Run-time compiled (PW_USER) right-hand side outside the fused pairs: the plain ODEProc evaluation
// dydt = f(t, y) every method can call through the stage / RHS / finish pipeline, and a whole RK4 step.
// Both exist only in the NVRTC translation unit; the argument block is shared with the host (jit.cu).
// ---------------------------------------------------------------------------------------------------
struct UserRhsArgs {
  const double* y;
  const double* p[kMaxUserVecs];
  double cs[kMaxUserScalars];
  double t;            // user_rhs_kernel: the time handed to f; user_rk4_kernel: step start time (solver coordinates)
  double tsign, rsign; // user_rk4_kernel: -1/-1 for the backward pass g(t, y) = -f(-t, y) (ode.nim:545), else +1/+1
  double hdt, dt, c6;  // user_rk4_kernel: 0.5*dt, dt, dt/6.0 computed by the host (ode.nim:185-188)
  double* out;
  size_t n;
};

#ifdef B200RK_JIT
template <int W, int U, int THREADS, int L2 = 0>
__global__ void __launch_bounds__(THREADS) user_rhs_kernel(const UserRhsArgs a) {
  constexpr int NP = PwTraits<PW_USER>::NP, NPX = PwTraits<PW_USER>::NPX;
  const size_t nvec = a.n / W;
  const size_t tile = (size_t)THREADS * U;
  for (size_t base = (size_t)blockIdx.x * tile; base < nvec; base += (size_t)gridDim.x * tile) {
    Pk<W> yv[U], pv[U][NPX];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t v = base + (size_t)u * THREADS + threadIdx.x;
      if (v < nvec) {
        yv[u] = ld_pol<W, L2 ? L2_EVICT_FIRST : L2_NORMAL>(a.y + v * W);
#pragma unroll
        for (int j = 0; j < NP; ++j) pv[u][j] = ld_stream<W>(a.p[j] + v * W);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t v = base + (size_t)u * THREADS + threadIdx.x;
      if (v < nvec) {
        Pk<W> o;
#pragma unroll
        for (int e = 0; e < W; ++e) {
          double pe[NPX] = {};
#pragma unroll
          for (int j = 0; j < NP; ++j) pe[j] = pv[u][j].v[e];
          o.v[e] = user_rhs(a.t, yv[u].v[e], pe, a.cs);
        }
        st_pol<W, L2 ? L2_EVICT_LAST : L2_NORMAL>(a.out + v * W, o);
      }
    }
  }
  if (blockIdx.x == 0) {
    const size_t i = nvec * W + threadIdx.x;
    if (i < a.n) {
      double pe[NPX] = {};
#pragma unroll
      for (int j = 0; j < NP; ++j) pe[j] = a.p[j][i];
      a.out[i] = user_rhs(a.t, a.y[i], pe, a.cs);
    }
  }
}

// k_s = g(t_s, in) with g = f forward and g(t, y) = -f(-t, y) backward; t_s = t + dt*c_s rounded like the host's
__device__ __forceinline__ double user_rk4_elem(double y, const double (&pe)[PwTraits<PW_USER>::NPX], const UserRhsArgs& a) {
  const double tm = __dadd_rn(a.t, __dmul_rn(a.dt, 0.5)), te = __dadd_rn(a.t, __dmul_rn(a.dt, 1.0));   // ode.nim:185-187
  const double k1 = flip_sign_by(user_rhs(flip_sign_by(a.t, a.tsign), y, pe, a.cs), a.rsign);
  const double k2 = flip_sign_by(user_rhs(flip_sign_by(tm, a.tsign), __dadd_rn(y, __dmul_rn(k1, a.hdt)), pe, a.cs), a.rsign);
  const double k3 = flip_sign_by(user_rhs(flip_sign_by(tm, a.tsign), __dadd_rn(y, __dmul_rn(k2, a.hdt)), pe, a.cs), a.rsign);
  const double k4 = flip_sign_by(user_rhs(flip_sign_by(te, a.tsign), __dadd_rn(y, __dmul_rn(k3, a.dt)), pe, a.cs), a.rsign);
  return rk4_elem(y, k1, k2, k3, k4, a.c6);
}
template <int W, int THREADS>
__global__ void __launch_bounds__(THREADS) user_rk4_kernel(const UserRhsArgs a) {
  constexpr int NP = PwTraits<PW_USER>::NP, NPX = PwTraits<PW_USER>::NPX;
  const size_t nvec = a.n / W;
  for (size_t v = (size_t)blockIdx.x * THREADS + threadIdx.x; v < nvec; v += (size_t)gridDim.x * THREADS) {
    const Pk<W> yv = ld_stream<W>(a.y + v * W);
    Pk<W> pv[NPX];
#pragma unroll
    for (int j = 0; j < NP; ++j) pv[j] = ld_stream<W>(a.p[j] + v * W);
    Pk<W> o;
#pragma unroll
    for (int e = 0; e < W; ++e) {
      double pe[NPX] = {};
#pragma unroll
      for (int j = 0; j < NP; ++j) pe[j] = pv[j].v[e];
      o.v[e] = user_rk4_elem(yv.v[e], pe, a);
    }
    st_stream<W>(a.out + v * W, o);
  }
  if (blockIdx.x == 0) {
    const size_t i = nvec * W + threadIdx.x;
    if (i < a.n) {
      double pe[NPX] = {};
#pragma unroll
      for (int j = 0; j < NP; ++j) pe[j] = a.p[j][i];
      a.out[i] = user_rk4_elem(a.y[i], pe, a);
    }
  }
}
#endif  // B200RK_JIT

// ---------------------------------------------------------------------------------------------------
// Dense output: hermiteSpline (utils.nim:273-279) with host-side scalars
//   out = ((h00*y1 + hA*dy1) + h01*y2) + hB*dy2,   hA = h10*(x2-x1), hB = h11*(x2-x1)
// ---------------------------------------------------------------------------------------------------
template <int W, int THREADS>
__global__ void __launch_bounds__(THREADS)
    hermite_kernel(const double* __restrict__ y1, const double* __restrict__ dy1, const double* __restrict__ y2,
                   const double* __restrict__ dy2, double h00, double hA, double h01, double hB,
                   double* __restrict__ out, size_t n) {
  const size_t nvec = n / W;
  for (size_t v = (size_t)blockIdx.x * THREADS + threadIdx.x; v < nvec; v += (size_t)gridDim.x * THREADS) {
    const Pk<W> a = ld_stream<W>(y1 + v * W), b = ld_stream<W>(dy1 + v * W), c = ld_stream<W>(y2 + v * W),
                d = ld_stream<W>(dy2 + v * W);
    Pk<W> o;
#pragma unroll
    for (int e = 0; e < W; ++e)
      o.v[e] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(a.v[e], h00), __dmul_rn(b.v[e], hA)), __dmul_rn(c.v[e], h01)),
                         __dmul_rn(d.v[e], hB));
    st_stream<W>(out + v * W, o);
  }
  if (blockIdx.x == 0) {
    const size_t i = nvec * W + threadIdx.x;
    if (i < n)
      out[i] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(y1[i], h00), __dmul_rn(dy1[i], hA)), __dmul_rn(y2[i], h01)),
                         __dmul_rn(dy2[i], hB));
  }
}

// ---------------------------------------------------------------------------------------------------
// Element-wise Vector[T] operators (utils.nim:59-223) and built-in right-hand sides
// ---------------------------------------------------------------------------------------------------
enum EwiseOp : int {
  EW_ADD = 0,        // a + b                 utils.nim:59-64
  EW_SUB = 1,        // a - b                 utils.nim:113-118
  EW_HMUL = 2,       // a *. b                utils.nim:186-191
  EW_HDIV = 3,       // a /. b                utils.nim:192-197
  EW_SCALE = 4,      // s * a                 utils.nim:176-180
  EW_ADD_SCALAR = 5, // s +. a                utils.nim:78-82
  EW_NEG = 6,        // -a                    utils.nim:214-218
  EW_ABS = 7,        // abs(a)                utils.nim:219-223
  EW_DIV_SCALAR = 8, // a / s                 utils.nim:166-170
  EW_NEG_HMUL = 9,   // -(a *. b)             diag-linear right-hand side k = -(lambda .* y)
  EW_FILL = 10,      // s
};

template <int OP>
__device__ __forceinline__ double ewise_elem(double a, double b, double s) {
  switch (OP) {
    case EW_ADD: return __dadd_rn(a, b);
    case EW_SUB: return __dadd_rn(a, -b);
    case EW_HMUL: return __dmul_rn(a, b);
    case EW_HDIV: return __ddiv_rn(a, b);
    case EW_SCALE: return __dmul_rn(a, s);
    case EW_ADD_SCALAR: return __dadd_rn(a, s);
    case EW_NEG: return -a;
    case EW_ABS: return fabs(a);
    case EW_DIV_SCALAR: return __ddiv_rn(a, s);
    case EW_NEG_HMUL: return -__dmul_rn(a, b);
    case EW_FILL: return s;
    default: return a;
  }
}
template <int OP>
struct EwiseArity { static constexpr bool binary = (OP == EW_ADD || OP == EW_SUB || OP == EW_HMUL || OP == EW_HDIV || OP == EW_NEG_HMUL); };

// L2 != 0 (right-hand-side use, out never aliases an input): inputs stream through with evict_first, the
// result is stored evict_last for the stage kernel that consumes it next.
template <int OP, int W, int U, int THREADS, int L2 = 0>
__global__ void __launch_bounds__(THREADS)
    ewise_kernel(const double* a, const double* b, double s, double* out, size_t n) {
  constexpr bool BIN = EwiseArity<OP>::binary;
  const size_t nvec = n / W;
  const size_t tile = (size_t)THREADS * U;
  for (size_t base = (size_t)blockIdx.x * tile; base < nvec; base += (size_t)gridDim.x * tile) {
    Pk<W> av[U], bv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t v = base + (size_t)u * THREADS + threadIdx.x;
      if (v < nvec) {
        av[u] = L2 ? ld_pol<W, L2_EVICT_FIRST>(a + v * W) : ld_plain<W>(a + v * W);
        if (BIN) bv[u] = L2 ? ld_pol<W, L2_EVICT_FIRST>(b + v * W) : ld_plain<W>(b + v * W);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t v = base + (size_t)u * THREADS + threadIdx.x;
      if (v < nvec) {
        Pk<W> o;
#pragma unroll
        for (int e = 0; e < W; ++e) o.v[e] = ewise_elem<OP>(av[u].v[e], BIN ? bv[u].v[e] : 0.0, s);
        st_pol<W, L2 ? L2_EVICT_LAST : L2_NORMAL>(out + v * W, o);
      }
    }
  }
  if (blockIdx.x == 0) {
    const size_t i = nvec * W + threadIdx.x;
    if (i < n) out[i] = ewise_elem<OP>(a[i], BIN ? b[i] : 0.0, s);
  }
}

// Plain sum(v) (utils.nim:243-250) — deterministic tree instead of the sequential CPU order.
template <int W, int THREADS>
__global__ void __launch_bounds__(THREADS) sum_kernel(const double* __restrict__ a, size_t n, ReduceScratch rs) {
  const size_t nvec = n / W;
  double acc = 0.0;
  for (size_t v = (size_t)blockIdx.x * THREADS + threadIdx.x; v < nvec; v += (size_t)gridDim.x * THREADS) {
    const Pk<W> x = ld_stream<W>(a + v * W);
#pragma unroll
    for (int e = 0; e < W; ++e) acc = __dadd_rn(acc, x.v[e]);
  }
  if (blockIdx.x == 0) {
    const size_t i = nvec * W + threadIdx.x;
    if (i < n) acc = __dadd_rn(acc, a[i]);
  }
  grid_sum_finish<THREADS>(acc, rs);
}

// Lorenz-96 right-hand side, cyclic: k[i] = ((y[i+1] - y[i-2]) * y[i-1] - y[i]) + F.
// `y` is the local block [lo, lo+n) of a cyclic global vector; left2[0..1] = y[lo-2], y[lo-1] and
// right1[0] = y[lo+n] (for a single GPU these are &y[n-2] and &y[0]: no copies). Two outputs per thread from three
// aligned 128-bit loads; the neighbour loads hit L1/L2, DRAM traffic stays at one read + one write.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
    lorenz96_kernel(const double* __restrict__ y, const double* __restrict__ left2, const double* __restrict__ right1,
                    double F, double* __restrict__ out, size_t n) {
  const size_t npair = n / 2;
  for (size_t p = (size_t)blockIdx.x * THREADS + threadIdx.x; p < npair; p += (size_t)gridDim.x * THREADS) {
    const size_t i = 2 * p;
    const double2 c = __ldg(reinterpret_cast<const double2*>(y + i));
    double2 l, r;
    if (p > 0) l = __ldg(reinterpret_cast<const double2*>(y + i - 2));
    else l = make_double2(left2[0], left2[1]);
    double rx;
    if (i + 2 < n) rx = __ldg(y + i + 2);
    else rx = right1[0];
    r.x = rx;
    double2 o;
    o.x = __dadd_rn(__dadd_rn(__dmul_rn(__dadd_rn(c.y, -l.x), l.y), -c.x), F);
    o.y = __dadd_rn(__dadd_rn(__dmul_rn(__dadd_rn(r.x, -l.y), c.x), -c.y), F);
    *reinterpret_cast<double2*>(out + i) = o;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {  // odd n: last element
    const size_t i = n - 1;
    out[i] = __dadd_rn(__dadd_rn(__dmul_rn(__dadd_rn(right1[0], -y[i - 2]), y[i - 1]), -y[i]), F);  // n >= 4
  }
}

// ---------------------------------------------------------------------------------------------------
// Stage accumulate FUSED with the Lorenz-96 stencil right-hand side (single GPU, cyclic):
//   in[i] = y[i] + c*(w1 k1[i] + ... + wM kM[i])          (same arithmetic as stage_kernel)
//   k[i]  = sgn * (((in[i+1] - in[i-2]) * in[i-1] - in[i]) + F)      (same as lorenz96_kernel [+ negate])
// The stage input never travels to HBM: each CTA stages its tile of `in` plus a 2+1 element halo (recomputed
// from y and the k's, wrapping cyclically) in shared memory and applies the stencil from there. Per stage
// this removes the write and the re-read of the intermediate vector: (M+2)+2 passes become M+2. `in` is
// stored only when the caller needs it (the last stage of DOPRI54/Tsit54, whose input is yNew).
// ---------------------------------------------------------------------------------------------------
template <int M, int THREADS>
__global__ void __launch_bounds__(THREADS)
    stage_l96_kernel(const StageArgs<M> a, double F, double sgn, double* __restrict__ kout) {
  constexpr int W = 4, TILE = THREADS * W;
  __shared__ double s[TILE + 4];  // s[0..1] left halo | s[2 .. 2+TILE) tile | right halo directly after the last valid element
  const size_t n = a.n;
  const size_t tile0 = (size_t)blockIdx.x * TILE;
  const size_t tile_len = (tile0 + TILE <= n) ? (size_t)TILE : n - tile0;
  const size_t i0 = tile0 + (size_t)threadIdx.x * W;
  double in[W];
  if (i0 + W <= n) {
    const Pk<W> yv = ld_stream<W>(a.y + i0);
    Pk<W> kv[M];
#pragma unroll
    for (int j = 0; j < M; ++j) kv[j] = ld_stream<W>(a.k[j] + i0);
#pragma unroll
    for (int e = 0; e < W; ++e) {
      double ke[M];
#pragma unroll
      for (int j = 0; j < M; ++j) ke[j] = kv[j].v[e];
      in[e] = stage_elem<M, false>(yv.v[e], ke, a.w, a.c);
    }
    if (a.out) {
      Pk<W> o;
#pragma unroll
      for (int e = 0; e < W; ++e) o.v[e] = in[e];
      st_stream<W>(a.out + i0, o);
    }
  } else {
#pragma unroll
    for (int e = 0; e < W; ++e) {
      const size_t i = i0 + e;
      in[e] = 0.0;
      if (i < n) {
        double ke[M];
#pragma unroll
        for (int j = 0; j < M; ++j) ke[j] = a.k[j][i];
        in[e] = stage_elem<M, false>(a.y[i], ke, a.w, a.c);
        if (a.out) a.out[i] = in[e];
      }
    }
  }
#pragma unroll
  for (int e = 0; e < W; ++e) s[2 + threadIdx.x * W + e] = in[e];
  if (threadIdx.x < 3) {  // halo: global indices tile0-2, tile0-1 (left) and tile0+tile_len (right), cyclic
    const size_t g = (threadIdx.x < 2) ? (tile0 + n - 2 + threadIdx.x) % n : (tile0 + tile_len) % n;
    double ke[M];
#pragma unroll
    for (int j = 0; j < M; ++j) ke[j] = a.k[j][g];
    const double h = stage_elem<M, false>(a.y[g], ke, a.w, a.c);
    // the right halo is written after the block-wide barrier below so it cannot race with a thread storing a
    // padded element of a ragged last tile into the same slot
    if (threadIdx.x < 2) s[threadIdx.x] = h;
    else in[0] = h;  // thread 2 keeps it in a register until the barrier (its own in[0] is already in shared memory)
  }
  __syncthreads();
  if (threadIdx.x == 2) s[2 + tile_len] = in[0];
  __syncthreads();
  if (i0 < n) {
    const int b = 2 + threadIdx.x * W;
    double k[W];
#pragma unroll
    for (int e = 0; e < W; ++e) {
      const double v = __dadd_rn(__dadd_rn(__dmul_rn(__dadd_rn(s[b + e + 1], -s[b + e - 2]), s[b + e - 1]), -s[b + e]), F);
      k[e] = __dmul_rn(v, sgn);  // sgn = +1, or -1 for the backward pass g = -f(-t, y): exact either way
    }
    if (i0 + W <= n) {
      Pk<W> o;
#pragma unroll
      for (int e = 0; e < W; ++e) o.v[e] = k[e];
      st_stream<W>(kout + i0, o);
    } else {
#pragma unroll
      for (int e = 0; e < W; ++e)
        if (i0 + e < n) kout[i0 + e] = k[e];
    }
  }
}

}  // namespace b200rk
