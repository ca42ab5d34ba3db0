// launch.cu — generic kernel launchers: template dispatch on the number of streams / vector width / launch
// geometry, the algorithmic-byte accounting of the profiler, and the read-back of a reduction result.
#include "internal.hpp"
#include "finish_pf.cuh"

template <int M, int W, bool CHAIN>
static int launch_stage_mw(b200rk_ctx* c, const StageArgs<M>& a) {
  constexpr int U = StageUnroll<M, W>::value;
  unsigned grid = grid_for(c, a.n / W, kThreads * U);
  if (l2_on(c, a.n) && !CHAIN) stage_kernel<M, W, U, CHAIN, kThreads, 1><<<grid, kThreads, 0, c->stream>>>(a);
  else stage_kernel<M, W, U, CHAIN, kThreads><<<grid, kThreads, 0, c->stream>>>(a);
  CUDA_TRY(c, cudaGetLastError());
  return B200RK_OK;
}

template <int M>
static int launch_stage_m(b200rk_ctx* c, const double* y, const double* const* k, const double* w, double cc,
                          bool chain, double* out, size_t n) {
  StageArgs<M> a;
  a.y = y; a.c = cc; a.out = out; a.n = n;
  for (int j = 0; j < M; ++j) { a.k[j] = k[j]; a.w[j] = w[j]; }
  ProfScope p(c, B200RK_K_STAGE, 8.0 * double(n) * (M + 2));
  if (chain) {
    if (c->vec_width == 4) return launch_stage_mw<M, 4, true>(c, a);
    return launch_stage_mw<M, 2, true>(c, a);
  }
  if (c->vec_width == 4) return launch_stage_mw<M, 4, false>(c, a);
  return launch_stage_mw<M, 2, false>(c, a);
}

// out = y + cc*(sum w_j k_j)  (or chain form); m >= 1 terms after zero-dropping
int launch_stage(b200rk_ctx* c, int m, const double* y, const double* const* k, const double* w, double cc,
                        bool chain, double* out, size_t n) {
  if (n == 0) return B200RK_OK;
  switch (m) {
    case 1: return launch_stage_m<1>(c, y, k, w, cc, chain, out, n);
    case 2: return launch_stage_m<2>(c, y, k, w, cc, chain, out, n);
    case 3: return launch_stage_m<3>(c, y, k, w, cc, chain, out, n);
    case 4: return launch_stage_m<4>(c, y, k, w, cc, chain, out, n);
    case 5: return launch_stage_m<5>(c, y, k, w, cc, chain, out, n);
    case 6: return launch_stage_m<6>(c, y, k, w, cc, chain, out, n);
    case 7: return launch_stage_m<7>(c, y, k, w, cc, chain, out, n);
    case 8: return launch_stage_m<8>(c, y, k, w, cc, chain, out, n);
    case 9: return launch_stage_m<9>(c, y, k, w, cc, chain, out, n);
  }
  return fail(c, B200RK_EINVAL, "stage_accum: m must be in 1..9");
}

template <int NK, int W, bool DIRECT, int MODE>
static int launch_finish_cfg(b200rk_ctx* c, const FinishPlan& p) {
  constexpr int U = StageUnroll<NK, W>::value;
  FinishArgs<NK> a;
  a.y = p.y;
  for (int j = 0; j < NK; ++j) { a.k[j] = p.k[j]; a.wb[j] = p.wb[j]; a.wbh[j] = p.wbh[j]; }
  a.mask_b = p.mask_b; a.mask_bh = p.mask_bh; a.cb = p.cb; a.cbh = p.cbh; a.absTol = p.absTol; a.relTol = p.relTol;
  a.ynew_out = p.ynew_out; a.err_out = p.err_out; a.n = p.n;
  unsigned grid = grid_for(c, p.n / W, kThreads * U, c->finish_ctas_per_sm);
  TRY(ensure_partials(c, grid));
  a.rs = reduce_scratch(c);
  if constexpr (NK >= 6 && U == 1) {  // software-pipelined form (finish_pf.cuh), knob "finish_prefetch": measured, no gain, off
    if (c->finish_prefetch) {
      finish_pf_kernel<NK, W, DIRECT, MODE, kThreads><<<grid, kThreads, 0, c->stream>>>(a);
      CUDA_TRY(c, cudaGetLastError());
      return B200RK_OK;
    }
  }
  finish_kernel<NK, W, U, DIRECT, MODE, kThreads><<<grid, kThreads, 0, c->stream>>>(a);
  CUDA_TRY(c, cudaGetLastError());
  return B200RK_OK;
}
template <int NK>
static int launch_finish_nk(b200rk_ctx* c, const FinishPlan& p) {
  const bool w4 = (c->vec_width == 4);
  if (p.direct && p.ynew_mode == 2) return w4 ? launch_finish_cfg<NK, 4, true, 2>(c, p) : launch_finish_cfg<NK, 2, true, 2>(c, p);
  if (!p.direct && p.ynew_mode == 0) return w4 ? launch_finish_cfg<NK, 4, false, 0>(c, p) : launch_finish_cfg<NK, 2, false, 0>(c, p);
  if (!p.direct && p.ynew_mode == 1) return w4 ? launch_finish_cfg<NK, 4, false, 1>(c, p) : launch_finish_cfg<NK, 2, false, 1>(c, p);
  return fail(c, B200RK_EINVAL, "finish: unsupported mode");
}
int launch_finish(b200rk_ctx* c, const FinishPlan& p) {
  // algorithmic bytes: y (or yNew) + nk derivative streams read, yNew written only in mode 1
  double bytes = 8.0 * double(p.n) * (p.nk + 1 + (p.ynew_mode == 1 ? 1 : 0) + (p.err_out ? 1 : 0));
  ProfScope ps(c, B200RK_K_FINISH, bytes);
  switch (p.nk) {
    case 1: return launch_finish_nk<1>(c, p);
    case 2: return launch_finish_nk<2>(c, p);
    case 3: return launch_finish_nk<3>(c, p);
    case 4: return launch_finish_nk<4>(c, p);
    case 5: return launch_finish_nk<5>(c, p);
    case 6: return launch_finish_nk<6>(c, p);
    case 7: return launch_finish_nk<7>(c, p);
    case 8: return launch_finish_nk<8>(c, p);
    case 9: return launch_finish_nk<9>(c, p);
  }
  return fail(c, B200RK_EINVAL, "finish: nk must be in 1..9");
}

// After a reducing kernel: (allreduce across shards) and bring the scalar to the host.
// Single GPU: the last CTA of the kernel wrote the sum and then the launch's sequence number into mapped
// pinned memory; the host polls that word (~2 us) instead of paying a stream synchronisation (~6-8 us) per
// attempt. A stuck or faulted kernel is caught by the bounded spin falling back to cudaStreamSynchronize.
int fetch_global_sum(b200rk_ctx* c, double* out) {
  if (c->world > 1 && !c->p2p) {  // fallback: one ncclAllReduce(sum, 1 x f64) per attempt + 8-byte D2H
    NCCL_TRY(c, g_nccl.AllReduce(c->d_result, c->d_result, 1, ncclDouble, ncclSum, c->comm, c->stream));
    c->collectives++;
    CUDA_TRY(c, cudaMemcpyAsync(c->h_result, c->d_result, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    *out = *(volatile double*)c->h_result;
    return B200RK_OK;
  }
  // single GPU, or sharded with the all-reduce done inside the kernel over the peer mailboxes
  const unsigned long long want = c->seq;
  volatile unsigned long long* flag = c->h_seq;
  if (c->spin_readback) {
    for (long spins = 0; (*flag & ~kPeerTimeoutFlag) != want; ++spins) {
      __builtin_ia32_pause();
      if (spins > 2000000L) {  // ~10 ms: far beyond any kernel of this path at sane sizes -> blocking wait
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        if ((*flag & ~kPeerTimeoutFlag) != want) return fail(c, B200RK_ECUDA, "reduction result never arrived");
        break;
      }
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
  } else {
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  }
  if (*flag & kPeerTimeoutFlag) return fail(c, B200RK_ENCCL, "peer mailbox all-reduce timed out: a rank did not publish its partial sum");
  *out = *(volatile double*)c->h_result;
  return B200RK_OK;
}

template <int OP>
static int launch_ewise_t(b200rk_ctx* c, const double* a, const double* b, double s, double* out, size_t n, int cls) {
  if (n == 0) return B200RK_OK;
  constexpr int W = 2, U = 2;
  const int streams = EwiseArity<OP>::binary ? 3 : 2;
  ProfScope ps(c, cls, 8.0 * double(n) * streams);
  unsigned grid = grid_for(c, n / W, kThreads * U);
  if (l2_on(c, n) && cls == B200RK_K_RHS && out != a && out != b) {  // 256-bit accesses: the only width the L2 modifiers accept
    grid = grid_for(c, n / 4, kThreads * 2);
    ewise_kernel<OP, 4, 2, kThreads, 1><<<grid, kThreads, 0, c->stream>>>(a, b, s, out, n);
  } else {
    ewise_kernel<OP, W, U, kThreads><<<grid, kThreads, 0, c->stream>>>(a, b, s, out, n);
  }
  CUDA_TRY(c, cudaGetLastError());
  return B200RK_OK;
}

int launch_hermite(b200rk_ctx* c, const double* y1, const double* dy1, const double* y2, const double* dy2,
                          double h00, double hA, double h01, double hB, double* out, size_t n) {
  if (n == 0) return B200RK_OK;
  ProfScope ps(c, B200RK_K_OTHER, 8.0 * double(n) * 5);
  unsigned grid = grid_for(c, n / 2, kThreads);
  hermite_kernel<2, kThreads><<<grid, kThreads, 0, c->stream>>>(y1, dy1, y2, dy2, h00, hA, h01, hB, out, n);
  CUDA_TRY(c, cudaGetLastError());
  return B200RK_OK;
}

int launch_ewise(b200rk_ctx* c, int op, const double* a, const double* b, double s, double* out, size_t n, int cls) {
  switch (op) {
    case EW_ADD: return launch_ewise_t<EW_ADD>(c, a, b, s, out, n, cls);
    case EW_SUB: return launch_ewise_t<EW_SUB>(c, a, b, s, out, n, cls);
    case EW_HMUL: return launch_ewise_t<EW_HMUL>(c, a, b, s, out, n, cls);
    case EW_HDIV: return launch_ewise_t<EW_HDIV>(c, a, b, s, out, n, cls);
    case EW_SCALE: return launch_ewise_t<EW_SCALE>(c, a, b, s, out, n, cls);
    case EW_ADD_SCALAR: return launch_ewise_t<EW_ADD_SCALAR>(c, a, b, s, out, n, cls);
    case EW_NEG: return launch_ewise_t<EW_NEG>(c, a, b, s, out, n, cls);
    case EW_ABS: return launch_ewise_t<EW_ABS>(c, a, b, s, out, n, cls);
    case EW_DIV_SCALAR: return launch_ewise_t<EW_DIV_SCALAR>(c, a, b, s, out, n, cls);
    case EW_NEG_HMUL: return launch_ewise_t<EW_NEG_HMUL>(c, a, b, s, out, n, cls);
    case EW_FILL: return launch_ewise_t<EW_FILL>(c, a, b, s, out, n, cls);
  }
  return fail(c, B200RK_EINVAL, "unknown element-wise op");
}

// y + dt/6*(k1 + 2*(k2+k3) + k4) (ode.nim:188); c6 = dt/6.0 computed by the caller
int launch_rk4_final(b200rk_ctx* c, const double* y, const double* k1, const double* k2, const double* k3, const double* k4,
                     double c6, double* out, size_t n) {
  if (!n) return B200RK_OK;
  ProfScope ps(c, B200RK_K_STAGE, 8.0 * double(n) * 6);
  if (c->vec_width == 4) {
    unsigned grid = grid_for(c, n / 4, kThreads);
    rk4_final_kernel<4, 1, kThreads><<<grid, kThreads, 0, c->stream>>>(y, k1, k2, k3, k4, c6, out, n);
  } else {
    unsigned grid = grid_for(c, n / 2, kThreads);
    rk4_final_kernel<2, 1, kThreads><<<grid, kThreads, 0, c->stream>>>(y, k1, k2, k3, k4, c6, out, n);
  }
  CUDA_TRY(c, cudaGetLastError());
  return B200RK_OK;
}

int launch_lorenz96(b200rk_ctx* c, const double* y, const double* left2, const double* right1, double F, double* out, size_t n) {
  ProfScope ps(c, B200RK_K_RHS, 8.0 * double(n) * 2);
  unsigned grid = grid_for(c, n / 2, kThreads);
  lorenz96_kernel<kThreads><<<grid, kThreads, 0, c->stream>>>(y, left2, right1, F, out, n);
  CUDA_TRY(c, cudaGetLastError());
  return B200RK_OK;
}

// sum(v) into the context's reduction scratch; fetch_global_sum() brings it to the host
int launch_sum(b200rk_ctx* c, const double* a, size_t n) {
  ProfScope ps(c, B200RK_K_OTHER, 8.0 * double(n));
  unsigned grid = grid_for(c, n / 2, kThreads * 4);
  grid = std::min(grid, (unsigned)(c->sm_count * 8));
  TRY(ensure_partials(c, grid));
  sum_kernel<2, kThreads><<<grid, kThreads, 0, c->stream>>>(a, n, reduce_scratch(c));
  CUDA_TRY(c, cudaGetLastError());
  return B200RK_OK;
}
