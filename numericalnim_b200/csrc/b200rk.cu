// b200rk.cu — host side of libb200rk.so: context, sharded device vectors, kernel launchers, the
// integrator executor (one attempt = S-1 right-hand-side callbacks + S-1 stage_kernel launches + one
// finish_kernel launch + an 8-byte read-back), the ODESolver driver and the extern "C" surface declared
// in include/b200rk.h. Host control flow restates numericalnim's ode.nim:57-76 (retry loop),
// ode.nim:471-586 (driver) and ode.nim:589-651 (dispatch); all O(N) arithmetic runs in kernels.cuh.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types only: the library is bound at run time, see NcclApi below

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200rk.h"
#include "kernels.cuh"
#include "methods.h"

using namespace b200rk;

// =====================================================================================================
// context / vector objects
// =====================================================================================================
static thread_local std::string g_thread_err;

struct ProfRec {
  int cls;
  double bytes;
  cudaEvent_t a, b;
};

struct b200rk_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  int rank = 0, world = 1;
  ncclComm_t comm = nullptr;
  // reduction scratch
  double* d_partials = nullptr;
  size_t partials_cap = 0;
  unsigned int* d_ticket = nullptr;
  double* d_result = nullptr;   // device scalar (allreduce buffer)
  double* h_result = nullptr;   // pinned + mapped host scalar
  double* h_result_dev = nullptr;  // device alias of h_result
  double* d_halo = nullptr;        // 3 doubles: stencil halo of the sharded Lorenz-96 right-hand side
  unsigned long long* h_seq = nullptr;      // pinned + mapped: sequence word of the last finished reduction
  unsigned long long* h_seq_dev = nullptr;  // device alias
  unsigned long long seq = 0;               // last sequence number handed to a reducing launch
  bool spin_readback = true;                // poll h_seq instead of cudaStreamSynchronize (single GPU)
  // in-kernel all-reduce of the error norm over peer mailboxes (multi-GPU)
  unsigned long long* d_mail = nullptr;     // this rank's mailbox
  unsigned long long* peer_mail[kMaxPeers] = {nullptr};
  bool peer_opened[kMaxPeers] = {false};
  bool p2p = false;
  std::string p2p_note;
  // workspace pool (free vectors by global length)
  std::vector<b200rk_vec*> pool;
  size_t pool_budget_bytes = (size_t)48 << 30;
  // knobs
  int vec_width = 4;
  int ctas_per_sm = 0;         // stage/element-wise kernels: 0 = one tile per CTA (measured best, profiles/)
  int finish_ctas_per_sm = 2;  // reducing kernels: persistent grid, one partial per CTA (measured best)
  int fused_ctas_per_sm = 4;   // fused pointwise attempt kernel (54 registers -> 4 CTAs/SM resident; measured best)
  bool fuse_pointwise = true;  // element-local built-in RHS: whole attempt in one kernel
  bool fuse_stencil = true;    // built-in Lorenz-96 (single GPU): stage accumulate + stencil RHS in one kernel
  int l2_hints = -1;           // producer stores evict_last / streams evict_first: -1 auto (vector <= 0.65 L2), 0 off, 1 on
  size_t l2_bytes = 126u << 20;
  bool strict_zeros = false;
  bool profile = false;
  // counters
  int64_t launches = 0, collectives = 0;
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_free;
  mutable std::string err;
};

struct b200rk_vec {
  b200rk_ctx* ctx;
  size_t n_global, offset, n_local;
  double* d;
};

static int fail(const b200rk_ctx* ctx, int code, const std::string& msg) {
  g_thread_err = msg;
  if (ctx) ctx->err = msg;
  return code;
}
#define CUDA_TRY(ctx, expr)                                                                         \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return fail(ctx, B200RK_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));           \
  } while (0)
#define NCCL_TRY(ctx, expr)                                                                         \
  do {                                                                                              \
    ncclResult_t _e = (expr);                                                                       \
    if (_e != ncclSuccess)                                                                          \
      return fail(ctx, B200RK_ENCCL, std::string(#expr) + ": " + g_nccl.GetErrorString(_e));        \
  } while (0)
#define TRY(expr)                      \
  do {                                 \
    int _rc = (expr);                  \
    if (_rc != B200RK_OK) return _rc;  \
  } while (0)

static inline double nim_min(double x, double y) { return (x <= y) ? x : y; }  // Nim system.min
static inline double nim_max(double x, double y) { return (y <= x) ? x : y; }  // Nim system.max

// ---- NCCL, bound lazily ------------------------------------------------------------------------------
// libb200rk.so does not link libnccl: a process may already hold a libnccl.so.2 (PyTorch bundles its own,
// newer than the system one, and resolves symbols against whichever copy was loaded first). The first
// distributed call binds, in order: a copy already in the process, $B200RK_NCCL_LIB, then the default
// search path. Single-GPU use never touches NCCL.
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  std::string where;
};
static NcclApi g_nccl;

static int nccl_bind(const b200rk_ctx* ctx) {
  if (g_nccl.handle) return B200RK_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  std::string where = "already loaded in the process";
  if (!h) {
    if (const char* p = getenv("B200RK_NCCL_LIB")) { h = dlopen(p, RTLD_NOW | RTLD_GLOBAL); where = p; }
  }
  if (!h) { h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL); where = "libnccl.so.2 (default search path)"; }
  if (!h) return fail(ctx, B200RK_ENCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
  NcclApi a;
  a.handle = h; a.where = where;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
  a.AllReduce = (decltype(a.AllReduce))dlsym(h, "ncclAllReduce");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
  a.GetVersion = (decltype(a.GetVersion))dlsym(h, "ncclGetVersion");
  a.Send = (decltype(a.Send))dlsym(h, "ncclSend");
  a.Recv = (decltype(a.Recv))dlsym(h, "ncclRecv");
  a.GroupStart = (decltype(a.GroupStart))dlsym(h, "ncclGroupStart");
  a.GroupEnd = (decltype(a.GroupEnd))dlsym(h, "ncclGroupEnd");
  a.AllGather = (decltype(a.AllGather))dlsym(h, "ncclAllGather");
  if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.GetErrorString || !a.Send || !a.Recv ||
      !a.GroupStart || !a.GroupEnd || !a.AllGather)
    return fail(ctx, B200RK_ENCCL, "libnccl.so.2 lacks a required symbol");
  g_nccl = a;
  return B200RK_OK;
}

// ---- profiling: one CUDA-event pair per launch on the context stream --------------------------------
struct ProfScope {
  b200rk_ctx* c;
  cudaEvent_t a = nullptr, b = nullptr;
  int cls;
  double bytes;
  ProfScope(b200rk_ctx* ctx, int cls_, double bytes_) : c(ctx), cls(cls_), bytes(bytes_) {
    c->launches++;
    if (!c->profile) return;
    a = take();
    b = take();
    cudaEventRecord(a, c->stream);
  }
  cudaEvent_t take() {
    if (!c->ev_free.empty()) { cudaEvent_t e = c->ev_free.back(); c->ev_free.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
  ~ProfScope() {
    if (!a) return;
    cudaEventRecord(b, c->stream);
    c->prof.push_back(ProfRec{cls, bytes, a, b});
  }
};

// =====================================================================================================
// kernel launchers
// =====================================================================================================
static constexpr int kThreads = 256;

template <int M, int W>
struct StageUnroll {  // keep ~16 doubles of loads in flight per thread
  static constexpr int raw = 16 / ((M + 1) * W);
  static constexpr int value = raw < 1 ? 1 : (raw > 4 ? 4 : raw);
};

// L2 hand-off between producer and consumer kernels (kernels.cuh: L2Policy): worthwhile while one vector
// fits comfortably in the L2 (measured: +13 % at 2^22, +16 % at 2^23, -2 % at 2^24 = 128 MiB per vector).
static inline bool l2_on(const b200rk_ctx* c, size_t n_local) {
  if (c->l2_hints >= 0) return c->l2_hints != 0;
  return n_local * sizeof(double) <= (size_t)(0.65 * (double)c->l2_bytes);
}

static inline unsigned grid_for(const b200rk_ctx* c, size_t nvec, int per_block, int ctas_per_sm = -1) {
  size_t tiles = (nvec + per_block - 1) / per_block;
  if (tiles == 0) tiles = 1;
  if (ctas_per_sm < 0) ctas_per_sm = c->ctas_per_sm;
  if (ctas_per_sm > 0) tiles = std::min(tiles, (size_t)ctas_per_sm * c->sm_count);
  return (unsigned)std::min(tiles, (size_t)0x7fffffff);
}

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
static int launch_stage(b200rk_ctx* c, int m, const double* y, const double* const* k, const double* w, double cc,
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

static int ensure_partials(b200rk_ctx* c, size_t blocks) {
  if (blocks <= c->partials_cap) return B200RK_OK;
  if (c->d_partials) CUDA_TRY(c, cudaFree(c->d_partials));
  size_t cap = std::max(blocks, (size_t)1 << 16);
  CUDA_TRY(c, cudaMalloc(&c->d_partials, cap * sizeof(double)));
  c->partials_cap = cap;
  return B200RK_OK;
}

static ReduceScratch reduce_scratch(b200rk_ctx* c) {
  ReduceScratch rs;
  rs.partials = c->d_partials;
  rs.ticket = c->d_ticket;
  rs.result = c->d_result;
  const bool in_kernel_collective = c->world > 1 && c->p2p;
  rs.result_host = (c->world == 1 || in_kernel_collective) ? c->h_result_dev : nullptr;
  rs.seq_host = c->h_seq_dev;
  rs.seq = ++c->seq;
  rs.mail.world = in_kernel_collective ? c->world : 1;
  rs.mail.rank = c->rank;
  for (int p = 0; p < kMaxPeers; ++p) rs.mail.box[p] = (in_kernel_collective && p < c->world) ? c->peer_mail[p] : nullptr;
  if (in_kernel_collective) c->collectives++;
  return rs;
}

// Peer mailboxes for the in-kernel all-reduce of the error norm (kernels.cuh: peer_allreduce). Every rank
// cudaMalloc's a 512-byte mailbox, the CUDA-IPC handles travel through one ncclAllGather on the
// communicator we already have, and each rank maps its peers' mailboxes (NVLink/NVSwitch peer access). The
// outcome is agreed by an ncclAllReduce(min): either every rank uses the mailboxes or every rank falls back to
// ncclAllReduce per attempt.
static int setup_p2p(b200rk_ctx* c) {
  c->p2p = false;
  if (const char* e = getenv("B200RK_P2P")) if (atoi(e) == 0) { c->p2p_note = "disabled by B200RK_P2P=0"; return B200RK_OK; }
  if (c->world > kMaxPeers) { c->p2p_note = "world larger than kMaxPeers"; return B200RK_OK; }
  const size_t mail_bytes = 2 * kMaxPeers * 2 * sizeof(unsigned long long);
  CUDA_TRY(c, cudaMalloc(&c->d_mail, mail_bytes));
  CUDA_TRY(c, cudaMemset(c->d_mail, 0, mail_bytes));
  cudaIpcMemHandle_t mine;
  int ok = 1;
  if (cudaIpcGetMemHandle(&mine, c->d_mail) != cudaSuccess) { ok = 0; cudaGetLastError(); std::memset(&mine, 0, sizeof(mine)); }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
  char* d_handles = nullptr;
  int* d_flag = nullptr;
  CUDA_TRY(c, cudaMalloc(&d_handles, 64 * (size_t)c->world));
  CUDA_TRY(c, cudaMalloc(&d_flag, sizeof(int)));
  CUDA_TRY(c, cudaMemcpyAsync(d_handles + 64 * (size_t)c->rank, &mine, 64, cudaMemcpyHostToDevice, c->stream));
  NCCL_TRY(c, g_nccl.AllGather(d_handles + 64 * (size_t)c->rank, d_handles, 64, ncclChar, c->comm, c->stream));
  std::vector<cudaIpcMemHandle_t> all(c->world);
  CUDA_TRY(c, cudaMemcpyAsync(all.data(), d_handles, 64 * (size_t)c->world, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  for (int p = 0; p < c->world && ok; ++p) {
    if (p == c->rank) { c->peer_mail[p] = c->d_mail; continue; }
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      c->p2p_note = std::string("cudaIpcOpenMemHandle failed: ") + cudaGetErrorString(cudaGetLastError());
      ok = 0;
    } else {
      c->peer_mail[p] = static_cast<unsigned long long*>(ptr);
      c->peer_opened[p] = true;
    }
  }
  CUDA_TRY(c, cudaMemcpyAsync(d_flag, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  NCCL_TRY(c, g_nccl.AllReduce(d_flag, d_flag, 1, ncclInt, ncclMin, c->comm, c->stream));
  int agreed = 0;
  CUDA_TRY(c, cudaMemcpyAsync(&agreed, d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  cudaFree(d_handles);
  cudaFree(d_flag);
  c->p2p = agreed != 0;
  if (c->p2p) c->p2p_note = "peer mailboxes mapped over CUDA IPC";
  else if (c->p2p_note.empty()) c->p2p_note = "a peer could not map the mailboxes";
  return B200RK_OK;
}

struct FinishPlan {
  int nk = 0;
  const double* k[kMaxTerms];
  double wb[kMaxTerms], wbh[kMaxTerms];
  uint32_t mask_b = 0, mask_bh = 0;
  double cb = 0, cbh = 0, absTol = 0, relTol = 0;
  const double* y = nullptr;
  double* ynew_out = nullptr;
  double* err_out = nullptr;
  bool direct = false;
  int ynew_mode = 0;
  size_t n = 0;
};

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
static int launch_finish(b200rk_ctx* c, const FinishPlan& p) {
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
static int fetch_global_sum(b200rk_ctx* c, double* out) {
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
static int launch_ewise(b200rk_ctx* c, const double* a, const double* b, double s, double* out, size_t n, int cls) {
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

static int launch_hermite(b200rk_ctx* c, const double* y1, const double* dy1, const double* y2, const double* dy2,
                          double h00, double hA, double h01, double hB, double* out, size_t n) {
  if (n == 0) return B200RK_OK;
  ProfScope ps(c, B200RK_K_OTHER, 8.0 * double(n) * 5);
  unsigned grid = grid_for(c, n / 2, kThreads);
  hermite_kernel<2, kThreads><<<grid, kThreads, 0, c->stream>>>(y1, dy1, y2, dy2, h00, hA, h01, hB, out, n);
  CUDA_TRY(c, cudaGetLastError());
  return B200RK_OK;
}

// =====================================================================================================
// vectors
// =====================================================================================================
static void shard_range(size_t n, int rank, int world, size_t* off, size_t* len) {
  size_t chunk = (n + world - 1) / world;
  chunk = (chunk + 3) / 4 * 4;  // keep shard boundaries 32-byte aligned in the global index space
  size_t lo = std::min(n, (size_t)rank * chunk), hi = std::min(n, (size_t)(rank + 1) * chunk);
  *off = lo;
  *len = hi - lo;
}

static int vec_alloc(b200rk_ctx* c, size_t n_global, b200rk_vec** out) {
  for (size_t i = 0; i < c->pool.size(); ++i) {
    if (c->pool[i]->n_global == n_global) {
      *out = c->pool[i];
      c->pool.erase(c->pool.begin() + i);
      return B200RK_OK;
    }
  }
  b200rk_vec* v = new b200rk_vec{c, n_global, 0, 0, nullptr};
  shard_range(n_global, c->rank, c->world, &v->offset, &v->n_local);
  cudaError_t e = cudaMalloc(&v->d, std::max<size_t>(v->n_local, 4) * sizeof(double));
  if (e != cudaSuccess) {
    delete v;
    return fail(c, e == cudaErrorMemoryAllocation ? B200RK_ENOMEM : B200RK_ECUDA,
                std::string("cudaMalloc: ") + cudaGetErrorString(e));
  }
  *out = v;
  return B200RK_OK;
}
static void vec_release(b200rk_vec* v) {  // back to the pool
  if (v) v->ctx->pool.push_back(v);
}
static int check_same(const b200rk_ctx* c, const b200rk_vec* a, const b200rk_vec* b) {
  if (!a || !b) return fail(c, B200RK_EINVAL, "null vector");
  if (a->n_global != b->n_global) return fail(c, B200RK_EINVAL, "Vectors must have the same size.");  // utils.nim:26
  return B200RK_OK;
}

// =====================================================================================================
// built-in right-hand sides
// =====================================================================================================
struct BuiltinRhs {
  b200rk_ctx* ctx;
  int kind;
  double scalar;
  const b200rk_vec* lambda;
};

static int builtin_rhs_fn(double /*t*/, const b200rk_vec* y, b200rk_vec* dydt, void* user) {
  BuiltinRhs* r = static_cast<BuiltinRhs*>(user);
  b200rk_ctx* c = r->ctx;
  const size_t n = y->n_local;
  switch (r->kind) {
    case B200RK_RHS_SCALE:
      return launch_ewise<EW_SCALE>(c, y->d, nullptr, r->scalar, dydt->d, n, B200RK_K_RHS);
    case B200RK_RHS_DIAG_LINEAR:
      TRY(check_same(c, y, r->lambda));
      return launch_ewise<EW_NEG_HMUL>(c, r->lambda->d, y->d, 0.0, dydt->d, n, B200RK_K_RHS);
    case B200RK_RHS_LORENZ96: {
      if (y->n_global < 4) return fail(c, B200RK_EINVAL, "lorenz96 needs n >= 4");
      const double *left2 = y->d + n - 2, *right1 = y->d;  // single GPU: the cyclic neighbours are in the vector itself
      if (c->world > 1) {
        // Sharded stencil: 3-element halo per evaluation over NVLink. Each rank sends its
        // first element to the left neighbour and its last two to the right neighbour (ring), on the
        // context stream, so the exchange is ordered with the producing and consuming kernels.
        if (n < 2) return fail(c, B200RK_EINVAL, "lorenz96: every shard needs at least 2 elements");
        const int left = (c->rank + c->world - 1) % c->world, right = (c->rank + 1) % c->world;
        NCCL_TRY(c, g_nccl.GroupStart());
        NCCL_TRY(c, g_nccl.Send(y->d, 1, ncclDouble, left, c->comm, c->stream));
        NCCL_TRY(c, g_nccl.Send(y->d + n - 2, 2, ncclDouble, right, c->comm, c->stream));
        NCCL_TRY(c, g_nccl.Recv(c->d_halo + 2, 1, ncclDouble, right, c->comm, c->stream));
        NCCL_TRY(c, g_nccl.Recv(c->d_halo, 2, ncclDouble, left, c->comm, c->stream));
        NCCL_TRY(c, g_nccl.GroupEnd());
        c->collectives++;
        left2 = c->d_halo;
        right1 = c->d_halo + 2;
      }
      ProfScope ps(c, B200RK_K_RHS, 8.0 * double(n) * 2);
      unsigned grid = grid_for(c, n / 2, kThreads);
      lorenz96_kernel<kThreads><<<grid, kThreads, 0, c->stream>>>(y->d, left2, right1, r->scalar, dydt->d, n);
      CUDA_TRY(c, cudaGetLastError());
      return B200RK_OK;
    }
  }
  return fail(c, B200RK_EINVAL, "unknown builtin rhs");
}

// =====================================================================================================
// integrator executor
// =====================================================================================================
struct RhsCall {
  b200rk_rhs_fn f;
  void* user;
  bool negate_time;  // backward pass: g(t, y) = -f(-t, y)   (ode.nim:545)
  int64_t* evals;
};

static int eval_rhs(b200rk_ctx* c, const RhsCall& r, double t, const b200rk_vec* y, b200rk_vec* out) {
  if (r.evals) ++*r.evals;
  int rc = r.f(r.negate_time ? -t : t, y, out, r.user);
  if (rc != 0) {
    if (r.f == &builtin_rhs_fn) return rc;  // our own launcher already recorded the error
    return fail(c, B200RK_ECALLBACK, "right-hand side callback returned " + std::to_string(rc));
  }
  if (r.negate_time) return launch_ewise<EW_NEG>(c, out->d, nullptr, 0.0, out->d, out->n_local, B200RK_K_OTHER);
  return B200RK_OK;
}

// Compact a reference row into launch arguments; zero weights dropped unless strict.
static int gather_row(const b200rk_ctx* c, const Row& row, b200rk_vec* const* k /*1-based*/, const double** kp, double* w) {
  int m = 0;
  for (int j = 0; j < row.m; ++j) {
    if (row.w[j] == 0.0 && !c->strict_zeros && row.m > 1) continue;
    kp[m] = k[row.idx[j]]->d;
    w[m] = row.w[j];
    ++m;
  }
  if (m == 0) {  // all-zero row: keep the first term so the kernel still has one stream
    kp[0] = k[row.idx[0]]->d; w[0] = row.w[0]; m = 1;
  }
  return m;
}

static int run_row(b200rk_ctx* c, const Row& row, double cfac, bool chain, double dt, const b200rk_vec* y,
                   b200rk_vec* const* k, b200rk_vec* out) {
  const double* kp[kMaxTerms];
  double w[kMaxTerms];
  int m = gather_row(c, row, k, kp, w);
  double cc = (cfac == 1.0) ? dt : cfac * dt;
  if (chain) {
    for (int j = 0; j < m; ++j) w[j] = w[j] * dt;  // (-1*dt), (2*dt): exact
    cc = 0.0;
  }
  return launch_stage(c, m, y->d, kp, w, cc, chain, out->d, y->n_local);
}

// Stage row fused with the built-in Lorenz-96 right-hand side (kernels.cuh: stage_l96_kernel): writes
// k_s = f(stage input) directly; the stage input itself is stored only when `in_out` is given.
template <int M>
static int launch_stage_l96_m(b200rk_ctx* c, const double* y, const double* const* kp, const double* w, double cc,
                              double F, double sgn, double* in_out, double* kout, size_t n) {
  StageArgs<M> a;
  a.y = y; a.c = cc; a.out = in_out; a.n = n;
  for (int j = 0; j < M; ++j) { a.k[j] = kp[j]; a.w[j] = w[j]; }
  ProfScope ps(c, B200RK_K_STAGE, 8.0 * double(n) * (M + 2 + (in_out ? 1 : 0)));
  const unsigned grid = (unsigned)((n + kThreads * 4 - 1) / (kThreads * 4));
  stage_l96_kernel<M, kThreads><<<grid, kThreads, 0, c->stream>>>(a, F, sgn, kout);
  CUDA_TRY(c, cudaGetLastError());
  return B200RK_OK;
}
static int run_row_l96(b200rk_ctx* c, const Row& row, double cfac, double dt, const b200rk_vec* y, b200rk_vec* const* k,
                       double F, bool negate, b200rk_vec* in_out, b200rk_vec* kout) {
  const double* kp[kMaxTerms];
  double w[kMaxTerms];
  const int m = gather_row(c, row, k, kp, w);
  const double cc = (cfac == 1.0) ? dt : cfac * dt;
  const double sgn = negate ? -1.0 : 1.0;
  double* io = in_out ? in_out->d : nullptr;
  const size_t n = y->n_local;
  switch (m) {
    case 1: return launch_stage_l96_m<1>(c, y->d, kp, w, cc, F, sgn, io, kout->d, n);
    case 2: return launch_stage_l96_m<2>(c, y->d, kp, w, cc, F, sgn, io, kout->d, n);
    case 3: return launch_stage_l96_m<3>(c, y->d, kp, w, cc, F, sgn, io, kout->d, n);
    case 4: return launch_stage_l96_m<4>(c, y->d, kp, w, cc, F, sgn, io, kout->d, n);
    case 5: return launch_stage_l96_m<5>(c, y->d, kp, w, cc, F, sgn, io, kout->d, n);
    case 6: return launch_stage_l96_m<6>(c, y->d, kp, w, cc, F, sgn, io, kout->d, n);
    case 7: return launch_stage_l96_m<7>(c, y->d, kp, w, cc, F, sgn, io, kout->d, n);
    case 8: return launch_stage_l96_m<8>(c, y->d, kp, w, cc, F, sgn, io, kout->d, n);
    case 9: return launch_stage_l96_m<9>(c, y->d, kp, w, cc, F, sgn, io, kout->d, n);
  }
  return fail(c, B200RK_EINVAL, "stage_l96: m must be in 1..9");
}

static int plan_finish(const b200rk_ctx* c, const MethodDef& md, double dt, double absTol, double relTol,
                       const b200rk_vec* y, b200rk_vec* const* k, b200rk_vec* ynew, bool ynew_ready, double* err_out,
                       FinishPlan* p) {
  // Union of the derivative streams the two rows touch, slots ascending in stage index so that the
  // kernel's left-to-right walk over slots reproduces the reference's association (both rows list their
  // terms in ascending stage order in ode.nim).
  p->direct = md.err_direct;
  const bool use_b = !(md.err_direct && ynew_ready);
  double wb_of[kMaxStages + 1] = {0}, wbh_of[kMaxStages + 1] = {0};
  bool in_b[kMaxStages + 1] = {false}, in_bh[kMaxStages + 1] = {false};
  int prev = 0;
  if (use_b)
    for (int j = 0; j < md.b.m; ++j) {
      if (md.b.idx[j] <= prev) return fail(c, B200RK_EINVAL, "finish: b row not in ascending stage order");
      prev = md.b.idx[j];
      if (md.b.w[j] == 0.0 && !c->strict_zeros) continue;
      in_b[md.b.idx[j]] = true; wb_of[md.b.idx[j]] = md.b.w[j];
    }
  prev = 0;
  for (int j = 0; j < md.bhat.m; ++j) {
    if (md.bhat.idx[j] <= prev) return fail(c, B200RK_EINVAL, "finish: bhat row not in ascending stage order");
    prev = md.bhat.idx[j];
    if (md.bhat.w[j] == 0.0 && !c->strict_zeros) continue;
    in_bh[md.bhat.idx[j]] = true; wbh_of[md.bhat.idx[j]] = md.bhat.w[j];
  }
  for (int s = 1; s <= md.stages; ++s) {
    if (!in_b[s] && !in_bh[s]) continue;
    p->k[p->nk] = k[s]->d; p->wb[p->nk] = wb_of[s]; p->wbh[p->nk] = wbh_of[s];
    if (in_b[s]) p->mask_b |= 1u << p->nk;
    if (in_bh[s]) p->mask_bh |= 1u << p->nk;
    ++p->nk;
  }
  if (p->nk == 0) return fail(c, B200RK_EINVAL, "finish: empty rows");
  p->cb = (md.b_cfac == 1.0) ? dt : dt * md.b_cfac;
  p->cbh = (md.bhat_cfac == 1.0) ? dt : dt * md.bhat_cfac;
  p->absTol = absTol; p->relTol = relTol; p->n = y->n_local; p->err_out = err_out;
  if (md.err_direct && ynew_ready) { p->ynew_mode = 2; p->y = ynew->d; }
  else if (ynew_ready) { p->ynew_mode = 0; p->y = y->d; }
  else { p->ynew_mode = 1; p->y = y->d; p->ynew_out = ynew->d; }
  if (md.err_direct && !ynew_ready) return fail(c, B200RK_EINVAL, "finish: direct-error methods need yNew first");
  return B200RK_OK;
}

struct Workspace {
  b200rk_ctx* c;
  std::vector<b200rk_vec*> held;
  explicit Workspace(b200rk_ctx* ctx) : c(ctx) {}
  int get(size_t n, b200rk_vec** out) {
    TRY(vec_alloc(c, n, out));
    held.push_back(*out);
    return B200RK_OK;
  }
  ~Workspace() { for (auto* v : held) vec_release(v); }
};

struct StepCounters { int64_t attempts = 0, rejected = 0, limiter_hits = 0; };

// ---- fused attempt for element-local built-in right-hand sides (kernels.cuh: fused_attempt_kernel) ----
static bool pointwise_kind(const RhsCall& rhs, int* pw_kind, const BuiltinRhs** br) {
  if (rhs.f != &builtin_rhs_fn) return false;
  const BuiltinRhs* r = static_cast<const BuiltinRhs*>(rhs.user);
  if (r->kind == B200RK_RHS_SCALE) *pw_kind = PW_SCALE;
  else if (r->kind == B200RK_RHS_DIAG_LINEAR) *pw_kind = PW_DIAG;
  else return false;
  *br = r;
  return true;
}
static bool method_fusable(const MethodDef& md) {
  if (md.rk4_final) return true;
  if (!(md.adaptive && md.k1_from_fsal && md.fsal_out == md.stages && (md.stages == 7 || md.stages == 9))) return false;
  for (int s = 2; s <= md.stages; ++s)
    if (md.a_cfac[s] != 1.0 || md.a_chain[s] || md.a[s].m != s - 1) return false;
  return md.b_cfac == 1.0 && md.bhat_cfac == 1.0;
}

static uint32_t row_mask(const b200rk_ctx* c, const Row& row, double* w_dense, int width) {
  uint32_t mask = 0;
  for (int j = 0; j < width; ++j) w_dense[j] = 0.0;
  int kept = 0;
  for (int j = 0; j < row.m; ++j) {
    w_dense[row.idx[j] - 1] = row.w[j];
    if (row.w[j] == 0.0 && !c->strict_zeros && row.m > 1) continue;
    mask |= 1u << (row.idx[j] - 1);
    ++kept;
  }
  if (!kept) mask |= 1u << (row.idx[0] - 1);  // same rule as gather_row: an all-zero row keeps its first term
  return mask;
}

template <int PAT, int KIND>
static int launch_fused_cfg(b200rk_ctx* c, const FusedArgs<Pattern<PAT>::S>& a) {
  const size_t n = a.n;
  if (c->vec_width == 4) {
    unsigned grid = grid_for(c, n / 4, kThreads, c->fused_ctas_per_sm);
    TRY(ensure_partials(c, grid));
    if (l2_on(c, n)) fused_attempt_kernel<PAT, KIND, 4, kThreads, 1><<<grid, kThreads, 0, c->stream>>>(a);
    else fused_attempt_kernel<PAT, KIND, 4, kThreads><<<grid, kThreads, 0, c->stream>>>(a);
  } else {
    unsigned grid = grid_for(c, n / 2, kThreads, c->fused_ctas_per_sm);
    TRY(ensure_partials(c, grid));
    fused_attempt_kernel<PAT, KIND, 2, kThreads><<<grid, kThreads, 0, c->stream>>>(a);
  }
  CUDA_TRY(c, cudaGetLastError());
  return B200RK_OK;
}

// Runtime masks of the method (same dropping rules as gather_row / plan_finish) must equal the kernel's
// compile-time pattern; otherwise the caller falls back to the pipeline.
template <int PAT>
static bool pattern_matches(const b200rk_ctx* c, const MethodDef& md) {
  constexpr int S = Pattern<PAT>::S;
  if (md.stages != S || md.err_direct != Pattern<PAT>::direct || md.ynew_is_last_stage_input != Pattern<PAT>::last) return false;
  double scratch[kMaxStages];
  for (int s = 2; s <= S; ++s)
    if (row_mask(c, md.a[s], scratch, S - 1) != Pattern<PAT>::a(s - 2)) return false;
  uint32_t bm = 0, bhm = 0;
  for (int j = 0; j < md.b.m; ++j) if (md.b.w[j] != 0.0 || c->strict_zeros) bm |= 1u << (md.b.idx[j] - 1);
  for (int j = 0; j < md.bhat.m; ++j) if (md.bhat.w[j] != 0.0 || c->strict_zeros) bhm |= 1u << (md.bhat.idx[j] - 1);
  if (!Pattern<PAT>::last && bm != Pattern<PAT>::b()) return false;
  return bhm == Pattern<PAT>::bh();
}

static int fused_pattern_of(const b200rk_ctx* c, const MethodDef& md) {
  if (pattern_matches<PAT_DOPRI54>(c, md) && !std::strcmp(md.name, "dopri54")) return PAT_DOPRI54;
  if (pattern_matches<PAT_DOPRI54_STRICT>(c, md) && !std::strcmp(md.name, "dopri54")) return PAT_DOPRI54_STRICT;
  if (pattern_matches<PAT_TSIT54>(c, md) && !std::strcmp(md.name, "tsit54")) return PAT_TSIT54;
  if (pattern_matches<PAT_VERN65>(c, md)) return PAT_VERN65;
  if (pattern_matches<PAT_VERN65_STRICT>(c, md)) return PAT_VERN65_STRICT;
  return -1;
}

template <int PAT>
static int launch_fused_pair(b200rk_ctx* c, const MethodDef& md, int pw_kind, const BuiltinRhs* br, bool negate, double dt,
                             const b200rk_options& o, const b200rk_vec* y, const b200rk_vec* fsal, b200rk_vec* y_new,
                             b200rk_vec* fsal_new) {
  constexpr int S = Pattern<PAT>::S;
  FusedArgs<S> a;
  std::memset(&a, 0, sizeof(a));
  a.y = y->d; a.k1 = fsal->d; a.lam = br->lambda ? br->lambda->d : nullptr;
  a.rhs_scalar = negate ? -br->scalar : br->scalar;   // -(y*c) == y*(-c) exactly
  a.rhs_sign = negate ? 1.0 : -1.0;                   // k = (lam*y)*sign
  for (int s = 2; s <= S; ++s) row_mask(c, md.a[s], a.a[s - 2], S - 1);
  row_mask(c, md.b, a.b, S);
  row_mask(c, md.bhat, a.bh, S);
  a.dt = dt; a.cb = dt; a.cbh = dt; a.absTol = o.absTol; a.relTol = o.relTol;
  a.ynew = y_new->d; a.ks_out = fsal_new->d; a.n = y->n_local;
  a.rs = reduce_scratch(c);
  const int streams = 4 + (pw_kind == PW_DIAG ? 1 : 0);  // y, k1 (+ lambda) read; yNew, k_S written
  ProfScope ps(c, B200RK_K_FUSED, 8.0 * double(y->n_local) * streams);
  if (pw_kind == PW_SCALE) return launch_fused_cfg<PAT, PW_SCALE>(c, a);
  return launch_fused_cfg<PAT, PW_DIAG>(c, a);
}

static int launch_fused_rk4(b200rk_ctx* c, int pw_kind, const BuiltinRhs* br, bool negate, double dt, const b200rk_vec* y,
                            b200rk_vec* y_new) {
  const size_t n = y->n_local;
  if (!n) return B200RK_OK;
  ProfScope ps(c, B200RK_K_FUSED, 8.0 * double(n) * (2 + (pw_kind == PW_DIAG ? 1 : 0)));
  const double hdt = 0.5 * dt, c6 = dt / 6.0;
  const double* lam = br->lambda ? br->lambda->d : nullptr;
  unsigned grid = grid_for(c, n / 4, kThreads, c->ctas_per_sm);
  const double cs = negate ? -br->scalar : br->scalar, sgn = negate ? 1.0 : -1.0;
  if (pw_kind == PW_SCALE) fused_rk4_kernel<PW_SCALE, 4, kThreads><<<grid, kThreads, 0, c->stream>>>(y->d, lam, cs, sgn, hdt, dt, c6, y_new->d, n);
  else fused_rk4_kernel<PW_DIAG, 4, kThreads><<<grid, kThreads, 0, c->stream>>>(y->d, lam, cs, sgn, hdt, dt, c6, y_new->d, n);
  CUDA_TRY(c, cudaGetLastError());
  return B200RK_OK;
}

// One IntegratorProc call. y, fsal read-only; y_new, fsal_new written.
static int do_step(b200rk_ctx* c, const MethodDef& md, const RhsCall& rhs, double t, const b200rk_vec* y,
                   const b200rk_vec* fsal, double dt_in, const b200rk_options& o, b200rk_vec* y_new,
                   b200rk_vec* fsal_new, double* dt_used, double* error_out, StepCounters* cnt) {
  const size_t N = y->n_global;
  const int S = md.stages;
  Workspace ws(c);
  b200rk_vec* k[kMaxStages + 1] = {nullptr};
  b200rk_vec* tmp = nullptr;
  int pw_kind = 0;
  const BuiltinRhs* br = nullptr;
  int fused_pat = -1;
  bool fused = c->fuse_pointwise && pointwise_kind(rhs, &pw_kind, &br) && method_fusable(md) &&
               (md.rk4_final || (fsal && fsal_new));
  if (fused && !md.rk4_final) { fused_pat = fused_pattern_of(c, md); fused = fused_pat >= 0; }
  const bool stencil_fused = !fused && c->fuse_stencil && c->world == 1 && rhs.f == &builtin_rhs_fn &&
                             static_cast<const BuiltinRhs*>(rhs.user)->kind == B200RK_RHS_LORENZ96 && y->n_global >= 4;
  if (fused && pw_kind == PW_DIAG) TRY(check_same(c, y, br->lambda));
  if (md.k1_from_fsal) {
    if (!fsal) return fail(c, B200RK_EINVAL, std::string(md.name) + ": FSAL vector required");
    TRY(check_same(c, y, fsal));
    k[1] = const_cast<b200rk_vec*>(fsal);
  } else if (!fused) {
    TRY(ws.get(N, &k[1]));
  }
  const bool last_input_is_ynew = md.ynew_is_last_stage_input;
  if (!fused) {
    for (int s = 2; s <= S; ++s) {
      if (s == md.fsal_out && fsal_new) k[s] = fsal_new;
      else TRY(ws.get(N, &k[s]));
    }
    if (!(last_input_is_ynew && S == 2)) TRY(ws.get(N, &tmp));
  }

  double dt = dt_in, error = 0.0;
  int limitCounter = 0;
  while (true) {
    if (cnt) cnt->attempts++;
    if (fused) {
      // element-local right-hand side: the whole attempt is one kernel (the callbacks it stands for are
      // still counted so rhs_evals matches the unfused path)
      if (rhs.evals) *rhs.evals += md.rk4_final ? 4 : (S - 1);
      if (md.rk4_final) { TRY(launch_fused_rk4(c, pw_kind, br, rhs.negate_time, dt, y, y_new)); break; }
      switch (fused_pat) {
        case PAT_DOPRI54: TRY(launch_fused_pair<PAT_DOPRI54>(c, md, pw_kind, br, rhs.negate_time, dt, o, y, fsal, y_new, fsal_new)); break;
        case PAT_DOPRI54_STRICT: TRY(launch_fused_pair<PAT_DOPRI54_STRICT>(c, md, pw_kind, br, rhs.negate_time, dt, o, y, fsal, y_new, fsal_new)); break;
        case PAT_TSIT54: TRY(launch_fused_pair<PAT_TSIT54>(c, md, pw_kind, br, rhs.negate_time, dt, o, y, fsal, y_new, fsal_new)); break;
        case PAT_VERN65: TRY(launch_fused_pair<PAT_VERN65>(c, md, pw_kind, br, rhs.negate_time, dt, o, y, fsal, y_new, fsal_new)); break;
        default: TRY(launch_fused_pair<PAT_VERN65_STRICT>(c, md, pw_kind, br, rhs.negate_time, dt, o, y, fsal, y_new, fsal_new)); break;
      }
    } else {
    if (!md.k1_from_fsal) TRY(eval_rhs(c, rhs, t, y, k[1]));
    for (int s = 2; s <= S; ++s) {
      const bool in_is_ynew = (s == S && last_input_is_ynew);
      b200rk_vec* in = in_is_ynew ? y_new : tmp;
      if (stencil_fused && !md.a_chain[s]) {
        // built-in Lorenz-96: stage input staged in shared memory, k_s written directly (no tmp round trip)
        if (rhs.evals) ++*rhs.evals;
        TRY(run_row_l96(c, md.a[s], md.a_cfac[s], dt, y, k, static_cast<const BuiltinRhs*>(rhs.user)->scalar, rhs.negate_time,
                        in_is_ynew ? y_new : nullptr, k[s]));
        continue;
      }
      TRY(run_row(c, md.a[s], md.a_cfac[s], md.a_chain[s], dt, y, k, in));
      TRY(eval_rhs(c, rhs, t + dt * md.c[s], in, k[s]));
    }
    if (!md.adaptive) {
      if (md.rk4_final) {
        ProfScope ps(c, B200RK_K_STAGE, 8.0 * double(y->n_local) * 6);
        const size_t n = y->n_local;
        if (n) {
          if (c->vec_width == 4) {
            unsigned grid = grid_for(c, n / 4, kThreads);
            rk4_final_kernel<4, 1, kThreads><<<grid, kThreads, 0, c->stream>>>(y->d, k[1]->d, k[2]->d, k[3]->d, k[4]->d, dt / 6.0, y_new->d, n);
          } else {
            unsigned grid = grid_for(c, n / 2, kThreads);
            rk4_final_kernel<2, 1, kThreads><<<grid, kThreads, 0, c->stream>>>(y->d, k[1]->d, k[2]->d, k[3]->d, k[4]->d, dt / 6.0, y_new->d, n);
          }
          CUDA_TRY(c, cudaGetLastError());
        }
      } else {
        TRY(run_row(c, md.b, md.b_cfac, false, dt, y, k, y_new));
      }
      break;
    }
    FinishPlan p;
    TRY(plan_finish(c, md, dt, o.absTol, o.relTol, y, k, y_new, last_input_is_ynew, nullptr, &p));
    TRY(launch_finish(c, p));
    }  // !fused
    double S2 = 0.0;
    TRY(fetch_global_sum(c, &S2));
    error = std::sqrt(1.0 / double(N) * S2);                                     // ode.nim:64-65
    if (error <= 1) break;                                                       // ode.nim:69-70
    if (std::isnan(error)) {
      *dt_used = dt; *error_out = error;
      return fail(c, B200RK_ENONFINITE, "error norm is NaN (the reference would loop forever here)");
    }
    if (cnt) cnt->rejected++;
    dt = dt * nim_min(4, nim_max(0.125, 0.9 * std::pow(1.0 / error, 1.0 / double(md.order_int))));  // ode.nim:71
    if (std::fabs(dt) < o.dtMin) {                                               // ode.nim:72-74
      dt = o.dtMin;
      limitCounter += 1;
      if (cnt) cnt->limiter_hits++;
    } else if (o.dtMax < std::fabs(dt)) {                                        // ode.nim:75-76
      dt = o.dtMax;
    }
    if (!(limitCounter < 2)) break;                                              // ode.nim:58
  }
  if (fsal_new && md.fsal_out == 0)  // non-FSAL steppers return (yNew, yNew, ...) (ode.nim:113,189,210)
    CUDA_TRY(c, cudaMemcpyAsync(fsal_new->d, y_new->d, y_new->n_local * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  *dt_used = dt;
  *error_out = error;
  return B200RK_OK;
}

// =====================================================================================================
// driver (ODESolver, ode.nim:471-586)
// =====================================================================================================
struct b200rk_solver {
  b200rk_ctx* c = nullptr;
  const MethodDef* md = nullptr;
  b200rk_options o{};
  RhsCall rhs{};
  b200rk_stats stats{};
  StepCounters cnt;
  // state
  b200rk_vec *Y[2] = {nullptr, nullptr}, *F[2] = {nullptr, nullptr}, *LDY = nullptr, *SCR = nullptr;
  int cur = 0, fcur = 0;
  double t = 0, dt = 0, dtInit = 0, tEnd = 0, error = 0;
  bool adaptive = false, dense = false;
  // dense output
  double sign = 1.0;
  std::vector<double> targets;  // tPositive, or tNegative (already reversed)
  long denseIndex = 0;
  double last_t = 0;
  const b200rk_vec *last_y = nullptr, *last_dy = nullptr;
  std::vector<b200rk_vec*>* emit = nullptr;
  bool finished = false;
  int64_t launches0 = 0, collectives0 = 0;

  ~b200rk_solver() {
    for (auto* v : {Y[0], Y[1], F[0], F[1], LDY, SCR}) if (v) vec_release(v);
  }
};

static int solver_alloc(b200rk_solver* s, size_t N) {
  b200rk_ctx* c = s->c;
  TRY(vec_alloc(c, N, &s->Y[0]));
  TRY(vec_alloc(c, N, &s->Y[1]));
  TRY(vec_alloc(c, N, &s->F[0]));
  if (s->md->use_fsal) TRY(vec_alloc(c, N, &s->F[1]));
  TRY(vec_alloc(c, N, &s->LDY));
  return B200RK_OK;
}

static int vec_copy_raw(b200rk_ctx* c, b200rk_vec* dst, const b200rk_vec* src) {
  CUDA_TRY(c, cudaMemcpyAsync(dst->d, src->d, src->n_local * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  return B200RK_OK;
}

static int hermite_into(b200rk_ctx* c, b200rk_vec* out, double x, double x1, double x2, const b200rk_vec* y1,
                        const b200rk_vec* y2, const b200rk_vec* dy1, const b200rk_vec* dy2) {
  // utils.nim:273-279 — scalars on the host in the reference's order
  const double t = (x - x1) / (x2 - x1);
  const double u = 1.0 - t;
  const double h00 = (1.0 + 2.0 * t) * (u * u);
  const double h10 = t * (u * u);
  const double h01 = (t * t) * (3.0 - 2.0 * t);
  const double h11 = (t * (t * t)) - (t * t);
  return launch_hermite(c, y1->d, dy1->d, y2->d, dy2->d, h00, h10 * (x2 - x1), h01, h11 * (x2 - x1), out->d, y1->n_local);
}

// Begin one time direction from (t0, y0). sign = +1 forward, -1 backward (t := -t, g = -f(-t, .)).
static int solver_begin(b200rk_solver* s, const b200rk_vec* y0, double sign, std::vector<double> targets,
                        bool dense, std::vector<b200rk_vec*>* emit) {
  b200rk_ctx* c = s->c;
  s->sign = sign;
  s->rhs.negate_time = (sign < 0);
  s->targets = std::move(targets);
  s->dense = dense;
  s->emit = emit;
  s->denseIndex = 0;
  s->cur = 0; s->fcur = 0;
  s->finished = false;
  const double t0 = s->o.tStart;
  s->t = (sign < 0) ? -t0 : t0;
  TRY(vec_copy_raw(c, s->Y[0], y0));
  if (sign > 0) {
    // ode.nim:498 and :506 — two separate evaluations at t0
    TRY(eval_rhs(c, s->rhs, s->t, s->Y[0], s->LDY));
    TRY(eval_rhs(c, s->rhs, s->t, s->Y[0], s->F[0]));
  } else {
    // ode.nim:546-548 — FSAL = g(-t0, y0); lastIter.dy = FSAL
    TRY(eval_rhs(c, s->rhs, s->t, s->Y[0], s->F[0]));
    TRY(vec_copy_raw(c, s->LDY, s->F[0]));
  }
  s->last_t = s->t; s->last_y = s->Y[0]; s->last_dy = s->LDY;
  s->dt = s->dtInit;
  if (s->targets.empty()) { s->finished = true; return B200RK_OK; }
  if (sign > 0) s->tEnd = *std::max_element(s->targets.begin(), s->targets.end());   // ode.nim:510
  else s->tEnd = -*std::min_element(s->targets.begin(), s->targets.end());           // ode.nim:549
  return B200RK_OK;
}

static int solver_emit_sample(b200rk_solver* s, double x) {
  b200rk_ctx* c = s->c;
  const size_t N = s->Y[0]->n_global;
  b200rk_vec* out = nullptr;
  TRY(vec_alloc(c, N, &out));
  const b200rk_vec* y = s->Y[s->cur];
  const b200rk_vec* dy2;
  if (s->md->use_fsal) dy2 = s->F[s->fcur];
  else {  // ode.nim:520-521 / 562-563: f(t, y) evaluated for every emitted sample
    if (!s->SCR) TRY(vec_alloc(c, N, &s->SCR));
    TRY(eval_rhs(c, s->rhs, s->t, y, s->SCR));
    dy2 = s->SCR;
  }
  int rc = hermite_into(c, out, x, s->last_t, s->t, s->last_y, y, s->last_dy, dy2);
  if (rc != B200RK_OK) { vec_release(out); return rc; }
  s->emit->push_back(out);
  return B200RK_OK;
}

// The `while t < tEnd` loop (ode.nim:511-541 / 553-583), resumable after max_steps accepted steps.
static int solver_advance(b200rk_solver* s, int64_t max_steps, int64_t* steps_done) {
  b200rk_ctx* c = s->c;
  const MethodDef& md = *s->md;
  int64_t done = 0;
  const long high = (long)s->targets.size() - 1;
  while (!s->finished && s->t < s->tEnd) {
    if (s->dense) {
      if (high < s->denseIndex) break;
      while (s->sign * s->targets[s->denseIndex] <= s->t) {
        TRY(solver_emit_sample(s, s->sign * s->targets[s->denseIndex]));
        s->denseIndex += 1;
        if (high < s->denseIndex) break;
      }
    }
    if (max_steps >= 0 && done >= max_steps) { if (steps_done) *steps_done = done; return B200RK_OK; }
    s->dt = nim_min(s->dt, s->tEnd - s->t);                                       // ode.nim:525
    if (s->dense) {                                                               // ode.nim:526-530
      s->last_t = s->t; s->last_y = s->Y[s->cur];
      if (md.use_fsal) s->last_dy = s->F[s->fcur];
      else { TRY(eval_rhs(c, s->rhs, s->t, s->Y[s->cur], s->LDY)); s->last_dy = s->LDY; }
    }
    double dt_used = 0, err = 0;
    b200rk_vec* fsal_new = md.use_fsal ? s->F[1 - s->fcur] : nullptr;
    TRY(do_step(c, md, s->rhs, s->t, s->Y[s->cur], s->F[s->fcur], s->dt, s->o, s->Y[1 - s->cur], fsal_new,
                &dt_used, &err, &s->cnt));                                        // ode.nim:531
    s->cur = 1 - s->cur;
    if (md.use_fsal) s->fcur = 1 - s->fcur;
    s->dt = dt_used; s->error = err;
    s->t += s->dt;                                                                // ode.nim:532
    s->stats.steps++;
    ++done;
    if (s->adaptive) {                                                            // ode.nim:533-541
      if (s->error == 0.0) s->dt *= 5;
      else s->dt = s->dt * nim_min(4, nim_max(0.125, 0.9 * std::pow(1.0 / s->error, 1.0 / md.order)));
      if (s->dt < s->o.dtMin) s->dt = s->o.dtMin;
      else if (s->o.dtMax < s->dt) s->dt = s->o.dtMax;
    }
  }
  if (!s->finished) {
    s->finished = true;
    if (s->emit) {                                                                // ode.nim:542 / 584
      b200rk_vec* out = nullptr;
      TRY(vec_alloc(c, s->Y[0]->n_global, &out));
      int rc = vec_copy_raw(c, out, s->Y[s->cur]);
      if (rc != B200RK_OK) { vec_release(out); return rc; }
      s->emit->push_back(out);
    }
  }
  if (steps_done) *steps_done = done;
  return B200RK_OK;
}

static void solver_fill_stats(const b200rk_solver* s, b200rk_stats* out, int64_t launches0, int64_t coll0) {
  *out = s->stats;
  out->attempts = s->cnt.attempts; out->rejected = s->cnt.rejected; out->limiter_hits = s->cnt.limiter_hits;
  out->launches = s->c->launches - launches0;
  out->collectives = s->c->collectives - coll0;
}

static int solver_create(b200rk_ctx* c, int method, b200rk_rhs_fn f, void* user, const b200rk_options* options,
                         size_t N, b200rk_solver** out) {
  if (method < 0 || method >= B200RK_METHOD_COUNT) return fail(c, B200RK_EINVAL, "bad method id");
  if (!f) return fail(c, B200RK_EINVAL, "null right-hand side");
  b200rk_solver* s = new b200rk_solver;
  s->c = c; s->md = &method_def(method);
  s->launches0 = c->launches; s->collectives0 = c->collectives;
  if (options) s->o = *options; else b200rk_options_default(&s->o);
  s->rhs = RhsCall{f, user, false, &s->stats.rhs_evals};
  s->adaptive = s->md->adaptive;
  s->dtInit = s->adaptive ? std::sqrt(s->o.dtMax * s->o.dtMin) : s->o.dt;       // ode.nim:491-496
  int rc = solver_alloc(s, N);
  if (rc != B200RK_OK) { delete s; return rc; }
  *out = s;
  return B200RK_OK;
}

// =====================================================================================================
// extern "C"
// =====================================================================================================
extern "C" {

const char* b200rk_last_error(const b200rk_ctx* ctx) { return ctx ? ctx->err.c_str() : g_thread_err.c_str(); }

static int ctx_common_init(b200rk_ctx* c) {
  CUDA_TRY(c, cudaSetDevice(c->device));
  cudaDeviceProp prop;
  CUDA_TRY(c, cudaGetDeviceProperties(&prop, c->device));
  c->sm_count = prop.multiProcessorCount;
  if (prop.l2CacheSize > 0) c->l2_bytes = (size_t)prop.l2CacheSize;
  CUDA_TRY(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CUDA_TRY(c, cudaMalloc(&c->d_ticket, sizeof(unsigned int)));
  CUDA_TRY(c, cudaMemset(c->d_ticket, 0, sizeof(unsigned int)));
  CUDA_TRY(c, cudaMalloc(&c->d_result, sizeof(double)));
  CUDA_TRY(c, cudaMalloc(&c->d_halo, 4 * sizeof(double)));
  CUDA_TRY(c, cudaHostAlloc(&c->h_result, sizeof(double), cudaHostAllocMapped));
  CUDA_TRY(c, cudaHostGetDevicePointer(&c->h_result_dev, c->h_result, 0));
  CUDA_TRY(c, cudaHostAlloc(&c->h_seq, sizeof(unsigned long long), cudaHostAllocMapped));
  *c->h_seq = 0;
  CUDA_TRY(c, cudaHostGetDevicePointer(&c->h_seq_dev, c->h_seq, 0));
  if (const char* e = getenv("B200RK_SPIN_READBACK")) c->spin_readback = atoi(e) != 0;
  TRY(ensure_partials(c, 1));
  if (const char* e = getenv("B200RK_VEC_WIDTH")) c->vec_width = (atoi(e) == 2) ? 2 : 4;
  if (const char* e = getenv("B200RK_CTAS_PER_SM")) c->ctas_per_sm = std::max(0, atoi(e));
  if (const char* e = getenv("B200RK_FINISH_CTAS_PER_SM")) c->finish_ctas_per_sm = std::max(0, atoi(e));
  if (const char* e = getenv("B200RK_STRICT_ZEROS")) c->strict_zeros = atoi(e) != 0;
  if (const char* e = getenv("B200RK_FUSE_POINTWISE")) c->fuse_pointwise = atoi(e) != 0;
  if (const char* e = getenv("B200RK_FUSE_STENCIL")) c->fuse_stencil = atoi(e) != 0;
  if (const char* e = getenv("B200RK_L2_HINTS")) c->l2_hints = atoi(e) < 0 ? -1 : (atoi(e) != 0);
  CUDA_TRY(c, cudaDeviceSynchronize());
  return B200RK_OK;
}

int b200rk_init(b200rk_ctx** out, int device) {
  if (!out) return fail(nullptr, B200RK_EINVAL, "null out");
  b200rk_ctx* c = new b200rk_ctx;
  c->device = device;
  int rc = ctx_common_init(c);
  if (rc != B200RK_OK) { g_thread_err = c->err; delete c; return rc; }
  *out = c;
  return B200RK_OK;
}

int b200rk_nccl_unique_id(void* out128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  TRY(nccl_bind(nullptr));
  ncclUniqueId id;
  NCCL_TRY(nullptr, g_nccl.GetUniqueId(&id));
  std::memcpy(out128, &id, sizeof(id));
  return B200RK_OK;
}

int b200rk_init_distributed(b200rk_ctx** out, int device, int rank, int world, const void* id128) {
  if (!out || world < 1 || rank < 0 || rank >= world) return fail(nullptr, B200RK_EINVAL, "bad rank/world");
  b200rk_ctx* c = new b200rk_ctx;
  c->device = device; c->rank = rank; c->world = world;
  int rc = ctx_common_init(c);
  if (rc == B200RK_OK && world > 1) {
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    rc = nccl_bind(c);
    if (rc == B200RK_OK) {
      ncclResult_t e = g_nccl.CommInitRank(&c->comm, world, id, rank);
      if (e != ncclSuccess) rc = fail(c, B200RK_ENCCL, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(e));
    }
    if (rc == B200RK_OK) rc = setup_p2p(c);
  }
  if (rc != B200RK_OK) { g_thread_err = c->err; delete c; return rc; }
  *out = c;
  return B200RK_OK;
}

void b200rk_destroy(b200rk_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto* v : c->pool) { cudaFree(v->d); delete v; }
  for (auto& r : c->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto e : c->ev_free) cudaEventDestroy(e);
  for (int p = 0; p < kMaxPeers; ++p) if (c->peer_opened[p]) cudaIpcCloseMemHandle(c->peer_mail[p]);
  if (c->d_mail) cudaFree(c->d_mail);
  if (c->comm) g_nccl.CommDestroy(c->comm);
  cudaFree(c->d_partials); cudaFree(c->d_ticket); cudaFree(c->d_result); cudaFree(c->d_halo); cudaFreeHost(c->h_result); cudaFreeHost(c->h_seq);
  cudaStreamDestroy(c->stream);
  delete c;
}

void* b200rk_stream(const b200rk_ctx* c) { return (void*)c->stream; }
int b200rk_synchronize(b200rk_ctx* c) { CUDA_TRY(c, cudaStreamSynchronize(c->stream)); return B200RK_OK; }
int b200rk_rank(const b200rk_ctx* c) { return c->rank; }
int b200rk_world(const b200rk_ctx* c) { return c->world; }

int b200rk_set(b200rk_ctx* c, const char* key, int64_t v) {
  std::string k = key ? key : "";
  if (k == "strict_zeros") c->strict_zeros = v != 0;
  else if (k == "vec_width") { if (v != 2 && v != 4) return fail(c, B200RK_EINVAL, "vec_width must be 2 or 4"); c->vec_width = (int)v; }
  else if (k == "ctas_per_sm") { if (v < 0) return fail(c, B200RK_EINVAL, "ctas_per_sm must be >= 0"); c->ctas_per_sm = (int)v; }
  else if (k == "finish_ctas_per_sm") { if (v < 0) return fail(c, B200RK_EINVAL, "finish_ctas_per_sm must be >= 0"); c->finish_ctas_per_sm = (int)v; }
  else if (k == "profile") c->profile = v != 0;
  else if (k == "fuse_pointwise") c->fuse_pointwise = v != 0;
  else if (k == "spin_readback") c->spin_readback = v != 0;
  else if (k == "fuse_stencil") c->fuse_stencil = v != 0;
  else if (k == "l2_hints") c->l2_hints = v < 0 ? -1 : (v != 0);
  else if (k == "fused_ctas_per_sm") { if (v < 0) return fail(c, B200RK_EINVAL, "fused_ctas_per_sm must be >= 0"); c->fused_ctas_per_sm = (int)v; }
  else if (k == "pool_budget_mb") {
    c->pool_budget_bytes = (size_t)std::max<int64_t>(0, v) << 20;
    if (v == 0) {  // trim now
      CUDA_TRY(c, cudaStreamSynchronize(c->stream));
      for (auto* p : c->pool) { cudaFree(p->d); delete p; }
      c->pool.clear();
    }
  }
  else return fail(c, B200RK_EINVAL, "unknown knob " + k);
  return B200RK_OK;
}
int b200rk_get(const b200rk_ctx* c, const char* key, int64_t* v) {
  std::string k = key ? key : "";
  if (k == "strict_zeros") *v = c->strict_zeros;
  else if (k == "vec_width") *v = c->vec_width;
  else if (k == "ctas_per_sm") *v = c->ctas_per_sm;
  else if (k == "finish_ctas_per_sm") *v = c->finish_ctas_per_sm;
  else if (k == "fuse_pointwise") *v = c->fuse_pointwise;
  else if (k == "spin_readback") *v = c->spin_readback;
  else if (k == "p2p") *v = c->p2p;
  else if (k == "fuse_stencil") *v = c->fuse_stencil;
  else if (k == "l2_hints") *v = c->l2_hints;
  else if (k == "fused_ctas_per_sm") *v = c->fused_ctas_per_sm;
  else if (k == "profile") *v = c->profile;
  else if (k == "sm_count") *v = c->sm_count;
  else if (k == "pool_budget_mb") *v = (int64_t)(c->pool_budget_bytes >> 20);
  else return fail(c, B200RK_EINVAL, "unknown knob " + k);
  return B200RK_OK;
}

int b200rk_profile_reset(b200rk_ctx* c) {
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  for (auto& r : c->prof) { c->ev_free.push_back(r.a); c->ev_free.push_back(r.b); }
  c->prof.clear();
  return B200RK_OK;
}
int b200rk_profile_read(b200rk_ctx* c, b200rk_profile* out) {
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  std::memset(out, 0, sizeof(*out));
  for (auto& r : c->prof) {
    float ms = 0;
    CUDA_TRY(c, cudaEventElapsedTime(&ms, r.a, r.b));
    out->launches[r.cls]++;
    out->ms[r.cls] += ms;
    out->algorithmic_bytes[r.cls] += r.bytes;
  }
  return B200RK_OK;
}
int b200rk_ctx_stats(const b200rk_ctx* c, b200rk_stats* out) {
  std::memset(out, 0, sizeof(*out));
  out->launches = c->launches;
  out->collectives = c->collectives;
  return B200RK_OK;
}

// ---- options / dispatch -----------------------------------------------------------------------------
int b200rk_options_new(b200rk_options* out, double dt, double absTol, double relTol, double dtMax, double dtMin,
                       double scaleMax, double scaleMin, double tStart) {
  if (std::fabs(dtMax) < std::fabs(dtMin)) return fail(nullptr, B200RK_EINVAL, "dtMin must be less than dtMax");   // ode.nim:95-96
  if (std::fabs(scaleMax) < 1) return fail(nullptr, B200RK_EINVAL, "scaleMax must be bigger than 1");              // ode.nim:97-98
  if (1 < std::fabs(scaleMin)) return fail(nullptr, B200RK_EINVAL, "scaleMin must be smaller than 1");             // ode.nim:99-100
  *out = b200rk_options{std::fabs(dt), std::fabs(dtMax), std::fabs(dtMin), tStart, std::fabs(absTol),
                        std::fabs(relTol), std::fabs(scaleMax), std::fabs(scaleMin)};                              // ode.nim:101-102
  return B200RK_OK;
}
void b200rk_options_default(b200rk_options* out) { b200rk_options_new(out, 1e-4, 1e-4, 1e-4, 1e-2, 1e-4, 4.0, 0.1, 0.0); }

int b200rk_method_from_name(const char* name, int* method) {
  std::string s = name ? name : "";
  std::string low = s;
  for (char& ch : low) ch = (char)std::tolower((unsigned char)ch);
  for (int i = 0; i < B200RK_METHOD_COUNT; ++i)
    if (low == method_def(i).name) { *method = i; return B200RK_OK; }
  return fail(nullptr, B200RK_EINVAL, s + " is not a valid integrator");  // ode.nim:651
}
const char* b200rk_method_name(int method) {
  return (method >= 0 && method < B200RK_METHOD_COUNT) ? method_def(method).name : nullptr;
}
int b200rk_method_info(int method, int* stages, int* use_fsal, double* order, int* adaptive) {
  if (method < 0 || method >= B200RK_METHOD_COUNT) return fail(nullptr, B200RK_EINVAL, "bad method id");
  const MethodDef& m = method_def(method);
  if (stages) *stages = m.stages;
  if (use_fsal) *use_fsal = m.use_fsal;
  if (order) *order = m.order;
  if (adaptive) *adaptive = m.adaptive;
  return B200RK_OK;
}
int b200rk_method_tableau(int method, double* c, double* a, double* b, double* bhat) {
  if (method < 0 || method > B200RK_VERN65) return fail(nullptr, B200RK_EINVAL, "tableau export covers the FSAL pairs only");
  const MethodDef& m = method_def(method);
  std::memset(c, 0, 10 * sizeof(double)); std::memset(a, 0, 90 * sizeof(double));
  std::memset(b, 0, 9 * sizeof(double)); std::memset(bhat, 0, 9 * sizeof(double));
  for (int s = 2; s <= m.stages; ++s) {
    c[s] = m.c[s];
    for (int j = 0; j < m.a[s].m; ++j) a[s * 9 + (m.a[s].idx[j] - 1)] = m.a[s].w[j];
  }
  for (int j = 0; j < m.b.m; ++j) b[m.b.idx[j] - 1] = m.b.w[j];
  for (int j = 0; j < m.bhat.m; ++j) bhat[m.bhat.idx[j] - 1] = m.bhat.w[j];
  return B200RK_OK;
}

int b200rk_shard_range(size_t n_global, int rank, int world, size_t* offset, size_t* len) {
  if (world < 1 || rank < 0 || rank >= world || !offset || !len) return fail(nullptr, B200RK_EINVAL, "bad rank/world");
  shard_range(n_global, rank, world, offset, len);
  return B200RK_OK;
}

// ---- vectors ----------------------------------------------------------------------------------------
int b200rk_vec_new(b200rk_ctx* c, size_t n_global, b200rk_vec** out) {
  if (!c || !out) return fail(c, B200RK_EINVAL, "null argument");
  return vec_alloc(c, n_global, out);
}
int b200rk_vec_free(b200rk_vec* v) {
  // Freed vectors go back to the context's pool (stream-ordered reuse is safe: one stream per context);
  // beyond the pool budget they are released to the driver.
  if (!v) return B200RK_OK;
  b200rk_ctx* c = v->ctx;
  size_t held = 0;
  for (auto* p : c->pool) held += p->n_local * sizeof(double);
  if (c->pool.size() < 512 && held + v->n_local * sizeof(double) <= c->pool_budget_bytes) {
    c->pool.push_back(v);
    return B200RK_OK;
  }
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  CUDA_TRY(c, cudaFree(v->d));
  delete v;
  return B200RK_OK;
}
size_t b200rk_vec_len(const b200rk_vec* v) { return v->n_global; }
size_t b200rk_vec_local_len(const b200rk_vec* v) { return v->n_local; }
size_t b200rk_vec_local_offset(const b200rk_vec* v) { return v->offset; }
double* b200rk_vec_data(const b200rk_vec* v) { return v->d; }

int b200rk_vec_upload_local(b200rk_vec* v, const double* h) {
  b200rk_ctx* c = v->ctx;
  CUDA_TRY(c, cudaMemcpyAsync(v->d, h, v->n_local * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return B200RK_OK;
}
int b200rk_vec_download_local(const b200rk_vec* v, double* h) {
  b200rk_ctx* c = v->ctx;
  CUDA_TRY(c, cudaMemcpyAsync(h, v->d, v->n_local * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return B200RK_OK;
}
int b200rk_vec_upload(b200rk_vec* v, const double* hg) { return b200rk_vec_upload_local(v, hg + v->offset); }
int b200rk_vec_download(const b200rk_vec* v, double* hg) { return b200rk_vec_download_local(v, hg + v->offset); }
int b200rk_vec_copy(b200rk_vec* dst, const b200rk_vec* src) {
  TRY(check_same(dst ? dst->ctx : nullptr, dst, src));
  return vec_copy_raw(dst->ctx, dst, src);
}
int b200rk_vec_fill(b200rk_vec* v, double value) {
  return launch_ewise<EW_FILL>(v->ctx, v->d, nullptr, value, v->d, v->n_local, B200RK_K_OTHER);
}

#define EW_BINARY(NAME, OP)                                                              \
  int NAME(b200rk_vec* out, const b200rk_vec* a, const b200rk_vec* b) {                  \
    TRY(check_same(a ? a->ctx : nullptr, a, b));                                         \
    TRY(check_same(a->ctx, a, out));                                                     \
    return launch_ewise<OP>(a->ctx, a->d, b->d, 0.0, out->d, a->n_local, B200RK_K_OTHER); \
  }
EW_BINARY(b200rk_vec_add, EW_ADD)
EW_BINARY(b200rk_vec_sub, EW_SUB)
EW_BINARY(b200rk_vec_hmul, EW_HMUL)
EW_BINARY(b200rk_vec_hdiv, EW_HDIV)
#undef EW_BINARY
int b200rk_vec_scale(b200rk_vec* out, double s, const b200rk_vec* a) {
  TRY(check_same(a ? a->ctx : nullptr, a, out));
  return launch_ewise<EW_SCALE>(a->ctx, a->d, nullptr, s, out->d, a->n_local, B200RK_K_OTHER);
}
int b200rk_vec_div_scalar(b200rk_vec* out, const b200rk_vec* a, double s) {
  TRY(check_same(a ? a->ctx : nullptr, a, out));
  return launch_ewise<EW_DIV_SCALAR>(a->ctx, a->d, nullptr, s, out->d, a->n_local, B200RK_K_OTHER);
}
int b200rk_vec_add_scalar(b200rk_vec* out, double s, const b200rk_vec* a) {
  TRY(check_same(a ? a->ctx : nullptr, a, out));
  return launch_ewise<EW_ADD_SCALAR>(a->ctx, a->d, nullptr, s, out->d, a->n_local, B200RK_K_OTHER);
}
int b200rk_vec_neg(b200rk_vec* out, const b200rk_vec* a) {
  TRY(check_same(a ? a->ctx : nullptr, a, out));
  return launch_ewise<EW_NEG>(a->ctx, a->d, nullptr, 0.0, out->d, a->n_local, B200RK_K_OTHER);
}
int b200rk_vec_abs(b200rk_vec* out, const b200rk_vec* a) {
  TRY(check_same(a ? a->ctx : nullptr, a, out));
  return launch_ewise<EW_ABS>(a->ctx, a->d, nullptr, 0.0, out->d, a->n_local, B200RK_K_OTHER);
}
int b200rk_vec_sum(const b200rk_vec* a, double* out) {
  b200rk_ctx* c = a->ctx;
  {
    ProfScope ps(c, B200RK_K_OTHER, 8.0 * double(a->n_local));
    unsigned grid = grid_for(c, a->n_local / 2, kThreads * 4);
    grid = std::min(grid, (unsigned)(c->sm_count * 8));
    TRY(ensure_partials(c, grid));
    sum_kernel<2, kThreads><<<grid, kThreads, 0, c->stream>>>(a->d, a->n_local, reduce_scratch(c));
    CUDA_TRY(c, cudaGetLastError());
  }
  return fetch_global_sum(c, out);
}
int b200rk_hermite(b200rk_vec* out, double x, double x1, double x2, const b200rk_vec* y1, const b200rk_vec* y2,
                   const b200rk_vec* dy1, const b200rk_vec* dy2) {
  b200rk_ctx* c = y1 ? y1->ctx : nullptr;
  TRY(check_same(c, y1, y2)); TRY(check_same(c, y1, dy1)); TRY(check_same(c, y1, dy2)); TRY(check_same(c, y1, out));
  return hermite_into(c, out, x, x1, x2, y1, y2, dy1, dy2);
}

// ---- built-in right-hand sides ----------------------------------------------------------------------
int b200rk_builtin_rhs_new(b200rk_ctx* c, int kind, double scalar, const b200rk_vec* lambda, b200rk_rhs_fn* fn, void** user) {
  if (kind < B200RK_RHS_SCALE || kind > B200RK_RHS_LORENZ96) return fail(c, B200RK_EINVAL, "unknown builtin rhs");
  if (kind == B200RK_RHS_DIAG_LINEAR && !lambda) return fail(c, B200RK_EINVAL, "diag-linear rhs needs lambda");
  *user = new BuiltinRhs{c, kind, scalar, lambda};
  *fn = &builtin_rhs_fn;
  return B200RK_OK;
}
int b200rk_builtin_rhs_free(void* user) { delete static_cast<BuiltinRhs*>(user); return B200RK_OK; }

// ---- hot path ---------------------------------------------------------------------------------------
int b200rk_step(b200rk_ctx* c, int method, b200rk_rhs_fn f, void* user, double t, const b200rk_vec* y,
                const b200rk_vec* fsal, double dt, const b200rk_options* options, b200rk_vec* y_new,
                b200rk_vec* fsal_new, double* dt_used, double* error) {
  if (!c || !y || !y_new || !f) return fail(c, B200RK_EINVAL, "null argument");
  if (method < 0 || method >= B200RK_METHOD_COUNT) return fail(c, B200RK_EINVAL, "bad method id");
  TRY(check_same(c, y, y_new));
  if (fsal_new) TRY(check_same(c, y, fsal_new));
  if (y_new == y || (fsal_new && (fsal_new == fsal || fsal_new == y)))
    return fail(c, B200RK_EINVAL, "step outputs must not alias inputs");
  b200rk_options o;
  if (options) o = *options; else b200rk_options_default(&o);
  const MethodDef& md = method_def(method);
  if (md.fsal_out != 0 && !fsal_new) return fail(c, B200RK_EINVAL, std::string(md.name) + ": fsal_new required");
  RhsCall rhs{f, user, false, nullptr};
  double du = dt, er = 0.0;
  int rc = do_step(c, md, rhs, t, y, fsal, dt, o, y_new, fsal_new, &du, &er, nullptr);
  if (dt_used) *dt_used = du;
  if (error) *error = er;
  if (rc == B200RK_OK && !md.adaptive) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return rc;
}

int b200rk_solve(b200rk_ctx* c, int method, b200rk_rhs_fn f, void* user, const b200rk_vec* y0, const double* tspan,
                 size_t n_tspan, const b200rk_options* options, double* t_out, b200rk_vec** y_out, size_t* n_y_out,
                 b200rk_stats* stats) {
  if (!c || !y0 || !tspan || !t_out || !y_out || !n_y_out) return fail(c, B200RK_EINVAL, "null argument");
  const int64_t l0 = c->launches, c0 = c->collectives;
  b200rk_solver* s = nullptr;
  TRY(solver_create(c, method, f, user, options, y0->n_global, &s));
  std::vector<double> ts(tspan, tspan + n_tspan);
  std::sort(ts.begin(), ts.end());                                                // ode.nim:609
  const double t0 = s->o.tStart;
  std::vector<double> tPos, tNeg;
  for (double x : ts) if (x > t0) tPos.push_back(x);                              // ode.nim:479
  for (double x : ts) if (x < t0) tNeg.push_back(x);                              // ode.nim:480
  std::reverse(tNeg.begin(), tNeg.end());
  const bool has_zero = std::find(ts.begin(), ts.end(), t0) != ts.end();          // ode.nim:485
  const bool dense = (n_tspan != 2);                                              // ode.nim:499-502
  std::vector<b200rk_vec*> yPos, yNeg, yZero;
  int rc = B200RK_OK;
  auto cleanup = [&]() {
    for (auto* v : yPos) vec_release(v);
    for (auto* v : yNeg) vec_release(v);
    for (auto* v : yZero) vec_release(v);
    delete s;
  };
  if (has_zero) {                                                                 // ode.nim:486-487
    b200rk_vec* z = nullptr;
    rc = vec_alloc(c, y0->n_global, &z);
    if (rc == B200RK_OK) { yZero.push_back(z); rc = vec_copy_raw(c, z, y0); }
  }
  // forward (the two RHS evaluations at t0 happen even when tPositive is empty, ode.nim:498,506)
  if (rc == B200RK_OK) rc = solver_begin(s, y0, +1.0, tPos, dense, &yPos);
  if (rc == B200RK_OK && !tPos.empty()) rc = solver_advance(s, -1, nullptr);
  if (rc == B200RK_OK && !tNeg.empty()) {                                         // ode.nim:544-584
    rc = solver_begin(s, y0, -1.0, tNeg, dense, &yNeg);
    if (rc == B200RK_OK) rc = solver_advance(s, -1, nullptr);
  }
  if (rc == B200RK_OK) { cudaError_t e = cudaStreamSynchronize(c->stream); if (e != cudaSuccess) rc = fail(c, B200RK_ECUDA, cudaGetErrorString(e)); }
  if (rc != B200RK_OK) { cleanup(); return rc; }
  size_t it = 0, iy = 0;                                                          // ode.nim:585-586
  for (auto r = tNeg.rbegin(); r != tNeg.rend(); ++r) t_out[it++] = *r;
  if (has_zero) t_out[it++] = t0;
  for (double x : tPos) t_out[it++] = x;
  for (auto r = yNeg.rbegin(); r != yNeg.rend(); ++r) y_out[iy++] = *r;
  for (auto* v : yZero) y_out[iy++] = v;
  for (auto* v : yPos) y_out[iy++] = v;
  *n_y_out = iy;
  if (stats) solver_fill_stats(s, stats, l0, c0);
  delete s;
  return B200RK_OK;
}

int b200rk_solve_host(b200rk_ctx* c, int method, b200rk_rhs_fn f, void* user, size_t n_global, const double* y0_local,
                      const double* tspan, size_t n_tspan, const b200rk_options* options, double* t_out,
                      double* y_out_local, size_t* n_y_out, b200rk_stats* stats) {
  if (!c || !y0_local || !y_out_local) return fail(c, B200RK_EINVAL, "null argument");
  b200rk_vec* y0 = nullptr;
  TRY(vec_alloc(c, n_global, &y0));
  int rc = B200RK_OK;
  cudaError_t e = cudaMemcpyAsync(y0->d, y0_local, y0->n_local * sizeof(double), cudaMemcpyHostToDevice, c->stream);
  if (e != cudaSuccess) rc = fail(c, B200RK_ECUDA, cudaGetErrorString(e));
  std::vector<b200rk_vec*> ys(n_tspan, nullptr);
  size_t ny = 0;
  if (rc == B200RK_OK) rc = b200rk_solve(c, method, f, user, y0, tspan, n_tspan, options, t_out, ys.data(), &ny, stats);
  if (rc == B200RK_OK) {
    for (size_t i = 0; i < ny; ++i) {
      e = cudaMemcpyAsync(y_out_local + i * y0->n_local, ys[i]->d, y0->n_local * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
      if (e != cudaSuccess) { rc = fail(c, B200RK_ECUDA, cudaGetErrorString(e)); break; }
    }
    e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess && rc == B200RK_OK) rc = fail(c, B200RK_ECUDA, cudaGetErrorString(e));
    for (size_t i = 0; i < ny; ++i) vec_release(ys[i]);
    if (n_y_out) *n_y_out = ny;
  }
  vec_release(y0);
  return rc;
}

int b200rk_solver_new(b200rk_ctx* c, int method, b200rk_rhs_fn f, void* user, const b200rk_vec* y0, double t_end,
                      const b200rk_options* options, b200rk_solver** out) {
  if (!c || !y0 || !out) return fail(c, B200RK_EINVAL, "null argument");
  b200rk_solver* s = nullptr;
  TRY(solver_create(c, method, f, user, options, y0->n_global, &s));
  if (!(t_end > s->o.tStart)) { delete s; return fail(c, B200RK_EINVAL, "solver_new: t_end must be > tStart"); }
  int rc = solver_begin(s, y0, +1.0, std::vector<double>{t_end}, false, nullptr);
  if (rc != B200RK_OK) { delete s; return rc; }
  *out = s;
  return B200RK_OK;
}
int b200rk_solver_advance(b200rk_solver* s, int64_t max_steps, int64_t* steps_done, int* finished) {
  int rc = solver_advance(s, max_steps, steps_done);
  if (rc == B200RK_OK) CUDA_TRY(s->c, cudaStreamSynchronize(s->c->stream));
  if (finished) *finished = s->finished ? 1 : 0;
  return rc;
}
int b200rk_solver_state(const b200rk_solver* s, double* t, double* dt_next, double* last_error, const b200rk_vec** y) {
  if (t) *t = s->t;
  if (dt_next) *dt_next = s->dt;
  if (last_error) *last_error = s->error;
  if (y) *y = s->Y[s->cur];
  return B200RK_OK;
}
int b200rk_solver_stats(const b200rk_solver* s, b200rk_stats* out) {
  solver_fill_stats(s, out, s->launches0, s->collectives0);
  return B200RK_OK;
}
int b200rk_solver_free(b200rk_solver* s) {
  if (s) { cudaStreamSynchronize(s->c->stream); delete s; }
  return B200RK_OK;
}

// ---- raw kernels ------------------------------------------------------------------------------------
int b200rk_stage_accum(b200rk_ctx* c, int m, const double* w, double cc, int chain, const b200rk_vec* y,
                       const b200rk_vec* const* k, b200rk_vec* out) {
  if (!c || !w || !y || !k || !out) return fail(c, B200RK_EINVAL, "null argument");
  if (m < 1 || m > kMaxTerms) return fail(c, B200RK_EINVAL, "stage_accum: m must be in 1..9");
  TRY(check_same(c, y, out));
  const double* kp[kMaxTerms];
  for (int j = 0; j < m; ++j) { TRY(check_same(c, y, k[j])); kp[j] = k[j]->d; }
  return launch_stage(c, m, y->d, kp, w, cc, chain != 0, out->d, y->n_local);
}

int b200rk_combine_err(b200rk_ctx* c, int method, double dt, double absTol, double relTol, const b200rk_vec* y,
                       const b200rk_vec* const* k, b200rk_vec* y_new, b200rk_vec* err_y, double* sumsq, double* error) {
  if (!c || !y || !k || !y_new) return fail(c, B200RK_EINVAL, "null argument");
  if (method < 0 || method >= B200RK_METHOD_COUNT || !method_def(method).adaptive)
    return fail(c, B200RK_EINVAL, "combine_err: adaptive method required");
  const MethodDef& md = method_def(method);
  TRY(check_same(c, y, y_new));
  b200rk_vec* kk[kMaxStages + 1] = {nullptr};
  for (int s = 1; s <= md.stages; ++s) { TRY(check_same(c, y, k[s - 1])); kk[s] = const_cast<b200rk_vec*>(k[s - 1]); }
  bool ynew_ready = false;
  if (md.err_direct) {  // yNew first (stage kernel on the b row), then the direct error row
    TRY(run_row(c, md.b, md.b_cfac, false, dt, y, kk, y_new));
    ynew_ready = true;
  }
  FinishPlan p;
  TRY(plan_finish(c, md, dt, absTol, relTol, y, kk, y_new, ynew_ready, err_y ? err_y->d : nullptr, &p));
  TRY(launch_finish(c, p));
  double S2 = 0.0;
  TRY(fetch_global_sum(c, &S2));
  if (sumsq) *sumsq = S2;
  if (error) *error = std::sqrt(1.0 / double(y->n_global) * S2);
  return B200RK_OK;
}

int b200rk_rk4_combine(b200rk_ctx* c, double dt, const b200rk_vec* y, const b200rk_vec* k1, const b200rk_vec* k2,
                       const b200rk_vec* k3, const b200rk_vec* k4, b200rk_vec* out) {
  TRY(check_same(c, y, k1)); TRY(check_same(c, y, k2)); TRY(check_same(c, y, k3)); TRY(check_same(c, y, k4)); TRY(check_same(c, y, out));
  const size_t n = y->n_local;
  if (!n) return B200RK_OK;
  ProfScope ps(c, B200RK_K_STAGE, 8.0 * double(n) * 6);
  if (c->vec_width == 4) {
    unsigned grid = grid_for(c, n / 4, kThreads);
    rk4_final_kernel<4, 1, kThreads><<<grid, kThreads, 0, c->stream>>>(y->d, k1->d, k2->d, k3->d, k4->d, dt / 6.0, out->d, n);
  } else {
    unsigned grid = grid_for(c, n / 2, kThreads);
    rk4_final_kernel<2, 1, kThreads><<<grid, kThreads, 0, c->stream>>>(y->d, k1->d, k2->d, k3->d, k4->d, dt / 6.0, out->d, n);
  }
  CUDA_TRY(c, cudaGetLastError());
  return B200RK_OK;
}

}  // extern "C"
