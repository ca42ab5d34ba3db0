// internal.hpp — declarations shared by the translation units of libb200rk.so (not installed).
//   runtime.cu   context, NCCL binding, peer mailboxes, vector pool, knobs, profiler
//   launch.cu    generic kernel launchers (stage / finish / element-wise / Hermite / sum / RK4) + read-back
//   executor.cu  built-in right-hand sides, stage rows, fused paths, one IntegratorProc call (do_step)
//   driver.cu    the resumable ODESolver loop and the step / solve entry points
//   capi.cu      the remaining extern "C" surface (options, dispatch, Vector operators, raw kernels)
//   jit.cu       element-local right-hand sides compiled at run time (NVRTC) into the fused kernels
//   quadrature.cu  hermiteInterpolate + cumulative quadrature over a trajectory (consumers of the solver's output)
#pragma once
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
struct ProfRec {
  int cls;
  double bytes;
  cudaEvent_t a, b;
};

struct PeerVecView;

struct b200rk_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // solve_host: returns the tStart state while the solve runs (created on first use)
  cudaEvent_t copy_event = nullptr;
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
  const PeerVecView* peer_view = nullptr;   // set by the solver that is advancing (driver.cu); null otherwise
  int* d_barrier = nullptr;                 // 1 int: payload of stream_barrier
  double* d_halo_stencil = nullptr;  // 2 x 8 doubles: halo of a sharded stencil right-hand side given as source (jit.cu)
  double* d_halo_attempt = nullptr;  // 2 x 24 doubles: halo of y and k1 for the one-kernel Lorenz-96 attempt (stencil_attempt.cuh), sharded
  unsigned long long* h_seq = nullptr;      // pinned + mapped: sequence word of the last finished reduction
  unsigned long long* h_seq_dev = nullptr;  // device alias
  unsigned long long seq = 0;               // last sequence number handed to a reducing launch
  bool spin_readback = true;                // poll h_seq instead of cudaStreamSynchronize (single GPU)
  int device_loop = -1;                     // persistent cooperative driver loop: -1 auto (= on, single GPU), 0 off, 1 on
  RunState* d_run_state = nullptr;          // device copy of the loop state
  RunState* h_run_state = nullptr;          // pinned + mapped mirror
  RunState* h_run_state_dev = nullptr;
  // in-kernel all-reduce of the error norm over peer mailboxes (multi-GPU)
  unsigned long long* d_mail = nullptr;     // this rank's mailbox
  unsigned long long* peer_mail[kMaxPeers] = {nullptr};
  bool peer_opened[kMaxPeers] = {false};
  bool p2p = false;
  std::string p2p_note;
  // How long a kernel waits for a peer's partial sum before it reports B200RK_ENCCL instead of hanging. Generous on
  // purpose: ranks can be seconds apart on the host (first-use NVRTC compile, a slow user callback, a large allocation),
  // and NCCL's own collectives have no such limit. Knob "peer_timeout_s" / env B200RK_PEER_TIMEOUT_S.
  int peer_timeout_s = 120;
  int clock_khz = 1965000;
  long long peer_timeout_cycles = 120ll * 1965000ll * 1000ll;
  // workspace pool (free vectors by global length)
  std::vector<b200rk_vec*> pool;
  size_t pool_budget_bytes = (size_t)48 << 30;
  // knobs
  int vec_width = 4;
  int ctas_per_sm = 0;         // stage/element-wise kernels: 0 = one tile per CTA (measured best, profiles/)
  int finish_ctas_per_sm = 2;  // reducing kernels: persistent grid, one partial per CTA (measured best)
  int fused_ctas_per_sm = 4;   // fused pointwise attempt kernel: persistent grid of 4 CTAs per SM (~100 registers: 2 resident, 2 waves; measured best, profiles/r01_tune_*)
  bool fuse_pointwise = true;  // element-local built-in RHS: whole attempt in one kernel
  bool fuse_stencil = true;    // built-in Lorenz-96 (single GPU): stage accumulate + stencil RHS in one kernel
  bool fuse_stencil_attempt = true;  // built-in Lorenz-96: a whole attempt in one kernel over overlapped tiles (default since round 2: 914 -> 3939 steps/s on config 3, profiles/r02_*)
  bool l96_peer_halo = true;   // sharded one-kernel Lorenz-96 attempt inside a solver: read the halo in place from the peer-mapped neighbours (false: ncclSend/ncclRecv)
  int l96_warp_tiles = 0;      // one-kernel Lorenz-96 attempt over warp-sized tiles (shuffles, no block barrier) instead of CTA-sized tiles: 0 off, 8 (or 1) = 8 elements per lane, 4 = 4 per lane
  int l96_ctas_per_sm = 0;     // l96_attempt_kernel's persistent grid: CTAs per SM, 0 = what the occupancy calculator allows
  int l96_attempt_threads = 128;   // l96_attempt_kernel: threads per CTA: 128 = 512-position tiles, 6 CTAs per SM, barrier domains of 4 warps (measured 150 vs 155 us for 256)
  int l96_attempt_pairs = 2;   // l96_attempt_kernel: 128-bit pairs per thread (tile = 512 * pairs positions); 1 or 2
  bool finish_prefetch = false; // register-prefetching finish kernel: measured on the B200 in round 2 — no gain (88.5 vs 90.1-91.8 us), stays off
  int stream_simpson = -1;     // cumsimpson(f, X, dx): -1 = stream the grid only when the composed form would not fit, 0 never, 1 always
  bool fuse_simpson = true;    // cumsimpson as one kernel (default since round 2: 1.166 -> 0.674 ms at 2^23 x 33 points, profiles/r02_quadrature_*)
  int tstart_copy = 0;         // solve_host's tStart output (= y0): 0 D2H on a second stream while the solve runs (measured best at 1..8 ranks), 1 host-side copy by helper threads
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

struct PeerVecView {
  static constexpr int kMax = 4;
  int count = 0;                       // 0 = not available
  const double* local[kMax] = {};      // this rank's vectors (device pointers)
  const double* left[kMax] = {};       // the left / right ring neighbour's vectors of the same role (peer-mapped base pointers)
  const double* right[kMax] = {};
  size_t n_left = 0, n_right = 0;      // the neighbours' shard lengths
  void* opened[2 * kMax] = {};         // what cudaIpcCloseMemHandle must release
  int n_opened = 0;
  int find(const double* d) const { for (int i = 0; i < count; ++i) if (local[i] == d) return i; return -1; }
};

struct b200rk_vec {
  b200rk_ctx* ctx;
  size_t n_global, offset, n_local;
  double* d;
};

int fail(const b200rk_ctx* ctx, int code, const std::string& msg);   // records the message, returns `code`
std::string& thread_error();                                          // last error of this thread (ctx == NULL)

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

inline double nim_min(double x, double y) { return (x <= y) ? x : y; }  // Nim system.min
inline double nim_max(double x, double y) { return (y <= x) ? x : y; }  // Nim system.max

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
};extern NcclApi g_nccl;
int nccl_bind(const b200rk_ctx* ctx);

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
// launch geometry (shared by every translation unit that launches kernels)
// =====================================================================================================
static constexpr int kThreads = 256;

template <int M, int W>
struct StageUnroll {  // keep ~16 doubles of loads in flight per thread
  static constexpr int raw = 16 / ((M + 1) * W);
  static constexpr int value = raw < 1 ? 1 : (raw > 4 ? 4 : raw);
};

// L2 hand-off between producer and consumer kernels (kernels.cuh: L2Policy): worthwhile while one vector
// fits comfortably in the L2 (measured: +13 % at 2^22, +16 % at 2^23, -2 % at 2^24 = 128 MiB per vector).
inline bool l2_on(const b200rk_ctx* c, size_t n_local) {
  if (c->l2_hints >= 0) return c->l2_hints != 0;
  return n_local * sizeof(double) <= (size_t)(0.65 * (double)c->l2_bytes);
}

inline unsigned grid_for(const b200rk_ctx* c, size_t nvec, int per_block, int ctas_per_sm = -1) {
  size_t tiles = (nvec + per_block - 1) / per_block;
  if (tiles == 0) tiles = 1;
  if (ctas_per_sm < 0) ctas_per_sm = c->ctas_per_sm;
  if (ctas_per_sm > 0) tiles = std::min(tiles, (size_t)ctas_per_sm * c->sm_count);
  return (unsigned)std::min(tiles, (size_t)0x7fffffff);
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

struct BuiltinRhs {
  b200rk_ctx* ctx;
  int kind;
  double scalar;
  const b200rk_vec* lambda;
};

struct RhsCall {
  b200rk_rhs_fn f;
  void* user;
  bool negate_time;  // backward pass: g(t, y) = -f(-t, y)   (ode.nim:545)
  int64_t* evals;
};

// ---- runtime.cu -------------------------------------------------------------------------------------
int ensure_partials(b200rk_ctx* c, size_t blocks);
int ensure_local_mailbox(b200rk_ctx* c);
ReduceScratch reduce_scratch(b200rk_ctx* c);                 // hands out the next reduction sequence number
int setup_p2p(b200rk_ctx* c);
// The same-role vectors of the two ring neighbours, peer-mapped (CUDA IPC over NVLink): the one-kernel Lorenz-96 attempt
// reads its halo from them in place. Opened per solver (driver.cu), used by do_step through ctx->peer_view while that
// solver advances. Both calls are collective over the communicator.
int peer_view_open(b200rk_ctx* c, b200rk_vec* const* vecs, int count, PeerVecView* out);   // out->count == 0: not available (agreed by all ranks)
int peer_view_close(b200rk_ctx* c, PeerVecView* v);
int stream_barrier(b200rk_ctx* c);   // every rank's stream has reached this point before any rank's stream continues
void shard_range(size_t n, int rank, int world, size_t* off, size_t* len);
int vec_alloc(b200rk_ctx* c, size_t n_global, b200rk_vec** out);
void vec_release(b200rk_vec* v);
void pool_trim(b200rk_ctx* c, size_t keep_bytes);            // free pooled vectors down to keep_bytes
bool pool_put(b200rk_vec* v);                                 // false: the pool is full (count or byte budget)
int check_same(const b200rk_ctx* c, const b200rk_vec* a, const b200rk_vec* b);
int vec_copy_raw(b200rk_ctx* c, b200rk_vec* dst, const b200rk_vec* src);

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

// ---- launch.cu --------------------------------------------------------------------------------------
int launch_stage(b200rk_ctx* c, int m, const double* y, const double* const* k, const double* w, double cc, bool chain,
                 double* out, size_t n);
int launch_finish(b200rk_ctx* c, const FinishPlan& p);
int fetch_global_sum(b200rk_ctx* c, double* out);
int launch_ewise(b200rk_ctx* c, int op, const double* a, const double* b, double s, double* out, size_t n, int cls);
int launch_hermite(b200rk_ctx* c, const double* y1, const double* dy1, const double* y2, const double* dy2, double h00,
                   double hA, double h01, double hB, double* out, size_t n);
int launch_rk4_final(b200rk_ctx* c, const double* y, const double* k1, const double* k2, const double* k3, const double* k4,
                     double c6, double* out, size_t n);
int launch_lorenz96(b200rk_ctx* c, const double* y, const double* left2, const double* right1, double F, double* out, size_t n);
int launch_sum(b200rk_ctx* c, const double* a, size_t n);

// ---- executor.cu ------------------------------------------------------------------------------------
int builtin_rhs_fn(double t, const b200rk_vec* y, b200rk_vec* dydt, void* user);
int eval_rhs(b200rk_ctx* c, const RhsCall& r, double t, const b200rk_vec* y, b200rk_vec* out);
int run_row(b200rk_ctx* c, const Row& row, double cfac, bool chain, double dt, const b200rk_vec* y, b200rk_vec* const* k,
            b200rk_vec* out);
int plan_finish(const b200rk_ctx* c, const MethodDef& md, double dt, double absTol, double relTol, const b200rk_vec* y,
                b200rk_vec* const* k, b200rk_vec* ynew, bool ynew_ready, double* err_out, FinishPlan* p);
int do_step(b200rk_ctx* c, const MethodDef& md, const RhsCall& rhs, double t, const b200rk_vec* y, const b200rk_vec* fsal,
            double dt_in, const b200rk_options& o, b200rk_vec* y_new, b200rk_vec* fsal_new, double* dt_used,
            double* error_out, StepCounters* cnt);
// Device-resident driver loop (kernels.cuh: fused_run_kernel): element-local built-in RHS, FSAL pair, one GPU.
struct DeviceLoopIO {
  b200rk_vec* Y[2];
  b200rk_vec* F[2];
  int cur;
  double t, dt, t_end, error;
  int64_t steps, attempts, rejected, limiter_hits;
};
bool l96_peer_halo_possible(const b200rk_ctx* c, const MethodDef& md, const RhsCall& rhs, size_t n_global);
bool device_loop_eligible(const b200rk_ctx* c, const MethodDef& md, const RhsCall& rhs, size_t n_local);
int run_device_loop(b200rk_ctx* c, const MethodDef& md, const RhsCall& rhs, const b200rk_options& o, DeviceLoopIO* io,
                    int64_t max_steps);
int hermite_into(b200rk_ctx* c, b200rk_vec* out, double x, double x1, double x2, const b200rk_vec* y1, const b200rk_vec* y2,
                 const b200rk_vec* dy1, const b200rk_vec* dy2);
// scalar factors of hermiteSpline (utils.nim:273-279): f = {h00, h10*(x2-x1), h01, h11*(x2-x1)}
void hermite_factors(double x, double x1, double x2, double* f);

// ---- jit.cu -----------------------------------------------------------------------------------------
// A right-hand side given as source (b200rk_jit_rhs_new). `pattern` is a FusedPattern (kernels.cuh) or -1 for
// the plain dydt / RK4 unit; `slot` picks the kernel inside that unit; `arg_block` is the kernel's single
// by-value argument (FusedArgs<S> / RunArgs<S> / UserRhsArgs), laid out by the same header on both sides.
struct JitRhs;
int jit_rhs_fn(double t, const b200rk_vec* y, b200rk_vec* dydt, void* user);
void jit_describe(const JitRhs* j, int* np, const b200rk_vec* const** vecs, const double** cs);
int jit_slot_attempt(int w);
int jit_slot_run(int w);
int jit_launch(b200rk_ctx* c, JitRhs* j, int pattern, int slot, unsigned grid, void* arg_block, bool cooperative);
bool jit_is_stencil(const JitRhs* j, int* radius_left, int* radius_right);
namespace b200rk { struct L96Halo; }   // stencil_attempt.cuh
int jit_launch_stencil_rk4(b200rk_ctx* c, JitRhs* j, bool negate, double t, double dt, const b200rk::L96Halo& halo, const b200rk_vec* y, b200rk_vec* y_new);   // stencil right-hand side from source (not element-local)
int jit_prepare(b200rk_ctx* c, JitRhs* j, int pattern);      // compile + load the base unit and (pattern >= 0) the fused unit now
int fused_pattern_for(const b200rk_ctx* c, const MethodDef& md);   // sparsity pattern of the method's fused kernels, -1: none
int jit_max_blocks_per_sm(b200rk_ctx* c, JitRhs* j, int pattern, int slot, int* per_sm);
int jit_launch_rk4(b200rk_ctx* c, JitRhs* j, bool negate, double t, double dt, const b200rk_vec* y, b200rk_vec* y_new);
