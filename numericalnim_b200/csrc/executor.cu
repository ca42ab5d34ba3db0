// executor.cu — one IntegratorProc call (ode.nim:38) on device vectors: built-in right-hand sides, stage rows,
// the fused paths (whole attempt for element-local RHS, stage + Lorenz-96 stencil) and the adaptive retry loop
// of commonAdaptiveMethodCode (ode.nim:57-76).
#include "internal.hpp"
#include "stencil_attempt.cuh"

int builtin_rhs_fn(double /*t*/, const b200rk_vec* y, b200rk_vec* dydt, void* user) {
  BuiltinRhs* r = static_cast<BuiltinRhs*>(user);
  b200rk_ctx* c = r->ctx;
  const size_t n = y->n_local;
  switch (r->kind) {
    case B200RK_RHS_SCALE:
      return launch_ewise(c, EW_SCALE, y->d, nullptr, r->scalar, dydt->d, n, B200RK_K_RHS);
    case B200RK_RHS_DIAG_LINEAR:
      TRY(check_same(c, y, r->lambda));
      return launch_ewise(c, EW_NEG_HMUL, r->lambda->d, y->d, 0.0, dydt->d, n, B200RK_K_RHS);
    case B200RK_RHS_LORENZ96: {
      if (y->n_global < 4) return fail(c, B200RK_EINVAL, "lorenz96 needs n >= 4");
      const double *left2 = y->d + n - 2, *right1 = y->d;  // single GPU: the cyclic neighbours are in the vector itself
      if (c->world > 1) {
        // Sharded stencil: 3-element halo per evaluation over NVLink. Each rank sends its
        // first element to the left neighbour and its last two to the right neighbour (ring), on the
        // context stream, so the exchange is ordered with the producing and consuming kernels.
        // the same verdict on every rank (shard_range is a pure function): a rank that bailed out alone would leave
        // its neighbours waiting in ncclSend/ncclRecv forever
        for (int r = 0; r < c->world; ++r) {
          size_t o_ = 0, l_ = 0;
          shard_range(y->n_global, r, c->world, &o_, &l_);
          if (l_ < 2) return fail(c, B200RK_EINVAL, "lorenz96: every shard needs at least 2 elements (rank " + std::to_string(r) + " would hold " + std::to_string(l_) + ")");
        }
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
      return launch_lorenz96(c, y->d, left2, right1, r->scalar, dydt->d, n);
    }
  }
  return fail(c, B200RK_EINVAL, "unknown builtin rhs");
}

int eval_rhs(b200rk_ctx* c, const RhsCall& r, double t, const b200rk_vec* y, b200rk_vec* out) {
  if (r.evals) ++*r.evals;
  int rc = r.f(r.negate_time ? -t : t, y, out, r.user);
  if (rc != 0) {
    if (r.f == &builtin_rhs_fn || r.f == &jit_rhs_fn) return rc;  // our own launcher already recorded the error
    return fail(c, B200RK_ECALLBACK, "right-hand side callback returned " + std::to_string(rc));
  }
  if (r.negate_time) return launch_ewise(c, EW_NEG, out->d, nullptr, 0.0, out->d, out->n_local, B200RK_K_OTHER);
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

int run_row(b200rk_ctx* c, const Row& row, double cfac, bool chain, double dt, const b200rk_vec* y,
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

int plan_finish(const b200rk_ctx* c, const MethodDef& md, double dt, double absTol, double relTol,
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

// ---- fused attempt for element-local right-hand sides (kernels.cuh: fused_attempt_kernel) -------------
// Element-local = the built-in c*y and -(lambda .* y), and every right-hand side given as source (jit.cu).
struct PwSpec {
  int kind = -1;                                      // PW_SCALE / PW_DIAG / PW_USER
  int np = 0;                                         // per-element parameter streams read next to y
  const b200rk_vec* pv[kMaxUserVecs] = {nullptr};
  double scalar = 0.0;                                // PW_SCALE
  const double* cs = nullptr;                         // PW_USER: c0..c7
  JitRhs* jit = nullptr;                              // PW_USER
};
static bool pointwise_spec(const RhsCall& rhs, PwSpec* pw) {
  if (rhs.f == &jit_rhs_fn) {
    if (jit_is_stencil(static_cast<JitRhs*>(rhs.user), nullptr, nullptr)) return false;   // reads its neighbours: not element-local
    pw->kind = PW_USER;
    pw->jit = static_cast<JitRhs*>(rhs.user);
    const b200rk_vec* const* vecs = nullptr;
    jit_describe(pw->jit, &pw->np, &vecs, &pw->cs);
    for (int j = 0; j < pw->np; ++j) pw->pv[j] = vecs[j];
    return true;
  }
  if (rhs.f != &builtin_rhs_fn) return false;
  const BuiltinRhs* r = static_cast<const BuiltinRhs*>(rhs.user);
  if (r->kind == B200RK_RHS_SCALE) { pw->kind = PW_SCALE; pw->scalar = r->scalar; }
  else if (r->kind == B200RK_RHS_DIAG_LINEAR) { pw->kind = PW_DIAG; pw->np = 1; pw->pv[0] = r->lambda; }
  else return false;
  return true;
}
static int check_pw_sizes(const b200rk_ctx* c, const PwSpec& pw, const b200rk_vec* y) {
  for (int j = 0; j < pw.np; ++j) TRY(check_same(c, y, pw.pv[j]));
  return B200RK_OK;
}
// Right-hand-side part of the fused kernels' argument block. `negate`: backward pass g(t, y) = -f(-t, y) (ode.nim:545).
template <int S>
static void fill_pw_args(FusedArgs<S>& a, const PwSpec& pw, const MethodDef& md, bool negate, double t) {
  for (int j = 0; j < pw.np; ++j) a.p[j] = pw.pv[j]->d;
  if (pw.kind == PW_USER) {
    a.rhs_sign = negate ? -1.0 : 1.0;                 // multiply by +-1: exact
    a.tsign = negate ? -1.0 : 1.0;
    a.t = t;
    for (int j = 0; j < kMaxUserScalars; ++j) a.cs[j] = pw.cs[j];
    for (int s = 2; s <= S; ++s) a.cnode[s - 1] = md.c[s];
  } else {
    a.rhs_scalar = negate ? -pw.scalar : pw.scalar;   // -(y*c) == y*(-c) exactly
    a.rhs_sign = negate ? 1.0 : -1.0;                 // k = (lam*y)*sign
  }
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

// The reduction scratch is taken AFTER ensure_partials: growing the partials array reallocates it, and a scratch
// block captured earlier would hand the kernel the freed pointer.
template <int PAT, int KIND>
static int launch_fused_cfg(b200rk_ctx* c, FusedArgs<Pattern<PAT>::S>& a) {
  const size_t n = a.n;
  if (c->vec_width == 4) {
    unsigned grid = grid_for(c, n / 4, kThreads, c->fused_ctas_per_sm);
    TRY(ensure_partials(c, grid));
    a.rs = reduce_scratch(c);
    // L2-resident lambda / yNew: measured neutral at 2^23 (64.8 vs 65.8 us; the kernel is co-limited by the fp64
    // pipe) and +3 % at 2^22, so only on explicit request, not under the auto policy
    if (c->l2_hints == 1) fused_attempt_kernel<PAT, KIND, 4, kThreads, 1><<<grid, kThreads, 0, c->stream>>>(a);
    else fused_attempt_kernel<PAT, KIND, 4, kThreads><<<grid, kThreads, 0, c->stream>>>(a);
  } else {
    unsigned grid = grid_for(c, n / 2, kThreads, c->fused_ctas_per_sm);
    TRY(ensure_partials(c, grid));
    a.rs = reduce_scratch(c);
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

int fused_pattern_for(const b200rk_ctx* c, const MethodDef& md) {
  return (!md.rk4_final && method_fusable(md)) ? fused_pattern_of(c, md) : -1;
}

template <int PAT>
static int launch_fused_pair(b200rk_ctx* c, const MethodDef& md, const PwSpec& pw, bool negate, double t, double dt,
                             const b200rk_options& o, const b200rk_vec* y, const b200rk_vec* fsal, b200rk_vec* y_new,
                             b200rk_vec* fsal_new) {
  constexpr int S = Pattern<PAT>::S;
  FusedArgs<S> a;
  std::memset(&a, 0, sizeof(a));
  a.y = y->d; a.k1 = fsal->d;
  fill_pw_args(a, pw, md, negate, t);
  for (int s = 2; s <= S; ++s) row_mask(c, md.a[s], a.a[s - 2], S - 1);
  row_mask(c, md.b, a.b, S);
  row_mask(c, md.bhat, a.bh, S);
  a.dt = dt; a.cb = dt; a.cbh = dt; a.absTol = o.absTol; a.relTol = o.relTol;
  a.ynew = y_new->d; a.ks_out = fsal_new->d; a.n = y->n_local;
  const int streams = 4 + pw.np;  // y, k1 (+ parameters) read; yNew, k_S written
  ProfScope ps(c, B200RK_K_FUSED, 8.0 * double(y->n_local) * streams);
  if (pw.kind == PW_USER) {  // the same kernel, compiled at run time around the caller's expression (jit.cu)
    const int W = (c->vec_width == 4) ? 4 : 2;
    const unsigned grid = grid_for(c, a.n / W, kThreads, c->fused_ctas_per_sm);
    TRY(ensure_partials(c, grid));
    a.rs = reduce_scratch(c);
    return jit_launch(c, pw.jit, PAT, jit_slot_attempt(W), grid, &a, false);
  }
  if (pw.kind == PW_SCALE) return launch_fused_cfg<PAT, PW_SCALE>(c, a);
  return launch_fused_cfg<PAT, PW_DIAG>(c, a);
}

static int launch_fused_rk4(b200rk_ctx* c, const PwSpec& pw, bool negate, double t, double dt, const b200rk_vec* y,
                            b200rk_vec* y_new) {
  const size_t n = y->n_local;
  if (!n) return B200RK_OK;
  if (pw.kind == PW_USER) return jit_launch_rk4(c, pw.jit, negate, t, dt, y, y_new);
  ProfScope ps(c, B200RK_K_FUSED, 8.0 * double(n) * (2 + pw.np));
  const double hdt = 0.5 * dt, c6 = dt / 6.0;
  const double* lam = pw.np ? pw.pv[0]->d : nullptr;
  unsigned grid = grid_for(c, n / 4, kThreads, c->ctas_per_sm);
  const double cs = negate ? -pw.scalar : pw.scalar, sgn = negate ? 1.0 : -1.0;
  if (pw.kind == PW_SCALE) fused_rk4_kernel<PW_SCALE, 4, kThreads><<<grid, kThreads, 0, c->stream>>>(y->d, lam, cs, sgn, hdt, dt, c6, y_new->d, n);
  else fused_rk4_kernel<PW_DIAG, 4, kThreads><<<grid, kThreads, 0, c->stream>>>(y->d, lam, cs, sgn, hdt, dt, c6, y_new->d, n);
  CUDA_TRY(c, cudaGetLastError());
  return B200RK_OK;
}

// Whole attempt of an FSAL pair with the built-in Lorenz-96 right-hand side in one kernel (stencil_attempt.cuh:
// overlapped tiles, stage inputs through shared memory). Knob "fuse_stencil_attempt" (default on).
// Sharded: every shard must hold at least the largest overlap, so that a halo comes from the immediate ring neighbour
// only (same answer on every rank: shard_range is a pure function of n, rank, world).
constexpr int kAttemptHaloMax = 2 * stencil_halo(kMaxStencilRadius, 9);   // built-in Lorenz-96 / Vern65: 16 + 8; stencils from source: up to 64 + 64
static bool stencil_shards_ok(const b200rk_ctx* c, size_t n_global, int need) {
  for (int r = 0; r < c->world; ++r) {
    size_t off = 0, len = 0;
    shard_range(n_global, r, c->world, &off, &len);
    if (len < (size_t)need) return false;
  }
  return true;
}
static bool l96_attempt_shards_ok(const b200rk_ctx* c, size_t n_global) { return stencil_shards_ok(c, n_global, StencilTile<9>::HL); }
// a stencil right-hand side given as source (jit.cu) and its radii
static JitRhs* jit_stencil_of(const RhsCall& rhs, int* rl, int* rr) {
  if (rhs.f != &jit_rhs_fn) return nullptr;
  JitRhs* j = static_cast<JitRhs*>(rhs.user);
  return jit_is_stencil(j, rl, rr) ? j : nullptr;
}
// The HL elements before this shard and the HR after it, of y and of k1 (FSAL): one grouped exchange with the ring
// neighbours on the context stream per IntegratorProc call — y and k1 do not change between the retries of an attempt.
// Layout of each halo array: [0, HL) = left neighbour's tail, [HL, HL + HR) = right neighbour's head.
static int exchange_attempt_halo(b200rk_ctx* c, const b200rk_vec* y, const b200rk_vec* fsal /* null: y only (RK4) */, int HL, int HR) {
  if (!c->d_halo_attempt) CUDA_TRY(c, cudaMalloc(&c->d_halo_attempt, 2 * kAttemptHaloMax * sizeof(double)));
  double *hy = c->d_halo_attempt, *hk = c->d_halo_attempt + kAttemptHaloMax;
  const size_t n = y->n_local;
  const int left = (c->rank + c->world - 1) % c->world, right = (c->rank + 1) % c->world;
  NCCL_TRY(c, g_nccl.GroupStart());   // same order on every rank: with world == 2 both neighbours are the same peer and the pairs match in order
  if (HR) NCCL_TRY(c, g_nccl.Send(y->d, HR, ncclDouble, left, c->comm, c->stream));     // (a one-sided stencil from source has HL or HR = 0)
  if (HL) NCCL_TRY(c, g_nccl.Send(y->d + n - HL, HL, ncclDouble, right, c->comm, c->stream));
  if (fsal) {
    if (HR) NCCL_TRY(c, g_nccl.Send(fsal->d, HR, ncclDouble, left, c->comm, c->stream));
    if (HL) NCCL_TRY(c, g_nccl.Send(fsal->d + n - HL, HL, ncclDouble, right, c->comm, c->stream));
  }
  if (HR) NCCL_TRY(c, g_nccl.Recv(hy + HL, HR, ncclDouble, right, c->comm, c->stream));
  if (HL) NCCL_TRY(c, g_nccl.Recv(hy, HL, ncclDouble, left, c->comm, c->stream));
  if (fsal) {
    if (HR) NCCL_TRY(c, g_nccl.Recv(hk + HL, HR, ncclDouble, right, c->comm, c->stream));
    if (HL) NCCL_TRY(c, g_nccl.Recv(hk, HL, ncclDouble, left, c->comm, c->stream));
  }
  NCCL_TRY(c, g_nccl.GroupEnd());
  c->collectives++;
  return B200RK_OK;
}

// Where the halo of this call's y (and k1 = fsal) lives. Inside an advancing solver whose ring neighbours' vectors are
// peer-mapped (ctx->peer_view) and for the adaptive pairs — whose all-reduce of the error norm keeps the ranks in
// lockstep, so a neighbour's y / FSAL of this attempt is complete and stays untouched until every rank has finished
// reading it — the halo is read in place over NVLink: no exchange, no collective. Otherwise (b200rk_step on arbitrary
// vectors, RK4, no peer mapping) one grouped ncclSend/ncclRecv fills ctx->d_halo_attempt.
static int l96_halo_for(b200rk_ctx* c, const MethodDef& md, const b200rk_vec* y, const b200rk_vec* fsal, int HL, int HR, L96Halo* h) {
  *h = L96Halo{nullptr, nullptr, nullptr, nullptr};
  if (c->world < 2) return B200RK_OK;
  const PeerVecView* pv = c->peer_view;
  const int iy = pv ? pv->find(y->d) : -1, ik = (pv && fsal) ? pv->find(fsal->d) : -1;
  if (!md.rk4_final && iy >= 0 && ik >= 0) {
    h->left_y = pv->left[iy] + (pv->n_left - HL); h->right_y = pv->right[iy];
    h->left_k = pv->left[ik] + (pv->n_left - HL); h->right_k = pv->right[ik];
    return B200RK_OK;
  }
  TRY(exchange_attempt_halo(c, y, md.rk4_final ? nullptr : fsal, HL, HR));
  h->left_y = c->d_halo_attempt; h->right_y = c->d_halo_attempt + HL;
  h->left_k = c->d_halo_attempt + kAttemptHaloMax; h->right_k = c->d_halo_attempt + kAttemptHaloMax + HL;
  return B200RK_OK;
}

bool l96_peer_halo_possible(const b200rk_ctx* c, const MethodDef& md, const RhsCall& rhs, size_t n_global) {
  if (!(c->world > 1 && c->p2p && c->fuse_stencil_attempt && c->l96_peer_halo && md.adaptive && md.use_fsal && method_fusable(md) &&
        !md.rk4_final && fused_pattern_of(c, md) >= 0))
    return false;
  int rl = 0, rr = 0;
  if (jit_stencil_of(rhs, &rl, &rr)) return stencil_shards_ok(c, n_global, std::max(stencil_halo(rl, md.stages), stencil_halo(rr, md.stages)));
  return rhs.f == &builtin_rhs_fn && static_cast<const BuiltinRhs*>(rhs.user)->kind == B200RK_RHS_LORENZ96 && l96_attempt_shards_ok(c, n_global);
}

template <int PAT, int J, bool HALF = false>   // HALF: 128-thread CTAs (half-width tiles, twice the CTAs per SM, barrier domains of 4 warps)
static int launch_l96_attempt_j(b200rk_ctx* c, const MethodDef& md, double F, bool negate, double dt, const b200rk_options& o,
                                const L96Halo& halo, const b200rk_vec* y, const b200rk_vec* fsal, b200rk_vec* y_new, b200rk_vec* fsal_new) {
  constexpr int S = Pattern<PAT>::S;
  constexpr int T = HALF ? kThreads / 2 : kThreads;
  constexpr int OUT = 2 * J * T - StencilTile<S>::HL - StencilTile<S>::HR;
  L96AttemptArgs<S> a;
  std::memset(&a, 0, sizeof(a));
  a.f.y = y->d; a.f.k1 = fsal->d;
  for (int s = 2; s <= S; ++s) row_mask(c, md.a[s], a.f.a[s - 2], S - 1);
  row_mask(c, md.b, a.f.b, S);
  row_mask(c, md.bhat, a.f.bh, S);
  a.f.dt = dt; a.f.cb = dt; a.f.cbh = dt; a.f.absTol = o.absTol; a.f.relTol = o.relTol;
  a.f.ynew = y_new->d; a.f.ks_out = fsal_new->d; a.f.n = y->n_local;
  a.F = F;
  a.halo = halo;
  // persistent grid: as many CTAs as fit on the device at once (3 per SM at 78-80 registers, 2 for Vern65), each walking
  // its share of the tiles with the next tile's bulk copies in flight
  static int per_sm = 0;   // per instantiation
  if (per_sm == 0) {
    CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, l96_attempt_kernel<PAT, J, T, false>, T, 0));
    if (per_sm < 1) per_sm = 1;
  }
  const size_t n_tiles = (a.f.n + OUT - 1) / OUT;
  const int want = c->l96_ctas_per_sm > 0 ? c->l96_ctas_per_sm : per_sm;
  const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>(n_tiles, (size_t)want * c->sm_count));
  TRY(ensure_partials(c, grid));
  a.f.rs = reduce_scratch(c);
  ProfScope ps(c, B200RK_K_FUSED, 8.0 * double(a.f.n) * 4);  // y, k1 read; yNew, k_S written
  if constexpr (HALF) {
    if (negate) l96_attempt_kernel<PAT, J, kThreads / 2, true><<<grid, kThreads / 2, 0, c->stream>>>(a);
    else l96_attempt_kernel<PAT, J, kThreads / 2, false><<<grid, kThreads / 2, 0, c->stream>>>(a);
  } else {
    if (negate) l96_attempt_kernel<PAT, J, kThreads, true><<<grid, kThreads, 0, c->stream>>>(a);
    else l96_attempt_kernel<PAT, J, kThreads, false><<<grid, kThreads, 0, c->stream>>>(a);
  }
  CUDA_TRY(c, cudaGetLastError());
  return B200RK_OK;
}

// Whole attempt of an FSAL pair for a stencil right-hand side given as SOURCE: the NVRTC-compiled ustencil_attempt_kernel
// (stencil_attempt.cuh) of this pattern — same tiling, same TMA prefetch, overlap from the stencil's radii.
template <int PAT>
static int launch_ustencil_attempt(b200rk_ctx* c, const MethodDef& md, JitRhs* jit, int rl, int rr, bool negate, double t, double dt,
                                   const b200rk_options& o, const L96Halo& halo, const b200rk_vec* y, const b200rk_vec* fsal, b200rk_vec* y_new,
                                   b200rk_vec* fsal_new) {
  constexpr int S = Pattern<PAT>::S;
  const int HL = stencil_halo(rl, S), HR = stencil_halo(rr, S), OUT = 4 * kThreads - HL - HR;
  UStencilAttemptArgs<S> a;
  std::memset(&a, 0, sizeof(a));
  PwSpec pw;
  pw.kind = PW_USER; pw.jit = jit;
  const b200rk_vec* const* vecs = nullptr;
  jit_describe(jit, &pw.np, &vecs, &pw.cs);
  for (int j = 0; j < pw.np; ++j) pw.pv[j] = vecs[j];
  TRY(check_pw_sizes(c, pw, y));
  a.f.y = y->d; a.f.k1 = fsal->d;
  fill_pw_args(a.f, pw, md, negate, t);
  for (int s = 2; s <= S; ++s) row_mask(c, md.a[s], a.f.a[s - 2], S - 1);
  row_mask(c, md.b, a.f.b, S);
  row_mask(c, md.bhat, a.f.bh, S);
  a.f.dt = dt; a.f.cb = dt; a.f.cbh = dt; a.f.absTol = o.absTol; a.f.relTol = o.relTol;
  a.f.ynew = y_new->d; a.f.ks_out = fsal_new->d; a.f.n = y->n_local;
  a.halo = halo;
  int per_sm = 0;
  TRY(jit_max_blocks_per_sm(c, jit, PAT, 0, &per_sm));
  if (per_sm < 1) return fail(c, B200RK_ECUDA, "stencil attempt kernel does not fit on an SM");
  const size_t n_tiles = (a.f.n + OUT - 1) / OUT;
  const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>(n_tiles, (size_t)per_sm * c->sm_count));
  TRY(ensure_partials(c, grid));
  a.f.rs = reduce_scratch(c);
  ProfScope ps(c, B200RK_K_FUSED, 8.0 * double(a.f.n) * (4 + pw.np));  // y, k1 (+ parameters) read; yNew, k_S written
  return jit_launch(c, jit, PAT, 0, grid, &a, false);
}

// Warp-sized tiles (stencil_attempt.cuh: l96_warp_attempt_kernel): same argument block, persistent grid of warps; E elements per lane.
template <int PAT, int E>
static int launch_l96_warp_e(b200rk_ctx* c, const MethodDef& md, double F, bool negate, double dt, const b200rk_options& o,
                             const L96Halo& halo, const b200rk_vec* y, const b200rk_vec* fsal, b200rk_vec* y_new, b200rk_vec* fsal_new) {
  constexpr int S = Pattern<PAT>::S;
  constexpr int OUT = WarpTile<S, E>::OUT, WPB = kThreads / 32;
  L96AttemptArgs<S> a;
  std::memset(&a, 0, sizeof(a));
  a.f.y = y->d; a.f.k1 = fsal->d;
  for (int s = 2; s <= S; ++s) row_mask(c, md.a[s], a.f.a[s - 2], S - 1);
  row_mask(c, md.b, a.f.b, S);
  row_mask(c, md.bhat, a.f.bh, S);
  a.f.dt = dt; a.f.cb = dt; a.f.cbh = dt; a.f.absTol = o.absTol; a.f.relTol = o.relTol;
  a.f.ynew = y_new->d; a.f.ks_out = fsal_new->d; a.f.n = y->n_local;
  a.F = F;
  a.halo = halo;
  static int per_sm = 0;   // per instantiation
  if (per_sm == 0) {
    CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, l96_warp_attempt_kernel<PAT, E, kThreads, false>, kThreads, 0));
    if (per_sm < 1) per_sm = 1;
  }
  const size_t n_tiles = (a.f.n + OUT - 1) / OUT;
  const int want = c->l96_ctas_per_sm > 0 ? c->l96_ctas_per_sm : per_sm;
  const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>((n_tiles + WPB - 1) / WPB, (size_t)want * c->sm_count));
  TRY(ensure_partials(c, grid));
  a.f.rs = reduce_scratch(c);
  ProfScope ps(c, B200RK_K_FUSED, 8.0 * double(a.f.n) * 4);  // y, k1 read; yNew, k_S written
  if (negate) l96_warp_attempt_kernel<PAT, E, kThreads, true><<<grid, kThreads, 0, c->stream>>>(a);
  else l96_warp_attempt_kernel<PAT, E, kThreads, false><<<grid, kThreads, 0, c->stream>>>(a);
  CUDA_TRY(c, cudaGetLastError());
  return B200RK_OK;
}
template <int PAT>
static int launch_l96_warp(b200rk_ctx* c, const MethodDef& md, double F, bool negate, double dt, const b200rk_options& o,
                           const L96Halo& halo, const b200rk_vec* y, const b200rk_vec* fsal, b200rk_vec* y_new, b200rk_vec* fsal_new) {
  if (c->l96_warp_tiles == 4) return launch_l96_warp_e<PAT, 4>(c, md, F, negate, dt, o, halo, y, fsal, y_new, fsal_new);
  return launch_l96_warp_e<PAT, 8>(c, md, F, negate, dt, o, halo, y, fsal, y_new, fsal_new);
}

// Tile width = 512 * J positions (knob "l96_attempt_pairs": 2 = 1024-wide tiles, 2 % overlap, 80-98 registers; 1 = 512-wide,
// 4 % overlap, fewer registers and more resident CTAs) — to be settled by measurement.
template <int PAT>
static int launch_l96_attempt(b200rk_ctx* c, const MethodDef& md, double F, bool negate, double dt, const b200rk_options& o,
                              const L96Halo& halo, const b200rk_vec* y, const b200rk_vec* fsal, b200rk_vec* y_new, b200rk_vec* fsal_new) {
  if (c->l96_warp_tiles) return launch_l96_warp<PAT>(c, md, F, negate, dt, o, halo, y, fsal, y_new, fsal_new);
  if (c->l96_attempt_threads == 128) return launch_l96_attempt_j<PAT, 2, true>(c, md, F, negate, dt, o, halo, y, fsal, y_new, fsal_new);
  if (c->l96_attempt_pairs == 1) return launch_l96_attempt_j<PAT, 1>(c, md, F, negate, dt, o, halo, y, fsal, y_new, fsal_new);
  return launch_l96_attempt_j<PAT, 2>(c, md, F, negate, dt, o, halo, y, fsal, y_new, fsal_new);
}

// A whole RK4 step with the built-in Lorenz-96 right-hand side in one kernel (stencil_attempt.cuh: l96_rk4_kernel).
static int launch_l96_rk4(b200rk_ctx* c, double F, bool negate, double dt, const L96Halo& halo, const b200rk_vec* y, b200rk_vec* y_new) {
  constexpr int J = 2, OUT = 2 * J * kThreads - 8 - 4;
  L96Rk4Args a;
  std::memset(&a, 0, sizeof(a));
  a.y = y->d; a.ynew = y_new->d; a.n = y->n_local;
  a.F = F; a.hdt = 0.5 * dt; a.dt = dt; a.c6 = dt / 6.0;   // same host scalars as launch_fused_rk4 / launch_rk4_final
  a.halo = halo;
  if (!a.n) return B200RK_OK;
  const unsigned grid = (unsigned)((a.n + OUT - 1) / OUT);
  ProfScope ps(c, B200RK_K_FUSED, 8.0 * double(a.n) * 2);   // y read, yNew written
  if (negate) l96_rk4_kernel<J, kThreads, true><<<grid, kThreads, 0, c->stream>>>(a);
  else l96_rk4_kernel<J, kThreads, false><<<grid, kThreads, 0, c->stream>>>(a);
  CUDA_TRY(c, cudaGetLastError());
  return B200RK_OK;
}

// One IntegratorProc call. y, fsal read-only; y_new, fsal_new written.
int do_step(b200rk_ctx* c, const MethodDef& md, const RhsCall& rhs, double t, const b200rk_vec* y,
                   const b200rk_vec* fsal, double dt_in, const b200rk_options& o, b200rk_vec* y_new,
                   b200rk_vec* fsal_new, double* dt_used, double* error_out, StepCounters* cnt) {
  const size_t N = y->n_global;
  const int S = md.stages;
  Workspace ws(c);
  b200rk_vec* k[kMaxStages + 1] = {nullptr};
  b200rk_vec* tmp = nullptr;
  PwSpec pw;
  int fused_pat = -1;
  bool fused = c->fuse_pointwise && pointwise_spec(rhs, &pw) && method_fusable(md) &&
               (md.rk4_final || (fsal && fsal_new));
  if (fused && !md.rk4_final) { fused_pat = fused_pattern_of(c, md); fused = fused_pat >= 0; }
  const bool stencil_fused = !fused && c->fuse_stencil && c->world == 1 && rhs.f == &builtin_rhs_fn &&
                             static_cast<const BuiltinRhs*>(rhs.user)->kind == B200RK_RHS_LORENZ96 && y->n_global >= 4;
  // the whole attempt of the stencil right-hand side in one kernel (overlapped tiles read y and FSAL far
  // beyond a CTA's own outputs, so the outputs must not alias the inputs)
  const bool l96_builtin = !fused && rhs.f == &builtin_rhs_fn && static_cast<const BuiltinRhs*>(rhs.user)->kind == B200RK_RHS_LORENZ96 &&
                           y->n_global >= 4;
  bool l96_attempt = l96_builtin && c->fuse_stencil_attempt && (c->world == 1 || l96_attempt_shards_ok(c, y->n_global)) && method_fusable(md) &&
                     y_new->d != y->d;
  if (l96_attempt && !md.rk4_final)
    l96_attempt = fsal && fsal_new && y_new->d != fsal->d && fsal_new->d != y->d && fsal_new->d != fsal->d;
  if (l96_attempt && !md.rk4_final) { fused_pat = fused_pattern_of(c, md); l96_attempt = fused_pat >= 0; }
  // a stencil right-hand side given as source: the same one-kernel attempt, compiled at run time around the expression
  int ust_rl = 0, ust_rr = 0;
  JitRhs* ust_jit = fused ? nullptr : jit_stencil_of(rhs, &ust_rl, &ust_rr);
  const int ust_HL = stencil_halo(ust_rl, S), ust_HR = stencil_halo(ust_rr, S);
  bool ust_attempt = ust_jit && c->fuse_stencil_attempt && !md.rk4_final && method_fusable(md) && fsal && fsal_new && y_new->d != y->d &&
                     y_new->d != fsal->d && fsal_new->d != y->d && fsal_new->d != fsal->d && y->n_global >= (size_t)(ust_rl + ust_rr + 1) &&
                     (c->world == 1 || stencil_shards_ok(c, y->n_global, std::max(ust_HL, ust_HR)));
  if (ust_attempt) { fused_pat = fused_pattern_of(c, md); ust_attempt = fused_pat >= 0; }
  // ... and a whole RK4 step in one kernel (4 evaluations: overlap 4 * radius)
  const int rk4_HL = stencil_halo(ust_rl, 5), rk4_HR = stencil_halo(ust_rr, 5);
  const bool ust_rk4 = ust_jit && c->fuse_stencil_attempt && md.rk4_final && y_new->d != y->d && y->n_global >= (size_t)(ust_rl + ust_rr + 1) &&
                       (c->world == 1 || stencil_shards_ok(c, y->n_global, std::max(rk4_HL, rk4_HR)));
  if (fused) TRY(check_pw_sizes(c, pw, y));
  fused = fused || l96_attempt || ust_attempt || ust_rk4;
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

  L96Halo l96_halo{nullptr, nullptr, nullptr, nullptr};
  if (l96_attempt) {  // y and k1 do not change between the retries of an attempt: once per call
    if (md.rk4_final) TRY(l96_halo_for(c, md, y, nullptr, 8, 4, &l96_halo));
    else TRY(l96_halo_for(c, md, y, fsal, S == 9 ? StencilTile<9>::HL : StencilTile<7>::HL, S == 9 ? StencilTile<9>::HR : StencilTile<7>::HR, &l96_halo));
  }
  if (ust_attempt) TRY(l96_halo_for(c, md, y, fsal, ust_HL, ust_HR, &l96_halo));
  if (ust_rk4) TRY(l96_halo_for(c, md, y, nullptr, rk4_HL, rk4_HR, &l96_halo));
  double dt = dt_in, error = 0.0;
  int limitCounter = 0;
  while (true) {
    if (cnt) cnt->attempts++;
    if (fused) {
      // element-local right-hand side: the whole attempt is one kernel (the callbacks it stands for are
      // still counted so rhs_evals matches the unfused path)
      if (rhs.evals) *rhs.evals += md.rk4_final ? 4 : (S - 1);
      if (md.rk4_final) {
        if (l96_attempt) TRY(launch_l96_rk4(c, static_cast<const BuiltinRhs*>(rhs.user)->scalar, rhs.negate_time, dt, l96_halo, y, y_new));
        else if (ust_rk4) TRY(jit_launch_stencil_rk4(c, ust_jit, rhs.negate_time, t, dt, l96_halo, y, y_new));
        else TRY(launch_fused_rk4(c, pw, rhs.negate_time, t, dt, y, y_new));
        break;
      }
      if (l96_attempt) {
        const double F = static_cast<const BuiltinRhs*>(rhs.user)->scalar;
        switch (fused_pat) {
          case PAT_DOPRI54: TRY(launch_l96_attempt<PAT_DOPRI54>(c, md, F, rhs.negate_time, dt, o, l96_halo, y, fsal, y_new, fsal_new)); break;
          case PAT_DOPRI54_STRICT: TRY(launch_l96_attempt<PAT_DOPRI54_STRICT>(c, md, F, rhs.negate_time, dt, o, l96_halo, y, fsal, y_new, fsal_new)); break;
          case PAT_TSIT54: TRY(launch_l96_attempt<PAT_TSIT54>(c, md, F, rhs.negate_time, dt, o, l96_halo, y, fsal, y_new, fsal_new)); break;
          case PAT_VERN65: TRY(launch_l96_attempt<PAT_VERN65>(c, md, F, rhs.negate_time, dt, o, l96_halo, y, fsal, y_new, fsal_new)); break;
          default: TRY(launch_l96_attempt<PAT_VERN65_STRICT>(c, md, F, rhs.negate_time, dt, o, l96_halo, y, fsal, y_new, fsal_new)); break;
        }
      } else if (ust_attempt) {
        switch (fused_pat) {
          case PAT_DOPRI54: TRY(launch_ustencil_attempt<PAT_DOPRI54>(c, md, ust_jit, ust_rl, ust_rr, rhs.negate_time, t, dt, o, l96_halo, y, fsal, y_new, fsal_new)); break;
          case PAT_DOPRI54_STRICT: TRY(launch_ustencil_attempt<PAT_DOPRI54_STRICT>(c, md, ust_jit, ust_rl, ust_rr, rhs.negate_time, t, dt, o, l96_halo, y, fsal, y_new, fsal_new)); break;
          case PAT_TSIT54: TRY(launch_ustencil_attempt<PAT_TSIT54>(c, md, ust_jit, ust_rl, ust_rr, rhs.negate_time, t, dt, o, l96_halo, y, fsal, y_new, fsal_new)); break;
          case PAT_VERN65: TRY(launch_ustencil_attempt<PAT_VERN65>(c, md, ust_jit, ust_rl, ust_rr, rhs.negate_time, t, dt, o, l96_halo, y, fsal, y_new, fsal_new)); break;
          default: TRY(launch_ustencil_attempt<PAT_VERN65_STRICT>(c, md, ust_jit, ust_rl, ust_rr, rhs.negate_time, t, dt, o, l96_halo, y, fsal, y_new, fsal_new)); break;
        }
      } else
      switch (fused_pat) {
        case PAT_DOPRI54: TRY(launch_fused_pair<PAT_DOPRI54>(c, md, pw, rhs.negate_time, t, dt, o, y, fsal, y_new, fsal_new)); break;
        case PAT_DOPRI54_STRICT: TRY(launch_fused_pair<PAT_DOPRI54_STRICT>(c, md, pw, rhs.negate_time, t, dt, o, y, fsal, y_new, fsal_new)); break;
        case PAT_TSIT54: TRY(launch_fused_pair<PAT_TSIT54>(c, md, pw, rhs.negate_time, t, dt, o, y, fsal, y_new, fsal_new)); break;
        case PAT_VERN65: TRY(launch_fused_pair<PAT_VERN65>(c, md, pw, rhs.negate_time, t, dt, o, y, fsal, y_new, fsal_new)); break;
        default: TRY(launch_fused_pair<PAT_VERN65_STRICT>(c, md, pw, rhs.negate_time, t, dt, o, y, fsal, y_new, fsal_new)); break;
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
        TRY(launch_rk4_final(c, y->d, k[1]->d, k[2]->d, k[3]->d, k[4]->d, dt / 6.0, y_new->d, y->n_local));
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
// device-resident driver loop (small N): one cooperative kernel for many accepted steps
// =====================================================================================================
bool device_loop_eligible(const b200rk_ctx* c, const MethodDef& md, const RhsCall& rhs, size_t n_local) {
  if (c->device_loop == 0 || !c->fuse_pointwise || n_local == 0) return false;
  if (c->world > 1 && !c->p2p) return false;  // sharded: needs the peer mailboxes for the in-kernel all-reduce
  // measured (profiles/r01_sweep_small_device_loop.json): 2.4x at 2^16 (5.8 vs 14.1 us/step), still +5 % at 2^23, so
  // the auto policy (-1) takes the device loop at every size
  PwSpec pw;
  if (!pointwise_spec(rhs, &pw) || md.rk4_final || !method_fusable(md)) return false;
  return fused_pattern_of(c, md) >= 0;
}

// all CTAs must be co-resident (grid barrier): at most per_sm * SMs, at most one tile each
static unsigned run_grid(const b200rk_ctx* c, size_t n, int W, int per_sm) {
  const size_t tiles = std::max<size_t>(1, (n / W + kThreads - 1) / kThreads);
  return (unsigned)std::min<size_t>(tiles, (size_t)per_sm * c->sm_count);
}
template <int PAT, int KIND, int W>
static int launch_run_cfg(b200rk_ctx* c, RunArgs<Pattern<PAT>::S>& a) {
  auto kernel = fused_run_kernel<PAT, KIND, W, kThreads>;
  int per_sm = 0;
  CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0));
  if (per_sm < 1) return fail(c, B200RK_ECUDA, "device loop: kernel does not fit on an SM");
  const unsigned grid = run_grid(c, a.f.n, W, per_sm);
  TRY(ensure_partials(c, 2 * (size_t)grid));
  a.partials = c->d_partials;
  void* args[] = {&a};
  CUDA_TRY(c, cudaLaunchCooperativeKernel((void*)kernel, dim3(grid), dim3(kThreads), args, 0, c->stream));
  return B200RK_OK;
}
template <int PAT>
static int launch_run_pat(b200rk_ctx* c, const PwSpec& pw, RunArgs<Pattern<PAT>::S>& a) {
  const bool w4 = a.f.n >= ((size_t)1 << 20);
  if (pw.kind == PW_USER) {  // run-time compiled instance of the same kernel (jit.cu)
    const int W = w4 ? 4 : 2, slot = jit_slot_run(W);
    int per_sm = 0;
    TRY(jit_max_blocks_per_sm(c, pw.jit, PAT, slot, &per_sm));
    if (per_sm < 1) return fail(c, B200RK_ECUDA, "device loop: kernel does not fit on an SM");
    const unsigned grid = run_grid(c, a.f.n, W, per_sm);
    TRY(ensure_partials(c, 2 * (size_t)grid));
    a.partials = c->d_partials;
    return jit_launch(c, pw.jit, PAT, slot, grid, &a, true);
  }
  if (pw.kind == PW_SCALE) return w4 ? launch_run_cfg<PAT, PW_SCALE, 4>(c, a) : launch_run_cfg<PAT, PW_SCALE, 2>(c, a);
  return w4 ? launch_run_cfg<PAT, PW_DIAG, 4>(c, a) : launch_run_cfg<PAT, PW_DIAG, 2>(c, a);
}

template <int PAT>
static int run_device_loop_pat(b200rk_ctx* c, const MethodDef& md, const PwSpec& pw, bool negate,
                               const b200rk_options& o, DeviceLoopIO* io, int64_t max_steps) {
  constexpr int S = Pattern<PAT>::S;
  RunArgs<S> a;
  std::memset(&a, 0, sizeof(a));
  fill_pw_args(a.f, pw, md, negate, io->t);  // the kernel refreshes f.t at every step
  for (int s = 2; s <= S; ++s) row_mask(c, md.a[s], a.f.a[s - 2], S - 1);
  row_mask(c, md.b, a.f.b, S);
  row_mask(c, md.bhat, a.f.bh, S);
  a.f.absTol = o.absTol; a.f.relTol = o.relTol; a.f.n = io->Y[0]->n_local;
  for (int i = 0; i < 2; ++i) { a.Y[i] = io->Y[i]->d; a.F[i] = io->F[i]->d; }
  a.dtMin = o.dtMin; a.dtMax = o.dtMax;
  a.inv_order_inner = 1.0 / double(md.order_int);   // ode.nim:71 (int order)
  a.inv_order_outer = 1.0 / md.order;               // ode.nim:537 (float order)
  a.n_global = double(io->Y[0]->n_global);
  a.max_steps = max_steps < 0 ? (long long)1 << 62 : (long long)max_steps;
  if (!c->d_run_state) {
    CUDA_TRY(c, cudaMalloc(&c->d_run_state, sizeof(RunState) + 64));   // + the CTA arrival counter of the in-kernel sum exchange
    CUDA_TRY(c, cudaMemset(c->d_run_state, 0, sizeof(RunState) + 64));
    CUDA_TRY(c, cudaHostAlloc(&c->h_run_state, sizeof(RunState), cudaHostAllocMapped));
    CUDA_TRY(c, cudaHostGetDevicePointer(&c->h_run_state_dev, c->h_run_state, 0));
  }
  RunState st{};
  st.t = io->t; st.dt = io->dt; st.t_end = io->t_end; st.error = io->error; st.cur = io->cur;
  CUDA_TRY(c, cudaMemcpyAsync(c->d_run_state, &st, sizeof(st), cudaMemcpyHostToDevice, c->stream));
  a.state = c->d_run_state;
  a.state_host = c->h_run_state_dev;
  a.seq_host = c->h_seq_dev;
  a.seq = c->seq + 1;  // attempt i of this launch uses sequence number seq + i (same on every rank: lockstep)
  // The attempt's sum travels through mailboxes (kernels.cuh: mail_put / mail_wait): the peers' when sharded, this GPU's
  // own single slot otherwise — the same code path, and no grid barrier in either case.
  TRY(ensure_local_mailbox(c));
  a.mail.world = (c->world > 1 && c->p2p) ? c->world : 1;
  a.mail.rank = a.mail.world > 1 ? c->rank : 0;
  a.mail.timeout_cycles = c->peer_timeout_cycles;
  for (int p = 0; p < kMaxPeers; ++p) a.mail.box[p] = (a.mail.world > 1 && p < c->world) ? c->peer_mail[p] : nullptr;
  if (a.mail.world == 1) a.mail.box[0] = c->d_mail;
  a.arrive = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(c->d_run_state) + ((sizeof(RunState) + 31) / 32) * 32);
  CUDA_TRY(c, cudaMemsetAsync(a.arrive, 0, sizeof(unsigned long long), c->stream));
  const size_t prof_slot = c->prof.size();
  {
    ProfScope ps(c, B200RK_K_FUSED, 0.0);  // bytes patched below: the number of attempts is data-dependent
    TRY(launch_run_pat<PAT>(c, pw, a));
  }
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));  // one wait per launch (many steps), not per attempt
  const RunState out = *c->h_run_state;
  c->seq += (unsigned long long)out.attempts;  // attempts used sequence numbers seq+1 .. seq+attempts
  if (a.mail.world > 1) c->collectives += out.attempts;
  if (c->profile && c->prof.size() > prof_slot)  // y, k1 (+ lambda) read, yNew and k_S written, per attempt
    c->prof[prof_slot].bytes = 8.0 * double(a.f.n) * (4 + pw.np) * double(out.attempts);
  io->t = out.t; io->dt = out.dt; io->error = out.error; io->cur = out.cur;
  io->steps = out.steps; io->attempts = out.attempts; io->rejected = out.rejected; io->limiter_hits = out.limiter_hits;
  if (out.status == 2) return fail(c, B200RK_ENCCL, "peer mailbox all-reduce timed out inside the device loop");
  if (out.status != 0) return fail(c, B200RK_ENONFINITE, "error norm is NaN (the reference would loop forever here)");
  return B200RK_OK;
}

int run_device_loop(b200rk_ctx* c, const MethodDef& md, const RhsCall& rhs, const b200rk_options& o, DeviceLoopIO* io,
                    int64_t max_steps) {
  PwSpec pw;
  if (!pointwise_spec(rhs, &pw)) return fail(c, B200RK_EINVAL, "device loop: right-hand side is not element-local");
  TRY(check_pw_sizes(c, pw, io->Y[0]));
  switch (fused_pattern_of(c, md)) {
    case PAT_DOPRI54: return run_device_loop_pat<PAT_DOPRI54>(c, md, pw, rhs.negate_time, o, io, max_steps);
    case PAT_DOPRI54_STRICT: return run_device_loop_pat<PAT_DOPRI54_STRICT>(c, md, pw, rhs.negate_time, o, io, max_steps);
    case PAT_TSIT54: return run_device_loop_pat<PAT_TSIT54>(c, md, pw, rhs.negate_time, o, io, max_steps);
    case PAT_VERN65: return run_device_loop_pat<PAT_VERN65>(c, md, pw, rhs.negate_time, o, io, max_steps);
    case PAT_VERN65_STRICT: return run_device_loop_pat<PAT_VERN65_STRICT>(c, md, pw, rhs.negate_time, o, io, max_steps);
  }
  return fail(c, B200RK_EINVAL, "device loop: unsupported method");
}
