// driver.cu — the ODESolver loop (ode.nim:471-586) as a resumable object, and the step / solve entry points.
#include "internal.hpp"

#include <limits>
#include <thread>

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
  // sharded one-kernel Lorenz-96 attempt: Y[0..1], F[0..1] of the two ring neighbours, peer-mapped (runtime.cu)
  PeerVecView peers;

  ~b200rk_solver() {
    if (c && c->peer_view == &peers) c->peer_view = nullptr;
    if (peers.count) peer_view_close(c, &peers);   // collective: every rank frees its solver at the same point of the program
    for (auto* v : {Y[0], Y[1], F[0], F[1], LDY, SCR}) if (v) vec_release(v);
  }
};

// ctx->peer_view names the advancing solver's mapping for the duration of one solver_begin / solver_advance call
struct PeerViewScope {
  b200rk_ctx* c;
  const PeerVecView* prev;
  PeerViewScope(b200rk_ctx* ctx, const PeerVecView* v) : c(ctx), prev(ctx->peer_view) { if (v->count) c->peer_view = v; }
  ~PeerViewScope() { c->peer_view = prev; }
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


void hermite_factors(double x, double x1, double x2, double* f) {
  // utils.nim:273-279 — scalars on the host in the reference's order
  const double t = (x - x1) / (x2 - x1);
  const double u = 1.0 - t;
  const double h00 = (1.0 + 2.0 * t) * (u * u);
  const double h10 = t * (u * u);
  const double h01 = (t * t) * (3.0 - 2.0 * t);
  const double h11 = (t * (t * t)) - (t * t);
  f[0] = h00; f[1] = h10 * (x2 - x1); f[2] = h01; f[3] = h11 * (x2 - x1);
}

int hermite_into(b200rk_ctx* c, b200rk_vec* out, double x, double x1, double x2, const b200rk_vec* y1,
                        const b200rk_vec* y2, const b200rk_vec* dy1, const b200rk_vec* dy2) {
  double f[4];
  hermite_factors(x, x1, x2, f);
  return launch_hermite(c, y1->d, dy1->d, y2->d, dy2->d, f[0], f[1], f[2], f[3], out->d, y1->n_local);
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
  // peer-read halos: a neighbour's first attempt reads this rank's Y[0] / F[0] in place, so they must be complete on every
  // rank before any rank goes on (from then on the per-attempt all-reduce of the error norm keeps the ranks in lockstep)
  if (s->peers.count) TRY(stream_barrier(c));
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
  PeerViewScope peer_scope(c, &s->peers);
  const MethodDef& md = *s->md;
  int64_t done = 0;
  const long high = (long)s->targets.size() - 1;
  // Small N, element-local built-in RHS, no dense output: the whole loop below runs inside ONE persistent
  // cooperative kernel (grid barrier per attempt, controller on the device) — kernels.cuh: fused_run_kernel.
  if (!s->finished && !s->dense && md.use_fsal && s->cur == s->fcur && s->t < s->tEnd &&
      device_loop_eligible(c, md, s->rhs, s->Y[0]->n_local)) {
    DeviceLoopIO io{{s->Y[0], s->Y[1]}, {s->F[0], s->F[1]}, s->cur, s->t, s->dt, s->tEnd, s->error, 0, 0, 0, 0};
    TRY(run_device_loop(c, md, s->rhs, s->o, &io, max_steps));
    s->cur = s->fcur = io.cur;
    s->t = io.t; s->dt = io.dt; s->error = io.error;
    s->stats.steps += io.steps;
    s->stats.rhs_evals += io.attempts * (md.stages - 1);
    s->cnt.attempts += io.attempts; s->cnt.rejected += io.rejected; s->cnt.limiter_hits += io.limiter_hits;
    done = io.steps;
    if (s->t < s->tEnd) { if (steps_done) *steps_done = done; return B200RK_OK; }  // stopped at max_steps
  }
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
  if (rc == B200RK_OK && f == &jit_rhs_fn) {
    // right-hand side from source: every rank compiles and loads its units now, then the ranks meet on the stream, so the
    // first attempt's in-kernel all-reduce does not have to absorb a multi-second NVRTC compile on one of them
    rc = jit_prepare(c, static_cast<JitRhs*>(user), c->fuse_pointwise ? fused_pattern_for(c, *s->md) : -1);
    if (rc == B200RK_OK && c->world > 1) { rc = stream_barrier(c); c->collectives--; }   // set-up, not a data-path collective: not counted
  }
  if (rc == B200RK_OK && l96_peer_halo_possible(c, *s->md, s->rhs, N)) {
    b200rk_vec* const vecs[4] = {s->Y[0], s->Y[1], s->F[0], s->F[1]};
    rc = peer_view_open(c, vecs, 4, &s->peers);   // peers.count stays 0 when the mapping is not available: ncclSend/ncclRecv halo
  }
  if (rc != B200RK_OK) { delete s; return rc; }
  *out = s;
  return B200RK_OK;
}

extern "C" {

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
  std::vector<double> ts;
  for (size_t i = 0; i < n_tspan; ++i) if (tspan[i] == tspan[i]) ts.push_back(tspan[i]);   // a NaN passes neither filter below
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
  // tStart is reported once however often it occurs in tspan (ode.nim:485-487): mark the unused slots
  while (it < n_tspan) t_out[it++] = std::numeric_limits<double>::quiet_NaN();
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
  if (!c || !y0_local || !y_out_local || !tspan || !t_out || !n_y_out) return fail(c, B200RK_EINVAL, "null argument");
  b200rk_vec* y0 = nullptr;
  TRY(vec_alloc(c, n_global, &y0));
  int rc = B200RK_OK;
  const size_t bytes = y0->n_local * sizeof(double);
  cudaError_t e = cudaMemcpyAsync(y0->d, y0_local, bytes, cudaMemcpyHostToDevice, c->stream);
  if (e != cudaSuccess) rc = fail(c, B200RK_ECUDA, cudaGetErrorString(e));
  // The state reported at tStart is y0 itself (ode.nim:485-487). When no requested time lies before tStart it
  // is output 0, known before the solve starts: send it back on a second stream now, so the copy rides the
  // device-to-host engine while the solve runs instead of queueing behind it.
  const double t0 = options ? options->tStart : 0.0;
  bool has_zero = false, has_neg = false;
  for (size_t i = 0; i < n_tspan; ++i) { has_zero |= (tspan[i] == t0); has_neg |= (tspan[i] < t0); }
  bool early = false;
  // knob "tstart_copy" = 1: the host already holds those bytes — fill output slot 0 with a host-side copy of y0_local (a few
  // helper threads, concurrent with the solve) instead of a second device-to-host transfer. MEASURED (profiles/r02_scale_cfg2_n*.json,
  // e2e.tstart_slot): the device-to-host copy wins at every rank count — 5174 vs 5082 steps/s on 1 GPU, 11.7k vs 10.7k on 4,
  // 14.0k vs 12.8k on 8 (host memory is the shared resource there, and a memcpy moves the bytes through it twice) — so the
  // D2H on the second stream stays the default and the host-side copy is opt-in.
  std::vector<std::thread> host_copy;
  const bool host_side = c->tstart_copy == 1;
  if (rc == B200RK_OK && has_zero && !has_neg && bytes && host_side) {
    const size_t nthreads = 4, n = y0->n_local, chunk = (n + nthreads - 1) / nthreads;
    for (size_t i = 0; i < nthreads; ++i) {
      const size_t lo = std::min(n, i * chunk), hi = std::min(n, lo + chunk);
      if (hi > lo) host_copy.emplace_back([=] { std::memcpy(y_out_local + lo, y0_local + lo, (hi - lo) * sizeof(double)); });
    }
    early = true;
  }
  if (rc == B200RK_OK && has_zero && !has_neg && bytes && !host_side) {
    if (!c->copy_stream) {
      if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&c->copy_event, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        if (c->copy_stream) { cudaStreamDestroy(c->copy_stream); c->copy_stream = nullptr; }
      }
    }
    if (c->copy_stream && cudaEventRecord(c->copy_event, c->stream) == cudaSuccess &&
        cudaStreamWaitEvent(c->copy_stream, c->copy_event, 0) == cudaSuccess &&
        cudaMemcpyAsync(y_out_local, y0->d, bytes, cudaMemcpyDeviceToHost, c->copy_stream) == cudaSuccess)
      early = true;
    else
      cudaGetLastError();
  }
  std::vector<b200rk_vec*> ys(n_tspan, nullptr);
  size_t ny = 0;
  if (rc == B200RK_OK) rc = b200rk_solve(c, method, f, user, y0, tspan, n_tspan, options, t_out, ys.data(), &ny, stats);
  if (rc == B200RK_OK) {
    for (size_t i = early ? 1 : 0; i < ny; ++i) {
      e = cudaMemcpyAsync(y_out_local + i * y0->n_local, ys[i]->d, bytes, cudaMemcpyDeviceToHost, c->stream);
      if (e != cudaSuccess) { rc = fail(c, B200RK_ECUDA, cudaGetErrorString(e)); break; }
    }
    e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess && rc == B200RK_OK) rc = fail(c, B200RK_ECUDA, cudaGetErrorString(e));
    for (size_t i = 0; i < ny; ++i) vec_release(ys[i]);
    if (n_y_out) *n_y_out = ny;
  }
  for (auto& th : host_copy) th.join();
  if (early && host_copy.empty()) {  // y0 must stay untouched until its copy has left the device
    e = cudaStreamSynchronize(c->copy_stream);
    if (e != cudaSuccess && rc == B200RK_OK) rc = fail(c, B200RK_ECUDA, cudaGetErrorString(e));
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
  if (!s) return fail(nullptr, B200RK_EINVAL, "null solver");
  int rc = solver_advance(s, max_steps, steps_done);
  if (rc == B200RK_OK) CUDA_TRY(s->c, cudaStreamSynchronize(s->c->stream));
  if (finished) *finished = s->finished ? 1 : 0;
  return rc;
}
int b200rk_solver_state(const b200rk_solver* s, double* t, double* dt_next, double* last_error, const b200rk_vec** y) {
  if (!s) return fail(nullptr, B200RK_EINVAL, "null solver");
  if (t) *t = s->t;
  if (dt_next) *dt_next = s->dt;
  if (last_error) *last_error = s->error;
  if (y) *y = s->Y[s->cur];
  return B200RK_OK;
}
int b200rk_solver_stats(const b200rk_solver* s, b200rk_stats* out) {
  if (!s || !out) return fail(nullptr, B200RK_EINVAL, "null argument");
  solver_fill_stats(s, out, s->launches0, s->collectives0);
  return B200RK_OK;
}
int b200rk_solver_free(b200rk_solver* s) {
  if (s) { cudaStreamSynchronize(s->c->stream); delete s; }
  return B200RK_OK;
}

}  // extern "C"
