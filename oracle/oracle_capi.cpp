// oracle_capi.cpp — flat C entry points over rk_oracle.hpp so the pytest suite (ctypes) and
// bench.py's cpu_baseline / --impl reference legs can drive the CPU oracle.
// TEST INFRASTRUCTURE ONLY: nothing under numericalnim_b200/ may load this library.
#include <chrono>
#include <thread>
#include <cstring>
#include <memory>

#include "rk_oracle.hpp"
#include "quad_oracle.hpp"

using namespace rk_oracle;

extern "C" {

struct oracle_options { double dt, dtMax, dtMin, tStart, absTol, relTol, scaleMax, scaleMin; };
struct oracle_stats { long steps, attempts, rejected, limiter_hits, rhs_evals, nan_guard; double seconds; };
struct oracle_step_record { double t, dt_used, error; int attempts; int _pad; };

// Host RHS callback for arbitrary closures (small cases only): writes dydt[0..n).
typedef void (*oracle_rhs_cb)(double t, const double* y, double* dydt, size_t n, void* user);

enum { ORACLE_RHS_SCALE = 0, ORACLE_RHS_DIAG_LINEAR = 1, ORACLE_RHS_LORENZ96 = 2, ORACLE_RHS_CALLBACK = 3 };

static thread_local std::string g_err;
const char* oracle_last_error() { return g_err.c_str(); }

static Options to_opt(const oracle_options* o) {
  return Options{o->dt, o->dtMax, o->dtMin, o->tStart, o->absTol, o->relTol, o->scaleMax, o->scaleMin};
}

// newODEoptions (ode.nim:78-102). Returns 0, or 1 (ValueError) with the message in oracle_last_error().
int oracle_new_options(oracle_options* out, double dt, double absTol, double relTol, double dtMax,
                       double dtMin, double scaleMax, double scaleMin, double tStart) {
  try {
    Options o = new_options(dt, absTol, relTol, dtMax, dtMin, scaleMax, scaleMin, tStart);
    *out = oracle_options{o.dt, o.dtMax, o.dtMin, o.tStart, o.absTol, o.relTol, o.scaleMax, o.scaleMin};
    return 0;
  } catch (const ValueError& e) { g_err = e.what(); return 1; }
}

int oracle_linspace(double x1, double x2, long n, double* out) {
  try {
    auto v = linspace(x1, x2, n);
    std::memcpy(out, v.data(), v.size() * sizeof(double));
    return 0;
  } catch (const ValueError& e) { g_err = e.what(); return 1; }
}

static OdeProc<Vector> make_rhs(int kind, const double* param, size_t n, double scalar,
                                oracle_rhs_cb cb, void* user) {
  switch (kind) {
    case ORACLE_RHS_SCALE: return rhs_scale_vec(scalar);
    case ORACLE_RHS_DIAG_LINEAR: return rhs_diag_linear(Vector(param, n));
    case ORACLE_RHS_LORENZ96: return rhs_lorenz96(scalar);
    case ORACLE_RHS_CALLBACK:
      return [cb, user](double t, const Vector& y, Context<Vector>* ctx) {
        if (ctx) ctx->rhs_evals++;
        std::vector<double> k(y.len());
        cb(t, y.components.data(), k.data(), y.len(), user);
        return Vector(k);
      };
  }
  throw ValueError("unknown rhs kind");
}

static void fill_stats(oracle_stats* s, const long steps, long attempts, long rejected, long lim,
                       long rhs, bool nan_guard, double seconds) {
  if (!s) return;
  s->steps = steps; s->attempts = attempts; s->rejected = rejected; s->limiter_hits = lim;
  s->rhs_evals = rhs; s->nan_guard = nan_guard ? 1 : 0; s->seconds = seconds;
}

// solveODE for T = Vector[float] (ode.nim:589-651). y_out receives *n_y_out vectors of length n, in
// the order of t_out (n_tspan sorted times). Note n_y_out can be < n_tspan.
int oracle_solve_vector(const char* integrator, int rhs_kind, const double* rhs_param, double rhs_scalar,
                        oracle_rhs_cb cb, void* user, size_t n, const double* y0, const double* tspan,
                        size_t n_tspan, const oracle_options* opt, double* t_out, double* y_out,
                        size_t y_out_cap, size_t* n_y_out, oracle_stats* stats,
                        oracle_step_record* trace, size_t trace_cap, size_t* n_trace, long max_steps) {
  try {
    OdeProc<Vector> f = make_rhs(rhs_kind, rhs_param, n, rhs_scalar, cb, user);
    Context<Vector> ctx;
    ctx.max_steps = max_steps;
    std::vector<Context<Vector>::StepRecord> tr;
    if (trace) ctx.trace = &tr;
    auto t0 = std::chrono::steady_clock::now();
    Solution<Vector> sol = solve_ode<Vector>(f, Vector(y0, n), std::vector<double>(tspan, tspan + n_tspan),
                                             to_opt(opt), &ctx, integrator);
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (size_t i = 0; i < sol.t.size(); ++i) t_out[i] = sol.t[i];
    *n_y_out = sol.y.size();
    for (size_t i = 0; i < sol.y.size() && i < y_out_cap; ++i)
      std::memcpy(y_out + i * n, sol.y[i].components.data(), n * sizeof(double));
    fill_stats(stats, ctx.steps, ctx.attempts, ctx.rejected, ctx.limiter_hits, ctx.rhs_evals,
               ctx.nan_guard_tripped, secs);
    if (trace) {
      size_t m = std::min(tr.size(), trace_cap);
      for (size_t i = 0; i < m; ++i) trace[i] = oracle_step_record{tr[i].t, tr[i].dt_used, tr[i].error, tr[i].attempts, 0};
      if (n_trace) *n_trace = tr.size();
    }
    return 0;
  } catch (const ValueError& e) { g_err = e.what(); return 1; }
}

// solveODE for T = float (scalar), rhs = c*y (tests/test_ode.nim:5) or a callback with n = 1.
int oracle_solve_scalar(const char* integrator, double rhs_scale, oracle_rhs_cb cb, void* user, double y0,
                        const double* tspan, size_t n_tspan, const oracle_options* opt, double* t_out,
                        double* y_out, size_t* n_y_out, oracle_stats* stats) {
  try {
    OdeProc<double> f;
    if (cb) f = [cb, user](double t, const double& y, Context<double>* ctx) {
      if (ctx) ctx->rhs_evals++;
      double k; cb(t, &y, &k, 1, user); return k; };
    else f = rhs_scale_scalar(rhs_scale);
    Context<double> ctx;
    auto t0 = std::chrono::steady_clock::now();
    Solution<double> sol = solve_ode<double>(f, y0, std::vector<double>(tspan, tspan + n_tspan), to_opt(opt), &ctx, integrator);
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (size_t i = 0; i < sol.t.size(); ++i) t_out[i] = sol.t[i];
    *n_y_out = sol.y.size();
    for (size_t i = 0; i < sol.y.size(); ++i) y_out[i] = sol.y[i];
    fill_stats(stats, ctx.steps, ctx.attempts, ctx.rejected, ctx.limiter_hits, ctx.rhs_evals, ctx.nan_guard_tripped, secs);
    return 0;
  } catch (const ValueError& e) { g_err = e.what(); return 1; }
}

// One IntegratorProc call (ode.nim:38) on Vectors: (yNew, newFSAL, dtUsed, error).
int oracle_step_vector(const char* integrator, int rhs_kind, const double* rhs_param, double rhs_scalar,
                       oracle_rhs_cb cb, void* user, size_t n, double t, const double* y, const double* fsal,
                       double dt, const oracle_options* opt, double* y_new, double* fsal_new,
                       double* dt_used, double* error, oracle_stats* stats) {
  try {
    OdeProc<Vector> f = make_rhs(rhs_kind, rhs_param, n, rhs_scalar, cb, user);
    std::string name = integrator;
    for (char& ch : name) ch = char(std::tolower((unsigned char)ch));
    for (const auto& m : method_table<Vector>()) {
      if (name != m.name) continue;
      Context<Vector> ctx;
      auto t0 = std::chrono::steady_clock::now();
      StepResult<Vector> r = m.step(f, t, Vector(y, n), Vector(fsal, n), dt, to_opt(opt), &ctx);
      double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      std::memcpy(y_new, r.y_new.components.data(), n * sizeof(double));
      std::memcpy(fsal_new, r.fsal.components.data(), n * sizeof(double));
      *dt_used = r.dt; *error = r.error;
      fill_stats(stats, 1, ctx.attempts, ctx.rejected, ctx.limiter_hits, ctx.rhs_evals, ctx.nan_guard_tripped, secs);
      return 0;
    }
    throw ValueError(std::string(integrator) + " is not a valid integrator");
  } catch (const ValueError& e) { g_err = e.what(); return 1; }
}

// ---- kernel-level restatements (the expressions each CUDA kernel replaces) --------------------
static const Pair* pair_by_name(const char* name) {
  std::string s = name;
  for (char& ch : s) ch = char(std::tolower((unsigned char)ch));
  if (s == "dopri54") return &dopri54_pair();
  if (s == "tsit54") return &tsit54_pair();
  if (s == "vern65") return &vern65_pair();
  return nullptr;
}

// y + c * (w0*k0 + w1*k1 + ... )  with the reference's association; m == 1 gives y + c*(w0*k0).
// This is the generic shape of ode.nim:294-299 (c = dt) — `k` holds m pointers to length-n arrays.
int oracle_weighted_stage(size_t n, int m, const double* w, double c, const double* y,
                          const double* const* k, double* out) {
  std::vector<Vector> kv; kv.reserve(m);
  std::vector<const Vector*> kp(m);
  for (int j = 0; j < m; ++j) { kv.emplace_back(k[j], n); }
  for (int j = 0; j < m; ++j) kp[j] = &kv[j];
  Vector r = Vector(y, n) + c * wsum<Vector>(w, kp.data(), m);
  std::memcpy(out, r.components.data(), n * sizeof(double));
  return 0;
}

// Stage-s input of a named pair (s = 2..stages), dense row including zero weights.
int oracle_pair_stage_input(const char* method, int s, size_t n, double dt, const double* y,
                            const double* const* k, double* out) {
  const Pair* p = pair_by_name(method);
  if (!p || s < 2 || s > p->stages) { g_err = "bad method/stage"; return 1; }
  std::vector<Vector> kv; kv.reserve(s - 1);
  std::vector<const Vector*> kp(s - 1);
  for (int j = 0; j < s - 1; ++j) kv.emplace_back(k[j], n);
  for (int j = 0; j < s - 1; ++j) kp[j] = &kv[j];
  Vector r = pair_stage_input<Vector>(*p, s, Vector(y, n), dt, kp.data());
  std::memcpy(out, r.components.data(), n * sizeof(double));
  return 0;
}

// Final combine + error norm of a named pair: yNew, error_y, sum(err1^2) (sequential), error.
int oracle_pair_finish(const char* method, size_t n, double dt, double absTol, double relTol,
                       const double* y, const double* const* k, double* y_new, double* error_y,
                       double* sumsq, double* error) {
  const Pair* p = pair_by_name(method);
  if (!p) { g_err = "bad method"; return 1; }
  std::vector<Vector> kv; kv.reserve(p->stages);
  std::vector<const Vector*> kp(p->stages);
  for (int j = 0; j < p->stages; ++j) kv.emplace_back(k[j], n);
  for (int j = 0; j < p->stages; ++j) kp[j] = &kv[j];
  Vector yv(y, n);
  Vector yN = pair_y_new<Vector>(*p, yv, dt, kp.data());
  Vector ey = pair_error_y<Vector>(*p, yv, yN, dt, kp.data());
  if (y_new) std::memcpy(y_new, yN.components.data(), n * sizeof(double));
  if (error_y) std::memcpy(error_y, ey.components.data(), n * sizeof(double));
  Vector totalTol = add_scalar(absTol, relTol * vabs(yN));
  Vector err1 = hdiv(ey, totalTol);
  Vector sq = hadamard(err1, err1);
  double S = vsum(sq);
  if (sumsq) *sumsq = S;
  if (error) *error = std::sqrt(1.0 / double(vsize(err1)) * S);
  return 0;
}

// RK4 pieces (ode.nim:185-188): stage inputs y + (cfac*dt)*k and the final combine.
int oracle_rk4_stage_input(size_t n, double cfac, double dt, const double* y, const double* k, double* out) {
  Vector r = (cfac == 1.0) ? Vector(y, n) + dt * Vector(k, n) : Vector(y, n) + cfac * dt * Vector(k, n);
  std::memcpy(out, r.components.data(), n * sizeof(double));
  return 0;
}
int oracle_rk4_combine(size_t n, double dt, const double* y, const double* k1, const double* k2,
                       const double* k3, const double* k4, double* out) {
  Vector r = Vector(y, n) + dt / 6.0 * (Vector(k1, n) + 2.0 * (Vector(k2, n) + Vector(k3, n)) + Vector(k4, n));
  std::memcpy(out, r.components.data(), n * sizeof(double));
  return 0;
}

// hermiteSpline on Vectors (utils.nim:273-279)
int oracle_hermite(size_t n, double x, double x1, double x2, const double* y1, const double* y2,
                   const double* dy1, const double* dy2, double* out) {
  Vector r = hermite_spline<Vector>(x, x1, x2, Vector(y1, n), Vector(y2, n), Vector(dy1, n), Vector(dy2, n));
  std::memcpy(out, r.components.data(), n * sizeof(double));
  return 0;
}

// Right-hand sides alone
int oracle_rhs_eval(int rhs_kind, const double* rhs_param, double rhs_scalar, size_t n, double t,
                    const double* y, double* out) {
  try {
    OdeProc<Vector> f = make_rhs(rhs_kind, rhs_param, n, rhs_scalar, nullptr, nullptr);
    Vector r = f(t, Vector(y, n), nullptr);
    std::memcpy(out, r.components.data(), n * sizeof(double));
    return 0;
  } catch (const ValueError& e) { g_err = e.what(); return 1; }
}

// Vector operator known-answers (tests/test_vector.nim): op 0:+ 1:- 2:*. 3:/. ; returns 1 on size mismatch.
int oracle_vector_binop(int op, const double* a, size_t na, const double* b, size_t nb, double* out) {
  try {
    Vector A(a, na), B(b, nb), r;
    switch (op) { case 0: r = A + B; break; case 1: r = A - B; break; case 2: r = hadamard(A, B); break;
                  case 3: r = hdiv(A, B); break; default: g_err = "bad op"; return 2; }
    std::memcpy(out, r.components.data(), r.len() * sizeof(double));
    return 0;
  } catch (const ValueError& e) { g_err = e.what(); return 1; }
}
// op 0: d*v  1: v/d  2: -v  3: abs v  4: d +. v
int oracle_vector_unop(int op, double d, const double* a, size_t n, double* out) {
  Vector A(a, n), r;
  switch (op) { case 0: r = d * A; break; case 1: r = A / d; break; case 2: r = -A; break;
                case 3: r = vabs(A); break; case 4: r = add_scalar(d, A); break; default: return 2; }
  std::memcpy(out, r.components.data(), n * sizeof(double));
  return 0;
}
double oracle_vector_sum(const double* a, size_t n) { return vsum(Vector(a, n)); }
double oracle_vector_norm(const double* a, size_t n, int p) { return vnorm(Vector(a, n), p); }
double oracle_vector_dot(const double* a, const double* b, size_t n) { return dot(Vector(a, n), Vector(b, n)); }
int oracle_is_close_vec(const double* a, const double* b, size_t n, double tol) { return is_close(Vector(a, n), Vector(b, n), tol) ? 1 : 0; }
int oracle_is_close_scalar(double a, double b, double tol) { return is_close(a, b, tol) ? 1 : 0; }

// Tableau export so tests can check order conditions / cross-check the product's own copy.
int oracle_pair_tableau(const char* method, int* stages, int* order, int* n_b, int* n_bhat, int* err_direct,
                        double* c /*10*/, double* a /*10*9*/, double* b /*9*/, double* bhat /*9*/) {
  const Pair* p = pair_by_name(method);
  if (!p) return 1;
  *stages = p->stages; *order = p->order; *n_b = p->n_b; *n_bhat = p->n_bhat; *err_direct = p->err_is_direct;
  std::memcpy(c, p->c, sizeof(p->c)); std::memcpy(a, p->a, sizeof(p->a));
  std::memcpy(b, p->b, sizeof(p->b)); std::memcpy(bhat, p->bhat, sizeof(p->bhat));
  return 0;
}

// ---- Hermite interpolation of a data set + cumulative quadrature (quad_oracle.hpp) ---------------------------
// Sequences of T are flat row-major arrays: item k occupies [k*n, (k+1)*n). scalar != 0 runs the T = float
// instantiation (n must be 1) — the reference's own tests use it. Return: 0 ok, 1 ValueError, 3 assert / bounds defect.
typedef void (*oracle_fn_cb)(double t, double* out, size_t n, void* user);
}  // extern "C"

template <class Fn>
static int guarded(Fn&& fn) {
  try { fn(); return 0; }
  catch (const ValueError& e) { g_err = e.what(); return 1; }
  catch (const AssertionDefect& e) { g_err = e.what(); return 3; }
  catch (const IndexDefect& e) { g_err = e.what(); return 3; }
}
static std::vector<Vector> rows(const double* a, size_t m, size_t n) {
  std::vector<Vector> r;
  for (size_t k = 0; k < m; ++k) r.emplace_back(a + k * n, n);
  return r;
}
static std::vector<double> scalars(const double* a, size_t m) { return std::vector<double>(a, a + m); }
static void store(const std::vector<Vector>& r, size_t n, double* out, size_t* n_out) {
  for (size_t k = 0; k < r.size(); ++k) std::memcpy(out + k * n, r[k].components.data(), n * sizeof(double));
  *n_out = r.size();
}
static void store(const std::vector<double>& r, size_t, double* out, size_t* n_out) {
  for (size_t k = 0; k < r.size(); ++k) out[k] = r[k];
  *n_out = r.size();
}
static FnOfT<Vector> vec_fn(oracle_fn_cb cb, void* user, size_t n, long* evals) {
  return [=](double t) { std::vector<double> v(n); cb(t, v.data(), n, user); if (evals) ++*evals; return Vector(v); };
}
static FnOfT<double> scalar_fn(oracle_fn_cb cb, void* user, long* evals) {
  return [=](double t) { double v = 0.0; cb(t, &v, 1, user); if (evals) ++*evals; return v; };
}

extern "C" {

int oracle_hermite_interpolate(int scalar, size_t n, const double* x, size_t nx, const double* t, size_t nt, const double* y,
                               const double* dy, double* out, size_t* n_out) {
  return guarded([&] {
    if (scalar) store(hermite_interpolate<double>(scalars(x, nx), scalars(t, nt), scalars(y, nt), scalars(dy, nt)), 1, out, n_out);
    else store(hermite_interpolate<Vector>(scalars(x, nx), scalars(t, nt), rows(y, nt, n), rows(dy, nt, n)), n, out, n_out);
  });
}
int oracle_cumtrapz(int scalar, size_t n, const double* Y, const double* X, size_t m, double* out, size_t* n_out) {
  return guarded([&] {
    if (scalar) store(cumtrapz<double>(scalars(Y, m), scalars(X, m)), 1, out, n_out);
    else store(cumtrapz<Vector>(rows(Y, m, n), scalars(X, m)), n, out, n_out);
  });
}
int oracle_cumsimpson(int scalar, size_t n, const double* Y, const double* X, size_t m, double* out, size_t* n_out) {
  return guarded([&] {
    if (scalar) store(cumsimpson<double>(scalars(Y, m), scalars(X, m)), 1, out, n_out);
    else store(cumsimpson<Vector>(rows(Y, m, n), scalars(X, m)), n, out, n_out);
  });
}
int oracle_cumtrapz_fn(int scalar, size_t n, oracle_fn_cb cb, void* user, const double* X, size_t m, double dx, double* out,
                       size_t* n_out, long* evals) {
  return guarded([&] {
    if (scalar) store(cumtrapz_fn<double>(scalar_fn(cb, user, evals), scalars(X, m), dx), 1, out, n_out);
    else store(cumtrapz_fn<Vector>(vec_fn(cb, user, n, evals), scalars(X, m), dx), n, out, n_out);
  });
}
int oracle_cumsimpson_fn(int scalar, size_t n, oracle_fn_cb cb, void* user, const double* X, size_t m, double dx, double* out,
                         size_t* n_out, long* evals) {
  return guarded([&] {
    if (scalar) store(cumsimpson_fn<double>(scalar_fn(cb, user, evals), scalars(X, m), dx), 1, out, n_out);
    else store(cumsimpson_fn<Vector>(vec_fn(cb, user, n, evals), scalars(X, m), dx), n, out, n_out);
  });
}
// the coefficient triples alone (host-side scalars of the product are checked against these)
void oracle_simpson_weights(int tail, double h1, double h2, double* alpha, double* beta, double* eta) {
  const SimpsonWeights w = tail ? simpson_tail_weights(h1, h2) : simpson_pair_weights(h1, h2);
  *alpha = w.alpha; *beta = w.beta; *eta = w.eta;
}

// ---- courtesy baseline (NOT the reference's behaviour) ---------------------------------------------------------------
// What a CPU can do for the same element-local IVP y' = -(lambda .* y) when the attempt is fused into one pass over
// the state and spread over all host cores (std::thread, contiguous chunks): per element the very same operations in
// the same order as pair_step (so element-wise results are bit-identical to the oracle's), the error norm summed per
// thread and combined in thread order. Reported
// next to the single-threaded port in bench.py's cpu_baseline so the GPU/CPU ratio can also be read against a
// well-written CPU code; numericalnim itself is single-threaded and allocates a Vector per operator.
struct oracle_fused_stats { long steps, attempts, rejected, limiter_hits; double seconds; int threads; int _pad; };

int oracle_fused_mt_solve_diag(const char* method, const double* lam, double* y_io, size_t n, double t_end,
                               const oracle_options* opt, long max_steps, int threads, oracle_fused_stats* out) {
  const Pair* pp = pair_by_name(method);
  if (!pp) { g_err = "fused baseline: dopri54 / tsit54 / vern65 only"; return 1; }
  const Pair p = *pp;
  const Options o = to_opt(opt);
  const int S = p.stages;
  int used = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (used < 1) used = 1;
  if ((size_t)used > n) used = (int)std::max<size_t>(n, 1);
  std::vector<double> Y[2] = {std::vector<double>(y_io, y_io + n), std::vector<double>(n)}, F[2] = {std::vector<double>(n), std::vector<double>(n)};
  for (size_t i = 0; i < n; ++i) F[0][i] = -(lam[i] * Y[0][i]);                       // FSAL = f(t0, y0)
  int cur = 0;
  double t = o.tStart, dt = std::sqrt(o.dtMax * o.dtMin), error = 0.0;                // ode.nim:491-493
  oracle_fused_stats st{0, 0, 0, 0, 0.0, used, 0};
  const auto t0 = std::chrono::steady_clock::now();
  while (t < t_end && (max_steps <= 0 || st.steps < max_steps)) {
    dt = nim_min(dt, t_end - t);                                                      // ode.nim:525
    int limit = 0;
    while (true) {                                                                    // ode.nim:57-76
      const double* y = Y[cur].data(); const double* k1 = F[cur].data();
      double* yn = Y[1 - cur].data(); double* ks = F[1 - cur].data();
      const double h = dt;
      std::vector<double> partial(used, 0.0);
      auto work = [&](int w) {
        const size_t lo = n * (size_t)w / (size_t)used, hi = n * (size_t)(w + 1) / (size_t)used;
        double acc_sum = 0.0;
        for (size_t i = lo; i < hi; ++i) {
          double k[9];
          k[0] = k1[i];
          double in = 0.0;
          for (int s = 2; s <= S; ++s) {
            double acc = p.a[s][0] * k[0];
            for (int j = 1; j < s - 1; ++j) acc = acc + p.a[s][j] * k[j];
            in = y[i] + h * acc;
            k[s - 1] = -(lam[i] * in);
          }
          double accb = p.b[0] * k[0];
          for (int j = 1; j < p.n_b; ++j) accb = accb + p.b[j] * k[j];
          const double ynew = y[i] + h * accb;
          double acch = p.bhat[0] * k[0];
          for (int j = 1; j < p.n_bhat; ++j) acch = acch + p.bhat[j] * k[j];
          const double e = p.err_is_direct ? h * acch : ynew - (y[i] + h * acch);
          const double r = e / (o.absTol + o.relTol * std::fabs(ynew));
          acc_sum += r * r;
          yn[i] = ynew; ks[i] = k[S - 1];
        }
        partial[w] = acc_sum;
      };
      std::vector<std::thread> pool;
      for (int w = 1; w < used; ++w) pool.emplace_back(work, w);
      work(0);
      for (auto& th : pool) th.join();
      double sum = 0.0;
      for (int w = 0; w < used; ++w) sum += partial[w];
      st.attempts++;
      error = std::sqrt(1.0 / double(n) * sum);
      if (error <= 1) break;
      if (std::isnan(error)) { g_err = "error norm is NaN"; return 2; }
      st.rejected++;
      dt = dt * nim_min(4, nim_max(0.125, 0.9 * std::pow(1.0 / error, 1.0 / double(p.order))));
      if (std::fabs(dt) < o.dtMin) { dt = o.dtMin; limit += 1; st.limiter_hits++; }
      else if (o.dtMax < std::fabs(dt)) dt = o.dtMax;
      if (!(limit < 2)) break;
    }
    cur = 1 - cur;
    t += dt;
    st.steps++;
    if (error == 0.0) dt *= 5;                                                        // ode.nim:533-541
    else dt = dt * nim_min(4, nim_max(0.125, 0.9 * std::pow(1.0 / error, 1.0 / double(p.order))));
    if (dt < o.dtMin) dt = o.dtMin;
    else if (o.dtMax < dt) dt = o.dtMax;
  }
  st.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::memcpy(y_io, Y[cur].data(), n * sizeof(double));
  *out = st;
  return 0;
}

}  // extern "C"
