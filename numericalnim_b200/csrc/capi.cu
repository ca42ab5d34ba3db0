// capi.cu — the rest of the extern "C" surface: options and dispatch (ode.nim:78-104, 607-651), the Vector[T]
// operators (utils.nim:59-250), hermiteSpline, built-in right-hand sides, and the raw fused kernels.
#include "internal.hpp"

extern "C" {

// ---- options / dispatch -----------------------------------------------------------------------------
int b200rk_options_new(b200rk_options* out, double dt, double absTol, double relTol, double dtMax, double dtMin,
                       double scaleMax, double scaleMin, double tStart) {
  if (!out) return fail(nullptr, B200RK_EINVAL, "null argument");
  if (std::fabs(dtMax) < std::fabs(dtMin)) return fail(nullptr, B200RK_EINVAL, "dtMin must be less than dtMax");   // ode.nim:95-96
  if (std::fabs(scaleMax) < 1) return fail(nullptr, B200RK_EINVAL, "scaleMax must be bigger than 1");              // ode.nim:97-98
  if (1 < std::fabs(scaleMin)) return fail(nullptr, B200RK_EINVAL, "scaleMin must be smaller than 1");             // ode.nim:99-100
  *out = b200rk_options{std::fabs(dt), std::fabs(dtMax), std::fabs(dtMin), tStart, std::fabs(absTol),
                        std::fabs(relTol), std::fabs(scaleMax), std::fabs(scaleMin)};                              // ode.nim:101-102
  return B200RK_OK;
}
void b200rk_options_default(b200rk_options* out) { if (out) b200rk_options_new(out, 1e-4, 1e-4, 1e-4, 1e-2, 1e-4, 4.0, 0.1, 0.0); }

int b200rk_method_from_name(const char* name, int* method) {
  if (!method) return fail(nullptr, B200RK_EINVAL, "null argument");
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
  if (!c || !a || !b || !bhat) return fail(nullptr, B200RK_EINVAL, "null argument");
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

int b200rk_vec_fill(b200rk_vec* v, double value) {
  if (!v) return fail(nullptr, B200RK_EINVAL, "null vector");
  return launch_ewise(v->ctx, EW_FILL, v->d, nullptr, value, v->d, v->n_local, B200RK_K_OTHER);
}

#define EW_BINARY(NAME, OP)                                                              \
  int NAME(b200rk_vec* out, const b200rk_vec* a, const b200rk_vec* b) {                  \
    TRY(check_same(a ? a->ctx : nullptr, a, b));                                         \
    TRY(check_same(a->ctx, a, out));                                                     \
    return launch_ewise(a->ctx, OP, a->d, b->d, 0.0, out->d, a->n_local, B200RK_K_OTHER); \
  }
EW_BINARY(b200rk_vec_add, EW_ADD)
EW_BINARY(b200rk_vec_sub, EW_SUB)
EW_BINARY(b200rk_vec_hmul, EW_HMUL)
EW_BINARY(b200rk_vec_hdiv, EW_HDIV)
#undef EW_BINARY
int b200rk_vec_scale(b200rk_vec* out, double s, const b200rk_vec* a) {
  TRY(check_same(a ? a->ctx : nullptr, a, out));
  return launch_ewise(a->ctx, EW_SCALE, a->d, nullptr, s, out->d, a->n_local, B200RK_K_OTHER);
}
int b200rk_vec_div_scalar(b200rk_vec* out, const b200rk_vec* a, double s) {
  TRY(check_same(a ? a->ctx : nullptr, a, out));
  return launch_ewise(a->ctx, EW_DIV_SCALAR, a->d, nullptr, s, out->d, a->n_local, B200RK_K_OTHER);
}
int b200rk_vec_add_scalar(b200rk_vec* out, double s, const b200rk_vec* a) {
  TRY(check_same(a ? a->ctx : nullptr, a, out));
  return launch_ewise(a->ctx, EW_ADD_SCALAR, a->d, nullptr, s, out->d, a->n_local, B200RK_K_OTHER);
}
int b200rk_vec_neg(b200rk_vec* out, const b200rk_vec* a) {
  TRY(check_same(a ? a->ctx : nullptr, a, out));
  return launch_ewise(a->ctx, EW_NEG, a->d, nullptr, 0.0, out->d, a->n_local, B200RK_K_OTHER);
}
int b200rk_vec_abs(b200rk_vec* out, const b200rk_vec* a) {
  TRY(check_same(a ? a->ctx : nullptr, a, out));
  return launch_ewise(a->ctx, EW_ABS, a->d, nullptr, 0.0, out->d, a->n_local, B200RK_K_OTHER);
}
int b200rk_vec_sum(const b200rk_vec* a, double* out) {
  if (!a || !out) return fail(a ? a->ctx : nullptr, B200RK_EINVAL, "null argument");
  b200rk_ctx* c = a->ctx;
  TRY(launch_sum(c, a->d, a->n_local));
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
  if (!c || !fn || !user) return fail(c, B200RK_EINVAL, "null argument");
  if (kind < B200RK_RHS_SCALE || kind > B200RK_RHS_LORENZ96) return fail(c, B200RK_EINVAL, "unknown builtin rhs");
  if (kind == B200RK_RHS_DIAG_LINEAR && !lambda) return fail(c, B200RK_EINVAL, "diag-linear rhs needs lambda");
  if (lambda && lambda->ctx != c) return fail(c, B200RK_EINVAL, "diag-linear rhs: lambda belongs to another context");
  *user = new BuiltinRhs{c, kind, scalar, lambda};
  *fn = &builtin_rhs_fn;
  return B200RK_OK;
}
int b200rk_builtin_rhs_free(void* user) { delete static_cast<BuiltinRhs*>(user); return B200RK_OK; }

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
  if (!c) return fail(nullptr, B200RK_EINVAL, "null argument");
  TRY(check_same(c, y, k1)); TRY(check_same(c, y, k2)); TRY(check_same(c, y, k3)); TRY(check_same(c, y, k4)); TRY(check_same(c, y, out));
  return launch_rk4_final(c, y->d, k1->d, k2->d, k3->d, k4->d, dt / 6.0, out->d, y->n_local);
}

}  // extern "C"