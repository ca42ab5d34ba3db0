// numericalnim_b200.hpp — header-only C++17 host mirror of numericalnim's ODE interface over the C-ABI
// (include/b200rk.h). The reference is compiled Nim whose extension point is generics; there is no Nim
// toolchain in this image, so this header is the compiled-language host side: the same names, argument
// meaning and error behaviour as `src/numericalnim/ode.nim` / `utils.nim`, for a device-resident vector.
//
//   auto f = [](double t, const GpuVector& y, NumContext& ctx) { return -0.1 * y; };   // tests/test_ode.nim:6
//   auto [ts, ys] = solveODE(f, newVector({1.0, 1.0, 1.0}), linspace(-10.0, 10.0, 100), newODEoptions(), nullptr, "tsit54");
//
//   ODEoptions / newODEoptions      ode.nim:26-34, 78-102   (ValueError on bad dtMax/dtMin, scaleMax, scaleMin)
//   ODEProc                         ode.nim:36              (t, y, ctx) -> dy
//   solveODE                        ode.nim:589-651         (ValueError on an unknown integrator)
//   fixedODE / adaptiveODE / allODE ode.nim:40-42
//   GpuVector (+ - * / abs ...)     utils.nim:14-271        value semantics: every operator returns a fresh vector
//   NumContext                      common/commonTypes.nim:3-39
//   linspace, hermiteSpline         utils.nim:498-507, 273-279
#pragma once
#include <cmath>
#include <exception>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "b200rk.h"

namespace numericalnim {

struct ValueError : std::invalid_argument {  // what the reference raises (Nim's ValueError)
  using std::invalid_argument::invalid_argument;
};
struct DeviceError : std::runtime_error {
  int code;
  DeviceError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline void check(int rc, const b200rk_ctx* ctx = nullptr) {
  if (rc == B200RK_OK) return;
  const char* m = b200rk_last_error(ctx);
  const std::string msg = m ? m : "";
  if (rc == B200RK_EINVAL) throw ValueError(msg);
  throw DeviceError(rc, "b200rk error " + std::to_string(rc) + ": " + msg);
}

// ---- device context (one GPU per process) --------------------------------------------------------------
class Device {
 public:
  explicit Device(int device = 0) { check(b200rk_init(&ctx_, device)); }
  Device(int device, int rank, int world, const void* nccl_id128) { check(b200rk_init_distributed(&ctx_, device, rank, world, nccl_id128)); }
  ~Device() { b200rk_destroy(ctx_); }
  Device(const Device&) = delete;
  Device& operator=(const Device&) = delete;
  b200rk_ctx* handle() const { return ctx_; }
  void set(const char* key, long long v) { check(b200rk_set(ctx_, key, v), ctx_); }
  static std::shared_ptr<Device>& defaultDevice() {
    static std::shared_ptr<Device> d;
    if (!d) d = std::make_shared<Device>(0);
    return d;
  }

 private:
  b200rk_ctx* ctx_ = nullptr;
};

// ---- GpuVector: stands where Vector[float] stands in the reference --------------------------------------
class GpuVector {
 public:
  GpuVector() = default;
  GpuVector(std::shared_ptr<Device> dev, size_t n) : dev_(std::move(dev)) { check(b200rk_vec_new(dev_->handle(), n, &h_), dev_->handle()); }
  GpuVector(const GpuVector& o) : dev_(o.dev_) {  // copy == clone (utils.nim:269)
    if (o.h_) { check(b200rk_vec_new(dev_->handle(), o.size(), &h_), dev_->handle()); check(b200rk_vec_copy(h_, o.h_), dev_->handle()); }
  }
  GpuVector(GpuVector&& o) noexcept : dev_(std::move(o.dev_)), h_(o.h_), borrowed_(o.borrowed_) { o.h_ = nullptr; }
  GpuVector& operator=(GpuVector o) noexcept { swap(o); return *this; }
  ~GpuVector() { if (h_ && !borrowed_) b200rk_vec_free(h_); }
  void swap(GpuVector& o) noexcept { std::swap(dev_, o.dev_); std::swap(h_, o.h_); std::swap(borrowed_, o.borrowed_); }

  static GpuVector borrow(std::shared_ptr<Device> dev, b200rk_vec* h) { GpuVector v; v.dev_ = std::move(dev); v.h_ = h; v.borrowed_ = true; return v; }
  static GpuVector adopt(std::shared_ptr<Device> dev, b200rk_vec* h) { GpuVector v; v.dev_ = std::move(dev); v.h_ = h; return v; }

  size_t size() const { return h_ ? b200rk_vec_len(h_) : 0; }  // utils.nim:57
  size_t len() const { return size(); }
  b200rk_vec* handle() const { return h_; }
  const std::shared_ptr<Device>& device() const { return dev_; }
  std::vector<double> components() const {  // `@v` (utils.nim:43): host copy
    std::vector<double> out(size());
    if (!out.empty()) check(b200rk_vec_download(h_, out.data()), dev_->handle());
    return out;
  }
  double sum() const { double s = 0; check(b200rk_vec_sum(h_, &s), dev_->handle()); return s; }  // utils.nim:243-250
  GpuVector clone() const { return GpuVector(*this); }

 private:
  std::shared_ptr<Device> dev_;
  b200rk_vec* h_ = nullptr;
  bool borrowed_ = false;
};

inline GpuVector newVector(const std::vector<double>& components, std::shared_ptr<Device> dev = Device::defaultDevice()) {  // utils.nim:19-20
  GpuVector v(std::move(dev), components.size());
  if (!components.empty()) check(b200rk_vec_upload(v.handle(), components.data()), v.device()->handle());
  return v;
}

namespace detail {
// b200rk_solve marks the unused tail of t_out with NaN when tStart is repeated in tspan (ode.nim:485-487).
inline void trim_times(std::vector<double>& t) {
  while (!t.empty() && t.back() != t.back()) t.pop_back();
}
template <class F>
inline GpuVector binary(const GpuVector& a, const GpuVector& b, F fn) {
  GpuVector out(a.device(), a.size());
  check(fn(out.handle(), a.handle(), b.handle()), a.device()->handle());  // size mismatch -> ValueError (utils.nim:22-26)
  return out;
}
}  // namespace detail
inline GpuVector operator+(const GpuVector& a, const GpuVector& b) { return detail::binary(a, b, b200rk_vec_add); }   // utils.nim:59-64
inline GpuVector operator-(const GpuVector& a, const GpuVector& b) { return detail::binary(a, b, b200rk_vec_sub); }   // utils.nim:113-118
inline GpuVector hadamard(const GpuVector& a, const GpuVector& b) { return detail::binary(a, b, b200rk_vec_hmul); }   // `*.` utils.nim:186-191
inline GpuVector hdiv(const GpuVector& a, const GpuVector& b) { return detail::binary(a, b, b200rk_vec_hdiv); }       // `/.` utils.nim:192-197
inline GpuVector operator*(double d, const GpuVector& a) {  // utils.nim:176-180
  GpuVector out(a.device(), a.size());
  check(b200rk_vec_scale(out.handle(), d, a.handle()), a.device()->handle());
  return out;
}
inline GpuVector operator*(const GpuVector& a, double d) { return d * a; }  // utils.nim:171-175
inline GpuVector operator/(const GpuVector& a, double d) {                  // utils.nim:166-170
  GpuVector out(a.device(), a.size());
  check(b200rk_vec_div_scalar(out.handle(), a.handle(), d), a.device()->handle());
  return out;
}
inline GpuVector operator-(const GpuVector& a) {  // utils.nim:214-218
  GpuVector out(a.device(), a.size());
  check(b200rk_vec_neg(out.handle(), a.handle()), a.device()->handle());
  return out;
}
inline GpuVector abs(const GpuVector& a) {  // utils.nim:219-223
  GpuVector out(a.device(), a.size());
  check(b200rk_vec_abs(out.handle(), a.handle()), a.device()->handle());
  return out;
}
inline GpuVector addScalar(double d, const GpuVector& a) {  // `+.` utils.nim:78-82
  GpuVector out(a.device(), a.size());
  check(b200rk_vec_add_scalar(out.handle(), d, a.handle()), a.device()->handle());
  return out;
}
inline double sum(const GpuVector& a) { return a.sum(); }
inline GpuVector hermiteSpline(double x, double x1, double x2, const GpuVector& y1, const GpuVector& y2, const GpuVector& dy1,
                               const GpuVector& dy2) {  // utils.nim:273-279
  GpuVector out(y1.device(), y1.size());
  check(b200rk_hermite(out.handle(), x, x1, x2, y1.handle(), y2.handle(), dy1.handle(), dy2.handle()), y1.device()->handle());
  return out;
}

// ---- NumContext (commonTypes.nim:3-39) -------------------------------------------------------------------
struct NumContext {
  std::map<std::string, double> fValues;
  std::map<std::string, GpuVector> tValues;
  GpuVector& operator[](const std::string& key) { return tValues[key]; }
  double getF(const std::string& key) const { return fValues.at(key); }
  void setF(const std::string& key, double v) { fValues[key] = v; }
};
inline std::shared_ptr<NumContext> newNumContext() { return std::make_shared<NumContext>(); }

// ---- options / helpers ----------------------------------------------------------------------------------
using ODEoptions = b200rk_options;  // field for field ode.nim:26-34
inline ODEoptions newODEoptions(double dt = 1e-4, double absTol = 1e-4, double relTol = 1e-4, double dtMax = 1e-2, double dtMin = 1e-4,
                                double scaleMax = 4.0, double scaleMin = 0.1, double tStart = 0.0) {  // ode.nim:78-102
  ODEoptions o;
  check(b200rk_options_new(&o, dt, absTol, relTol, dtMax, dtMin, scaleMax, scaleMin, tStart));
  return o;
}
inline std::vector<double> linspace(double x1, double x2, long N) {  // utils.nim:498-507
  if (N <= 0) throw ValueError("Number of samples " + std::to_string(N) + " must be greater then 0");
  std::vector<double> r;
  const double dx = (x2 - x1) / double(N - 1);
  r.push_back(x1);
  for (long i = 1; i <= N - 2; ++i) r.push_back(x1 + dx * double(i));
  r.push_back(x2);
  return r;
}
inline const std::vector<std::string>& fixedODE() { static const std::vector<std::string> v = {"heun2", "ralston2", "kutta3", "heun3", "ralston3", "ssprk3", "ralston4", "kutta4", "rk4"}; return v; }
inline const std::vector<std::string>& adaptiveODE() { static const std::vector<std::string> v = {"rk21", "bs32", "dopri54", "tsit54", "vern65"}; return v; }
inline std::vector<std::string> allODE() { auto v = fixedODE(); v.insert(v.end(), adaptiveODE().begin(), adaptiveODE().end()); return v; }

// ---- ODEProc / solveODE -----------------------------------------------------------------------------------
using ODEProc = std::function<GpuVector(double, const GpuVector&, NumContext&)>;  // ode.nim:36

namespace detail {
struct RhsEnv {
  const ODEProc* f;
  NumContext* ctx;
  std::shared_ptr<Device> dev;
  std::exception_ptr err;
};
// C callback: runs the user's closure on borrowed handles and copies its result into dydt. Exceptions never
// cross the ABI; they are re-thrown by solveODE after the C call returns.
inline int trampoline(double t, const b200rk_vec* y, b200rk_vec* dydt, void* user) {
  auto* env = static_cast<RhsEnv*>(user);
  try {
    GpuVector yv = GpuVector::borrow(env->dev, const_cast<b200rk_vec*>(y));
    GpuVector r = (*env->f)(t, yv, *env->ctx);
    if (r.handle() != dydt) check(b200rk_vec_copy(dydt, r.handle()), env->dev->handle());
    return 0;
  } catch (...) {
    env->err = std::current_exception();
    return 1;
  }
}
}  // namespace detail

struct Solution {
  std::vector<double> t;
  std::vector<GpuVector> y;
  b200rk_stats stats{};
};

// solveODE (ode.nim:589-651): sorted tspan out, one state per returned time (see SURVEY.md A.4 for the
// reference's quirks, reproduced). Unknown integrator -> ValueError before any work.
inline Solution solveODE(const ODEProc& f, const GpuVector& y0, const std::vector<double>& tspan, const ODEoptions& options = newODEoptions(),
                         std::shared_ptr<NumContext> ctx = nullptr, const std::string& integrator = "dopri54") {
  int method = 0;
  check(b200rk_method_from_name(integrator.c_str(), &method));  // ode.nim:650-651
  if (!ctx) ctx = newNumContext();                               // ode.nim:604-606
  detail::RhsEnv env{&f, ctx.get(), y0.device(), nullptr};
  Solution sol;
  sol.t.resize(tspan.size());
  std::vector<b200rk_vec*> slots(tspan.size() ? tspan.size() : 1, nullptr);
  size_t n_out = 0;
  const int rc = b200rk_solve(y0.device()->handle(), method, &detail::trampoline, &env, y0.handle(), tspan.data(), tspan.size(), &options,
                              sol.t.data(), slots.data(), &n_out, &sol.stats);
  if (env.err) std::rethrow_exception(env.err);
  check(rc, y0.device()->handle());
  for (size_t i = 0; i < n_out; ++i) sol.y.push_back(GpuVector::adopt(y0.device(), slots[i]));
  detail::trim_times(sol.t);
  return sol;
}

// Built-in device-side right-hand sides (fused paths): returns the C callback + its state.
class BuiltinRhs {
 public:
  BuiltinRhs(std::shared_ptr<Device> dev, int kind, double scalar, const GpuVector* lambda) : dev_(std::move(dev)) {
    if (lambda) lambda_ = std::make_unique<GpuVector>(*lambda);
    check(b200rk_builtin_rhs_new(dev_->handle(), kind, scalar, lambda_ ? lambda_->handle() : nullptr, &fn_, &user_), dev_->handle());
  }
  ~BuiltinRhs() { b200rk_builtin_rhs_free(user_); }
  BuiltinRhs(const BuiltinRhs&) = delete;
  b200rk_rhs_fn fn() const { return fn_; }
  void* user() const { return user_; }

 private:
  std::shared_ptr<Device> dev_;
  std::unique_ptr<GpuVector> lambda_;
  b200rk_rhs_fn fn_ = nullptr;
  void* user_ = nullptr;
};
// Element-local right-hand side given as source (b200rk_jit_rhs_new): dydt[i] = expr(t, y[i], p0[i].., c0..),
// compiled at run time into the fused kernels. E.g. JitRhs(dev, "c0*y*(1.0 - y/p0)", {&K}, {r}).
class JitRhs {
 public:
  JitRhs(std::shared_ptr<Device> dev, const std::string& expr, const std::vector<const GpuVector*>& vecs = {},
         const std::vector<double>& scalars = {}) : dev_(std::move(dev)) {
    std::vector<const b200rk_vec*> hs;
    for (const GpuVector* v : vecs) { keep_.push_back(std::make_unique<GpuVector>(*v)); hs.push_back(keep_.back()->handle()); }
    check(b200rk_jit_rhs_new(dev_->handle(), expr.c_str(), (int)hs.size(), hs.empty() ? nullptr : hs.data(), (int)scalars.size(),
                             scalars.empty() ? nullptr : scalars.data(), &fn_, &user_), dev_->handle());
  }
  ~JitRhs() { b200rk_jit_rhs_free(user_); }
  JitRhs(const JitRhs&) = delete;
  void setScalars(const std::vector<double>& scalars) {
    check(b200rk_jit_rhs_set_scalars(user_, (int)scalars.size(), scalars.empty() ? nullptr : scalars.data()), dev_->handle());
  }
  b200rk_rhs_fn fn() const { return fn_; }
  void* user() const { return user_; }

 private:
  std::shared_ptr<Device> dev_;
  std::vector<std::unique_ptr<GpuVector>> keep_;
  b200rk_rhs_fn fn_ = nullptr;
  void* user_ = nullptr;
};

// Stencil right-hand side given as source (b200rk_jit_stencil_rhs_new): dydt[i] = expr(t, Y(-rl)..Y(+rr), p0[i].., c0..) with
// Y(d) = y[(i + d) mod N]. E.g. JitStencilRhs(dev, "((Y(1) - Y(-2)) * Y(-1) - Y(0)) + c0", 2, 1, {}, {8.0}) is Lorenz-96: compiled
// into a dydt kernel for every method and into the one-kernel attempt over overlapped tiles for DOPRI54 / Tsit54 / Vern65.
class JitStencilRhs {
 public:
  JitStencilRhs(std::shared_ptr<Device> dev, const std::string& expr, int radiusLeft, int radiusRight, const std::vector<const GpuVector*>& vecs = {},
                const std::vector<double>& scalars = {}) : dev_(std::move(dev)) {
    std::vector<const b200rk_vec*> hs;
    for (const GpuVector* v : vecs) { keep_.push_back(std::make_unique<GpuVector>(*v)); hs.push_back(keep_.back()->handle()); }
    check(b200rk_jit_stencil_rhs_new(dev_->handle(), expr.c_str(), radiusLeft, radiusRight, (int)hs.size(), hs.empty() ? nullptr : hs.data(),
                                     (int)scalars.size(), scalars.empty() ? nullptr : scalars.data(), &fn_, &user_), dev_->handle());
  }
  ~JitStencilRhs() { b200rk_jit_rhs_free(user_); }
  JitStencilRhs(const JitStencilRhs&) = delete;
  void setScalars(const std::vector<double>& scalars) {
    check(b200rk_jit_rhs_set_scalars(user_, (int)scalars.size(), scalars.empty() ? nullptr : scalars.data()), dev_->handle());
  }
  b200rk_rhs_fn fn() const { return fn_; }
  void* user() const { return user_; }

 private:
  std::shared_ptr<Device> dev_;
  std::vector<std::unique_ptr<GpuVector>> keep_;
  b200rk_rhs_fn fn_ = nullptr;
  void* user_ = nullptr;
};

template <class DeviceRhs, class = decltype(std::declval<const DeviceRhs&>().fn()), class = decltype(std::declval<const DeviceRhs&>().user())>
inline Solution solveODE(const DeviceRhs& f, const GpuVector& y0, const std::vector<double>& tspan, const ODEoptions& options = newODEoptions(),
                         const std::string& integrator = "dopri54") {
  int method = 0;
  check(b200rk_method_from_name(integrator.c_str(), &method));
  Solution sol;
  sol.t.resize(tspan.size());
  std::vector<b200rk_vec*> slots(tspan.size() ? tspan.size() : 1, nullptr);
  size_t n_out = 0;
  check(b200rk_solve(y0.device()->handle(), method, f.fn(), f.user(), y0.handle(), tspan.data(), tspan.size(), &options, sol.t.data(),
                     slots.data(), &n_out, &sol.stats), y0.device()->handle());
  for (size_t i = 0; i < n_out; ++i) sol.y.push_back(GpuVector::adopt(y0.device(), slots[i]));
  detail::trim_times(sol.t);
  return sol;
}

// ---- consumers of a trajectory: hermiteInterpolate (utils.nim:282-312), cumtrapz / cumsimpson (integrate.nim) ----
namespace detail {
inline std::vector<const b200rk_vec*> handles(const std::vector<GpuVector>& v) {
  std::vector<const b200rk_vec*> h;
  for (const GpuVector& x : v) h.push_back(x.handle());
  return h;
}
inline std::vector<GpuVector> adopt_all(const std::shared_ptr<Device>& dev, const std::vector<b200rk_vec*>& slots, size_t n) {
  std::vector<GpuVector> r;
  for (size_t i = 0; i < n; ++i) r.push_back(GpuVector::adopt(dev, slots[i]));
  return r;
}
}  // namespace detail

inline std::vector<GpuVector> hermiteInterpolate(const std::vector<double>& x, const std::vector<double>& t, const std::vector<GpuVector>& y,
                                                 const std::vector<GpuVector>& dy) {
  if (y.empty() || y.size() != t.size() || dy.size() != t.size()) throw ValueError("t, y and dy must have the same non-zero length");
  std::vector<b200rk_vec*> slots(x.size() ? x.size() : 1, nullptr);
  size_t n = 0;
  const auto hy = detail::handles(y), hdy = detail::handles(dy);
  check(b200rk_hermite_interpolate(y[0].device()->handle(), x.data(), x.size(), t.data(), t.size(), hy.data(), hdy.data(), slots.data(), &n),
        y[0].device()->handle());
  return detail::adopt_all(y[0].device(), slots, n);
}
inline std::vector<GpuVector> cumtrapz(const std::vector<GpuVector>& Y, const std::vector<double>& X) {  // integrate.nim:119-135
  if (Y.empty() || Y.size() != X.size()) throw ValueError("X and Y must have the same non-zero length");
  std::vector<b200rk_vec*> slots(X.size(), nullptr);
  size_t n = 0;
  const auto h = detail::handles(Y);
  check(b200rk_cumtrapz(Y[0].device()->handle(), h.data(), X.data(), X.size(), slots.data(), &n), Y[0].device()->handle());
  return detail::adopt_all(Y[0].device(), slots, n);
}
inline std::vector<GpuVector> cumsimpson(const std::vector<GpuVector>& Y, const std::vector<double>& X) {  // integrate.nim:330-378
  if (Y.empty() || Y.size() != X.size()) throw ValueError("X and Y must have the same non-zero length");
  std::vector<b200rk_vec*> slots(X.size(), nullptr);
  size_t n = 0;
  const auto h = detail::handles(Y);
  check(b200rk_cumsimpson(Y[0].device()->handle(), h.data(), X.data(), X.size(), slots.data(), &n), Y[0].device()->handle());
  return detail::adopt_all(Y[0].device(), slots, n);
}

// NumContextProc[T, float]: the integrand at one point
using NumContextProc = std::function<GpuVector(double, NumContext&)>;
namespace detail {
struct FnEnv {
  const NumContextProc* f;
  NumContext* ctx;
  std::shared_ptr<Device> dev;
  std::exception_ptr err;
};
inline int fn_trampoline(double t, b200rk_vec* out, void* user) {
  auto* env = static_cast<FnEnv*>(user);
  try {
    GpuVector r = (*env->f)(t, *env->ctx);
    return b200rk_vec_copy(out, r.handle());
  } catch (...) {
    env->err = std::current_exception();
    return 1;
  }
}
using CumFn = int (*)(b200rk_ctx*, b200rk_fn_of_t, void*, size_t, const double*, size_t, double, b200rk_vec**, size_t*);
inline std::vector<GpuVector> cumulative_fn(CumFn api, const NumContextProc& f, const std::vector<double>& X, const GpuVector& like,
                                            std::shared_ptr<NumContext> ctx, double dx) {
  if (!ctx) ctx = newNumContext();
  FnEnv env{&f, ctx.get(), like.device(), nullptr};
  std::vector<b200rk_vec*> slots(X.size() ? X.size() : 1, nullptr);
  size_t n = 0;
  const int rc = api(like.device()->handle(), &fn_trampoline, &env, like.size(), X.data(), X.size(), dx, slots.data(), &n);
  if (env.err) std::rethrow_exception(env.err);
  check(rc, like.device()->handle());
  return adopt_all(like.device(), slots, n);
}
}  // namespace detail
// cumtrapz(f, X, ctx, dx) (integrate.nim:138-175) / cumsimpson(f, X, ctx, dx) (integrate.nim:379-400); `like` fixes T's size
inline std::vector<GpuVector> cumtrapz(const NumContextProc& f, const std::vector<double>& X, const GpuVector& like,
                                       std::shared_ptr<NumContext> ctx = nullptr, double dx = 1e-5) {
  return detail::cumulative_fn(&b200rk_cumtrapz_fn, f, X, like, std::move(ctx), dx);
}
inline std::vector<GpuVector> cumsimpson(const NumContextProc& f, const std::vector<double>& X, const GpuVector& like,
                                         std::shared_ptr<NumContext> ctx = nullptr, double dx = 1e-5) {
  return detail::cumulative_fn(&b200rk_cumsimpson_fn, f, X, like, std::move(ctx), dx);
}

}  // namespace numericalnim
