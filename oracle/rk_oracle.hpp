// rk_oracle.hpp — CPU ORACLE (TEST INFRASTRUCTURE ONLY; never linked into or called by the product).
//
// A C++ restatement of the arithmetic of SciNim/numericalnim's explicit Runge–Kutta path:
//   src/numericalnim/ode.nim   (steppers, adaptive retry loop, ODESolver driver, solveODE dispatch)
//   src/numericalnim/utils.nim (Vector[T] element-wise operators, sum, hermiteSpline, linspace, isClose)
// The reference is Nim and there is no Nim toolchain in this image, so it cannot be compiled or run
// here. Every function below cites the reference file:line whose *arithmetic* it restates; the
// structure (stage-row tables, a generic left-associated weighted sum, a shared retry loop) is ours.
//
// PINNING STATUS
//   * Pinned against the reference's own tests: every in-scope case of tests/test_ode.nim
//     (t == tspan exactly, |y - exp(-0.1 t)| <= per-test tol), the operator known-answers of
//     tests/test_vector.nim and the linspace/isClose cases of tests/test_utils.nim — see
//     tests/test_oracle_reference_cases.py.
//   * Cross-checked bit-for-bit against an independent pure-Python restatement
//     (tests/pyref.py) and, for the tableaux, against scipy's Dormand–Prince coefficients and the
//     Runge–Kutta order conditions.
//   * PARITY UNPINNED for the step-rejection path and the dtMin limiter (ode.nim:71-76): no
//     reference test exercises them; they are restated from reading the code only.
//
// Build: g++ -O2 -ffp-contract=off (no -march=native, no -ffast-math) — what a default `nim c`
// build gives: plain SSE2 mul/add, no FMA contraction.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace rk_oracle {

// ---------------------------------------------------------------------------------------------
// Vector[T]  (utils.nim:14-20).  Every operator allocates a zero-filled buffer, fills it in one
// serial pass and then COPIES it into the result (newSeq at utils.nim:61, newVector's `@components`
// at utils.nim:20). The copy is kept so the CPU baseline has the reference's cost structure.
// ---------------------------------------------------------------------------------------------
struct Vector {
  std::vector<double> components;
  Vector() = default;
  explicit Vector(const std::vector<double>& c) : components(c) {}  // copy == `@components`
  Vector(const double* p, size_t n) : components(p, p + n) {}
  size_t len() const { return components.size(); }
  double operator[](size_t i) const { return components[i]; }
};

struct ValueError : std::invalid_argument {  // Nim's ValueError
  using std::invalid_argument::invalid_argument;
};

inline void check_sizes(const Vector& a, const Vector& b) {  // utils.nim:22-26
  if (a.len() != b.len()) throw ValueError("Vectors must have the same size.");
}

namespace detail {
template <class F>
inline Vector map1(const Vector& v, F f) {
  std::vector<double> tmp(v.len());  // zero-filled, like newSeq
  for (size_t i = 0; i < v.len(); ++i) tmp[i] = f(v.components[i]);
  return Vector(tmp);  // copy, like newVector(@...)
}
template <class F>
inline Vector map2(const Vector& a, const Vector& b, F f) {
  check_sizes(a, b);
  std::vector<double> tmp(a.len());
  for (size_t i = 0; i < a.len(); ++i) tmp[i] = f(a.components[i], b.components[i]);
  return Vector(tmp);
}
}  // namespace detail

inline Vector operator+(const Vector& a, const Vector& b) {  // utils.nim:59-64
  return detail::map2(a, b, [](double x, double y) { return x + y; });
}
inline Vector operator-(const Vector& a, const Vector& b) {  // utils.nim:113-118
  return detail::map2(a, b, [](double x, double y) { return x - y; });
}
inline Vector operator*(double d, const Vector& v) {  // utils.nim:176-180
  return detail::map1(v, [d](double x) { return x * d; });
}
inline Vector operator*(const Vector& v, double d) {  // utils.nim:171-175
  return detail::map1(v, [d](double x) { return x * d; });
}
inline Vector operator/(const Vector& v, double d) {  // utils.nim:166-170
  return detail::map1(v, [d](double x) { return x / d; });
}
inline Vector operator-(const Vector& v) {  // utils.nim:214-218
  return detail::map1(v, [](double x) { return -x; });
}
inline Vector vabs(const Vector& v) {  // utils.nim:219-223
  return detail::map1(v, [](double x) { return std::fabs(x); });
}
inline Vector add_scalar(double d, const Vector& v) {  // `+.` utils.nim:78-79 -> utils.nim:72-76
  return detail::map1(v, [d](double x) { return x + d; });
}
inline Vector hadamard(const Vector& a, const Vector& b) {  // `*.` utils.nim:186-191
  return detail::map2(a, b, [](double x, double y) { return x * y; });
}
inline Vector hdiv(const Vector& a, const Vector& b) {  // `/.` utils.nim:192-197
  return detail::map2(a, b, [](double x, double y) { return x / y; });
}
inline double dot(const Vector& a, const Vector& b) {  // utils.nim:181-185
  check_sizes(a, b);
  double r = 0.0;
  for (size_t i = 0; i < a.len(); ++i) r += a.components[i] * b.components[i];
  return r;
}
inline size_t vsize(const Vector& v) { return v.components.size(); }  // utils.nim:57
// sum(v) -> norm(v, 1) -> math.sum(@v): signed, sequential, left to right from 0
// (utils.nim:243-250, utils.nim:233-235; `@v` copies the seq, utils.nim:43).
inline double vsum(const Vector& v) {
  std::vector<double> copy(v.components);
  double r = 0.0;
  for (double x : copy) r = r + x;
  return r;
}
inline Vector vpow_nat(const Vector& v, unsigned p) {  // `^` utils.nim:254-259 (Nim math.`^`)
  return detail::map1(v, [p](double x) {
    double base = x, r = 1.0;
    unsigned e = p;
    while (true) {
      if (e & 1u) r *= base;
      e >>= 1;
      if (e == 0) break;
      base *= base;
    }
    return r;
  });
}
inline double vnorm(const Vector& v, int p = 2) {  // utils.nim:225-241
  if (p == 0) return *std::max_element(v.components.begin(), v.components.end());
  if (p == 1) return vsum(v);
  if (p == 2) return std::sqrt(vsum(vpow_nat(v, 2)));
  return std::pow(vsum(vpow_nat(v, (unsigned)p)), 1.0 / double(p));
}
inline Vector clone(const Vector& v) { return v; }  // utils.nim:269

// Scalar shims so the same templates instantiate for T = double (ode.nim:45-55).
inline double vabs(double x) { return std::fabs(x); }
inline double add_scalar(double d, double x) { return d + x; }
inline double hadamard(double a, double b) { return a * b; }
inline double hdiv(double a, double b) { return a / b; }
inline size_t vsize(double) { return 1; }
inline double vsum(double x) { return x; }
inline double clone(double x) { return x; }

// calcError / isClose (utils.nim:252, 270-271, 474-479)
inline double calc_error(double a, double b) { return std::fabs(a - b); }
inline double calc_error(const Vector& a, const Vector& b) { return vnorm(a - b) / double(a.len()); }
template <class T>
inline bool is_close(const T& a, const T& b, double tol = 1e-3) { return calc_error(a, b) <= tol; }

// linspace (utils.nim:498-507)
inline std::vector<double> linspace(double x1, double x2, long n) {
  if (n <= 0) throw ValueError("Number of samples must be greater then 0");
  std::vector<double> r;
  const double dx = (x2 - x1) / double(n - 1);
  r.push_back(x1);
  for (long i = 1; i <= n - 2; ++i) r.push_back(x1 + dx * double(i));
  r.push_back(x2);
  return r;
}

// hermiteSpline (utils.nim:273-279). `^` binds tighter than `*`; sums are left-associated.
template <class T>
inline T hermite_spline(double x, double x1, double x2, const T& y1, const T& y2, const T& dy1,
                        const T& dy2) {
  const double t = (x - x1) / (x2 - x1);
  const double u = 1.0 - t;
  const double h00 = (1.0 + 2.0 * t) * (u * u);
  const double h10 = t * (u * u);
  const double h01 = (t * t) * (3.0 - 2.0 * t);
  const double h11 = (t * (t * t)) - (t * t);
  return h00 * y1 + h10 * (x2 - x1) * dy1 + h01 * y2 + h11 * (x2 - x1) * dy2;
}

// Nim's min/max on floats (`if x <= y: x else: y` / `if y <= x: x else: y`): NaN-propagation differs
// from std::fmin/fmax, so restate them.
inline double nim_min(double x, double y) { return (x <= y) ? x : y; }
inline double nim_max(double x, double y) { return (y <= x) ? x : y; }

// ---------------------------------------------------------------------------------------------
// ODEoptions / newODEoptions (ode.nim:26-34, 78-102)
// ---------------------------------------------------------------------------------------------
struct Options {
  double dt, dtMax, dtMin, tStart, absTol, relTol, scaleMax, scaleMin;
};
inline Options new_options(double dt = 1e-4, double absTol = 1e-4, double relTol = 1e-4,
                           double dtMax = 1e-2, double dtMin = 1e-4, double scaleMax = 4.0,
                           double scaleMin = 0.1, double tStart = 0.0) {
  if (std::fabs(dtMax) < std::fabs(dtMin)) throw ValueError("dtMin must be less than dtMax");
  if (std::fabs(scaleMax) < 1) throw ValueError("scaleMax must be bigger than 1");
  if (1 < std::fabs(scaleMin)) throw ValueError("scaleMin must be smaller than 1");
  return Options{std::fabs(dt), std::fabs(dtMax), std::fabs(dtMin), tStart, std::fabs(absTol),
                 std::fabs(relTol), std::fabs(scaleMax), std::fabs(scaleMin)};
}

// NumContext (commonTypes.nim:3-15) plus oracle-only instrumentation counters.
template <class T>
struct Context {
  std::map<std::string, double> fValues;
  std::map<std::string, T> tValues;
  // instrumentation (not in the reference)
  long rhs_evals = 0, attempts = 0, rejected = 0, limiter_hits = 0, steps = 0;
  long max_steps = 0;  // > 0: stop the driver loop after this many accepted steps (bounded CPU timing)
  bool nan_guard_tripped = false;
  struct StepRecord { double t, dt_used, error; int attempts; };
  std::vector<StepRecord>* trace = nullptr;
};

template <class T>
using OdeProc = std::function<T(double, const T&, Context<T>*)>;  // ode.nim:36

template <class T>
struct StepResult {  // (yNew, newFSAL, dtUsed, error) — ode.nim:38
  T y_new, fsal;
  double dt, error;
};

template <class T>
using IntegratorProc = StepResult<T> (*)(const OdeProc<T>&, double, const T&, const T&, double,
                                         const Options&, Context<T>*);

// Left-associated weighted sum  w0*k0 + w1*k1 + ...  — the shape of every bracket in ode.nim's stage
// and solution rows (e.g. ode.nim:294-302): each product is one Vector pass, each `+` another,
// zero weights are multiplied through exactly as the reference does.
template <class T>
inline T wsum(const double* w, const T* const* k, int m) {
  T acc = w[0] * (*k[0]);
  for (int j = 1; j < m; ++j) acc = acc + w[j] * (*k[j]);
  return acc;
}

// ---------------------------------------------------------------------------------------------
// Fixed-step methods (ode.nim:107-189). Each returns (yNew, yNew, dt, 0.0).
// ---------------------------------------------------------------------------------------------
#define RK_ORACLE_FIXED_SIG(NAME)                                                              \
  template <class T>                                                                           \
  StepResult<T> NAME(const OdeProc<T>& f, double t, const T& y, const T& /*FSAL*/, double dt,  \
                     const Options& /*o*/, Context<T>* ctx)

RK_ORACLE_FIXED_SIG(heun2_step) {  // ode.nim:107-113
  T k1 = f(t, y, ctx);
  T k2 = f(t + dt, y + dt * k1, ctx);
  T yNew = y + 0.5 * dt * (k1 + k2);
  return {yNew, yNew, dt, 0.0};
}
RK_ORACLE_FIXED_SIG(ralston2_step) {  // ode.nim:115-121 (2/3 is float division in Nim)
  T k1 = f(t, y, ctx);
  T k2 = f(t + 2.0 / 3.0 * dt, y + 2.0 / 3.0 * dt * k1, ctx);
  T yNew = y + dt * (0.25 * k1 + 0.75 * k2);
  return {yNew, yNew, dt, 0.0};
}
RK_ORACLE_FIXED_SIG(kutta3_step) {  // ode.nim:123-130
  T k1 = f(t, y, ctx);
  T k2 = f(t + 0.5 * dt, y + 0.5 * dt * k1, ctx);
  T k3 = f(t + dt, y - dt * k1 + 2.0 * dt * k2, ctx);
  T yNew = y + dt * (1.0 / 6.0 * k1 + 2.0 / 3.0 * k2 + 1.0 / 6.0 * k3);
  return {yNew, yNew, dt, 0.0};
}
RK_ORACLE_FIXED_SIG(heun3_step) {  // ode.nim:132-139
  T k1 = f(t, y, ctx);
  T k2 = f(t + 1.0 / 3.0 * dt, y + 1.0 / 3.0 * dt * k1, ctx);
  T k3 = f(t + 2.0 / 3.0 * dt, y + 2.0 / 3.0 * dt * k2, ctx);
  T yNew = y + dt * (0.25 * k1 + 0.75 * k3);
  return {yNew, yNew, dt, 0.0};
}
RK_ORACLE_FIXED_SIG(ralston3_step) {  // ode.nim:141-148
  T k1 = f(t, y, ctx);
  T k2 = f(t + 1.0 / 2.0 * dt, y + 1.0 / 2.0 * dt * k1, ctx);
  T k3 = f(t + 3.0 / 4.0 * dt, y + 3.0 / 4.0 * dt * k2, ctx);
  T yNew = y + dt * (2.0 / 9.0 * k1 + 1.0 / 3.0 * k2 + 4.0 / 9.0 * k3);
  return {yNew, yNew, dt, 0.0};
}
RK_ORACLE_FIXED_SIG(ssprk3_step) {  // ode.nim:150-157
  T k1 = f(t, y, ctx);
  T k2 = f(t + dt, y + dt * k1, ctx);
  T k3 = f(t + 0.5 * dt, y + 0.25 * dt * (k1 + k2), ctx);
  T yNew = y + dt * (1.0 / 6.0 * k1 + 1.0 / 6.0 * k2 + 2.0 / 3.0 * k3);
  return {yNew, yNew, dt, 0.0};
}
RK_ORACLE_FIXED_SIG(ralston4_step) {  // ode.nim:160-168
  T k1 = f(t, y, ctx);
  T k2 = f(t + 0.4 * dt, y + 0.4 * dt * k1, ctx);
  T k3 = f(t + 0.45573725 * dt, y + dt * (0.29697761 * k1 + 0.15875964 * k2), ctx);
  T k4 = f(t + dt, y + dt * (0.21810040 * k1 - 3.05096516 * k2 + 3.83286476 * k3), ctx);
  T yNew = y + dt * (0.17476028 * k1 - 0.55148066 * k2 + 1.20553560 * k3 + 0.17118478 * k4);
  return {yNew, yNew, dt, 0.0};
}
RK_ORACLE_FIXED_SIG(kutta4_step) {  // ode.nim:170-178
  T k1 = f(t, y, ctx);
  T k2 = f(t + 1.0 / 3.0 * dt, y + 1.0 / 3.0 * dt * k1, ctx);
  T k3 = f(t + 2.0 / 3.0 * dt, y + dt * (-1.0 / 3.0 * k1 + k2), ctx);
  T k4 = f(t + dt, y + dt * (k1 - k2 + k3), ctx);
  T yNew = y + dt * (1.0 / 8.0 * k1 + 3.0 / 8.0 * k2 + 3.0 / 8.0 * k3 + 1.0 / 8.0 * k4);
  return {yNew, yNew, dt, 0.0};
}
RK_ORACLE_FIXED_SIG(rk4_step) {  // ode.nim:180-189
  T k1 = f(t, y, ctx);
  T k2 = f(t + 0.5 * dt, y + 0.5 * dt * k1, ctx);
  T k3 = f(t + 0.5 * dt, y + 0.5 * dt * k2, ctx);
  T k4 = f(t + dt, y + dt * k3, ctx);
  T yNew = y + dt / 6.0 * (k1 + 2.0 * (k2 + k3) + k4);
  return {yNew, yNew, dt, 0.0};
}
#undef RK_ORACLE_FIXED_SIG

// ---------------------------------------------------------------------------------------------
// Adaptive retry loop (commonAdaptiveMethodCode, ode.nim:57-76).
// `attempt(dt, yNew&, error_y&)` evaluates the stages and fills yNew / error_y.
// `order` is an int in the reference and 1/order is Nim int/int -> float.
// ---------------------------------------------------------------------------------------------
template <class T>
inline double scaled_rms(const T& yNew, const T& error_y, double absTol, double relTol) {
  T totalTol = add_scalar(absTol, relTol * vabs(yNew));      // ode.nim:61
  T err1 = hdiv(error_y, totalTol);                          // ode.nim:62
  T err1_square = hadamard(err1, err1);                      // ode.nim:63
  const double size = double(vsize(err1));                   // ode.nim:64
  return std::sqrt(1.0 / size * vsum(err1_square));          // ode.nim:65  (1/size)*sum, not sum/size
}

template <class T, class Attempt>
inline double adaptive_retry(Attempt&& attempt, T& yNew, double& dt, int order, const Options& o,
                             Context<T>* ctx) {
  const double absTol = o.absTol, relTol = o.relTol, dtMax = o.dtMax, dtMin = o.dtMin;
  int limitCounter = 0;
  double error = 0.0;
  int n_attempts = 0;
  while (limitCounter < 2) {                                  // ode.nim:58
    T error_y;
    attempt(dt, yNew, error_y);
    ++n_attempts;
    if (ctx) ctx->attempts++;
    error = scaled_rms(yNew, error_y, absTol, relTol);        // ode.nim:61-65
    if (error <= 1) break;                                    // ode.nim:69-70
    if (ctx) ctx->rejected++;
    dt = dt * nim_min(4, nim_max(0.125, 0.9 * std::pow(1.0 / error, 1.0 / double(order))));  // :71
    if (std::fabs(dt) < dtMin) {                              // ode.nim:72-74
      dt = dtMin;
      limitCounter += 1;
      if (ctx) ctx->limiter_hits++;
    } else if (dtMax < std::fabs(dt)) {                       // ode.nim:75-76
      dt = dtMax;
    }
    if (std::isnan(dt)) {  // the reference would spin forever here; the oracle stops.
      if (ctx) ctx->nan_guard_tripped = true;
      break;
    }
  }
  return error;
}

// RK21 (ode.nim:191-210): Heun with embedded Euler; k1 recomputed every attempt; FSAL out = yNew.
template <class T>
StepResult<T> rk21_step(const OdeProc<T>& f, double t, const T& y, const T& /*FSAL*/, double dt_in,
                        const Options& o, Context<T>* ctx) {
  T yNew;
  double dt = dt_in;
  double error = adaptive_retry<T>(
      [&](double h, T& yN, T& err) {
        T k1 = f(t, y, ctx);
        T k2 = f(t + h, y + h * k1, ctx);
        yN = y + h * 0.5 * (k1 + k2);
        T yLow = y + h * k1;
        err = yN - yLow;
      },
      yNew, dt, 2, o, ctx);
  return {yNew, yNew, dt, error};
}

// BS32 (ode.nim:212-234): ignores the incoming FSAL (k1 recomputed, :225) but returns k4 as FSAL.
template <class T>
StepResult<T> bs32_step(const OdeProc<T>& f, double t, const T& y, const T& /*FSAL*/, double dt_in,
                        const Options& o, Context<T>* ctx) {
  T yNew, k4;
  double dt = dt_in;
  double error = adaptive_retry<T>(
      [&](double h, T& yN, T& err) {
        T k1 = f(t, y, ctx);
        T k2 = f(t + 0.5 * h, y + 0.5 * h * k1, ctx);
        T k3 = f(t + 0.75 * h, y + 0.75 * h * k2, ctx);
        yN = y + h * (2.0 / 9.0 * k1 + 1.0 / 3.0 * k2 + 4.0 / 9.0 * k3);
        k4 = f(t + h, yN, ctx);
        T yLow = y + h * (7.0 / 24.0 * k1 + 1.0 / 4.0 * k2 + 1.0 / 3.0 * k3 + 1.0 / 8.0 * k4);
        err = yN - yLow;
      },
      yNew, dt, 3, o, ctx);
  return {yNew, k4, dt, error};
}

// ---------------------------------------------------------------------------------------------
// Tableaux of the three high-order FSAL pairs. Literals are transcribed digit for digit from
// ode.nim:241-282 (DOPRI54), ode.nim:311-352 (Tsit54), ode.nim:381-443 (Vern65). Rows are dense
// lower-triangular including the zero entries, which the reference multiplies through.
// ---------------------------------------------------------------------------------------------
struct Pair {
  int stages;            // 7 or 9 (last stage is the FSAL evaluation)
  int order;             // int `order` handed to the retry loop (ode.nim:292,362,453)
  int n_b, n_bhat;       // number of terms in the yNew row and in the bHat row
  bool err_is_direct;    // Tsit54: error_y = dt*(bHat . k) (ode.nim:372); else yNew - yLow
  double c[10];          // c[s] for stage s (1-based; c[1] unused)
  double a[10][9];       // a[s][j-1] = a_sj, s = 2..stages
  double b[9], bhat[9];
};

inline const Pair& dopri54_pair() {
  static const Pair p = [] {
    Pair q{};
    q.stages = 7; q.order = 5; q.n_b = 6; q.n_bhat = 7; q.err_is_direct = false;
    q.c[2] = 1.0 / 5.0; q.c[3] = 3.0 / 10.0; q.c[4] = 4.0 / 5.0; q.c[5] = 8.0 / 9.0; q.c[6] = 1.0; q.c[7] = 1.0;
    q.a[2][0] = 1.0 / 5.0;
    q.a[3][0] = 3.0 / 40.0; q.a[3][1] = 9.0 / 40.0;
    q.a[4][0] = 44.0 / 45.0; q.a[4][1] = -56.0 / 15.0; q.a[4][2] = 32.0 / 9.0;
    q.a[5][0] = 19372.0 / 6561.0; q.a[5][1] = -25360.0 / 2187.0; q.a[5][2] = 64448.0 / 6561.0; q.a[5][3] = -212.0 / 729.0;
    q.a[6][0] = 9017.0 / 3168.0; q.a[6][1] = -355.0 / 33.0; q.a[6][2] = 46732.0 / 5247.0; q.a[6][3] = 49.0 / 176.0; q.a[6][4] = -5103.0 / 18656.0;
    q.a[7][0] = 35.0 / 384.0; q.a[7][1] = 0.0; q.a[7][2] = 500.0 / 1113.0; q.a[7][3] = 125.0 / 192.0; q.a[7][4] = -2187.0 / 6784.0; q.a[7][5] = 11.0 / 84.0;
    for (int j = 0; j < 6; ++j) q.b[j] = q.a[7][j];  // b_i = a7i (ode.nim:269-274)
    q.bhat[0] = 5179.0 / 57600.0; q.bhat[1] = 0.0; q.bhat[2] = 7571.0 / 16695.0; q.bhat[3] = 393.0 / 640.0;
    q.bhat[4] = -92097.0 / 339200.0; q.bhat[5] = 187.0 / 2100.0; q.bhat[6] = 1.0 / 40.0;
    return q;
  }();
  return p;
}

inline const Pair& tsit54_pair() {
  static const Pair p = [] {
    Pair q{};
    q.stages = 7; q.order = 5; q.n_b = 6; q.n_bhat = 7; q.err_is_direct = true;
    q.c[2] = 0.161; q.c[3] = 0.327; q.c[4] = 0.9; q.c[5] = 0.9800255409045097; q.c[6] = 1.0; q.c[7] = 1.0;
    q.a[2][0] = 0.161;
    q.a[3][0] = -0.008480655492356989; q.a[3][1] = 0.335480655492357;
    q.a[4][0] = 2.8971530571054935; q.a[4][1] = -6.359448489975075; q.a[4][2] = 4.3622954328695815;
    q.a[5][0] = 5.325864828439257; q.a[5][1] = -11.748883564062828; q.a[5][2] = 7.4955393428898365; q.a[5][3] = -0.09249506636175525;
    q.a[6][0] = 5.86145544294642; q.a[6][1] = -12.92096931784711; q.a[6][2] = 8.159367898576159; q.a[6][3] = -0.071584973281401; q.a[6][4] = -0.028269050394068383;
    q.a[7][0] = 0.09646076681806523; q.a[7][1] = 0.01; q.a[7][2] = 0.4798896504144996; q.a[7][3] = 1.379008574103742; q.a[7][4] = -3.290069515436081; q.a[7][5] = 2.324710524099774;
    for (int j = 0; j < 6; ++j) q.b[j] = q.a[7][j];  // ode.nim:339-344
    q.bhat[0] = -0.001780011052226; q.bhat[1] = -0.000816434459657; q.bhat[2] = 0.007880878010262; q.bhat[3] = -0.144711007173263;
    q.bhat[4] = 0.582357165452555; q.bhat[5] = -0.458082105929187; q.bhat[6] = 1.0 / 66.0;
    return q;
  }();
  return p;
}

inline const Pair& vern65_pair() {
  static const Pair p = [] {
    Pair q{};
    q.stages = 9; q.order = 6; q.n_b = 8; q.n_bhat = 9; q.err_is_direct = false;
    q.c[2] = 0.06; q.c[3] = 0.09593333333333333; q.c[4] = 0.1439; q.c[5] = 0.4973; q.c[6] = 0.9725; q.c[7] = 0.9995; q.c[8] = 1.0; q.c[9] = 1.0;
    q.a[2][0] = 0.06;
    q.a[3][0] = 0.019239962962962962; q.a[3][1] = 0.07669337037037037;
    q.a[4][0] = 0.035975; q.a[4][1] = 0.0; q.a[4][2] = 0.107925;
    q.a[5][0] = 1.3186834152331484; q.a[5][1] = 0.0; q.a[5][2] = -5.042058063628562; q.a[5][3] = 4.220674648395414;
    q.a[6][0] = -41.87259166432751; q.a[6][1] = 0.0; q.a[6][2] = 159.43256216313748; q.a[6][3] = -122.11921356501004; q.a[6][4] = 5.531743066200053;
    q.a[7][0] = -54.430156935316504; q.a[7][1] = 0.0; q.a[7][2] = 207.06725136501848; q.a[7][3] = -158.61081378459; q.a[7][4] = 6.991816585950242; q.a[7][5] = -0.01859723106220323;
    q.a[8][0] = -54.66374178728198; q.a[8][1] = 0.0; q.a[8][2] = 207.95280625538936; q.a[8][3] = -159.2889574744995; q.a[8][4] = 7.018743740796944; q.a[8][5] = -0.018338785905045722; q.a[8][6] = -0.0005119484997882099;
    q.a[9][0] = 0.03438957868357036; q.a[9][1] = 0.0; q.a[9][2] = 0.0; q.a[9][3] = 0.25826245556335037; q.a[9][4] = 0.4209371189673537; q.a[9][5] = 4.405396469669310; q.a[9][6] = -176.48311902429865; q.a[9][7] = 172.36413340141507;
    // sixth-order weights are separate literals here (ode.nim:426-433), NOT aliases of row 9:
    q.b[0] = 0.03438957868357036; q.b[1] = 0.0; q.b[2] = 0.0; q.b[3] = 0.25826245556335034; q.b[4] = 0.42093711896735372;
    q.b[5] = 4.4053964696693102; q.b[6] = -176.48311902429866; q.b[7] = 172.36413340141507;
    q.bhat[0] = 0.04909967648382; q.bhat[1] = 0.0; q.bhat[2] = 0.0; q.bhat[3] = 0.22511122295165; q.bhat[4] = 0.46946822530296;
    q.bhat[5] = 0.80657922499889; q.bhat[6] = 0.0; q.bhat[7] = -0.60711948917780; q.bhat[8] = 0.05686113944048;
    return q;
  }();
  return p;
}

// Stage input  y + dt * (a_s1*k1 + ... + a_s,s-1*k_{s-1})   (ode.nim:294-299, 364-369, 455-462)
template <class T>
inline T pair_stage_input(const Pair& p, int s, const T& y, double dt, const T* const* k) {
  return y + dt * wsum<T>(p.a[s], k, s - 1);
}
// yNew = y + dt * (b . k)   (ode.nim:301, 371, 464)
template <class T>
inline T pair_y_new(const Pair& p, const T& y, double dt, const T* const* k) {
  return y + dt * wsum<T>(p.b, k, p.n_b);
}
// error_y (ode.nim:302-303, 372, 465-466)
template <class T>
inline T pair_error_y(const Pair& p, const T& y, const T& yNew, double dt, const T* const* k) {
  if (p.err_is_direct) return dt * wsum<T>(p.bhat, k, p.n_bhat);
  T yLow = y + dt * wsum<T>(p.bhat, k, p.n_bhat);
  return yNew - yLow;
}

// One step of an FSAL pair: DOPRI54_step / TSIT54_step / VERN65_step (ode.nim:237-305, 307-374,
// 377-468). k1 = FSAL on every attempt (never recomputed on retry); returns (yNew, k_last, dt, error).
template <class T>
StepResult<T> pair_step(const Pair& p, const OdeProc<T>& f, double t, const T& y, const T& FSAL,
                        double dt_in, const Options& o, Context<T>* ctx) {
  T yNew;
  std::vector<T> k(p.stages);
  std::vector<const T*> kp(p.stages);
  for (int s = 0; s < p.stages; ++s) kp[s] = &k[s];
  double dt = dt_in;
  double error = adaptive_retry<T>(
      [&](double h, T& yN, T& err) {
        k[0] = FSAL;
        for (int s = 2; s <= p.stages; ++s)
          k[s - 1] = f(t + h * p.c[s], pair_stage_input<T>(p, s, y, h, kp.data()), ctx);
        yN = pair_y_new<T>(p, y, h, kp.data());
        err = pair_error_y<T>(p, y, yN, h, kp.data());
      },
      yNew, dt, p.order, o, ctx);
  return {yNew, k[p.stages - 1], dt, error};
}
template <class T>
StepResult<T> dopri54_step(const OdeProc<T>& f, double t, const T& y, const T& F, double dt, const Options& o, Context<T>* c) {
  return pair_step<T>(dopri54_pair(), f, t, y, F, dt, o, c);
}
template <class T>
StepResult<T> tsit54_step(const OdeProc<T>& f, double t, const T& y, const T& F, double dt, const Options& o, Context<T>* c) {
  return pair_step<T>(tsit54_pair(), f, t, y, F, dt, o, c);
}
template <class T>
StepResult<T> vern65_step(const OdeProc<T>& f, double t, const T& y, const T& F, double dt, const Options& o, Context<T>* c) {
  return pair_step<T>(vern65_pair(), f, t, y, F, dt, o, c);
}

// ---------------------------------------------------------------------------------------------
// Driver: ODESolver (ode.nim:471-586)
// ---------------------------------------------------------------------------------------------
template <class T>
struct Solution {
  std::vector<double> t;
  std::vector<T> y;
};

template <class T>
Solution<T> ode_solver(const OdeProc<T>& f, const T& y0, const std::vector<double>& tspan,
                       const Options& options, IntegratorProc<T> integrator, bool useFSAL,
                       double order, bool adaptive, Context<T>* ctx) {
  const double t0 = options.tStart;
  double t = t0;
  std::vector<double> tPositive, tNegative;
  for (double x : tspan) if (x > t0) tPositive.push_back(x);          // ode.nim:479
  for (double x : tspan) if (x < t0) tNegative.push_back(x);          // ode.nim:480
  std::reverse(tNegative.begin(), tNegative.end());
  std::vector<T> yPositive, yNegative, yZero;
  std::vector<double> tZero;
  T y = clone(y0);                                                     // ode.nim:482
  if (std::find(tspan.begin(), tspan.end(), t0) != tspan.end()) {      // ode.nim:485-487
    yZero.push_back(y);
    tZero.push_back(t0);
  }
  const double dtMax = options.dtMax, dtMin = options.dtMin;
  double dt, dtInit;
  if (adaptive) dtInit = std::sqrt(dtMax * dtMin);                     // ode.nim:491-493
  else dtInit = options.dt;                                            // ode.nim:495-496
  dt = dtInit;
  struct Last { double t; T y; T dy; };
  Last lastIter{t0, y, f(t0, y, ctx)};                                 // ode.nim:498
  const bool useDense = (tspan.size() != 2);                           // ode.nim:499-502
  long denseIndex = 0;
  double error = 0.0;
  T FSAL = f(t0, y, ctx);                                              // ode.nim:506
  double tEnd;

  auto controller = [&](double& h) {                                   // ode.nim:533-541 / 575-583
    if (error == 0.0) h *= 5;
    else h = h * nim_min(4, nim_max(0.125, 0.9 * std::pow(1.0 / error, 1.0 / order)));
    if (h < dtMin) h = dtMin;
    else if (dtMax < h) h = dtMax;
  };
  auto record = [&](double t_before, double h_used, long attempts_before) {
    if (!ctx) return;
    ctx->steps++;
    if (ctx->trace)
      ctx->trace->push_back({t_before, h_used, error, int(ctx->attempts - attempts_before)});
  };

  if (!tPositive.empty()) {                                            // ode.nim:508-542
    dt = dtInit;
    tEnd = *std::max_element(tPositive.begin(), tPositive.end());
    const long high = long(tPositive.size()) - 1;
    while (t < tEnd) {
      if (useDense) {
        if (high < denseIndex) break;
        while (tPositive[denseIndex] <= t) {
          if (useFSAL)
            yPositive.push_back(hermite_spline<T>(tPositive[denseIndex], lastIter.t, t, lastIter.y, y, lastIter.dy, FSAL));
          else
            yPositive.push_back(hermite_spline<T>(tPositive[denseIndex], lastIter.t, t, lastIter.y, y, lastIter.dy, f(t, y, ctx)));
          denseIndex += 1;
          if (high < denseIndex) break;
        }
      }
      dt = nim_min(dt, tEnd - t);                                      // ode.nim:525
      if (useDense) {
        if (useFSAL) lastIter = Last{t, y, FSAL};
        else lastIter = Last{t, y, f(t, y, ctx)};                      // ode.nim:530
      }
      const long a0 = ctx ? ctx->attempts : 0;
      const double t_before = t;
      StepResult<T> r = integrator(f, t, y, FSAL, dt, options, ctx);   // ode.nim:531
      y = r.y_new; FSAL = r.fsal; dt = r.dt; error = r.error;
      t += dt;                                                         // ode.nim:532
      record(t_before, dt, a0);
      if (adaptive) controller(dt);
      if (ctx && ctx->nan_guard_tripped) break;
      if (ctx && ctx->max_steps > 0 && ctx->steps >= ctx->max_steps) break;
    }
    yPositive.push_back(y);                                            // ode.nim:542
  }

  if (!tNegative.empty()) {                                            // ode.nim:544-584
    OdeProc<T> g = [&f](double tt, const T& yy, Context<T>* c) -> T { return -f(-tt, yy, c); };  // :545
    FSAL = g(-t0, clone(y0), ctx);
    dt = dtInit;
    lastIter = Last{-t0, clone(y0), FSAL};
    tEnd = -*std::min_element(tNegative.begin(), tNegative.end());
    t = -t0;
    y = clone(y0);
    denseIndex = 0;
    const long high = long(tNegative.size()) - 1;
    while (t < tEnd) {
      if (useDense) {
        if (high < denseIndex) break;
        while (-tNegative[denseIndex] <= t) {
          if (useFSAL)
            yNegative.push_back(hermite_spline<T>(-tNegative[denseIndex], lastIter.t, t, lastIter.y, y, lastIter.dy, FSAL));
          else
            yNegative.push_back(hermite_spline<T>(-tNegative[denseIndex], lastIter.t, t, lastIter.y, y, lastIter.dy, g(t, y, ctx)));
          denseIndex += 1;
          if (high < denseIndex) break;
        }
      }
      dt = nim_min(dt, tEnd - t);
      if (useDense) {
        if (useFSAL) lastIter = Last{t, y, FSAL};
        else lastIter = Last{t, y, g(t, y, ctx)};
      }
      const long a0 = ctx ? ctx->attempts : 0;
      const double t_before = t;
      StepResult<T> r = integrator(g, t, y, FSAL, dt, options, ctx);
      y = r.y_new; FSAL = r.fsal; dt = r.dt; error = r.error;
      t += dt;
      record(-t_before, dt, a0);
      if (adaptive) controller(dt);
      if (ctx && ctx->nan_guard_tripped) break;
      if (ctx && ctx->max_steps > 0 && ctx->steps >= ctx->max_steps) break;
    }
    yNegative.push_back(y);
  }

  Solution<T> out;                                                     // ode.nim:585-586
  out.t.assign(tNegative.rbegin(), tNegative.rend());
  out.t.insert(out.t.end(), tZero.begin(), tZero.end());
  out.t.insert(out.t.end(), tPositive.begin(), tPositive.end());
  out.y.assign(yNegative.rbegin(), yNegative.rend());
  out.y.insert(out.y.end(), yZero.begin(), yZero.end());
  out.y.insert(out.y.end(), yPositive.begin(), yPositive.end());
  return out;
}

// Dispatch table == solveODE's `case integrator.toLower()` (ode.nim:607-651).
template <class T>
struct MethodEntry {
  const char* name;
  IntegratorProc<T> step;
  bool useFSAL;
  double order;
  bool adaptive;
};
template <class T>
inline const std::vector<MethodEntry<T>>& method_table() {
  static const std::vector<MethodEntry<T>> tab = {
      {"dopri54", &dopri54_step<T>, true, 5.0, true},   {"rk21", &rk21_step<T>, false, 2.0, true},
      {"bs32", &bs32_step<T>, true, 3.0, true},         {"rk4", &rk4_step<T>, false, 4.0, false},
      {"heun2", &heun2_step<T>, false, 2.0, false},     {"ralston2", &ralston2_step<T>, false, 2.0, false},
      {"kutta3", &kutta3_step<T>, false, 3.0, false},   {"heun3", &heun3_step<T>, false, 3.0, false},
      {"ralston3", &ralston3_step<T>, false, 3.0, false}, {"ssprk3", &ssprk3_step<T>, false, 3.0, false},
      {"ralston4", &ralston4_step<T>, false, 4.0, false}, {"kutta4", &kutta4_step<T>, false, 4.0, false},
      {"vern65", &vern65_step<T>, true, 6.0, true},     {"tsit54", &tsit54_step<T>, true, 5.0, true},
  };
  return tab;
}

template <class T>
Solution<T> solve_ode(const OdeProc<T>& f, const T& y0, std::vector<double> tspan,
                      const Options& options, Context<T>* ctx, const std::string& integrator = "dopri54") {
  Context<T> local;
  if (!ctx) ctx = &local;                                              // ode.nim:604-606
  std::string name = integrator;
  for (char& ch : name) ch = char(std::tolower((unsigned char)ch));    // ode.nim:607
  std::sort(tspan.begin(), tspan.end());                               // tspan.sorted(), ode.nim:609
  for (const auto& m : method_table<T>())
    if (name == m.name) return ode_solver<T>(f, y0, tspan, options, m.step, m.useFSAL, m.order, m.adaptive, ctx);
  throw ValueError(integrator + " is not a valid integrator");        // ode.nim:650-651
}

// ---------------------------------------------------------------------------------------------
// Benchmark right-hand sides. Only `scale` (tests/test_ode.nim:5-7: -0.1 * y) comes from the
// reference; the other two are the synthetic IVPs of BASELINE.json written with the reference's
// Vector operators (diag-linear) or as the plain per-element loop a user closure would contain.
// ---------------------------------------------------------------------------------------------
inline OdeProc<Vector> rhs_scale_vec(double c) {
  return [c](double, const Vector& y, Context<Vector>* ctx) { if (ctx) ctx->rhs_evals++; return c * y; };
}
inline OdeProc<double> rhs_scale_scalar(double c) {
  return [c](double, const double& y, Context<double>* ctx) { if (ctx) ctx->rhs_evals++; return c * y; };
}
inline OdeProc<Vector> rhs_diag_linear(const Vector& lambda) {  // k = -(lambda *. y)
  return [lambda](double, const Vector& y, Context<Vector>* ctx) {
    if (ctx) ctx->rhs_evals++;
    return -hadamard(lambda, y);
  };
}
inline OdeProc<Vector> rhs_lorenz96(double F) {  // k[i] = ((y[i+1] - y[i-2]) * y[i-1] - y[i]) + F, cyclic
  return [F](double, const Vector& y, Context<Vector>* ctx) {
    if (ctx) ctx->rhs_evals++;
    const size_t n = y.len();
    std::vector<double> k(n);
    for (size_t i = 0; i < n; ++i) {
      const double yp1 = y[(i + 1) % n], ym1 = y[(i + n - 1) % n], ym2 = y[(i + n - 2) % n];
      k[i] = ((yp1 - ym2) * ym1 - y[i]) + F;
    }
    return Vector(k);
  };
}

}  // namespace rk_oracle
