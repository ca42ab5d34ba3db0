// quad_oracle.hpp — CPU restatement of the reference's Hermite interpolation of a data set and of the
// cumulative quadrature routines that consume it, generic over T = double and
// T = Vector like the reference's own generics.
//
// TEST INFRASTRUCTURE ONLY (same rules as rk_oracle.hpp: nothing under numericalnim_b200/ may use it).
//
// Follows, operation for operation (Nim precedence: `^` binds tighter than `*`, `* /` before `+ -`, everything
// left-associated; scalar * Vector multiplies every component BY the scalar, utils.nim:171-180):
//   hermiteInterpolate       src/numericalnim/utils.nim:282-312
//   sortDataset / removeDuplicates / sortAndTrimDataset   utils.nim:349-420
//   cumtrapz(Y, X)           src/numericalnim/integrate.nim:119-135
//   cumtrapz(f, X, ctx, dx)  integrate.nim:138-175
//   cumsimpson(Y, X)         integrate.nim:330-378
//   cumsimpson(f, X, ctx, dx) integrate.nim:379-400
// Third-party arithmetic not under /root/reference: Nim's stdlib (unpinned, `requires "nim >= 1.0"`):
//   math.`^`(x, y: Natural) — y = 2: x*x, y = 3: x*x*x, else square-and-multiply (restated in ipow);
//   algorithm.sort on (x, index) tuples — lexicographic, so equal x keep their input order;
//   algorithm.isSorted — non-strict ascending; system.toInt(float) — round half away from zero.
// Pinning: the reference's own cases tests/test_integrate.nim:67-95 (cumtrapz / cumsimpson, discrete and function
// variants, X = linspace(0, 3pi/2, 17), Y = 2cos x against 2 sin x with its tolerances) — see
// tests/test_oracle_quadrature.py, which also compares every routine bit for bit with an independent pure-Python
// restatement (tests/pyref_quad.py). The branch of hermiteInterpolate for unsorted x and the duplicate handling of
// sortAndTrimDataset are reached by no reference test: PARITY UNPINNED there, restated from the code.
#pragma once
#include <algorithm>
#include <cmath>
#include <functional>
#include <numeric>
#include <string>
#include <vector>

#include "rk_oracle.hpp"

namespace rk_oracle {

inline double ipow(double x, unsigned y) {  // Nim math.`^`
  if (y == 0) return 1.0;
  if (y == 1) return x;
  if (y == 2) return x * x;
  if (y == 3) return x * x * x;
  double r = 1.0;
  while (true) {
    if (y & 1u) r *= x;
    y >>= 1;
    if (y == 0) break;
    x *= x;
  }
  return r;
}
inline long nim_to_int(double x) { return (long)std::llround(x); }  // system.toInt: round half away from zero

inline bool t_equal(double a, double b) { return a == b; }
inline bool t_equal(const Vector& a, const Vector& b) {  // utils.nim:51-55: walks v1's length, no size check
  for (size_t i = 0; i < a.len(); ++i)
    if (a.components[i] != b.components[i]) return false;
  return true;
}
inline std::string t_repr(double a) { return std::to_string(a); }
inline std::string t_repr(const Vector&) { return "Vector(...)"; }

struct AssertionDefect : std::logic_error {  // Nim `assert` (utils.nim:362, 387-388, 400)
  using std::logic_error::logic_error;
};
struct IndexDefect : std::out_of_range {  // Nim bounds check
  using std::out_of_range::out_of_range;
};

template <class T>
struct Dataset {
  std::vector<double> x;
  std::vector<T> y;
};

// sortDataset (utils.nim:385-409) then removeDuplicates (utils.nim:360-383)
template <class T>
inline Dataset<T> sort_and_trim(const std::vector<double>& X, const std::vector<T>& Y) {
  if (X.empty()) throw AssertionDefect("x is empty!");
  if (Y.size() != X.size()) throw AssertionDefect("y seq at index 0 has length " + std::to_string(Y.size()) + " while the first seq has length " +
                                                   std::to_string(X.size()) + ". They must match!");
  std::vector<size_t> idx(X.size());
  std::iota(idx.begin(), idx.end(), size_t(0));
  std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return X[a] < X[b]; });  // (x, index) tuples
  Dataset<T> s;
  for (size_t i : idx) { s.x.push_back(X[i]); s.y.push_back(Y[i]); }
  // groups of equal x (getIndexTable / findDuplicates); the first index of each group is kept
  // (after the sort equal x are adjacent, so the table of the reference reduces to runs)
  std::vector<bool> drop(s.x.size(), false);
  for (size_t i = 0; i < s.x.size(); ++i) {
    if (drop[i]) continue;
    for (size_t j = i + 1; j < s.x.size() && s.x[j] == s.x[i]; ++j) {
      if (!t_equal(s.y[j], s.y[i]))
        throw ValueError("impure y-duplicates was found: " + t_repr(s.y[i]) + " at index " + std::to_string(i) + " and " + t_repr(s.y[j]) +
                         " at index " + std::to_string(j));
      drop[j] = true;
    }
  }
  Dataset<T> r;
  for (size_t i = 0; i < s.x.size(); ++i)
    if (!drop[i]) { r.x.push_back(s.x[i]); r.y.push_back(s.y[i]); }
  return r;
}

// hermiteInterpolate (utils.nim:282-312)
template <class T>
inline std::vector<T> hermite_interpolate(const std::vector<double>& x, const std::vector<double>& t, const std::vector<T>& y,
                                          const std::vector<T>& dy) {
  std::vector<T> result;
  if (x.empty() || t.empty()) throw IndexDefect("index out of bounds, the container is empty");
  const long thigh = (long)t.size() - 1, xhigh = (long)x.size() - 1;
  long xIndex = 0;
  if (std::is_sorted(x.begin(), x.end())) {
    for (long i = 0; i <= thigh - 1; ++i) {
      while (t[i] <= x[xIndex] && x[xIndex] < t[i + 1]) {
        result.push_back(hermite_spline<T>(x[xIndex], t[i], t[i + 1], y[i], y[i + 1], dy[i], dy[i + 1]));
        xIndex += 1;
        if (xhigh < xIndex) break;
      }
      if (xhigh < xIndex) break;
    }
    if (x[xhigh] == t[thigh]) result.push_back(y[y.size() - 1]);
  } else {
    for (double a : x) {
      bool found = false;
      for (long i = 0; i <= thigh - 1; ++i) {
        if (t[i] <= a && a < t[i + 1]) {
          result.push_back(hermite_spline<T>(a, t[i], t[i + 1], y[i], y[i + 1], dy[i], dy[i + 1]));
          found = true;
          break;
        }
      }
      if (found) continue;
      if (a == t[thigh]) result.push_back(y[y.size() - 1]);
      else
        throw ValueError(std::to_string(a) + " not in interval " + std::to_string(*std::min_element(t.begin(), t.end())) + " - " +
                         std::to_string(*std::max_element(t.begin(), t.end())));
    }
  }
  return result;
}

// cumtrapz(Y, X) (integrate.nim:119-135)
template <class T>
inline std::vector<T> cumtrapz(const std::vector<T>& Y, const std::vector<double>& X) {
  Dataset<T> s = sort_and_trim(X, Y);
  std::vector<T> result;
  result.push_back(s.y[0] - s.y[0]);  // the right kind of zero
  T integral = s.y[0] - s.y[0];
  for (size_t i = 0; i + 1 < s.x.size(); ++i) {
    integral = integral + 0.5 * (s.x[i + 1] - s.x[i]) * (s.y[i + 1] + s.y[i]);
    result.push_back(integral);
  }
  return result;
}

template <class T>
using FnOfT = std::function<T(double)>;  // NumContextProc[T, float] with the context captured

// cumtrapz(f, X, ctx, dx) (integrate.nim:138-175): integrates to max(X) + 1.0 in steps of dx, then interpolates
template <class T>
inline std::vector<T> cumtrapz_fn(const FnOfT<T>& f, const std::vector<double>& X, double dx = 1e-5) {
  if (X.empty()) throw IndexDefect("index out of bounds, the container is empty");
  std::vector<double> times;
  std::vector<T> dy, y;
  double t = *std::min_element(X.begin(), X.end());
  const double tEnd = *std::max_element(X.begin(), X.end()) + 1.0;
  T dyTemp = f(t);
  T integral = dyTemp - dyTemp;
  times.push_back(t); dy.push_back(dyTemp); y.push_back(integral);
  t += dx;
  while (t <= tEnd) {
    T dyPrev = dyTemp;
    dyTemp = f(t);
    integral = integral + 0.5 * dx * (dyPrev + dyTemp);
    times.push_back(t); dy.push_back(dyTemp); y.push_back(integral);
    t += dx;
  }
  return hermite_interpolate<T>(X, times, y, dy);
}

// coefficient triples of the discrete Simpson rule on two unequal intervals (integrate.nim:357-359, 367-369)
struct SimpsonWeights { double alpha, beta, eta; };
inline SimpsonWeights simpson_pair_weights(double h1, double h2) {
  SimpsonWeights w;
  w.alpha = (2.0 * ipow(h2, 3) - ipow(h1, 3) + 3.0 * h1 * ipow(h2, 2)) / (6.0 * h2 * (h2 + h1));
  w.beta = (ipow(h2, 3) + ipow(h1, 3) + 3.0 * h1 * h2 * (h2 + h1)) / (6.0 * h2 * h1);
  w.eta = (2.0 * ipow(h1, 3) - ipow(h2, 3) + 3.0 * h2 * ipow(h1, 2)) / (6.0 * h1 * (h2 + h1));
  return w;
}
inline SimpsonWeights simpson_tail_weights(double h1, double h2) {
  SimpsonWeights w;
  w.alpha = (2.0 * ipow(h2, 2) + 3.0 * h1 * h2) / (6.0 * (h1 + h2));
  w.beta = (ipow(h2, 2) + 3.0 * h1 * h2) / (6.0 * h1);
  w.eta = -(ipow(h2, 3)) / (6.0 * h1 * (h1 + h2));
  return w;
}

// cumsimpson(Y, X) (integrate.nim:330-378)
template <class T>
inline std::vector<T> cumsimpson(const std::vector<T>& Y, const std::vector<double>& X) {
  Dataset<T> s = sort_and_trim(X, Y);
  long N = (long)s.x.size();
  bool evenN = false;
  if (N < 3) throw ValueError("X and Y must have at least 3 elements to perform Simpson, use cumtrapz instead");
  if (N % 2 == 0) { evenN = true; N -= 1; }
  std::vector<T> y, dy;
  std::vector<double> xs;
  T integral = s.y[0] - s.y[0];
  y.push_back(integral); dy.push_back(s.y[0]); xs.push_back(s.x[0]);
  const long pairs = nim_to_int(double(N - 1) / 2.0);
  for (long i = 0; i < pairs; ++i) {
    const double h1 = s.x[2 * i + 1] - s.x[2 * i];
    const double h2 = s.x[2 * i + 2] - s.x[2 * i + 1];
    const SimpsonWeights w = simpson_pair_weights(h1, h2);
    integral = integral + (w.alpha * s.y[2 * i + 2] + w.beta * s.y[2 * i + 1] + w.eta * s.y[2 * i]);
    y.push_back(integral); dy.push_back(s.y[2 * i + 2]); xs.push_back(s.x[2 * i + 2]);
  }
  if (evenN) {
    const long last = (long)s.x.size() - 1;
    const double h1 = s.x[last - 1] - s.x[last - 2];
    const double h2 = s.x[last] - s.x[last - 1];
    const SimpsonWeights w = simpson_tail_weights(h1, h2);
    integral = integral + (w.eta * s.y[last - 2] + w.beta * s.y[last - 1] + w.alpha * s.y[last]);
    y.push_back(integral); dy.push_back(s.y[last]); xs.push_back(s.x[last]);
  }
  return hermite_interpolate<T>(X, xs, y, dy);
}

// cumsimpson(f, X, ctx, dx) (integrate.nim:379-400)
template <class T>
inline std::vector<T> cumsimpson_fn(const FnOfT<T>& f, const std::vector<double>& X, double dx = 1e-5) {
  if (X.empty()) throw IndexDefect("index out of bounds, the container is empty");
  const double lo = *std::min_element(X.begin(), X.end()), hi = *std::max_element(X.begin(), X.end());
  const std::vector<double> t = linspace(lo, hi, nim_to_int((hi - lo) / dx) + 2);
  std::vector<T> dy;
  for (double x : t) dy.push_back(f(x));
  const std::vector<T> ys = cumsimpson<T>(dy, t);
  return hermite_interpolate<T>(X, t, ys, dy);
}

}  // namespace rk_oracle
