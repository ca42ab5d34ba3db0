// cpp_quadrature_demo.cpp — the consumers of a trajectory through the C++ host mirror: solve
// y' = -0.1 y like tests/test_ode.nim:139-197, integrate the trajectory over time with cumtrapz / cumsimpson
// (integrate.nim:119-135, 330-378), and the function variant with a closure integrand as tests/test_integrate.nim:6
// writes fVector.
//   g++ -std=c++17 -O2 -Iinclude examples/cpp_quadrature_demo.cpp -Lnumericalnim_b200/lib -lb200rk -o cpp_quadrature_demo
#include <cmath>
#include <cstdio>

#include "numericalnim_b200.hpp"

using namespace numericalnim;

static double meanSquaredError(const std::vector<double>& a, double c) {  // utils.nim:252: norm(v1 - v2) / len
  double s = 0.0;
  for (double x : a) s += (x - c) * (x - c);
  return std::sqrt(s) / double(a.size());
}

int main() {
  const ODEProc fVector = [](double, const GpuVector& y, NumContext&) { return -0.1 * y; };
  const GpuVector y0 = newVector({1.0, 1.0, 1.0});
  const ODEoptions ooVector = newODEoptions(1e-2, 1e-4, 1e-8);
  int quad_ok = 0;
  {
  // The consumers on the far side: integrate the solved trajectory over time (tests/test_integrate.nim:67-95 with the
  // trajectory y(t) = exp(-0.1 t) from above: cumulative integral 10 (1 - exp(-0.1 t)) + const).
    Solution sol = solveODE(fVector, y0, linspace(0.0, 10.0, 41), ooVector, nullptr, "tsit54");
    const std::vector<GpuVector> It = cumtrapz(sol.y, sol.t), Is = cumsimpson(sol.y, sol.t);
    bool ok = It.size() == 41 && Is.size() == 41;
    for (size_t i = 0; ok && i < 41; ++i) {
      const double exact = 10.0 * (1.0 - std::exp(-0.1 * sol.t[i]));
      ok = meanSquaredError(It[i].components(), exact) <= 1e-2 && meanSquaredError(Is[i].components(), exact) <= 1e-5;
    }
    // cumtrapz(f, X, ctx, dx) with a closure integrand, as tests/test_integrate.nim:6 writes fVector
    auto ctx = newNumContext();
    ctx->tValues.emplace("a", newVector({2.0, 2.0, 2.0}));
    const NumContextProc g = [](double x, NumContext& c) { return std::cos(x) * c.tValues.at("a"); };
    const std::vector<double> X = linspace(0.0, 4.71238898038469, 17);
    const std::vector<GpuVector> G = cumsimpson(g, X, y0, ctx, 0.01);
    bool ok2 = G.size() == 17;
    for (size_t i = 0; ok2 && i < 17; ++i) ok2 = meanSquaredError(G[i].components(), 2.0 * std::sin(X[i])) <= 1e-3;
    quad_ok = ok && ok2;
    std::printf("quadrature trajectory_ok=%d function_variant_ok=%d\n", ok ? 1 : 0, ok2 ? 1 : 0);
  }
  return quad_ok ? 0 : 1;
}
