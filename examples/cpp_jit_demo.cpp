// cpp_jit_demo.cpp — a right-hand side handed over as SOURCE through the C++ host mirror (numericalnim::JitRhs):
// compiled at run time into the fused kernels (one kernel per RK4 step / per adaptive attempt instead of one per Vector
// operator); fixed-step results are bit-identical to the same right-hand side written as a closure.
//   g++ -std=c++17 -O2 -Iinclude examples/cpp_jit_demo.cpp -Lnumericalnim_b200/lib -lb200rk -o cpp_jit_demo
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
  const ODEProc fVector = [](double, const GpuVector& y, NumContext&) { return -0.1 * y; };  // tests/test_ode.nim:6
  const std::vector<double> tspan = linspace(-10.0, 10.0, 100);                                  // tests/test_ode.nim:15
  const GpuVector y0 = newVector({1.0, 1.0, 1.0});
  const ODEoptions ooVector = newODEoptions(1e-2, 1e-4, 1e-8);
  int jit_ok = 0;
  {
    JitRhs fJit(y0.device(), "c0*y", {}, {-0.1});
    Solution a = solveODE(fVector, y0, tspan, ooVector, nullptr, "rk4");
    Solution b = solveODE(fJit, y0, tspan, ooVector, "rk4");
    bool same = a.t == b.t && a.y.size() == b.y.size();
    for (size_t i = 0; same && i < a.y.size(); ++i) same = a.y[i].components() == b.y[i].components();
    Solution d = solveODE(fJit, y0, tspan, newODEoptions(), "dopri54");
    bool close = d.t == tspan;
    for (size_t i = 0; close && i < d.y.size(); ++i) close = meanSquaredError(d.y[i].components(), std::exp(-0.1 * d.t[i])) <= 1e-4;
    int bad_expr = 0;
    try { JitRhs bad(y0.device(), "c0*z", {}, {1.0}); } catch (const ValueError&) { bad_expr = 1; }
    jit_ok = same && close && bad_expr;
    std::printf("jit same_bits_as_closure=%d dopri54_ok=%d bad_expression_raises=%d launches_closure=%lld launches_jit=%lld\n", same ? 1 : 0,
                close ? 1 : 0, bad_expr, (long long)a.stats.launches, (long long)b.stats.launches);
  }
  return jit_ok ? 0 : 1;
}
