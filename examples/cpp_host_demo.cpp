// cpp_host_demo.cpp — the reference's own Vector ODE tests (tests/test_ode.nim:139-197), written against the C++
// host mirror include/numericalnim_b200.hpp the way they are written in Nim against numericalnim:
// right-hand side as a closure `-0.1 * y`, y0 = [1,1,1], tspan = linspace(-10, 10, 100), `check t == tspan`,
// `check isClose(val, correct, tol)`. Prints one line per case plus y(10) as hex floats for the pytest wrapper.
//   g++ -std=c++17 -O2 -Iinclude examples/cpp_host_demo.cpp -Lnumericalnim_b200/lib -lb200rk -o cpp_host_demo
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
  const ODEoptions ooVector = newODEoptions(1e-2, 1e-4, 1e-8);                                   // relTol=1e-8, dt=1e-2 (tests/test_ode.nim:10)
  struct Case { const char* integrator; bool oo; double tol; };
  const Case cases[] = {{"dopri54", false, 1e-4}, {"dopri54", true, 1e-8}, {"rk4", true, 1e-8}, {"heun2", true, 1e-5},
                        {"tsit54", false, 1e-4}, {"tsit54", true, 1e-8}, {"vern65", false, 1e-4}, {"vern65", true, 1e-8}};
  int failures = 0;
  for (const Case& c : cases) {
    Solution sol = solveODE(fVector, y0, tspan, c.oo ? ooVector : newODEoptions(), nullptr, c.integrator);
    bool ok = (sol.t == tspan) && sol.y.size() == tspan.size();
    double worst = 0.0;
    for (size_t i = 0; ok && i < sol.y.size(); ++i) {
      const double e = meanSquaredError(sol.y[i].components(), std::exp(-0.1 * sol.t[i]));
      if (e > worst) worst = e;
      if (!(e <= c.tol)) ok = false;  // isClose (utils.nim:474-479)
    }
    const std::vector<double> last = sol.y.back().components();
    std::printf("case integrator=%s options=%s tol=%g ok=%d worst=%.3e steps=%lld rhs_evals=%lld y10=%a\n", c.integrator,
                c.oo ? "ooVector" : "default", c.tol, ok ? 1 : 0, worst, (long long)sol.stats.steps, (long long)sol.stats.rhs_evals, last[0]);
    failures += ok ? 0 : 1;
  }
  // error behaviour of the reference: ValueError on an unknown integrator and on bad options
  int raised = 0;
  try { solveODE(fVector, y0, tspan, newODEoptions(), nullptr, "rk5"); } catch (const ValueError& e) { raised += std::string(e.what()) == "rk5 is not a valid integrator"; }
  try { newODEoptions(1e-4, 1e-4, 1e-4, 1e-5, 1e-4); } catch (const ValueError&) { raised += 1; }
  try { (void)(newVector({1.0, 2.0}) + newVector({1.0, 2.0, 3.0})); } catch (const ValueError&) { raised += 1; }
  // an exception thrown inside the right-hand side comes back to the caller unchanged
  try {
    solveODE([](double, const GpuVector&, NumContext&) -> GpuVector { throw std::domain_error("boom"); }, y0, {0.0, 1.0});
  } catch (const std::domain_error&) { raised += 1; }
  std::printf("errors raised=%d of 4\n", raised);
  return (failures == 0 && raised == 4) ? 0 : 1;
}
