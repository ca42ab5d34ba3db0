// emul_main.cpp — the product's kernels, compiled for the host (emul.hpp) and run one emulated thread at a time,
// against the CPU oracle. Every element-wise result must agree BIT FOR BIT; sums within 1e-12 (different order).
// Prints one line per case; exit code = number of failing cases. Built and run by tests/test_kernel_host_emulation.py.
#include "emul.hpp"
#include <cstdlib>
#include "quad_group.hpp"

#include "../../include/b200rk.h"  // host-only planning entry points of the library (b200rk_hermite_plan)

using rk_oracle::Vector;
using namespace b200rk;

static int g_fail = 0, g_cases = 0;
static void report(const std::string& name, bool ok) {
  ++g_cases;
  if (!ok) ++g_fail;
  std::printf("case %s ok=%d\n", name.c_str(), ok ? 1 : 0);
}
static uint64_t g_seed = 0x9E3779B97F4A7C15ull;
static double urand(double lo, double hi) {
  g_seed ^= g_seed << 13; g_seed ^= g_seed >> 7; g_seed ^= g_seed << 17;
  return lo + (hi - lo) * double(g_seed >> 11) / 9007199254740992.0;
}
static std::vector<double> rvec(size_t n, double lo = -1.0, double hi = 1.0) {
  std::vector<double> v(n);
  for (auto& x : v) x = urand(lo, hi);
  return v;
}
static bool same_bits(const std::vector<double>& a, const std::vector<double>& b) {
  return a.size() == b.size() && (a.empty() || std::memcmp(a.data(), b.data(), a.size() * sizeof(double)) == 0);
}
static bool close_rel(double a, double b, double rtol) { return std::fabs(a - b) <= rtol * std::fabs(b) || (a == 0.0 && b == 0.0); }
static ReduceScratch scratch(double* sum) {
  static double partials[64];      // one per emulated CTA (threaded mode: the real two-stage reduction runs)
  static unsigned ticket = 0;
  ReduceScratch rs;
  std::memset(&rs, 0, sizeof(rs));
  rs.result = sum;
  rs.partials = partials;
  rs.ticket = &ticket;
  rs.mail.world = 1;
  return rs;
}
static const size_t kSizes[] = {1, 3, 4, 5, 8, 1023, 2051};
static constexpr int T = 64;  // emulated threads per block (the kernels take it as a template parameter)

// ---- stage accumulate (kernels.cuh: stage_kernel) vs y + c*(w0*k0 + ...) --------------------------------------------
template <int M, int W>
static void test_stage() {
  bool ok = true;
  for (size_t n : kSizes) {
    const auto y = rvec(n);
    std::vector<std::vector<double>> k(M);
    for (auto& v : k) v = rvec(n);
    const auto w = rvec(M, -3.0, 3.0);
    const double c = 0.0123;
    std::vector<double> out(n, -7.0);
    StageArgs<M> a;
    a.y = y.data(); a.c = c; a.out = out.data(); a.n = n;
    for (int j = 0; j < M; ++j) { a.k[j] = k[j].data(); a.w[j] = w[j]; }
    emul_launch(3, T, [&] { stage_kernel<M, W, 2, false, T, 0>(a); });
    std::vector<Vector> kv;
    std::vector<const Vector*> kp;
    for (auto& v : k) kv.emplace_back(v);
    for (auto& v : kv) kp.push_back(&v);
    const Vector ref = Vector(y) + c * rk_oracle::wsum<Vector>(w.data(), kp.data(), M);
    ok = ok && same_bits(out, ref.components);
    if (M == 2) {  // chain form (Kutta3's stage 3, ode.nim:128): ((y + w0*k0) + w1*k1)
      std::vector<double> oc(n), rc(n);
      a.out = oc.data();
      emul_launch(2, T, [&] { stage_kernel<M, W, 1, true, T, 0>(a); });
      for (size_t i = 0; i < n; ++i) rc[i] = (y[i] + k[0][i] * w[0]) + k[1][i] * w[1];
      ok = ok && same_bits(oc, rc);
    }
  }
  report("stage_kernel M=" + std::to_string(M) + " W=" + std::to_string(W), ok);
}

// ---- combine + error (finish_kernel), all tableau terms kept (what the reference multiplies through) ----------------
template <int NK, int W, bool DIRECT>
static void test_finish(const char* name, const rk_oracle::Pair& p) {
  bool ok = true;
  for (size_t n : kSizes) {
    const auto y = rvec(n, 0.5, 1.5);
    std::vector<std::vector<double>> k(NK);
    for (auto& v : k) v = rvec(n);
    const double dt = 0.0371, absTol = 1e-6, relTol = 1e-5;
    std::vector<double> ynew(n), err(n);
    double sum = 0.0;
    FinishArgs<NK> a;
    std::memset(&a, 0, sizeof(a));
    a.y = y.data();
    for (int j = 0; j < NK; ++j) {
      a.k[j] = k[j].data();
      a.wb[j] = j < p.n_b ? p.b[j] : 0.0;
      a.wbh[j] = j < p.n_bhat ? p.bhat[j] : 0.0;
      if (j < p.n_b) a.mask_b |= 1u << j;
      if (j < p.n_bhat) a.mask_bh |= 1u << j;
    }
    a.cb = dt; a.cbh = dt; a.absTol = absTol; a.relTol = relTol;
    a.ynew_out = ynew.data(); a.err_out = err.data(); a.n = n; a.rs = scratch(&sum);
    emul_launch(2, T, [&] { finish_kernel<NK, W, 1, DIRECT, 1, T>(a); });
    std::vector<Vector> kv;
    std::vector<const Vector*> kp;
    for (auto& v : k) kv.emplace_back(v);
    for (auto& v : kv) kp.push_back(&v);
    const Vector yv(y);
    const Vector yN = rk_oracle::pair_y_new<Vector>(p, yv, dt, kp.data());
    const Vector ey = rk_oracle::pair_error_y<Vector>(p, yv, yN, dt, kp.data());
    const Vector tol = rk_oracle::add_scalar(absTol, relTol * rk_oracle::vabs(yN));
    const Vector e1 = rk_oracle::hdiv(ey, tol);
    const double S = rk_oracle::vsum(rk_oracle::hadamard(e1, e1));
    ok = ok && same_bits(ynew, yN.components) && same_bits(err, ey.components) && close_rel(sum, S, 1e-12);
    {  // the software-pipelined form (finish_pf.cuh, experimental): same elements, and the SAME sum bits on the same grid
      std::vector<double> ynew_pf(n), err_pf(n);
      double sum_pf = 0.0;
      a.ynew_out = ynew_pf.data(); a.err_out = err_pf.data(); a.rs = scratch(&sum_pf);
      emul_launch(2, T, [&] { finish_pf_kernel<NK, W, DIRECT, 1, T>(a); });
      ok = ok && same_bits(ynew_pf, yN.components) && same_bits(err_pf, ey.components) && std::memcmp(&sum_pf, &sum, 8) == 0;
      a.ynew_out = ynew.data(); a.err_out = err.data();
    }
    if (DIRECT) {  // Tsit54 as the solver runs it: yNew is loaded instead of y (mode 2), the error row is direct
      std::vector<double> err2(n);
      double sum2 = 0.0;
      a.y = yN.components.data(); a.ynew_out = nullptr; a.err_out = err2.data(); a.rs = scratch(&sum2);
      emul_launch(2, T, [&] { finish_kernel<NK, W, 1, true, 2, T>(a); });
      ok = ok && same_bits(err2, ey.components) && close_rel(sum2, S, 1e-12);
      std::vector<double> err2p(n);
      double sum2p = 0.0;
      a.err_out = err2p.data(); a.rs = scratch(&sum2p);
      emul_launch(2, T, [&] { finish_pf_kernel<NK, W, true, 2, T>(a); });
      ok = ok && same_bits(err2p, ey.components) && std::memcmp(&sum2p, &sum2, 8) == 0;
    } else {      // DOPRI54 as the solver runs it: yNew recomputed in registers, not stored (mode 0)
      std::vector<double> err0(n);
      double sum0 = 0.0;
      a.ynew_out = nullptr; a.err_out = err0.data(); a.rs = scratch(&sum0);
      emul_launch(2, T, [&] { finish_kernel<NK, W, 1, false, 0, T>(a); });
      ok = ok && same_bits(err0, ey.components) && close_rel(sum0, S, 1e-12);
      std::vector<double> err0p(n);
      double sum0p = 0.0;
      a.err_out = err0p.data(); a.rs = scratch(&sum0p);
      emul_launch(2, T, [&] { finish_pf_kernel<NK, W, false, 0, T>(a); });
      ok = ok && same_bits(err0p, ey.components) && std::memcmp(&sum0p, &sum0, 8) == 0;
    }
  }
  report(std::string("finish_kernel + finish_pf_kernel ") + name + " W=" + std::to_string(W), ok);
}

// ---- whole attempt in one kernel (fused_attempt_kernel) vs the oracle's X_step with a first attempt that is accepted -
template <int S>
static void fill_tableau(FusedArgs<S>& a, const rk_oracle::Pair& p) {
  for (int s = 2; s <= S; ++s)
    for (int j = 0; j < s - 1; ++j) a.a[s - 2][j] = p.a[s][j];
  for (int j = 0; j < p.n_b; ++j) a.b[j] = p.b[j];
  for (int j = 0; j < p.n_bhat; ++j) a.bh[j] = p.bhat[j];
  for (int s = 2; s <= S; ++s) a.cnode[s - 1] = p.c[s];
}
template <int PAT, int KIND, int W>
static void test_fused(const char* name, const rk_oracle::Pair& p, bool backward) {
  constexpr int S = Pattern<PAT>::S;
  bool ok = true;
  for (size_t n : kSizes) {
    const auto y = rvec(n, 0.5, 1.5), lam = rvec(n, 0.1, 5.0);
    const double t0 = backward ? -0.3 : 0.3, dt = 0.005, cs[2] = {0.7, 0.05}, scale = -0.37;
    const double ts = backward ? -1.0 : 1.0;  // solver coordinate t := -t on the backward pass; g(t, y) = -f(-t, y) (ode.nim:545)
    rk_oracle::OdeProc<Vector> f;
    if (KIND == PW_SCALE) f = rk_oracle::rhs_scale_vec(scale);
    else if (KIND == PW_DIAG) f = rk_oracle::rhs_diag_linear(Vector(lam));
    else f = [&](double t, const Vector& v, rk_oracle::Context<Vector>*) {
      std::vector<double> r(v.len());
      for (size_t i = 0; i < v.len(); ++i) r[i] = user_rhs(t, v.components[i], &lam[i], cs);
      return Vector(r);
    };
    rk_oracle::OdeProc<Vector> g = f;
    if (backward) g = [&](double t, const Vector& v, rk_oracle::Context<Vector>* c) { return -f(-t, v, c); };
    const Vector yv(y);
    const Vector fsal = g(t0, yv, nullptr);
    const rk_oracle::Options o = rk_oracle::new_options(1e-4, 1.0, 1.0, 1.0, 1e-8);  // loose: the first attempt is accepted
    const auto ref = rk_oracle::pair_step<Vector>(p, g, t0, yv, fsal, dt, o, nullptr);
    if (ref.dt != dt) { ok = false; break; }
    std::vector<double> ynew(n), ks(n);
    double sum = 0.0;
    FusedArgs<S> a;
    std::memset(&a, 0, sizeof(a));
    a.y = y.data(); a.k1 = fsal.components.data(); a.p[0] = lam.data();
    if (KIND == PW_USER) { a.rhs_sign = ts; a.tsign = ts; a.t = t0; a.cs[0] = cs[0]; a.cs[1] = cs[1]; }
    else { a.rhs_scalar = backward ? -scale : scale; a.rhs_sign = backward ? 1.0 : -1.0; }
    fill_tableau(a, p);
    a.dt = dt; a.cb = dt; a.cbh = dt; a.absTol = o.absTol; a.relTol = o.relTol;
    a.ynew = ynew.data(); a.ks_out = ks.data(); a.n = n; a.rs = scratch(&sum);
    emul_launch(2, T, [&] { fused_attempt_kernel<PAT, KIND, W, T, 0>(a); });
    ok = ok && same_bits(ynew, ref.y_new.components) && same_bits(ks, ref.fsal.components) &&
         close_rel(std::sqrt(1.0 / double(n) * sum), ref.error, 1e-12);
  }
  report(std::string("fused_attempt ") + name + " kind=" + std::to_string(KIND) + " W=" + std::to_string(W) + (backward ? " backward" : ""), ok);
}

// ---- RK4: final combine, whole step for built-in and source right-hand sides ----------------------------------------
static void test_rk4() {
  bool ok = true, ok_user = true, ok_rhs = true;
  for (size_t n : kSizes) {
    const auto y = rvec(n, 0.5, 1.5), lam = rvec(n, 0.1, 5.0);
    const double t0 = 0.25, dt = 0.01, cs[2] = {0.7, 0.05};
    const rk_oracle::Options o = rk_oracle::new_options(dt);
    const Vector yv(y);
    const auto ref = rk_oracle::rk4_step<Vector>(rk_oracle::rhs_diag_linear(Vector(lam)), t0, yv, yv, dt, o, nullptr);
    std::vector<double> out(n);
    emul_launch(2, T, [&] { fused_rk4_kernel<PW_DIAG, 4, T>(y.data(), lam.data(), 0.0, -1.0, 0.5 * dt, dt, dt / 6.0, out.data(), n); });
    ok = ok && same_bits(out, ref.y_new.components);
    for (int backward = 0; backward < 2; ++backward) {
      rk_oracle::OdeProc<Vector> f = [&](double t, const Vector& v, rk_oracle::Context<Vector>*) {
        std::vector<double> r(v.len());
        for (size_t i = 0; i < v.len(); ++i) r[i] = user_rhs(t, v.components[i], &lam[i], cs);
        return Vector(r);
      };
      rk_oracle::OdeProc<Vector> g = f;
      if (backward) g = [&](double t, const Vector& v, rk_oracle::Context<Vector>* c) { return -f(-t, v, c); };
      const auto ru = rk_oracle::rk4_step<Vector>(g, t0, yv, yv, dt, o, nullptr);
      UserRhsArgs a;
      std::memset(&a, 0, sizeof(a));
      a.y = y.data(); a.p[0] = lam.data(); a.cs[0] = cs[0]; a.cs[1] = cs[1];
      a.t = t0; a.tsign = backward ? -1.0 : 1.0; a.rsign = backward ? -1.0 : 1.0;
      a.hdt = 0.5 * dt; a.dt = dt; a.c6 = dt / 6.0; a.out = out.data(); a.n = n;
      emul_launch(2, T, [&] { user_rk4_kernel<4, T>(a); });
      ok_user = ok_user && same_bits(out, ru.y_new.components);
    }
    UserRhsArgs r;
    std::memset(&r, 0, sizeof(r));
    r.y = y.data(); r.p[0] = lam.data(); r.cs[0] = cs[0]; r.cs[1] = cs[1]; r.t = 0.625; r.out = out.data(); r.n = n;
    std::vector<double> want(n);
    for (size_t i = 0; i < n; ++i) want[i] = user_rhs(0.625, y[i], &lam[i], cs);
    emul_launch(2, T, [&] { user_rhs_kernel<4, 2, T, 1>(r); });
    ok_rhs = ok_rhs && same_bits(out, want);
    emul_launch(3, T, [&] { user_rhs_kernel<2, 2, T, 0>(r); });
    ok_rhs = ok_rhs && same_bits(out, want);
  }
  report("fused_rk4_kernel diag", ok);
  report("user_rk4_kernel forward+backward", ok_user);
  report("user_rhs_kernel both widths", ok_rhs);
}

// ---- trajectory consumers (quad_kernels.cuh) with the library's own host-side plan -----------------------------------
template <int W>
static void test_quadrature() {
  bool ok_t = true, ok_s = true, ok_h = true, ok_f = true;
  for (int m : {1, 2, 3, 4, 5, 8, 9, 17, 18}) {
    for (size_t n : {size_t(1), size_t(5), size_t(1023)}) {
      std::vector<double> X(m);
      double x = 0.0;
      for (auto& v : X) { x += urand(0.05, 0.4); v = x; }
      std::vector<std::vector<double>> Y(m);
      for (auto& v : Y) v = rvec(n, -2.0, 2.0);
      std::vector<Vector> Yv;
      for (auto& v : Y) Yv.emplace_back(v);
      std::vector<const double*> yp;
      for (auto& v : Y) yp.push_back(v.data());
      {  // cumtrapz
        std::vector<std::vector<double>> out(m, std::vector<double>(n));
        std::vector<double*> op;
        for (auto& v : out) op.push_back(v.data());
        std::vector<double> h;
        for (int k = 0; k + 1 < m; ++k) h.push_back(0.5 * (X[k + 1] - X[k]));
        CumTrapzArgs a{yp.data(), op.data(), h.data(), m, n};
        emul_launch(2, T, [&] { cumtrapz_kernel<W, 4, T>(a); });
        const auto ref = rk_oracle::cumtrapz<Vector>(Yv, X);
        for (int k = 0; k < m; ++k) ok_t = ok_t && same_bits(out[k], ref[k].components);
      }
      if (m >= 3) {  // cumsimpson = Simpson scan onto the knots + Hermite interpolation back onto X
        long N = m;
        const bool even = (N % 2 == 0);
        if (even) N -= 1;
        std::vector<SimpsonStep> steps;
        std::vector<double> xs{X[0]};
        std::vector<const double*> dyp{Y[0].data()};
        for (long i = 0; i < (N - 1) / 2; ++i) {
          const auto w = rk_oracle::simpson_pair_weights(X[2 * i + 1] - X[2 * i], X[2 * i + 2] - X[2 * i + 1]);
          steps.push_back(SimpsonStep{int(2 * i + 2), int(2 * i + 1), int(2 * i), 1, w.alpha, w.beta, w.eta});
          xs.push_back(X[2 * i + 2]); dyp.push_back(Y[2 * i + 2].data());
        }
        if (even) {
          const long last = m - 1;
          const auto w = rk_oracle::simpson_tail_weights(X[last - 1] - X[last - 2], X[last] - X[last - 1]);
          steps.push_back(SimpsonStep{int(last - 2), int(last - 1), int(last), 0, w.eta, w.beta, w.alpha});
          xs.push_back(X[last]); dyp.push_back(Y[last].data());
        }
        std::vector<std::vector<double>> knots(xs.size(), std::vector<double>(n));
        std::vector<double*> kp;
        for (auto& v : knots) kp.push_back(v.data());
        SimpsonScanArgs sa{yp.data(), kp.data(), steps.data(), (int)steps.size(), 0, n};
        emul_launch(2, T, [&] { simpson_scan_kernel<W, T>(sa); });
        std::vector<int> j(m), kind(m);
        std::vector<double> fac(4 * m);
        size_t n_out = 0;
        if (b200rk_hermite_plan(X.data(), m, xs.data(), xs.size(), j.data(), kind.data(), fac.data(), &n_out) != 0) { ok_s = false; continue; }
        std::vector<HermiteOut> plan(n_out);
        for (size_t o = 0; o < n_out; ++o) plan[o] = HermiteOut{j[o], kind[o], fac[4 * o], fac[4 * o + 1], fac[4 * o + 2], fac[4 * o + 3]};
        std::vector<std::vector<double>> out(n_out, std::vector<double>(n));
        std::vector<double*> op;
        for (auto& v : out) op.push_back(v.data());
        std::vector<const double*> kcp(kp.begin(), kp.end());
        HermiteManyArgs ha{kcp.data(), dyp.data(), op.data(), plan.data(), (int)n_out, n};
        emul_launch(2, T, [&] { hermite_many_kernel<W, T>(ha); });
        const auto ref = rk_oracle::cumsimpson<Vector>(Yv, X);
        ok_s = ok_s && ref.size() == n_out;
        for (size_t o = 0; ok_s && o < n_out; ++o) ok_s = same_bits(out[o], ref[o].components);
        // the same call in ONE kernel (simpson_fused_kernel): samples grouped by the knot interval that completes them,
        // here asked for in a scrambled order with a repeat, so the output-slot indirection is exercised too
        std::vector<double> Xq(X.rbegin(), X.rend());
        Xq.push_back(X[m / 2]);
        std::swap(Xq[0], Xq[m / 3]);
        std::vector<int> jq(Xq.size()), kq(Xq.size());
        std::vector<double> fq(4 * Xq.size());
        size_t nq = 0;
        if (b200rk_hermite_plan(Xq.data(), Xq.size(), xs.data(), xs.size(), jq.data(), kq.data(), fq.data(), &nq) != 0) { ok_f = false; continue; }
        const int nsteps = (int)steps.size();
        std::vector<int> tail(nsteps, 0), begin, slot;
        if (even) tail[nsteps - 1] = 1;
        std::vector<HermiteOut> planq(nq), grouped;
        for (size_t o = 0; o < nq; ++o) planq[o] = HermiteOut{jq[o], kq[o], fq[4 * o], fq[4 * o + 1], fq[4 * o + 2], fq[4 * o + 3]};
        group_samples_by_interval(planq, nsteps, &begin, &grouped, &slot);   // the product's own grouping (quad_group.hpp)
        std::vector<std::vector<double>> outf(nq, std::vector<double>(n, -3.0));
        std::vector<double*> opf;
        for (auto& v : outf) opf.push_back(v.data());
        SimpsonFusedArgs fa{yp.data(), opf.data(), steps.data(), tail.data(), begin.data(), grouped.data(), slot.data(), nsteps, 0, n};
        emul_launch(2, T, [&] { simpson_fused_kernel<W, T>(fa); });
        const auto refq = rk_oracle::cumsimpson<Vector>(Yv, X);   // the data set is the same; only the samples differ
        std::vector<Vector> kn;                                   // knots of the oracle: re-derive the samples with its own interpolation
        {
          std::vector<Vector> yk, dk;
          for (size_t q = 0; q < xs.size(); ++q) { yk.emplace_back(knots[q]); dk.emplace_back(std::vector<double>(dyp[q], dyp[q] + n)); }
          kn = rk_oracle::hermite_interpolate<Vector>(Xq, xs, yk, dk);
        }
        ok_f = ok_f && kn.size() == nq && grouped.size() == nq;
        for (size_t o = 0; ok_f && o < nq; ++o) ok_f = same_bits(outf[o], kn[o].components);
      }
      if (m >= 2) {  // hermiteInterpolate at unsorted samples (jumps between intervals, repeats, the end point)
        std::vector<std::vector<double>> dY(m);
        for (auto& v : dY) v = rvec(n);
        std::vector<Vector> dYv;
        for (auto& v : dY) dYv.emplace_back(v);
        std::vector<const double*> dyp;
        for (auto& v : dY) dyp.push_back(v.data());
        std::vector<double> xq;
        for (int q = 0; q < 9; ++q) xq.push_back(urand(X[0], X[m - 1]));
        xq.push_back(X[m - 1]); xq.push_back(X[0]); xq.push_back(xq[2]);
        std::vector<int> j(xq.size()), kind(xq.size());
        std::vector<double> fac(4 * xq.size());
        size_t n_out = 0;
        if (b200rk_hermite_plan(xq.data(), xq.size(), X.data(), m, j.data(), kind.data(), fac.data(), &n_out) != 0) { ok_h = false; continue; }
        std::vector<HermiteOut> plan(n_out);
        for (size_t o = 0; o < n_out; ++o) plan[o] = HermiteOut{j[o], kind[o], fac[4 * o], fac[4 * o + 1], fac[4 * o + 2], fac[4 * o + 3]};
        std::vector<std::vector<double>> out(n_out, std::vector<double>(n));
        std::vector<double*> op;
        for (auto& v : out) op.push_back(v.data());
        HermiteManyArgs ha{yp.data(), dyp.data(), op.data(), plan.data(), (int)n_out, n};
        emul_launch(2, T, [&] { hermite_many_kernel<W, T>(ha); });
        const auto ref = rk_oracle::hermite_interpolate<Vector>(xq, X, Yv, dYv);
        ok_h = ok_h && ref.size() == n_out;
        for (size_t o = 0; ok_h && o < n_out; ++o) ok_h = same_bits(out[o], ref[o].components);
      }
    }
  }
  report("cumtrapz_kernel W=" + std::to_string(W), ok_t);
  report("simpson_scan_kernel + hermite_many_kernel (cumsimpson) W=" + std::to_string(W), ok_s);
  report("hermite_many_kernel unsorted samples W=" + std::to_string(W), ok_h);
  report("simpson_fused_kernel (cumsimpson in one pass) W=" + std::to_string(W), ok_f);
}

// ---- the remaining thread-independent kernels: Vector operators, RK4 combine, Hermite, Lorenz-96, one trapezoid step ----
template <int OP>
static bool ewise_case(const std::vector<double>& a, const std::vector<double>& b, double s, const std::vector<double>& want) {
  bool ok = true;
  const size_t n = a.size();
  std::vector<double> out(n);
  emul_launch(2, T, [&] { ewise_kernel<OP, 2, 2, T, 0>(a.data(), b.data(), s, out.data(), n); });
  ok = ok && same_bits(out, want);
  emul_launch(3, T, [&] { ewise_kernel<OP, 4, 2, T, 1>(a.data(), b.data(), s, out.data(), n); });
  ok = ok && same_bits(out, want);
  std::vector<double> inplace = a;  // out may alias an input (utils operators used in place)
  emul_launch(2, T, [&] { ewise_kernel<OP, 2, 2, T, 0>(inplace.data(), b.data(), s, inplace.data(), n); });
  return ok && same_bits(inplace, want);
}
static void test_elementwise() {
  bool ok_ops = true, ok_rk4 = true, ok_h = true, ok_l = true, ok_tz = true;
  for (size_t n : kSizes) {
    const auto a = rvec(n, -2.0, 2.0), b = rvec(n, 0.5, 3.0);
    const Vector A(a), B(b);
    const double s = -0.37;
    ok_ops = ok_ops && ewise_case<EW_ADD>(a, b, s, (A + B).components) && ewise_case<EW_SUB>(a, b, s, (A - B).components) &&
             ewise_case<EW_HMUL>(a, b, s, rk_oracle::hadamard(A, B).components) && ewise_case<EW_HDIV>(a, b, s, rk_oracle::hdiv(A, B).components) &&
             ewise_case<EW_SCALE>(a, b, s, (s * A).components) && ewise_case<EW_ADD_SCALAR>(a, b, s, rk_oracle::add_scalar(s, A).components) &&
             ewise_case<EW_NEG>(a, b, s, (-A).components) && ewise_case<EW_ABS>(a, b, s, rk_oracle::vabs(A).components) &&
             ewise_case<EW_DIV_SCALAR>(a, b, s, (A / s).components) && ewise_case<EW_NEG_HMUL>(a, b, s, (-rk_oracle::hadamard(A, B)).components);
    // RK4 final combine (ode.nim:188)
    const auto k1 = rvec(n), k2 = rvec(n), k3 = rvec(n), k4 = rvec(n);
    const double dt = 0.0123;
    std::vector<double> out(n);
    emul_launch(2, T, [&] { rk4_final_kernel<4, 1, T>(a.data(), k1.data(), k2.data(), k3.data(), k4.data(), dt / 6.0, out.data(), n); });
    const Vector r4 = A + dt / 6.0 * (Vector(k1) + 2.0 * (Vector(k2) + Vector(k3)) + Vector(k4));
    ok_rk4 = ok_rk4 && same_bits(out, r4.components);
    // hermiteSpline (utils.nim:273-279) with the host-side factors the driver computes
    const double x1 = 0.3, x2 = 0.8, x = 0.4321;
    const double tt = (x - x1) / (x2 - x1), u = 1.0 - tt;
    const double h00 = (1.0 + 2.0 * tt) * (u * u), h10 = tt * (u * u), h01 = (tt * tt) * (3.0 - 2.0 * tt), h11 = (tt * (tt * tt)) - (tt * tt);
    emul_launch(2, T, [&] { hermite_kernel<2, T>(a.data(), k1.data(), b.data(), k2.data(), h00, h10 * (x2 - x1), h01, h11 * (x2 - x1), out.data(), n); });
    ok_h = ok_h && same_bits(out, rk_oracle::hermite_spline<Vector>(x, x1, x2, A, B, Vector(k1), Vector(k2)).components);
    // Lorenz-96 right-hand side, cyclic (single GPU: the neighbours are in the vector itself)
    if (n >= 4) {
      const auto y = rvec(n, 7.0, 9.0);
      emul_launch(2, T, [&] { lorenz96_kernel<T>(y.data(), y.data() + n - 2, y.data(), 8.0, out.data(), n); });
      ok_l = ok_l && same_bits(out, rk_oracle::rhs_lorenz96(8.0)(0.0, Vector(y), nullptr).components);
    }
    // one step of the streaming cumtrapz(f, X, ctx, dx) (integrate.nim:170)
    emul_launch(2, T, [&] { trapz_step_kernel<2, T>(a.data(), k1.data(), k2.data(), 0.5 * 0.01, out.data(), n); });
    ok_tz = ok_tz && same_bits(out, (A + 0.5 * 0.01 * (Vector(k1) + Vector(k2))).components);
  }
  report("ewise_kernel: + - *. /. scalar* +. neg abs /scalar -(a*.b), both widths, in place", ok_ops);
  report("rk4_final_kernel", ok_rk4);
  report("hermite_kernel", ok_h);
  report("lorenz96_kernel", ok_l);
  report("trapz_step_kernel", ok_tz);
}

int main(int argc, char** argv) {
  if (argc > 1) g_seed ^= std::strtoull(argv[1], nullptr, 10) * 0x2545F4914F6CDD1Dull + 1;  // other random inputs, same cases
  test_elementwise();
  test_stage<1, 4>(); test_stage<2, 4>(); test_stage<3, 2>(); test_stage<5, 4>(); test_stage<6, 2>(); test_stage<8, 4>(); test_stage<9, 4>();
  test_finish<7, 4, false>("dopri54", rk_oracle::dopri54_pair());
  test_finish<7, 2, true>("tsit54", rk_oracle::tsit54_pair());
  test_finish<9, 4, false>("vern65", rk_oracle::vern65_pair());
  test_fused<PAT_DOPRI54, PW_DIAG, 4>("dopri54", rk_oracle::dopri54_pair(), false);
  test_fused<PAT_DOPRI54_STRICT, PW_DIAG, 2>("dopri54 strict", rk_oracle::dopri54_pair(), false);
  test_fused<PAT_DOPRI54, PW_SCALE, 4>("dopri54", rk_oracle::dopri54_pair(), true);
  test_fused<PAT_TSIT54, PW_DIAG, 4>("tsit54", rk_oracle::tsit54_pair(), false);
  test_fused<PAT_TSIT54, PW_SCALE, 2>("tsit54", rk_oracle::tsit54_pair(), false);
  test_fused<PAT_VERN65, PW_DIAG, 4>("vern65", rk_oracle::vern65_pair(), false);
  test_fused<PAT_VERN65_STRICT, PW_DIAG, 4>("vern65 strict", rk_oracle::vern65_pair(), true);
  test_fused<PAT_DOPRI54, PW_USER, 4>("dopri54 source rhs", rk_oracle::dopri54_pair(), false);
  test_fused<PAT_TSIT54, PW_USER, 2>("tsit54 source rhs", rk_oracle::tsit54_pair(), true);
  test_fused<PAT_VERN65, PW_USER, 4>("vern65 source rhs", rk_oracle::vern65_pair(), false);
  test_fused<PAT_VERN65_STRICT, PW_USER, 4>("vern65 strict source rhs", rk_oracle::vern65_pair(), true);
  test_rk4();
  test_quadrature<4>();
  test_quadrature<2>();
  std::printf("cases=%d failures=%d\n", g_cases, g_fail);
  return g_fail;
}
