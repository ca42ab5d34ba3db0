// emul_mt_main.cpp — the kernels whose threads INTERACT, under the threaded host emulation (emul.hpp with EMUL_MT: one
// host thread per CUDA thread, real barriers / shuffles / atomics): the shared-memory Lorenz-96 stage kernel, the
// deterministic two-stage reduction, and the device-resident driver loop run as a one-block cooperative grid — each
// against the CPU oracle. Built a second time with -fsanitize=thread, it is the suite's race detector for the kernels'
// shared- and global-memory traffic; `racy` runs a deliberately broken kernel that ThreadSanitizer must flag.
#ifndef EMUL_MT
#define EMUL_MT 1
#endif
#include "emul.hpp"

using rk_oracle::Vector;
using namespace b200rk;

static int g_fail = 0, g_cases = 0;
static void report(const std::string& name, bool ok) {
  ++g_cases;
  if (!ok) ++g_fail;
  std::printf("case %s ok=%d\n", name.c_str(), ok ? 1 : 0);
}
static uint64_t g_seed = 0xD1B54A32D192ED03ull;
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
static constexpr int T = 64;

struct Scratch {
  std::vector<double> partials = std::vector<double>(256, 0.0);
  unsigned ticket = 0;
  double result = 0.0;
  ReduceScratch rs() {
    ReduceScratch r;
    std::memset(&r, 0, sizeof(r));
    r.partials = partials.data(); r.ticket = &ticket; r.result = &result; r.mail.world = 1;
    return r;
  }
};

// ---- sum(v) and the duplicate test through the real reduction: same bits on every run, any grid ---------------------
static void test_reductions() {
  bool ok = true;
  for (size_t n : {size_t(1), size_t(2), size_t(63), size_t(64), size_t(1000), size_t(5003)}) {
    const auto a = rvec(n);
    auto b = a;
    if (n > 2) { b[n / 2] += 1.0; b[n - 1] = std::nan(""); }
    double want = 0.0;
    for (double x : a) want += x;
    double first = 0.0;
    for (unsigned grid : {1u, 3u, 3u}) {
      Scratch s;
      emul_launch(grid, T, [&] { sum_kernel<2, T>(a.data(), n, s.rs()); });
      ok = ok && close_rel(s.result, want, 1e-12) && s.ticket == 0u;
      if (grid == 3u) { if (first == 0.0) first = s.result; else ok = ok && std::memcmp(&first, &s.result, 8) == 0; }
    }
    Scratch s;
    emul_launch(2, T, [&] { neq_count_kernel<2, T>(a.data(), b.data(), n, s.rs()); });
    ok = ok && s.result == (n > 2 ? 2.0 : 0.0);
  }
  report("sum_kernel / neq_count_kernel (two-stage reduction, deterministic)", ok);
}

// ---- stage accumulate fused with the Lorenz-96 stencil through a shared-memory tile ----------------------------------
static std::vector<double> l96(const std::vector<double>& v, double F) {  // k[i] = ((v[i+1] - v[i-2]) * v[i-1] - v[i]) + F, cyclic
  const size_t n = v.size();
  std::vector<double> k(n);
  for (size_t i = 0; i < n; ++i) k[i] = ((v[(i + 1) % n] - v[(i + n - 2) % n]) * v[(i + n - 1) % n] - v[i]) + F;
  return k;
}
template <int M>
static void test_stage_l96() {
  bool ok = true;
  for (size_t n : {size_t(4), size_t(5), size_t(7), size_t(64), size_t(255), size_t(256), size_t(257), size_t(258), size_t(1000)}) {
    const auto y = rvec(n, 7.0, 9.0);
    std::vector<std::vector<double>> k(M);
    for (auto& v : k) v = rvec(n);
    const auto w = rvec(M, -2.0, 2.0);
    const double c = 0.0173, F = 8.0;
    std::vector<Vector> kv;
    std::vector<const Vector*> kp;
    for (auto& v : k) kv.emplace_back(v);
    for (auto& v : kv) kp.push_back(&v);
    const Vector in_ref = Vector(y) + c * rk_oracle::wsum<Vector>(w.data(), kp.data(), M);
    for (int keep_input = 0; keep_input < 2; ++keep_input)
      for (double sgn : {1.0, -1.0}) {
        std::vector<double> kout(n, -5.0), in_out(n, -5.0);
        StageArgs<M> a;
        a.y = y.data(); a.c = c; a.out = keep_input ? in_out.data() : nullptr; a.n = n;
        for (int j = 0; j < M; ++j) { a.k[j] = k[j].data(); a.w[j] = w[j]; }
        const unsigned grid = (unsigned)((n + T * 4 - 1) / (T * 4));
        emul_launch(grid, T, [&] { stage_l96_kernel<M, T>(a, F, sgn, kout.data()); });
        auto want = l96(in_ref.components, F);
        for (auto& x : want) x = x * sgn;
        ok = ok && same_bits(kout, want) && (!keep_input || same_bits(in_out, in_ref.components));
      }
  }
  report("stage_l96_kernel (shared-memory tile + cyclic halo) M=" + std::to_string(M), ok);
}

// ---- the device-resident driver loop as a one-block cooperative grid vs the oracle's ODESolver -----------------------
template <int PAT, int KIND, int W>
static void test_device_loop(const char* name, const rk_oracle::Pair& p, rk_oracle::IntegratorProc<Vector> step) {
  constexpr int S = Pattern<PAT>::S;
  bool ok = true;
  for (size_t n : {size_t(1), size_t(5), size_t(300)}) {
    const auto y0 = rvec(n, 0.5, 1.5), lam = rvec(n, 0.1, 9.0);
    const double cs[2] = {0.7, 0.05};
    rk_oracle::Options o = rk_oracle::new_options(1e-4, 1e-6, 1e-6, 1.0, 1e-8);
    rk_oracle::OdeProc<Vector> f;
    if (KIND == PW_DIAG) f = rk_oracle::rhs_diag_linear(Vector(lam));
    else f = [&](double t, const Vector& v, rk_oracle::Context<Vector>* c) {
      if (c) c->rhs_evals++;
      std::vector<double> r(v.len());
      for (size_t i = 0; i < v.len(); ++i) r[i] = user_rhs(t, v.components[i], &lam[i], cs);
      return Vector(r);
    };
    rk_oracle::Context<Vector> ctx;
    std::vector<rk_oracle::Context<Vector>::StepRecord> trace;
    ctx.trace = &trace;
    const auto ref = rk_oracle::ode_solver<Vector>(f, Vector(y0), {0.0, 2.0}, o, step, true, double(p.order), true, &ctx);
    // device side: Y[0] = y0, F[0] = f(t0, y0)
    std::vector<double> Y[2] = {y0, std::vector<double>(n)}, Fv[2] = {f(0.0, Vector(y0), nullptr).components, std::vector<double>(n)};
    std::vector<double> partials(2 * 4, 0.0);
    RunState st{}, st_host{};
    st.t = 0.0; st.dt = std::sqrt(o.dtMax * o.dtMin); st.t_end = 2.0; st.cur = 0;
    unsigned long long seq_host = 0;
    RunArgs<S> a;
    std::memset(&a, 0, sizeof(a));
    a.f.p[0] = lam.data();
    if (KIND == PW_USER) { a.f.rhs_sign = 1.0; a.f.tsign = 1.0; a.f.cs[0] = cs[0]; a.f.cs[1] = cs[1]; }
    else a.f.rhs_sign = -1.0;
    for (int s = 2; s <= S; ++s) {
      for (int j = 0; j < s - 1; ++j) a.f.a[s - 2][j] = p.a[s][j];
      a.f.cnode[s - 1] = p.c[s];
    }
    for (int j = 0; j < p.n_b; ++j) a.f.b[j] = p.b[j];
    for (int j = 0; j < p.n_bhat; ++j) a.f.bh[j] = p.bhat[j];
    a.f.absTol = o.absTol; a.f.relTol = o.relTol; a.f.n = n;
    for (int i = 0; i < 2; ++i) { a.Y[i] = Y[i].data(); a.F[i] = Fv[i].data(); }
    a.dtMin = o.dtMin; a.dtMax = o.dtMax; a.inv_order_inner = 1.0 / double(p.order); a.inv_order_outer = 1.0 / double(p.order);
    a.n_global = double(n); a.max_steps = 1ll << 40;
    a.partials = partials.data(); a.state = &st; a.state_host = &st_host; a.seq_host = &seq_host; a.seq = 1;
    std::vector<unsigned long long> mailbox(2 * kMaxPeers * 2, 0ull);   // world == 1: this "GPU's" own mailbox carries the grid-wide sum
    unsigned long long arrive = 0;
    a.mail.world = 1; a.mail.rank = 0; a.mail.box[0] = mailbox.data(); a.arrive = &arrive;
    emul_launch(1, T, [&] { fused_run_kernel<PAT, KIND, W, T>(a); });
    const auto& yend = Y[st_host.cur];
    bool states = yend.size() == ref.y.back().components.size();
    double ymax = 0.0;
    for (double v : ref.y.back().components) ymax = std::max(ymax, std::fabs(v));
    for (size_t i = 0; states && i < n; ++i)
      states = std::fabs(yend[i] - ref.y.back().components[i]) <= 1e-9 * std::fabs(ref.y.back().components[i]) + 1e-13 * ymax;
    ok = ok && states && st_host.status == 0 && st_host.steps == ctx.steps && st_host.attempts == ctx.attempts && st_host.rejected == ctx.rejected &&
         st_host.limiter_hits == ctx.limiter_hits && close_rel(st_host.t, 2.0, 1e-14);
  }
  report(std::string("fused_run_kernel (device-resident loop, one-block grid) ") + name, ok);
}

// ---- a whole Lorenz-96 attempt in one kernel (stencil_attempt.cuh: overlapped tiles, two shared-memory buffers, one
// barrier per stage) vs the oracle's IntegratorProc: yNew and the new FSAL bit for bit, error norm through the real
// two-stage reduction ---------------------------------------------------------------------------------------------------
template <int PAT>
static void test_l96_attempt(const char* name, const rk_oracle::Pair& p, rk_oracle::IntegratorProc<Vector> step) {
  constexpr int S = Pattern<PAT>::S, J = 2;
  constexpr int OUT = 2 * J * T - StencilTile<S>::HL - StencilTile<S>::HR;
  bool ok = true;
  int splits = 0, in_place = 0;
  for (size_t n : {size_t(4), size_t(5), size_t(7), size_t(40), size_t(OUT - 1), size_t(OUT), size_t(OUT + 1), size_t(2 * OUT + 3), size_t(1000)})
    for (double sgn : {1.0, -1.0}) {
      const double F = 8.0, dt = 0.004;
      const auto y = rvec(n, 7.0, 9.0);
      rk_oracle::OdeProc<Vector> f = rk_oracle::rhs_lorenz96(F);
      if (sgn < 0) f = [F](double t, const Vector& v, rk_oracle::Context<Vector>* c) { return -rk_oracle::rhs_lorenz96(F)(-t, v, c); };  // ode.nim:545
      const Vector fsal = f(0.0, Vector(y), nullptr);
      const rk_oracle::Options o = rk_oracle::new_options(1e-4, 1e-3, 1e-3, 1.0, 1e-8);
      rk_oracle::Context<Vector> ctx;
      const auto ref = step(f, 0.0, Vector(y), fsal, dt, o, &ctx);
      std::vector<double> ynew(n, -5.0), ks(n, -5.0);
      Scratch sc;
      L96AttemptArgs<S> a;
      std::memset(&a, 0, sizeof(a));
      a.f.y = y.data(); a.f.k1 = fsal.components.data();
      for (int s = 2; s <= S; ++s)
        for (int j = 0; j < s - 1; ++j) a.f.a[s - 2][j] = p.a[s][j];
      for (int j = 0; j < p.n_b; ++j) a.f.b[j] = p.b[j];
      for (int j = 0; j < p.n_bhat; ++j) a.f.bh[j] = p.bhat[j];
      a.f.dt = dt; a.f.cb = dt; a.f.cbh = dt; a.f.absTol = o.absTol; a.f.relTol = o.relTol;
      a.f.ynew = ynew.data(); a.f.ks_out = ks.data(); a.f.n = n; a.f.rs = sc.rs();
      a.F = F;
      const unsigned grid = (unsigned)((n + OUT - 1) / OUT);
      if (sgn < 0) emul_launch(grid, T, [&] { l96_attempt_kernel<PAT, J, T, true>(a); });
      else emul_launch(grid, T, [&] { l96_attempt_kernel<PAT, J, T, false>(a); });
      const double err = std::sqrt(1.0 / double(n) * sc.result);
      ok = ok && ref.dt == dt && same_bits(ynew, ref.y_new.components) && same_bits(ks, ref.fsal.components) && close_rel(err, ref.error, 1e-13);
      // the same attempt SHARDED: G contiguous blocks of the ring, each launched on its own with the HL / HR elements of
      // y and k1 around the block as halos (what executor.cu's exchange_attempt_halo delivers from the ring neighbours)
      constexpr int HL = StencilTile<S>::HL, HR = StencilTile<S>::HR;
      for (size_t G : {size_t(2), size_t(3)}) {
        const size_t chunk = ((n + G - 1) / G + 3) / 4 * 4;           // runtime.cu: shard_range
        if (n < chunk * (G - 1) + HL || chunk < (size_t)HL) continue;  // every shard at least HL long
        std::vector<double> yn2(n, -5.0), ks2(n, -5.0);
        double s2 = 0.0;
        ++splits;
        for (size_t r = 0; r < G; ++r) {
          const size_t lo = r * chunk, len = std::min(n, lo + chunk) - lo;
          std::vector<double> hy(HL + HR), hk(HL + HR);
          for (int i = 0; i < HL; ++i) { hy[i] = y[(lo + n - HL + i) % n]; hk[i] = fsal.components[(lo + n - HL + i) % n]; }
          for (int i = 0; i < HR; ++i) { hy[HL + i] = y[(lo + len + i) % n]; hk[HL + i] = fsal.components[(lo + len + i) % n]; }
          Scratch sc2;
          L96AttemptArgs<S> b = a;
          b.f.y = y.data() + lo; b.f.k1 = fsal.components.data() + lo; b.f.ynew = yn2.data() + lo; b.f.ks_out = ks2.data() + lo;
          b.f.n = len; b.f.rs = sc2.rs(); b.halo = L96Halo{hy.data(), hy.data() + HL, hk.data(), hk.data() + HL};
          if (sgn < 0) emul_launch((unsigned)((len + OUT - 1) / OUT), T, [&] { l96_attempt_kernel<PAT, J, T, true>(b); });
          else emul_launch((unsigned)((len + OUT - 1) / OUT), T, [&] { l96_attempt_kernel<PAT, J, T, false>(b); });
          s2 += sc2.result;                                             // the all-reduce of the shards' partial sums
          // the same shard with its halo read IN PLACE from the neighbours' vectors (what the peer-mapped form does over
          // NVLink; here all shards live in one array): left tail = the HL elements before the block, right head = after it
          const size_t lt = (lo + n - HL) % n, rh = (lo + len) % n;
          if (sgn > 0 && lt + HL <= n && rh + HR <= n) {
            std::vector<double> yn3(len, -5.0), ks3(len, -5.0);
            Scratch sc3;
            L96AttemptArgs<S> d = b;
            d.f.ynew = yn3.data(); d.f.ks_out = ks3.data(); d.f.rs = sc3.rs();
            d.halo = L96Halo{y.data() + lt, y.data() + rh, fsal.components.data() + lt, fsal.components.data() + rh};
            emul_launch((unsigned)((len + OUT - 1) / OUT), T, [&] { l96_attempt_kernel<PAT, J, T, false>(d); });
            ok = ok && std::memcmp(yn3.data(), yn2.data() + lo, len * sizeof(double)) == 0 &&
                 std::memcmp(ks3.data(), ks2.data() + lo, len * sizeof(double)) == 0 && sc3.result == sc2.result;
            ++in_place;
          }
        }
        ok = ok && same_bits(yn2, ref.y_new.components) && same_bits(ks2, ref.fsal.components) &&
             close_rel(std::sqrt(1.0 / double(n) * s2), ref.error, 1e-13);
      }
    }
  report(std::string("l96_attempt_kernel (whole attempt, overlapped tiles; ") + std::to_string(splits) + " sharded splits, " + std::to_string(in_place) + " shards with the halo read in place) " + name, ok && splits >= 16 && in_place >= 16);
}

// ---- a whole RK4 step with the Lorenz-96 stencil in one kernel (l96_rk4_kernel) vs the oracle's RK4_step: bit for bit,
// the ring as one block and shard by shard with halos ------------------------------------------------------------------------
static void test_l96_rk4() {
  constexpr int J = 2, HL = 8, HR = 4, OUT = 2 * J * T - HL - HR;
  bool ok = true;
  int splits = 0;
  for (size_t n : {size_t(4), size_t(5), size_t(7), size_t(40), size_t(OUT - 1), size_t(OUT), size_t(OUT + 1), size_t(2 * OUT + 3), size_t(1000)})
    for (double sgn : {1.0, -1.0}) {
      const double F = 8.0, dt = 0.004;
      const auto y = rvec(n, 7.0, 9.0);
      rk_oracle::OdeProc<Vector> f = rk_oracle::rhs_lorenz96(F);
      if (sgn < 0) f = [F](double t, const Vector& v, rk_oracle::Context<Vector>* c) { return -rk_oracle::rhs_lorenz96(F)(-t, v, c); };  // ode.nim:545
      const rk_oracle::Options o = rk_oracle::new_options(dt, 1e-3, 1e-3, 1.0, 1e-8);
      rk_oracle::Context<Vector> ctx;
      const auto ref = rk_oracle::rk4_step<Vector>(f, 0.0, Vector(y), Vector(y), dt, o, &ctx);
      auto run = [&](const L96Rk4Args& a, unsigned grid) {
        if (sgn < 0) emul_launch(grid, T, [&] { l96_rk4_kernel<J, T, true>(a); });
        else emul_launch(grid, T, [&] { l96_rk4_kernel<J, T, false>(a); });
      };
      std::vector<double> ynew(n, -5.0);
      L96Rk4Args a;
      std::memset(&a, 0, sizeof(a));
      a.y = y.data(); a.ynew = ynew.data(); a.n = n; a.F = F; a.hdt = 0.5 * dt; a.dt = dt; a.c6 = dt / 6.0;
      run(a, (unsigned)((n + OUT - 1) / OUT));
      ok = ok && same_bits(ynew, ref.y_new.components);
      for (size_t G : {size_t(2), size_t(3)}) {
        const size_t chunk = ((n + G - 1) / G + 3) / 4 * 4;           // runtime.cu: shard_range
        if (n < chunk * (G - 1) + HL || chunk < (size_t)HL) continue;
        std::vector<double> yn2(n, -5.0);
        ++splits;
        for (size_t r = 0; r < G; ++r) {
          const size_t lo = r * chunk, len = std::min(n, lo + chunk) - lo;
          std::vector<double> hy(HL + HR);
          for (int i = 0; i < HL; ++i) hy[i] = y[(lo + n - HL + i) % n];
          for (int i = 0; i < HR; ++i) hy[HL + i] = y[(lo + len + i) % n];
          L96Rk4Args b = a;
          b.y = y.data() + lo; b.ynew = yn2.data() + lo; b.n = len; b.halo = L96Halo{hy.data(), hy.data() + HL, nullptr, nullptr};
          run(b, (unsigned)((len + OUT - 1) / OUT));
        }
        ok = ok && same_bits(yn2, ref.y_new.components);
      }
    }
  report("l96_rk4_kernel (whole RK4 step, overlapped tiles; " + std::to_string(splits) + " sharded splits)", ok && splits >= 16);
}

// ---- positive control for the race detector: a tile kernel with its barrier removed ---------------------------------
template <int THREADS>
__global__ void racy_tile_kernel(const double* in, double* out) {
  __shared__ double tile[THREADS];
  tile[threadIdx.x] = in[threadIdx.x];
  // (missing __syncthreads)
  out[threadIdx.x] = tile[(threadIdx.x + 1) % THREADS];
}

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "racy") {
    const auto in = rvec(T);
    std::vector<double> out(T);
    emul_launch(1, T, [&] { racy_tile_kernel<T>(in.data(), out.data()); });
    std::printf("racy control ran\n");
    return 0;
  }
  test_reductions();
  test_stage_l96<1>();
  test_stage_l96<3>();
  test_stage_l96<6>();
  test_l96_attempt<PAT_DOPRI54>("dopri54", rk_oracle::dopri54_pair(), &rk_oracle::dopri54_step<Vector>);
  test_l96_attempt<PAT_TSIT54>("tsit54", rk_oracle::tsit54_pair(), &rk_oracle::tsit54_step<Vector>);
  test_l96_attempt<PAT_VERN65>("vern65", rk_oracle::vern65_pair(), &rk_oracle::vern65_step<Vector>);
  test_l96_rk4();
  test_device_loop<PAT_DOPRI54, PW_DIAG, 2>("dopri54 diag W=2", rk_oracle::dopri54_pair(), &rk_oracle::dopri54_step<Vector>);
  test_device_loop<PAT_TSIT54, PW_DIAG, 4>("tsit54 diag W=4", rk_oracle::tsit54_pair(), &rk_oracle::tsit54_step<Vector>);
  test_device_loop<PAT_VERN65, PW_DIAG, 2>("vern65 diag W=2", rk_oracle::vern65_pair(), &rk_oracle::vern65_step<Vector>);
  test_device_loop<PAT_DOPRI54, PW_USER, 4>("dopri54 source rhs W=4", rk_oracle::dopri54_pair(), &rk_oracle::dopri54_step<Vector>);
  std::printf("cases=%d failures=%d\n", g_cases, g_fail);
  return g_fail;
}
