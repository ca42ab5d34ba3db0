// quadrature.cu — the consumers of a trajectory on the far side of the hot path:
// hermiteInterpolate (utils.nim:282-312) and the cumulative quadrature routines built on it — cumtrapz(Y, X)
// (integrate.nim:119-135), cumsimpson(Y, X) (integrate.nim:330-378) and their function variants
// (integrate.nim:138-175, 379-400) — for T = device vector. The host does what is scalar in the reference
// (sorting X, the duplicate rule, which interval each sample falls into, Simpson's coefficient triples, the
// spline's scalar factors — same expressions, same rounding); every O(N) operation runs in quad_kernels.cuh.
#include "internal.hpp"
#include "quad_kernels.cuh"
#include "quad_group.hpp"

#include <map>
#include <numeric>

namespace {

// A host array living in device memory for the duration of one call, allocated and freed in stream order.
template <class T>
struct DeviceTable {
  b200rk_ctx* c;
  T* d = nullptr;
  explicit DeviceTable(b200rk_ctx* ctx) : c(ctx) {}
  int upload(const std::vector<T>& h) {
    if (h.empty()) return B200RK_OK;
    CUDA_TRY(c, cudaMallocAsync((void**)&d, h.size() * sizeof(T), c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));  // pageable source: staged before return
    return B200RK_OK;
  }
  ~DeviceTable() { if (d) cudaFreeAsync(d, c->stream); }
  DeviceTable(const DeviceTable&) = delete;
  DeviceTable& operator=(const DeviceTable&) = delete;
};

// result vectors of one call: handed to the caller on success, back to the pool on failure
struct Outputs {
  std::vector<b200rk_vec*> v;
  bool keep = false;
  ~Outputs() { if (!keep) for (auto* p : v) vec_release(p); }
  int add(b200rk_ctx* c, size_t n_global, b200rk_vec** out) {
    TRY(vec_alloc(c, n_global, out));
    v.push_back(*out);
    return B200RK_OK;
  }
};

unsigned quad_grid(const b200rk_ctx* c, size_t n, int W) { return grid_for(c, n / W, kThreads, 0); }

int count_neq(b200rk_ctx* c, const b200rk_vec* a, const b200rk_vec* b, double* count) {
  ProfScope ps(c, B200RK_K_QUAD, 8.0 * double(a->n_local) * 2);
  unsigned grid = std::min(grid_for(c, a->n_local / 2, kThreads * 4, 0), (unsigned)(c->sm_count * 8));
  TRY(ensure_partials(c, grid));
  neq_count_kernel<2, kThreads><<<grid, kThreads, 0, c->stream>>>(a->d, b->d, a->n_local, reduce_scratch(c));
  CUDA_TRY(c, cudaGetLastError());
  return fetch_global_sum(c, count);
}

int check_series(b200rk_ctx* c, const b200rk_vec* const* Y, size_t m, const char* what) {
  if (!Y) return fail(c, B200RK_EINVAL, std::string(what) + ": null vector list");
  for (size_t i = 0; i < m; ++i) {
    if (!Y[i]) return fail(c, B200RK_EINVAL, std::string(what) + ": null vector");
    if (Y[i]->ctx != c) return fail(c, B200RK_EINVAL, std::string(what) + ": vector belongs to another context");
    TRY(check_same(c, Y[0], Y[i]));
  }
  return B200RK_OK;
}

// sortDataset + removeDuplicates (utils.nim:360-420): sort by (x, input position); of every run of equal x keep
// the first, and raise if any member of the run differs from it in any component (`!=`, so a NaN component
// always counts as different, utils.nim:371-373).
struct Series {
  std::vector<double> x;
  std::vector<const b200rk_vec*> y;
};
int sort_and_trim(b200rk_ctx* c, const b200rk_vec* const* Y, const double* X, size_t m, Series* out) {
  if (m == 0) return fail(c, B200RK_EINVAL, "x is empty!");                          // assert, utils.nim:387
  for (size_t i = 0; i < m; ++i)
    if (X[i] != X[i]) return fail(c, B200RK_EINVAL, "X contains NaN (the reference's sort order is undefined there)");
  std::vector<size_t> idx(m);
  std::iota(idx.begin(), idx.end(), size_t(0));
  std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return X[a] < X[b]; });
  for (size_t i = 0; i < m;) {
    size_t j = i + 1;
    for (; j < m && X[idx[j]] == X[idx[i]]; ++j) {
      double differing = 0.0;
      TRY(count_neq(c, Y[idx[j]], Y[idx[i]], &differing));
      if (differing != 0.0)
        return fail(c, B200RK_EINVAL, "impure y-duplicates was found: the vectors at x = " + std::to_string(X[idx[i]]) + " (positions " +
                                          std::to_string(idx[i]) + " and " + std::to_string(idx[j]) + ") differ");   // utils.nim:373
    }
    out->x.push_back(X[idx[i]]);
    out->y.push_back(Y[idx[i]]);
    i = j;
  }
  return B200RK_OK;
}

// Which spline each sample of hermiteInterpolate(x, t, ..) evaluates, in the order the reference appends its
// results (utils.nim:287-312) — including what it silently drops in the sorted branch and the ValueError of the
// unsorted one.
int hermite_plan(const b200rk_ctx* c, const double* x, size_t nx, const double* t, size_t nt, std::vector<HermiteOut>* plan) {
  if (nx == 0 || nt == 0) return fail(c, B200RK_EINVAL, "index out of bounds, the container is empty");
  const long thigh = (long)nt - 1, xhigh = (long)nx - 1;
  auto spline = [&](double a, long i) {
    HermiteOut o;
    o.j = (int)i; o.kind = 0;
    double f[4];
    hermite_factors(a, t[i], t[i + 1], f);
    o.h00 = f[0]; o.hA = f[1]; o.h01 = f[2]; o.hB = f[3];
    return o;
  };
  const HermiteOut last{(int)thigh, 1, 0.0, 0.0, 0.0, 0.0};   // result.add(y[y.high])
  if (std::is_sorted(x, x + nx)) {
    long xi = 0;
    for (long i = 0; i <= thigh - 1; ++i) {
      while (t[i] <= x[xi] && x[xi] < t[i + 1]) {
        plan->push_back(spline(x[xi], i));
        xi += 1;
        if (xhigh < xi) break;
      }
      if (xhigh < xi) break;
    }
    if (x[xhigh] == t[thigh]) plan->push_back(last);
  } else {
    double tmin = t[0], tmax = t[0];
    for (size_t i = 1; i < nt; ++i) { tmin = std::min(tmin, t[i]); tmax = std::max(tmax, t[i]); }
    // The reference scans the intervals from the left for every sample. For ascending t (every internal caller, and
    // any sensible data set) the first interval with t[i] <= a < t[i+1] is the only one, so it is found by bisection;
    // anything else keeps the literal scan.
    const bool t_sorted = std::is_sorted(t, t + nt);
    for (size_t k = 0; k < nx; ++k) {
      const double a = x[k];
      bool found = false;
      if (t_sorted) {
        const long i = (long)(std::upper_bound(t, t + nt, a) - t) - 1;   // last i with t[i] <= a (NaN: none)
        if (i >= 0 && i <= thigh - 1 && t[i] <= a && a < t[i + 1]) { plan->push_back(spline(a, i)); found = true; }
      } else {
        for (long i = 0; i <= thigh - 1; ++i)
          if (t[i] <= a && a < t[i + 1]) { plan->push_back(spline(a, i)); found = true; break; }
      }
      if (found) continue;
      if (a == t[thigh]) plan->push_back(last);
      else return fail(c, B200RK_EINVAL, std::to_string(a) + " not in interval " + std::to_string(tmin) + " - " + std::to_string(tmax));  // utils.nim:312
    }
  }
  return B200RK_OK;
}

template <class T>
std::vector<const T*> device_ptrs(const std::vector<const b200rk_vec*>& v) {
  std::vector<const T*> p;
  for (auto* x : v) p.push_back(x->d);
  return p;
}
std::vector<double*> device_ptrs_mut(const std::vector<b200rk_vec*>& v) {
  std::vector<double*> p;
  for (auto* x : v) p.push_back(x->d);
  return p;
}

// all samples of one hermiteInterpolate call in one launch
int run_hermite_many(b200rk_ctx* c, const std::vector<HermiteOut>& plan, const std::vector<const b200rk_vec*>& y,
                     const std::vector<const b200rk_vec*>& dy, Outputs* outs) {
  if (plan.empty()) return B200RK_OK;
  const size_t n_global = y[0]->n_global, n = y[0]->n_local;
  for (size_t o = 0; o < plan.size(); ++o) { b200rk_vec* v = nullptr; TRY(outs->add(c, n_global, &v)); }
  if (n == 0) return B200RK_OK;
  DeviceTable<const double*> dy_t(c), y_t(c);
  DeviceTable<double*> out_t(c);
  DeviceTable<HermiteOut> plan_t(c);
  TRY(y_t.upload(device_ptrs<double>(y)));
  TRY(dy_t.upload(device_ptrs<double>(dy)));
  TRY(out_t.upload(device_ptrs_mut(std::vector<b200rk_vec*>(outs->v.end() - plan.size(), outs->v.end()))));
  TRY(plan_t.upload(plan));
  // algorithmic traffic: one write per sample; reads: the distinct (vector, role) pairs the samples touch
  std::vector<char> seen_y(y.size(), 0), seen_dy(y.size(), 0);
  double reads = 0;
  for (auto& p : plan) {
    if (p.kind == 1) { reads += 1; continue; }
    for (int d = 0; d < 2; ++d) {
      if (!seen_y[p.j + d]) { seen_y[p.j + d] = 1; reads += 1; }
      if (!seen_dy[p.j + d]) { seen_dy[p.j + d] = 1; reads += 1; }
    }
  }
  ProfScope ps(c, B200RK_K_QUAD, 8.0 * double(n) * (reads + double(plan.size())));
  HermiteManyArgs a{y_t.d, dy_t.d, out_t.d, plan_t.d, (int)plan.size(), n};
  if (c->vec_width == 4) hermite_many_kernel<4, kThreads><<<quad_grid(c, n, 4), kThreads, 0, c->stream>>>(a);
  else hermite_many_kernel<2, kThreads><<<quad_grid(c, n, 2), kThreads, 0, c->stream>>>(a);
  CUDA_TRY(c, cudaGetLastError());
  return B200RK_OK;
}

// Simpson's coefficient triples on two unequal intervals (integrate.nim:357-359) and on the last interval of an
// even-length data set (integrate.nim:367-369). `^` is Nim's math.`^`: x^2 = x*x, x^3 = x*x*x.
inline double p2(double x) { return x * x; }
inline double p3(double x) { return x * x * x; }
void simpson_pair(double h1, double h2, double* alpha, double* beta, double* eta) {
  *alpha = (2.0 * p3(h2) - p3(h1) + 3.0 * h1 * p2(h2)) / (6.0 * h2 * (h2 + h1));
  *beta = (p3(h2) + p3(h1) + 3.0 * h1 * h2 * (h2 + h1)) / (6.0 * h2 * h1);
  *eta = (2.0 * p3(h1) - p3(h2) + 3.0 * h2 * p2(h1)) / (6.0 * h1 * (h2 + h1));
}
void simpson_tail(double h1, double h2, double* alpha, double* beta, double* eta) {
  *alpha = (2.0 * p2(h2) + 3.0 * h1 * h2) / (6.0 * (h1 + h2));
  *beta = (p2(h2) + 3.0 * h1 * h2) / (6.0 * h1);
  *eta = -(p3(h2)) / (6.0 * h1 * (h1 + h2));
}

void hand_over(Outputs& outs, b200rk_vec** out, size_t* n_out) {
  for (size_t i = 0; i < outs.v.size(); ++i) out[i] = outs.v[i];
  *n_out = outs.v.size();
  outs.keep = true;
}

int cumtrapz_impl(b200rk_ctx* c, const b200rk_vec* const* Y, const double* X, size_t m, Outputs* outs) {
  TRY(check_series(c, Y, m, "cumtrapz"));
  Series s;
  TRY(sort_and_trim(c, Y, X, m, &s));
  const size_t M = s.x.size(), n_global = s.y[0]->n_global, n = s.y[0]->n_local;
  for (size_t k = 0; k < M; ++k) { b200rk_vec* v = nullptr; TRY(outs->add(c, n_global, &v)); }
  if (n == 0) return B200RK_OK;
  std::vector<double> h;
  for (size_t k = 0; k + 1 < M; ++k) h.push_back(0.5 * (s.x[k + 1] - s.x[k]));                    // integrate.nim:134
  DeviceTable<const double*> y_t(c);
  DeviceTable<double*> out_t(c);
  DeviceTable<double> h_t(c);
  TRY(y_t.upload(device_ptrs<double>(s.y)));
  TRY(out_t.upload(device_ptrs_mut(outs->v)));
  TRY(h_t.upload(h));
  ProfScope ps(c, B200RK_K_QUAD, 8.0 * double(n) * 2.0 * double(M));   // every point read once, every integral written once
  CumTrapzArgs a{y_t.d, out_t.d, h_t.d, (int)M, n};
  if (c->vec_width == 4) cumtrapz_kernel<4, 4, kThreads><<<quad_grid(c, n, 4), kThreads, 0, c->stream>>>(a);
  else cumtrapz_kernel<2, 4, kThreads><<<quad_grid(c, n, 2), kThreads, 0, c->stream>>>(a);
  CUDA_TRY(c, cudaGetLastError());
  return B200RK_OK;
}

int cumsimpson_impl(b200rk_ctx* c, const b200rk_vec* const* Y, const double* X, size_t m, Outputs* outs) {
  TRY(check_series(c, Y, m, "cumsimpson"));
  Series s;
  TRY(sort_and_trim(c, Y, X, m, &s));
  long N = (long)s.x.size();
  if (N < 3) return fail(c, B200RK_EINVAL, "X and Y must have at least 3 elements to perform Simpson, use cumtrapz instead");   // integrate.nim:347-348
  const bool evenN = (N % 2 == 0);
  if (evenN) N -= 1;
  const size_t n_global = s.y[0]->n_global, n = s.y[0]->n_local;
  // integrals at every second data point (+ the last point of an even-length set): the knots of the spline
  std::vector<SimpsonStep> steps;
  std::vector<double> xs{s.x[0]};
  std::vector<const b200rk_vec*> dy{s.y[0]};
  const long pairs = (N - 1) / 2;
  for (long i = 0; i < pairs; ++i) {
    SimpsonStep st;
    st.ia = (int)(2 * i + 2); st.ib = (int)(2 * i + 1); st.ic = (int)(2 * i); st.reuse = 1;       // y[2i] was the previous y[2i'+2] (or the first point)
    simpson_pair(s.x[2 * i + 1] - s.x[2 * i], s.x[2 * i + 2] - s.x[2 * i + 1], &st.ca, &st.cb, &st.cc);
    steps.push_back(st);
    xs.push_back(s.x[2 * i + 2]);
    dy.push_back(s.y[2 * i + 2]);
  }
  if (evenN) {
    const long last = (long)s.x.size() - 1;
    double alpha, beta, eta;
    simpson_tail(s.x[last - 1] - s.x[last - 2], s.x[last] - s.x[last - 1], &alpha, &beta, &eta);
    SimpsonStep st;                                                                                // eta*y[last-2] + beta*y[last-1] + alpha*y[last]
    st.ia = (int)(last - 2); st.ib = (int)(last - 1); st.ic = (int)last; st.reuse = 0;
    st.ca = eta; st.cb = beta; st.cc = alpha;
    steps.push_back(st);
    xs.push_back(s.x[last]);
    dy.push_back(s.y[last]);
  }
  // because Simpson uses several input points per integral point, the integral at all input points comes from
  // hermiteInterpolate(X, xs, y, dy) with the ORIGINAL X (integrate.nim:377-378): plan it before any device work
  std::vector<HermiteOut> plan;
  TRY(hermite_plan(c, X, m, xs.data(), xs.size(), &plan));
  if (c->fuse_simpson && n) {
    // single-pass form (knob "fuse_simpson", default): scan + interpolation in one kernel, the knot
    // integrals stay in registers — see simpson_fused_kernel
    std::vector<int> tail(steps.size(), 0), emit_begin, slot;
    if (evenN) tail.back() = 1;
    std::vector<HermiteOut> grouped;
    group_samples_by_interval(plan, (int)steps.size(), &emit_begin, &grouped, &slot);
    for (size_t o = 0; o < plan.size(); ++o) { b200rk_vec* v = nullptr; TRY(outs->add(c, n_global, &v)); }
    if (plan.empty()) return B200RK_OK;
    DeviceTable<const double*> y_t(c);
    DeviceTable<double*> out_t(c);
    DeviceTable<SimpsonStep> step_t(c);
    DeviceTable<int> tail_t(c), begin_t(c), slot_t(c);
    DeviceTable<HermiteOut> emit_t(c);
    TRY(y_t.upload(device_ptrs<double>(s.y)));
    TRY(out_t.upload(device_ptrs_mut(outs->v)));
    TRY(step_t.upload(steps));
    TRY(tail_t.upload(tail));
    TRY(begin_t.upload(emit_begin));
    TRY(emit_t.upload(grouped));
    TRY(slot_t.upload(slot));
    ProfScope ps(c, B200RK_K_QUAD, 8.0 * double(n) * double(s.x.size() + plan.size()));   // every point read once, every sample written once
    SimpsonFusedArgs a{y_t.d, out_t.d, step_t.d, tail_t.d, begin_t.d, emit_t.d, slot_t.d, (int)steps.size(), 0, n};
    if (c->vec_width == 4) simpson_fused_kernel<4, kThreads><<<quad_grid(c, n, 4), kThreads, 0, c->stream>>>(a);
    else simpson_fused_kernel<2, kThreads><<<quad_grid(c, n, 2), kThreads, 0, c->stream>>>(a);
    CUDA_TRY(c, cudaGetLastError());
    return B200RK_OK;
  }
  Outputs knots;   // released at the end of the call (stream-ordered reuse: the pool belongs to this stream)
  for (size_t k = 0; k < xs.size(); ++k) { b200rk_vec* v = nullptr; TRY(knots.add(c, n_global, &v)); }
  if (n) {
    DeviceTable<const double*> y_t(c);
    DeviceTable<double*> node_t(c);
    DeviceTable<SimpsonStep> step_t(c);
    TRY(y_t.upload(device_ptrs<double>(s.y)));
    TRY(node_t.upload(device_ptrs_mut(knots.v)));
    TRY(step_t.upload(steps));
    ProfScope ps(c, B200RK_K_QUAD, 8.0 * double(n) * double(s.x.size() + xs.size()));   // every point read once, every knot written once
    SimpsonScanArgs a{y_t.d, node_t.d, step_t.d, (int)steps.size(), 0, n};
    if (c->vec_width == 4) simpson_scan_kernel<4, kThreads><<<quad_grid(c, n, 4), kThreads, 0, c->stream>>>(a);
    else simpson_scan_kernel<2, kThreads><<<quad_grid(c, n, 2), kThreads, 0, c->stream>>>(a);
    CUDA_TRY(c, cudaGetLastError());
  }
  std::vector<const b200rk_vec*> knot_y(knots.v.begin(), knots.v.end());
  return run_hermite_many(c, plan, knot_y, dy, outs);
}

// cumsimpson(f, X, ctx, dx) without keeping every evaluation alive. The reference's composition
//   dy = f(t);  ys = cumsimpson(dy, t) = hermiteInterpolate(t, knots, I, dy[knots]);  result = hermiteInterpolate(X, t, ys, dy)
// only ever looks two grid points back: the Simpson step of pair j needs f at t[2j], t[2j+1], t[2j+2]; ys[2j] and
// ys[2j+1] are splines on the knot interval j; the samples of X inside [t[k], t[k+1]) need ys and f at k and k+1. So the
// grid is walked once with a window of a few vectors, each launch being one of the kernels the composed form uses
// (stage accumulate for the Simpson step — y + 1.0*((ya*ca + yb*cb) + yc*cc) is the scan's arithmetic bit for bit —,
// the Hermite kernel for both interpolation levels): same results, same order of callbacks, O(1) memory.
// Requires a strictly increasing grid (always the case unless min(X) == max(X)); returns 1 if it cannot apply.
int cumsimpson_fn_streaming(b200rk_ctx* c, b200rk_fn_of_t f, void* user, size_t n_global, const double* X, size_t m,
                            const std::vector<double>& t, Outputs* outs, bool* applied) {
  *applied = false;
  const long nt = (long)t.size();
  if (nt < 3) return B200RK_OK;
  for (long k = 0; k + 1 < nt; ++k) if (!(t[k] < t[k + 1])) return B200RK_OK;
  // knots and Simpson steps of the discrete rule on (f(t), t) (integrate.nim:341-376)
  long N = nt;
  const bool evenN = (N % 2 == 0);
  if (evenN) N -= 1;
  const long pairs = (N - 1) / 2;
  std::vector<SimpsonStep> steps;
  std::vector<double> xs{t[0]};
  std::vector<long> knot_idx{0};
  for (long i = 0; i < pairs; ++i) {
    SimpsonStep st;
    st.ia = (int)(2 * i + 2); st.ib = (int)(2 * i + 1); st.ic = (int)(2 * i); st.reuse = 0;
    simpson_pair(t[2 * i + 1] - t[2 * i], t[2 * i + 2] - t[2 * i + 1], &st.ca, &st.cb, &st.cc);
    steps.push_back(st); xs.push_back(t[2 * i + 2]); knot_idx.push_back(2 * i + 2);
  }
  if (evenN) {
    const long last = nt - 1;
    double alpha, beta, eta;
    simpson_tail(t[last - 1] - t[last - 2], t[last] - t[last - 1], &alpha, &beta, &eta);
    SimpsonStep st;
    st.ia = (int)(last - 2); st.ib = (int)(last - 1); st.ic = (int)last; st.reuse = 0;
    st.ca = eta; st.cb = beta; st.cc = alpha;
    steps.push_back(st); xs.push_back(t[last]); knot_idx.push_back(last);
  }
  std::vector<HermiteOut> inner, outer;   // ys[k] from the knots; the result from (ys, f) on the grid
  TRY(hermite_plan(c, t.data(), (size_t)nt, xs.data(), xs.size(), &inner));
  if ((long)inner.size() != nt) return B200RK_OK;   // cannot happen on an increasing grid; let the composed form decide
  TRY(hermite_plan(c, X, m, t.data(), (size_t)nt, &outer));
  std::vector<size_t> order(outer.size());
  std::iota(order.begin(), order.end(), size_t(0));
  auto need = [&](const HermiteOut& p) { return p.kind == 1 ? nt - 1 : (long)p.j + 1; };   // highest grid point a sample needs
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return need(outer[a]) < need(outer[b]); });
  for (size_t o = 0; o < outer.size(); ++o) { b200rk_vec* v = nullptr; TRY(outs->add(c, n_global, &v)); }
  *applied = true;

  std::map<long, b200rk_vec*> dyv, ysv;   // the window
  struct Window {
    std::map<long, b200rk_vec*>&a, &b;
    b200rk_vec* I[2] = {nullptr, nullptr};
    ~Window() { for (auto& kv : a) vec_release(kv.second); for (auto& kv : b) vec_release(kv.second); for (auto* v : I) if (v) vec_release(v); }
  } win{dyv, ysv};
  TRY(vec_alloc(c, n_global, &win.I[0]));
  TRY(vec_alloc(c, n_global, &win.I[1]));
  const size_t n = win.I[0]->n_local;
  auto eval = [&](long k) -> int {   // f(t[k]), once, in increasing k like the reference's loop (integrate.nim:397-398)
    if (dyv.count(k)) return B200RK_OK;
    b200rk_vec* v = nullptr;
    TRY(vec_alloc(c, n_global, &v));
    dyv[k] = v;
    const int rc = f(t[k], v, user);
    return rc == 0 ? B200RK_OK : fail(c, B200RK_ECALLBACK, "integrand callback returned " + std::to_string(rc));
  };
  size_t next_out = 0;
  long avail = -1;   // ys[0..avail] computed
  auto emit_ready = [&]() -> int {
    for (; next_out < order.size(); ++next_out) {
      const HermiteOut& p = outer[order[next_out]];
      if (need(p) > avail) break;
      b200rk_vec* dst = outs->v[order[next_out]];
      if (p.kind == 1) { TRY(vec_copy_raw(c, dst, ysv.at(nt - 1))); continue; }
      TRY(launch_hermite(c, ysv.at(p.j)->d, dyv.at(p.j)->d, ysv.at(p.j + 1)->d, dyv.at(p.j + 1)->d, p.h00, p.hA, p.h01, p.hB, dst->d, n));
    }
    return B200RK_OK;
  };
  TRY(eval(0));
  int cur = 0;
  TRY(launch_ewise(c, EW_SUB, dyv[0]->d, dyv[0]->d, 0.0, win.I[0]->d, n, B200RK_K_QUAD));   // the right kind of zero (integrate.nim:351)
  long in_k = 0;   // next grid point whose ys is due
  for (size_t j = 0; j < steps.size(); ++j) {
    const SimpsonStep& st = steps[j];
    const long lo_k = std::min({st.ia, st.ib, st.ic}), hi_k = std::max({st.ia, st.ib, st.ic});
    for (long k = lo_k; k <= hi_k; ++k) TRY(eval(k));
    const double* kp[3] = {dyv.at(st.ia)->d, dyv.at(st.ib)->d, dyv.at(st.ic)->d};
    const double w[3] = {st.ca, st.cb, st.cc};
    TRY(launch_stage(c, 3, win.I[cur]->d, kp, w, 1.0, false, win.I[1 - cur]->d, n));   // I_{j+1} = I_j + ((ya*ca + yb*cb) + yc*cc)
    // ys at the grid points of knot interval j
    for (; in_k < nt && inner[in_k].kind == 0 && inner[in_k].j == (int)j; ++in_k) {
      const HermiteOut& p = inner[in_k];
      b200rk_vec* v = nullptr;
      TRY(vec_alloc(c, n_global, &v));
      ysv[in_k] = v;
      TRY(launch_hermite(c, win.I[cur]->d, dyv.at(knot_idx[j])->d, win.I[1 - cur]->d, dyv.at(knot_idx[j + 1])->d, p.h00, p.hA, p.h01, p.hB, v->d, n));
      avail = in_k;
      TRY(emit_ready());
    }
    cur = 1 - cur;
    // drop what no later step, spline or sample can need: everything below the last grid point whose ys exists
    for (auto* mp : {&dyv, &ysv})
      for (auto it = mp->begin(); it != mp->end() && it->first < avail;) { vec_release(it->second); it = mp->erase(it); }
  }
  if (in_k == nt - 1 && inner[in_k].kind == 1) {   // the last grid point is the last knot: its ys is the integral itself
    b200rk_vec* v = nullptr;
    TRY(vec_alloc(c, n_global, &v));
    ysv[in_k] = v;
    TRY(vec_copy_raw(c, v, win.I[cur]));
    avail = in_k; ++in_k;
    TRY(emit_ready());
  }
  if (in_k != nt || next_out != order.size()) return fail(c, B200RK_EINVAL, "cumsimpson(f, X, dx): internal streaming plan mismatch");
  return B200RK_OK;
}

}  // namespace

extern "C" {

// Host only (no device, no context): what the two routines above decide on the host, exposed so that the CPU test-suite
// can pin it against the oracle — which data interval every returned sample of hermiteInterpolate(x, t, ..) uses (or
// that it is the copy of the last data point) with the spline's four scalar factors, and Simpson's coefficient triples.
int b200rk_hermite_plan(const double* x, size_t nx, const double* t, size_t nt, int* interval, int* is_copy, double* factors,
                        size_t* n_out) {
  if (!x || !t || !interval || !is_copy || !factors || !n_out) return fail(nullptr, B200RK_EINVAL, "null argument");
  std::vector<HermiteOut> plan;
  TRY(hermite_plan(nullptr, x, nx, t, nt, &plan));
  for (size_t o = 0; o < plan.size(); ++o) {
    interval[o] = plan[o].j; is_copy[o] = plan[o].kind;
    factors[4 * o] = plan[o].h00; factors[4 * o + 1] = plan[o].hA; factors[4 * o + 2] = plan[o].h01; factors[4 * o + 3] = plan[o].hB;
  }
  *n_out = plan.size();
  return B200RK_OK;
}
int b200rk_simpson_weights(int tail, double h1, double h2, double* alpha, double* beta, double* eta) {
  if (!alpha || !beta || !eta) return fail(nullptr, B200RK_EINVAL, "null argument");
  if (tail) simpson_tail(h1, h2, alpha, beta, eta);
  else simpson_pair(h1, h2, alpha, beta, eta);
  return B200RK_OK;
}

int b200rk_hermite_interpolate(b200rk_ctx* c, const double* x, size_t nx, const double* t, size_t nt, const b200rk_vec* const* y,
                               const b200rk_vec* const* dy, b200rk_vec** out, size_t* n_out) {
  if (!c || !x || !t || !out || !n_out) return fail(c, B200RK_EINVAL, "null argument");
  TRY(check_series(c, y, nt, "hermiteInterpolate"));
  TRY(check_series(c, dy, nt, "hermiteInterpolate"));
  if (nt) TRY(check_same(c, y[0], dy[0]));
  std::vector<HermiteOut> plan;
  TRY(hermite_plan(c, x, nx, t, nt, &plan));
  Outputs outs;
  TRY(run_hermite_many(c, plan, std::vector<const b200rk_vec*>(y, y + nt), std::vector<const b200rk_vec*>(dy, dy + nt), &outs));
  hand_over(outs, out, n_out);
  return B200RK_OK;
}

int b200rk_cumtrapz(b200rk_ctx* c, const b200rk_vec* const* Y, const double* X, size_t m, b200rk_vec** out, size_t* n_out) {
  if (!c || !X || !out || !n_out) return fail(c, B200RK_EINVAL, "null argument");
  Outputs outs;
  TRY(cumtrapz_impl(c, Y, X, m, &outs));
  hand_over(outs, out, n_out);
  return B200RK_OK;
}

int b200rk_cumsimpson(b200rk_ctx* c, const b200rk_vec* const* Y, const double* X, size_t m, b200rk_vec** out, size_t* n_out) {
  if (!c || !X || !out || !n_out) return fail(c, B200RK_EINVAL, "null argument");
  Outputs outs;
  TRY(cumsimpson_impl(c, Y, X, m, &outs));
  hand_over(outs, out, n_out);
  return B200RK_OK;
}

// cumtrapz(f, X, ctx, dx) (integrate.nim:138-175). The reference stores f, the running integral and the time at
// EVERY step (from min(X) to max(X) + 1.0 in steps of dx: 10^5 vectors per unit of time at the default dx) and
// interpolates afterwards. The samples each step interval contributes are known beforehand, so this streams:
// four vectors live, each step = 1 callback + 1 kernel, samples are emitted when their interval is complete.
int b200rk_cumtrapz_fn(b200rk_ctx* c, b200rk_fn_of_t f, void* user, size_t n_global, const double* X, size_t m, double dx,
                       b200rk_vec** out, size_t* n_out) {
  if (!c || !f || !X || !out || !n_out) return fail(c, B200RK_EINVAL, "null argument");
  if (m == 0) return fail(c, B200RK_EINVAL, "index out of bounds, the container is empty");
  if (!(dx > 0.0)) return fail(c, B200RK_EINVAL, "cumtrapz: dx must be > 0 (the reference would never terminate)");
  double lo = X[0], hi = X[0];
  for (size_t i = 0; i < m; ++i) {
    if (X[i] != X[i]) return fail(c, B200RK_EINVAL, "X contains NaN");
    lo = std::min(lo, X[i]); hi = std::max(hi, X[i]);
  }
  const double tEnd = hi + 1.0;                                                                    // integrate.nim:160
  if ((tEnd - lo) / dx > 5e7) return fail(c, B200RK_EINVAL, "cumtrapz: more than 5e7 steps; choose a larger dx");
  std::vector<double> times;
  double t = lo;
  times.push_back(t);
  t += dx;
  while (t <= tEnd) { times.push_back(t); t += dx; }                                               // integrate.nim:167-174
  std::vector<HermiteOut> plan;
  TRY(hermite_plan(c, X, m, times.data(), times.size(), &plan));
  std::vector<size_t> order(plan.size());                                                          // samples by the interval that completes them
  std::iota(order.begin(), order.end(), size_t(0));
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
    return (plan[a].j + (plan[a].kind == 1 ? 0 : 1)) < (plan[b].j + (plan[b].kind == 1 ? 0 : 1));
  });
  Outputs outs;
  for (size_t o = 0; o < plan.size(); ++o) { b200rk_vec* v = nullptr; TRY(outs.add(c, n_global, &v)); }
  Workspace ws(c);
  b200rk_vec *dy[2], *I[2];
  for (int i = 0; i < 2; ++i) { TRY(ws.get(n_global, &dy[i])); TRY(ws.get(n_global, &I[i])); }
  const size_t n = dy[0]->n_local;
  auto call = [&](double tt, b200rk_vec* dst) -> int {
    const int rc = f(tt, dst, user);
    return rc == 0 ? B200RK_OK : fail(c, B200RK_ECALLBACK, "integrand callback returned " + std::to_string(rc));
  };
  int cur = 0;
  TRY(call(times[0], dy[0]));
  TRY(launch_ewise(c, EW_SUB, dy[0]->d, dy[0]->d, 0.0, I[0]->d, n, B200RK_K_QUAD));                // the right kind of zero
  size_t next = 0;
  auto emit_done = [&](long completed_points) -> int {   // samples whose data (points <= completed_points - 1) is ready
    for (; next < order.size(); ++next) {
      const HermiteOut& p = plan[order[next]];
      const long needs = p.j + (p.kind == 1 ? 0 : 1);
      if (needs > completed_points - 1) break;
      b200rk_vec* dst = outs.v[order[next]];
      if (p.kind == 1) { TRY(vec_copy_raw(c, dst, I[cur])); continue; }
      // the interval (j, j+1) is (previous, current)
      TRY(launch_hermite(c, I[1 - cur]->d, dy[1 - cur]->d, I[cur]->d, dy[cur]->d, p.h00, p.hA, p.h01, p.hB, dst->d, n));
    }
    return B200RK_OK;
  };
  for (size_t k = 0; k + 1 < times.size(); ++k) {
    TRY(call(times[k + 1], dy[1 - cur]));
    if (n) {
      ProfScope ps(c, B200RK_K_QUAD, 8.0 * double(n) * 4);
      trapz_step_kernel<2, kThreads><<<quad_grid(c, n, 2), kThreads, 0, c->stream>>>(I[cur]->d, dy[cur]->d, dy[1 - cur]->d, 0.5 * dx, I[1 - cur]->d, n);   // integrate.nim:170
      CUDA_TRY(c, cudaGetLastError());
    }
    cur = 1 - cur;
    TRY(emit_done((long)k + 2));
  }
  TRY(emit_done((long)times.size()));   // a data set of one point: only copies of it can be asked for
  if (next != order.size()) return fail(c, B200RK_EINVAL, "cumtrapz: internal plan mismatch");
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));   // the callback's buffers may go away after return
  hand_over(outs, out, n_out);
  return B200RK_OK;
}

// cumsimpson(f, X, ctx, dx) (integrate.nim:379-400): f on linspace(min X, max X, round((max-min)/dx) + 2), the
// discrete rule on those values, then hermiteInterpolate(X, t, ys, dy). Up to 4096 grid points that fit in memory it is
// composed exactly like the reference (all evaluations alive at once); beyond that the grid is streamed through a
// window of a few vectors (cumsimpson_fn_streaming above) with the same results.
int b200rk_cumsimpson_fn(b200rk_ctx* c, b200rk_fn_of_t f, void* user, size_t n_global, const double* X, size_t m, double dx,
                         b200rk_vec** out, size_t* n_out) {
  if (!c || !f || !X || !out || !n_out) return fail(c, B200RK_EINVAL, "null argument");
  if (m == 0) return fail(c, B200RK_EINVAL, "index out of bounds, the container is empty");
  if (!(dx > 0.0)) return fail(c, B200RK_EINVAL, "cumsimpson: dx must be > 0");
  double lo = X[0], hi = X[0];
  for (size_t i = 0; i < m; ++i) {
    if (X[i] != X[i]) return fail(c, B200RK_EINVAL, "X contains NaN");
    lo = std::min(lo, X[i]); hi = std::max(hi, X[i]);
  }
  const double count = std::round((hi - lo) / dx) + 2.0;                                           // toInt rounds half away from zero
  if (count > 5e7) return fail(c, B200RK_EINVAL, "cumsimpson(f, X, dx): more than 5e7 grid points; choose a larger dx");
  size_t off = 0, len = 0;
  shard_range(n_global, c->rank, c->world, &off, &len);
  size_t free_b = 0, total_b = 0;
  CUDA_TRY(c, cudaMemGetInfo(&free_b, &total_b));
  const bool fits = count * 3.0 * double(std::max<size_t>(len, 4)) * 8.0 <= 0.25 * double(free_b) && count <= 4096;
  const long nt = (long)count;
  std::vector<double> t;                                                                           // linspace, utils.nim:498-507
  const double step = (hi - lo) / double(nt - 1);
  t.push_back(lo);
  for (long i = 1; i <= nt - 2; ++i) t.push_back(lo + step * double(i));
  t.push_back(hi);
  // The composed form below keeps all evaluations alive like the reference (3 vectors per grid point). When that does
  // not fit — or on request (knob "stream_simpson" = 1; 0 = never) — the grid is streamed through a window instead.
  if (c->stream_simpson == 1 || (c->stream_simpson < 0 && !fits)) {
    Outputs souts;
    bool applied = false;
    TRY(cumsimpson_fn_streaming(c, f, user, n_global, X, m, t, &souts, &applied));
    if (applied) {
      CUDA_TRY(c, cudaStreamSynchronize(c->stream));
      hand_over(souts, out, n_out);
      return B200RK_OK;
    }
  }
  if (count * 3.0 * double(std::max<size_t>(len, 4)) * 8.0 > 0.9 * double(free_b))
    return fail(c, B200RK_ENOMEM, "cumsimpson(f, X, dx): " + std::to_string((long long)count) + " grid points of " + std::to_string(len) +
                                      " elements do not fit (3 vectors per point are alive at once, as in the reference); choose a larger dx");
  Outputs dy;   // released when the call returns
  for (double x : t) {
    b200rk_vec* v = nullptr;
    TRY(dy.add(c, n_global, &v));
    const int rc = f(x, v, user);
    if (rc != 0) return fail(c, B200RK_ECALLBACK, "integrand callback returned " + std::to_string(rc));
  }
  std::vector<const b200rk_vec*> dyc(dy.v.begin(), dy.v.end());
  Outputs ys;
  TRY(cumsimpson_impl(c, dyc.data(), t.data(), t.size(), &ys));
  if (ys.v.size() != t.size()) return fail(c, B200RK_EINVAL, "cumsimpson(f, X, dx): the grid collapsed (min(X) == max(X)?)");
  std::vector<HermiteOut> plan;
  TRY(hermite_plan(c, X, m, t.data(), t.size(), &plan));
  Outputs outs;
  TRY(run_hermite_many(c, plan, std::vector<const b200rk_vec*>(ys.v.begin(), ys.v.end()), dyc, &outs));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  hand_over(outs, out, n_out);
  return B200RK_OK;
}

}  // extern "C"
