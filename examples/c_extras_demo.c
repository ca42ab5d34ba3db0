/* c_extras_demo.c — the newer entry points of the C-ABI from plain C99 (what the Nim {.importc, cdecl.} shim binds):
 * a right-hand side handed over as SOURCE (b200rk_jit_rhs_new; a stencil: b200rk_jit_stencil_rhs_new) and the consumers of the
 * trajectory (b200rk_cumsimpson, b200rk_hermite_interpolate). Logistic growth  y' = r y (1 - y/K)  with a per-element carrying
 * capacity K[i]; checks the solution, its cumulative integral and an interpolated state against the closed forms.
 *   gcc -std=c99 -O2 -Iinclude examples/c_extras_demo.c -Lnumericalnim_b200/lib -lb200rk -lm -o c_extras_demo
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "b200rk.h"

#define CHECK(call)                                                                      \
  do {                                                                                   \
    int rc_ = (call);                                                                    \
    if (rc_ != B200RK_OK) {                                                              \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, b200rk_last_error(ctx));       \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

#define NT 21

static double logistic(double t, double y0, double K, double r) { return K / (1.0 + (K / y0 - 1.0) * exp(-r * t)); }
static double logistic_integral(double t, double y0, double K, double r) { /* int_0^t y ds */
  const double A = K / y0 - 1.0;
  return K / r * log((exp(r * t) + A) / (1.0 + A));
}

int main(void) {
  const size_t n = 1000;
  const double r = 0.7, y0v = 0.5;
  b200rk_ctx* ctx = NULL;
  CHECK(b200rk_init(&ctx, 0));
  double* K_h = (double*)malloc(n * sizeof(double));
  double* y0_h = (double*)malloc(n * sizeof(double));
  double* buf = (double*)malloc(n * sizeof(double));
  for (size_t i = 0; i < n; ++i) { K_h[i] = 2.0 + 3.0 * (double)i / (double)(n - 1); y0_h[i] = y0v; }
  b200rk_vec *K = NULL, *y0 = NULL;
  CHECK(b200rk_vec_new(ctx, n, &K));
  CHECK(b200rk_vec_upload(K, K_h));
  CHECK(b200rk_vec_new(ctx, n, &y0));
  CHECK(b200rk_vec_upload(y0, y0_h));

  /* the closure  f(t, y) = r*y*(1 - y/K)  as source: compiled at run time into the fused kernels */
  b200rk_rhs_fn f = NULL;
  void* user = NULL;
  const b200rk_vec* params[1];
  const double scalars[1] = {r};
  params[0] = K;
  CHECK(b200rk_jit_rhs_new(ctx, "c0*y*(1.0 - y/p0)", 1, params, 1, scalars, &f, &user));
  /* a typo is a ValueError with the compiler log */
  {
    b200rk_rhs_fn f2 = NULL;
    void* user2 = NULL;
    const int rc = b200rk_jit_rhs_new(ctx, "c0*yy", 0, NULL, 1, scalars, &f2, &user2);
    printf("bad_expression_rc=%d is_einval=%d\n", rc, rc == B200RK_EINVAL);
    if (rc != B200RK_EINVAL) return 3;
  }

  b200rk_options opt;
  CHECK(b200rk_options_new(&opt, 1e-2, 1e-9, 1e-9, 0.05, 1e-8, 4.0, 0.1, 0.0));
  double tspan[NT], t_out[NT];
  for (int k = 0; k < NT; ++k) tspan[k] = 0.1 * (double)k;
  b200rk_vec* ys[NT];
  size_t n_out = 0;
  b200rk_stats st;
  int method = 0;
  CHECK(b200rk_method_from_name("tsit54", &method));
  CHECK(b200rk_solve(ctx, method, f, user, y0, tspan, NT, &opt, t_out, ys, &n_out, &st));
  if (n_out != NT) { fprintf(stderr, "n_out=%zu\n", n_out); return 4; }

  double err_y = 0.0, err_i = 0.0, err_h = 0.0;
  CHECK(b200rk_vec_download(ys[NT - 1], buf));
  for (size_t i = 0; i < n; ++i) err_y = fmax(err_y, fabs(buf[i] - logistic(t_out[NT - 1], y0v, K_h[i], r)));

  /* cumulative integral of the trajectory (integrate.nim:330-378) */
  b200rk_vec* I[NT];
  size_t n_i = 0;
  CHECK(b200rk_cumsimpson(ctx, (const b200rk_vec* const*)ys, t_out, NT, I, &n_i));
  if (n_i != NT) { fprintf(stderr, "n_i=%zu\n", n_i); return 5; }
  for (int k = 0; k < NT; k += 5) {
    CHECK(b200rk_vec_download(I[k], buf));
    for (size_t i = 0; i < n; ++i) err_i = fmax(err_i, fabs(buf[i] - logistic_integral(t_out[k], y0v, K_h[i], r)));
  }

  /* state between the output times: hermiteInterpolate(x, t, y, dy) with dy = f(t, y) evaluated by the same callback */
  b200rk_vec* dys[NT];
  for (int k = 0; k < NT; ++k) {
    CHECK(b200rk_vec_new(ctx, n, &dys[k]));
    if (f(t_out[k], ys[k], dys[k], user) != 0) return 6;
  }
  const double xq[2] = {0.4321, 1.775};
  b200rk_vec* H[2];
  size_t n_h = 0;
  CHECK(b200rk_hermite_interpolate(ctx, xq, 2, t_out, NT, (const b200rk_vec* const*)ys, (const b200rk_vec* const*)dys, H, &n_h));
  if (n_h != 2) { fprintf(stderr, "n_h=%zu\n", n_h); return 7; }
  for (int q = 0; q < 2; ++q) {
    CHECK(b200rk_vec_download(H[q], buf));
    for (size_t i = 0; i < n; ++i) err_h = fmax(err_h, fabs(buf[i] - logistic(xq[q], y0v, K_h[i], r)));
  }
  printf("steps=%lld launches=%lld err_solution=%.3e err_integral=%.3e err_interpolated=%.3e\n", (long long)st.steps, (long long)st.launches, err_y,
         err_i, err_h);

  /* a closure that reads its NEIGHBOURS, as source (b200rk_jit_stencil_rhs_new): the heat equation on a ring,
   *   u_i' = c0 * ((u_{i-1} - 2 u_i) + u_{i+1}),   Y(d) = u[(i + d) mod N].
   * A single Fourier mode sin(2 pi m i / N) of the semi-discrete system decays exactly like exp(-4 c0 sin^2(pi m / N) t). For the
   * FSAL pairs the whole attempt is one kernel over overlapped tiles: launches ~ attempts. */
  double err_s = 0.0;
  long long st_launches = 0, st_attempts = 0;
  {
    const size_t n2 = 4000;
    const int m = 3;
    const double c0 = 100.0, t_end = 2.0, pi = 3.14159265358979323846;
    const double lam = 4.0 * c0 * sin(pi * m / (double)n2) * sin(pi * m / (double)n2);
    double* u_h = (double*)malloc(n2 * sizeof(double));
    for (size_t i = 0; i < n2; ++i) u_h[i] = sin(2.0 * pi * m * (double)i / (double)n2);
    b200rk_vec* u0 = NULL;
    CHECK(b200rk_vec_new(ctx, n2, &u0));
    CHECK(b200rk_vec_upload(u0, u_h));
    b200rk_rhs_fn fs = NULL;
    void* us = NULL;
    CHECK(b200rk_jit_stencil_rhs_new(ctx, "c0*((Y(-1) - 2.0*Y(0)) + Y(1))", 1, 1, 0, NULL, 1, &c0, &fs, &us));
    {
      b200rk_rhs_fn f3 = NULL;
      void* u3 = NULL;
      const int rc = b200rk_jit_stencil_rhs_new(ctx, "Y(2) - Y(0)", 1, 1, 0, NULL, 0, NULL, &f3, &u3);   /* offset outside the declared radii */
      printf("bad_stencil_rc=%d is_einval=%d\n", rc, rc == B200RK_EINVAL);
      if (rc != B200RK_EINVAL) return 8;
    }
    b200rk_options o2;
    CHECK(b200rk_options_new(&o2, 1e-2, 1e-9, 1e-9, 0.05, 1e-8, 4.0, 0.1, 0.0));
    const double ts2[2] = {0.0, t_end};
    double to2[2];
    b200rk_vec* us_out[2];
    size_t n_o2 = 0;
    b200rk_stats st2;
    CHECK(b200rk_solve(ctx, method, fs, us, u0, ts2, 2, &o2, to2, us_out, &n_o2, &st2));
    if (n_o2 != 2) return 9;
    CHECK(b200rk_vec_download(us_out[1], u_h));
    for (size_t i = 0; i < n2; ++i) err_s = fmax(err_s, fabs(u_h[i] - exp(-lam * t_end) * sin(2.0 * pi * m * (double)i / (double)n2)));
    st_launches = (long long)st2.launches; st_attempts = (long long)st2.attempts;
    b200rk_vec_free(us_out[0]); b200rk_vec_free(us_out[1]); b200rk_vec_free(u0);
    b200rk_jit_rhs_free(us);
    free(u_h);
  }
  printf("stencil_attempts=%lld stencil_launches=%lld err_stencil=%.3e\n", st_attempts, st_launches, err_s);

  for (int k = 0; k < NT; ++k) { b200rk_vec_free(ys[k]); b200rk_vec_free(I[k]); b200rk_vec_free(dys[k]); }
  b200rk_vec_free(H[0]); b200rk_vec_free(H[1]);
  b200rk_jit_rhs_free(user);
  b200rk_vec_free(K); b200rk_vec_free(y0);
  b200rk_destroy(ctx);
  free(K_h); free(y0_h); free(buf);
  return (err_y < 1e-6 && err_i < 1e-5 && err_h < 1e-5 && err_s < 1e-7 && st_launches <= st_attempts + 8) ? 0 : 2;
}
