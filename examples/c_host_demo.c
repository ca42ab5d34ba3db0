/* c_host_demo.c — the C-ABI used from plain C, the way the Nim {.importc, cdecl.} shim uses it.
 * Solves the diag-linear IVP  y' = -lambda .* y,  y0[i] = 1 + 0.5 sin(2 pi i/N),  lambda[i] = 0.1 + 9.9 i/(N-1)
 * over [0, 2] with the named integrator on cuda:0, host buffers in and out (b200rk_solve_host), and prints the
 * step statistics plus y(2) so tests/test_gpu_c_host.py can compare them with the CPU oracle.
 *   gcc -std=c99 -O2 -Iinclude examples/c_host_demo.c -Lnumericalnim_b200/lib -lb200rk -lm -o c_host_demo
 *   ./c_host_demo dopri54 4096
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

int main(int argc, char** argv) {
  const char* integrator = argc > 1 ? argv[1] : "dopri54";
  const size_t n = argc > 2 ? (size_t)strtoull(argv[2], NULL, 10) : 4096;
  b200rk_ctx* ctx = NULL;
  int method = 0;
  CHECK(b200rk_method_from_name(integrator, &method)); /* ValueError path of solveODE (ode.nim:650-651) */
  CHECK(b200rk_init(&ctx, 0));

  double* lam_h = (double*)malloc(n * sizeof(double));
  double* y0_h = (double*)malloc(n * sizeof(double));
  double* out_h = (double*)malloc(2 * n * sizeof(double));
  const double pi = 3.14159265358979323846;
  for (size_t i = 0; i < n; ++i) {
    lam_h[i] = 0.1 + 9.9 * (double)i / (double)(n - 1);
    y0_h[i] = 1.0 + 0.5 * sin(2.0 * pi * (double)i / (double)n);
  }
  b200rk_vec* lam = NULL;
  CHECK(b200rk_vec_new(ctx, n, &lam));
  CHECK(b200rk_vec_upload(lam, lam_h));
  b200rk_rhs_fn f = NULL;
  void* user = NULL;
  CHECK(b200rk_builtin_rhs_new(ctx, B200RK_RHS_DIAG_LINEAR, 0.0, lam, &f, &user));

  b200rk_options opt;
  CHECK(b200rk_options_new(&opt, 1e-2, 1e-6, 1e-6, 1.0, 1e-8, 4.0, 0.1, 0.0)); /* newODEoptions(dt, absTol, relTol, dtMax, dtMin, ...) */
  const double tspan[2] = {0.0, 2.0};
  double t_out[2];
  size_t n_out = 0;
  b200rk_stats st;
  CHECK(b200rk_solve_host(ctx, method, f, user, n, y0_h, tspan, 2, &opt, t_out, out_h, &n_out, &st));

  printf("integrator=%s n=%zu n_out=%zu steps=%lld attempts=%lld rejected=%lld launches=%lld\n", integrator, n, n_out,
         (long long)st.steps, (long long)st.attempts, (long long)st.rejected, (long long)st.launches);
  const double* y_end = out_h + (n_out - 1) * n;
  for (size_t i = 0; i < n; i += (n / 8 ? n / 8 : 1)) printf("y[%zu]=%a\n", i, y_end[i]);
  double max_err = 0.0;
  for (size_t i = 0; i < n; ++i) {
    const double e = fabs(y_end[i] - y0_h[i] * exp(-lam_h[i] * 2.0));
    if (e > max_err) max_err = e;
  }
  printf("max_abs_err_vs_exact=%.3e\n", max_err);

  b200rk_builtin_rhs_free(user);
  b200rk_vec_free(lam);
  b200rk_destroy(ctx);
  free(lam_h); free(y0_h); free(out_h);
  return max_err < 1e-4 ? 0 : 2;
}
