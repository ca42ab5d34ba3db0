/* b200rk.h — C-ABI of the B200-native explicit Runge–Kutta stepper (libb200rk.so).
 *
 * Drop-in boundary for ONE path of SciNim/numericalnim: the explicit RK stage loop of
 * src/numericalnim/ode.nim, its adaptive error norm / step-size controller, and the Vector[T]
 * element-wise operators of src/numericalnim/utils.nim underneath it. The reference has no FFI of its
 * own (it is pure Nim generics), so every entry point below names the reference procedure whose role it
 * takes; nim/b200rk.nim binds these symbols 1:1 with {.importc, cdecl.} and re-exposes
 * solveODE / newODEoptions / IntegratorProc for a device-resident vector type.
 *
 * Conventions
 *   - plain C types only; every function returns an int status (0 = OK) and never throws;
 *     B200RK_EINVAL plays the role of Nim's ValueError; b200rk_last_error() returns the message.
 *     A NULL handle or output pointer, and vectors of two different contexts in one call, are B200RK_EINVAL.
 *   - all device work is enqueued on the context's CUDA stream; functions that return scalars to the
 *     host (step, solve, sum) synchronise that stream, the others do not.
 *   - one context = one GPU = one host thread. Multi-GPU = one process (rank) per GPU; a vector of
 *     global length N is sharded contiguously, rank r owning [r*chunk, min(N,(r+1)*chunk)), and the only
 *     collective on the path is one all-reduce(sum, 1 x fp64) of the squared error norm per attempt. It is
 *     fused into the reducing kernel: the last CTA stores the shard's partial into every peer's mailbox over
 *     NVLink (CUDA-IPC mapped peer memory) and adds the `world` partials in rank order. If the mailboxes
 *     cannot be mapped (or B200RK_P2P=0) every rank falls back to one ncclAllReduce per attempt.
 *   - step functions never write their inputs; outputs must not alias inputs.
 */
#ifndef B200RK_H
#define B200RK_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200RK_API __attribute__((visibility("default")))

typedef struct b200rk_ctx b200rk_ctx;       /* device, stream, NCCL communicator, scratch, counters */
typedef struct b200rk_vec b200rk_vec;       /* sharded fp64 device vector == Vector[float] (utils.nim:14-17) */
typedef struct b200rk_solver b200rk_solver; /* resumable ODESolver state (ode.nim:471-586) */

enum b200rk_status {
  B200RK_OK = 0,
  B200RK_EINVAL = 1,     /* ValueError: bad options (ode.nim:95-100), unknown integrator (ode.nim:651), size mismatch (utils.nim:22-26) */
  B200RK_ECUDA = 2,
  B200RK_ENCCL = 3,
  B200RK_ENOMEM = 4,
  B200RK_ECALLBACK = 5,  /* the user's right-hand side returned non-zero */
  B200RK_ENONFINITE = 6  /* error norm became NaN: the reference would loop forever (ode.nim:69-76) */
};

/* ODEoptions, field for field (ode.nim:26-34). */
typedef struct b200rk_options {
  double dt, dtMax, dtMin, tStart, absTol, relTol, scaleMax, scaleMin;
} b200rk_options;

/* Integrators == the `case` of solveODE (ode.nim:607-651). */
enum b200rk_method {
  B200RK_DOPRI54 = 0, B200RK_TSIT54 = 1, B200RK_VERN65 = 2, B200RK_RK4 = 3,
  B200RK_RK21 = 4, B200RK_BS32 = 5, B200RK_HEUN2 = 6, B200RK_RALSTON2 = 7, B200RK_KUTTA3 = 8,
  B200RK_HEUN3 = 9, B200RK_RALSTON3 = 10, B200RK_SSPRK3 = 11, B200RK_RALSTON4 = 12, B200RK_KUTTA4 = 13,
  B200RK_METHOD_COUNT = 14
};

/* ODEProc[T] (ode.nim:36): dydt = f(t, y, ctx). Must ENQUEUE its work on b200rk_stream(ctx) and write
 * only `dydt`; `user` stands for the NumContext (commonTypes.nim:3-6). Return 0 on success. */
typedef int (*b200rk_rhs_fn)(double t, const b200rk_vec* y, b200rk_vec* dydt, void* user);

typedef struct b200rk_stats {
  int64_t steps;         /* accepted steps (driver iterations) */
  int64_t attempts;      /* stage-loop executions, retries included */
  int64_t rejected;      /* attempts with error > 1 */
  int64_t limiter_hits;  /* limitCounter increments (ode.nim:72-74) */
  int64_t rhs_evals;     /* right-hand-side callbacks issued */
  int64_t launches;      /* library kernels launched (RHS built-ins included) */
  int64_t collectives;   /* ncclAllReduce calls */
} b200rk_stats;

/* Per-kernel-class device timing (CUDA events on the context stream) for the roofline report. */
enum b200rk_kernel_class { B200RK_K_STAGE = 0, B200RK_K_FINISH = 1, B200RK_K_RHS = 2, B200RK_K_OTHER = 3, B200RK_K_FUSED = 4, B200RK_K_QUAD = 5, B200RK_K_COUNT = 6 };
typedef struct b200rk_profile {
  int64_t launches[B200RK_K_COUNT];
  double ms[B200RK_K_COUNT];              /* summed event time */
  double algorithmic_bytes[B200RK_K_COUNT]; /* summed algorithmic bytes */
} b200rk_profile;

/* ---- context -------------------------------------------------------------------------------- */
B200RK_API int b200rk_init(b200rk_ctx** out, int device);
/* rank 0 obtains an id, the host distributes the 128 bytes, every rank calls init_distributed. */
B200RK_API int b200rk_nccl_unique_id(void* out128);
B200RK_API int b200rk_init_distributed(b200rk_ctx** out, int device, int rank, int world, const void* id128);
B200RK_API void b200rk_destroy(b200rk_ctx* ctx);
B200RK_API const char* b200rk_last_error(const b200rk_ctx* ctx); /* ctx may be NULL: last error of this thread */
B200RK_API void* b200rk_stream(const b200rk_ctx* ctx);            /* cudaStream_t */
B200RK_API int b200rk_synchronize(b200rk_ctx* ctx);
B200RK_API int b200rk_rank(const b200rk_ctx* ctx);
B200RK_API int b200rk_world(const b200rk_ctx* ctx);
/* knobs: "strict_zeros" (1 = multiply zero Butcher weights through like the reference instead of
 * skipping the read), "vec_width" (2|4 doubles per access), "ctas_per_sm" (0 = one tile per CTA,
 * k = persistent grid of k*SMs CTAs), "finish_ctas_per_sm" / "fused_ctas_per_sm" (same for the reducing
 * kernels), "fuse_pointwise" (1 = element-local built-in right-hand sides run a whole attempt as one
 * kernel; 0 = always the stage / RHS / finish pipeline), "fuse_stencil" (built-in Lorenz-96: stage accumulate + stencil
 * in one kernel), "device_loop" (-1 auto | 0 | 1: the whole adaptive loop in one persistent kernel), "spin_readback",
 * "l2_hints" (-1 auto | 0 | 1: producer/consumer hand-off through the L2), "fuse_simpson" (default 1: cumsimpson as one kernel), "fuse_stencil_attempt" (default 1: a whole attempt of the built-in Lorenz-96 right-hand side in one
 * kernel, with "l96_attempt_pairs" 1|2 = 512- or 1024-wide tiles, "l96_ctas_per_sm" (0 = occupancy), "l96_warp_tiles" (default 0; 8 | 4 = warp-sized tiles exchanged by shuffle with that many elements per lane — measured slower) and,
 * sharded, "l96_peer_halo" 1|0 = halo read in place from the peer-mapped ring neighbours | one ncclSend/ncclRecv per step), "finish_prefetch" (default 0: measured, no gain),
 * "tstart_copy" (solve_host: 0 = the tStart state returns by D2H on a second stream, 1 = host-side copy), "peer_timeout_s" (how long a kernel waits for a peer's partial sum, default 120), "profile" (0|1), "pool_budget_mb" (bytes of freed vectors the context keeps for reuse;
 * 0 = release everything now). Read-only: "p2p", "sm_count". */
B200RK_API int b200rk_set(b200rk_ctx* ctx, const char* key, int64_t value);
B200RK_API int b200rk_get(const b200rk_ctx* ctx, const char* key, int64_t* value);
B200RK_API int b200rk_profile_reset(b200rk_ctx* ctx);
B200RK_API int b200rk_profile_read(b200rk_ctx* ctx, b200rk_profile* out); /* synchronises */
B200RK_API int b200rk_ctx_stats(const b200rk_ctx* ctx, b200rk_stats* out); /* launches / collectives since init */

/* ---- options / dispatch ----------------------------------------------------------------------- */
/* newODEoptions (ode.nim:78-102): validation + abs() of every field except tStart. */
B200RK_API int b200rk_options_new(b200rk_options* out, double dt, double absTol, double relTol, double dtMax,
                                  double dtMin, double scaleMax, double scaleMin, double tStart);
B200RK_API void b200rk_options_default(b200rk_options* out); /* DEFAULT_ODEoptions (ode.nim:104) */
/* integrator.toLower() lookup (ode.nim:607); unknown name -> B200RK_EINVAL "<name> is not a valid integrator". */
B200RK_API int b200rk_method_from_name(const char* name, int* method);
B200RK_API const char* b200rk_method_name(int method);
/* method properties as handed to ODESolver (ode.nim:609-649) */
B200RK_API int b200rk_method_info(int method, int* stages, int* use_fsal, double* order, int* adaptive);
/* dense tableau of the three FSAL pairs for cross-checks: c[10], a[10*9] (a[s*9+j-1] = a_sj), b[9], bhat[9] */
B200RK_API int b200rk_method_tableau(int method, double* c, double* a, double* b, double* bhat);

/* Contiguous shard of rank `rank` of `world` for a vector of n_global elements: [offset, offset+len).
 * chunk = ceil(n_global/world) rounded up to a multiple of 4 elements (32-byte aligned boundaries). Host-only. */
B200RK_API int b200rk_shard_range(size_t n_global, int rank, int world, size_t* offset, size_t* len);

/* ---- vectors (Vector[float], utils.nim:14-271) ------------------------------------------------ */
B200RK_API int b200rk_vec_new(b200rk_ctx* ctx, size_t n_global, b200rk_vec** out); /* newVector; contents undefined */
B200RK_API int b200rk_vec_free(b200rk_vec* v);
B200RK_API size_t b200rk_vec_len(const b200rk_vec* v);          /* global length == Vector.len */
B200RK_API size_t b200rk_vec_local_len(const b200rk_vec* v);
B200RK_API size_t b200rk_vec_local_offset(const b200rk_vec* v);
B200RK_API double* b200rk_vec_data(const b200rk_vec* v);        /* device pointer of the local shard */
/* host arrays of GLOBAL length: each rank moves only its own slice */
B200RK_API int b200rk_vec_upload(b200rk_vec* v, const double* host_global);
B200RK_API int b200rk_vec_download(const b200rk_vec* v, double* host_global);
/* host arrays of LOCAL length */
B200RK_API int b200rk_vec_upload_local(b200rk_vec* v, const double* host_local);
B200RK_API int b200rk_vec_download_local(const b200rk_vec* v, double* host_local);
/* same as upload_local without the trailing stream synchronisation: the copy is ordered on the context stream in front of
 * whatever is enqueued next (host_local must be pinned and stay untouched until b200rk_synchronize or a solve returns) */
B200RK_API int b200rk_vec_upload_local_async(b200rk_vec* v, const double* host_local);
B200RK_API int b200rk_vec_copy(b200rk_vec* dst, const b200rk_vec* src);          /* clone, utils.nim:269 */
B200RK_API int b200rk_vec_fill(b200rk_vec* v, double value);
/* element-wise operators; size mismatch -> B200RK_EINVAL (utils.nim:22-26). out may alias an input. */
B200RK_API int b200rk_vec_add(b200rk_vec* out, const b200rk_vec* a, const b200rk_vec* b);   /* utils.nim:59-64 */
B200RK_API int b200rk_vec_sub(b200rk_vec* out, const b200rk_vec* a, const b200rk_vec* b);   /* utils.nim:113-118 */
B200RK_API int b200rk_vec_hmul(b200rk_vec* out, const b200rk_vec* a, const b200rk_vec* b);  /* `*.` utils.nim:186-191 */
B200RK_API int b200rk_vec_hdiv(b200rk_vec* out, const b200rk_vec* a, const b200rk_vec* b);  /* `/.` utils.nim:192-197 */
B200RK_API int b200rk_vec_scale(b200rk_vec* out, double s, const b200rk_vec* a);            /* utils.nim:171-180 */
B200RK_API int b200rk_vec_div_scalar(b200rk_vec* out, const b200rk_vec* a, double s);       /* utils.nim:166-170 */
B200RK_API int b200rk_vec_add_scalar(b200rk_vec* out, double s, const b200rk_vec* a);       /* `+.` utils.nim:78-82 */
B200RK_API int b200rk_vec_neg(b200rk_vec* out, const b200rk_vec* a);                        /* utils.nim:214-218 */
B200RK_API int b200rk_vec_abs(b200rk_vec* out, const b200rk_vec* a);                        /* utils.nim:219-223 */
B200RK_API int b200rk_vec_sum(const b200rk_vec* a, double* out);                            /* utils.nim:243-250 (+ allreduce) */
/* hermiteSpline (utils.nim:273-279) */
B200RK_API int b200rk_hermite(b200rk_vec* out, double x, double x1, double x2, const b200rk_vec* y1,
                              const b200rk_vec* y2, const b200rk_vec* dy1, const b200rk_vec* dy2);

/* ---- built-in right-hand sides (device-side user closures for the benchmark IVPs) -------------- */
enum b200rk_rhs_kind {
  B200RK_RHS_SCALE = 0,       /* dydt = c * y              (tests/test_ode.nim:5-7 with c = -0.1) */
  B200RK_RHS_DIAG_LINEAR = 1, /* dydt = -(lambda .* y)     (BASELINE.json configs 2, 4) */
  B200RK_RHS_LORENZ96 = 2     /* dydt[i] = (y[i+1]-y[i-2])*y[i-1] - y[i] + F, cyclic (config 3) */
};
B200RK_API int b200rk_builtin_rhs_new(b200rk_ctx* ctx, int kind, double scalar, const b200rk_vec* lambda,
                                      b200rk_rhs_fn* fn, void** user);
B200RK_API int b200rk_builtin_rhs_free(void* user);

/* ---- right-hand sides from source: element-local ODEProc closures compiled INTO the fused kernels ---- */
/* The reference's extension point is the user's closure f(t, y, ctx) (ode.nim:36). An element-local closure,
 *     dydt[i] = expr(t, y[i], p0[i], .., p3[i], c0, .., c7),
 * can be handed over as a CUDA C++ expression in the variables t, y, p0..p<n_vec-1> (per-element parameter
 * vectors, same length as y) and c0..c<n_scalar-1> (scalars): e.g. "-(p0*y)", "c0*y*(1.0 - y/p0)". It is
 * compiled at run time (NVRTC, sm_100a, --fmad=false: a*b+c stays a multiply and an add like the reference's CPU
 * arithmetic) into the same kernels the built-in right-hand sides use: the whole-attempt kernel of
 * DOPRI54 / Tsit54 / Vern65, the device-resident driver loop, the one-kernel RK4 step, and a plain dydt kernel for
 * every other method. The returned (fn, user) pair is an ordinary b200rk_rhs_fn for b200rk_step / b200rk_solve /
 * b200rk_solver_new. A wrong expression -> B200RK_EINVAL with the compiler log in b200rk_last_error().
 * The parameter vectors must outlive `user`, and `user` must be freed before its context is destroyed. */
B200RK_API int b200rk_jit_rhs_new(b200rk_ctx* ctx, const char* expr, int n_vec, const b200rk_vec* const* vecs,
                                  int n_scalar, const double* scalars, b200rk_rhs_fn* fn, void** user);
B200RK_API int b200rk_jit_rhs_set_scalars(void* user, int n_scalar, const double* scalars); /* no recompilation */
B200RK_API int b200rk_jit_rhs_free(void* user);
/* Host only (no device needed): compile one translation unit and return the sm_100a cubin (cubin_out may be NULL;
 * *cubin_bytes receives the size) and the compiler log + kernel names. pattern: -1 = dydt / RK4 kernels,
 * 0..4 = fused attempt + device-loop kernels (dopri54, dopri54 strict, tsit54, vern65, vern65 strict). */
B200RK_API int b200rk_jit_compile_only(const char* expr, int n_vec, int n_scalar, int pattern, void* cubin_out,
                                       size_t cubin_cap, size_t* cubin_bytes, char* log, size_t log_cap);
/* STENCIL right-hand sides from source: an ODEProc closure (ode.nim:36) whose dydt[i] reads a neighbourhood of y,
 *     dydt[i] = expr(t, Y(-radius_left) .. Y(+radius_right), p0[i] .., c0 ..),     Y(d) = y[(i + d) mod N]   (cyclic),
 * e.g. Lorenz-96 "((Y(1) - Y(-2)) * Y(-1) - Y(0)) + c0" with radii 2 / 1, diffusion "c0*((Y(-1) - 2.0*Y(0)) + Y(1))" with 1 / 1.
 * Radii 0..8; a Y(d) outside them is a compile error. NVRTC compiles the expression into (i) a plain dydt = f(t, y) kernel
 * every method calls through the stage / RHS / finish pipeline, (ii) a one-kernel RK4 step and (iii) — for DOPRI54 / Tsit54 / Vern65 — the ONE-KERNEL attempt
 * over overlapped tiles (the built-in Lorenz-96's kernel with the overlap the radii ask for: 4 + n_vec vector passes per attempt
 * instead of ~55). Sharded, the halo travels once per step (or is read in place from the peer-mapped ring neighbours inside a
 * solver). Everything else as for b200rk_jit_rhs_new; freed with b200rk_jit_rhs_free. */
B200RK_API int b200rk_jit_stencil_rhs_new(b200rk_ctx* ctx, const char* expr, int radius_left, int radius_right, int n_vec,
                                          const b200rk_vec* const* vecs, int n_scalar, const double* scalars, b200rk_rhs_fn* fn, void** user);
/* Host only: compile a stencil unit (pattern -1 = the dydt kernel, 0..4 = the whole-attempt kernel of that pair). */
B200RK_API int b200rk_jit_stencil_compile_only(const char* expr, int radius_left, int radius_right, int n_vec, int n_scalar, int pattern,
                                               size_t* cubin_bytes, char* log, size_t log_cap);

/* ---- consumers of a trajectory: Hermite interpolation and cumulative quadrature -------- */
/* A trajectory is a list of device vectors (what b200rk_solve returns) with its times. All results are newly
 * allocated vectors (caller frees each); `out` needs room for nx (interpolation) / m (quadrature) handles and
 * *n_out receives how many the reference returns — fewer than asked where it drops samples (below). */
/* hermiteInterpolate(x, t, y, dy) (utils.nim:282-312): cubic Hermite spline through (t[i], y[i]) with slopes dy[i],
 * evaluated at every x[k], all samples in ONE kernel. x sorted: samples outside [t[0], t[high]) are silently
 * dropped and x == t[high] is returned once (utils.nim:290-300); x unsorted: a sample outside the data ->
 * B200RK_EINVAL "<x> not in interval .." (utils.nim:312). */
B200RK_API int b200rk_hermite_interpolate(b200rk_ctx* ctx, const double* x, size_t nx, const double* t, size_t nt,
                                          const b200rk_vec* const* y, const b200rk_vec* const* dy, b200rk_vec** out,
                                          size_t* n_out);
/* cumtrapz(Y, X) (integrate.nim:119-135): X is sorted, duplicates are trimmed (utils.nim:360-420; same x with different
 * vectors -> B200RK_EINVAL "impure y-duplicates"), then out[k] = integral from X_sorted[0] to X_sorted[k]. One kernel:
 * every input read once, every output written once. */
B200RK_API int b200rk_cumtrapz(b200rk_ctx* ctx, const b200rk_vec* const* Y, const double* X, size_t m, b200rk_vec** out,
                               size_t* n_out);
/* cumsimpson(Y, X) (integrate.nim:330-378): Simpson's rule on pairs of unequal intervals, then Hermite interpolation of
 * the running integral back onto the ORIGINAL X. Fewer than 3 distinct points -> B200RK_EINVAL. */
B200RK_API int b200rk_cumsimpson(b200rk_ctx* ctx, const b200rk_vec* const* Y, const double* X, size_t m, b200rk_vec** out,
                                 size_t* n_out);
/* NumContextProc[T, float] (commonTypes.nim): the integrand at one point. Must ENQUEUE on b200rk_stream(ctx), write only
 * `out` (a vector of n_global elements) and return 0. */
typedef int (*b200rk_fn_of_t)(double t, b200rk_vec* out, void* user);
/* cumtrapz(f, X, ctx, dx) (integrate.nim:138-175): trapezoidal rule from min(X) to max(X) + 1.0 in steps of dx,
 * interpolated at X. Streams: four vectors alive whatever the number of steps. */
B200RK_API int b200rk_cumtrapz_fn(b200rk_ctx* ctx, b200rk_fn_of_t f, void* user, size_t n_global, const double* X, size_t m,
                                  double dx, b200rk_vec** out, size_t* n_out);
/* cumsimpson(f, X, ctx, dx) (integrate.nim:379-400): f on linspace(min X, max X, round((max-min)/dx) + 2), the discrete
 * rule, interpolated at X. Up to 4096 grid points that fit, all evaluations are alive at once like in the reference;
 * finer grids are streamed through a window of a few vectors (same results; knob "stream_simpson": 1 always, 0 never). */
B200RK_API int b200rk_cumsimpson_fn(b200rk_ctx* ctx, b200rk_fn_of_t f, void* user, size_t n_global, const double* X,
                                    size_t m, double dx, b200rk_vec** out, size_t* n_out);

/* Host only (no device needed): the host-side decisions of the routines above, for tests and tooling. hermite_plan:
 * for every sample hermiteInterpolate(x, t, ..) returns, in order, the data interval [t[j], t[j+1]] it is evaluated on
 * (is_copy = 1: it is the copy of data point j = nt-1, utils.nim:299-300) and {h00, h10*(x2-x1), h01, h11*(x2-x1)}
 * (utils.nim:273-279); arrays need room for nx entries (4*nx factors). simpson_weights: (alpha, beta, eta) of
 * integrate.nim:357-359 (tail = 0) or :367-369 (tail = 1). */
B200RK_API int b200rk_hermite_plan(const double* x, size_t nx, const double* t, size_t nt, int* interval, int* is_copy,
                                   double* factors, size_t* n_out);
B200RK_API int b200rk_simpson_weights(int tail, double h1, double h2, double* alpha, double* beta, double* eta);

/* ---- the hot path ------------------------------------------------------------------------------ */
/* One IntegratorProc call (ode.nim:38): (yNew, newFSAL, dtUsed, error) = X_step(f, t, y, FSAL, dt, options, ctx).
 * Runs the adaptive retry loop of commonAdaptiveMethodCode (ode.nim:57-76) for adaptive methods.
 * fsal may be NULL for methods that ignore it; fsal_new may be NULL for non-FSAL methods (the reference
 * returns yNew there). */
B200RK_API int b200rk_step(b200rk_ctx* ctx, int method, b200rk_rhs_fn f, void* user, double t,
                           const b200rk_vec* y, const b200rk_vec* fsal, double dt, const b200rk_options* options,
                           b200rk_vec* y_new, b200rk_vec* fsal_new, double* dt_used, double* error);

/* solveODE (ode.nim:589-651) on device vectors. t_out (room for n_tspan) receives the sorted times; y_out
 * receives *n_y_out newly allocated vectors (caller frees each), in the order of t_out. As in the reference,
 * n_y_out can be smaller than n_tspan. The reference reports tStart once however
 * often tspan holds it (ode.nim:485-487) and a NaN in tspan passes neither of its filters (ode.nim:479-480):
 * the returned time list is then shorter than n_tspan and the unused tail of t_out is set to NaN.
 * stats may be NULL. */
B200RK_API int b200rk_solve(b200rk_ctx* ctx, int method, b200rk_rhs_fn f, void* user, const b200rk_vec* y0,
                            const double* tspan, size_t n_tspan, const b200rk_options* options, double* t_out,
                            b200rk_vec** y_out, size_t* n_y_out, b200rk_stats* stats);

/* solveODE with HOST buffers (the end-to-end call): y0_local and y_out_local hold this rank's shard
 * (for one GPU: the whole vector). y_out_local has room for n_tspan * local_len doubles. Host<->device
 * copies happen inside. */
B200RK_API int b200rk_solve_host(b200rk_ctx* ctx, int method, b200rk_rhs_fn f, void* user, size_t n_global,
                                 const double* y0_local, const double* tspan, size_t n_tspan,
                                 const b200rk_options* options, double* t_out, double* y_out_local,
                                 size_t* n_y_out, b200rk_stats* stats);

/* Resumable forward driver: the `while t < tEnd` loop of ODESolver (ode.nim:508-542) for tspan =
 * [tStart, t_end] (no dense output), advanced a bounded number of accepted steps at a time. */
B200RK_API int b200rk_solver_new(b200rk_ctx* ctx, int method, b200rk_rhs_fn f, void* user, const b200rk_vec* y0,
                                 double t_end, const b200rk_options* options, b200rk_solver** out);
B200RK_API int b200rk_solver_advance(b200rk_solver* s, int64_t max_steps, int64_t* steps_done, int* finished);
B200RK_API int b200rk_solver_state(const b200rk_solver* s, double* t, double* dt_next, double* last_error,
                                   const b200rk_vec** y);
B200RK_API int b200rk_solver_stats(const b200rk_solver* s, b200rk_stats* out);
B200RK_API int b200rk_solver_free(b200rk_solver* s);

/* ---- raw kernels (bandwidth sweep, BASELINE.json config 5, and kernel-level parity tests) ------ */
/* out = y + c*(w[0]*k[0] + ... + w[m-1]*k[m-1]), left-associated (ode.nim:294-299); 1 <= m <= 9.
 * chain != 0: out = ((y + w[0]*k[0]) + w[1]*k[1]) + ... (ode.nim:128), c ignored. */
B200RK_API int b200rk_stage_accum(b200rk_ctx* ctx, int m, const double* w, double c, int chain,
                                  const b200rk_vec* y, const b200rk_vec* const* k, b200rk_vec* out);
/* final combine + error norm of DOPRI54 / TSIT54 / VERN65 / RK21 / BS32 given all stage derivatives
 * k[0..stages): writes y_new, optionally the element-wise error_y (err_y may be NULL), and returns the
 * global sum of squares S and error = sqrt(1/N * S) (ode.nim:61-65). */
B200RK_API int b200rk_combine_err(b200rk_ctx* ctx, int method, double dt, double absTol, double relTol,
                                  const b200rk_vec* y, const b200rk_vec* const* k, b200rk_vec* y_new,
                                  b200rk_vec* err_y, double* sumsq, double* error);
/* y + dt/6*(k1 + 2*(k2+k3) + k4) (ode.nim:188) */
B200RK_API int b200rk_rk4_combine(b200rk_ctx* ctx, double dt, const b200rk_vec* y, const b200rk_vec* k1,
                                  const b200rk_vec* k2, const b200rk_vec* k3, const b200rk_vec* k4, b200rk_vec* out);

#ifdef __cplusplus
}
#endif
#endif /* B200RK_H */
