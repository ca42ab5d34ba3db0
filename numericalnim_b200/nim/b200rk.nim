## b200rk.nim — thin {.importc, cdecl.} shim over libb200rk.so (include/b200rk.h), re-exposing
## numericalnim's `solveODE` / `newODEoptions` / `IntegratorProc` for a device-resident vector type.
##
## STATUS: written against the C header, NOT compiled — this image has no Nim toolchain (nim, nimble and
## choosenim are absent; see DESIGN.md §2). There is one `importc` declaration per B200RK_API symbol of the header
## (71; tests/test_nim_shim.py parses both files and compares names and arities), and the behaviour of every symbol
## is exercised through the same ABI by the pytest suite (ctypes). Usage from reference code:
##
##   import numericalnim            # for ODEoptions, NumContext, linspace, ...
##   import b200rk                  # adds the GpuVector overloads
##   proc f(t: float, y: GpuVector, ctx: NumContext[GpuVector, float]): GpuVector = -0.1 * y
##   let (ts, ys) = solveODE(f, newGpuVector(@[1.0, 1.0, 1.0]), linspace(-10.0, 10.0, 100), integrator = "tsit54")
##
## Nim's overload resolution prefers this non-generic `solveODE` over the reference's generic
## `solveODE*[T]` (ode.nim:589) when `y0` is a GpuVector, so existing call sites switch by changing the
## type of `y0` only.
import std/[strutils]
import numericalnim/ode            # ODEoptions, newODEoptions, ODEProc, IntegratorProc (ode.nim:26-38)
import numericalnim/common/commonTypes   # NumContext (commonTypes.nim:3-15)

const lib = "libb200rk.so"

type
  CtxObj {.incompleteStruct.} = object
  VecObj {.incompleteStruct.} = object
  B200rkCtx* = ptr CtxObj
  VecHandle = ptr VecObj
  COptions {.bycopy.} = object       ## b200rk_options == ODEoptions field for field (ode.nim:26-34)
    dt, dtMax, dtMin, tStart, absTol, relTol, scaleMax, scaleMin: cdouble
  CStats {.bycopy.} = object
    steps, attempts, rejected, limiterHits, rhsEvals, launches, collectives: int64
  RhsFn = proc(t: cdouble, y: VecHandle, dydt: VecHandle, user: pointer): cint {.cdecl.}

  GpuVector* = object                ## stands where Vector[float] stands in the reference (utils.nim:14-17)
    h: VecHandle
    ctx: B200rkCtx
    borrowed: bool                   ## handles lent by the library inside a callback are not freed

const
  B200RK_OK = 0.cint
  B200RK_EINVAL = 1.cint

# ---- 1:1 declarations (include/b200rk.h) ------------------------------------------------------------
proc b200rk_init(outCtx: ptr B200rkCtx, device: cint): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_nccl_unique_id(out128: pointer): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_init_distributed(outCtx: ptr B200rkCtx, device, rank, world: cint, id128: pointer): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_destroy(ctx: B200rkCtx) {.importc, cdecl, dynlib: lib.}
proc b200rk_last_error(ctx: B200rkCtx): cstring {.importc, cdecl, dynlib: lib.}
proc b200rk_stream(ctx: B200rkCtx): pointer {.importc, cdecl, dynlib: lib.}
proc b200rk_method_from_name(name: cstring, methodId: ptr cint): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_vec_new(ctx: B200rkCtx, n: csize_t, outVec: ptr VecHandle): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_vec_free(v: VecHandle): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_vec_len(v: VecHandle): csize_t {.importc, cdecl, dynlib: lib.}
proc b200rk_vec_data(v: VecHandle): ptr cdouble {.importc, cdecl, dynlib: lib.}
proc b200rk_vec_upload(v: VecHandle, host: ptr cdouble): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_vec_download(v: VecHandle, host: ptr cdouble): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_vec_copy(dst, src: VecHandle): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_vec_add(o, a, b: VecHandle): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_vec_sub(o, a, b: VecHandle): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_vec_hmul(o, a, b: VecHandle): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_vec_hdiv(o, a, b: VecHandle): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_vec_scale(o: VecHandle, s: cdouble, a: VecHandle): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_vec_div_scalar(o, a: VecHandle, s: cdouble): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_vec_add_scalar(o: VecHandle, s: cdouble, a: VecHandle): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_vec_neg(o, a: VecHandle): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_vec_abs(o, a: VecHandle): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_vec_sum(a: VecHandle, outSum: ptr cdouble): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_hermite(o: VecHandle, x, x1, x2: cdouble, y1, y2, dy1, dy2: VecHandle): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_step(ctx: B200rkCtx, methodId: cint, f: RhsFn, user: pointer, t: cdouble, y, fsal: VecHandle, dt: cdouble,
                 options: ptr COptions, yNew, fsalNew: VecHandle, dtUsed, error: ptr cdouble): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_solve(ctx: B200rkCtx, methodId: cint, f: RhsFn, user: pointer, y0: VecHandle, tspan: ptr cdouble, nTspan: csize_t,
                  options: ptr COptions, tOut: ptr cdouble, yOut: ptr VecHandle, nYOut: ptr csize_t, stats: ptr CStats): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_jit_rhs_new(ctx: B200rkCtx, expr: cstring, nVec: cint, vecs: ptr VecHandle, nScalar: cint, scalars: ptr cdouble,
                        fn: ptr RhsFn, user: ptr pointer): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_jit_stencil_rhs_new(ctx: B200rkCtx, expr: cstring, radiusLeft, radiusRight, nVec: cint, vecs: ptr VecHandle, nScalar: cint, scalars: ptr cdouble,
                                fn: ptr RhsFn, user: ptr pointer): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_jit_stencil_compile_only(expr: cstring, radiusLeft, radiusRight, nVec, nScalar, pattern: cint, cubinBytes: ptr csize_t,
                                     log: cstring, logCap: csize_t): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_jit_rhs_set_scalars(user: pointer, nScalar: cint, scalars: ptr cdouble): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_jit_rhs_free(user: pointer): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_hermite_interpolate(ctx: B200rkCtx, x: ptr cdouble, nx: csize_t, t: ptr cdouble, nt: csize_t, y, dy: ptr VecHandle,
                                outVecs: ptr VecHandle, nOut: ptr csize_t): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_cumtrapz(ctx: B200rkCtx, Y: ptr VecHandle, X: ptr cdouble, m: csize_t, outVecs: ptr VecHandle, nOut: ptr csize_t): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_cumsimpson(ctx: B200rkCtx, Y: ptr VecHandle, X: ptr cdouble, m: csize_t, outVecs: ptr VecHandle, nOut: ptr csize_t): cint {.importc, cdecl, dynlib: lib.}
type FnOfT = proc(t: cdouble, outVec: VecHandle, user: pointer): cint {.cdecl.}
proc b200rk_cumtrapz_fn(ctx: B200rkCtx, f: FnOfT, user: pointer, nGlobal: csize_t, X: ptr cdouble, m: csize_t, dx: cdouble,
                        outVecs: ptr VecHandle, nOut: ptr csize_t): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_cumsimpson_fn(ctx: B200rkCtx, f: FnOfT, user: pointer, nGlobal: csize_t, X: ptr cdouble, m: csize_t, dx: cdouble,
                          outVecs: ptr VecHandle, nOut: ptr csize_t): cint {.importc, cdecl, dynlib: lib.}

# the rest of the ABI, one declaration per B200RK_API symbol (tests/test_nim_shim.py compares this list with the header)
type
  SolverObj {.incompleteStruct.} = object
  SolverHandle = ptr SolverObj
  CProfile {.bycopy.} = object       ## b200rk_profile: per kernel class (stage, finish, rhs, other, fused, quad)
    launches: array[6, int64]
    ms: array[6, cdouble]
    algorithmicBytes: array[6, cdouble]
proc b200rk_synchronize(ctx: B200rkCtx): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_rank(ctx: B200rkCtx): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_world(ctx: B200rkCtx): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_set(ctx: B200rkCtx, key: cstring, value: int64): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_get(ctx: B200rkCtx, key: cstring, value: ptr int64): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_profile_reset(ctx: B200rkCtx): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_profile_read(ctx: B200rkCtx, outProfile: ptr CProfile): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_ctx_stats(ctx: B200rkCtx, outStats: ptr CStats): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_options_new(outOptions: ptr COptions, dt, absTol, relTol, dtMax, dtMin, scaleMax, scaleMin, tStart: cdouble): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_options_default(outOptions: ptr COptions) {.importc, cdecl, dynlib: lib.}
proc b200rk_method_name(methodId: cint): cstring {.importc, cdecl, dynlib: lib.}
proc b200rk_method_info(methodId: cint, stages, useFsal: ptr cint, order: ptr cdouble, adaptive: ptr cint): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_method_tableau(methodId: cint, c, a, b, bhat: ptr cdouble): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_shard_range(nGlobal: csize_t, rank, world: cint, offset, len: ptr csize_t): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_vec_local_len(v: VecHandle): csize_t {.importc, cdecl, dynlib: lib.}
proc b200rk_vec_local_offset(v: VecHandle): csize_t {.importc, cdecl, dynlib: lib.}
proc b200rk_vec_upload_local(v: VecHandle, hostLocal: ptr cdouble): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_vec_download_local(v: VecHandle, hostLocal: ptr cdouble): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_vec_upload_local_async(v: VecHandle, hostLocal: ptr cdouble): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_vec_fill(v: VecHandle, value: cdouble): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_builtin_rhs_new(ctx: B200rkCtx, kind: cint, scalar: cdouble, lambda: VecHandle, fn: ptr RhsFn, user: ptr pointer): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_builtin_rhs_free(user: pointer): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_jit_compile_only(expr: cstring, nVec, nScalar, pattern: cint, cubinOut: pointer, cubinCap: csize_t, cubinBytes: ptr csize_t,
                             log: cstring, logCap: csize_t): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_hermite_plan(x: ptr cdouble, nx: csize_t, t: ptr cdouble, nt: csize_t, interval, isCopy: ptr cint, factors: ptr cdouble,
                         nOut: ptr csize_t): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_simpson_weights(tail: cint, h1, h2: cdouble, alpha, beta, eta: ptr cdouble): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_solve_host(ctx: B200rkCtx, methodId: cint, f: RhsFn, user: pointer, nGlobal: csize_t, y0Local, tspan: ptr cdouble, nTspan: csize_t,
                       options: ptr COptions, tOut, yOutLocal: ptr cdouble, nYOut: ptr csize_t, stats: ptr CStats): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_solver_new(ctx: B200rkCtx, methodId: cint, f: RhsFn, user: pointer, y0: VecHandle, tEnd: cdouble, options: ptr COptions,
                       outSolver: ptr SolverHandle): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_solver_advance(s: SolverHandle, maxSteps: int64, stepsDone: ptr int64, finished: ptr cint): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_solver_state(s: SolverHandle, t, dtNext, lastError: ptr cdouble, y: ptr VecHandle): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_solver_stats(s: SolverHandle, outStats: ptr CStats): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_solver_free(s: SolverHandle): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_stage_accum(ctx: B200rkCtx, m: cint, w: ptr cdouble, c: cdouble, chain: cint, y: VecHandle, k: ptr VecHandle, outVec: VecHandle): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_combine_err(ctx: B200rkCtx, methodId: cint, dt, absTol, relTol: cdouble, y: VecHandle, k: ptr VecHandle, yNew, errY: VecHandle,
                        sumsq, error: ptr cdouble): cint {.importc, cdecl, dynlib: lib.}
proc b200rk_rk4_combine(ctx: B200rkCtx, dt: cdouble, y, k1, k2, k3, k4, outVec: VecHandle): cint {.importc, cdecl, dynlib: lib.}

# ---- error convention: status code -> Nim exception (SURVEY.md §8b) ---------------------------------
proc check(rc: cint, ctx: B200rkCtx = nil) =
  if rc == B200RK_OK: return
  let msg = $b200rk_last_error(ctx)
  if rc == B200RK_EINVAL: raise newException(ValueError, msg)     # what the reference raises
  raise newException(IOError, "b200rk: " & msg)

# ---- context ----------------------------------------------------------------------------------------
var defaultCtx: B200rkCtx
proc b200rkContext*(device = 0): B200rkCtx =
  if defaultCtx.isNil: check b200rk_init(addr defaultCtx, device.cint)
  defaultCtx

# ---- GpuVector: value semantics like Vector[T] (every operator returns a fresh vector) --------------
proc `=destroy`*(v: var GpuVector) =
  if not v.h.isNil and not v.borrowed: discard b200rk_vec_free(v.h)
  v.h = nil
proc `=copy`*(dst: var GpuVector, src: GpuVector) =
  if dst.h == src.h: return
  `=destroy`(dst)
  if src.h.isNil: return
  dst.ctx = src.ctx; dst.borrowed = false
  check(b200rk_vec_new(src.ctx, b200rk_vec_len(src.h), addr dst.h), src.ctx)
  check(b200rk_vec_copy(dst.h, src.h), src.ctx)

proc newLike(v: GpuVector): GpuVector =
  result.ctx = v.ctx
  check(b200rk_vec_new(v.ctx, b200rk_vec_len(v.h), addr result.h), v.ctx)

proc newGpuVector*(components: openArray[float], ctx: B200rkCtx = b200rkContext()): GpuVector =
  ## newVector (utils.nim:19-20): copies the host data to the device.
  result.ctx = ctx
  check(b200rk_vec_new(ctx, components.len.csize_t, addr result.h), ctx)
  if components.len > 0: check(b200rk_vec_upload(result.h, cast[ptr cdouble](unsafeAddr components[0])), ctx)

proc toSeq*(v: GpuVector): seq[float] =
  result = newSeq[float](b200rk_vec_len(v.h).int)
  if result.len > 0: check(b200rk_vec_download(v.h, cast[ptr cdouble](addr result[0])), v.ctx)

proc len*(v: GpuVector): int = b200rk_vec_len(v.h).int
proc size*(v: GpuVector): int = v.len                                         # utils.nim:57
proc clone*(v: GpuVector): GpuVector = (result = newLike(v); check(b200rk_vec_copy(result.h, v.h), v.ctx))  # utils.nim:269
proc `+`*(a, b: GpuVector): GpuVector = (result = newLike(a); check(b200rk_vec_add(result.h, a.h, b.h), a.ctx))   # utils.nim:59-64
proc `-`*(a, b: GpuVector): GpuVector = (result = newLike(a); check(b200rk_vec_sub(result.h, a.h, b.h), a.ctx))   # utils.nim:113-118
proc `*`*(d: float, a: GpuVector): GpuVector = (result = newLike(a); check(b200rk_vec_scale(result.h, d, a.h), a.ctx))  # utils.nim:176-180
proc `*`*(a: GpuVector, d: float): GpuVector = d * a                                                              # utils.nim:171-175
proc `/`*(a: GpuVector, d: float): GpuVector = (result = newLike(a); check(b200rk_vec_div_scalar(result.h, a.h, d), a.ctx))  # utils.nim:166-170
proc `-`*(a: GpuVector): GpuVector = (result = newLike(a); check(b200rk_vec_neg(result.h, a.h), a.ctx))          # utils.nim:214-218
proc abs*(a: GpuVector): GpuVector = (result = newLike(a); check(b200rk_vec_abs(result.h, a.h), a.ctx))          # utils.nim:219-223
proc `+.`*(d: float, a: GpuVector): GpuVector = (result = newLike(a); check(b200rk_vec_add_scalar(result.h, d, a.h), a.ctx))  # utils.nim:78-82
proc `*.`*(a, b: GpuVector): GpuVector = (result = newLike(a); check(b200rk_vec_hmul(result.h, a.h, b.h), a.ctx))  # utils.nim:186-191
proc `/.`*(a, b: GpuVector): GpuVector = (result = newLike(a); check(b200rk_vec_hdiv(result.h, a.h, b.h), a.ctx))  # utils.nim:192-197
proc sum*(a: GpuVector): float = check(b200rk_vec_sum(a.h, addr result), a.ctx)                                  # utils.nim:243-250
proc hermiteSpline*(x, x1, x2: float, y1, y2, dy1, dy2: GpuVector): GpuVector =                                  # utils.nim:273-279
  result = newLike(y1)
  check(b200rk_hermite(result.h, x, x1, x2, y1.h, y2.h, dy1.h, dy2.h), y1.ctx)

# ---- ODEProc[GpuVector] closure -> b200rk_rhs_fn ------------------------------------------------------
type RhsEnv = object
  f: ODEProc[GpuVector]
  ctx: NumContext[GpuVector, float]
  dev: B200rkCtx
  err: ref Exception

proc rhsTrampoline(t: cdouble, y: VecHandle, dydt: VecHandle, user: pointer): cint {.cdecl.} =
  ## Calls the user's Nim closure with borrowed handles; the closure's GpuVector operators enqueue their
  ## kernels on the context stream; the result is copied into `dydt`. Exceptions never cross the ABI.
  let env = cast[ptr RhsEnv](user)
  try:
    let yv = GpuVector(h: y, ctx: env.dev, borrowed: true)
    let r = env.f(t.float, yv, env.ctx)
    if r.h != dydt: check(b200rk_vec_copy(dydt, r.h), env.dev)
    return 0
  except CatchableError as e:
    env.err = e
    return 1

proc toC(o: ODEoptions): COptions =
  COptions(dt: o.dt, dtMax: o.dtMax, dtMin: o.dtMin, tStart: o.tStart, absTol: o.absTol, relTol: o.relTol,
           scaleMax: o.scaleMax, scaleMin: o.scaleMin)

proc methodId(integrator: string): cint =
  check b200rk_method_from_name(integrator.cstring, addr result)   # ValueError "<name> is not a valid integrator" (ode.nim:651)

# ---- IntegratorProc[GpuVector] (ode.nim:38): one step of the named integrator ------------------------
proc gpuIntegrator*(integrator: string): IntegratorProc[GpuVector] =
  let mid = methodId(integrator)
  result = proc(f: ODEProc[GpuVector], t: float, y, FSAL: GpuVector, dt: float, options: ODEoptions,
                ctx: NumContext[GpuVector, float]): (GpuVector, GpuVector, float, float) =
    var env = RhsEnv(f: f, ctx: ctx, dev: y.ctx)
    var yNew = newLike(y)
    var fsalNew = newLike(y)
    var co = toC(options)
    var dtUsed, error: cdouble
    let rc = b200rk_step(y.ctx, mid, rhsTrampoline, addr env, t, y.h, FSAL.h, dt, addr co, yNew.h, fsalNew.h, addr dtUsed, addr error)
    if not env.err.isNil: raise env.err
    check(rc, y.ctx)
    (yNew, fsalNew, dtUsed.float, error.float)

# ---- solveODE (ode.nim:589-651) for T = GpuVector -----------------------------------------------------
proc solveODE*(f: ODEProc[GpuVector], y0: GpuVector, tspan: openArray[float],
               options: ODEoptions = newODEoptions(), ctx: NumContext[GpuVector, float] = nil,
               integrator = "dopri54"): (seq[float], seq[GpuVector]) =
  var nctx = ctx
  if nctx.isNil: nctx = newNumContext[GpuVector, float]()          # ode.nim:604-606
  let mid = methodId(integrator.toLower())
  var env = RhsEnv(f: f, ctx: nctx, dev: y0.ctx)
  var co = toC(options)
  var ts = @tspan
  var tOut = newSeq[cdouble](ts.len)
  var slots = newSeq[VecHandle](max(ts.len, 1))
  var nOut: csize_t
  let rc = b200rk_solve(y0.ctx, mid, rhsTrampoline, addr env, y0.h, cast[ptr cdouble](addr ts[0]), ts.len.csize_t, addr co,
                        cast[ptr cdouble](addr tOut[0]), addr slots[0], addr nOut, nil)
  if not env.err.isNil: raise env.err
  check(rc, y0.ctx)
  var ys = newSeq[GpuVector](nOut.int)
  for i in 0 ..< nOut.int: ys[i] = GpuVector(h: slots[i], ctx: y0.ctx, borrowed: false)
  while tOut.len > 0 and tOut[^1] != tOut[^1]: tOut.setLen(tOut.len - 1)   # tStart repeated in tspan: NaN-marked tail
  result = (@tOut, ys)

# ---- element-local right-hand side from source, fused into the RK kernels (b200rk_jit_rhs_new) -------
type JitRhs* = ref object
  ## dydt[i] = expr(t, y[i], p0[i].., c0..) as a CUDA C++ expression; compiled at run time INTO the fused kernels.
  fn: RhsFn
  user: pointer
  ctx: B200rkCtx
  vecs: seq[GpuVector]   # kept alive

proc newJitRhs*(expr: string, vecs: openArray[GpuVector] = [], scalars: openArray[float] = [],
                ctx: B200rkCtx = b200rkContext()): JitRhs =
  ## e.g. newJitRhs("c0*y*(1.0 - y/p0)", [K], [r]); a wrong expression raises ValueError with the compiler log.
  new(result, proc(r: JitRhs) = (if not r.user.isNil: discard b200rk_jit_rhs_free(r.user)))
  result.ctx = ctx
  result.vecs = @vecs                       # value semantics: the right-hand side owns its own copies of the parameters
  var hs = newSeq[VecHandle](vecs.len)
  for i in 0 ..< result.vecs.len: hs[i] = result.vecs[i].h
  var cs = @scalars
  check(b200rk_jit_rhs_new(ctx, expr.cstring, hs.len.cint, (if hs.len > 0: addr hs[0] else: nil), cs.len.cint,
                           (if cs.len > 0: cast[ptr cdouble](addr cs[0]) else: nil), addr result.fn, addr result.user), ctx)

proc newJitStencilRhs*(expr: string, radiusLeft, radiusRight: int, vecs: openArray[GpuVector] = [], scalars: openArray[float] = [],
                       ctx: B200rkCtx = b200rkContext()): JitRhs =
  ## dydt[i] = expr(t, Y(-radiusLeft)..Y(+radiusRight), p0[i].., c0..), Y(d) = y[(i + d) mod N]; e.g.
  ## newJitStencilRhs("((Y(1) - Y(-2)) * Y(-1) - Y(0)) + c0", 2, 1, scalars = [8.0]) is Lorenz-96. Used like a JitRhs.
  new(result, proc(r: JitRhs) = (if not r.user.isNil: discard b200rk_jit_rhs_free(r.user)))
  result.ctx = ctx
  result.vecs = @vecs
  var hs = newSeq[VecHandle](vecs.len)
  for i in 0 ..< result.vecs.len: hs[i] = result.vecs[i].h
  var cs = @scalars
  check(b200rk_jit_stencil_rhs_new(ctx, expr.cstring, radiusLeft.cint, radiusRight.cint, hs.len.cint, (if hs.len > 0: addr hs[0] else: nil),
                                   cs.len.cint, (if cs.len > 0: cast[ptr cdouble](addr cs[0]) else: nil), addr result.fn, addr result.user), ctx)

proc solveODE*(f: JitRhs, y0: GpuVector, tspan: openArray[float], options: ODEoptions = newODEoptions(),
               integrator = "dopri54"): (seq[float], seq[GpuVector]) =
  ## solveODE (ode.nim:589-651) with the right-hand side evaluated inside the fused kernels.
  let mid = methodId(integrator.toLower())
  var co = toC(options)
  var ts = @tspan
  var tOut = newSeq[cdouble](ts.len)
  var slots = newSeq[VecHandle](max(ts.len, 1))
  var nOut: csize_t
  check(b200rk_solve(y0.ctx, mid, f.fn, f.user, y0.h, cast[ptr cdouble](addr ts[0]), ts.len.csize_t, addr co,
                     cast[ptr cdouble](addr tOut[0]), addr slots[0], addr nOut, nil), y0.ctx)
  var ys = newSeq[GpuVector](nOut.int)
  for i in 0 ..< nOut.int: ys[i] = GpuVector(h: slots[i], ctx: y0.ctx, borrowed: false)
  while tOut.len > 0 and tOut[^1] != tOut[^1]: tOut.setLen(tOut.len - 1)
  result = (@tOut, ys)

# ---- consumers of a trajectory: hermiteInterpolate (utils.nim:282-312), cumtrapz / cumsimpson (integrate.nim) --
proc handles(v: openArray[GpuVector]): seq[VecHandle] =
  result = newSeq[VecHandle](v.len)
  for i, x in v: result[i] = x.h
proc adoptAll(ctx: B200rkCtx, slots: seq[VecHandle], n: csize_t): seq[GpuVector] =
  result = newSeq[GpuVector](n.int)
  for i in 0 ..< n.int: result[i] = GpuVector(h: slots[i], ctx: ctx, borrowed: false)

proc hermiteInterpolate*(x, t: openArray[float], y, dy: openArray[GpuVector]): seq[GpuVector] =
  ## Same name and argument order as utils.nim:282; all samples are evaluated by one kernel.
  if y.len == 0 or y.len != t.len or dy.len != t.len: raise newException(ValueError, "t, y and dy must have the same non-zero length")
  var (xs, ts, hy, hdy) = (@x, @t, handles(y), handles(dy))
  var slots = newSeq[VecHandle](max(xs.len, 1))
  var n: csize_t
  check(b200rk_hermite_interpolate(y[0].ctx, cast[ptr cdouble](addr xs[0]), xs.len.csize_t, cast[ptr cdouble](addr ts[0]), ts.len.csize_t,
                                   addr hy[0], addr hdy[0], addr slots[0], addr n), y[0].ctx)
  adoptAll(y[0].ctx, slots, n)

template cumulativeDiscrete(api: untyped, yArg: openArray[GpuVector], xArg: openArray[float]): seq[GpuVector] =
  if yArg.len == 0 or yArg.len != xArg.len: raise newException(ValueError, "X and Y must have the same non-zero length")
  var (xs, hy) = (@xArg, handles(yArg))
  var slots = newSeq[VecHandle](xs.len)
  var n: csize_t
  let dev = yArg[0].ctx
  check(api(dev, addr hy[0], cast[ptr cdouble](addr xs[0]), xs.len.csize_t, addr slots[0], addr n), dev)
  adoptAll(dev, slots, n)

proc cumtrapz*(Y: openArray[GpuVector], X: openArray[float]): seq[GpuVector] = cumulativeDiscrete(b200rk_cumtrapz, Y, X)      # integrate.nim:119-135
proc cumsimpson*(Y: openArray[GpuVector], X: openArray[float]): seq[GpuVector] = cumulativeDiscrete(b200rk_cumsimpson, Y, X)  # integrate.nim:330-378

type FnEnv = object
  f: NumContextProc[GpuVector, float]
  ctx: NumContext[GpuVector, float]
  err: ref Exception
proc fnTrampoline(t: cdouble, outVec: VecHandle, user: pointer): cint {.cdecl.} =
  let env = cast[ptr FnEnv](user)
  try:
    let r = env.f(t.float, env.ctx)
    result = b200rk_vec_copy(outVec, r.h)
  except CatchableError as e:
    env.err = e
    result = 1

template cumulativeFn(api: untyped, fArg, xArg, likeArg, ctxArg, dxArg: untyped): seq[GpuVector] =
  # (parameter names chosen not to collide with the field names used below: untyped parameters replace every
  # identifier of the same name, also after a dot and inside object constructors)
  var nctx = ctxArg
  if nctx.isNil: nctx = newNumContext[GpuVector, float]()
  var env = FnEnv(f: fArg, ctx: nctx)
  var xs = @xArg
  var slots = newSeq[VecHandle](max(xs.len, 1))
  var n: csize_t
  let dev = likeArg.ctx
  let rc = api(dev, fnTrampoline, addr env, likeArg.len.csize_t, cast[ptr cdouble](addr xs[0]), xs.len.csize_t, dxArg.cdouble, addr slots[0], addr n)
  if not env.err.isNil: raise env.err
  check(rc, dev)
  adoptAll(dev, slots, n)

proc cumtrapz*(f: NumContextProc[GpuVector, float], X: openArray[float], like: GpuVector,
               ctx: NumContext[GpuVector, float] = nil, dx = 1e-5): seq[GpuVector] =      # integrate.nim:138-175; `like` fixes T's size
  cumulativeFn(b200rk_cumtrapz_fn, f, X, like, ctx, dx)
proc cumsimpson*(f: NumContextProc[GpuVector, float], X: openArray[float], like: GpuVector,
                 ctx: NumContext[GpuVector, float] = nil, dx = 1e-5): seq[GpuVector] =    # integrate.nim:379-400
  cumulativeFn(b200rk_cumsimpson_fn, f, X, like, ctx, dx)

# ---- solveODE on HOST sequences (b200rk_solve_host): seq[float] in, seq[seq[float]] out -----------------
proc solveODEHost*(f: ODEProc[GpuVector], y0: openArray[float], tspan: openArray[float], options: ODEoptions = newODEoptions(),
                   ctx: NumContext[GpuVector, float] = nil, integrator = "dopri54", dev: B200rkCtx = b200rkContext()): (seq[float], seq[seq[float]]) =
  ## The call a maintainer makes when the state lives in host memory: y0 goes up, the states at `tspan` come down, the
  ## right-hand side closure sees device vectors (its GpuVector operators enqueue kernels). ode.nim:589-591 for T = seq[float].
  var nctx = ctx
  if nctx.isNil: nctx = newNumContext[GpuVector, float]()
  let mid = methodId(integrator.toLower())
  var env = RhsEnv(f: f, ctx: nctx, dev: dev)
  var co = toC(options)
  var (ts, y0s) = (@tspan, @y0)
  var tOut = newSeq[cdouble](ts.len)
  var yOut = newSeq[cdouble](max(ts.len, 1) * y0s.len)
  var nOut: csize_t
  let rc = b200rk_solve_host(dev, mid, rhsTrampoline, addr env, y0s.len.csize_t, cast[ptr cdouble](addr y0s[0]), cast[ptr cdouble](addr ts[0]),
                             ts.len.csize_t, addr co, cast[ptr cdouble](addr tOut[0]), cast[ptr cdouble](addr yOut[0]), addr nOut, nil)
  if not env.err.isNil: raise env.err
  check(rc, dev)
  var ys = newSeq[seq[float]](nOut.int)
  for i in 0 ..< nOut.int: ys[i] = yOut[i * y0s.len ..< (i + 1) * y0s.len]
  while tOut.len > 0 and tOut[^1] != tOut[^1]: tOut.setLen(tOut.len - 1)
  result = (@tOut, ys)

# ---- built-in device right-hand sides (b200rk_builtin_rhs_new): c*y, -(lambda .* y), Lorenz-96 -------------
type BuiltinRhs* = ref object
  fn: RhsFn
  user: pointer
  lam: GpuVector        # kept alive
proc newBuiltinRhs(kind: cint, scalar: float, lam: GpuVector, ctx: B200rkCtx): BuiltinRhs =
  new(result, proc(r: BuiltinRhs) = (if not r.user.isNil: discard b200rk_builtin_rhs_free(r.user)))
  result.lam = lam
  check(b200rk_builtin_rhs_new(ctx, kind, scalar, lam.h, addr result.fn, addr result.user), ctx)
proc rhsScale*(c: float, ctx: B200rkCtx = b200rkContext()): BuiltinRhs = newBuiltinRhs(0, c, GpuVector(), ctx)
proc rhsDiagLinear*(lam: GpuVector): BuiltinRhs = newBuiltinRhs(1, 0.0, lam, lam.ctx)
proc rhsLorenz96*(F = 8.0, ctx: B200rkCtx = b200rkContext()): BuiltinRhs = newBuiltinRhs(2, F, GpuVector(), ctx)

proc solveODE*(f: BuiltinRhs, y0: GpuVector, tspan: openArray[float], options: ODEoptions = newODEoptions(),
               integrator = "dopri54"): (seq[float], seq[GpuVector]) =
  let mid = methodId(integrator.toLower())
  var co = toC(options)
  var ts = @tspan
  var tOut = newSeq[cdouble](ts.len)
  var slots = newSeq[VecHandle](max(ts.len, 1))
  var nOut: csize_t
  check(b200rk_solve(y0.ctx, mid, f.fn, f.user, y0.h, cast[ptr cdouble](addr ts[0]), ts.len.csize_t, addr co,
                     cast[ptr cdouble](addr tOut[0]), addr slots[0], addr nOut, nil), y0.ctx)
  var ys = newSeq[GpuVector](nOut.int)
  for i in 0 ..< nOut.int: ys[i] = GpuVector(h: slots[i], ctx: y0.ctx, borrowed: false)
  while tOut.len > 0 and tOut[^1] != tOut[^1]: tOut.setLen(tOut.len - 1)
  result = (@tOut, ys)

# ---- the ODESolver loop as a resumable object (b200rk_solver_*): advance K accepted steps at a time ---------
type GpuSolver* = ref object
  h: SolverHandle
  ctx: B200rkCtx
  env: ref RhsEnv       # kept alive: the library calls back into it
proc newGpuSolver*(integrator: string, f: ODEProc[GpuVector], y0: GpuVector, tEnd: float, options: ODEoptions = newODEoptions(),
                   ctx: NumContext[GpuVector, float] = nil): GpuSolver =
  new(result, proc(s: GpuSolver) = (if not s.h.isNil: discard b200rk_solver_free(s.h)))
  var nctx = ctx
  if nctx.isNil: nctx = newNumContext[GpuVector, float]()
  result.ctx = y0.ctx
  new(result.env)
  result.env[] = RhsEnv(f: f, ctx: nctx, dev: y0.ctx)
  var co = toC(options)
  check(b200rk_solver_new(y0.ctx, methodId(integrator.toLower()), rhsTrampoline, addr result.env[], y0.h, tEnd, addr co, addr result.h), y0.ctx)
proc advance*(s: GpuSolver, maxSteps = -1): (int, bool) =
  var done: int64
  var fin: cint
  let rc = b200rk_solver_advance(s.h, maxSteps.int64, addr done, addr fin)
  if not s.env.err.isNil: raise s.env.err
  check(rc, s.ctx)
  (done.int, fin != 0)
proc state*(s: GpuSolver): (float, float, float, GpuVector) =
  var t, dtNext, err: cdouble
  var y: VecHandle
  check(b200rk_solver_state(s.h, addr t, addr dtNext, addr err, addr y), s.ctx)
  (t.float, dtNext.float, err.float, GpuVector(h: y, ctx: s.ctx, borrowed: true))
proc stats*(s: GpuSolver): tuple[steps, attempts, rejected, limiterHits, rhsEvals, launches, collectives: int] =
  var st: CStats
  check(b200rk_solver_stats(s.h, addr st), s.ctx)
  (st.steps.int, st.attempts.int, st.rejected.int, st.limiterHits.int, st.rhsEvals.int, st.launches.int, st.collectives.int)

# ---- IntegratorProc[GpuVector] values under the reference's step names (ode.nim:180, 237, 307, 377) -----------
# (the reference keeps its *_step procs private and selects them by name in solveODE; here they can also be passed around)
let
  RK4_step* = gpuIntegrator("rk4")
  DOPRI54_step* = gpuIntegrator("dopri54")
  TSIT54_step* = gpuIntegrator("tsit54")
  VERN65_step* = gpuIntegrator("vern65")

# ---- context knobs, sharding, device synchronisation ----------------------------------------------------------------
proc setKnob*(ctx: B200rkCtx, key: string, value: int) = check(b200rk_set(ctx, key.cstring, value.int64), ctx)
proc getKnob*(ctx: B200rkCtx, key: string): int =
  var v: int64
  check(b200rk_get(ctx, key.cstring, addr v), ctx)
  v.int
proc synchronize*(ctx: B200rkCtx) = check(b200rk_synchronize(ctx), ctx)
proc rank*(ctx: B200rkCtx): int = b200rk_rank(ctx).int
proc world*(ctx: B200rkCtx): int = b200rk_world(ctx).int
proc shardRange*(nGlobal, rank, world: int): (int, int) =
  var off, ln: csize_t
  check b200rk_shard_range(nGlobal.csize_t, rank.cint, world.cint, addr off, addr ln)
  (off.int, ln.int)
proc localLen*(v: GpuVector): int = b200rk_vec_local_len(v.h).int
proc localOffset*(v: GpuVector): int = b200rk_vec_local_offset(v.h).int
proc fill*(v: GpuVector, value: float) = check(b200rk_vec_fill(v.h, value), v.ctx)
proc uploadLocal*(v: GpuVector, host: openArray[float]) = check(b200rk_vec_upload_local(v.h, cast[ptr cdouble](unsafeAddr host[0])), v.ctx)
proc downloadLocal*(v: GpuVector): seq[float] =
  result = newSeq[float](v.localLen)
  if result.len > 0: check(b200rk_vec_download_local(v.h, cast[ptr cdouble](addr result[0])), v.ctx)
