import time, numpy as np, sys
sys.path.insert(0, "/root/repo")
import numericalnim_b200 as nn
ctx = nn.default_context()
n, m = 1 << 23, 33
rng = np.random.default_rng(0)
base = rng.uniform(-1, 1, n)
X = np.linspace(0, 2, m)
Y = [nn.newVector(np.roll(base, k)) for k in range(m)]
dY = [nn.newVector(np.roll(base, k + 5)) for k in range(m)]
xs = np.sort(rng.uniform(0, 2, 2 * m))
for name, fn in (("hermite", lambda: nn.hermiteInterpolate(xs, X, Y, dY)), ("cumtrapz", lambda: nn.cumtrapz(Y, X)), ("cumsimpson", lambda: nn.cumsimpson(Y, X))):
    ts = []
    for it in range(8):
        ctx.synchronize(); t0 = time.perf_counter()
        out = fn()
        t1 = time.perf_counter(); ctx.synchronize(); t2 = time.perf_counter()
        for v in out: v.free()
        t3 = time.perf_counter()
        ts.append((round(1e3*(t1-t0),2), round(1e3*(t2-t1),2), round(1e3*(t3-t2),2)))
    print(name, "call/sync/free ms:", ts)
