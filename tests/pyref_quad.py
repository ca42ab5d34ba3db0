"""Independent pure-Python restatement (floats only, no numpy arithmetic) of the reference's trajectory consumers for
T = float: hermiteSpline / hermiteInterpolate (utils.nim:273-312), sortAndTrimDataset (utils.nim:360-420),
cumtrapz (integrate.nim:119-175) and cumsimpson (integrate.nim:330-400). Written from the Nim source, separately from
oracle/quad_oracle.hpp, so that the two restatements can be compared bit for bit (tests/test_oracle_quadrature.py) —
the same role tests/pyref.py plays for the ODE path. TEST INFRASTRUCTURE ONLY."""
import math


def ipow(x: float, y: int) -> float:  # Nim math.`^`
    if y == 0:
        return 1.0
    if y == 1:
        return x
    if y == 2:
        return x * x
    if y == 3:
        return x * x * x
    r = 1.0
    while True:
        if y & 1:
            r *= x
        y >>= 1
        if y == 0:
            break
        x *= x
    return r


def hermite_spline(x, x1, x2, y1, y2, dy1, dy2):  # utils.nim:273-279
    t = (x - x1) / (x2 - x1)
    h00 = (1.0 + 2.0 * t) * ipow(1.0 - t, 2)
    h10 = t * ipow(1.0 - t, 2)
    h01 = ipow(t, 2) * (3.0 - 2.0 * t)
    h11 = ipow(t, 3) - ipow(t, 2)
    return h00 * y1 + h10 * (x2 - x1) * dy1 + h01 * y2 + h11 * (x2 - x1) * dy2


def hermite_interpolate(x, t, y, dy):  # utils.nim:282-312
    result, xi = [], 0
    if all(x[i] <= x[i + 1] for i in range(len(x) - 1)):
        for i in range(0, len(t) - 1):
            while t[i] <= x[xi] and x[xi] < t[i + 1]:
                result.append(hermite_spline(x[xi], t[i], t[i + 1], y[i], y[i + 1], dy[i], dy[i + 1]))
                xi += 1
                if len(x) - 1 < xi:
                    break
            if len(x) - 1 < xi:
                break
        if x[-1] == t[-1]:
            result.append(y[-1])
    else:
        for a in x:
            for i in range(0, len(t) - 1):
                if t[i] <= a and a < t[i + 1]:
                    result.append(hermite_spline(a, t[i], t[i + 1], y[i], y[i + 1], dy[i], dy[i + 1]))
                    break
            else:
                if a == t[-1]:
                    result.append(y[-1])
                else:
                    raise ValueError(f"{a} not in interval {min(t)} - {max(t)}")
    return result


def sort_and_trim(X, Y):  # utils.nim:385-420 via 360-383
    order = sorted(range(len(X)), key=lambda i: (X[i], i))
    xs, ys = [X[i] for i in order], [Y[i] for i in order]
    groups = {}
    for i, v in enumerate(xs):
        groups.setdefault(v, []).append(i)
    delete = set()
    for idx in groups.values():
        if len(idx) > 1:
            for i in idx:
                if ys[i] != ys[idx[0]]:
                    raise ValueError("impure y-duplicates was found")
            delete.update(idx[1:])
    keep = [i for i in range(len(xs)) if i not in delete]
    return [xs[i] for i in keep], [ys[i] for i in keep]


def cumtrapz(Y, X):  # integrate.nim:119-135
    xs, ys = sort_and_trim(list(X), list(Y))
    result = [ys[0] - ys[0]]
    integral = ys[0] - ys[0]
    for i in range(0, len(xs) - 1):
        integral += 0.5 * (xs[i + 1] - xs[i]) * (ys[i + 1] + ys[i])
        result.append(integral)
    return result


def cumsimpson(Y, X):  # integrate.nim:330-378
    xs, ys = sort_and_trim(list(X), list(Y))
    N = len(xs)
    even = False
    if N < 3:
        raise ValueError("X and Y must have at least 3 elements to perform Simpson, use cumtrapz instead")
    if N % 2 == 0:
        even = True
        N -= 1
    integral = ys[0] - ys[0]
    y, dy, knots = [integral], [ys[0]], [xs[0]]
    for i in range(0, int((N - 1) / 2)):
        h1 = xs[2 * i + 1] - xs[2 * i]
        h2 = xs[2 * i + 2] - xs[2 * i + 1]
        alpha = (2.0 * ipow(h2, 3) - ipow(h1, 3) + 3.0 * h1 * ipow(h2, 2)) / (6.0 * h2 * (h2 + h1))
        beta = (ipow(h2, 3) + ipow(h1, 3) + 3.0 * h1 * h2 * (h2 + h1)) / (6.0 * h2 * h1)
        eta = (2.0 * ipow(h1, 3) - ipow(h2, 3) + 3.0 * h2 * ipow(h1, 2)) / (6.0 * h1 * (h2 + h1))
        integral += alpha * ys[2 * i + 2] + beta * ys[2 * i + 1] + eta * ys[2 * i]
        y.append(integral)
        dy.append(ys[2 * i + 2])
        knots.append(xs[2 * i + 2])
    if even:
        last = len(xs) - 1
        h1 = xs[last - 1] - xs[last - 2]
        h2 = xs[last] - xs[last - 1]
        alpha = (2.0 * ipow(h2, 2) + 3.0 * h1 * h2) / (6.0 * (h1 + h2))
        beta = (ipow(h2, 2) + 3.0 * h1 * h2) / (6.0 * h1)
        eta = -(ipow(h2, 3)) / (6.0 * h1 * (h1 + h2))
        integral += eta * ys[last - 2] + beta * ys[last - 1] + alpha * ys[last]
        y.append(integral)
        dy.append(ys[last])
        knots.append(xs[last])
    return hermite_interpolate(list(X), knots, y, dy)


def cumtrapz_fn(f, X, dx=1e-5):  # integrate.nim:138-175
    times, dy, y = [], [], []
    t = min(X)
    t_end = max(X) + 1.0
    dy_temp = f(t)
    integral = dy_temp - dy_temp
    times.append(t); dy.append(dy_temp); y.append(integral)
    t += dx
    while t <= t_end:
        dy_prev = dy_temp
        dy_temp = f(t)
        integral += 0.5 * dx * (dy_prev + dy_temp)
        times.append(t); dy.append(dy_temp); y.append(integral)
        t += dx
    return hermite_interpolate(list(X), times, y, dy)


def linspace(x1, x2, n):  # utils.nim:498-507
    dx = (x2 - x1) / float(n - 1)
    return [x1] + [x1 + dx * float(i) for i in range(1, n - 1)] + [x2]


def nim_to_int(x: float) -> int:  # system.toInt: round half away from zero
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def cumsimpson_fn(f, X, dx=1e-5):  # integrate.nim:379-400
    t = linspace(min(X), max(X), nim_to_int((max(X) - min(X)) / dx) + 2)
    dy = [f(x) for x in t]
    ys = cumsimpson(dy, t)
    return hermite_interpolate(list(X), t, ys, dy)
