"""Pin the C++ oracle: (1) its tableaux equal, bit for bit, the coefficients extracted mechanically from
the reference source (tests/golden/tableaux.json); (2) its trajectories and step sequences equal, bit for
bit, those of the independent pure-Python restatement (tests/golden/trajectories.json); (3) the tableaux
satisfy the Runge-Kutta order conditions and match scipy's Dormand-Prince pair."""
import numpy as np
import pytest
from conftest import assert_bitwise_equal, unhex

import oracle as O

PAIRS = ("dopri54", "tsit54", "vern65")


@pytest.mark.parametrize("method", PAIRS)
def test_oracle_tableau_matches_reference_literals(method, golden_tableaux):
    g = golden_tableaux[method]
    T = O.pair_tableau(method)
    S = T["stages"]
    for s in range(2, S + 1):
        assert T["c"][s].hex() == g[f"c{s}"].hex()
        for j in range(1, s):
            assert T["a"][s][j - 1].hex() == g[f"a{s}{j}"].hex(), (s, j)
    for j in range(1, T["n_b"] + 1):
        assert T["b"][j - 1].hex() == g[f"b{j}"].hex()
    for j in range(1, T["n_bhat"] + 1):
        assert T["bhat"][j - 1].hex() == g[f"bHat{j}"].hex()
    # every literal of the reference's const block is covered
    assert len(g) == (S - 1) + S * (S - 1) // 2 + T["n_b"] + T["n_bhat"]


@pytest.mark.parametrize("method,order_hi", [("dopri54", 5), ("tsit54", 5), ("vern65", 6)])
def test_tableau_order_conditions(method, order_hi):
    T = O.pair_tableau(method)
    S = T["stages"]
    A = np.zeros((S, S))
    for s in range(2, S + 1):
        A[s - 1, : s - 1] = T["a"][s][: s - 1]
    c = np.zeros(S)
    c[1:] = T["c"][2 : S + 1]
    b = np.zeros(S)
    b[: T["n_b"]] = T["b"][: T["n_b"]]
    # coefficients are rounded to double; cancellation scales with max|b| (176 for Vern65)
    tol = {"dopri54": 5e-14, "tsit54": 5e-13, "vern65": 2e-12}[method]
    assert np.allclose(A.sum(1), c, atol=1e-14)  # row sums
    # bushy-tree conditions b.c^(q-1) = 1/q and the second tree family b.A.c^(q-2) = 1/(q(q-1))
    for q in range(1, order_hi + 1):
        assert abs(b @ c ** (q - 1) - 1.0 / q) < tol, q
    for q in range(2, order_hi + 1):
        assert abs(b @ A @ c ** (q - 2) - 1.0 / (q * (q - 1))) < tol * 10, q
    bh = T["bhat"][:S].copy()
    if T["err_direct"]:  # Tsit54's bHat are difference weights: they sum to ~0 (ode.nim:346-352)
        assert abs(bh.sum()) < 1e-14
    else:  # embedded lower-order solution is consistent
        assert abs(bh.sum() - 1.0) < 1e-13
        for q in range(2, order_hi):
            assert abs(bh @ c ** (q - 1) - 1.0 / q) < 1e-11, q


def test_dopri54_matches_scipy_tableau():
    rk = pytest.importorskip("scipy.integrate._ivp.rk")
    T = O.pair_tableau("dopri54")
    A = rk.RK45.A
    for s in range(2, 7):
        assert np.allclose(T["a"][s][: s - 1], A[s - 1, : s - 1], rtol=0, atol=1e-16)
    assert np.allclose(T["b"][:6], rk.RK45.B, rtol=0, atol=1e-16)
    # scipy's E = bhat - b (7 entries; sign convention of its error estimator)
    b7 = np.append(T["b"][:6], 0.0)
    assert np.allclose(T["bhat"][:7] - b7, rk.RK45.E, rtol=0, atol=2e-16)


def _rhs_from(desc):
    if desc["kind"] == "scale":
        return O.rhs_scale(desc["c"])
    if desc["kind"] == "diag":
        return O.rhs_diag_linear(unhex(desc["lam"]))
    if desc["kind"] == "l96":
        return O.rhs_lorenz96(desc["F"])
    raise KeyError(desc)


def test_oracle_equals_python_restatement_bitwise(golden_trajectories):
    assert len(golden_trajectories) >= 20
    for name, g in golden_trajectories.items():
        opts = O.new_options(**g["options"])
        y0, ts = unhex(g["y0"]), unhex(g["tspan"])
        if g["vector"]:
            sol = O.solve_vector(g["integrator"], _rhs_from(g["rhs"]), y0, ts, opts, trace=True)
            t, y, st, tr = sol.t, sol.y, sol.stats, sol.trace
        else:
            t, y, st = O.solve_scalar(g["integrator"], float(y0[0]), ts, opts, rhs_scale_c=g["rhs"]["c"])
            y = y.reshape(-1, 1)
            tr = None
        assert_bitwise_equal(t, unhex(g["t"]), name + " t")
        gy = np.array([unhex(r) for r in g["y"]])
        assert_bitwise_equal(y, gy, name + " y")
        assert (st.steps, st.attempts, st.rejected, st.limiter_hits) == (g["steps"], g["attempts"], g["rejected"], g["limiter_hits"]), name
        if tr is not None:
            n = len(g["trace_dt"])
            assert_bitwise_equal([r[1] for r in tr[:n]], unhex(g["trace_dt"]), name + " dt sequence")
            assert_bitwise_equal([r[2] for r in tr[:n]], unhex(g["trace_err"]), name + " error sequence")
            assert [r[3] for r in tr[:n]] == g["trace_attempts"], name


def test_golden_covers_rejections_and_limiter(golden_trajectories):
    """The fixtures exercise the branches the reference's own tests never reach."""
    assert golden_trajectories["l96_40_tsit54"]["rejected"] > 0
    assert golden_trajectories["limiter4_dopri54"]["limiter_hits"] > 0
    assert any(a > 1 for a in golden_trajectories["limiter4_vern65"]["trace_attempts"])
    g = golden_trajectories["dense_diag8_dopri54"]
    assert len(g["y"]) == len(g["t"]) == 11


def test_survey_appendix_b_anchors(golden_trajectories):
    """Step counts reported in SURVEY.md Appendix B (derived there from a third, throw-away restatement)."""
    expect = {"diag8_dopri54": (29, 0), "diag8_tsit54": (26, 0), "diag8_vern65": (21, 0), "l96_40_tsit54": (23, 4),
              "limiter4_dopri54": (34, 25), "limiter4_tsit54": (31, 23), "limiter4_vern65": (25, 19)}
    for k, (steps, rej) in expect.items():
        assert (golden_trajectories[k]["steps"], golden_trajectories[k]["rejected"]) == (steps, rej), k
