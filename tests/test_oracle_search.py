"""Oracle flip-angle search: Hermite interpolator minimiser, surrogate exactness, probe order,
bounding-box descent, and the NNLS loss gradient vs central differences.
Ported from test/splines.jl:165-203, 249-285, 344-388, 404-435."""
import numpy as np
import pytest


def hermite_eval(a, b, u0, u1, m0, m1, x):
    t = (x - a) / (b - a)
    h00, h10 = 2 * t ** 3 - 3 * t ** 2 + 1, t ** 3 - 2 * t ** 2 + t
    h01, h11 = -2 * t ** 3 + 3 * t ** 2, t ** 3 - t ** 2
    return h00 * u0 + h10 * (b - a) * m0 + h01 * u1 + h11 * (b - a) * m1


def test_hermite_minimize_not_worse_than_dense_grid(orc):
    # test/splines.jl:249-285: 1000 random boundary conditions, 1025-point grid
    rng = np.random.default_rng(0)
    for _ in range(1000):
        a = rng.uniform(-2, 2)
        b = a + rng.uniform(0.1, 3)
        u0, u1, m0, m1 = rng.standard_normal(4) * rng.choice([0.1, 1.0, 10.0])
        x, u = orc.hermite_minimize(a, b, u0, u1, m0, m1)
        assert a <= x <= b
        xs = np.linspace(a, b, 1025)
        us = hermite_eval(a, b, u0, u1, m0, m1, xs)
        assert u <= us.min() + 1e-9 * max(1.0, np.abs(us).max())
        assert u == pytest.approx(hermite_eval(a, b, u0, u1, m0, m1, x), rel=1e-9, abs=1e-9)


def test_probe_order_default_grid(orc):
    # seeds are grid indices 1, 64, 32, 16, 48 (splines.jl:716-734)
    grid = orc.linrange(50.0, 180.0, 64)
    f = lambda I: ((grid[I - 1] - 120.0) ** 2, 2 * (grid[I - 1] - 120.0))
    x, u, order = orc.surrogate_search(f, grid, 5, 64)
    assert order[:5] == [1, 64, 32, 16, 48]
    assert len(order) == len(set(order))
    assert abs(x - 120.0) < 1e-9 and abs(u) < 1e-12  # quadratic is reproduced exactly by Hermite cubics


def test_surrogate_exact_on_cubic(orc):
    # test/splines.jl:165-203
    grid = orc.linrange(50.0, 180.0, 64)
    c = [3.0, -0.5, 0.004, 1e-5]
    p = lambda x: c[0] + c[1] * (x - 100) + c[2] * (x - 100) ** 2 + c[3] * (x - 100) ** 3
    dp = lambda x: c[1] + 2 * c[2] * (x - 100) + 3 * c[3] * (x - 100) ** 2
    f = lambda I: (p(grid[I - 1]), dp(grid[I - 1]))
    x, u, order = orc.surrogate_search(f, grid, 5, 64)
    xs = np.linspace(50, 180, 200001)
    assert u == pytest.approx(p(xs).min(), abs=1e-9)
    assert x == pytest.approx(xs[np.argmin(p(xs))], abs=1e-3)


@pytest.mark.parametrize("xstar", [50.0, 51.0, 83.7, 114.5, 115.0, 147.2, 179.9, 180.0])
def test_search_converges_to_bracketing_box(orc, xstar):
    grid = orc.linrange(50.0, 180.0, 64)
    f = lambda I: ((grid[I - 1] - xstar) ** 2, 2 * (grid[I - 1] - xstar))
    x, u, order = orc.surrogate_search(f, grid, 5, 64)
    assert abs(x - xstar) < 1e-9
    # halving 63 -> 32/31 -> 16 -> 8 -> 4 -> 2 -> 1: at most 5 seeds + ~2 per level
    assert 5 <= len(order) <= 5 + 2 * 6
    # the final box has width 1 and contains x*: its corners were probed
    k = min(int(np.searchsorted(grid, xstar, side="right")), 63)
    assert k in order or (k + 1) in order


def test_all_points_when_mineval_equals_maxeval(orc):
    grid = orc.linrange(0.0, 7.0, 8)
    f = lambda I: (np.cos(grid[I - 1]), -np.sin(grid[I - 1]))
    x, u, order = orc.surrogate_search(f, grid, 8, 8)
    assert sorted(order) == list(range(1, 9))


def test_maxeval_budget_respected(orc):
    grid = orc.linrange(50.0, 180.0, 64)
    f = lambda I: (np.sin(grid[I - 1] / 3.0), np.cos(grid[I - 1] / 3.0) / 3.0)
    x, u, order = orc.surrogate_search(f, grid, 5, 7)
    assert len(order) <= 7


def test_nnls_loss_gradient_vs_central_differences(orc):
    # test/splines.jl:344-388, rtol 1e-6: d/dalpha ||A(alpha) x+ - b||^2 with the analytic basis Jacobian
    nTE, nT2, TE = 32, 40, 10e-3
    o = orc.make_t2map_opts((1, 1, 1), nTE, nT2, TE)
    _, t2, ang, basis, dbasis = orc.setup_tables(o)
    img = orc.mock_image(4, nTE, TE, seed=7)
    for v in range(4):
        b = img[v] / img[v].max()
        for k in [5, 20, 40, 55]:
            A, dA = basis[:, :, k], dbasis[:, :, k]
            r = orc.nnls(A, b, warm=True)
            P = r.x > 0
            grad = 2 * (dA[:, P] @ r.x[P]) @ (A[:, P] @ r.x[P] - b)

            def loss(alpha):
                Aa = np.column_stack([orc.epg(nTE, alpha, TE, t2[j], 1.0) for j in range(nT2)])
                rr = orc.nnls(Aa, b, warm=True)
                return rr.rnorm ** 2
            h = 1e-4
            fd = (loss(ang[k] + h) - loss(ang[k] - h)) / (2 * h)
            assert grad == pytest.approx(fd, rel=1e-4, abs=1e-10)
