"""Oracle checks for the `legacy = true` algorithms (src/splines.jl:419-500, src/lsqnonneg.jl:595-636).

The interpolating spline is third-party FITPACK (Dierckx.jl); scipy.interpolate wraps the same Fortran
routines (curfit / splev), so splrep(..., s=0) / splev pin the restatement in oracle/orc_spline.c."""
import numpy as np
import pytest
from scipy import interpolate

import orc


@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("m", [2, 3, 4, 5, 8, 13, 32, 64])
def test_fitpack_interpolating_spline_matches_scipy(m, k):
    if m <= k:
        pytest.skip("curfit needs m > k")
    rng = np.random.default_rng(100 * m + k)
    x = np.sort(rng.uniform(50, 180, m)) if m != 8 else np.linspace(50, 180, 8)
    y = rng.standard_normal(m)
    t, c = orc.fitpack_interp(x, y, k)
    ts, cs, _ = interpolate.splrep(x, y, k=k, s=0)
    assert np.array_equal(t, ts)  # knot placement of the interpolating spline (fpcurf)
    assert np.allclose(c, cs[:m], rtol=1e-11, atol=1e-12)
    xs = np.linspace(x[0], x[-1], 2001)
    ys = orc.fitpack_splev(t, c, k, xs)
    assert np.allclose(ys, interpolate.splev(xs, (ts, cs, k)), rtol=1e-10, atol=1e-11)
    assert np.allclose(orc.fitpack_splev(t, c, k, x), y, rtol=1e-10, atol=1e-11)  # interpolates the data


def test_exponentially_spaced_abscissae_like_the_chi2_search():
    # mu_cache = [0, 1e-3, 2e-3, ...] (src/lsqnonneg.jl:597-609): poorly conditioned but well defined
    for n in (2, 3, 4, 6, 10, 14):
        x = np.concatenate([[0.0], 1e-3 * 2.0 ** np.arange(n - 1)])
        y = 1.0 + 0.05 * x ** 0.7
        k = min(3, n - 1)
        t, c = orc.fitpack_interp(x, y, k)
        ts, cs, _ = interpolate.splrep(x, y, k=k, s=0)
        assert np.array_equal(t, ts)
        assert np.allclose(c, cs[:n], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("start,step,stop,n,last", [
    (50.0, 0.001, 180.0, 130001, 180.0),         # the default flip-angle grid
    (0.0, 0.001, 0.064, 65, 0.064),              # mu_cache after six doublings
    (0.0, 0.001, 0.001, 2, 0.001),
    (100.0, 0.001, 180.0, 80001, 180.0),
    (0.0, 0.1, 1.0, 11, 1.0),                    # Julia docs: 0:0.1:1 has 11 elements ending at 1.0
    (1.0, 0.001, 1.0005, 1, 1.0),
])
def test_julia_float_range(start, step, stop, n, last):
    v = orc.jl_range_values(start, step, stop)
    assert len(v) == n
    assert v[0] == start and v[-1] == last
    # every element is the correctly rounded rational start + i*step (what TwicePrecision delivers)
    from fractions import Fraction
    fs, ft = Fraction(start), Fraction(str(step))
    for i in (0, 1, 2, 7, n // 3, n // 2, n - 2, n - 1):
        if 0 <= i < n:
            assert v[i] == float(fs + i * ft)


def test_julia_range_known_values():
    # collect(0:0.1:0.5) in Julia prints [0.0, 0.1, 0.2, 0.3, 0.4, 0.5] (not 0.30000000000000004)
    assert orc.jl_range_values(0.0, 0.1, 0.5).tolist() == [0.0, 0.1, 0.2, 0.3, 0.4, 0.5]
    # irrational-looking step: falls back to the literal start + i*step with the documented length rule
    r = orc.jl_range(0.0, np.pi / 7, 3.0)
    assert not r.rational and r.len == 7


def test_spline_opt_legacy_is_the_sampled_minimum():
    rng = np.random.default_rng(7)
    X = np.linspace(50, 180, 8)
    for trial in range(5):
        Y = (X - rng.uniform(60, 170)) ** 2 * 1e-4 + rng.uniform(0, 1e-3, 8)
        x, y = orc.spline_opt_legacy(X, Y)
        tck = interpolate.splrep(X, Y, k=3, s=0)
        xs = (50000 + np.arange(130001)) / 1000.0
        ys = interpolate.splev(xs, tck)
        i = int(np.argmin(ys))
        assert abs(x - xs[i]) <= 0.001 + 1e-12 and abs(y - ys[i]) <= 1e-12 * max(1, abs(ys[i]))
        assert round(x * 1000) / 1000 == x  # a grid point of knots[1]:0.001:knots[end]
    # two and three points: degree min(3, n-1)
    assert orc.spline_opt_legacy([50.0, 180.0], [2.0, 1.0])[0] == 180.0
    x3, y3 = orc.spline_opt_legacy([50.0, 115.0, 180.0], [1.0, 0.0, 1.0])
    assert x3 == 115.0 and abs(y3) < 1e-12


def test_spline_root_legacy_nearest_sample():
    X = np.array([0.0, 1e-3, 2e-3, 4e-3, 8e-3, 16e-3, 32e-3])
    Y = 1.0 + 3.0 * X  # exactly linear data: the cubic interpolant is the line
    x = orc.spline_root_legacy(X, Y, 1.02 * 1.0)
    assert x == pytest.approx(0.007, abs=1e-15) or x == pytest.approx(0.006, abs=1e-15)
    assert abs((1 + 3 * x) - 1.02) <= 0.0015 * 3


def test_chi2_search_legacy_doubling():
    calls = []

    def f(mu):
        calls.append(mu)
        return 1.0 + 40.0 * mu * mu

    mu, r2 = orc.chi2_search_legacy(f, 1.0, 1.02)
    # doubling 1e-3, 2e-3, ... until 1 + 40 mu^2 >= 1.02  -> mu = 0.032, then one more call at the root
    assert calls[:6] == [1e-3 * 2 ** i for i in range(6)] and len(calls) == 7
    assert calls[-1] == mu and r2 == f(mu)
    assert abs(mu - np.sqrt(0.02 / 40)) <= 2e-3  # spline root sampled on the 0.001 grid
    assert round(mu * 1000) / 1000 == mu


def test_legacy_surrogate_search_evaluates_every_angle():
    grid = np.linspace(50, 180, 8)
    u = (grid - 141.3) ** 2 * 1e-5 + 1e-3
    x, ux, order = orc.surrogate_search(lambda I: (u[I - 1], 0.0), grid, 8, 8, legacy=True)
    assert sorted(order) == list(range(1, 9))
    xs, us = orc.spline_opt_legacy(grid, u)
    assert (x, ux) == (xs, us)
    assert abs(x - 141.3) < 0.5


@pytest.mark.parametrize("Reg,extra", [("none", {}), ("chi2", {"Chi2Factor": 1.02}), ("lcurve", {})])
def test_legacy_pipeline(Reg, extra):
    nvox, nTE, nT2, TE = 48, 32, 40, 10e-3
    img = orc.mock_image(nvox, nTE, TE, seed=11)
    o = orc.make_t2map_opts((nvox, 1, 1), nTE, nT2, TE, Reg=Reg, legacy=True, nRefAngles=8, nRefAnglesMin=8, **extra)
    p = orc.make_t2part_opts((nvox, 1, 1), nT2)
    res, st = orc.t2map(img, o, p)
    o2 = orc.make_t2map_opts((nvox, 1, 1), nTE, nT2, TE, Reg=Reg, **extra)
    ref, _ = orc.t2map(img, o2, p)
    a = res["alpha"]
    assert np.all(np.isfinite(a)) and np.all((a >= 50) & (a <= 180))
    assert np.allclose(np.round(a * 1000) / 1000, a, atol=1e-9)       # sampled on the 0.001 degree grid
    assert np.median(np.abs(a - ref["alpha"])) < 2.0                  # same minimum as the modern search, roughly
    assert np.all(res["dist"] >= 0) and np.all(np.isfinite(res["gdn"]))
    if Reg == "chi2":
        mu = res["mu"]
        assert np.allclose(np.round(mu * 1000) / 1000, mu, atol=1e-12)  # legacy mu is a multiple of 0.001
        ok = mu > 0
        assert ok.mean() > 0.5
        assert np.all(res["chi2factor"][ok] > 1.0) and np.median(res["chi2factor"][ok]) < 1.2


def test_four_points_make_the_spline_an_exact_surrogate_of_a_cubic():
    """test/splines.jl:123-163 (test_cubic_spline_surrogate) for the legacy suggest_point: with four points the
    FITPACK spline is the cubic itself, so the sampled minimum sits within half a grid step of the analytic one."""
    x = np.linspace(-0.5, 2.0, 4)
    f = lambda t: 3 * t ** 3 - 5 * t ** 2 - t - 2  # noqa: E731
    u = f(x)
    xs, us, order = orc.surrogate_search(lambda I: (u[I - 1], 0.0), x, 4, 4, legacy=True)
    assert sorted(order) == [1, 2, 3, 4]
    xtrue = (10 + np.sqrt(100 + 36)) / 18  # root of 9x^2 - 10x - 1 inside the interval, a minimum (f'' > 0)
    assert abs(xs - xtrue) <= 0.0005 + 1e-12
    assert abs(us - f(xs)) <= 1e-12 and us <= f(xtrue) + 9 * 0.0005 ** 2  # f'' = 18x - 10 < 18 around xtrue
    t, c = orc.fitpack_interp(x, u, 3)
    grid = np.linspace(-0.5, 2.0, 1001)
    assert np.allclose(orc.fitpack_splev(t, c, 3, grid), f(grid), rtol=0, atol=1e-12)


@pytest.mark.parametrize("npts,deg", [(n, d) for n in range(2, 6) for d in range(1, min(n - 1, 3) + 1)])
def test_cubic_splines_minimum_and_root(npts, deg):
    """test/splines.jl:100-121 (test_cubic_splines) restated for the sampled legacy searches; the legacy functions
    always take deg_spline = min(3, npts - 1) (src/splines.jl:419, 446), so only that degree applies."""
    if deg != min(3, npts - 1):
        pytest.skip("legacy searches use deg_spline = min(3, npts - 1)")
    rng = np.random.default_rng(10 * npts + deg)
    X = np.linspace(-0.5, 2.0, npts)
    Y = rng.standard_normal(npts)
    t, c = orc.fitpack_interp(X, Y, deg)
    x, y = orc.spline_opt_legacy(X, Y)
    assert X[0] <= x <= X[-1]
    assert abs(orc.fitpack_splev(t, c, deg, [x])[0] - y) <= 1e-14 * max(1.0, abs(y))
    dense = orc.fitpack_splev(t, c, deg, np.linspace(X[0], X[-1], 100))
    assert dense.min() >= y - 5e-3  # 100 points against 2501 samples: the sampled minimum can only be lower, up to slope * step
    ybar = (Y.min() + Y.max()) / 2
    xbar = orc.spline_root_legacy(X, Y, ybar)
    assert X[0] <= xbar <= X[-1]
    slope = np.abs(np.diff(orc.fitpack_splev(t, c, deg, np.linspace(X[0], X[-1], 2501)))).max() / 0.001
    assert abs(orc.fitpack_splev(t, c, deg, [xbar])[0] - ybar) <= slope * 0.0005 + 1e-9
