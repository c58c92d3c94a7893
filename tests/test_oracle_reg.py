"""Oracle regularisation choosers: L-curve known answer, chi2 / MDP targets, GCV identities,
singular values against LAPACK gesdd (numpy), 1-D optimiser known answers.
Ported from test/nnls.jl:342-519, test/utils.jl:185-202, test/optimization.jl."""
import itertools

import numpy as np
import pytest

SIZES = [1, 2, 5, 8, 13, 16, 25, 32]


def rand_data(rng, m, n):
    x = rng.random(n)
    x[rng.integers(0, 2)::2] *= -1
    A = rng.random((m, n))
    return A, A @ x


# ---------------------------------------------------------------- optimisers (test/optimization.jl)
def test_brent_root_known_answers(orc):
    x, fx = orc.brent_root(np.sin, 3.0, 4.5, xatol=1e-6)
    assert abs(x - np.pi) <= 1e-6
    x, fx = orc.brent_root(np.sin, 2.5, 4.0, xrtol=1e-7)
    assert abs(x - np.pi) <= 1e-7 * np.pi
    x, fx = orc.brent_root(np.sin, 2.0, 3.5, ftol=1e-8)
    assert abs(fx) <= 1e-8


def test_brent_minimize_known_answers(orc):
    f = lambda x: np.sin(x) ** 2
    x, fx = orc.brent_minimize(f, 3.0, 4.5, xatol=1e-6, xrtol=0.0)
    assert abs(x - np.pi) <= 1e-6
    x, fx = orc.brent_minimize(f, 2.5, 4.0, xatol=0.0, xrtol=1e-7)
    assert abs(x - np.pi) <= 1e-7 * np.pi
    g = lambda x: np.exp(x) - x / 2
    x, fx = orc.brent_minimize(g, -2.0, 1.0, xatol=1e-8, xrtol=0.0)
    assert abs(x - np.log(0.5)) <= 1e-7


def test_bracket_root_monotonic(orc):
    f = lambda x: x - 2.3
    a, b, fa, fb = orc.bracket_root_monotonic(f, -4.0, 1.0, dilate=1.5, mono=+1, maxiters=6)
    assert a < 2.3 < b and fa < 0 < fb
    # steps: -4 -> -3 (d=1), then 1.5, 2.25, 3.375 ...
    assert a == pytest.approx(-4 + 1 + 1.5 + 2.25) and b == pytest.approx(-4 + 1 + 1.5 + 2.25 + 3.375)
    f = lambda x: x + 6.1  # root on the other side: steps go down
    a, b, fa, fb = orc.bracket_root_monotonic(f, -4.0, 1.0, dilate=1.5, mono=+1, maxiters=6)
    assert a < -6.1 < b
    # budget exhausted: no sign change within 1 + 6 steps
    f = lambda x: x - 1e6
    a, b, fa, fb = orc.bracket_root_monotonic(f, -4.0, 1.0, dilate=1.5, mono=+1, maxiters=6)
    assert fa * fb > 0


# ---------------------------------------------------------------- L-curve
def test_lcurve_corner_known_answer(orc):
    # test/nnls.jl:342-365: (xi, eta) = (mu, 1/mu) has its corner at log(mu) = 0
    f = lambda t: (np.exp(t), np.exp(-t))
    x, nf = orc.lcurve_corner(f, np.log(0.1), np.log(10.0), xtol=1e-6, Ptol=1e-6, Ctol=0.0)
    assert abs(x) <= 1e-3
    x2, _ = orc.lcurve_corner(f, np.log(0.1), np.log(10.0), xtol=1e-6, Ptol=1e-6, Ctol=0.0, backtracking=False)
    assert abs(x2) <= 1e-2


def test_lcurve_eval_count_bound(orc):
    # 4 initial points + one per golden-section step until 10 * phi^-k < 1e-4  (k ~ 24)
    f = lambda t: (np.log(1 + np.exp(2 * t)), -np.log(1 + np.exp(2 * t)) + 0.1 * t)
    x, nf = orc.lcurve_corner(f, -8.0, 2.0)
    assert 5 <= nf <= 64
    assert -8.0 <= x <= 2.0


@pytest.mark.parametrize("m,n", list(itertools.product(SIZES, SIZES)))
def test_lsqnonneg_lcurve_runs_everywhere(orc, m, n):
    rng = np.random.default_rng(m * 100 + n)
    A, b = rand_data(rng, m, n)
    R = orc.Reg(A, b)
    x, mu, chi2 = R.lcurve()
    assert np.all(x >= 0) and np.exp(-8) * (1 - 1e-12) <= mu <= np.exp(2) * (1 + 1e-12)
    # the returned x is the Tikhonov solution at the returned mu
    x2, r2, s2 = R.tikh(mu)
    np.testing.assert_allclose(x, x2, rtol=0, atol=0)


# ---------------------------------------------------------------- chi2 / MDP
@pytest.mark.parametrize("m,n", list(itertools.product(SIZES, SIZES)))
def test_lsqnonneg_chi2(orc, m, n):
    rng = np.random.default_rng(m * 31 + n)
    A, b = rand_data(rng, m, n)
    R = orc.Reg(A, b)
    x_unreg = R.none()
    res2_min = float(np.sum((A @ x_unreg - b) ** 2))
    res2_max = float(b @ b)
    target = min(np.sqrt(res2_min * res2_max) / res2_min if res2_min > 0 else 2.0, 1.01 + 0.99 * rng.random())
    x, mu, chi2, early = R.chi2(target)
    if res2_min <= 1e-12 or np.sum(x_unreg ** 2) == 0:
        assert mu >= 0
    else:
        assert mu > 0
        assert chi2 == pytest.approx(target, rel=1e-3)  # test/nnls.jl:423
        np.testing.assert_allclose(np.sum((A @ x - b) ** 2) / res2_min, chi2, rtol=1e-6)


@pytest.mark.parametrize("m,n", [(8, 5), (13, 8), (16, 16), (32, 25), (25, 32)])
def test_lsqnonneg_mdp(orc, m, n):
    rng = np.random.default_rng(m * 17 + n)
    A, b = rand_data(rng, m, n)
    b = b + 0.05 * rng.standard_normal(m)
    R = orc.Reg(A, b)
    x_unreg = R.none()
    res_min = np.linalg.norm(A @ x_unreg - b)
    res_max = np.linalg.norm(b)
    if 1e-8 < res_min < 0.9 * res_max:
        delta = np.sqrt(res_min * res_max)
        x, mu, chi2, early = R.mdp(delta)
        assert early == 0 and mu > 0
        assert np.sum((A @ x - b) ** 2) == pytest.approx(delta ** 2, rel=2e-3)  # test/nnls.jl:504
    # edge cases  test/nnls.jl:498-501
    x, mu, chi2, early = R.mdp(0.5 * res_min if res_min > 0 else 1e-300)
    assert early == 1 and mu == 0 and chi2 == 1
    np.testing.assert_array_equal(x, x_unreg)
    x, mu, chi2, early = R.mdp(2 * res_max)
    assert early == 2 and np.isinf(mu) and np.all(x == 0)


# ---------------------------------------------------------------- GCV / SVD
@pytest.mark.parametrize("m,n", list(itertools.product(SIZES, SIZES)))
def test_svdvals_match_lapack(orc, m, n):
    # the reference calls LAPACK dgesdd_ (src/utils.jl:117); numpy.linalg.svd is the same routine
    rng = np.random.default_rng(m * 13 + n)
    A = rng.random((m, n))
    ref = np.linalg.svd(A, compute_uv=False)
    np.testing.assert_allclose(orc.svdvals(A), ref, rtol=1e-10, atol=1e-13 * ref[0])


def test_svdvals_ill_conditioned_basis(orc):
    o = orc.make_t2map_opts((1, 1, 1), 48, 60, 8e-3)
    _, _, _, basis, _ = orc.setup_tables(o)
    A = basis[:, :, 40]
    ref = np.linalg.svd(A, compute_uv=False)
    np.testing.assert_allclose(orc.svdvals(A), ref, rtol=0, atol=1e-12 * ref[0])


@pytest.mark.parametrize("m,n", [(5, 5), (8, 5), (5, 8), (32, 13), (13, 32)])
def test_gcv_dof_trace_identity(orc, m, n):
    # test/nnls.jl:447: dof == tr(I - A (A'A + mu^2 I)^-1 A')
    rng = np.random.default_rng(m + 100 * n)
    A = rng.random((m, n))
    g = np.linalg.svd(A, compute_uv=False)
    for mu in [1e-3, 0.1, 1.0, 10.0]:
        H = A @ np.linalg.solve(A.T @ A + mu ** 2 * np.eye(n), A.T)
        dof = orc.lib().orc_gcv_dof(m, n, g.ctypes.data_as(orc.dp), mu)
        assert dof == pytest.approx(m - np.trace(H), rel=1e-8)


@pytest.mark.parametrize("m,n", [(8, 5), (16, 13), (32, 25), (25, 32), (13, 13)])
def test_lsqnonneg_gcv(orc, m, n):
    rng = np.random.default_rng(m * 3 + n)
    A, b = rand_data(rng, m, n)
    b = b + 0.02 * rng.standard_normal(m)
    R = orc.Reg(A, b)
    x, mu, chi2 = R.gcv()
    assert np.exp(-8) <= mu <= np.exp(2)
    g = np.linalg.svd(A, compute_uv=False)

    def loggcv(t):
        xx, r2, _ = R.tikh(np.exp(t))
        dof = max(m - n, 0) + np.sum(np.exp(2 * t) / (g ** 2 + np.exp(2 * t)))
        return np.log(max(r2 / dof ** 2, np.finfo(float).eps ** 2 / m))
    # returned mu is the best evaluated point of a Brent search: not worse than a coarse scan by much
    ts = np.linspace(-8, 2, 41)
    assert loggcv(np.log(mu)) <= min(loggcv(t) for t in ts) + 0.05
    assert chi2 >= 1 - 1e-9


# ---------------------------------------------------------------- Tikhonov helpers
def test_tikhonov_resnorm_and_seminorm(orc):
    rng = np.random.default_rng(11)
    A, b = rand_data(rng, 16, 13)
    R = orc.Reg(A, b)
    for mu in [1e-3, 1e-1, 1.0]:
        x, r2, s2 = R.tikh(mu)
        assert r2 == pytest.approx(np.sum((A @ x - b) ** 2), rel=1e-9)
        assert s2 == pytest.approx(np.sum(x ** 2), rel=1e-12)
