"""Oracle NNLS: the invariant battery of test/nnls.jl:30-328 (KKT, permutations, QR/Cholesky
identities, residual norm), plus agreement with an independent Lawson-Hanson
implementation (scipy.optimize.nnls)."""
import itertools

import numpy as np
import pytest
import scipy.optimize

SIZES = [1, 2, 5, 8, 13, 16, 25, 32]  # test/nnls.jl:1
MUS = [0.0, 1e-6, 1e-2, 10.0, 1e4]     # test/nnls.jl:331


def rand_data(rng, m, n):  # test/nnls.jl:3-9
    x = rng.random(n)
    x[rng.integers(0, 2)::2] *= -1
    A = rng.random((m, n))
    return A, A @ x


def check_solution(A0, b0, mu, r, n):
    """A0, b0: the (possibly padded) problem actually solved; r: oracle result."""
    m = A0.shape[0]
    assert r.mode == 0
    assert sorted(r.idx.tolist()) == list(range(1, n + 1))
    assert np.all(r.idx[r.invidx - 1] == np.arange(1, n + 1))
    npos = r.nsetp
    P = r.idx[:npos] - 1
    Z = r.idx[npos:] - 1
    x = r.x
    assert np.all(x[P] > 0) and np.all(x[Z] == 0)
    # x+ == least squares on the active columns
    if npos:
        xls = np.linalg.lstsq(A0[:, P], b0, rcond=None)[0]
        np.testing.assert_allclose(x[P], xls, rtol=1e-7, atol=1e-10 * max(1.0, np.abs(xls).max()))
    # dual
    w_true = -A0.T @ (A0 @ x - b0)
    scale = max(1.0, np.abs(A0).max() * np.abs(b0).max() * m)
    assert np.all(w_true[Z] <= 1e-9 * scale)
    assert np.all(np.abs(w_true[P]) <= 1e-8 * scale)
    if npos < min(m, n) and npos < m:
        np.testing.assert_allclose(r.w[r.invidx - 1][Z], w_true[Z], rtol=1e-6, atol=1e-9 * scale)
    # residual norm  test/nnls.jl:128
    np.testing.assert_allclose(r.rnorm, np.linalg.norm(A0 @ x - b0), rtol=1e-9, atol=1e-12 * scale)
    # U = work.A[1:n+, 1:n+], U x+ = b[1:n+], U'U = A+'A+   test/nnls.jl:121-123, 150-170
    if npos:
        U = np.triu(r.A[:npos, :npos])
        np.testing.assert_allclose(U @ x[P], r.b[:npos], rtol=1e-9, atol=1e-12 * scale)
        G = A0[:, P].T @ A0[:, P]
        np.testing.assert_allclose(U.T @ U, G, rtol=1e-9, atol=1e-12 * np.abs(G).max())
        R = np.linalg.qr(A0[:, P], mode="r")
        np.testing.assert_allclose(np.abs(np.diag(U)), np.abs(np.diag(R)), rtol=1e-7)


@pytest.mark.parametrize("m,n", list(itertools.product(SIZES, SIZES)))
def test_nnls_invariants(orc, m, n):
    rng = np.random.default_rng(1000 * m + n)
    A, b = rand_data(rng, m, n)
    for mu in MUS:
        if mu > 0:
            Ap = np.vstack([A, mu * np.eye(n)])
            bp = np.concatenate([b, np.zeros(n)])
            r = orc.nnls(A, b, mu=mu)               # lazily padded Tikhonov variant (NNLS.jl:827-1061)
            r_dense = orc.nnls(Ap, bp)              # plain algorithm on the explicit padded system
            check_solution(Ap, bp, mu, r, n)
            np.testing.assert_allclose(r.x, r_dense.x, rtol=1e-7, atol=1e-10 * max(1.0, np.abs(r_dense.x).max()))
            np.testing.assert_allclose(r.rnorm, r_dense.rnorm, rtol=1e-9, atol=1e-12)
            # Tikhonov stationarity  test/nnls.jl:102
            P = r.idx[:r.nsetp] - 1
            g = A[:, P].T @ (A[:, P] @ r.x[P] - b) + mu ** 2 * r.x[P]
            assert np.all(np.abs(g) <= 1e-8 * max(1.0, mu ** 2) * max(1.0, np.abs(b).max()) * m * n)
        else:
            r = orc.nnls(A, b)
            check_solution(A, b, 0.0, r, n)


@pytest.mark.parametrize("m,n", [(5, 8), (8, 5), (13, 13), (32, 16), (16, 32), (48, 40), (32, 60)])
def test_against_scipy_nnls(orc, m, n):
    rng = np.random.default_rng(m * 77 + n)
    for _ in range(5):
        A, b = rand_data(rng, m, n)
        b = b + 0.01 * rng.standard_normal(m)  # generic position: unique solution
        r = orc.nnls(A, b)
        xs, rn = scipy.optimize.nnls(A, b, maxiter=50 * n)
        # both are exact active-set methods: same objective value; same x when A has full column rank
        assert abs(r.rnorm - rn) <= 1e-9 * max(1.0, rn)
        if m >= n:
            np.testing.assert_allclose(r.x, xs, rtol=1e-6, atol=1e-9)


def test_warm_start_driver_equals_cold_solution(orc):
    # lsqnonneg.jl:30-84 only changes the first pivot; the minimiser is the same
    rng = np.random.default_rng(5)
    for (m, n) in [(8, 5), (16, 13), (32, 25), (48, 40)]:
        A, b = rand_data(rng, m, n)
        b = b + 0.01 * rng.standard_normal(m)
        cold, warm = orc.nnls(A, b), orc.nnls(A, b, warm=True)
        np.testing.assert_allclose(warm.x, cold.x, rtol=1e-7, atol=1e-10)
        np.testing.assert_allclose(warm.rnorm, cold.rnorm, rtol=1e-10)
        for mu in [1e-3, 1e-1, 1.0]:
            cold, warm = orc.nnls(A, b, mu=mu), orc.nnls(A, b, mu=mu, warm=True)
            np.testing.assert_allclose(warm.x, cold.x, rtol=1e-7, atol=1e-10)
            np.testing.assert_allclose(warm.rnorm, cold.rnorm, rtol=1e-10)


def test_warm_start_first_pivot_quirk(orc):
    # w[n] is forced to 0, and to 1.0 only when every other dual is <= 0 (lsqnonneg.jl:69-70)
    A = np.array([[1.0, 0.0], [0.0, 1.0], [0.0, 0.0]])
    b = np.array([-1.0, 2.0, 0.0])
    r = orc.nnls(A, b, warm=True)
    np.testing.assert_allclose(r.x, [0.0, 2.0])
    b = np.array([3.0, 2.0, 0.0])
    r = orc.nnls(A, b, warm=True)
    np.testing.assert_allclose(r.x, [3.0, 2.0])


def test_zero_rhs_and_zero_column(orc):
    A = np.random.default_rng(0).random((6, 4))
    r = orc.nnls(A, np.zeros(6))
    assert np.all(r.x == 0) and r.nsetp == 0 and r.rnorm == 0
    A[:, 2] = 0
    r = orc.nnls(A, A @ np.array([1.0, 2.0, 3.0, 4.0]))
    assert r.x[2] == 0


def test_triangular_solves(orc):
    import ctypes as C
    rng = np.random.default_rng(3)
    L = orc.lib()
    for n in [1, 2, 5, 13]:
        U = np.asfortranarray(np.triu(rng.random((n, n))) + n * np.eye(n))
        z = rng.random(n)
        for transp in (0, 1):
            y = z.copy()
            L.orc_solve_triangular(y.ctypes.data_as(orc.dp), U.ctypes.data_as(orc.dp), n, n, transp)
            ref = np.linalg.solve(U.T if transp else U, z)
            np.testing.assert_allclose(y, ref, rtol=1e-12)


def test_hypot(orc):
    L = orc.lib()
    rng = np.random.default_rng(1)
    for a, b in rng.standard_normal((200, 2)) * 10.0 ** rng.integers(-8, 8, (200, 1)):
        assert L.orc_hypot(a, b) == pytest.approx(np.hypot(a, b), rel=4e-16)
    assert L.orc_hypot(3.0, 4.0) == 5.0
    assert L.orc_hypot(0.0, 0.0) == 0.0
