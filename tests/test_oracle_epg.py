"""Oracle EPG: known answers from the reference docstring, the analytic alpha=180 curve, an
independent numpy restatement of the readable spec EPGWork_Basic_Cplx
(src/EPGdecaycurve.jl:254-322), and Jacobian-vs-finite-difference checks (test/epg.jl)."""
import numpy as np
import pytest

EPS = np.finfo(float).eps


def epg_spec_numpy(ETL, alpha, TE, T2, T1, beta=180.0):
    """EPGWork_Basic_Cplx: complex 3-vector phase states, flip matrix of element_flipmat
    (src/EPGdecaycurve.jl:6-9), relax TE/2 - flip - shift - relax TE/2."""
    sind = lambda d: np.sin(np.deg2rad(d))
    cosd = lambda d: np.cos(np.deg2rad(d))

    def flipmat(a):
        return np.array([[cosd(a / 2) ** 2, sind(a / 2) ** 2, -1j * sind(a)],
                         [sind(a / 2) ** 2, cosd(a / 2) ** 2, 1j * sind(a)],
                         [-0.5j * sind(a), 0.5j * sind(a), cosd(a)]])

    A = alpha / 180
    a_ex, a1, ai = A * 90, A * 180, A * beta
    E1, E2 = np.exp(-(TE / 2) / T1), np.exp(-(TE / 2) / T2)
    E = np.array([E2, E2, E1])
    R1, Ri = flipmat(a1), flipmat(ai)
    M = np.zeros((ETL, 3), dtype=complex)
    M[0, 0] = sind(a_ex)
    dc = np.zeros(ETL)
    for i in range(ETL):
        R = R1 if i == 0 else Ri
        M = (R @ (E[None, :] * M).T).T
        Mn = np.zeros_like(M)
        Mn[0] = [M[0, 1], M[1, 1], M[0, 2]]
        for j in range(1, ETL - 1):
            Mn[j] = [M[j - 1, 0], M[j + 1, 1], M[j, 2]]
        Mn[ETL - 1] = [M[ETL - 2, 0], 0, M[ETL - 1, 2]]
        M = E[None, :] * Mn
        dc[i] = abs(M[0, 0])
    return dc


def test_docstring_known_answers(orc):
    # src/T2mapSEcorr.jl:126 and :134 (TE = 10 ms, nT2 = 40, T2Range = (10 ms, 2 s), first angle 50 deg)
    t2 = orc.logrange(10e-3, 2.0, 40)
    np.testing.assert_allclose(t2[:5], [0.01, 0.0114551, 0.013122, 0.0150315, 0.0172188], rtol=5e-6)
    assert t2[0] == 10e-3 and t2[-1] == 2.0
    kat = {0: 0.0277684, 1: 0.0315296, 38: 0.0750511, 39: 0.0751058}
    for j, val in kat.items():
        dc = orc.epg(48, 50.0, 10e-3, t2[j], 1.0)
        assert abs(dc[0] - val) < 5e-8
    # refangleset printed in the docstring (32-angle grid of that release): 50.0, 54.1935, 58.3871, ...
    np.testing.assert_allclose(orc.linrange(50.0, 180.0, 32)[:4], [50.0, 54.1935, 58.3871, 62.5806], atol=5e-5)


def test_alpha_180_is_pure_exponential(orc):
    for ETL, TE, T2 in [(32, 10e-3, 0.05), (48, 8e-3, 0.0123), (56, 7e-3, 1.7)]:
        dc = orc.epg(ETL, 180.0, TE, T2, 1.0)
        ref = np.exp(-np.arange(1, ETL + 1) * TE / T2)
        np.testing.assert_allclose(dc, ref, rtol=64 * EPS)


@pytest.mark.parametrize("ETL", list(range(4, 65)))
def test_fast_kernel_matches_spec(orc, ETL):
    # test/epg.jl:77-82 — default workspace vs EPGWork_Basic_Cplx, rtol sqrt(eps), atol 10 eps
    al, TE, T2, T1 = 165.0, 39e-3, 1.1, 151.0
    ref = epg_spec_numpy(ETL, al, TE, T2, T1)
    got = orc.epg(ETL, al, TE, T2, T1)
    np.testing.assert_allclose(got, ref, rtol=np.sqrt(EPS), atol=10 * EPS)


@pytest.mark.parametrize("ETL", [4, 5, 6, 7, 16, 32, 47, 48, 56, 64])
@pytest.mark.parametrize("beta", [150.0, 180.0, 90.0])
def test_general_beta_kernel_matches_spec(orc, ETL, beta):
    al, TE, T2, T1 = 165.0, 39e-3, 1.1, 151.0
    ref = epg_spec_numpy(ETL, al, TE, T2, T1, beta)
    got = orc.epg(ETL, al, TE, T2, T1, beta=beta)
    np.testing.assert_allclose(got, ref, rtol=np.sqrt(EPS), atol=10 * EPS)
    if beta == 180.0:  # constant-flip kernel vs EPGOptions(beta = 180)  test/epg.jl:86-121
        np.testing.assert_allclose(orc.epg(ETL, al, TE, T2, T1), got, rtol=np.sqrt(EPS), atol=10 * EPS)


def test_random_parameters_match_spec(orc):
    rng = np.random.default_rng(0)
    for _ in range(50):
        ETL = int(rng.integers(4, 65))
        al = rng.uniform(50, 180)
        TE = rng.uniform(5e-3, 15e-3)
        T2 = np.exp(rng.uniform(np.log(10e-3), np.log(2.0)))
        ref = epg_spec_numpy(ETL, al, TE, T2, 1.0)
        np.testing.assert_allclose(orc.epg(ETL, al, TE, T2, 1.0), ref, rtol=np.sqrt(EPS), atol=10 * EPS)


@pytest.mark.parametrize("ETL", [4, 5, 8, 32, 47, 56])
def test_jacobian_vs_finite_differences(orc, ETL):
    # test/epg.jl:140-173 — ForwardDiff Jacobian vs central differences at three step sizes
    TE, T2, T1 = 10e-3, 0.08, 1.0
    for al in [50.0, 77.3, 120.0, 165.0, 179.0]:
        dc, ddc = orc.epg_jac(ETL, al, TE, T2, T1)
        np.testing.assert_allclose(dc, orc.epg(ETL, al, TE, T2, T1), rtol=4 * EPS, atol=4 * EPS)
        best = np.inf
        for h in [1e-3, 1e-4, 1e-5]:
            fd = (orc.epg(ETL, al + h, TE, T2, T1) - orc.epg(ETL, al - h, TE, T2, T1)) / (2 * h)
            best = min(best, np.max(np.abs(fd - ddc)))
        assert best < 1e-9


def test_jacobian_at_180_is_zero_slope_sign(orc):
    # at alpha = 180 the curve is at its maximum in alpha: derivative ~ 0
    dc, ddc = orc.epg_jac(32, 180.0, 10e-3, 0.05, 1.0)
    assert np.max(np.abs(ddc)) < 1e-12


def test_sind_exact_cases(orc):
    L = orc.lib()
    assert L.orc_sind(90.0) == 1.0
    assert L.orc_sind(30.0) == pytest.approx(0.5, abs=1e-16)
    assert L.orc_sind(180.0) == 0.0
    assert L.orc_sind(45.0) == pytest.approx(np.sqrt(0.5), abs=1.2e-16)


@pytest.mark.parametrize("ETL", [4, 5, 8, 17, 32, 56, 64])
@pytest.mark.parametrize("beta", [90.0, 120.0, 150.0, 179.0])
def test_beta_jacobian_forward_mode(orc, ETL, beta):
    """Forward-mode Jacobian of the general-beta kernel: values are bitwise those of the double-buffered
    restatement of src/EPGdecaycurve.jl:722-818, derivatives match central differences
    (the reference's own check of its ForwardDiff pass, test/epg.jl:161-172)."""
    rng = np.random.default_rng(ETL)
    for _ in range(5):
        alpha, T2 = rng.uniform(50, 180), rng.uniform(0.01, 2.0)
        dc, ddc = orc.epg_beta_jac(ETL, alpha, 10e-3, T2, 1.0, beta)
        np.testing.assert_array_equal(dc, orc.epg(ETL, alpha, 10e-3, T2, 1.0, beta=beta))
        h = 1e-5
        fd = (orc.epg(ETL, alpha + h, 10e-3, T2, 1.0, beta=beta) - orc.epg(ETL, alpha - h, 10e-3, T2, 1.0, beta=beta)) / (2 * h)
        np.testing.assert_allclose(ddc, fd, rtol=2e-6, atol=1e-9)
