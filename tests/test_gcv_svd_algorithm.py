"""CPU restatement (numpy) of the singular-value algorithm the GPU kernel runs for Reg = gcv (csrc/voxel.cuh:
gcv_svdvals_bidiag): Golub-Kahan bidiagonalisation with unnormalised Householder reflectors, then trisection on the Sturm counts
of the Golub-Kahan tridiagonal form with two pivots per reciprocal.  Checked against LAPACK (numpy.linalg.svd, the routine family
behind the reference's svdvals!, src/utils.jl:103-134) on the EPG bases of the benchmark configurations: every singular value
within a few eps * sigma_max.  The GPU side is covered by the gcv parity tests (tests/test_gpu_parity.py)."""
import numpy as np
import pytest


def bidiag_unnormalised(B):
    B = B.copy()
    R, C = B.shape
    for k in range(C):
        x = B[k:, k]
        xn2, x0 = np.dot(x[1:], x[1:]), x[0]
        if xn2 != 0.0:  # left reflector: annihilate B[k+1:, k]
            nrm = np.sqrt(x0 * x0 + xn2)
            v0, g = x0 + np.copysign(nrm, x0), 1.0 / (nrm * (nrm + abs(x0)))
            B[k, k] = -np.copysign(nrm, x0)
            if k + 1 < C:
                w = g * (v0 * B[k, k + 1:] + B[k + 1:, k] @ B[k + 1:, k + 1:])
                B[k, k + 1:] -= v0 * w
                B[k + 1:, k + 1:] -= np.outer(B[k + 1:, k], w)
        if k < C - 2:  # right reflector: annihilate B[k, k+2:]
            x = B[k, k + 1:]
            xn2, x0 = np.dot(x[1:], x[1:]), x[0]
            if xn2 != 0.0:
                nrm = np.sqrt(x0 * x0 + xn2)
                u0, g = x0 + np.copysign(nrm, x0), 1.0 / (nrm * (nrm + abs(x0)))
                B[k, k + 1] = -np.copysign(nrm, x0)
                z = g * (u0 * B[k + 1:, k + 1] + B[k + 1:, k + 2:] @ B[k, k + 2:])
                B[k + 1:, k + 1] -= u0 * z
                B[k + 1:, k + 2:] -= np.outer(z, B[k, k + 2:])
    return np.array([B[k, k] for k in range(C)]), np.array([B[k, k + 1] for k in range(C - 1)])


def sturm_count(b2, x, pivmin):
    """Number of eigenvalues of the Golub-Kahan tridiagonal form below x (vectorised over x), two pivots per reciprocal."""
    nb = len(b2)
    q = np.where(x < pivmin, -pivmin, -x)
    cnt, i = np.ones(len(x), dtype=int), 0
    while i + 1 < nb:
        n = -x * q - b2[i]
        n = np.where(np.abs(n) < pivmin * np.abs(q), -pivmin * q, n)
        cnt += (n < 0) != (q < 0)
        q = -(b2[i + 1] * q) * (1.0 / n) - x
        q = np.where(np.abs(q) < pivmin, -pivmin, q)
        cnt += q < 0
        i += 2
    cnt += (-b2[nb - 1] * (1.0 / q) - x) < 0
    return cnt


def svdvals_multisection(d, e):
    C = len(d)
    b2 = np.zeros(2 * C - 1)
    b2[0::2], b2[1::2] = d * d, e * e
    bound, pivmin = np.sqrt(b2.sum()) * 1.0000001, 1e-150 * max(1.0, b2.max())
    lo, hi, target = np.zeros(C), np.full(C, bound), C + np.arange(C)
    for _ in range(34):  # trisection: counts at the two interior points
        third = (hi - lo) * (1.0 / 3.0)
        xa, xb = lo + third, hi - third
        ca, cb = sturm_count(b2, xa, pivmin), sturm_count(b2, xb, pivmin)
        lo, hi = np.where(ca > target, lo, np.where(cb > target, xa, xb)), np.where(ca > target, xa, np.where(cb > target, xb, hi))
    return 0.5 * (lo + hi)


@pytest.mark.parametrize("nTE,nT2,TE", [(48, 60, 8e-3), (32, 60, 10e-3), (56, 40, 7e-3), (32, 40, 10e-3), (64, 64, 5e-3), (8, 5, 1e-2)])
def test_bidiagonalisation_and_bisection_match_lapack(orc, nTE, nT2, TE):
    T2 = orc.logrange(10e-3, 2.0, nT2)
    for alpha in (50.0, 142.1, 180.0):
        A = np.stack([orc.epg(nTE, alpha, TE, t2, 1.0) for t2 in T2], axis=1)
        B = A if nTE >= nT2 else A.T.copy()
        s = np.sort(svdvals_multisection(*bidiag_unnormalised(B)))
        ref = np.sort(np.linalg.svd(A, compute_uv=False))
        assert np.max(np.abs(s - ref)) <= 2e-15 * ref[-1], (nTE, nT2, alpha)
