"""Parity metric between two runs of the pipeline (north_star tolerances).

  - T2 distributions: relative 1e-6, absolute 1e-9 for near-zero bins
  - flip angle, MWF (sfr), gmT2 (ggm): absolute 1e-6
  - voxels whose NNLS active set (support of the distribution) differs are counted and reported
"""
import numpy as np

DIST_RTOL, DIST_ATOL = 1e-6, 1e-9
SCALAR_ATOL = 1e-6


def compare(ref, got, scalars=("alpha", "sfr", "ggm"), extra=("gdn", "gva", "fnr", "snr", "sgm", "mfr", "mgm", "mu",
                                                              "chi2factor", "resnorm")):
    """ref/got: dicts with 'dist' (nvox, nT2) and map arrays (nvox,).  Returns a report dict."""
    d0, d1 = np.asarray(ref["dist"]), np.asarray(got["dist"])
    nvox = d0.shape[0]
    nan0, nan1 = np.isnan(d0).any(1), np.isnan(d1).any(1)
    rep = {"nvox": int(nvox), "nan_mismatch": int((nan0 != nan1).sum())}
    valid = ~nan0 & ~nan1
    scale = np.maximum(np.abs(d0), np.abs(d1))
    ok_bins = np.abs(d0 - d1) <= np.maximum(DIST_ATOL * np.maximum(1.0, scale.max(1, keepdims=True)), DIST_RTOL * scale)
    dist_ok = ok_bins.all(1) | ~valid
    support_diff = ((d0 > 0) != (d1 > 0)).any(1) & valid
    rep["dist_fail"] = int((~dist_ok).sum())
    rep["support_diff"] = int(support_diff.sum())
    rep["dist_fail_same_support"] = int((~dist_ok & ~support_diff).sum())
    with np.errstate(invalid="ignore", divide="ignore"):
        rel = np.abs(d0 - d1) / np.maximum(scale, 1e-300)
    rep["dist_max_rel_same_support"] = float(np.nanmax(np.where((~support_diff & valid)[:, None] & (scale > 0), rel, 0.0))) if nvox else 0.0
    fails = ~dist_ok
    for k in scalars:
        if k in ref and k in got:
            a, b = np.asarray(ref[k]), np.asarray(got[k])
            bad = ~((np.abs(a - b) <= SCALAR_ATOL) | (np.isnan(a) & np.isnan(b)))
            rep[k + "_fail"] = int(bad.sum())
            rep[k + "_max_abs"] = float(np.nanmax(np.abs(a - b))) if nvox else 0.0
            fails |= bad
    for k in extra:
        if k in ref and k in got:
            a, b = np.asarray(ref[k]), np.asarray(got[k])
            with np.errstate(invalid="ignore", divide="ignore"):
                r = np.abs(a - b) / np.maximum(np.abs(a), 1e-300)
            r = np.where(np.isnan(a) & np.isnan(b), 0.0, r)
            r = np.where(np.isinf(a) & (a == b), 0.0, r)
            rep[k + "_median_rel"] = float(np.nanmedian(r)) if nvox else 0.0
    rep["voxels_out_of_tolerance"] = int(fails.sum())
    rep["frac_out_of_tolerance"] = float(fails.mean()) if nvox else 0.0
    # Regularisation-parameter "path flips": the L-curve / GCV / Brent searches take discrete decisions on
    # quantities that are noisy at the 1e-13 level, so two faithful implementations (or the oracle
    # itself under a 1-ulp input perturbation, tests/test_oracle_sensitivity.py) pick a slightly
    # different mu for a few percent of voxels.  Those voxels are counted and reported separately.
    if "mu" in ref and "mu" in got:
        m0, m1 = np.asarray(ref["mu"], dtype=float), np.asarray(got["mu"], dtype=float)
        with np.errstate(invalid="ignore", divide="ignore"):
            dlog = np.abs(np.log(m0) - np.log(m1))
        same = (m0 == m1) | (np.isnan(m0) & np.isnan(m1)) | (dlog <= 1e-7)
        flip = ~same
        rep["mu_flips"] = int(flip.sum())
        rep["mu_flip_frac"] = float(flip.mean()) if nvox else 0.0
        rep["mu_flip_median_dlog"] = float(np.nanmedian(dlog[flip])) if flip.any() else 0.0
        rep["mu_flip_max_dlog"] = float(np.nanmax(dlog[flip])) if flip.any() else 0.0
        rep["support_diff_same_mu"] = int((support_diff & ~flip).sum())  # (a voxel at another mu may well have another support)
        rep["out_of_tolerance_same_mu"] = int((fails & ~flip).sum())
        rep["frac_out_of_tolerance_same_mu"] = float((fails & ~flip).mean()) if nvox else 0.0
    else:
        rep["mu_flips"], rep["mu_flip_frac"] = 0, 0.0
        rep["support_diff_same_mu"] = rep["support_diff"]
        rep["out_of_tolerance_same_mu"] = rep["voxels_out_of_tolerance"]
        rep["frac_out_of_tolerance_same_mu"] = rep["frac_out_of_tolerance"]
    return rep


def flip_bound(own, n):
    """GPU mu flips allowed on n voxels when two CPU builds of the oracle flip a fraction `own` of the same voxels: the
    same rate + 15 % + three sigmas of binomial sampling noise (the flips of two faithful implementations are
    independent draws from the same chaotic process, tests/test_oracle_sensitivity.py)."""
    import math
    return 1.15 * own + 3.0 * math.sqrt(max(own, 1.0 / n) * (1 - own) / n)
