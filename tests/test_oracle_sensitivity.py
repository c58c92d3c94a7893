"""The reference's regularisation-parameter searches are chaotic at the 1e-4 level.

lcurve_corner (src/lsqnonneg.jl:872-919) stops when the bracket is 1e-4 wide and moves on
comparisons of Menger curvatures of nearly collinear point triples; a 1-ulp perturbation of the
input therefore changes the selected mu for a few percent of voxels (by ~xtol), and with it the
distribution at the 1e-4..1e-2 level.  This test measures that intrinsic sensitivity on the
oracle itself; tests/test_gpu_parity.py bounds the GPU-vs-oracle "mu flip" rate by the same
figure and requires north_star tolerances on all voxels that selected the same mu."""
import numpy as np

import parity


def test_lcurve_mu_selection_is_intrinsically_sensitive(orc):
    nvox, nTE, TE, nT2 = 768, 48, 8e-3, 40
    img = orc.mock_image(nvox, nTE, TE, seed=2)
    o = orc.make_t2map_opts((nvox, 1, 1), nTE, nT2, TE, Reg="lcurve")
    p = orc.make_t2part_opts((nvox, 1, 1), nT2)
    ref, _ = orc.t2map(img, o, p)
    rng = np.random.default_rng(0)
    img2 = np.asfortranarray(img * (1 + 1e-16 * rng.standard_normal(img.shape)))  # ~1 ulp on some samples
    assert 0 < np.mean(img2 != img) < 1
    got, _ = orc.t2map(img2, o, p)
    rep = parity.compare(ref, got)
    # same active sets everywhere, yet a few percent of voxels choose another mu ...
    assert rep["support_diff"] <= 2
    assert 0.005 <= rep["mu_flip_frac"] <= 0.08, rep
    # ... by roughly the search tolerance
    assert rep["mu_flip_median_dlog"] < 2e-3
    # and every voxel that kept its mu agrees to the north_star tolerance
    assert rep["out_of_tolerance_same_mu"] == 0, rep


def test_unregularised_path_is_stable_under_the_same_perturbation(orc):
    nvox, nTE, TE, nT2 = 768, 32, 10e-3, 40
    img = orc.mock_image(nvox, nTE, TE, seed=1)
    o = orc.make_t2map_opts((nvox, 1, 1), nTE, nT2, TE, Reg="none")
    ref, _ = orc.t2map(img, o)
    rng = np.random.default_rng(0)
    got, _ = orc.t2map(np.asfortranarray(img * (1 + 1e-16 * rng.standard_normal(img.shape))), o)
    rep = parity.compare(ref, got)
    assert rep["voxels_out_of_tolerance"] <= 1, rep
