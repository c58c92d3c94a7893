"""Oracle end-to-end: self-consistency of T2mapSEcorr/T2partSEcorr outputs on a small mock
volume for every regulariser and option permutation (modelled on test/cli.jl:255-361)."""
import numpy as np
import pytest

NTE, NT2, TE = 32, 40, 10e-3


@pytest.fixture(scope="module")
def image(orc):
    return orc.mock_image(48, NTE, TE, seed=1)


def run(orc, image, Reg, **kw):
    part_kw = kw.pop("part_kw", {})
    alloc = {k: kw.pop(k) for k in list(kw) if k in ("save_curve", "save_basis", "alpha_init")}
    o = orc.make_t2map_opts((image.shape[0], 1, 1), image.shape[1], kw.pop("nT2", NT2), TE, Reg=Reg, **kw)
    p = orc.make_t2part_opts((image.shape[0], 1, 1), o.nT2, **part_kw)
    return orc.t2map(image, o, p, **alloc), o, p


@pytest.mark.parametrize("Reg,extra", [("none", {}), ("lcurve", {}), ("gcv", {}), ("chi2", {"Chi2Factor": 1.02}),
                                       ("mdp", {"NoiseLevel": 1e-3})])
def test_outputs_are_consistent(orc, image, Reg, extra):
    (m, st), o, p = run(orc, image, Reg, save_curve=True, save_basis=True, **extra)
    nvox = image.shape[0]
    assert st.voxels_processed == nvox
    t2 = orc.logrange(10e-3, 2.0, NT2)
    dist = m["dist"]
    assert np.all(np.isfinite(dist)) and np.all(dist >= 0)
    np.testing.assert_allclose(m["gdn"], dist.sum(1), rtol=1e-12)
    np.testing.assert_allclose(m["ggm"], np.exp((dist * np.log(t2)).sum(1) / dist.sum(1)), rtol=1e-10)
    assert np.all((m["alpha"] >= 50) & (m["alpha"] <= 180))
    # fit = decaybasis * dist; resnorm, fnr, snr definitions (src/T2mapSEcorr.jl:527-551)
    basis = m["decaybasis"].reshape(NT2, NTE, nvox).transpose(2, 1, 0)
    fit = np.einsum("vij,vj->vi", basis, dist)
    np.testing.assert_allclose(m["decaycurve"], fit, rtol=1e-10, atol=1e-14)
    res = fit - image
    np.testing.assert_allclose(m["resnorm"], np.linalg.norm(res, axis=1), rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(m["fnr"], m["gdn"] / np.sqrt((res ** 2).sum(1) / (NTE - 1)), rtol=1e-6)
    np.testing.assert_allclose(m["snr"], image.max(1) / res.std(1, ddof=1), rtol=1e-6)
    # the saved basis is the EPG basis at the fitted angle
    for v in (0, 17):
        A = np.column_stack([orc.epg(NTE, m["alpha"][v], TE, t2[j], 1.0) for j in range(NT2)])
        np.testing.assert_array_equal(basis[v], A)
    if Reg == "none":
        assert np.all(m["mu"] == 0) and np.all(m["chi2factor"] == 1)
    else:
        assert np.all(m["mu"][~np.isinf(m["mu"])] >= 0)
        ok = m["mu"] > 0
        assert np.all(m["chi2factor"][ok] >= 1 - 1e-9)
    if Reg == "chi2":
        ok = m["mu"] > 0
        np.testing.assert_allclose(m["chi2factor"][ok], 1.02, rtol=1e-3)
    # fitted angles recover the simulated ones reasonably (true alpha ~ U(120,180), SNR 60 dB)
    # fused T2part == standalone T2part on the same distributions
    parts = orc.t2part(dist, p)
    for k in orc.PART_NAMES:
        np.testing.assert_array_equal(parts[k], m[k])
    sp = (t2 >= 10e-3) & (t2 <= 25e-3)
    np.testing.assert_allclose(m["sfr"], dist[:, sp].sum(1) / dist.sum(1), rtol=1e-12)
    assert 0.02 < np.median(m["sfr"]) < 0.35  # simulated sfr ~ U(0.05, 0.25)


def test_flip_angle_recovery(orc):
    img = orc.mock_image(64, 48, 8e-3, SNR=80.0, seed=3)
    o = orc.make_t2map_opts((64, 1, 1), 48, 40, 8e-3, Reg="none")
    m, st = orc.t2map(img, o)
    assert np.all((m["alpha"] >= 110) & (m["alpha"] <= 180))
    # spread of true angles is U(120,180): the fit must not collapse to a single value
    assert m["alpha"].std() > 5.0


def test_threshold_skips_voxels_and_keeps_nan(orc, image):
    img = image.copy()
    img[::3, 0] = 0.0
    o = orc.make_t2map_opts((img.shape[0], 1, 1), NTE, NT2, TE, Reg="none", Threshold=0.0)
    p = orc.make_t2part_opts((img.shape[0], 1, 1), NT2)
    m, st = orc.t2map(img, o, p)
    skipped = np.zeros(img.shape[0], bool)
    skipped[::3] = True
    assert st.voxels_processed == (~skipped).sum()
    for k in orc.MAP_NAMES + orc.PART_NAMES:
        assert np.all(np.isnan(m[k][skipped])) and np.all(np.isfinite(m[k][~skipped]))
    assert np.all(np.isnan(m["dist"][skipped]))
    # Threshold = +huge: nothing processed, everything NaN (src/T2mapSEcorr.jl:178-181)
    o = orc.make_t2map_opts((img.shape[0], 1, 1), NTE, NT2, TE, Reg="none", Threshold=1e30)
    m, st = orc.t2map(img, o)
    assert st.voxels_processed == 0 and np.all(np.isnan(m["gdn"]))
    # Threshold = -Inf processes every voxel
    o = orc.make_t2map_opts((img.shape[0], 1, 1), NTE, NT2, TE, Reg="none", Threshold=-np.inf)
    m, st = orc.t2map(img, o)
    assert st.voxels_processed == img.shape[0]


def test_set_flip_angle_and_b1_map(orc, image):
    (m, _), o, p = run(orc, image, "none", SetFlipAngle=170.0, save_basis=True)
    assert np.all(m["alpha"] == 170.0)
    assert np.all(np.isnan(m["decaybasis"]))  # single shared basis is not written per voxel (:581-587)
    b1 = np.linspace(130.0, 179.0, image.shape[0])
    (m2, _), o, p = run(orc, image, "none", alpha_provided=True, alpha_init=b1)
    np.testing.assert_array_equal(m2["alpha"], b1)
    # a B1 map equal to the fixed angle reproduces the fixed-angle run
    (m3, _), o, p = run(orc, image, "none", alpha_provided=True, alpha_init=np.full(image.shape[0], 170.0))
    np.testing.assert_allclose(m3["dist"], m["dist"], rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("nTE,nT2", [(4, 2), (5, 3), (8, 8), (47, 47)])
def test_odd_sizes(orc, nTE, nT2):
    # test/cli.jl:264-267 exercises nTE in {4,5,8,47}, nT2 in {2,3,8,47}
    img = orc.mock_image(8, nTE, 10e-3, seed=nTE)
    for Reg, extra in [("none", {}), ("lcurve", {}), ("gcv", {}), ("chi2", {"Chi2Factor": 1.05}),
                       ("mdp", {"NoiseLevel": 1e-2})]:
        o = orc.make_t2map_opts((8, 1, 1), nTE, nT2, 10e-3, Reg=Reg, **extra)
        m, st = orc.t2map(img, o)
        assert np.all(np.isfinite(m["dist"])) and np.all(m["dist"] >= 0)


def test_refcon_angle_path(orc, image):
    (m, _), o, p = run(orc, image, "none", RefConAngle=150.0)
    assert np.all(np.isfinite(m["gdn"])) and np.all((m["alpha"] >= 50) & (m["alpha"] <= 180))


def test_sigmoid_weights(orc, image):
    from scipy.special import erfc, erfinv
    (m, _), o, p = run(orc, image, "none", part_kw={"Sigmoid": 5e-3})
    t2 = orc.logrange(10e-3, 2.0, NT2)
    sigma = abs(5e-3 / (np.sqrt(2) * erfinv(2 * 0.1 - 1)))
    w = erfc(((t2 - 25e-3) / sigma) / np.sqrt(2)) / 2
    w[w <= np.finfo(float).eps] = 0
    np.testing.assert_allclose(m["sfr"], (m["dist"] * w).sum(1) / m["dist"].sum(1), rtol=1e-10)


def test_option_validation(orc):
    import ctypes as C
    L = orc.lib()
    msg = C.create_string_buffer(256)
    good = orc.make_t2map_opts((2, 2, 2), 32, 40, 10e-3, Reg="chi2", Chi2Factor=1.02)
    assert L.orc_validate_t2map_opts(C.byref(good), msg, 256) == 0
    bad = [dict(nTE=3), dict(nT2=1), dict(TE=-1.0), dict(T2Range=(2.0, 1.0)), dict(T1=0.0), dict(Threshold=-1.0),
           dict(MinRefAngle=200.0), dict(nRefAngles=1), dict(nRefAnglesMin=1), dict(Reg="chi2", Chi2Factor=1.0),
           dict(Reg="chi2"), dict(Reg="mdp"), dict(Reg="mdp", NoiseLevel=0.0), dict(RefConAngle=181.0),
           dict(SetFlipAngle=-5.0)]
    base = dict(shape=(2, 2, 2), nTE=32, nT2=40, TE=10e-3)
    for kw in bad:
        args = {**base, **kw}
        o = orc.make_t2map_opts(args.pop("shape"), args.pop("nTE"), args.pop("nT2"), args.pop("TE"), **args)
        assert L.orc_validate_t2map_opts(C.byref(o), msg, 256) == -1, kw
    o = orc.make_t2map_opts((2, 2, 2), 32, 40, 10e-3, legacy=True, nRefAngles=8, nRefAnglesMin=8)
    assert L.orc_validate_t2map_opts(C.byref(o), msg, 256) == 0


def test_thread_count_does_not_change_results(orc, image):
    o = orc.make_t2map_opts((image.shape[0], 1, 1), NTE, NT2, TE, Reg="lcurve")
    m1, _ = orc.t2map(image, o, nthreads=1)
    m4, _ = orc.t2map(image, o, nthreads=4)
    for k in m1:
        np.testing.assert_array_equal(m1[k], m4[k])


def test_flop_counter_is_in_expected_range(orc, image):
    # SURVEY 8(d) pre-measurement estimate for the 32-echo / none config: ~0.9 MFLOP per voxel
    o = orc.make_t2map_opts((image.shape[0], 1, 1), NTE, NT2, TE, Reg="none")
    m, st = orc.t2map(image, o)
    per_voxel = st.flops / st.voxels_processed
    assert 0.2e6 < per_voxel < 3e6
    assert st.nnls_unreg >= 6 * st.voxels_processed
