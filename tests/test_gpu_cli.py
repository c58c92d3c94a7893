"""End to end through the command line (mirror of DECAES.main, src/main.jl:333-420) on a small synthetic volume:
NIfTI in, mask applied, T2map + fused T2part, MAT files out; then --T2part alone on the saved distribution; then a
B1-map run.  Results must equal the direct API calls on the same arrays."""
import importlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_cli_end_to_end(pkg, orc, tmp_path):
    from scipy.io import loadmat
    cli = importlib.import_module("decaes_jl_b200.cli")
    fio = importlib.import_module("decaes_jl_b200.fileio")
    nx, ny, nz, nTE, nT2, TE = 8, 6, 4, 32, 40, 10e-3
    nvox = nx * ny * nz
    img = orc.mock_image(nvox, nTE, TE, seed=31).reshape((nx, ny, nz, nTE), order="F")
    mask = np.ones((nx, ny, nz))
    mask[:2] = 0.0
    f_img, f_mask = str(tmp_path / "brain.nii.gz"), str(tmp_path / "mask.nii")
    fio.save_nifti(f_img, img)
    fio.save_nifti(f_mask, mask.astype(np.uint8))
    out = str(tmp_path / "results")
    common = ["--T2Range", "10e-3", "2.0", "--SPWin", "10e-3", "25e-3", "--MPWin", "25e-3", "200e-3"]
    res = cli.main([f_img, "--mask", f_mask, "--output", out, "--T2map", "--T2part", "--TE", str(TE), "--nT2", str(nT2), "--Reg", "lcurve",
                    "--SaveRegParam", "--quiet", "--ngpus", "1"] + common)
    assert len(res) == 1
    for suffix in (".t2dist.mat", ".t2maps.mat", ".t2parts.mat"):
        assert os.path.isfile(os.path.join(out, "brain" + suffix)), suffix
    maps_f, dist_f, parts_f = (loadmat(os.path.join(out, "brain" + s)) for s in (".t2maps.mat", ".t2dist.mat", ".t2parts.mat"))
    # direct API on the same masked array
    masked = np.asfortranarray(img * mask[..., None])
    o = pkg.T2mapOptions(MatrixSize=(nx, ny, nz), nTE=nTE, TE=TE, nT2=nT2, T2Range=(10e-3, 2.0), Reg="lcurve", SaveRegParam=True,
                         Silent=True, ngpus=1)
    p = pkg.T2partOptions(MatrixSize=(nx, ny, nz), nT2=nT2, T2Range=(10e-3, 2.0), SPWin=(10e-3, 25e-3), MPWin=(25e-3, 200e-3), Silent=True)
    maps, dist = pkg.T2mapSEcorr(masked, o, t2part=p)
    np.testing.assert_array_equal(dist_f["dist"], dist)
    for k in ("gdn", "ggm", "gva", "fnr", "snr", "alpha", "mu", "chi2factor"):
        np.testing.assert_array_equal(maps_f[k], maps[k], err_msg=k)
    for k in ("sfr", "sgm", "mfr", "mgm"):
        np.testing.assert_array_equal(parts_f[k], maps[k], err_msg=k)
    assert np.isnan(maps_f["gdn"][:2]).all() and np.isfinite(maps_f["gdn"][2:]).all()  # masked voxels are skipped -> NaN
    np.testing.assert_allclose(maps_f["t2times"].ravel()[[0, -1]], [10e-3, 2.0])
    assert maps_f["decaybasisset"].shape == (nTE, nT2, 64) and maps_f["refangleset"].size == 64

    # --T2part alone on the saved distribution (standalone kernel; summation order differs from the fused epilogue)
    res2 = cli.main([os.path.join(out, "brain.t2dist.mat"), "--T2part", "--quiet", "--dry"] + common)
    for k in ("sfr", "sgm", "mfr", "mgm"):
        np.testing.assert_allclose(res2[0]["t2parts"][k], maps[k], rtol=1e-12, equal_nan=True, err_msg=k)
    assert not os.path.exists(os.path.join(out, "brain.t2dist.t2parts.mat"))  # --dry saves nothing

    # B1 map: the fitted angle is the map itself
    b1 = np.linspace(140.0, 178.0, nvox).reshape((nx, ny, nz), order="F")
    f_b1 = str(tmp_path / "b1.nii")
    fio.save_nifti(f_b1, b1)
    res3 = cli.main([f_img, "--B1map", f_b1, "--T2map", "--TE", str(TE), "--nT2", str(nT2), "--Reg", "none", "--quiet", "--dry",
                     "--T2Range", "10e-3", "2.0"])
    np.testing.assert_array_equal(res3[0]["t2maps"]["alpha"], b1)
