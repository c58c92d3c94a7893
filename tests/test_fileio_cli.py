"""Host side next to the hot path: file formats (src/main.jl:577-633) and the command line (src/main.jl:1-575).
CPU only: nothing here launches a kernel."""
import gzip
import importlib
import os
import struct

import numpy as np
import pytest


@pytest.fixture(scope="module")
def fio(pkg):
    return importlib.import_module("decaes_jl_b200.fileio")


@pytest.fixture(scope="module")
def cli(pkg):
    return importlib.import_module("decaes_jl_b200.cli")


@pytest.mark.parametrize("dtype", ["u1", "i2", "i4", "f4", "f8", "u2"])
@pytest.mark.parametrize("gz", [False, True])
def test_nifti_round_trip(fio, tmp_path, dtype, gz):
    rng = np.random.default_rng(1)
    a = (rng.random((5, 4, 3, 6)) * 100).astype(dtype)
    f = str(tmp_path / ("x.nii.gz" if gz else "x.nii"))
    fio.save_nifti(f, a)
    raw, slope, inter = fio.read_nifti(f)
    assert raw.dtype.kind == np.dtype(dtype).kind and raw.shape == a.shape
    np.testing.assert_array_equal(raw, a)
    img = fio.load_image(f)
    assert img.dtype == np.float64 and img.flags["F_CONTIGUOUS"] and img.flags["WRITEABLE"] and img.flags["OWNDATA"]
    np.testing.assert_array_equal(img, a.astype(np.float64))
    # voxel index fastest: element (1, 0, 0, 0) follows (0, 0, 0, 0) in memory
    assert img.ravel(order="K")[1] == float(a[1, 0, 0, 0])


def test_nifti_scaling_rule_and_big_endian(fio, tmp_path):
    a = np.arange(24, dtype=np.int16).reshape(2, 3, 4)
    f = str(tmp_path / "s.nii")
    fio.save_nifti(f, a, scl_slope=0.5, scl_inter=10.0)
    np.testing.assert_allclose(fio.load_image(f, 3), a * 0.5 + 10.0)
    # the reference scales in promote_type(eltype(raw), Float32) (src/main.jl:599): Float32 arithmetic for integer volumes
    fio.save_nifti(f, a, scl_slope=0.1, scl_inter=0.3)
    want = (a.astype(np.float32) * np.float32(0.1) + np.float32(0.3)).astype(np.float64)
    np.testing.assert_array_equal(fio.load_image(f, 3), want)
    assert not np.array_equal(want, a * 0.1 + 0.3)  # ... which is not what Float64 arithmetic gives
    fio.save_nifti(f, a, scl_slope=0.0, scl_inter=7.0)  # slope 0: "data is not scaled", raw data returned
    np.testing.assert_array_equal(fio.load_image(f, 3), a.astype(float))
    # hand-made big-endian file
    hdr = bytearray(348)
    struct.pack_into(">i", hdr, 0, 348)
    struct.pack_into(">8h", hdr, 40, 3, 2, 3, 4, 1, 1, 1, 1)
    struct.pack_into(">hh", hdr, 70, 4, 16)
    struct.pack_into(">3f", hdr, 108, 352.0, 1.0, 0.0)
    hdr[344:348] = b"n+1\0"
    with open(f, "wb") as fh:
        fh.write(bytes(hdr) + b"\0" * 4 + a.astype(">i2").tobytes(order="F"))
    np.testing.assert_array_equal(fio.load_image(f, 3), a.astype(float))
    with open(f, "wb") as fh:
        fh.write(b"not a nifti file")
    with pytest.raises(ValueError):
        fio.load_image(f, 3)


def test_ensure_ndims_and_mat_selection(fio, tmp_path):
    from scipy.io import savemat
    a3 = np.arange(24.0).reshape(2, 3, 4)
    assert fio.ensure_ndims("f", a3, 4).shape == (2, 3, 4, 1)
    a5 = np.arange(48.0).reshape(2, 3, 4, 2, 1)
    np.testing.assert_array_equal(fio.ensure_ndims("f", a5, 4), a5[..., 0])
    a6 = np.arange(96.0).reshape(2, 3, 4, 2, 2, 1)
    with pytest.warns(UserWarning, match="selecting the first 4-D volume"):
        np.testing.assert_array_equal(fio.ensure_ndims("f", a6, 4), a6[:, :, :, :, 0, 0])
    f = str(tmp_path / "m.mat")
    b = np.random.default_rng(0).random((2, 3, 4, 5))
    savemat(f, {"zeta": b + 1, "alpha": b, "vec": np.arange(3.0)})
    with pytest.warns(UserWarning, match="Choosing variable 'alpha'"):
        np.testing.assert_array_equal(fio.load_image(f), b)
    savemat(f, {"vec": np.arange(3.0)})
    with pytest.raises(ValueError, match="No 4-D array was found"):
        fio.load_image(f)
    g = str(tmp_path / "out" / "r.t2maps.mat")
    fio.save_mat(g, {"gdn": b[..., 0], "t2times": np.arange(5.0), "refangleset": 170.0})
    from scipy.io import loadmat
    back = loadmat(g)
    np.testing.assert_array_equal(back["gdn"], b[..., 0])
    assert back["t2times"].shape == (5, 1) and float(back["refangleset"].squeeze()) == 170.0


def test_suffix_helpers(fio):
    assert fio.maybe_get_suffix("A/B/img.NII.GZ") == ".nii.gz" and fio.maybe_get_suffix("x.nii") == ".nii"
    assert fio.chop_allowed_suffix("brain.nii.gz") == "brain" and fio.chop_allowed_suffix("d.t2dist.mat") == "d.t2dist"
    assert not fio.is_allowed_suffix("settings.txt")
    with pytest.raises(ValueError):
        fio.chop_allowed_suffix("x.txt")


def test_cli_parsing_settings_file_and_reg_params(cli, tmp_path):
    s = tmp_path / "settings.txt"
    s.write_text("\n".join(["a.nii", "b.nii.gz", "--T2map", "--T2part", "--TE", "7e-3", "--nT2", "40", "--T2Range", "10e-3", "2.0",
                            "--SPWin", "10e-3", "25e-3", "--MPWin", "25e-3", "200e-3", "--Reg", "chi2", "--RegParams", "1.02",
                            "--output", str(tmp_path / "o")]))
    o = cli.parse_cli(["@" + str(s), "--SaveRegParam"])
    assert o["input"] == ["a.nii", "b.nii.gz"] and o["T2map"] and o["T2part"] and o["Chi2Factor"] == 1.02 and o["SaveRegParam"]
    infos = cli.get_file_infos(o)
    assert [i["choppedinputfile"] for i in infos] == ["a", "b"] and all(i["outputfolder"] == str(tmp_path / "o") for i in infos)
    o = cli.parse_cli(["x.mat", "--T2map", "--Reg", "mdp", "--RegParams", "1e-3"])
    assert o["NoiseLevel"] == 1e-3 and "Chi2Factor" not in o
    with pytest.warns(UserWarning, match="deprecated"):
        o = cli.parse_cli(["x.mat", "--T2map", "--Reg", "chi2", "--Chi2Factor", "1.05"])
    assert o["Chi2Factor"] == 1.05
    with pytest.raises(SystemExit, match="both passed"):
        cli.parse_cli(["x.mat", "--T2map", "--Reg", "chi2", "--Chi2Factor", "1.05", "--RegParams", "1.02"])
    with pytest.raises(SystemExit, match="At least one of --T2map or --T2part"):
        cli.parse_cli(["x.mat"])
    with pytest.raises(AssertionError, match="Must set chi2 factor"):
        cli.parse_cli(["x.mat", "--T2map", "--Reg", "chi2"])
    with pytest.warns(UserWarning, match="--legacy is deprecated"):  # warn_deprecated_future_removed  src/main.jl:441-443
        o = cli.parse_cli(["x.mat", "--T2map", "--legacy"])
    assert o["legacy"] is True


def test_cli_file_infos_rules(cli, tmp_path):
    base = dict(mask=[], B1map=[], SetFlipAngle=None)
    infos = cli.get_file_infos(dict(base, input=["d/a.nii", "b.mat", "notes.txt"], output=[]))
    assert [(i["inputfile"], i["outputfolder"]) for i in infos] == [("d/a.nii", "d"), ("b.mat", ".")]
    infos = cli.get_file_infos(dict(base, input=["a.nii", "b.nii"], output=["o1", "o2"], mask=["m1.nii", "m2.nii"]))
    assert [(i["outputfolder"], i["maskfile"]) for i in infos] == [("o1", "m1.nii"), ("o2", "m2.nii")]
    with pytest.raises(SystemExit, match="Incorrect number of output files"):
        cli.get_file_infos(dict(base, input=["a.nii", "b.nii", "c.nii"], output=["o1", "o2"]))
    with pytest.raises(SystemExit, match="Number of mask files"):
        cli.get_file_infos(dict(base, input=["a.nii", "b.nii"], output=[], mask=["m.nii"]))
    with pytest.raises(AssertionError, match="Cannot set a fixed flip angle"):
        cli.get_file_infos(dict(base, input=["a.nii"], output=[], B1map=["b1.nii"], SetFlipAngle=170.0))
    f = tmp_path / "settings.txt"
    f.write_text("--T2map")
    with pytest.raises(SystemExit, match="prepend an '@' character"):
        cli.get_file_infos(dict(base, input=[str(f)], output=[]))
    with pytest.raises(SystemExit, match="No valid files were found"):
        cli.get_file_infos(dict(base, input=["nothing.txt"], output=[]))


def test_cli_option_structs(cli):
    o = cli.parse_cli(["x.nii", "--T2map", "--T2part", "--TE", "8e-3", "--nT2", "60", "--T2Range", "0.01", "2", "--SPWin", "0.01", "0.025",
                       "--MPWin", "0.025", "0.2", "--Reg", "mdp", "--RegParams", "1e-3", "--Sigmoid", "1e-3", "--RefConAngle", "150",
                       "--Threshold", "5", "--nRefAngles", "32", "--SaveDecayCurve", "--ngpus", "2"])
    m = cli.t2map_options(np.zeros((3, 4, 5, 48)), o)
    assert (m.MatrixSize, m.nTE, m.TE, m.nT2, m.T2Range, m.Reg, m.NoiseLevel) == ((3, 4, 5), 48, 8e-3, 60, (0.01, 2.0), "mdp", 1e-3)
    assert m.RefConAngle == 150 and m.Threshold == 5 and m.nRefAngles == 32 and m.SaveDecayCurve and not m.SaveRegParam and m.ngpus == 2
    p = cli.t2part_options(np.zeros((3, 4, 5, 60)), o)
    assert (p.nT2, p.SPWin, p.MPWin, p.Sigmoid) == (60, (0.01, 0.025), (0.025, 0.2), 1e-3)
    # --T2part alone: nT2 comes from the distribution, a passed --nT2 is ignored
    o2 = cli.parse_cli(["d.t2dist.mat", "--T2part", "--nT2", "7", "--T2Range", "0.01", "2", "--SPWin", "0.01", "0.025", "--MPWin", "0.025", "0.2"])
    assert cli.t2part_options(np.zeros((3, 4, 5, 40)), o2).nT2 == 40
    with pytest.raises(AssertionError, match="Echo spacing"):
        cli.t2map_options(np.zeros((3, 4, 5, 48)), dict(o, TE=-1.0))
