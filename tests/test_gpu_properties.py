"""Size-independent properties of the GPU path at the headline size (config 3: 56 echoes, 240x240x113 =
6,508,800 voxels, nT2 = 40, Reg = lcurve + fused T2part), where the CPU oracle would need minutes:

  * slab invariance: the volume processed as one launch equals, bit for bit, the same volume processed
    as three slabs (how volumes are sharded over GPUs) — voxels are independent;
  * determinism: a second run reproduces every byte (dynamic work distribution must not leak into results);
  * the fused T2part epilogue equals (to 1e-12: different summation order) the standalone T2partSEcorr kernel
    applied to the stored distributions;
  * sanity of the maps: every voxel processed, finite outputs, sfr in [0, 1], distributions non-negative;
  * exact scale covariance on a sub-volume: image * 2^k gives dist * 2^k, gdn * 2^k and identical
    scale-free maps (the per-voxel normalisation divides by a power of two exactly), src/T2mapSEcorr.jl:205-218.

Device buffers come from torch (plumbing only); everything computed goes through the C ABI.
"""
import pytest

pytestmark = pytest.mark.gpu

NAMES = ["gdn", "ggm", "gva", "fnr", "snr", "alpha", "sfr", "sgm", "mfr", "mgm"]


def _alloc(torch, dev, nvox, nT2):
    outs = {k: torch.full((nvox,), float("nan"), dtype=torch.float64, device=dev) for k in NAMES}
    outs["dist"] = torch.full((nT2, nvox), float("nan"), dtype=torch.float64, device=dev)
    return outs


def _run(pkg, torch, img, nvox_total, v0, v1, outs, o_kw, nT2):
    """Process voxels [v0, v1) of the (nTE, nvox_total) device image into the matching range of `outs`."""
    n = v1 - v0
    o = pkg.T2mapOptions(MatrixSize=(n, 1, 1), nTE=img.shape[0], nT2=nT2, T2Range=(10e-3, 2.0), Silent=True, ngpus=1,
                         **o_kw).to_c()
    p = pkg.T2partOptions(MatrixSize=(n, 1, 1), nT2=nT2, T2Range=(10e-3, 2.0), SPWin=(10e-3, 25e-3),
                          MPWin=(25e-3, 200e-3), Silent=True).to_c()
    ptrs = {k: outs[k].data_ptr() + 8 * v0 for k in NAMES}
    ptrs["dist"] = outs["dist"].data_ptr() + 8 * v0
    out = pkg.make_out(ptrs)
    pkg.t2map_device(img.data_ptr() + 8 * v0, n, nvox_total, o, p, out, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return pkg.last_stats(), p


def _same(torch, a, b):
    return bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all().item())


def test_full_size_slab_invariance_determinism_and_fused_part(pkg):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    dev = torch.device("cuda", 0)
    nTE, TE, nT2 = 56, 7e-3, 40
    nvox = 240 * 240 * 113
    img = torch.empty((nTE, nvox), dtype=torch.float64, device=dev)
    pkg.mock_image_device(img.data_ptr(), nvox, nvox, 0, nTE, TE, seed=3, stream=torch.cuda.current_stream().cuda_stream)
    kw = dict(TE=TE, Reg="lcurve")
    whole = _alloc(torch, dev, nvox, nT2)
    st, p = _run(pkg, torch, img, nvox, 0, nvox, whole, kw, nT2)
    assert st["voxels_processed"] == nvox
    # sanity of the maps
    for k in ("gdn", "ggm", "gva", "fnr", "snr", "alpha", "sfr", "mfr"):  # sgm / mgm stay NaN where their window is empty
        assert bool(torch.isfinite(whole[k]).all().item()), k
    for k, w in (("sgm", "sfr"), ("mgm", "mfr")):  # ... and only there (src/T2partSEcorr.jl:126-135)
        assert _same(torch, torch.isnan(whole[k]).double(), (whole[w] == 0).double()), k
    assert bool((whole["dist"] >= 0).all().item())
    assert bool(((whole["sfr"] >= 0) & (whole["sfr"] <= 1)).all().item())
    assert bool(((whole["alpha"] >= 50.0) & (whole["alpha"] <= 180.0)).all().item())
    assert abs(float((whole["dist"].sum(0) - whole["gdn"]).abs().max().item())) < 1e-9 * float(whole["gdn"].max().item())
    # slabs of unequal, 4-voxel-aligned and unaligned sizes
    parts = _alloc(torch, dev, nvox, nT2)
    cuts = [0, pkg.slab_bounds(nvox, 3, 0)[1], pkg.slab_bounds(nvox, 3, 1)[1] + 3, nvox]
    for a, b in zip(cuts[:-1], cuts[1:]):
        _run(pkg, torch, img, nvox, a, b, parts, kw, nT2)
    for k in NAMES + ["dist"]:
        assert _same(torch, whole[k], parts[k]), f"slab invariance: {k}"
    del parts
    # determinism
    again = _alloc(torch, dev, nvox, nT2)
    _run(pkg, torch, img, nvox, 0, nvox, again, kw, nT2)
    for k in NAMES + ["dist"]:
        assert _same(torch, whole[k], again[k]), f"determinism: {k}"
    del again
    # fused epilogue == standalone T2partSEcorr on the stored distributions
    alone = {k: torch.full((nvox,), float("nan"), dtype=torch.float64, device=dev) for k in ("sfr", "sgm", "mfr", "mgm")}
    pkg.t2part_device(whole["dist"].data_ptr(), nvox, nvox, p, *[alone[k].data_ptr() for k in ("sfr", "sgm", "mfr", "mgm")],
                      stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for k in alone:  # same formulas, different summation order (warp butterfly vs one thread per voxel): ~1 ulp
        assert _same(torch, torch.isnan(whole[k]).double(), torch.isnan(alone[k]).double()), k
        ok = torch.isclose(whole[k], alone[k], rtol=1e-12, atol=0.0, equal_nan=True)
        assert bool(ok.all().item()), f"fused vs standalone T2part: {k}"


@pytest.mark.parametrize("Reg,extra", [("lcurve", {}), ("chi2", {"Chi2Factor": 1.02})])
def test_power_of_two_scale_covariance(pkg, Reg, extra):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    dev = torch.device("cuda", 0)
    nTE, TE, nT2, nvox = 48, 8e-3, 40, 100_000
    img = torch.empty((nTE, nvox), dtype=torch.float64, device=dev)
    pkg.mock_image_device(img.data_ptr(), nvox, nvox, 0, nTE, TE, seed=5, stream=torch.cuda.current_stream().cuda_stream)
    kw = dict(TE=TE, Reg=Reg, **extra)
    a = _alloc(torch, dev, nvox, nT2)
    _run(pkg, torch, img, nvox, 0, nvox, a, kw, nT2)
    img2 = img * 1024.0
    b = _alloc(torch, dev, nvox, nT2)
    _run(pkg, torch, img2, nvox, 0, nvox, b, kw, nT2)
    assert _same(torch, a["dist"] * 1024.0, b["dist"])
    assert _same(torch, a["gdn"] * 1024.0, b["gdn"])
    for k in ("ggm", "gva", "fnr", "snr", "alpha", "sfr", "sgm", "mfr", "mgm"):
        assert _same(torch, a[k], b[k]), k
