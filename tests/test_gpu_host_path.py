"""Host-pointer path of the C ABI as the real caller uses it: pageable host buffers (ordinary
arrays, what Julia hands over, src/T2mapSEcorr.jl:24-54) staged through the library's pinned ring, page-locked buffers
copied directly, Float32 volumes converted on the device, several sub-slabs per device, masked volumes."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(pkg, orc, img, o, p, entry="decaes_t2map", arrs_out=None, **kw):
    nvox, nTE = img.shape
    arrs, out = arrs_out if arrs_out is not None else orc.alloc_outputs(nvox, nTE, o.nT2, part=p is not None, **kw)
    rc = getattr(pkg.lib(), entry)(img.ctypes.data, C.byref(o), C.byref(p) if p is not None else None, C.byref(out))
    assert rc == 0, pkg.lib().decaes_last_error().decode()
    return arrs, pkg.last_stats()


def _pinned_like(pkg, a):
    """Copy of `a` in page-locked memory from the library's own allocator (decaes_host_alloc)."""
    L = pkg.lib()
    ptr = L.decaes_host_alloc(a.nbytes)
    assert ptr
    buf = (C.c_char * a.nbytes).from_address(ptr)
    out = np.frombuffer(buf, dtype=a.dtype).reshape(a.shape, order="F" if a.flags.f_contiguous else "C")
    out[...] = a
    return out, ptr


@pytest.mark.parametrize("nvox", [1000, 300_000])  # one sub-slab / eight sub-slabs with small first and last ones
def test_pageable_equals_pinned_bit_for_bit(pkg, orc, nvox):
    nTE, nT2, TE = 32, 40, 10e-3
    img = orc.mock_image(nvox, nTE, TE, seed=31)
    img[::9, 0] = 0.0
    o = orc.make_t2map_opts((nvox, 1, 1), nTE, nT2, TE, Reg="none", ngpus=1)
    p = orc.make_t2part_opts((nvox, 1, 1), nT2)
    a, st = _run(pkg, orc, img, o, p, save_curve=True)
    assert st["pinned_staging"] == 1 and st["voxels_processed"] == int((img[:, 0] > 0).sum())
    # page-locked input and outputs: direct copies
    pimg, hp = _pinned_like(pkg, img)
    arrs, out = orc.alloc_outputs(nvox, nTE, nT2, part=True, save_curve=True)
    keep = []
    for k in list(arrs):
        arrs[k], ptr = _pinned_like(pkg, arrs[k])
        keep.append(ptr)
        setattr(out, k, arrs[k].ctypes.data)
    b, st = _run(pkg, orc, pimg, o, p, arrs_out=(arrs, out))
    assert st["pinned_staging"] == 0
    for k in a:
        np.testing.assert_array_equal(a[k].ravel(), b[k].ravel(), err_msg=k)
    for ptr in keep + [hp]:
        pkg.lib().decaes_host_free(ptr)


def test_b1_map_through_the_staging_ring(pkg, orc):
    nvox, nTE, nT2, TE = 300_000, 32, 40, 10e-3
    img = orc.mock_image(nvox, nTE, TE, seed=32)
    b1 = np.linspace(120.0, 180.0, nvox)
    o = orc.make_t2map_opts((nvox, 1, 1), nTE, nT2, TE, Reg="none", alpha_provided=True, ngpus=1)
    got, st = _run(pkg, orc, img, o, None, alpha_init=b1)
    assert st["pinned_staging"] == 1
    np.testing.assert_array_equal(got["alpha"], b1)
    ref, _ = orc.t2map(img[:2048], orc.make_t2map_opts((2048, 1, 1), nTE, nT2, TE, Reg="none", alpha_provided=True),
                       alpha_init=b1[:2048])
    np.testing.assert_allclose(got["dist"].reshape(nT2, nvox).T[:2048], ref["dist"], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("Reg,extra", [("none", {}), ("lcurve", {})])
def test_float32_volume_entry_point(pkg, orc, Reg, extra):
    """decaes_t2map_f32 == decaes_t2map on the widened image, bit for bit (the conversion is exact)."""
    nvox, nTE, nT2, TE = 70_000, 48, 40, 8e-3
    img32 = np.asfortranarray(orc.mock_image(nvox, nTE, TE, seed=33).astype(np.float32))
    img32[::11, 0] = 0.0
    o = orc.make_t2map_opts((nvox, 1, 1), nTE, nT2, TE, Reg=Reg, ngpus=1, **extra)
    p = orc.make_t2part_opts((nvox, 1, 1), nT2)
    a, st = _run(pkg, orc, img32, o, p, entry="decaes_t2map_f32")
    assert st["voxels_processed"] == int((img32[:, 0] > 0).sum())
    b, _ = _run(pkg, orc, np.asfortranarray(img32.astype(np.float64)), o, p)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    # and through the Python mirror of T2mapSEcorr(image::Array{Float32,4})
    maps, dist = pkg.T2mapSEcorr(img32[:600].reshape(10, 10, 6, nTE, order="F"), TE=TE, nT2=nT2, T2Range=(10e-3, 2.0),
                                 Reg=Reg, Silent=True, ngpus=1)
    np.testing.assert_array_equal(dist.reshape(600, nT2, order="F"), a["dist"].reshape(nT2, nvox).T[:600])


def test_run_stats_report_counted_voxels(pkg, orc):
    """ABI v2: early returns of the chi2 / MDP choosers, L-curve cache overflows and NNLS iteration caps are reported
    (north_star: "counted and reported"), and the caches never overflow on the benchmark configurations."""
    for nTE, TE, nT2, Reg, extra in [(32, 10e-3, 40, "none", {}), (48, 8e-3, 40, "lcurve", {}), (56, 7e-3, 40, "lcurve", {}),
                                     (48, 8e-3, 60, "chi2", {"Chi2Factor": 1.02}), (48, 8e-3, 60, "gcv", {}),
                                     (32, 10e-3, 60, "mdp", {"NoiseLevel": 1e-3})]:
        nvox = 512 if Reg == "gcv" else 4096
        img = orc.mock_image(nvox, nTE, TE, seed=7)
        o = orc.make_t2map_opts((nvox, 1, 1), nTE, nT2, TE, Reg=Reg, ngpus=1, **extra)
        _, st = _run(pkg, orc, img, o, None)
        ref, ost = orc.t2map(img, o)
        assert st["lcurve_overflow"] == 0 and st["nnls_itercap"] == 0, (Reg, st)
        assert st["early_returns"] == ost.early_returns, (Reg, st["early_returns"], ost.early_returns)
        if Reg == "mdp":
            assert st["early_returns"] > 0  # the delta <= sqrt(res2_min) branch is taken by a good share of cfg5's voxels
