"""Multi-GPU: decaes_t2map with ngpus = all visible devices must equal the single-GPU result bit for bit
(voxels are independent; slabs are only a partition).  Skipped when fewer than two devices are visible."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_host_api_multi_gpu_equals_single(pkg, orc):
    ndev = pkg.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    nvox, nTE, nT2, TE = 4099, 32, 40, 10e-3   # odd size: last slab takes the remainder
    img = orc.mock_image(nvox, nTE, TE, seed=21)
    img[::7, 0] = 0.0
    res = []
    for ng in (1, ndev):
        o = orc.make_t2map_opts((nvox, 1, 1), nTE, nT2, TE, Reg="chi2", Chi2Factor=1.02, ngpus=ng)
        p = orc.make_t2part_opts((nvox, 1, 1), nT2)
        arrs, out = orc.alloc_outputs(nvox, nTE, nT2, part=True)
        rc = pkg.lib().decaes_t2map(img.ctypes.data, C.byref(o), C.byref(p), C.byref(out))
        assert rc == 0, pkg.lib().decaes_last_error().decode()
        st = pkg.last_stats()
        assert st["ngpus_used"] == ng and st["voxels_processed"] == int((img[:, 0] > 0).sum())
        res.append(arrs)
    for k in res[0]:
        np.testing.assert_array_equal(res[0][k], res[1][k], err_msg=k)


def test_release_gives_back_the_cached_workspaces(pkg, orc):
    """decaes_release frees the grow-only device caches (slab copy, scratch, tables); the next call re-allocates
    and reproduces the result bit for bit.  Sizes shrink and grow between the calls on purpose."""
    import torch
    nTE, nT2, TE = 32, 40, 10e-3
    res = {}
    for tag, nvox in (("a", 2048), ("b", 512), ("released", 2048)):
        if tag == "released":
            used_before = torch.cuda.mem_get_info()[0]
            pkg.release()
            assert torch.cuda.mem_get_info()[0] >= used_before  # free memory did not shrink
        img = orc.mock_image(nvox, nTE, TE, seed=5)
        o = orc.make_t2map_opts((nvox, 1, 1), nTE, nT2, TE, Reg="lcurve", ngpus=1)
        arrs, out = orc.alloc_outputs(nvox, nTE, nT2, part=False)
        rc = pkg.lib().decaes_t2map(img.ctypes.data, C.byref(o), None, C.byref(out))
        assert rc == 0, pkg.lib().decaes_last_error().decode()
        res[tag] = arrs
    for k in res["a"]:
        np.testing.assert_array_equal(res["a"][k], res["released"][k], err_msg=k)
