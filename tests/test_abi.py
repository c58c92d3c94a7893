"""CPU-side checks of the C ABI and the host mirror: the library loads, exports every symbol
declared in include/decaes_cuda.h, struct layouts match, option validation mirrors the
reference's assertions, and compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "decaes_cuda.h")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_symbols_are_exported(pkg):
    pkg.build()
    L = pkg.lib()
    src = open(HEADER).read()
    declared = set(re.findall(r"\b(decaes_[a-z0-9_]+)\s*\(", src))
    assert declared == set(pkg._abi.DECLARED_SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.decaes_abi_version() == 2


def test_struct_layouts_match_header(pkg):
    code = r"""
    #include <stdio.h>
    #include <stddef.h>
    #include "decaes_cuda.h"
    int main(void) {
      printf("%zu %zu %zu %zu\n", sizeof(decaes_t2map_opts), sizeof(decaes_t2part_opts), sizeof(decaes_t2map_out), sizeof(decaes_run_stats));
      printf("%zu %zu %zu\n", offsetof(decaes_t2map_opts, TE), offsetof(decaes_t2map_opts, SetFlipAngle), offsetof(decaes_t2part_opts, Sigmoid));
      printf("%zu %zu\n", offsetof(decaes_t2map_out, dist), offsetof(decaes_run_stats, total_ms));
      printf("%zu %zu\n", offsetof(decaes_run_stats, early_returns), offsetof(decaes_run_stats, pinned_staging));
      return 0;
    }"""
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(code)
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.run([cc, "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()
    a = pkg._abi
    got = [C.sizeof(a.T2mapOpts), C.sizeof(a.T2partOpts), C.sizeof(a.T2mapOut), C.sizeof(a.RunStats),
           a.T2mapOpts.TE.offset, a.T2mapOpts.SetFlipAngle.offset, a.T2partOpts.Sigmoid.offset,
           a.T2mapOut.dist.offset, a.RunStats.total_ms.offset, a.RunStats.early_returns.offset,
           a.RunStats.pinned_staging.offset]
    assert got == [int(x) for x in out]


def test_validation_through_the_abi(pkg, orc):
    L = pkg.lib()
    nvox = 8
    img = np.ones((32, nvox))
    arrs, out = orc.alloc_outputs(nvox, 32, 40, part=False)
    bad = orc.make_t2map_opts((nvox, 1, 1), 3, 40, 10e-3)  # nTE < 4
    assert L.decaes_t2map(img.ctypes.data, C.byref(bad), None, C.byref(out)) == -1
    assert b"four echoes" in L.decaes_last_error()
    bad = orc.make_t2map_opts((nvox, 1, 1), 32, 40, 10e-3, Reg="chi2")  # Chi2Factor unset
    assert L.decaes_t2map(img.ctypes.data, C.byref(bad), None, C.byref(out)) == -1
    big = orc.make_t2map_opts((nvox, 1, 1), 32, 65, 10e-3)  # nT2 > 64: outside the accelerated path
    assert L.decaes_t2map(img.ctypes.data, C.byref(big), None, C.byref(out)) == -3
    assert b"nT2 > 64" in L.decaes_last_error()
    long_train = orc.make_t2map_opts((nvox, 1, 1), 97, 40, 10e-3)  # nTE > 96: outside the accelerated path
    assert L.decaes_t2map(img.ctypes.data, C.byref(long_train), None, C.byref(out)) == -3
    assert b"nTE > 96" in L.decaes_last_error()
    good = orc.make_t2map_opts((nvox, 1, 1), 32, 40, 10e-3)
    noout = pkg._abi.T2mapOut()
    assert L.decaes_t2map(img.ctypes.data, C.byref(good), None, C.byref(noout)) == -1
    p = orc.make_t2part_opts((nvox, 1, 1), 40, SPWin=(25e-3, 10e-3))
    d = np.zeros((40, nvox))
    o4 = [np.zeros(nvox) for _ in range(4)]
    assert L.decaes_t2part(d.ctypes.data, C.byref(p), *[x.ctypes.data for x in o4]) == -1


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback(pkg, orc):
    L = pkg.lib()
    nvox = 8
    img = np.ones((32, nvox))
    arrs, out = orc.alloc_outputs(nvox, 32, 40, part=False)
    good = orc.make_t2map_opts((nvox, 1, 1), 32, 40, 10e-3)
    assert L.decaes_device_count() == 0
    assert L.decaes_t2map(img.ctypes.data, C.byref(good), None, C.byref(out)) == -2
    assert b"no CUDA device" in L.decaes_last_error()
    assert np.all(np.isnan(arrs["gdn"]))  # nothing was computed behind the caller's back
    with pytest.raises(pkg.DecaesError):
        pkg.T2partSEcorr(np.zeros((2, 2, 2, 40)), T2Range=(10e-3, 2.0), SPWin=(10e-3, 25e-3), MPWin=(25e-3, 0.2))


def test_options_mirror_reference_asserts(pkg):
    O = pkg.T2mapOptions
    base = dict(MatrixSize=(2, 2, 2), nTE=32, TE=10e-3, nT2=40, T2Range=(10e-3, 2.0), Reg="none")
    o = O(**base)
    assert (o.nRefAngles, o.nRefAnglesMin, o.T1, o.Threshold, o.MinRefAngle, o.RefConAngle) == (64, 5, 1.0, 0.0, 50.0, 180.0)
    assert O(**{**base, "legacy": True}).nRefAngles == 8
    for kw in [dict(nTE=3), dict(nT2=1), dict(TE=0.0), dict(T2Range=(1.0, 0.5)), dict(T1=-1.0), dict(Threshold=-0.5),
               dict(MinRefAngle=181.0), dict(nRefAngles=1), dict(nRefAnglesMin=70), dict(Reg="foo"), dict(Reg="chi2"),
               dict(Reg="chi2", Chi2Factor=1.0), dict(Reg="mdp"), dict(RefConAngle=-1.0), dict(SetFlipAngle=200.0),
               dict(MatrixSize=(0, 1, 1))]:
        with pytest.raises(AssertionError):
            O(**{**base, **kw})
    assert O(**{**base, "Threshold": -np.inf}).Threshold == -np.inf
    c = O(**{**base, "Reg": "mdp", "NoiseLevel": 1e-3}).to_c()
    assert c.reg == 4 and c.NoiseLevel == 1e-3 and np.isnan(c.Chi2Factor) and np.isnan(c.SetFlipAngle)
    P = pkg.T2partOptions
    pb = dict(MatrixSize=(2, 2, 2), nT2=40, T2Range=(10e-3, 2.0), SPWin=(10e-3, 25e-3), MPWin=(25e-3, 0.2))
    assert np.isnan(P(**pb).to_c().Sigmoid)
    for kw in [dict(nT2=1), dict(SPWin=(1.0, 0.5)), dict(MPWin=(1.0, 1.0)), dict(Sigmoid=0.0)]:
        with pytest.raises(AssertionError):
            P(**{**pb, **kw})


def test_julia_shim_names_every_abi_field():
    """julia/DECAESCUDA.jl is untestable here (no Julia); at least keep it in sync with the header."""
    shim = open(os.path.join(ROOT, "julia", "DECAESCUDA.jl")).read()
    hdr = open(HEADER).read()
    m = re.search(r"typedef struct \{(.*?)\} decaes_t2map_opts;", hdr, re.S)
    fields = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*[;,]", re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S))
    for f in fields:
        assert re.search(r"\b%s\b" % f, shim), f
    for sym in ("decaes_t2map", "decaes_t2part", "decaes_last_error"):
        assert sym in shim
