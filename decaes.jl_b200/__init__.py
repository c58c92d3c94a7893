"""decaes.jl_b200 — host-side mirror of DECAES.jl's T2mapSEcorr / T2partSEcorr API on top of the
C-ABI library libdecaes_cuda.so (B200 / sm_100a).

The reference is Julia (no runtime in this image), so this module plays the role of the Julia
shim (julia/DECAESCUDA.jl): same entry-point names, keyword options, defaults, assertions and
output dictionaries as src/T2mapSEcorr.jl:148-196, src/T2partSEcorr.jl:40-71 and
src/types.jl:19-190.  All numerics happen in the CUDA library; there is no CPU fallback and
importing the compute entry points without the built library raises.

Arrays follow Julia's memory order: pass `image` with shape (nx, ny, nz, nTE); it is converted
to Fortran order (voxel index fastest, echo slowest) before the call.
"""
import ctypes as C
import math
import os
from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np

from . import _abi
from ._abi import REG_CODES, RunStats, T2mapOpts, T2mapOut, T2partOpts

__all__ = ["T2mapOptions", "T2partOptions", "T2mapSEcorr", "T2partSEcorr", "lib", "build", "DecaesError",
           "last_stats", "device_count"]

_HERE = os.path.dirname(os.path.abspath(__file__))
# DECAES_LIB: developer override (e.g. an instrumented -DDECAES_PROFILE build next to the product library)
LIB_PATH = os.environ.get("DECAES_LIB") or os.path.join(_HERE, "libdecaes_cuda.so")
_lib = None


class DecaesError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"libdecaes_cuda status {status}: {message}")
        self.status = status


def build(force=False, verbose=False):
    from .build import build as _build
    return _build(force=force, verbose=verbose)


def lib():
    """Load libdecaes_cuda.so (fails loudly when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -m decaes_jl_b200.build` / __graft_entry__.build(); "
                              "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        dp, vp = C.POINTER(C.c_double), C.c_void_p
        L.decaes_last_error.restype = C.c_char_p
        L.decaes_t2map.argtypes = [vp, C.POINTER(T2mapOpts), C.POINTER(T2partOpts), C.POINTER(T2mapOut)]
        L.decaes_t2map_f32.argtypes = [vp, C.POINTER(T2mapOpts), C.POINTER(T2partOpts), C.POINTER(T2mapOut)]
        L.decaes_host_alloc.restype = vp
        L.decaes_host_alloc.argtypes = [C.c_size_t]
        L.decaes_host_free.restype = None
        L.decaes_host_free.argtypes = [vp]
        L.decaes_slab_bounds_masked.argtypes = [vp, C.c_int64, C.c_double, C.c_int32, C.POINTER(C.c_int64)]
        L.decaes_t2part.argtypes = [vp, C.POINTER(T2partOpts), vp, vp, vp, vp]
        L.decaes_setup_tables.argtypes = [C.POINTER(T2mapOpts), vp, vp, vp, vp]
        L.decaes_t2map_device.argtypes = [vp, C.c_int64, C.c_int64, C.POINTER(T2mapOpts), C.POINTER(T2partOpts),
                                          C.POINTER(T2mapOut), vp]
        L.decaes_t2part_device.argtypes = [vp, C.c_int64, C.c_int64, C.POINTER(T2partOpts), vp, vp, vp, vp, vp]
        L.decaes_mock_image_device.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_double, C.c_double,
                                               C.c_double, C.c_uint64, vp]
        L.decaes_get_stats.argtypes = [C.POINTER(RunStats)]
        L.decaes_release.restype = None
        L.decaes_measure_fp64_peak.argtypes = [dp]
        L.decaes_slab_bounds.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        for name in _abi.DECLARED_SYMBOLS:
            getattr(L, name)  # AttributeError if the header and the library disagree
        _lib = L
    return _lib


def _check(status):
    if status != 0:
        raise DecaesError(status, lib().decaes_last_error().decode(errors="replace"))


def device_count():
    return lib().decaes_device_count()


def last_stats():
    st = RunStats()
    lib().decaes_get_stats(C.byref(st))
    return {name: getattr(st, name) for name, _ in RunStats._fields_}


def release():
    """Give back the device workspaces the library caches between calls (decaes_release)."""
    lib().decaes_release()


# ------------------------------------------------------------------------------------ options
@dataclass
class T2mapOptions:
    """Mirror of DECAES.T2mapOptions (src/types.jl:19-100): same fields, defaults and assertions."""
    MatrixSize: Tuple[int, int, int]
    nTE: int
    TE: float
    nT2: int
    T2Range: Tuple[float, float]
    Reg: str
    legacy: bool = False
    Threaded: bool = True
    T1: float = 1.0
    Threshold: float = 0.0
    MinRefAngle: float = 50.0
    nRefAngles: Optional[int] = None
    nRefAnglesMin: Optional[int] = None
    Chi2Factor: Optional[float] = None
    NoiseLevel: Optional[float] = None
    RefConAngle: float = 180.0
    SetFlipAngle: Optional[float] = None
    SaveResidualNorm: bool = False
    SaveDecayCurve: bool = False
    SaveRegParam: bool = False
    SaveNNLSBasis: bool = False
    Silent: bool = False
    ngpus: int = 0  # extension: 0 = all visible devices

    def __post_init__(self):
        if self.nRefAngles is None:
            self.nRefAngles = 64 if not self.legacy else 8
        if self.nRefAnglesMin is None:
            self.nRefAnglesMin = min(5, self.nRefAngles) if not self.legacy else self.nRefAngles
        self.MatrixSize = tuple(int(s) for s in self.MatrixSize)
        a = self._assert
        a(len(self.MatrixSize) == 3 and all(s >= 1 for s in self.MatrixSize),
          f"MatrixSize must be a tuple of 3 positive integers, but MatrixSize = {self.MatrixSize}.")
        a(self.nTE >= 4, f"At least four echoes are required for T2 mapping, but nTE = {self.nTE}.")
        a(self.TE > 0.0, f"Echo spacing must be positive, but TE = {self.TE}.")
        a(self.nT2 >= 2, f"At least two T2 components are required for T2 mapping, but nT2 = {self.nT2}.")
        a(0.0 < self.T2Range[0] < self.T2Range[1],
          f"T2Range must a sorted 2-tuple of positive values, but T2Range = {self.T2Range}.")
        a(self.T1 > 0.0, f"T1 must be positive, but T1 = {self.T1}.")
        a(self.Threshold >= 0.0 or self.Threshold == -math.inf,
          f"First echo signal threshold must be non-negative or -Inf to force processing of every voxel, but Threshold = {self.Threshold}.")
        a(0.0 <= self.MinRefAngle <= 180.0,
          f"Minimum refocusing angle must be in the range [0, 180], but MinRefAngle = {self.MinRefAngle}.")
        a(self.nRefAngles >= 2,
          f"Maximum number of angles to check during flip angle optimization must be at least 2, but nRefAngles = {self.nRefAngles}.")
        a(2 <= self.nRefAnglesMin <= self.nRefAngles,
          f"Minimum number of angles to check during flip angle optimization must be in the range [2, nRefAngles], but nRefAngles = {self.nRefAngles} and nRefAnglesMin = {self.nRefAnglesMin}.")
        a(self.Reg in ("none", "lcurve", "gcv", "chi2", "mdp"), f"Unrecognized regularization method: {self.Reg}")
        a(self.Reg != "chi2" or (self.Chi2Factor is not None and self.Chi2Factor > 1.0),
          f"Chi2Factor must be greater than 1.0, but Chi2Factor = {self.Chi2Factor}.")
        a(self.Reg != "mdp" or (self.NoiseLevel is not None and self.NoiseLevel > 0.0),
          f"Noise level must be positive, but NoiseLevel = {self.NoiseLevel}.")
        a(0.0 <= self.RefConAngle <= 180.0,
          f"Refocusing control angle must be in the range [0, 180], but RefConAngle = {self.RefConAngle}.")
        a(self.SetFlipAngle is None or 0.0 <= self.SetFlipAngle <= 180.0,
          f"Fixed flip angle must be in the range [0, 180], but SetFlipAngle = {self.SetFlipAngle}.")

    @staticmethod
    def _assert(cond, msg):
        if not cond:
            raise AssertionError(msg)

    def to_c(self, alpha_provided=False):
        o = T2mapOpts()
        o.nx, o.ny, o.nz = self.MatrixSize
        o.nTE, o.nT2 = self.nTE, self.nT2
        o.nRefAngles, o.nRefAnglesMin = self.nRefAngles, self.nRefAnglesMin
        o.reg = REG_CODES[self.Reg]
        o.legacy = int(self.legacy)
        o.alpha_provided = int(alpha_provided)
        o.ngpus = self.ngpus
        o.TE, o.T2min, o.T2max, o.T1 = self.TE, self.T2Range[0], self.T2Range[1], self.T1
        o.Threshold, o.MinRefAngle, o.RefConAngle = self.Threshold, self.MinRefAngle, self.RefConAngle
        nan = float("nan")
        o.Chi2Factor = nan if self.Chi2Factor is None else self.Chi2Factor
        o.NoiseLevel = nan if self.NoiseLevel is None else self.NoiseLevel
        o.SetFlipAngle = nan if self.SetFlipAngle is None else self.SetFlipAngle
        return o


@dataclass
class T2partOptions:
    """Mirror of DECAES.T2partOptions (src/types.jl:139-172)."""
    MatrixSize: Tuple[int, int, int]
    nT2: int
    T2Range: Tuple[float, float]
    SPWin: Tuple[float, float]
    MPWin: Tuple[float, float]
    legacy: bool = False
    Threaded: bool = True
    Sigmoid: Optional[float] = None
    Silent: bool = False

    def __post_init__(self):
        self.MatrixSize = tuple(int(s) for s in self.MatrixSize)
        a = T2mapOptions._assert
        a(len(self.MatrixSize) == 3 and all(s >= 1 for s in self.MatrixSize), "MatrixSize must be positive")
        a(self.nT2 >= 2, "nT2 >= 2")
        a(0.0 < self.T2Range[0] < self.T2Range[1], "0.0 < T2Range[1] < T2Range[2]")
        a(self.SPWin[0] < self.SPWin[1], "SPWin[1] < SPWin[2]")
        a(self.MPWin[0] < self.MPWin[1], "MPWin[1] < MPWin[2]")
        a(self.Sigmoid is None or self.Sigmoid > 0, "Sigmoid === nothing || Sigmoid > 0")

    def to_c(self):
        p = T2partOpts()
        p.nx, p.ny, p.nz = self.MatrixSize
        p.nT2 = self.nT2
        p.T2min, p.T2max = self.T2Range
        p.SPWin_lo, p.SPWin_hi = self.SPWin
        p.MPWin_lo, p.MPWin_hi = self.MPWin
        p.Sigmoid = float("nan") if self.Sigmoid is None else self.Sigmoid
        return p


# ------------------------------------------------------------------------------------ API
def _nanfill(shape):
    """tfill(NaN, ...)  src/utils.jl:379-381 — Fortran order so that the voxel index is fastest."""
    a = np.empty(shape, dtype=np.float64, order="F")
    a.fill(np.nan)
    return a


def _ptr(a):
    return a.ctypes.data if a is not None else None


def T2mapSEcorr(image, opts: Optional[T2mapOptions] = None, B1map=None, t2part: Optional[T2partOptions] = None,
                **kwargs):
    """T2mapSEcorr(image; kwargs...) / T2mapSEcorr(image, opts)  (src/T2mapSEcorr.jl:148-196).

    Returns (maps, distributions) like the reference: `maps` is a dict with "echotimes",
    "t2times", "refangleset", "decaybasisset", "gdn", "ggm", "gva", "fnr", "snr", "alpha" and the
    optional "resnorm", "decaycurve", "mu", "chi2factor", "decaybasis".  Extension: passing
    `t2part=T2partOptions(...)` fuses the T2part epilogue into the same kernel and adds
    "sfr", "sgm", "mfr", "mgm" to `maps`.  `B1map` plays the role of load_B1map! (:56-60).
    """
    image = np.asarray(image)
    if image.ndim != 4:
        raise AssertionError("image must be a 4D array (row, column, slice, echo)")
    if opts is None:
        opts = T2mapOptions(MatrixSize=image.shape[:3], nTE=image.shape[3], **kwargs)
    elif kwargs:
        raise TypeError("pass either an options struct or keyword arguments")
    if tuple(image.shape) != (*opts.MatrixSize, opts.nTE):
        raise AssertionError(f"size(image) == (opts.MatrixSize..., opts.nTE) failed: {image.shape}")
    # Float32 volumes go to the library as they are (converted on the device, decaes_t2map_f32)
    is_f32 = image.dtype == np.float32
    img = np.asfortranarray(image, dtype=np.float32 if is_f32 else np.float64)
    L = lib()
    msz, nTE, nT2 = opts.MatrixSize, opts.nTE, opts.nT2
    fixed = opts.SetFlipAngle is not None
    copts = opts.to_c(alpha_provided=B1map is not None)

    maps = {}
    # table fields  (src/T2mapSEcorr.jl:28-33)
    nA = 1 if fixed else opts.nRefAngles
    maps["echotimes"] = np.empty(nTE)
    maps["t2times"] = np.empty(nT2)
    refangleset = np.empty(nA)
    basisset = np.empty((nTE, nT2, nA), order="F")
    _check(L.decaes_setup_tables(C.byref(copts), _ptr(maps["echotimes"]), _ptr(maps["t2times"]), _ptr(refangleset),
                                 _ptr(basisset)))
    maps["refangleset"] = float(refangleset[0]) if fixed else refangleset
    maps["decaybasisset"] = basisset[:, :, 0].copy(order="F") if fixed else basisset

    for k in ("gdn", "ggm", "gva", "fnr", "snr", "alpha"):
        maps[k] = _nanfill(msz)
    if B1map is not None:
        maps["alpha"][...] = np.asarray(B1map, dtype=np.float64)
    dist = _nanfill((*msz, nT2))
    if opts.SaveResidualNorm:
        maps["resnorm"] = _nanfill(msz)
    if opts.SaveDecayCurve:
        maps["decaycurve"] = _nanfill((*msz, nTE))
    if opts.SaveRegParam:
        maps["mu"] = _nanfill(msz)
        maps["chi2factor"] = _nanfill(msz)
    if opts.SaveNNLSBasis:
        maps["decaybasis"] = maps["decaybasisset"].copy(order="F") if fixed else _nanfill((*msz, nTE, nT2))
    cpart = None
    if t2part is not None:
        cpart = t2part.to_c()
        for k in ("sfr", "sgm", "mfr", "mgm"):
            maps[k] = _nanfill(msz)

    out = T2mapOut()
    for name in _abi.OUT_FIELDS:
        arr = dist if name == "dist" else maps.get(name)
        if name == "decaybasis" and fixed:
            arr = None
        setattr(out, name, _ptr(arr))
    entry = L.decaes_t2map_f32 if is_f32 else L.decaes_t2map
    _check(entry(_ptr(img), C.byref(copts), C.byref(cpart) if cpart is not None else None, C.byref(out)))
    return maps, dist


def T2partSEcorr(T2distributions, opts: Optional[T2partOptions] = None, **kwargs):
    """T2partSEcorr(T2distributions; kwargs...)  (src/T2partSEcorr.jl:40-71) -> dict sfr/sgm/mfr/mgm."""
    d = np.asarray(T2distributions)
    if d.ndim != 4:
        raise AssertionError("T2distributions must be a 4D array (row, column, slice, T2 amplitude)")
    if opts is None:
        opts = T2partOptions(MatrixSize=d.shape[:3], nT2=d.shape[3], **kwargs)
    elif kwargs:
        raise TypeError("pass either an options struct or keyword arguments")
    if tuple(d.shape) != (*opts.MatrixSize, opts.nT2):
        raise AssertionError(f"size(T2distributions) == (opts.MatrixSize..., opts.nT2) failed: {d.shape}")
    d = np.asfortranarray(d, dtype=np.float64)
    maps = {k: _nanfill(opts.MatrixSize) for k in ("sfr", "sgm", "mfr", "mgm")}
    cp = opts.to_c()
    _check(lib().decaes_t2part(_ptr(d), C.byref(cp), _ptr(maps["sfr"]), _ptr(maps["sgm"]), _ptr(maps["mfr"]),
                               _ptr(maps["mgm"])))
    return maps


# ------------------------------------------------------------------------------------ device-level helpers
def make_out(ptrs: dict) -> T2mapOut:
    """Build the output bundle from {name: raw pointer (int) or None}."""
    out = T2mapOut()
    for name in _abi.OUT_FIELDS:
        setattr(out, name, ptrs.get(name))
    return out


def t2map_device(d_image_ptr, nvox, stride, copts: T2mapOpts, cpart: Optional[T2partOpts], out: T2mapOut, stream=0):
    _check(lib().decaes_t2map_device(d_image_ptr, nvox, stride, C.byref(copts),
                                     C.byref(cpart) if cpart is not None else None, C.byref(out), stream))


def t2part_device(d_dist_ptr, nvox, stride, cpart: T2partOpts, sfr, sgm, mfr, mgm, stream=0):
    _check(lib().decaes_t2part_device(d_dist_ptr, nvox, stride, C.byref(cpart), sfr, sgm, mfr, mgm, stream))


def mock_image_device(d_image_ptr, nvox, stride, first_voxel, nTE, TE, T1=1.0, SNR=60.0, seed=1, stream=0):
    _check(lib().decaes_mock_image_device(d_image_ptr, nvox, stride, first_voxel, nTE, TE, T1, SNR, seed, stream))


def slab_bounds(nvox, nshards, index):
    """Voxel range [v0, v1) of shard `index` (decaes_slab_bounds): how volumes are split over GPUs / ranks."""
    v0, v1 = C.c_int64(), C.c_int64()
    _check(lib().decaes_slab_bounds(nvox, nshards, index, C.byref(v0), C.byref(v1)))
    return v0.value, v1.value


def slab_bounds_masked(first_echo, threshold, nshards):
    """Cuts of decaes_slab_bounds_masked: equal numbers of voxels above `threshold` per shard."""
    fe = np.ascontiguousarray(first_echo, dtype=np.float64).ravel()
    cuts = (C.c_int64 * (nshards + 1))()
    _check(lib().decaes_slab_bounds_masked(fe.ctypes.data, fe.size, float(threshold), nshards, cuts))
    return list(cuts)


def measure_fp64_peak():
    v = C.c_double()
    _check(lib().decaes_measure_fp64_peak(C.byref(v)))
    return v.value
