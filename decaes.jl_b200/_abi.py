"""ctypes mirror of include/decaes_cuda.h (flat option structs, output bundle, stats).

Field order and types must match the header exactly; tests/test_abi.py checks sizes and
that every declared symbol is exported by libdecaes_cuda.so.
"""
import ctypes as C

c_double_p = C.POINTER(C.c_double)

REG_CODES = {"none": 0, "lcurve": 1, "gcv": 2, "chi2": 3, "mdp": 4}

DECAES_OK = 0
DECAES_EINVAL = -1
DECAES_ECUDA = -2
DECAES_EUNSUPPORTED = -3
DECAES_ENOMEM = -4


class T2mapOpts(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
        ("nTE", C.c_int32), ("nT2", C.c_int32),
        ("nRefAngles", C.c_int32), ("nRefAnglesMin", C.c_int32),
        ("reg", C.c_int32), ("legacy", C.c_int32), ("alpha_provided", C.c_int32),
        ("ngpus", C.c_int32), ("reserved", C.c_int32),
        ("TE", C.c_double), ("T2min", C.c_double), ("T2max", C.c_double),
        ("T1", C.c_double), ("Threshold", C.c_double), ("MinRefAngle", C.c_double),
        ("RefConAngle", C.c_double), ("Chi2Factor", C.c_double), ("NoiseLevel", C.c_double),
        ("SetFlipAngle", C.c_double),
    ]


class T2partOpts(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("nT2", C.c_int32),
        ("T2min", C.c_double), ("T2max", C.c_double),
        ("SPWin_lo", C.c_double), ("SPWin_hi", C.c_double),
        ("MPWin_lo", C.c_double), ("MPWin_hi", C.c_double),
        ("Sigmoid", C.c_double),
    ]


OUT_FIELDS = ["gdn", "ggm", "gva", "fnr", "snr", "alpha", "dist", "resnorm", "decaycurve",
              "mu", "chi2factor", "decaybasis", "sfr", "sgm", "mfr", "mgm"]


class T2mapOut(C.Structure):
    _fields_ = [(name, C.c_void_p) for name in OUT_FIELDS]


class RunStats(C.Structure):
    _fields_ = [
        ("voxels_total", C.c_int64), ("voxels_processed", C.c_int64),
        ("ngpus_used", C.c_int32), ("kernel_launches", C.c_int32),
        ("setup_ms", C.c_double), ("pipeline_ms", C.c_double),
        ("h2d_ms", C.c_double), ("d2h_ms", C.c_double), ("total_ms", C.c_double),
        # ABI v2 (appended): counted voxels, see include/decaes_cuda.h
        ("early_returns", C.c_int64), ("lcurve_overflow", C.c_int64), ("nnls_itercap", C.c_int64),
        ("pinned_staging", C.c_int64),
    ]


# every extern "C" symbol declared in include/decaes_cuda.h
DECLARED_SYMBOLS = [
    "decaes_t2map", "decaes_t2part", "decaes_setup_tables", "decaes_t2map_device",
    "decaes_t2part_device", "decaes_mock_image_device", "decaes_last_error",
    "decaes_device_count", "decaes_abi_version", "decaes_get_stats", "decaes_measure_fp64_peak",
    "decaes_slab_bounds", "decaes_release", "decaes_t2map_f32", "decaes_host_alloc", "decaes_host_free", "decaes_slab_bounds_masked",
]
