"""ctypes binding of the CPU oracle (oracle/_build/liborc.so) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.
"""
import ctypes as C
import importlib.util
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "liborc.so")

def _load_package():
    """The package directory `decaes.jl_b200/` has a dot in its name; import it as `decaes_jl_b200`
    (same module name as tests/conftest.py and __graft_entry__.py so the ctypes classes are shared)."""
    name = "decaes_jl_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg_dir = os.path.join(ROOT, "decaes.jl_b200")
    spec = importlib.util.spec_from_file_location(name, os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


abi = _load_package()._abi  # struct definitions only; importing the package does not load the CUDA library

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


def build(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h"))]
    srcs.append(os.path.join(ROOT, "include", "decaes_cuda.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        have_cc = any(os.access(os.path.join(p, "gcc"), os.X_OK) for p in os.environ.get("PATH", "").split(":"))
        if not have_cc and os.path.exists(LIB_PATH):
            return LIB_PATH
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


class NnlsWork(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("A", dp), ("b", dp), ("x", dp), ("w", dp), ("zz", dp),
                ("idx", ip), ("invidx", ip), ("diag", C.POINTER(C.c_ubyte)), ("rnorm", C.c_double),
                ("mode", C.c_int), ("nsetp", C.c_int), ("n_enter", C.c_int64), ("n_exit", C.c_int64),
                ("n_reject", C.c_int64), ("flops", C.c_double)]


class RegWork(C.Structure):
    _fields_ = [("m", C.c_int), ("n", C.c_int), ("A", dp), ("b", dp), ("nnls", C.POINTER(NnlsWork)),
                ("cache", C.c_void_p), ("gamma", dp), ("svd_work", dp), ("lcurve", C.c_void_p),
                ("n_solves_unreg", C.c_int64), ("n_solves_tikh", C.c_int64), ("n_cache_hits", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [("voxels_processed", C.c_int64), ("nnls_unreg", C.c_int64), ("nnls_tikh", C.c_int64),
                ("cols_entered", C.c_int64), ("cols_exited", C.c_int64), ("cols_rejected", C.c_int64),
                ("cache_hits", C.c_int64), ("early_returns", C.c_int64), ("flops", C.c_double),
                ("seconds", C.c_double), ("threads", C.c_int)]


class JlRange(C.Structure):
    _fields_ = [("rational", C.c_int), ("start_n", C.c_int64), ("step_n", C.c_int64), ("den", C.c_int64),
                ("len", C.c_int64), ("start", C.c_double), ("step", C.c_double)]


LCURVE_FN = C.CFUNCTYPE(None, C.c_double, dp, C.c_void_p)
FN1 = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)
FG_FN = C.CFUNCTYPE(None, C.c_int, dp, dp, C.c_void_p)

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_sind.restype = C.c_double
        L.orc_sind.argtypes = [C.c_double]
        L.orc_hypot.restype = C.c_double
        L.orc_hypot.argtypes = [C.c_double, C.c_double]
        L.orc_nnls_alloc.restype = C.POINTER(NnlsWork)
        L.orc_nnls_alloc.argtypes = [C.c_int, C.c_int]
        L.orc_nnls_free.argtypes = [C.POINTER(NnlsWork)]
        L.orc_reg_alloc.restype = C.POINTER(RegWork)
        L.orc_reg_alloc.argtypes = [C.c_int, C.c_int]
        L.orc_reg_free.argtypes = [C.POINTER(RegWork)]
        L.orc_reg_bind.argtypes = [C.POINTER(RegWork), dp, dp]
        for name in ("orc_lsqnonneg",):
            getattr(L, name).restype = dp
            getattr(L, name).argtypes = [C.POINTER(RegWork)]
        L.orc_lsqnonneg_tikh.restype = dp
        L.orc_lsqnonneg_tikh.argtypes = [C.POINTER(RegWork), C.c_double, dp, dp]
        for name in ("orc_lsqnonneg_lcurve", "orc_lsqnonneg_gcv"):
            getattr(L, name).restype = dp
            getattr(L, name).argtypes = [C.POINTER(RegWork), dp, dp]
        for name in ("orc_lsqnonneg_chi2", "orc_lsqnonneg_mdp"):
            getattr(L, name).restype = dp
            getattr(L, name).argtypes = [C.POINTER(RegWork), C.c_double, dp, dp, ip]
        L.orc_gcv_dof.restype = C.c_double
        L.orc_gcv_dof.argtypes = [C.c_int, C.c_int, dp, C.c_double]
        L.orc_lcurve_corner.restype = C.c_double
        L.orc_lcurve_corner.argtypes = [LCURVE_FN, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double,
                                        C.c_double, C.c_int, ip]
        L.orc_brent_root.argtypes = [FN1, C.c_void_p] + [C.c_double] * 7 + [C.c_int, dp, dp]
        L.orc_bracket_root_monotonic.argtypes = [FN1, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int,
                                                 C.c_int, dp, dp, dp, dp]
        L.orc_brent_minimize.argtypes = [FN1, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                                         dp, dp]
        L.orc_hermite_minimize.argtypes = [C.c_double] * 6 + [dp, dp]
        L.orc_surrogate_search.argtypes = [FG_FN, C.c_void_p, dp, C.c_int, C.c_int, C.c_int, dp, dp, ip, ip]
        L.orc_surrogate_search_legacy.argtypes = L.orc_surrogate_search.argtypes
        L.orc_fitpack_interp.argtypes = [dp, dp, C.c_int, C.c_int, dp, dp]
        L.orc_fitpack_splev.argtypes = [dp, C.c_int, dp, C.c_int, dp, C.c_int, dp]
        L.orc_jl_range.argtypes = [C.c_double, C.c_double, C.c_double, C.POINTER(JlRange)]
        L.orc_jl_range_at.restype = C.c_double
        L.orc_jl_range_at.argtypes = [C.POINTER(JlRange), C.c_int64]
        L.orc_spline_opt_legacy.argtypes = [dp, dp, C.c_int, dp, dp]
        L.orc_spline_root_legacy.argtypes = [dp, dp, C.c_int, C.c_double, dp]
        L.orc_chi2_search_legacy.argtypes = [FN1, C.c_void_p, C.c_double, C.c_double, dp, dp]
        L.orc_lsqnonneg_chi2_legacy.restype = dp
        L.orc_lsqnonneg_chi2_legacy.argtypes = [C.POINTER(RegWork), C.c_double, dp, dp, ip]
        L.orc_t2map.argtypes = [dp, C.c_int64, C.c_int64, C.POINTER(abi.T2mapOpts), C.POINTER(abi.T2partOpts),
                                C.POINTER(abi.T2mapOut), C.c_int, C.POINTER(Stats)]
        L.orc_t2part.argtypes = [dp, C.c_int64, C.c_int64, C.POINTER(abi.T2partOpts), dp, dp, dp, dp]
        L.orc_setup_tables.argtypes = [C.POINTER(abi.T2mapOpts), dp, dp, dp, dp, dp]
        L.orc_validate_t2map_opts.argtypes = [C.POINTER(abi.T2mapOpts), C.c_char_p, C.c_int]
        L.orc_validate_t2part_opts.argtypes = [C.POINTER(abi.T2partOpts), C.c_char_p, C.c_int]
        L.orc_mock_image.argtypes = [dp, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_double, C.c_double,
                                     C.c_double, C.c_uint64]
        L.orc_nnls.argtypes = [C.POINTER(NnlsWork), dp, dp]
        L.orc_nnls_tikh_explicit.argtypes = [C.POINTER(NnlsWork), dp, dp, C.c_double]
        L.orc_nnls_solve.argtypes = [C.POINTER(NnlsWork), dp, C.c_int, dp, C.c_int, C.c_int]
        L.orc_nnls_solve_tikh.argtypes = [C.POINTER(NnlsWork), dp, C.c_int, dp, C.c_int, C.c_int, C.c_double]
        L.orc_logrange.argtypes = [C.c_double, C.c_double, C.c_int, dp]
        L.orc_linrange.argtypes = [C.c_double, C.c_double, C.c_int, dp]
        L.orc_epg_decay_curve.argtypes = [C.c_int] + [C.c_double] * 4 + [dp, dp]
        L.orc_epg_decay_curve_jac.argtypes = [C.c_int] + [C.c_double] * 4 + [dp, dp, dp]
        L.orc_epg_decay_curve_beta_jac.argtypes = [C.c_int] + [C.c_double] * 5 + [dp, dp, dp]
        L.orc_epg_decay_curve_beta.argtypes = [C.c_int] + [C.c_double] * 5 + [dp, dp]
        L.orc_svdvals.argtypes = [C.c_int, C.c_int, dp, C.c_int, dp, dp]
        L.orc_solve_triangular.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(dp) if a is not None else None


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ----------------------------------------------------------------------------- helpers
def logrange(a, b, n):
    out = np.empty(n)
    lib().orc_logrange(C.c_double(a), C.c_double(b), n, _p(out))
    return out


def linrange(a, b, n):
    out = np.empty(n)
    lib().orc_linrange(C.c_double(a), C.c_double(b), n, _p(out))
    return out


def epg(ETL, alpha, TE, T2, T1, beta=None):
    dc = np.empty(ETL)
    work = np.empty(6 * ETL)
    if beta is None:
        lib().orc_epg_decay_curve(ETL, C.c_double(alpha), C.c_double(TE), C.c_double(T2), C.c_double(T1), _p(dc), _p(work))
    else:
        lib().orc_epg_decay_curve_beta(ETL, C.c_double(alpha), C.c_double(TE), C.c_double(T2), C.c_double(T1),
                                       C.c_double(beta), _p(dc), _p(work))
    return dc


def epg_beta_jac(ETL, alpha, TE, T2, T1, beta):
    dc, ddc, work = np.zeros(ETL), np.zeros(ETL), np.zeros(12 * ETL)
    lib().orc_epg_decay_curve_beta_jac(ETL, C.c_double(alpha), C.c_double(TE), C.c_double(T2), C.c_double(T1),
                                       C.c_double(beta), _p(dc), _p(ddc), _p(work))
    return dc, ddc


def epg_jac(ETL, alpha, TE, T2, T1):
    dc, ddc = np.empty(ETL), np.empty(ETL)
    work = np.empty(12 * ETL)
    lib().orc_epg_decay_curve_jac(ETL, C.c_double(alpha), C.c_double(TE), C.c_double(T2), C.c_double(T1), _p(dc),
                                  _p(ddc), _p(work))
    return dc, ddc


class NnlsResult:
    pass


def _nnls_result(wp, M, N):
    w = wp.contents
    r = NnlsResult()
    r.x = np.ctypeslib.as_array(w.x, (N,)).copy()
    r.w = np.ctypeslib.as_array(w.w, (N,)).copy()
    r.idx = np.ctypeslib.as_array(w.idx, (N,)).copy()
    r.invidx = np.ctypeslib.as_array(w.invidx, (N,)).copy()
    r.A = np.ctypeslib.as_array(w.A, (N, M)).T.copy()  # column-major M x N
    r.b = np.ctypeslib.as_array(w.b, (M,)).copy()
    r.rnorm, r.mode, r.nsetp = w.rnorm, w.mode, w.nsetp
    r.n_enter, r.n_exit, r.n_reject = w.n_enter, w.n_exit, w.n_reject
    return r


def nnls(A, b, mu=None, warm=False):
    """mu=None: plain problem.  warm=False -> NNLS.nnls! (cold dual);  warm=True -> lsqnonneg.jl solve!"""
    A = np.asfortranarray(A, dtype=np.float64)
    b = f64(b)
    m, n = A.shape
    L = lib()
    if mu is None:
        wp = L.orc_nnls_alloc(m, n)
        if warm:
            L.orc_nnls_solve(wp, _p(A), m, _p(b), m, n)
        else:
            L.orc_nnls(wp, _p(A), _p(b))
        r = _nnls_result(wp, m, n)
    else:
        wp = L.orc_nnls_alloc(m + n, n)
        if warm:
            L.orc_nnls_solve_tikh(wp, _p(A), m, _p(b), m, n, C.c_double(mu))
        else:
            Ap = np.asfortranarray(np.vstack([A, mu * np.eye(n)]))
            bp = np.concatenate([b, np.zeros(n)])
            L.orc_nnls_tikh_explicit(wp, _p(Ap), _p(bp), C.c_double(mu))
        r = _nnls_result(wp, m + n, n)
    L.orc_nnls_free(wp)
    return r


def svdvals(A):
    A = np.asfortranarray(A, dtype=np.float64)
    m, n = A.shape
    S = np.empty(min(m, n))
    work = np.empty(m * n + m + n)
    lib().orc_svdvals(m, n, _p(A), m, _p(S), _p(work))
    return S


class Reg:
    """Regularised NNLS problem bound to (A, b) — mirrors lsqnonneg_*_work + lsqnonneg_*!"""

    def __init__(self, A, b):
        self.A = np.asfortranarray(A, dtype=np.float64)
        self.b = f64(b)
        self.m, self.n = self.A.shape
        self.L = lib()
        self.w = self.L.orc_reg_alloc(self.m, self.n)
        self.L.orc_reg_bind(self.w, _p(self.A), _p(self.b))

    def __del__(self):
        try:
            self.L.orc_reg_free(self.w)
        except Exception:
            pass

    def _x(self, ptr):
        return np.ctypeslib.as_array(ptr, (self.n,)).copy()

    def none(self):
        return self._x(self.L.orc_lsqnonneg(self.w))

    def tikh(self, mu):
        r2, s2 = C.c_double(), C.c_double()
        x = self._x(self.L.orc_lsqnonneg_tikh(self.w, C.c_double(mu), C.byref(r2), C.byref(s2)))
        return x, r2.value, s2.value

    def lcurve(self):
        mu, chi2 = C.c_double(), C.c_double()
        x = self._x(self.L.orc_lsqnonneg_lcurve(self.w, C.byref(mu), C.byref(chi2)))
        return x, mu.value, chi2.value

    def gcv(self):
        mu, chi2 = C.c_double(), C.c_double()
        x = self._x(self.L.orc_lsqnonneg_gcv(self.w, C.byref(mu), C.byref(chi2)))
        return x, mu.value, chi2.value

    def chi2(self, target):
        mu, chi2, early = C.c_double(), C.c_double(), C.c_int()
        x = self._x(self.L.orc_lsqnonneg_chi2(self.w, C.c_double(target), C.byref(mu), C.byref(chi2), C.byref(early)))
        return x, mu.value, chi2.value, early.value

    def mdp(self, delta):
        mu, chi2, early = C.c_double(), C.c_double(), C.c_int()
        x = self._x(self.L.orc_lsqnonneg_mdp(self.w, C.c_double(delta), C.byref(mu), C.byref(chi2), C.byref(early)))
        return x, mu.value, chi2.value, early.value

    @property
    def stats(self):
        w = self.w.contents
        return dict(unreg=w.n_solves_unreg, tikh=w.n_solves_tikh, hits=w.n_cache_hits)


def lcurve_corner(f, xlow, xhigh, xtol=1e-4, Ptol=1e-4, Ctol=1e-4, backtracking=True):
    def cb(t, P, _ctx):
        a, b = f(t)
        P[0], P[1] = a, b
    n = C.c_int()
    x = lib().orc_lcurve_corner(LCURVE_FN(cb), None, xlow, xhigh, xtol, Ptol, Ctol, int(backtracking), C.byref(n))
    return x, n.value


def brent_root(f, x0, x1, xatol=0.0, xrtol=0.0, ftol=0.0, maxiters=100):
    cb = FN1(lambda x, _c: f(x))
    xo, fo = C.c_double(), C.c_double()
    lib().orc_brent_root(cb, None, x0, x1, f(x0), f(x1), xatol, xrtol, ftol, maxiters, C.byref(xo), C.byref(fo))
    return xo.value, fo.value


def bracket_root_monotonic(f, a, delta, dilate=1.0, mono=+1, maxiters=100):
    cb = FN1(lambda x, _c: f(x))
    o = [C.c_double() for _ in range(4)]
    lib().orc_bracket_root_monotonic(cb, None, a, delta, dilate, mono, maxiters, *[C.byref(v) for v in o])
    return tuple(v.value for v in o)


def brent_minimize(f, x1, x2, xrtol=np.sqrt(np.finfo(float).eps), xatol=np.sqrt(np.finfo(float).eps), maxiters=100):
    cb = FN1(lambda x, _c: f(x))
    xo, yo = C.c_double(), C.c_double()
    lib().orc_brent_minimize(cb, None, x1, x2, xrtol, xatol, maxiters, C.byref(xo), C.byref(yo))
    return xo.value, yo.value


def hermite_minimize(a, b, u0, u1, m0, m1):
    x, u = C.c_double(), C.c_double()
    lib().orc_hermite_minimize(a, b, u0, u1, m0, m1, C.byref(x), C.byref(u))
    return x.value, u.value


def surrogate_search(fg, grid, mineval, maxeval, legacy=False):
    grid = f64(grid)

    def cb(I, u, du, _c):
        a, b = fg(I)
        u[0], du[0] = a, b
    order = np.zeros(len(grid) + 4, dtype=np.int32)
    norder = C.c_int()
    x, u = C.c_double(), C.c_double()
    fn = lib().orc_surrogate_search_legacy if legacy else lib().orc_surrogate_search
    fn(FG_FN(cb), None, _p(grid), len(grid), mineval, maxeval, C.byref(x), C.byref(u),
                               order.ctypes.data_as(ip), C.byref(norder))
    return x.value, u.value, order[:norder.value].tolist()


# ----------------------------------------------------------------------------- legacy = true helpers
def fitpack_interp(x, y, k):
    x, y = f64(x), f64(y)
    t, c = np.zeros(len(x) + k + 1), np.zeros(len(x))
    rc = lib().orc_fitpack_interp(_p(x), _p(y), len(x), k, _p(t), _p(c))
    if rc != 0:
        raise ValueError("orc_fitpack_interp failed")
    return t, c


def fitpack_splev(t, c, k, x):
    t, c, x = f64(t), f64(c), f64(x)
    y = np.zeros(len(x))
    lib().orc_fitpack_splev(_p(t), len(t), _p(c), k, _p(x), len(x), _p(y))
    return y


def jl_range(start, step, stop):
    r = JlRange()
    lib().orc_jl_range(start, step, stop, C.byref(r))
    return r


def jl_range_values(start, step, stop):
    r = jl_range(start, step, stop)
    return np.array([lib().orc_jl_range_at(C.byref(r), i) for i in range(r.len)])


def spline_opt_legacy(X, Y):
    X, Y = f64(X), f64(Y)
    x, y = C.c_double(), C.c_double()
    if lib().orc_spline_opt_legacy(_p(X), _p(Y), len(X), C.byref(x), C.byref(y)):
        raise ValueError("orc_spline_opt_legacy failed")
    return x.value, y.value


def spline_root_legacy(X, Y, value=0.0):
    X, Y = f64(X), f64(Y)
    x = C.c_double()
    if lib().orc_spline_root_legacy(_p(X), _p(Y), len(X), value, C.byref(x)):
        raise ValueError("orc_spline_root_legacy failed")
    return x.value


def chi2_search_legacy(f, res2min, chi2fact):
    mu, r2 = C.c_double(), C.c_double()
    rc = lib().orc_chi2_search_legacy(FN1(lambda m, _c: f(m)), None, res2min, chi2fact, C.byref(mu), C.byref(r2))
    if rc:
        raise ValueError("doubling did not terminate")
    return mu.value, r2.value


# ----------------------------------------------------------------------------- options
def make_t2map_opts(shape, nTE, nT2, TE, T2Range=(10e-3, 2.0), Reg="none", T1=1.0, Threshold=0.0,
                    MinRefAngle=50.0, nRefAngles=64, nRefAnglesMin=None, RefConAngle=180.0, Chi2Factor=None,
                    NoiseLevel=None, SetFlipAngle=None, legacy=False, alpha_provided=False, ngpus=0):
    o = abi.T2mapOpts()
    o.nx, o.ny, o.nz = shape
    o.nTE, o.nT2 = nTE, nT2
    o.nRefAngles = nRefAngles
    o.nRefAnglesMin = min(5, nRefAngles) if nRefAnglesMin is None else nRefAnglesMin
    o.reg = abi.REG_CODES[Reg]
    o.legacy = int(legacy)
    o.alpha_provided = int(alpha_provided)
    o.ngpus = ngpus
    o.TE, o.T2min, o.T2max, o.T1 = TE, T2Range[0], T2Range[1], T1
    o.Threshold, o.MinRefAngle, o.RefConAngle = Threshold, MinRefAngle, RefConAngle
    o.Chi2Factor = float("nan") if Chi2Factor is None else Chi2Factor
    o.NoiseLevel = float("nan") if NoiseLevel is None else NoiseLevel
    o.SetFlipAngle = float("nan") if SetFlipAngle is None else SetFlipAngle
    return o


def make_t2part_opts(shape, nT2, T2Range=(10e-3, 2.0), SPWin=(10e-3, 25e-3), MPWin=(25e-3, 200e-3), Sigmoid=None):
    p = abi.T2partOpts()
    p.nx, p.ny, p.nz = shape
    p.nT2 = nT2
    p.T2min, p.T2max = T2Range
    p.SPWin_lo, p.SPWin_hi = SPWin
    p.MPWin_lo, p.MPWin_hi = MPWin
    p.Sigmoid = float("nan") if Sigmoid is None else Sigmoid
    return p


MAP_NAMES = ["gdn", "ggm", "gva", "fnr", "snr", "alpha"]
PART_NAMES = ["sfr", "sgm", "mfr", "mgm"]


def alloc_outputs(nvox, nTE, nT2, part=True, save_reg=True, save_resnorm=True, save_curve=False, save_basis=False,
                  alpha_init=None):
    """NaN-filled outputs like T2Maps(opts) / T2Distributions(opts) (src/T2mapSEcorr.jl:36-52, 68-72)."""
    arrs = {k: np.full(nvox, np.nan) for k in MAP_NAMES}
    if alpha_init is not None:
        arrs["alpha"][:] = alpha_init
    arrs["dist"] = np.full(nvox * nT2, np.nan)
    if save_resnorm:
        arrs["resnorm"] = np.full(nvox, np.nan)
    if save_curve:
        arrs["decaycurve"] = np.full(nvox * nTE, np.nan)
    if save_reg:
        arrs["mu"] = np.full(nvox, np.nan)
        arrs["chi2factor"] = np.full(nvox, np.nan)
    if save_basis:
        arrs["decaybasis"] = np.full(nvox * nTE * nT2, np.nan)
    if part:
        for k in PART_NAMES:
            arrs[k] = np.full(nvox, np.nan)
    out = abi.T2mapOut()
    for k in abi.OUT_FIELDS:
        setattr(out, k, arrs[k].ctypes.data if k in arrs else None)
    return arrs, out


_variants = {}


def lib_variant(name):
    """Another build of the same oracle sources (oracle/Makefile): "simd" (liborc_simd.so: the reference's @simd
    reductions vectorised / reassociated) or "native" (the same with -march=native, built on this machine if gcc is
    here; falls back to "simd").  Only orc_t2map is declared: these builds are timed / compared, never the checker."""
    if name in _variants:
        return _variants[name]
    build()
    path = os.path.join(ORACLE_DIR, "_build", f"liborc_{name}.so")
    if name == "native":
        try:
            subprocess.run(["make", "-C", ORACLE_DIR, "native"], check=True, stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL)
        except Exception:
            pass
        if not os.path.exists(path):
            return lib_variant("simd")
    if not os.path.exists(path):
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)
    L = C.CDLL(path)
    L.orc_t2map.argtypes = lib().orc_t2map.argtypes
    L._decaes_variant = os.path.basename(path)
    _variants[name] = L
    return L


def t2map(image, opts, part=None, nthreads=0, L=None, **alloc_kw):
    """image: (nvox, nTE) array-like in Julia memory order, i.e. image.T.ravel() is [echo][voxel]."""
    img = np.asfortranarray(image, dtype=np.float64)  # (nvox, nTE) column-major -> v + e*nvox
    nvox, nTE = img.shape
    arrs, out = alloc_outputs(nvox, nTE, opts.nT2, part=part is not None, **alloc_kw)
    st = Stats()
    rc = (L or lib()).orc_t2map(img.ctypes.data_as(dp), nvox, nvox, C.byref(opts), C.byref(part) if part is not None else None,
                                C.byref(out), nthreads, C.byref(st))
    if rc != 0:
        raise ValueError(f"orc_t2map failed with status {rc}")
    arrs["dist"] = arrs["dist"].reshape(opts.nT2, nvox).T
    if "decaycurve" in arrs:
        arrs["decaycurve"] = arrs["decaycurve"].reshape(nTE, nvox).T
    return arrs, st


def t2part(dist, part):
    d = np.asfortranarray(dist, dtype=np.float64)
    nvox, nT2 = d.shape
    outs = [np.full(nvox, np.nan) for _ in range(4)]
    rc = lib().orc_t2part(d.ctypes.data_as(dp), nvox, nvox, C.byref(part), *[_p(o) for o in outs])
    if rc != 0:
        raise ValueError(f"orc_t2part failed with status {rc}")
    return dict(zip(PART_NAMES, outs))


def setup_tables(opts):
    nA = 1 if not np.isnan(opts.SetFlipAngle) else opts.nRefAngles
    et, t2, ang = np.empty(opts.nTE), np.empty(opts.nT2), np.empty(nA)
    basis = np.empty(nA * opts.nT2 * opts.nTE)
    dbasis = np.empty_like(basis)
    lib().orc_setup_tables(C.byref(opts), _p(et), _p(t2), _p(ang), _p(basis), _p(dbasis))
    shp = (nA, opts.nT2, opts.nTE)
    return et, t2, ang, basis.reshape(shp).transpose(2, 1, 0), dbasis.reshape(shp).transpose(2, 1, 0)


def mock_image(nvox, nTE, TE, T1=1.0, SNR=60.0, seed=1, first_voxel=0):
    img = np.empty((nTE, nvox))
    lib().orc_mock_image(_p(img), nvox, nvox, first_voxel, nTE, C.c_double(TE), C.c_double(T1), C.c_double(SNR),
                         C.c_uint64(seed))
    return np.asfortranarray(img.T)  # (nvox, nTE), column-major
