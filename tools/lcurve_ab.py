#!/usr/bin/env python
"""L-curve parity A/B: on the SAME >= 65,536 seeded voxels of cfg2 / cfg3

  oracle            vs  oracle on the image perturbed by one ulp   -> the reference algorithm's own noise floor
  oracle            vs  oracle built with vectorised @simd reductions (liborc_simd.so): two faithful CPU builds
  oracle            vs  GPU, Tikhonov solves unrefined  (DECAES_REFINE=0)
  oracle            vs  GPU, Tikhonov solves refined    (DECAES_REFINE=1)
  oracle            vs  GPU, QR port                    (DECAES_SOLVER=qr; follows the reference's pivoting path)

For every pair: mu-flip rate, |dlog mu| of the flips, and for the flipped voxels the distribution of the
differences that matter downstream (|dMWF|, |dggm| relative, max rel ddist): median / 95th percentile / max.
Writes one JSON document (stdout + --out).  Run on a GPU box:  python tools/lcurve_ab.py --out gpurun_out/x.json
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc  # noqa: E402
import parity  # noqa: E402

pkg = orc._load_package()

CFG = {"cfg2": (48, 8e-3, 40, 2), "cfg3": (56, 7e-3, 40, 3)}


def gpu_run(img, o, p, env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        nvox, nTE = img.shape
        arrs, out = orc.alloc_outputs(nvox, nTE, o.nT2, part=True)
        L = pkg.lib()
        t0 = time.perf_counter()
        rc = L.decaes_t2map(img.ctypes.data, C.byref(o), C.byref(p), C.byref(out))
        dt = time.perf_counter() - t0
        assert rc == 0, L.decaes_last_error().decode()
        st = pkg.last_stats()
        arrs["dist"] = arrs["dist"].reshape(o.nT2, nvox).T
        return arrs, st, dt
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def pct(a):
    a = np.asarray(a, dtype=float)
    a = a[np.isfinite(a)]
    if a.size == 0:
        return None
    return {"median": float(np.median(a)), "p95": float(np.percentile(a, 95)), "max": float(a.max())}


def flip_report(ref, got):
    rep = parity.compare(ref, got)
    m0, m1 = ref["mu"], got["mu"]
    with np.errstate(invalid="ignore", divide="ignore"):
        dlog = np.abs(np.log(m0) - np.log(m1))
    flip = ~((m0 == m1) | (dlog <= 1e-7))
    d0, d1 = ref["dist"], got["dist"]
    scale = np.maximum(np.abs(d0), np.abs(d1)).max(1, keepdims=True)
    with np.errstate(invalid="ignore", divide="ignore"):
        # per-voxel max difference of the distribution relative to its largest bin
        ddist = (np.abs(d0 - d1) / np.maximum(scale, 1e-300)).max(1)
        dggm = np.abs(ref["ggm"] - got["ggm"]) / np.abs(ref["ggm"])
    out = {k: rep[k] for k in ("nvox", "nan_mismatch", "support_diff", "voxels_out_of_tolerance", "mu_flips", "mu_flip_frac",
                               "mu_flip_median_dlog", "mu_flip_max_dlog", "out_of_tolerance_same_mu", "dist_max_rel_same_support")}
    out["flipped"] = {"abs_dMWF": pct(np.abs(ref["sfr"] - got["sfr"])[flip]), "rel_dggm": pct(dggm[flip]),
                      "max_rel_ddist": pct(ddist[flip]), "abs_dlog_mu": pct(dlog[flip]),
                      "abs_dalpha": pct(np.abs(ref["alpha"] - got["alpha"])[flip])}
    out["same_mu"] = {"abs_dMWF": pct(np.abs(ref["sfr"] - got["sfr"])[~flip]), "max_rel_ddist": pct(ddist[~flip])}
    return out, flip


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--voxels", type=int, default=65536)
    ap.add_argument("--configs", default="cfg2,cfg3")
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-qr", action="store_true")
    args = ap.parse_args()
    doc = {"voxels": args.voxels, "configs": {}}
    for name in args.configs.split(","):
        nTE, TE, nT2, seed = CFG[name]
        nvox = args.voxels
        img = orc.mock_image(nvox, nTE, TE, seed=seed)
        o = orc.make_t2map_opts((nvox, 1, 1), nTE, nT2, TE, Reg="lcurve", ngpus=1)
        p = orc.make_t2part_opts((nvox, 1, 1), nT2)
        t0 = time.perf_counter()
        ref, st = orc.t2map(img, o, p)
        t_orc = time.perf_counter() - t0
        ulp, _ = orc.t2map(np.asfortranarray(np.nextafter(img, np.inf)), o, p)
        rng = np.random.default_rng(0)
        ulp_r, _ = orc.t2map(np.asfortranarray(img * (1 + 1e-16 * rng.standard_normal(img.shape))), o, p)
        res = {"oracle_seconds": t_orc, "oracle_threads": int(st.threads)}
        # the same C sources with the reference's @simd reductions vectorised (reassociated), as LLVM does for Julia
        t0 = time.perf_counter()
        simd, _ = orc.t2map(img, o, p, L=orc.lib_variant("simd"))
        res["oracle_simd_seconds"] = time.perf_counter() - t0
        res["oracle_vs_oracle_simd_build"], f_simd = flip_report(ref, simd)
        res["oracle_vs_oracle_1ulp_up"], f_up = flip_report(ref, ulp)
        res["oracle_vs_oracle_1ulp_random"], f_r = flip_report(ref, ulp_r)
        flips = {}
        variants = [("gpu_unrefined", {"DECAES_REFINE": "0"}), ("gpu_refined", {"DECAES_REFINE": "1"})]
        if not args.no_qr:
            variants.append(("gpu_qr", {"DECAES_SOLVER": "qr"}))
        for tag, env in variants:
            gpu_run(img, o, p, env)  # warm-up (workspaces, clocks)
            got, gst, dt = gpu_run(img, o, p, env)
            res[tag], flips[tag] = flip_report(ref, got)
            res[tag]["pipeline_ms"] = gst["pipeline_ms"]
            res[tag]["voxels_per_s_kernel"] = nvox / (gst["pipeline_ms"] * 1e-3)
            for k in ("early_returns", "lcurve_overflow", "nnls_itercap"):
                if k in gst:
                    res[tag][k] = gst[k]
        # are the GPU flips the same voxels the oracle itself is unsure about?
        unsure = f_up | f_r | f_simd
        for tag, f in flips.items():
            res[tag]["flips_also_oracle_unsure"] = int((f & unsure).sum())
        if "gpu_unrefined" in flips and "gpu_refined" in flips:
            res["flips_common_unrefined_refined"] = int((flips["gpu_unrefined"] & flips["gpu_refined"]).sum())
        doc["configs"][name] = res
        print(name, json.dumps(res, indent=1), flush=True)
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(doc, fh, indent=1)


if __name__ == "__main__":
    main()
