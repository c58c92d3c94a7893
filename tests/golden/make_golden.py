#!/usr/bin/env python
"""Generate the golden fixtures tests/golden/*.npz from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).

DECAES.jl itself cannot run in this image (no Julia runtime), so these vectors are NOT reference outputs: they
freeze the oracle — which is pinned to the reference by the known answers and invariants of tests/test_oracle_*.py —
so that (a) any later change of the oracle shows up as a diff of committed numbers (tests/test_golden.py, CPU) and
(b) the GPU path is checked against numbers that do not move with the code under test (tests/test_golden.py, -m gpu).
Each file holds the seeded input image of 64 voxels and the oracle's maps for one benchmark configuration.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orc  # noqa: E402

# name: (nTE, TE, nT2, Reg, t2map extras, t2part extras)
CASES = {
    "cfg1_none": (32, 10e-3, 40, "none", {}, {}),
    "cfg2_lcurve48": (48, 8e-3, 40, "lcurve", {}, {}),
    "cfg3_lcurve56": (56, 7e-3, 40, "lcurve", {}, {}),
    "cfg4_chi2": (48, 8e-3, 60, "chi2", {"Chi2Factor": 1.02}, {}),
    "cfg4_gcv": (48, 8e-3, 60, "gcv", {}, {}),
    "cfg5_mdp": (32, 10e-3, 60, "mdp", {"NoiseLevel": 1e-3}, {"SPWin": (10e-3, 200e-3), "MPWin": (200e-3, 2.0)}),
    "refcon150_chi2": (32, 10e-3, 40, "chi2", {"Chi2Factor": 1.02, "RefConAngle": 150.0}, {}),
    # legacy = true (src/types.jl:20-21, 59-63): 8 refocusing angles, all probed; sampled-spline flip angle and chi2 root
    "legacy_chi2": (32, 10e-3, 40, "chi2", {"Chi2Factor": 1.02, "legacy": True, "nRefAngles": 8, "nRefAnglesMin": 8}, {}),
    "legacy_none": (48, 8e-3, 40, "none", {"legacy": True, "nRefAngles": 8, "nRefAnglesMin": 8}, {}),
}
# seeds of the first seven cases are 100 + their rank in the sorted list of those seven names (kept as generated)
_FIRST = sorted(["cfg1_none", "cfg2_lcurve48", "cfg3_lcurve56", "cfg4_chi2", "cfg4_gcv", "cfg5_mdp", "refcon150_chi2"])
SEEDS = {n: 100 + i for i, n in enumerate(_FIRST)}
SEEDS.update({"legacy_chi2": 121, "legacy_none": 122})
NVOX = 64
KEYS = ["dist", "gdn", "ggm", "gva", "fnr", "snr", "alpha", "mu", "chi2factor", "resnorm", "sfr", "sgm", "mfr", "mgm"]


def compute(name):
    nTE, TE, nT2, Reg, extra, pextra = CASES[name]
    seed = SEEDS[name]
    img = orc.mock_image(NVOX, nTE, TE, seed=seed)
    o = orc.make_t2map_opts((NVOX, 1, 1), nTE, nT2, TE, Reg=Reg, ngpus=1, **extra)
    p = orc.make_t2part_opts((NVOX, 1, 1), nT2, **pextra)
    ref, _ = orc.t2map(img, o, p, nthreads=1)
    return img, {k: np.asarray(ref[k]) for k in KEYS}


if __name__ == "__main__":
    for name in (sys.argv[1:] or CASES):
        img, ref = compute(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), image=img, **ref)
        print(name, "written:", {k: v.shape for k, v in ref.items() if k in ("dist", "alpha")})
