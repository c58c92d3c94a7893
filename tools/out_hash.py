#!/usr/bin/env python
"""SHA-256 of every output map of one run of decaes_t2map on a seeded synthetic volume (bit-equality checks between builds:
   DECAES_LIB=... python tools/out_hash.py [nvox] [Reg] [nTE] [nT2])."""
import ctypes as C, hashlib, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import orc

pkg = orc._load_package()
nvox = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
Reg = sys.argv[2] if len(sys.argv) > 2 else "lcurve"
nTE = int(sys.argv[3]) if len(sys.argv) > 3 else 56
nT2 = int(sys.argv[4]) if len(sys.argv) > 4 else 40
TE = {32: 10e-3, 48: 8e-3, 56: 7e-3}.get(nTE, 8e-3)
extra = {"chi2": {"Chi2Factor": 1.02}, "mdp": {"NoiseLevel": 1e-3}}.get(Reg, {})
img = np.asfortranarray(orc.mock_image(nvox, nTE, TE, seed=3), dtype=np.float64)
o = orc.make_t2map_opts((nvox, 1, 1), nTE, nT2, TE, Reg=Reg, ngpus=1, **extra)
p = orc.make_t2part_opts((nvox, 1, 1), nT2)
arrs, out = orc.alloc_outputs(nvox, nTE, nT2, part=True)
rc = pkg.lib().decaes_t2map(img.ctypes.data, C.byref(o), C.byref(p), C.byref(out))
assert rc == 0, pkg.lib().decaes_last_error().decode()
h = hashlib.sha256()
for k in sorted(arrs):
    h.update(np.ascontiguousarray(arrs[k]).tobytes())
print(os.environ.get("DECAES_LIB", "default"), nvox, Reg, nTE, nT2, h.hexdigest()[:16])
