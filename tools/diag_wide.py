"""Diagnostic: voxels where the Gram solver misses the tolerance on the wide-parity data families (run on a GPU box)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import orc, parity
import test_gpu_parity_wide as w

pkg = orc._load_package()
nvox, nTE, TE = 2048, 48, 8e-3
for name, pools, SNR, nT2 in w.CASES:
    if len(sys.argv) > 1 and name not in sys.argv[1:]:
        continue
    img = w.synth(orc, nvox, nTE, TE, pools, SNR, seed=sum(map(ord, name)))
    p = orc.make_t2part_opts((nvox, 1, 1), nT2)
    T2 = orc.logrange(10e-3, 2.0, nT2)
    for Reg, extra in (("none", {}), ("chi2", {"Chi2Factor": 1.02})):
        o = orc.make_t2map_opts((nvox, 1, 1), nTE, nT2, TE, Reg=Reg, ngpus=1, **extra)
        ref, _ = orc.t2map(img, o, p)
        got = w.gpu_t2map(pkg, orc, img, o, p)
        qr = w.gpu_t2map(pkg, orc, img, o, p, env={"DECAES_SOLVER": "qr"})
        rep = parity.compare(ref, got)
        repq = parity.compare(ref, qr)
        print(f"== {name} {Reg}: gram oot {rep['voxels_out_of_tolerance']} support {rep['support_diff']} | qr oot {repq['voxels_out_of_tolerance']} support {repq['support_diff']}")
        d0, d1 = ref["dist"], got["dist"]
        scale = np.maximum(np.abs(d0), np.abs(d1))
        bad = ~(np.abs(d0 - d1) <= np.maximum(1e-9 * np.maximum(1.0, scale.max(1, keepdims=True)), 1e-6 * scale)).all(1)
        for v in np.where(bad)[0][:6]:
            s0, s1 = np.where(d0[v] > 0)[0], np.where(d1[v] > 0)[0]
            A = np.stack([orc.epg(nTE, ref["alpha"][v], TE, t, 1.0) for t in T2], 1)
            b = img[v]
            r0, r1 = np.linalg.norm(A @ d0[v] - b), np.linalg.norm(A @ d1[v] - b)
            print(f" voxel {v}: alpha d {abs(ref['alpha'][v]-got['alpha'][v]):.1e} support ref {s0.tolist()} gpu {s1.tolist()} cond(A_P ref) {np.linalg.cond(A[:, s0]):.2e}"
                  f" resid ref {r0:.15e} gpu {r1:.15e}")
            print("   x ref", d0[v][s0], "\n   x gpu", d1[v][s1])
            # KKT of both answers
            for tag, x in (("ref", d0[v]), ("gpu", d1[v])):
                wd = A.T @ (b - A @ x)
                print(f"   {tag}: max dual off-support {wd[x == 0].max():.3e}, max |dual| on support {np.abs(wd[x > 0]).max():.3e}")
