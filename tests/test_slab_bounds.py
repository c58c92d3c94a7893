"""Slab arithmetic of the multi-GPU host path (CPU only: pure host functions of the C ABI)."""
import numpy as np


def test_masked_cuts_balance_the_foreground(pkg):
    rng = np.random.default_rng(0)
    nx, ny, nz = 96, 96, 60
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    # ellipsoidal "brain": ~35 % of the volume, nothing in the first and last slices
    inside = ((x - nx / 2) / (0.42 * nx)) ** 2 + ((y - ny / 2) / (0.46 * ny)) ** 2 + ((z - nz / 2) / (0.45 * nz)) ** 2 < 1
    first_echo = np.where(inside, 100.0 + rng.random(inside.shape), 0.0).ravel()  # x fastest, z slowest: Julia's order
    nvox = first_echo.size
    for ng in (2, 4, 8):
        cuts = pkg.slab_bounds_masked(first_echo, 0.0, ng)
        assert cuts[0] == 0 and cuts[-1] == nvox and all(b >= a for a, b in zip(cuts, cuts[1:]))
        assert all(c % 4 == 0 for c in cuts[:-1])
        fg = [int((first_echo[a:b] > 0).sum()) for a, b in zip(cuts, cuts[1:])]
        assert sum(fg) == int(inside.sum())
        assert max(fg) - min(fg) <= 2 * 1024, fg          # equal to within a block
        naive = [int((first_echo[nvox * d // ng: nvox * (d + 1) // ng] > 0).sum()) for d in range(ng)]
        assert max(fg) < 0.8 * max(naive) or ng == 2, (fg, naive)  # the naive split leaves the outer slabs idle
    # an unmasked volume: equal lengths (to within a block); a fully masked one: still a valid partition
    cuts = pkg.slab_bounds_masked(np.ones(100_000), 0.0, 4)
    assert all(abs((b - a) - 25_000) <= 1024 for a, b in zip(cuts, cuts[1:]))
    cuts = pkg.slab_bounds_masked(np.zeros(10_000), 0.0, 3)
    assert cuts[0] == 0 and cuts[-1] == 10_000 and all(b >= a for a, b in zip(cuts, cuts[1:]))


def test_plain_slab_bounds_partition(pkg):
    for nvox in (0, 1, 3, 4099, 6_508_800):
        for ng in (1, 2, 3, 8):
            b = [pkg.slab_bounds(nvox, ng, d) for d in range(ng)]
            assert b[0][0] == 0 and b[-1][1] == nvox
            assert all(b[d][1] == b[d + 1][0] for d in range(ng - 1))
            assert all(lo % 4 == 0 for lo, _ in b)
