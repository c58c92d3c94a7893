"""N > 1 path on CPU (gloo, world_size 2): the path shards voxel slabs over ranks with no data-path
collective.  Each rank takes its slab from decaes_slab_bounds (the same function the library uses to
split a volume over GPUs), runs the pipeline on it (here: the oracle stands in for the device), and
rank 0 gathers the slabs; the result must equal the single-rank run bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nvox, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    pkg = orc._load_package()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    nTE, nT2, TE = 32, 40, 10e-3
    v0, v1 = pkg.slab_bounds(nvox, world, rank)
    img = orc.mock_image(v1 - v0, nTE, TE, seed=5, first_voxel=v0)   # counter-based RNG keyed by the global voxel id
    o = orc.make_t2map_opts((v1 - v0, 1, 1), nTE, nT2, TE, Reg="chi2", Chi2Factor=1.02)
    p = orc.make_t2part_opts((v1 - v0, 1, 1), nT2)
    m, st = orc.t2map(img, o, p, nthreads=1)
    # host-side gather of disjoint slabs (no reduction of voxel data)
    local = torch.from_numpy(np.concatenate([m["dist"], m["alpha"][:, None], m["sfr"][:, None]], axis=1))
    sizes = [pkg.slab_bounds(nvox, world, r) for r in range(world)]
    nmax = max(b - a for a, b in sizes)  # gloo gathers equal-sized tensors: pad the slabs, trim on rank 0
    padded = torch.zeros((nmax, local.shape[1]), dtype=torch.float64)
    padded[: local.shape[0]] = local
    bufs = [torch.empty_like(padded) for _ in sizes] if rank == 0 else None
    dist.gather(padded, bufs, dst=0)
    if rank == 0:
        bufs = [buf[: b - a] for buf, (a, b) in zip(bufs, sizes)]
    t = torch.tensor([float(st.voxels_processed)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)        # bookkeeping only (bench.py reduces its timing the same way)
    if rank == 0:
        q.put((torch.cat(bufs).numpy(), t.item()))
    dist.destroy_process_group()


def test_slab_bounds_partition(pkg):
    for nvox in [0, 1, 3, 4, 5, 63, 64, 65, 1000, 65536, 6508800]:
        for n in [1, 2, 3, 4, 8]:
            edges = [pkg.slab_bounds(nvox, n, i) for i in range(n)]
            assert edges[0][0] == 0 and edges[-1][1] == nvox
            for (a, b), (c, d) in zip(edges, edges[1:]):
                assert b == c and a <= b
            assert all(a % 4 == 0 for a, _ in edges)          # work groups never straddle shards
            if nvox >= 64 * n:
                sizes = [b - a for a, b in edges]
                assert max(sizes) - min(sizes) <= 8
    with pytest.raises(pkg.DecaesError):
        pkg.slab_bounds(10, 2, 2)


def test_two_ranks_equal_one_rank(orc):
    import torch.multiprocessing as mp
    nvox, world = 203, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nvox, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    gathered, nproc = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert nproc == nvox
    img = orc.mock_image(nvox, 32, 10e-3, seed=5)
    o = orc.make_t2map_opts((nvox, 1, 1), 32, 40, 10e-3, Reg="chi2", Chi2Factor=1.02)
    p = orc.make_t2part_opts((nvox, 1, 1), 40)
    m, _ = orc.t2map(img, o, p, nthreads=1)
    ref = np.concatenate([m["dist"], m["alpha"][:, None], m["sfr"][:, None]], axis=1)
    np.testing.assert_array_equal(gathered, ref)
