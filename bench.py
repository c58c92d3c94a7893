#!/usr/bin/env python
"""bench.py — voxels/s of T2mapSEcorr + T2partSEcorr (56-echo 240x240x113, Reg=lcurve) on N B200s.

    python bench.py --gpus N --steps K --warmup W          (N > 1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference ...                   CPU arm: the oracle port on all host cores

A step = one pass of the hot path over ONE full synthetic volume (6,508,800 voxels for cfg3).
  N = 1   `value`: the volume resident in HBM, decaes_t2map_device (kernels only inside the timed region).
  N > 1   `value`: the SAME volume sharded as contiguous voxel slabs, rank r owns slab r on its GPU (strong scaling,
          no collective on the data path; barrier + max over ranks).  `replicas_weak` keeps the one-volume-per-rank
          figure of round 1 as an extra.
  `e2e`   one blocking decaes_t2map call (the reference-facing C ABI) on the whole volume in PINNED host memory,
          driving all N GPUs from rank 0 (ngpus = N): H2D + kernels + D2H inside the timed region.
  `e2e_pageable`  the same call on plain pageable numpy buffers (what a Julia caller hands over): the library stages
          them through its pinned ring.  At N > 1 the N-GPU outputs are compared byte for byte with a 1-GPU run.
  `parity`  a sample of the workload checked against the CPU oracle in this very run (north_star tolerances).
PyTorch is used for device memory, the barrier / max-over-ranks reduction and CUDA events only.
"""
import argparse
import ctypes as C
import importlib.util
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (shape, nTE, TE, nT2, Reg, extra opts, part windows)
    "cfg3": ((240, 240, 113), 56, 7e-3, 40, "lcurve", {}, {}),
    "cfg2": ((240, 240, 48), 48, 8e-3, 40, "lcurve", {}, {}),
    "cfg1": ((64, 64, 16), 32, 10e-3, 40, "none", {}, {}),
    "cfg4": ((240, 240, 48), 48, 8e-3, 60, "chi2", {"Chi2Factor": 1.02}, {}),
    "cfg4gcv": ((240, 240, 48), 48, 8e-3, 60, "gcv", {}, {}),
    "cfg5": ((256, 256, 160), 32, 10e-3, 60, "mdp", {"NoiseLevel": 1e-3}, {"SPWin": (10e-3, 200e-3), "MPWin": (200e-3, 2.0)}),
}
METRIC = "voxels/s T2map+T2part (56-echo 240x240x113, lcurve) at 1/2/4/8 B200 vs CPU"
NAMES = ["gdn", "ggm", "gva", "fnr", "snr", "alpha", "sfr", "sgm", "mfr", "mgm"]


def load_package():
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


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_name(wl_name):
    """config.workload — the same string on both arms (ours and --impl reference)."""
    shape, nTE, TE, nT2, Reg, extra, pwin = WORKLOADS[wl_name]
    return f"{wl_name}: {nTE}-echo {'x'.join(map(str, shape))}, nT2={nT2}, Reg={Reg} + T2part"


def config_dict(args, world):
    """config of the JSON line — identical on both arms for the same command line."""
    shape = WORKLOADS[args.workload][0]
    nvox = args.voxels if args.voxels > 0 else shape[0] * shape[1] * shape[2]
    return {"workload": workload_name(args.workload),
            "sharding": "ONE volume; rank r owns the contiguous voxel slab r (decaes_slab_bounds), no collective; T2part fused into the same kernel",
            "voxels": nvox, "voxels_per_rank": nvox // world, "l2": "inputs+outputs per step exceed L2 (no flush needed)",
            "background_fraction": args.mask, "debug_voxels_override": bool(args.voxels) or args.identical >= 0}


def oracle_opts(orc, wl, nvox, **kw):
    shape, nTE, TE, nT2, Reg, extra, pwin = wl
    o = orc.make_t2map_opts((nvox, 1, 1), nTE, nT2, TE, Reg=Reg, **extra, **kw)
    p = orc.make_t2part_opts((nvox, 1, 1), nT2, **pwin)
    return o, p


def cpu_build(orc):
    """The CPU build that is TIMED: the oracle sources with the reference's @simd reductions vectorised, -O3, built
    with -march=native on this machine when gcc is here (oracle/Makefile `native`), else the shipped x86-64-v3 one.
    (The strict sequential build, liborc.so, is the parity checker; it is ~2.5x slower and is not the baseline.)"""
    L = orc.lib_variant("native")
    name = getattr(L, "_decaes_variant", "liborc_native.so")
    flags = "-O3 -fopenmp -DORC_SIMD -ffp-contract=off " + ("-march=native" if "native" in name else "-march=x86-64-v3")
    return L, name, flags


def oracle_sample(orc, wl, nvox, seed, threads, L=None):
    """Run the CPU oracle on `nvox` voxels of the workload; returns (voxels/s, flops/voxel, stats)."""
    img = orc.mock_image(nvox, wl[1], wl[2], seed=seed)
    o, p = oracle_opts(orc, wl, nvox)
    _, st = orc.t2map(img, o, p, nthreads=threads, L=L, save_reg=False, save_resnorm=False)
    return nvox / st.seconds, st.flops / max(st.voxels_processed, 1), st


def run_reference(args, wl_name, out):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  DECAES.jl
    is Julia and cannot run in this image, so this is the oracle port (kind = "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import orc
    wl = WORKLOADS[wl_name]
    threads = os.cpu_count() or 1
    sample = args.ref_sample
    L, libname, flags = cpu_build(orc)
    for _ in range(args.warmup):
        oracle_sample(orc, wl, sample, 1, threads, L)
    t0 = time.perf_counter()
    tot = 0.0
    for k in range(args.steps):
        vps, fpv, st = oracle_sample(orc, wl, sample, 1 + k, threads, L)
        tot += st.seconds
    wall = time.perf_counter() - t0
    value = sample * args.steps / tot
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": threads, "kind": "port",
                         "sample": f"each step is a bounded sample of {sample} voxels of the same synthetic workload on all host cores; C restatement of DECAES.jl "
                                   f"(no Julia runtime in the image), OpenMP over voxels, {libname}, gcc {flags}"},
        "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": wall,
    }
    print(json.dumps(line), file=out, flush=True)


def claim_stdout():
    """The driver reads ONE JSON line from stdout.  Libraries write there too (NCCL prints its version banner on
    init), so everything else is routed to stderr at the file-descriptor level and the JSON line goes to the
    original stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def parity_block(pkg, orc, wl, nvox, seed):
    """GPU vs CPU oracle on `nvox` seeded voxels of the workload, through the host-pointer C ABI, north_star tolerances.
    mu_flips are judged against the rate at which two CPU builds of the oracle disagree on the same voxels."""
    import parity
    img = orc.mock_image(nvox, wl[1], wl[2], seed=seed)
    o, p = oracle_opts(orc, wl, nvox, ngpus=1)
    ref, ost = orc.t2map(img, o, p)
    alt, _ = orc.t2map(img, o, p, L=orc.lib_variant("simd"))
    arrs, out = orc.alloc_outputs(nvox, wl[1], wl[3], part=True)
    rc = pkg.lib().decaes_t2map(img.ctypes.data, C.byref(o), C.byref(p), C.byref(out))
    if rc != 0:
        raise RuntimeError(pkg.lib().decaes_last_error().decode())
    st = pkg.last_stats()
    arrs["dist"] = arrs["dist"].reshape(wl[3], nvox).T
    rep, own = parity.compare(ref, arrs), parity.compare(ref, alt)
    return {"voxels_compared": nvox, "against": "CPU oracle (oracle/liborc.so), same seeded voxels, through decaes_t2map",
            "tolerances": "dist rel 1e-6 / abs 1e-9; alpha, MWF, gmT2 abs 1e-6",
            "out_of_tolerance": rep["voxels_out_of_tolerance"], "out_of_tolerance_same_mu": rep["out_of_tolerance_same_mu"],
            "mu_flips": rep["mu_flips"], "mu_flip_frac": rep["mu_flip_frac"], "mu_flip_median_dlog": rep["mu_flip_median_dlog"],
            "mu_flips_between_two_cpu_builds": own["mu_flips"], "support_diff": rep["support_diff"], "support_diff_same_mu": rep["support_diff_same_mu"],
            "alpha_max_abs": rep["alpha_max_abs"], "sfr_max_abs": rep["sfr_max_abs"],
            "early_returns": st["early_returns"], "early_returns_oracle": int(ost.early_returns),
            "lcurve_overflow": st["lcurve_overflow"], "nnls_itercap": st["nnls_itercap"]}


def main():
    out = claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--voxels", type=int, default=0, help="debug: override voxels of the volume (invalidates the headline)")
    ap.add_argument("--ref-sample", type=int, default=65536)
    ap.add_argument("--cpu-sample", type=int, default=131072)
    ap.add_argument("--parity-sample", type=int, default=8192)
    ap.add_argument("--mask", type=float, default=0.0, help="fraction of background voxels (first echo zeroed, ellipsoidal mask)")
    ap.add_argument("--identical", type=int, default=-1,
                    help="debug: every voxel is a copy of voxel K (all warps follow the same control flow; invalidates the headline)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-replicas", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args, args.workload, out)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    pkg = load_package()
    pkg.lib()  # fail loudly if the CUDA library is missing — there is no fallback

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    wl = WORKLOADS[args.workload]
    shape, nTE, TE, nT2, Reg, extra, pwin = wl
    nvox = args.voxels if args.voxels > 0 else shape[0] * shape[1] * shape[2]
    stream = torch.cuda.current_stream().cuda_stream

    def options(n, ngpus=1):
        o = pkg.T2mapOptions(MatrixSize=(n, 1, 1), nTE=nTE, TE=TE, nT2=nT2, T2Range=(10e-3, 2.0), Reg=Reg, ngpus=ngpus,
                             Silent=True, **extra).to_c()
        p = pkg.T2partOptions(MatrixSize=(n, 1, 1), nT2=nT2, T2Range=(10e-3, 2.0), SPWin=pwin.get("SPWin", (10e-3, 25e-3)),
                              MPWin=pwin.get("MPWin", (25e-3, 200e-3)), Silent=True).to_c()
        return o, p

    def mask_volume(img_t, v0, n):
        """Zero the first echo outside an ellipsoid (brain-mask stand-in; src/T2mapSEcorr.jl:177 skips those voxels)."""
        if args.mask <= 0:
            return
        r = (1.0 - args.mask) ** (1.0 / 3.0) * (6.0 / math.pi) ** (1.0 / 3.0) / 2.0  # ellipsoid semi-axis / box side for that volume fraction
        v = torch.arange(v0, v0 + n, device=img_t.device)
        x = (v % shape[0]).double() / shape[0] - 0.5
        y = ((v // shape[0]) % shape[1]).double() / shape[1] - 0.5
        z = (v // (shape[0] * shape[1])).double() / shape[2] - 0.5
        img_t[0][(x * x + y * y + z * z) > r * r] = 0.0

    def alloc_dev(n):
        t = {k: torch.empty((n,), dtype=torch.float64, device=dev) for k in NAMES}
        t["dist"] = torch.empty((nT2, n), dtype=torch.float64, device=dev)
        return t

    # ---- `value`: the volume resident in HBM; rank r owns slab r (strong scaling; N = 1: the whole volume) ----
    v0, v1 = pkg.slab_bounds(nvox, world, rank)
    nloc = v1 - v0
    img = torch.empty((nTE, nloc), dtype=torch.float64, device=dev)
    pkg.mock_image_device(img.data_ptr(), nloc, nloc, v0, nTE, TE, seed=3, stream=stream)
    mask_volume(img, v0, nloc)
    if args.identical >= 0:
        img[:] = img[:, args.identical:args.identical + 1].clone()
    outs = alloc_dev(nloc)
    out_struct = pkg.make_out({k: v.data_ptr() for k, v in outs.items()})
    o_loc, p_loc = options(nloc)
    in_bytes = nvox * nTE * 8
    out_bytes = nvox * (nT2 + len(NAMES)) * 8
    # inputs + outputs are ~5.5 GB per volume for cfg3: far larger than the 126 MB L2, no flush needed

    def step_device():
        pkg.t2map_device(img.data_ptr(), nloc, nloc, o_loc, p_loc, out_struct, stream)

    kernel_ms = []
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
        st = pkg.last_stats()  # resolves the library's own CUDA events (kernel-only duration)
        kernel_ms.append(st["pipeline_ms"])
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    elapsed_ms = max_over_ranks(e0.elapsed_time(e1))
    value = nvox * args.steps / (elapsed_ms * 1e-3)
    st_dev = pkg.last_stats()
    processed = torch.tensor([float(st_dev["voxels_processed"])], dtype=torch.float64, device=dev)
    chk = outs["gdn"].nan_to_num().sum().reshape(1).clone()
    if world > 1:
        dist.all_reduce(processed)
        dist.all_reduce(chk)
    processed, checksum = int(processed.item()), float(chk.item())
    mean_kernel_s = max_over_ranks(1e-3 * sum(kernel_ms) / max(len(kernel_ms), 1))

    # ---- extra at N > 1: one volume PER rank (round-1 `value`, weak scaling) ----
    replicas = None
    if world > 1 and not args.no_replicas:
        rimg = torch.empty((nTE, nvox), dtype=torch.float64, device=dev)
        pkg.mock_image_device(rimg.data_ptr(), nvox, nvox, rank * nvox, nTE, TE, seed=3, stream=stream)
        routs = alloc_dev(nvox)
        rstruct = pkg.make_out({k: v.data_ptr() for k, v in routs.items()})
        o_r, p_r = options(nvox)
        pkg.t2map_device(rimg.data_ptr(), nvox, nvox, o_r, p_r, rstruct, stream)
        barrier()
        e0.record()
        for _ in range(2):
            pkg.t2map_device(rimg.data_ptr(), nvox, nvox, o_r, p_r, rstruct, stream)
        e1.record()
        barrier()
        rms = max_over_ranks(e0.elapsed_time(e1))
        replicas = {"value": nvox * world * 2 / (rms * 1e-3), "unit": "voxels/s", "scaling": "weak", "steps": 2,
                    "what": "one full volume per rank, device resident (the `value` of round 1)"}
        del rimg, routs

    # ---- end to end through the host-pointer C-ABI call: ONE call on the whole volume drives all N GPUs ----
    e2e = e2e_pageable = equal_1gpu = None
    if not args.no_e2e:
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier(group=cpu_group)
        if rank == 0:
            o_all, p_all = options(nvox, ngpus=world)
            h_img = torch.empty((nTE, nvox), dtype=torch.float64).pin_memory()
            if world == 1:
                h_img.copy_(img)
            else:  # the same bytes as the ranks' slabs: generated on this device in pieces
                piece = 1 << 20
                tmp = torch.empty((nTE, piece), dtype=torch.float64, device=dev)
                for a in range(0, nvox, piece):
                    n = min(piece, nvox - a)
                    pkg.mock_image_device(tmp.data_ptr(), n, piece, a, nTE, TE, seed=3, stream=stream)
                    mask_volume(tmp[:, :n], a, n)
                    h_img[:, a:a + n].copy_(tmp[:, :n])
                del tmp
            h_outs = {k: torch.empty((nvox,), dtype=torch.float64).pin_memory() for k in NAMES}
            h_outs["dist"] = torch.empty((nT2, nvox), dtype=torch.float64).pin_memory()
            h_struct = pkg.make_out({k: v.data_ptr() for k, v in h_outs.items()})
            e2e_steps = max(1, min(args.steps, 2))

            def timed(call, steps):
                call()  # warm-up (allocates / grows the cached workspaces)
                t0 = time.perf_counter()
                for _ in range(steps):
                    call()
                return (time.perf_counter() - t0) / steps

            def step_host(o=o_all, image_ptr=h_img.data_ptr(), struct=h_struct):
                rc = pkg.lib().decaes_t2map(image_ptr, C.byref(o), C.byref(p_all), C.byref(struct))
                if rc != 0:
                    raise RuntimeError(pkg.lib().decaes_last_error().decode())
            dt = timed(step_host, e2e_steps)
            hst = pkg.last_stats()
            assert hst["pinned_staging"] == 0 and hst["ngpus_used"] == world, hst
            hsum = float(h_outs["gdn"].nan_to_num().sum().item())
            assert abs(hsum - checksum) <= 1e-9 * abs(checksum) + 1e-9, (hsum, checksum)
            e2e = {"value": nvox / dt, "unit": "voxels/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                   "steps": e2e_steps, "seconds_per_volume": dt, "host_buffers": "pinned", "ngpus": hst["ngpus_used"],
                   "host_stats": hst}
            # pageable buffers: plain numpy arrays, as a Julia caller's Arrays are
            n_img = h_img.numpy().copy()
            n_outs = {k: np.full(v.shape, np.nan) for k, v in h_outs.items()}
            n_struct = pkg.make_out({k: v.ctypes.data for k, v in n_outs.items()})
            dtp = timed(lambda: step_host(o_all, n_img.ctypes.data, n_struct), e2e_steps)
            pst = pkg.last_stats()
            assert pst["pinned_staging"] == 1, pst
            same = all(np.array_equal(n_outs[k], h_outs[k].numpy(), equal_nan=True) for k in n_outs)
            e2e_pageable = {"value": nvox / dtp, "unit": "voxels/s", "seconds_per_volume": dtp, "steps": e2e_steps,
                            "host_buffers": "pageable numpy arrays, staged through the library's pinned ring",
                            "relative_to_pinned": dt / dtp, "bytes_equal_to_pinned_run": bool(same), "host_stats": pst}
            if world > 1:  # the sharded result must be the single-GPU result, byte for byte
                o_one, _ = options(nvox, ngpus=1)
                step_host(o_one, n_img.ctypes.data, n_struct)
                nbytes = sum(v.nbytes for v in n_outs.values())
                eq = all(np.array_equal(n_outs[k], h_outs[k].numpy(), equal_nan=True) for k in n_outs)
                equal_1gpu = {"equal": bool(eq), "bytes_compared": nbytes,
                              "what": f"decaes_t2map(ngpus={world}) vs decaes_t2map(ngpus=1) on the same volume, every output array"}
                assert eq, "N-GPU result differs from the single-GPU result"
            del n_img, n_outs, h_img, h_outs
        if world > 1:
            dist.barrier(group=cpu_group)

    if rank == 0:
        import orc
        # ---- parity sample, CPU baseline + algorithmic FLOPs per voxel from the oracle's instrumented counters ----
        par = parity_block(pkg, orc, wl, args.parity_sample, seed=3) if args.parity_sample > 0 else None
        cpu = None
        flops_per_voxel = None
        if not args.no_cpu:
            threads = os.cpu_count() or 1
            L, libname, flags = cpu_build(orc)
            oracle_sample(orc, wl, 4096, 2, threads, L)
            vps, flops_per_voxel, st = oracle_sample(orc, wl, args.cpu_sample, 3, threads, L)
            cpu = {"value": vps, "unit": "voxels/s", "cores": threads, "kind": "port", "build": f"{libname}: gcc {flags}",
                   "sample": f"{args.cpu_sample} voxels of the same synthetic workload (seed 3); C restatement of DECAES.jl "
                             "(no Julia runtime in the image), OpenMP over voxels, the reference's @simd reductions vectorised"}
        peak = pkg.measure_fp64_peak()
        roofline = None
        # DRAM traffic of the pipeline kernel: measured with ncu (profiles/*_traffic.json, dram__bytes_read.sum +
        # dram__bytes_write.sum of one full-size launch), scaled to this launch's voxel count
        traffic, tsrc = None, None
        for cand in ("r02_traffic.json", "r01_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", cand)
            if os.path.exists(tpath) and args.workload == "cfg3":
                with open(tpath) as fh:
                    tj = json.load(fh)
                traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["voxels"] * (nvox / world)
                tsrc = f"ncu capture committed as profiles/{cand} (bytes per launch, scaled by voxels per launch)"
                break
        if flops_per_voxel:
            achieved = flops_per_voxel * (nvox / world) / mean_kernel_s  # per GPU: one launch = one slab
            roofline = {"bound": "fp64", "achieved": achieved / 1e12, "peak": peak / 1e12, "unit": "TFLOP/s",
                        "frac": achieved / peak, "traffic": traffic, "traffic_source": tsrc,
                        "peak_source": "measured in this run by decaes_measure_fp64_peak (independent DFMA chains on all SMs); MEASURED_PEAKS.json has no FP64 entry",
                        "flops_per_voxel": flops_per_voxel, "flops_definition": "reference-algorithm FLOPs counted by the instrumented oracle (SURVEY 8d); the Gram solver executes fewer",
                        "kernel": "voxel_pipeline_kernel", "kernel_ms": 1e3 * mean_kernel_s, "per": "GPU (one launch = one slab)",
                        "hbm": {"algorithmic_bytes_per_voxel": 8 * (nTE + nT2 + 10),
                                "achieved_GBps": 8 * (nTE + nT2 + 10) * (nvox / world) / mean_kernel_s / 1e9}}
        line = {
            "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, world),
            "clocks": clocks, "e2e": e2e, "e2e_pageable": e2e_pageable, "sharded_equals_single_gpu": equal_1gpu,
            "gpu_launches": 3 * args.steps,  # per rank: basis_setup + gram_setup + voxel_pipeline per step
            "kernel_ms_per_step": 1e3 * mean_kernel_s,
            "roofline": roofline, "cpu_baseline": cpu, "parity": par, "replicas_weak": replicas,
            "voxels_processed_last_step": processed, "checksum_gdn": checksum,
            "counted_last_step_rank0": {k: st_dev[k] for k in ("early_returns", "lcurve_overflow", "nnls_itercap")},
        }
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.barrier(group=cpu_group)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
