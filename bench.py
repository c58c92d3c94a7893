#!/usr/bin/env python
"""bench.py — voxels/s of T2mapSEcorr + T2partSEcorr (56-echo 240x240x113, Reg=lcurve) on N B200s.

    python bench.py --gpus N --steps K --warmup W          (N > 1: launched under torchrun)
    python bench.py --impl reference ...                   CPU arm: the oracle port on all host cores

A step = one pass of the hot path over one full synthetic volume per rank (weak scaling: every
rank owns an independent volume, no collective on the data path).  `value` is device-resident
throughput (decaes_t2map_device), `e2e` goes through the host-pointer C-ABI call with pinned host
buffers (H2D + kernels + D2H inside the timed region).  PyTorch is used for device memory, the
barrier / max-over-ranks reduction and CUDA events only.
"""
import argparse
import ctypes as C
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (shape, nTE, TE, nT2, Reg, extra opts)
    "cfg3": ((240, 240, 113), 56, 7e-3, 40, "lcurve", {}),
    "cfg2": ((240, 240, 48), 48, 8e-3, 40, "lcurve", {}),
    "cfg1": ((64, 64, 16), 32, 10e-3, 40, "none", {}),
    "cfg4": ((240, 240, 48), 48, 8e-3, 60, "chi2", {"Chi2Factor": 1.02}),
    "cfg5": ((256, 256, 160), 32, 10e-3, 60, "mdp", {"NoiseLevel": 1e-3}),
}
METRIC = "voxels/s T2map+T2part (56-echo 240x240x113, lcurve) at 1/2/4/8 B200 vs CPU"


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
    shape, nTE, TE, nT2, Reg, extra = WORKLOADS[wl_name]
    return f"{wl_name}: {nTE}-echo {'x'.join(map(str, shape))}, nT2={nT2}, Reg={Reg} + T2part"


def oracle_sample(orc, wl, nvox, seed, threads):
    """Run the CPU oracle on `nvox` voxels of the workload; returns (voxels/s, flops/voxel, stats)."""
    shape, nTE, TE, nT2, Reg, extra = wl
    img = orc.mock_image(nvox, nTE, TE, seed=seed)
    o = orc.make_t2map_opts((nvox, 1, 1), nTE, nT2, TE, Reg=Reg, **extra)
    p = orc.make_t2part_opts((nvox, 1, 1), nT2)
    _, st = orc.t2map(img, o, p, nthreads=threads, save_reg=False, save_resnorm=False)
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
    for _ in range(args.warmup):
        oracle_sample(orc, wl, sample, 1, threads)
    t0 = time.perf_counter()
    tot = 0.0
    for k in range(args.steps):
        vps, fpv, st = oracle_sample(orc, wl, sample, 1 + k, threads)
        tot += st.seconds
    wall = time.perf_counter() - t0
    value = sample * args.steps / tot
    shape, nTE, TE, nT2, Reg, extra = wl
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(wl_name), "sample_voxels_per_step": sample,
                   "note": "each step is a bounded sample of the workload on all host cores"},
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} voxels of the same synthetic workload per step, OpenMP C restatement of DECAES.jl (no Julia runtime in the image)"},
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


def main():
    out = claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--voxels", type=int, default=0, help="debug: override voxels per rank (invalidates the headline)")
    ap.add_argument("--ref-sample", type=int, default=16384)
    ap.add_argument("--cpu-sample", type=int, default=32768)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args, args.workload, out)
        return

    import torch
    import torch.distributed as dist
    pkg = load_package()
    pkg.lib()  # fail loudly if the CUDA library is missing — there is no fallback

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    shape, nTE, TE, nT2, Reg, extra = WORKLOADS[args.workload]
    nvox = args.voxels if args.voxels > 0 else shape[0] * shape[1] * shape[2]
    o = pkg.T2mapOptions(MatrixSize=(nvox, 1, 1), nTE=nTE, TE=TE, nT2=nT2, T2Range=(10e-3, 2.0), Reg=Reg, ngpus=1,
                         Silent=True, **extra).to_c()
    p = pkg.T2partOptions(MatrixSize=(nvox, 1, 1), nT2=nT2, T2Range=(10e-3, 2.0), SPWin=(10e-3, 25e-3),
                          MPWin=(25e-3, 200e-3), Silent=True).to_c()

    # ---- synthetic volume, resident in HBM (generated on the device, seeded per rank) ----
    img = torch.empty((nTE, nvox), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    pkg.mock_image_device(img.data_ptr(), nvox, nvox, rank * nvox, nTE, TE, seed=3, stream=stream)
    names = ["gdn", "ggm", "gva", "fnr", "snr", "alpha", "sfr", "sgm", "mfr", "mgm"]
    outs = {k: torch.empty((nvox,), dtype=torch.float64, device=dev) for k in names}
    outs["dist"] = torch.empty((nT2, nvox), dtype=torch.float64, device=dev)
    out_struct = pkg.make_out({k: v.data_ptr() for k, v in outs.items()})
    in_bytes = img.numel() * 8
    out_bytes = sum(v.numel() for v in outs.values()) * 8
    # inputs + outputs are ~5.5 GB per step for cfg3: far larger than the 126 MB L2, no flush needed

    def step_device():
        pkg.t2map_device(img.data_ptr(), nvox, nvox, o, p, out_struct, stream)

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
    elapsed_ms = e0.elapsed_time(e1)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = nvox * world * args.steps / (elapsed_ms * 1e-3)
    processed = pkg.last_stats()["voxels_processed"]
    checksum = float(outs["gdn"].sum().item())

    # ---- end to end through the host-pointer C-ABI call, pinned host buffers ----
    e2e = None
    one_volume = None
    if not args.no_e2e:
        h_img = torch.empty((nTE, nvox), dtype=torch.float64).pin_memory()
        h_img.copy_(img)
        h_outs = {k: torch.empty(v.shape, dtype=torch.float64).pin_memory() for k, v in outs.items()}
        h_struct = pkg.make_out({k: v.data_ptr() for k, v in h_outs.items()})
        e2e_steps = max(1, min(args.steps, 2))

        def step_host():
            rc = pkg.lib().decaes_t2map(h_img.data_ptr(), C.byref(o), C.byref(p), C.byref(h_struct))
            if rc != 0:
                raise RuntimeError(pkg.lib().decaes_last_error().decode())
        step_host()  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_host()
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        assert abs(float(h_outs["gdn"].sum().item()) - checksum) <= 1e-6 * abs(checksum) + 1e-9
        e2e = {"value": nvox * world * e2e_steps / dt, "unit": "voxels/s", "h2d_bytes_per_step": in_bytes,
               "d2h_bytes_per_step": out_bytes, "steps": e2e_steps, "host_stats": pkg.last_stats()}

        # ---- N > 1: ONE volume sharded over all N GPUs through the host-pointer call (north_star: "the 56-echo
        #      240x240x113 volume end to end in well under 1 s on 8 GPUs").  Rank 0 drives every device from its own
        #      process (decaes_t2map, ngpus = N, contiguous voxel slabs, no collective); the other ranks wait on a
        #      CPU-side (gloo) barrier so that their GPUs are idle.
        if world > 1:
            cpu_group = dist.new_group(backend="gloo")
            torch.cuda.synchronize()
            dist.barrier(group=cpu_group)
            if rank == 0:
                o_all = pkg.T2mapOptions(MatrixSize=(nvox, 1, 1), nTE=nTE, TE=TE, nT2=nT2, T2Range=(10e-3, 2.0), Reg=Reg,
                                         ngpus=world, Silent=True, **extra).to_c()

                def step_all():
                    rc = pkg.lib().decaes_t2map(h_img.data_ptr(), C.byref(o_all), C.byref(p), C.byref(h_struct))
                    if rc != 0:
                        raise RuntimeError(pkg.lib().decaes_last_error().decode())
                step_all()  # warm-up (allocates the workspaces of the other devices)
                times = []
                for _ in range(3):
                    t0 = time.perf_counter()
                    step_all()
                    times.append(time.perf_counter() - t0)
                assert abs(float(h_outs["gdn"].sum().item()) - checksum) <= 1e-6 * abs(checksum) + 1e-9
                st1 = pkg.last_stats()
                one_volume = {"seconds": min(times), "seconds_all": times, "voxels": nvox, "ngpus": st1["ngpus_used"],
                              "voxels_per_s": nvox / min(times), "host_stats": st1,
                              "what": "one volume in pinned host memory, sharded as contiguous slabs over all GPUs by decaes_t2map (H2D + kernels + D2H)"}
            dist.barrier(group=cpu_group)

    if rank == 0:
        # ---- CPU baseline (oracle port) + algorithmic FLOPs per voxel from its instrumented counters ----
        cpu = None
        flops_per_voxel = None
        if not args.no_cpu:
            import orc
            threads = os.cpu_count() or 1
            vps, flops_per_voxel, st = oracle_sample(orc, WORKLOADS[args.workload], args.cpu_sample, 3, threads)
            cpu = {"value": vps, "unit": "voxels/s", "cores": threads, "kind": "port",
                   "sample": f"{args.cpu_sample} voxels of the same synthetic workload (seed 3), OpenMP C restatement of DECAES.jl"}
        peak = pkg.measure_fp64_peak()
        mean_kernel_s = 1e-3 * sum(kernel_ms) / max(len(kernel_ms), 1)
        roofline = None
        # DRAM traffic of the pipeline kernel: measured once per round with ncu (profiles/r01_traffic.json,
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch), scaled to this launch's voxel count
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath) and args.workload == "cfg3":
            with open(tpath) as fh:
                tj = json.load(fh)
            traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["voxels"] * nvox
        if flops_per_voxel:
            achieved = flops_per_voxel * nvox / mean_kernel_s
            roofline = {"bound": "fp64", "achieved": achieved / 1e12, "peak": peak / 1e12, "unit": "TFLOP/s",
                        "frac": achieved / peak, "traffic": traffic,
                        "traffic_source": "ncu capture committed as profiles/r01_traffic.json (bytes per launch, scaled by voxels)",
                        "peak_source": "measured in this run by decaes_measure_fp64_peak (independent DFMA chains on all SMs); MEASURED_PEAKS.json has no FP64 entry",
                        "flops_per_voxel": flops_per_voxel, "kernel": "voxel_pipeline_kernel",
                        "kernel_ms": 1e3 * mean_kernel_s,
                        "hbm": {"algorithmic_bytes_per_voxel": 8 * (nTE + nT2 + 10),
                                "achieved_GBps": 8 * (nTE + nT2 + 10) * nvox / mean_kernel_s / 1e9}}
        line = {
            "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "sharding": "one volume per rank, T2part fused into the same kernel",
                       "voxels_per_rank": nvox, "l2": "inputs+outputs per step exceed L2 (no flush needed)",
                       "debug_voxels_override": bool(args.voxels)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": 3 * args.steps,  # basis_setup + gram_setup + voxel_pipeline per step
            "kernel_ms_per_step": sum(kernel_ms) / max(len(kernel_ms), 1),
            "roofline": roofline, "cpu_baseline": cpu, "one_volume_all_gpus": one_volume,
            "voxels_processed_last_step": processed, "checksum_gdn": checksum,
        }
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
