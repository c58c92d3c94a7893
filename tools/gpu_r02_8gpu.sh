#!/bin/bash
# eight-GPU run: ONE cfg3 volume sharded over the 8 B200s of the box (bench.py under torchrun: `value` = strong scaling,
# e2e from pinned and from pageable host memory, bit-equality with the 1-GPU result)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 3 --warmup 3 --no-replicas > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err
tail -c 2500 gpurun_out/r02_bench_8gpu.json | head -c 1200; tail -n 2 gpurun_out/r02_bench_8gpu.err
