#!/bin/bash
# two-GPU validation: N-GPU == 1-GPU bit equality (tests), bench under torchrun at N = 2 (one volume sharded over both GPUs)
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_host_path.py -m gpu -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
tail -c 1500 gpurun_out/r02_bench_2gpu.json; tail -n 3 gpurun_out/r02_bench_2gpu.err
