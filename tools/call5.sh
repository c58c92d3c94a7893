#!/bin/bash
# two-GPU validation: full GPU suite (multi-GPU test not skipped) + the N = 2 bench line (weak scaling + one volume over both GPUs)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c5_smi.txt
( time timeout 500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c5_pytest.log 2>&1
tail -n 6 gpurun_out/c5_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 > gpurun_out/c5_bench_2gpu.json 2> gpurun_out/c5_bench_2gpu.err
tail -c 1500 gpurun_out/c5_bench_2gpu.json
tail -n 5 gpurun_out/c5_bench_2gpu.err
