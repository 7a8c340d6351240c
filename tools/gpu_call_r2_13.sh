#!/bin/bash
# 2-GPU validation of bench.py's multi-rank path (torchrun, NCCL barrier, per-rank NUMA binding)
mkdir -p gpurun_out
exec > gpurun_out/call13.log 2>&1
nvidia-smi -L
echo "=== cpu quota"; cat /sys/fs/cgroup/cpu.max; nproc
echo "=== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
echo rc=$?; tail -5 gpurun_out/r2_bench_2gpu.err; cat gpurun_out/r2_bench_2gpu.json
echo "=== reference arm N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>&1 | tail -3
echo "=== bench N=1 on the same box"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_1gpu_b.json 2> gpurun_out/r2_bench_1gpu_b.err
echo rc=$?; tail -3 gpurun_out/r2_bench_1gpu_b.err; cat gpurun_out/r2_bench_1gpu_b.json
echo done
