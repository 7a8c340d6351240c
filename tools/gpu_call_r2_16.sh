#!/bin/bash
# 8-GPU box: bare D2H ceiling (bound / unbound), bench at N = 8 and N = 4 (both arms at 8)
mkdir -p gpurun_out
exec > gpurun_out/call16.log 2>&1
nvidia-smi -L | wc -l
echo "=== cpu quota"; cat /sys/fs/cgroup/cpu.max; nproc
nvidia-smi topo -m 2>/dev/null | head -12
echo "=== D2H ceiling, 8 ranks"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/d2h_ceiling.py 2>/dev/null | tail -1 | tee gpurun_out/r2_d2h_ceiling_8gpu.json
echo "=== D2H ceiling, 4 ranks"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 tools/d2h_ceiling.py 2>/dev/null | tail -1 | tee gpurun_out/r2_d2h_ceiling_4gpu.json
echo "=== bench N=8"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
echo rc=$?; tail -3 gpurun_out/r2_bench_8gpu.err; cut -c1-1500 gpurun_out/r2_bench_8gpu.json
echo "=== reference arm N=8"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 2>/dev/null | tail -1 | tee gpurun_out/r2_bench_8gpu_reference.json | cut -c1-600
echo "=== bench N=4"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err
echo rc=$?; tail -3 gpurun_out/r2_bench_4gpu.err; cut -c1-1500 gpurun_out/r2_bench_4gpu.json
echo "=== bench N=1 on this box"
timeout 900 python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/r2_bench_1gpu_on8.json 2> gpurun_out/r2_bench_1gpu_on8.err
echo rc=$?; cut -c1-1500 gpurun_out/r2_bench_1gpu_on8.json
echo done
