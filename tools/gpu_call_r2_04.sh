#!/bin/bash
# Round-2 GPU call 4: stage times of recon v4 / ring v2, parity at the bench configuration, bench.py end to end.
mkdir -p gpurun_out
exec > gpurun_out/call04.log 2>&1
echo "=== stage times"
echo "--- default"; timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -2
echo "--- sparse fetch forced"; JSMPEG_B200_RECON_DENSE=0 timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -2
echo "--- unforked"; JSMPEG_B200_PARSE_GROUPS=1 timeout 300 python tools/time_stages.py 64 60 2 2>&1 | tail -1
echo "--- r3 (lanes kernel 3 CTAs/SM)"; JSMPEG_B200_LIB=$PWD/variants/lib_r3.so timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -2
echo "--- 720p 64 streams"; BENCH_WIDTH=1280 BENCH_HEIGHT=720 timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -2
echo "=== parity at the bench configuration"
timeout 900 python -m pytest tests/test_gpu_bench_config.py -m gpu -x -q 2>&1 | tail -8
echo "=== bench.py end to end"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err; echo "rc=$?"
tail -c 3000 gpurun_out/bench_r2_a.json; tail -5 gpurun_out/bench_r2_a.err
echo "=== ncu --set full: reconstruct, 3 launches"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:reconstruct -s 70 -c 3 \
   -o gpurun_out/prof_recon_r2a python tools/time_stages.py 64 60 2 > gpurun_out/ncu_recon_r2a.log 2>&1
tail -2 gpurun_out/ncu_recon_r2a.log
echo done
