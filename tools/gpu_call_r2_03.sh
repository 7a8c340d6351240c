#!/bin/bash
# Round-2 GPU call 3: stage 2 with early record requests + L2 prefetches, simplified ring refill.
mkdir -p gpurun_out
exec > gpurun_out/call03.log 2>&1
echo "=== pytest -m gpu (default build)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "=== pytest -m gpu, sparse record fetch forced"
JSMPEG_B200_RECON_DENSE=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
echo "=== stage times"
echo "--- default"; timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -2
echo "--- sparse fetch forced"; JSMPEG_B200_RECON_DENSE=0 timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -2
echo "--- unforked"; JSMPEG_B200_PARSE_GROUPS=1 timeout 300 python tools/time_stages.py 64 60 2 2>&1 | tail -1
echo "--- r3 (lanes kernel 3 CTAs/SM)"; JSMPEG_B200_LIB=$PWD/variants/lib_r3.so timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -2
echo "--- 720p 64 streams"; BENCH_WIDTH=1280 BENCH_HEIGHT=720 timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -2
echo "=== ncu --set full: reconstruct, 3 launches"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:reconstruct -s 70 -c 3 \
   -o gpurun_out/prof_recon_r2a python tools/time_stages.py 64 60 2 > gpurun_out/ncu_recon_r2a.log 2>&1
tail -2 gpurun_out/ncu_recon_r2a.log
echo done
