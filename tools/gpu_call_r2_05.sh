#!/bin/bash
# Round-2 GPU call 5: lane walk with staged relative records + fix-up (no second pass).
mkdir -p gpurun_out
exec > gpurun_out/call05.log 2>&1
echo "=== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "=== stage times"
echo "--- default (64 regs)"; timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -2
echo "--- r3 (lanes kernel 3 CTAs/SM, 80 regs)"; JSMPEG_B200_LIB=$PWD/variants/lib_r3.so timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -2
echo "--- unforked default"; JSMPEG_B200_PARSE_GROUPS=1 timeout 300 python tools/time_stages.py 64 60 2 2>&1 | tail -1
echo "--- unforked r3"; JSMPEG_B200_LIB=$PWD/variants/lib_r3.so JSMPEG_B200_PARSE_GROUPS=1 timeout 300 python tools/time_stages.py 64 60 2 2>&1 | tail -1
echo "--- 64 x 1, 64 x 12 (latency)"; timeout 200 python tools/time_stages.py 64 1 3 2>&1 | tail -1; timeout 200 python tools/time_stages.py 64 12 3 2>&1 | tail -1
echo "=== ncu --set full: walk (unforked wave)"
JSMPEG_B200_PARSE_GROUPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk_pictures_lanes -s 1 -c 1 \
   -o gpurun_out/prof_walk_r2b python tools/time_stages.py 64 60 2 > gpurun_out/ncu_walk_r2b.log 2>&1
tail -2 gpurun_out/ncu_walk_r2b.log
echo done
