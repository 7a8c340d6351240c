#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call07.log 2>&1
echo "=== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "=== stage times"
echo "--- default"; timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -2
echo "--- unforked"; JSMPEG_B200_PARSE_GROUPS=1 timeout 300 python tools/time_stages.py 64 60 2 2>&1 | tail -1
echo "--- 4 groups"; JSMPEG_B200_PARSE_GROUPS=4 timeout 300 python tools/time_stages.py 64 60 2 2>&1 | tail -1
echo "--- 2 groups"; JSMPEG_B200_PARSE_GROUPS=2 timeout 300 python tools/time_stages.py 64 60 2 2>&1 | tail -1
echo "=== ncu walk (unforked)"
JSMPEG_B200_PARSE_GROUPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk_pictures_lanes -s 1 -c 1 \
   -o gpurun_out/prof_walk_r2c python tools/time_stages.py 64 60 2 > gpurun_out/ncu_walk_r2c.log 2>&1
tail -1 gpurun_out/ncu_walk_r2c.log
echo done
