#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call12.log 2>&1
echo "=== pytest -m gpu (everything)"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "=== stage times"
echo "--- default"; timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -2
echo "--- unforked"; JSMPEG_B200_PARSE_GROUPS=1 timeout 300 python tools/time_stages.py 64 60 2 2>&1 | tail -1
echo "--- 64 x 1"; timeout 200 python tools/time_stages.py 64 1 3 2>&1 | tail -1
echo "--- 720p"; BENCH_WIDTH=1280 BENCH_HEIGHT=720 timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -1
echo done
