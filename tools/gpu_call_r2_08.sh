#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call08.log 2>&1
echo "=== pytest -m gpu (subset)"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
echo "=== stage times"
echo "--- default"; timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -2
for g in 30 20 15; do echo "--- chunk $g"; JSMPEG_B200_CHUNK=$g timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -1; done
echo "--- chunk 30, 4 parse groups"; JSMPEG_B200_CHUNK=30 JSMPEG_B200_PARSE_GROUPS=4 timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -1
echo "--- chunk 20, 4 parse groups"; JSMPEG_B200_CHUNK=20 JSMPEG_B200_PARSE_GROUPS=4 timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -1
echo "=== d2h ceiling, 1 GPU"
timeout 120 python tools/d2h_ceiling.py 2>&1 | tail -1
echo done
