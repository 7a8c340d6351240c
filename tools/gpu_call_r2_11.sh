#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call11.log 2>&1
echo "=== chunked parity (bench config test incl. chunk 6)"
timeout 900 python -m pytest tests/test_gpu_bench_config.py -m gpu -x -q 2>&1 | tail -2
JSMPEG_B200_CHUNK=3 JSMPEG_B200_CHUNK_MIN_WAVE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
echo "=== stage times"
echo "--- default"; timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -1
for cs in 2 3 4; do for g in 30 20 12 6; do echo "--- chunk $g streams $cs"; JSMPEG_B200_CHUNK=$g JSMPEG_B200_CHUNK_STREAMS=$cs timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -1; done; done
echo done
