#!/bin/bash
# Round-2 GPU call 2: new engine (exceptions, pipeline), ring bit reader, dense side array, expand v2.
mkdir -p gpurun_out
exec > gpurun_out/call02.log 2>&1
echo "=== pytest -m gpu (default build)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "=== pytest -m gpu with the chunked pipeline forced on small waves (G=3)"
JSMPEG_B200_CHUNK=3 JSMPEG_B200_CHUNK_MIN_WAVE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8
echo "=== stage times, default build"
for chunk in 0 12 6 20; do
  echo "--- chunk $chunk"
  JSMPEG_B200_CHUNK=$chunk timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -2
done
echo "--- 64 x 1 picture (I pictures only), 64 x 12"
timeout 200 python tools/time_stages.py 64 1 3 2>&1 | tail -2
timeout 200 python tools/time_stages.py 64 12 3 2>&1 | tail -2
echo "--- unforked (PARSE_GROUPS=1), chunk 0"
JSMPEG_B200_PARSE_GROUPS=1 timeout 300 python tools/time_stages.py 64 60 2 2>&1 | tail -1
echo "=== stage times, lanes kernel at 3 CTAs/SM (80 registers)"
for chunk in 0 12; do
  echo "--- r3 chunk $chunk"
  JSMPEG_B200_LIB=$PWD/variants/lib_r3.so JSMPEG_B200_CHUNK=$chunk timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -2
done
echo "=== ncu --set full: walk + expand, one launch each (unforked wave)"
JSMPEG_B200_PARSE_GROUPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk_pictures_lanes -s 1 -c 1 \
   -o gpurun_out/prof_walk_r2a python tools/time_stages.py 64 60 2 > gpurun_out/ncu_walk_r2a.log 2>&1
tail -2 gpurun_out/ncu_walk_r2a.log
JSMPEG_B200_PARSE_GROUPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:expand_blocks -s 1 -c 1 \
   -o gpurun_out/prof_expand_r2a python tools/time_stages.py 64 60 2 > gpurun_out/ncu_expand_r2a.log 2>&1
tail -2 gpurun_out/ncu_expand_r2a.log
echo done
