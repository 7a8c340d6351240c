#!/bin/bash
# expand: three-word reader; CTAs that take 1 / 4 / 8 runs of slots
mkdir -p gpurun_out
exec > gpurun_out/call18.log 2>&1
for v in default groups4 groups8; do
  echo "--- $v: parity"; JSMPEG_B200_LIB=$PWD/variants/lib_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_config.py -m gpu -x -q 2>&1 | tail -1
  echo "--- $v"; JSMPEG_B200_LIB=$PWD/variants/lib_$v.so timeout 200 python tools/time_stages.py 64 60 3 2>&1 | tail -2
  echo "--- $v unforked"; JSMPEG_B200_PARSE_GROUPS=1 JSMPEG_B200_LIB=$PWD/variants/lib_$v.so timeout 200 python tools/time_stages.py 64 60 2 2>&1 | tail -1
  echo "--- $v 720p"; BENCH_WIDTH=1280 BENCH_HEIGHT=720 JSMPEG_B200_LIB=$PWD/variants/lib_$v.so timeout 200 python tools/time_stages.py 64 60 2 2>&1 | tail -1
done
echo done
