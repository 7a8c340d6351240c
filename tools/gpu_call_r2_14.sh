#!/bin/bash
# warp-uniform IDCT path in stage 2 + length-sorted expand: parity, stage times, variants, one ncu capture of expand
mkdir -p gpurun_out
exec > gpurun_out/call14.log 2>&1
echo "=== pytest -m gpu (everything)"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "=== stage times"
echo "--- default"; timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -2
echo "--- unforked"; JSMPEG_B200_PARSE_GROUPS=1 timeout 300 python tools/time_stages.py 64 60 2 2>&1 | tail -1
echo "--- 720p"; BENCH_WIDTH=1280 BENCH_HEIGHT=720 timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -1
for v in expand128 expand512 recon6; do
  echo "--- variant $v"; JSMPEG_B200_LIB=$PWD/variants/lib_$v.so timeout 200 python tools/time_stages.py 64 60 2 2>&1 | tail -1
  echo "--- variant $v unforked"; JSMPEG_B200_PARSE_GROUPS=1 JSMPEG_B200_LIB=$PWD/variants/lib_$v.so timeout 200 python tools/time_stages.py 64 60 2 2>&1 | tail -1
done
echo "=== ncu expand (unforked)"
JSMPEG_B200_PARSE_GROUPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:expand_blocks -s 1 -c 1 \
   -o gpurun_out/prof_expand_r2c python tools/time_stages.py 64 60 1 > gpurun_out/ncu_expand_r2c.log 2>&1
tail -2 gpurun_out/ncu_expand_r2c.log
echo "=== ncu --set full: reconstruct, 3 launches"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:reconstruct -s 70 -c 3 \
   -o gpurun_out/prof_recon_r2b python tools/time_stages.py 64 60 1 > gpurun_out/ncu_recon_r2b.log 2>&1
tail -2 gpurun_out/ncu_recon_r2b.log
echo done
