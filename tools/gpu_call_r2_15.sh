#!/bin/bash
# Final build, evidence pass 1: parity, stage times, sanitizers, ncu --set full of the three kernels
mkdir -p gpurun_out
exec > gpurun_out/call15.log 2>&1
echo "=== pytest -m gpu (everything)"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "=== stage times"
echo "--- default"; timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -2
echo "--- unforked"; JSMPEG_B200_PARSE_GROUPS=1 timeout 300 python tools/time_stages.py 64 60 2 2>&1 | tail -1
echo "--- 720p"; BENCH_WIDTH=1280 BENCH_HEIGHT=720 timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -1
echo "=== memcheck (final build): lane walk, TS demux, fused RGBA, goldens, corrupt streams, a 1080p clip"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_ts_cases.py tests/test_rgba_cases.py tests/test_gpu_parity.py -m gpu -x -q -k "ts or rgba or golden or lane_parallel or corrupted or whole_clip" > gpurun_out/r2_memcheck_final.log 2>&1
tail -4 gpurun_out/r2_memcheck_final.log
echo "=== racecheck (final build): lane walk (shared ring, staging), fused RGBA, goldens"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_rgba_cases.py tests/test_gpu_parity.py -m gpu -x -q -k "rgba or golden or lane_parallel" > gpurun_out/r2_racecheck_final.log 2>&1
tail -4 gpurun_out/r2_racecheck_final.log
echo "=== synccheck (final build)"
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 1 python -m pytest tests/test_rgba_cases.py tests/test_gpu_parity.py -m gpu -x -q -k "rgba or golden or lane_parallel" > gpurun_out/r2_synccheck_final.log 2>&1
tail -4 gpurun_out/r2_synccheck_final.log
echo "=== ncu --set full: reconstruct, 13 launches (I, 11 P, I)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:reconstruct_kernel -s 60 -c 13 \
   -o gpurun_out/prof_recon_r2f python tools/time_stages.py 64 60 2 > gpurun_out/ncu_recon_r2f.log 2>&1
tail -1 gpurun_out/ncu_recon_r2f.log
echo "=== ncu --set full: walk (unforked wave)"
JSMPEG_B200_PARSE_GROUPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk_pictures_lanes -s 1 -c 1 \
   -o gpurun_out/prof_walk_r2f python tools/time_stages.py 64 60 2 > gpurun_out/ncu_walk_r2f.log 2>&1
tail -1 gpurun_out/ncu_walk_r2f.log
echo "=== ncu --set full: expand (unforked wave)"
JSMPEG_B200_PARSE_GROUPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:expand_blocks -s 1 -c 1 \
   -o gpurun_out/prof_expand_r2f python tools/time_stages.py 64 60 2 > gpurun_out/ncu_expand_r2f.log 2>&1
tail -1 gpurun_out/ncu_expand_r2f.log
echo done
