#!/bin/bash
# sanitizers on the round's final build (stage 1b rewritten since the last run)
mkdir -p gpurun_out
exec > gpurun_out/call20.log 2>&1
echo "=== memcheck: lane walk, stage 1b, TS demux, fused RGBA, goldens, corrupt streams, a 1080p clip"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_ts_cases.py tests/test_rgba_cases.py tests/test_gpu_parity.py -m gpu -x -q -k "ts or rgba or golden or lane_parallel or corrupted or whole_clip" > gpurun_out/r2_memcheck_final.log 2>&1
tail -4 gpurun_out/r2_memcheck_final.log
echo "=== racecheck"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_rgba_cases.py tests/test_gpu_parity.py -m gpu -x -q -k "rgba or golden or lane_parallel" > gpurun_out/r2_racecheck_final.log 2>&1
tail -4 gpurun_out/r2_racecheck_final.log
echo "=== synccheck"
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 1 python -m pytest tests/test_rgba_cases.py tests/test_gpu_parity.py -m gpu -x -q -k "rgba or golden or lane_parallel" > gpurun_out/r2_synccheck_final.log 2>&1
tail -4 gpurun_out/r2_synccheck_final.log
echo done
