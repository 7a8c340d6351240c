#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call10.log 2>&1
echo "=== pytest -m gpu (everything)"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "=== stage times"
echo "--- default"; timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -2
echo "=== sanitizer: memcheck on the new TS demux + fused RGBA + golden"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_ts_cases.py tests/test_rgba_cases.py tests/test_gpu_parity.py -m gpu -x -q -k "ts or rgba or golden or lane_parallel" > gpurun_out/r2_memcheck_ts_rgba.log 2>&1
tail -4 gpurun_out/r2_memcheck_ts_rgba.log
echo "=== racecheck: fused RGBA (shared-memory chroma exchange) + golden"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_rgba_cases.py tests/test_gpu_parity.py -m gpu -x -q -k "rgba or golden" > gpurun_out/r2_racecheck_rgba.log 2>&1
tail -4 gpurun_out/r2_racecheck_rgba.log
echo done
