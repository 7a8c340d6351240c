#!/bin/bash
# the slice walk (option "slice_walk") on the B200, with whatever is left of the round's GPU minutes: its GPU tests,
# the B tests again (their routing code was generalised), the timed 64 x 720p row-sliced I/P/B wave with and without it
mkdir -p gpurun_out
exec > gpurun_out/call24.log 2>&1
echo "=== pytest tests/test_gpu_zz_slice_walk.py tests/test_gpu_zz_b_pictures.py"
timeout 70 python -m pytest tests/test_gpu_zz_slice_walk.py tests/test_gpu_zz_b_pictures.py -q 2>&1 | tail -8
echo "=== tools/time_b.py 64 3"
TIME_B_OUT=r2_slice_walk_720p.json timeout 50 python tools/time_b.py 64 3 2>&1 | tail -14
echo done
