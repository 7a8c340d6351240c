#!/bin/bash
# B-picture extension on the B200: its GPU tests, a timed wave of 64 x 720p I/P/B streams (checked against the
# oracle), then -- as far as the round's remaining GPU minutes reach -- the existing parity suite and the stage
# times of the default workload on this build (SASS of every pre-existing kernel is unchanged, tools/sass_hash.py)
mkdir -p gpurun_out
exec > gpurun_out/call22.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "=== pytest tests/test_gpu_zz_b_pictures.py"
timeout 150 python -m pytest tests/test_gpu_zz_b_pictures.py -q 2>&1 | tail -25
echo "=== tools/time_b.py 64 4"
timeout 90 python tools/time_b.py 64 4 2>&1 | tail -12
echo "=== smoke()"
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== pytest tests/test_gpu_parity.py"
timeout 240 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4
echo "=== stage times, default workload"
timeout 150 python tools/time_stages.py 64 60 3 2>&1 | tail -3
echo done
