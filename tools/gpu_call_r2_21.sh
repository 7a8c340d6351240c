#!/bin/bash
# batched start-code scan (one launch + five copies for all streams): parity, stage times, and the round's
# bench lines / launch lists again on this build
mkdir -p gpurun_out
exec > gpurun_out/call21.log 2>&1
echo "=== pytest -m gpu (everything)"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "=== stage times"
echo "--- default"; timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -2
echo "--- unforked"; JSMPEG_B200_PARSE_GROUPS=1 timeout 300 python tools/time_stages.py 64 60 2 2>&1 | tail -1
echo "--- 720p"; BENCH_WIDTH=1280 BENCH_HEIGHT=720 timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -1
echo "--- 64 x 1"; timeout 200 python tools/time_stages.py 64 1 3 2>&1 | tail -1
echo "=== smoke()"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== bench.py (default flags)"
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
echo rc=$?; tail -2 gpurun_out/r2_bench_1gpu.err; cut -c1-400 gpurun_out/r2_bench_1gpu.json
echo "=== bench.py --impl reference"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_1gpu_reference.json 2>/dev/null
cut -c1-300 gpurun_out/r2_bench_1gpu_reference.json
echo "=== ncu launch list (forked = default)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv \
   python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extras --no-verify > gpurun_out/ncu_bench_r2.log 2>&1
tail -1 gpurun_out/ncu_bench_r2.log | cut -c1-200
echo "=== ncu launch list (unforked)"
JSMPEG_B200_PARSE_GROUPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_unforked.csv \
   python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extras --no-verify > gpurun_out/ncu_bench_r2u.log 2>&1
tail -1 gpurun_out/ncu_bench_r2u.log | cut -c1-200
echo "=== the same command, unforked, not under ncu"
JSMPEG_B200_PARSE_GROUPS=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --no-verify 2>/dev/null > gpurun_out/r2_bench_unforked_live.json
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_unforked_live.json').read()); print(d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['avg_launch_ms'])" 2>&1 | tail -1
echo done
