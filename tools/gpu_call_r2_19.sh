#!/bin/bash
# final build of the round: parity, stage times, bench lines, launch lists, expand capture
# round's bench lines and the ncu launch lists of the same command
mkdir -p gpurun_out
exec > gpurun_out/call19.log 2>&1
echo "=== pytest -m gpu (everything)"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "=== stage times"
echo "--- default"; timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -2
echo "--- unforked"; JSMPEG_B200_PARSE_GROUPS=1 timeout 300 python tools/time_stages.py 64 60 2 2>&1 | tail -1
echo "--- 720p"; BENCH_WIDTH=1280 BENCH_HEIGHT=720 timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -1
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
echo "=== ncu --set full: expand (unforked wave)"
JSMPEG_B200_PARSE_GROUPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:expand_blocks -s 1 -c 1 \
   -o gpurun_out/prof_expand_r2h python tools/time_stages.py 64 60 2 > gpurun_out/ncu_expand_r2h.log 2>&1
tail -1 gpurun_out/ncu_expand_r2h.log
echo done
