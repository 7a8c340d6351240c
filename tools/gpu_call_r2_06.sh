#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call06.log 2>&1
echo "=== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "=== stage times"
echo "--- default"; timeout 300 python tools/time_stages.py 64 60 3 2>&1 | tail -2
echo "--- unforked"; JSMPEG_B200_PARSE_GROUPS=1 timeout 300 python tools/time_stages.py 64 60 2 2>&1 | tail -1
echo "=== bench value leg with 1 / 2 / 4 decoders sharing the 64 streams"
for g in 1 2 4; do
  echo "--- value groups $g"
  BENCH_VALUE_GROUPS=$g timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['verified'], d['roofline']['frac'], d['stage_ms_per_step'])"
done
echo "=== ncu expand (unforked)"
JSMPEG_B200_PARSE_GROUPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:expand_blocks -s 1 -c 1 \
   -o gpurun_out/prof_expand_r2b python tools/time_stages.py 64 60 2 > gpurun_out/ncu_expand_r2b.log 2>&1
tail -1 gpurun_out/ncu_expand_r2b.log
echo done
