#!/bin/bash
# Round-2 GPU call 1: host/NUMA probe, A/B of the macro-gated walk candidates, sanitizer runs of the
# lane-parallel walk.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
exec > gpurun_out/call01.log 2>&1
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nvidia-smi topo -m
nproc; python - <<'EOF'
import os
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us", "/sys/fs/cgroup/cpuset.cpus.effective"):
    try: print(p, open(p).read().strip())
    except Exception as e: print(p, "n/a", e)
EOF
lscpu | head -25
cat /sys/devices/system/node/node*/cpulist 2>/dev/null
free -g | head -2
set +x
echo "=== A/B variants (time_stages 64 streams x 60 pictures x 3 reps)"
for v in default fixup emit wide fixup_wide emit_wide; do
  echo "--- $v"
  JSMPEG_B200_LIB=$PWD/variants/lib_$v.so timeout 200 python tools/time_stages.py 64 60 3 2>&1 | tail -2
done
echo "=== parity of the variants (golden + stage1 + whole clip)"
for v in fixup emit wide; do
  echo "--- $v"
  JSMPEG_B200_LIB=$PWD/variants/lib_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or stage1 or whole_clip or lane_parallel" 2>&1 | tail -3
done
echo "=== sanitizer: memcheck, lane-parallel walk (default build)"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or corrupted or whole_clip or lane_parallel or rgba or ts_demux" > gpurun_out/r2_memcheck_lanes.log 2>&1
tail -5 gpurun_out/r2_memcheck_lanes.log
echo "=== sanitizer: racecheck, lane-parallel walk"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or lane_parallel" > gpurun_out/r2_racecheck_lanes.log 2>&1
tail -5 gpurun_out/r2_racecheck_lanes.log
echo "=== sanitizer: synccheck, lane-parallel walk"
timeout 400 compute-sanitizer --tool synccheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or lane_parallel" > gpurun_out/r2_synccheck_lanes.log 2>&1
tail -5 gpurun_out/r2_synccheck_lanes.log
echo done
