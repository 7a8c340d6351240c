#!/bin/bash
# B-picture extension, second call: the B walk now runs beside the I/P walk (own stream).  Tests again, the timed
# wave, bench.py's b_pictures_720p leg on its own, an ncu launch list + full captures of the two new kernels,
# and compute-sanitizer on the new kernels (as far as the round's last GPU minutes reach).
mkdir -p gpurun_out
exec > gpurun_out/call23.log 2>&1
echo "=== pytest tests/test_gpu_zz_b_pictures.py"
timeout 40 python -m pytest tests/test_gpu_zz_b_pictures.py -q 2>&1 | tail -5
echo "=== tools/time_b.py 64 4"
timeout 40 python tools/time_b.py 64 4 2>&1 | tail -10
echo "=== bench.b_pictures_leg"
timeout 40 python -c "
import json, bench
print(json.dumps(bench.b_pictures_leg(0, 64)))" 2>&1 | tail -1
echo "=== ncu launch list (time_b, extension on, one repetition)"
TIME_B_QUICK=1 timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_b_launches.csv \
   python tools/time_b.py 64 1 > gpurun_out/ncu_b1.log 2>&1
tail -2 gpurun_out/ncu_b1.log | cut -c1-200
echo "=== ncu --set full: reconstruct_b_kernel, walk_pictures_b_kernel"
TIME_B_QUICK=1 timeout 90 ncu --set full --clock-control none --import-source on -k regex:'reconstruct_b_kernel|walk_pictures_b_kernel' -c 3 \
   -o gpurun_out/r2_b_kernels python tools/time_b.py 64 1 > gpurun_out/ncu_b2.log 2>&1
tail -2 gpurun_out/ncu_b2.log | cut -c1-200
echo "=== compute-sanitizer memcheck on the B tests (reference ABI cases)"
timeout 70 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_zz_b_pictures.py -q -k "reference_abi or fused or mixes" > gpurun_out/r2_memcheck_b.txt 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/r2_memcheck_b.txt
echo done
