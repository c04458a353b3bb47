#!/bin/bash
# One GPU call: A/B the lane-refill time-of-impact kernel against the one-pair-per-thread kernel (timing, CRC of the full output,
# oracle check on a prefix), then run every GPU test that reaches the time-of-impact path with the refill kernel switched on.
mkdir -p gpurun_out
L=bam3d_b200/libbam3d_gpu.so
timeout -s KILL 100 python tests/tools/ab_time.py toi $L $L@BAM3D_TOI_REFILL=1 $L@BAM3D_TOI_REFILL=8 $L@BAM3D_TOI_REFILL=16 --check 65536 --steps 10 \
  > gpurun_out/toi_refill_ab.log 2>&1
echo "ab rc=$?" >> gpurun_out/toi_refill_ab.log
cat gpurun_out/toi_refill_ab.log
BAM3D_TOI_REFILL=${1:-1} timeout -s KILL 75 python -m pytest tests/test_narrow_gpu.py tests/test_golden_gpu.py tests/test_polyhedron_faces_gpu.py tests/test_fullsize_gpu.py \
  -m gpu -x -q -k "time_of_impact or toi or compound or narrow_against_golden or half_edge" > gpurun_out/toi_refill_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/toi_refill_tests.log
tail -5 gpurun_out/toi_refill_tests.log
