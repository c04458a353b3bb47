#!/bin/bash
# round 2, GPU call 1: full GPU suite, EPA cohort A/B, distance refill A/B, the new default bench (short)
mkdir -p gpurun_out
L=bam3d_b200/libbam3d_gpu.so
timeout -s KILL 300 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -15 gpurun_out/c1_pytest.log
timeout -s KILL 240 python tests/tools/ab_time.py c2 $L@BAM3D_EPA_COHORT=1 $L@BAM3D_EPA_COHORT=16 $L@BAM3D_EPA_COHORT=24 $L@BAM3D_EPA_COHORT=28 $L@BAM3D_EPA_COHORT=31 $L@BAM3D_EPA_COHORT=32 \
  --check 65536 --steps 10 > gpurun_out/c1_cohort_ab.log 2>&1; echo "rc=$?" >> gpurun_out/c1_cohort_ab.log
cat gpurun_out/c1_cohort_ab.log
timeout -s KILL 120 python tests/tools/ab_time.py dist $L@BAM3D_DISTANCE_REFILL=0 $L@BAM3D_DISTANCE_REFILL=1 $L@BAM3D_DISTANCE_REFILL=8 $L@BAM3D_DISTANCE_REFILL=16 \
  --check 65536 --steps 10 > gpurun_out/c1_dist_ab.log 2>&1; echo "rc=$?" >> gpurun_out/c1_dist_ab.log
cat gpurun_out/c1_dist_ab.log
( time timeout -s KILL 420 python bench.py --steps 5 --warmup 3 ) > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/c1_bench.err
head -c 3000 gpurun_out/c1_bench.json
