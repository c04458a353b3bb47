#!/bin/bash
# One GPU call: A/B the lane-refill intersect kernel against the one-pair-per-thread kernel on C1 (timing, CRC, oracle check on a
# prefix), run the GPU tests that reach GJK::intersect with it switched on, and print a bench line for the time-of-impact workload.
mkdir -p gpurun_out
L=bam3d_b200/libbam3d_gpu.so
K=${1:-16}
timeout -s KILL 60 python tests/tools/ab_time.py c1 $L $L@BAM3D_INTERSECT_REFILL=8 $L@BAM3D_INTERSECT_REFILL=16 $L@BAM3D_INTERSECT_REFILL=24 --check 65536 --steps 20 \
  > gpurun_out/intersect_refill_ab.log 2>&1
echo "ab rc=$?" >> gpurun_out/intersect_refill_ab.log
cat gpurun_out/intersect_refill_ab.log
BAM3D_INTERSECT_REFILL=$K timeout -s KILL 60 python -m pytest tests/test_narrow_gpu.py tests/test_golden_gpu.py tests/test_fullsize_gpu.py tests/test_polyhedron_faces_gpu.py \
  -m gpu -x -q -k "intersect or kat_gjk or edge_cases or narrow_against_golden or c1_full or mixed_polyhedron or shared_shapes or half_edge or compound" \
  > gpurun_out/intersect_refill_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/intersect_refill_tests.log
tail -4 gpurun_out/intersect_refill_tests.log
timeout -s KILL 60 python bench.py --workload toi > gpurun_out/bench_toi_r1h.json 2> gpurun_out/bench_toi_r1h.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_toi_r1h.json
