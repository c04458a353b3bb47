#!/bin/bash
# ncu --set full of the device-resident BVH query kernels inside `bench.py --workload c4` (2^20 rays, 1 024 frusta over 4M boxes).
mkdir -p gpurun_out
timeout -s KILL 300 python bench.py --workload c4 --steps 1 --warmup 1 > /dev/null 2> gpurun_out/bvhq_plain.err || { echo "plain run failed"; tail -3 gpurun_out/bvhq_plain.err; exit 1; }
for k in bvh_query_append_kernel bvh_brute_append_kernel bvh_ray_closest_kernel; do
  timeout -s KILL 600 ncu --clock-control none --set full --import-source on -k regex:$k --launch-count 1 -f -o gpurun_out/r2_$k python bench.py --workload c4 --steps 1 --warmup 1 > gpurun_out/bvhq_$k.log 2>&1
  ncu -i gpurun_out/r2_$k.ncu-rep --page raw --csv > gpurun_out/r2_${k}_raw.csv 2>/dev/null
  rm -f gpurun_out/r2_$k.ncu-rep
  ls -la gpurun_out/r2_${k}_raw.csv | awk '{print $5, $9}'
done
