#!/bin/bash
mkdir -p gpurun_out
for b in 1 2 3 4; do
BAM3D_EPA_BLOCKS_PER_SM=$b timeout -s KILL 100 python tests/tools/bin_time.py 4194304 sph-sph,sph-cyl,cub-cyl,all 2>&1 | grep -v "^sum"
done | tee gpurun_out/c3_blocks.log
