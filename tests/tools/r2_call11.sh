#!/bin/bash
mkdir -p gpurun_out
bash tests/tools/r2_call10.sh 2>&1 | sort -t= -k2 -n | tail -8 | cut -c1-330
L=bam3d_b200/libbam3d_gpu.so
timeout -s KILL 200 python tests/tools/ab_time.py c2 $L@BAM3D_EPA_HUGE=0 $L --steps 10 2>&1 | tee gpurun_out/c11_huge_ab.log
timeout -s KILL 200 python -m pytest tests/test_narrow_gpu.py -m gpu -x -q -k "pinched" 2>&1 | tail -3
