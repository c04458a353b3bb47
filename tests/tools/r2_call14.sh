#!/bin/bash
mkdir -p gpurun_out
L=bam3d_b200/libbam3d_gpu.so
timeout -s KILL 300 python -m pytest tests/test_narrow_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q -k "pinched or c2" 2>&1 | tail -5 | tee gpurun_out/c14_pinched.log
grep -q passed gpurun_out/c14_pinched.log || exit 1
BAM3D_LIB=build/variants/lib_huge1024.so timeout -s KILL 300 python -m pytest tests/test_narrow_gpu.py -m gpu -x -q -k "pinched" 2>&1 | tail -2
bash tests/tools/r2_call10.sh > /dev/null 2>&1; mv gpurun_out/c10_trace.log gpurun_out/c14_trace512.log; sort -t= -k2 -n gpurun_out/c14_trace512.log | grep huge | awk 'NR%8==1' | cut -c1-300; sort -t= -k2 -n gpurun_out/c14_trace512.log | tail -2 | cut -c1-300
sed -i 's/lib_hugetrace.so/lib_hugetrace1024.so/' tests/tools/r2_call10.sh
bash tests/tools/r2_call10.sh > /dev/null 2>&1; mv gpurun_out/c10_trace.log gpurun_out/c14_trace1024.log; sort -t= -k2 -n gpurun_out/c14_trace1024.log | grep huge | awk 'NR%8==1' | cut -c1-300; sort -t= -k2 -n gpurun_out/c14_trace1024.log | tail -2 | cut -c1-300
timeout -s KILL 200 python tests/tools/ab_time.py c2 $L@BAM3D_EPA_HUGE=0 $L build/variants/lib_huge1024.so --steps 10 2>&1 | tee gpurun_out/c14_huge_ab.log
