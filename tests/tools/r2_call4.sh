#!/bin/bash
mkdir -p gpurun_out
L=bam3d_b200/libbam3d_gpu.so
timeout -s KILL 200 python tests/tools/ab_time.py c2 $L@BAM3D_EPA_PIPELINE=0 $L@BAM3D_EPA_PIPELINE=1 build/variants/lib_ring1.so build/variants/lib_ring3.so --check 131072 --steps 10 2>&1 | tee gpurun_out/c4_pipe_ab.log
timeout -s KILL 100 python tests/tools/bin_time.py 4194304 sph-sph,sph-cyl,cub-cyl,cub-cub,all 2>&1 | grep -v "^sum" | tee gpurun_out/c4_bins.log
timeout -s KILL 200 python -m pytest tests/test_narrow_gpu.py tests/test_fullsize_gpu.py tests/test_golden_gpu.py -m gpu -x -q -k "contact or golden or frame or c2 or compound" 2>&1 | tail -5 | tee gpurun_out/c4_pytest.log
