#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_narrow_gpu.py -m gpu -x -q -k "pinched or undecided or out_of_range or execution_options" 2>&1 | tail -25 | tee gpurun_out/c8_pytest_a.log
L=bam3d_b200/libbam3d_gpu.so
timeout -s KILL 200 python tests/tools/ab_time.py c2 $L@BAM3D_EPA_HUGE=0 $L --check 131072 --steps 10 2>&1 | tee gpurun_out/c8_huge_ab.log
timeout -s KILL 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/c8_pytest.log
