#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_broad_gpu.py -m gpu -x -q -k "sphere_bounds" 2>&1 | tail -25 | tee gpurun_out/c29_sphere.log
timeout -s KILL 400 python -m pytest tests/test_broad_gpu.py -m gpu -q 2>&1 | tail -3
