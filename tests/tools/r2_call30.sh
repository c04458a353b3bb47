#!/bin/bash
mkdir -p gpurun_out
g++ -std=c++17 -O1 -ffp-contract=off tests/cpp/kat_gpu.cpp -o /tmp/kat_gpu -Lbam3d_b200 -lbam3d_gpu -Wl,-rpath,$PWD/bam3d_b200 && /tmp/kat_gpu 2>&1 | tail -12 | tee gpurun_out/c30_kat.log
