#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_narrow_gpu.py tests/test_broad_gpu.py -m gpu -x -q -k "pinched or undecided or degenerate" 2>&1 | tail -8 | tee gpurun_out/c9_pytest_a.log
L=bam3d_b200/libbam3d_gpu.so
timeout -s KILL 200 python tests/tools/ab_time.py c2 $L@BAM3D_EPA_HUGE=0 $L --steps 10 2>&1 | tee gpurun_out/c9_huge_ab.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"huge|overflow|epa_refill" -c 6 --csv --log-file gpurun_out/c9_ncu.csv python tests/tools/ab_time.py c2 --child --pairs 1048576 --steps 1 > gpurun_out/c9_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/c9_ncu.csv')) if len(r)>10]
h=rows[0]
for r in rows[1:]: print(r[h.index('Kernel Name')][:50], r[h.index('Metric Value')])
PY
