#!/bin/bash
mkdir -p gpurun_out
L=bam3d_b200/libbam3d_gpu.so
timeout -s KILL 120 python -m pytest tests/test_narrow_gpu.py -m gpu -x -q -k "native_gather or distance_refill or frame_queue" > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
tail -12 gpurun_out/c2_pytest.log
timeout -s KILL 200 python tests/tools/bin_time.py 4194304 > gpurun_out/c2_bin_time.log 2>&1; cat gpurun_out/c2_bin_time.log
M=smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct
for c in 1 32; do
BAM3D_EPA_COHORT=$c timeout -s KILL 200 ncu --metrics $M --clock-control none -k regex:epa_refill -c 2 --csv --log-file gpurun_out/c2_ncu_cohort$c.csv python tests/tools/ab_time.py c2 --child --pairs 1048576 --steps 1 > gpurun_out/c2_ncu_cohort$c.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/c2_ncu_cohort$c.csv')) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    if r[h.index('ID')]=='1': print('cohort $c', r[h.index('Metric Name')], r[h.index('Metric Value')])
PY
done
