#!/bin/bash
mkdir -p gpurun_out
L=bam3d_b200/libbam3d_gpu.so
timeout -s KILL 200 python tests/tools/ab_time.py c2 $L@BAM3D_EPA_PIPELINE=0 $L build/variants/lib_pipeA1.so --check 131072 --steps 10 2>&1 | tee gpurun_out/c6_pipe_ab.log
timeout -s KILL 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/c6_pytest.log
M=smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,smsp__thread_inst_executed_per_inst_executed.ratio
BAM3D_EPA_PIPELINE=1 timeout -s KILL 200 ncu --metrics $M --clock-control none -k regex:epa_refill -c 2 --csv --log-file gpurun_out/c6_ncu_pipe.csv python tests/tools/ab_time.py c2 --child --pairs 1048576 --steps 1 > gpurun_out/c6_ncu_pipe.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/c6_ncu_pipe.csv')) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    if r[h.index('ID')]=='1': print(f"{r[h.index('Metric Name')]:85s} {r[h.index('Metric Value')]}")
PY
