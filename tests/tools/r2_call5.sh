#!/bin/bash
mkdir -p gpurun_out
L=bam3d_b200/libbam3d_gpu.so
timeout -s KILL 200 python tests/tools/ab_time.py c2 $L@BAM3D_EPA_PIPELINE=0 $L build/variants/lib_pipeA0.so --steps 10 2>&1 | tee gpurun_out/c5_pipe_ab.log
M=smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,lts__t_sectors.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_drain_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum
for v in old:$L:0 pipe:$L:1 pipeA0:build/variants/lib_pipeA0.so:1; do
IFS=: read name lib pipe <<< "$v"
BAM3D_LIB=$lib BAM3D_EPA_PIPELINE=$pipe timeout -s KILL 200 ncu --metrics $M --clock-control none -k regex:epa_refill -c 2 --csv --log-file gpurun_out/c5_ncu_$name.csv python tests/tools/ab_time.py c2 --child --pairs 1048576 --steps 1 > gpurun_out/c5_ncu_$name.log 2>&1
done
python - <<'PY'
import csv
cols={}
for name in ("old","pipe","pipeA0"):
    rows=[r for r in csv.reader(open(f'gpurun_out/c5_ncu_{name}.csv')) if len(r)>10]
    h=rows[0]
    for r in rows[1:]:
        if r[h.index('ID')]=='1': cols.setdefault(r[h.index('Metric Name')],{})[name]=r[h.index('Metric Value')]
for k,v in cols.items(): print(f"{k:85s} " + "  ".join(f"{v.get(n,'-'):>16s}" for n in ("old","pipe","pipeA0")))
PY
