"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.
  python profiles/summarize.py launches <launches.csv> <out.txt>        (gpu__time_duration per launch -> per-kernel totals and shares)
  python profiles/summarize.py raw <report.ncu-rep> <out.txt>            (key metrics of every captured launch)
"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__sass_average_branch_targets_threads_uniform.pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum", "sm__sass_thread_inst_executed_op_fp32_pred_on.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg, total = {}, 0.0
    for r in rows[1:]:
        k = r[ki].split("(")[0]
        ms = float(r[vi]) / 1e6
        agg.setdefault(k, [0.0, 0])
        agg[k][0] += ms
        agg[k][1] += 1
        total += ms
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none: {len(rows) - 1} launches, {total:.3f} ms total (cold-cache, serialised: compare SHARES)\n")
        f.write(f"{'kernel':70s} {'launches':>8s} {'total ms':>10s} {'avg ms':>10s} {'share':>7s}\n")
        for k, (ms, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write(f"{k[:70]:70s} {c:8d} {ms:10.3f} {ms / c:10.4f} {100 * ms / total:6.1f}%\n")


def raw(rep, dst):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none, report {rep.split('/')[-1]}\n")
        for r in rows[2:]:
            f.write("\n== " + r[hdr.index("Kernel Name")][:100] + "\n")
            for i, h in enumerate(hdr):
                if h in KEYS:
                    f.write(f"  {h:70s} {r[i]:>18s} {units[i]}\n")
            for i, h in enumerate(hdr):
                if "issue_stalled" in h and h.endswith("per_warp_active.pct"):
                    try:
                        if float(r[i]) > 5:
                            f.write(f"  stall {h[26:-20]:64s} {r[i]:>18s} %\n")
                    except ValueError:
                        pass


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
