"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.
  python profiles/summarize.py launches <launches.csv> <out.txt>        (gpu__time_duration per launch -> per-kernel totals and shares)
  python profiles/summarize.py raw <report.ncu-rep | raw.csv> <out.txt>  (key metrics of every captured launch)
  python profiles/summarize.py step <launches.csv> <out.txt> <closing kernel> [counters.json [name]]
        (one step cut out of a launch list: per-kernel time shares and DRAM traffic; optionally added to the counters file)
  python profiles/summarize.py counters <out.json> name=raw.csv[:kernel substring] ...
        (the counters bench.py's roofline object quotes — DRAM bytes, FP32-pipe, branch and lane efficiency of the named
         kernel's launch — from `ncu -i X.ncu-rep --page raw --csv` files; written as profiles/ncu_counters.json)
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


def read_launches(src):
    """-> launches in order: dicts with k (kernel name), ms, dram (bytes; None when the list has no DRAM metrics)."""
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, ui, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    by_id, order = {}, []
    for r in rows[1:]:
        if r[0] not in by_id:
            by_id[r[0]] = {"k": r[ki].split("(")[0].replace("void ", "").replace("b3::", ""), "ms": 0.0, "dram": None}
            order.append(r[0])
        d, v = by_id[r[0]], float(r[vi].replace(",", ""))
        if r[mi] == "gpu__time_duration.sum":
            d["ms"] = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[ui], 1e-6)
        elif r[mi].startswith("dram__bytes"):
            d["dram"] = (d["dram"] or 0.0) + v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1.0)
    return [by_id[i] for i in order]


def table(f, ls):
    agg, total = {}, sum(d["ms"] for d in ls)
    for d in ls:
        a = agg.setdefault(d["k"], [0.0, 0, 0.0])
        a[0] += d["ms"]; a[1] += 1; a[2] += d["dram"] or 0.0
    f.write(f"{'kernel':60s} {'launches':>8s} {'total ms':>10s} {'avg ms':>10s} {'share':>7s} {'DRAM MB/launch':>15s}\n")
    for k, (ms, c, by) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        f.write(f"{k[:60]:60s} {c:8d} {ms:10.3f} {ms / c:10.4f} {100 * ms / total:6.1f}% {by / c / 1e6:15.1f}\n")


def launches(src, dst):
    ls = read_launches(src)
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --clock-control none: {len(ls)} launches, {sum(d['ms'] for d in ls):.3f} ms total "
                f"(cold-cache, serialised: compare SHARES)\n")
        table(f, ls)


def step(src, dst, last_kernel, json_dst=None, name="c5"):
    """One step out of a launch list: from the first launch of our own kernels after the set-up (pack_shapes / torch fills) to
    the first launch of `last_kernel` (the step's closing kernel), e.g. the C5 mixed step ends with compact_contacts_kernel."""
    import json
    ls = read_launches(src)
    i0 = next(i for i, d in enumerate(ls) if not (d["k"].startswith("pack_shapes") or d["k"].startswith("at::") or "mesh_" in d["k"]))
    i1 = next(i for i in range(i0, len(ls)) if last_kernel in ls[i]["k"])
    st = ls[i0:i1 + 1]
    ms, by = sum(d["ms"] for d in st), sum(d["dram"] or 0.0 for d in st)
    with open(dst, "w") as f:
        f.write(f"# one step of {src.split('/')[-1]}: launches {i0}..{i1} ({len(st)} launches), {ms:.3f} ms serialised, DRAM traffic {by / 1e9:.3f} GB\n")
        table(f, st)
        f.write("\n# in launch order\n")
        for d in st:
            f.write(f"{d['k'][:60]:60s} {d['ms']:10.4f} ms {(d['dram'] or 0.0) / 1e6:12.1f} MB\n")
    if json_dst:
        try:
            out = json.load(open(json_dst))
        except Exception:
            out = {}
        top = max(st, key=lambda d: d["ms"])
        out[name] = {"source": "profiles/" + dst.split("/")[-1], "dram_bytes": by, "step_ms_under_ncu": ms, "launches_per_step": len(st),
                     "dram_gbs": by / (ms * 1e-3) / 1e9, "top_kernel": top["k"], "top_kernel_share": top["ms"] / ms,
                     "kernels": {k: v for k, v in out.items() if k in ("c2", "c3")} and {"epa_refill_kernel": "see c2", "gjk_contact_kernel<true>": "see c3"}}
        json.dump(out, open(json_dst, "w"), indent=1, sort_keys=True)


def read_raw(rep):
    if rep.endswith(".csv"):
        text = open(rep).read()
    else:
        text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    lines = text.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))  # skip ncu's own log lines
    return list(csv.reader(lines[start:]))


def counters(dst, *specs):
    import json
    import os
    out = {}
    for spec in specs:
        name, _, rest = spec.partition("=")
        path, _, kernel = rest.partition(":")
        rows = read_raw(path)
        hdr = rows[0]
        ki = hdr.index("Kernel Name")
        cand = [r for r in rows[2:] if kernel in r[ki]]
        if not cand:
            raise SystemExit(f"{path}: no launch of a kernel matching {kernel!r}")

        def val(r, key):
            try:
                return float(r[hdr.index(key)].replace(",", ""))
            except (ValueError, IndexError):
                return None

        def unit(key):
            return rows[1][hdr.index(key)] if key in hdr else ""
        r = max(cand, key=lambda r: val(r, "gpu__time_duration.sum") or 0.0)  # the longest matching launch
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        rd = (val(r, "dram__bytes_read.sum") or 0.0) * scale.get(unit("dram__bytes_read.sum"), 1.0)
        wr = (val(r, "dram__bytes_write.sum") or 0.0) * scale.get(unit("dram__bytes_write.sum"), 1.0)
        tscale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}
        ms = (val(r, "gpu__time_duration.sum") or 0.0) * tscale.get(unit("gpu__time_duration.sum"), 1e-6)
        out[name] = {
            "kernel": r[ki].split("(")[0], "source": "profiles/" + os.path.basename(path).replace("_raw.csv", "_ncu_full.txt"),
            "duration_ms_under_ncu": ms, "dram_bytes": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
            "dram_gbs": (rd + wr) / (ms * 1e-3) / 1e9 if ms else None,
            "fp32_pipe_pct": val(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
            "alu_pipe_pct": val(r, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
            "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "branch_uniform_pct": val(r, "smsp__sass_average_branch_targets_threads_uniform.pct"),
            "lanes_active": val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
            "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "l1_hit_pct": val(r, "l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": val(r, "lts__t_sector_hit_rate.pct"),
            "registers": val(r, "launch__registers_per_thread"), "grid": val(r, "launch__grid_size"), "block": val(r, "launch__block_size"),
        }
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    print(f"{dst}: {', '.join(out)}")


def raw(rep, dst):
    rows = read_raw(rep)
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
    {"launches": launches, "raw": raw, "counters": counters, "step": step}[sys.argv[1]](*sys.argv[2:])
