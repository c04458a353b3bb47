#!/bin/bash
# ncu evidence of a build (profiles/README.md): launch list + DRAM bytes of one bench command, then --set full captures of the
# dominant kernels through tests/tools/ab_time.py; outputs in gpurun_out/ (summarise with profiles/summarize.py). Run with gpurun.
# ncu evidence of the round-2 build: launch list + DRAM bytes of one bench command, --set full captures of the dominant kernels
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# the command exits 0 without ncu first
timeout -s KILL 300 python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/cap_bench_plain.json 2> gpurun_out/cap_bench_plain.err || { echo "plain bench failed"; exit 1; }
timeout -s KILL 1200 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -c 700 --csv --log-file gpurun_out/r2_c5_launches.csv python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/cap_ncu_bench.log 2>&1; echo "launch list rc=$? lines $(wc -l < gpurun_out/r2_c5_launches.csv)"
cap() { # name workload kernel-regex [pairs]
  name=$1; wl=$2; k=$3; n=${4:-1048576}
  timeout -s KILL 600 python tests/tools/ab_time.py $wl --child --steps 1 --pairs $n > /dev/null 2>&1 || { echo "$name: plain run failed"; return; }
  timeout -s KILL 900 $NCU --set full --import-source on -k regex:$k --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2_$name python tests/tools/ab_time.py $wl --child --steps 1 --pairs $n > gpurun_out/cap_$name.log 2>&1
  ncu -i gpurun_out/r2_$name.ncu-rep --page raw --csv > gpurun_out/r2_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2_$name.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r2_${name}_source.csv 2>/dev/null
  python profiles/src_hotspots.py gpurun_out/r2_${name}_source.csv 40 > gpurun_out/r2_${name}_source_hotspots.txt 2>/dev/null
  rm -f gpurun_out/r2_${name}_source.csv
  ls -la gpurun_out/r2_$name.ncu-rep | awk '{print $5, $9}'
  [ "$name" != "c2" ] && rm -f gpurun_out/r2_$name.ncu-rep
}
cap c2 c2 epa_refill_kernel
cap c2_huge c2 gjk_contact_huge_kernel
cap c2_phase c2 gjk_phase_kernel
cap c3 c3 gjk_contact_kernel
cap dist dist gjk_distance_refill_kernel
cap toi toi gjk_toi_refill_kernel
cap c1 c1 gjk_intersect_kernel
ls -la gpurun_out | grep r2_ | awk '{print $5, $9}'
