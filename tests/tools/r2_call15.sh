#!/bin/bash
mkdir -p gpurun_out
L=bam3d_b200/libbam3d_gpu.so
timeout -s KILL 300 python -m pytest tests/test_narrow_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q -k "pinched or c2" 2>&1 | tail -5 | tee gpurun_out/c15_pinched.log
grep -q passed gpurun_out/c15_pinched.log || exit 1
g++ -std=c++17 -O1 -ffp-contract=off tests/cpp/kat_gpu.cpp -o /tmp/kat_gpu -Lbam3d_b200 -lbam3d_gpu -Wl,-rpath,$PWD/bam3d_b200 && /tmp/kat_gpu 2>&1 | tail -12 | tee gpurun_out/c15_kat.log
bash tests/tools/r2_call10.sh > /dev/null 2>&1; mv gpurun_out/c10_trace.log gpurun_out/c15_trace512.log; sort -t= -k2 -n gpurun_out/c15_trace512.log | grep huge | awk 'NR%11==1' | cut -c1-300; sort -t= -k2 -n gpurun_out/c15_trace512.log | tail -2 | cut -c1-300
timeout -s KILL 200 python tests/tools/ab_time.py c2 $L@BAM3D_EPA_HUGE=0 $L build/variants/lib_huge1024.so --steps 10 2>&1 | tee gpurun_out/c15_huge_ab.log
for prio in 0 1; do
BAM3D_SIDE_PRIORITY=$prio timeout -s KILL 600 python bench.py > gpurun_out/c15_bench_p$prio.json 2> gpurun_out/c15_bench_p$prio.err; echo "bench prio=$prio rc=$?"
python - $prio <<'PY'
import json, sys
try:
    d = json.loads(open(f'gpurun_out/c15_bench_p{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print('c5', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'contact_all', d['contact_all']['value'], 'intersect_all', d['intersect_all']['value'])
    for k, v in d['extra'].items():
        print(k, v.get('value'), v.get('ms_per_step'), 'e2e', (v.get('e2e') or {}).get('value'))
except Exception as e:
    print('bench parse failed', e)
PY
done
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/c15_pytest.log
