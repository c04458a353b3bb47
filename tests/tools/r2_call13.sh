#!/bin/bash
mkdir -p gpurun_out
L=bam3d_b200/libbam3d_gpu.so
timeout -s KILL 300 python -m pytest tests/test_narrow_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q -k "pinched or c2" 2>&1 | tail -5 | tee gpurun_out/c13_pinched.log
grep -q passed gpurun_out/c13_pinched.log || exit 1
bash tests/tools/r2_call10.sh > /dev/null 2>&1; mv gpurun_out/c10_trace.log gpurun_out/c13_trace.log; sort -t= -k2 -n gpurun_out/c13_trace.log | tail -4 | cut -c1-300
timeout -s KILL 200 python tests/tools/ab_time.py c2 $L@BAM3D_EPA_HUGE=0 $L --steps 10 2>&1 | tee gpurun_out/c13_huge_ab.log
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/c13_pytest.log
timeout -s KILL 600 python bench.py > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/c13_bench.json').read().strip().splitlines()[-1])
    print('c5', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'contact_all', d['contact_all']['value'], 'intersect_all', d['intersect_all']['value'])
    for k, v in d['extra'].items():
        print(k, v.get('value'), v.get('ms_per_step'), 'e2e', (v.get('e2e') or {}).get('value'))
except Exception as e:
    print('bench parse failed', e)
PY
tail -3 gpurun_out/c13_bench.err
