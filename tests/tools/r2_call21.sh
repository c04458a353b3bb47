#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/c21_pytest.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
for m in 0 1; do
if [ $m = 1 ]; then export BAM3D_BENCH_MAT4=1; fi
timeout -s KILL 600 python bench.py --no-extra > gpurun_out/c21_bench_m$m.json 2> gpurun_out/c21_bench_m$m.err; echo "bench mat4=$m rc=$?"
python - $m <<'PY'
import json, sys
try:
    d = json.loads(open(f'gpurun_out/c21_bench_m{sys.argv[1]}.json').read().strip().splitlines()[-1])
    e = d['e2e']
    print('c5 %.4g %.2f ms' % (d['value'], d['ms_per_step']), 'e2e %.4g' % e['value'], 'h2d MB %.0f d2h MB %.0f' % (e['h2d_bytes_per_step']/1e6, e['d2h_bytes_per_step']/1e6), 'contact_all %.4g' % d['contact_all']['value'])
except Exception as ex:
    print('bench parse failed', ex)
PY
done
