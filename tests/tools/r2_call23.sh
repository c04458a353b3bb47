#!/bin/bash
mkdir -p gpurun_out
for L in 2 3 4; do
BAM3D_BENCH_LANES=$L timeout -s KILL 600 python bench.py --steps 10 > gpurun_out/c23_bench_l$L.json 2> gpurun_out/c23_bench_l$L.err; echo "bench lanes=$L rc=$?"
python - $L <<'PY'
import json, sys
try:
    d = json.loads(open(f'gpurun_out/c23_bench_l{sys.argv[1]}.json').read().strip().splitlines()[-1])
    e = d['e2e']
    print('c5 %.4g %.2f ms' % (d['value'], d['ms_per_step']), 'e2e %.4g' % e['value'], 'contact_all %.4g' % d['contact_all']['value'], 'intersect_all %.4g' % d['intersect_all']['value'])
    print('  ' + ' | '.join('%s %.4g (%.2f ms) e2e %.4g' % (k, v.get('value'), v.get('ms_per_step'), (v.get('e2e') or {}).get('value', 0)) for k, v in d['extra'].items()))
except Exception as ex:
    print('bench parse failed', ex)
PY
tail -2 gpurun_out/c23_bench_l$L.err | cut -c1-200
done
