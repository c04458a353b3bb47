#!/bin/bash
# Final check of a build on one GPU: pytest -m gpu, smoke(), the default bench line and the reference arm. Run with gpurun.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/val_pytest.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -s KILL 600 python bench.py > gpurun_out/bench_c5_r2_final.json 2> gpurun_out/val_bench.err; echo "bench rc=$?"
timeout -s KILL 600 python bench.py --impl reference > gpurun_out/bench_c5_reference_r2_final.json 2> gpurun_out/val_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_c5_r2_final.json').read().strip().splitlines()[-1])
e = d['e2e']
print('c5 %.4g %.2f ms' % (d['value'], d['ms_per_step']), 'e2e %.4g' % e['value'], 'contact_all %.4g' % d['contact_all']['value'], 'intersect_all %.4g' % d['intersect_all']['value'], 'launches', d['gpu_launches'])
print('roofline', {k: d['roofline'][k] for k in ('achieved', 'frac', 'traffic')}, 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
for k, v in d['extra'].items():
    print(' ', k, '%.4g (%.3f ms) e2e %.4g' % (v.get('value'), v.get('ms_per_step'), (v.get('e2e') or {}).get('value', 0)), 'frac %.4f' % v['roofline']['frac'], 'cpu %.4g' % v['cpu_baseline']['value'], (v['roofline'].get('dominant_kernel_alone') or {}).get('frac'))
r = json.loads(open('gpurun_out/bench_c5_reference_r2_final.json').read().strip().splitlines()[-1])
print('reference arm', r['value'], r['cpu_baseline']['sample'][:80])
PY
